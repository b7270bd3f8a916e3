#!/bin/bash
# round 2, call X: k_expand_sorted with the warp's transposition through shared memory (SPR_EXPAND_STAGE=1) against shuffles
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for W in c3 c4 c2; do
B="python bench.py --workload $W --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5"
timeout 600 $B > $O/r2x_${W}_shfl.json 2> $O/r2x_${W}_shfl.err; echo "$W shfl exit $?"; python tools/bench_summary.py < $O/r2x_${W}_shfl.json | grep -E "ms/step|expand_sorted"
SPR_EXPAND_STAGE=1 timeout 600 $B > $O/r2x_${W}_stage.json 2> $O/r2x_${W}_stage.err; echo "$W stage exit $?"; python tools/bench_summary.py < $O/r2x_${W}_stage.json | grep -E "ms/step|expand_sorted"
done
