#!/bin/bash
# round 2, call W: current per-kernel breakdown of C5 (high churn) and C2, with the host timing of C5
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for W in c5 c2; do
  SPR_HOST_TIMING=1 timeout 600 python bench.py --workload $W --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5 > $O/r2w_$W.json 2> $O/r2w_$W.err; echo "$W exit $?"
  python tools/bench_summary.py < $O/r2w_$W.json; grep "spr host" $O/r2w_$W.err | tail -3
done
