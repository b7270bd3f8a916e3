#!/bin/bash
# round 2, call V: where the depth-bucket kernels spend their time (timing-only variants: SPR_BUCKET_DBG)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="python bench.py --workload c3 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5"
for D in 0 16 1 2 4 8; do
  SPR_BUCKET_DBG=$D timeout 600 $B > $O/r2v_dbg$D.json 2> $O/r2v_dbg$D.err; echo "dbg $D exit $?"; python tools/bench_summary.py < $O/r2v_dbg$D.json | grep -E "ms/step|bucket_(scatter|finish|hist)"
done
