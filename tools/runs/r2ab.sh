#!/bin/bash
# round 2, call AB: k_plan with the timer draws of a sub-step evaluated side by side, one-granule effects with transposed stream stores
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2ab_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2ab_pytest.log; tail -4 $O/r2ab_pytest.log
for W in c5 c2; do
  SPR_HOST_TIMING=1 timeout 600 python bench.py --workload $W --no-cpu-baseline --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5 > $O/r2ab_$W.json 2> $O/r2ab_$W.err; echo "$W exit $?"
  python tools/bench_summary.py < $O/r2ab_$W.json; grep "spr host" $O/r2ab_$W.err | tail -1
done
