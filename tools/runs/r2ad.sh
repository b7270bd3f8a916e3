#!/bin/bash
# round 2, call AD: per-system camera in batched updates -- full GPU suite, then C3 / C4 / C5 for regressions
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2ad_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2ad_pytest.log; tail -6 $O/r2ad_pytest.log
for W in c3 c4 c5; do
  timeout 600 python bench.py --workload $W --no-cpu-baseline --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5 > $O/r2ad_$W.json 2> $O/r2ad_$W.err; echo "$W exit $?"
  python tools/bench_summary.py < $O/r2ad_$W.json | grep -E "ms/step|k_integrate|k_emit|k_plan"
done
