#!/bin/bash
# round 2, call AI: k_sort_small on a cluster of two CTAs -- GPU suite, then C2 (64 x 16k) and C5 / C3 for regressions
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2ai_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2ai_pytest.log; tail -6 $O/r2ai_pytest.log
for W in c2 c3; do
  timeout 600 python bench.py --workload $W --no-cpu-baseline --no-billboards --no-vertex-stage --no-c4-slice --no-stream-mode --steps 20 --warmup 5 > $O/r2ai_$W.json 2> $O/r2ai_$W.err; echo "$W exit $?"
  python tools/bench_summary.py < $O/r2ai_$W.json
done
