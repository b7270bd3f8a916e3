#!/bin/bash
# round 2, call AR: histogram CTAs over 4 consecutive tiles, work-list reads and table clears ahead of griddepcontrol.wait in the sort kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2ar_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2ar_pytest.log; tail -4 $O/r2ar_pytest.log
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5"
for W in c3 c4 c2; do
  timeout 300 python bench.py --workload $W $B > $O/r2ar_$W.json 2> $O/r2ar_$W.err; echo "== $W exit $?"
  python tools/bench_summary.py < $O/r2ar_$W.json | egrep "ms/step|k_sort|k_expand_sorted"
done
