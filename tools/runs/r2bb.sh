#!/bin/bash
# round 2, call BB: final build -- GPU suite, smoke(), default bench + reference arm, then C2 / C4 / C5 for the README numbers
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=3 > $O/r2bb_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2bb_pytest.log; tail -4 $O/r2bb_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2bb_smoke.log 2>&1; echo "smoke exit $?" >> $O/r2bb_smoke.log; tail -2 $O/r2bb_smoke.log
timeout 900 python bench.py > $O/r2bb_bench_default.json 2> $O/r2bb_bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2bb_bench_ref.json 2> $O/r2bb_bench_ref.err; echo "ref exit $?"
python tools/bench_summary.py < $O/r2bb_bench_default.json
for W in c2 c4 c5; do
  timeout 600 python bench.py --workload $W --no-cpu-baseline --no-billboards --no-vertex-stage --steps 20 --warmup 5 > $O/r2bb_$W.json 2> $O/r2bb_$W.err; echo "== $W exit $?"
  python tools/bench_summary.py < $O/r2bb_$W.json | head -3
done
