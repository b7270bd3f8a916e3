#!/bin/bash
# round 2, call BI: final build (bulk-async assembly at small shares by default) -- GPU suite with the assembly tests at share 1 and 8, smoke, default bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2bi_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2bi_pytest.log; tail -3 $O/r2bi_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2bi_smoke.log 2>&1; echo "smoke exit $?" >> $O/r2bi_smoke.log; tail -2 $O/r2bi_smoke.log
timeout 900 python bench.py --no-cpu-baseline --no-billboards > $O/r2bi_bench_default.json 2> $O/r2bi_bench_default.err; echo "bench exit $?"
python tools/bench_summary.py < $O/r2bi_bench_default.json | head -2
