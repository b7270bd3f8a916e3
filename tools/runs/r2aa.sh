#!/bin/bash
# round 2, call AA: verification of the final build -- GPU suite, smoke(), default bench + reference arm, workload matrix
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 > $O/r2aa_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2aa_pytest.log; tail -10 $O/r2aa_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2aa_smoke.log 2>&1; echo "smoke exit $?" >> $O/r2aa_smoke.log; tail -2 $O/r2aa_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2aa_bench_default.json 2> $O/r2aa_bench_default.err; echo "bench exit $?"
python tools/bench_summary.py < $O/r2aa_bench_default.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2aa_bench_ref.json 2> $O/r2aa_bench_ref.err; echo "ref exit $?"; tail -c 600 $O/r2aa_bench_ref.json
for W in c2 c4 c5; do
  timeout 600 python bench.py --workload $W --no-cpu-baseline --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5 > $O/r2aa_$W.json 2> $O/r2aa_$W.err; echo "$W exit $?"
  python tools/bench_summary.py < $O/r2aa_$W.json | head -1
done
