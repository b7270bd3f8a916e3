#!/bin/bash
# round 2, call AV: reference measurement of the build after the GLSL pin, deferred expansion and the hoisted prologues -- GPU suite, smoke(), default bench + reference arm, then the ncu
# launch list and --set full capture of one steady-state sorted C3 step (same command, after it exited 0 without ncu)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > $O/r2av_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2av_pytest.log; tail -4 $O/r2av_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2av_smoke.log 2>&1; echo "smoke exit $?" >> $O/r2av_smoke.log; tail -2 $O/r2av_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2av_bench_default.json 2> $O/r2av_bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2av_bench_ref.json 2> $O/r2av_bench_ref.err; echo "ref exit $?"
python tools/bench_summary.py < $O/r2av_bench_default.json
CMD="python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --no-kernel-timing"
timeout 300 $CMD > $O/r2av_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2av_launches_c3_sorted.csv $CMD > $O/r2av_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_expand_sorted|k_sort_pass|k_integrate_pass|k_sort_hist' -s 40 -c 7 -o $O/r2av_c3_sorted_step -f $CMD > $O/r2av_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 $O/r2av_ncu_full.log
