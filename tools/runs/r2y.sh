#!/bin/bash
# round 2, call Y: persistent host workers (Context::hostPool) -- tests that cover the threaded passes, then C5
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "batch or many_systems or billboard or c5 or churn or refused or random_api" > $O/r2y_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2y_pytest.log; tail -5 $O/r2y_pytest.log
for T in 1 4; do
  SPR_HOST_THREADS=$T SPR_HOST_TIMING=1 timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5 > $O/r2y_c5_t$T.json 2> $O/r2y_c5_t$T.err; echo "c5 threads=$T exit $?"
  python tools/bench_summary.py < $O/r2y_c5_t$T.json | head -1; grep "spr host" $O/r2y_c5_t$T.err | tail -2
done
timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5 > $O/r2y_c5.json 2> $O/r2y_c5.err; echo "c5 default exit $?"; python tools/bench_summary.py < $O/r2y_c5.json | head -1
