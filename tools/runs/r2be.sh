#!/bin/bash
# round 2, call BE: ncu launch list and --set full capture of one steady-state sorted C3 step on the final build (k_expand_sorted at one record per thread)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --no-kernel-timing"
timeout 300 $CMD > $O/r2be_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2be_launches_c3_sorted.csv $CMD > $O/r2be_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_expand_sorted|k_sort_pass|k_integrate_pass|k_sort_hist' -s 40 -c 7 -o $O/r2be_c3_sorted_step -f $CMD > $O/r2be_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 $O/r2be_ncu_full.log
