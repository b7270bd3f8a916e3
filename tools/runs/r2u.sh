#!/bin/bash
# round 2, call U: ncu --set full of the depth-bucket kernels in one steady-state sorted C3 step
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 6 --repeats 1 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --no-kernel-timing"
timeout 300 $CMD > $O/r2u_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_bucket_hist|k_bucket_scatter|k_bucket_finish|k_bucket_scan' -s 25 -c 5 -o $O/r2u_c3_buckets -f $CMD > $O/r2u_ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 $O/r2u_ncu_full.log
