#!/bin/bash
# round 2, call AC: ncu --set full of the C5 (100k one-granule effects) kernels of one frame
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
CMD="python bench.py --workload c5 --steps 2 --warmup 4 --repeats 1 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --no-kernel-timing"
timeout 300 $CMD > $O/r2ac_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_plan|k_emit_warp|k_integrate_pass' -s 40 -c 5 -o $O/r2ac_c5 -f $CMD > $O/r2ac_ncu.log 2>&1
echo "ncu exit $?"; tail -2 $O/r2ac_ncu.log
timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5 > $O/r2ac_c5.json 2> $O/r2ac_c5.err; python tools/bench_summary.py < $O/r2ac_c5.json
