#!/bin/bash
# round 2, call K: why is the C4 slice slow in stream mode (k_stream_fused 1.1 ms for 8.4M particles, C3 0.35 ms for 10M)?
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-sort --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 3"
SPR_STREAM_TWO_PASS=1 timeout 300 python bench.py --workload c4 $B > $O/r2k_c4_twopass.json 2> $O/r2k_c4_twopass.err
echo "== c4 stream, two-pass"; python tools/bench_summary.py < $O/r2k_c4_twopass.json
timeout 300 python bench.py --workload c4 $B > $O/r2k_c4_fused.json 2> $O/r2k_c4_fused.err
echo "== c4 stream, fused"; python tools/bench_summary.py < $O/r2k_c4_fused.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_fused -s 6 -c 1 -o $O/r2k_c4_stream_fused -f python bench.py --workload c4 $B --steps 3 --no-kernel-timing > $O/r2k_ncu.log 2>&1
echo "ncu exit $?"; tail -3 $O/r2k_ncu.log
