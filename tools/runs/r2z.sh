#!/bin/bash
# round 2, call Z: host workers that poll for the next job (SPR_HOST_SPIN_US) against parked ones, C5
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nproc; 
for SP in 0 3000; do for T in 1 4; do
  SPR_HOST_SPIN_US=$SP SPR_HOST_THREADS=$T SPR_HOST_TIMING=1 timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5 > $O/r2z_c5_s${SP}_t$T.json 2> $O/r2z_c5_s${SP}_t$T.err; echo "c5 spin=$SP threads=$T exit $?"
  python tools/bench_summary.py < $O/r2z_c5_s${SP}_t$T.json | head -1; grep "spr host" $O/r2z_c5_s${SP}_t$T.err | tail -1
done; done
