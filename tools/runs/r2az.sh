#!/bin/bash
# round 2, call AZ: where a C5 frame's host time goes on this build (SPR_HOST_TIMING=1), kernel times beside it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5"
SPR_HOST_TIMING=1 timeout 300 python bench.py --workload c5 $B --no-kernel-timing > $O/r2az_c5_host.json 2> $O/r2az_c5_host.err; grep "spr host" $O/r2az_c5_host.err | tail -4
python -c "import json;d=json.load(open('$O/r2az_c5_host.json'));print('c5: %.4f ms'%d['ms_per_step'], ['%.4f'%x for x in d['repeats']['ms_per_step']], 'e2e', d['e2e']['ms_per_step'])"
timeout 300 python bench.py --workload c5 $B > $O/r2az_c5.json 2> $O/r2az_c5.err; python tools/bench_summary.py < $O/r2az_c5.json
nproc; grep -c processor /proc/cpuinfo; grep "model name" /proc/cpuinfo | head -1
