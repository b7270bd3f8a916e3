#!/bin/bash
# round 2, call BH N: the peer-memory assembly through bulk async loads + tensor stores (SPR_PEER_TMA=1: no data registers, 128-thread CTAs) against the LSU kernel, N GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
N=${1:-8}
for T in 0 1; do
  SPR_PEER_TMA=$T timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$T bench.py --gpus $N --steps 10 --warmup 3 --assembly-shares ${2:-1,2,4} > $O/r2bh_n${N}_tma$T.json 2> $O/r2bh_n${N}_tma$T.err
  echo "n$N tma $T exit $?"
  python - <<P
import json
d=json.load(open('$O/r2bh_n${N}_tma$T.json'))
g=d['gather']
print('tma $T: value', d['value']/1e9, 'ms', d['ms_per_step'])
for k in ("fused","fused_overlapped","fused_deferred"):
    print('  ', k, g[k].get("ms_per_step_by_assembly_share"), g[k].get("ms_per_step"), g[k].get("matches_nccl_path_bit_for_bit"))
print('   kernel alone', g['fused']['nvlink_roofline']['ms_per_launch'], g['fused']['nvlink_roofline']['achieved_gbs'])
P
done
