#!/bin/bash
# round 2, call BG N: the peer-memory assembly kernel with 8 records per thread in flight (SPR_PEER_UNROLL=8) against 4, N GPUs, one CTA per SM
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
N=${1:-8}
for U in 4 8; do
  SPR_PEER_UNROLL=$U timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$U bench.py --gpus $N --steps 10 --warmup 3 --assembly-shares 1,2 > $O/r2bg_n${N}_u$U.json 2> $O/r2bg_n${N}_u$U.err
  echo "n$N unroll $U exit $?"
  python - <<P
import json
d=json.load(open('$O/r2bg_n${N}_u$U.json'))
g=d['gather']
print('unroll $U: value', d['value']/1e9, 'ms', d['ms_per_step'])
for k in ("fused","fused_overlapped","fused_deferred"):
    print('  ', k, g[k].get("ms_per_step_by_assembly_share"), g[k].get("ms_per_step"), g[k].get("matches_nccl_path_bit_for_bit"))
print('   kernel alone', g['fused']['nvlink_roofline']['ms_per_launch'], g['fused']['nvlink_roofline']['achieved_gbs'])
P
done
