#!/bin/bash
# round 2, call AQ N: the multi-GPU bench with the deferred-expansion leg (gather.fused_deferred) on N GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --assembly-shares ${2:-1,2,4} > $O/r2aq_n$N.json 2> $O/r2aq_n$N.err
echo "n$N exit $?"; tail -3 $O/r2aq_n$N.err
python - <<P
import json
d=json.load(open('$O/r2aq_n$N.json'))
print('value', d['value']/1e9, 'ms', d['ms_per_step'], 'compute_only', d['compute_only']['ms_per_step'])
g=d['gather']
for k in ("nccl","fused","fused_overlapped","fused_deferred"):
    print(k, g.get(k, {}).get("assembly_kernel"), g.get(k, {}).get("ms_per_step_by_assembly_share"), g.get(k, {}).get("ms_per_step"), (g.get(k, {}).get("value_with_gather") or 0)/1e9, g.get(k, {}).get("matches_nccl_path_bit_for_bit"))
print(d['config'].get('assembly'))
P
