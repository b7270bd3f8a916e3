#!/bin/bash
# round 2, call O (N GPUs): bench with the draw-stream assembly in the headline; assembly share sweep
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --assembly-shares ${2:-2,3,4,8} > $O/r2o_n$N.json 2> $O/r2o_n$N.err
echo "exit $?"; tail -5 $O/r2o_n$N.err
python - <<PY
import json
d = json.loads(open("$O/r2o_n$N.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "gpu_launches")}); print("compute_only", d["compute_only"]["value"], d["compute_only"]["ms_per_step"])
g = d["gather"]; print("nccl", g["nccl"]["ms_per_step"], "fused", g["fused"]["ms_per_step"], g["fused"]["nvlink_roofline"], "match", g["fused"]["matches_nccl_path_bit_for_bit"])
print("overlapped", g["fused_overlapped"])
PY
