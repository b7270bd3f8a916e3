#!/bin/bash
# round 2, call AY: A/B on one box -- hoisted prologues only in k_emit_warp / k_finalize / k_sort_small (current) against the build before any hoist (prehoist)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --no-kernel-timing --steps 20 --warmup 5"
for R in 1 2; do
  for W in c3 c4 c2; do
    for V in "" prehoist; do
      SPR_LIB_VARIANT=$V timeout 300 python bench.py --workload $W $B > $O/r2ay_${W}_${V:-now}_$R.json 2> $O/r2ay_${W}_${V:-now}_$R.err
      python -c "import json;d=json.load(open('$O/r2ay_${W}_${V:-now}_$R.json'));print('$W ${V:-now} run $R: %.4f ms'%d['ms_per_step'], ['%.4f'%x for x in d['repeats']['ms_per_step']])"
    done
  done
done
