#!/bin/bash
# round 2, call AW (run twice: 4-tile histogram CTAs, then 1-tile): A/B on one box -- the current build against the build before the hoisted prologues / 4-tile histogram CTAs (prehoist), alternating
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --no-kernel-timing --steps 20 --warmup 5"
for R in 1 2; do
  for W in c3 c4; do
    for V in "" prehoist; do
      SPR_LIB_VARIANT=$V timeout 300 python bench.py --workload $W $B > $O/r2aw_${W}_${V:-now}_$R.json 2> $O/r2aw_${W}_${V:-now}_$R.err
      python -c "import json;d=json.load(open('$O/r2aw_${W}_${V:-now}_$R.json'));print('$W ${V:-now} run $R: %.4f ms'%d['ms_per_step'], ['%.4f'%x for x in d['repeats']['ms_per_step']])"
    done
  done
done
