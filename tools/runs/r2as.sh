#!/bin/bash
# round 2, call AS: A/B of the work hoisted ahead of griddepcontrol.wait (default build) against the same build with the wait first (nohoist), alternating
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --no-kernel-timing --steps 20 --warmup 5"
for R in 1 2 3; do
  for W in c3 c4; do
    for V in "" nohoist; do
      SPR_LIB_VARIANT=$V timeout 300 python bench.py --workload $W $B > $O/r2as_${W}_${V:-hoist}_$R.json 2> $O/r2as_${W}_${V:-hoist}_$R.err
      python -c "import json;d=json.load(open('$O/r2as_${W}_${V:-hoist}_$R.json'));print('$W ${V:-hoist} run $R: %.4f ms'%d['ms_per_step'], ['%.4f'%x for x in d['repeats']['ms_per_step']])"
    done
  done
done
