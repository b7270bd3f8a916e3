#!/bin/bash
# round 2, call BA: k_sort_small shared by 2 or 4 CTAs per effect by key range -- GPU suite, C2 A/B (SPR_SMALL_SPLIT=0 / 1), alternating
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2ba_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2ba_pytest.log; tail -6 $O/r2ba_pytest.log
B="--no-cpu-baseline --no-billboards --no-vertex-stage --no-c4-slice --no-stream-mode --steps 20 --warmup 5"
for R in 1 2; do
  for V in 1 0; do
    SPR_SMALL_SPLIT=$V timeout 300 python bench.py --workload c2 $B > $O/r2ba_c2_split${V}_$R.json 2> $O/r2ba_c2_split${V}_$R.err
    python -c "import json;d=json.load(open('$O/r2ba_c2_split${V}_$R.json'));print('c2 split=$V run $R: %.4f ms'%d['ms_per_step'], ['%.4f'%x for x in d['repeats']['ms_per_step']], 'k_sort_small %.4f'%d['kernels']['k_sort_small']['ms_per_launch'])"
  done
done
