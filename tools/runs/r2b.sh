#!/bin/bash
# round 2, call B: per-phase cycle profile of k_sort_pass (instrumented build, -DSPR_SORT_PROFILE)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-kernel-timing --steps 20 --warmup 3"
for R in 1 2; do
  SPR_LARGE_SLOTS=0 SPR_SORT_RANK=$R timeout 300 python bench.py --workload c3 $B > $O/r2b_c3_rank$R.json 2> $O/r2b_c3_rank$R.err
  grep "sort profile" $O/r2b_c3_rank$R.err
done
SPR_LARGE_SLOTS=0 SPR_SORT_RANK=1 timeout 300 python bench.py --workload c4 $B > $O/r2b_c4_rank1.json 2> $O/r2b_c4_rank1.err
grep "sort profile" $O/r2b_c4_rank1.err
