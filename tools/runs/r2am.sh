#!/bin/bash
# round 2, call AM: timing-only -- the radix passes with the chained look-back skipped (upper bound of what a cheaper look-back buys)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5"
for W in c3 c4; do
  for V in "" nolb; do
    SPR_LIB_VARIANT=$V timeout 300 python bench.py --workload $W $B > $O/r2am_${W}_${V:-base}.json 2> $O/r2am_${W}_${V:-base}.err; echo "== $W ${V:-base} exit $?"
    python tools/bench_summary.py < $O/r2am_${W}_${V:-base}.json | egrep "ms/step|k_sort|k_expand_sorted"
  done
done
