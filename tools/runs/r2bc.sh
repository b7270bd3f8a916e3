#!/bin/bash
# round 2, call BC: ncu --set full of k_sort_small (two CTAs per effect, population cuts) and k_expand_sorted in a C2 frame, with the one-CTA build of the same kernel beside it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
CMD="python bench.py --workload c2 --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --no-kernel-timing"
timeout 300 $CMD > $O/r2bc_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sort_small|k_expand_sorted' -s 6 -c 2 -o $O/r2bc_c2_split -f $CMD > $O/r2bc_ncu_split.log 2>&1
echo "ncu split exit $?"
SPR_SMALL_SPLIT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sort_small' -s 3 -c 1 -o $O/r2bc_c2_nosplit -f $CMD > $O/r2bc_ncu_nosplit.log 2>&1
echo "ncu nosplit exit $?"
