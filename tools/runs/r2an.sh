#!/bin/bash
# round 2, call AN: k_sort_small with the MSD shortcut (top-8-bit ranking pass + one warp per depth bucket) -- sort tests, full suite, C2 A/B
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_more.py -m gpu -x -q -k "small_effects or c2_full or pass_skipping or sorted_stream" > $O/r2an_pytest_sort.log 2>&1; echo "pytest exit $?" >> $O/r2an_pytest_sort.log; tail -5 $O/r2an_pytest_sort.log
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2an_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2an_pytest.log; tail -4 $O/r2an_pytest.log
B="--no-cpu-baseline --no-billboards --no-vertex-stage --no-c4-slice --no-stream-mode --steps 20 --warmup 5"
for V in 1 0; do
  SPR_SMALL_MSD=$V timeout 300 python bench.py --workload c2 $B > $O/r2an_c2_msd$V.json 2> $O/r2an_c2_msd$V.err; echo "== c2 msd=$V exit $?"
  python tools/bench_summary.py < $O/r2an_c2_msd$V.json
done
