#!/bin/bash
# round 2, call S: first run of the depth-bucket path -- sort tests, then the C3 sorted step with and without it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_more.py tests/test_gpu_full_size.py tests/test_gpu_parity.py -m gpu -x -q -k "sort or full_size or large or c3 or c4 or tile" > $O/r2s_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2s_pytest.log; tail -15 $O/r2s_pytest.log
B="python bench.py --workload c3 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5"
SPR_DEPTH_BUCKETS=0 timeout 600 $B > $O/r2s_c3_radix.json 2> $O/r2s_c3_radix.err; echo "radix exit $?"; python tools/bench_summary.py < $O/r2s_c3_radix.json
timeout 600 $B > $O/r2s_c3_buckets.json 2> $O/r2s_c3_buckets.err; echo "buckets exit $?"; python tools/bench_summary.py < $O/r2s_c3_buckets.json; tail -3 $O/r2s_c3_buckets.err
