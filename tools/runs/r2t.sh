#!/bin/bash
# round 2, call T: depth buckets after the first tuning (LUT in the histogram, 128-thread scatter), cost of the radix launches behind them
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_more.py tests/test_gpu_full_size.py -m gpu -x -q -k "sort or full_size or large or c3 or c4 or tile" > $O/r2t_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2t_pytest.log; tail -5 $O/r2t_pytest.log
B="python bench.py --workload c3 --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5"
timeout 600 $B > $O/r2t_c3_buckets.json 2> $O/r2t_c3_buckets.err; echo "buckets exit $?"; python tools/bench_summary.py < $O/r2t_c3_buckets.json; tail -3 $O/r2t_c3_buckets.err
timeout 600 $B --no-kernel-timing > $O/r2t_c3_buckets_nt.json 2> $O/r2t_c3_buckets_nt.err; echo "buckets (no events) exit $?"; python -c "
import json,sys; d=json.loads(open('$O/r2t_c3_buckets_nt.json').read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], d.get('repeats'))"
