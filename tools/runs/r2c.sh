#!/bin/bash
# round 2, call C: 4096- against 8192-key sort tiles (SPR_LIB_VARIANT builds), sort tests on both, phase profiles
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --steps 20 --warmup 3"
timeout 600 python -m pytest tests/test_gpu_more.py -m gpu -x -q -k "sorted or sort" > $O/r2c_pytest_4096.log 2>&1; echo "exit $?" >> $O/r2c_pytest_4096.log
SPR_LIB_VARIANT=t8192 timeout 600 python -m pytest tests/test_gpu_more.py tests/test_gpu_parity.py -m gpu -x -q > $O/r2c_pytest_8192.log 2>&1; echo "exit $?" >> $O/r2c_pytest_8192.log
tail -2 $O/r2c_pytest_4096.log $O/r2c_pytest_8192.log
for W in c3 c4; do
  timeout 300 python bench.py --workload $W $B > $O/r2c_${W}_4096.json 2> $O/r2c_${W}_4096.err
  SPR_LIB_VARIANT=t8192 timeout 300 python bench.py --workload $W $B > $O/r2c_${W}_8192.json 2> $O/r2c_${W}_8192.err
done
for V in t4096p t8192p; do
  SPR_LIB_VARIANT=$V timeout 300 python bench.py --workload c3 $B --no-kernel-timing > $O/r2c_c3_$V.json 2> $O/r2c_c3_$V.err
  grep "sort profile" $O/r2c_c3_$V.err
done
for f in $O/r2c_c?_*96.json $O/r2c_c?_8192.json; do echo $f; python tools/bench_summary.py < $f 2>/dev/null | grep -E "particle-steps|sort_|expand"; done
