#!/bin/bash
# round 2, call AO: radix pass CTAs with 8 keys per thread (512 threads per 4096-key tile, 3 or 4 CTAs per SM) and with
# 16 keys per thread at 5 CTAs per SM (48 registers) against the product (16 keys, 4 CTAs, 64 registers)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5"
for V in i8c3 i8c4 i16c5; do
  SPR_LIB_VARIANT=$V timeout 600 python -m pytest tests/test_gpu_more.py tests/test_gpu_parity.py -m gpu -x -q -k "sort or large or parity" > $O/r2ao_pytest_$V.log 2>&1; echo "pytest $V exit $?"; tail -1 $O/r2ao_pytest_$V.log
done
for W in c3 c4; do
  for V in "" i8c3 i8c4 i16c5; do
    SPR_LIB_VARIANT=$V timeout 300 python bench.py --workload $W $B > $O/r2ao_${W}_${V:-base}.json 2> $O/r2ao_${W}_${V:-base}.err; echo "== $W ${V:-base} exit $?"
    python tools/bench_summary.py < $O/r2ao_${W}_${V:-base}.json | egrep "ms/step|k_sort_pass"
  done
done
