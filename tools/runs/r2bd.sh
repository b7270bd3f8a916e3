#!/bin/bash
# round 2, call BD: k_expand_sorted with more resident warps -- 2 records per thread (512 threads, 3 CTAs per SM, 40 registers) and 1 record
# per thread (1024 threads, 2 CTAs, 32 registers) against the product's 4 records per thread (256 threads, 64 registers), alternating
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5"
for V in e2c3 e1c2; do
  SPR_LIB_VARIANT=$V timeout 600 python -m pytest tests/test_gpu_more.py tests/test_gpu_parity.py -m gpu -x -q -k "sort or large or parity" > $O/r2bd_pytest_$V.log 2>&1; echo "pytest $V exit $?"; tail -1 $O/r2bd_pytest_$V.log
done
for R in 1 2; do
  for W in c2 c4 c3; do
    for V in "" e2c3 e1c2; do
      SPR_LIB_VARIANT=$V timeout 300 python bench.py --workload $W $B > $O/r2bd_${W}_${V:-base}_$R.json 2> $O/r2bd_${W}_${V:-base}_$R.err
      python -c "import json;d=json.load(open('$O/r2bd_${W}_${V:-base}_$R.json'));print('$W ${V:-base} run $R: %.4f ms'%d['ms_per_step'], 'k_expand_sorted %.4f'%d['kernels']['k_expand_sorted']['ms_per_launch'])"
    done
  done
done
