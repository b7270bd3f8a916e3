#!/bin/bash
# round 2, call BF: the one-granule (TINY) instantiation of k_integrate_pass at 4 CTAs per SM (64 registers, 52 B spilled) against the product's 3 (80 registers), C5, alternating
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 5"
SPR_LIB_VARIANT=tiny4 timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py -m gpu -x -q -k "c5 or parity" > $O/r2bf_pytest.log 2>&1; echo "pytest exit $?"; tail -1 $O/r2bf_pytest.log
for R in 1 2; do
  for V in "" tiny4; do
    SPR_LIB_VARIANT=$V timeout 300 python bench.py --workload c5 $B > $O/r2bf_c5_${V:-base}_$R.json 2> $O/r2bf_c5_${V:-base}_$R.err
    python -c "import json;d=json.load(open('$O/r2bf_c5_${V:-base}_$R.json'));print('c5 ${V:-base} run $R: %.4f ms'%d['ms_per_step'], 'k_integrate %.4f'%d['kernels']['k_integrate']['ms_per_launch'])"
  done
done
