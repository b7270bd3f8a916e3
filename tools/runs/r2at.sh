#!/bin/bash
# round 2, call AT: host-list reads ahead of the wait in k_emit_warp / k_finalize / k_sort_small -- GPU suite, C2 / C5 / C3 timing
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2at_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2at_pytest.log; tail -4 $O/r2at_pytest.log
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --no-kernel-timing --steps 20 --warmup 5"
for W in c2 c5 c3; do
  timeout 300 python bench.py --workload $W $B > $O/r2at_$W.json 2> $O/r2at_$W.err
  python -c "import json;d=json.load(open('$O/r2at_$W.json'));print('$W: %.4f ms'%d['ms_per_step'], ['%.4f'%x for x in d['repeats']['ms_per_step']])"
done
