#!/bin/bash
# round 2, call AK: k_shade_vertices with the division-free UNORM8 conversion and the zero-row-weight fast path
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_vertex_stage.py -m gpu -x -q > $O/r2ak_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2ak_pytest.log; tail -4 $O/r2ak_pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-billboards --no-c4-slice --no-stream-mode --steps 20 --warmup 5 > $O/r2ak_c3.json 2> $O/r2ak_c3.err; echo "c3 exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2ak_c3.json'))
print(d['ms_per_step'], json.dumps(d.get('vertex_stage')))
P
