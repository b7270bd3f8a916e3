#!/bin/bash
# round 2, call AJ: vertex stage on GL's LINEAR filter formula, pinned against the reference's shipped GLSL executed on the CPU
# (oracle/_ref/libspear_ref_glslvs.so) -- GPU suite, then the default bench (k_shade_vertices timing in "vertex_stage")
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2aj_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2aj_pytest.log; tail -6 $O/r2aj_pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-billboards --no-c4-slice --no-stream-mode --steps 20 --warmup 5 > $O/r2aj_c3.json 2> $O/r2aj_c3.err; echo "c3 exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2aj_c3.json'))
print(d['ms_per_step'], json.dumps(d.get('vertex_stage')))
P
