#!/bin/bash
# round 2, call AE: packed / device billboard input -- billboard tests, then the billboard bench leg
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_billboards.py -m gpu -x -q > $O/r2ae_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2ae_pytest.log; tail -6 $O/r2ae_pytest.log
timeout 900 python bench.py --workload c2 --no-stream-mode --no-vertex-stage --no-c4-slice --steps 5 --warmup 3 --repeats 1 > $O/r2ae_bench.json 2> $O/r2ae_bench.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('$O/r2ae_bench.json').read().strip().splitlines()[-1]); print(json.dumps(d.get('billboards'), indent=1))"
