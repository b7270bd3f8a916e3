#!/bin/bash
# round 2, call AP: deferred expansion (sort without the local 4x stream, pack from the slot state) -- its test, the full GPU suite
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_more.py -m gpu -x -q -k "deferred or pack_and_fused" > $O/r2ap_pytest_def.log 2>&1; echo "pytest exit $?" >> $O/r2ap_pytest_def.log; tail -15 $O/r2ap_pytest_def.log
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r2ap_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2ap_pytest.log; tail -4 $O/r2ap_pytest.log
