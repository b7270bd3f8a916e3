#!/bin/bash
# round 2, call E: full GPU suite (adapter renderer side, DrawRecord upload fields)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2e_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2e_pytest.log
tail -15 $O/r2e_pytest.log
