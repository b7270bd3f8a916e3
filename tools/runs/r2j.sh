#!/bin/bash
# round 2, call J: k_stream_fused phase B through bulk tensor stores (TMA, default) against the LSU loop (SPR_STREAM_TMA=0)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2j_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2j_pytest.log; tail -3 $O/r2j_pytest.log
B="--no-sort --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --steps 20 --warmup 3"
for W in c3 c4 c2 c5; do
  for T in 1 0; do
    SPR_STREAM_TMA=$T timeout 300 python bench.py --workload $W $B > $O/r2j_${W}_tma$T.json 2> $O/r2j_${W}_tma$T.err
    echo "== $W stream mode, TMA=$T"; python tools/bench_summary.py < $O/r2j_${W}_tma$T.json | grep -E "particle-steps|k_stream_fused"
  done
done
