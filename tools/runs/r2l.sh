#!/bin/bash
# round 2, call L: k_stream_fused consumes unowned tiles at claim time; stream-mode numbers of every workload; ncu A/B of
# phase B (bulk tensor stores against the LSU loop) on the 10M-particle effect
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2l_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2l_pytest.log; tail -3 $O/r2l_pytest.log
B="--no-sort --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 3"
for W in c3 c4 c2 c5; do
  timeout 300 python bench.py --workload $W $B > $O/r2l_${W}_stream.json 2> $O/r2l_${W}_stream.err
  echo "== $W stream mode"; python tools/bench_summary.py < $O/r2l_${W}_stream.json | grep -E "particle-steps|k_stream_fused"
done
for T in 1 0; do
  SPR_STREAM_TMA=$T timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_fused -s 6 -c 1 -o $O/r2l_c3_stream_fused_tma$T -f python bench.py --workload c3 $B --steps 3 --no-kernel-timing > $O/r2l_ncu_tma$T.log 2>&1
  echo "ncu tma=$T exit $?"
done
