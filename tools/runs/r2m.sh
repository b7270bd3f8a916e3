#!/bin/bash
# round 2, call M (2 GPUs): peer-memory pull + 4x expansion micro-benchmark; stream-mode check of the scan-warp claim
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 200 ./tools/peer_expand_bench > $O/r2m_peer_2gpu.txt 2>&1; cat $O/r2m_peer_2gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_full_size.py -m gpu -x -q > $O/r2m_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2m_pytest.log; tail -3 $O/r2m_pytest.log
B="--no-sort --no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --no-c4-slice --steps 20 --warmup 3"
for W in c3 c4 c5; do
  timeout 300 python bench.py --workload $W $B > $O/r2m_${W}_stream.json 2> $O/r2m_${W}_stream.err
  echo "== $W stream mode"; python tools/bench_summary.py < $O/r2m_${W}_stream.json | grep -E "particle-steps|k_stream_fused|k_finalize|k_plan"
done
