#!/bin/bash
# round 2, call A: GPU suite on the new sort passes / scatter expansion, GL probe, C3 A/B runs
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
python tools/probe_gl.py > $O/r2a_probe_gl.json 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2a_pytest.log
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --steps 20 --warmup 3"
timeout 300 python bench.py --workload c3 $B > $O/r2a_c3_default.json 2> $O/r2a_c3_default.err
SPR_SORT_RANK=1 timeout 300 python bench.py --workload c3 $B > $O/r2a_c3_rank1.json 2> $O/r2a_c3_rank1.err
SPR_LARGE_SLOTS=0 timeout 300 python bench.py --workload c3 $B > $O/r2a_c3_gather.json 2> $O/r2a_c3_gather.err
SPR_LARGE_SLOTS=0 SPR_SORT_RANK=1 timeout 300 python bench.py --workload c3 $B > $O/r2a_c3_gather_rank1.json 2> $O/r2a_c3_gather_rank1.err
timeout 300 python bench.py --workload c4 $B > $O/r2a_c4_default.json 2> $O/r2a_c4_default.err
SPR_SORT_RANK=1 timeout 300 python bench.py --workload c4 $B > $O/r2a_c4_rank1.json 2> $O/r2a_c4_rank1.err
timeout 300 python bench.py --workload c2 $B > $O/r2a_c2_default.json 2> $O/r2a_c2_default.err
tail -3 $O/r2a_pytest.log
for f in $O/r2a_c*.json; do echo $f; python tools/bench_summary.py < $f 2>/dev/null | head -20; done
