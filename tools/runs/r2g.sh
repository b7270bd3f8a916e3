#!/bin/bash
# round 2, call G: GPU suite with ticketed tile order + argument refusal tests; timing-only experiment: the sorted
# expansion of frame k on a second stream under frame k+1 (SPR_EXP_OVERLAP=1 low priority, =2 high priority)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > $O/r2g_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2g_pytest.log
tail -12 $O/r2g_pytest.log
B="--no-cpu-baseline --no-stream-mode --no-billboards --no-vertex-stage --steps 50 --warmup 3"
for W in c3 c4; do
  for X in 0 1 2; do
    SPR_EXP_OVERLAP=$X timeout 300 python bench.py --workload $W $B --no-kernel-timing > $O/r2g_${W}_x$X.json 2> $O/r2g_${W}_x$X.err
    echo "$W overlap=$X: $(python -c "import json,sys; d=json.loads(open('$O/r2g_${W}_x$X.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['repeats']['ms_per_step'])")"
  done
done
timeout 300 python bench.py --workload c3 --no-cpu-baseline --no-billboards --no-vertex-stage --steps 20 --warmup 3 > $O/r2g_c3_full.json 2> $O/r2g_c3_full.err
python tools/bench_summary.py < $O/r2g_c3_full.json
