#!/bin/bash
# round 2, call H: tile = block index for the one-CTA-per-tile chained scans (default) against tickets everywhere
# (libspear_particles.strict.so = -DSPR_STRICT_TILE_TICKETS=1); k_stream_fused is ticketed in both
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="--no-cpu-baseline --no-billboards --no-vertex-stage --steps 20 --warmup 3"
timeout 900 python -m pytest tests/test_gpu_more.py tests/test_gpu_parity.py -m gpu -x -q > $O/r2h_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2h_pytest.log; tail -3 $O/r2h_pytest.log
for W in c3 c4 c2; do
  for V in default strict; do
    if [ $V = strict ]; then export SPR_LIB_VARIANT=strict; else unset SPR_LIB_VARIANT; fi
    timeout 300 python bench.py --workload $W $B > $O/r2h_${W}_$V.json 2> $O/r2h_${W}_$V.err
    echo "== $W $V"; python tools/bench_summary.py < $O/r2h_${W}_$V.json | head -12
    python -c "import json; d=json.loads(open('$O/r2h_${W}_$V.json').read().strip().splitlines()[-1]); print('stream mode ms', d.get('stream_mode',{}).get('ms_per_step'))"
  done
done
