#!/bin/bash
# round 2, call AL: L2 persistence window over the slot state (SPR_L2_PERSIST_MB) on the sorted C3 step, A/B
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 600 python bench.py --workload c3 --no-cpu-baseline --no-billboards --no-vertex-stage --no-c4-slice --no-stream-mode --steps 20 --warmup 5 > $O/r2al_$tag.json 2> $O/r2al_$tag.err
  echo "== $tag exit $?"; grep "L2 window" $O/r2al_$tag.err | head -1
  python tools/bench_summary.py < $O/r2al_$tag.json | egrep "ms/step|k_integrate|k_expand_sorted|k_sort_pass"
}
run off SPR_L2_PERSIST_MB=0
run p40 SPR_L2_PERSIST_MB=40
run p80 SPR_L2_PERSIST_MB=80
run p80s SPR_L2_PERSIST_MB=80 SPR_L2_PERSIST_MISS=s
run p60w320 SPR_L2_PERSIST_MB=60 SPR_L2_WINDOW_MB=320
run p120 SPR_L2_PERSIST_MB=120
