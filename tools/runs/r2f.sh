#!/bin/bash
# round 2, call F: GPU suite with the full-size goldens, smoke(), bench JSON sanity (median of 5 regions), reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > $O/r2f_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2f_pytest.log
tail -16 $O/r2f_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; echo "smoke exit $?" >> $O/r2f_smoke.log; tail -3 $O/r2f_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2f_bench_default.json 2> $O/r2f_bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2f_bench_ref.json 2> $O/r2f_bench_ref.err; echo "ref exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f_bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "repeats", "gpu_launches")}); print(d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["step"]["frac"], d["cpu_baseline"]["value"], d.get("stream_mode", {}).get("value"))
r = json.loads(open("gpurun_out/r2f_bench_ref.json").read().strip().splitlines()[-1]); print(r["value"], r["cpu_baseline"])
PY
