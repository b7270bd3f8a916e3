import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
frac = d["roofline"]["step"]["frac"] if d.get("roofline") else float("nan")
print("G particle-steps/s %.3f  ms/step %.4f  step frac %.3f  e2e %.3f" % (d["value"] / 1e9, d["ms_per_step"], frac, d["e2e"]["value"] / 1e9))
for k, v in d["kernels"].items():
    print("  %-16s %.4f ms x %d  %s" % (k, v["ms_per_launch"], v["launches"] // d["steps"], ("%.0f GB/s" % v["achieved_gbs"]) if "achieved_gbs" in v else ""))
