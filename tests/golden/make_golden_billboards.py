#!/usr/bin/env python3
"""Generate tests/golden/billboards.json from the reference's own cl::BillboardSystem
(oracle/_ref/libspear_ref_billboard.so). Needs /root/reference to build that library; the JSON travels."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import billboards as B  # noqa: E402
from tests_render_args import make_render_args  # noqa: E402

FRAMES = [(1, 17), (2, 300), (3, 1024), (4, 1025), (5, 4096), (6, 5000), (7, 20000)]


def main():
    out = {"frames": []}
    o = B.BillboardOracle()
    ra = make_render_args()
    for seed, n in FRAMES:
        o.add(B.random_frame(seed, n))
        verts, draws, w2c = o.render(ra)
        out["frames"].append({"seed": seed, "n": n, "quads": len(verts) // 4, "draws": [list(d) for d in draws],
                              "vertices_sha256": hashlib.sha256(verts.tobytes()).hexdigest(),
                              "first_vertices": [[float(x) for x in v["position"]] + [float(x) for x in v["texCoord"]] + [int(v["color"])] for v in verts[:8]]})
    o.close()
    json.dump(out, open(os.path.join(HERE, "billboards.json"), "w"), indent=1)
    print("wrote billboards.json:", [(f["n"], f["quads"], len(f["draws"])) for f in out["frames"]])


if __name__ == "__main__":
    main()
