#!/usr/bin/env python3
"""Generate tests/golden/areas.json from the reference's own cl::AreaSystem
(oracle/_ref/libspear_ref_area.so). Needs /root/reference to build that library; the JSON travels."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import areas as A  # noqa: E402
from tests_render_args import area_test_frusta  # noqa: E402

SCENES = [(1, 1), (2, 40), (3, 333), (4, 2000), (5, 20000)]


def main():
    frusta = area_test_frusta()
    out = {"scenes": []}
    for seed, n in SCENES:
        res = A.run_scene_reference(A.random_scene(seed, n), frusta)
        out["scenes"].append({"seed": seed, "n": n, "counts": [int(len(r)) for r in res],
                              "sha256": [hashlib.sha256(r.astype("<u4").tobytes()).hexdigest() for r in res],
                              "first_ids": [[int(x) for x in r[:8]] for r in res]})
    json.dump(out, open(os.path.join(HERE, "areas.json"), "w"), indent=1)
    print("wrote areas.json:", [(s["n"], s["counts"]) for s in out["scenes"]])


if __name__ == "__main__":
    main()
