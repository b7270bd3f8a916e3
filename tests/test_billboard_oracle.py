"""CPU: the numpy restatement of cl::BillboardSystem::renderMain (oracle/billboards.py) against the reference's own
BillboardSystem.cpp (oracle/_ref/libspear_ref_billboard.so, when built) and against the golden vectors generated
from it (tests/golden/billboards.json) -- SURVEY.md 8(f) row N4."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import billboards as B
from spear_b200 import abi
from tests_render_args import make_render_args

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "billboards.json")))


@pytest.mark.parametrize("frame", GOLDEN["frames"], ids=lambda f: "n%d" % f["n"])
def test_port_matches_golden(frame):
    verts, draws = B.port_render(B.random_frame(frame["seed"], frame["n"]))
    assert len(verts) // 4 == frame["quads"]
    assert [list(d) for d in draws] == frame["draws"]
    assert hashlib.sha256(verts.tobytes()).hexdigest() == frame["vertices_sha256"]


@pytest.mark.skipif(not B.have_reference(), reason="oracle/_ref/libspear_ref_billboard.so not built (needs /root/reference)")
def test_port_matches_reference_build_frame_by_frame():
    o = B.BillboardOracle()
    ra = make_render_args()
    try:
        for seed, n in [(11, 1), (12, 50), (13, 700), (14, 4500), (15, 0)]:
            frame = B.random_frame(seed, n)
            o.add(frame)
            verts, draws, w2c = o.render(ra)
            pv, pd = B.port_render(frame)
            assert np.array_equal(verts.view(np.uint8), pv.view(np.uint8)) if len(pv) else len(verts) == 0
            assert draws == pd
            if len(verts):   # the uniform block is only applied when something is drawn (:205-210)
                assert w2c == list(ra.worldToClip.v)
        # the frame's billboards are consumed by renderMain (:227-228)
        verts, draws, _ = o.render(ra)
        assert len(verts) == 0 and draws == []
    finally:
        o.close()


def test_port_known_answers():
    # identity transform, anchor 0.5, full crop: a unit quad centred on the origin, uv = vert * scale + bias
    b = abi.make_billboard(atlas=5, uv_scale=(0.5, 0.25), uv_bias=(0.25, 0.5), color=(1.0, 0.5, 0.0, 2.0), depth=1.0)
    v, d = B.port_render([b])
    assert d == [(5, 1, 0)]
    assert [tuple(x) for x in v["position"]] == [(-0.5, -0.5, 0.0), (0.5, -0.5, 0.0), (-0.5, 0.5, 0.0), (0.5, 0.5, 0.0)]
    assert [tuple(x) for x in v["texCoord"]] == [(0.25, 0.5), (0.75, 0.5), (0.25, 0.75), (0.75, 0.75)]
    assert int(v["color"][0]) == 255 | 127 << 8 | 0 << 16 | 255 << 24          # (uint32)(f * 255), clamped to 255
    # sort by depth, ties (and -0 / +0) by insertion order; the budget cuts the far end
    bbs = [abi.make_billboard(atlas=a, depth=z) for a, z in [(1, 2.0), (2, 0.0), (3, -0.0), (4, -1.0), (5, 2.0)]]
    v, d = B.port_render(bbs)
    assert [x[0] for x in d] == [4, 2, 3, 1, 5]
    v, d = B.port_render(bbs, max_per_frame=3)
    assert [x[0] for x in d] == [4, 2, 3]
    # an empty crop and a null sprite produce nothing; equal neighbouring atlases merge into one draw
    bbs = [abi.make_billboard(atlas=9, depth=0), abi.make_billboard(atlas=9, depth=1, crop_min=(0.5, 0), crop_max=(0.5, 1)),
           abi.make_billboard(atlas=0, depth=2), abi.make_billboard(atlas=9, depth=3)]
    v, d = B.port_render(bbs)
    assert d == [(9, 2, 0)] and len(v) == 8
