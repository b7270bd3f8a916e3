"""CPU: the numpy restatement of sf::Frustum::intersects(Sphere) over a flat sphere list (oracle/areas.py) against the
reference's own AreaSystem.cpp (oracle/_ref/libspear_ref_area.so, when built: spatial tree, optimize passes) and
against the golden vectors generated from it (tests/golden/areas.json) -- SURVEY.md 8(f) row N5."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import areas as A
from spear_b200 import abi
from tests_render_args import area_test_frusta, make_render_args

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "areas.json")))


@pytest.mark.parametrize("scene", GOLDEN["scenes"], ids=lambda s: "n%d" % s["n"])
def test_port_matches_golden(scene):
    res = A.run_scene_port(A.random_scene(scene["seed"], scene["n"]), area_test_frusta())
    assert [len(r) for r in res] == scene["counts"]
    assert [[int(x) for x in r[:8]] for r in res] == scene["first_ids"]
    assert [hashlib.sha256(r.astype("<u4").tobytes()).hexdigest() for r in res] == scene["sha256"]


@pytest.mark.skipif(not A.have_reference(), reason="oracle/_ref/libspear_ref_area.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed,n", [(21, 0), (22, 3), (23, 150), (24, 5000)])
def test_port_matches_reference_build(seed, n):
    ops = A.random_scene(seed, n, rounds=6)
    frusta = area_test_frusta()
    ref = A.run_scene_reference(ops, frusta)
    port = A.run_scene_port(ops, frusta)
    assert len(ref) == len(port)
    for q, (a, b) in enumerate(zip(ref, port)):
        assert np.array_equal(a, b), "query %d" % q


@pytest.mark.skipif(not A.have_reference(), reason="oracle/_ref/libspear_ref_area.so not built (needs /root/reference)")
def test_reference_is_a_subset_outside_the_slab():
    # spheres that reach outside AreaMinY..AreaMaxY: the reference's clamped tree boxes may drop some that pass the
    # sphere test, never the other way round
    ops = A.random_scene(31, 3000, in_slab=False)
    frusta = area_test_frusta()
    dropped = 0
    for a, b in zip(A.run_scene_reference(ops, frusta), A.run_scene_port(ops, frusta)):
        assert np.all(np.isin(a, b))
        dropped += len(b) - len(a)
    assert dropped < 100   # and it is rare (measured: a handful per query at this size)


def test_port_known_answers():
    fr = make_render_args().frustum   # camera (16, 20, -30) looking at (14, 0, 14)
    ids = [0, 1, 2, 3, 4]
    origins = [(14.0, 0.0, 14.0), (16.0, 20.0, -31.0), (16.0, 20.0, -31.0), (14.0, 0.0, 14.0), (1e4, 0.0, 0.0)]
    radii = [0.0, 0.5, 2.0, float("nan"), 100.0]
    # the look-at point is inside; a point 1 m behind the camera is outside with r = 0.5 and reaches in with r = 2;
    # a NaN radius fails every plane; a far sphere off to the side is outside
    assert A.port_query(fr, ids, origins, radii).tolist() == [0, 2]
    # a plane exactly through the origin: distance + r == 0 is NOT inside (strictly greater, Frustum.h:57)
    f = abi.Frustum()
    for j in range(4):
        f.side[0][j] = 1.0
        f.caps[0][j] = 1.0
    assert A.port_query(f, [7, 8, 9], [(0, 0, 0), (0, 5, 5), (-1, 0, 0)], [0.0, 1e-30, 1.0]).tolist() == [8]
