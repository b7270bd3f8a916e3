"""SURVEY.md section 8(f) row N1, the literal drop-in: integration/ParticleSystemB200.cpp implements the
reference's abstract class cl::ParticleSystem (src/client/ParticleSystem.h:9-22) over the C ABI. It is
compiled against the reference's own headers (oracle/build_ref.py -> oracle/_ref/libspear_adapter.so) and
driven here through the reference's own types -- sf::Box<sv::ParticleSystemComponent>, cl::Transform,
cl::VisibleAreas, cl::FrameArgs, cl::RenderArgs, the sg_* draw calls -- by the same harness that drives
the reference implementation. Expected values: the golden vectors generated from the reference."""
import json
import os

import pytest

import scenarios as S
from oracle import orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def adapter():
    if not orc.have("adapter"):
        pytest.skip("oracle/_ref/libspear_adapter.so is built only where the reference tree is available")
    return "adapter"


@pytest.mark.parametrize("sc", S.scenarios(), ids=lambda s: s.name)
def test_adapter_scenario_matches_reference_golden(adapter, sc):
    g = json.load(open(os.path.join(GOLDEN, "scenarios.json")))[sc.name]
    o = orc.OracleSystem(adapter, sc.seed)
    assert o.h, "adapter could not create its context (no GPU?)"
    try:
        snaps = S.run_scenario(o, sc)
    finally:
        o.close()
    for f, (snap, ref) in enumerate(zip(snaps, g["frames"])):
        assert S.snapshot_digest(snap) == ref, "%s frame %d" % (sc.name, f)


def test_adapter_render_main_matches_reference_golden(adapter):
    from golden.make_golden import render_scene
    g = json.load(open(os.path.join(GOLDEN, "render.json")))
    o = orc.OracleSystem(adapter, (1, 5, 3, 4))
    try:
        assert render_scene(o) == g
    finally:
        o.close()


def test_adapter_renderer_side_matches_reference_capture(adapter):
    """Everything the adapter hands to sokol, against the reference's own calls captured by the same harness stubs
    (tests/golden/render_capture.json): the bytes appended to the 16-deep vertex-buffer ring and the offsets returned
    (ParticleSystem.cpp:861-867), the buffer / offset / index buffer / images bound for every draw (:892-899), the
    draw sizes, the spline atlas texels after initEffectType's sg_make_image + sg_bqq_copy_subimages (:421-449, gradient
    alpha bytes masked: uninitialised in the reference) and the sphere areas after updateTransform (:782-783)."""
    from golden.make_golden import render_capture_scene
    g = json.load(open(os.path.join(GOLDEN, "render_capture.json")))
    o = orc.OracleSystem(adapter, (1, 6, 3, 4))
    try:
        got = render_capture_scene(o)
    finally:
        o.close()
    assert got["atlasCells"] == g["atlasCells"] and got["atlas"] == g["atlas"]
    assert got["spheres"] == g["spheres"]
    for f, (a, b) in enumerate(zip(g["frames"], got["frames"])):
        assert a["records"] == b["records"], "frame %d draw records" % f
        assert a["appends"] == b["appends"], "frame %d sg_append_buffer calls" % f
        assert a["draws"] == b["draws"], "frame %d bindings" % f
    assert len(g["frames"]) == len(got["frames"]) == 40
