"""GPU parity against the golden vectors generated from the reference itself (tests/golden/):
this is what pins the CUDA path on the GPU box, where /root/reference does not exist."""
import hashlib
import json
import os

import pytest

import scenarios as S
from spear_b200 import abi

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    from spear_b200 import particles
    c = particles.ParticleContext(device=0, initial_slot_capacity=1 << 16)
    yield c
    c.close()


@pytest.mark.parametrize("sc", S.scenarios(), ids=lambda s: s.name)
def test_scenario_matches_reference_golden(ctx, sc):
    g = json.load(open(os.path.join(GOLDEN, "scenarios.json")))[sc.name]
    s = ctx.create_system(sc.seed)
    try:
        snaps = S.run_scenario(s, sc)
    finally:
        s.close()
    for f, (snap, ref) in enumerate(zip(snaps, g["frames"])):
        assert S.snapshot_digest(snap) == ref, "%s frame %d" % (sc.name, f)


def test_c1_full_600_steps(ctx):
    """BASELINE configs[0] against the reference's end state (SURVEY.md Appendix C known answers)."""
    g = json.load(open(os.path.join(GOLDEN, "c1_full.json")))
    s = ctx.create_system((1, 2, 3, 4))
    e = s.add_effect(S.comp_c1(), abi.make_transform((1, 2, 3)))
    dt = s.effect_info(e).timeStep
    for i in range(600):
        s.update([e], abi.make_frame_args(dt, game_time=i * dt, frame_index=i))
    info = s.effect_info(e)
    assert (info.slotCount, info.aliveCount, info.freeCount) == (g["slotCount"], g["aliveCount"], g["freeCount"]) == (13224, 4012, 9212)
    assert "%016x" % s.rng_state()[0] == g["rng"]
    assert hashlib.sha256(s.read_stream(e).tobytes()).hexdigest() == g["stream"]
    assert hashlib.sha256(s.read_free(e).tobytes()).hexdigest() == g["free"]
    s.close()


def test_render_matches_reference_golden(ctx):
    from golden.make_golden import render_scene
    g = json.load(open(os.path.join(GOLDEN, "render.json")))
    s = ctx.create_system((1, 5, 3, 4))
    assert render_scene(s) == g
    s.close()
