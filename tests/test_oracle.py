"""CPU tests of the oracle itself (no GPU): the plain-C restatement against the golden vectors
generated from the reference (tests/golden/make_golden.py) and, when oracle/_ref can be built
here, against the reference's own code frame by frame."""
import hashlib
import json
import os

import numpy as np
import pytest

import scenarios as S
from oracle import orc
from spear_b200 import abi
from tests_render_args import make_render_args

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.mark.parametrize("sc", S.scenarios(), ids=lambda s: s.name)
def test_port_matches_golden_scenarios(oracle_port, sc):
    g = golden("scenarios.json")[sc.name]
    snaps = S.run_scenario(orc.OracleSystem(oracle_port, sc.seed), sc)
    assert len(snaps) == len(g["frames"])
    for f, (snap, ref) in enumerate(zip(snaps, g["frames"])):
        assert S.snapshot_digest(snap) == ref, "%s frame %d" % (sc.name, f)


def test_port_matches_golden_c1_full(oracle_port):
    """BASELINE configs[0]: single emitter, 10k burst, 600 fixed-dt steps."""
    g = golden("c1_full.json")
    o = orc.OracleSystem(oracle_port, (1, 2, 3, 4))
    e = o.add_effect(S.comp_c1(), abi.make_transform((1, 2, 3)))
    info = o.effect_info(e)
    assert S.f32_bits(info.timeStep) == g["timeStepBits"] == 0x3c888889
    dt, total = info.timeStep, 0
    for i in range(600):
        o.update([e], abi.make_frame_args(dt, game_time=i * dt, frame_index=i))
        total += o.effect_info(e).aliveCount
    info = o.effect_info(e)
    assert (info.slotCount, info.aliveCount, info.freeCount, total) == (13224, 4012, 9212, 4799013)  # Appendix C
    assert "%016x" % o.rng_state()[0] == g["rng"] == "3508c3884b97803a"
    assert hashlib.sha256(o.read_stream(e).tobytes()).hexdigest() == g["stream"]
    assert hashlib.sha256(o.read_free(e).tobytes()).hexdigest() == g["free"]


def test_port_types_match_golden(oracle_port):
    g = golden("types.json")
    comps = dict(c1=S.comp_c1(), fountain=S.comp_fountain(), c2_0=S.comp_c2(0), c2_1=S.comp_c2(1), c2_2=S.comp_c2(2),
                 c2_3=S.comp_c2(3), defaults=abi.make_component())
    o = orc.OracleSystem(oracle_port)
    for name, comp in comps.items():
        t = o.reserve_type(comp)
        assert t == g[name]["typeId"]
        assert o.type_lut(t).tobytes().hex() == g[name]["lut"], name
        assert bytes(o.type_ubo(t)).hex() == g[name]["ubo"], name
    # known answers of SURVEY.md Appendix C for the C1 descriptor
    lut = np.frombuffer(bytes.fromhex(g["c1"]["lut"]), dtype=np.uint8)
    assert lut[0] == 0 and tuple(lut[256:259]) == (255, 188, 89)
    assert lut[27 * 4] == 178 and tuple(lut[256 + 27 * 4:256 + 27 * 4 + 3]) == (206, 155, 89)
    assert lut[63 * 4] == 51 and tuple(lut[256 + 63 * 4:256 + 63 * 4 + 3]) == (89, 89, 89)


def test_rng_known_answers(oracle_port):
    """sf::Random(2, 581271): burst of 6 + 1 timer spawn = 50 draws."""
    g = golden("rng.json")
    assert g["rows"][1] == dict(state="b0a3e85a9933dcf1", u32="e2393051", floatBits="3f623930")
    sc = [s for s in S.scenarios() if s.name == "burst6"][0]
    o = orc.OracleSystem(oracle_port, sc.seed)
    e = o.add_effect(*sc.effects[0])
    o.update([e], abi.make_frame_args(o.effect_info(e).timeStep))
    info = o.effect_info(e)
    assert (info.aliveCount, info.slotCount, info.freeCount) == (7, 8, 1)
    assert list(o.read_free(e)) == [4]
    assert "%016x" % o.rng_state()[0] == g["after50"] == "ecc7bcc0d6d44264"
    assert S.f32_bits(info.spawnTimer) == S.f32_bits(float(np.float32(999.9)))


def test_port_render_matches_golden(oracle_port):
    from golden.make_golden import render_scene
    g = golden("render.json")
    assert render_scene(orc.OracleSystem(oracle_port, (1, 5, 3, 4))) == g


def test_frustum_restatements_agree(oracle_port):
    ra = make_render_args()
    f = orc.make_frustum(list(ra.worldToClip.v), 0.0, kind=oracle_port)
    assert bytes(f) == bytes(ra.frustum)


@pytest.mark.parametrize("sc", S.scenarios(), ids=lambda s: s.name)
def test_port_matches_reference_build(oracle_port, oracle_ref, sc):
    """Frame by frame against the reference's own code (both Float4 back ends)."""
    port = S.run_scenario(orc.OracleSystem(oracle_port, sc.seed), sc)
    simd = S.run_scenario(orc.OracleSystem("reference", sc.seed), sc)
    scalar = S.run_scenario(orc.OracleSystem("reference_scalar", sc.seed), sc)
    for f, (a, b, c) in enumerate(zip(port, simd, scalar)):
        S.assert_snapshots_equal(c, a, "%s frame %d scalar" % (sc.name, f))
        # gravityPoints: the SSE build uses rsqrtps + one Newton step (src/sf/Float4.h:180-186)
        S.assert_snapshots_equal(b, a, "%s frame %d simd" % (sc.name, f), float_exact=sc.float_exact, rtol=1e-4)


def test_sorted_order_oracle_is_a_permutation(oracle_port):
    sc = [s for s in S.scenarios() if s.name == "fountain"][0]
    snaps = S.run_scenario(orc.OracleSystem(oracle_port, sc.seed), sc)
    st = snaps[-1][0]["stream"]
    exp, order = S.expected_sorted(st)
    assert sorted(order.tolist()) == list(range(len(st) // 4))
    k = S.sort_key(exp[::4])
    assert (k[1:] >= k[:-1]).all()


def test_reference_renderer_capture_is_reproducible(oracle_ref):
    """tests/golden/render_capture.json (row N1: what cl::ParticleSystem::renderMain hands to sokol) comes out of the
    reference build again, bit for bit -- the adapter on the GPU box is compared with this file."""
    import json
    from golden.make_golden import render_capture_scene
    got = render_capture_scene(orc.OracleSystem(oracle_ref, (1, 6, 3, 4)))
    assert json.loads(json.dumps(got)) == golden("render_capture.json")
