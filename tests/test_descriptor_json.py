"""CPU (host-only product code): sprDescriptorsFromJson (spear_b200/csrc/descriptor_json.cpp) against what
the reference's own reader -- lenient jsi dialect + reflection tables + sp::readJson -- makes of the same JSON texts:
tests/golden/descriptors.json (generated from oracle/_ref/libspear_ref_json.so) and, when that library is built, the
reference itself on further texts. Field values are compared exactly (as float.hex)."""
import json
import os

import pytest

import scenarios as S
from oracle import descriptors as D

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "descriptors.json")))


def _particles():
    from spear_b200 import particles
    return particles


def _load(text):
    """Product loader -> list of field vectors (float.hex), or None when the load fails."""
    P = _particles()
    try:
        ds = P.DescriptorSet(text)
    except Exception:
        return None
    try:
        out = []
        for i in range(len(ds)):
            c = ds[i]
            name = ds.texture_name(i)
            assert c.texture == (D.fnv1a64(name.encode()) if name else 0)
            out.append([x.hex() for x in D.flatten(c)])
        return out
    finally:
        ds.close()


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: c["name"])
def test_loader_matches_reference_reader_golden(case):
    assert _load(case["json"]) == case["components"]


@pytest.mark.skipif(not D.have_reference(), reason="oracle/_ref/libspear_ref_json.so not built (needs /root/reference)")
def test_loader_matches_reference_reader_on_every_scenario_descriptor():
    o = D.JsonOracle()
    comps = [S.comp_c1(), S.comp_c5(), S.comp_fountain(), S.comp_gravity_points(), S.comp_local_space(), S.comp_timed_emitter(),
             S.comp_no_emit(), S.comp_point()] + [S.comp_c2(k, burst=100 + k, churn=bool(k & 1)) for k in range(4)]
    for pretty in (True, False):
        text = o.write_prefab(comps, ["tex%d" % i if i % 3 else None for i in range(len(comps))], pretty=pretty)
        want = o.read(text)
        assert want is not None and len(want) == len(comps)
        assert "emitVelocityAttractorStrength" in text   # the writer always emits it, so the reader never leaves it uninitialised here
        assert _load(text) == [[x.hex() for x in v] for v in want]


def test_loaded_descriptor_equals_the_one_it_was_written_from_up_to_text_precision():
    """sp::writeJson prints floats with 6 decimals, so a round trip is exact only for values that survive that."""
    P = _particles()
    text = [c for c in GOLDEN["cases"] if c["name"] == "canonical_pretty"][0]["json"]
    ds = P.DescriptorSet(text)
    try:
        c = ds[2]                                     # comp_c1(burst=500)
        assert (c.burstAmount, c.lifeTime, c.lifeTimeVariance, c.drag) == (500, 4.0, 2.0, pytest.approx(0.3, abs=1e-6))
        assert c.scaleSpline.numPoints == 3 and (c.scaleSpline.points[1].x, c.scaleSpline.points[1].y) == (pytest.approx(0.3, abs=1e-6), 1.0)
        assert c.gradient.numPoints == 2 and ds.texture_name(2) == "a" and ds.texture_name(1) == ""
        assert ds[1].numGravityPoints >= 1
    finally:
        ds.close()


def test_errors_are_reported():
    P = _particles()
    with pytest.raises(Exception) as e:
        P.DescriptorSet('{ "timeStep": "fast" }')
    assert "timeStep" in str(e.value)
    with pytest.raises(Exception):
        P.DescriptorSet("[1, 2]")
