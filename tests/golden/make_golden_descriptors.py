#!/usr/bin/env python3
"""Generate tests/golden/descriptors.json: JSON texts with what the REFERENCE's own reader (lenient jsi dialect +
reflection + sp::readJson, oracle/_ref/libspear_ref_json.so) makes of them. Canonical texts come from the reference's
own writer (sp::writeJson); the hand-written ones exercise the dialect and the failure rules."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenarios as S  # noqa: E402
from oracle import descriptors as D  # noqa: E402

HAND_WRITTEN = {
    "bare_component_partial": '{ "timeStep": 0.02, "burstAmount": 12, "lifeTime": 3, "gravity": { "y": -9.81 } }',
    "dialect": """// a prefab written by hand
{
  name: "fx/sparks"   /* bare keys, no comma after this line */
  components: [
    { type: "PointLight", radius: 4.0, },
    null,
    { type: "ParticleSystem"
      texture: "Assets/fx/spark.png"
      frameCount: { x: 4 y: 2 },
      timeStep: 0.016667, spawnTime: 0.01 spawnTimeVariance: 5e-3,
      burstAmount: 25.0, localSpace: 1, instantDelete: true, updateOutOfCamera: 0.0,
      renderOrder: -2,
      emitPosition: { boxExtent: { x: 1, y: 0.5, z: 1 }, sphere: { maxRadius: 2, scale: { y: 0.25 } } },
      emitVelocity: { offset: { y: 3 }, rotation: { x: 10, y: 20, z: 30 }, },
      gravityPoints: [ { position: { x: 1, y: 2, z: 3 }, strength: -2.5 }, { radius: 0.01 }, ],
      scaleSpline: { points: [ { x: 0, y: 0 }, { x: 0.3, y: 1 } { x: 1, y: 0.2 } ] },
      gradient: { defaultColor: { x: 0.5, y: 0.25, z: 1 }, points: [ { t: 0, color: { x: 1 } }, { t: 1 } ] },
      someFutureField: [1, 2, 3],
    },
    { "type": "ParticleSystem", "lifeTime": 1e1, "spin": -90, },
  ],
}
""",
    # a string value that opens with a raw line break is a multi-line string: sp::readJson takes the first line's indentation
    # off every line (Json.cpp:134-166); an escaped \\n or a line break further in does not make one
    "multiline_texture": '{ "type": "ParticleSystem", "texture": "\n      Assets/fx/\n        smoke sheet.png\n\n      tail\n    ", "size": 3 }',
    "multiline_texture_crlf_tabs": '{ "texture": "\r\n\t\t  first\r\n\t\t  second\r\n\t third\r\n\t\t\t\t  fourth\r\n", "lifeTime": 2 }',
    "not_multiline_texture": '{ "texture": "Assets/fx/a\n   b.png", "burstAmount": 3 }',
    "escaped_newline_texture": '{ "texture": "\\n   Assets/fx/c.png" }',
    "tagged_component": '{ "type": "ParticleSystem", "size": 2.5, "sizeVariance": 1, "alphaSpline": { "points": [] } }',
    "no_particle_components": '{ "name": "x", "components": [ { "type": "Door" } ] }',
    "fail_wrong_type_number": '{ "timeStep": "fast" }',
    "fail_wrong_type_array": '{ "type": "ParticleSystem", "gravityPoints": { "position": { "x": 1 } } }',
    "fail_bool_from_string": '{ "localSpace": "yes" }',
    "fail_unknown_component": '{ "components": [ { "type": "Teleporter" } ] }',
    "fail_number_for_object": '{ "gravity": 3 }',
    "fail_syntax": '{ "timeStep": 0.1 ',
}


def main():
    o = D.JsonOracle()
    comps = [S.comp_fountain(), S.comp_gravity_points(), S.comp_c1(burst=500), S.comp_c2(1, burst=64), S.comp_local_space(), S.comp_timed_emitter(), S.comp_point()]
    textures = ["Assets/fx/spark.png", None, "a", "Assets/fx/smoke sheet.png", None, "t", None]
    cases = {}
    cases["canonical_pretty"] = o.write_prefab(comps, textures, pretty=True)
    cases["canonical_compact"] = o.write_prefab(comps, textures, pretty=False)
    cases.update(HAND_WRITTEN)
    out = {"cases": []}
    for name, text in cases.items():
        got = o.read(text)
        # sv::ParticleSystemComponent::emitVelocityAttractorStrength has no default initialiser (ServerState.h:212): when the
        # text does not set it the reference reads back whatever the allocation held. Recorded as 0, the product's default.
        if got is not None and "emitVelocityAttractorStrength" not in text:
            for v in got:
                v[D.ATTRACTOR_STRENGTH_INDEX] = 0.0
        out["cases"].append({"name": name, "json": text, "components": None if got is None else [[x.hex() for x in v] for v in got]})
        print(name, "->", "FAIL" if got is None else "%d component(s)" % len(got))
    json.dump(out, open(os.path.join(HERE, "descriptors.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
