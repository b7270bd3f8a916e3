"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical
descriptors and seeds. Bit-exact: counts, RNG state, free lists, emission indices,
compaction order, vertex floats, bounds, timers. The only
tolerance is for gravityPoints against the SSE reference (rsqrtps approximation)."""
import numpy as np
import pytest

import scenarios as S
from oracle import orc
from spear_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from spear_b200 import particles
    c = particles.ParticleContext(device=0, initial_slot_capacity=1 << 16)
    yield c
    c.close()


@pytest.mark.parametrize("sc", S.scenarios(), ids=lambda s: s.name)
def test_scenario_matches_oracle(ctx, oracle_port, sc):
    ref = S.run_scenario(orc.OracleSystem(oracle_port, sc.seed), sc)
    sys_ = ctx.create_system(sc.seed)
    try:
        got = S.run_scenario(sys_, sc)
    finally:
        sys_.close()
    assert len(ref) == len(got)
    for f, (a, b) in enumerate(zip(ref, got)):
        S.assert_snapshots_equal(a, b, "%s frame %d" % (sc.name, f))
