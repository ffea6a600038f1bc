"""Pin the CPU oracle to the executed reference (CPU-only test).

tests/golden/<scenario>.npz are dumps of the unmodified reference running on a
B200 through Numba-CUDA; tiny_sim.npz is the same scenario through
NUMBA_ENABLE_CUDASIM in the CPU container (tests/golden/make_golden.py).
"""
import pytest

import replay
from oracle.gvom_oracle import OracleGvom


@pytest.mark.parametrize("name", ["tiny", "small_moving", "small_quirks", "small_eigen2"])
def test_oracle_matches_reference_small(name):
    bad, _ = replay.replay(lambda P: OracleGvom(*P), name, replay.golden(name), what="oracle")
    assert not bad, "\n".join(bad[:20])


@pytest.mark.parametrize("name", ["os1_64", "os1_128", "long_range"])
def test_oracle_matches_reference_full_size(name):
    bad, _ = replay.replay(lambda P: OracleGvom(*P), name, replay.golden(name), what="oracle")
    assert not bad, "\n".join(bad[:20])


def test_oracle_matches_cudasim_tiny():
    """CUDASIM types the DDA differently from compiled PTX; the tiny
    scenario is chosen so both agree, which pins the host-side call order."""
    bad, _ = replay.replay(lambda P: OracleGvom(*P), "tiny", replay.golden("tiny", "_sim"), what="oracle-vs-sim")
    assert not bad, "\n".join(bad[:20])
