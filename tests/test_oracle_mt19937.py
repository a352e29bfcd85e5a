"""Pin oracle/mt19937.py against numpy's own legacy RandomState (the third-party holder of the algorithm)."""
import numpy as np
import pytest

from oracle.mt19937 import LegacyMT19937, LegacyMT19937Bank


@pytest.mark.parametrize("seed", [0, 1, 1707366464, 1707366464 + 12345, 4294967295])
def test_raw_stream_and_draw_protocol(seed):
    rs = np.random.RandomState(seed)
    m = LegacyMT19937(seed)
    raw = np.frombuffer(rs.bytes(4 * 1500), dtype="<u4")
    mine = np.array([m.random_uint32() for _ in range(1500)], dtype=np.uint32)
    assert (raw == mine).all()
    rs = np.random.RandomState(seed)
    m = LegacyMT19937(seed)
    assert rs.choice([0, 1, 2, 3]) == m.interval(3)                     # environment.py:29
    assert rs.uniform(0, 75) == m.uniform(0, 75)                        # environment.py:33
    u = rs.uniform(5, 95, 2)                                            # environment.py:54
    assert u[0] == m.uniform(5, 95) and u[1] == m.uniform(5, 95)
    n1 = rs.normal(0, 5, size=(2,))                                     # robot.py:640
    assert n1[0] == m.normal(0, 5) and n1[1] == m.normal(0, 5)
    assert rs.normal(0.5, 2.5) == m.normal(0.5, 2.5)
    for n, B in [(988, 100), (10000, 100), (100, 100), (257, 256), (2, 1)]:
        assert (rs.choice(n, B, replace=False) == m.choice_no_replace(n, B)).all()   # robot.py:111
    c = rs.choice([-5, 5], 2)                                           # environment.py:156
    assert list(c) == [[-5, 5][m.interval(1)] for _ in range(2)]
    assert rs.uniform(0, 1) == m.uniform(0, 1)                          # streams still aligned


def test_state_roundtrip_with_numpy():
    rs = np.random.RandomState(42)
    rs.normal(size=3)           # leaves a cached gauss
    m = LegacyMT19937()
    m.set_state(rs.get_state())
    assert rs.normal() == m.gauss()
    assert rs.random_sample() == m.random_double()


def test_bank_matches_scalar():
    seeds = 1707366464 + np.arange(7)
    bank = LegacyMT19937Bank(seeds)
    ms = [LegacyMT19937(int(s)) for s in seeds]
    rs = np.random.RandomState(0)
    for _ in range(700):
        act = rs.rand(7) < 0.7
        out = bank.random_uint32(act)
        for i, m in enumerate(ms):
            if act[i]:
                assert out[i] == m.random_uint32()
    d = bank.random_double()
    assert all(d[i] == m.random_double() for i, m in enumerate(ms))
