"""CPU: the shape-bucketed, least-recently-used, byte-bounded plan cache (zipvoice_b200/engine.py) with a
stand-in plan factory -- the policy is host logic; the real plans are exercised by the -m gpu tests."""
import random

from zipvoice_b200.engine import PlanCache, round_up


class _Packed:
    class device:
        type = "cpu"


class _FakePlan:
    def __init__(self, packed, N, T):
        self.N, self.T = N, T
        self.graphs = {}
        self.nbytes = N * T * 1000


def _cache(**kw):
    return PlanCache(_Packed(), factory=_FakePlan, **kw)


def test_round_up():
    assert round_up(1219, 64) == 1280 and round_up(1280, 64) == 1280 and round_up(7, 0) == 7 and round_up(7, 1) == 7


def test_exact_shapes_by_default_and_lru_eviction():
    c = _cache(max_plans=3)
    a = c.get(2, 100)
    assert (a.N, a.T) == (2, 100) and c.get(2, 100) is a
    c.get(2, 101); c.get(2, 102)
    c.get(2, 100)                    # refresh a
    c.get(2, 103)                    # evicts (2, 101), the least recently used
    assert c.get(2, 100) is a and len(c) == 3
    assert c.created == 4
    c.get(2, 101)
    assert c.created == 5


def test_bucketed_shapes_bound_the_number_of_plans():
    c = _cache(max_plans=64, frame_bucket=128, row_bucket=8)
    rnd = random.Random(0)
    for _ in range(200):
        B, T = rnd.randint(1, 64), rnd.randint(600, 1219)
        p = c.get(B, T)
        assert p.N >= B and p.T >= T and p.N % 8 == 0 and p.T % 128 == 0 and p.N - B < 8 and p.T - T < 128
    assert c.created <= 8 * 6        # 8 row buckets x 6 frame buckets at most


def test_byte_budget_evicts_but_keeps_the_newest():
    c = _cache(max_plans=100, max_bytes=250_000)
    c.get(1, 100); c.get(1, 120)
    assert len(c) == 2
    big = c.get(4, 100)              # 400 kB alone exceeds the budget: everything else goes, it stays
    assert len(c) == 1 and c.get(4, 100) is big
