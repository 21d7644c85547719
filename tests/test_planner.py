"""Launch planner (vw_plan_levels through vw_plan_query / vw_describe_plan): host logic only, runs without a GPU.
The plan decides which consecutive levels share a launch; whatever it picks, the groups must tile 1..J exactly, respect
the halo rule (a fused group's dilated halo fits the signal and the shared-memory tile), and be deterministic."""
import re

import pytest

from vectorwave_b200 import _native

SHAPES = [(2, 4, 4096), (8, 4, 4096), (16, 8, 65536), (30, 10, 1 << 28), (16, 6, 1 << 20), (2, 10, 1 << 28),
          (8, 10, 1 << 28), (4, 5, 10000), (12, 6, 1 << 16), (20, 3, 512), (6, 1, 64), (30, 9, 1 << 14), (2, 1, 2)]


@pytest.mark.parametrize("l,levels,n", SHAPES)
@pytest.mark.parametrize("forward", [1, 0])
def test_groups_tile_the_levels(l, levels, n, forward):
    groups = _native.plan_groups(forward, l, levels, n)
    assert groups == _native.plan_groups(forward, l, levels, n)          # deterministic
    nxt = 1
    for first, nlev in groups:
        assert first == nxt and nlev >= 1
        nxt = first + nlev
        if l >= 24:
            assert nlev == 1                                             # FP64-bound filters never share a launch
        assert nlev <= 4
    assert nxt == levels + 1


@pytest.mark.parametrize("l,levels,n", SHAPES)
def test_described_plan_is_consistent(l, levels, n):
    for forward in (1, 0):
        text = _native.describe_plan(forward, l, levels, n)
        rows = re.findall(r"levels (\d+)-(\d+): (\w+) tile=(-?\d+) halo=(\d+) cost=([\d.]+)", text)
        assert [(int(a), int(b) - int(a) + 1) for a, b, *_ in rows] == _native.plan_groups(forward, l, levels, n)
        for a, b, kind, tile, halo, cost in rows:
            a, b, tile, halo = int(a), int(b), int(tile), int(halo)
            assert float(cost) > 0
            if kind == "fused":
                # dilated halo of the group: (l-1) * 2^(first-1) * (2^nlev - 1), rounded to even for 16-byte bulk copies
                exact = (l - 1) * (1 << (a - 1)) * ((1 << (b - a + 1)) - 1)
                assert halo in (exact, exact + 1) and halo <= n + 1
                assert tile > 0 and tile % 2 == 0 and tile <= n + 1
                assert (4 * (tile + halo)) * 8 <= 227 * 1024                 # fits the opt-in shared memory either direction
            elif kind == "column":
                assert a == b and halo == (l - 1) * (1 << (a - 1))


def test_forced_tile_and_fuse_are_honoured():
    text = _native.describe_plan(1, 8, 4, 4096, 1024, 2)
    rows = re.findall(r"levels (\d+)-(\d+): (\w+) tile=(-?\d+)", text)
    assert all(int(b) - int(a) + 1 <= 2 for a, b, _, _ in rows)
    assert all(int(t) == 1024 for _, _, k, t in rows if k == "fused")
