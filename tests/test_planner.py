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
            # FP64-bound filters never share a TILE launch; from the column threshold (level 3) on, two levels may share one
            # pass of the lattice pair kernels (csrc/vw_column.cu)
            assert nlev == 1 or (nlev == 2 and first >= 3)
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
                assert b - a in (0, 1) and halo == (l - 1) * (1 << (a - 1)) * ((1 << (b - a + 1)) - 1)
                if b > a:
                    # lattice pairs (30 taps) or direct-form pairs (16-20 taps: synthesis from level 4, analysis from level 5)
                    assert (l == 30 or (l in (16, 18, 20) and a >= (5 if forward else 4))) and 3 * (l - 1) * (1 << (a - 1)) <= n


def test_forced_tile_and_fuse_are_honoured():
    text = _native.describe_plan(1, 8, 4, 4096, 1024, 2)
    rows = re.findall(r"levels (\d+)-(\d+): (\w+) tile=(-?\d+)", text)
    assert all(int(b) - int(a) + 1 <= 2 for a, b, _, _ in rows)
    assert all(int(t) == 1024 for _, _, k, t in rows if k == "fused")


@pytest.mark.parametrize("l,levels,n_local,world", [(30, 10, 1 << 25, 8), (30, 10, 1 << 28, 1), (8, 6, 1 << 16, 2),
                                                    (16, 7, 1 << 16, 1), (2, 9, 1 << 13, 3), (8, 5, 1 << 14, 4)])
def test_span_plan_layout_invariants(l, levels, n_local, world):
    """vw_span_plan_query (host logic only): groups tile 1..J in both directions and equal the planner's groups for the
    WHOLE signal, every halo covers its group and is a multiple of 32 samples, lead / lead_w / pad / inverse_msg are the
    sums the cascades and the message packing rely on."""
    p = _native.span_plan(l, levels, n_local, world)
    gf = [(p.first_f[g], p.nlev_f[g]) for g in range(p.ngroups_f)]
    gi = [(p.first_i[g], p.nlev_i[g]) for g in range(p.ngroups_i)]
    assert gf == _native.plan_groups(1, l, levels, n_local * world)
    assert gi == _native.plan_groups(0, l, levels, n_local * world)
    for groups, halos in ((gf, p.halo_f), (gi, p.halo_i)):
        nxt = 1
        for g, (first, nlev) in enumerate(groups):
            assert first == nxt
            nxt = first + nlev
            need = (l - 1) * (1 << (first - 1)) * ((1 << nlev) - 1)
            assert halos[g] >= need and halos[g] % 32 == 0 and halos[g] - need < 32
        assert nxt == levels + 1
    assert p.lead == sum(p.halo_f[:p.ngroups_f]) and p.lead_w == sum(p.halo_f[1:p.ngroups_f])
    assert p.pad == sum(p.halo_i[:p.ngroups_i])
    si_in = [sum(p.halo_i[:g + 1]) for g in range(p.ngroups_i)]
    assert p.inverse_msg == si_in[-1] + sum(n * s for (_, n), s in zip(gi, si_in))
    assert max(p.lead, p.pad) <= n_local and (p.l, p.levels, p.world, p.n_local) == (l, levels, world, n_local)


def test_span_plan_rejects_what_cannot_be_sharded():
    from vectorwave_b200.errors import IllegalArgumentException, InvalidArgumentException
    with pytest.raises(IllegalArgumentException):       # total halo 29 * 255 > 256 samples per rank
        _native.span_plan(30, 8, 256, 64)
    with pytest.raises(InvalidArgumentException):       # upsampled filter longer than the whole signal
        _native.span_plan(30, 8, 256, 4)
