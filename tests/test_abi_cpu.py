"""CPU-side checks: the C-ABI library loads and exports every symbol include/sphb200.h declares,
the ctypes structs match the header, the host geometry logic, and the product never touches oracle/."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import build_native
    build_native.build()
    import _native
    return _native


def test_header_symbols_exported(native):
    hdr = open(os.path.join(ROOT, "include", "sphb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(sphb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    lib = C.CDLL(native.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/sphb200.h but not exported"
    assert names == set(native.SIGNATURES), names ^ set(native.SIGNATURES)


def test_struct_sizes_match_header(native):
    # sizes implied by include/sphb200.h (LP64)
    assert C.sizeof(native.Consts) == 8 * 3 + 4 * 2 + 8 * 8 + 8   # + the per-particle mass pointer
    assert C.sizeof(native.Grid) == 8 * 4 + 4 * 4
    assert C.sizeof(native.Status) == 32
    assert C.sizeof(native.Cells) == 48 + 8 + 5 * 8 + 8 + 2 * 4 + 2 * 8   # + nlevels, exp0, levels, level_hmax
    assert C.sizeof(native.Level) == 48 + 2 * 4
    assert C.sizeof(native.Lists) == 3 * 8 + 8
    assert C.sizeof(native.NList) == native.load().sphb_sizeof_nlist()


def test_no_compute_entry_points_without_gpu(native):
    lib = native.load()
    assert lib.sphb_abi_version() == 1
    assert lib.sphb_launch_count() == 0
    g = native.Grid(0, 0, 0, 1.0, 4, 5, 6, 0)
    assert lib.sphb_grid_ncell(C.byref(g)) == 120
    assert lib.sphb_grid_workspace_bytes(1000, C.byref(g)) > 0
    # argument validation happens before any CUDA call
    assert lib.sphb_grid_build(None, None, -1, C.byref(g), 0.0, None, None, None, None, None, None, None, 0, None) == -1
    assert lib.sphb_force(None, None, None, None, None, None, None, None, None, None, None, None) == -1


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pyfluid as sph
    import numpy as np
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sph.density_arr(np.zeros((4, 3)), np.ones(4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "python-sph_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "sph_oracle" not in src, f


def test_choose_grid():
    import engine
    g, delta = engine.choose_grid([0, 0, 0], [10, 5, 0], 0.5, 1000, cell=1.0, xdiv=2)
    assert g.cell == 1.0 and g.xdiv == 2 and g.nz == 3 and g.nx == 25 and g.ny == 8
    assert g.ox == -1.0 and delta > 0
    # cell count is capped relative to n
    g2, _ = engine.choose_grid([0, 0, 0], [1000, 1000, 1000], 0.01, 100)
    assert g2.ncell <= 8 * 100
    # degenerate h
    g3, _ = engine.choose_grid([0, 0, 0], [1, 1, 1], float("nan"), 10)
    assert g3.ncell >= 1 and g3.cell > 0


def test_gridref_stable_sort():
    import numpy as np
    import gridref
    rng = np.random.default_rng(0)
    pos = rng.random((500, 3)) * 4 - 0.5
    ids, perm, start = gridref.build(pos, (0, 0, 0), 1.0, (6, 3, 3), xdiv=2)
    assert start[-1] == 500 and (np.diff(ids[perm]) >= 0).all()
    for c in range(54):
        seg = perm[start[c]:start[c + 1]]
        assert (np.diff(seg) > 0).all()


def test_choose_levels_and_level_table():
    """nested grids: levels from per-octave statistics (what sphb_level_stats returns), numbered level-major; the
    device table has the layout of sphb_level (56 bytes); the numpy restatement bins by the same octaves"""
    import ctypes
    import numpy as np
    import _native
    import engine
    import gridref
    assert ctypes.sizeof(_native.Level) == 56 and ctypes.sizeof(_native.Cells) == 128
    import workloads
    n = 40000
    w = workloads.plummer(n, seed=1)
    pos, h = w["pos"], w["h_guess"] / 0.8                 # ~7 octaves of h, small h in the core only
    hmax = float(h.max())
    exp_top = int(np.floor(np.log2(hmax)))
    bins = gridref.level_of(h, exp_top - 15, 16)
    stats = []
    for b in range(16):
        m = bins == b
        if m.any():
            stats.append([float(m.sum()), float(h[m].sum())] + pos[m].min(axis=0).tolist() + pos[m].max(axis=0).tolist())
        else:
            stats.append([0.0] * 8)
    lv, delta = engine.choose_levels(stats, exp_top, pos.min(axis=0).tolist(), pos.max(axis=0).tolist(), hmax)
    occupied = [b for b in range(16) if stats[b][0] > 0]
    assert lv.nlevels == 16 - occupied[0] and lv.exp0 == exp_top - 15 + occupied[0] and delta > 0
    assert lv.bases[0] == 0 and lv.ncell == sum(g.ncell for g in lv.grids) and len(lv.table()) == 56 * lv.nlevels
    lev = gridref.level_of(h, lv.exp0, lv.nlevels)
    for b, g in enumerate(lv.grids):
        p = pos[lev == b]
        if len(p):   # every level's grid contains its particles, with about one support radius per cell
            c = (p - np.array([g.ox, g.oy, g.oz])) / g.cell
            assert (c >= 0).all() and (c < np.array([g.nx, g.ny, g.nz])).all()
            assert 0.5 < g.cell / (2 * h[lev == b].mean()) < 2.0, (b, len(p), g.cell, h[lev == b].mean())
    ids, perm, start = gridref.build_levels(pos, h, lv.exp0, [((g.ox, g.oy, g.oz), g.cell, (g.nx, g.ny, g.nz)) for g in lv.grids])
    assert start[-1] == n and (np.diff(ids[perm]) >= 0).all() and (np.diff(lev[perm]) >= 0).all()
    # one occupied octave: no nesting
    assert engine.choose_levels([[0.0] * 8] * 15 + [[10.0, 5.0, 0, 0, 0, 1, 1, 1]], 0, [0, 0, 0], [1, 1, 1], 0.6) is None


def test_per_particle_margin_plan_is_order_independent():
    """ListMargins.plan: every record gets margins from its OWN measured motion; the planned life comes from integer
    histogram counts, so it does not depend on the order of the records (slot order, rank decomposition)"""
    import torch
    import engine
    torch.manual_seed(0)
    n = 50000
    rd = torch.rand(n, dtype=torch.float64) ** 8 * 0.04        # a few fast records, a quiet majority
    rg = torch.rand(n, dtype=torch.float64) ** 8 * 0.02
    rd[7] = 0.0
    a, b = engine.ListMargins(), engine.ListMargins()
    a.age = b.age = 2
    ma = a.plan(rd, rg, 20)
    perm = torch.randperm(n)
    mb = b.plan(rd[perm], rg[perm], 20)
    assert a.life == b.life >= 1 and torch.equal(mb, ma[perm])
    assert ma.dtype == torch.float32 and ma.shape == (n, 2)
    assert float(ma[:, 0].max()) <= a.cover_h and float(ma[:, 1].max()) <= a.skin          # the uniform bounds cover all
    assert float(ma[:, 1].mean()) < 0.25 * float(ma[:, 1].max())                            # the majority keeps tight lists
    assert float(ma[7, 1]) == pytest.approx(a.PP_FLOOR)
    # two "ranks": the same plan from all-reduced counts
    c = engine.ListMargins()
    c.age = 2
    half = n // 2
    other = {}

    def reduce(counts, maxes):
        cd = engine.ListMargins()
        cd.age = 2
        def grab(c2, m2):
            other["c"], other["m"] = c2.clone(), m2.clone()
        cd.plan(rd[half:], rg[half:], 20, reduce=grab)
        counts += other["c"]
        torch.maximum(maxes, other["m"], out=maxes)
    mc = c.plan(rd[:half], rg[:half], 20, reduce=reduce)
    assert c.life == a.life and torch.equal(mc, ma[:half])
    # a faster flow plans a shorter life
    d = engine.ListMargins()
    d.age = 2
    d.plan(rd * 8, rg * 8, 20)
    assert d.life <= a.life


def test_range_decomposition_host_helpers():
    """ranges.py: weighted slot cuts (aligned to warps, monotone, degenerate weights) and the geometry that rank 0 draws
    and broadcasts survives its plain-tuple form, for one grid and for nested grids"""
    import torch
    import engine
    import ranges
    w = torch.ones(1000, dtype=torch.float64)
    c = ranges.weighted_cuts(torch.cumsum(w, 0), 4)
    assert c[0] == 0 and c[-1] == 1000 and c == sorted(c) and all(x % 32 == 0 for x in c[:-1])
    assert max(b - a for a, b in zip(c[:-1], c[1:])) <= 250 + 32
    w[:10] = 1e6                                      # all the work in the first ten slots: later ranks get empty ranges, no crash
    c = ranges.weighted_cuts(torch.cumsum(w, 0), 4)
    assert c[0] == 0 and c[-1] == 1000 and c == sorted(c)
    assert ranges.weighted_cuts(torch.cumsum(torch.ones(64, dtype=torch.float64), 0), 1) == [0, 64]
    g, delta = engine.choose_grid([0, 0, 0], [10, 5, 2], 0.5, 1000)
    g2, d2 = ranges._geom_from_plain(ranges._geom_to_plain((g, delta)))
    assert d2 == delta and bytes(g2) == bytes(g)
    lv = engine.Levels(g, -7, [g, engine.choose_grid([1, 1, 1], [3, 3, 2], 0.05, 500)[0]])
    lv2, d3 = ranges._geom_from_plain(ranges._geom_to_plain((lv, 0.125)))
    assert d3 == 0.125 and lv2.exp0 == -7 and lv2.ncell == lv.ncell and lv2.table() == lv.table() and bytes(lv2.frame) == bytes(lv.frame)


def test_status_prints_are_summarised_not_dropped_and_crowded_cells_are_seen():
    import numpy as np
    import torch
    import engine
    out = []
    st = np.zeros((2, 8), dtype=np.int32)
    st[0, 1] = 3                       # three Newton warnings
    st[1, 2] = engine.MAX_REPEAT_PRINTS + 500
    engine.raise_for_status(st, printer=out.append)
    assert out[:3] == ["WARNING: Failure to converge in newton_new_h."] * 3
    assert out.count("ERROR: Root out of bisection bounds.") == engine.MAX_REPEAT_PRINTS
    assert out[-1].startswith("[sph-b200] ... and 500 more: ERROR: Root out of bisection bounds.")
    st[1, 0] = 2
    with pytest.raises(RuntimeError, match="Obtained negative pressure."):
        engine.raise_for_status(st, printer=lambda *_: None)
    assert engine.max_cell_population(torch.tensor([0, 3, 3, 10000, 10001], dtype=torch.int32)) == 9997
    assert engine.max_cell_population(torch.tensor([0], dtype=torch.int32)) == 0
