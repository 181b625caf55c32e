"""Nested cell grids and BASELINE config 5 (the Plummer sphere): h spans decades, one uniform grid
cannot serve it.  Cell assignment is pinned by the numpy restatement (tests/gridref.py: the reference has no cell
list); neighbour sets are exact; h, Omega, a, du/dt within 1e-10 of the CPU oracle (full arrays at small N, sampled
particles at 1 M)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.where(both_nan | (a == b), 0.0, np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
    return float(np.max(d)) if d.size else 0.0


def vrel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-300)))


@pytest.fixture(scope="module")
def sph():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import build_native
    build_native.build()
    import pyfluid
    pyfluid.G = 0.0
    yield pyfluid
    pyfluid.PARTICLE_MASS = 1


def _plummer(n, seed=0):
    import workloads
    w = workloads.plummer(n, seed=seed)
    rng = np.random.default_rng(seed + 17)
    w["vel"] = (rng.random((n, 3)) - 0.5) * 0.2
    w["u"] = w["u"] * (1 + 0.1 * rng.random(n))
    return w


def _levels_of(dg):
    lv = dg.levels
    return [((g.ox, g.oy, g.oz), g.cell, (g.nx, g.ny, g.nz)) for g in lv.grids]


@pytest.mark.parametrize("n,seed", [(3000, 0), (200000, 1)])
def test_nested_grid_build_bit_exact(sph, n, seed):
    """cells, permutation, cell_start of the nested grids == the numpy restatement; every level's grid contains its
    particles; level_hmax bounds the h of the level"""
    import torch
    import engine as E
    import gridref
    e = sph._eng()
    w = _plummer(n, seed)
    pos, h = w["pos"], w["h_guess"] / 0.8
    h[5] = -1.0          # unbounded support: last level
    posh = e.pack_posh(torch.from_numpy(pos).cuda(), torch.from_numpy(h).cuda())
    grid, delta = e.grid_for(posh)
    assert isinstance(grid, E.Levels) and grid.nlevels >= 4
    dg = e.build(posh, grid, delta)
    ids, perm, start = gridref.build_levels(pos, h, grid.exp0, _levels_of(dg))
    assert (dg.perm.cpu().numpy().astype(np.int64) == perm).all()
    assert (dg.cell_start.cpu().numpy().astype(np.int64) == start).all()
    assert (dg.cell_id.cpu().numpy().astype(np.int64) == ids[perm]).all()
    assert (dg.posh.cpu().numpy() == np.concatenate([pos, h[:, None]], 1)[perm]).all()
    lv = gridref.level_of(h, grid.exp0, grid.nlevels)
    assert lv[5] == grid.nlevels - 1
    for b, (o, cell, dims) in enumerate(_levels_of(dg)):
        p = pos[lv == b]
        if len(p):
            c = (p - np.asarray(o)) / cell
            assert (c >= 0).all() and (c < np.asarray(dims)).all(), b     # no clamped outliers
    hm = dg.level_hmax.cpu().numpy().astype(np.float64)
    for b in range(grid.nlevels):
        hb = h[(lv == b) & (h > 0)]
        if b == grid.nlevels - 1:
            assert np.isinf(hm[b])                                         # the h <= 0 record
        elif len(hb):
            assert hb.max() <= hm[b] <= hb.max() * (1 + 2e-7)
        else:
            assert hm[b] == 0.0


def test_nested_neighbour_sets_exact(sph, oracle_mod):
    """NB(j) = {i : distance/h_j < 2} and the symmetric support, found through the nested grids == brute force"""
    import torch
    import engine as E
    e = sph._eng()
    n = 20000
    w = _plummer(n, 3)
    pos, h = w["pos"], w["h_guess"] / 0.8
    posh = e.pack_posh(torch.from_numpy(pos).cuda(), torch.from_numpy(h).cuda())
    grid, delta = e.grid_for(posh)
    assert isinstance(grid, E.Levels)
    dg = e.build(posh, grid, delta)
    perm = dg.perm.cpu().numpy().astype(np.int64)
    o = oracle_mod.Oracle(G=0.0, PARTICLE_MASS=w["m"])
    rng = np.random.default_rng(4)
    r = np.linalg.norm(pos, axis=1)
    js = np.concatenate([rng.choice(n, 40, replace=False), np.argsort(r)[:4], np.argsort(r)[-4:]])
    slot_of = np.empty(n, dtype=np.int64)
    slot_of[perm] = np.arange(n)
    q = torch.from_numpy(slot_of[js].astype(np.int32)).cuda()
    for sym in (False, True):
        cap = n
        out, cnt = e.neighbours(dg, q, cap, symmetric=sym)
        out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
        for k, j in enumerate(js):
            assert cnt[k] <= cap
            got = perm[out[k, :cnt[k]]]
            assert (np.diff(out[k, :cnt[k]].astype(np.int64)) > 0).all()        # ascending slots
            ref = o.neighbours_sym(pos, h, int(j)) if sym else o.neighbours(pos, int(j), h[j])
            assert (np.sort(got) == np.sort(np.asarray(ref))).all(), (sym, j)


def test_plummer_stage_and_step_vs_oracle(sph, oracle_mod):
    """cold h solve (Newton, bisection where it gives up), Omega, a, du/dt and one leapfrog step on a small Plummer
    sphere: scanning kernels and kept-list kernels, both on nested grids, against the oracle's full arrays"""
    import torch
    import engine as E
    n = 3000
    w = _plummer(n, 7)
    sph.PARTICLE_MASS = w["m"]
    o = oracle_mod.Oracle(G=0.0, PARTICLE_MASS=w["m"])
    o.set_num_threads(os.cpu_count() or 1)
    pos, vel, u = w["pos"], w["vel"], w["u"]
    h = sph.newton_smoothlength_arr(pos, w["h_guess"])
    h_ref = o.newton_smoothlength_arr(pos, w["h_guess"])
    assert rel(h, h_ref) < RTOL
    assert h.max() / h.min() > 30.0
    eng = sph._eng()
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    grid, _ = eng.grid_for(eng.pack_posh(to(pos), to(h)))
    assert isinstance(grid, E.Levels)
    rho = sph.var_density_arr(h)
    P = sph.pressure_arr(u, rho)
    om = sph.omega_arr(pos, rho, h)
    assert rel(om, o.omega_arr(pos, rho, h)) < RTOL
    acc = sph.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(n))
    assert vrel(acc, o.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(n))) < RTOL
    dudt = sph.energy_rate_arr(pos, vel, P, rho, h, om)
    d_ref = o.energy_rate_arr(pos, vel, P, rho, h, om)
    assert np.abs(dudt - d_ref).max() < RTOL * np.abs(d_ref).max()
    # one step: the scanning Stepper and the FastStepper (kept lists) against the oracle
    dt = 0.2 * w["dt"]
    ref = o.var_smoothlength_leapfrog(pos, vel, u, h, dt)
    c = sph._consts()
    for stepper in (E.Stepper(eng), E.FastStepper(eng)):
        got = stepper.step(to(pos), to(vel), to(u), to(h), dt, c)
        got = [x.cpu().numpy() for x in got]
        assert rel(got[0], ref[0]) < RTOL or np.abs(got[0] - ref[0]).max() < RTOL * np.abs(ref[0]).max()
        assert vrel(got[1], ref[1]) < RTOL
        assert rel(got[2], ref[2]) < RTOL and rel(got[3], ref[3]) < RTOL
        if isinstance(stepper, E.FastStepper):
            assert stepper.stats["safe_steps"] == 0 and stepper.nl is not None and stepper.nl.dg.levels is not None


def test_C5_plummer_1M_sampled(sph, oracle_mod):
    """config 5 at 1 M particles: conservation, idempotence of the solve, neighbour counts, and 48 sampled particles
    (core, rim and a seeded sample) against the oracle's per-particle functions"""
    import torch
    import engine as E
    import workloads
    n = 1 << 20
    w = workloads.plummer(n, seed=0)
    sph.PARTICLE_MASS = w["m"]
    o = oracle_mod.Oracle(G=0.0, PARTICLE_MASS=w["m"])
    o.set_num_threads(os.cpu_count() or 1)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pos = w["pos"]
    rng = np.random.default_rng(5)
    vel = (rng.random((n, 3)) - 0.5) * 0.2
    u = w["u"] * (1 + 0.1 * rng.random(n))
    h_t, it, code = sph.newton_smoothlength_arr(to(pos), to(w["h_guess"]), _details=True)
    h = h_t.cpu().numpy()
    assert np.isfinite(h).all() and h.max() / h.min() > 300.0            # the variable-h stress: > 2.5 decades at 1 M
    eng = sph._eng()
    grid, _ = eng.grid_for(eng.pack_posh(to(pos), h_t))
    assert isinstance(grid, E.Levels) and grid.nlevels >= 8
    rho_t = sph.var_density_arr(h_t)
    P_t = sph.pressure_arr(to(u), rho_t)
    om_t = sph.omega_arr(to(pos), rho_t, h_t)
    acc_t = sph.acceleration_arr(to(pos), rho_t, to(vel), P_t, h_t, om_t, torch.zeros_like(h_t))
    dudt_t = sph.energy_rate_arr(to(pos), to(vel), P_t, rho_t, h_t, om_t)
    rho, P, om, acc, dudt = (x.cpu().numpy() for x in (rho_t, P_t, om_t, acc_t, dudt_t))
    assert np.abs((w["m"] * acc).sum(axis=0)).max() < 1e-11 * np.abs(w["m"] * acc).sum()
    assert abs(((vel * acc).sum(axis=1) + dudt).sum()) < 1e-11 * (np.abs(vel * acc).sum() + np.abs(dudt).sum())
    _, it2, code2 = sph.newton_smoothlength_arr(to(pos), h_t, _details=True)
    assert ((it2 == 1) | (code != 0)).all()
    r = np.linalg.norm(pos, axis=1)
    order = np.argsort(r)
    js = np.concatenate([order[:8], order[-8:], rng.choice(n, 32, replace=False)])
    h_ref, it_ref, code_ref, _ = o.newton_smoothlength_while(js, pos, w["h_guess"][js])
    assert rel(h[js], h_ref) < RTOL and (it[js] == it_ref).all()
    assert rel(om[js], o.omega_j(js, pos, rho[js], h[js])) < RTOL
    assert vrel(acc[js], o.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(n), j=js)) < RTOL
    d_ref = o.energy_rate_arr(pos, vel, P, rho, h, om, j=js)
    assert np.abs(dudt[js] - d_ref).max() < RTOL * np.abs(dudt[js]).max()
    # neighbour counts: scanning kernel == list drain == size of the reference-arithmetic neighbour set
    c = sph._consts()
    posh = eng.pack_posh(to(pos), h_t)
    dg = eng.build(posh, *eng.grid_for(posh))
    _, _, _, ng_scan = eng.density_omega(dg, c, want_rho=False, want_ngb=True)
    nl = eng.nl_build(dg, 1.03, 0.03)
    _, _, _, ng_list = eng.nl_density_omega(nl, c, dg.posh, want_omega=False, want_ngb=True)
    assert nl.read_flags() == 0 and torch.equal(ng_scan, ng_list)
    ng = eng.scatter(ng_scan.double(), dg.perm, 1).cpu().numpy()
    for j in js[:24]:
        assert int(ng[j]) == len(o.neighbours(pos, int(j), h[j])), j


def test_plummer_kept_lists_with_per_particle_margins_vs_oracle(sph, oracle_mod):
    """four steps of the resident FastStepper on a small Plummer sphere: from the second build on the lists carry
    per-particle margins ON nested grids (level-wise and cell-wise allowance bounds in the builder's search); every step
    is checked against the oracle's leapfrog, so a pair missing from a list would show as an error far above 1e-10"""
    import torch
    import engine as E
    n = 3000
    w = _plummer(n, 11)
    sph.PARTICLE_MASS = w["m"]
    o = oracle_mod.Oracle(G=0.0, PARTICLE_MASS=w["m"])
    o.set_num_threads(os.cpu_count() or 1)
    pos, vel, u = w["pos"], w["vel"] * (0.2 + 4.0 * np.random.default_rng(3).random((n, 1))), w["u"]   # slow and fast movers
    h = sph.newton_smoothlength_arr(pos, w["h_guess"])
    eng = sph._eng()
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    c = sph._consts()
    fs = E.FastStepper(eng)
    fs.load(to(pos), to(vel), to(u), to(h), consts=c)
    dt = 2.0 * w["dt"]
    s = (pos, vel, u, h)
    saw_margins = False
    for _ in range(4):
        fs.advance(dt, c, defer_status=False)
        s = o.var_smoothlength_leapfrog(*s, dt)
        if fs.nl is not None:
            saw_margins = saw_margins or (fs.nl.margins is not None and fs.nl.dg.levels is not None)
        r = [x.cpu().numpy() for x in fs.state()]
        assert np.abs(r[0] - s[0]).max() < RTOL * np.abs(s[0]).max()
        assert vrel(r[1], s[1]) < 1e-9 and rel(r[2], s[2]) < RTOL and rel(r[3], s[3]) < RTOL
    assert saw_margins or fs.stats["safe_steps"] > 0
