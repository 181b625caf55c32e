"""Parity at BASELINE.json's full sizes (configs C2, C3, C4): size-independent properties of the whole arrays
(momentum / energy conservation of the pair sums, idempotence of the h solve, neighbour counts) plus sampled-particle
parity against the CPU oracle's per-particle functions (each O(N), so a handful of particles).  Tolerance 1e-10."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.where(a == b, 0.0, np.abs(a - b) / np.maximum(np.abs(b), 1e-300))))


def vrel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-300)))


@pytest.fixture(scope="module")
def sph():
    import torch
    assert torch.cuda.is_available()
    import build_native
    build_native.build()
    import pyfluid
    pyfluid.G = 0.0
    return pyfluid


def _check(sph, oracle_mod, w, js, n_expect=None, vel_noise_seed=5):
    import torch
    n = len(w["pos"])
    sph.PARTICLE_MASS = w["m"]
    try:
        o = oracle_mod.Oracle(G=0.0, PARTICLE_MASS=w["m"])
        o.set_num_threads(os.cpu_count() or 1)    # the sampled particles are spread over the host threads
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        pos = w["pos"]
        rng = np.random.default_rng(vel_noise_seed)
        vel = (rng.random((n, 3)) - 0.5) * 0.2
        if (pos[:, 2] == 0).all():
            vel[:, 2] = 0.0
        u = w["u"] * (1 + 0.1 * rng.random(n))
        # cold solve from the config's guess, on device tensors (results stay on the GPU)
        h_t, it, code = sph.newton_smoothlength_arr(to(pos), to(w["h_guess"]), _details=True)
        assert (code == 0).all()
        rho_t = sph.var_density_arr(h_t)
        P_t = sph.pressure_arr(to(u), rho_t)
        om_t = sph.omega_arr(to(pos), rho_t, h_t)
        acc_t = sph.acceleration_arr(to(pos), rho_t, to(vel), P_t, h_t, om_t, torch.zeros_like(h_t))
        dudt_t = sph.energy_rate_arr(to(pos), to(vel), P_t, rho_t, h_t, om_t)
        h, rho, P, om, acc, dudt = (x.cpu().numpy() for x in (h_t, rho_t, P_t, om_t, acc_t, dudt_t))
        # conservation (Newton's third law of the pair terms) and first-law consistency
        assert np.abs(acc.sum(axis=0)).max() < 1e-11 * np.abs(acc).sum()
        assert abs(((vel * acc).sum(axis=1) + dudt).sum()) < 1e-11 * (np.abs(vel * acc).sum() + np.abs(dudt).sum())
        # idempotence of the solve
        _, it2, code2 = sph.newton_smoothlength_arr(to(pos), h_t, _details=True)
        assert (it2 == 1).all() and (code2 == 0).all()
        # sampled particles against the oracle
        js = np.asarray(js)
        h_ref, it_ref, code_ref, _ = o.newton_smoothlength_while(js, pos, w["h_guess"][js])
        assert rel(h[js], h_ref) < RTOL and (it[js] == it_ref).all()
        assert rel(om[js], o.omega_j(js, pos, rho[js], h[js])) < RTOL
        assert vrel(acc[js], o.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(n), j=js)) < RTOL
        d_ref = o.energy_rate_arr(pos, vel, P, rho, h, om, j=js)
        assert np.abs(dudt[js] - d_ref).max() < RTOL * np.abs(dudt).max()
        if n_expect is not None:
            jn, hn, nn = n_expect
            assert len(o.neighbours(pos, jn, h[jn])) == nn and h[jn] == pytest.approx(hn, rel=1e-6)
        # the neighbour COUNT the hot kernels see (q < 2 decided with their own r = d2 rsqrt(d2), q = r / h) equals the
        # size of the reference-arithmetic neighbour set, for the scanning kernel and for the list drain
        import engine as E
        eng = sph._eng()
        c = sph._consts()
        dg = eng.build(eng.pack_posh(to(pos), h_t), *eng.grid_for(eng.pack_posh(to(pos), h_t)))
        _, _, _, ng_scan = eng.density_omega(dg, c, want_rho=False, want_ngb=True)
        nl = eng.nl_build(dg, 1.03, 0.03)
        _, _, _, ng_list = eng.nl_density_omega(nl, c, dg.posh, want_omega=False, want_ngb=True)
        assert nl.read_flags() == 0 and torch.equal(ng_scan, ng_list)
        ng = eng.scatter(ng_scan.double(), dg.perm, 1).cpu().numpy()
        for j in js[:24]:
            assert int(ng[j]) == len(o.neighbours(pos, int(j), h[j])), j
        del nl, dg
        return it
    finally:
        sph.PARTICLE_MASS = 1


def test_C2_planar_262144(sph, oracle_mod):
    """pyfluidex2d_var_h.py settings scaled: z = 0 sheet, cold solve from h = 1"""
    import workloads
    w = workloads.planar_uniform(262144, seed=0, u=1.0)
    it = _check(sph, oracle_mod, w, js=[0, 1000, 131072, 262143])
    assert 5 < it.mean() < 14   # SURVEY: ~9.5 cold Newton iterations


def test_C3_lattice_1M(sph, oracle_mod):
    """pyfluid_newton_h_testing.py settings scaled (C3): 100^3 lattice at number density 3.75, guess 0.8"""
    import workloads
    w = workloads.lattice_cube(100)
    centre = 50 * 100 * 100 + 50 * 100 + 50
    it = _check(sph, oracle_mod, w, js=[0, centre, 999999], n_expect=(centre, 0.837525, 81))
    assert it[centre] == 3   # SURVEY: 3 cold iterations in the interior


def test_C4_sod_16M(sph, oracle_mod):
    """the benchmark configuration itself (C4): 15 925 248 particles, 8:1 density jump"""
    import workloads
    w = workloads.sod_tube(384)
    assert len(w["pos"]) == 15925248
    # left bulk edge, both sides of the interface, and a seeded sample of 61 more
    js = [7, 14155776 - 3, 14155776 + 5] + np.random.default_rng(9).choice(15925248, size=61, replace=False).tolist()
    _check(sph, oracle_mod, w, js=js)
