"""Per-particle masses (the reference authors' own NOTE at pyfluid.py:3-4).  The reference has ONE
scalar PARTICLE_MASS, so there is no reference output for two mass species: the oracle's mass array is pinned by
"reduces to the golden vectors when uniform" (tests/test_oracle_golden.py) and the CUDA path is compared with it here.
Tolerance 1e-10 as everywhere."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = np.where(a == b, 0.0, np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
    return float(np.max(d)) if d.size else 0.0


def vrel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-300)))


@pytest.fixture()
def sph():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import build_native
    build_native.build()
    import pyfluid
    pyfluid.G = 0.0
    yield pyfluid
    pyfluid.PARTICLE_MASS = 1


def _cloud(n, seed):
    rng = np.random.default_rng(seed)
    L = (n / 3.75) ** (1 / 3)
    pos = rng.random((n, 3)) * L
    vel = (rng.random((n, 3)) - 0.5) * 0.5
    u = 1 + rng.random(n)
    m = np.where(rng.random(n) < 0.5, 1.0, 2.5)      # two species
    return pos, vel, u, m


def test_two_species_stage_and_step_vs_oracle(sph, oracle_mod):
    import torch
    import engine as E
    n = 3000
    pos, vel, u, m = _cloud(n, 21)
    o = oracle_mod.Oracle(G=0.0)
    o.set_num_threads(os.cpu_count() or 1)
    o.set_masses(m)
    try:
        sph.PARTICLE_MASS = m
        h = sph.newton_smoothlength_arr(pos, np.full(n, 0.8))
        h_ref = o.newton_smoothlength_arr(pos, np.full(n, 0.8))
        assert rel(h, h_ref) < RTOL
        # the heavy species needs the larger h at equal number density: rho_j = m_j (eta / h_j)^3
        assert h[m > 2].mean() > 1.05 * h[m < 2].mean()
        rho = sph.var_density_arr(h)
        assert rel(rho, m * (1.3 / h) ** 3) < 1e-14
        assert rel(sph.density_arr(pos, h), o.density_arr(pos, h)) < RTOL
        P = sph.pressure_arr(u, rho)
        om = sph.omega_arr(pos, rho, h)
        assert rel(om, o.omega_arr(pos, rho, h)) < RTOL
        acc = sph.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(n))
        assert vrel(acc, o.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(n))) < RTOL
        dudt = sph.energy_rate_arr(pos, vel, P, rho, h, om)
        d_ref = o.energy_rate_arr(pos, vel, P, rho, h, om)
        assert np.abs(dudt - d_ref).max() < RTOL * np.abs(d_ref).max()
        # momentum: sum_j m_j a_j = 0 (the pair terms are antisymmetric once weighted by both masses)
        assert np.abs((m[:, None] * acc).sum(axis=0)).max() < 1e-11 * np.abs(m[:, None] * acc).sum()
        # one step: the module call (kept lists), the scanning Stepper and the FastStepper against the oracle
        dt = 1e-3
        ref = o.var_smoothlength_leapfrog(pos, vel, u, h, dt)
        got = sph.var_smoothlength_leapfrog(pos, vel, u, h, dt)
        for a, b in zip(got, ref):
            assert (rel(a, b) < RTOL) if a.ndim == 1 else (vrel(a, b) < RTOL)
        eng = sph._eng()
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        c = E.make_consts(dict({k: getattr(sph, k) for k in sph._CONST_NAMES}, PARTICLE_MASS=float("nan")))
        for stepper in (E.Stepper(eng), E.FastStepper(eng)):
            r = [x.cpu().numpy() for x in stepper.step(to(pos), to(vel), to(u), to(h), dt, c, mass=to(m))]
            for a, b in zip(r, ref):
                assert (rel(a, b) < RTOL) if a.ndim == 1 else (vrel(a, b) < RTOL)
        # several steps on kept lists (masses travel through the rebuilds) against the oracle
        fs = E.FastStepper(eng)
        fs.load(to(pos), to(vel), to(u), to(h), consts=c, mass=to(m))
        s = (pos, vel, u, h)
        for _ in range(3):
            fs.advance(dt, c, defer_status=False)
            s = o.var_smoothlength_leapfrog(*s, dt)
        r = [x.cpu().numpy() for x in fs.state()]
        assert rel(r[3], s[3]) < RTOL and rel(r[2], s[2]) < RTOL and vrel(r[1], s[1]) < 1e-9 and vrel(r[0], s[0]) < RTOL
        assert fs.stats["builds"] >= 2 and fs.stats["safe_steps"] == 0
    finally:
        o.set_masses(None)


def test_equal_mass_array_matches_the_scalar(sph, oracle_mod):
    """an array of equal masses gives the scalar's fields to rounding (m_i inside the sums instead of one factor after
    them); the scalar path itself is untouched: the same kernels, the same bits as before"""
    n = 2000
    pos, vel, u, _ = _cloud(n, 22)
    sph.PARTICLE_MASS = 0.37
    h0 = sph.newton_smoothlength_arr(pos, np.full(n, 0.8))
    s0 = sph.var_smoothlength_leapfrog(pos, vel, u, h0, 1e-3)
    sph.PARTICLE_MASS = np.full(n, 0.37)
    h1 = sph.newton_smoothlength_arr(pos, np.full(n, 0.8))
    s1 = sph.var_smoothlength_leapfrog(pos, vel, u, h1, 1e-3)
    assert rel(h1, h0) < 1e-13
    for a, b in zip(s1, s0):
        assert (rel(a, b) < 1e-12) if a.ndim == 1 else (vrel(a, b) < 1e-12)
    with pytest.raises(ValueError):
        sph.PARTICLE_MASS = np.full(n - 1, 0.37)
        sph.density_arr(pos, h0)
