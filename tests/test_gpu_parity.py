"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the drop-in module and
the C-ABI, against (1) the golden fixtures produced by the reference itself and (2) the CPU oracle on
the same seeded inputs.

Tolerances (BASELINE.json north_star): neighbour sets and cell assignments bit-exact; density, h,
Omega, acceleration, du/dt within 1e-10 relative in fp64.  Scalars compare relative to the value;
vectors relative to the largest component of the row.
"""
import json
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
    scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-300)
    return float(np.max(np.abs(a - b) / scale))


@pytest.fixture(scope="module")
def sph():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import build_native
    build_native.build()
    import pyfluid
    for k, v in dict(PARTICLE_MASS=1, COUPLING_CONST=1.3, G=0.0).items():
        setattr(pyfluid, k, v)
    return pyfluid


@pytest.fixture(scope="module")
def orc(oracle_mod):
    return oracle_mod.Oracle(G=0.0)


def uniform_cloud(n, seed, density=3.75, vscale=1.0):
    rng = np.random.default_rng(seed)
    L = (n / density) ** (1 / 3)
    return rng.random((n, 3)) * L, (rng.random((n, 3)) - 0.5) * vscale, 1 + rng.random(n), L


def test_direction_and_dellM4_golden(sph):
    """direction_i_j (pyfluid.py:50-54) and dellM4 (:81-82) against outputs of the reference (make_golden_r2.py)"""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pyfluid_golden_r2.npz"))
    for k in range(len(g["r2_h"])):
        d = sph.direction_i_j(g["r2_rj"][k], g["r2_ri"][k])
        w = sph.dellM4(g["r2_rj"][k], g["r2_ri"][k], float(g["r2_h"][k]))
        assert isinstance(d, np.ndarray) and d.shape == (3,)
        assert np.array_equal(d, g["r2_direction"][k])              # three IEEE divisions: the same bits
        assert vrel(w[None], g["r2_dellM4"][k][None]) < RTOL or not g["r2_dellM4"][k].any()
    assert not g["r2_direction"][5].any() and not sph.direction_i_j(g["r2_rj"][5], g["r2_ri"][5]).any()


# ---- kernel function KATs ---------------------------------------------------------------------------
def test_m4_kats(sph, golden):
    d, h = golden["kat_dist"], golden["kat_h"]
    for fn, key in ((sph.M4, "kat_M4"), (sph.M4_d1, "kat_M4_d1"), (sph.M4_smoothlength_derivative, "kat_M4_dh")):
        got = fn(d, h)
        ref = golden[key]
        # dW/dh cancels to ~0 at q = 1 and values at the branch points may come from the neighbouring
        # (continuous) branch: measure against the size of the individual terms, 1/h^4
        scale = np.maximum(np.abs(ref), 1.0 / np.abs(h) ** 4)
        assert np.all((np.abs(got - ref) <= 1e-13 * scale) | (got == ref)), (key, np.max(np.abs(got - ref) / scale))
    assert sph.M4(0.5, 1) == pytest.approx(0.22878523069459955, rel=1e-15)
    assert sph.M4(2, 1) == 0
    assert sph.distance(np.array([1.0, 2.0, 3.0]), np.array([0.5, -1.0, 2.0])) == float(
        np.sqrt(np.array([0.5, 3.0, 1.0]).dot(np.array([0.5, 3.0, 1.0]))))


# ---- K1: cell list is bit-exact against the numpy restatement ------------------------------------------
@pytest.mark.parametrize("n,seed,planar", [(1, 0, False), (37, 1, False), (5000, 2, False), (3000, 3, True), (200000, 4, False)])
def test_grid_build_bit_exact(sph, n, seed, planar):
    import torch
    import engine as E
    import gridref
    e = sph._eng()
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3)) * np.array([7.0, 3.0, 5.0]) - 1.0
    if planar:
        pos[:, 2] = 0.0
    h = 0.2 + 0.3 * rng.random(n)
    posh = e.pack_posh(torch.from_numpy(pos).cuda(), torch.from_numpy(h).cuda())
    grid, delta = e.grid_for(posh)
    dg = e.build(posh, grid, delta)
    dims = (grid.nx, grid.ny, grid.nz)
    ids, perm, start = gridref.build(pos, (grid.ox, grid.oy, grid.oz), grid.cell, dims, grid.xdiv)
    assert (dg.perm.cpu().numpy().astype(np.int64) == perm).all()
    assert (dg.cell_start.cpu().numpy().astype(np.int64) == start).all()
    assert (dg.cell_id.cpu().numpy().astype(np.int64) == ids[perm]).all()
    assert (dg.posh.cpu().numpy() == np.concatenate([pos, h[:, None]], 1)[perm]).all()
    hm = np.zeros(grid.ncell)
    np.maximum.at(hm, ids, h)
    got = dg.cell_hmax.cpu().numpy().astype(np.float64)
    assert (got >= hm).all() and (got <= hm * (1 + 2e-7)).all()
    # permutation helpers round-trip
    x = torch.from_numpy(rng.random((n, 3))).cuda()
    assert torch.equal(e.scatter(e.gather(x, dg.perm, 3), dg.perm, 3), x)


def test_grid_outliers_are_clamped(sph):
    """particles outside the box land in border cells and are still found"""
    import torch
    import engine as E
    e = sph._eng()
    pos, _, _, L = uniform_cloud(600, 9)
    h = np.full(600, 0.9)
    posh = e.pack_posh(torch.from_numpy(pos).cuda(), torch.from_numpy(h).cuda())
    grid, delta = E.choose_grid([0.3 * L] * 3, [0.6 * L] * 3, 0.9, 600, pad_cells=0.0)   # a box much smaller than the cloud
    dg = e.build(posh, grid, delta)
    rho, _, _, _ = e.density_omega(dg, sph._consts())
    rho = e.scatter(rho, dg.perm, 1).cpu().numpy()
    ref = sph.density_arr(pos, h)
    assert rel(rho, ref) < 1e-13


# ---- golden fixtures (outputs of the reference itself) ---------------------------------------------------
def test_golden_B_stage(sph, golden):
    g = golden
    pos, vel, e, h = g["B_pos"], g["B_vel"], g["B_e"], g["B_h"]
    assert rel(sph.newton_smoothlength_arr(pos, np.full(24, 0.8)), h) < RTOL
    assert rel(sph.density_arr(pos, h), g["B_density_sum"]) < RTOL
    rho = sph.var_density_arr(h)
    assert rel(rho, g["B_den"]) < 1e-15
    om = sph.omega_arr(pos, rho, h)
    assert rel(om, g["B_omega"]) < RTOL
    P = sph.pressure_arr(e, rho)
    assert rel(P, g["B_P"]) < 1e-15
    assert rel(sph.energy_rate_arr(pos, vel, P, rho, h, om), g["B_dudt"]) < RTOL
    assert vrel(sph.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(24)), g["B_acc_G0"]) < RTOL
    # per-particle entry points
    for j in (0, 7, 23):
        assert sph.density(j, pos, h[j]) == pytest.approx(g["B_density_sum"][j], rel=RTOL)
        assert sph.omega_j(j, pos, rho[j], h[j]) == pytest.approx(g["B_omega"][j], rel=RTOL)
        assert sph.energy_rate(j, pos, vel, P, rho, h, om) == pytest.approx(g["B_dudt"][j], rel=RTOL)
        assert vrel(sph.acceleration(j, pos, rho, vel, P, h, om, np.zeros(24)), g["B_acc_G0"][j]) < RTOL
        assert sph.newton_smoothlength_while(j, pos, 0.8) == pytest.approx(h[j], rel=RTOL)


def test_golden_B_leapfrog(sph, golden):
    g = golden
    p1, v1, e1, h1 = sph.var_smoothlength_leapfrog(g["B_pos"], g["B_vel"], g["B_e"], g["B_h"], 1e-3)
    assert rel(p1, g["B_lf_G0_pos"]) < RTOL and vrel(v1, g["B_lf_G0_vel"]) < RTOL
    assert rel(e1, g["B_lf_G0_e"]) < RTOL and rel(h1, g["B_lf_G0_h"]) < RTOL


def test_golden_A_newton_paths(sph, golden, capsys):
    g = golden
    pos = g["A_pos"]
    for guess, h_ref, it_ref, warned in zip(g["A_guess"], g["A_h"], g["A_iters"], g["A_warned"]):
        h, it, code = sph.newton_smoothlength_while(0, pos, float(guess), _details=True)
        out = capsys.readouterr().out
        assert it == it_ref and (code != 0) == bool(warned)
        assert ("WARNING: Failure to converge in newton_new_h." in out) == bool(warned)
        assert h == pytest.approx(h_ref, rel=RTOL)
    hb = sph.bisection_h(0, pos)
    assert hb == float(g["A_bisection_h0"])          # dyadic midpoints: bit-exact
    assert sph.density(0, pos, 0.8159916383864595) == pytest.approx(float(g["A_density_j0"]), rel=RTOL)
    assert sph.newton_smoothlength_iteration(0, pos, 0.8) == pytest.approx(
        0.8 - (sph.zeta_j(0.8, sph.density(0, pos, 0.8))) / sph.zetaprime(
            sph.density(0, pos, 0.8), sph.omega_j(0, pos, sph.density(0, pos, 0.8), 0.8), 0.8), rel=1e-14)


def test_golden_C_sim_and_neighbours(sph, golden, capsys):
    import torch
    g = golden
    pos, vel, e = g["C_pos"], g["C_vel"], g["C_e"]
    P, V, E, H = sph.var_smoothlength_sim(np.array([0.0, 2e-3, 4e-3]), pos, vel, e, np.full(64, 0.8))
    assert capsys.readouterr().out == "time step: 0 of 3\ntime step: 1 of 3\n"
    assert rel(H, g["C_sim_h"]) < RTOL and rel(P, g["C_sim_pos"]) < RTOL
    assert vrel(V, g["C_sim_vel"]) < RTOL and rel(E, g["C_sim_e"]) < RTOL
    # neighbour sets: exact
    eng = sph._eng()
    h = g["C_h"]
    posh = eng.pack_posh(torch.from_numpy(pos).cuda(), torch.from_numpy(h).cuda())
    grid, delta = eng.grid_for(posh)
    dg = eng.build(posh, grid, delta)
    perm = dg.perm.cpu().numpy().astype(np.int64)
    out, cnt = eng.neighbours(dg, torch.arange(64, dtype=torch.int32, device="cuda"), 64)
    out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
    for s in range(64):
        got = np.sort(perm[out[s, :cnt[s]]])
        assert (got == np.flatnonzero(g["C_nb"][perm[s]])).all()


def test_golden_D_planar(sph, golden):
    g = golden
    assert rel(sph.newton_smoothlength_arr(g["D_pos"], np.full(50, 1.0)), g["D_h"]) < RTOL
    p1, v1, e1, h1 = sph.var_smoothlength_leapfrog(g["D_pos"], g["D_vel"], g["D_e"], g["D_h"], 0.1)
    assert rel(p1, g["D_lf_pos"]) < RTOL and rel(e1, g["D_lf_e"]) < RTOL and rel(h1, g["D_lf_h"]) < RTOL
    np.testing.assert_allclose(v1, g["D_lf_vel"], rtol=1e-9, atol=1e-70)


def test_golden_E_bisection_fallback(sph, golden, capsys):
    g = golden
    h, it, code = sph.newton_smoothlength_arr(g["C_pos"], np.full(64, float(g["E_guess"])), _details=True)
    out = capsys.readouterr().out
    assert ((code != 0) == g["E_warned"]).all() and (it == g["E_iters"]).all()
    assert out.count("WARNING: Failure to converge in newton_new_h.") == int(g["E_warned"].sum())
    assert rel(h, g["E_h"]) < RTOL


def test_golden_F_no_root_gives_nan(sph, golden, capsys):
    meta = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pyfluid_golden.json")))
    h = sph.newton_smoothlength_arr(golden["F_pos"], np.full(4, 0.2))
    out = capsys.readouterr().out
    assert np.isnan(h).all()
    for line, count in meta["F_stdout_counts"].items():
        assert out.count(line) == count, line
    assert sph.bisection_h(0, golden["F_pos"]) is None


def test_golden_G_coincident(sph, golden):
    g = golden
    pos, vel, e, h = g["G_pos"], g["G_vel"], g["G_e"], g["G_h"]
    assert rel(sph.newton_smoothlength_arr(pos, np.full(64, 0.8)), h) < RTOL
    rho = sph.var_density_arr(h); P = sph.pressure_arr(e, rho); om = sph.omega_arr(pos, rho, h)
    assert rel(om, g["G_omega"]) < RTOL
    assert rel(sph.energy_rate_arr(pos, vel, P, rho, h, om), g["G_dudt"]) < RTOL
    assert vrel(sph.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(64)), g["G_acc_G0"]) < RTOL


def test_error_conventions(sph, golden):
    g = golden
    with pytest.raises(RuntimeError, match="Obtained negative energy."):
        sph.var_smoothlength_leapfrog(g["B_pos"], g["B_vel"] * 5.0, np.full(24, 1e-3), g["B_h"], 0.5)
    with pytest.raises(RuntimeError, match="Obtained negative pressure."):
        sph.pressure_arr(np.array([1.0, -1.0]), np.array([1.0, 1.0]))


def test_constants_read_at_call_time(sph, golden, oracle_mod):
    g = golden
    sph.PARTICLE_MASS, sph.COUPLING_CONST = 2.5, 1.2
    try:
        o = oracle_mod.Oracle(G=0.0, PARTICLE_MASS=2.5, COUPLING_CONST=1.2)
        h = sph.newton_smoothlength_arr(g["C_pos"], np.full(64, 0.7))
        assert rel(h, o.newton_smoothlength_arr(g["C_pos"], np.full(64, 0.7))) < RTOL
        assert rel(sph.density_arr(g["C_pos"], h), o.density_arr(g["C_pos"], h)) < RTOL
    finally:
        sph.PARTICLE_MASS, sph.COUPLING_CONST = 1, 1.3


def test_torch_tensors_in_and_out(sph, golden):
    import torch
    g = golden
    pos = torch.from_numpy(g["B_pos"]).cuda()
    h = torch.from_numpy(g["B_h"]).cuda()
    rho = sph.density_arr(pos, h)
    assert isinstance(rho, torch.Tensor) and rho.is_cuda
    assert rel(rho.cpu().numpy(), g["B_density_sum"]) < RTOL


# ---- oracle on seeded inputs, sizes the O(N^2) oracle finishes in seconds -----------------------------------
@pytest.mark.parametrize("n,seed", [(33, 11), (1000, 12), (6000, 13)])
def test_stage_vs_oracle(sph, orc, n, seed):
    pos, vel, e, L = uniform_cloud(n, seed)
    h = sph.newton_smoothlength_arr(pos, np.full(n, 0.8))
    h_ref, it_ref, code_ref = orc.newton_smoothlength_arr(pos, np.full(n, 0.8), details=True)
    assert rel(h, h_ref) < RTOL
    _, it, code = sph.newton_smoothlength_arr(pos, np.full(n, 0.8), _details=True)
    assert (it == it_ref).all() and (code == code_ref).all()
    rho = sph.var_density_arr(h_ref)
    om = sph.omega_arr(pos, rho, h_ref)
    om_ref = orc.omega_arr(pos, rho, h_ref)
    assert rel(om, om_ref) < RTOL
    assert rel(sph.density_arr(pos, h_ref), orc.density_arr(pos, h_ref)) < RTOL
    P = sph.pressure_arr(e, rho)
    assert vrel(sph.acceleration_arr(pos, rho, vel, P, h_ref, om_ref, np.zeros(n)),
                orc.acceleration_arr(pos, rho, vel, P, h_ref, om_ref, np.zeros(n))) < RTOL
    d_ref = orc.energy_rate_arr(pos, vel, P, rho, h_ref, om_ref)
    d = sph.energy_rate_arr(pos, vel, P, rho, h_ref, om_ref)
    assert np.max(np.abs(d - d_ref)) < RTOL * np.max(np.abs(d_ref))


def test_variable_h_two_density_regions(sph, orc):
    """a Sod-like 8:1 density jump: h differs by 2x across the interface (symmetric-support search)"""
    a = 0.1
    gl = np.arange(12) * a
    left = np.stack(np.meshgrid(gl, gl, gl, indexing="ij"), -1).reshape(-1, 3) + a * 0.5
    left[:, 0] -= 12 * a
    gr = np.arange(6) * 2 * a
    right = np.stack(np.meshgrid(gr, gr, gr, indexing="ij"), -1).reshape(-1, 3) + a
    pos = np.concatenate([left, right])
    rng = np.random.default_rng(3)
    pos += (rng.random(pos.shape) - 0.5) * 0.2 * a
    n = len(pos)
    vel = (rng.random((n, 3)) - 0.5) * 0.3
    e = 1 + rng.random(n)
    sph.PARTICLE_MASS = a ** 3
    try:
        o = type(orc)(G=0.0, PARTICLE_MASS=a ** 3)
        guess = np.where(pos[:, 0] < 0, 1.2 * a, 2.4 * a)
        h = sph.newton_smoothlength_arr(pos, guess)
        h_ref = o.newton_smoothlength_arr(pos, guess)
        assert rel(h, h_ref) < RTOL
        assert h_ref.max() / h_ref.min() > 1.7
        p1, v1, e1, h1 = sph.var_smoothlength_leapfrog(pos, vel, e, h_ref, 1e-3)
        q1, w1, f1, g1 = o.var_smoothlength_leapfrog(pos, vel, e, h_ref, 1e-3)
        assert rel(p1, q1) < RTOL and vrel(v1, w1) < RTOL and rel(e1, f1) < RTOL and rel(h1, g1) < RTOL
        # symmetric neighbour sets, exact
        import torch
        eng = sph._eng()
        posh = eng.pack_posh(torch.from_numpy(pos).cuda(), torch.from_numpy(h_ref).cuda())
        grid, delta = eng.grid_for(posh)
        dg = eng.build(posh, grid, delta)
        perm = dg.perm.cpu().numpy().astype(np.int64)
        q = torch.arange(0, n, 7, dtype=torch.int32, device="cuda")
        out, cnt = eng.neighbours(dg, q, 512, symmetric=True)
        out, cnt, q = out.cpu().numpy(), cnt.cpu().numpy(), q.cpu().numpy()
        for k, s in enumerate(q):
            assert (np.sort(perm[out[k, :cnt[k]]]) == o.neighbours_sym(pos, h_ref, perm[s])).all()
    finally:
        sph.PARTICLE_MASS = 1


def test_leapfrog_vs_oracle_2000(sph, orc):
    n = 2000
    pos, vel, e, L = uniform_cloud(n, 21, vscale=2.0)
    h = orc.newton_smoothlength_arr(pos, np.full(n, 0.8))
    p1, v1, e1, h1 = sph.var_smoothlength_leapfrog(pos, vel, e, h, 2e-3)
    q1, w1, f1, g1 = orc.var_smoothlength_leapfrog(pos, vel, e, h, 2e-3)
    assert rel(p1, q1) < RTOL and vrel(v1, w1) < RTOL and rel(e1, f1) < RTOL and rel(h1, g1) < RTOL


def test_conservation_properties(sph):
    """size-independent properties: sum m a = 0 and sum (v.a + du/dt) = 0 with G = 0"""
    n = 50000
    pos, vel, e, L = uniform_cloud(n, 31)
    h = sph.newton_smoothlength_arr(pos, np.full(n, 0.8))
    rho = sph.var_density_arr(h); P = sph.pressure_arr(e, rho); om = sph.omega_arr(pos, rho, h)
    acc = sph.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(n))
    dudt = sph.energy_rate_arr(pos, vel, P, rho, h, om)
    scale = np.abs(acc).sum()
    assert np.abs(acc.sum(axis=0)).max() < 1e-11 * scale
    assert abs(((vel * acc).sum(axis=1) + dudt).sum()) < 1e-11 * (np.abs(vel * acc).sum() + np.abs(dudt).sum())
    # idempotence of the h solve: starting from the converged h it stops after one iteration
    h2, it, code = sph.newton_smoothlength_arr(pos, h, _details=True)
    assert (it == 1).all() and (code == 0).all()
    # sampled-j parity at this size against the per-particle entry points of the oracle
    import oracle
    o = oracle.Oracle(G=0.0)
    js = np.array([0, 1234, 49999])
    assert rel(h[js], o.newton_smoothlength_while(js, pos, 0.8)[0]) < RTOL
    assert rel(om[js], o.omega_j(js, pos, rho[js], h[js])) < RTOL
    assert vrel(acc[js], o.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(n), j=js)) < RTOL


def test_golden_P_pair_functions(sph, golden):
    g = golden
    pos, vel, h, P, den, om = g["C_pos"], g["C_vel"], g["C_h"], g["C_P"], g["C_den"], g["C_omega"]
    for (j, i), pi_ref, acc_ref in zip(g["P_pairs"], g["P_Pi"], g["P_acc"]):
        pi = sph.Pi(int(j), int(i), pos, vel, P, den, h)
        assert pi == pytest.approx(pi_ref, rel=RTOL, abs=0.0)
        a = sph.fluid_accel_viscosity_comp(int(j), int(i), pos, vel, den, P, h, om)
        if j == i:
            assert isinstance(a, int) and a == 0
        else:
            assert vrel(a, acc_ref) < RTOL or np.abs(acc_ref).max() == 0.0
    assert (g["P_Pi"] > 0).any() and (g["P_Pi"] == 0).any()


@pytest.mark.parametrize("guess", [0.8, 0.45, 1.15])
def test_newton_fused_omega(sph, orc, guess):
    """Omega returned by the Newton kernel (one more pass over the converged iteration's neighbour list, or the
    bisection kernel's own sum) equals omega_arr at the solved h"""
    import torch
    import engine as E
    n = 3000
    pos, vel, e, L = uniform_cloud(n, 41)
    eng = sph._eng()
    dg = sph._grid_and_order(eng, torch.from_numpy(pos).cuda(), torch.full((n,), guess, dtype=torch.float64, device="cuda"))
    st = E.Status(1, eng.device)
    h_s, om_s = eng.newton(dg, sph._consts(), dg.posh[:, 3].contiguous(), st.ptr(0), want_omega=True)
    h = eng.scatter(h_s, dg.perm, 1).cpu().numpy()
    om = eng.scatter(om_s, dg.perm, 1).cpu().numpy()
    assert rel(h, orc.newton_smoothlength_arr(pos, np.full(n, guess))) < RTOL
    ok = np.isfinite(h)
    rho = orc.var_density_arr(h[ok])
    om_ref = orc.omega_j(np.flatnonzero(ok), pos, rho, h[ok])
    assert rel(om[ok], om_ref) < RTOL


def test_sums_do_not_depend_on_warp_chunking(sph):
    """every per-particle sum has ONE order and ONE rounding: prepending far-away particles (which shifts how
    targets fall into warps) or masking targets must not change a single bit of rho, Omega, a, du/dt.
    This is what makes k-rank runs bit-identical to 1-rank runs."""
    import torch
    import engine as E
    import workloads
    eng = sph._eng()
    w = workloads.sod_tube(48)
    sph.PARTICLE_MASS = w["m"]
    try:
        c = sph._consts()
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        pos, vel, u = to(w["pos"]), to(w["vel"]), to(w["u"])
        h0 = to(w["h_guess"] * 1.0843)
        lo, hi, hmax, hmean = eng.bounds(eng.pack_posh(pos, h0))
        grid, delta = E.choose_grid(lo, hi, hmean, pos.shape[0])

        def run(shift=0, mask_every=0):
            p, hh, vv, uu = pos, h0, vel, u
            if shift:
                far = torch.full((shift, 3), -10.0, dtype=torch.float64, device="cuda")
                far += torch.arange(shift, device="cuda", dtype=torch.float64)[:, None] * 1e-3
                p = torch.cat([far, p]); hh = torch.cat([h0[:shift], hh]); vv = torch.cat([vel[:shift] * 0, vv]); uu = torch.cat([u[:shift], uu])
            dg = eng.build(eng.pack_posh(p, hh), grid, delta)
            st = E.Status(1, eng.device)
            tgt = None
            if mask_every:
                tgt = torch.ones(p.shape[0], dtype=torch.uint8, device="cuda")
                tgt[::mask_every] = 0
                tgt = tgt[dg.perm.long()].contiguous()
            rho, _, om, _ = eng.density_omega(dg, c, target=tgt, want_rho=True, want_omega=True)
            _, _, om_all, _ = eng.density_omega(dg, c, want_rho=False, want_omega=True)
            _, _, rB, rC = eng.eos(dg.posh, eng.gather(vv, dg.perm, 3), eng.gather(uu, dg.perm, 1), om_all, c, st.ptr(0))
            acc, dudt = eng.force(dg, c, rB, rC, eng.hmax(dg.posh), st.ptr(0), target=tgt)
            hs, oms = eng.newton(dg, c, dg.posh[:, 3].contiguous(), st.ptr(0), target=tgt, want_omega=True)
            return [eng.scatter(x, dg.perm, k)[shift:] for x, k in ((acc, 3), (dudt, 1), (rho, 1), (om, 1), (hs, 1), (oms, 1))]

        ref = run()
        for shift, me in ((5, 0), (17, 0), (0, 3)):
            got = run(shift, me)
            keep = torch.ones(pos.shape[0], dtype=torch.bool, device="cuda")
            if me:
                keep[::me] = False
            for name, a, b in zip(("acc", "dudt", "rho", "omega", "h", "omega_fused"), ref, got):
                assert torch.equal(a[keep], b[keep]), (name, shift, me)
    finally:
        sph.PARTICLE_MASS = 1


def test_omega_carry_does_not_change_a_bit(sph):
    """the time loop hands Omega(pos1, h1) from one step's last Newton kernel to the next step (var_smoothlength_sim);
    a step that receives it must return exactly what a stand-alone step returns"""
    import torch
    import engine as E
    n = 5000
    pos, vel, e, L = uniform_cloud(n, 77, vscale=0.7)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    h = to(sph.newton_smoothlength_arr(pos, np.full(n, 0.8)))
    stp = E.Stepper(sph._eng())
    c = sph._consts()
    s1 = stp.step(to(pos), to(vel), to(e), h, 2e-3, c, want_omega1=True)
    plain = stp.step(s1[0], s1[1], s1[2], s1[3], 2e-3, c)
    carried = stp.step(s1[0], s1[1], s1[2], s1[3], 2e-3, c, omega0=s1[4])
    for a, b in zip(plain, carried):
        assert torch.equal(a, b)


# ---- softened self-gravity: the reference's default G -------------------------------------
def test_gravity_kats_and_golden(sph, golden):
    g = golden
    d = g["kat_grav_dist"] * 1.3
    h = np.full(len(d), 1.3)
    with np.errstate(all="ignore"):
        for fn, key in ((sph.grav_potential, "kat_grav_phi"), (sph.grav_force, "kat_grav_force"),
                        (sph.grav_potential_smoothlength_derivative, "kat_grav_dphidh")):
            got, ref = fn(d, h), g[key]
            ok = np.isclose(got, ref, rtol=1e-13, atol=1e-13) | (np.isinf(got) & np.isinf(ref)) | (np.isnan(got) & np.isnan(ref))
            assert ok.all(), (key, got, ref)
    sph.G = 39.4227
    try:
        pos, vel, h, P, den, om = g["C_pos"], g["C_vel"], g["C_h"], g["C_P"], g["C_den"], g["C_omega"]
        xi = sph.xi_arr(den, h, pos)
        assert rel(xi, g["C_xi"]) < RTOL
        assert sph.xi(5, pos, den[5], h[5]) == pytest.approx(g["C_xi"][5], rel=RTOL)
        assert vrel(sph.acceleration_arr(pos, den, vel, P, h, om, xi), g["C_acc_G"]) < RTOL
        assert vrel(sph.acceleration(9, pos, den, vel, P, h, om, xi), g["C_acc_G"][9]) < RTOL
        for (j, i), gr, co in zip(g["P_pairs"], g["P_grav"], g["P_corr"]):
            a = sph.smoothed_gravity_acceleration_comp(int(j), int(i), pos, h)
            b = sph.smoothed_gravity_correction_comp(int(j), int(i), pos, h, om, xi)
            if j == i:
                assert isinstance(a, int) and a == 0
            else:
                assert np.abs(a - gr).max() <= RTOL * max(np.abs(gr).max(), 1e-300)
            assert np.abs(b - co).max() <= RTOL * max(np.abs(co).max(), 1e-300) or np.abs(co).max() == 0.0
        # case B: stage and one leapfrog step with the default G
        pos, vel, e, h = g["B_pos"], g["B_vel"], g["B_e"], g["B_h"]
        xi = sph.xi_arr(g["B_den"], h, pos)
        assert rel(xi, g["B_xi"]) < RTOL
        assert vrel(sph.acceleration_arr(pos, g["B_den"], vel, g["B_P"], h, g["B_omega"], xi), g["B_acc_G"]) < RTOL
        p1, v1, e1, h1 = sph.var_smoothlength_leapfrog(pos, vel, e, h, 1e-3)
        assert rel(p1, g["B_lf_G_pos"]) < RTOL and vrel(v1, g["B_lf_G_vel"]) < RTOL
        assert rel(e1, g["B_lf_G_e"]) < RTOL and rel(h1, g["B_lf_G_h"]) < RTOL
    finally:
        sph.G = 0.0


def test_gravity_leapfrog_vs_oracle(sph, oracle_mod):
    n = 1500
    pos, vel, e, L = uniform_cloud(n, 91, vscale=0.5)
    o = oracle_mod.Oracle()   # default G = 39.4227
    h = o.newton_smoothlength_arr(pos, np.full(n, 0.8))
    sph.G = 39.4227
    try:
        p1, v1, e1, h1 = sph.var_smoothlength_leapfrog(pos, vel, e, h, 1e-3)
    finally:
        sph.G = 0.0
    q1, w1, f1, g1 = o.var_smoothlength_leapfrog(pos, vel, e, h, 1e-3)
    assert rel(p1, q1) < RTOL and vrel(v1, w1) < RTOL and rel(e1, f1) < RTOL and rel(h1, g1) < RTOL


def test_golden_S_static_h_config1(sph, golden):
    """BASELINE config 1 (pyfluidex2d_static_h.py settings) through the repaired static path:
    three steps against the same shim applied to the reference's own functions, with G = 0 and with the default G"""
    g = golden
    for gname, gval in (("G0", 0.0), ("G", 39.4227)):
        sph.G = gval
        try:
            p, v, e = g["S_%s_pos" % gname][0], g["S_%s_vel" % gname][0], g["S_%s_e" % gname][0]
            for t in range(1, 4):
                p, v, e = sph.static_smoothlength_leapfrog(p, v, e, 0.1)
                assert rel(p, g["S_%s_pos" % gname][t]) < RTOL, (gname, t)
                assert vrel(v, g["S_%s_vel" % gname][t]) < RTOL, (gname, t)
                assert rel(e, g["S_%s_e" % gname][t]) < RTOL, (gname, t)
        finally:
            sph.G = 0.0
    # the in-place [particle][dim][time] driver of the reference (pyfluid.py:623-630)
    n, nt = 20, 3
    P = np.zeros((n, 3, nt)); V = np.zeros((n, 3, nt)); E = np.zeros((n, nt))
    P[:, :, 0], V[:, :, 0], E[:, 0] = g["S_G0_pos"][0], g["S_G0_vel"][0], g["S_G0_e"][0]
    sph.static_smoothlength_sim(np.array([0.0, 0.1, 0.2]), P, V, E)
    assert rel(P[:, :, 2], g["S_G0_pos"][2]) < RTOL and rel(E[:, 2], g["S_G0_e"][2]) < RTOL
    e_one = sph.energy_evolve(3, g["S_G0_pos"][0], g["S_G0_vel"][0], g["S_G0_e"][0],
                              sph.pressure_arr(g["S_G0_e"][0], sph.density_arr(g["S_G0_pos"][0], np.full(20, 2.0))),
                              sph.density_arr(g["S_G0_pos"][0], np.full(20, 2.0)), 2.0, 0.1)
    assert e_one == pytest.approx(g["S_G0_e"][1][3], rel=RTOL)


def test_pinned_host_tensors_overlapped_path(sph, golden):
    """pinned CPU tensors in -> pinned CPU tensors out (transfers overlapped with the kernels): same bits as the
    device-tensor path"""
    import torch
    n = 20000
    pos, vel, e, L = uniform_cloud(n, 55, vscale=0.6)
    h = sph.newton_smoothlength_arr(pos, np.full(n, 0.8))
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (pos, vel, e, h)]
    ref = sph.var_smoothlength_leapfrog(*dev, 1e-3)
    host = [t.cpu().pin_memory() for t in dev]
    out = sph.var_smoothlength_leapfrog(*host, 1e-3)
    for a, b in zip(ref, out):
        assert (not b.is_cuda) and b.is_pinned() and torch.equal(a.cpu(), b)
    # caller-provided pinned result buffers (the time-loop form: two sets used in turn)
    mine = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in host]
    got = sph.var_smoothlength_leapfrog(*host, 1e-3, out=mine)
    for a, b, m in zip(ref, got, mine):
        assert b.data_ptr() == m.data_ptr() and torch.equal(a.cpu(), b)


def test_sim_stride_and_staged_history(sph, golden, capsys):
    """stride keeps every k-th snapshot of the same run; numpy histories arrive through the pinned staging writer"""
    import torch
    g = golden
    t = np.arange(7) * 2e-3
    full = sph.var_smoothlength_sim(t, g["C_pos"], g["C_vel"], g["C_e"], np.full(64, 0.8))
    thin = sph.var_smoothlength_sim(t, g["C_pos"], g["C_vel"], g["C_e"], np.full(64, 0.8), stride=3)
    capsys.readouterr()
    for a, b in zip(full, thin):
        assert isinstance(b, np.ndarray) and b.shape[0] == 3 and np.array_equal(a[::3], b)
    assert rel(full[3][:3], g["C_sim_h"]) < RTOL            # the first three snapshots are the golden sim
    dev = sph.var_smoothlength_sim(t, *(torch.from_numpy(g[k]).cuda() for k in ("C_pos", "C_vel", "C_e")),
                                   torch.full((64,), 0.8, dtype=torch.float64, device="cuda"))
    capsys.readouterr()
    for a, b in zip(full, dev):
        assert b.is_cuda and np.array_equal(a, b.cpu().numpy())


def test_example_drivers_run(tmp_path):
    """the corrected 2-D example drivers run headless and write their histories"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for script, extra in (("ex2d_var_h.py", ["--steps", "6", "--stride", "2"]), ("ex2d_static_h.py", ["--steps", "5"])):
        out = tmp_path / (script + ".npz")
        r = subprocess.run([sys.executable, os.path.join(root, "examples", script), "--out", str(out)] + extra,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        z = np.load(out)
        assert np.isfinite(z["pos"]).all() and np.isfinite(z["u"]).all()
