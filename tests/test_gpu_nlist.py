"""GPU tests of the kept-neighbour-list path (csrc/sphb_nlist.cu, engine.FastStepper), through the C-ABI.

The drain kernels run the same FP64 pair arithmetic in the same (cell, caller index) order as the scanning kernels, so
on one cell list they must return the SAME BITS; across steps (lists reused, slot order frozen) they must agree with
the scanning Stepper -- and with the CPU oracle -- within the 1e-10 bar of BASELINE.json (observed: ~1e-15).
"""
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
    if not torch.cuda.is_available():
        pytest.skip("-m gpu tests need a CUDA device")
    import build_native
    build_native.build()
    import pyfluid
    for k, v in dict(PARTICLE_MASS=1, COUPLING_CONST=1.3, G=0.0).items():
        setattr(pyfluid, k, v)
    return pyfluid


def cloud(n, seed, density=3.75, vscale=1.0):
    rng = np.random.default_rng(seed)
    L = (n / density) ** (1 / 3)
    return rng.random((n, 3)) * L, (rng.random((n, 3)) - 0.5) * vscale, 1 + rng.random(n), L


def two_density(seed=3, a=0.1, nl=14):
    """8:1 density jump (h differs 2x across the interface): sparse particles next to the dense block have several
    hundred candidates -- more than the 128-entry staging area: the builder's heavy-warp path"""
    gl = np.arange(nl) * a
    left = np.stack(np.meshgrid(gl, gl, gl, indexing="ij"), -1).reshape(-1, 3) + a * 0.5
    left[:, 0] -= nl * a
    gr = np.arange(nl // 2) * 2 * a
    right = np.stack(np.meshgrid(gr, gr, gr, indexing="ij"), -1).reshape(-1, 3) + a
    pos = np.concatenate([left, right])
    rng = np.random.default_rng(seed)
    pos += (rng.random(pos.shape) - 0.5) * 0.2 * a
    n = len(pos)
    vel = (rng.random((n, 3)) - 0.5) * 0.3
    e = 1 + rng.random(n)
    guess = np.where(pos[:, 0] < 0, 1.2 * a, 2.4 * a)
    return pos, vel, e, guess, a ** 3


def _to(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("case", ["cloud", "two_density"])
@pytest.mark.parametrize("cover_h,skin", [(1.0, 0.0), (1.03, 0.03), (1.25, 0.25)])
def test_drains_equal_scanning_kernels_bit_for_bit(sph, case, cover_h, skin):
    import torch
    import engine as E
    eng = sph._eng()
    if case == "cloud":
        pos, vel, e, L = cloud(20000, 5)
        guess, m = np.full(len(pos), 0.8), 1.0
    else:
        pos, vel, e, guess, m = two_density()
    sph.PARTICLE_MASS = m
    try:
        c = sph._consts()
        h = _to(sph.newton_smoothlength_arr(pos, guess))
        posh = eng.pack_posh(_to(pos), h)
        grid, delta = eng.grid_for(posh)
        dg = eng.build(posh, grid, delta)
        st = E.Status(2, eng.device)
        nl = eng.nl_build(dg, cover_h, skin)
        fl = nl.read_flags()
        assert fl == 0, fl
        rows = nl.warp_rows.cpu().numpy()
        if case == "two_density" and cover_h >= 1.2:
            assert rows.max() > 192, "the interface warps were meant to overflow the builder's staging area"
        # exact symmetric neighbour sets are subsets of the rows
        q = torch.arange(0, dg.n, 97, dtype=torch.int32, device="cuda")
        out, cnt = eng.neighbours(dg, q, 1024, symmetric=True)
        out, cnt, qq = out.cpu().numpy(), cnt.cpu().numpy(), q.cpu().numpy()
        off = nl.warp_off.cpu().numpy().astype(np.int64)
        for k, s in enumerate(qq):
            w, lane = s // 32, s % 32
            mine = nl.pool[off[w]:off[w] + rows[w], lane].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            mine = mine[mine != 0xFFFFFFFF]
            assert (np.diff(mine) > 0).all(), "rows must be in ascending slot order"
            assert set(out[k, :cnt[k]].tolist()) <= set(mine.tolist())
        # K2
        rho_a, dw_a, om_a, ng_a = eng.density_omega(dg, c, want_rho=True, want_dwdh=True, want_omega=True, want_ngb=True)
        rho_b, dw_b, om_b, ng_b = eng.nl_density_omega(nl, c, dg.posh, want_rho=True, want_dwdh=True, want_omega=True,
                                                       want_ngb=True)
        for a, b in ((rho_a, rho_b), (dw_a, dw_b), (om_a, om_b), (ng_a, ng_b)):
            assert torch.equal(a, b)
        nl2 = eng.nl_build(dg, cover_h, skin, consts=c)      # Omega formed by the build itself
        assert nl2.read_flags() == 0 and torch.equal(nl2.omega0, om_a)
        assert torch.equal(nl2.warp_rows, nl.warp_rows)
        del nl2
        # K4
        vel_s, u_s = eng.gather(_to(vel), dg.perm, 3), eng.gather(_to(e), dg.perm, 1)
        _, _, rB, rC = eng.eos(dg.posh, vel_s, u_s, om_a, c, st.ptr(0))
        acc_a, du_a = eng.force(dg, c, rB, rC, eng.hmax(dg.posh), st.ptr(0))
        acc_b, du_b = eng.nl_force(nl, c, dg.posh, rB, rC, st.ptr(1))
        assert torch.equal(acc_a, acc_b) and torch.equal(du_a, du_b)
        rB2, rD = eng.nl_eos(dg.posh, vel_s, u_s, om_a, c, st.ptr(1))          # dense {rho, cs} records
        acc_c, du_c = eng.nl_force(nl, c, dg.posh, rB2, None, st.ptr(1), recD=rD)
        assert torch.equal(rB, rB2) and torch.equal(acc_a, acc_c) and torch.equal(du_a, du_c)
        # K3 from a slightly wrong guess that stays inside the cover (Omega fused)
        if cover_h > 1.0:
            g = dg.posh[:, 3].contiguous() * 0.995
            h_a, it_a, _, om2_a = eng.newton(dg, c, g, st.ptr(0), details=True, want_omega=True)
            h_b, om2_b, posh_b, it_b = eng.nl_newton(nl, c, dg.posh, st.ptr(1), h_guess=g, want_omega=True, want_iters=True)
            assert nl.read_flags() & ~E.N.NL_AGING == 0
            assert torch.equal(h_a, h_b) and torch.equal(om2_a, om2_b) and torch.equal(it_a, it_b)
            assert torch.equal(posh_b[:, 3], h_b) and torch.equal(posh_b[:, :3], dg.posh[:, :3])
        assert (st.read()[:, 0] == 0).all()
    finally:
        sph.PARTICLE_MASS = 1


def test_check_flags_displacement_and_h_growth(sph):
    import engine as E
    eng = sph._eng()
    pos, vel, e, L = cloud(3000, 9)
    h = _to(sph.newton_smoothlength_arr(pos, np.full(len(pos), 0.8)))
    posh = eng.pack_posh(_to(pos), h)
    dg = eng.build(posh, *eng.grid_for(posh))
    nl = eng.nl_build(dg, 1.03, 0.04)
    eng.nl_check(nl, dg.posh)
    assert nl.read_flags() == 0
    p = dg.posh.clone()
    p[17, 0] += 0.025 * float(dg.posh[17, 3])          # more than half of the skin, less than all of it
    eng.nl_check(nl, p)
    assert nl.read_flags() == E.N.NL_AGING
    p[17, 0] += 0.03 * float(dg.posh[17, 3])
    eng.nl_check(nl, p)
    assert nl.read_flags() == E.N.NL_AGING | E.N.NL_BROKEN
    nl.flags.zero_()
    p = dg.posh.clone()
    p[5, 3] *= 1.031
    eng.nl_check(nl, p)
    assert nl.read_flags() == E.N.NL_AGING | E.N.NL_BROKEN


def _run_steps(stepper, state, dt, c, nsteps):
    for _ in range(nsteps):
        state = stepper.step(state[0], state[1], state[2], state[3], dt, c)
    return state


@pytest.mark.parametrize("case", ["cloud", "two_density"])
def test_fast_stepper_matches_scanning_stepper(sph, case):
    """six steps: the first on fresh lists, the others on kept ones (slot order frozen, positions and h rewritten)"""
    import engine as E
    eng = sph._eng()
    if case == "cloud":
        pos, vel, e, L = cloud(20000, 41, vscale=1.0)
        guess, m, dt = np.full(len(pos), 0.8), 1.0, 1e-3
    else:
        pos, vel, e, guess, m = two_density(seed=8)
        dt = 2e-4
    sph.PARTICLE_MASS = m
    try:
        c = sph._consts()
        h = sph.newton_smoothlength_arr(pos, guess)
        s0 = (_to(pos), _to(vel), _to(e), _to(h))
        fast, safe = E.FastStepper(eng), E.Stepper(eng)
        a = _run_steps(fast, s0, dt, c, 6)
        b = _run_steps(safe, s0, dt, c, 6)
        assert fast.stats["safe_steps"] == 0 and fast.stats["redos"] == 0, fast.stats
        assert fast.stats["builds"] < 6, fast.stats          # the lists really were reused
        names = ("pos", "vel", "u", "h")
        for nm, x, y in zip(names, a, b):
            x, y = x.cpu().numpy(), y.cpu().numpy()
            err = vrel(x, y) if x.ndim == 2 else rel(x, y)
            assert err < 1e-12, (nm, err)
        # resident-state API: same lists, same bits as the pure-function calls
        res = E.FastStepper(eng)
        res.load(*s0)
        for _ in range(6):
            res.advance(dt, c)
        for x, y in zip(res.state(), a):
            assert rel(x.cpu().numpy(), y.cpu().numpy()) < 1e-13
    finally:
        sph.PARTICLE_MASS = 1


def test_fast_stepper_vs_oracle_two_steps(sph, oracle_mod):
    n = 2000
    pos, vel, e, L = cloud(n, 21, vscale=2.0)
    o = oracle_mod.Oracle(G=0.0)
    h = o.newton_smoothlength_arr(pos, np.full(n, 0.8))
    s = (pos, vel, e, h)
    r = s
    for _ in range(2):
        s = sph.var_smoothlength_leapfrog(*s, 2e-3)
        r = o.var_smoothlength_leapfrog(*r, 2e-3)
    assert rel(s[0], r[0]) < RTOL and vrel(s[1], r[1]) < RTOL and rel(s[2], r[2]) < RTOL and rel(s[3], r[3]) < RTOL


def test_broken_lists_redo_and_fallback(sph, capsys):
    """(1) a step that moves particles by more than the skin is redone on lists with wider margins and still equals the
    scanning path; (2) a step with absurd h guesses (lists far beyond the pool, Newton solves that may need the
    bisection) ends up wherever it must -- larger pool, or the scanning Stepper -- with the same prints and results"""
    import engine as E
    eng = sph._eng()
    pos, vel, e, L = cloud(6000, 63, vscale=1.0)
    c = sph._consts()
    h = sph.newton_smoothlength_arr(pos, np.full(len(pos), 0.8))
    s0 = (_to(pos), _to(vel), _to(e), _to(h))
    fast, safe = E.FastStepper(eng, cover_h=1.004, skin=0.0005), E.Stepper(eng)
    a = fast.step(*s0, 2e-3, c)          # 0.5 * 2e-3 * |v| ~ 5e-4 >> skin * h
    b = safe.step(*s0, 2e-3, c)
    assert fast.stats["redos"] >= 1 and fast.stats.get("widened", 0) >= 1 and fast.stats["safe_steps"] == 0, fast.stats
    assert fast.skin >= 0.02
    for x, y in zip(a, b):
        assert rel(x.cpu().numpy(), y.cpu().numpy()) < 1e-12
    # bad guesses: 2.5x too large -> Newton fails for some particles -> WARNING + bisection on the scanning path
    fast2 = E.FastStepper(eng)
    s1 = (s0[0], s0[1], s0[2], s0[3] * 2.5)
    capsys.readouterr()
    a = fast2.step(*s1, 1e-4, c)
    out_fast = capsys.readouterr().out
    b = safe.step(*s1, 1e-4, c)
    out_safe = capsys.readouterr().out
    assert fast2.stats["redos"] >= 1, fast2.stats
    assert out_fast == out_safe
    for x, y in zip(a, b):
        assert rel(x.cpu().numpy(), y.cpu().numpy()) < 1e-12
