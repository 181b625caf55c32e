#!/usr/bin/env python
"""Generate golden fixtures by RUNNING THE REFERENCE (TRU-astrophysics/python-sph, pyfluid.py).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes tests/golden/pyfluid_golden.npz (+ pyfluid_golden.json with provenance).  The reference
ships no golden vectors of its own (no asserts, unseeded RNG), so these
outputs of the reference itself are what pins the oracle (oracle/sph_oracle.c) and, through it
and directly, the CUDA path.  Every array below is produced by calling a function of the
unmodified reference module; only module constants are set the way its run scripts are meant
to (pyfluid.py:5 "This number must be modified in the run script").
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
from fractions import Fraction

import numpy as np

REF = "/root/reference/pyfluid.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_pyfluid", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def fma(a, b, c):
    return float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def fma_selftest(n=20000):
    """numpy's d.dot(d) (OpenBLAS ddot, n=3) vs fma(z,z,fma(y,y,x*x)) -- SURVEY 8(c)."""
    rng = np.random.default_rng(1)
    ok_self = ok_cross = 0
    for _ in range(n):
        d = rng.standard_normal(3) * rng.random() * 10
        ok_self += float(d.dot(d)) == fma(d[2], d[2], fma(d[1], d[1], d[0] * d[0]))
        a, b = rng.standard_normal(3), rng.standard_normal(3)
        ok_cross += float(np.dot(a, b)) == fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0]))
    return ok_self, ok_cross, n


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


def stage_arrays(sph, pos, vel, e, h):
    den = sph.var_density_arr(h)
    om = sph.omega_arr(pos, den, h)
    P = sph.pressure_arr(e, den)
    dudt = sph.energy_rate_arr(pos, vel, P, den, h, om)
    return den, om, P, dudt


def main():
    sph = load_ref()
    G_DEFAULT = sph.G
    out = {}
    meta = {"numpy": np.__version__, "reference": REF}
    s, c, n = fma_selftest()
    meta["fma_selftest"] = {"dot_self_matches": s, "dot_cross_matches": c, "n": n}
    assert s == n and c == n, "numpy dot is not fma(z,z',fma(y,y',x*x')) on this host: oracle needs another order"

    # ---- KAT: scalar kernel functions (pyfluid.py:56,64,85) --------------------------------
    hs = np.array([1.0, 0.5, 0.8159916383864595, 2.0, 3.7, -0.7, 1e3, 1e-3])
    qs = np.array([0.0, 0.25, 0.5, 0.999999, 1.0, 1.000001, 1.5, 1.9999999, 2.0, 2.0000001, 3.0])
    kd, kh = [], []
    for h in hs:
        for q in qs:
            kd.append(q * abs(h)); kh.append(h)
    kd, kh = np.array(kd), np.array(kh)
    with np.errstate(all="ignore"):
        out["kat_dist"], out["kat_h"] = kd, kh
        out["kat_M4"] = np.array([float(sph.M4(np.float64(d), np.float64(h))) for d, h in zip(kd, kh)])
        out["kat_M4_d1"] = np.array([float(sph.M4_d1(np.float64(d), np.float64(h))) for d, h in zip(kd, kh)])
        out["kat_M4_dh"] = np.array([float(sph.M4_smoothlength_derivative(np.float64(d), np.float64(h))) for d, h in zip(kd, kh)])

    # ---- case A: pyfluid_newton_h_testing.py:7-16 settings, seeded ---------------------------
    rng = np.random.default_rng(0)
    posA = rng.random((30, 3)) * 2
    posA[0] = 1.0
    out["A_pos"] = posA
    out["A_h_from_0p8"] = np.float64(sph.newton_smoothlength_while(0, posA, 0.8))
    (hA2, txt) = quiet(sph.newton_smoothlength_while, 0, posA, 2.0)
    out["A_h_from_2p0"] = np.float64(hA2)
    meta["A_from_2p0_stdout"] = txt
    out["A_bisection_h0"] = np.float64(quiet(sph.bisection_h, 0, posA)[0])
    out["A_density_j0"] = np.float64(sph.density(0, posA, 0.8159916383864595))
    # iteration counts, by counting calls of the per-iteration function
    counts = []
    orig_it = sph.newton_smoothlength_iteration

    def counting(j, p, h):
        counts[-1] += 1
        return orig_it(j, p, h)

    sph.newton_smoothlength_iteration = counting
    its, hs_out, warned = [], [], []
    for g in (0.2, 0.5, 0.8, 1.0, 1.2, 2.0):
        counts.append(0)
        (hh, txt) = quiet(sph.newton_smoothlength_while, 0, posA, g)
        its.append(counts[-1]); hs_out.append(hh); warned.append("WARNING" in txt)
    sph.newton_smoothlength_iteration = orig_it
    out["A_guess"] = np.array([0.2, 0.5, 0.8, 1.0, 1.2, 2.0])
    out["A_iters"] = np.array(its); out["A_h"] = np.array(hs_out); out["A_warned"] = np.array(warned)

    # ---- case B: the survey's N=24 set -------------------------------------------
    rng = np.random.default_rng(0)
    N = 24
    L = (N / 3.75) ** (1 / 3)
    pos = rng.random((N, 3)) * L
    vel = rng.random((N, 3)) - 0.5
    e = 1 + rng.random(N)
    h = sph.newton_smoothlength_arr(pos, np.full(N, 0.8))
    sph.G = 0.0
    den, om, P, dudt = stage_arrays(sph, pos, vel, e, h)
    acc0 = sph.acceleration_arr(pos, den, vel, P, h, om, np.zeros(N))
    p1, v1, e1, h1 = sph.var_smoothlength_leapfrog(pos, vel, e, h, 1e-3)
    sph.G = G_DEFAULT
    xi = sph.xi_arr(den, h, pos)
    accG = sph.acceleration_arr(pos, den, vel, P, h, om, xi)
    pG, vG, eG, hG = sph.var_smoothlength_leapfrog(pos, vel, e, h, 1e-3)
    for k, v in dict(pos=pos, vel=vel, e=e, h=h, den=den, omega=om, P=P, dudt=dudt, acc_G0=acc0, xi=xi, acc_G=accG,
                     lf_G0_pos=p1, lf_G0_vel=v1, lf_G0_e=e1, lf_G0_h=h1,
                     lf_G_pos=pG, lf_G_vel=vG, lf_G_e=eG, lf_G_h=hG).items():
        out["B_" + k] = v
    out["B_density_sum"] = sph.density_arr(pos, h)

    # ---- case C: N=64 3-D uniform random at n=3.75, convergent flows, 2 steps (G=0) ----------
    rng = np.random.default_rng(7)
    N = 64
    L = (N / 3.75) ** (1 / 3)
    pos = rng.random((N, 3)) * L
    vel = (rng.random((N, 3)) - 0.5) * 2.0
    e = 1 + rng.random(N)
    sph.G = 0.0
    (hist, txt) = quiet(sph.var_smoothlength_sim, np.array([0.0, 2e-3, 4e-3]), pos, vel, e, np.full(N, 0.8))
    hC = hist[3][0]
    den, om, P, dudt = stage_arrays(sph, pos, vel, e, hC)
    acc0 = sph.acceleration_arr(pos, den, vel, P, hC, om, np.zeros(N))
    nb = np.zeros((N, N), dtype=bool)
    for j in range(N):
        for i in range(N):
            nb[j, i] = sph.distance(pos[j], pos[i]) / hC[j] < 2
    sph.G = G_DEFAULT
    for k, v in dict(pos=pos, vel=vel, e=e, h=hC, den=den, omega=om, P=P, dudt=dudt, acc_G0=acc0, nb=nb,
                     sim_pos=hist[0], sim_vel=hist[1], sim_e=hist[2], sim_h=hist[3]).items():
        out["C_" + k] = v
    meta["C_sim_stdout"] = txt

    # ---- case P: per-pair functions Pi (pyfluid.py:261) and fluid_accel_viscosity_comp (:459) on case C ----
    pairs = np.array([[0, 1], [3, 17], [5, 5], [10, 42], [63, 0], [7, 8], [20, 21], [33, 2]])
    sph.G = 0.0
    out["P_pairs"] = pairs
    out["P_Pi"] = np.array([float(sph.Pi(j, i, pos, vel, P, den, hC)) for j, i in pairs])
    out["P_acc"] = np.array([np.zeros(3) + sph.fluid_accel_viscosity_comp(j, i, pos, vel, den, P, hC, om) for j, i in pairs])
    sph.G = G_DEFAULT
    #      gravity pieces on the same pairs with the default G (pyfluid.py:401-437)
    xiC = sph.xi_arr(den, hC, pos)
    out["C_xi"] = xiC
    out["P_grav"] = np.array([np.zeros(3) + sph.smoothed_gravity_acceleration_comp(j, i, pos, hC) for j, i in pairs])
    out["P_corr"] = np.array([np.zeros(3) + sph.smoothed_gravity_correction_comp(j, i, pos, hC, om, xiC) for j, i in pairs])
    out["C_acc_G"] = sph.acceleration_arr(pos, den, vel, P, hC, om, xiC)
    kd2 = np.array([0.0, 0.3, 0.999, 1.0, 1.5, 1.999, 2.0, 2.5, 7.0])
    with np.errstate(all="ignore"):
        out["kat_grav_dist"] = kd2
        out["kat_grav_phi"] = np.array([float(sph.grav_potential(np.float64(d), np.float64(1.3))) for d in kd2 * 1.3])
        out["kat_grav_force"] = np.array([float(sph.grav_force(np.float64(d), np.float64(1.3))) for d in kd2 * 1.3])
        out["kat_grav_dphidh"] = np.array([float(sph.grav_potential_smoothlength_derivative(np.float64(d), np.float64(1.3))) for d in kd2 * 1.3])

    # ---- case D: pyfluidex2d_var_h.py:10-25 settings (planar z=0, u=1e-59, h guess 1, dt 0.1),
    #      seeded, with the (N,3) layout var_smoothlength_sim actually takes (pyfluid.py:544) -----
    rng = np.random.default_rng(3)
    N = 50
    pos = np.zeros((N, 3)); pos[:, :2] = rng.random((N, 2)) * 20
    vel = np.zeros((N, 3)); e = np.full(N, 1e-59)
    sph.G = 0.0
    (hD, txt) = quiet(sph.newton_smoothlength_arr, pos, np.full(N, 1.0))
    pD, vD, eD, hD1 = quiet(sph.var_smoothlength_leapfrog, pos, vel, e, hD, 0.1)[0]
    sph.G = G_DEFAULT
    for k, v in dict(pos=pos, vel=vel, e=e, h=hD, lf_pos=pD, lf_vel=vD, lf_e=eD, lf_h=hD1).items():
        out["D_" + k] = v
    meta["D_newton_stdout"] = txt

    # ---- case E: over-guessed Newton on case C positions -> WARNING + bisection for some ------
    pos = out["C_pos"]; N = len(pos)
    hE = np.zeros(N); wE = np.zeros(N, dtype=bool); itE = np.zeros(N, dtype=np.int64)
    sph.newton_smoothlength_iteration = counting
    for j in range(N):
        counts.append(0)
        (hh, txt) = quiet(sph.newton_smoothlength_while, j, pos, 1.15)
        hE[j] = hh; wE[j] = "WARNING" in txt; itE[j] = counts[-1]
    sph.newton_smoothlength_iteration = orig_it
    out["E_guess"] = np.float64(1.15); out["E_h"] = hE; out["E_warned"] = wE; out["E_iters"] = itE

    # ---- case F: var_h_testing.py:4 four-particle fixture: no root -> None -> TypeError -------
    posF = np.array([[0.0, 0, 0], [1, 1, 1], [-1, -1, -1], [1, 0, 0]])
    #      (numpy >= 2 stores None into a float64 array as nan, so no exception is raised here)
    try:
        (hF, txt) = quiet(sph.newton_smoothlength_arr, posF, np.full(4, 0.2))
        meta["F_exception"] = "none"
        out["F_h"] = hF
        meta["F_stdout_counts"] = {k: txt.count(k) for k in ("WARNING: Failure to converge in newton_new_h.",
                                                             "ERROR: Root out of bisection bounds.",
                                                             "ERROR: Failure to converge in Bisection_new_h.")}
    except Exception as ex:  # noqa: BLE001
        meta["F_exception"] = type(ex).__name__
    out["F_pos"] = posF

    # ---- case G: coincident distinct particles (Q9) + lattice block -----------------------------
    a = (1 / 3.75) ** (1 / 3)
    g = np.arange(4) * a
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).copy()
    pos[5] = pos[4]                         # two particles on the same point
    rng = np.random.default_rng(11)
    vel = (rng.random((64, 3)) - 0.5); e = 1 + rng.random(64)
    sph.G = 0.0
    hG_ = quiet(sph.newton_smoothlength_arr, pos, np.full(64, 0.8))[0]
    den, om, P, dudt = stage_arrays(sph, pos, vel, e, hG_)
    accg = sph.acceleration_arr(pos, den, vel, P, hG_, om, np.zeros(64))
    sph.G = G_DEFAULT
    for k, v in dict(pos=pos, vel=vel, e=e, h=hG_, omega=om, dudt=dudt, acc_G0=accg).items():
        out["G_" + k] = v

    # ---- case S: BASELINE config 1, pyfluidex2d_static_h.py:10-31 settings (N=20 planar, u=10, h=2, dt=0.1, two fast
    #      intruders), seeded, through the MINIMAL REPAIR of the static path (the shipped
    #      static_smoothlength_leapfrog calls acceleration_arr / energy_evolve_arr with the wrong arity, pyfluid.py:605-621).
    #      Every arithmetic step below is a call of an unmodified reference function.
    def static_leapfrog_repaired(pos0, vel0, energy0, dt):
        N = pos0.shape[0]
        hs = sph.DEFAULT_SMOOTHLENGTH * np.ones(N)
        one, zero = np.ones(N), np.zeros(N)
        den0 = sph.density_arr(pos0, hs)
        press0 = sph.pressure_arr(energy0, den0)
        acc0 = sph.acceleration_arr(pos0, den0, vel0, press0, hs, one, zero)
        pos1 = pos0 + vel0 * dt + 0.5 * acc0 * dt**2
        energy1 = sph.energy_evolve_arr(pos0, vel0, energy0, press0, den0, hs, None, dt)
        den1 = sph.density_arr(pos1, hs)
        press1 = sph.pressure_arr(energy1, den1)
        acc1 = sph.acceleration_arr(pos1, den1, vel0, press1, hs, one, zero)
        vel1 = vel0 + 0.5 * (acc0 + acc1) * dt
        return pos1, vel1, energy1

    rng = np.random.default_rng(0)
    N = 20
    posS = np.zeros((N, 3)); posS[:, :2] = rng.random((N, 2)) * 10
    velS = np.zeros((N, 3))
    posS[0] = [10, 0, 0]; velS[0] = [0, 10, 0]
    posS[1] = [0, 10, 0]; velS[1] = [-10, 0, 0]
    eS = np.full(N, 10.0)
    for gname, gval in (("G", G_DEFAULT), ("G0", 0.0)):
        sph.G = gval
        p, v, e_ = posS, velS, eS
        hist = [(p, v, e_)]
        for _ in range(3):
            p, v, e_ = static_leapfrog_repaired(p, v, e_, 0.1)
            hist.append((p, v, e_))
        out["S_%s_pos" % gname] = np.array([hh[0] for hh in hist])
        out["S_%s_vel" % gname] = np.array([hh[1] for hh in hist])
        out["S_%s_e" % gname] = np.array([hh[2] for hh in hist])
    sph.G = G_DEFAULT

    # ---- error convention: negative energy at the predictor (pyfluid.py:514-515) -------------
    sph.G = 0.0
    try:
        quiet(sph.var_smoothlength_leapfrog, out["B_pos"], out["B_vel"] * 5.0, np.full(24, 1e-3), out["B_h"], 0.5)
        meta["neg_energy_exception"] = "none"
    except RuntimeError as ex:
        meta["neg_energy_exception"] = str(ex)
    sph.G = G_DEFAULT

    np.savez_compressed(os.path.join(HERE, "pyfluid_golden.npz"), **out)
    with open(os.path.join(HERE, "pyfluid_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(out), "arrays;", json.dumps(meta["fma_selftest"]))


if __name__ == "__main__":
    sys.exit(main())
