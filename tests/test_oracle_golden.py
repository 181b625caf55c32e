"""Pin the CPU oracle (oracle/sph_oracle.c) against outputs of the reference itself.

The fixtures in tests/golden/pyfluid_golden.npz were produced by running the unmodified
reference (pyfluid.py) -- see tests/golden/make_golden.py.  Tolerance: the oracle restates the
reference operation by operation, so agreement is expected at rounding level; we assert
1e-13 relative (and report bit-exactness separately in test_oracle_bit_exact_report).
"""
import json
import os

import numpy as np
import pytest

RTOL = 1e-13


def close(a, b, rtol=RTOL, atol=0.0):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = np.maximum(np.abs(b), atol)
    err = np.abs(a - b)
    ok = (err <= rtol * np.maximum(scale, 1e-300)) | ((a == b)) | (np.isnan(a) & np.isnan(b))
    assert ok.all(), f"max rel err {np.nanmax(err / np.maximum(scale, 1e-300)):.3e}"


def vclose(a, b, rtol=RTOL):
    """vector fields: error relative to the largest component magnitude of each row"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = np.abs(b).max(axis=-1, keepdims=True)
    assert (np.abs(a - b) <= rtol * scale).all(), f"max rel err {np.max(np.abs(a - b) / scale):.3e}"


def test_kernel_kats(golden, oracle_mod):
    o = oracle_mod.Oracle()
    d, h = golden["kat_dist"], golden["kat_h"]
    close([o.M4(x, y) for x, y in zip(d, h)], golden["kat_M4"], 1e-15)
    close([o.M4_d1(x, y) for x, y in zip(d, h)], golden["kat_M4_d1"], 1e-15)
    close([o.M4_smoothlength_derivative(x, y) for x, y in zip(d, h)], golden["kat_M4_dh"], 1e-15)


def test_survey_known_answers(oracle_mod):
    """SURVEY.md 8(c) survey-measured known answers of the reference."""
    o = oracle_mod.Oracle()
    assert o.M4(0.5, 1) == 0.22878523069459955
    assert o.M4(1.5, 1) == 0.009947183943243459
    assert o.M4(0, 1) == 0.3183098861837907
    assert o.M4(2, 1) == 0
    assert o.M4_d1(0.5, 1) == -0.2984155182973038
    assert o.M4_d1(1, 1) == -0.238732414637843
    assert o.M4_d1(1.5, 1) == -0.05968310365946075
    assert o.M4_smoothlength_derivative(0, 1) == -0.954929658551372
    assert o.M4_smoothlength_derivative(0.5, 1) == -0.5371479329351467
    assert o.M4_smoothlength_derivative(1, 1) == 0
    assert o.M4_smoothlength_derivative(1.5, 1) == 0.05968310365946075


def test_case_A_newton_and_bisection(golden, oracle_mod):
    o = oracle_mod.Oracle()
    pos = golden["A_pos"]
    h, it, code, _ = o.newton_smoothlength_while([0] * 6, pos, golden["A_guess"])
    close(h, golden["A_h"])
    assert (it == golden["A_iters"]).all()
    assert ((code != 0) == golden["A_warned"]).all()
    assert float(golden["A_h_from_0p8"]) == 0.8159916383864595
    hb, codeb, _ = o.bisection_h([0], pos)
    assert codeb[0] == 1 and hb[0] == float(golden["A_bisection_h0"])  # dyadic midpoints: bit-exact
    close(o.density([0], pos, [0.8159916383864595]), golden["A_density_j0"])


def test_case_B_stage_and_step(golden, oracle_mod):
    g = golden
    o = oracle_mod.Oracle(G=0.0)
    pos, vel, e = g["B_pos"], g["B_vel"], g["B_e"]
    h = o.newton_smoothlength_arr(pos, np.full(24, 0.8))
    close(h, g["B_h"])
    assert list(g["B_h"][:3]) == [1.0567999602726978, 1.3546743185444694, 1.0091701482800675]
    rho = o.var_density_arr(g["B_h"])
    close(rho, g["B_den"], 1e-15)
    om = o.omega_arr(pos, rho, g["B_h"])
    close(om, g["B_omega"])
    P = o.pressure_arr(e, rho)
    close(P, g["B_P"], 1e-15)
    close(o.energy_rate_arr(pos, vel, P, rho, g["B_h"], om), g["B_dudt"], 1e-12)
    vclose(o.acceleration_arr(pos, rho, vel, P, g["B_h"], om, np.zeros(24)), g["B_acc_G0"], 1e-12)
    close(o.density_arr(pos, g["B_h"]), g["B_density_sum"])
    p1, v1, e1, h1 = o.var_smoothlength_leapfrog(pos, vel, e, g["B_h"], 1e-3)
    close(p1, g["B_lf_G0_pos"]); vclose(v1, g["B_lf_G0_vel"]); close(e1, g["B_lf_G0_e"]); close(h1, g["B_lf_G0_h"])


def test_case_B_default_gravity(golden, oracle_mod):
    g = golden
    o = oracle_mod.Oracle()  # default G = 39.4227
    pos, vel, e, h = g["B_pos"], g["B_vel"], g["B_e"], g["B_h"]
    xi = o.xi_arr(g["B_den"], h, pos)
    close(xi, g["B_xi"])
    vclose(o.acceleration_arr(pos, g["B_den"], vel, g["B_P"], h, g["B_omega"], xi), g["B_acc_G"], 1e-12)
    p1, v1, e1, h1 = o.var_smoothlength_leapfrog(pos, vel, e, h, 1e-3)
    close(p1, g["B_lf_G_pos"]); vclose(v1, g["B_lf_G_vel"]); close(e1, g["B_lf_G_e"]); close(h1, g["B_lf_G_h"])
    assert g["B_lf_G_h"][0] == 1.0562858987682495 and g["B_lf_G_e"][0] == 1.408523737660402


def test_case_C_gravity_pieces(golden, oracle_mod):
    g = golden
    o = oracle_mod.Oracle()   # default G
    xi = o.xi_arr(g["C_den"], g["C_h"], g["C_pos"])
    close(xi, g["C_xi"])
    vclose(o.acceleration_arr(g["C_pos"], g["C_den"], g["C_vel"], g["C_P"], g["C_h"], g["C_omega"], xi), g["C_acc_G"], 1e-12)


def test_case_C_sim_and_neighbours(golden, oracle_mod):
    g = golden
    o = oracle_mod.Oracle(G=0.0)
    pos, vel, e = g["C_pos"], g["C_vel"], g["C_e"]
    P, V, E, H = o.var_smoothlength_sim(np.array([0.0, 2e-3, 4e-3]), pos, vel, e, np.full(64, 0.8))
    close(H, g["C_sim_h"]); close(P, g["C_sim_pos"]); vclose(V, g["C_sim_vel"]); close(E, g["C_sim_e"])
    rho = o.var_density_arr(g["C_h"])
    om = o.omega_arr(pos, rho, g["C_h"])
    close(om, g["C_omega"])
    close(o.energy_rate_arr(pos, vel, g["C_P"], rho, g["C_h"], om), g["C_dudt"], 1e-12)
    vclose(o.acceleration_arr(pos, rho, vel, g["C_P"], g["C_h"], om, np.zeros(64)), g["C_acc_G0"], 1e-12)
    for j in range(64):  # neighbour sets: exact
        assert (np.flatnonzero(g["C_nb"][j]) == o.neighbours(pos, j, g["C_h"][j])).all()


def test_case_D_planar_example_settings(golden, oracle_mod):
    g = golden
    o = oracle_mod.Oracle(G=0.0)
    h = o.newton_smoothlength_arr(g["D_pos"], np.full(50, 1.0))
    close(h, g["D_h"])
    p1, v1, e1, h1 = o.var_smoothlength_leapfrog(g["D_pos"], g["D_vel"], g["D_e"], g["D_h"], 0.1)
    close(p1, g["D_lf_pos"]); close(e1, g["D_lf_e"]); close(h1, g["D_lf_h"])
    np.testing.assert_allclose(v1, g["D_lf_vel"], rtol=1e-12, atol=1e-70)


def test_case_E_overguess_falls_back_to_bisection(golden, oracle_mod):
    g = golden
    o = oracle_mod.Oracle()
    h, it, code, _ = o.newton_smoothlength_while(np.arange(64), g["C_pos"], float(g["E_guess"]))
    assert ((code != 0) == g["E_warned"]).all()
    assert (it == g["E_iters"]).all()
    close(h, g["E_h"])
    assert g["E_warned"].any() and not g["E_warned"].all()


def test_case_F_no_root_gives_nan(golden, oracle_mod):
    o = oracle_mod.Oracle()
    h, it, code = o.newton_smoothlength_arr(golden["F_pos"], np.full(4, 0.2), details=True)
    assert np.isnan(h).all() and np.isnan(golden["F_h"]).all() and (code == 2).all()
    meta = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pyfluid_golden.json")))
    _, _, oob = o.bisection_h(np.arange(4), golden["F_pos"])
    assert int(oob.sum()) == meta["F_stdout_counts"]["ERROR: Root out of bisection bounds."]


def test_case_G_coincident_particles(golden, oracle_mod):
    g = golden
    o = oracle_mod.Oracle(G=0.0)
    pos, vel, e, h = g["G_pos"], g["G_vel"], g["G_e"], g["G_h"]
    close(o.newton_smoothlength_arr(pos, np.full(64, 0.8)), h)
    rho = o.var_density_arr(h); P = o.pressure_arr(e, rho); om = o.omega_arr(pos, rho, h)
    close(om, g["G_omega"])
    close(o.energy_rate_arr(pos, vel, P, rho, h, om), g["G_dudt"], 1e-12)
    vclose(o.acceleration_arr(pos, rho, vel, P, h, om, np.zeros(64)), g["G_acc_G0"], 1e-12)


def test_case_S_static_path_repaired(golden, oracle_mod):
    """BASELINE config 1 through the documented repair, with and without gravity: three steps"""
    g = golden
    for tag, G in (("G0", 0.0), ("G", None)):
        o = oracle_mod.Oracle(G=0.0) if G == 0.0 else oracle_mod.Oracle()
        p, v, e = g["S_%s_pos" % tag][0], g["S_%s_vel" % tag][0], g["S_%s_e" % tag][0]
        for t in range(1, 4):
            p, v, e = o.static_smoothlength_leapfrog(p, v, e, 0.1)
            close(p, g["S_%s_pos" % tag][t], 1e-12, atol=1e-12)
            np.testing.assert_allclose(v, g["S_%s_vel" % tag][t], rtol=1e-11, atol=1e-12)
            close(e, g["S_%s_e" % tag][t], 1e-12)


def test_negative_energy_raises(golden, oracle_mod):
    g = golden
    o = oracle_mod.Oracle(G=0.0)
    with pytest.raises(RuntimeError, match="Obtained negative energy."):
        o.var_smoothlength_leapfrog(g["B_pos"], g["B_vel"] * 5.0, np.full(24, 1e-3), g["B_h"], 0.5)


def test_oracle_bit_exact_report(golden, oracle_mod, capsys):
    """Not an assertion of the 1e-10 contract: records how many fields the operation-by-operation
    restatement reproduces bit for bit."""
    g = golden
    o = oracle_mod.Oracle(G=0.0)
    h = o.newton_smoothlength_arr(g["C_pos"], np.full(64, 0.8))
    rho = o.var_density_arr(g["C_h"]); om = o.omega_arr(g["C_pos"], rho, g["C_h"])
    acc = o.acceleration_arr(g["C_pos"], rho, g["C_vel"], g["C_P"], g["C_h"], om, np.zeros(64))
    exact = dict(h=(h == g["C_sim_h"][0]).mean(), omega=(om == g["C_omega"]).mean(), acc=(acc == g["C_acc_G0"]).mean())
    with capsys.disabled():
        print("\n[oracle bit-exact fractions vs reference]", exact)
    assert exact["h"] > 0.9


def test_mass_array_reduces_to_the_golden_vectors_when_uniform(oracle_mod, golden):
    """per-particle masses in the oracle: with every m_i = PARTICLE_MASS it reproduces the
    reference's outputs like the scalar does (the reference itself cannot take an array), and the array is really used"""
    g = golden
    o = oracle_mod.Oracle(G=0.0)
    pos, vel, e, h = g["B_pos"], g["B_vel"], g["B_e"], g["B_h"]
    n = len(h)
    try:
        o.set_masses(np.ones(n))
        assert np.max(np.abs(o.newton_smoothlength_arr(pos, np.full(n, 0.8)) - h) / h) < 1e-14
        rho = o.var_density_arr(h)
        om = o.omega_arr(pos, rho, h)
        assert np.max(np.abs(om - g["B_omega"]) / np.abs(g["B_omega"])) < 1e-13
        o.set_masses(np.full(n, 2.0))
        rho2 = o.var_density_arr(h)
        assert np.allclose(rho2, 2.0 * rho, rtol=1e-15)
    finally:
        o.set_masses(None)
    d1 = o.density_arr(pos, h)
    o.set_masses(np.full(n, 2.0))
    try:
        assert np.allclose(o.density_arr(pos, h), 2.0 * d1, rtol=1e-15)
    finally:
        o.set_masses(None)
