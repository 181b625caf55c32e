"""Synthetic particle distributions of BASELINE.json's configs; numpy, seeded.

Each generator returns dict(pos (N,3), vel (N,3), u (N,), h_guess (N,), m, dt, name).  The reference's
example scripts only fix the *settings* (number density, eta, guesses); sizes are scaled as SURVEY says.
"""
import numpy as np


def _noise_vel(n, amp, seed):
    if amp == 0.0:
        return np.zeros((n, 3))
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3)) - 0.5) * (2.0 * amp)


def planar_uniform(n=262144, seed=0, u=1.0, vel_noise=0.0):
    """C2: pyfluidex2d_var_h.py:10-25 settings scaled: z = 0 sheet, surface density 50/(20*20), h guess 1."""
    rng = np.random.default_rng(seed)
    L = float(np.sqrt(n / 0.125))
    pos = np.zeros((n, 3))
    pos[:, :2] = rng.random((n, 2)) * L
    return dict(pos=pos, vel=_noise_vel(n, vel_noise, seed + 1), u=np.full(n, u), h_guess=np.full(n, 1.0), m=1.0, dt=0.1,
                name=f"C2 planar uniform-random sheet, N={n}")


def lattice_cube(side=100, vel_noise=0.0, seed=0):
    """C3: pyfluid_newton_h_testing.py:7-14 settings scaled: number density 3.75 simple-cubic lattice, guess 0.8."""
    a = (1.0 / 3.75) ** (1.0 / 3.0)
    g = np.arange(side) * a
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).copy()
    n = len(pos)
    return dict(pos=pos, vel=_noise_vel(n, vel_noise, seed + 1), u=np.full(n, 1.0), h_guess=np.full(n, 0.8), m=1.0, dt=1e-3,
                name=f"C3 lattice cube {side}^3, N={n}")


def sod_tube(nx_left=384, vel_noise=0.01, seed=0):
    """C4: 3-D Sod tube.  Left x in [-0.5,0): nx x nx/2 x nx/2 lattice, spacing a, rho=1, P=1; right x in [0,0.5):
    nx/2 x nx/4 x nx/4, spacing 2a, rho=0.125, P=0.1; equal masses m = a^3; gamma = 5/3 -> u = 1.5 / 1.2;
    h guess 1.2 x spacing.  nx_left = 384 -> N = 15 925 248.  vel_noise: uniform random velocity components of
    that amplitude (in units of c_s ~ 1.3) so that half of the pairs approach and the viscosity branch runs."""
    a = 0.5 / nx_left
    nl = (nx_left, nx_left // 2, nx_left // 2)
    nr = (nx_left // 2, nx_left // 4, nx_left // 4)
    gl = [(np.arange(k) + 0.5) * a for k in nl]
    gr = [(np.arange(k) + 0.5) * 2 * a for k in nr]
    left = np.stack(np.meshgrid(*gl, indexing="ij"), -1).reshape(-1, 3)
    left[:, 0] -= 0.5
    right = np.stack(np.meshgrid(*gr, indexing="ij"), -1).reshape(-1, 3)
    pos = np.concatenate([left, right])
    n = len(pos)
    u = np.concatenate([np.full(len(left), 1.5), np.full(len(right), 1.2)])
    hg = np.concatenate([np.full(len(left), 1.2 * a), np.full(len(right), 2.4 * a)])
    cs = 1.29
    return dict(pos=pos, vel=_noise_vel(n, vel_noise * cs, seed + 1), u=u, h_guess=hg, m=a ** 3, dt=0.05 * a,
                name=f"C4 3-D Sod tube {nl[0]}x{nl[1]}x{nl[2]} + {nr[0]}x{nr[1]}x{nr[2]}, N={n}")


def uniform_random(n, seed=0, density=3.75, vscale=1.0):
    """3-D uniform random cloud at the reference test script's number density (pyfluid_newton_h_testing.py:7)."""
    rng = np.random.default_rng(seed)
    L = (n / density) ** (1 / 3)
    return dict(pos=rng.random((n, 3)) * L, vel=(rng.random((n, 3)) - 0.5) * vscale, u=1 + rng.random(n),
                h_guess=np.full(n, 0.8), m=1.0, dt=1e-3, name=f"uniform random cloud, N={n}")


def plummer(n=67_108_864, seed=0, u=1.0, guess=0.8, eta=1.3):
    """C5: Plummer sphere: scale a = 1, r = a / sqrt(U^(-2/3) - 1) with U ~ U(0, 0.999] (r < 38.7 a),
    isotropic directions, equal masses m = 1/N, v = 0, u = const, G = 0 for the metric.  h guess = `guess` x the h of the
    analytic local density, eta (m/rho)^(1/3) with rho = 3/(4 pi) (1 + r^2)^(-5/2) / 0.999: from 5e-3 in the core to > 2
    at the edge for 64 M particles -- nearly three decades, more once the sparse rim is solved."""
    rng = np.random.default_rng(seed)
    U = 0.999 * (1.0 - rng.random(n))                      # (0, 0.999]
    r = 1.0 / np.sqrt(U ** (-2.0 / 3.0) - 1.0)
    cz = 2.0 * rng.random(n) - 1.0
    ph = 2.0 * np.pi * rng.random(n)
    sz = np.sqrt(np.maximum(0.0, 1.0 - cz * cz))
    pos = np.empty((n, 3))
    pos[:, 0] = r * sz * np.cos(ph)
    pos[:, 1] = r * sz * np.sin(ph)
    pos[:, 2] = r * cz
    m = 1.0 / n
    rho = 3.0 / (4.0 * np.pi) * (1.0 + r * r) ** -2.5 / 0.999
    hg = guess * eta * (m / rho) ** (1.0 / 3.0)
    cs = np.sqrt(5.0 / 3.0 * (2.0 / 3.0) * u)
    dt = 0.05 * float(hg.min() / guess) / cs               # a twentieth of the sound crossing time of the smallest h
    return dict(pos=pos, vel=np.zeros((n, 3)), u=np.full(n, float(u)), h_guess=hg, m=m, dt=dt,
                name=f"C5 Plummer sphere, N={n}")


WORKLOADS = {"plummer": plummer, "sod": sod_tube, "lattice": lattice_cube, "planar": planar_uniform}
