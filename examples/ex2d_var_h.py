"""Corrected driver for the reference's 2-D variable-h example (pyfluidex2d_var_h.py:10-54).

The shipped script cannot run against its own library: it passes [particle][dim][time] arrays and a 2-D h to
`var_smoothlength_sim`, which takes the t = 0 state as (N,3), (N,3), (N), (N) and RETURNS the histories.  This
driver keeps the script's settings -- 50 particles at random in [0,20]^2 of the z = 0 plane, at rest, u = 1e-59,
h guess 1, t = 0 .. 30 in steps of 0.1 -- seeds the generator, and works without a display.

    python examples/ex2d_var_h.py [--steps 300] [--stride 10] [--seed 0] [--out ex2d_var_h.npz] [--plot frames.png]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-sph_b200"))
import pyfluid as sph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--dt", type=float, default=0.1)
    ap.add_argument("--stride", type=int, default=1, help="keep every stride-th snapshot")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None, help="write the histories to this .npz")
    ap.add_argument("--plot", default=None, help="write first/last frame scatter plots to this image (needs matplotlib)")
    a = ap.parse_args()

    rng = np.random.default_rng(a.seed)
    t = np.arange(a.steps + 1) * a.dt
    pos0 = np.zeros((a.n, 3))
    pos0[:, :2] = rng.random((a.n, 2)) * 20.0
    vel0 = np.zeros((a.n, 3))
    u0 = np.full(a.n, 1e-59)
    h_guess = np.ones(a.n)

    sph.G = 0.0   # the example's point is the hydro step; the reference default G would add the all-pairs gravity sum
    P, V, E, H = sph.var_smoothlength_sim(t, pos0, vel0, u0, h_guess, stride=a.stride)
    print(f"{P.shape[0]} snapshots of {a.n} particles; h range at the end {H[-1].min():.4f} .. {H[-1].max():.4f}; "
          f"largest displacement {np.abs(P[-1] - P[0]).max():.3e}")
    if a.out:
        np.savez(a.out, t=t[::a.stride], pos=P, vel=V, u=E, h=H)
    if a.plot:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        fig, ax = plt.subplots(1, 2, figsize=(10, 5))
        for axis, k, title in ((ax[0], 0, "t = 0"), (ax[1], -1, f"t = {t[::a.stride][-1]:.1f}")):
            axis.scatter(P[k, :, 0], P[k, :, 1])
            axis.set_xlim(-25, 35)
            axis.set_ylim(-25, 35)
            axis.set_title(title)
        fig.savefig(a.plot, dpi=120)


if __name__ == "__main__":
    main()
