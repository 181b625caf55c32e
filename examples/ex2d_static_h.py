"""Corrected driver for the reference's 2-D fixed-h example (pyfluidex2d_static_h.py:10-31, BASELINE config 1).

The shipped script calls `sph.static_h_sim`, which does not exist, and the function it means
(`static_smoothlength_sim`, pyfluid.py:623-630) cannot run as shipped; this driver uses the documented repair
(`pyfluid.static_smoothlength_leapfrog` in this package) with the script's settings: 20 particles
at random in [0,10]^2 of the z = 0 plane, particle 0 at (10,0,0) moving +y at 10, particle 1 at (0,10,0) moving -x at
10, u = 10, h = DEFAULT_SMOOTHLENGTH = 2, dt = 0.1, 100 steps.  Arrays are [particle][dim][time], as in the script.

    python examples/ex2d_static_h.py [--steps 100] [--seed 0] [--out ex2d_static_h.npz]
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
    ap.add_argument("--n", type=int, default=20)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--dt", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()

    rng = np.random.default_rng(a.seed)
    t = np.arange(a.steps + 1) * a.dt
    pos = np.zeros((a.n, 3, len(t)))
    vel = np.zeros((a.n, 3, len(t)))
    eng = np.ones((a.n, len(t))) * 10.0
    pos[:, :2, 0] = rng.random((a.n, 2)) * 10.0
    pos[0, :, 0] = [10, 0, 0]
    vel[0, :, 0] = [0, 10, 0]
    pos[1, :, 0] = [0, 10, 0]
    vel[1, :, 0] = [-10, 0, 0]

    sph.G = 0.0
    sph.static_smoothlength_sim(t, pos, vel, eng)
    print(f"{len(t)} snapshots of {a.n} particles; particle 0 ends at {pos[0, :, -1]}, energies {eng[:, -1].min():.4f} .. "
          f"{eng[:, -1].max():.4f}")
    if a.out:
        np.savez(a.out, t=t, pos=pos, vel=vel, u=eng)


if __name__ == "__main__":
    main()
