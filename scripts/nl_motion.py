"""Per-step motion of the Sod tube as the kept lists see it (tuning aid): max displacement / h and max h growth since
the last build, the margins chosen, builds.  python scripts/nl_motion.py [nx=192] [steps=40] [dt_scale=1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-sph_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import engine as E
import workloads

a1 = sys.argv[1] if len(sys.argv) > 1 else "192"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dts = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
w = workloads.plummer(int(a1.split(":")[1])) if a1.startswith("plummer:") else workloads.sod_tube(int(a1))   # or plummer:N
eng = E.Engine()
k = dict(PARTICLE_MASS=w["m"], COUPLING_CONST=1.3, SMOOTHLENGTH_VARIATION_TOLERANCE=1e-3, NEWTON_ITERATION_LIMIT=20,
         BISECTION_ITERATION_LIMIT=100, ALPHA_SPH=1.0, BETA_SPH=2.0, EPSILON=0.01, BISECTION_TOLERANCE=1e-10,
         INITIAL_BISECTION_SMOOTHLENGTH_A=1e3, INITIAL_BISECTION_SMOOTHLENGTH_B=1e-3, G=0.0, ADIABATIC_INDEX=5 / 3)
c = E.make_consts(k)
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
pos, vel, u, hg = to(w["pos"]), to(w["vel"]), to(w["u"]), to(w["h_guess"])
st0 = E.Status(1, eng.device)
posh = eng.pack_posh(pos, hg)
dg = eng.build(posh, *eng.grid_for(posh))
h = eng.scatter(eng.newton(dg, c, dg.posh[:, 3].contiguous(), st0.ptr(0)), dg.perm, 1)
fs = E.FastStepper(eng)
fs.load(pos, vel, u, h)
dt = w["dt"] * dts
p0 = fs.state()
for s in range(steps):
    b0 = fs.stats["builds"]
    fs.advance(dt, c)
    nl = fs.nl
    if nl is None:
        print(s, "safe path", fs.stats)
        continue
    P = fs.P
    d = (P[:, :3] - nl.dg.posh[:, :3]).norm(dim=1) / nl.dg.posh[:, 3]
    g = P[:, 3] / nl.dg.posh[:, 3] - 1
    mm = nl.margins
    ms = f"mean cover {mm[:, 0].double().mean().item():.4f} skin {mm[:, 1].double().mean().item():.4f}" if mm is not None else "uniform"
    print(f"step {s:3d} built {fs.stats['builds'] - b0} age {fs.age}/{fs.margins.life} max cover_h {nl.cover_h:.4f} skin {nl.skin:.4f} {ms} "
          f"used {nl.used_d:.3f} {nl.used_g:.3f} | max moved {d.max().item():.5f} grown {g.max().item():.5f} "
          f"rows/warp {nl.warp_rows.double().mean().item():.1f} redos {fs.stats['redos']} safety {fs.margins.pp_safety}")
    if a1.startswith("plummer:") and s < 3:
        k = int(torch.argmax(g))
        print("      largest growth at slot", k, "r", float(P[k, :3].norm()), "h_build", float(nl.dg.posh[k, 3]), "h_now", float(P[k, 3]))
E.raise_for_status(fs.last_status.read(), printer=lambda *_: None)
