"""Time the list-drain kernels one by one on the Sod tube (tuning aid): python scripts/nl_kernels.py [nx=256] [reps=3]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-sph_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import engine as E
import workloads

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = workloads.sod_tube(nx)
n = len(w["pos"])
eng = E.Engine()
k = dict(PARTICLE_MASS=w["m"], COUPLING_CONST=1.3, SMOOTHLENGTH_VARIATION_TOLERANCE=1e-3, NEWTON_ITERATION_LIMIT=20,
         BISECTION_ITERATION_LIMIT=100, ALPHA_SPH=1.0, BETA_SPH=2.0, EPSILON=0.01, BISECTION_TOLERANCE=1e-10,
         INITIAL_BISECTION_SMOOTHLENGTH_A=1e3, INITIAL_BISECTION_SMOOTHLENGTH_B=1e-3, G=0.0, ADIABATIC_INDEX=5 / 3)
c = E.make_consts(k)
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
pos, vel, u, hg = to(w["pos"]), to(w["vel"]), to(w["u"]), to(w["h_guess"])
st = E.Status(4, eng.device)
posh = eng.pack_posh(pos, hg)
dg = eng.build(posh, *eng.grid_for(posh))
h = eng.scatter(eng.newton(dg, c, dg.posh[:, 3].contiguous(), st.ptr(0)), dg.perm, 1)
posh = eng.pack_posh(pos, h)
dg = eng.build(posh, *eng.grid_for(posh))
vel_s, u_s = eng.gather(vel, dg.perm, 3), eng.gather(u, dg.perm, 1)


def t(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 3), r


out = {"n": n}
out["build_ms"], nl = t(lambda: eng.nl_build(dg, 1.03, 0.03))
out["rows_mean"] = float(nl.warp_rows.double().mean().item())
out["omega_ms"], r = t(lambda: eng.nl_density_omega(nl, c, dg.posh))
om = r[2]
_, _, rB, rC = eng.eos(dg.posh, vel_s, u_s, om, c, st.ptr(1))
out["force_ms"], r = t(lambda: eng.nl_force(nl, c, dg.posh, rB, rC, st.ptr(1)))
rB2, rD = eng.nl_eos(dg.posh, vel_s, u_s, om, c, st.ptr(1))
out["force_dense_ms"], r2 = t(lambda: eng.nl_force(nl, c, dg.posh, rB2, None, st.ptr(1), recD=rD))
out["dense_same_bits"] = bool(torch.equal(r[0], r2[0]) and torch.equal(r[1], r2[1]))
out["newton_fused_ms"], r = t(lambda: eng.nl_newton(nl, c, dg.posh, st.ptr(2), want_omega=True))
out["newton_ms"], r = t(lambda: eng.nl_newton(nl, c, dg.posh, st.ptr(3)))
out["acc_checksum"] = float(torch.nan_to_num(r[0]).sum().item())
print(json.dumps(out))
