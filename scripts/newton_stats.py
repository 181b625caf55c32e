import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "python-sph_b200")]
import numpy as np, torch
import engine as E, workloads
eng = E.Engine()
w = workloads.sod_tube(int(sys.argv[1]) if len(sys.argv) > 1 else 128)
k = dict(PARTICLE_MASS=w["m"], COUPLING_CONST=1.3, SMOOTHLENGTH_VARIATION_TOLERANCE=1e-3, NEWTON_ITERATION_LIMIT=20,
         BISECTION_ITERATION_LIMIT=100, ALPHA_SPH=1.0, BETA_SPH=2.0, EPSILON=0.01, BISECTION_TOLERANCE=1e-10,
         INITIAL_BISECTION_SMOOTHLENGTH_A=1e3, INITIAL_BISECTION_SMOOTHLENGTH_B=1e-3, G=0.0, ADIABATIC_INDEX=5 / 3)
c = E.make_consts(k)
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
pos, vel, u, hg = to(w["pos"]), to(w["vel"]), to(w["u"]), to(w["h_guess"])
st = E.Status(1, eng.device)
posh = eng.pack_posh(pos, hg); grid, delta = eng.grid_for(posh); dg = eng.build(posh, grid, delta)
h, it, code = eng.newton(dg, c, dg.posh[:, 3].contiguous(), st.ptr(0), details=True)
print("cold solve iterations:", np.bincount(it.cpu().numpy())[:12], "codes", np.bincount(code.cpu().numpy()))
h = eng.scatter(h, dg.perm, 1)
stp = E.Stepper(eng)
s = (pos, vel, u, h)
for _ in range(3):
    s = stp.step(s[0], s[1], s[2], s[3], w["dt"], c)
# steady state: solve at advanced positions with the previous h as the guess
p2 = s[0] + s[1] * (0.5 * w["dt"])
posh = eng.pack_posh(p2, s[3]); dg = eng.build(posh, grid, delta)
h2, it, code = eng.newton(dg, c, dg.posh[:, 3].contiguous(), st.ptr(0), details=True)
itn = it.cpu().numpy()
print("steady-state iterations:", np.bincount(itn)[:8])
warps = itn[: len(itn) // 32 * 32].reshape(-1, 32).max(axis=1)
print("warps by max iterations:", np.bincount(warps)[:8])
