"""Invariance of the force / density sums under re-chunking of targets into warps (prepend far-away dummies) and
under target masks, plus run-to-run determinism.  Bit-identical sums are what makes k-rank == 1-rank."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "python-sph_b200")]
import numpy as np, torch
import engine as E, workloads
eng = E.Engine()
w = workloads.sod_tube(64)
k = dict(PARTICLE_MASS=w["m"], COUPLING_CONST=1.3, SMOOTHLENGTH_VARIATION_TOLERANCE=1e-3, NEWTON_ITERATION_LIMIT=20,
         BISECTION_ITERATION_LIMIT=100, ALPHA_SPH=1.0, BETA_SPH=2.0, EPSILON=0.01, BISECTION_TOLERANCE=1e-10,
         INITIAL_BISECTION_SMOOTHLENGTH_A=1e3, INITIAL_BISECTION_SMOOTHLENGTH_B=1e-3, G=0.0, ADIABATIC_INDEX=5 / 3)
c = E.make_consts(k)
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
pos, vel, u = to(w["pos"]), to(w["vel"]), to(w["u"])
h0 = to(w["h_guess"] * 1.0843)
lo, hi, hmax, hmean = eng.bounds(eng.pack_posh(pos, h0))
grid, delta = E.choose_grid(lo, hi, hmean, pos.shape[0])
def force(shift=0, mask_every=0):
    p, hh, vv, uu = pos, h0, vel, u
    if shift:
        far = torch.full((shift, 3), -10.0, dtype=torch.float64, device="cuda") + torch.arange(shift, device="cuda", dtype=torch.float64)[:, None] * 1e-3
        p = torch.cat([far, p]); hh = torch.cat([h0[:shift] * 0 + float(h0[0]), hh]); vv = torch.cat([vel[:shift] * 0, vv]); uu = torch.cat([u[:shift], uu])
    dg = eng.build(eng.pack_posh(p, hh), grid, delta)
    st = E.Status(1, eng.device)
    tgt = None
    if mask_every:
        tgt = torch.ones(p.shape[0], dtype=torch.uint8, device="cuda"); tgt[::mask_every] = 0
        tgt = tgt[dg.perm.long()].contiguous()
    rho, _, om, _ = eng.density_omega(dg, c, target=tgt, want_rho=True, want_omega=True)
    _, _, om_all, _ = eng.density_omega(dg, c, want_rho=False, want_omega=True)
    _, _, rB, rC = eng.eos(dg.posh, eng.gather(vv, dg.perm, 3), eng.gather(uu, dg.perm, 1), om_all, c, st.ptr(0))
    acc, dudt = eng.force(dg, c, rB, rC, eng.hmax(dg.posh), st.ptr(0), target=tgt)
    return eng.scatter(acc, dg.perm, 3)[shift:], eng.scatter(dudt, dg.perm, 1)[shift:], eng.scatter(rho, dg.perm, 1)[shift:], eng.scatter(om, dg.perm, 1)[shift:]
a0, d0, r0, o0 = force()
for s, me in ((0, 0), (5, 0), (17, 0), (0, 3)):
    a1, d1, r1, o1 = force(s, me)
    keep = torch.ones(a0.shape[0], dtype=torch.bool, device="cuda")
    if me: keep[::me] = False
    print(f"shift {s} mask {me}: mismatches acc {int((a0[keep] != a1[keep]).any(dim=1).sum())} dudt {int((d0[keep] != d1[keep]).sum())} "
          f"rho {int((r0[keep] != r1[keep]).sum())} omega {int((o0[keep] != o1[keep]).sum())}  max|dacc|/|acc| {float(((a0[keep]-a1[keep]).abs().max()/a0.abs().max()))}")

a1, d1, r1, o1 = force(5, 0)
bad = (a0 != a1).any(dim=1)
xb = pos[bad, 0].cpu().numpy(); xa = pos[:, 0].cpu().numpy()
hist_b, edges = np.histogram(xb, bins=10, range=(-0.5, 0.5)); hist_a, _ = np.histogram(xa, bins=10, range=(-0.5, 0.5))
print("fraction mismatching per x-bin:", np.round(hist_b / np.maximum(hist_a, 1), 3))
yb = pos[bad, 1].cpu().numpy(); ya = pos[:, 1].cpu().numpy()
hb, _ = np.histogram(yb, bins=8, range=(0, 0.25)); ha, _ = np.histogram(ya, bins=8, range=(0, 0.25))
print("per y-bin:", np.round(hb / np.maximum(ha, 1), 3))
rel = ((a0 - a1).abs().max(dim=1).values / a0.abs().max(dim=1).values)[bad]
if rel.numel():
    print("relative diff of mismatching: median", float(rel.median()), "max", float(rel.max()))
