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
pos, hg = to(w["pos"]), to(w["h_guess"])
def run(pos, hg, shift=0):
    if shift:
        far = torch.full((shift, 3), -10.0, dtype=torch.float64, device="cuda") + torch.arange(shift, device="cuda", dtype=torch.float64)[:, None] * 1e-3
        pos = torch.cat([far, pos]); hg = torch.cat([torch.full((shift,), float(hg[0]), dtype=torch.float64, device="cuda"), hg])
    posh = eng.pack_posh(pos, hg)
    lo, hi, hmax, hmean = eng.bounds(eng.pack_posh(pos[shift:], hg[shift:]))
    grid, delta = E.choose_grid(lo, hi, hmean, pos.shape[0] - shift)
    dg = eng.build(posh, grid, delta)
    st = E.Status(1, eng.device)
    h_s, om_s = eng.newton(dg, c, dg.posh[:, 3].contiguous(), st.ptr(0), want_omega=True)
    eng.set_h(dg, h_s)
    _, _, om2, _ = eng.density_omega(dg, c, want_rho=False, want_omega=True)
    h = eng.scatter(h_s, dg.perm, 1)[shift:]; om = eng.scatter(om_s, dg.perm, 1)[shift:]; om2 = eng.scatter(om2, dg.perm, 1)[shift:]
    return h, om, om2
h0, om0, om0b = run(pos, hg)
print("fused vs separate pass: mismatches", int((om0 != om0b).sum()), "max rel", float(((om0 - om0b).abs() / om0b.abs()).max()))
for s in (5, 17):
    h1, om1, om1b = run(pos, hg, s)
    print(f"shift {s}: h mismatches", int((h0 != h1).sum()), "omega fused mismatches", int((om0 != om1).sum()), "omega separate mismatches", int((om0b != om1b).sum()))

# force invariance under re-chunking and target masks
vel = to(w["vel"]); u = to(w["u"])
def force(shift=0, mask_every=0):
    p, hh, vv, uu = pos, h0, vel, u
    if shift:
        far = torch.full((shift, 3), -10.0, dtype=torch.float64, device="cuda") + torch.arange(shift, device="cuda", dtype=torch.float64)[:, None] * 1e-3
        p = torch.cat([far, p]); hh = torch.cat([h0[:shift] * 0 + float(h0[0]), hh]); vv = torch.cat([vel[:shift] * 0, vv]); uu = torch.cat([u[:shift], uu])
    posh = eng.pack_posh(p, hh)
    lo, hi, hmax, hmean = eng.bounds(eng.pack_posh(pos, h0))
    grid, delta = E.choose_grid(lo, hi, hmean, pos.shape[0])
    dg = eng.build(posh, grid, delta)
    st = E.Status(1, eng.device)
    tgt = None
    if mask_every:
        tgt = torch.ones(p.shape[0], dtype=torch.uint8, device="cuda"); tgt[::mask_every] = 0
        tgt = tgt[dg.perm.long()].contiguous()
    _, _, om, _ = eng.density_omega(dg, c, want_rho=False, want_omega=True)
    _, _, rB, rC = eng.eos(dg.posh, eng.gather(vv, dg.perm, 3), eng.gather(uu, dg.perm, 1), om, c, st.ptr(0))
    acc, dudt = eng.force(dg, c, rB, rC, eng.hmax(dg.posh), st.ptr(0), target=tgt)
    return eng.scatter(acc, dg.perm, 3)[shift:], eng.scatter(dudt, dg.perm, 1)[shift:]
a0, d0 = force()
for s, me in ((5, 0), (17, 0), (0, 3)):
    a1, d1 = force(s, me)
    keep = torch.ones(a0.shape[0], dtype=torch.bool, device="cuda")
    if me: keep[::me] = False
    print(f"shift {s} mask {me}: acc mismatches", int((a0[keep] != a1[keep]).any(dim=1).sum()), "dudt mismatches", int((d0[keep] != d1[keep]).sum()))
