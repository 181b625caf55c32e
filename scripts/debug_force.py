"""debug aid: lists of contributing candidates of one target under two chunkings (needs -DSPHB_DEBUG_FORCE)"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "python-sph_b200")]
import numpy as np, torch
import engine as E, workloads
eng = E.Engine()
lib = eng.lib
lib.sphb_debug_force_set.argtypes = [C.c_longlong]
lib.sphb_debug_force_get.argtypes = [C.c_void_p, C.c_void_p]
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
def force(shift, dbg_caller=None):
    p, hh, vv, uu = pos, h0, vel, u
    if shift:
        far = torch.full((shift, 3), -10.0, dtype=torch.float64, device="cuda") + torch.arange(shift, device="cuda", dtype=torch.float64)[:, None] * 1e-3
        p = torch.cat([far, p]); hh = torch.cat([h0[:shift] * 0 + float(h0[0]), hh]); vv = torch.cat([vel[:shift] * 0, vv]); uu = torch.cat([u[:shift], uu])
    dg = eng.build(eng.pack_posh(p, hh), grid, delta)
    st = E.Status(1, eng.device)
    _, _, om_all, _ = eng.density_omega(dg, c, want_rho=False, want_omega=True)
    _, _, rB, rC = eng.eos(dg.posh, eng.gather(vv, dg.perm, 3), eng.gather(uu, dg.perm, 1), om_all, c, st.ptr(0))
    perm = dg.perm.cpu().numpy().astype(np.int64)
    if dbg_caller is not None:
        slot = int(np.nonzero(perm == dbg_caller + shift)[0][0])
        lib.sphb_debug_force_set(slot)
    acc, dudt = eng.force(dg, c, rB, rC, eng.hmax(dg.posh), st.ptr(0))
    out = eng.scatter(acc, dg.perm, 3)[shift:]
    lst = None
    if dbg_caller is not None:
        li = np.zeros(4096, np.uint32); lv = np.zeros(4096)
        n = lib.sphb_debug_force_get(li.ctypes.data, lv.ctypes.data)
        lst = (perm[li[:n]] - shift, lv[:n])
    return out, lst
a0, _ = force(0); a1, _ = force(5)
bad = torch.nonzero((a0 != a1).any(dim=1)).flatten().cpu().numpy()
print("mismatching targets:", len(bad))
for t in bad[:3]:
    _, (l0, v0) = force(0, int(t)); _, (l1, v1) = force(5, int(t))
    print("target", t, "pos", w["pos"][t], "n0", len(l0), "n1", len(l1), "same set", set(l0) == set(l1), "same order", list(l0) == list(l1))
    if set(l0) != set(l1):
        d01 = sorted(set(l0) - set(l1)); d10 = sorted(set(l1) - set(l0))
        print("  only in run0:", d01[:10], [float(v0[list(l0).index(i)]) for i in d01[:5]])
        print("  only in run1:", d10[:10], [float(v1[list(l1).index(i)]) for i in d10[:5]])
    elif list(l0) != list(l1):
        idx = [k for k in range(len(l0)) if l0[k] != l1[k]][:6]
        print("  first order differences at", idx, [(int(l0[k]), int(l1[k])) for k in idx])
    else:
        dv = np.abs(v0 - v1).max()
        print("  same lists; max |value diff|", dv)
