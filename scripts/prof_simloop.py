"""Per-kernel times of the plain step vs the Omega-carrying time loop (Sod tube)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "python-sph_b200")]
import numpy as np, torch
import engine as E, workloads
eng = E.Engine()
w = workloads.sod_tube(int(sys.argv[1]) if len(sys.argv) > 1 else 384)
k = dict(PARTICLE_MASS=w["m"], COUPLING_CONST=1.3, SMOOTHLENGTH_VARIATION_TOLERANCE=1e-3, NEWTON_ITERATION_LIMIT=20,
         BISECTION_ITERATION_LIMIT=100, ALPHA_SPH=1.0, BETA_SPH=2.0, EPSILON=0.01, BISECTION_TOLERANCE=1e-10,
         INITIAL_BISECTION_SMOOTHLENGTH_A=1e3, INITIAL_BISECTION_SMOOTHLENGTH_B=1e-3, G=0.0, ADIABATIC_INDEX=5 / 3)
c = E.make_consts(k)
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
pos, vel, u, hg = to(w["pos"]), to(w["vel"]), to(w["u"]), to(w["h_guess"])
st = E.Status(1, eng.device)
posh = eng.pack_posh(pos, hg); grid, delta = eng.grid_for(posh); dg = eng.build(posh, grid, delta)
h = eng.scatter(eng.newton(dg, c, dg.posh[:, 3].contiguous(), st.ptr(0)), dg.perm, 1)
del dg, posh
stp = E.Stepper(eng)
s = (pos, vel, u, h)


def timed(fn, s, n=3):
    for _ in range(2):
        s = fn(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        s = fn(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    eng.newton = newton_twice
    eng.prof_start()
    fn(s)
    torch.cuda.synchronize()
    print("   newton calls, each launched twice:", [round(a.elapsed_time(b), 2) for nm, a, b in eng.prof if nm == "newton_h"])
    eng.prof_stop()
    eng.newton = _newton
    eng.prof_start()
    s = fn(s)
    torch.cuda.synchronize()
    print("   newton calls:", [round(a.elapsed_time(b), 2) for nm, a, b in eng.prof if nm == "newton_h"])
    prof = eng.prof_stop()
    return s, ms, prof


_newton = eng.newton
_its = []


def newton_spy(dg, consts, h_guess, status_ptr, target=None, details=False, want_omega=False, lists=None):
    r = _newton(dg, consts, h_guess, status_ptr, target=target, details=True, want_omega=want_omega, lists=lists)
    _its.append(r[1])
    M = 1 << 21
    print("      ptrs MB (off KB):", {k: (v.data_ptr() >> 20, (v.data_ptr() % M) >> 10) for k, v in
                                     dict(posh=dg.posh, frec=dg.frec, cs=dg.cell_start, hm=dg.cell_hmax, g=h_guess, h=r[0],
                                          it=r[1], code=r[2], om=r[3] if want_omega else r[0], cid=dg.cell_id, perm=dg.perm).items()})
    return (r[0], r[3]) if want_omega else r[0]


def newton_twice(dg, consts, h_guess, status_ptr, **kw):
    _newton(dg, consts, h_guess, status_ptr, **kw)
    return _newton(dg, consts, h_guess, status_ptr, **kw)


plain = lambda s: stp.step(s[0], s[1], s[2], s[3], w["dt"], c, defer_status=True)
carry = lambda s: stp.step(s[0], s[1], s[2], s[3], w["dt"], c, defer_status=True, omega0=s[4] if len(s) > 4 else None,
                           want_omega1=True)
_build = eng.build
_mode = [None]
_nb = [0]
_hold = []


def build_spy(*a, **k):
    dg = _build(*a, **k)
    if _nb[0] % 3 == 0:
        del _hold[:]
        if _mode[0] == "omega":
            eng.density_omega(dg, c, want_rho=False, want_omega=True)
        elif _mode[0] == "alloc":
            _hold.append(torch.empty(dg.n, dtype=torch.float64, device="cuda"))
    _nb[0] += 1
    return dg


eng.build = build_spy


def carry_with(mode):
    def fn(s):
        _mode[0] = mode
        r = carry(s)
        _mode[0] = None
        return r
    return fn


want1 = lambda s: stp.step(s[0], s[1], s[2], s[3], w["dt"], c, defer_status=True, want_omega1=True)
for name, fn in (("plain", plain), ("carry", carry), ("plain", plain)):
    s, ms, prof = timed(fn, s)
    eng.newton = newton_spy
    del _its[:]
    fn(s)
    eng.newton = _newton
    print("   iterations per call:", [np.bincount(i.cpu().numpy())[:6].tolist() for i in _its])
    print(f"{name}: {ms:.2f} ms/step; kernels {sum(t for _, t in prof.values()):.2f} ms")
    for kname, (calls, t) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        print(f"    {kname:16s} x{calls}  {t:8.3f}")
    s = s[:4] if name != "carry" else s
