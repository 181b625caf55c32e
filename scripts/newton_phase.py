"""Per-step Newton kernel times against what the data asked of them (iterations per warp, h drift), Sod tube."""
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
_newton = eng.newton
rec = []


def spy(dg, consts, h_guess, status_ptr, target=None, details=False, want_omega=False, lists=None):
    r = _newton(dg, consts, h_guess, status_ptr, target=target, details=True, want_omega=want_omega, lists=lists)
    it = r[1]
    if os.environ.get("TRACE"):
        n32 = it.numel() // 32 * 32
        cyc = r[2][:n32].view(-1, 32)[:, 0].double() * 16
        info = it[:n32].view(-1, 32)
        scans, rounds = (info[:, 0] >> 8) & 255, (info[:, 0] >> 16) & 255
        cntmax = ((info >> 24) & 127).max(dim=1).values
        print(f"   warps {cyc.numel()}  mean cycles {cyc.mean():.0f}  p50 {cyc.median():.0f}  p99 {cyc.quantile(0.99) if cyc.numel() < 1e7 else -1:.0f} max {cyc.max():.0f}")
        for sc in range(1, 4):
            for ro in range(1, 6):
                m = (scans == sc) & (rounds == ro)
                if m.any():
                    print(f"      scans {sc} rounds {ro}: warps {int(m.sum()):8d}  mean cycles {cyc[m].mean():9.0f}  share of cycles {float(cyc[m].sum() / cyc.sum()):.3f} mean max-list {cntmax[m].float().mean():.1f}")
        wi = int(cyc.argmax())
        sl = slice(wi * 32, wi * 32 + 32)
        print(f"      slowest warp {wi}: scans {int(scans[wi])} rounds {int(rounds[wi])} cycles {cyc[wi]:.0f}")
        print("        iters", (it[sl] & 255).tolist())
        print("        h_guess", [round(x, 5) for x in h_guess[sl].tolist()])
        print("        h_out  ", [round(x, 5) for x in r[0][sl].tolist()])
        print("        x      ", [round(x, 4) for x in dg.posh[sl, 0].tolist()])
        print("        y      ", [round(x, 4) for x in dg.posh[sl, 1].tolist()])
        print("        z      ", [round(x, 4) for x in dg.posh[sl, 2].tolist()])
        print("        cell   ", dg.cell_id[sl].tolist(), "grid", dg.grid.nx, dg.grid.ny, dg.grid.nz, "cell edge", dg.grid.cell)
        nw = cyc.numel()
        parts = cyc[: nw // 16 * 16].view(16, -1).mean(dim=1)
        print("      mean cycles by sixteenth of the slot range:", [int(x) for x in parts.tolist()])
        it = it & 255
    n32 = it.numel() // 32 * 32
    wmax = it[:n32].view(-1, 32).max(dim=1).values
    ratio = r[0] / h_guess
    rec.append((float((it >= 2).float().mean()), float((wmax >= 2).float().mean()), float((wmax >= 3).float().mean()),
                float(ratio.max()), float(ratio.min()), float((ratio > 1.004).float().mean())))
    return (r[0], r[3]) if want_omega else r[0]


for step in range(int(sys.argv[2]) if len(sys.argv) > 2 else 36):
    eng.prof_start()
    s2 = stp.step(s[0], s[1], s[2], s[3], w["dt"], c, defer_status=True)
    torch.cuda.synchronize()
    tn = [round(a.elapsed_time(b), 2) for nm, a, b in eng.prof if nm == "newton_h"]
    tf = [round(a.elapsed_time(b), 2) for nm, a, b in eng.prof if nm == "force"]
    eng.prof_stop()
    del rec[:]
    eng.newton = spy
    stp.step(s[0], s[1], s[2], s[3], w["dt"], c, defer_status=True)
    eng.newton = _newton
    s = s2
    m = rec[0]
    print(f"step {step:2d} newton {tn} force {tf} | mid: lanes>=2it {m[0]:.3f} warps>=2it {m[1]:.3f} warps>=3it {m[2]:.4f} "
          f"h/h0 max {m[3]:.4f} min {m[4]:.4f} frac>1.004 {m[5]:.4f}")
