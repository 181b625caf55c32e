"""Sod tube timing of the kept-list path against the scanning path (tuning aid, not a test).
usage: python scripts/nl_eval.py [nx=384] [steps=10] [cover_h=1.03] [skin=0.03]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-sph_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import engine as E
import workloads

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cover_h = float(sys.argv[3]) if len(sys.argv) > 3 else E.NL_COVER_H
skin = float(sys.argv[4]) if len(sys.argv) > 4 else E.NL_SKIN
do_safe = os.environ.get("NL_EVAL_SAFE", "1") == "1"

w = workloads.sod_tube(nx)
n = len(w["pos"])
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
del dg, posh
dt = w["dt"]
out = {"n": n, "steps": steps, "cover_h": cover_h, "skin": skin}


def timed(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


fs = E.FastStepper(eng, cover_h=cover_h, skin=skin)
fs.load(pos, vel, u, h)
out["rows_used"] = fs.nl.rows_used()
out["rows_per_warp_mean"] = float(fs.nl.warp_rows.double().mean().item())
out["rows_per_warp_max"] = int(fs.nl.warp_rows.max().item())
out["build_flags"] = fs.nl.read_flags()
for _ in range(2):
    fs.advance(dt, c)
out["fast_ms_per_step"] = timed(lambda: fs.advance(dt, c), steps)
out["fast_stats"] = dict(fs.stats)
eng.prof_start()
fs.advance(dt, c)
out["fast_kernels"] = {kk: [v[0], round(v[1], 3)] for kk, v in sorted(eng.prof_stop().items(), key=lambda kv: -kv[1][1])}
# cost of one rebuild
eng.prof_start()
fs.stale = True
fs.advance(dt, c)
out["fast_rebuild_step_kernels"] = {kk: [v[0], round(v[1], 3)] for kk, v in sorted(eng.prof_stop().items(), key=lambda kv: -kv[1][1])}
E.raise_for_status(fs.last_status.read(), printer=lambda *_: None)
fast_state = fs.state()
nsteps_fast = fs.stats["steps"]

if do_safe:
    sp = E.Stepper(eng)
    s = (pos, vel, u, h)
    for _ in range(2):
        s = sp.step(*s, dt, c, defer_status=True)

    def one():
        global s
        s = sp.step(*s, dt, c, defer_status=True)
    out["safe_ms_per_step"] = timed(one, min(steps, 5))
    # same number of steps on both paths -> compare the states
    s = (pos, vel, u, h)
    for _ in range(nsteps_fast):
        s = sp.step(*s, dt, c, defer_status=True)
    errs = {}
    for nm, a, b in zip(("pos", "vel", "u", "h"), fast_state, s):
        d = (a - b).abs().max().item()
        errs[nm] = d / max(b.abs().max().item(), 1e-300)
    out["fast_vs_safe_rel_after_steps"] = errs
print(json.dumps(out))
