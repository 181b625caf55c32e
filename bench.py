#!/usr/bin/env python
"""bench.py -- particle-steps/s of one var_smoothlength_leapfrog-equivalent step (pyfluid.py:498-532).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sod|lattice|planar] [--nx 384]
    python bench.py --impl reference ...      # the CPU arm (oracle port of the reference, all host threads)

Metric (BASELINE.json): particle-steps/sec (density + h-solve + force) at 16 M particles.  Workload at
N=1: configs[3], the 3-D Sod tube with 15 925 248 particles (fits one B200).  One step = 3 cell-list
builds, 1 Omega pass + 1 Omega sum inside a Newton kernel, 2 Newton h solves, 2 force passes, predictor,
corrector, fp64.
`value` is timed with CUDA events with the state resident in HBM; `e2e` goes through the public drop-in
API (pyfluid.var_smoothlength_leapfrog semantics on pinned host arrays: H2D + step + D2H per step).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "python-sph_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "particle-steps/sec"
FLOP_OWN = 23.0    # per accepted density / dW/dh interaction
FLOP_FORCE = 66.0  # per interaction of the symmetric set


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sod", choices=["sod", "lattice", "planar", "plummer"])
    ap.add_argument("--nx", type=int, default=384, help="sod: left-half lattice points along x (384 -> 15.9 M)")
    ap.add_argument("--side", type=int, default=100, help="lattice: cube side")
    ap.add_argument("--n", "--particles", dest="n", type=int, default=None, help="planar (default 262144) / plummer (default 67108864): particle count")
    ap.add_argument("--path", default="lists", choices=["lists", "scan"],
                    help="lists: neighbour lists kept across kernels and steps (default); scan: the scanning kernels only")
    ap.add_argument("--decomp", default=None, choices=["slab", "range"],
                    help="multi-GPU decomposition: slabs along x (default; sod, lattice, planar) or ranges of sorted slots of one "
                         "global cell list (default for plummer)")
    ap.add_argument("--warm-state-steps", type=int, default=0, help="untimed steps at setup so that the timed steps see a developed flow")
    ap.add_argument("--dt-scale", type=float, default=1.0, help="multiply the workload's dt (sod: 0.05 a; 6 is a CFL-sized step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def make_workload(a):
    import workloads
    if a.workload == "sod":
        return workloads.sod_tube(a.nx)
    if a.workload == "lattice":
        return workloads.lattice_cube(a.side, vel_noise=0.01)
    if a.workload == "plummer":
        return workloads.plummer(a.n or 67_108_864, seed=0)
    return workloads.planar_uniform(a.n or 262144, vel_noise=0.01)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        self.lines = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_start(self):
        """the timed region begins now (the sampler itself was started before the warm-up, so that nvidia-smi's
        start-up -- NVML initialisation takes driver locks -- does not fall into the timed steps)"""
        self.t0 = time.time()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.time()
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = getattr(self, "t0", 0.0)
        inside = [ln for ts, ln in self.lines if t0 <= ts <= t1 + 0.25]
        for ln in (inside or [ln for _, ln in self.lines[-1:]]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_sample(w, h, seconds, consts_over):
    """oracle port of the reference timed on the host cores for a bounded sample of particles j at the full N
    (each j costs O(N), like the reference's per-particle loops).  -> (particle-steps/s, threads, nj, elapsed)"""
    import oracle
    oracle.build()
    o = oracle.Oracle(**consts_over)
    o.set_num_threads(os.cpu_count() or 1)
    o.set_wide(True)
    n = len(w["pos"])
    rng = np.random.default_rng(123)
    cand = rng.choice(n, size=min(n, 4096), replace=False)
    t0 = time.perf_counter()
    o.step_work_sample(w["pos"], w["vel"], w["u"], h, cand[:1])
    t1 = time.perf_counter() - t0
    nj = int(max(1, min(len(cand) - 1, round(seconds / max(t1, 1e-6)))))
    t0 = time.perf_counter()
    o.step_work_sample(w["pos"], w["vel"], w["u"], h, cand[1:1 + nj])
    el = time.perf_counter() - t0
    o.set_wide(False)
    return nj / el, o.num_threads(), nj, el


def run_reference(a):
    """--impl reference: the reference's algorithm (O(N) sums per particle, pyfluid.py:498-532) as restated by
    oracle/sph_oracle.c (the reference itself is pure Python and cannot be compiled), on all host
    threads.  Each step evaluates the per-particle share of one leapfrog step for a bounded sample of particles
    against the full-N state."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    w = make_workload(a)
    n = len(w["pos"])
    o = oracle.Oracle(G=0.0, PARTICLE_MASS=w["m"])
    o.set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1: take the box's cores at every N
    o.set_wide(True)
    h = w["h_guess"] * {"sod": 1.3012 / 1.2, "plummer": 1.0 / 0.8}.get(a.workload, 1.0)   # about the solved h
    rng = np.random.default_rng(7)
    cand = rng.choice(n, size=min(n, 65536), replace=False)
    t0 = time.perf_counter()
    o.step_work_sample(w["pos"], w["vel"], w["u"], h, cand[:1])
    t1 = time.perf_counter() - t0
    budget = 150.0 / max(1, a.steps + a.warmup)
    nj = int(max(1, min(4096, budget / max(t1, 1e-6), (len(cand) - 1) // max(1, a.steps + a.warmup))))   # distinct particles per step
    k = 1
    for _ in range(a.warmup):
        o.step_work_sample(w["pos"], w["vel"], w["u"], h, cand[k:k + nj]); k += nj
    t0 = time.perf_counter()
    for _ in range(a.steps):
        o.step_work_sample(w["pos"], w["vel"], w["u"], h, cand[k:k + nj]); k += nj
    el = time.perf_counter() - t0
    val = nj * a.steps / el
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "particle-steps/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": el / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "n_particles": n, "sample_per_step": nj},
            "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": o.num_threads(), "kind": "port",
                             "sample": f"{nj} particles/step x {a.steps} steps, each against all {n} particles "
                                       "(per-particle share of one leapfrog step: 2x(omega, energy_rate, acceleration, newton))"},
            "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bit_checksum(tensors):
    """order-independent exact fingerprint of a particle state: the sum (mod 2^64) of the IEEE bit patterns of every
    value -- it does not depend on how particles are ordered or split over ranks, and any flipped bit changes it"""
    import torch
    tot = torch.zeros((), dtype=torch.int64, device=tensors[0].device)
    for t in tensors:
        tot = tot + t.contiguous().view(torch.int64).sum()
    return tot


def interactions(eng, E, consts, pos, h, n_total, dev):
    """mean interaction counts of a state: own support (q_j < 2, incl. self) and symmetric support"""
    import torch
    posh = eng.pack_posh(pos, h)
    grid, delta = eng.grid_for(posh)
    dg = eng.build(posh, grid, delta)
    _, _, _, ng = eng.density_omega(dg, consts, want_rho=False, want_ngb=True)
    n_own = float(ng.double().sum().item())
    q = torch.arange(0, n_total, max(1, n_total // 200000), dtype=torch.int32, device=dev)
    _, cs = eng.neighbours(dg, q, 1, symmetric=True)
    n_sym = float(cs.double().mean().item()) * n_total
    return n_own, n_sym


def run_ours(a):
    import torch
    import torch.distributed as dist
    import engine as E

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = E.Engine(dev)

    w = make_workload(a)
    n_total = len(w["pos"])
    k = dict(PARTICLE_MASS=w["m"], COUPLING_CONST=1.3, SMOOTHLENGTH_VARIATION_TOLERANCE=1e-3, NEWTON_ITERATION_LIMIT=20,
             BISECTION_ITERATION_LIMIT=100, ALPHA_SPH=1.0, BETA_SPH=2.0, EPSILON=0.01, BISECTION_TOLERANCE=1e-10,
             INITIAL_BISECTION_SMOOTHLENGTH_A=1e3, INITIAL_BISECTION_SMOOTHLENGTH_B=1e-3, G=0.0, ADIABATIC_INDEX=5 / 3)
    consts = E.make_consts(k)
    dt = w["dt"] * a.dt_scale
    use_lists = a.path == "lists"

    if world > 1:
        import domain
        decomp = a.decomp or ("range" if a.workload == "plummer" else "slab")
        runner = domain.make_runner(eng, consts, world, rank, lists=use_lists, decomp=decomp)
        state = runner.scatter_initial(w)
    else:
        runner = None

    # ---- setup (untimed): state to HBM, cold h solve, warm-state steps ------------------------------
    host = {key: torch.from_numpy(np.ascontiguousarray(w[key])).pin_memory() for key in ("pos", "vel", "u", "h_guess")}
    fs = None
    if runner is None:
        pos = host["pos"].to(dev, non_blocking=True)
        vel = host["vel"].to(dev, non_blocking=True)
        u = host["u"].to(dev, non_blocking=True)
        hg = host["h_guess"].to(dev, non_blocking=True)
        st0 = E.Status(1, dev)
        posh = eng.pack_posh(pos, hg)
        grid, delta = eng.grid_for(posh)
        dg = eng.build(posh, grid, delta)
        h_s = eng.newton(dg, consts, dg.posh[:, 3].contiguous(), st0.ptr(0))
        h = eng.scatter(h_s, dg.perm, 1)
        E.raise_for_status(st0.read(), printer=lambda *_: None)
        del dg, posh, h_s
        statuses = []
        if use_lists:
            fs = E.FastStepper(eng)
            fs.load(pos, vel, u, h)

            def step(s):
                fs.advance(dt, consts, defer_status=True)
                statuses.append(fs.last_status)
                return s
            get_state = fs.state
        else:
            stepper = E.Stepper(eng)

            def step(s):
                r = stepper.step(s[0], s[1], s[2], s[3], dt, consts, defer_status=True)
                statuses.append(stepper.last_status)
                return r
            get_state = None
        state = (pos, vel, u, h)
    else:
        state = runner.cold_solve(state)
        state = runner.calibrate(state, dt)   # setup: throw-away steps to balance the cut planes by measured time
        step = lambda s: runner.step(s, dt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_status():
        if world > 1:
            runner.check_status()
        else:
            for st in statuses:
                if st is not None:
                    E.raise_for_status(st.read(), printer=lambda *_: None)
            statuses.clear()

    for _ in range(a.warm_state_steps):       # untimed: let the flow develop before anything is measured
        state = step(state)
    flush_status()
    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(a.warmup):
        state = step(state)
    barrier()
    lib = eng.lib
    launches0 = lib.sphb_launch_count()
    stats0 = dict(fs.stats) if fs is not None else (dict(runner.stats) if runner is not None and hasattr(runner, "stats") else {})
    if clocks:
        clocks.mark_start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        state = step(state)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib.sphb_launch_count() - launches0
    stats1 = dict(fs.stats) if fs is not None else (dict(runner.stats) if runner is not None and hasattr(runner, "stats") else {})
    list_stats = {kk: stats1[kk] - stats0.get(kk, 0) for kk in stats1}
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    flush_status()
    clk = clocks.stop() if clocks else None
    value = n_total * a.steps / (ms * 1e-3)

    # ---- state fingerprint after warm-state + warm-up + timed steps (the same at every N when the runs agree bit for bit)
    if runner is None:
        cur = get_state() if get_state is not None else state
        csum = bit_checksum([cur[0], cur[1], cur[2], cur[3]])
        sums = [float(cur[3].sum()), float(cur[2].sum()), float(cur[1].norm(dim=1).sum()), float(cur[0].norm(dim=1).sum())]
    else:
        csum = bit_checksum([state["pos"], state["vel"], state["u"], state["h"]])
        dist.all_reduce(csum, op=dist.ReduceOp.SUM)
        sv = torch.stack([state["h"].sum(), state["u"].sum(), state["vel"].norm(dim=1).sum(), state["pos"].norm(dim=1).sum()])
        dist.all_reduce(sv, op=dist.ReduceOp.SUM)
        sums = sv.tolist()
    checksum = {"bits_mod_2_64": int(csum.item()) & 0xFFFFFFFFFFFFFFFF, "sum_h": sums[0], "sum_u": sums[1], "sum_speed": sums[2],
                "sum_radius": sums[3], "steps_taken": a.warm_state_steps + a.warmup + a.steps}

    # ---- e2e: public API semantics on pinned host buffers (H2D + step + D2H each step) ----------------
    e2e = None
    if not a.no_e2e and runner is None:
        # the public call with pinned host tensors: pyfluid.var_smoothlength_leapfrog(pos, vel, u, h, dt)
        import pyfluid as sph
        sph.G = 0.0
        sph.PARTICLE_MASS = w["m"]
        if not use_lists:
            os.environ["SPHB_NO_NLIST"] = "1"
        cur = get_state() if get_state is not None else state
        hin = [t.cpu().pin_memory() for t in cur]
        hout = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in hin]   # two pinned sets used in turn
        nb = sum(t.numel() * 8 for t in hin)
        ke = max(1, min(a.steps, 10))
        if fs is not None:
            fs.nl = None          # the public module keeps its own lists: free these (8 GB) first
            torch.cuda.empty_cache()
        for _ in range(2):   # warm-up (device allocations, streams, the module's own neighbour lists)
            sph.var_smoothlength_leapfrog(hin[0], hin[1], hin[2], hin[3], dt, out=hout)
            hin, hout = hout, hin
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(ke):
            sph.var_smoothlength_leapfrog(hin[0], hin[1], hin[2], hin[3], dt, out=hout)
            hin, hout = hout, hin
        el = time.perf_counter() - t0
        e2e = {"value": n_total * ke / el, "unit": "particle-steps/s", "h2d_bytes_per_step": nb, "d2h_bytes_per_step": nb,
               "steps": ke, "api": "pyfluid.var_smoothlength_leapfrog(pos, vel, u, h, dt, out=...) on pinned host tensors, two sets used in turn (transfers overlapped with kernels)"}
        pst = sph._get_stepper()
        if hasattr(pst, "stats"):
            e2e["list_stats"] = dict(pst.stats)
            pst.nl = None
            pst.stale = True
            torch.cuda.empty_cache()
    elif runner is not None and not a.no_e2e:
        e2e = runner.e2e(state, dt, max(1, min(a.steps, 3)))

    # ---- per-kernel profile + roofline ---------------------------------------------------------------------
    roof, per_kernel = None, None
    if runner is None:
        fp64_peak = max(eng.fp64_peak() for _ in range(3))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        if fs is not None and fs.nl is None and fs.hold == 0:
            fs.stale = True
        # the same number of steps again with CUDA events around every C-ABI call (includes the rebuilds they need)
        b0 = fs.stats["builds"] if fs is not None else 0
        eng.prof_start()
        for _ in range(a.steps):
            state = step(state)
        prof = eng.prof_stop()
        flush_status()
        prof_builds = (fs.stats["builds"] - b0) if fs is not None else 3 * a.steps
        cur = get_state() if get_state is not None else state
        n_own, n_sym = interactions(eng, E, consts, cur[0], cur[3], n_total, dev)
        # Newton evaluations per particle (iterations, + 1 for the fused Omega of the mid-step solve)
        if fs is not None and fs.nl is not None and fs.hold == 0:
            stx = E.Status(1, dev)
            _, _, _, its = eng.nl_newton(fs.nl, consts, fs.P, stx.ptr(0), want_posh=False, want_iters=True)
            newton_its = float(its.double().mean().item())
        else:
            newton_its = 1.0
        step_ms = sum(v[1] for v in prof.values()) / a.steps
        per_kernel = {}
        for name, (calls, tms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
            per_kernel[name] = {"calls_per_step": calls / a.steps, "ms_per_step": round(tms / a.steps, 4), "ms_per_call": round(tms / calls, 4)}

        def fp64_line(name, flops_per_call):
            if name in prof:
                ms1 = prof[name][1] / prof[name][0]
                tf = flops_per_call / (ms1 * 1e-3) / 1e12
                per_kernel[name].update({"tflops": tf, "frac_fp64": tf / fp64_peak})

        def hbm_line(name, bytes_per_call):
            if name in prof:
                ms1 = prof[name][1] / prof[name][0]
                gbs = bytes_per_call / (ms1 * 1e-3) / 1e9
                per_kernel[name].update({"gbs": gbs, "frac_hbm": gbs / hbm_peak})

        evals_newton = 0.5 * ((newton_its + 1.0) + newton_its)       # mean over the two launches of a step
        fp64_line("nl_force", FLOP_FORCE * n_sym)
        fp64_line("force", FLOP_FORCE * n_sym)
        fp64_line("nl_density_omega", FLOP_OWN * n_own)
        fp64_line("density_omega", FLOP_OWN * n_own)
        fp64_line("nl_newton_h", FLOP_OWN * n_own * evals_newton)
        fp64_line("newton_h", FLOP_OWN * n_own * evals_newton)
        hbm_line("grid_build", 188.0 * n_total)
        hbm_line("predict", 144.0 * n_total)
        hbm_line("correct", 144.0 * n_total)
        hbm_line("eos", (32.0 + 24.0 + 8.0 + 8.0 + 64.0) * n_total)
        hbm_line("nl_check", 64.0 * n_total)
        if fs is not None and fs.nl is not None:
            rows = fs.nl.rows_used()
            hbm_line("nl_build", 128.0 * rows + 56.0 * n_total)   # list rows written + fp32 records / cell ids read
            per_kernel.get("nl_build", {}).update({"list_rows": rows, "list_bytes": 128 * rows})
        top = max(prof.items(), key=lambda kv: kv[1][1])[0]
        topk = per_kernel[top]
        # step fraction = (F / peak_fp64 + B_K1K5 / peak_hbm) / t_step per particle
        F = FLOP_OWN * (n_own / n_total) * (2.0 * newton_its + 2.0) + FLOP_FORCE * (n_sym / n_total) * 2.0
        builds_per_step = prof_builds / a.steps
        B15 = 288.0 + 188.0 * builds_per_step
        lower_ms = (F / (fp64_peak * 1e12) + B15 / (hbm_peak * 1e9)) * n_total * 1e3
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            ent = tr.get("k_" + top) or tr.get(top)
            if ent:
                traffic = ent["dram_bytes"] / ent["n_particles"] * n_total
                traffic_src = f"{ent['source']}: {ent['dram_bytes']:.4g} B at {ent['n_particles']} particles, scaled per particle"
        except (OSError, ValueError, KeyError):
            pass
        roof = {"kernel": "k_" + top, "bound": "fp64" if "tflops" in topk else "hbm",
                "achieved": topk.get("tflops", topk.get("gbs")), "peak": fp64_peak if "tflops" in topk else hbm_peak,
                "unit": "TFLOP/s" if "tflops" in topk else "GB/s", "frac": topk.get("frac_fp64", topk.get("frac_hbm")),
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "DFMA microbenchmark in this run (sphb_fp64_peak); MEASURED_PEAKS.json has no fp64 entry; HBM: " + hbm_src,
                "launch_ms": topk["ms_per_call"], "share_of_step": prof[top][1] / a.steps / step_ms,
                "step_frac": lower_ms / (ms / a.steps), "step_lower_bound_ms": lower_ms,
                "step_flops_per_particle": F, "step_hbm_bytes_per_particle": B15,
                "hbm_peak": hbm_peak, "newton_iterations_per_solve": newton_its, "builds_per_step": builds_per_step,
                "interactions_per_particle": {"own": n_own / n_total, "symmetric": n_sym / n_total}}

    if runner is not None:
        # rank-0 view of the dominant kernel in the decomposed run: its own particles, its own launch time
        prof_build = None
        if hasattr(runner, "stale"):
            runner.stale = True           # one step that rebuilds the lists (and migrates, re-plans the halo) ...
            eng.prof_start()
            state = step(state)
            prof_build = eng.prof_stop()
            runner.stale = False          # ... and one that runs on kept lists
        eng.prof_start()
        state = step(state)
        prof = eng.prof_stop()
        runner.statuses.clear()
        # per-rank time of the step that rebuilt: how well the cut planes balance the ranks
        names = ["nl_build", "nl_force", "nl_newton_h", "nl_density_omega", "grid_build"]
        pb = prof_build if prof_build is not None else prof
        mine = [pb.get(nm, (0, 0.0))[1] for nm in names]
        mine.append(sum(t_ for nm, (c_, t_) in pb.items() if not nm.startswith("host:")))
        mine.append(sum(t_ for nm, (c_, t_) in pb.items() if nm.startswith("host:")))
        mine.append(float(state["gid"].numel()))
        mine.append(float(runner.nl.n - state["gid"].numel()) if getattr(runner, "nl", None) is not None else 0.0)
        mine.append(float(runner.nl.warp_rows.double().mean().item()) if getattr(runner, "nl", None) is not None else 0.0)
        tv = torch.tensor(mine, dtype=torch.float64, device=dev)
        allv = [torch.empty_like(tv) for _ in range(world)]
        dist.all_gather(allv, tv)
        rank_table = {"columns": names + ["kernels_ms", "host_sections_ms", "owned", "ghosts", "rows_per_warp"],
                      "rows": [[round(x, 3) for x in v.tolist()] for v in allv]}
        if rank == 0:
            fp64_peak = max(eng.fp64_peak() for _ in range(3))
            n_own_r = int(state["gid"].numel())
            fname = "nl_force" if "nl_force" in prof else "force"
            f_ms = prof[fname][1] / prof[fname][0]
            ach = FLOP_FORCE * 80.37 * n_own_r / (f_ms * 1e-3) / 1e12
            roof = {"kernel": "k_" + fname, "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach / fp64_peak, "traffic": None, "scope": f"rank 0 of {world}: {n_own_r} owned particles",
                    "launch_ms": f_ms, "peak_source": "DFMA microbenchmark in this run (sphb_fp64_peak)"}
            per_kernel = {nm: {"calls": c_, "ms": round(t_, 4)} for nm, (c_, t_) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
            if prof_build is not None:
                per_kernel = {"step_on_kept_lists": per_kernel,
                              "step_with_rebuild": {nm: {"calls": c_, "ms": round(t_, 4)}
                                                    for nm, (c_, t_) in sorted(prof_build.items(), key=lambda kv: -kv[1][1])}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        cur = get_state() if get_state is not None else state
        h_host = cur[3].cpu().numpy()
        wcur = dict(pos=cur[0].cpu().numpy(), vel=cur[1].cpu().numpy(), u=cur[2].cpu().numpy())
        v, cores, nj, el = cpu_sample(wcur, h_host, a.cpu_seconds, dict(G=0.0, PARTICLE_MASS=w["m"]))
        cpu = {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": "port",
               "sample": f"{nj} particles against all {n_total} (per-particle share of one leapfrog step, oracle/sph_oracle.c), {el:.1f} s"}

    line = {"metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": w["name"], "n_particles": n_total, "dt": dt, "warm_state_steps": a.warm_state_steps,
                       "l2": f"inputs_larger_than_L2 (state {n_total * 64 / 1e9:.2f} GB vs 126 MB L2)",
                       "parallelism": f"{decomp}{world}" if world > 1 else "single",
                       "path": "kept neighbour lists (rebuilt when the device-side check says so)" if use_lists else "scanning kernels (3 cell-list builds and 3 scans per step)"},
            "clocks": clk, "gpu_launches": int(launches), "neighbour_lists": list_stats, "state_checksum": checksum,
            "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "kernels": per_kernel}
    if world > 1:
        line["ranks"] = rank_table
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
