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
    ap.add_argument("--workload", default="sod", choices=["sod", "lattice", "planar"])
    ap.add_argument("--nx", type=int, default=384, help="sod: left-half lattice points along x (384 -> 15.9 M)")
    ap.add_argument("--side", type=int, default=100, help="lattice: cube side")
    ap.add_argument("--n", type=int, default=262144, help="planar: particle count")
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
    return workloads.planar_uniform(a.n, vel_noise=0.01)


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
    o.set_wide(True)
    h = w["h_guess"] * (1.3012 / 1.2) if a.workload == "sod" else w["h_guess"]
    rng = np.random.default_rng(7)
    cand = rng.choice(n, size=min(n, 65536), replace=False)
    t0 = time.perf_counter()
    o.step_work_sample(w["pos"], w["vel"], w["u"], h, cand[:1])
    t1 = time.perf_counter() - t0
    budget = 150.0 / max(1, a.steps + a.warmup)
    nj = int(max(1, min(4096, budget / max(t1, 1e-6))))
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
    dt = w["dt"]

    if world > 1:
        import domain
        runner = domain.SlabRunner(eng, consts, world, rank)
        state = runner.scatter_initial(w)
    else:
        runner = None

    # ---- setup (untimed): state to HBM, cold h solve ------------------------------------------------
    host = {key: torch.from_numpy(np.ascontiguousarray(w[key])).pin_memory() for key in ("pos", "vel", "u", "h_guess")}
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
        stepper = E.Stepper(eng)
        step = lambda s: stepper.step(s[0], s[1], s[2], s[3], dt, consts, defer_status=True)
        state = (pos, vel, u, h)
        statuses = []
    else:
        state = runner.cold_solve(state)
        state = runner.calibrate(state, dt)   # setup: two throw-away steps to balance the cut planes by measured time
        step = lambda s: runner.step(s, dt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(a.warmup):
        state = step(state)
        if runner is None:
            statuses.append(stepper.last_status)
    barrier()
    lib = eng.lib
    launches0 = lib.sphb_launch_count()
    if clocks:
        clocks.mark_start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        state = step(state)
        if runner is None:
            statuses.append(stepper.last_status)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib.sphb_launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        runner.check_status()
    else:
        for st in statuses:
            E.raise_for_status(st.read(), printer=lambda *_: None)
    clk = clocks.stop() if clocks else None
    value = n_total * a.steps / (ms * 1e-3)

    # ---- e2e: public API semantics on pinned host buffers (H2D + step + D2H each step) ----------------
    e2e = None
    if not a.no_e2e and runner is None:
        # the public call with pinned host tensors: pyfluid.var_smoothlength_leapfrog(pos, vel, u, h, dt)
        import pyfluid as sph
        sph.G = 0.0
        sph.PARTICLE_MASS = w["m"]
        hin = [t.cpu().pin_memory() for t in state]
        hout = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in hin]   # two pinned sets used in turn
        nb = sum(t.numel() * 8 for t in hin)
        ke = max(1, min(a.steps, 3))
        for _ in range(2):   # warm-up (device allocations, streams)
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
    elif runner is not None and not a.no_e2e:
        e2e = runner.e2e(state, dt, max(1, min(a.steps, 3)))

    # ---- per-kernel profile of one more step + roofline -------------------------------------------------
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
        eng.prof_start()
        state2 = step(state)
        prof = eng.prof_stop()
        # interaction counts of the current state (own support and symmetric support)
        posh = eng.pack_posh(state[0], state[3])
        grid, delta = eng.grid_for(posh)
        dg = eng.build(posh, grid, delta)
        _, _, _, ng = eng.density_omega(dg, consts, want_rho=False, want_ngb=True)
        n_own = float(ng.double().sum().item())
        q = torch.arange(0, n_total, max(1, n_total // 200000), dtype=torch.int32, device=dev)
        _, cs = eng.neighbours(dg, q, 1, symmetric=True)
        n_sym = float(cs.double().mean().item()) * n_total
        del dg, posh, state2
        per_kernel = {}
        for name, (calls, tms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
            per_kernel[name] = {"calls": calls, "ms": round(tms, 4)}
        step_ms = sum(v[1] for v in prof.values())
        f_ms = prof["force"][1] / prof["force"][0]
        flops_force = FLOP_FORCE * n_sym
        ach = flops_force / (f_ms * 1e-3) / 1e12
        bytes_force = 104.0 * n_total
        # DRAM bytes per launch: dram__bytes_read+write of k_force from the committed ncu --set full capture
        # (profiles/r01_ncu_final_summary.txt: 570.9 MB read + 73.4 MB written for 1 990 656 particles; the parked
        # neighbour lists are most of the reads), scaled per particle
        traffic_force = 644.308e6 / 1990656 * n_total
        roof = {"kernel": "k_force", "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                "traffic": traffic_force, "traffic_source": "ncu capture at 1 990 656 particles scaled per particle (profiles/)", "peak_source": "DFMA microbenchmark in this run (sphb_fp64_peak); not in MEASURED_PEAKS.json",
                "launch_ms": f_ms, "share_of_step": prof["force"][1] / step_ms,
                "hbm": {"achieved": bytes_force / (f_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_force / (f_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
                "interactions_per_particle": {"own": n_own / n_total, "symmetric": n_sym / n_total}}
        own_ms = prof["density_omega"][1] / prof["density_omega"][0]
        per_kernel["density_omega"]["tflops"] = FLOP_OWN * n_own / (own_ms * 1e-3) / 1e12
        per_kernel["density_omega"]["frac_fp64"] = per_kernel["density_omega"]["tflops"] / fp64_peak
        nw_ms = prof["newton_h"][1] / prof["newton_h"][0]
        per_kernel["newton_h"]["tflops"] = FLOP_OWN * n_own / (nw_ms * 1e-3) / 1e12
        per_kernel["newton_h"]["frac_fp64"] = per_kernel["newton_h"]["tflops"] / fp64_peak
        per_kernel["force"]["tflops"] = ach
        per_kernel["force"]["frac_fp64"] = ach / fp64_peak
        gb_ms = prof["grid_build"][1] / prof["grid_build"][0]
        per_kernel["grid_build"]["gbs"] = 188.0 * n_total / (gb_ms * 1e-3) / 1e9
        per_kernel["grid_build"]["frac_hbm"] = per_kernel["grid_build"]["gbs"] / hbm_peak
        for nm in ("predict", "correct"):
            pm = prof[nm][1] / prof[nm][0]
            per_kernel[nm]["gbs"] = 144.0 * n_total / (pm * 1e-3) / 1e9
            per_kernel[nm]["frac_hbm"] = per_kernel[nm]["gbs"] / hbm_peak

    if runner is not None:
        # rank-0 view of the dominant kernel in the decomposed run: its own particles, its own launch time.
        # Interactions per particle are the single-GPU measurement of this workload (80.37 symmetric pairs).
        eng.prof_start()
        state = step(state)
        prof = eng.prof_stop()
        runner.statuses.clear()
        if rank == 0:
            fp64_peak = max(eng.fp64_peak() for _ in range(3))
            n_own = int(state["gid"].numel())
            f_ms = prof["force"][1] / prof["force"][0]
            ach = FLOP_FORCE * 80.37 * n_own / (f_ms * 1e-3) / 1e12
            roof = {"kernel": "k_force", "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach / fp64_peak, "traffic": None, "scope": f"rank 0 of {world}: {n_own} owned particles",
                    "launch_ms": f_ms, "peak_source": "DFMA microbenchmark in this run (sphb_fp64_peak)"}
            per_kernel = {nm: {"calls": c_, "ms": round(t_, 4)} for nm, (c_, t_) in sorted(prof.items(), key=lambda kv: -kv[1][1])}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        h_host = state[3].cpu().numpy()
        wcur = dict(pos=state[0].cpu().numpy(), vel=state[1].cpu().numpy(), u=state[2].cpu().numpy())
        v, cores, nj, el = cpu_sample(wcur, h_host, a.cpu_seconds, dict(G=0.0, PARTICLE_MASS=w["m"]))
        cpu = {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": "port",
               "sample": f"{nj} particles against all {n_total} (per-particle share of one leapfrog step, oracle/sph_oracle.c), {el:.1f} s"}

    line = {"metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": w["name"], "n_particles": n_total, "dt": dt, "l2": "inputs_larger_than_L2 (state ~1 GB vs 126 MB L2)",
                       "parallelism": f"slab{world}" if world > 1 else "single"},
            "clocks": clk, "gpu_launches": int(launches), "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
            "kernels": per_kernel}
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
