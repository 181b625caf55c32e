"""host/device time split of SlabRunner.step (run under torchrun)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "python-sph_b200")]
import numpy as np, torch, torch.distributed as dist
import domain, engine as E, workloads

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
eng = E.Engine(dev)
w = workloads.sod_tube(int(sys.argv[1]) if len(sys.argv) > 1 else 384)
k = dict(PARTICLE_MASS=w["m"], COUPLING_CONST=1.3, SMOOTHLENGTH_VARIATION_TOLERANCE=1e-3, NEWTON_ITERATION_LIMIT=20,
         BISECTION_ITERATION_LIMIT=100, ALPHA_SPH=1.0, BETA_SPH=2.0, EPSILON=0.01, BISECTION_TOLERANCE=1e-10,
         INITIAL_BISECTION_SMOOTHLENGTH_A=1e3, INITIAL_BISECTION_SMOOTHLENGTH_B=1e-3, G=0.0, ADIABATIC_INDEX=5 / 3)
consts = E.make_consts(k)
r = domain.SlabRunner(eng, consts, world, rank)
st = r.cold_solve(r.scatter_initial(w))
st = r.calibrate(st, w["dt"])
for _ in range(2):
    st = r.step(st, w["dt"])
torch.cuda.synchronize(); dist.barrier()
# kernel time via engine profiling, wall time around the step
eng.prof_start()
t0 = time.perf_counter()
st = r.step(st, w["dt"])
torch.cuda.synchronize()
wall = time.perf_counter() - t0
prof = eng.prof_stop()
kern = sum(v[1] for v in prof.values())
allk = [None] * world
dist.all_gather_object(allk, {"rank": rank, "n_own": int(st["gid"].numel()), "kernel_ms": round(kern, 2), "wall_ms": round(wall * 1e3, 2),
                              "force": round(prof["force"][1], 2), "newton": round(prof["newton_h"][1], 2)})
if rank == 0:
    print(json.dumps(allk))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
    st = r.step(st, w["dt"])
    torch.cuda.synchronize()
if rank == 0:
    print(json.dumps({"world": world, "n_own": int(st["gid"].numel()), "wall_ms": wall * 1e3, "sphb_kernel_ms": kern,
                      "kernels": {k_: round(v[1], 2) for k_, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}}))
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=50))
dist.barrier()
dist.destroy_process_group()
