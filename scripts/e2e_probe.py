"""Where does the host-buffer step spend its time?  (pinned allocation, H2D, step, D2H; several repetitions)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "python-sph_b200")]
import numpy as np, torch
import engine as E, workloads
import pyfluid as sph
w = workloads.sod_tube(int(sys.argv[1]) if len(sys.argv) > 1 else 384)
sph.G = 0.0
sph.PARTICLE_MASS = w["m"]
t = time.perf_counter()
host = [torch.from_numpy(np.ascontiguousarray(w[k])).pin_memory() for k in ("pos", "vel", "u", "h_guess")]
print(f"pin 1 GB of inputs: {time.perf_counter() - t:.3f} s")
dev = [x.cuda() for x in host]
h = sph.newton_smoothlength_arr(dev[0], dev[3])
host[3] = h.cpu().pin_memory()
torch.cuda.synchronize()
for rep in range(3):
    t = time.perf_counter()
    d = [x.to("cuda", non_blocking=True) for x in host]
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    back = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in host]
    t2 = time.perf_counter()
    for b, x in zip(back, d):
        b.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    nb = sum(x.numel() * 8 for x in host) / 1e9
    print(f"H2D {nb / (t1 - t):.1f} GB/s ({(t1 - t) * 1e3:.1f} ms)  pinned alloc {(t2 - t1) * 1e3:.1f} ms  D2H {nb / (t3 - t2):.1f} GB/s ({(t3 - t2) * 1e3:.1f} ms)")
    del back, d
hin = host
hout = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in host]
for rep in range(8):
    torch.cuda.synchronize()
    t = time.perf_counter()
    sph.var_smoothlength_leapfrog(hin[0], hin[1], hin[2], hin[3], w["dt"], out=hout)
    hin, hout = hout, hin
    print(f"call {rep}: {(time.perf_counter() - t) * 1e3:.1f} ms")
