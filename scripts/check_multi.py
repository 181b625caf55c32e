"""k-rank vs 1-rank comparison (SURVEY.md 8(e) determinism target), run under torchrun on k GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_multi.py
Every rank steps its slab; rank 0 also steps the whole set on its GPU with the same grid geometry and compares
(bit-identical fields expected)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "python-sph_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import domain  # noqa: E402
import engine as E  # noqa: E402
import workloads  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = E.Engine(dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    nsteps = int(os.environ.get("CHECK_STEPS", "2"))
    lists = os.environ.get("CHECK_PATH", "lists") == "lists"
    decomp = os.environ.get("CHECK_DECOMP", "slab")
    if len(sys.argv) > 2 and sys.argv[2] == "plummer":
        w = workloads.plummer(n, seed=3)
    else:
        w = workloads.sod_tube(int(sys.argv[2])) if len(sys.argv) > 2 else workloads.uniform_random(n, seed=3, vscale=0.5)
    k = dict(PARTICLE_MASS=w["m"], COUPLING_CONST=1.3, SMOOTHLENGTH_VARIATION_TOLERANCE=1e-3, NEWTON_ITERATION_LIMIT=20,
             BISECTION_ITERATION_LIMIT=100, ALPHA_SPH=1.0, BETA_SPH=2.0, EPSILON=0.01, BISECTION_TOLERANCE=1e-10,
             INITIAL_BISECTION_SMOOTHLENGTH_A=1e3, INITIAL_BISECTION_SMOOTHLENGTH_B=1e-3, G=float(os.environ.get("CHECK_G", "0")),
             ADIABATIC_INDEX=5 / 3)
    consts = E.make_consts(k)
    dt = w["dt"]
    runner = domain.make_runner(eng, consts, world, rank, lists=lists, decomp=decomp)
    st = runner.scatter_initial(w)
    geom0 = runner._global_geom(st["pos"], st["h"]) if decomp == "slab" else None
    st = runner.cold_solve(st)
    geom1 = runner.geom
    if geom0 is None:
        geom0 = geom1          # range decomposition: the cold solve's own (global) cell list
    if os.environ.get("CHECK_CALIBRATE"):
        st = runner.calibrate(st, dt)          # re-cuts the slabs by measured time: must not change a bit of the physics
    for _ in range(nsteps):
        st = runner.step(st, dt)
    runner.check_status()
    full = runner.gather_state(st)
    res = {}
    if rank == 0:
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        pos, vel, u, hg = to(w["pos"]), to(w["vel"]), to(w["u"]), to(w["h_guess"])
        s0 = E.Status(1, dev)
        dg = eng.build(eng.pack_posh(pos, hg), *geom0)
        h = eng.scatter(eng.newton(dg, consts, dg.posh[:, 3].contiguous(), s0.ptr(0)), dg.perm, 1)
        if lists:
            # the single-GPU form of the same path: its builds take the geometry from the same (global) bounding box
            stp = E.FastStepper(eng)
            stp.load(pos, vel, u, h, consts)
            for _ in range(nsteps):
                stp.advance(dt, consts, defer_status=False)
            s = stp.state()
            res["_lists"] = {"ranks": dict(runner.stats), "single": dict(stp.stats)}
        else:
            stp = E.Stepper(eng)
            stp.geom = geom1
            s = (pos, vel, u, h)
            for _ in range(nsteps):
                s = stp.step(s[0], s[1], s[2], s[3], dt, consts)
        for name, a, b in (("pos", full["pos"], s[0]), ("vel", full["vel"], s[1]), ("u", full["u"], s[2]), ("h", full["h"], s[3])):
            res[name] = {"max_abs_diff": float((a - b).abs().max().item()), "bit_identical": bool(torch.equal(a, b)),
                         "scale": float(b.abs().max().item())}
        stats = res.pop("_lists", None)
        print(json.dumps({"world": world, "n": int(pos.shape[0]), "steps": nsteps, "path": "lists" if lists else "scan", "decomp": decomp,
                          "list_stats": stats, "compare": res}))
        ok = all(v["max_abs_diff"] <= 1e-12 * max(v["scale"], 1e-300) for v in res.values())
        if not ok:
            dist.destroy_process_group()
            sys.exit(1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
