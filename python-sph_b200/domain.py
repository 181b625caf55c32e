"""Slab decomposition of the per-step particle loop over the GPUs of one box.

One process per GPU.  Rank r owns the particles with cuts[r] <= x < cuts[r+1] (cut planes chosen so
that ranks own equal particle counts).  The reference has no distributed code; the only data-parallel
axis is "every particle's sums are independent given a frozen snapshot" (pyfluid.py:203-207,:487-491),
so each gather phase needs exactly one exchange of ghost particles:

    halo(pos, h, vel, u)   before every cell-list build      all_to_all_single (NCCL over NVLink)
    halo(h_new)            after a Newton solve               same send lists
    halo(omega)            after an Omega pass (force needs Omega_i of ghost neighbours)
    migration + re-plan    every `replan_every` steps         full 64-byte records + global id, new send lists
    allreduce(max/min/sum) h_max, bounding box, status words  a few scalars, only at a re-plan / status check

The send lists (the halo "plan") are made with a skin and reused for every snapshot until the next re-plan, so a
steady-state step contains no device->host read; drift and h growth are checked against the plan on the device.
Ghost width = 2 h_max (1 + margin) + skin: the force support is 2 max(h_i, h_j).  Local arrays are kept sorted
by global particle id before each build, so the stable cell sort visits neighbours in the same order as
a single-GPU run with the same grid geometry: k-rank results are bit-identical to 1-rank results.

`HaloExchanger` is pure plumbing on torch tensors (any device, any backend) and is what the world-size-2
gloo tests on CPU exercise; `SlabRunner` drives the CUDA kernels through engine.Engine.
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist

import engine as E

F64 = torch.float64


def equal_count_cuts(x_sorted_sample, world):
    """Cut planes from a sorted sample of x coordinates: world+1 values, ends at -inf/+inf."""
    n = x_sorted_sample.numel()
    cuts = [-math.inf]
    for r in range(1, world):
        k = (n * r) // world
        # midway between neighbours so that equal coordinates (lattice planes) stay on one side
        a, b = float(x_sorted_sample[max(k - 1, 0)]), float(x_sorted_sample[min(k, n - 1)])
        cuts.append(0.5 * (a + b))
    cuts.append(math.inf)
    return cuts


class HaloExchanger:
    """Send lists + all_to_all plumbing for one snapshot of owned x coordinates."""

    def __init__(self, cuts, rank, world, group=None):
        self.cuts, self.rank, self.world, self.group = list(cuts), rank, world, group

    def owner_of(self, x):
        inner = torch.tensor(self.cuts[1:-1], dtype=x.dtype, device=x.device)
        return torch.bucketize(x, inner, right=True)

    def plan(self, x_owned, width):
        """indices (into the owned arrays) to send to every other rank: everything within `width` of its slab"""
        lo = torch.tensor([c - width for c in self.cuts[:-1]], dtype=x_owned.dtype, device=x_owned.device)
        hi = torch.tensor([c + width for c in self.cuts[1:]], dtype=x_owned.dtype, device=x_owned.device)
        m = (x_owned[None, :] >= lo[:, None]) & (x_owned[None, :] < hi[:, None])   # (world, n_owned)
        m[self.rank] = False
        self._from_mask(m)
        return self

    def plan_migration(self, x_owned):
        """send lists of the particles that left the slab.  self.moved = how many do, over all ranks (0: nothing to do,
        and then no index array is made at all)"""
        own = self.owner_of(x_owned)
        leaving = own != self.rank
        cnt = leaving.sum().reshape(1)
        if self.world > 1:
            dist.all_reduce(cnt, group=self.group)
        self.moved = int(cnt.item())
        if self.moved == 0:
            self.stay = None
            self.all_idx = torch.empty(0, dtype=torch.int64, device=x_owned.device)
            self.send_counts = self.recv_counts = [0] * self.world
            return self
        idx = torch.nonzero(leaving).flatten()
        dest = own.index_select(0, idx)
        order = torch.argsort(dest, stable=True)              # destination-major, owned order within a destination
        self.all_idx = idx.index_select(0, order)
        self.stay = torch.nonzero(~leaving).flatten()
        send = torch.bincount(dest, minlength=self.world)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        both = torch.stack([send, recv]).tolist()
        self.send_counts, self.recv_counts = both[0], both[1]
        return self

    def _from_mask(self, m):
        """send lists from a (world, n_owned) membership mask: destination-major, owned order within a destination;
        one count exchange, one D2H of 2*world integers"""
        send = m.sum(dim=1)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        self.all_idx = torch.nonzero(m)[:, 1].contiguous()   # row-major: sorted by destination, then owned index
        both = torch.stack([send, recv]).tolist()
        self.send_counts, self.recv_counts = both[0], both[1]

    @property
    def n_recv(self):
        return int(sum(self.recv_counts))

    def exchange(self, arrays):
        """arrays: list of (n_owned, k) or (n_owned,) float64 tensors -> list of received arrays, sources in rank
        order, each source's particles in the sender's owned order (deterministic)."""
        cols = [a.reshape(a.shape[0], -1) for a in arrays]
        widths = [c.shape[1] for c in cols]
        K = sum(widths)
        packed = torch.cat([c.index_select(0, self.all_idx) for c in cols], dim=1) if K else None
        out = torch.empty((self.n_recv, K), dtype=packed.dtype, device=packed.device)
        dist.all_to_all_single(out, packed, output_split_sizes=self.recv_counts, input_split_sizes=self.send_counts,
                               group=self.group)
        res, o = [], 0
        for a, wd in zip(arrays, widths):
            r = out[:, o:o + wd]
            res.append(r.reshape(-1) if a.dim() == 1 else r.contiguous())
            o += wd
        return [r.contiguous() for r in res]


class SlabRunner:
    """var_smoothlength_leapfrog on a slab-decomposed particle set (one rank per GPU)."""

    def __init__(self, eng: E.Engine, consts, world, rank, margin=0.05, balance=True, calib_steps=0):
        self.e, self.consts, self.world, self.rank, self.margin = eng, consts, world, rank, margin
        self.grav = consts.G != 0.0     # the reference's default (pyfluid.py:35): scanning kernels + all-gathered pair sum
        self.balance = balance
        self.park = os.environ.get("SPHB_NO_PARK") is None   # share neighbour lists between kernels of one snapshot
        self.calib_left = calib_steps if balance else 0   # optional: time the first steps per rank and re-balance
        self.dev = eng.device
        self.statuses = []
        self.geom = None
        self._halo_viol = None   # device flag: a solved h / the accumulated drift outgrew the halo plan
        # Halo plan (send lists + ghost ids), reused for every snapshot until the next re-plan: in steady state a step
        # then has NO device->host read at all.  The plan is made with a skin; the drift since it was made and the
        # current h_max are checked against it on the device, and check_status() reports a violation.
        self.plan = None
        self.replan_every = 8     # steps between migrations / re-plans
        self.skin = 0.1           # extra halo width, in units of h_max, that particles may drift before a re-plan

    # ---- setup -------------------------------------------------------------------------------------
    def scatter_initial(self, w):
        """every rank generates the same synthetic workload and keeps its slab"""
        import numpy as np
        x = w["pos"][:, 0]
        xs = np.sort(x[:: max(1, len(x) // 1000000)])
        self.cuts = equal_count_cuts(torch.from_numpy(xs), self.world)
        own = (x >= self.cuts[self.rank]) & (x < self.cuts[self.rank + 1])
        idx = np.nonzero(own)[0]
        self.n_total = len(x)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a[idx])).to(self.dev)
        return dict(gid=torch.from_numpy(idx.astype(np.int64)).to(self.dev), pos=to(w["pos"]), vel=to(w["vel"]), u=to(w["u"]),
                    h=to(w["h_guess"]), wfac=torch.ones(len(idx), dtype=F64, device=self.dev))

    def _global_geom(self, pos, h):
        lo = pos.min(dim=0).values
        hi = pos.max(dim=0).values
        s = torch.stack([h.sum(), torch.tensor(float(h.numel()), dtype=F64, device=self.dev)])
        if self.world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
        lo, hi, s = lo.tolist(), hi.tolist(), s.tolist()
        return E.choose_grid(lo, hi, s[0] / s[1], int(s[1]))

    def _hmax(self, h):
        t = h.max().reshape(1).clone()
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def _make_plan(self, st, extra=1.0):
        """migrate, then fix the halo send lists and the ghost ids until the next re-plan (the only place with
        device->host reads: send counts and h_max)"""
        st = self._migrate(st)
        # the cells follow the gas: the fp32 pre-filter's error bound (delta) only holds while positions stay within a few
        # box lengths of the grid origin, and a free expansion does not (host reads are fine here)
        self.geom = self._global_geom(st["pos"], st["h"])
        hm = self._hmax(st["h"])
        width = 2.0 * hm * (1.0 + self.margin) * extra + 2.0 * self.skin * hm + 4.0 * self.geom[1]
        hx = HaloExchanger(self.cuts, self.rank, self.world).plan(st["pos"][:, 0].contiguous(), width)
        gid_recv = hx.exchange([st["gid"].to(F64)])[0].to(torch.int64)
        self.plan = dict(hx=hx, width=width, gid_all=torch.cat([st["gid"], gid_recv]), pos_ref=st["pos"].clone(), age=0)
        return st

    def _note_plan_use(self, pos, h):
        """device-side validity check of the current plan: 2 h_max + 2 drift must stay inside the planned width"""
        drift = (pos - self.plan["pos_ref"]).abs().amax()
        bad = (2.0 * h.max() + 2.0 * drift > self.plan["width"]).to(torch.int32).reshape(1)
        self._halo_viol = bad if self._halo_viol is None else torch.maximum(self._halo_viol, bad)

    def _snapshot(self, st, fields):
        """exchange ghosts of `fields` (names in st) along the current plan and build the cell list over
        [owned..., ghosts...]; ties inside a cell are broken by global particle id, so the neighbour order does not
        depend on the decomposition."""
        hx = self.plan["hx"]
        recv = hx.exchange([st[f] for f in fields])
        n_own = st["gid"].numel()
        gid_all = self.plan["gid_all"]
        snap = dict(hx=hx, n_own=n_own, n_loc=gid_all.numel())
        for f, g in zip(fields, recv):
            snap[f] = torch.cat([st[f], g])
        posh = self.e.pack_posh(snap["pos"], snap["h"])
        dg = self.e.build(posh, *self.geom, tie_key=gid_all)
        snap["dg"] = dg
        snap["target"] = (dg.perm < n_own).to(torch.uint8)   # grid order: owned particles are the targets
        return snap

    def _to_grid(self, snap, arr_L, ncomp):
        return self.e.gather(arr_L, snap["dg"].perm, ncomp)

    def _owned_from_grid(self, snap, arr_S, ncomp):
        return self.e.scatter(arr_S, snap["dg"].perm, ncomp)[:snap["n_own"]].contiguous()

    def _fill_ghosts(self, snap, owned_vals):
        """owned values (owned order) -> local array [owned..., ghosts...] with the ghosts' values from their owners"""
        return torch.cat([owned_vals, snap["hx"].exchange([owned_vals])[0]])

    # ---- physics stages -----------------------------------------------------------------------------------
    def _stage(self, snap, status_ptr, om_own=None, lists=None):
        """the six calls of pyfluid.py:502-507 for the owned particles -> acc, dudt (owned order)
        lists: neighbour lists parked by the Newton kernel of this snapshot (else the Omega pass parks its own)"""
        e, dg = self.e, snap["dg"]
        if self.grav:
            return self._stage_gravity(snap, status_ptr)
        if om_own is None:
            lists = e.park_lists(dg) if self.park else None
            _, _, om_S, _ = e.density_omega(dg, self.consts, target=snap["target"], want_rho=False, want_omega=True,
                                            lists=lists)
            om_own = self._owned_from_grid(snap, om_S, 1)
        om_L = self._fill_ghosts(snap, om_own)
        vel_S = self._to_grid(snap, snap["vel"], 3)
        u_S = self._to_grid(snap, snap["u"], 1)
        _, _, rB, rC = e.eos(dg.posh, vel_S, u_S, self._to_grid(snap, om_L, 1), self.consts, status_ptr)
        acc_S, dudt_S = e.force(dg, self.consts, rB, rC, e.hmax(dg.posh), status_ptr, target=snap["target"], lists=lists)
        return self._owned_from_grid(snap, acc_S, 3), self._owned_from_grid(snap, dudt_S, 1)

    def _stage_gravity(self, snap, status_ptr):
        """the stage with the softened self-gravity of the reference's default G (pyfluid.py:481-483): xi rides in the
        Omega pass (:401-418), its compact grad-h correction in the force kernel (:434-437), and the 1/r^2 pair sum
        (:420-432) runs over ALL particles -- every rank's records are all-gathered (32 bytes each) and each rank sums
        for its own targets.  O(N^2 / ranks): the N <~ 1e6 regime the exact sum is meant for."""
        e, dg, c = self.e, snap["dg"], self.consts
        _, _, om_S, _, xi_S = e.density_omega(dg, c, target=snap["target"], want_rho=False, want_omega=True, want_xi=True)
        om_own, xi_own = self._owned_from_grid(snap, om_S, 1), self._owned_from_grid(snap, xi_S, 1)
        g = snap["hx"].exchange([om_own, xi_own])
        om_L, xi_L = torch.cat([om_own, g[0]]), torch.cat([xi_own, g[1]])
        vel_S, u_S = self._to_grid(snap, snap["vel"], 3), self._to_grid(snap, snap["u"], 1)
        _, _, rB, rC, bg = e.eos(dg.posh, vel_S, u_S, self._to_grid(snap, om_L, 1), c, status_ptr,
                                 xi=self._to_grid(snap, xi_L, 1), want_bgrav=True)
        acc_S, dudt_S = e.force(dg, c, rB, rC, e.hmax(dg.posh), status_ptr, target=snap["target"], bgrav=bg)
        acc = self._owned_from_grid(snap, acc_S, 3)
        n_own = snap["n_own"]
        posh_own = self._owned_from_grid(snap, dg.posh, 4)    # the grid's records: they carry the solved h at the mid stage
        if self.world > 1:
            n = torch.tensor([n_own], dtype=torch.int64, device=self.dev)
            ns = [torch.empty_like(n) for _ in range(self.world)]
            dist.all_gather(ns, n)
            ns = [int(t.item()) for t in ns]
            src = torch.empty((sum(ns), 4), dtype=F64, device=self.dev)
            dist.all_gather(list(src.split(ns)), posh_own)
        else:
            src = posh_own
        e.gravity_direct_targets(posh_own, src, c, acc)
        return acc, self._owned_from_grid(snap, dudt_S, 1)

    def _newton(self, snap, status_ptr, want_omega=False, lists=None):
        e, dg = self.e, snap["dg"]
        r = e.newton(dg, self.consts, dg.posh[:, 3].contiguous(), status_ptr, target=snap["target"], want_omega=want_omega,
                     lists=lists)
        h_own = self._owned_from_grid(snap, r[0] if want_omega else r, 1)
        e.check_neg_h(h_own, status_ptr)
        return (h_own, self._owned_from_grid(snap, r[1], 1)) if want_omega else h_own

    # cost model of one particle-step, fitted on B200 (profiles/): a fixed part plus the cells its search box
    # covers; sparse particles (large h on a grid sized for the dense majority) cost ~1.65x a dense one
    COST_A, COST_B = 5.2, 0.032

    def rebalance(self, st, nbins=16384):
        """work-balanced cut planes: equalise the summed per-particle cost model between ranks,
        then migrate.  One all-reduce of a weighted x-histogram."""
        cell = self.geom[0].cell
        wgt = (self.COST_A + self.COST_B * (4.0 * st["h"] / cell + 1.0) ** 3) * st["wfac"]
        x = st["pos"][:, 0].contiguous()
        ext = torch.stack([x.min(), -x.max()])
        dist.all_reduce(ext, op=dist.ReduceOp.MIN)
        xlo, xhi = float(ext[0].item()), float(-ext[1].item())
        span = max(xhi - xlo, 1e-300)
        b = ((x - xlo) * (nbins / span)).long().clamp_(0, nbins - 1)
        hist = torch.bincount(b, weights=wgt, minlength=nbins)
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        cum = torch.cumsum(hist, 0).cpu()
        total = float(cum[-1])
        cuts = [-math.inf]
        for r in range(1, self.world):
            k = int(torch.searchsorted(cum, torch.tensor(total * r / self.world, dtype=cum.dtype)).item())
            # cut at the upper edge of bin k, nudged off any lattice plane by half a bin
            cuts.append(xlo + (min(k, nbins - 1) + 1) * span / nbins)
        cuts.append(math.inf)
        self.cuts = cuts
        self.plan = None
        return self._migrate(st)

    def _calibrate(self, st, my_ms):
        """feedback: scale every owned particle's weight by (measured time / modelled weight) of its rank, normalised
        over the job, then re-balance (the weight factor travels with the particle)"""
        cell = self.geom[0].cell
        w_model = float(((self.COST_A + self.COST_B * (4.0 * st["h"] / cell + 1.0) ** 3) * st["wfac"]).sum().item())
        v = torch.tensor([my_ms, w_model], dtype=F64, device=self.dev)
        allv = [torch.empty_like(v) for _ in range(self.world)]
        dist.all_gather(allv, v)
        tt = sum(float(a[0]) for a in allv)
        ww = sum(float(a[1]) for a in allv)
        fac = (my_ms / max(w_model, 1e-300)) / (tt / max(ww, 1e-300))
        st = dict(st)
        st["wfac"] = st["wfac"] * min(max(fac, 0.5), 2.0)
        return self.rebalance(st)

    def calibrate(self, st, dt, rounds=2):
        """setup-time load balancing: run `rounds` throw-away steps from `st`, time this rank's kernels in each, fold the
        measured time into the per-particle cost factors and re-cut.  Returns `st` (unchanged physics) redistributed."""
        if not self.balance or self.world == 1:
            return st
        for _ in range(rounds):
            self.e.prof_start()
            self._step(st, dt)
            ms = sum(v[1] for v in self.e.prof_stop().values())
            self.statuses.pop()          # the throw-away step's status block
            st = self._calibrate(st, ms)
        return st

    def cold_solve(self, st):
        self.geom = self._global_geom(st["pos"], st["h"])
        # the guess is usually an under-estimate and the halo must cover the support of the SOLVED h: widen and
        # repeat until the plan's own validity check passes (setup path: host reads are fine here)
        extra, guess = 1.6, st["h"]
        while True:
            stt = E.Status(1, self.dev)
            st = self._make_plan(dict(st, h=guess), extra=extra)
            guess = st["h"]
            snap = self._snapshot(st, ["pos", "h"])
            h_new = self._newton(snap, stt.ptr(0))
            self._halo_viol = None
            self._note_plan_use(st["pos"], h_new)
            bad = self._halo_viol.clone()
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            self._halo_viol = None
            if not int(bad.item()):
                break
            extra *= 1.6
        st["h"] = h_new
        self.statuses.append(stt)
        self.geom = self._global_geom(st["pos"], st["h"])
        self.plan = None
        if self.balance:
            st = self.rebalance(st)
        return st

    def step(self, st, dt):
        if self.calib_left > 0 and self.e.prof is None:
            self.calib_left -= 1
            self.e.prof_start()
            new = self._step(st, dt)
            ms = sum(v[1] for v in self.e.prof_stop().values())
            return self._calibrate(new, ms)
        return self._step(st, dt)

    def _step(self, st, dt):
        e = self.e
        stt = E.Status(E.S_COUNT, self.dev)
        if self.plan is None or self.plan["age"] >= self.replan_every:
            st = self._make_plan(st)
        self.plan["age"] += 1
        # stage 0
        s0 = self._snapshot(st, ["pos", "h", "vel", "u"])
        acc0, dudt0 = self._stage(s0, stt.ptr(E.S_STAGE0))
        # predictor (owned only)
        posh0 = e.pack_posh(st["pos"], st["h"])
        posh_m, vel_m, u_m = e.predict(posh0, st["vel"], st["u"], acc0, dudt0, dt, stt.ptr(E.S_PRED))
        mid = dict(gid=st["gid"], pos=posh_m[:, :3].contiguous(), h=st["h"], vel=vel_m, u=u_m)
        wfac = st.get("wfac")
        if wfac is None:
            wfac = torch.ones_like(st["h"])
        # h_mid
        s1 = self._snapshot(mid, ["pos", "h", "vel", "u"])
        # the Newton kernel parks symmetric-support lists that stay complete while no h (own or ghost, hence the
        # all-reduce of the flag) grows by more than PARK_COVER; the mid-stage force kernel drains them
        l1 = e.park_lists(s1["dg"], cover=E.PARK_COVER) if (self.park and not self.grav) else None
        h_mid, om_mid = self._newton(s1, stt.ptr(E.S_NEWTON_MID), want_omega=True, lists=l1)
        if l1 is not None:
            dist.all_reduce(l1.broken, op=dist.ReduceOp.MAX)
        self._note_plan_use(mid["pos"], h_mid)
        h_mid_L = self._fill_ghosts(s1, h_mid)
        e.set_h(s1["dg"], self._to_grid(s1, h_mid_L, 1))
        acc1, dudt1 = self._stage(s1, stt.ptr(E.S_STAGE_MID), om_own=om_mid, lists=l1)
        # corrector (owned only, t0 state is still in owned order)
        posh_1, vel_1, u_1 = e.correct(posh0, st["vel"], st["u"], acc1, dudt1, dt, stt.ptr(E.S_CORR))
        new = dict(gid=st["gid"], pos=posh_1[:, :3].contiguous(), h=h_mid, vel=vel_1, u=u_1, wfac=wfac)
        # h1 (migration happens at the next re-plan)
        s2 = self._snapshot(new, ["pos", "h"])
        new["h"] = self._newton(s2, stt.ptr(E.S_NEWTON_1))
        self._note_plan_use(new["pos"], new["h"])
        self.statuses.append(stt)
        return new

    def _migrate(self, st):
        hx = HaloExchanger(self.cuts, self.rank, self.world).plan_migration(st["pos"][:, 0].contiguous())
        if hx.moved == 0:
            return st
        names = [f for f in ("pos", "h", "vel", "u", "wfac", "rows", "rd", "rg") if f in st]
        recv = hx.exchange([st["gid"].to(F64)] + [st[f] for f in names])
        out = dict(gid=torch.cat([st["gid"].index_select(0, hx.stay), recv[0].to(torch.int64)]))
        for f, g in zip(names, recv[1:]):
            out[f] = torch.cat([st[f].index_select(0, hx.stay), g]).contiguous()
        return out

    def check_status(self, printer=lambda *_: None):
        """OR the flag words / sum the counters of every step since the last call, over all ranks"""
        acc = torch.zeros((E.S_COUNT, 8), dtype=torch.int32, device=self.dev)
        for s in self.statuses:
            t = s.t if s.t.shape[0] == E.S_COUNT else torch.cat([s.t, torch.zeros((E.S_COUNT - s.t.shape[0], 8),
                                                                                  dtype=torch.int32, device=self.dev)])
            acc[:, 0] |= t[:, 0]
            acc[:, 1:] += t[:, 1:]
        self.statuses = []
        allr = [torch.empty_like(acc) for _ in range(self.world)]
        dist.all_gather(allr, acc)
        tot = allr[0].clone()
        for t in allr[1:]:
            tot[:, 0] |= t[:, 0]
            tot[:, 1:] += t[:, 1:]
        if self._halo_viol is not None:
            v = self._halo_viol.clone()
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            self._halo_viol = None
            if int(v.item()):
                raise RuntimeError("the halo plan was outgrown (h or particle drift): lower SlabRunner.replan_every or raise "
                                   ".margin / .skin")
        E.raise_for_status(tot.cpu().numpy(), printer=printer)

    def gather_state(self, st):
        """all particles in global-id order on every rank (tests / small runs)"""
        names = ["pos", "h", "vel", "u"]
        n = torch.tensor([st["gid"].numel()], dtype=torch.int64, device=self.dev)
        ns = [torch.empty_like(n) for _ in range(self.world)]
        dist.all_gather(ns, n)
        ns = [int(t.item()) for t in ns]
        packed = torch.cat([st["gid"].to(F64).reshape(-1, 1)] + [st[f].reshape(st[f].shape[0], -1) for f in names], dim=1)
        bufs = [torch.empty((k, packed.shape[1]), dtype=F64, device=self.dev) for k in ns]
        dist.all_gather(bufs, packed.contiguous())
        allp = torch.cat(bufs)
        allp = allp.index_select(0, torch.argsort(allp[:, 0].to(torch.int64)))
        return dict(pos=allp[:, 1:4].contiguous(), h=allp[:, 4].contiguous(), vel=allp[:, 5:8].contiguous(), u=allp[:, 8].contiguous())

    def e2e(self, st, dt, steps):
        """public-API-style step with HOST buffers on every rank: H2D of the owned state from pinned memory, step,
        D2H of the result into pinned memory, every step (buffers have 10 % headroom for migration)"""
        import time
        n0 = st["gid"].numel()
        cap = int(n0 * 1.1) + 1024
        pinned = {k: torch.empty((cap,) + tuple(v.shape[1:]), dtype=v.dtype).pin_memory() for k, v in st.items()}
        for k, v in st.items():
            pinned[k][:n0].copy_(v)
        n = n0
        nbytes = 0
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            d = {k: v[:n].to(self.dev, non_blocking=True) for k, v in pinned.items()}
            nbytes += sum(v.numel() * v.element_size() for v in d.values())
            r = self.step(d, dt)
            n = r["gid"].numel()
            assert n <= cap, "migration exceeded the host buffer headroom"
            for k, v in r.items():
                pinned[k][:n].copy_(v, non_blocking=True)
            torch.cuda.synchronize()
        dist.barrier()
        el = torch.tensor([time.perf_counter() - t0], dtype=F64, device=self.dev)
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        tot = torch.tensor([float(nbytes) / steps], dtype=F64, device=self.dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        return {"value": self.n_total * steps / float(el.item()), "unit": "particle-steps/s",
                "h2d_bytes_per_step": int(tot.item()), "d2h_bytes_per_step": int(tot.item()), "steps": steps}



class SlotExchanger:
    """Ghost refresh in a FROZEN slot order (pure torch plumbing: any device, any backend).

    Between two rebuilds of the neighbour lists the local arrays [owned..., ghosts...] keep their sorted slots; what
    changes is the values.  `refresh` sends the owners' current values along the send lists of the halo plan and writes
    what arrives straight into the ghosts' slots: one pack, one all_to_all_single, one unpack, for any number of arrays."""

    def __init__(self, hx: HaloExchanger, slot_of_loc, n_own, engine=None):
        """engine: an engine.Engine whose rows_pack kernel packs / unpacks in one launch each (CUDA tensors); without
        it the same is done with torch indexing (any device: what the gloo tests run)"""
        self.hx = hx
        self.engine = engine
        self.send_slots = slot_of_loc.index_select(0, hx.all_idx)          # destination-major, owned order within
        self.ghost_slots = slot_of_loc[n_own:].contiguous()                # arrival order: source rank, sender's order
        self.n_recv = hx.n_recv
        self.calls = 0

    def refresh(self, arrays):
        """arrays: slot-order tensors (n_loc,) or (n_loc, k); the ghosts' rows are overwritten in place"""
        cols = [a.reshape(a.shape[0], -1) for a in arrays]
        widths = [c.shape[1] for c in cols]
        K = sum(widths)
        if self.hx.world == 1 or K == 0:
            return
        fast = self.engine is not None and len(arrays) <= 4 and all(a.is_cuda and a.is_contiguous() and a.dtype == F64 for a in arrays)
        if fast:
            packed = torch.empty((self.send_slots.numel(), K), dtype=F64, device=arrays[0].device)
            self.engine.rows_pack(arrays, self.send_slots, packed)
        else:
            packed = torch.cat([c.index_select(0, self.send_slots) for c in cols], dim=1)
        out = torch.empty((self.n_recv, K), dtype=packed.dtype, device=packed.device)
        dist.all_to_all_single(out, packed, output_split_sizes=self.hx.recv_counts, input_split_sizes=self.hx.send_counts,
                               group=self.hx.group)
        self.calls += 1
        if fast:
            self.engine.rows_pack(arrays, self.ghost_slots, out, unpack=True)
            return
        o = 0
        for c, wd in zip(cols, widths):
            c.index_copy_(0, self.ghost_slots, out[:, o:o + wd])
            o += wd


class FastSlabRunner(SlabRunner):
    """SlabRunner on neighbour lists that are kept across kernels and steps (engine.FastStepper's multi-GPU form).

    At a rebuild (all ranks together, decided by the all-reduced list flags): migrate, refresh the grid geometry from
    the global bounding box, fix the ghost set -- everything within 2 h_max (cover_h + skin) of the slab --, build ONE
    cell list over [owned..., ghosts...] with ties broken by global id, and scan it once for the owned targets.  Until
    the next rebuild the local arrays keep that slot order; a step is five list drains, the elementwise kernels, and
    FOUR ghost refreshes (h + Omega, the predicted state, h_mid + Omega_mid, the corrected state) in the frozen order.
    Sums are formed in (cell, global id) order on the same cells as a single-GPU FastStepper run: same bits."""

    def __init__(self, eng: E.Engine, consts, world, rank, margin=0.0, balance=True, calib_steps=0, cover_h=None, skin=None,
                 adaptive=True):
        if consts.G != 0.0:
            raise ValueError("kept neighbour lists run the G = 0 path; make_runner() gives the scanning SlabRunner for G != 0")
        super().__init__(eng, consts, world, rank, margin=margin, balance=balance, calib_steps=calib_steps)
        self._margins0 = (cover_h, skin, adaptive)
        self.margins = E.ListMargins(cover_h, skin, adaptive)
        self.nl = None
        self.fast = None          # slot-order state between rebuilds: dict(P, V, U, gid_S, wfac_S, ...)
        self.stale = True
        self.hold = 0
        self.stats = dict(steps=0, builds=0, redos=0, safe_steps=0, exchanges=0)

    # ---- collectives on a few scalars ----------------------------------------------------------------------
    def _allmax(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    # ---- rebuild ----------------------------------------------------------------------------------------------
    def _rows(self, src, idx):
        """src[idx] for an fp64 / int64 array of row width <= 4, in one launch of the library's row kernel"""
        a = src.view(F64) if src.dtype == torch.int64 else src
        out = torch.empty((idx.numel(),) + tuple(a.shape[1:]), dtype=F64, device=a.device)
        self.e.rows_pack([a], idx, out)
        return out.view(torch.int64) if src.dtype == torch.int64 else out

    def _to_owned(self):
        """slot-order state -> the dict SlabRunner works with (owned particles only, in ascending slot order)"""
        f = self.fast
        own = f["own_sorted"]
        with self.e.section("host:to_owned"):
            P = self._rows(f["P"], own)
            out = dict(gid=self._rows(f["gid_S"], own), pos=P[:, :3].contiguous(), h=P[:, 3].contiguous(),
                       vel=self._rows(f["V"], own), u=self._rows(f["U"], own), wfac=self._rows(f["wfac_S"], own),
                       rows=self._rows(f["rows_S"], own))
            # every record's own motion per step, measured against the lists' build-time records: what the next lists'
            # per-particle margins are sized with (it migrates with the particle)
            if self.margins.per_particle and self.nl is not None:
                if self.margins.age >= 1:
                    rd, rg = E.slot_rates(self.nl.posh_build, f["P"], self.margins.age)
                    out["rd"], out["rg"] = self._rows(rd, own), self._rows(rg, own)
                elif f.get("rd_S") is not None:
                    out["rd"], out["rg"] = self._rows(f["rd_S"], own), self._rows(f["rg_S"], own)
            return out

    def _rebuild(self, st):
        """st: owned-order dict.  Migrate, plan the halo for the lists' whole life, sort, scan."""
        e = self.e
        self.nl = None
        self.fast = None
        recut = self.balance and self.world > 1 and self.geom is not None and "rows" in st and \
            (self._force_recut or self.stats["builds"] % self.REBALANCE_EVERY == 1)
        self._force_recut = False
        with e.section("host:geometry"):
            # bounding box, mean and max h, and how many particles left their slab: two all-reduces, one read
            lo, hi, hmean, ntot, hm, moved = self._global_stats(st)
            self.geom = E.choose_grid(lo, hi, hmean, ntot)          # refreshed at every rebuild (cells follow the gas)
            pp = self.margins.per_particle and "rd" in st
            if pp:
                # the life of the new lists and the bounds of their margins, agreed over all ranks
                self.margins.plan(st["rd"], st["rg"], self.stats["steps"], reduce=self._plan_reduce)
                cover_h, skin = self.margins.cover_h, self.margins.skin
            else:
                cover_h, skin = self.margins.choose()
                self.margins.life = 1 if self.margins.per_particle else 10 ** 9
            self.margins.built()
        with e.section("host:migrate"):
            if recut:
                st = self._rebalance_rows(st)
            elif moved:
                st = self._migrate(st)
        width = 2.0 * hm * (cover_h + skin) * (1.0 + 1e-6 + self.margin) + 4.0 * self.geom[1]
        with e.section("host:halo_plan"):
            hx = HaloExchanger(self.cuts, self.rank, self.world).plan(st["pos"][:, 0].contiguous(), width) if self.world > 1 \
                else None
        n_own = st["gid"].numel()
        with e.section("host:halo_exchange"):
            rd_L = rg_L = None
            if hx is not None:
                g = hx.exchange([st["gid"].to(F64), st["pos"], st["h"], st["vel"], st["u"]] + ([st["rd"], st["rg"]] if pp else []))
                gid_all = torch.cat([st["gid"], g[0].to(torch.int64)])
                pos_L, h_L = torch.cat([st["pos"], g[1]]), torch.cat([st["h"], g[2]])
                vel_L, u_L = torch.cat([st["vel"], g[3]]), torch.cat([st["u"], g[4]])
                if pp:
                    rd_L, rg_L = torch.cat([st["rd"], g[5]]), torch.cat([st["rg"], g[6]])
            else:
                gid_all, pos_L, h_L, vel_L, u_L = st["gid"], st["pos"], st["h"], st["vel"], st["u"]
                if pp:
                    rd_L, rg_L = st["rd"], st["rg"]
        dg = e.build(e.pack_posh(pos_L, h_L), *self.geom, tie_key=gid_all)
        perm = dg.perm.long()                                   # slot -> local index
        target = (perm < n_own).to(torch.uint8)
        rd_S = e.gather(rd_L, dg.perm, 1) if pp else None
        rg_S = e.gather(rg_L, dg.perm, 1) if pp else None
        self.nl = e.nl_build(dg, cover_h, skin, target=target if self.world > 1 else None, consts=self.consts,
                             margins=self.margins.margins_for(rd_S, rg_S) if pp else None)
        with e.section("host:slot_state"):
            slot_of_loc = torch.empty_like(perm)
            slot_of_loc[perm] = torch.arange(perm.numel(), device=perm.device)
            wf = st.get("wfac")
            if wf is None:
                wf = torch.ones_like(st["h"])
            wf_L = torch.cat([wf, torch.ones(perm.numel() - n_own, dtype=F64, device=self.dev)])
            self.fast = dict(P=dg.posh, V=e.gather(vel_L, dg.perm, 3), U=e.gather(u_L, dg.perm, 1),
                             gid_S=e.gather(gid_all.view(F64), dg.perm, 1).view(torch.int64), wfac_S=e.gather(wf_L, dg.perm, 1),
                             target=target if self.world > 1 else None,
                             own_sorted=torch.nonzero(target).flatten() if self.world > 1 else slot_of_loc,
                             n_own=n_own, sx=SlotExchanger(hx, slot_of_loc, n_own, engine=e) if hx is not None else None, fresh=True,
                             ghost_h_stale=False, rd_S=rd_S, rg_S=rg_S)
        self.stats["builds"] += 1
        self.stale = False
        # (a build that could not list its input leaves SPHB_NL_POOL / _UNLISTABLE in the lists' flag word: the step
        # runs anyway, and the all-reduced flags at its end send every rank down the same road)
        self._note_rows()

    def _plan_reduce(self, sums, maxes):
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            dist.all_reduce(maxes, op=dist.ReduceOp.MAX)

    def _note(self, nl, flags):
        """after a successful step: should the lists be rebuilt before the next one?"""
        if nl.margins is not None or self.margins.per_particle:
            return self.margins.note_used(nl.used_d, nl.used_g)
        return self.margins.note(nl.moved, nl.grown, flags & E.N.NL_AGING)

    def _failed(self):
        """a step ended with broken lists -> False: give up on lists for a while"""
        if self.nl.margins is not None:
            return self.margins.failed_pp()
        return self.margins.failed(self.fresh_failed, self.nl.moved, self.nl.grown)

    def _global_stats(self, st):
        pos, h = st["pos"], st["h"]
        own = HaloExchanger(self.cuts, self.rank, self.world).owner_of(pos[:, 0].contiguous()) if self.world > 1 else None
        left = (own != self.rank).sum().to(F64) if own is not None else torch.zeros((), dtype=F64, device=self.dev)
        vmax = torch.cat([-pos.min(dim=0).values, pos.max(dim=0).values, h.max().reshape(1)])
        vsum = torch.stack([h.sum(), torch.tensor(float(h.numel()), dtype=F64, device=self.dev), left])
        if self.world > 1:
            dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(vsum, op=dist.ReduceOp.SUM)
        v = torch.cat([vmax, vsum]).tolist()
        return [-v[0], -v[1], -v[2]], v[3:6], v[7] / v[8], int(v[8]), v[6], int(v[9])

    def _agree(self, flags4):
        """OR of the SPHB_NL_* flags and MAX of the two maxima over all ranks, in one all-reduce (bits as a MAX per bit;
        non-negative float bit patterns order like integers)"""
        if self.world == 1:
            return
        if not hasattr(self, "_bitpos"):
            self._bitpos = torch.arange(8, dtype=torch.int32, device=self.dev)
        v = torch.cat([(flags4[0:1] >> self._bitpos) & 1, flags4[1:3]])
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        flags4[0:1] = (v[:8] << self._bitpos).sum().to(torch.int32)
        flags4[1:3] = v[8:10]

    _force_recut = False

    def _rebalance_rows(self, st, nbins=16384):
        """work-balanced cut planes from what the drains actually walk: the list rows of each particle's warp (measured
        on the previous lists, travelling with the particle as `wfac`), one all-reduced weighted x-histogram"""
        x = st["pos"][:, 0].contiguous()
        wgt = (st["rows"] + 8.0) * st["wfac"]     # rows walked by the drains (+ the elementwise kernels), times the
        ext = torch.stack([x.min(), -x.max()])    # measured cost factor of the rank the particle was calibrated on
        dist.all_reduce(ext, op=dist.ReduceOp.MIN)
        xlo, xhi = float(ext[0].item()), float(-ext[1].item())
        span = max(xhi - xlo, 1e-300)
        b = ((x - xlo) * (nbins / span)).long().clamp_(0, nbins - 1)
        hist = torch.bincount(b, weights=wgt, minlength=nbins)
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        cum = torch.cumsum(hist, 0).cpu()
        total = float(cum[-1])
        cuts = [-math.inf]
        for r in range(1, self.world):
            k = int(torch.searchsorted(cum, torch.tensor(total * r / self.world, dtype=cum.dtype)).item())
            cuts.append(xlo + (min(k, nbins - 1) + 1) * span / nbins)
        cuts.append(math.inf)
        self.cuts = cuts
        return self._migrate(st)

    def _note_rows(self):
        """per-particle work for the next re-cut: the rows of the particle's warp"""
        f, nl = self.fast, self.nl
        f["rows_S"] = nl.warp_rows.to(F64).repeat_interleave(32)[:nl.n].contiguous()

    # ---- one step in slot order ------------------------------------------------------------------------------------
    def _slots_step(self, dt, stt):
        e, nl, f, c = self.e, self.nl, self.fast, self.consts
        P0, V0, U0, tgt, sx0 = f["P"], f["V"], f["U"], f["target"], f["sx"]

        class _R:   # the ghost refresh, visible to the per-call profile
            @staticmethod
            def refresh(arrs):
                with e.section("host:ghost_refresh"):
                    sx0.refresh(arrs)
        sx = _R if sx0 is not None else None
        if f["fresh"] and nl.omega0 is not None:
            om0 = nl.omega0                                                             # pyfluid.py:503, formed by the build
        else:
            _, _, om0, _ = e.nl_density_omega(nl, c, P0, target=tgt)
        if sx is not None:
            if f["ghost_h_stale"]:
                hcol = P0[:, 3].contiguous()
                sx.refresh([hcol, om0])                                                 # ghosts' h1 (last step) and Omega
                P0[:, 3] = hcol
            else:
                sx.refresh([om0])
        rB, rD = e.nl_eos(P0, V0, U0, om0, c, stt.ptr(E.S_STAGE0))
        acc0, dudt0 = e.nl_force(nl, c, P0, rB, None, stt.ptr(E.S_STAGE0), target=tgt, recD=rD)
        Pm, Vm, Um = e.predict(P0, V0, U0, acc0, dudt0, dt, stt.ptr(E.S_PRED))
        if sx is not None:
            sx.refresh([Pm, Vm, Um])
        e.nl_check(nl, Pm)
        hm, omm, Pm2, _ = e.nl_newton(nl, c, Pm, stt.ptr(E.S_NEWTON_MID), target=tgt, want_omega=True)
        if sx is not None:
            sx.refresh([hm, omm])
            Pm2[:, 3] = hm
        rB, rD = e.nl_eos(Pm2, Vm, Um, omm, c, stt.ptr(E.S_STAGE_MID))
        acc1, dudt1 = e.nl_force(nl, c, Pm2, rB, None, stt.ptr(E.S_STAGE_MID), target=tgt, recD=rD)
        P1, V1, U1 = e.correct(P0, V0, U0, acc1, dudt1, dt, stt.ptr(E.S_CORR))
        if sx is not None:
            sx.refresh([P1, V1, U1])
        e.nl_check(nl, P1)
        _, _, P1b, _ = e.nl_newton(nl, c, P1, stt.ptr(E.S_NEWTON_1), target=tgt, h_guess=hm)
        return P1b, V1, U1

    _BAD = E.N.NL_BROKEN | E.N.NL_POOL | E.N.NL_UNLISTABLE

    def _try(self, dt):
        """one step on the current lists -> (ok, P1, V1, U1, status)"""
        nl, f = self.nl, self.fast
        stt = E.Status(E.S_COUNT, self.dev)
        P1, V1, U1 = self._slots_step(dt, stt)
        self._agree(nl.flags)
        flags = self.last_flags = nl.read_flags()
        ok = (flags & self._BAD) == 0
        self.fresh_failed = f["fresh"] and not ok
        if ok and self._note(nl, flags):
            self.stale = True
        return ok, P1, V1, U1, stt

    MAX_ATTEMPTS = 4
    SAFE_HOLD = 8
    REBALANCE_EVERY = 16      # builds between two re-cuts of the slabs (the first one right after the calibration step)

    def _step(self, st, dt):
        """st: an owned-order dict, or None when the state is resident in slot order (self.fast)"""
        self.stats["steps"] += 1
        if self.hold > 0:
            self.hold -= 1
            self.stats["safe_steps"] += 1
            return SlabRunner._step(self, st, dt)
        for attempt in range(self.MAX_ATTEMPTS):
            if self.fast is None:
                self._rebuild(st)
            elif self.stale or attempt > 0:
                self._rebuild(self._to_owned())
            ok, P1, V1, U1, stt = self._try(dt)
            if ok:
                f = self.fast
                f["P"], f["V"], f["U"], f["fresh"], f["ghost_h_stale"] = P1, V1, U1, False, f["sx"] is not None
                self.statuses.append(stt)
                if f["sx"] is not None:
                    self.stats["exchanges"] += f["sx"].calls
                    f["sx"].calls = 0
                self._version += 1
                return _LazyOwned(self, self._version)
            self.stats["redos"] += 1
            if self.last_flags & (E.N.NL_UNLISTABLE | E.N.NL_POOL):
                break
            if not self._failed():
                break
        # the lists cannot serve this step: scanning path for a while
        st = self._to_owned()
        self.nl = None
        self.fast = None
        self.plan = None
        self.hold = self.SAFE_HOLD
        self.stats["steps"] -= 1
        return self._step(st, dt)

    _version = 0

    def step(self, st, dt):
        """st: an owned-order dict (a new state), or what the previous step() returned"""
        if isinstance(st, _LazyOwned):
            if st.runner is self and st.version == self._version and self.fast is not None:
                return self._step(None, dt)
            st = dict(st.resolve())
        self.fast = None
        self.nl = None
        return self._step(st, dt)

    def calibrate(self, st, dt, rounds=2):
        """setup-time load balancing.  Each round builds the lists, runs ONE throw-away step with CUDA events around every
        kernel, and folds this rank's time per modelled unit of work (rows walked) into the cost factors of its
        particles -- a sparse particle on a grid sized for the dense majority costs several times more to SCAN than its
        rows say --, then re-cuts the slabs.  Returns the state it was given (not advanced), redistributed."""
        if not self.balance or self.world == 1:
            return st
        if isinstance(st, _LazyOwned):
            st = dict(st.resolve())
        for _ in range(rounds):
            self.fast, self.nl = None, None
            self.e.prof_start()                        # the build counts: it is where sparse particles cost most
            self._rebuild(st)
            pre = self._to_owned()                     # the same physical state, in this build's owned order, with rows
            self._try(dt)                              # result discarded
            prof = self.e.prof_stop()
            ms = sum(t for nm, (c, t) in prof.items() if not nm.startswith("host:"))
            work = float(((pre["rows"] + 8.0) * pre["wfac"]).sum().item())
            v = torch.tensor([ms, work], dtype=F64, device=self.dev)
            allv = [torch.empty_like(v) for _ in range(self.world)]
            dist.all_gather(allv, v)
            tt, ww = sum(float(a[0]) for a in allv), sum(float(a[1]) for a in allv)
            fac = (ms / max(work, 1e-300)) / (tt / max(ww, 1e-300))
            pre["wfac"] = pre["wfac"] * min(max(fac, 0.5), 2.0)
            st = pre
            self._force_recut = True
        self.fast, self.nl = None, None
        self.stale = True
        self.margins = E.ListMargins(*self._margins0)     # the real run starts like a single-GPU run does
        self.stats = dict(steps=0, builds=0, redos=0, safe_steps=0, exchanges=0)
        self._force_recut = True
        return st


class _LazyOwned(dict):
    """The owned-order view of a FastSlabRunner's state, materialised only when somebody looks at it (bench checksums,
    gather_state, tests): between rebuilds the state lives in the lists' slot order."""

    def __init__(self, runner, version):
        super().__init__()
        self.runner, self.version = runner, version
        self._done = False

    def resolve(self):
        if not self._done:
            if self.version != self.runner._version or self.runner.fast is None:
                raise RuntimeError("this state has been stepped past; keep what step() returned last")
            super().update(self.runner._to_owned())
            self._done = True
        return self

    def __getitem__(self, k):
        return dict.__getitem__(self.resolve(), k)

    def get(self, k, d=None):
        return dict.get(self.resolve(), k, d)

    def items(self):
        return dict.items(self.resolve())

    def keys(self):
        return dict.keys(self.resolve())

    def __iter__(self):
        return dict.__iter__(self.resolve())



def make_runner(eng, consts, world, rank, lists=True, decomp="slab", **kw):
    """the runner bench.py and the tests use.  decomp "slab": cut planes along x, kept neighbour lists when `lists` (and
    G = 0), else the scanning kernels; "range": ranges of sorted slots on one global cell list (ranges.RangeRunner),
    for distributions whose h spans decades"""
    if decomp == "range":
        import ranges
        return ranges.RangeRunner(eng, consts, world, rank, **kw)
    if lists and consts.G == 0.0:
        return FastSlabRunner(eng, consts, world, rank, **kw)
    return SlabRunner(eng, consts, world, rank, **kw)
