"""Decomposition by ranges of sorted slots, for distributions no plane can cut (BASELINE config 5, the Plummer sphere).

A slab's ghost layer is 2 h_max wide, and in a Plummer sphere h_max belongs to the sparse rim: every slab through the
core would need the whole core as ghosts.  Here every rank sorts the SAME global cell list (the nested grids of
engine.Levels; 64 M records are 2 GB of the 180 GB) and owns a contiguous range of its sorted slots -- level-major, then
z, y, x: a space-filling order with compact pieces, cut where the cumulative work (list rows walked per particle, measured
on the previous lists) is equal.  The rank lists only its own targets; the slots its rows name ARE its ghost set
(sphb_nlist_mark), exact, with no geometric halo width at all.  Own range and ghosts are then renamed to a compact local
numbering that is increasing in the global slot (sphb_nlist_remap), so every sum still runs in ascending global slot
order: the fields equal the single-GPU FastStepper's bit for bit.  Between rebuilds a step is FastSlabRunner's: five
drains and four ghost refreshes along fixed send lists (one all_to_all_single each, NCCL over NVLink).

At a rebuild the owned records travel to every rank (one all-gather of 88 bytes per particle); the new owner of a slot
simply keeps what it needs: migration and re-balancing need no extra exchange.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

import domain
import engine as E

F64 = torch.float64
COLS = ("gid", "pos", "h", "vel", "u", "rows", "wfac", "rd", "rg")   # what travels at a rebuild: 13 doubles per particle


def _geom_to_plain(g):
    grid, delta = g
    f = lambda q: (q.ox, q.oy, q.oz, q.cell, q.nx, q.ny, q.nz, q.xdiv)
    if isinstance(grid, E.Levels):
        return ("levels", f(grid.frame), grid.exp0, [f(q) for q in grid.grids], delta)
    return ("grid", f(grid), delta)


def _geom_from_plain(p):
    if p[0] == "levels":
        return E.Levels(E.N.Grid(*p[1]), p[2], [E.N.Grid(*q) for q in p[3]]), p[4]
    return E.N.Grid(*p[1]), p[2]


def weighted_cuts(cum, world, align=32):
    """slot ranges of equal cumulative weight: cum = inclusive prefix sums of the per-slot weights (1-D tensor) ->
    world + 1 boundaries, multiples of `align` (a warp of 32 slots is the unit of list work), first 0, last n"""
    n = cum.numel()
    total = float(cum[-1])
    want = torch.tensor([total * r / world for r in range(1, world)], dtype=cum.dtype, device=cum.device)
    k = torch.searchsorted(cum, want).tolist() if world > 1 else []
    cuts = [0]
    for v in k:
        v = min(n, (int(v) + align // 2) // align * align)
        cuts.append(max(v, cuts[-1]))
    cuts.append(n)
    return cuts


def ghost_plan(ghost_slots, cuts, rank, world, group=None):
    """ghost_slots: ascending global slots this rank reads but does not own.  Every owner learns what it has to send.
    -> (requests received: global slots of MY range, grouped by requesting rank), send_counts, recv_counts
    Pure torch plumbing (any device, any backend): what the gloo test runs."""
    inner = torch.tensor(cuts[1:-1], dtype=torch.int64, device=ghost_slots.device)
    owner = torch.bucketize(ghost_slots, inner, right=True)
    want = torch.bincount(owner, minlength=world)
    asked = torch.empty_like(want)
    dist.all_to_all_single(asked, want, group=group)
    both = torch.stack([want, asked]).tolist()
    recv_counts, send_counts = both[0], both[1]          # I receive what I want; I send what I am asked for
    req = torch.empty(int(sum(send_counts)), dtype=torch.int64, device=ghost_slots.device)
    dist.all_to_all_single(req, ghost_slots.contiguous(), output_split_sizes=send_counts, input_split_sizes=recv_counts, group=group)
    return req, send_counts, recv_counts


class _Plan:
    """what domain.SlotExchanger needs to know about the exchange"""

    def __init__(self, world, send_counts, recv_counts, group=None):
        self.world, self.send_counts, self.recv_counts, self.group = world, send_counts, recv_counts, group

    @property
    def n_recv(self):
        return int(sum(self.recv_counts))


def slot_exchanger(plan, send_local, ghost_local, engine=None):
    sx = domain.SlotExchanger.__new__(domain.SlotExchanger)
    sx.hx, sx.engine = plan, engine
    sx.send_slots, sx.ghost_slots = send_local.contiguous(), ghost_local.contiguous()
    sx.n_recv, sx.calls = plan.n_recv, 0
    return sx


class RangeRunner(domain.FastSlabRunner):
    """var_smoothlength_leapfrog (pyfluid.py:498-532, G = 0) on kept neighbour lists, decomposed by slot ranges."""

    def __init__(self, eng: E.Engine, consts, world, rank, cover_h=None, skin=None, adaptive=True):
        super().__init__(eng, consts, world, rank, margin=0.0, balance=True, cover_h=cover_h, skin=skin, adaptive=adaptive)
        self.cuts = None
        self.safe = E.Stepper(eng)

    # ---- setup ------------------------------------------------------------------------------------------------------
    def scatter_initial(self, w):
        """every rank generates the same synthetic workload and keeps one block of caller indices"""
        import numpy as np
        n = len(w["pos"])
        self.n_total = n
        a, b = n * self.rank // self.world, n * (self.rank + 1) // self.world
        to = lambda x: torch.from_numpy(np.ascontiguousarray(x[a:b])).to(self.dev)
        one = torch.ones(b - a, dtype=F64, device=self.dev)
        return dict(gid=torch.arange(a, b, dtype=torch.int64, device=self.dev), pos=to(w["pos"]), vel=to(w["vel"]), u=to(w["u"]),
                    h=to(w["h_guess"]), wfac=one, rows=one.clone())

    # ---- the global sort every rank makes ------------------------------------------------------------------------------
    def _allgather(self, st):
        """-> dict of the same fields for ALL particles, ranks in order (the same on every rank)"""
        names = [c for c in COLS if c in st]
        cols = [st[c].view(F64).reshape(st[c].shape[0], -1) if st[c].dtype == torch.int64 else st[c].reshape(st[c].shape[0], -1)
                for c in names]
        packed = torch.cat(cols, dim=1).contiguous()
        if self.world == 1:
            full = packed
        else:
            n = torch.tensor([packed.shape[0]], dtype=torch.int64, device=self.dev)
            ns = [torch.empty_like(n) for _ in range(self.world)]
            dist.all_gather(ns, n)
            ns = [int(t.item()) for t in ns]
            full = torch.empty((sum(ns), packed.shape[1]), dtype=F64, device=self.dev)
            dist.all_gather(list(full.split(ns)), packed)
        out, o = {}, 0
        for c, col in zip(names, cols):
            wd = col.shape[1]
            v = full[:, o:o + wd]
            out[c] = v.reshape(-1).contiguous() if st[c].dim() == 1 else v.contiguous()
            if st[c].dtype == torch.int64:
                out[c] = out[c].view(torch.int64)
            o += wd
        return out

    def _sorted_world(self, st):
        """all records, one geometry (rank 0 draws it, everyone uses it), one cell list: identical on every rank"""
        e = self.e
        with e.section("host:allgather"):
            full = self._allgather(st)
        posh = e.pack_posh(full["pos"], full["h"])
        with e.section("host:geometry"):
            box = [_geom_to_plain(e.grid_for(posh)) if self.rank == 0 else None]
            if self.world > 1:
                dist.broadcast_object_list(box, src=0)
            self.geom = _geom_from_plain(box[0])
        dg = e.build(posh, *self.geom, tie_key=full["gid"])
        return full, dg

    def _cut(self, weight_S):
        """slot ranges of equal weight, agreed by broadcast"""
        box = [weighted_cuts(torch.cumsum(weight_S, 0), self.world) if self.rank == 0 else None]
        if self.world > 1:
            dist.broadcast_object_list(box, src=0)
        self.cuts = box[0]
        return self.cuts[self.rank], self.cuts[self.rank + 1]

    def cold_solve(self, st):
        """newton_smoothlength_arr from the workload's guess (scanning kernel, bisection fallback), each rank its range"""
        e = self.e
        full, dg = self._sorted_world(st)
        n = dg.n
        s0, s1 = self._cut(torch.ones(n, dtype=F64, device=self.dev))
        mask = torch.zeros(n, dtype=torch.uint8, device=self.dev)
        mask[s0:s1] = 1
        stt = E.Status(1, self.dev)
        h_S = e.newton(dg, self.consts, dg.posh[:, 3].contiguous(), stt.ptr(0), target=mask)
        e.check_neg_h(h_S[s0:s1].contiguous(), stt.ptr(0))
        self.statuses.append(stt)
        src = dg.perm[s0:s1].long()
        return dict(gid=full["gid"].index_select(0, src), pos=dg.posh[s0:s1, :3].contiguous(), h=h_S[s0:s1].contiguous(),
                    vel=full["vel"].index_select(0, src), u=full["u"].index_select(0, src),
                    wfac=full["wfac"].index_select(0, src), rows=full["rows"].index_select(0, src))

    def calibrate(self, st, dt, rounds=1):
        """setup: one build with equal counts measures the rows every particle's warp walks; the next build cuts by them"""
        if self.world == 1:
            return st
        if isinstance(st, domain._LazyOwned):
            st = dict(st.resolve())
        for _ in range(rounds):
            self.fast, self.nl = None, None
            self._rebuild(st)
            st = self._to_owned()
        self.fast, self.nl = None, None
        self.stale = True
        self.margins = E.ListMargins(*self._margins0)
        self.stats = dict(steps=0, builds=0, redos=0, safe_steps=0, exchanges=0)
        return st

    # ---- rebuild --------------------------------------------------------------------------------------------------------
    def _rebuild(self, st):
        e = self.e
        self.nl = None
        self.fast = None
        if "rows" not in st:
            st = dict(st, rows=torch.ones_like(st["h"]))
        if "wfac" not in st:
            st = dict(st, wfac=torch.ones_like(st["h"]))
        full, dg = self._sorted_world(st)
        n = dg.n
        perm = dg.perm.long()
        pp = self.margins.per_particle and "rd" in full
        if pp:
            # every rank holds every record's rates: the same plan everywhere, no reduction needed
            m_all = self.margins.plan(full["rd"], full["rg"], self.stats["steps"])
            cover_h, skin = self.margins.cover_h, self.margins.skin
            m_S = m_all.index_select(0, perm)
        else:
            cover_h, skin = self.margins.choose()
            self.margins.life = 1 if self.margins.per_particle else 10 ** 9
            m_S = None
        self.margins.built()
        with e.section("host:cuts"):
            w_S = (full["rows"] + 8.0) * full["wfac"]
            s0, s1 = self._cut(w_S.index_select(0, perm))
        n_own = s1 - s0
        mask = torch.zeros(n, dtype=torch.uint8, device=self.dev)
        mask[s0:s1] = 1
        # lists of the own targets only, in global slot names
        pool_rows = self.pool_rows_hint if getattr(self, "pool_rows_hint", None) else ((n_own + 31) // 32) * 176 + 65536
        nl_g = e.nl_build(dg, cover_h, skin, target=mask if self.world > 1 else None, consts=self.consts, pool_rows=pool_rows,
                          margins=m_S)
        self.pool_rows_hint = max(nl_g.pool_rows, ((n_own + 31) // 32) * 176 + 65536)
        with e.section("host:local_numbering"):
            mark = e.nl_mark(nl_g, n)
            mark[s0:s1] = 1
            loc = torch.nonzero(mark).flatten()                      # ascending global slots: ghosts below, own, ghosts above
            n_low = int((loc < s0).sum().item())
            pad = (-n_low) % 32                                       # own slots start on a warp boundary
            n_loc = pad + loc.numel()
            if pad:
                loc = torch.cat([loc[:1].expand(pad), loc])           # dummy records: never named, never targets
            slot_map = torch.full((n,), -1, dtype=torch.int32, device=self.dev)
            slot_map[loc[pad:]] = torch.arange(pad, n_loc, dtype=torch.int32, device=self.dev)
            e.nl_remap(nl_g, slot_map)
            o0 = pad + n_low                                           # local index of the first own slot
            nw = (n_loc + 31) // 32
            w_off = torch.zeros(nw, dtype=torch.int32, device=self.dev)
            w_rows = torch.zeros(nw, dtype=torch.int32, device=self.dev)
            nwo = (n_own + 31) // 32
            w_off[o0 // 32:o0 // 32 + nwo] = nl_g.warp_off[s0 // 32:s0 // 32 + nwo]
            w_rows[o0 // 32:o0 // 32 + nwo] = nl_g.warp_rows[s0 // 32:s0 // 32 + nwo]
            P = dg.posh.index_select(0, loc)
            self.nl = E.KeptLists.view(nl_g, n_loc, w_off, w_rows, P, margins=m_S.index_select(0, loc) if m_S is not None else None)
            if nl_g.omega0 is not None:
                om = torch.ones(n_loc, dtype=F64, device=self.dev)
                om[o0:o0 + n_own] = nl_g.omega0[s0:s1]
                self.nl.omega0 = om
            src = perm.index_select(0, loc)                            # where the local records sit in the gathered arrays
            V, U = full["vel"].index_select(0, src), full["u"].index_select(0, src)
            target = torch.zeros(n_loc, dtype=torch.uint8, device=self.dev)
            target[o0:o0 + n_own] = 1
        with e.section("host:ghost_plan"):
            sx = None
            if self.world > 1:
                # the dummies are refreshed like the ghost they copy: every non-target record is then current at every stage
                ghosts = torch.cat([loc[:o0], loc[o0 + n_own:]])
                req, send_counts, recv_counts = ghost_plan(ghosts, self.cuts, self.rank, self.world)
                ghost_local = torch.cat([torch.arange(0, o0, device=self.dev), torch.arange(o0 + n_own, n_loc, device=self.dev)])
                sx = slot_exchanger(_Plan(self.world, send_counts, recv_counts), req - s0 + o0, ghost_local, engine=e)
        self.fast = dict(P=P, V=V, U=U, gid_S=full["gid"].index_select(0, src), wfac_S=full["wfac"].index_select(0, src),
                         target=target if self.world > 1 else None,
                         own_sorted=torch.arange(o0, o0 + n_own, device=self.dev), n_own=n_own, sx=sx, fresh=True,
                         ghost_h_stale=False, range=(s0, s1), n_ghost=n_loc - pad - n_own,
                         rd_S=full["rd"].index_select(0, src) if pp else None, rg_S=full["rg"].index_select(0, src) if pp else None)
        self.stats["builds"] += 1
        self.stale = False
        self._note_rows()

    # ---- what the lists cannot serve: every rank replays the step on the whole set (scanning kernels) --------------------
    def _replicated_step(self, st, dt):
        full = self._allgather({k: st[k] for k in ("gid", "pos", "h", "vel", "u")})
        order = torch.argsort(full["gid"])
        pos, vel, u, h = (full[k].index_select(0, order) for k in ("pos", "vel", "u", "h"))
        p1, v1, u1, h1 = self.safe.step(pos, vel, u, h, dt, self.consts, defer_status=True)
        self.statuses.append(self.safe.last_status)
        gid = st["gid"]
        return dict(gid=gid, pos=p1.index_select(0, gid), h=h1.index_select(0, gid), vel=v1.index_select(0, gid),
                    u=u1.index_select(0, gid), wfac=st.get("wfac", torch.ones_like(st["h"])),
                    rows=st.get("rows", torch.ones_like(st["h"])))

    def _step(self, st, dt):
        self.stats["steps"] += 1
        if self.hold > 0:
            self.hold -= 1
            self.stats["safe_steps"] += 1
            return self._replicated_step(st, dt)
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
                return domain._LazyOwned(self, self._version)
            self.stats["redos"] += 1
            if self.last_flags & (E.N.NL_UNLISTABLE | E.N.NL_POOL):
                break
            if not self._failed():
                break
        st = self._to_owned()
        self.nl = None
        self.fast = None
        self.hold = self.SAFE_HOLD
        self.stats["steps"] -= 1
        return self._step(st, dt)
