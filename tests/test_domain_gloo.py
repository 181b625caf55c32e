"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: cut planes, halo send lists, the all_to_all
exchange and particle migration (python-sph_b200/domain.py).  No kernels run here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import domain
        rng = np.random.default_rng(0)
        n = 4000
        pos = rng.random((n, 3)) * np.array([10.0, 2.0, 2.0])
        val = rng.random(n)
        xs = torch.from_numpy(np.sort(pos[:, 0]))
        cuts = domain.equal_count_cuts(xs, world)
        own = np.nonzero((pos[:, 0] >= cuts[rank]) & (pos[:, 0] < cuts[rank + 1]))[0]
        gid = torch.from_numpy(own.astype(np.int64))
        p = torch.from_numpy(pos[own])
        v = torch.from_numpy(val[own])
        w = 0.7
        hx = domain.HaloExchanger(cuts, rank, world).plan(p[:, 0], w)
        g_gid, g_pos, g_val = hx.exchange([gid.to(torch.float64), p, v])
        g_gid = g_gid.to(torch.int64).numpy()
        # every particle of another rank within w of my slab arrived, exactly once, with the right payload
        lo, hi = cuts[rank], cuts[rank + 1]
        others = np.setdiff1d(np.arange(n), own)
        need = others[(pos[others, 0] >= lo - w) & (pos[others, 0] < hi + w)]
        ok = (np.sort(g_gid) == np.sort(need)).all() and (g_pos.numpy() == pos[g_gid]).all() and (g_val.numpy() == val[g_gid]).all()
        # second exchange with the same plan returns values in the same order
        g2 = hx.exchange([v * 2.0])[0]
        ok = ok and (g2.numpy() == 2.0 * val[g_gid]).all()
        # migration conserves the particle set
        p_moved = p.clone()
        p_moved[:, 0] += 1.5
        mg = domain.HaloExchanger(cuts, rank, world).plan_migration(p_moved[:, 0])
        r_gid, r_pos = mg.exchange([gid.to(torch.float64), p_moved])
        mine = torch.cat([gid[mg.stay], r_gid.to(torch.int64)])
        mine_x = torch.cat([p_moved[mg.stay, 0], r_pos[:, 0]])
        ok = ok and bool(((mine_x >= lo) & (mine_x < hi)).all())
        cnt = torch.tensor([mine.numel()])
        dist.all_reduce(cnt)
        ok = ok and int(cnt.item()) == n
        q.put((rank, bool(ok), int(len(need))))
    finally:
        dist.destroy_process_group()


def _worker_slots(rank, world, port, q):
    """SlotExchanger: ghost refresh in a frozen (shuffled) slot order"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import domain
        rng = np.random.default_rng(1)
        n = 3000
        pos = rng.random((n, 3)) * np.array([10.0, 2.0, 2.0])
        xs = torch.from_numpy(np.sort(pos[:, 0]))
        cuts = domain.equal_count_cuts(xs, world)
        own = np.nonzero((pos[:, 0] >= cuts[rank]) & (pos[:, 0] < cuts[rank + 1]))[0]
        gid = torch.from_numpy(own.astype(np.int64))
        p = torch.from_numpy(pos[own])
        hx = domain.HaloExchanger(cuts, rank, world).plan(p[:, 0], 0.9)
        g_gid = hx.exchange([gid.to(torch.float64)])[0].to(torch.int64)
        gid_all = torch.cat([gid, g_gid])
        n_own, n_loc = gid.numel(), gid_all.numel()
        perm = torch.from_numpy(np.random.default_rng(5 + rank).permutation(n_loc))      # slot -> local index
        slot_of_loc = torch.empty_like(perm)
        slot_of_loc[perm] = torch.arange(n_loc)
        sx = domain.SlotExchanger(hx, slot_of_loc, n_own)
        gid_S = gid_all[perm]
        owned_S = perm < n_own
        ok = True
        for round_ in range(3):
            # "truth" values are functions of (global id, round); owners hold them, ghosts hold garbage before the refresh
            f = lambda g, c: (g.to(torch.float64) + 1.0) * (c + 1) + 0.25 * round_
            a = torch.where(owned_S, f(gid_S, 0), torch.full((n_loc,), -1.0, dtype=torch.float64))
            b = torch.stack([torch.where(owned_S, f(gid_S, c), torch.full((n_loc,), -1.0, dtype=torch.float64)) for c in (1, 2, 3)], 1)
            sx.refresh([a, b])
            ok = ok and bool((a == f(gid_S, 0)).all()) and all(bool((b[:, c - 1] == f(gid_S, c)).all()) for c in (1, 2, 3))
        ok = ok and sx.calls == 3
        q.put((rank, bool(ok), int(n_loc - n_own)))
    finally:
        dist.destroy_process_group()


def test_slot_exchanger_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_slots, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(ng > 0 for _, _, ng in res)


def test_halo_and_migration_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(nneed > 0 for _, _, nneed in res)


def test_equal_count_cuts():
    import domain
    x = torch.sort(torch.rand(1000, dtype=torch.float64)).values
    c = domain.equal_count_cuts(x, 4)
    assert len(c) == 5 and c[0] == -np.inf and c[-1] == np.inf
    counts = [int(((x >= c[r]) & (x < c[r + 1])).sum()) for r in range(4)]
    assert counts == [250] * 4


def _worker_ranges(rank, world, port, q):
    """ranges.py: weighted slot cuts, the ghost request exchange, refresh in the compact local numbering"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ranges
        n = 5000
        rng = np.random.default_rng(2)
        wgt = torch.from_numpy(1.0 + 9.0 * (np.arange(n) < 1500) + rng.random(n))       # the first slots are heavy
        cuts = ranges.weighted_cuts(torch.cumsum(wgt, 0), world)
        s0, s1 = cuts[rank], cuts[rank + 1]
        ok = cuts[0] == 0 and cuts[-1] == n and all(c % 32 == 0 for c in cuts[:-1])
        part = [float(wgt[cuts[r]:cuts[r + 1]].sum()) for r in range(world)]
        ok = ok and max(part) < 1.05 * min(part)
        # a ghost set: random slots of the other ranges, ascending, some low ones repeated like the padding dummies
        others = np.setdiff1d(np.arange(n), np.arange(s0, s1))
        g = np.sort(np.random.default_rng(10 + rank).choice(others, size=400, replace=False))
        n_low = int((g < s0).sum())
        pad = (-n_low) % 32
        ghosts = torch.from_numpy(np.concatenate([np.repeat(g[:1], pad) if n_low else np.empty(0, np.int64), g]).astype(np.int64))
        if not n_low:
            pad = 0
        req, send_counts, recv_counts = ranges.ghost_plan(ghosts, cuts, rank, world)
        ok = ok and bool(((req >= s0) & (req < s1)).all()) and sum(recv_counts) == ghosts.numel()
        o0, n_own = pad + n_low, s1 - s0
        n_loc = o0 + n_own + (len(g) - n_low)
        ghost_local = torch.cat([torch.arange(0, o0), torch.arange(o0 + n_own, n_loc)])
        sx = ranges.slot_exchanger(ranges._Plan(world, send_counts, recv_counts), req - s0 + o0, ghost_local)
        glob = torch.cat([ghosts[:o0], torch.arange(s0, s1), ghosts[o0:]])             # global slot of every local record
        own_l = torch.zeros(n_loc, dtype=torch.bool)
        own_l[o0:o0 + n_own] = True
        f = lambda s, c: (s.to(torch.float64) + 1.0) * (c + 1)
        a = torch.where(own_l, f(glob, 0), torch.full((n_loc,), -1.0, dtype=torch.float64))
        b = torch.stack([torch.where(own_l, f(glob, c), torch.full((n_loc,), -1.0, dtype=torch.float64)) for c in (1, 2)], 1)
        sx.refresh([a, b])
        ok = ok and bool((a == f(glob, 0)).all()) and bool((b[:, 0] == f(glob, 1)).all()) and bool((b[:, 1] == f(glob, 2)).all())
        q.put((rank, bool(ok), int(ghosts.numel())))
    finally:
        dist.destroy_process_group()


def test_range_decomposition_plumbing_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ranges, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(ng > 0 for _, _, ng in res)
