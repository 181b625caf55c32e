"""numpy restatement of the cell assignment / stable sort of K1 (SURVEY.md 8(c) "self-oracle":
the reference has no cell list, so this file, not pyfluid.py, defines the expected values)."""
import numpy as np


def cell_ids(pos, origin, cell, dims, xdiv=1):
    edges = np.array([cell / max(int(xdiv), 1), cell, cell])
    c = np.floor((pos - np.asarray(origin)[None, :]) / edges[None, :])
    c = np.clip(c, 0, np.asarray(dims)[None, :] - 1).astype(np.int64)
    return (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0]


def build(pos, origin, cell, dims, xdiv=1):
    ids = cell_ids(pos, origin, cell, dims, xdiv)
    perm = np.argsort(ids, kind="stable")
    ncell = int(dims[0]) * int(dims[1]) * int(dims[2])
    counts = np.bincount(ids, minlength=ncell)
    start = np.concatenate([[0], np.cumsum(counts)])
    return ids, perm, start


def level_of(h, exp0, nlevels):
    """nested grids (include/sphb200.h sphb_level): level = clamp(floor(log2 h) - exp0, 0, nlevels - 1); h <= 0 or
    non-finite -> the last level"""
    h = np.asarray(h, dtype=np.float64)
    ok = (h > 0) & np.isfinite(h)
    e = np.frexp(np.where(ok, h, 1.0))[1] - 1
    lv = np.clip(e - exp0, 0, nlevels - 1)
    return np.where(ok, lv, nlevels - 1).astype(np.int64)


def build_levels(pos, h, exp0, grids):
    """grids: list of (origin[3], cell, dims[3]) per level -> global cell ids, stable permutation, cell_start"""
    lv = level_of(h, exp0, len(grids))
    ids = np.zeros(len(pos), dtype=np.int64)
    base = 0
    for b, (origin, cell, dims) in enumerate(grids):
        m = lv == b
        if m.any():
            ids[m] = base + cell_ids(pos[m], origin, cell, dims)
        base += int(dims[0]) * int(dims[1]) * int(dims[2])
    perm = np.argsort(ids, kind="stable")
    counts = np.bincount(ids, minlength=base)
    start = np.concatenate([[0], np.cumsum(counts)])
    return ids, perm, start
