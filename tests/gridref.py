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
