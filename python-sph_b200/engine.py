"""Device-side driver: torch tensors carry the HBM buffers, libsphb200 does all the arithmetic.

Layout in HBM (all fp64 unless noted), in "grid order" (particles sorted by cell, x fastest):
    posh  (n,4)  x, y, z, h           the exact record every pair evaluation reads
    frec  (n,4)  fp32 x-ox, y-oy, z-oz, (2h+delta)^2   prefilter record
    cell_start (ncell+1) u32, cell_hmax (ncell) f32, cell_id (n) u32, perm (n) u32
    vel (n,3), u (n), recB (n,4) = vx,vy,vz,A, recC (n,4) = rho,cs,1/h,1/(pi h^4)

The step driver mirrors var_smoothlength_leapfrog (pyfluid.py:498-532): three cell-list builds
(pos0, pos_mid, pos1), one Omega pass, two Newton solves (+ an Omega pass at the new h), two force
passes, predictor, corrector.  Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

import _native as N

F64 = torch.float64


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class ParkedLists:
    """sphb_lists: one producer kernel's neighbour lists, kept for the force kernel of the same snapshot"""

    def __init__(self, store, broken, hmax, cover):
        self.store, self.broken, self.hmax = store, broken, hmax   # keep the tensors alive
        self.c = N.Lists(store.data_ptr(), broken.data_ptr(), hmax.data_ptr(), float(cover))


def _lists(pl):
    return C.byref(pl.c) if pl is not None else None


class KeptLists:
    """sphb_nlist: neighbour lists that outlive kernels and steps, with the buffers they point to.
    Built from ONE cell list (dg); the drain kernels then only need the current sorted records."""

    def __init__(self, engine, dg, cover_h, skin, pool_rows=None, margins=None, level_dmax=None, cell_dmax=None):
        """margins (n,2) float32 = per-particle {cover, skin} in slot order (cover_h / skin are then their upper bounds),
        level_dmax: float32 upper bounds of skin_i h_i per level of the cell list"""
        lib, dev = engine.lib, engine.device
        self.margins, self.level_dmax, self.cell_dmax = margins, level_dmax, cell_dmax
        self.posh_build = dg.posh
        n = dg.n
        nw = max(1, (n + 31) // 32)
        self.n = n
        self.dg = dg                      # keeps posh_build (dg.posh) and perm alive
        self.cover_h, self.skin = float(cover_h), float(skin)
        rows = int(pool_rows if pool_rows is not None else lib.sphb_nlist_pool_rows(n))
        self.pool_rows = rows
        self.warp_off = torch.empty(nw, dtype=torch.int32, device=dev)
        self.warp_rows = torch.zeros(nw, dtype=torch.int32, device=dev)
        self.pool = torch.empty((rows, 32), dtype=torch.int32, device=dev)
        self.cursor = torch.zeros(1, dtype=torch.int64, device=dev)
        self.flags = torch.zeros(4, dtype=torch.int32, device=dev)   # [0] SPHB_NL_* flags, [1..2] the maxima
        self.c = N.NList(n, self.warp_off.data_ptr(), self.warp_rows.data_ptr(), self.pool.data_ptr(), rows,
                         self.cursor.data_ptr(), self.flags.data_ptr(), dg.posh.data_ptr(), self.cover_h, self.skin,
                         self.flags.data_ptr() + 4, margins.data_ptr() if margins is not None else None,
                         level_dmax.data_ptr() if level_dmax is not None else None,
                         cell_dmax.data_ptr() if cell_dmax is not None else None)
        self.build_flags = 0
        self.moved = self.grown = 0.0
        self.used_d = self.used_g = 0.0

    @classmethod
    def view(cls, src, n, warp_off, warp_rows, posh_build, margins=None):
        """the lists of `src` seen through renamed slots (sphb_nlist_remap): same pool, cursor and flag words, other
        warp tables, build-time records and per-particle margins"""
        nl = cls.__new__(cls)
        nl.margins, nl.level_dmax, nl.cell_dmax = margins, src.level_dmax, None   # (the builder alone reads cell_dmax)
        nl.posh_build = posh_build
        nl.n, nl.dg, nl.cover_h, nl.skin = int(n), src.dg, src.cover_h, src.skin
        nl.pool_rows, nl.warp_off, nl.warp_rows, nl.pool, nl.cursor, nl.flags = src.pool_rows, warp_off, warp_rows, src.pool, src.cursor, src.flags
        nl.c = N.NList(nl.n, warp_off.data_ptr(), warp_rows.data_ptr(), nl.pool.data_ptr(), nl.pool_rows, nl.cursor.data_ptr(),
                       nl.flags.data_ptr(), posh_build.data_ptr(), nl.cover_h, nl.skin, nl.flags.data_ptr() + 4,
                       margins.data_ptr() if margins is not None else None,
                       src.level_dmax.data_ptr() if (margins is not None and src.level_dmax is not None) else None, None)
        nl.build_flags, nl.moved, nl.grown, nl.omega0 = src.build_flags, 0.0, 0.0, None
        nl.used_d = nl.used_g = 0.0
        return nl

    def ref(self):
        return C.byref(self.c)

    def read_flags(self):
        """one 16-byte D2H read (synchronises the stream): the flags; also refreshes .moved / .grown, the largest
        displacement / h_build and h growth the checks have seen since the build"""
        v = self.flags.cpu()
        mx = v[1:3].view(torch.float32)
        # the device reports the largest FRACTIONS of their allowances the records have used; with uniform margins
        # .moved / .grown are those in absolute terms (displacement / h_build, h growth)
        self.used_d, self.used_g = float(mx[0]), float(mx[1])
        self.moved = self.used_d * self.skin if self.used_d > 0.0 else 0.0
        self.grown = self.used_g * (self.cover_h - 1.0) if self.used_g > 0.0 else 0.0
        return int(v[0])

    def rows_used(self):
        return int(self.cursor.item())


def require_cuda():
    lib = N.load()
    if not torch.cuda.is_available():
        raise N.NativeError("sph-b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return lib


class Status:
    """A device array of sphb_status blocks (one per stage), read back in one copy."""

    def __init__(self, nblocks, device):
        self.t = torch.zeros((nblocks, N.STATUS_INTS), dtype=torch.int32, device=device)

    def ptr(self, i):
        return C.c_void_p(self.t.data_ptr() + i * N.STATUS_INTS * 4)

    def read(self):
        return self.t.cpu().numpy()


class Levels:
    """Nested cell grids (include/sphb200.h sphb_level): particles binned by the octave of h, one grid per octave,
    cells numbered level-major.  frame: a Grid whose origin is the origin of the fp32 prefilter frame."""

    def __init__(self, frame, exp0, grids):
        self.frame = frame
        self.exp0 = int(exp0)
        self.grids = list(grids)
        self.bases = []
        tot = 0
        for g in self.grids:
            self.bases.append(tot)
            tot += g.ncell
        self.ncell = tot

    @property
    def nlevels(self):
        return len(self.grids)

    def table(self):
        """the device table as bytes (sphb_level[nlevels])"""
        arr = (N.Level * self.nlevels)()
        for b, g in enumerate(self.grids):
            arr[b].grid = g
            arr[b].grid.xdiv = 1
            arr[b].cell_base = self.bases[b]
            arr[b].ncell = g.ncell
        return bytes(arr)


class DeviceGrid:
    """One cell list over one snapshot of positions (K1 output)."""

    def __init__(self, n, grid, delta, device):
        self.n = int(n)
        self.grid = grid
        self.delta = float(delta)
        self.levels = grid if isinstance(grid, Levels) else None
        if self.levels is not None:
            tb = self.levels.table()
            self.level_table = torch.frombuffer(bytearray(tb), dtype=torch.uint8).to(device)
            self.level_hmax = torch.zeros(self.levels.nlevels, dtype=torch.float32, device=device)
        nc = grid.ncell
        self.perm = torch.empty(self.n, dtype=torch.int32, device=device)
        self.cell_start = torch.empty(nc + 1, dtype=torch.int32, device=device)
        self.cell_hmax = torch.empty(nc, dtype=torch.float32, device=device)
        self.cell_id = torch.empty(self.n, dtype=torch.int32, device=device)
        self.posh = torch.empty((self.n, 4), dtype=F64, device=device)
        self.frec = torch.empty((self.n, 4), dtype=torch.float32, device=device)
        self._cells = None

    def cells(self):
        if self._cells is None:
            c = N.Cells()
            if self.levels is not None:
                c.grid = self.levels.frame
                c.nlevels = self.levels.nlevels
                c.exp0 = self.levels.exp0
                c.levels = self.level_table.data_ptr()
                c.level_hmax = self.level_hmax.data_ptr()
            else:
                c.grid = self.grid
            c.n = self.n
            c.cell_start = self.cell_start.data_ptr()
            c.cell_hmax = self.cell_hmax.data_ptr()
            c.cell_id = self.cell_id.data_ptr()
            c.posh = self.posh.data_ptr()
            c.frec = self.frec.data_ptr()
            c.delta = self.delta
            self._cells = c
        return C.byref(self._cells)


def make_consts(k) -> N.Consts:
    """k: mapping with the reference's constant names (pyfluid.py:5-41)."""
    return N.Consts(
        float(k["PARTICLE_MASS"]), float(k["COUPLING_CONST"]), float(k["SMOOTHLENGTH_VARIATION_TOLERANCE"]),
        int(k["NEWTON_ITERATION_LIMIT"]), int(k["BISECTION_ITERATION_LIMIT"]),
        float(k["ALPHA_SPH"]), float(k["BETA_SPH"]), float(k["EPSILON"]), float(k["BISECTION_TOLERANCE"]),
        float(k["INITIAL_BISECTION_SMOOTHLENGTH_A"]), float(k["INITIAL_BISECTION_SMOOTHLENGTH_B"]),
        float(k["G"]), float(k["ADIABATIC_INDEX"]))


def with_mass(consts, mass_t):
    """a copy of `consts` whose sphb_consts.mass points at mass_t (a device fp64 tensor indexed like the particle arrays
    of the calls it is used for), or `consts` itself when mass_t is None.  The copy keeps the tensor alive."""
    if mass_t is None:
        return consts
    c = N.Consts.from_buffer_copy(consts)
    c.mass = mass_t.data_ptr()
    c._mass_t = mass_t
    return c


# cell edge (y,z) = CELL_FACTOR * 2 * mean(h): one cell ~ one support radius of a typical particle; x edge = that / XDIV
CELL_FACTOR = float(__import__('os').environ.get('SPHB_CELL_FACTOR', '0.9'))
XDIV = int(__import__('os').environ.get('SPHB_XDIV', '1'))
MAX_CELLS_PER_PARTICLE = 8


def choose_grid(lo, hi, h_mean, n, pad_cells=1.0, cell=None, xdiv=None):
    """Grid geometry from the bounding box [lo, hi] (3-vectors of python floats) and the mean h.
    Outliers are legal (they are clamped into border cells), so the box only affects speed."""
    xdiv = max(1, int(xdiv if xdiv is not None else XDIV))
    ext = [max(hi[k] - lo[k], 0.0) for k in range(3)]
    if cell is None:
        cell = 2.0 * h_mean * CELL_FACTOR
        if cell > 0.0 and math.isfinite(cell):
            # 12 significant bits: the mean h of a decomposed run (per-rank sums, then an all-reduce) differs from the
            # single-GPU one in the last bits, and the two runs must still draw the very same cells
            m, ex = math.frexp(cell)
            cell = math.ldexp(round(m * 4096.0) / 4096.0, ex)
    if not (cell > 0.0) or not math.isfinite(cell):
        cell = max(max(ext), 1.0)
    cell = max(cell, max(ext) * 1e-4, 1e-300)   # at most 1e4 cells per axis (y,z)
    max_cells = max(64, MAX_CELLS_PER_PARTICLE * n)
    while True:
        dims = [int(math.floor((ext[0] + 2 * pad_cells * cell) / (cell / xdiv))) + 1,
                int(math.floor((ext[1] + 2 * pad_cells * cell) / cell)) + 1,
                int(math.floor((ext[2] + 2 * pad_cells * cell) / cell)) + 1]
        if dims[0] * dims[1] * dims[2] <= max_cells:
            break
        if xdiv > 1:
            xdiv //= 2
        else:
            cell *= 1.26
    g = N.Grid(lo[0] - pad_cells * cell, lo[1] - pad_cells * cell, lo[2] - pad_cells * cell, cell, dims[0], dims[1], dims[2], xdiv)
    # fp32 positions relative to the origin: |x| <= L  ->  error <= 2^-24 L per coordinate; 4*2^-23*L covers a
    # three-component difference with room; 2x slack for particles that drift outside the box
    L = max(dims[0] / xdiv, dims[1], dims[2]) * cell
    delta = 2.0 * 4.0 * 2.0 ** -23 * L
    return g, delta


# nested grids (one per octave of h) when the largest h exceeds this multiple of the mean h; 0 = never
LEVEL_SPREAD = float(os.environ.get("SPHB_LEVEL_SPREAD", "6.0"))


def choose_levels(stats, exp_top, lo, hi, hmax):
    """Nested grids from sphb_level_stats' per-octave rows [count, sum h, lo xyz, hi xyz] (16 octaves ending at exp_top)
    and the global bounding box.  -> (Levels, delta), or None when fewer than two octaves are occupied."""
    occ = [b for b in range(16) if stats[b][0] > 0]
    if not occ or 16 - occ[0] < 2:
        return None
    b0 = occ[0]
    exp0 = exp_top - 15 + b0
    pad = 2.0 * hmax if math.isfinite(hmax) and hmax > 0 else 0.0
    ext = max(max(hi[k] - lo[k], 0.0) for k in range(3)) + 2.0 * pad
    ext = max(ext, 1e-300)
    frame = N.Grid(lo[0] - pad, lo[1] - pad, lo[2] - pad, ext, 1, 1, 1, 1)
    delta = 2.0 * 4.0 * 2.0 ** -23 * ext
    grids = []
    for b in range(b0, 16):
        cnt, sh = stats[b][0], stats[b][1]
        if cnt <= 0:
            grids.append(N.Grid(frame.ox, frame.oy, frame.oz, ext, 1, 1, 1, 1))
            continue
        hm = sh / cnt if sh > 0 else 2.0 ** (exp0 + (b - b0))
        g, _ = choose_grid(list(stats[b][2:5]), list(stats[b][5:8]), hm, int(cnt), xdiv=1)
        grids.append(g)
    return Levels(frame, exp0, grids), delta


class _Section:
    def __init__(self, eng, name):
        self.eng, self.name = eng, name

    def __enter__(self):
        if self.eng.prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.eng.prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.eng.prof.append((self.name, self.e0, e1))
        return False


class Engine:
    """Stateless helpers around the C-ABI for one device."""

    def __init__(self, device=None):
        self.lib = require_cuda()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._ws = None
        self.prof = None
        self._bws = torch.empty(int(self.lib.sphb_bounds_workspace_bytes(0)), dtype=torch.uint8, device=self.device)
        self._b8 = torch.empty(8, dtype=F64, device=self.device)

    def _run(self, name, fn, *args):
        """call one C-ABI entry point; with profiling on, bracket it with CUDA events on the launch stream"""
        if self.prof is None:
            N.check(fn(*args), name)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(fn(*args), name)
        e1.record()
        self.prof.append((name, e0, e1))

    def section(self, name):
        """with eng.section("host:plan"): ...  -- CUDA events around a stretch of host-side torch work while profiling
        (sections may contain C-ABI calls, which are then counted in both)"""
        return _Section(self, name)

    def prof_start(self):
        self.prof = []

    def prof_stop(self):
        """-> {name: (calls, total ms)}; synchronises"""
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.prof or []:
            c, t = out.get(name, (0, 0.0))
            out[name] = (c + 1, t + e0.elapsed_time(e1))
        self.prof = None
        return out

    # -- K1 ------------------------------------------------------------------------------------
    def pack_posh(self, pos, h):
        n = pos.shape[0]
        posh = torch.empty((n, 4), dtype=F64, device=self.device)
        self._run("pack_posh", self.lib.sphb_pack_posh, _ptr(pos), _ptr(h), n, _ptr(posh), _stream())
        return posh

    def bounds(self, posh):
        """-> (lo[3], hi[3], hmax, hmean) as python floats (one tiny D2H sync)."""
        n = posh.shape[0]
        self._run("bounds", self.lib.sphb_bounds, _ptr(posh), n, _ptr(self._b8), _ptr(self._bws), self._bws.numel(), _stream())
        b = self._b8.cpu().tolist()
        return b[0:3], b[3:6], b[6], b[7] / n

    def grid_for(self, posh, cell=None):
        """-> (grid, delta): one Grid, or nested Levels when h spans decades (LEVEL_SPREAD)"""
        lo, hi, hmax, hmean = self.bounds(posh)
        if not all(math.isfinite(v) for v in lo + hi):
            raise ValueError("non-finite particle positions")
        if (cell is None and LEVEL_SPREAD > 0 and math.isfinite(hmax) and math.isfinite(hmean) and hmean > 0
                and hmax > LEVEL_SPREAD * hmean):
            lv = self.levels_for(posh, lo, hi, hmax)
            if lv is not None:
                return lv
        return choose_grid(lo, hi, hmean if math.isfinite(hmean) and hmean > 0 else hmax, posh.shape[0], cell=cell)

    def levels_for(self, posh, lo, hi, hmax):
        """per-octave statistics of h (one small kernel, one 1 KB D2H read) -> (Levels, delta) or None"""
        exp_top = math.frexp(hmax)[1] - 1            # floor(log2 hmax)
        out = torch.empty((16, 8), dtype=F64, device=self.device)
        self._run("level_stats", self.lib.sphb_level_stats, _ptr(posh), posh.shape[0], exp_top, _ptr(out), _stream())
        return choose_levels(out.cpu().tolist(), exp_top, lo, hi, hmax)

    def build(self, posh_in, grid, delta, tie_key=None, out=None) -> DeviceGrid:
        """out: a DeviceGrid of the same n and cell count to rebuild in place (the Stepper's per-stage pool)"""
        n = posh_in.shape[0]
        nested = isinstance(grid, Levels)
        if out is not None and out.n == n and out.grid.ncell == grid.ncell and not nested and out.levels is None:
            dg = out
            dg.grid, dg.delta, dg._cells = grid, float(delta), None
        else:
            dg = DeviceGrid(n, grid, delta, self.device)
        need = int(self.lib.sphb_levels_workspace_bytes(n, grid.ncell))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        if nested:
            self._run("grid_build", self.lib.sphb_grid_build_levels, _ptr(posh_in), _ptr(tie_key), n, dg.cells(), grid.ncell,
                      _ptr(dg.perm), _ptr(self._ws), self._ws.numel(), _stream())
            return dg
        self._run("grid_build", self.lib.sphb_grid_build, _ptr(posh_in), _ptr(tie_key), n, C.byref(grid), C.c_float(delta), _ptr(dg.perm), _ptr(dg.cell_start),
                                         _ptr(dg.cell_hmax), _ptr(dg.cell_id), _ptr(dg.posh), _ptr(dg.frec),
                                         _ptr(self._ws), self._ws.numel(), _stream())
        return dg

    def set_h(self, dg: DeviceGrid, h_sorted):
        if dg.levels is not None:
            self._run("grid_set_h", self.lib.sphb_grid_set_h_levels, _ptr(h_sorted), dg.cells(), dg.grid.ncell, _stream())
            return
        self._run("grid_set_h", self.lib.sphb_grid_set_h, _ptr(h_sorted), dg.n, C.byref(dg.grid), C.c_float(dg.delta), _ptr(dg.cell_start),
                                         _ptr(dg.cell_hmax), _ptr(dg.posh), _ptr(dg.frec), _stream())

    def gather(self, src, perm, ncomp):
        n = perm.shape[0]
        dst = torch.empty((n, ncomp) if ncomp > 1 else (n,), dtype=F64, device=self.device)
        self._run("gather", self.lib.sphb_gather_f64, _ptr(src), _ptr(perm), n, ncomp, _ptr(dst), _stream())
        return dst

    def scatter(self, src, perm, ncomp):
        n = perm.shape[0]
        dst = torch.empty((n, ncomp) if ncomp > 1 else (n,), dtype=F64, device=self.device)
        self._run("scatter", self.lib.sphb_scatter_f64, _ptr(src), _ptr(perm), n, ncomp, _ptr(dst), _stream())
        return dst

    def gather_u32(self, src, perm):
        dst = torch.empty(perm.shape[0], dtype=torch.int32, device=self.device)
        self._run("gather_u32", self.lib.sphb_gather_u32, _ptr(src), _ptr(perm), perm.shape[0], _ptr(dst), _stream())
        return dst

    # -- neighbour lists parked between kernels (sphb_lists) ------------------------------------
    def park_lists(self, dg, cover=1.0):
        """-> ParkedLists for this cell list: pass it to ONE producer (density_omega or newton) and then to force.
        cover: how much any h may grow between the two (1 when h does not change).  The store lives in `dg`."""
        need = int(self.lib.sphb_lists_bytes(dg.n))
        if getattr(dg, "_lists_store", None) is None or dg._lists_store.numel() < need:
            dg._lists_store = torch.empty(need, dtype=torch.uint8, device=self.device)
            dg._lists_broken = torch.zeros(1, dtype=torch.int32, device=self.device)
        dg._lists_broken.zero_()
        return ParkedLists(dg._lists_store, dg._lists_broken, self.hmax(dg.posh), cover)

    # -- K2 ------------------------------------------------------------------------------------
    def density_omega(self, dg, consts, target=None, h_target=None, rho_in=None, want_rho=True, want_dwdh=False,
                      want_omega=False, want_ngb=False, want_xi=False, lists=None):
        """-> rho_sum, dwdh_sum, omega, ngb_count (, xi when want_xi); entries not asked for are None
        lists: a ParkedLists to fill for the force call on the same cell list"""
        n = dg.n
        xi = torch.empty(n, dtype=F64, device=self.device) if want_xi else None
        rho = torch.empty(n, dtype=F64, device=self.device) if want_rho else None
        dw = torch.empty(n, dtype=F64, device=self.device) if want_dwdh else None
        om = torch.empty(n, dtype=F64, device=self.device) if want_omega else None
        ng = torch.empty(n, dtype=torch.int32, device=self.device) if want_ngb else None
        self._run("density_omega", self.lib.sphb_density_omega, dg.cells(), C.byref(consts), _ptr(target), _ptr(h_target), _ptr(rho_in), _ptr(rho),
                                            _ptr(dw), _ptr(om), _ptr(xi), _ptr(ng), _lists(lists), _stream())
        return (rho, dw, om, ng, xi) if want_xi else (rho, dw, om, ng)

    # -- K3 ------------------------------------------------------------------------------------
    def newton(self, dg, consts, h_guess, status_ptr, target=None, details=False, want_omega=False, lists=None):
        """-> h (, iters, code) (, omega at the returned h when want_omega)
        lists: a ParkedLists to fill for the force call on the same cell list (after set_h with the result)"""
        n = dg.n
        om = torch.empty(n, dtype=F64, device=self.device) if want_omega else None
        h = torch.full((n,), float("nan"), dtype=F64, device=self.device) if target is not None else \
            torch.empty(n, dtype=F64, device=self.device)
        it = torch.zeros(n, dtype=torch.int32, device=self.device) if details else None
        code = torch.zeros(n, dtype=torch.int32, device=self.device) if details else None
        fail = torch.empty(n, dtype=torch.int32, device=self.device)
        self._run("newton_h", self.lib.sphb_newton_h, dg.cells(), C.byref(consts), _ptr(target), _ptr(h_guess), _ptr(h), _ptr(om), _ptr(it), _ptr(code),
                                       _ptr(fail), status_ptr, _lists(lists), _stream())
        out = (h, it, code) if details else (h,)
        if want_omega:
            out = out + (om,)
        return out if len(out) > 1 else out[0]

    def bisect(self, dg, consts, slots, status_ptr):
        m = slots.shape[0]
        h = torch.empty(m, dtype=F64, device=self.device)
        code = torch.empty(m, dtype=torch.int32, device=self.device)
        oob = torch.empty(m, dtype=torch.int32, device=self.device)
        self._run("bisect_h", self.lib.sphb_bisect_h, dg.cells(), C.byref(consts), _ptr(slots), m, _ptr(h), None, _ptr(code), _ptr(oob), 0,
                                       status_ptr, _stream())
        return h, code, oob

    def hmax(self, posh):
        out = torch.empty(1, dtype=torch.float32, device=self.device)
        self._run("hmax", self.lib.sphb_hmax, _ptr(posh), posh.shape[0], _ptr(out), _stream())
        return out

    def neighbours(self, dg, query_slots, cap, symmetric=False):
        m = query_slots.shape[0]
        out = torch.empty((m, cap), dtype=torch.int32, device=self.device)
        cnt = torch.empty(m, dtype=torch.int32, device=self.device)
        hm = self.hmax(dg.posh) if symmetric else None
        self._run("neighbours", self.lib.sphb_neighbours, dg.cells(), _ptr(query_slots), m, 1 if symmetric else 0, _ptr(hm), _ptr(out), cap,
                                         _ptr(cnt), _stream())
        return out, cnt

    # -- kept neighbour lists (sphb_nlist) -------------------------------------------------------
    def level_dmax(self, dg, margins):
        """float32 upper bounds of skin_i h_i (as the builder forms it) over the records of each level of the cell list,
        and over the records of each cell"""
        h = dg.posh[:, 3]
        d = margins[:, 1].double() * (h * 1.00002 + dg.delta)
        d = torch.where((h > 0) & torch.isfinite(h), d, torch.full_like(d, float("inf")))
        up = lambda t: torch.nextafter(t.float(), torch.full_like(t, float("inf")).float())
        cell = torch.zeros(dg.grid.ncell, dtype=torch.float32, device=self.device)
        cell.scatter_reduce_(0, dg.cell_id.long(), up(d), reduce="amax", include_self=True)
        self._cell_dmax = cell
        if dg.levels is None:
            m = d.max().reshape(1) if d.numel() else torch.zeros(1, dtype=F64, device=self.device)
        else:
            idx = torch.tensor(dg.levels.bases + [dg.levels.ncell], dtype=torch.int64, device=self.device)
            edges = dg.cell_start.long().index_select(0, idx).tolist()
            zero = torch.zeros((), dtype=F64, device=self.device)
            m = torch.stack([d[a:b].max() if b > a else zero for a, b in zip(edges[:-1], edges[1:])])
        return torch.nextafter(m.float(), torch.full_like(m, float("inf")).float())

    def nl_build(self, dg, cover_h, skin, target=None, hmax=None, pool_rows=None, max_bytes=96 * 2 ** 30,
                 consts=None, margins=None) -> KeptLists:
        """scan the cell list once and keep every warp's candidate rows.  Reads the build's flags (one 4-byte D2H
        after the kernel): a pool that turned out too small is enlarged to what the build asked for and the build is
        repeated.  The returned lists may still carry SPHB_NL_UNLISTABLE / SPHB_NL_POOL (nl.build_flags): the caller then
        has to use the scanning kernels.
        consts: also form Omega of the build-time records while the lists are in the staging area -> nl.omega0"""
        hm = hmax if hmax is not None else self.hmax(dg.posh)
        om = torch.empty(dg.n, dtype=F64, device=self.device) if consts is not None else None
        ldm = self.level_dmax(dg, margins) if margins is not None else None
        cdm = self._cell_dmax if margins is not None else None
        self._cell_dmax = None
        for _ in range(3):
            nl = KeptLists(self, dg, cover_h, skin, pool_rows=pool_rows, margins=margins, level_dmax=ldm, cell_dmax=cdm)
            nl._hmax = hm
            nl.omega0 = om
            self._run("nl_build", self.lib.sphb_nlist_build, dg.cells(), _ptr(target), _ptr(hm), nl.ref(),
                      C.byref(consts) if consts is not None else None, _ptr(om), _stream())
            nl.build_flags = nl.read_flags()
            if not (nl.build_flags & N.NL_POOL) or (nl.build_flags & N.NL_UNLISTABLE):
                break
            need = int(nl.rows_used() * 1.05) + 4096
            if need * 128 > max_bytes:
                break
            pool_rows = need
            del nl
        return nl

    def nl_mark(self, nl, n_slots):
        """-> uint8 (n_slots): which sorted slots the lists name"""
        mark = torch.zeros(n_slots, dtype=torch.uint8, device=self.device)
        self._run("nl_mark", self.lib.sphb_nlist_mark, nl.ref(), _ptr(mark), _stream())
        return mark

    def nl_remap(self, nl, slot_map):
        self._run("nl_remap", self.lib.sphb_nlist_remap, nl.ref(), _ptr(slot_map), _stream())

    def nl_check(self, nl, posh):
        self._run("nl_check", self.lib.sphb_nlist_check, nl.ref(), _ptr(posh), _stream())

    def nl_density_omega(self, nl, consts, posh, target=None, rho_in=None, want_rho=False, want_dwdh=False, want_omega=True,
                         want_ngb=False):
        n = nl.n
        rho = torch.empty(n, dtype=F64, device=self.device) if want_rho else None
        dw = torch.empty(n, dtype=F64, device=self.device) if want_dwdh else None
        om = torch.empty(n, dtype=F64, device=self.device) if want_omega else None
        ng = torch.empty(n, dtype=torch.int32, device=self.device) if want_ngb else None
        self._run("nl_density_omega", self.lib.sphb_nlist_density_omega, nl.ref(), C.byref(consts), _ptr(target), _ptr(posh),
                  _ptr(rho_in), _ptr(rho), _ptr(dw), _ptr(om), _ptr(ng), _stream())
        return rho, dw, om, ng

    def nl_newton(self, nl, consts, posh, status_ptr, target=None, h_guess=None, want_omega=False, want_posh=True,
                  want_iters=False):
        """-> h, omega (or None), posh_out (or None), iters (or None)"""
        n = nl.n
        h = torch.empty(n, dtype=F64, device=self.device)
        om = torch.empty(n, dtype=F64, device=self.device) if want_omega else None
        po = torch.empty((n, 4), dtype=F64, device=self.device) if want_posh else None
        it = torch.zeros(n, dtype=torch.int32, device=self.device) if want_iters else None
        self._run("nl_newton_h", self.lib.sphb_nlist_newton_h, nl.ref(), C.byref(consts), _ptr(target), _ptr(posh),
                  _ptr(h_guess), _ptr(h), _ptr(om), _ptr(po), _ptr(it), status_ptr, _stream())
        return h, om, po, it

    def nl_eos(self, posh, vel, u, omega, consts, status_ptr):
        """-> recB (n,4) = vx, vy, vz, A and recD (n,2) = rho, cs for nl_force(..., recD=)"""
        n = posh.shape[0]
        recB = torch.empty((n, 4), dtype=F64, device=self.device)
        recD = torch.empty((n, 2), dtype=F64, device=self.device)
        self._run("eos", self.lib.sphb_nlist_eos, _ptr(posh), _ptr(vel), _ptr(u), _ptr(omega), n, C.byref(consts), _ptr(recB),
                  _ptr(recD), status_ptr, _stream())
        return recB, recD

    def nl_force(self, nl, consts, posh, recB, recC, status_ptr, target=None, recD=None):
        n = nl.n
        acc = torch.zeros((n, 3), dtype=F64, device=self.device) if target is not None else \
            torch.empty((n, 3), dtype=F64, device=self.device)
        dudt = torch.zeros(n, dtype=F64, device=self.device) if target is not None else \
            torch.empty(n, dtype=F64, device=self.device)
        self._run("nl_force", self.lib.sphb_nlist_force, nl.ref(), C.byref(consts), _ptr(target), _ptr(posh), _ptr(recB),
                  _ptr(recC), _ptr(recD), _ptr(acc), _ptr(dudt), status_ptr, _stream())
        return acc, dudt

    def rows_pack(self, arrays, idx, buf, unpack=False):
        """rows idx of up to four fp64 arrays <-> one packed (len(idx), K) buffer, in one launch (ghost refresh)"""
        a = [(t, t.shape[1] if t.dim() > 1 else 1) for t in arrays] + [(None, 0)] * (4 - len(arrays))
        self._run("rows_pack", self.lib.sphb_rows_pack, _ptr(a[0][0]), a[0][1], _ptr(a[1][0]), a[1][1], _ptr(a[2][0]), a[2][1],
                  _ptr(a[3][0]), a[3][1], _ptr(idx), idx.shape[0], _ptr(buf), 1 if unpack else 0, _stream())

    def gather_state(self, perm, pos=None, h=None, vel=None, u=None):
        """caller order -> slot order: posh (n,4), vel (n,3), u (n) (None for what was not given)"""
        n = perm.shape[0]
        po = torch.empty((n, 4), dtype=F64, device=self.device) if pos is not None else None
        ve = torch.empty((n, 3), dtype=F64, device=self.device) if vel is not None else None
        uu = torch.empty(n, dtype=F64, device=self.device) if u is not None else None
        self._run("gather_state", self.lib.sphb_gather_state, _ptr(pos), _ptr(h), _ptr(vel), _ptr(u), _ptr(perm), n, _ptr(po),
                  _ptr(ve), _ptr(uu), _stream())
        return po, ve, uu

    def scatter_state(self, perm, posh=None, vel=None, u=None):
        """slot order -> caller order: pos (n,3), h (n), vel (n,3), u (n)"""
        n = perm.shape[0]
        pos = torch.empty((n, 3), dtype=F64, device=self.device) if posh is not None else None
        h = torch.empty(n, dtype=F64, device=self.device) if posh is not None else None
        ve = torch.empty((n, 3), dtype=F64, device=self.device) if vel is not None else None
        uu = torch.empty(n, dtype=F64, device=self.device) if u is not None else None
        self._run("scatter_state", self.lib.sphb_scatter_state, _ptr(posh), _ptr(vel), _ptr(u), _ptr(perm), n, _ptr(pos), _ptr(h),
                  _ptr(ve), _ptr(uu), _stream())
        return pos, h, ve, uu

    # -- K4 ------------------------------------------------------------------------------------
    def eos(self, posh, vel, u, omega, consts, status_ptr, rho_in=None, P_in=None, records=True, want_rho_P=False, xi=None,
            want_bgrav=False):
        """-> rho, P, recB, recC (, bgrav = (G/2) xi/omega when want_bgrav)"""
        n = posh.shape[0]
        bg = torch.empty(n, dtype=F64, device=self.device) if want_bgrav else None
        rho = torch.empty(n, dtype=F64, device=self.device) if want_rho_P else None
        P = torch.empty(n, dtype=F64, device=self.device) if want_rho_P else None
        recB = torch.empty((n, 4), dtype=F64, device=self.device) if records else None
        recC = torch.empty((n, 4), dtype=F64, device=self.device) if records else None
        self._run("eos", self.lib.sphb_eos, _ptr(posh), _ptr(vel), _ptr(u), _ptr(omega), _ptr(rho_in), _ptr(P_in), _ptr(xi), n, C.byref(consts),
                                  _ptr(rho), _ptr(P), _ptr(recB), _ptr(recC), _ptr(bg), status_ptr, _stream())
        return (rho, P, recB, recC, bg) if want_bgrav else (rho, P, recB, recC)

    def gravity_direct(self, posh, consts, acc=None):
        """long-range softened gravity (all pairs, fp64); adds into `acc` (n,3) when given (same order as posh)"""
        n = posh.shape[0]
        out = acc if acc is not None else torch.empty((n, 3), dtype=F64, device=self.device)
        self._run("gravity_direct", self.lib.sphb_gravity_direct, _ptr(posh), n, C.byref(consts), 1 if acc is not None else 0,
                  _ptr(out), _stream())
        return out

    def gravity_direct_targets(self, posh_targets, posh_sources, consts, acc):
        """the long-range sum for a subset of targets against all sources (decomposed runs); adds into acc"""
        self._run("gravity_direct", self.lib.sphb_gravity_direct_targets, _ptr(posh_targets), posh_targets.shape[0],
                  _ptr(posh_sources), posh_sources.shape[0], C.byref(consts), 1, _ptr(acc), _stream())
        return acc

    def grav_eval(self, dist, h):
        n = dist.shape[0]
        phi, f, dp = (torch.empty(n, dtype=F64, device=self.device) for _ in range(3))
        self._run("grav_eval", self.lib.sphb_grav_eval, _ptr(dist), _ptr(h), n, _ptr(phi), _ptr(f), _ptr(dp), _stream())
        return phi, f, dp

    def force(self, dg, consts, recB, recC, hmax, status_ptr, target=None, bgrav=None, lists=None):
        """lists: the ParkedLists a producer filled on this cell list (drained instead of scanning)"""
        n = dg.n
        acc = torch.zeros((n, 3), dtype=F64, device=self.device) if target is not None else \
            torch.empty((n, 3), dtype=F64, device=self.device)
        dudt = torch.zeros(n, dtype=F64, device=self.device) if target is not None else \
            torch.empty(n, dtype=F64, device=self.device)
        self._run("force", self.lib.sphb_force, dg.cells(), C.byref(consts), _ptr(target), _ptr(recB), _ptr(recC), _ptr(bgrav), _ptr(hmax), _ptr(acc),
                                    _ptr(dudt), status_ptr, _lists(lists), _stream())
        return acc, dudt

    def pair_terms(self, posh, recB, recC, pairs, consts, status_ptr, bgrav=None, want_grav=False):
        """pairs: int32 (m,2) slot pairs -> Pi (m), fluid_accel_viscosity_comp (m,3)
        (, smoothed_gravity_acceleration_comp (m,3), smoothed_gravity_correction_comp (m,3) when want_grav)"""
        m = pairs.shape[0]
        pi = torch.empty(m, dtype=F64, device=self.device)
        acc = torch.empty((m, 3), dtype=F64, device=self.device)
        gr = torch.empty((m, 3), dtype=F64, device=self.device) if want_grav else None
        co = torch.empty((m, 3), dtype=F64, device=self.device) if want_grav else None
        self._run("pair_terms", self.lib.sphb_pair_terms, _ptr(posh), _ptr(recB), _ptr(recC), _ptr(pairs), m, C.byref(consts),
                  _ptr(pi), _ptr(acc), _ptr(bgrav), _ptr(gr), _ptr(co), status_ptr, _stream())
        return (pi, acc, gr, co) if want_grav else (pi, acc)

    # -- K5 ------------------------------------------------------------------------------------
    def predict(self, posh0, vel0, u0, acc0, dudt0, dt, status_ptr):
        n = posh0.shape[0]
        posh = torch.empty((n, 4), dtype=F64, device=self.device)
        vel = torch.empty((n, 3), dtype=F64, device=self.device)
        u = torch.empty(n, dtype=F64, device=self.device)
        self._run("predict", self.lib.sphb_predict, _ptr(posh0), _ptr(vel0), _ptr(u0), _ptr(acc0), _ptr(dudt0), float(dt), n, _ptr(posh),
                                      _ptr(vel), _ptr(u), status_ptr, _stream())
        return posh, vel, u

    def correct(self, posh0, vel0, u0, acc_mid, dudt_mid, dt, status_ptr):
        n = posh0.shape[0]
        posh = torch.empty((n, 4), dtype=F64, device=self.device)
        vel = torch.empty((n, 3), dtype=F64, device=self.device)
        u = torch.empty(n, dtype=F64, device=self.device)
        self._run("correct", self.lib.sphb_correct, _ptr(posh0), _ptr(vel0), _ptr(u0), _ptr(acc_mid), _ptr(dudt_mid), float(dt), n, _ptr(posh),
                                      _ptr(vel), _ptr(u), status_ptr, _stream())
        return posh, vel, u

    def check_neg_h(self, h, status_ptr):
        self._run("check_neg_h", self.lib.sphb_check_neg_h, _ptr(h), h.shape[0], status_ptr, _stream())

    def distance(self, rj, ri):
        n = rj.shape[0]
        out = torch.empty(n, dtype=F64, device=self.device)
        self._run("distance", self.lib.sphb_distance, _ptr(rj), _ptr(ri), n, _ptr(out), _stream())
        return out

    def m4_eval(self, dist, h):
        n = dist.shape[0]
        W, dW, dWdh = (torch.empty(n, dtype=F64, device=self.device) for _ in range(3))
        self._run("m4_eval", self.lib.sphb_m4_eval, _ptr(dist), _ptr(h), n, _ptr(W), _ptr(dW), _ptr(dWdh), _stream())
        return W, dW, dWdh

    def fp64_peak(self, blocks=148 * 8, iters=20000):
        ms, fl = C.c_double(), C.c_double()
        self._run("fp64_peak", self.lib.sphb_fp64_peak, blocks, iters, C.byref(ms), C.byref(fl))
        return fl.value / (ms.value * 1e-3) / 1e12   # TFLOP/s


# status stage indices of one leapfrog step
S_STAGE0, S_PRED, S_NEWTON_MID, S_STAGE_MID, S_CORR, S_NEWTON_1, S_COUNT = range(7)


MAX_REPEAT_PRINTS = 1000


def max_cell_population(cell_start):
    """largest number of particles in one cell of a cell list (cell_start: the [ncell+1] start table).  Above 8192 the
    build keeps the atomics' order inside that cell instead of the stable (input index / global id) order, and sums
    over such a cell are no longer reproducible bit for bit; nested grids (engine.Levels) keep cells at ~ a dozen"""
    if cell_start.numel() < 2:
        return 0
    return int(torch.diff(cell_start.long()).max().item())


CROWDED_CELL = 8192   # FIN_SORT_MAX of csrc/sphb_grid.cu


def raise_for_status(st, printer=print):
    """Map the device status blocks of one step to the reference's prints and RuntimeErrors,
    in program order (pyfluid.py:498-532)."""
    def emit(count, text):
        # the reference prints one line per event; beyond MAX_REPEAT_PRINTS the rest is summarised, never dropped silently
        n = int(count)
        for _ in range(min(n, MAX_REPEAT_PRINTS)):
            printer(text)
        if n > MAX_REPEAT_PRINTS:
            printer("[sph-b200] ... and %d more: %s" % (n - MAX_REPEAT_PRINTS, text))

    def soft(row):
        emit(row[1], "WARNING: Failure to converge in newton_new_h.")          # pyfluid.py:178
        emit(row[2], "ERROR: Root out of bisection bounds.")                    # pyfluid.py:136
        emit(row[3], "ERROR: Failure to converge in Bisection_new_h.")          # pyfluid.py:153

    def hard(flags):
        if flags & N.ST_NEG_P:
            raise RuntimeError("Obtained negative pressure.")                   # pyfluid.py:258
        if flags & N.ST_NEG_PI:
            raise RuntimeError("Obtained negative value for Pi")                # pyfluid.py:282
        if flags & N.ST_NEG_VISC:
            raise RuntimeError("Obtained negative energy rate due to viscosity.")  # pyfluid.py:350
        if flags & N.ST_NEG_E:
            raise RuntimeError("Obtained negative energy.")                     # pyfluid.py:515,528
        if flags & N.ST_NEG_H:
            raise RuntimeError("Obtained negative smoothing length.")           # pyfluid.py:210

    for row in st:
        soft(row)
        hard(int(row[0]))


PARK_COVER = float(os.environ.get("SPHB_PARK_COVER", "1.016"))


class Stepper:
    """var_smoothlength_leapfrog on device tensors (pyfluid.py:498-532); with consts.G != 0 the softened self-gravity
    terms (:420-437, :481-483) are included (single GPU, all-pairs long-range sum).

    step(pos0 (n,3), vel0 (n,3), u0 (n), h0 (n), dt) -> pos1, vel1, u1, h1 in caller order; inputs are
    not modified.  `defer_status=True` skips the end-of-step status read (bench loops read it once)."""

    def __init__(self, engine: Engine):
        self.e = engine
        self.last_status = None
        # (grid, delta) reused across steps when set by the caller.  A pinned geometry is the caller's to keep valid: delta
        # bounds the fp32 pre-filter error only while |x - origin| stays within ~4 box lengths (outliers are clamped into
        # border cells, which is exact but slow); leave it None and every step draws its cells from the bounding box
        self.geom = None
        self._pool = [None, None, None]   # the three cell lists of a step, rebuilt in place while n and the grid size hold

    def step(self, pos0, vel0, u0, h0, dt, consts, defer_status=False, printer=print, omega0=None, want_omega1=False,
             hooks=None, mass=None):
        """omega0 (caller order, optional): Omega(pos0, h0) if the caller already has it -- the time loop does, because
        the previous step's last Newton kernel returns it (want_omega1=True appends omega1 to the result).  Without
        it the step computes Omega0 itself, exactly like a stand-alone var_smoothlength_leapfrog call."""
        e = self.e
        st = Status(S_COUNT, e.device)
        posh0 = e.pack_posh(pos0, h0)
        grid, delta = self.geom if self.geom is not None else e.grid_for(posh0)

        def geom_of(posh):
            """one grid serves the three snapshots of a step (outliers are clamped into its border cells); nested grids
            must contain their particles, so they are drawn again for every snapshot"""
            return e.grid_for(posh) if (self.geom is None and isinstance(grid, Levels)) else (grid, delta)

        # State stays in CALLER order between stages; every cell list is built from caller-order positions, so
        # ties inside a cell are always broken by the caller index and every sum has one fixed order (the
        # multi-GPU runner does the same with global ids: k-rank results equal 1-rank results bit for bit).
        grav = consts.G != 0.0
        consts0 = consts
        of = lambda dg: with_mass(consts0, e.gather(mass, dg.perm, 1)) if mass is not None else consts0   # masses in grid order

        park = os.environ.get("SPHB_NO_PARK") is None

        def stage(dg, vel_c, u_c, sidx, om=None, lists=None):
            """the six calls at pyfluid.py:502-507 / :517-522 -> acc, dudt in caller order
            (om: Omega already produced by the Newton kernel for this snapshot; lists: its parked neighbour lists)"""
            consts = of(dg)
            if not grav and om is None:
                lists = e.park_lists(dg) if park else None   # the Omega pass scans once for itself and for the force
                _, _, om, _ = e.density_omega(dg, consts, want_rho=False, want_omega=True, lists=lists)
            if hooks is not None and sidx == S_STAGE0:
                hooks.before_vel_u()          # velocities and energies are first needed here
            vel_s, u_s = e.gather(vel_c, dg.perm, 3), e.gather(u_c, dg.perm, 1)
            if grav:
                # xi_arr (:504) rides in the Omega pass; its compact correction (:434-437) rides in the force kernel;
                # the long-range pair force (:420-432) is an exact all-pairs sum
                _, _, om, _, xi = e.density_omega(dg, consts, want_rho=False, want_omega=True, want_xi=True)
                _, _, rB, rC, bg = e.eos(dg.posh, vel_s, u_s, om, consts, st.ptr(sidx), xi=xi, want_bgrav=True)
                acc, dudt = e.force(dg, consts, rB, rC, e.hmax(dg.posh), st.ptr(sidx), bgrav=bg)
                e.gravity_direct(dg.posh, consts, acc=acc)
            else:
                _, _, rB, rC = e.eos(dg.posh, vel_s, u_s, om, consts, st.ptr(sidx))
                acc, dudt = e.force(dg, consts, rB, rC, e.hmax(dg.posh), st.ptr(sidx), lists=lists)
            return e.scatter(acc, dg.perm, 3), e.scatter(dudt, dg.perm, 1)

        def solve_h(dg, sidx, want_omega=False, lists=None):
            """newton_smoothlength_arr(pos, guess = the h carried in the snapshot), grid order"""
            r = e.newton(dg, of(dg), dg.posh[:, 3].contiguous(), st.ptr(sidx), want_omega=want_omega, lists=lists)
            e.check_neg_h(r[0] if want_omega else r, st.ptr(sidx))
            return r

        # ---- stage 0 at (pos0, h0): pyfluid.py:502-507 ------------------------------------------
        pool = self._pool if os.environ.get("SPHB_NO_POOL") is None else [None, None, None]
        g0 = pool[0] = e.build(posh0, grid, delta, out=pool[0])
        acc0, dudt0 = stage(g0, vel0, u0, S_STAGE0, om=None if omega0 is None else e.gather(omega0, g0.perm, 1))
        # ---- predictor: pyfluid.py:511-515 ---------------------------------------------------------
        posh_m, vel_m, u_m = e.predict(posh0, vel0, u0, acc0, dudt0, dt, st.ptr(S_PRED))
        # ---- h_mid = newton(pos_mid, h0), stage mid: pyfluid.py:516-522 -------------------------------
        g1 = pool[1] = e.build(posh_m, *geom_of(posh_m), out=pool[1])
        # the mid-step Newton kernel scans with the symmetric support and a cover of PARK_COVER on every h: its lists
        # serve the force kernel at h_mid as long as no h grew by more than that (else the force kernel scans itself)
        l1 = e.park_lists(g1, cover=PARK_COVER) if (park and not grav) else None
        h_mid_s, om_mid = solve_h(g1, S_NEWTON_MID, want_omega=True, lists=l1)
        e.set_h(g1, h_mid_s)
        acc1, dudt1 = stage(g1, vel_m, u_m, S_STAGE_MID, om=om_mid, lists=l1)
        h_mid = e.scatter(h_mid_s, g1.perm, 1)
        # ---- corrector: pyfluid.py:524-528 ---------------------------------------------------------------
        posh_1, vel1, u1 = e.correct(posh0, vel0, u0, acc1, dudt1, dt, st.ptr(S_CORR))
        if hooks is not None:
            hooks.after_corrector(posh_1, vel1, u1)   # pos1, vel1, u1 are final: their copy-out can start now
        # ---- h1 = newton(pos1, h_mid): pyfluid.py:530 ------------------------------------------------------
        posh_1 = e.pack_posh(posh_1[:, :3].contiguous(), h_mid)
        g2 = pool[2] = e.build(posh_1, *geom_of(posh_1), out=pool[2])
        r = solve_h(g2, S_NEWTON_1, want_omega=want_omega1)
        h1 = e.scatter(r[0] if want_omega1 else r, g2.perm, 1)
        pos1 = posh_1[:, :3].contiguous()
        self.last_status = st
        if not defer_status:
            raise_for_status(st.read(), printer)
        if want_omega1:
            return pos1, vel1, u1, h1, e.scatter(r[1], g2.perm, 1)
        return pos1, vel1, u1, h1


NL_COVER_H = float(os.environ.get("SPHB_NL_COVER_H", "1.03"))
NL_SKIN = float(os.environ.get("SPHB_NL_SKIN", "0.03"))


class ListMargins:
    """How wide the next neighbour lists are built (cover_h, skin) and when they are rebuilt.

    Lists that are to live L steps need margins of L x the per-step motion; their drains cost ~ (cover_h + skin)^3,
    the rebuild 1/L of BUILD_OVER_DRAIN: take the cheapest L.  The per-step motion (largest displacement / h and
    largest relative h growth over all particles -- and over all ranks) is measured by the lists' own checks and
    extrapolated linearly; a step that would leave the margins makes the lists stale BEFORE it runs.  Shared by the
    single-GPU FastStepper and the slab runner: the same measurements give the same decisions on every rank."""
    MAX_SKIN = 0.25
    MAX_GROW = 0.25
    SAFETY = float(os.environ.get("SPHB_NL_SAFETY", "1.25"))                  # margin on the extrapolated motion
    BUILD_OVER_DRAIN = float(os.environ.get("SPHB_NL_BUILD_COST", "0.4"))    # one rebuild / one step's drains

    def __init__(self, cover_h=None, skin=None, adaptive=True):
        self.cover_h = NL_COVER_H if cover_h is None else float(cover_h)
        self.skin = NL_SKIN if skin is None else float(skin)
        self.adaptive = adaptive
        self.rate_d = self.rate_g = None   # per-step displacement / h and h growth
        self.slope_d = self.slope_g = 0.0  # their change from one step to the next
        self.age = 0                       # steps run on the current lists
        self._moved = self._grown = 0.0
        self.widened = 0
        self.per_particle = adaptive and self.PER_PARTICLE
        self.pp_safety = self.PP_SAFETY
        self.life = 1                      # planned life of the current lists (per-particle margins)

    # ---- per-particle margins -----------------------------------------------------------------------------------------
    # The largest motion belongs to a few per cent of the particles (free surfaces, the contact of the Sod tube, a close
    # pair in the Plummer core).  Uniform margins sized for them make every list long, or every life short; with
    # sphb_nlist.margins each record gets the cover and skin ITS OWN measured motion asks for over the planned life.
    PER_PARTICLE = os.environ.get("SPHB_NL_PP", "1") != "0"
    PP_SAFETY = float(os.environ.get("SPHB_NL_PP_SAFETY", "2.0"))
    PP_FLOOR = 0.003      # every record may move / grow by at least this much (fractions of h)
    PP_MAX = 0.45         # no record is given more; a life that would need more is not planned
    LIVES = (1, 2, 3, 4, 6, 8, 12, 16, 24, 32)
    PP_BINS = 128

    def plan(self, rd, rg, steps_seen, reduce=None):
        """rd, rg: per-record per-step displacement / h and relative h growth, measured over the `self.age` steps the
        last lists lived (device tensors, any order).  Chooses the life L (steps) of the next lists by the modelled cost
        -- mean list volume (1 + cover + skin)^3 + BUILD_OVER_DRAIN / L --, and returns margins (n,2) float32 =
        {cover_i, skin_i} in the order of rd / rg, plus their upper bounds.  A start-up flow accelerates: rates measured
        over steps [T - age, T] are scaled by (T + L/2) / (T - age/2) (constant acceleration from rest), at most 6x.
        reduce(counts, maxes): all-reduce over ranks (sum of int64 / max) so that every rank plans the same life."""
        f = self.pp_safety
        dev = rd.device
        age = max(self.age, 1)
        T = max(float(steps_seen), float(age))
        Ls = torch.tensor(self.LIVES, dtype=F64, device=dev)
        ramp = ((T + 0.5 * Ls) / max(T - 0.5 * age, 0.5)).clamp(max=6.0)
        # a 2-D histogram of the two rates (three bins per octave, 2^-36 .. 2^6): integer counts do not depend on the
        # order of the records or on how they are spread over ranks, so every run of the same physics plans the same life
        nb = self.PP_BINS
        bd = ((torch.log2(rd) + 36.0) * 3.0).floor().clamp_(0, nb - 1).long()
        bg = ((torch.log2(rg) + 36.0) * 3.0).floor().clamp_(0, nb - 1).long()
        sums = torch.bincount(bd * nb + bg, minlength=nb * nb)
        maxes = torch.stack([rd.max(), rg.max()]) if rd.numel() else torch.zeros(2, dtype=F64, device=dev)
        if reduce is not None:
            reduce(sums, maxes)
        rep = torch.exp2((torch.arange(nb, dtype=F64, device=dev) + 1.0) / 3.0 - 36.0)      # upper edge of each bin
        k = (f * ramp * Ls)[:, None]
        s_b = (k * rep[None, :] + self.PP_FLOOR).clamp(max=self.PP_MAX)                     # (lives, bins)
        vol = (1.0 + s_b[:, :, None] + s_b[:, None, :]) ** 3                                # (lives, bins_d, bins_g)
        cnt = sums.to(F64).reshape(1, nb, nb)
        mean_vol = ((vol * cnt).sum(dim=(1, 2)) / cnt.sum().clamp(min=1.0)).tolist()
        maxes, kk = maxes.tolist(), (f * ramp * Ls).tolist()
        best = None
        for i, L in enumerate(self.LIVES):
            need = kk[i] * max(maxes[0], maxes[1]) + self.PP_FLOOR
            if best is not None and need > self.PP_MAX:
                break
            cost = mean_vol[i] + self.BUILD_OVER_DRAIN / L
            if best is None or cost < best[0]:
                best = (cost, i)
        i = best[1]
        self.life = self.LIVES[i]
        self.pp_k = kk[i]
        self.cover_h = 1.0 + min(self.PP_MAX, kk[i] * maxes[1] + self.PP_FLOOR) * 1.000001 + 1e-7
        self.skin = min(self.PP_MAX, kk[i] * maxes[0] + self.PP_FLOOR) * 1.000001 + 1e-7
        return self.margins_for(rd, rg)

    def margins_for(self, rd, rg):
        """{cover_i, skin_i} (n,2) float32 for the planned life: a pure function of each record's own rates, so a ghost
        copy gets the very margins its owner gives the record"""
        skin = (self.pp_k * rd + self.PP_FLOOR).clamp(max=self.PP_MAX)
        cover = 1.0 + (self.pp_k * rg + self.PP_FLOOR).clamp(max=self.PP_MAX)
        m = torch.stack([cover, skin], dim=1).float()
        return torch.nan_to_num(m, nan=1.0 + self.PP_MAX, posinf=1.0 + self.PP_MAX).contiguous()

    def note_used(self, used_d, used_g):
        """after a successful step on per-particle margins: used_* = the largest fraction of its allowance any record has
        used up.  -> True when the lists should be rebuilt before the next step (planned life over, or the next step --
        at the pace so far, with a quarter in hand -- would break them)"""
        self.age += 1
        u = max(used_d, used_g)
        return self.age >= self.life or u * (self.age + 1.0) / self.age * 1.25 > 1.0

    def failed_pp(self):
        """lists with per-particle margins broke before their planned life: plan with more in hand.  -> False: give up"""
        self.pp_safety *= 1.5
        self.widened += 1
        return self.pp_safety <= 12.0

    def choose(self):
        """-> cover_h, skin of the next build"""
        if not self.adaptive or self.rate_d is None:
            return self.cover_h, self.skin
        best = None
        sd, sg = max(self.slope_d, 0.0), max(self.slope_g, 0.0)   # the per-step motion itself grows (accelerating flow)
        f = self.SAFETY
        for L in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32):
            skin = f * (self.rate_d * L + sd * L * (L + 1) / 2) + 0.001
            grow = f * (self.rate_g * L + sg * L * (L + 1) / 2) + 0.001
            if best is not None and (skin > self.MAX_SKIN or grow > self.MAX_GROW):
                break
            skin, grow = min(skin, self.MAX_SKIN), min(grow, self.MAX_GROW)
            cost = (1.0 + grow + skin) ** 3 + self.BUILD_OVER_DRAIN / L
            if best is None or cost < best[0]:
                best = (cost, 1.0 + grow, skin)
        self.cover_h, self.skin = best[1], best[2]
        return self.cover_h, self.skin

    def built(self):
        self.age = 0
        self._moved = self._grown = 0.0

    def note(self, moved, grown, aging):
        """after a successful step on lists with the current margins: moved / grown = the largest displacement / h
        and h growth since their build.  -> True when the lists should be rebuilt before the next step"""
        self.age += 1
        dd, dg_ = moved - self._moved, grown - self._grown
        self._moved, self._grown = moved, grown
        rd, rg = max(dd, moved / self.age), max(dg_, grown / self.age)
        if self.rate_d is not None:
            self.slope_d, self.slope_g = rd - self.rate_d, rg - self.rate_g
        self.rate_d, self.rate_g = rd, rg
        if not self.adaptive:
            return bool(aging)
        f = self.SAFETY
        return (moved + f * (rd + max(self.slope_d, 0.0)) > self.skin or
                grown + f * (rg + max(self.slope_g, 0.0)) > self.cover_h - 1.0)

    def failed(self, fresh, moved, grown):
        """a step ended with SPHB_NL_BROKEN.  fresh: its lists had been built for that very step.  -> False: give up"""
        if fresh:
            # even lists built for this state broke within the step: the motion per step exceeds the margins -- widen
            if self.skin >= self.MAX_SKIN:
                return False
            self.skin = min(self.MAX_SKIN, max(2.0 * self.skin, 0.02))
            self.cover_h = 1.0 + min(self.MAX_GROW, max(2.0 * (self.cover_h - 1.0), 0.02))
            # the measured rates said less: do not let the next choice shrink the margins again
            self.rate_d = max(self.rate_d or 0.0, self.skin / self.SAFETY)
            self.rate_g = max(self.rate_g or 0.0, (self.cover_h - 1.0) / self.SAFETY)
            self.widened += 1
        elif self.adaptive and self.rate_d is not None:
            # kept lists broke: the motion since the build is faster than the rates they were sized for
            self.rate_d = max(self.rate_d, 2.0 * moved / max(self.age + 1, 1))
            self.rate_g = max(self.rate_g, 2.0 * grown / max(self.age + 1, 1))
        return True


def slot_rates(posh_build, posh_now, age):
    """per-record motion per step since the lists were built: displacement / h_build and relative h growth (non-finite:
    a large number, i.e. the widest margins)"""
    hb = posh_build[:, 3]
    rd = (posh_now[:, :3] - posh_build[:, :3]).norm(dim=1) / hb / age
    rg = (posh_now[:, 3] / hb - 1.0).clamp(min=0.0) / age
    big = 10.0
    return torch.nan_to_num(rd, nan=big, posinf=big).clamp_(max=big), torch.nan_to_num(rg, nan=big, posinf=big).clamp_(max=big)


class FastStepper:
    """var_smoothlength_leapfrog (pyfluid.py:498-532, G = 0) on neighbour lists that are kept across kernels and steps.

    One cell-list build + one scan (sphb_nlist_build) serve every gather kernel of every step until a device-side check
    says the lists no longer cover the state (h grew past cover_h x, or a particle moved more than skin x, its h at build
    time).  Between builds the state lives in the build's slot order; the five gather kernels of a step are pure drains.
    A step runs speculatively: if it ends with SPHB_NL_BROKEN it is redone after a rebuild from its (unmodified) input,
    and if that fails too -- or the input cannot be listed at all -- by the scanning Stepper, which replays the
    reference to the letter (bisection fallback, unbounded h).  Sums are formed in (cell, caller index) order like the
    scanning path: on the same grid geometry the results are the same bits.

    Two ways to drive it:
      * step(pos0, vel0, u0, h0, dt, consts) -> pos1, vel1, u1, h1   pure function on caller-order device tensors, same
        contract as Stepper.step; the lists of the previous call are reused when the new input passes the check;
      * load(pos, vel, u, h) / advance(dt, consts) / state()          the state stays resident in slot order (time loops).
    """

    def __init__(self, engine, cover_h=None, skin=None, adaptive=True):
        self.e = engine
        self.margins = ListMargins(cover_h, skin, adaptive)   # how wide the lists are built, when they are rebuilt
        self.safe = Stepper(engine)
        self.geom = None          # (grid, delta) pinned by the caller; None: from the bounding box at every build
        self.nl = None
        self.perm = None          # int32 (n): caller index of slot s
        self.ids = None           # int64 copy of perm (tie keys of the next build)
        self.P = self.V = self.U = None
        self.stale = True
        self.last_status = None
        self.pool_rows = None     # rows of the list pool (None: sphb_nlist_pool_rows)
        self.fresh = False
        self.fresh_failed = False
        self.last_flags = 0
        self.hold = 0             # steps left on the scanning path after the lists failed on one step
        self.caller_state = None  # resident state in caller order while on hold
        self.stats = dict(steps=0, builds=0, redos=0, safe_steps=0, reuse_rejected=0)
        self._consts = None       # the constants of the current call (the build forms Omega with them)
        self.M = None             # per-particle masses in slot order (None: the uniform consts.particle_mass)

    # ---- builds ------------------------------------------------------------------------------------
    def _rates(self):
        """per-record motion per step over the life of the current lists, in their slot order: (displacement / h_build,
        relative h growth), or None when nothing has been measured"""
        nl, age = self.nl, self.margins.age
        if nl is None or age < 1 or not self.margins.per_particle or self.P is None or self.P.shape[0] != nl.n:
            return None
        return slot_rates(nl.posh_build, self.P, age)

    def _rebuild_from(self, posh_in, vel_in, u_in, ids_in, rates=None, mass_in=None):
        """sort records (any order; ids_in = caller index of each input row, None = identity) into a fresh cell list,
        scan it once, and take the state along in the new slot order.  rates: _rates() of the lists being replaced, in
        the order of posh_in -- the new lists then get per-particle margins"""
        e = self.e
        if self.geom is not None:
            grid, delta = self.geom
            self.safe.geom = self.geom
        else:
            grid, delta = e.grid_for(posh_in)
        self.nl = None                       # drop the old pool before the new one is allocated
        dg = e.build(posh_in, grid, delta, tie_key=ids_in)
        self.V = e.gather(vel_in, dg.perm, 3) if vel_in is not None else None
        self.U = e.gather(u_in, dg.perm, 1) if u_in is not None else None
        self.M = e.gather(mass_in, dg.perm, 1) if mass_in is not None else None   # per-particle masses, slot order
        if ids_in is None:
            self.perm = dg.perm
            self.ids = dg.perm.long()
        else:
            self.ids = e.gather(ids_in.view(F64), dg.perm, 1).view(torch.int64)   # a bit-pattern copy
            self.perm = self.ids.to(torch.int32)
        self.P = dg.posh                     # never written in place: every kernel below returns fresh arrays
        pp = None
        if rates is not None:
            pp = self.margins.plan(e.gather(rates[0], dg.perm, 1), e.gather(rates[1], dg.perm, 1), self.stats["steps"])
            cover_h, skin = self.margins.cover_h, self.margins.skin
        else:
            cover_h, skin = self.margins.choose()
            self.margins.life = 1 if self.margins.per_particle else 10 ** 9
        self.margins.built()
        # Omega(pos0, h0) of the step that follows comes out of the build itself (the lists are in the staging area)
        self.nl = e.nl_build(dg, cover_h, skin, pool_rows=self.pool_rows, consts=self._consts if self.M is None else None,
                             margins=pp)
        self.pool_rows = max(self.nl.pool_rows, self.pool_rows or 0)
        self.stats["builds"] += 1
        try:   # a diagnostic, never a reason to fail a step
            if max_cell_population(dg.cell_start) > CROWDED_CELL:
                import warnings
                warnings.warn("a cell holds more than %d particles: the order inside it (and the last bits of sums over it) "
                              "is not reproducible; the input is far more clustered than its cell size" % CROWDED_CELL)
                self.stats["crowded_builds"] = self.stats.get("crowded_builds", 0) + 1
        except Exception:
            pass
        self.stale = False
        self.fresh = True                    # no step has run on these lists yet

    @property
    def skin(self):
        return self.margins.skin

    @property
    def cover_h(self):
        return self.margins.cover_h

    @property
    def age(self):
        return self.margins.age

    def _adapt(self, flags):
        """a step failed on its lists (flags = SPHB_NL_*): what to change before the rebuild.  -> False: give up"""
        if flags & (N.NL_UNLISTABLE | N.NL_POOL):
            return False      # Engine.nl_build already enlarged the pool as far as it may
        if self.nl.margins is not None:
            ok = self.margins.failed_pp()
        else:
            ok = self.margins.failed(self.fresh_failed, self.nl.moved, self.nl.grown)
        self.stats["widened"] = self.margins.widened
        return ok

    # ---- one step in slot order ---------------------------------------------------------------------
    def _slots_step(self, P0, V0, U0, dt, consts, st, hooks=None):
        """pyfluid.py:502-530 on the kept lists -> P1 (x, y, z, h1), V1, U1 in slot order; inputs untouched"""
        e, nl = self.e, self.nl
        consts = with_mass(consts, self.M)
        if self.fresh and nl.omega0 is not None and P0 is nl.dg.posh:
            om0 = nl.omega0                                                             # :503, formed by the build
        else:
            _, _, om0, _ = e.nl_density_omega(nl, consts, P0)                           # :503
        if hooks is not None:
            V0, U0 = hooks.vel_u()
        rB, rD = e.nl_eos(P0, V0, U0, om0, consts, st.ptr(S_STAGE0))                    # :502,:505
        acc0, dudt0 = e.nl_force(nl, consts, P0, rB, None, st.ptr(S_STAGE0), recD=rD)   # :506-507
        Pm, Vm, Um = e.predict(P0, V0, U0, acc0, dudt0, dt, st.ptr(S_PRED))             # :511-515
        e.nl_check(nl, Pm)
        hm, omm, Pm2, _ = e.nl_newton(nl, consts, Pm, st.ptr(S_NEWTON_MID), want_omega=True)   # :516-518
        rB, rD = e.nl_eos(Pm2, Vm, Um, omm, consts, st.ptr(S_STAGE_MID))                # :517,:520
        acc1, dudt1 = e.nl_force(nl, consts, Pm2, rB, None, st.ptr(S_STAGE_MID), recD=rD)   # :521-522
        P1, V1, U1 = e.correct(P0, V0, U0, acc1, dudt1, dt, st.ptr(S_CORR))             # :524-528
        if hooks is not None:
            hooks.after_corrector(P1, V1, U1)
        e.nl_check(nl, P1)
        _, _, P1b, _ = e.nl_newton(nl, consts, P1, st.ptr(S_NEWTON_1), h_guess=hm)      # :530
        return P1b, V1, U1

    _BAD = N.NL_BROKEN | N.NL_POOL | N.NL_UNLISTABLE

    MAX_ATTEMPTS = 4

    def _try(self, dt, consts, hooks=None):
        """run one step on the current lists; -> (ok, P1, V1, U1, status); self.last_flags = the lists' flags"""
        if self.nl.build_flags & self._BAD:      # the input could not be listed at all
            self.last_flags, self.fresh_failed, self.fresh = self.nl.build_flags, False, False
            return False, None, None, None, None
        st = Status(S_COUNT, self.e.device)
        P1, V1, U1 = self._slots_step(self.P, self.V, self.U, dt, consts, st, hooks)
        fl = self.last_flags = self.nl.read_flags()
        ok = (fl & self._BAD) == 0
        self.fresh_failed = self.fresh and not ok
        self.fresh = False
        if ok:
            if self.nl.margins is not None or self.margins.per_particle:
                # (uniform first lists of a per-particle run: one step, to measure every record's motion)
                stale = self.margins.note_used(self.nl.used_d, self.nl.used_g)
            else:
                stale = self.margins.note(self.nl.moved, self.nl.grown, fl & N.NL_AGING)
            if stale:
                self.stale = True
        return ok, P1, V1, U1, st

    # ---- resident-state API ---------------------------------------------------------------------------
    SAFE_HOLD = 8

    def load(self, pos, vel, u, h, consts=None, mass=None):
        """take a caller-order state (device tensors); it stays resident, in slot order, until state() is asked for.
        mass (n,): per-particle masses; None: the uniform consts.particle_mass"""
        self.caller_state = None
        self.hold = 0
        if consts is not None:
            self._consts = consts
        self.mass_caller = mass
        self._rebuild_from(self.e.pack_posh(pos, h), vel, u, None, mass_in=mass)

    def state(self):
        """-> pos, vel, u, h in caller order"""
        if self.caller_state is not None:
            return self.caller_state
        pos, h, vel, u = self.e.scatter_state(self.perm, self.P, self.V, self.U)
        return pos, vel, u, h

    def _safe_advance(self, dt, consts, defer_status, printer):
        self.stats["safe_steps"] += 1
        pos, vel, u, h = self.caller_state
        mass = getattr(self, "mass_caller", None)
        self.caller_state = self.safe.step(pos, vel, u, h, dt, consts, defer_status=defer_status, printer=printer, mass=mass)
        self.last_status = self.safe.last_status
        self.hold -= 1
        if self.hold <= 0:
            self.load(*self.caller_state, mass=mass)

    def advance(self, dt, consts, defer_status=True, printer=print):
        if consts.G != 0.0:
            raise ValueError("FastStepper runs the G = 0 path; use Stepper for the softened-gravity terms")
        self._consts = consts
        self.stats["steps"] += 1
        if self.hold > 0:
            return self._safe_advance(dt, consts, defer_status, printer)
        for attempt in range(self.MAX_ATTEMPTS):
            if self.stale or attempt > 0:
                self._rebuild_from(self.P, self.V, self.U, self.ids, rates=self._rates(), mass_in=self.M)
            ok, P1, V1, U1, st = self._try(dt, consts)
            if ok:
                self.P, self.V, self.U = P1, V1, U1
                self.last_status = st
                if not defer_status:
                    raise_for_status(st.read(), printer)
                return
            self.stats["redos"] += 1
            if not self._adapt(self.last_flags):
                break
        # the lists cannot serve this step (bisection fallback, unbounded h ...): scanning path for a while
        self.caller_state = self.state()
        self.nl = None
        self.hold = self.SAFE_HOLD
        self._safe_advance(dt, consts, defer_status, printer)

    # ---- pure-function API (same contract as Stepper.step) -----------------------------------------------
    def step(self, pos0, vel0, u0, h0, dt, consts, defer_status=False, printer=print, hooks=None, mass=None):
        if consts.G != 0.0:
            r = self.safe.step(pos0, vel0, u0, h0, dt, consts, defer_status=defer_status, printer=printer, hooks=hooks, mass=mass)
            self.last_status = self.safe.last_status
            return r
        self.mass_caller = mass
        e = self.e
        n = pos0.shape[0]
        self._consts = consts
        self.stats["steps"] += 1
        if self.hold > 0:
            self.hold -= 1
            self.stats["safe_steps"] += 1
            r = self.safe.step(pos0, vel0, u0, h0, dt, consts, defer_status=defer_status, printer=printer, hooks=hooks, mass=mass)
            self.last_status = self.safe.last_status
            return r
        fh = _FastHooks(self, hooks, vel0, u0) if hooks is not None else None
        have_old = self.nl is not None and self.nl.n == n and self.perm is not None
        reuse = have_old and not self.stale
        if have_old:
            # the new input in the old slot order; does it pass the lists' own validity check?
            self.P, _, _ = e.gather_state(self.perm, pos=pos0, h=h0)
        if reuse:
            e.nl_check(self.nl, self.P)
            if self.nl.read_flags() & (self._BAD | (0 if self.margins.adaptive else N.NL_AGING)):
                reuse = False
                self.stats["reuse_rejected"] += 1
            elif fh is None:
                _, self.V, self.U = e.gather_state(self.perm, vel=vel0, u=u0)
            if reuse:
                self.M = e.gather(mass, self.perm, 1) if mass is not None else None
        for attempt in range(self.MAX_ATTEMPTS):
            if not reuse:
                if fh is not None:
                    fh.wait_inputs()
                rates = self._rates() if have_old else None
                if rates is not None:
                    # the lists being replaced measured every record's motion: the new ones get per-particle margins
                    _, V_old, U_old = e.gather_state(self.perm, vel=vel0, u=u0)
                    M_old = e.gather(mass, self.perm, 1) if mass is not None else None
                    self._rebuild_from(self.P, V_old, U_old, self.ids, rates=rates, mass_in=M_old)
                else:
                    self._rebuild_from(e.pack_posh(pos0, h0), vel0, u0, None, mass_in=mass)
                have_old = False
                if fh is not None:
                    fh.have_state()
            ok, P1, V1, U1, st = self._try(dt, consts, fh)
            if ok:
                self.last_status = st
                if fh is not None:
                    out = fh.finish(P1)
                else:
                    pos1, h1, vel1, u1 = e.scatter_state(self.perm, P1, V1, U1)
                    out = (pos1, vel1, u1, h1)
                self.P, self.V, self.U = P1, V1, U1
                if not defer_status:
                    raise_for_status(st.read(), printer)
                return out
            self.stats["redos"] += 1
            reuse = False
            if not self._adapt(self.last_flags):
                break
        self.stats["safe_steps"] += 1
        self.nl = None
        self.stale = True
        self.hold = self.SAFE_HOLD
        r = self.safe.step(pos0, vel0, u0, h0, dt, consts, defer_status=defer_status, printer=printer, hooks=hooks, mass=mass)
        self.last_status = self.safe.last_status
        return r


class _FastHooks:
    """adapts the pinned-transfer hooks (before_vel_u / after_corrector take caller-order tensors) to the slot-order step"""

    def __init__(self, fs, hooks, vel0, u0):
        self.fs, self.hooks, self.vel0, self.u0 = fs, hooks, vel0, u0
        self.waited = False
        self.state_ready = False

    def wait_inputs(self):
        if not self.waited:
            self.hooks.before_vel_u()
            self.waited = True

    def have_state(self):
        self.state_ready = True   # _rebuild_from gathered V, U

    def vel_u(self):
        fs = self.fs
        if not self.state_ready:
            self.wait_inputs()
            _, fs.V, fs.U = fs.e.gather_state(fs.perm, vel=self.vel0, u=self.u0)
            self.state_ready = True
        return fs.V, fs.U

    def after_corrector(self, P1, V1, U1):
        # pos1, vel1, u1 are final: back to caller order, their copy-out can start while the last Newton solve runs
        e = self.fs.e
        pos1, _, vel1, u1 = e.scatter_state(self.fs.perm, P1, V1, U1)
        self.pos1, self.vel1, self.u1 = pos1, vel1, u1
        self.hooks.after_corrector_caller(pos1, vel1, u1)

    def finish(self, P1b):
        e = self.fs.e
        _, h1, _, _ = e.scatter_state(self.fs.perm, P1b)
        return self.pos1, self.vel1, self.u1, h1


class _PinnedHooks:
    """stream choreography of Stepper.step_pinned"""

    def __init__(self, copy_stream, ev_vel_u, out):
        self.cs, self.ev, self.out = copy_stream, ev_vel_u, out
        self.ev_out = torch.cuda.Event()

    def before_vel_u(self):
        torch.cuda.current_stream().wait_event(self.ev)

    def after_corrector(self, posh_1, vel1, u1):
        self.after_corrector_caller(posh_1[:, :3], vel1, u1)

    def after_corrector_caller(self, pos1, vel1, u1):
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self.cs):
            self.cs.wait_event(ev)
            self.out[0].copy_(pos1, non_blocking=True)
            self.out[1].copy_(vel1, non_blocking=True)
            self.out[2].copy_(u1, non_blocking=True)
            self.ev_out.record(self.cs)
        for t in (pos1, vel1, u1):
            t.record_stream(self.cs)


def step_pinned(stepper, pos_h, vel_h, u_h, h_h, dt, consts, out=None, printer=print):
    """var_smoothlength_leapfrog for HOST tensors (pinned memory): the step with its transfers overlapped.
    pos and h go up first; vel and u follow on a copy stream while the cell list and the Omega pass run; pos1, vel1, u1
    come down while the last Newton solve runs; h1 follows.  Returns pinned host tensors (`out` is reused if given)."""
    e = stepper.e
    dev = e.device
    main = torch.cuda.current_stream()
    cs = getattr(stepper, "_copy_stream", None)
    if cs is None:
        cs = stepper._copy_stream = torch.cuda.Stream(device=dev)
    if out is None:
        out = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (pos_h, vel_h, u_h, h_h)]
    pos_d = pos_h.to(dev, non_blocking=True)
    h_d = h_h.to(dev, non_blocking=True)
    ev = torch.cuda.Event()
    with torch.cuda.stream(cs):
        vel_d = vel_h.to(dev, non_blocking=True)
        u_d = u_h.to(dev, non_blocking=True)
        ev.record(cs)
    vel_d.record_stream(main)
    u_d.record_stream(main)
    hooks = _PinnedHooks(cs, ev, out)
    r = stepper.step(pos_d, vel_d, u_d, h_d, dt, consts, defer_status=True, hooks=hooks)
    out[3].copy_(r[3], non_blocking=True)
    main.wait_event(hooks.ev_out)
    torch.cuda.synchronize(dev)
    raise_for_status(stepper.last_status.read(), printer)
    return out
