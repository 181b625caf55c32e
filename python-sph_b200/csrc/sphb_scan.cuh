// sphb_scan.cuh -- the candidate scan shared by every gather kernel (K2 density/omega, K3 Newton,
// K4 force).
//
// Mapping: one lane = one target particle, one warp = 32 consecutive slots of the cell-sorted
// particle array.  For each group of lanes that share a cell row, the warp
//   1. takes the group's bounding box + search radius and turns it into rows of cells; every row is ONE
//      contiguous span of the sorted array (x is the fastest cell axis), resolved by all lanes in parallel;
//   2. stages the span through a 32-entry shared-memory tile: each lane loads one fp32 prefilter record
//      (coalesced float4), re-centres it on the group's box centre O and stores (c', |c'|^2); then every
//      lane tests every staged candidate with 3 FFMA + 1 compare:
//          |c' - t'|^2 <= thr   <=>   |c'|^2 - 2 c'.t'  <=  thr - |t'|^2   (+ rounding allowance)
//      (broadcast LDS.128, no bank conflicts);
//   3. appends survivors as 16-bit (segment, offset) codes to a per-lane list in shared memory
//      ([k][lane]); `flush(cnt)` evaluates the listed pairs in FP64 when a list is about to overflow, when
//      the 32-entry segment table is full, and at the end.
// The FP32 test only has to be conservative (positions carry an absolute error <= delta, thresholds are
// rounded up, `err` covers the re-centred arithmetic); membership is decided in FP64 by the consumer.
// Candidates are always visited in ascending sorted-slot order, so every per-target sum has one fixed
// order regardless of how targets are grouped into warps or ranks.
#pragma once
#include "sphb_common.cuh"

#ifndef SPHB_LIST
#define SPHB_LIST 128 /* per-lane list capacity (16-bit entries) */
#endif
#ifndef SPHB_SUB
#define SPHB_SUB 8 /* candidates filtered between two overflow checks */
#endif
#ifndef SPHB_SEG_BITS
#define SPHB_SEG_BITS 9   /* list code = (segment slot << 9) | offset in segment */
#endif
#define SPHB_SEG_LEN (1u << SPHB_SEG_BITS)
#define SPHB_NSEG (1 << (16 - SPHB_SEG_BITS)) /* live segments a 16-bit code can address */

#ifndef SPHB_FILTER_MASK
#define SPHB_FILTER_MASK 0 /* pre-filter results collected in a register mask, entries written once per tile: measured slower, off */
#endif
#ifndef SPHB_FILTER_X2
#define SPHB_FILTER_X2 0 /* fp32 pre-filter two candidates per instruction (FFMA2): measured no faster, off */
#endif

#if SPHB_FILTER_MASK && SPHB_FILTER_X2
#error "SPHB_FILTER_X2 uses its own tile layout: build it with -DSPHB_FILTER_MASK=0"
#endif

__device__ __forceinline__ unsigned long long sphb_pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void sphb_unpack2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sphb_fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// LIST: per-lane list capacity of the staging area (the scanning kernels use SPHB_LIST; the list builder of
// sphb_nlist.cu a larger one, so that fewer warps have to flush)
template <int LIST>
struct SphbWarpSharedT {
    static constexpr int kList = LIST;
    float4 tile[32];      // x', y', z', |c'|^2 of the staged candidates (NaN-padded)
    float wthr[32];       // nested grids, SYM: the staged candidates' own squared thresholds (PPM: their radii A)
    float wd[32];         // PPM: the staged candidates' displacement allowances D
    uint32_t segbase[SPHB_NSEG]; // first sorted slot of each live segment
    uint16_t list[LIST * 32];
};
using SphbWarpShared = SphbWarpSharedT<SPHB_LIST>;

// One warp's lists parked in global memory between two kernels that work on the same snapshot (sphb_lists,
// include/sphb200.h): the producer scans once with the symmetric support, the consumer drains without scanning.
struct SphbWarpSaved {
    uint32_t nrows;   // rows of `list` in use = largest per-lane count; SPHB_SAVED_NONE: this warp must scan
    uint32_t nseg;
    uint32_t pad[2];
    uint16_t count[32];
    uint32_t segbase[SPHB_NSEG];
    uint16_t list[SPHB_LIST * 32];
};
#define SPHB_SAVED_NONE 0xffffffffu

__device__ __forceinline__ void sphb_lists_save(const SphbWarpShared &ws, int lane, int cnt, int nseg, bool complete,
                                                SphbWarpSaved *dst)
{
    __syncwarp();
    const int rows = __reduce_max_sync(SPHB_FULL, cnt);
    if (lane == 0) {
        dst->nrows = complete ? (uint32_t)rows : SPHB_SAVED_NONE;
        dst->nseg = (uint32_t)nseg;
    }
    if (!complete) return;
    dst->count[lane] = (uint16_t)cnt;
    for (int s = lane; s < nseg; s += 32) dst->segbase[s] = ws.segbase[s];
    const uint4 *src = reinterpret_cast<const uint4 *>(ws.list);
    uint4 *d = reinterpret_cast<uint4 *>(dst->list);
    for (int c = lane; c < rows * 4; c += 32) d[c] = src[c];
}

// -> the lane's count, or -1 (warp-uniform) when the warp has no parked lists
__device__ __forceinline__ int sphb_lists_load(SphbWarpShared &ws, int lane, const SphbWarpSaved *src)
{
    const uint32_t rows = __ldg(&src->nrows);
    const int nseg = (int)__ldg(&src->nseg);
    if (rows > (uint32_t)SPHB_LIST || nseg > SPHB_NSEG) return -1;   // SPHB_SAVED_NONE, or not a parked record
    for (int s = lane; s < nseg; s += 32) ws.segbase[s] = __ldg(src->segbase + s);
    const uint4 *g = reinterpret_cast<const uint4 *>(src->list);
    uint4 *d = reinterpret_cast<uint4 *>(ws.list);
    for (int c = lane; c < (int)rows * 4; c += 32) d[c] = __ldg(g + c);
    const int cnt = (int)__ldg(src->count + lane);
    __syncwarp();
    return cnt;
}

struct SphbScanCtx {
    const uint32_t *cell_start;
    const float *cell_hmax;
    const float4 *frec;
    int nx, ny, nz;
    double cell, inv_cell;     // y,z edge
    double cellx, inv_cellx;   // x edge = cell / xdiv
    float delta;
    // nested grids (nlev > 1): the fields above are replaced level by level from this table
    int nlev;
    const sphb_level *levels;
    const float *level_hmax;
    double ox, oy, oz;         // origin of the fp32 frame
};

__device__ __forceinline__ SphbScanCtx sphb_make_ctx(const sphb_cells &c)
{
    SphbScanCtx s;
    s.cell_start = c.cell_start;
    s.cell_hmax = c.cell_hmax;
    s.frec = reinterpret_cast<const float4 *>(c.frec);
    s.nx = c.grid.nx;
    s.ny = c.grid.ny;
    s.nz = c.grid.nz;
    s.cell = c.grid.cell;
    s.inv_cell = 1.0 / c.grid.cell;
    s.cellx = sphb_cellx(c.grid);
    s.inv_cellx = 1.0 / s.cellx;
    s.delta = c.delta;
    s.nlev = c.nlevels > 1 ? c.nlevels : 1;
    s.levels = c.levels;
    s.level_hmax = c.level_hmax;
    s.ox = c.grid.ox;
    s.oy = c.grid.oy;
    s.oz = c.grid.oz;
    return s;
}

// one level of the search: the grid to walk, where it sits in the fp32 frame, and how far its particles reach
struct SphbLevelCtx {
    const uint32_t *cell_start;
    const float *cell_hmax;
    int nx, ny, nz;
    double cell, inv_cell, cellx, inv_cellx;
    double sx, sy, sz;   // level origin minus frame origin: frame coordinate - s = coordinate in the level's grid
    bool bounded;        // the grid contains all of its particles: a search that does not touch its box skips it
};

template <bool MULTI = true>
__device__ __forceinline__ SphbLevelCtx sphb_level_ctx(const SphbScanCtx &g, int b)
{
    SphbLevelCtx l;
    if (MULTI && g.nlev > 1) {
        const sphb_level L = sphb_load_level(g.levels, b);
        l.cell_start = g.cell_start + L.cell_base;
        l.cell_hmax = g.cell_hmax + L.cell_base;
        l.nx = L.grid.nx;
        l.ny = L.grid.ny;
        l.nz = L.grid.nz;
        l.cell = l.cellx = L.grid.cell;
        l.inv_cell = l.inv_cellx = 1.0 / L.grid.cell;
        l.sx = L.grid.ox - g.ox;
        l.sy = L.grid.oy - g.oy;
        l.sz = L.grid.oz - g.oz;
        l.bounded = true;
    } else {
        l.cell_start = g.cell_start;
        l.cell_hmax = g.cell_hmax;
        l.nx = g.nx;
        l.ny = g.ny;
        l.nz = g.nz;
        l.cell = g.cell;
        l.inv_cell = g.inv_cell;
        l.cellx = g.cellx;
        l.inv_cellx = g.inv_cellx;
        l.sx = l.sy = l.sz = 0.0;
        l.bounded = false;   // outliers are clamped into border cells: those are unbounded
    }
    return l;
}

// does [lo - R, hi + R] miss the axis range [0, n cell] of a bounded level?
__device__ __forceinline__ bool sphb_axis_miss(double lo, double hi, double R, int n, double cell)
{
    return (hi + R < -1e-6 * cell) || (lo - R > ((double)n + 1e-6) * cell);
}

// lowest / highest cell index that can hold a coordinate >= v / <= v (clamped like the build clamps)
__device__ __forceinline__ int sphb_cell_lo(double v, double inv_cell, int n)
{
    double c = floor(v * inv_cell - 1e-6);
    c = fmin(fmax(c, 0.0), (double)(n - 1));
    return (int)c;
}
__device__ __forceinline__ int sphb_cell_hi(double v, double inv_cell, int n)
{
    double c = floor(v * inv_cell + 1e-6);
    c = fmin(fmax(c, 0.0), (double)(n - 1));
    return (int)c;
}

// distance from the interval [lo, hi] to cell c of an axis with n cells (border cells are unbounded:
// the build clamps outliers into them)
__device__ __forceinline__ double sphb_axis_gap(double lo, double hi, int c, int n, double cell, bool bounded = false)
{
    double g = 0.0;
    if (c > 0 || bounded) g = fmax(g, (double)c * cell - hi);
    if (c < n - 1 || bounded) g = fmax(g, lo - (double)(c + 1) * cell);
    return g;
}

// per-lane list cursor: 32-bit shared-window addresses (a generic 64-bit pointer would cost two more
// integer instructions per accepted candidate)
struct SphbList {
    uint32_t p;    // address of the lane's next free entry (stride 32 entries = 64 bytes)
    uint32_t p0;   // address of list[lane]
    template <class WS>
    __device__ __forceinline__ void init(const WS &ws, int lane)
    {
        p0 = (uint32_t)__cvta_generic_to_shared(ws.list + lane);
        p = p0;
    }
    __device__ __forceinline__ int count() const { return (int)((p - p0) >> 6); }
    __device__ __forceinline__ void reset() { p = p0; }
    __device__ __forceinline__ void push(uint32_t code)
    {
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(p), "h"((uint16_t)code) : "memory");
        p += 64;
    }
};

// decode entry k of the calling lane's list
template <class WS>
__device__ __forceinline__ uint32_t sphb_list_get(const WS &ws, int lane, int k)
{
    const uint32_t code = ws.list[k * 32 + lane];
    return ws.segbase[code >> SPHB_SEG_BITS] + (code & (SPHB_SEG_LEN - 1u));
}

// Per-particle list margins (the list builder only; PPM = true).  Particle i may move by D_i = s_i h_i and grow to
// c_i h_i while the lists are in use; a pair belongs on the lists iff
//       |r_j - r_i|  <=  max(A_j, A_i) + D_j + D_i,        A = 2 c h (+ the fp32 position error)
// which covers every later true pair: r_build <= r_now + |dx_j| + |dx_i| < 2 max(c_j h_j, c_i h_i) + D_j + D_i.  The sum
// form lets a fast particle pay for itself: its partners keep their tight margins.  margins: [n][2] = {c_i, s_i}, or
// NULL for the uniform {cu, su}.  Aj, Dj: the calling lane's own.  dpad = (largest D of the level searched) + (largest
// D of the group of targets): how far beyond its own support a
// candidate may sit and still reach a target (bounds the cells searched) -- taken level by level on nested grids, so
// that the allowance of a large-h record of the rim does not widen the searches of the dense core.
struct SphbPpm {
    const float2 *margins;
    float cu, su;              // uniform margins, or (margins != NULL) upper bounds of every c_i, s_i
    float Aj, Dj;
    const float *level_dmax;   // [levels] largest D_i of each level's records (one entry for a single grid); NULL: su h_max
    const float *cell_dmax;    // [ncell] largest D_i of each cell's records, or NULL (then the level's bound is used)
    float hmax_u;              // single grid: h_max of all records
};

__device__ __forceinline__ void sphb_ppm_of(const SphbPpm &pm, uint32_t i, float w, float &A, float &D)
{
    float c = pm.cu, sk = pm.su;
    if (pm.margins) {
        const float2 m = __ldg(pm.margins + i);
        c = m.x;
        sk = m.y;
    }
    const float rw = __fsqrt_ru(w);                 // >= 2 h + delta (frec.w is that, squared, rounded up); +inf: unbounded h
    A = __fmul_ru(rw, c);
    D = (sk > 0.f) ? __fmul_ru(sk, 0.5f * rw) : 0.f;
}

// One full scan for the calling warp.  Every lane must call it (inactive lanes pass tgt = false).
//   fx,fy,fz : the lane's own fp32 record position;  Rj: its search radius (float, rounded up, may be +inf)
//   thr      : its squared threshold (float, rounded up, may be +inf)
//   cid      : its linear cell id, used to group lanes (0xffffffff for lanes without a particle)
//   SYM      : also accept candidates whose own support reaches the target (frec.w), using RH =
//              radius of the global h_max to bound the cells that can hold such candidates
//   flush(n) : consume entries 0..n-1 of the lane's list (sphb_list_get); called warp-uniformly
// The caller must call flush(L.count()) once more after the scan returns (lists persist across groups).
//
// Lanes are grouped into runs of consecutive slots that share a row of cells AND follow each other closely in x:
// a lane whose cell lies more than one search diameter beyond its predecessor's starts a new group.  (Merging two
// lanes costs the cells between them on every row, splitting costs one more diameter: the break-even gap.  Without
// the rule a warp over a nearly empty row -- a few strays at both ends of a boundary layer -- would scan the whole
// row for all of them.)
//   FLUSH_ROWS: 0 = drain only when a list fills up; n > 0 = also after every n rows of cells; -1 = after every z layer
//              of rows.  Early drains keep the lanes of the warp inside the same few rows while they gather their
//              neighbours' records (shorter reuse distance in L1, fewer distinct lines per gather).
//   symscale : SYM only, >= 1: the candidates' supports are taken as symscale times what the grid stores (lists
//              that have to stay complete while every h may still grow by that factor)
// Returns the number of live segments (what sphb_lists_save needs).
//   MULTI    : the cell list may be a set of nested grids (sphb_cells.nlevels > 1); false compiles the level loop away
//   PPM      : per-particle margins (see SphbPpm): Rj = Aj + Dj + max D, thr is unused, symscale = max c, RH includes dpad
template <bool SYM, int FLUSH_ROWS = 0, bool MULTI = false, bool PPM = false, class WS, class Flush>
__device__ __forceinline__ int sphb_warp_scan(const SphbScanCtx &g, WS &ws, int lane, bool tgt, float fx,
                                              float fy, float fz, float Rj, float thr, uint32_t cid, float RH,
                                              SphbList &L, Flush &&flush, float symscale = 1.f, SphbPpm pm = SphbPpm())
{
    int nseg = 0;   // live entries of ws.segbase (warp-uniform)
    unsigned remaining = __ballot_sync(SPHB_FULL, tgt);
    // group key: number of breaks at or before the lane
    uint32_t gkey;
    {
        // the lane's own level, and its cell within it
        uint32_t local = cid, own_nx = (uint32_t)g.nx;
        float own_inv_cellx = (float)g.inv_cellx, own_RH = SYM ? RH : 0.f;
        int lev = 0;
        if (MULTI && g.nlev > 1) {
            lev = (cid != 0xffffffffu) ? sphb_level_of_cell(g.levels, g.nlev, cid) : g.nlev;
            if (cid != 0xffffffffu) {
                const sphb_level Lo = sphb_load_level(g.levels, lev);
                local = cid - Lo.cell_base;
                own_nx = (uint32_t)Lo.grid.nx;
                own_inv_cellx = (float)(1.0 / Lo.grid.cell);
                // lanes are merged by their OWN radii only: a merged box is walked through every level, and on a finer
                // level the cells between two far-apart lanes of a sparse coarse level can hold the whole dense core
                own_RH = 0.f;
            }
        }
        const uint32_t row = local / own_nx, cx = local - row * own_nx;
        const float rmax = fmaxf(sphb_warp_max(tgt ? Rj : 0.f), own_RH);
        const float gapf = ceilf(2.f * rmax * own_inv_cellx);                    // +inf, NaN -> no x breaks
        const uint32_t gap = (gapf < 1e9f) ? (uint32_t)fmaxf(gapf, 1.f) : 0xffffffffu;
        const uint32_t prow = __shfl_up_sync(SPHB_FULL, row, 1), pcx = __shfl_up_sync(SPHB_FULL, cx, 1);
        const int plev = __shfl_up_sync(SPHB_FULL, lev, 1);
        bool brk = (lane == 0) || (row != prow) || (lev != plev) || (cx > pcx && cx - pcx > gap);
        if (MULTI && g.nlev > 1) {
            // Nested grids: a sparse coarse level puts every lane of a warp into another row of cells, and 32 one-lane
            // groups would each walk their (large) sphere through the fine levels, row by row.  Lanes follow each other
            // in one group while they stay within one search diameter of their predecessor in every coordinate.
            const float px = __shfl_up_sync(SPHB_FULL, fx, 1), py = __shfl_up_sync(SPHB_FULL, fy, 1);
            const float pz = __shfl_up_sync(SPHB_FULL, fz, 1);
            const float far = 2.f * rmax;   // +inf: never
            brk = (lane == 0) || (lev != plev) || !(fabsf(fx - px) <= far && fabsf(fy - py) <= far && fabsf(fz - pz) <= far);
        }
        gkey = (uint32_t)__popc(__ballot_sync(SPHB_FULL, brk) & (0xffffffffu >> (31 - lane)));
    }
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const uint32_t lkey = __shfl_sync(SPHB_FULL, gkey, leader);
        const bool ing = tgt && (gkey == lkey);
        const unsigned grp = __ballot_sync(SPHB_FULL, ing);
        remaining &= ~grp;

        // bounding box and largest radius of the group
        const float inf = CUDART_INF_F;
        const float xmin_f = sphb_warp_min(ing ? fx : inf), xmax_f = sphb_warp_max(ing ? fx : -inf);
        const float ymin_f = sphb_warp_min(ing ? fy : inf), ymax_f = sphb_warp_max(ing ? fy : -inf);
        const float zmin_f = sphb_warp_min(ing ? fz : inf), zmax_f = sphb_warp_max(ing ? fz : -inf);
        const double xmin0 = xmin_f, xmax0 = xmax_f, ymin0 = ymin_f, ymax0 = ymax_f, zmin0 = zmin_f, zmax0 = zmax_f;
        const double Rg0 = sphb_warp_max(ing ? Rj : 0.f);        // PPM: Rj = Aj + Dj
        const float Dg = PPM ? sphb_warp_max(ing ? pm.Dj : 0.f) : 0.f;
        const double Ag = PPM ? (double)sphb_warp_max(ing ? pm.Aj : 0.f) : 0.0;

        // group-local frame: centre of the box; err bounds the fp32 rounding of the re-centred test
        const float ox = 0.5f * (xmin_f + xmax_f), oy = 0.5f * (ymin_f + ymax_f), oz = 0.5f * (zmin_f + zmax_f);
        const float tx = fx - ox, ty = fy - oy, tz = fz - oz;
        const float tt = fmaf(tz, tz, fmaf(ty, ty, tx * tx));
        const float hdx = xmax_f - xmin_f, hdy = ymax_f - ymin_f, hdz = zmax_f - zmin_f;
        const float m2x = -2.f * tx, m2y = -2.f * ty, m2z = -2.f * tz;
#if SPHB_FILTER_X2
        const unsigned long long m2x2 = sphb_pack2(m2x, m2x), m2y2 = sphb_pack2(m2y, m2y), m2z2 = sphb_pack2(m2z, m2z);
#endif

      // nested grids: the levels one after the other (ascending cell_base = ascending sorted slot)
      const int nlev = MULTI ? g.nlev : 1;
      for (int lb = 0; lb < nlev; ++lb) {
        const SphbLevelCtx G = sphb_level_ctx<MULTI>(g, lb);
        double Rg = Rg0;
        float dpad = 0.f;
        const float hmax_b = (MULTI && g.nlev > 1) ? __ldg(g.level_hmax + lb) : (PPM ? pm.hmax_u : 0.f);
        const float *cdm = nullptr;   // the level's slice of cell_dmax
        double Rwide = 0.0;
        if (PPM) {
            const float dl = pm.level_dmax ? __ldg(pm.level_dmax + lb)
                                           : ((pm.su > 0.f) ? __fmul_ru(pm.su, 0.5f * sphb_radius_of_h((double)hmax_b, g.delta)) : 0.f);
            dpad = __fadd_ru(dl, Dg);
            if (pm.cell_dmax) {
                // cells within the targets' own reach are always walked; farther ones only if THEIR records' allowances
                // (or supports) bridge the gap -- the search of a quiet region is not widened by a fast record elsewhere
                cdm = pm.cell_dmax + (G.cell_hmax - g.cell_hmax);
                Rwide = Ag + (double)Dg + (double)dl;   // the farthest any cell's records can matter
            } else {
                Rg = Rg0 + (double)dl;
            }
        }
        double Rw = Rg;
        if (SYM) {
            const float RHb = (PPM || (MULTI && g.nlev > 1))
                                  ? __fadd_ru(sphb_radius_of_h((double)hmax_b * (double)symscale, g.delta), dpad) : RH;
            Rw = fmax(fmax(Rg, Rwide), (double)RHb);
        }
        // the group's box in the level's own coordinates
        const double xmin = xmin0 - G.sx, xmax = xmax0 - G.sx, ymin = ymin0 - G.sy, ymax = ymax0 - G.sy;
        const double zmin = zmin0 - G.sz, zmax = zmax0 - G.sz;
        if (MULTI && G.bounded && (sphb_axis_miss(xmin, xmax, Rw, G.nx, G.cellx) || sphb_axis_miss(ymin, ymax, Rw, G.ny, G.cell) ||
                          sphb_axis_miss(zmin, zmax, Rw, G.nz, G.cell)))
            continue;
        const float dmax = 0.5f * (hdx + hdy + hdz) + (float)Rw + 1.8f * (float)G.cell;   // >= |c'| of anything in range
        const float err = 4e-6f * dmax * dmax;
        // pass  <=>  |c'|^2 - 2 c'.t' <= thr - |t'|^2 + err ; lanes outside the group never pass
        const float thrp = ing ? (thr - tt + err) : -inf;
        const float offp = ing ? (err - tt) : -inf;   // SYM: candidate threshold + offp

        const int cyl = sphb_cell_lo(ymin - Rw, G.inv_cell, G.ny), cyh = sphb_cell_hi(ymax + Rw, G.inv_cell, G.ny);
        const int czl = sphb_cell_lo(zmin - Rw, G.inv_cell, G.nz), czh = sphb_cell_hi(zmax + Rw, G.inv_cell, G.nz);
        const int nry = cyh - cyl + 1;
        const long long nrows = (long long)nry * (long long)(czh - czl + 1);

        for (long long rb = 0; rb < nrows; rb += 32) {
            // ---- every lane resolves one row of cells to a span [s, e) of the sorted array ----------
            uint32_t s = 0, e = 0;
            const long long rr = rb + lane;
            if (rr < nrows) {
                const int ry = cyl + (int)(rr % nry), rz = czl + (int)(rr / nry);
                const double gy = sphb_axis_gap(ymin, ymax, ry, G.ny, G.cell, G.bounded);
                const double gz = sphb_axis_gap(zmin, zmax, rz, G.nz, G.cell, G.bounded);
                const double g2 = gy * gy + gz * gz;
                int lo = 0x7fffffff, hi = -1;
                if (g2 < Rg * Rg) {
                    const double hx = sqrt(Rg * Rg - g2);   // +inf stays +inf
                    lo = sphb_cell_lo(xmin - hx, G.inv_cellx, G.nx);
                    hi = sphb_cell_hi(xmax + hx, G.inv_cellx, G.nx);
                }
                const size_t rowbase = ((size_t)rz * G.ny + ry) * (size_t)G.nx;
                if (SYM && g2 < Rw * Rw) {
                    const double hxw = sqrt(Rw * Rw - g2);
                    const int wl = sphb_cell_lo(xmin - hxw, G.inv_cellx, G.nx);
                    const int wh = sphb_cell_hi(xmax + hxw, G.inv_cellx, G.nx);
                    // extend the tight range outwards to the farthest cell whose h_max reaches the group
                    for (int cx = wl; cx < lo && cx <= wh; ++cx) {
                        const double hm = (double)G.cell_hmax[rowbase + cx] * (double)symscale;
                        const double Rc = (PPM && cdm) ? fmax(Ag, (2.0 * hm * 1.000001 + (double)g.delta) * 1.00001) + (double)Dg + (double)cdm[rowbase + cx] * 1.000001
                                                       : (2.0 * hm * 1.000001 + (double)g.delta) * 1.00001 + (double)dpad;
                        const double gx = sphb_axis_gap(xmin, xmax, cx, G.nx, G.cellx, G.bounded);
                        if (hm > 0.0 && g2 + gx * gx < Rc * Rc) {
                            lo = cx;
                            if (hi < lo) hi = lo;
                            break;
                        }
                    }
                    for (int cx = wh; cx > hi && cx >= wl; --cx) {
                        const double hm = (double)G.cell_hmax[rowbase + cx] * (double)symscale;
                        const double Rc = (PPM && cdm) ? fmax(Ag, (2.0 * hm * 1.000001 + (double)g.delta) * 1.00001) + (double)Dg + (double)cdm[rowbase + cx] * 1.000001
                                                       : (2.0 * hm * 1.000001 + (double)g.delta) * 1.00001 + (double)dpad;
                        const double gx = sphb_axis_gap(xmin, xmax, cx, G.nx, G.cellx, G.bounded);
                        if (hm > 0.0 && g2 + gx * gx < Rc * Rc) {
                            hi = cx;
                            if (lo > hi) lo = hi;
                            break;
                        }
                    }
                }
                if (hi >= lo) {
                    s = G.cell_start[rowbase + lo];
                    e = G.cell_start[rowbase + hi + 1];
                }
            }
            const int nb = (int)((nrows - rb) < 32 ? (nrows - rb) : 32);
            // ---- walk the spans, one segment (<= SPHB_SEG_LEN slots) at a time -----------------------------
            for (int k = 0; k < nb; ++k) {
                if (FLUSH_ROWS != 0) {
                    const long long r = rb + k;
                    const int every = FLUSH_ROWS > 0 ? FLUSH_ROWS : nry;
                    if (r > 0 && r % every == 0 && __any_sync(SPHB_FULL, L.count() > 0)) {
                        flush(L.count());
                        L.reset();
                        nseg = 0;
                    }
                }
                const uint32_t sk = __shfl_sync(SPHB_FULL, s, k), ek = __shfl_sync(SPHB_FULL, e, k);
                for (uint32_t t0 = sk; t0 < ek; t0 += SPHB_SEG_LEN) {
                    if (nseg == SPHB_NSEG) {   // segment table full: drain the lists that reference it
                        flush(L.count());
                        L.reset();
                        nseg = 0;
                    }
                    __syncwarp();
                    if (lane == 0) ws.segbase[nseg] = t0;
                    uint32_t code = (uint32_t)nseg << SPHB_SEG_BITS;
                    ++nseg;
                    const uint32_t t1 = (ek - t0) < SPHB_SEG_LEN ? ek : t0 + SPHB_SEG_LEN;
                    for (uint32_t t = t0; t < t1; t += 32) {
                        const uint32_t ci = t + lane;
                        float4 rec = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f);
                        float wthr = 0.f, ac = 0.f, dc = 0.f;
                        if (ci < t1) {
                            const float4 f = __ldg(g.frec + ci);
                            rec.x = f.x - ox;
                            rec.y = f.y - oy;
                            rec.z = f.z - oz;
                            rec.w = fmaf(rec.z, rec.z, fmaf(rec.y, rec.y, rec.x * rec.x));
                            wthr = f.w;
                            if (PPM) sphb_ppm_of(pm, ci, f.w, ac, dc);
                        }
#if SPHB_FILTER_X2
                        {
                            // pair layout for the packed filter: tile[2p] = (x_a, x_b, y_a, y_b), tile[2p+1] =
                            // (z_a, z_b, w_a, w_b) for candidates a = 2p, b = 2p+1; the even lane takes its partner's
                            // x, y, the odd lane its partner's z, w -- two shuffles, one conflict-free STS.128 each
                            const bool odd = lane & 1;
                            const float r1 = __shfl_xor_sync(SPHB_FULL, odd ? rec.x : rec.z, 1);
                            const float r2 = __shfl_xor_sync(SPHB_FULL, odd ? rec.y : rec.w, 1);
                            ws.tile[lane] = odd ? make_float4(r1, rec.z, r2, rec.w) : make_float4(rec.x, r1, rec.y, r2);
                        }
#else
                        ws.tile[lane] = rec;
#endif
                        // SYM: the candidates' own supports enter through the largest threshold of the tile (one
                        // warp reduction per 32 candidates instead of a load + max per candidate; thresholds are
                        // non-negative floats, so their bit patterns order like integers)
                        float th = thrp;
                        if (PPM) {
                            // the tile's largest radius and allowance bound every candidate's criterion; survivors are
                            // tested again with their own
                            const float at = __int_as_float(__reduce_max_sync(SPHB_FULL, __float_as_int(ac)));
                            const float dt_ = __int_as_float(__reduce_max_sync(SPHB_FULL, __float_as_int(dc)));
                            const float R = __fadd_ru(__fadd_ru(fmaxf(pm.Aj, at), pm.Dj), dt_);
                            th = __fmul_ru(R, R) + offp;      // non-group lanes: offp = -inf (inf - inf = NaN never passes)
                            ws.wthr[lane] = ac;
                            ws.wd[lane] = dc;
                        } else if (SYM) {
                            const float ttile = __int_as_float(__reduce_max_sync(SPHB_FULL, __float_as_int(wthr)));
                            th = fmaxf(thrp, __fmul_ru(ttile, symscale * symscale) + offp);
                            // within one level h spans an octave: under the tile's largest threshold alone a lane would
                            // list up to 8x the volume it needs; survivors are tested again with their own threshold
                            if (MULTI) ws.wthr[lane] = __fmul_ru(wthr, symscale * symscale);
                        }
                        __syncwarp();
                        const int nt = (int)((t1 - t) < 32u ? (t1 - t) : 32u);
#if SPHB_FILTER_MASK
                        // the tests only set bits of a per-lane mask (no shared-memory traffic); the mask becomes
                        // list entries once per tile, in ascending order
                        uint32_t hit = 0;
                        for (int c0 = 0; c0 < nt; c0 += SPHB_SUB) {
                            float4 f[SPHB_SUB];
#pragma unroll
                            for (int c = 0; c < SPHB_SUB; ++c) f[c] = ws.tile[c0 + c];
                            float sv[SPHB_SUB];
#pragma unroll
                            for (int c = 0; c < SPHB_SUB; ++c)
                                sv[c] = fmaf(f[c].x, m2x, fmaf(f[c].y, m2y, fmaf(f[c].z, m2z, f[c].w)));
                            uint32_t m = 0;
#pragma unroll
                            for (int c = 0; c < SPHB_SUB; ++c) m |= (sv[c] <= th) ? (1u << c) : 0u;
                            hit |= m << c0;
                        }
                        if (__any_sync(SPHB_FULL, L.count() + __popc(hit) > WS::kList)) {
                            flush(L.count());
                            L.reset();
                        }
                        while (hit) {
                            L.push(code + (uint32_t)(__ffs(hit) - 1));
                            hit &= hit - 1;
                        }
                        code += 32;
#else
                        for (int c0 = 0; c0 < nt; c0 += SPHB_SUB) {
                            if (__any_sync(SPHB_FULL, L.count() > WS::kList - SPHB_SUB)) {
                                flush(L.count());
                                L.reset();
                            }
#if SPHB_FILTER_X2
                            // two candidates per FFMA2 (fma.rn.f32x2, sm_100): same roundings as the scalar form
                            const ulonglong2 *t2 = reinterpret_cast<const ulonglong2 *>(ws.tile) + c0;
                            ulonglong2 xy[SPHB_SUB / 2], zw[SPHB_SUB / 2];
#pragma unroll
                            for (int q = 0; q < SPHB_SUB / 2; ++q) {
                                xy[q] = t2[2 * q];
                                zw[q] = t2[2 * q + 1];
                            }
                            unsigned long long sv2[SPHB_SUB / 2];
#pragma unroll
                            for (int q = 0; q < SPHB_SUB / 2; ++q)
                                sv2[q] = sphb_fma2(xy[q].x, m2x2, sphb_fma2(xy[q].y, m2y2, sphb_fma2(zw[q].x, m2z2, zw[q].y)));
#pragma unroll
                            for (int q = 0; q < SPHB_SUB / 2; ++q) {
                                float sa, sb;
                                sphb_unpack2(sv2[q], sa, sb);
                                if (sa <= th) L.push(code + 2 * q);
                                if (sb <= th) L.push(code + 2 * q + 1);
                            }
#else
                            // all loads of the sub-tile first (independent LDS.128), then the tests
                            float4 f[SPHB_SUB];
#pragma unroll
                            for (int c = 0; c < SPHB_SUB; ++c) f[c] = ws.tile[c0 + c];
                            float sv[SPHB_SUB];
#pragma unroll
                            for (int c = 0; c < SPHB_SUB; ++c)
                                sv[c] = fmaf(f[c].x, m2x, fmaf(f[c].y, m2y, fmaf(f[c].z, m2z, f[c].w)));
#pragma unroll
                            for (int c = 0; c < SPHB_SUB; ++c)
                                if (sv[c] <= th) {
                                    if (PPM) {
                                        const float Rc = __fadd_ru(__fadd_ru(fmaxf(pm.Aj, ws.wthr[c0 + c]), pm.Dj), ws.wd[c0 + c]);
                                        if (sv[c] <= __fmul_ru(Rc, Rc) + offp) L.push(code + c);
                                    } else if (!(SYM && MULTI) || sv[c] <= fmaxf(thrp, ws.wthr[c0 + c] + offp)) {
                                        L.push(code + c);
                                    }
                                }
#endif
                            code += SPHB_SUB;
                        }
#endif
                        __syncwarp();
                    }
                }
            }
        }
      }   // levels
    }
    return nseg;
}
