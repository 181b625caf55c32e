// sphb_scan.cuh -- the candidate scan shared by every gather kernel (K2 density/omega, K3 Newton,
// K4 force, neighbour sets).
//
// Mapping: one lane = one target particle, one warp = 32 consecutive slots of the cell-sorted
// particle array.  For each group of lanes that share a cell row, the warp
//   1. takes the group's bounding box + search radius and turns it into rows of cells; every row is ONE
//      contiguous span of the sorted array (x is the fastest cell axis), found by all lanes in parallel;
//   2. stages the span's fp32 prefilter records through a 32-entry shared-memory tile (coalesced
//      float4 loads) and lets every lane test every staged candidate in FP32 (broadcast LDS);
//   3. appends survivors to a per-lane index list in shared memory ([k][lane]: conflict free); when a
//      list is about to overflow, or at the end, `flush(cnt)` evaluates the listed pairs in FP64.
// The FP32 test only has to be conservative (positions carry an absolute error <= delta, thresholds are
// rounded up); membership is decided in FP64 by the consumer.  Candidates are always visited in
// ascending sorted-slot order, so every per-target sum has one fixed order regardless of how targets
// are grouped into warps or ranks.
#pragma once
#include "sphb_common.cuh"

#ifndef SPHB_LIST
#define SPHB_LIST 64 /* per-lane list capacity */
#endif
#define SPHB_SUB 8 /* candidates filtered between two overflow checks */

struct SphbWarpShared {
    float4 tile[32];
    uint32_t list[SPHB_LIST * 32];
};

struct SphbScanCtx {
    const uint32_t *cell_start;
    const float *cell_hmax;
    const float4 *frec;
    int nx, ny, nz;
    double cell, inv_cell;
    float delta;
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
    s.delta = c.delta;
    return s;
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
__device__ __forceinline__ double sphb_axis_gap(double lo, double hi, int c, int n, double cell)
{
    double g = 0.0;
    if (c > 0) g = fmax(g, (double)c * cell - hi);
    if (c < n - 1) g = fmax(g, lo - (double)(c + 1) * cell);
    return g;
}

// One full scan for the calling warp.  Every lane must call it (inactive lanes pass tgt = false).
//   fx,fy,fz : the lane's own fp32 record position;  Rj: its search radius (float, rounded up, may be +inf)
//   thr      : its squared threshold (float, rounded up, may be +inf)
//   row      : its cell row id (cell_id / nx), used to group lanes
//   SYM      : also accept candidates whose own support reaches the target (frec.w), using RH =
//              radius of the global h_max to bound the cells that can hold such candidates
//   cnt      : per-lane list fill, in/out;  flush(cnt) must consume ws.list[k*32+lane], k < cnt
template <bool SYM, class Flush>
__device__ __forceinline__ void sphb_warp_scan(const SphbScanCtx &g, SphbWarpShared &ws, int lane, bool tgt, float fx,
                                               float fy, float fz, float Rj, float thr, uint32_t row, float RH,
                                               int &cnt, Flush &&flush)
{
    unsigned remaining = __ballot_sync(SPHB_FULL, tgt);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const uint32_t lrow = __shfl_sync(SPHB_FULL, row, leader);
        const bool ing = tgt && (row == lrow);
        const unsigned grp = __ballot_sync(SPHB_FULL, ing);
        remaining &= ~grp;

        // bounding box and largest radius of the group
        const float inf = CUDART_INF_F;
        const double xmin = sphb_warp_min(ing ? fx : inf), xmax = sphb_warp_max(ing ? fx : -inf);
        const double ymin = sphb_warp_min(ing ? fy : inf), ymax = sphb_warp_max(ing ? fy : -inf);
        const double zmin = sphb_warp_min(ing ? fz : inf), zmax = sphb_warp_max(ing ? fz : -inf);
        const double Rg = sphb_warp_max(ing ? Rj : 0.f);
        const double Rw = SYM ? fmax(Rg, (double)RH) : Rg;
        const float thr_l = ing ? thr : -1.f;

        const int cyl = sphb_cell_lo(ymin - Rw, g.inv_cell, g.ny), cyh = sphb_cell_hi(ymax + Rw, g.inv_cell, g.ny);
        const int czl = sphb_cell_lo(zmin - Rw, g.inv_cell, g.nz), czh = sphb_cell_hi(zmax + Rw, g.inv_cell, g.nz);
        const int nry = cyh - cyl + 1;
        const long long nrows = (long long)nry * (long long)(czh - czl + 1);

        for (long long rb = 0; rb < nrows; rb += 32) {
            // ---- every lane resolves one row of cells to a span [s, e) of the sorted array ----------
            uint32_t s = 0, e = 0;
            const long long rr = rb + lane;
            if (rr < nrows) {
                const int ry = cyl + (int)(rr % nry), rz = czl + (int)(rr / nry);
                const double gy = sphb_axis_gap(ymin, ymax, ry, g.ny, g.cell);
                const double gz = sphb_axis_gap(zmin, zmax, rz, g.nz, g.cell);
                const double g2 = gy * gy + gz * gz;
                int lo = 0x7fffffff, hi = -1;
                if (g2 < Rg * Rg) {
                    const double hx = sqrt(Rg * Rg - g2);   // +inf stays +inf
                    lo = sphb_cell_lo(xmin - hx, g.inv_cell, g.nx);
                    hi = sphb_cell_hi(xmax + hx, g.inv_cell, g.nx);
                }
                const size_t rowbase = ((size_t)rz * g.ny + ry) * (size_t)g.nx;
                if (SYM && g2 < Rw * Rw) {
                    const double hxw = sqrt(Rw * Rw - g2);
                    const int wl = sphb_cell_lo(xmin - hxw, g.inv_cell, g.nx);
                    const int wh = sphb_cell_hi(xmax + hxw, g.inv_cell, g.nx);
                    // extend the tight range outwards to the farthest cell whose h_max reaches the group
                    for (int cx = wl; cx < lo && cx <= wh; ++cx) {
                        const double hm = (double)g.cell_hmax[rowbase + cx];
                        const double Rc = (2.0 * hm * 1.000001 + (double)g.delta) * 1.00001;
                        const double gx = sphb_axis_gap(xmin, xmax, cx, g.nx, g.cell);
                        if (hm > 0.0 && g2 + gx * gx < Rc * Rc) {
                            lo = cx;
                            if (hi < lo) hi = lo;
                            break;
                        }
                    }
                    for (int cx = wh; cx > hi && cx >= wl; --cx) {
                        const double hm = (double)g.cell_hmax[rowbase + cx];
                        const double Rc = (2.0 * hm * 1.000001 + (double)g.delta) * 1.00001;
                        const double gx = sphb_axis_gap(xmin, xmax, cx, g.nx, g.cell);
                        if (hm > 0.0 && g2 + gx * gx < Rc * Rc) {
                            hi = cx;
                            if (lo > hi) lo = hi;
                            break;
                        }
                    }
                }
                if (hi >= lo) {
                    s = g.cell_start[rowbase + lo];
                    e = g.cell_start[rowbase + hi + 1];
                }
            }
            const int nb = (int)((nrows - rb) < 32 ? (nrows - rb) : 32);
            // ---- walk the spans ---------------------------------------------------------------------
            for (int k = 0; k < nb; ++k) {
                const uint32_t sk = __shfl_sync(SPHB_FULL, s, k), ek = __shfl_sync(SPHB_FULL, e, k);
                for (uint32_t t = sk; t < ek; t += 32) {
                    const uint32_t ci = t + lane;
                    if (ci < ek) ws.tile[lane] = __ldg(g.frec + ci);
                    __syncwarp();
                    const int nt = (int)((ek - t) < 32u ? (ek - t) : 32u);
                    for (int c0 = 0; c0 < nt; c0 += SPHB_SUB) {
                        if (__any_sync(SPHB_FULL, cnt > SPHB_LIST - SPHB_SUB)) {
                            flush(cnt);
                            cnt = 0;
                        }
#pragma unroll
                        for (int c = 0; c < SPHB_SUB; ++c) {
                            if (c0 + c < nt) {
                                const float4 f = ws.tile[c0 + c];
                                const float dx = f.x - fx, dy = f.y - fy, dz = f.z - fz;
                                const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                                const float th = SYM ? fmaxf(thr_l, f.w) : thr_l;
                                if (ing && d2 <= th) {
                                    ws.list[cnt * 32 + lane] = t + (uint32_t)(c0 + c);
                                    ++cnt;
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
}
