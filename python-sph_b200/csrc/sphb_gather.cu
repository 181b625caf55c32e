// sphb_gather.cu -- K2 (M4 density + grad-h sums), K3 (Newton-Raphson h solve + bisection fallback)
// and the exact neighbour-set kernel.  See include/sphb200.h for the reference lines each replaces.
#include <stdio.h>

#include "sphb_scan.cuh"
#include "sphb_pair.cuh"

#define SPHB_WARPS 4
#ifndef SPHB_MINB_OWN
#define SPHB_MINB_OWN 4
#endif
#define SPHB_THREADS (SPHB_WARPS * 32)
#ifndef SPHB_OWN_FLUSH_ROWS
#define SPHB_OWN_FLUSH_ROWS 0
#endif

#ifndef SPHB_OWN_ILP
#define SPHB_OWN_ILP 1
#endif

// drain the lane's list, two entries at a time so that two independent FP64 chains are in flight per lane
// (SPHB_OWN_ILP 1: the records of the next two entries are loaded while the current two are evaluated;
//  2: no look-ahead, fewer registers; 0: one entry at a time with one record of look-ahead)
template <bool XI = false>
__device__ __forceinline__ void sphb_own_flush(const double *posh, const SphbWarpShared &ws, int lane, int cnt,
                                               SphbOwnAcc &a, const double *mass = nullptr)
{
#if SPHB_OWN_ILP == 1
    double A[4], B[4], C[4], D[4];
    A[0] = A[1] = A[2] = B[0] = B[1] = B[2] = 0.0;
    if (cnt > 0) sphb_load_posw(posh, mass, sphb_list_get(ws, lane, 0), A[0], A[1], A[2], A[3]);
    if (cnt > 1) sphb_load_posw(posh, mass, sphb_list_get(ws, lane, 1), B[0], B[1], B[2], B[3]);
    int k = 0;
    while (k < cnt) {
        C[0] = C[1] = C[2] = D[0] = D[1] = D[2] = 0.0;
        if (k + 2 < cnt) sphb_load_posw(posh, mass, sphb_list_get(ws, lane, k + 2), C[0], C[1], C[2], C[3]);
        if (k + 3 < cnt) sphb_load_posw(posh, mass, sphb_list_get(ws, lane, k + 3), D[0], D[1], D[2], D[3]);
        sphb_own_pair2<XI, true>(A, B, k + 1 < cnt, a);
        k += 2;
        if (k >= cnt) break;
        if (k + 2 < cnt) sphb_load_posw(posh, mass, sphb_list_get(ws, lane, k + 2), A[0], A[1], A[2], A[3]);
        if (k + 3 < cnt) sphb_load_posw(posh, mass, sphb_list_get(ws, lane, k + 3), B[0], B[1], B[2], B[3]);
        sphb_own_pair2<XI, true>(C, D, k + 1 < cnt, a);
        k += 2;
    }
#elif SPHB_OWN_ILP == 2
    for (int k = 0; k < cnt; k += 2) {
        double A[4], B[4];
        B[0] = B[1] = B[2] = 0.0;
        sphb_load_posw(posh, mass, sphb_list_get(ws, lane, k), A[0], A[1], A[2], A[3]);
        if (k + 1 < cnt) sphb_load_posw(posh, mass, sphb_list_get(ws, lane, k + 1), B[0], B[1], B[2], B[3]);
        sphb_own_pair2<XI, true>(A, B, k + 1 < cnt, a);
    }
#else
    double ax, ay, az, ah, bx, by, bz, bh;
    if (cnt > 0) sphb_load_posw(posh, mass, sphb_list_get(ws, lane, 0), ax, ay, az, ah);
    int k = 0;
    while (k < cnt) {
        if (k + 1 < cnt) sphb_load_posw(posh, mass, sphb_list_get(ws, lane, k + 1), bx, by, bz, bh);
        sphb_own_pair<XI>(ax, ay, az, a, ah);
        ++k;
        if (k >= cnt) break;
        if (k + 1 < cnt) sphb_load_posw(posh, mass, sphb_list_get(ws, lane, k + 1), ax, ay, az, ah);
        sphb_own_pair<XI>(bx, by, bz, a, bh);
        ++k;
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------------
// PRODUCE: scan with the symmetric support and park the lists for sphb_force (sphb_lists, include/sphb200.h)
template <bool XI, bool PRODUCE, bool MULTI>
__global__ void __launch_bounds__(SPHB_THREADS, SPHB_MINB_OWN)
    k_density_omega(sphb_cells c, sphb_consts k, const uint8_t *__restrict__ target, const double *__restrict__ h_target,
                    const double *__restrict__ rho_in, double *__restrict__ rho_sum, double *__restrict__ dwdh_sum,
                    double *__restrict__ omega, double *__restrict__ xi_out, int32_t *__restrict__ ngb_count,
                    sphb_lists lists)
{
    __shared__ SphbWarpShared sh[SPHB_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long base = ((long long)blockIdx.x * SPHB_WARPS + warp) * 32;
    if (base >= c.n) return;
    const long long j = base + lane;
    const bool valid = j < c.n;
    const bool tgt = valid && (target == nullptr || target[j] != 0);
    const SphbScanCtx g = sphb_make_ctx(c);

    SphbOwnAcc a;
    double h = 1.0;
    a.xj = a.yj = a.zj = 0.0;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t cid = 0xffffffffu;
    if (valid) {
        sphb_load_posh(c.posh, (size_t)j, a.xj, a.yj, a.zj, h);
        if (h_target) h = h_target[j];
        f = __ldg(g.frec + j);
        cid = c.cell_id[j];
    }
    a.hinv = 1.0 / h;
    a.sumF = a.sumG = a.sumX = 0.0;
    a.nn = 0;
    const float Rj = sphb_radius_of_h(h, g.delta), thr = sphb_thr_of_h(h, g.delta);
    SphbWarpShared &ws = sh[warp];
    SphbList L;
    L.init(ws, lane);
    bool flushed = false;   // warp-uniform
    auto flush = [&](int n_) {
        flushed = true;
        sphb_own_flush<XI>(c.posh, ws, lane, n_, a, k.mass);
    };
    if (PRODUCE) {
        const float cov = (float)lists.cover * 1.000001f;
        const float RH = sphb_radius_of_h((double)lists.hmax_global[0] * (double)cov, g.delta);
        const int nseg = sphb_warp_scan<true, 0, MULTI>(g, ws, lane, tgt, f.x, f.y, f.z, Rj, thr, cid, RH, L, flush, cov);
        sphb_lists_save(ws, lane, L.count(), nseg, !flushed,
                        reinterpret_cast<SphbWarpSaved *>(lists.store) + ((long long)blockIdx.x * SPHB_WARPS + warp));
    } else {
        sphb_warp_scan<false, SPHB_OWN_FLUSH_ROWS, MULTI>(g, ws, lane, tgt, f.x, f.y, f.z, Rj, thr, cid, 0.f, L, flush);
    }
    sphb_own_flush<XI>(c.posh, ws, lane, L.count(), a, k.mass);
    if (tgt) {
        const double pref = sphb_mass_after(k) / (SPHB_PI * (h * h * h));
        const double rs = pref * a.sumF;
        const double ds = -(pref / h) * a.sumG;
        if (rho_sum) rho_sum[j] = rs;
        if (dwdh_sum) dwdh_sum[j] = ds;
        const double rho = rho_in ? rho_in[j] : sphb_var_density_m(k, sphb_mass_of(k, (size_t)j), h);
        if (omega) omega[j] = 1.0 + h / (3.0 * rho) * ds;   // pyfluid.py:97
        if (XI) {
            // xi_j = -h/(3 rho) sum_i m dphi/dh = -h/(3 rho) * (-(m/h^2) sum pX), pyfluid.py:401-410
            const double sx = sphb_mass_after(k) * (-(1.0 / (h * h)) * a.sumX);
            xi_out[j] = -h / (3.0 * rho) * sx;
        }
        if (ngb_count) ngb_count[j] = a.nn;
    }
}

// ------------------------------------------------------------------------------------------------
// K3: per-lane Newton loop with warp-vote convergence mask (pyfluid.py:155-179).
// The cell scan runs at the guess with a small radius margin and leaves each lane a neighbour list that covers
// every h up to (1 + margin) * guess.  Each further evaluation the lane needs -- the sums of the next Newton
// iterate, or (omega_out != NULL) the sum m dW/dh at the converged h that Omega is made of (pyfluid.py:91-97 with
// the algebraic density, as the step uses it at :503/:518) -- is one more pass over the SAME list while the new h
// stays inside that cover; entries beyond the support of the new h add exact zeros.  Only a lane whose h leaves
// the cover (cold solves, bad guesses), or whose list overflowed, makes the warp scan again.
// ------------------------------------------------------------------------------------------------
#ifndef SPHB_NEWTON_MARGIN
#define SPHB_NEWTON_MARGIN 4e-3
#endif

template <bool FUSED, bool PRODUCE, bool MULTI>
__global__ void __launch_bounds__(SPHB_THREADS, SPHB_MINB_OWN)
    k_newton_h(sphb_cells c, sphb_consts k, const uint8_t *__restrict__ target, const double *__restrict__ h_guess,
               double *__restrict__ h_out, double *__restrict__ omega_out, int32_t *__restrict__ iters,
               int32_t *__restrict__ code, uint32_t *__restrict__ fail_list, sphb_status *__restrict__ status,
               sphb_lists lists)
{
    __shared__ SphbWarpShared sh[SPHB_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long base = ((long long)blockIdx.x * SPHB_WARPS + warp) * 32;
    if (base >= c.n) return;
    const long long j = base + lane;
    const bool valid = j < c.n;
    const bool tgt = valid && (target == nullptr || target[j] != 0);
    const SphbScanCtx g = sphb_make_ctx(c);
    SphbWarpShared &ws = sh[warp];
    const double rmargin = PRODUCE ? lists.cover : 1.0 + SPHB_NEWTON_MARGIN;
    bool first_scan = true;   // PRODUCE: the lists of the first scan (every target at its guess) are the ones parked

    SphbOwnAcc a;
    a.xj = a.yj = a.zj = 0.0;
    double h_cur = 1.0, hdummy;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t cid = 0xffffffffu;
    if (valid) {
        sphb_load_posh(c.posh, (size_t)j, a.xj, a.yj, a.zj, hdummy);
        f = __ldg(g.frec + j);
        cid = c.cell_id[j];
    }
    if (tgt) h_cur = h_guess[j];
    // what the lane needs next: 1 = the sums of a Newton iterate at h_cur, 2 = sum m dW/dh at the converged h_cur
    // (FUSED), 0 = nothing
    int want = tgt ? 1 : 0;
    double h_cover = 0.0;    // the lane's list holds every neighbour of any h <= h_cover
    bool have_lists = false; // warp-uniform: the lists of the last scan are intact (no overflow flush)
    int cnt = 0, it = 0;
    if (PRODUCE && !__any_sync(SPHB_FULL, tgt))   // a warp without targets never scans: park empty lists
        sphb_lists_save(ws, lane, 0, 0, true,
                        reinterpret_cast<SphbWarpSaved *>(lists.store) + ((long long)blockIdx.x * SPHB_WARPS + warp));
#ifdef SPHB_NEWTON_TRACE
    const long long trace_t0 = clock64();
    int trace_scans = 0, trace_rounds = 0;
#endif
    while (__any_sync(SPHB_FULL, want != 0)) {
        const bool pend = want != 0;
        a.hinv = 1.0 / h_cur;
        a.sumF = a.sumG = a.sumX = 0.0;
        a.nn = 0;
        const bool covered = have_lists && h_cur > 0.0 && h_cur <= h_cover;
        if (__any_sync(SPHB_FULL, pend && !covered)) {
            // scan (again) for every lane that still needs something
#ifdef SPHB_NEWTON_TRACE
            ++trace_scans;
#endif
            const double hs = h_cur * rmargin;
            const float Rj = sphb_radius_of_h(hs, g.delta), thr = sphb_thr_of_h(hs, g.delta);
            SphbList L;
            L.init(ws, lane);
            bool flushed = false;   // warp-uniform: did a list overflow during this scan?
            auto flush = [&](int n_) {
                flushed = true;
                sphb_own_flush(c.posh, ws, lane, n_, a, k.mass);
            };
            if (PRODUCE && first_scan) {
                // symmetric support, every h taken at its cover: complete for sphb_force at any h below the cover
                const float cov = (float)lists.cover * 1.000001f;
                const float RH = sphb_radius_of_h((double)lists.hmax_global[0] * (double)cov, g.delta);
                const int nseg = sphb_warp_scan<true, 0, MULTI>(g, ws, lane, pend, f.x, f.y, f.z, Rj, thr, cid, RH, L, flush, cov);
                sphb_lists_save(ws, lane, L.count(), nseg, !flushed,
                                reinterpret_cast<SphbWarpSaved *>(lists.store) + ((long long)blockIdx.x * SPHB_WARPS + warp));
                first_scan = false;
            } else {
                sphb_warp_scan<false, 0, MULTI>(g, ws, lane, pend, f.x, f.y, f.z, Rj, thr, cid, 0.f, L, flush);
            }
            cnt = L.count();   // 0 for lanes that were not scanned for
            have_lists = !flushed;
            h_cover = pend ? hs : 0.0;
        }
        sphb_own_flush(c.posh, ws, lane, pend ? cnt : 0, a, k.mass);
        if (want == 1) {
            // newton_smoothlength_iteration, pyfluid.py:155-163 (zetaprime with the rho-free identity)
            const double pref = sphb_mass_after(k) / (SPHB_PI * (h_cur * h_cur * h_cur));
            const double rho = pref * a.sumF;
            const double sdh = -(pref / h_cur) * a.sumG;
            const double vd = sphb_var_density_m(k, sphb_mass_of(k, (size_t)j), h_cur);
            const double zeta = vd - rho;
            const double zetap = -sdh - 3.0 * vd / h_cur;
            const double h_new = h_cur - zeta / zetap;
            ++it;
            // smoothlength_variation(old, new) = |old - new| / |new|, pyfluid.py:119-120,172
            if (fabs(h_cur - h_new) / fabs(h_new) < k.h_tol && h_new > 0.0) {
                h_out[j] = h_new;
                if (iters) iters[j] = it;
                if (code) code[j] = 0;
                want = FUSED ? 2 : 0;
                if (PRODUCE && !(h_new <= h_guess[j] * lists.cover)) atomicOr(lists.broken, 1);
            } else if (it >= k.newton_limit) {
                // "WARNING: Failure to converge in newton_new_h." -> bisection_h(j, pos), pyfluid.py:178-179
                atomicAdd(&status->newton_warn, 1);
                const int slot = atomicAdd(&status->nfail, 1);
                fail_list[slot] = (uint32_t)j;
                h_out[j] = CUDART_NAN;
                if (FUSED) omega_out[j] = CUDART_NAN;
                if (iters) iters[j] = it;
                if (code) code[j] = 1;
                want = 0;
                if (PRODUCE) atomicOr(lists.broken, 1);   // the bisection may return any h
            }
            h_cur = h_new;
        } else if (want == 2) {
            omega_out[j] = sphb_omega_from_sumG(k, h_cur, a.sumG, sphb_mass_of(k, (size_t)j), sphb_mass_after(k));
            want = 0;
        }
        __syncwarp();
#ifdef SPHB_NEWTON_TRACE
        ++trace_rounds;
#endif
    }
#ifdef SPHB_NEWTON_TRACE
    // tuning aid (scripts/newton_phase.py): per-warp cost and what it was spent on, in place of iters / code
    if (valid && iters && code) {
        iters[j] = it | (trace_scans << 8) | (trace_rounds << 16) | (cnt << 24 & 0x7f000000);
        code[j] = (int)((clock64() - trace_t0) >> 4);
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// block-wide brute-force density at an arbitrary h for ONE particle (bisection path).  All threads
// of the block call it; returns the same value in every thread.
// ------------------------------------------------------------------------------------------------
__device__ double sphb_block_density(const sphb_cells &c, const sphb_consts &k, double xj, double yj, double zj,
                                     double fxj, double fyj, double fzj, double h, double *red /*[2*SPHB_WARPS]*/,
                                     double *sumG_out = nullptr)
{
    const SphbScanCtx gc = sphb_make_ctx(c);
    const double R = (h > 0.0) ? (2.0 * h * 1.000001 + (double)gc.delta) * 1.00001 : (double)CUDART_INF_F;
    const double hinv = 1.0 / h;
    double sum = 0.0, sumg = 0.0;
  for (int lb = 0; lb < gc.nlev; ++lb) {   // nested grids: level by level, ascending slots
    const SphbLevelCtx g = sphb_level_ctx(gc, lb);
    const double fxl = fxj - g.sx, fyl = fyj - g.sy, fzl = fzj - g.sz;
    if (g.bounded && (sphb_axis_miss(fxl, fxl, R, g.nx, g.cellx) || sphb_axis_miss(fyl, fyl, R, g.ny, g.cell) ||
                      sphb_axis_miss(fzl, fzl, R, g.nz, g.cell)))
        continue;
    const int cxl = sphb_cell_lo(fxl - R, g.inv_cellx, g.nx), cxh = sphb_cell_hi(fxl + R, g.inv_cellx, g.nx);
    const int cyl = sphb_cell_lo(fyl - R, g.inv_cell, g.ny), cyh = sphb_cell_hi(fyl + R, g.inv_cell, g.ny);
    const int czl = sphb_cell_lo(fzl - R, g.inv_cell, g.nz), czh = sphb_cell_hi(fzl + R, g.inv_cell, g.nz);
    const bool fullx = (cxl == 0 && cxh == g.nx - 1);
    const bool fully = fullx && (cyl == 0 && cyh == g.ny - 1);
    // spans: merge rows when the x range (and the y range) cover the whole axis
    const int nzs = fully ? 1 : (czh - czl + 1);
    const int nys = fullx ? 1 : (cyh - cyl + 1);
    for (int iz = 0; iz < nzs; ++iz) {
        for (int iy = 0; iy < nys; ++iy) {
            size_t c0, c1;
            if (fully) {
                c0 = ((size_t)czl * g.ny) * g.nx;
                c1 = ((size_t)(czh + 1) * g.ny) * g.nx;
            } else if (fullx) {
                c0 = ((size_t)(czl + iz) * g.ny + cyl) * g.nx;
                c1 = ((size_t)(czl + iz) * g.ny + cyh + 1) * g.nx;
            } else {
                const size_t rowbase = ((size_t)(czl + iz) * g.ny + (cyl + iy)) * g.nx;
                c0 = rowbase + cxl;
                c1 = rowbase + cxh + 1;
            }
            const uint32_t s = g.cell_start[c0], e = g.cell_start[c1];
            for (uint32_t i = s + threadIdx.x; i < e; i += blockDim.x) {
                double x, y, z, hh;
                sphb_load_posh(c.posh, i, x, y, z, hh);
                const double dx = xj - x, dy = yj - y, dz = zj - z;
                const double d2 = sphb_d2(dx, dy, dz);
                const double q = sqrt(d2) * hinv;
                double f, fp;
                sphb_m4(q, f, fp);
                const double w = sphb_mass_w(k, i);
                sum += w * f;
                sumg += w * fma(q, fp, 3.0 * f);
            }
        }
    }
  }
    sum = sphb_warp_sum(sum);
    sumg = sphb_warp_sum(sumg);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = sum;
        red[SPHB_WARPS + (threadIdx.x >> 5)] = sumg;
    }
    __syncthreads();
    double tot = 0.0, totg = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        tot += red[w];
        totg += red[SPHB_WARPS + w];
    }
    if (sumG_out) *sumG_out = totg;
    return sphb_mass_after(k) / (SPHB_PI * (h * h * h)) * tot;
}

// bisection_h, pyfluid.py:123-153 (density at h_a / h_b is re-used instead of recomputed: same value)
__global__ void __launch_bounds__(SPHB_THREADS)
    k_bisect_h(sphb_cells c, sphb_consts k, const uint32_t *__restrict__ list, long long n_list_arg,
               double *__restrict__ h_out, double *__restrict__ omega_out, int32_t *__restrict__ code,
               int32_t *__restrict__ oob_out, int indexed_by_slot, sphb_status *__restrict__ status)
{
    __shared__ double red[2 * SPHB_WARPS];
    const long long n_list = n_list_arg >= 0 ? n_list_arg : (long long)status->nfail;
    for (long long e = blockIdx.x; e < n_list; e += gridDim.x) {
        const uint32_t j = list[e];
        double xj, yj, zj, hj;
        sphb_load_posh(c.posh, j, xj, yj, zj, hj);
        const float4 f = __ldg(reinterpret_cast<const float4 *>(c.frec) + j);
        double h_a = k.bisect_a, h_b = k.bisect_b;
        const double mj = sphb_mass_of(k, (size_t)j);
        double zeta_a = sphb_var_density_m(k, mj, h_a) - sphb_block_density(c, k, xj, yj, zj, f.x, f.y, f.z, h_a, red);
        double zeta_b = sphb_var_density_m(k, mj, h_b) - sphb_block_density(c, k, xj, yj, zj, f.x, f.y, f.z, h_b, red);
        int i = 0, oob = 0, res = 2;
        double root = CUDART_NAN;
        while (i < k.bisect_limit) {
            if (zeta_a * zeta_b > 0.0) ++oob;   // "ERROR: Root out of bisection bounds."
            const double cand = (h_a + h_b) / 2.0;
            const double zeta_root =
                sphb_var_density_m(k, mj, cand) - sphb_block_density(c, k, xj, yj, zj, f.x, f.y, f.z, cand, red);
            if (fabs(zeta_root) < k.bisect_tol) {
                root = cand;
                res = 1;
                break;
            }
            if (zeta_a * zeta_root < 0.0) {
                h_b = cand;
                zeta_b = zeta_root;
            } else if (zeta_b * zeta_root < 0.0) {
                h_a = cand;
                zeta_a = zeta_root;
            }
            ++i;
        }
        double om = CUDART_NAN;
        if (omega_out && res == 1) {   // Omega at the bisection root (algebraic density, like the step)
            double sg = 0.0;
            sphb_block_density(c, k, xj, yj, zj, f.x, f.y, f.z, root, red, &sg);
            om = sphb_omega_from_sumG(k, root, sg, mj, sphb_mass_after(k));
        }
        if (threadIdx.x == 0) {
            const long long o = indexed_by_slot ? (long long)j : e;
            h_out[o] = root;
            if (omega_out) omega_out[o] = om;
            if (code) code[o] = res;
            if (oob_out) oob_out[o] = oob;
            if (oob) atomicAdd(&status->bisect_oob, oob);
            if (res == 2) {
                atomicAdd(&status->bisect_none, 1);   // "ERROR: Failure to converge in Bisection_new_h."
                atomicOr(&status->flags, SPHB_ST_BISECT_NONE);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// exact neighbour sets: one warp per query slot, reference arithmetic
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPHB_THREADS)
    k_neighbours(sphb_cells c, const uint32_t *__restrict__ query, long long n_query, int symmetric,
                 const float *__restrict__ hmax_global, uint32_t *__restrict__ out, long long cap,
                 int32_t *__restrict__ count)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long qi = (long long)blockIdx.x * SPHB_WARPS + warp;
    if (qi >= n_query) return;
    const SphbScanCtx gc = sphb_make_ctx(c);
    const uint32_t j = query[qi];
    double xj, yj, zj, hj;
    sphb_load_posh(c.posh, j, xj, yj, zj, hj);
    const float4 fj = __ldg(gc.frec + j);
    long long total = 0;
  for (int lb = 0; lb < gc.nlev; ++lb) {   // nested grids: level by level, ascending slots
    const SphbLevelCtx g = sphb_level_ctx(gc, lb);
    double hr = hj;
    if (symmetric) {
        const double hm = (gc.nlev > 1) ? (double)__ldg(gc.level_hmax + lb) : (hmax_global ? (double)hmax_global[0] : hj);
        if (!(hm <= hr)) hr = hm;
    }
    const double R = (hr > 0.0) ? (2.0 * hr * 1.000001 + (double)gc.delta) * 1.00001 : (double)CUDART_INF_F;
    const double fxl = (double)fj.x - g.sx, fyl = (double)fj.y - g.sy, fzl = (double)fj.z - g.sz;
    if (g.bounded && (sphb_axis_miss(fxl, fxl, R, g.nx, g.cellx) || sphb_axis_miss(fyl, fyl, R, g.ny, g.cell) ||
                      sphb_axis_miss(fzl, fzl, R, g.nz, g.cell)))
        continue;
    const int cxl = sphb_cell_lo(fxl - R, g.inv_cellx, g.nx), cxh = sphb_cell_hi(fxl + R, g.inv_cellx, g.nx);
    const int cyl = sphb_cell_lo(fyl - R, g.inv_cell, g.ny), cyh = sphb_cell_hi(fyl + R, g.inv_cell, g.ny);
    const int czl = sphb_cell_lo(fzl - R, g.inv_cell, g.nz), czh = sphb_cell_hi(fzl + R, g.inv_cell, g.nz);
    for (int rz = czl; rz <= czh; ++rz)
        for (int ry = cyl; ry <= cyh; ++ry) {
            const size_t rowbase = ((size_t)rz * g.ny + ry) * g.nx;
            const uint32_t s = g.cell_start[rowbase + cxl], e = g.cell_start[rowbase + cxh + 1];
            for (uint32_t t = s; t < e; t += 32) {
                const uint32_t i = t + lane;
                bool in = false;
                if (i < e) {
                    double x, y, z, hi;
                    sphb_load_posh(c.posh, i, x, y, z, hi);
                    const double d = __dsqrt_rn(sphb_d2(xj - x, yj - y, zj - z));   // distance(), pyfluid.py:43-47
                    in = __ddiv_rn(d, hj) < 2.0;                                    // complement of M4's q >= 2 branch
                    if (symmetric) in = in || (__ddiv_rn(d, hi) < 2.0);
                }
                const unsigned m = __ballot_sync(SPHB_FULL, in);
                if (in) {
                    const long long pos = total + __popc(m & ((1u << lane) - 1u));
                    if (pos < cap) out[qi * cap + pos] = i;
                }
                total += __popc(m);
            }
        }
  }
    if (lane == 0) count[qi] = (int32_t)total;
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
static inline int sphb_cells_ok(const sphb_cells *c)
{
    return c && c->n >= 0 && c->cell_start && c->cell_hmax && c->cell_id && c->posh && c->frec && c->grid.nx > 0 &&
           c->grid.ny > 0 && c->grid.nz > 0 && c->grid.cell > 0.0 &&
           (c->nlevels <= 1 || (c->nlevels <= SPHB_MAX_LEVELS && c->levels && c->level_hmax));
}

static bool sphb_lists_ok(const sphb_lists *l)
{
    return !l || (l->store && l->broken && l->hmax_global && l->cover >= 1.0 && l->cover < 2.0);
}

extern "C" int sphb_density_omega(const sphb_cells *cells, const sphb_consts *k, const uint8_t *target,
                                  const double *h_target, const double *rho_in, double *rho_sum, double *dwdh_sum,
                                  double *omega, double *xi_out, int32_t *ngb_count, const sphb_lists *lists,
                                  void *stream)
{
    SPHB_RANGE();
    if (!sphb_cells_ok(cells) || !k || !sphb_lists_ok(lists)) return SPHB_E_ARG;
    if (cells->n == 0) return SPHB_OK;
    const long long blocks = (cells->n + SPHB_THREADS - 1) / SPHB_THREADS;
    const sphb_lists ls = lists ? *lists : sphb_lists{nullptr, nullptr, nullptr, 1.0};
    const bool multi = cells->nlevels > 1;
#define SPHB_K2M(XI_, PR_, MU_)                                                                   \
    k_density_omega<XI_, PR_, MU_><<<(unsigned)blocks, SPHB_THREADS, 0, sphb_stream(stream)>>>(   \
        *cells, *k, target, h_target, rho_in, rho_sum, dwdh_sum, omega, xi_out, ngb_count, ls)
#define SPHB_K2(XI_, PR_)                                                                         \
    do {                                                                                          \
        if (multi) SPHB_K2M(XI_, PR_, true); else SPHB_K2M(XI_, PR_, false);                      \
    } while (0)
    if (xi_out) {
        if (lists) SPHB_K2(true, true); else SPHB_K2(true, false);
    } else {
        if (lists) SPHB_K2(false, true); else SPHB_K2(false, false);
    }
#undef SPHB_K2
#undef SPHB_K2M
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" size_t sphb_lists_bytes(int64_t n)
{
    const long long warps = (n + 31) / 32;
    return (size_t)(warps > 0 ? warps : 1) * sizeof(SphbWarpSaved);
}

extern "C" int sphb_bisect_h(const sphb_cells *cells, const sphb_consts *k, const uint32_t *list, int64_t n_list,
                             double *h_out, double *omega_out, int32_t *code, int32_t *oob, int32_t indexed_by_slot,
                             sphb_status *status, void *stream)
{
    SPHB_RANGE();
    if (!sphb_cells_ok(cells) || !k || !list || !h_out || !status) return SPHB_E_ARG;
    if (n_list == 0 || cells->n == 0) return SPHB_OK;
    long long blocks = n_list > 0 ? n_list : 296;
    if (blocks > 1184) blocks = 1184;
    k_bisect_h<<<(unsigned)blocks, SPHB_THREADS, 0, sphb_stream(stream)>>>(*cells, *k, list, (long long)n_list, h_out,
                                                                           omega_out, code, oob, indexed_by_slot, status);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_newton_h(const sphb_cells *cells, const sphb_consts *k, const uint8_t *target,
                             const double *h_guess, double *h_out, double *omega_out, int32_t *iters, int32_t *code,
                             uint32_t *fail_list, sphb_status *status, const sphb_lists *lists, void *stream)
{
    SPHB_RANGE();
    if (!sphb_cells_ok(cells) || !k || !h_guess || !h_out || !fail_list || !status || !sphb_lists_ok(lists))
        return SPHB_E_ARG;
    if (cells->n == 0) return SPHB_OK;
    const long long blocks = (cells->n + SPHB_THREADS - 1) / SPHB_THREADS;
    SPHB_CUDA(cudaMemsetAsync(&status->nfail, 0, sizeof(int32_t), sphb_stream(stream)));
    const sphb_lists ls = lists ? *lists : sphb_lists{nullptr, nullptr, nullptr, 1.0};
    const bool multi = cells->nlevels > 1;
#define SPHB_K3M(FU_, PR_, MU_)                                                                             \
    k_newton_h<FU_, PR_, MU_><<<(unsigned)blocks, SPHB_THREADS, 0, sphb_stream(stream)>>>(*cells, *k, target, h_guess, h_out, \
                                                                                          omega_out, iters, code, fail_list, status, ls)
#define SPHB_K3(FU_, PR_)                                                                                   \
    do {                                                                                                    \
        if (multi) SPHB_K3M(FU_, PR_, true); else SPHB_K3M(FU_, PR_, false);                                \
    } while (0)
    if (omega_out) {
        if (lists) SPHB_K3(true, true); else SPHB_K3(true, false);
    } else {
        if (lists) SPHB_K3(false, true); else SPHB_K3(false, false);
    }
#undef SPHB_K3
#undef SPHB_K3M
    SPHB_LAUNCHED();
    // the fallback reads its list length from status->nfail on the device: no host round trip
    return sphb_bisect_h(cells, k, fail_list, -1, h_out, omega_out, code, nullptr, 1, status, stream);
}

extern "C" int sphb_neighbours(const sphb_cells *cells, const uint32_t *query, int64_t n_query, int32_t symmetric,
                               const float *hmax_global, uint32_t *out, int64_t cap, int32_t *count, void *stream)
{
    SPHB_RANGE();
    if (!sphb_cells_ok(cells) || !query || !out || !count || cap < 0 || n_query < 0) return SPHB_E_ARG;
    if (symmetric && !hmax_global) return SPHB_E_ARG;
    if (n_query == 0) return SPHB_OK;
    const long long blocks = (n_query + SPHB_WARPS - 1) / SPHB_WARPS;
    k_neighbours<<<(unsigned)blocks, SPHB_THREADS, 0, sphb_stream(stream)>>>(*cells, query, (long long)n_query,
                                                                             symmetric, hmax_global, out,
                                                                             (long long)cap, count);
    SPHB_LAUNCHED();
    return SPHB_OK;
}
