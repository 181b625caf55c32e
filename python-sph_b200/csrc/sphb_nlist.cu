// sphb_nlist.cu -- neighbour lists that outlive a kernel AND a step, and the gather kernels that only drain them.
//
// The scanning kernels (sphb_gather.cu, sphb_force.cu) find every target's candidates again in each launch; one
// var_smoothlength_leapfrog step (pyfluid.py:498-532) launches five of them on three snapshots that differ by
// v dt -- a small fraction of h.  Here the candidates are found ONCE (sphb_nlist_build: the same warp scan, with the
// symmetric support 2 max(c_j h_j, c_i h_i) plus the allowances s_j h_j + s_i h_i) and stored as rows of sorted-slot indices; the gather
// kernels below (density / Omega, Newton h solve, momentum + energy) are pure drains: no cell list, no scan code, no
// shared-memory lists, far fewer registers.  The particle state stays in the build's slot order, positions and h are
// rewritten in place, and the lists stay complete while
//       h_now <= cover_h * h_build      and      |x_now - x_build| <= skin * h_build        for every record
// (sphb_nlist_check and the Newton drain set SPHB_NL_BROKEN otherwise; the caller then rebuilds and redoes the step).
// Entries outside the current support add exact zeros (membership is still decided pair by pair in FP64 with the
// reference's q < 2 rule), and rows are in ascending slot order, so every sum has the same bits as the scanning
// kernels give on the same cell list.
#include <stdio.h>

#include "sphb_scan.cuh"
#include "sphb_pair.cuh"

#define NL_WARPS 4
#define NL_THREADS (NL_WARPS * 32)
#ifndef SPHB_NL_RESERVE
#define SPHB_NL_RESERVE 768 /* rows a warp reserves when its lists outgrow the staging area (4 chunks) */
#endif
#ifndef SPHB_NL_MAX_ROWS
#define SPHB_NL_MAX_ROWS (1 << 18) /* rows one warp may take (32 MiB): beyond that the input is "unlistable" */
#endif
#ifndef SPHB_MINB_NL_BUILD
#define SPHB_MINB_NL_BUILD 4
#endif
#ifndef SPHB_NL_LIST
#define SPHB_NL_LIST 192 /* per-lane staging capacity of the builder: 4 warps x 13.3 KB, four blocks per SM still fit */
#endif
typedef SphbWarpSharedT<SPHB_NL_LIST> NlWarpShared;
#ifndef SPHB_MINB_NL_OWN
#define SPHB_MINB_NL_OWN 8
#endif
#ifndef SPHB_NL_EARLY_B
#define SPHB_NL_EARLY_B 0 /* force drain: load the {v, A} records with the positions, two rows ahead */
#endif
#ifndef SPHB_MINB_NL_FORCE
#define SPHB_MINB_NL_FORCE 4
#endif

// ------------------------------------------------------------------------------------------------
// build: one warp = 32 consecutive slots; its lists become `rows` rows of 32 slot indices (row k = entry k of every
// lane, SPHB_NL_NONE where a lane has fewer), handed out from the pool by one atomicAdd per warp.  A warp whose lists
// do not fit the shared-memory staging area (SPHB_NL_LIST entries per lane) counts its rows first and scans twice.
// ------------------------------------------------------------------------------------------------
// the density / grad-h sums of a chunk of staged lists (what sphb_own_flush does in the scanning kernels): two entries
// side by side, the records of the next two in flight
template <class WS>
__device__ __forceinline__ void nl_flush_own(const double *__restrict__ posh, const WS &ws, int lane, int cnt, SphbOwnAcc &a)
{
    double A[4], B[4], C[4], D[4];
    A[0] = A[1] = A[2] = B[0] = B[1] = B[2] = 0.0;
    if (cnt > 0) sphb_load_posh(posh, sphb_list_get(ws, lane, 0), A[0], A[1], A[2], A[3]);
    if (cnt > 1) sphb_load_posh(posh, sphb_list_get(ws, lane, 1), B[0], B[1], B[2], B[3]);
    int k = 0;
    while (k < cnt) {
        C[0] = C[1] = C[2] = D[0] = D[1] = D[2] = 0.0;
        if (k + 2 < cnt) sphb_load_posh(posh, sphb_list_get(ws, lane, k + 2), C[0], C[1], C[2], C[3]);
        if (k + 3 < cnt) sphb_load_posh(posh, sphb_list_get(ws, lane, k + 3), D[0], D[1], D[2], D[3]);
        sphb_own_pair2<false>(A, B, k + 1 < cnt, a);
        k += 2;
        if (k >= cnt) break;
        if (k + 2 < cnt) sphb_load_posh(posh, sphb_list_get(ws, lane, k + 2), A[0], A[1], A[2], A[3]);
        if (k + 3 < cnt) sphb_load_posh(posh, sphb_list_get(ws, lane, k + 3), B[0], B[1], B[2], B[3]);
        sphb_own_pair2<false>(C, D, k + 1 < cnt, a);
        k += 2;
    }
}

// OMEGA: also return Omega(pos, h) of the build-time records (pyfluid.py:91-105 with the algebraic density, :503):
// every chunk is drained once more from the staging area while it is stored -- the first gather pass of the step that
// follows a rebuild comes for the price of its arithmetic, without reading a single row back.
template <bool OMEGA, bool MULTI>
__global__ void __launch_bounds__(NL_THREADS, SPHB_MINB_NL_BUILD)
    k_nl_build(sphb_cells c, const uint8_t *__restrict__ target, const float *__restrict__ hmax_global, sphb_nlist nl,
               sphb_consts kc, double *__restrict__ omega)
{
    extern __shared__ __align__(16) unsigned char nl_smem[];
    NlWarpShared *sh = reinterpret_cast<NlWarpShared *>(nl_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * NL_WARPS + warp;
    const long long base = w * 32;
    if (base >= c.n) return;
    const long long j = base + lane;
    const bool valid = j < c.n;
    const bool tgt = valid && (target == nullptr || target[j] != 0);
    const SphbScanCtx g = sphb_make_ctx(c);
    NlWarpShared &ws = sh[warp];

    double x, y, z, h = 1.0;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t cid = 0xffffffffu;
    if (valid) {
        sphb_load_posh(c.posh, (size_t)j, x, y, z, h);
        f = __ldg(g.frec + j);
        cid = c.cell_id[j];
    }
    // complete for any later state with h_i <= c_i h_build and |dx_i| <= s_i h_build for every record (c, s per particle,
    // or the uniform cover_h, skin): |r_build| <= |r_now| + |dx_j| + |dx_i| < 2 max(c_j h_j, c_i h_i) + s_j h_j + s_i h_i
    SphbPpm pm;
    pm.margins = reinterpret_cast<const float2 *>(nl.margins);
    pm.cu = __double2float_ru(nl.cover_h);
    pm.su = __double2float_ru(nl.skin);
    pm.Aj = pm.Dj = 0.f;
    if (valid) sphb_ppm_of(pm, (uint32_t)j, f.w, pm.Aj, pm.Dj);
    pm.level_dmax = nl.level_dmax;
    pm.cell_dmax = nl.cell_dmax;
    pm.hmax_u = hmax_global[0];
    const float cov = pm.cu * 1.000001f;                                          // >= every c_i
    const float Rj = __fadd_ru(pm.Aj, pm.Dj);
    const float thr = 0.f, RH = 0.f;                                              // (formed level by level inside the scan)

    // mode 0: the lists are expected to fit the staging area: one exact allocation at the end, one store.
    // A warp that has to flush before the end of its scan (heavy: a sparse target next to a dense region) takes a
    // reserve of SPHB_NL_RESERVE rows and appends its chunks there (mode 3).  If even that is outgrown it only COUNTS the
    // rows of every further chunk (mode 1), allocates exactly the total, and scans once more to store (mode 2; the scan
    // is deterministic, the chunks come out the same).  Rows of a chunk are padded to its longest lane.
    SphbOwnAcc oa;
    oa.xj = x;
    oa.yj = y;
    oa.zj = z;
    oa.hinv = 1.0 / h;
    oa.sumF = oa.sumG = oa.sumX = 0.0;
    oa.nn = 0;
    uint32_t row0 = 0, used = 0, total = 0;   // warp-uniform
    int mode = 0;                             // warp-uniform
    bool over = false;                        // warp-uniform
    auto alloc = [&](uint32_t rows) {
        unsigned long long r0 = 0;
        if (lane == 0 && rows) r0 = atomicAdd(nl.cursor, (unsigned long long)rows);
        r0 = __shfl_sync(SPHB_FULL, r0, 0);
        if (r0 + rows > (unsigned long long)nl.pool_rows) {
            over = true;
            if (lane == 0) atomicOr(nl.flags, SPHB_NL_POOL);
        }
        row0 = (uint32_t)r0;
        used = 0;
    };
    auto emit = [&](int cnt, bool final) {
        __syncwarp();
        const uint32_t rows = (uint32_t)__reduce_max_sync(SPHB_FULL, cnt);
        if (mode == 0 && !final) {   // first overflow: an optimistic reserve, so that most heavy warps still scan once
            alloc((uint32_t)SPHB_NL_RESERVE);
            mode = 3;
        }
        if (mode == 3 && used + rows > (uint32_t)SPHB_NL_RESERVE) {   // outgrown: count the rest, then scan again
            mode = 1;
            total = used;
        }
        if (mode == 1) {
            total += rows;
            if (total > (uint32_t)SPHB_NL_MAX_ROWS && !over) {
                over = true;
                if (lane == 0) atomicOr(nl.flags, SPHB_NL_UNLISTABLE);
            }
            return;
        }
        if (mode == 0) alloc(rows);
        if (!over) {
            uint32_t *dst = nl.pool + ((size_t)row0 + used) * 32 + lane;
            for (uint32_t r = 0; r < rows; ++r)
                dst[(size_t)r * 32] = ((int)r < cnt) ? sphb_list_get(ws, lane, (int)r) : SPHB_NL_NONE;
            used += rows;
            if (OMEGA) nl_flush_own(c.posh, ws, lane, cnt, oa);
        }
        __syncwarp();
    };
    auto flush = [&](int n_) { emit(n_, false); };
    for (int pass = 0; pass < 2; ++pass) {
        SphbList L;
        L.init(ws, lane);
        sphb_warp_scan<true, 0, MULTI, true>(g, ws, lane, tgt, f.x, f.y, f.z, Rj, thr, cid, RH, L, flush, cov, pm);
        emit(L.count(), true);
        if (mode != 1 || over) break;
        alloc(total);
        mode = 2;
        oa.sumF = oa.sumG = 0.0;   // the chunks drained so far are drained again by the second scan
        oa.nn = 0;
        __syncwarp();
    }
    if (lane == 0) {
        nl.warp_off[w] = row0;
        nl.warp_rows[w] = over ? 0u : used;
    }
    if (OMEGA && tgt) omega[j] = over ? CUDART_NAN : sphb_omega_from_sumG(kc, h, oa.sumG, kc.particle_mass, kc.particle_mass);
}

// did any record leave what the lists were built to cover?  (half of a bound used up: SPHB_NL_AGING)
// Grid-stride, one atomic per block and quantity: half a million warps updating two addresses would serialise.
#define NL_CHECK_THREADS 256
#define NL_CHECK_BLOCKS 1184
__global__ void __launch_bounds__(NL_CHECK_THREADS)
    k_nl_check(const double *__restrict__ posh_build, const double *__restrict__ posh_now, long long n, double skin_u,
               double cover_u, const float *__restrict__ margins, int32_t *__restrict__ flags, int32_t *__restrict__ maxima)
{
    __shared__ int sh[3][NL_CHECK_THREADS / 32];
    int fl = 0;
    float dr = 0.f, gr = 0.f;   // displacement / h_build and h_now / h_build - 1 (what the caller sizes margins with)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double x0, y0, z0, h0, x1, y1, z1, h1;
        sphb_load_posh(posh_build, (size_t)i, x0, y0, z0, h0);
        sphb_load_posh(posh_now, (size_t)i, x1, y1, z1, h1);
        double skin = skin_u, cover_h = cover_u;
        if (margins) {
            const float2 m = __ldg(reinterpret_cast<const float2 *>(margins) + i);
            cover_h = (double)m.x;
            skin = (double)m.y;
        }
        const double dx = x1 - x0, dy = y1 - y0, dz = z1 - z0;
        const double d2 = dx * dx + dy * dy + dz * dz, lim = skin * h0;
        if (!(d2 <= lim * lim) || !(h1 <= cover_h * h0)) fl |= SPHB_NL_BROKEN;
        if (!(d2 <= 0.25 * (lim * lim)) || !(h1 <= (1.0 + 0.5 * (cover_h - 1.0)) * h0)) fl |= SPHB_NL_AGING;
        if (h0 > 0.0) {   // non-negative floats order like their bit patterns; a NaN (> +inf as an integer) sticks
            // the fractions of the two allowances this record has used up
            const double fd = (lim > 0.0) ? sqrt(d2) / lim : ((d2 > 0.0) ? (double)CUDART_INF_F : 0.0);
            const double gw = cover_h - 1.0, gg = h1 / h0 - 1.0;
            const double fg = (gg > 0.0) ? ((gw > 0.0) ? gg / gw : (double)CUDART_INF_F) : ((gg == gg) ? 0.0 : gg);
            dr = __int_as_float(max(__float_as_int(dr), __float_as_int(__double2float_ru(fd))));
            gr = __int_as_float(max(__float_as_int(gr), __float_as_int(__double2float_ru(fg))));
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wf = __reduce_or_sync(SPHB_FULL, fl);
    const int md = __reduce_max_sync(SPHB_FULL, __float_as_int(dr)), mg = __reduce_max_sync(SPHB_FULL, __float_as_int(gr));
    if (lane == 0) {
        sh[0][warp] = wf;
        sh[1][warp] = md;
        sh[2][warp] = mg;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int f = 0, d = 0, g = 0;
        for (int q = 0; q < NL_CHECK_THREADS / 32; ++q) {
            f |= sh[0][q];
            d = max(d, sh[1][q]);
            g = max(g, sh[2][q]);
        }
        if (f) atomicOr(flags, f);
        if (maxima) {
            if (d > 0) atomicMax(maxima, d);
            if (g > 0) atomicMax(maxima + 1, g);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// drains
// ------------------------------------------------------------------------------------------------
// one list row for this lane: the neighbour's record, or (padding) a record so far away that every kernel term
// is an exact zero
__device__ __forceinline__ void nl_fetch(const double *__restrict__ posh, uint32_t s, uint32_t self, double (&p)[4])
{
    const bool real = s != SPHB_NL_NONE;
    sphb_load_posh(posh, (size_t)(real ? s : self), p[0], p[1], p[2], p[3]);
    if (!real) p[0] = 1e150;
}

// the same for the own-support sums with per-particle masses: m_i rides in the record's fourth component (h_i is not
// used by those sums)
__device__ __forceinline__ void nl_fetch_m(const double *__restrict__ posh, const double *__restrict__ mass, uint32_t s,
                                           uint32_t self, double (&p)[4])
{
    const bool real = s != SPHB_NL_NONE;
    sphb_load_posh(posh, (size_t)(real ? s : self), p[0], p[1], p[2], p[3]);
    p[3] = real ? __ldg(mass + s) : 0.0;
    if (!real) p[0] = 1e150;
}

__device__ __forceinline__ uint32_t nl_row(const uint32_t *__restrict__ p, int k, int nrows)
{
    return (k < nrows) ? __ldg(p + (size_t)k * 32) : SPHB_NL_NONE;
}

// sumF, sumG (, nn) over the lane's rows: two entries side by side (two independent FP64 chains), the records of the
// next two rows and the indices of the two after them in flight
template <bool MASS = false>
__device__ __forceinline__ void nl_own_drain(const double *__restrict__ posh, const uint32_t *__restrict__ rows, int nrows,
                                             int lane, uint32_t self, SphbOwnAcc &a, const double *__restrict__ mass = nullptr)
{
    const uint32_t *p = rows + lane;
    uint32_t i0 = nl_row(p, 0, nrows), i1 = nl_row(p, 1, nrows), i2 = nl_row(p, 2, nrows), i3 = nl_row(p, 3, nrows);
    double A[4], B[4], C[4], D[4];
    auto fetch = [&](uint32_t s, double (&r)[4]) {
        if (MASS) nl_fetch_m(posh, mass, s, self, r); else nl_fetch(posh, s, self, r);
    };
    fetch(i0, A);
    fetch(i1, B);
    int k = 0;
    while (k < nrows) {
        fetch(i2, C);
        fetch(i3, D);
        i0 = nl_row(p, k + 4, nrows);
        i1 = nl_row(p, k + 5, nrows);
        sphb_own_pair2<false, MASS>(A, B, true, a);
        k += 2;
        if (k >= nrows) break;
        fetch(i0, A);
        fetch(i1, B);
        i2 = nl_row(p, k + 4, nrows);
        i3 = nl_row(p, k + 5, nrows);
        sphb_own_pair2<false, MASS>(C, D, true, a);
        k += 2;
    }
}

// K2 on kept lists: density / density_arr (pyfluid.py:218-232), omega_j / omega_arr (:91-105)
template <bool MASS>
__global__ void __launch_bounds__(NL_THREADS, SPHB_MINB_NL_OWN)
    k_nl_density_omega(sphb_nlist nl, sphb_consts k, const uint8_t *__restrict__ target, const double *__restrict__ posh,
                       const double *__restrict__ rho_in, double *__restrict__ rho_sum, double *__restrict__ dwdh_sum,
                       double *__restrict__ omega, int32_t *__restrict__ ngb_count)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * NL_WARPS + warp;
    const long long base = w * 32;
    if (base >= nl.n) return;
    const long long j = base + lane;
    const bool valid = j < nl.n;
    const bool tgt = valid && (target == nullptr || target[j] != 0);
    SphbOwnAcc a;
    double h = 1.0;
    a.xj = a.yj = a.zj = 0.0;
    if (valid) sphb_load_posh(posh, (size_t)j, a.xj, a.yj, a.zj, h);
    a.hinv = 1.0 / h;
    a.sumF = a.sumG = a.sumX = 0.0;
    a.nn = 0;
    const int nrows = (int)nl.warp_rows[w];
    const uint32_t *rows = nl.pool + (size_t)nl.warp_off[w] * 32;
    nl_own_drain<MASS>(posh, rows, nrows, lane, valid ? (uint32_t)j : 0u, a, k.mass);
    if (tgt) {
        const double pref = (MASS ? 1.0 : k.particle_mass) / (SPHB_PI * (h * h * h));
        const double rs = pref * a.sumF;
        const double ds = -(pref / h) * a.sumG;
        if (rho_sum) rho_sum[j] = rs;
        if (dwdh_sum) dwdh_sum[j] = ds;
        const double rho = rho_in ? rho_in[j] : sphb_var_density_m(k, MASS ? __ldg(k.mass + j) : k.particle_mass, h);
        if (omega) omega[j] = 1.0 + h / (3.0 * rho) * ds;   // pyfluid.py:97
        if (ngb_count) ngb_count[j] = a.nn;
    }
}

// K3 on kept lists: newton_smoothlength_iteration / _while / _arr (pyfluid.py:155-179, 203-211), guess = h_guess, or
// the h the records carry.  Same sequence of iterates and the same stop rule as sphb_newton_h; every evaluation is one more
// pass over the rows.  What the lists cannot serve -- an iterate outside (0, cover_h h_build], or no convergence
// within the limit (the reference then bisects over all particles, :178-179) -- sets SPHB_NL_BROKEN: the caller
// redoes the step with the scanning kernels, which replay the reference to the letter.
template <bool FUSED, bool MASS>
__global__ void __launch_bounds__(NL_THREADS, SPHB_MINB_NL_OWN)
    k_nl_newton_h(sphb_nlist nl, sphb_consts k, const uint8_t *__restrict__ target, const double *__restrict__ posh,
                  const double *__restrict__ h_guess, double *__restrict__ h_out, double *__restrict__ omega_out, double *__restrict__ posh_out,
                  int32_t *__restrict__ iters, sphb_status *__restrict__ status)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * NL_WARPS + warp;
    const long long base = w * 32;
    if (base >= nl.n) return;
    const long long j = base + lane;
    const bool valid = j < nl.n;
    const bool tgt = valid && (target == nullptr || target[j] != 0);
    SphbOwnAcc a;
    a.xj = a.yj = a.zj = 0.0;
    double h_cur = 1.0, hb = 1.0;
    if (valid) {
        sphb_load_posh(posh, (size_t)j, a.xj, a.yj, a.zj, h_cur);
        if (h_guess) h_cur = h_guess[j];
        hb = nl.posh_build[4 * j + 3];
    }
    const double h_in = h_cur;
    const double cj = (nl.margins && valid) ? (double)__ldg(nl.margins + 2 * j) : nl.cover_h;   // this record's own cover
    const double h_cover = hb * cj, h_old = hb * (1.0 + 0.5 * (cj - 1.0));
    const int nrows = (int)nl.warp_rows[w];
    const uint32_t *rows = nl.pool + (size_t)nl.warp_off[w] * 32;
    const uint32_t self = valid ? (uint32_t)j : 0u;
    const double mj = (MASS && valid) ? __ldg(k.mass + j) : k.particle_mass;   // (uniform: a constant-bank operand)
    // what the lane needs next: 1 = the sums of a Newton iterate at h_cur, 2 = sum m dW/dh at the converged h_cur
    // (FUSED: Omega at the solved h, pyfluid.py:518), 0 = nothing
    int want = tgt ? 1 : 0, it = 0, fl = 0;
    double h_res = h_in;
    while (__any_sync(SPHB_FULL, want != 0)) {
        if (want != 0 && !(h_cur > 0.0 && h_cur <= h_cover)) fl |= SPHB_NL_BROKEN;
        a.hinv = 1.0 / h_cur;
        a.sumF = a.sumG = a.sumX = 0.0;
        a.nn = 0;
        nl_own_drain<MASS>(posh, rows, nrows, lane, self, a, k.mass);
        if (want == 1) {
            // newton_smoothlength_iteration, pyfluid.py:155-163 (zetaprime with the rho-free identity)
            const double pref = (MASS ? 1.0 : k.particle_mass) / (SPHB_PI * (h_cur * h_cur * h_cur));
            const double rho = pref * a.sumF;
            const double sdh = -(pref / h_cur) * a.sumG;
            const double vd = sphb_var_density_m(k, mj, h_cur);
            const double zeta = vd - rho;
            const double zetap = -sdh - 3.0 * vd / h_cur;
            const double h_new = h_cur - zeta / zetap;
            ++it;
            // smoothlength_variation(old, new) = |old - new| / |new|, pyfluid.py:119-120,172
            if (fabs(h_cur - h_new) / fabs(h_new) < k.h_tol && h_new > 0.0) {
                h_res = h_new;
                want = FUSED ? 2 : 0;
                if (!(h_new <= h_cover)) fl |= SPHB_NL_BROKEN;
                if (!(h_new <= h_old)) fl |= SPHB_NL_AGING;
            } else if (it >= k.newton_limit) {
                h_res = CUDART_NAN;
                if (FUSED) omega_out[j] = CUDART_NAN;
                want = 0;
                fl |= SPHB_NL_BROKEN;   // the bisection (pyfluid.py:178-179) needs all particles: scanning path
            }
            h_cur = h_new;
        } else if (want == 2) {
            omega_out[j] = sphb_omega_from_sumG(k, h_cur, a.sumG, mj, MASS ? 1.0 : k.particle_mass);
            want = 0;
        }
        __syncwarp();
    }
    if (tgt) {
        h_out[j] = h_res;
        if (iters) iters[j] = it;
        if (h_res < 0.0) fl |= 0x10000;   // pyfluid.py:209-210
    }
    if (valid && posh_out) {
        double2 *o = reinterpret_cast<double2 *>(posh_out) + 2 * j;
        o[0] = make_double2(a.xj, a.yj);
        o[1] = make_double2(a.zj, tgt ? h_res : h_in);
    }
    const int any = __reduce_or_sync(SPHB_FULL, fl);
    if (lane == 0 && any) {
        if (any & 0xffff) atomicOr(nl.flags, any & 0xffff);
        if (any & 0x10000) atomicOr(&status->flags, SPHB_ST_NEG_H);
    }
    if (nl.maxima) {
        float gr = 0.f;
        if (tgt && hb > 0.0 && h_res > hb) gr = (cj > 1.0) ? __double2float_ru((h_res / hb - 1.0) / (cj - 1.0)) : CUDART_INF_F;
        const int mg = __reduce_max_sync(SPHB_FULL, __float_as_int(gr));
        // most warps do not raise the maximum: look before the atomic (half a million of them on one address serialise)
        if (lane == 0 && mg > 0 && mg > __ldcg(nl.maxima + 1)) atomicMax(nl.maxima + 1, mg);
    }
}

// K4 on kept lists: energy_rate / _arr (pyfluid.py:316-361), Pi (:261-285), fluid_accel_viscosity_comp (:459-474),
// acceleration / _arr (:477-493) with G = 0.  posh carries the CURRENT h; recB / recC are sphb_eos's records, or
// (DENSE) recB and sphb_nlist_eos's dense {rho, cs}.
template <bool DENSE, bool MASS>
__global__ void __launch_bounds__(NL_THREADS, SPHB_MINB_NL_FORCE)
    k_nl_force(sphb_nlist nl, sphb_consts k, const uint8_t *__restrict__ target, const double *__restrict__ posh,
               const double *__restrict__ recB, const double *__restrict__ recC, double *__restrict__ acc,
               double *__restrict__ dudt, sphb_status *__restrict__ status)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * NL_WARPS + warp;
    const long long base = w * 32;
    if (base >= nl.n) return;
    const long long j = base + lane;
    const bool valid = j < nl.n;
    const bool tgt = valid && (target == nullptr || target[j] != 0);
    SphbForceAcc a;
    a.xj = a.yj = a.zj = 0.0;
    a.hj = 1.0;
    a.vxj = a.vyj = a.vzj = a.Aj = 0.0;
    a.rhoj = a.csj = a.hinvj = a.gj = 0.0;
    if (valid) {
        sphb_load4(posh, (size_t)j, a.xj, a.yj, a.zj, a.hj);
        sphb_load4(recB, (size_t)j, a.vxj, a.vyj, a.vzj, a.Aj);
        if (DENSE) {
            const double2 d = reinterpret_cast<const double2 *>(recC)[j];
            a.rhoj = d.x;
            a.csj = d.y;
            const double h2 = a.hj * a.hj;
            a.hinvj = 1.0 / a.hj;                       // the same two divisions as k_eos: same bits
            a.gj = 1.0 / (SPHB_PI * (h2 * h2));
        } else {
            sphb_load4(recC, (size_t)j, a.rhoj, a.csj, a.hinvj, a.gj);
        }
    }
    a.ax = a.ay = a.az = a.dens = a.visc = 0.0;
    a.negpi = 0;
    a.Bj = 0.0;
    const int nrows = (int)nl.warp_rows[w];
    const uint32_t *p = nl.pool + (size_t)nl.warp_off[w] * 32 + lane;
    const uint32_t self = valid ? (uint32_t)j : 0u;
#if SPHB_NL_EARLY_B
    {
        // the {v, A} records travel with the positions: both are in flight two rows ahead of the arithmetic
        uint32_t i0 = nl_row(p, 0, nrows), i1 = nl_row(p, 1, nrows), i2 = nl_row(p, 2, nrows), i3 = nl_row(p, 3, nrows);
        double A[4], B[4], C[4], D[4], VA[4], VB[4], VC[4], VD[4];
        nl_fetch(posh, i0, self, A);
        nl_fetch(posh, i1, self, B);
        nl_fetch(recB, i0, self, VA);
        nl_fetch(recB, i1, self, VB);
        int kk = 0;
        while (kk < nrows) {
            nl_fetch(posh, i2, self, C);
            nl_fetch(posh, i3, self, D);
            nl_fetch(recB, i2, self, VC);
            nl_fetch(recB, i3, self, VD);
            uint32_t n0 = nl_row(p, kk + 4, nrows), n1 = nl_row(p, kk + 5, nrows);
            sphb_force_pair2<false, DENSE, true, MASS>(recB, recC, nullptr, i0, i1, A, B, i1 != SPHB_NL_NONE, k, a,
                                                 i0 != SPHB_NL_NONE, VA, VB);
            kk += 2;
            if (kk >= nrows) break;
            i0 = n0;
            i1 = n1;
            nl_fetch(posh, i0, self, A);
            nl_fetch(posh, i1, self, B);
            nl_fetch(recB, i0, self, VA);
            nl_fetch(recB, i1, self, VB);
            n0 = nl_row(p, kk + 4, nrows);
            n1 = nl_row(p, kk + 5, nrows);
            sphb_force_pair2<false, DENSE, true, MASS>(recB, recC, nullptr, i2, i3, C, D, i3 != SPHB_NL_NONE, k, a,
                                                 i2 != SPHB_NL_NONE, VC, VD);
            i2 = n0;
            i3 = n1;
            kk += 2;
        }
    }
#else
    {
        uint32_t i0 = nl_row(p, 0, nrows), i1 = nl_row(p, 1, nrows), i2 = nl_row(p, 2, nrows), i3 = nl_row(p, 3, nrows);
        double A[4], B[4], C[4], D[4];
        nl_fetch(posh, i0, self, A);
        nl_fetch(posh, i1, self, B);
        int kk = 0;
        while (kk < nrows) {
            nl_fetch(posh, i2, self, C);
            nl_fetch(posh, i3, self, D);
            uint32_t n0 = nl_row(p, kk + 4, nrows), n1 = nl_row(p, kk + 5, nrows);
            sphb_force_pair2<false, DENSE, false, MASS>(recB, recC, nullptr, i0, i1, A, B, i1 != SPHB_NL_NONE, k, a, i0 != SPHB_NL_NONE);
            kk += 2;
            if (kk >= nrows) break;
            i0 = n0;
            i1 = n1;
            nl_fetch(posh, i0, self, A);
            nl_fetch(posh, i1, self, B);
            n0 = nl_row(p, kk + 4, nrows);
            n1 = nl_row(p, kk + 5, nrows);
            sphb_force_pair2<false, DENSE, false, MASS>(recB, recC, nullptr, i2, i3, C, D, i3 != SPHB_NL_NONE, k, a, i2 != SPHB_NL_NONE);
            i2 = n0;
            i3 = n1;
            kk += 2;
        }
    }
#endif
    if (tgt) {
        const double m = MASS ? 1.0 : k.particle_mass;
        acc[3 * j] = m * a.ax;
        acc[3 * j + 1] = m * a.ay;
        acc[3 * j + 2] = m * a.az;
        const double visc = 0.5 * (m * a.visc);                 // viscosity_sum *= 0.5, pyfluid.py:346
        dudt[j] = fma(a.Aj, m * a.dens, visc);                  // P_j/(rho_j^2 omega_j) * density_change + viscosity_sum
        int fl = 0;
        if (a.negpi) fl |= SPHB_ST_NEG_PI;
        if (visc < 0.0) fl |= SPHB_ST_NEG_VISC;
        if (fl) atomicOr(&status->flags, fl);
    }
}

// caller order -> slot order in one pass: posh (n,4) from pos (n,3) + h (n), vel (n,3), u (n) through perm
__global__ void k_nl_gather_state(const double *__restrict__ pos, const double *__restrict__ h,
                                  const double *__restrict__ vel, const double *__restrict__ u,
                                  const uint32_t *__restrict__ perm, long long n, double *__restrict__ posh_s,
                                  double *__restrict__ vel_s, double *__restrict__ u_s)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const size_t i = perm[s];
    if (posh_s) {
        double2 *o = reinterpret_cast<double2 *>(posh_s) + 2 * s;
        o[0] = make_double2(pos[3 * i], pos[3 * i + 1]);
        o[1] = make_double2(pos[3 * i + 2], h[i]);
    }
    if (vel_s) {
        vel_s[3 * s] = vel[3 * i];
        vel_s[3 * s + 1] = vel[3 * i + 1];
        vel_s[3 * s + 2] = vel[3 * i + 2];
    }
    if (u_s) u_s[s] = u[i];
}

// slot order -> caller order: pos (n,3), h (n) from posh_s, vel, u through perm
__global__ void k_nl_scatter_state(const double *__restrict__ posh_s, const double *__restrict__ vel_s,
                                   const double *__restrict__ u_s, const uint32_t *__restrict__ perm, long long n,
                                   double *__restrict__ pos, double *__restrict__ h, double *__restrict__ vel,
                                   double *__restrict__ u)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const size_t i = perm[s];
    if (posh_s) {
        const double2 a = reinterpret_cast<const double2 *>(posh_s)[2 * s];
        const double2 b = reinterpret_cast<const double2 *>(posh_s)[2 * s + 1];
        if (pos) {
            pos[3 * i] = a.x;
            pos[3 * i + 1] = a.y;
            pos[3 * i + 2] = b.x;
        }
        if (h) h[i] = b.y;
    }
    if (vel_s && vel) {
        vel[3 * i] = vel_s[3 * s];
        vel[3 * i + 1] = vel_s[3 * s + 1];
        vel[3 * i + 2] = vel_s[3 * s + 2];
    }
    if (u_s && u) u[i] = u_s[s];
}

// rows of up to four fp64 arrays (widths w0..w3, K = their sum) gathered into / scattered from one packed buffer:
// what a ghost refresh sends and receives.  One launch instead of a dozen tensor operations.
struct NlRowSet {
    double *p[4];
    int w[4];
};

template <bool PACK>
__global__ void k_nl_rows(NlRowSet rs, int K, const long long *__restrict__ idx, long long n_idx, double *__restrict__ buf)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * K) return;
    const long long r = t / K;
    int c = (int)(t - r * K);
    const long long row = idx[r];
    int q = 0;
    while (c >= rs.w[q]) {
        c -= rs.w[q];
        ++q;
    }
    double *e = rs.p[q] + row * rs.w[q] + c;
    if (PACK)
        buf[t] = *e;
    else
        *e = buf[t];
}

extern "C" int sphb_rows_pack(double *s0, int32_t w0, double *s1, int32_t w1, double *s2, int32_t w2, double *s3,
                              int32_t w3, const int64_t *idx, int64_t n_idx, double *buf, int32_t unpack, void *stream)
{
    SPHB_RANGE();
    NlRowSet rs{{s0, s1, s2, s3}, {w0, w1, w2, w3}};
    int K = 0;
    for (int q = 0; q < 4; ++q) {
        if (rs.w[q] < 0 || (rs.w[q] > 0 && !rs.p[q])) return SPHB_E_ARG;
        K += rs.w[q];
    }
    if (n_idx < 0 || K <= 0 || (n_idx > 0 && (!idx || !buf))) return SPHB_E_ARG;
    if (n_idx == 0) return SPHB_OK;
    const long long tot = (long long)n_idx * K;
    if (unpack)
        k_nl_rows<false><<<(unsigned)((tot + 255) / 256), 256, 0, sphb_stream(stream)>>>(rs, K, (const long long *)idx, n_idx, buf);
    else
        k_nl_rows<true><<<(unsigned)((tot + 255) / 256), 256, 0, sphb_stream(stream)>>>(rs, K, (const long long *)idx, n_idx, buf);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
static inline bool nl_ok(const sphb_nlist *nl)
{
    return nl && nl->n >= 0 && nl->warp_off && nl->warp_rows && nl->pool && nl->pool_rows > 0 &&
           nl->pool_rows < (int64_t)0xffffffffll && nl->cursor && nl->flags && nl->posh_build && nl->cover_h >= 1.0 &&
           nl->skin >= 0.0 && nl->cover_h + nl->skin < 4.0 && (!nl->margins || nl->level_dmax);
}

static inline unsigned nl_blocks(int64_t n) { return (unsigned)((n + NL_THREADS - 1) / NL_THREADS); }

extern "C" int64_t sphb_sizeof_nlist(void) { return (int64_t)sizeof(sphb_nlist); }

extern "C" int64_t sphb_nlist_pool_rows(int64_t n)
{
    // a first guess (a lattice at eta = 1.3 takes ~90 rows per warp, a random cloud ~130); a build that needs more sets
    // SPHB_NL_POOL and leaves the number of rows it asked for in *cursor: the caller enlarges the pool and builds again
    const int64_t warps = (n + 31) / 32;
    return warps * 128 + 16384;
}

extern "C" int sphb_nlist_build(const sphb_cells *cells, const uint8_t *target, const float *hmax_global,
                                const sphb_nlist *nl, const sphb_consts *k, double *omega, void *stream)
{
    SPHB_RANGE();
    if (omega && (!k || k->mass)) return SPHB_E_ARG;   // per-particle masses: Omega comes from sphb_nlist_density_omega
    if (!cells || cells->n < 0 || !cells->cell_start || !cells->cell_hmax || !cells->cell_id || !cells->posh ||
        !cells->frec || !hmax_global || !nl_ok(nl) || nl->n != cells->n || nl->posh_build != cells->posh ||
        (cells->nlevels > 1 && (cells->nlevels > SPHB_MAX_LEVELS || !cells->levels || !cells->level_hmax)))
        return SPHB_E_ARG;
    cudaStream_t st = sphb_stream(stream);
    SPHB_CUDA(cudaMemsetAsync(nl->cursor, 0, sizeof(unsigned long long), st));
    SPHB_CUDA(cudaMemsetAsync(nl->flags, 0, sizeof(int32_t), st));
    if (nl->maxima) SPHB_CUDA(cudaMemsetAsync(nl->maxima, 0, 2 * sizeof(int32_t), st));
    if (cells->n == 0) return SPHB_OK;
    const int smem = (int)(NL_WARPS * sizeof(NlWarpShared));   // > 48 KB of dynamic shared memory needs the opt-in
#define SPHB_NLB(OM_, MU_, KC_, OMP_)                                                                               \
    do {                                                                                                             \
        SPHB_CUDA(cudaFuncSetAttribute(k_nl_build<OM_, MU_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));    \
        k_nl_build<OM_, MU_><<<nl_blocks(cells->n), NL_THREADS, smem, st>>>(*cells, target, hmax_global, *nl, KC_, OMP_); \
    } while (0)
    if (cells->nlevels > 1) {
        if (omega) SPHB_NLB(true, true, *k, omega); else SPHB_NLB(false, true, sphb_consts{}, nullptr);
    } else {
        if (omega) SPHB_NLB(true, false, *k, omega); else SPHB_NLB(false, false, sphb_consts{}, nullptr);
    }
#undef SPHB_NLB
    SPHB_LAUNCHED();
    return SPHB_OK;
}

// which sorted slots do the lists name?  (mark[s] = 1)  A rank that lists only its own range of targets learns its exact
// ghost set from this: the decomposition then needs no geometric halo width at all.  Walks every warp's own rows (a
// heavy warp's reserve holds rows it never wrote).
__global__ void k_nl_mark(sphb_nlist nl, uint8_t *__restrict__ mark)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w * 32 >= nl.n) return;
    const uint32_t rows = nl.warp_rows[w];
    const uint32_t *p = nl.pool + (size_t)nl.warp_off[w] * 32 + lane;
    for (uint32_t r = 0; r < rows; ++r) {
        const uint32_t v = __ldg(p + (size_t)r * 32);
        if (v != SPHB_NL_NONE) mark[v] = 1;   // idempotent: racing writers store the same byte
    }
}

// rename the slots in place (pool entry v -> map[v]); map must be increasing on the named slots so that rows stay in
// ascending order, i.e. every sum keeps its order
__global__ void k_nl_remap(sphb_nlist nl, const int32_t *__restrict__ map)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w * 32 >= nl.n) return;
    const uint32_t rows = nl.warp_rows[w];
    uint32_t *p = nl.pool + (size_t)nl.warp_off[w] * 32 + lane;
    for (uint32_t r = 0; r < rows; ++r) {
        const uint32_t v = p[(size_t)r * 32];
        if (v != SPHB_NL_NONE) p[(size_t)r * 32] = (uint32_t)__ldg(map + v);
    }
}

extern "C" int sphb_nlist_mark(const sphb_nlist *nl, uint8_t *mark, void *stream)
{
    SPHB_RANGE();
    if (!nl_ok(nl) || !mark) return SPHB_E_ARG;
    if (nl->n == 0) return SPHB_OK;
    k_nl_mark<<<nl_blocks(nl->n), NL_THREADS, 0, sphb_stream(stream)>>>(*nl, mark);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_nlist_remap(const sphb_nlist *nl, const int32_t *map, void *stream)
{
    SPHB_RANGE();
    if (!nl_ok(nl) || !map) return SPHB_E_ARG;
    if (nl->n == 0) return SPHB_OK;
    k_nl_remap<<<nl_blocks(nl->n), NL_THREADS, 0, sphb_stream(stream)>>>(*nl, map);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_nlist_check(const sphb_nlist *nl, const double *posh_now, void *stream)
{
    SPHB_RANGE();
    if (!nl_ok(nl) || !posh_now) return SPHB_E_ARG;
    if (nl->n == 0) return SPHB_OK;
    long long nb = (nl->n + NL_CHECK_THREADS - 1) / NL_CHECK_THREADS;
    if (nb > NL_CHECK_BLOCKS) nb = NL_CHECK_BLOCKS;
    k_nl_check<<<(unsigned)nb, NL_CHECK_THREADS, 0, sphb_stream(stream)>>>(nl->posh_build, posh_now, (long long)nl->n, nl->skin,
                                                                          nl->cover_h, nl->margins, nl->flags, nl->maxima);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_nlist_density_omega(const sphb_nlist *nl, const sphb_consts *k, const uint8_t *target,
                                        const double *posh, const double *rho_in, double *rho_sum, double *dwdh_sum,
                                        double *omega, int32_t *ngb_count, void *stream)
{
    SPHB_RANGE();
    if (!nl_ok(nl) || !k || !posh) return SPHB_E_ARG;
    if (nl->n == 0) return SPHB_OK;
    if (k->mass)
        k_nl_density_omega<true><<<nl_blocks(nl->n), NL_THREADS, 0, sphb_stream(stream)>>>(*nl, *k, target, posh, rho_in, rho_sum,
                                                                                           dwdh_sum, omega, ngb_count);
    else
        k_nl_density_omega<false><<<nl_blocks(nl->n), NL_THREADS, 0, sphb_stream(stream)>>>(*nl, *k, target, posh, rho_in, rho_sum,
                                                                                            dwdh_sum, omega, ngb_count);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_nlist_newton_h(const sphb_nlist *nl, const sphb_consts *k, const uint8_t *target, const double *posh,
                                   const double *h_guess, double *h_out, double *omega_out, double *posh_out, int32_t *iters,
                                   sphb_status *status, void *stream)
{
    SPHB_RANGE();
    if (!nl_ok(nl) || !k || !posh || !h_out || !status || posh_out == posh) return SPHB_E_ARG;
    if (nl->n == 0) return SPHB_OK;
#define SPHB_NLN(FU_, MA_)                                                                                            \
    k_nl_newton_h<FU_, MA_><<<nl_blocks(nl->n), NL_THREADS, 0, sphb_stream(stream)>>>(*nl, *k, target, posh, h_guess, h_out, \
                                                                                      omega_out, posh_out, iters, status)
    if (k->mass) {
        if (omega_out) SPHB_NLN(true, true); else SPHB_NLN(false, true);
    } else {
        if (omega_out) SPHB_NLN(true, false); else SPHB_NLN(false, false);
    }
#undef SPHB_NLN
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_nlist_force(const sphb_nlist *nl, const sphb_consts *k, const uint8_t *target, const double *posh,
                                const double *recB, const double *recC, const double *recD, double *acc, double *dudt,
                                sphb_status *status, void *stream)
{
    SPHB_RANGE();
    if (!nl_ok(nl) || !k || !posh || !recB || (!recC == !recD) || !acc || !dudt || !status) return SPHB_E_ARG;
    if (k->G != 0.0) return SPHB_E_ARG;   // the softened-gravity terms ride in the scanning kernels only
    if (nl->n == 0) return SPHB_OK;
#define SPHB_NLF(DE_, MA_, RC_)                                                                                       \
    k_nl_force<DE_, MA_><<<nl_blocks(nl->n), NL_THREADS, 0, sphb_stream(stream)>>>(*nl, *k, target, posh, recB, RC_, acc, \
                                                                                   dudt, status)
    if (k->mass) {
        if (recD) SPHB_NLF(true, true, recD); else SPHB_NLF(false, true, recC);
    } else {
        if (recD) SPHB_NLF(true, false, recD); else SPHB_NLF(false, false, recC);
    }
#undef SPHB_NLF
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_gather_state(const double *pos, const double *h, const double *vel, const double *u,
                                 const uint32_t *perm, int64_t n, double *posh_s, double *vel_s, double *u_s, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || (n > 0 && !perm)) return SPHB_E_ARG;
    if (n > 0 && ((posh_s && (!pos || !h)) || (vel_s && !vel) || (u_s && !u))) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_nl_gather_state<<<(unsigned)((n + 255) / 256), 256, 0, sphb_stream(stream)>>>(pos, h, vel, u, perm, (long long)n, posh_s,
                                                                                  vel_s, u_s);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_scatter_state(const double *posh_s, const double *vel_s, const double *u_s, const uint32_t *perm,
                                  int64_t n, double *pos, double *h, double *vel, double *u, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || (n > 0 && !perm)) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_nl_scatter_state<<<(unsigned)((n + 255) / 256), 256, 0, sphb_stream(stream)>>>(posh_s, vel_s, u_s, perm, (long long)n, pos,
                                                                                   h, vel, u);
    SPHB_LAUNCHED();
    return SPHB_OK;
}
