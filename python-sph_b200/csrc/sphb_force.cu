// sphb_force.cu -- K4: EOS prologue and the fused momentum + energy sums (pressure gradient with
// grad-h terms + Monaghan artificial viscosity), and K5: the predictor / corrector updates.
#include <stdio.h>

#include "sphb_scan.cuh"
#include "sphb_pair.cuh"

#define SPHB_WARPS 4
#ifndef SPHB_MINB_FORCE
#define SPHB_MINB_FORCE 4
#endif
#define SPHB_THREADS (SPHB_WARPS * 32)
#ifndef SPHB_FORCE_FLUSH_ROWS
#define SPHB_FORCE_FLUSH_ROWS 0
#endif

// ------------------------------------------------------------------------------------------------
// EOS prologue: var_density (pyfluid.py:237-245), pressure (:248-259), plus the per-particle factors
// every pair term needs: A = P/(omega rho^2) (:466-467,:352), cs = (gamma P/rho)^0.5 (:274-275),
// 1/h and 1/(pi h^4) (:66)
// ------------------------------------------------------------------------------------------------
__global__ void k_eos(const double *__restrict__ posh, const double *__restrict__ vel, const double *__restrict__ u,
                      const double *__restrict__ omega, const double *__restrict__ rho_in,
                      const double *__restrict__ P_in, const double *__restrict__ xi, long long n, sphb_consts k,
                      double *__restrict__ rho_out, double *__restrict__ P_out, double *__restrict__ recB,
                      double *__restrict__ recC, double *__restrict__ bgrav, sphb_status *__restrict__ status,
                      double *__restrict__ recD)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double h = posh[4 * i + 3];
    const double rho = rho_in ? rho_in[i] : sphb_var_density_m(k, sphb_mass_of(k, (size_t)i), h);
    // (ADIABATIC_INDEX - 1) * energy * density
    const double P = P_in ? P_in[i] : __dmul_rn(__dmul_rn(k.gamma - 1.0, u[i]), rho);
    if (P < 0.0) atomicOr(&status->flags, SPHB_ST_NEG_P);
    if (rho_out) rho_out[i] = rho;
    if (P_out) P_out[i] = P;
    if (bgrav) bgrav[i] = xi ? (k.G / 2.0) * (xi[i] / omega[i]) : 0.0;   // (G/2) xi/Omega, pyfluid.py:436
    if (!recB || (!recC && !recD)) return;
    const double om = omega[i];
    const double A = P / (om * (rho * rho));
    const double cs = sqrt(k.gamma * P / rho);
    double2 *b = reinterpret_cast<double2 *>(recB) + 2 * i;
    b[0] = make_double2(vel[3 * i], vel[3 * i + 1]);
    b[1] = make_double2(vel[3 * i + 2], A);
    if (recD) reinterpret_cast<double2 *>(recD)[i] = make_double2(rho, cs);
    if (!recC) return;
    const double h2 = h * h;
    double2 *c = reinterpret_cast<double2 *>(recC) + 2 * i;
    c[0] = make_double2(rho, cs);
    c[1] = make_double2(1.0 / h, 1.0 / (SPHB_PI * (h2 * h2)));
}

#ifndef SPHB_FORCE_ILP
#define SPHB_FORCE_ILP 1
#endif

// drain the lane's list (SPHB_FORCE_ILP 1: two entries side by side, the next two position records in flight;
// 2: two side by side without look-ahead; 0: one at a time with one record of look-ahead)
template <bool GRAV>
__device__ __forceinline__ void sphb_force_flush(const double *posh, const double *recB, const double *recC,
                                                 const double *bgrav, const SphbWarpShared &ws, int lane, int cnt,
                                                 const sphb_consts &k, SphbForceAcc &a)
{
#if SPHB_FORCE_ILP == 1
    double A[4], B[4], C[4], D[4];
    uint32_t ia = 0, ib = 0, ic = 0, id = 0;
    A[0] = A[1] = A[2] = B[0] = B[1] = B[2] = 0.0;
    A[3] = B[3] = 1.0;
    if (cnt > 0) {
        ia = sphb_list_get(ws, lane, 0);
        sphb_load4(posh, ia, A[0], A[1], A[2], A[3]);
    }
    if (cnt > 1) {
        ib = sphb_list_get(ws, lane, 1);
        sphb_load4(posh, ib, B[0], B[1], B[2], B[3]);
    }
    int n_ = 0;
    while (n_ < cnt) {
        C[0] = C[1] = C[2] = D[0] = D[1] = D[2] = 0.0;
        C[3] = D[3] = 1.0;
        if (n_ + 2 < cnt) {
            ic = sphb_list_get(ws, lane, n_ + 2);
            sphb_load4(posh, ic, C[0], C[1], C[2], C[3]);
        }
        if (n_ + 3 < cnt) {
            id = sphb_list_get(ws, lane, n_ + 3);
            sphb_load4(posh, id, D[0], D[1], D[2], D[3]);
        }
        sphb_force_pair2<GRAV, false, false, false, true>(recB, recC, bgrav, ia, ib, A, B, n_ + 1 < cnt, k, a);
        n_ += 2;
        if (n_ >= cnt) break;
        A[0] = A[1] = A[2] = B[0] = B[1] = B[2] = 0.0;
        A[3] = B[3] = 1.0;
        if (n_ + 2 < cnt) {
            ia = sphb_list_get(ws, lane, n_ + 2);
            sphb_load4(posh, ia, A[0], A[1], A[2], A[3]);
        }
        if (n_ + 3 < cnt) {
            ib = sphb_list_get(ws, lane, n_ + 3);
            sphb_load4(posh, ib, B[0], B[1], B[2], B[3]);
        }
        sphb_force_pair2<GRAV, false, false, false, true>(recB, recC, bgrav, ic, id, C, D, n_ + 1 < cnt, k, a);
        n_ += 2;
    }
#elif SPHB_FORCE_ILP == 2
    for (int n_ = 0; n_ < cnt; n_ += 2) {
        double A[4], B[4];
        B[0] = B[1] = B[2] = 0.0;
        B[3] = 1.0;
        const uint32_t ia = sphb_list_get(ws, lane, n_);
        uint32_t ib = ia;
        sphb_load4(posh, ia, A[0], A[1], A[2], A[3]);
        if (n_ + 1 < cnt) {
            ib = sphb_list_get(ws, lane, n_ + 1);
            sphb_load4(posh, ib, B[0], B[1], B[2], B[3]);
        }
        sphb_force_pair2<GRAV, false, false, false, true>(recB, recC, bgrav, ia, ib, A, B, n_ + 1 < cnt, k, a);
    }
#else
    double ax, ay, az, ah, bx, by, bz, bh;
    uint32_t ia = 0, ib = 0;
    if (cnt > 0) {
        ia = sphb_list_get(ws, lane, 0);
        sphb_load4(posh, ia, ax, ay, az, ah);
    }
    int n_ = 0;
    while (n_ < cnt) {
        if (n_ + 1 < cnt) {
            ib = sphb_list_get(ws, lane, n_ + 1);
            sphb_load4(posh, ib, bx, by, bz, bh);
        }
        sphb_force_pair<GRAV>(recB, recC, bgrav, ia, ax, ay, az, ah, k, a, sphb_mass_w(k, ia));
        ++n_;
        if (n_ >= cnt) break;
        if (n_ + 1 < cnt) {
            ia = sphb_list_get(ws, lane, n_ + 1);
            sphb_load4(posh, ia, ax, ay, az, ah);
        }
        sphb_force_pair<GRAV>(recB, recC, bgrav, ib, bx, by, bz, bh, k, a, sphb_mass_w(k, ib));
        ++n_;
    }
#endif
}

template <bool GRAV, bool MULTI>
__global__ void __launch_bounds__(SPHB_THREADS, SPHB_MINB_FORCE)
    k_force(sphb_cells c, sphb_consts k, const uint8_t *__restrict__ target, const double *__restrict__ recB,
            const double *__restrict__ recC, const double *__restrict__ bgrav, const float *__restrict__ hmax_global,
            double *__restrict__ acc, double *__restrict__ dudt, sphb_status *__restrict__ status,
            const SphbWarpSaved *__restrict__ saved, const int32_t *__restrict__ broken)
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

    SphbForceAcc a;
    a.xj = a.yj = a.zj = 0.0;
    a.hj = 1.0;
    a.vxj = a.vyj = a.vzj = a.Aj = 0.0;
    a.rhoj = a.csj = a.hinvj = a.gj = 0.0;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t cid = 0xffffffffu;
    if (valid) {
        sphb_load4(c.posh, (size_t)j, a.xj, a.yj, a.zj, a.hj);
        sphb_load4(recB, (size_t)j, a.vxj, a.vyj, a.vzj, a.Aj);
        sphb_load4(recC, (size_t)j, a.rhoj, a.csj, a.hinvj, a.gj);
        f = __ldg(g.frec + j);
        cid = c.cell_id[j];
    }
    a.ax = a.ay = a.az = a.dens = a.visc = 0.0;
    a.negpi = 0;
    a.Bj = (GRAV && valid) ? bgrav[j] : 0.0;
#ifdef SPHB_DEBUG_FORCE
    a.jslot = tgt ? j : -2;
#endif
    const float Rj = sphb_radius_of_h(a.hj, g.delta), thr = f.w;
    const float RH = sphb_radius_of_h((double)hmax_global[0], g.delta);
    SphbList L;
    L.init(ws, lane);
    auto flush = [&](int n_) { sphb_force_flush<GRAV>(c.posh, recB, recC, bgrav, ws, lane, n_, k, a); };
    // lists parked by the producer kernel of this snapshot (sphb_lists): drain them, no scan
    int parked = -1;
    if (saved != nullptr && __ldg(broken) == 0)
        parked = sphb_lists_load(ws, lane, saved + ((long long)blockIdx.x * SPHB_WARPS + warp));
    if (parked >= 0) {
        flush(tgt ? parked : 0);
    } else {
        sphb_warp_scan<true, SPHB_FORCE_FLUSH_ROWS, MULTI>(g, ws, lane, tgt, f.x, f.y, f.z, Rj, thr, cid, RH, L, flush);
        flush(L.count());
    }
    if (tgt) {
        const double m = sphb_mass_after(k);
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

// Pi(j,i) and fluid_accel_viscosity_comp(j,i) for explicit pairs, written in the reference's own form
// (pyfluid.py:261-285, :459-474): this is the per-pair entry point, not the hot loop.
// forward declarations of the softened-gravity scalar functions (defined further down)
__device__ __forceinline__ double sphb_grav_force(double dist, double h);

__global__ void k_pair_terms(const double *__restrict__ posh, const double *__restrict__ recB,
                             const double *__restrict__ recC, const uint32_t *__restrict__ pairs, long long n_pairs,
                             sphb_consts k, double *__restrict__ pi_out, double *__restrict__ acc_out,
                             const double *__restrict__ bgrav, double *__restrict__ grav_out,
                             double *__restrict__ corr_out, sphb_status *__restrict__ status)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const uint32_t j = pairs[2 * p], i = pairs[2 * p + 1];
    double xj, yj, zj, hj, xi, yi, zi, hi, vxj, vyj, vzj, Aj, vxi, vyi, vzi, Ai, rhoj, csj, hinvj, gj, rhoi, csi, hinvi, gi;
    sphb_load4(posh, j, xj, yj, zj, hj);
    sphb_load4(posh, i, xi, yi, zi, hi);
    sphb_load4(recB, j, vxj, vyj, vzj, Aj);
    sphb_load4(recB, i, vxi, vyi, vzi, Ai);
    sphb_load4(recC, j, rhoj, csj, hinvj, gj);
    sphb_load4(recC, i, rhoi, csi, hinvi, gi);
    const double dx = xj - xi, dy = yj - yi, dz = zj - zi;
    const double vdr = sphb_dot3(dx, dy, dz, vxj - vxi, vyj - vyi, vzj - vzi);
    const double dist = __dsqrt_rn(sphb_d2(dx, dy, dz));
    double pi = 0.0;
    if (vdr < 0.0) {
        const double mean_h = 0.5 * (hj + hi);
        const double mu = mean_h * vdr / (dist * dist + k.epsilon * (mean_h * mean_h));
        const double cs = 0.5 * (csi + csj);
        const double mean_rho = 0.5 * (rhoj + rhoi);
        pi = (-k.alpha * cs * mu + k.beta * (mu * mu)) / mean_rho;
        if (pi < 0.0) atomicOr(&status->flags, SPHB_ST_NEG_PI);
    }
    if (pi_out) pi_out[p] = pi;
    if (acc_out) {
        double a[3] = {0.0, 0.0, 0.0};
        if (!(dx == 0.0 && dy == 0.0 && dz == 0.0)) {
            const double wj = sphb_m4_fp(dist / hj) * gj, wi = sphb_m4_fp(dist / hi) * gi;   // M4_d1
            const double d[3] = {dx / dist, dy / dist, dz / dist};                            // direction_i_j
            for (int c = 0; c < 3; ++c) {
                const double gJ = wj * d[c], gI = wi * d[c];
                a[c] = -sphb_mass_of(k, i) * ((Aj * gJ + Ai * gI) + pi * 0.5 * (gJ + gI));
            }
        }
        acc_out[3 * p] = a[0];
        acc_out[3 * p + 1] = a[1];
        acc_out[3 * p + 2] = a[2];
    }
    const bool same = (dx == 0.0 && dy == 0.0 && dz == 0.0);
    if (grav_out) {   // smoothed_gravity_acceleration_comp, pyfluid.py:420-432
        double gq[3] = {0.0, 0.0, 0.0};
        if (!same) {
            const double f = sphb_grav_force(dist, hj) + sphb_grav_force(dist, hi);
            const double d[3] = {dx / dist, dy / dist, dz / dist};
            for (int c = 0; c < 3; ++c) gq[c] = -k.G / 2.0 * sphb_mass_of(k, i) * (f * d[c]);
        }
        for (int c = 0; c < 3; ++c) grav_out[3 * p + c] = gq[c];
    }
    if (corr_out) {   // smoothed_gravity_correction_comp, pyfluid.py:434-437 (bgrav = (G/2) xi/omega)
        double cq[3] = {0.0, 0.0, 0.0};
        if (!same && bgrav) {
            const double wj = sphb_m4_fp(dist / hj) * gj, wi = sphb_m4_fp(dist / hi) * gi;
            const double d[3] = {dx / dist, dy / dist, dz / dist};
            for (int c = 0; c < 3; ++c) cq[c] = -sphb_mass_of(k, i) * (bgrav[j] * (wj * d[c]) + bgrav[i] * (wi * d[c]));
        }
        for (int c = 0; c < 3; ++c) corr_out[3 * p + c] = cq[c];
    }
}

extern "C" int sphb_pair_terms(const double *posh, const double *recB, const double *recC, const uint32_t *pairs,
                               int64_t n_pairs, const sphb_consts *k, double *pi_out, double *acc_out,
                               const double *bgrav, double *grav_out, double *corr_out, sphb_status *status,
                               void *stream)
{
    SPHB_RANGE();
    if (n_pairs < 0 || !k || !status || (n_pairs > 0 && (!posh || !recB || !recC || !pairs))) return SPHB_E_ARG;
    if (n_pairs == 0) return SPHB_OK;
    k_pair_terms<<<(unsigned)((n_pairs + 127) / 128), 128, 0, sphb_stream(stream)>>>(
        posh, recB, recC, pairs, n_pairs, *k, pi_out, acc_out, bgrav, grav_out, corr_out, status);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

// ------------------------------------------------------------------------------------------------
// Softened self-gravity (pyfluid.py:378-389, :420-432).  The pair force is 1/r^2 beyond 2h: it has
// no compact support, so this is an exact all-pairs sum in fp64, tiled through shared memory (one thread per target,
// 128 candidates per tile, candidates in index order).  O(N^2): meant for the N <~ 1e6 runs where the reference's
// default G is kept; the 16 M benchmark path runs with G = 0.
// ------------------------------------------------------------------------------------------------
// grav_force(dist, h), pyfluid.py:378-389; np.piecewise order [x < 1, 1 <= x < 2, x >= 2], NaN -> 0
__device__ __forceinline__ double sphb_grav_force(double dist, double h)
{
    const double x = dist / h;
    const double ih2 = 1.0 / (h * h);
    const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
    const double in = ih2 * (4.0 / 3.0 * x - 6.0 / 5.0 * x3 + 0.5 * x4);
    const double mid = ih2 * (8.0 / 3.0 * x - 3.0 * x2 + 6.0 / 5.0 * x3 - 1.0 / 6.0 * x4 - 1.0 / (15.0 * x2));
    const double far = 1.0 / (dist * dist);
    return (x < 1.0) ? in : ((x >= 1.0 && x < 2.0) ? mid : ((x >= 2.0) ? far : 0.0));
}
// grav_potential(dist, h), pyfluid.py:365-376
__device__ __forceinline__ double sphb_grav_potential(double dist, double h)
{
    const double x = dist / h;
    const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2, x5 = x4 * x;
    const double in = 1.0 / h * (2.0 / 3.0 * x2 - 3.0 / 10.0 * x3 + 1.0 / 10.0 * x5 - 7.0 / 5.0);
    const double mid = 1.0 / h * (4.0 / 3.0 * x2 - x3 + 3.0 / 10.0 * x4 - 1.0 / 30.0 * x5 - 8.0 / 5.0 + 1.0 / (15.0 * x));
    return (x < 1.0) ? in : ((x >= 1.0 && x < 2.0) ? mid : ((x >= 2.0) ? -1.0 / dist : 0.0));
}
// grav_potential_smoothlength_derivative(dist, h), pyfluid.py:391-399
__device__ __forceinline__ double sphb_grav_dphidh(double dist, double h)
{
    const double x = dist / h;
    const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2, x5 = x4 * x;
    const double in = -1.0 / (h * h) * (2.0 * x2 - 12.0 / 10.0 * x3 + 6.0 / 10.0 * x5 - 7.0 / 5.0);
    const double mid = -1.0 / (h * h) * (4.0 * x2 - 4.0 * x3 + 3.0 / 2.0 * x4 - 1.0 / 5.0 * x5 - 8.0 / 5.0);
    return (x < 1.0) ? in : ((x >= 1.0 && x < 2.0) ? mid : 0.0);
}

#define GRAV_TILE 128
// acc[j] (+)= sum_i -G/2 m (grav_force(d, h_j) + grav_force(d, h_i)) (r_j - r_i)/d   for r_i != r_j
__global__ void __launch_bounds__(GRAV_TILE)
    k_gravity_direct(const double *__restrict__ posh_t, long long n_t, const double *__restrict__ posh, long long n,
                     sphb_consts k, int accumulate, double *__restrict__ acc)
{
    // targets posh_t[0..n_t), sources posh[0..n) (k.mass, if any, indexes the sources)
    __shared__ double4 tile[GRAV_TILE];
    __shared__ double tmass[GRAV_TILE];   // the candidates' weights: m_i with per-particle masses, else exactly 1
    const long long j = (long long)blockIdx.x * GRAV_TILE + threadIdx.x;
    double xj = 0.0, yj = 0.0, zj = 0.0, hj = 1.0;
    if (j < n_t) sphb_load4(posh_t, (size_t)j, xj, yj, zj, hj);
    double ax = 0.0, ay = 0.0, az = 0.0;
    for (long long t0 = 0; t0 < n; t0 += GRAV_TILE) {
        const long long i = t0 + threadIdx.x;
        double4 c = make_double4(0.0, 0.0, 0.0, 1.0);
        if (i < n) sphb_load4(posh, (size_t)i, c.x, c.y, c.z, c.w);
        __syncthreads();
        tile[threadIdx.x] = c;
        tmass[threadIdx.x] = (i < n) ? sphb_mass_w(k, (size_t)i) : 0.0;
        __syncthreads();
        const int nt = (int)((n - t0) < GRAV_TILE ? (n - t0) : GRAV_TILE);
        for (int q = 0; q < nt; ++q) {
            const double4 p = tile[q];
            const double dx = xj - p.x, dy = yj - p.y, dz = zj - p.z;
            if (dx == 0.0 && dy == 0.0 && dz == 0.0) continue;   // np.array_equal(rj, ri) -> 0, pyfluid.py:422
            const double d = __dsqrt_rn(sphb_d2(dx, dy, dz));
            const double f = sphb_grav_force(d, hj) + sphb_grav_force(d, p.w);
            const double s = tmass[q] * (f / d);
            ax = fma(s, dx, ax);
            ay = fma(s, dy, ay);
            az = fma(s, dz, az);
        }
    }
    if (j < n_t) {
        const double w = -k.G / 2.0 * sphb_mass_after(k);
        if (accumulate) {
            acc[3 * j] += w * ax;
            acc[3 * j + 1] += w * ay;
            acc[3 * j + 2] += w * az;
        } else {
            acc[3 * j] = w * ax;
            acc[3 * j + 1] = w * ay;
            acc[3 * j + 2] = w * az;
        }
    }
}

extern "C" int sphb_gravity_direct(const double *posh, int64_t n, const sphb_consts *k, int32_t accumulate, double *acc,
                                   void *stream)
{
    SPHB_RANGE();
    if (n < 0 || !k || (n > 0 && (!posh || !acc))) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_gravity_direct<<<(unsigned)((n + GRAV_TILE - 1) / GRAV_TILE), GRAV_TILE, 0, sphb_stream(stream)>>>(posh, n, posh, n, *k,
                                                                                                      accumulate, acc);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_gravity_direct_targets(const double *posh_targets, int64_t n_targets, const double *posh_sources,
                                           int64_t n_sources, const sphb_consts *k, int32_t accumulate, double *acc,
                                           void *stream)
{
    SPHB_RANGE();
    if (n_targets < 0 || n_sources < 0 || !k || (n_targets > 0 && (!posh_targets || !acc)) || (n_sources > 0 && !posh_sources))
        return SPHB_E_ARG;
    if (n_targets == 0) return SPHB_OK;
    k_gravity_direct<<<(unsigned)((n_targets + GRAV_TILE - 1) / GRAV_TILE), GRAV_TILE, 0, sphb_stream(stream)>>>(
        posh_targets, n_targets, posh_sources, n_sources, *k, accumulate, acc);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

__global__ void k_grav_eval(const double *__restrict__ dist, const double *__restrict__ h, long long n,
                            double *__restrict__ phi, double *__restrict__ force, double *__restrict__ dphidh)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (phi) phi[i] = sphb_grav_potential(dist[i], h[i]);
    if (force) force[i] = sphb_grav_force(dist[i], h[i]);
    if (dphidh) dphidh[i] = sphb_grav_dphidh(dist[i], h[i]);
}

extern "C" int sphb_grav_eval(const double *dist, const double *h, int64_t n, double *phi, double *force, double *dphidh,
                              void *stream)
{
    SPHB_RANGE();
    if (n < 0 || (n > 0 && (!dist || !h))) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_grav_eval<<<(unsigned)((n + 255) / 256), 256, 0, sphb_stream(stream)>>>(dist, h, n, phi, force, dphidh);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

// ------------------------------------------------------------------------------------------------
// K5: predictor (pyfluid.py:511-515) and corrector (:524-528); arithmetic in the reference's order
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double sphb_axpy_half(double a, double b, double dt)
{
    return __dadd_rn(a, __dmul_rn(__dmul_rn(b, 0.5), dt));   // a + b * 0.5 * dt
}

__global__ void k_predict(const double *__restrict__ posh0, const double *__restrict__ vel0, const double *__restrict__ u0,
                          const double *__restrict__ acc0, const double *__restrict__ dudt0, double dt, long long n,
                          double *__restrict__ posh_mid, double *__restrict__ vel_mid, double *__restrict__ u_mid,
                          sphb_status *__restrict__ status)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v[3], p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        v[k] = vel0[3 * i + k];
        p[k] = sphb_axpy_half(posh0[4 * i + k], v[k], dt);                 // pos0 + vel0 * 0.5 * dt
        vel_mid[3 * i + k] = sphb_axpy_half(v[k], acc0[3 * i + k], dt);    // vel0 + acc0 * 0.5 * dt
    }
    const double e = sphb_axpy_half(u0[i], dudt0[i], dt);                  // energy0 + energy_rate0 * 0.5 * dt
    if (e < 0.0) atomicOr(&status->flags, SPHB_ST_NEG_E);
    u_mid[i] = e;
    double2 *o = reinterpret_cast<double2 *>(posh_mid) + 2 * i;
    o[0] = make_double2(p[0], p[1]);
    o[1] = make_double2(p[2], posh0[4 * i + 3]);
}

__global__ void k_correct(const double *__restrict__ posh0, const double *__restrict__ vel0, const double *__restrict__ u0,
                          const double *__restrict__ acc_mid, const double *__restrict__ dudt_mid, double dt, long long n,
                          double *__restrict__ posh1, double *__restrict__ vel1, double *__restrict__ u1,
                          sphb_status *__restrict__ status)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double v0 = vel0[3 * i + k];
        const double v1 = __dadd_rn(v0, __dmul_rn(acc_mid[3 * i + k], dt));                       // vel0 + acc_mid * dt
        p[k] = __dadd_rn(posh0[4 * i + k], __dmul_rn(__dmul_rn(0.5, __dadd_rn(v0, v1)), dt));     // pos0 + 0.5*(vel0+vel1)*dt
        vel1[3 * i + k] = v1;
    }
    const double e = __dadd_rn(u0[i], __dmul_rn(dudt_mid[i], dt));                                // energy0 + rate_mid * dt
    if (e < 0.0) atomicOr(&status->flags, SPHB_ST_NEG_E);
    u1[i] = e;
    double2 *o = reinterpret_cast<double2 *>(posh1) + 2 * i;
    o[0] = make_double2(p[0], p[1]);
    o[1] = make_double2(p[2], posh0[4 * i + 3]);
}

// ------------------------------------------------------------------------------------------------
// helpers: bounds / h_max reductions, negative-h check, DFMA peak probe
// ------------------------------------------------------------------------------------------------
#define RED_THREADS 256

__global__ void __launch_bounds__(RED_THREADS)
    k_bounds_partial(const double *__restrict__ posh, long long n, double *__restrict__ part /*[grid][8]*/)
{
    __shared__ double sh[RED_THREADS / 32][8];
    double mn[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, mx[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
    double hm = -CUDART_INF, hs = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double2 a = reinterpret_cast<const double2 *>(posh)[2 * i];
        const double2 b = reinterpret_cast<const double2 *>(posh)[2 * i + 1];
        mn[0] = fmin(mn[0], a.x); mx[0] = fmax(mx[0], a.x);
        mn[1] = fmin(mn[1], a.y); mx[1] = fmax(mx[1], a.y);
        mn[2] = fmin(mn[2], b.x); mx[2] = fmax(mx[2], b.x);
        hm = fmax(hm, b.y);
        hs += b.y;
    }
    double v[8] = {mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], hm, hs};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = fmin(v[k], __shfl_xor_sync(SPHB_FULL, v[k], o));
#pragma unroll
        for (int k = 3; k < 7; ++k) v[k] = fmax(v[k], __shfl_xor_sync(SPHB_FULL, v[k], o));
        v[7] += __shfl_xor_sync(SPHB_FULL, v[7], o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int k = 0; k < 8; ++k) sh[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < RED_THREADS / 32; ++w) {
            for (int k = 0; k < 3; ++k) v[k] = fmin(v[k], sh[w][k]);
            for (int k = 3; k < 7; ++k) v[k] = fmax(v[k], sh[w][k]);
            v[7] += sh[w][7];
        }
        for (int k = 0; k < 8; ++k) part[(size_t)blockIdx.x * 8 + k] = v[k];
    }
}

__global__ void k_bounds_final(const double *__restrict__ part, int nb, double *__restrict__ out8, float *__restrict__ hmax_f)
{
    // one warp folds the per-block partials
    const int lane = threadIdx.x;
    double v[8] = {CUDART_INF, CUDART_INF, CUDART_INF, -CUDART_INF, -CUDART_INF, -CUDART_INF, -CUDART_INF, 0.0};
    for (int b = lane; b < nb; b += 32) {
        for (int k = 0; k < 3; ++k) v[k] = fmin(v[k], part[(size_t)b * 8 + k]);
        for (int k = 3; k < 7; ++k) v[k] = fmax(v[k], part[(size_t)b * 8 + k]);
        v[7] += part[(size_t)b * 8 + 7];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = fmin(v[k], __shfl_xor_sync(SPHB_FULL, v[k], o));
#pragma unroll
        for (int k = 3; k < 7; ++k) v[k] = fmax(v[k], __shfl_xor_sync(SPHB_FULL, v[k], o));
        v[7] += __shfl_xor_sync(SPHB_FULL, v[7], o);
    }
    if (lane == 0) {
        if (out8)
            for (int k = 0; k < 8; ++k) out8[k] = v[k];
        if (hmax_f) hmax_f[0] = sphb_f_up(v[6]);
    }
}

#define BOUNDS_BLOCKS 592

extern "C" int64_t sphb_bounds_workspace_bytes(int64_t n)
{
    (void)n;
    return (int64_t)BOUNDS_BLOCKS * 8 * sizeof(double);
}

extern "C" int sphb_bounds(const double *posh, int64_t n, double *out8, void *workspace, int64_t workspace_bytes,
                           void *stream)
{
    SPHB_RANGE();
    if (n <= 0 || !posh || !out8 || !workspace || workspace_bytes < sphb_bounds_workspace_bytes(n)) return SPHB_E_ARG;
    int nb = (int)((n + RED_THREADS - 1) / RED_THREADS);
    if (nb > BOUNDS_BLOCKS) nb = BOUNDS_BLOCKS;
    k_bounds_partial<<<nb, RED_THREADS, 0, sphb_stream(stream)>>>(posh, n, (double *)workspace);
    SPHB_LAUNCHED();
    k_bounds_final<<<1, 32, 0, sphb_stream(stream)>>>((const double *)workspace, nb, out8, nullptr);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

// atomicMax on a non-negative float through its int pattern
__global__ void k_hmax(const double *__restrict__ posh, long long n, int *__restrict__ hmax_bits)
{
    float hm = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double h = posh[4 * i + 3];
        hm = (h > 0.0) ? fmaxf(hm, sphb_f_up(h)) : CUDART_INF_F;
    }
    hm = sphb_warp_max(hm);
    if ((threadIdx.x & 31) == 0) atomicMax(hmax_bits, __float_as_int(hm));
}

extern "C" int sphb_hmax(const double *posh, int64_t n, float *hmax_out, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || !hmax_out || (n > 0 && !posh)) return SPHB_E_ARG;
    SPHB_CUDA(cudaMemsetAsync(hmax_out, 0, sizeof(float), sphb_stream(stream)));
    if (n == 0) return SPHB_OK;
    int nb = (int)((n + RED_THREADS - 1) / RED_THREADS);
    if (nb > BOUNDS_BLOCKS) nb = BOUNDS_BLOCKS;
    k_hmax<<<nb, RED_THREADS, 0, sphb_stream(stream)>>>(posh, n, (int *)hmax_out);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

__global__ void k_check_neg_h(const double *__restrict__ h, long long n, sphb_status *__restrict__ status)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && h[i] < 0.0) atomicOr(&status->flags, SPHB_ST_NEG_H);
}

extern "C" int sphb_check_neg_h(const double *h, int64_t n, sphb_status *status, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || !status || (n > 0 && !h)) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_check_neg_h<<<(unsigned)((n + 255) / 256), 256, 0, sphb_stream(stream)>>>(h, n, status);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

// M4 / M4_d1 / M4_smoothlength_derivative, pyfluid.py:56-70,85-87 (true division for q, like the reference)
__global__ void k_m4_eval(const double *__restrict__ dist, const double *__restrict__ h, long long n,
                          double *__restrict__ W, double *__restrict__ dW, double *__restrict__ dWdh)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double hh = h[i], q = dist[i] / hh;
    double f, fp;
    sphb_m4(q, f, fp);
    const double h3 = hh * hh * hh;
    const double w = 1.0 / (SPHB_PI * h3) * f;
    const double wp = 1.0 / (SPHB_PI * (h3 * hh)) * fp;
    if (W) W[i] = w;
    if (dW) dW[i] = wp;
    if (dWdh) dWdh[i] = -q * wp - 3.0 / hh * w;
}

extern "C" int sphb_m4_eval(const double *dist, const double *h, int64_t n, double *W, double *dW, double *dWdh,
                            void *stream)
{
    SPHB_RANGE();
    if (n < 0 || (n > 0 && (!dist || !h))) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_m4_eval<<<(unsigned)((n + 255) / 256), 256, 0, sphb_stream(stream)>>>(dist, h, n, W, dW, dWdh);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

// distance(), pyfluid.py:43-47
__global__ void k_distance(const double *__restrict__ rj, const double *__restrict__ ri, long long n,
                           double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = __dsqrt_rn(sphb_d2(rj[3 * i] - ri[3 * i], rj[3 * i + 1] - ri[3 * i + 1], rj[3 * i + 2] - ri[3 * i + 2]));
}

extern "C" int sphb_distance(const double *rj, const double *ri, int64_t n, double *out, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || (n > 0 && (!rj || !ri || !out))) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_distance<<<(unsigned)((n + 255) / 256), 256, 0, sphb_stream(stream)>>>(rj, ri, n, out);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

// 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double *__restrict__ sink)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) sink[0] = s;
}

extern "C" int sphb_fp64_peak(int32_t blocks, int32_t iters, double *ms_host, double *flops_host)
{
    if (blocks <= 0 || iters <= 0 || !ms_host || !flops_host) return SPHB_E_ARG;
    double *sink = nullptr;
    SPHB_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    SPHB_CUDA(cudaEventCreate(&e0));
    SPHB_CUDA(cudaEventCreate(&e1));
    k_fp64_peak<<<blocks, 256>>>(iters / 8 + 1, sink);   // warm-up
    SPHB_CUDA(cudaEventRecord(e0));
    k_fp64_peak<<<blocks, 256>>>(iters, sink);
    SPHB_CUDA(cudaEventRecord(e1));
    SPHB_CUDA(cudaEventSynchronize(e1));
    g_sphb_launches += 2;
    float ms = 0.f;
    SPHB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_host = (double)ms;
    *flops_host = 2.0 * 8.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return SPHB_OK;
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" int sphb_eos(const double *posh, const double *vel, const double *u, const double *omega,
                        const double *rho_in, const double *P_in, const double *xi, int64_t n, const sphb_consts *k,
                        double *rho, double *P, double *recB, double *recC, double *bgrav, sphb_status *status,
                        void *stream)
{
    SPHB_RANGE();
    if (n < 0 || !k || !status) return SPHB_E_ARG;
    if (n > 0 && (!posh || (!u && !P_in))) return SPHB_E_ARG;
    if ((recB != nullptr) != (recC != nullptr)) return SPHB_E_ARG;
    if (n > 0 && recB && (!vel || !omega)) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    if (bgrav && !omega) return SPHB_E_ARG;
    k_eos<<<(unsigned)((n + 255) / 256), 256, 0, sphb_stream(stream)>>>(posh, vel, u, omega, rho_in, P_in, xi, n, *k, rho,
                                                                        P, recB, recC, bgrav, status, nullptr);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

// the force records of the list-draining kernel (sphb_nlist_force): recB = {vx, vy, vz, A} as above, recD = dense
// {rho, cs} (the kernel takes 1/h and 1/(pi h^4) of its own target from h, with the same two divisions)
extern "C" int sphb_nlist_eos(const double *posh, const double *vel, const double *u, const double *omega, int64_t n,
                              const sphb_consts *k, double *recB, double *recD, sphb_status *status, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || !k || !status) return SPHB_E_ARG;
    if (n > 0 && (!posh || !vel || !u || !omega || !recB || !recD)) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_eos<<<(unsigned)((n + 255) / 256), 256, 0, sphb_stream(stream)>>>(posh, vel, u, omega, nullptr, nullptr, nullptr, n, *k,
                                                                        nullptr, nullptr, recB, nullptr, nullptr, status, recD);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_force(const sphb_cells *cells, const sphb_consts *k, const uint8_t *target, const double *recB,
                          const double *recC, const double *bgrav, const float *hmax_global, double *acc, double *dudt,
                          sphb_status *status, const sphb_lists *lists, void *stream)
{
    SPHB_RANGE();
    if (!cells || !k || !recB || !recC || !hmax_global || !acc || !dudt || !status) return SPHB_E_ARG;
    if (lists && (!lists->store || !lists->broken)) return SPHB_E_ARG;
    if (cells->n < 0 || !cells->cell_start || !cells->cell_hmax || !cells->cell_id || !cells->posh || !cells->frec ||
        (cells->nlevels > 1 && (cells->nlevels > SPHB_MAX_LEVELS || !cells->levels || !cells->level_hmax)))
        return SPHB_E_ARG;
    if (k->G != 0.0 && !bgrav) return SPHB_E_ARG;   // with gravity on, the grad-h correction factors are required
    if (cells->n == 0) return SPHB_OK;
    const long long blocks = (cells->n + SPHB_THREADS - 1) / SPHB_THREADS;
    const SphbWarpSaved *saved = lists ? reinterpret_cast<const SphbWarpSaved *>(lists->store) : nullptr;
    const int32_t *broken = lists ? lists->broken : nullptr;
#define SPHB_K4(GR_, MU_)                                                                                          \
    k_force<GR_, MU_><<<(unsigned)blocks, SPHB_THREADS, 0, sphb_stream(stream)>>>(*cells, *k, target, recB, recC, bgrav, \
                                                                                  hmax_global, acc, dudt, status, saved, broken)
    if (cells->nlevels > 1) {
        if (bgrav) SPHB_K4(true, true); else SPHB_K4(false, true);
    } else {
        if (bgrav) SPHB_K4(true, false); else SPHB_K4(false, false);
    }
#undef SPHB_K4
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_predict(const double *posh0, const double *vel0, const double *u0, const double *acc0,
                            const double *dudt0, double dt, int64_t n, double *posh_mid, double *vel_mid, double *u_mid,
                            sphb_status *status, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || !status || (n > 0 && (!posh0 || !vel0 || !u0 || !acc0 || !dudt0 || !posh_mid || !vel_mid || !u_mid)))
        return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_predict<<<(unsigned)((n + 255) / 256), 256, 0, sphb_stream(stream)>>>(posh0, vel0, u0, acc0, dudt0, dt, n,
                                                                            posh_mid, vel_mid, u_mid, status);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_correct(const double *posh0, const double *vel0, const double *u0, const double *acc_mid,
                            const double *dudt_mid, double dt, int64_t n, double *posh1, double *vel1, double *u1,
                            sphb_status *status, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || !status || (n > 0 && (!posh0 || !vel0 || !u0 || !acc_mid || !dudt_mid || !posh1 || !vel1 || !u1)))
        return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_correct<<<(unsigned)((n + 255) / 256), 256, 0, sphb_stream(stream)>>>(posh0, vel0, u0, acc_mid, dudt_mid, dt, n,
                                                                            posh1, vel1, u1, status);
    SPHB_LAUNCHED();
    return SPHB_OK;
}
