// sphb_pair.cuh -- the FP64 pair arithmetic shared by the scanning kernels (sphb_gather.cu, sphb_force.cu) and the
// list-draining kernels (sphb_nlist.cu): one definition, so that a sum is the same bits whichever kernel forms it.
// Reference lines are cited per function (pyfluid.py of TRU-astrophysics/python-sph).
#pragma once
#include "sphb_common.cuh"

__device__ __forceinline__ void sphb_load_posh(const double *posh, size_t i, double &x, double &y, double &z, double &h)
{
    sphb_ld4(posh, i, x, y, z, h);
}

// the record of an own-support pair with its weight in the fourth component (those sums never use h_i): m_i with
// per-particle masses, else exactly 1
__device__ __forceinline__ void sphb_load_posw(const double *posh, const double *mass, size_t i, double &x, double &y, double &z,
                                               double &w)
{
    sphb_ld4(posh, i, x, y, z, w);
    w = mass ? __ldg(mass + i) : 1.0;
}

// One evaluation of sumF = sum f(q), sumG = sum (q f'(q) + 3 f(q)) over the lane's list.
// rho_sum = m sumF/(pi h^3) (pyfluid.py:218-224), sum m dW/dh = -m sumG/(pi h^4) (pyfluid.py:85-87,91-97).
struct SphbOwnAcc {
    double xj, yj, zj, hinv;
    double sumF, sumG, sumX;
    int nn;
};

// -h^2 dphi/dh of the softened potential, dimensionless: grav_potential_smoothlength_derivative(dist, h) =
// -(1/h^2) pX(x), x = dist/h (pyfluid.py:391-399); np.piecewise order [x < 1, 1 <= x < 2, x >= 2], NaN -> 0
__device__ __forceinline__ double sphb_grav_pX(double x)
{
    const double x2 = x * x, x3 = x2 * x;
    const double in = fma(0.6, x2 * x3, fma(-1.2, x3, fma(2.0, x2, -1.4)));
    const double mid = fma(-0.2, x2 * x3, fma(1.5, x2 * x2, fma(-4.0, x3, fma(4.0, x2, -1.6))));
    return (x < 1.0) ? in : ((x >= 1.0 && x < 2.0) ? mid : 0.0);
}

// w: the weight of the pair inside the sums (m_i with per-particle masses; 1 -- a multiplication the compiler drops, or an
// exact one -- when the uniform mass is applied after the sum)
template <bool XI = false>
__device__ __forceinline__ void sphb_own_pair(double x, double y, double z, SphbOwnAcc &a, double w = 1.0)
{
    const double dx = a.xj - x, dy = a.yj - y, dz = a.zj - z;
    const double d2 = sphb_d2(dx, dy, dz);
    const double r = sphb_nonzero(d2) ? d2 * sphb_rsqrt(d2) : 0.0;
    const double q = r * a.hinv;
    double f, fp;
    sphb_m4c(q, f, fp);
    a.sumF += w * f;
    a.sumG += w * fma(q, fp, 3.0 * f);
    a.nn += (q < 2.0) ? 1 : 0;
    if (XI) a.sumX += w * sphb_grav_pX(q);
}

// the pair's terms without touching the accumulators: two of these are independent instruction streams
template <bool XI = false>
__device__ __forceinline__ void sphb_own_terms(double x, double y, double z, const SphbOwnAcc &a, double &f, double &g,
                                               int &in, double &px, double w = 1.0)
{
    const double dx = a.xj - x, dy = a.yj - y, dz = a.zj - z;
    const double d2 = sphb_d2(dx, dy, dz);
    const double r = sphb_nonzero(d2) ? d2 * sphb_rsqrt(d2) : 0.0;
    const double q = r * a.hinv;
    double fp;
    sphb_m4c(q, f, fp);
    g = w * fma(q, fp, 3.0 * f);
    f = w * f;
    in = (q < 2.0) ? 1 : 0;
    px = XI ? w * sphb_grav_pX(q) : 0.0;
}

// entries k and k+1 evaluated side by side (the second only counted when it exists), accumulated in list order
// W: the records' fourth component is the pair's weight (the loader put m_i there: own-support sums never use h_i)
template <bool XI = false, bool W = false>
__device__ __forceinline__ void sphb_own_pair2(const double (&p)[4], const double (&q)[4], bool second, SphbOwnAcc &a)
{
    double f0, g0, x0, f1, g1, x1;
    int n0, n1;
    sphb_own_terms<XI>(p[0], p[1], p[2], a, f0, g0, n0, x0, W ? p[3] : 1.0);
    sphb_own_terms<XI>(q[0], q[1], q[2], a, f1, g1, n1, x1, W ? q[3] : 1.0);
    a.sumF += f0;
    a.sumG += g0;
    a.nn += n0;
    if (XI) a.sumX += x0;
    if (second) {
        a.sumF += f1;
        a.sumG += g1;
        a.nn += n1;
        if (XI) a.sumX += x1;
    }
}


// Omega from the grad-h sum at h with the algebraic density (pyfluid.py:91-97 as the step uses it, :503/:518)
// mj: the target's own mass (sphb_mass_of)
// mafter: the factor still to be applied to the sum (sphb_mass_after)
__device__ __forceinline__ double sphb_omega_from_sumG(const sphb_consts &k, double h, double sumG, double mj, double mafter)
{
    const double pref = mafter / (SPHB_PI * (h * h * h));
    const double ds = -(pref / h) * sumG;                           // sum_i m_i dW/dh
    return 1.0 + h / (3.0 * sphb_var_density_m(k, mj, h)) * ds;     // pyfluid.py:97 with rho = m_j (eta/h)^3
}


__device__ __forceinline__ void sphb_load4(const double *rec, size_t i, double &a, double &b, double &c, double &d)
{
    sphb_ld4(rec, i, a, b, c, d);
}


// ------------------------------------------------------------------------------------------------
// K4 pair sums.  For target j over i with q_j < 2 or q_i < 2 (pair terms vanish otherwise) and
// r_ji != 0 (coincident particles contribute exactly 0: pyfluid.py:51-52,461):
//   gradW(h) = W'(r,h) (r_j - r_i)/r = s (r_j - r_i),  s = fp(q)/(pi h^4)/r
//   dens  += s_j (v_ji . r_ji)                                        :327-331
//   Pi_ji  per :261-285 (convergent pairs only)
//   acc   -= (A_j s_j + A_i s_i + Pi/2 (s_j + s_i)) r_ji              :464-474
//   visc  += Pi (1/2 (s_j + s_i)) (v_ji . r_ji)                       :333-343
// m, and the final 1/2 of :346, are applied once per target.
// ------------------------------------------------------------------------------------------------
struct SphbForceAcc {
    double xj, yj, zj, hj;
    double vxj, vyj, vzj, Aj;
    double rhoj, csj, hinvj, gj;
    double ax, ay, az, dens, visc;
    double Bj;   // GRAV: (G/2) xi_j / Omega_j
    int negpi;
#ifdef SPHB_DEBUG_FORCE
    long long jslot;
#endif
};

#ifdef SPHB_DEBUG_FORCE
// debug aid (not built by default): record the contributing candidates of one target slot, in evaluation order
__device__ long long g_dbg_target = -1;
__device__ int g_dbg_count = 0;
__device__ uint32_t g_dbg_list[4096];
__device__ double g_dbg_val[4096];
extern "C" int sphb_debug_force_set(long long t)
{
    int z = 0;
    cudaMemcpyToSymbol(g_dbg_target, &t, sizeof(t));
    cudaMemcpyToSymbol(g_dbg_count, &z, sizeof(z));
    return 0;
}
extern "C" int sphb_debug_force_get(uint32_t *list_host, double *val_host)
{
    int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_dbg_count, sizeof(n));
    cudaMemcpyFromSymbol(list_host, g_dbg_list, sizeof(uint32_t) * 4096);
    cudaMemcpyFromSymbol(val_host, g_dbg_val, sizeof(double) * 4096);
    return n;
}
#define SPHB_DBG(jslot, islot, v)                                   \
    if ((long long)(jslot) == g_dbg_target) {                       \
        const int p_ = atomicAdd(&g_dbg_count, 1);                  \
        if (p_ < 4096) {                                            \
            g_dbg_list[p_] = (islot);                               \
            g_dbg_val[p_] = (v);                                    \
        }                                                           \
    }
#else
#define SPHB_DBG(jslot, islot, v)
#endif

// one pair, given the candidate's position record; the other two records are only fetched for real pairs.
// GRAV: the compact grad-h correction of the softened gravity (smoothed_gravity_correction_comp, pyfluid.py:434-437)
// has the same shape as the pressure term, -m (B_j gradW(h_j) + B_i gradW(h_i)) with B = (G/2) xi/Omega, and rides
// in the acceleration coefficient only (du/dt keeps the pure A_j).
// wi: the pair's weight inside the sums (m_i with per-particle masses, else 1: the uniform mass is applied per target);
// every term of the sums is linear in (s_j, s_i), so the weight enters there and nowhere else
template <bool GRAV>
__device__ __forceinline__ void sphb_force_pair(const double *recB, const double *recC, const double *bgrav, uint32_t i,
                                                double x, double y, double z, double hi, const sphb_consts &k,
                                                SphbForceAcc &a, double wi = 1.0)
{
    const double dx = a.xj - x, dy = a.yj - y, dz = a.zj - z;
    const double d2 = sphb_d2(dx, dy, dz);
    if (!sphb_nonzero(d2)) return;   // coincident positions contribute exactly 0 (pyfluid.py:51-52,461)
    // 1/h_i and 1/(pi h_i^4) are recomputed (7 FP64 instructions) rather than gathered: the third record is then
    // only needed by the viscous branch, i.e. for about half of the pairs
    const double hinvi = sphb_rcp(hi);
    const double hi2 = hinvi * hinvi;
    const double gi = (hi2 * hi2) * 0.3183098861837907;   // 1/pi
    const double rinv = sphb_rsqrt(d2);
    const double r = d2 * rinv;
    const double qj = r * a.hinvj, qi = r * hinvi;
    if (!(qj < 2.0) && !(qi < 2.0)) return;   // prefilter false positive
    double vxi, vyi, vzi, Ai;
    sphb_load4(recB, i, vxi, vyi, vzi, Ai);
    const double sj = wi * (sphb_m4c_fp(qj) * a.gj * rinv);
    const double si = wi * (sphb_m4c_fp(qi) * gi * rinv);
    const double dvx = a.vxj - vxi, dvy = a.vyj - vyi, dvz = a.vzj - vzi;
    const double vdr = sphb_dot3(dx, dy, dz, dvx, dvy, dvz);
    a.dens = fma(sj, vdr, a.dens);
    const double sm = sj + si;
    double pi_half = 0.0;
    if (vdr < 0.0) {
        double rhoi, csi, hinv_unused, g_unused;
        sphb_load4(recC, i, rhoi, csi, hinv_unused, g_unused);
        const double mean_h = 0.5 * (a.hj + hi);
        const double den = fma(k.epsilon * mean_h, mean_h, d2);
        const double mean_rho = 0.5 * (a.rhoj + rhoi);
        const double inv = sphb_rcp(den * mean_rho);
        const double mu = (mean_h * vdr) * (inv * mean_rho);
        const double cs = 0.5 * (csi + a.csj);
        const double pi = (mu * fma(k.beta, mu, -k.alpha * cs)) * (inv * den);
        if (pi < 0.0) a.negpi = 1;
        pi_half = 0.5 * pi;
    }
    double coef = fma(a.Aj, sj, fma(Ai, si, pi_half * sm));
    if (GRAV) coef = fma(a.Bj, sj, fma(__ldg(bgrav + i), si, coef));
    SPHB_DBG(a.jslot, i, coef * dx);
    a.ax = fma(-coef, dx, a.ax);
    a.ay = fma(-coef, dy, a.ay);
    a.az = fma(-coef, dz, a.az);
    a.visc = fma(pi_half * sm, vdr, a.visc);
}

// ---- two list entries side by side ---------------------------------------------------------------------
// The same operations per pair as sphb_force_pair, arranged as straight-line code for two candidates so that two
// independent FP64 dependency chains are in flight per lane; the accumulators are then updated in list order, which
// keeps every sum bit-identical to the one-at-a-time loop.  `ok` replaces the early returns.
struct SphbPairTerms {
    double dx, dy, dz, d2, hi, sj, si, vdr, Ai, pism, coef;
    bool ok;
};

// PRE: the candidate's {v, A} record was loaded ahead of time (vb) instead of after the membership test -- one
// dependent load latency less per pair, a few wasted loads for the candidates that fail the test; same bits
template <bool PRE = false>
__device__ __forceinline__ void sphb_force_geom(const double *recB, uint32_t i, const double (&p)[4], bool present,
                                                const SphbForceAcc &a, SphbPairTerms &t, const double *vb = nullptr,
                                                double wi = 1.0)
{
    t.dx = a.xj - p[0];
    t.dy = a.yj - p[1];
    t.dz = a.zj - p[2];
    t.hi = p[3];
    t.d2 = sphb_d2(t.dx, t.dy, t.dz);
    const double hinvi = sphb_rcp(t.hi);
    const double hi2 = hinvi * hinvi;
    const double gi = (hi2 * hi2) * 0.3183098861837907;   // 1/pi
    const double rinv = sphb_rsqrt(t.d2);
    const double r = t.d2 * rinv;
    const double qj = r * a.hinvj, qi = r * hinvi;
    t.ok = present && sphb_nonzero(t.d2) && ((qj < 2.0) || (qi < 2.0));
    double vxi = 0.0, vyi = 0.0, vzi = 0.0;
    t.Ai = 0.0;
    if (PRE) {
        vxi = vb[0];
        vyi = vb[1];
        vzi = vb[2];
        t.Ai = vb[3];
    } else if (t.ok) {
        sphb_load4(recB, i, vxi, vyi, vzi, t.Ai);
    }
    t.sj = wi * (sphb_m4c_fp(qj) * a.gj * rinv);
    t.si = wi * (sphb_m4c_fp(qi) * gi * rinv);
    const double dvx = a.vxj - vxi, dvy = a.vyj - vyi, dvz = a.vzj - vzi;
    t.vdr = sphb_dot3(t.dx, t.dy, t.dz, dvx, dvy, dvz);
}

__device__ __forceinline__ double sphb_force_pi(const sphb_consts &k, const SphbForceAcc &a, const SphbPairTerms &t,
                                                double rhoi, double csi)
{
    const double mean_h = 0.5 * (a.hj + t.hi);
    const double den = fma(k.epsilon * mean_h, mean_h, t.d2);
    const double mean_rho = 0.5 * (a.rhoj + rhoi);
    const double inv = sphb_rcp(den * mean_rho);
    const double mu = (mean_h * t.vdr) * (inv * mean_rho);
    const double cs = 0.5 * (csi + a.csj);
    return (mu * fma(k.beta, mu, -k.alpha * cs)) * (inv * den);
}

template <bool GRAV>
__device__ __forceinline__ void sphb_force_accumulate(const double *bgrav, uint32_t i, double pi_half,
                                                      const SphbPairTerms &t, SphbForceAcc &a)
{
    if (!t.ok) return;
    a.dens = fma(t.sj, t.vdr, a.dens);
    const double sm = t.sj + t.si;
    double coef = fma(a.Aj, t.sj, fma(t.Ai, t.si, pi_half * sm));
    if (GRAV) coef = fma(a.Bj, t.sj, fma(__ldg(bgrav + i), t.si, coef));
    SPHB_DBG(a.jslot, i, coef * t.dx);
    a.ax = fma(-coef, t.dx, a.ax);
    a.ay = fma(-coef, t.dy, a.ay);
    a.az = fma(-coef, t.dz, a.az);
    a.visc = fma(pi_half * sm, t.vdr, a.visc);
}

// DENSE: recC is a dense [n][2] array of (rho, cs) instead of the 32-byte record (a gathered 16-byte entry shares its
// cache line with seven neighbours instead of three)
// MASS: per-particle masses (k.mass != NULL): one more 8-byte gather per pair
// RTM: decide at run time (k.mass), for the scanning kernels: one predictable branch and two exact multiplications by 1
template <bool GRAV, bool DENSE = false, bool PRE = false, bool MASS = false, bool RTM = false>
__device__ __forceinline__ void sphb_force_pair2(const double *recB, const double *recC, const double *bgrav, uint32_t i0,
                                                 uint32_t i1, const double (&p0)[4], const double (&p1)[4], bool second,
                                                 const sphb_consts &k, SphbForceAcc &a, bool first = true,
                                                 const double *vb0 = nullptr, const double *vb1 = nullptr)
{
    SphbPairTerms t0, t1;
    const bool wm = MASS || (RTM && k.mass != nullptr);
    const double w0 = (wm && first) ? __ldg(k.mass + i0) : 1.0, w1 = (wm && second) ? __ldg(k.mass + i1) : 1.0;
    sphb_force_geom<PRE>(recB, i0, p0, first, a, t0, vb0, w0);
    sphb_force_geom<PRE>(recB, i1, p1, second, a, t1, vb1, w1);
    const bool v0 = t0.ok && t0.vdr < 0.0, v1 = t1.ok && t1.vdr < 0.0;
    double ph0 = 0.0, ph1 = 0.0;
    if (v0 || v1) {
        double rho0 = 1.0, cs0 = 0.0, rho1 = 1.0, cs1 = 0.0, u0, u1;
        if (DENSE) {
            if (v0) {
                const double2 d = __ldg(reinterpret_cast<const double2 *>(recC) + i0);
                rho0 = d.x;
                cs0 = d.y;
            }
            if (v1) {
                const double2 d = __ldg(reinterpret_cast<const double2 *>(recC) + i1);
                rho1 = d.x;
                cs1 = d.y;
            }
        } else {
            if (v0) sphb_load4(recC, i0, rho0, cs0, u0, u1);
            if (v1) sphb_load4(recC, i1, rho1, cs1, u0, u1);
        }
        const double pi0 = sphb_force_pi(k, a, t0, rho0, cs0);
        const double pi1 = sphb_force_pi(k, a, t1, rho1, cs1);
        if ((v0 && pi0 < 0.0) || (v1 && pi1 < 0.0)) a.negpi = 1;
        ph0 = v0 ? 0.5 * pi0 : 0.0;
        ph1 = v1 ? 0.5 * pi1 : 0.0;
    }
    sphb_force_accumulate<GRAV>(bgrav, i0, ph0, t0, a);
    sphb_force_accumulate<GRAV>(bgrav, i1, ph1, t1, a);
}

