/*
 * sph_oracle.c -- CPU oracle for the python-sph per-step particle loop.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C restatement of the reference's
 * O(N^2) algorithm (TRU-astrophysics/python-sph, pyfluid.py).  It exists so the CUDA
 * path can be checked on a box where /root/reference is absent, and so bench.py can
 * time a CPU baseline.  Nothing in the product path (python-sph_b200/) may import,
 * link or execute it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Parity pinning: the reference ships no golden vectors, so this
 * restatement is pinned against outputs of the reference itself, generated in the
 * build container by tests/golden/make_golden.py (which imports /root/reference/pyfluid.py)
 * and committed under tests/golden/.  tests/test_oracle_golden.py checks every fixture.
 *
 * Arithmetic follows the reference expression by expression, in Python's evaluation
 * order, all sums in index order 0..N-1 like the reference's `for i in range(N)` loops:
 *   - `x**k`            -> pow(x, k)      (numpy/python scalar power calls libm pow)
 *   - `d.dot(d)`, np.dot -> fma(z,z', fma(y,y', x*x'))  (OpenBLAS ddot, n=3; verified by
 *                           make_golden.py's FMA self-test on this image's numpy)
 *   - np.piecewise(q, [q>=2, 1<=q<2, q<1], ...) -> 0 when no condition holds (NaN q)
 * Each function cites the reference line range it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* pyfluid.py:5-41 -- the module-level constants, passed per call (the reference reads
 * its globals at call time). */
typedef struct {
    double particle_mass;   /* PARTICLE_MASS                       :5  */
    double coupling_const;  /* COUPLING_CONST (eta)                :9  */
    double h_tol;           /* SMOOTHLENGTH_VARIATION_TOLERANCE    :13 */
    int32_t newton_limit;   /* NEWTON_ITERATION_LIMIT              :14 */
    int32_t bisect_limit;   /* BISECTION_ITERATION_LIMIT           :23 */
    double alpha;           /* ALPHA_SPH                           :18 */
    double beta;            /* BETA_SPH                            :19 */
    double epsilon;         /* EPSILON                             :20 */
    double bisect_tol;      /* BISECTION_TOLERANCE                 :24 */
    double bisect_a;        /* INITIAL_BISECTION_SMOOTHLENGTH_A    :25 */
    double bisect_b;        /* INITIAL_BISECTION_SMOOTHLENGTH_B    :26 */
    double G;               /* G                                   :35 */
    double gamma;           /* ADIABATIC_INDEX                     :41 */
} oracle_consts;

/* status bits shared with the product's status word (include/sphb200.h) */
enum {
    ORC_NEG_H = 1,          /* "Obtained negative smoothing length."              :210 */
    ORC_NEG_P = 2,          /* "Obtained negative pressure."                      :258 */
    ORC_NEG_PI = 4,         /* "Obtained negative value for Pi"                   :282 */
    ORC_NEG_VISC = 8,       /* "Obtained negative energy rate due to viscosity."  :350 */
    ORC_NEG_E = 16,         /* "Obtained negative energy."                        :515,528 */
    ORC_BISECT_NONE = 32    /* bisection_h fell off its loop and returned None    :153 */
};

#define PI_D 3.141592653589793 /* np.pi */

/* "wide" mode (bench.py only): the O(N) inner loop of each per-particle function is split over the
 * OpenMP threads, so that a handful of sampled particles keeps every host core busy at N = 16 M.
 * It changes the summation order (chunked partial sums), nothing else; tests run with it off. */
static int g_wide = 0;
void oracle_set_wide(int on) { g_wide = on; }

/* Per-particle masses (the authors' own NOTE at pyfluid.py:3-4).  The reference has ONE scalar
 * PARTICLE_MASS; with a mass array set, every "PARTICLE_MASS" of a sum over i becomes m_i (pyfluid.py:95,216,327,340,474)
 * and the one in var_density / zetaprime becomes the target's m_j (:115,238).  NULL (the default) is the reference. */
static const double *g_mass = NULL;
void oracle_set_masses(const double *m) { g_mass = m; }
#define MASS_OF(c, i) (g_mass ? g_mass[i] : (c)->particle_mass)

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* pyfluid.py:43-47 distance(): sqrt(diff.dot(diff)) */
static inline double dot3(const double *a, const double *b)
{
    return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0]));
}

double oracle_distance(const double *rj, const double *ri)
{
    double d[3] = { rj[0] - ri[0], rj[1] - ri[1], rj[2] - ri[2] };
    return sqrt(dot3(d, d));
}

static inline int same_pos(const double *rj, const double *ri)
{
    /* np.array_equal: elementwise ==, so NaN != NaN and -0.0 == 0.0 */
    return rj[0] == ri[0] && rj[1] == ri[1] && rj[2] == ri[2];
}

/* pyfluid.py:50-54 direction_i_j(): zeros for identical positions, else (rj-ri)/dist */
static inline void direction_i_j(const double *rj, const double *ri, double *out)
{
    if (same_pos(rj, ri)) {
        out[0] = out[1] = out[2] = 0.0;
        return;
    }
    double dist = oracle_distance(rj, ri);
    out[0] = (rj[0] - ri[0]) / dist;
    out[1] = (rj[1] - ri[1]) / dist;
    out[2] = (rj[2] - ri[2]) / dist;
}

/* np.piecewise over the three M4 branches; value 0 if q is NaN */
static inline double pick3(double q, double v_ge2, double v_mid, double v_lt1)
{
    double out = 0.0;
    if (q >= 2.0) out = v_ge2;
    if (q >= 1.0 && q < 2.0) out = v_mid;
    if (q < 1.0) out = v_lt1;
    return out;
}

/* pyfluid.py:56-62 M4() */
double oracle_M4(double dist, double h)
{
    double q = dist / h;
    double mid = 0.25 * pow(2.0 - q, 3.0);
    double in = 1.0 - 1.5 * pow(q, 2.0) + 0.75 * pow(q, 3.0);
    return 1.0 / (PI_D * pow(h, 3.0)) * pick3(q, 0.0, mid, in);
}

/* pyfluid.py:64-70 M4_d1() */
double oracle_M4_d1(double dist, double h)
{
    double q = dist / h;
    double mid = -0.75 * pow(2.0 - q, 2.0);
    double in = 2.25 * pow(q, 2.0) - 3.0 * q;
    return 1.0 / (PI_D * pow(h, 4.0)) * pick3(q, 0.0, mid, in);
}

/* pyfluid.py:85-87 M4_smoothlength_derivative() (Cossins 3.107) */
double oracle_M4_dh(double dist, double h)
{
    double x = dist / h;
    return -x * oracle_M4_d1(dist, h) - 3.0 / h * oracle_M4(dist, h);
}

/* pyfluid.py:81-82 dellM4(): M4_d1(distance) * direction */
static inline void dellM4(const double *rj, const double *ri, double h, double *out)
{
    double w = oracle_M4_d1(oracle_distance(rj, ri), h);
    double dir[3];
    direction_i_j(rj, ri, dir);
    out[0] = w * dir[0];
    out[1] = w * dir[1];
    out[2] = w * dir[2];
}

/* pyfluid.py:237-238 var_density() */
static inline double var_density_m(const oracle_consts *c, double m, double h)
{
    return m * pow(c->coupling_const / h, 3.0);
}

/* pyfluid.py:218-224 density(): index-order sum including i == j */
static double density_j(const oracle_consts *c, const double *pos, int64_t n, int64_t j, double h)
{
    double rho = 0.0;
#pragma omp parallel for reduction(+ : rho) schedule(static) if (g_wide)
    for (int64_t i = 0; i < n; ++i)
        rho += MASS_OF(c, i) * oracle_M4(oracle_distance(pos + 3 * j, pos + 3 * i), h);
    return rho;
}

/* pyfluid.py:91-97 omega_j() */
static double omega_j(const oracle_consts *c, const double *pos, int64_t n, int64_t j, double rho, double h)
{
    double sum = 0.0;
#pragma omp parallel for reduction(+ : sum) schedule(static) if (g_wide)
    for (int64_t i = 0; i < n; ++i)
        sum += MASS_OF(c, i) * oracle_M4_dh(oracle_distance(pos + 3 * j, pos + 3 * i), h);
    return 1.0 + h / (3.0 * rho) * sum;
}

/* pyfluid.py:108-117 zeta_j(), zetaprime() */
static inline double zeta_j(const oracle_consts *c, double mj, double h, double rho)
{
    return var_density_m(c, mj, h) - rho;
}

static inline double zetaprime(const oracle_consts *c, double mj, double rho, double omega, double h)
{
    return (-3.0 * rho / h * (omega - 1.0) - 3.0 * mj * pow(c->coupling_const / h, 3.0) / h);
}

/* pyfluid.py:155-163 newton_smoothlength_iteration() */
static double newton_iteration(const oracle_consts *c, const double *pos, int64_t n, int64_t j, double h_old)
{
    double rho = density_j(c, pos, n, j, h_old);
    double om = omega_j(c, pos, n, j, rho, h_old);
    double zeta = zeta_j(c, MASS_OF(c, j), h_old, rho);
    double zetap = zetaprime(c, MASS_OF(c, j), rho, om, h_old);
    return h_old - zeta / zetap; /* newton_new_h :199-201 */
}

/* pyfluid.py:123-153 bisection_h().  Returns 1 and *h_out on success, 0 for the
 * fall-through `None`.  *oob counts the "ERROR: Root out of bisection bounds." prints. */
static int bisection_h(const oracle_consts *c, const double *pos, int64_t n, int64_t j, double *h_out, int32_t *oob)
{
    double h_a = c->bisect_a, h_b = c->bisect_b;
    int i = 0;
    while (i < c->bisect_limit) {
        double zeta_a = zeta_j(c, MASS_OF(c, j), h_a, density_j(c, pos, n, j, h_a));
        double zeta_b = zeta_j(c, MASS_OF(c, j), h_b, density_j(c, pos, n, j, h_b));
        if (zeta_a * zeta_b > 0.0) (*oob)++;
        double root = (h_a + h_b) / 2.0;
        double zeta_root = zeta_j(c, MASS_OF(c, j), root, density_j(c, pos, n, j, root));
        if (fabs(zeta_root) < c->bisect_tol) {
            *h_out = root;
            return 1;
        }
        if (zeta_a * zeta_root < 0.0)
            h_b = root;
        else if (zeta_b * zeta_root < 0.0)
            h_a = root;
        i += 1;
    }
    return 0;
}

/* pyfluid.py:165-179 newton_smoothlength_while().
 * iters: Newton iterations performed (1-based count of newton_smoothlength_iteration calls);
 * code: 0 = Newton converged, 1 = WARNING + bisection converged, 2 = bisection returned None. */
static double newton_while(const oracle_consts *c, const double *pos, int64_t n, int64_t j, double h0,
                           int32_t *iters, int32_t *code, int32_t *oob)
{
    double h_old = h0;
    int i = 0;
    while (i < c->newton_limit) {
        double h_cur = newton_iteration(c, pos, n, j, h_old);
        /* smoothlength_variation(old, cur) = |old - cur| / |cur|   :119-120, :172 */
        if (fabs(h_old - h_cur) / fabs(h_cur) < c->h_tol && h_cur > 0.0) {
            *iters = i + 1;
            *code = 0;
            return h_cur;
        }
        h_old = h_cur;
        i += 1;
    }
    *iters = i;
    double h = NAN;
    *code = bisection_h(c, pos, n, j, &h, oob) ? 1 : 2;
    return h;
}

/* ---- list entry points: evaluate the reference's per-particle functions for jlist[0..nj) ---- */

/* pyfluid.py:218 density(j, position_arr, smoothlength_j) */
void oracle_density_list(const oracle_consts *c, const double *pos, int64_t n, const int64_t *jlist, int64_t nj,
                         const double *h_j, double *rho_out)
{
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t t = 0; t < nj; ++t)
        rho_out[t] = density_j(c, pos, n, jlist[t], h_j[t]);
}

/* pyfluid.py:91 omega_j(j, position_arr, density_j, smoothlength_j) */
void oracle_omega_list(const oracle_consts *c, const double *pos, int64_t n, const int64_t *jlist, int64_t nj,
                       const double *rho_j, const double *h_j, double *omega_out)
{
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t t = 0; t < nj; ++t)
        omega_out[t] = omega_j(c, pos, n, jlist[t], rho_j[t], h_j[t]);
}

/* pyfluid.py:165 newton_smoothlength_while(j, position_arr, initial_smoothlength_j) */
void oracle_newton_list(const oracle_consts *c, const double *pos, int64_t n, const int64_t *jlist, int64_t nj,
                        const double *h_guess_j, double *h_out, int32_t *iters_out, int32_t *code_out, int32_t *oob_out)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < nj; ++t) {
        int32_t it = 0, code = 0, oob = 0;
        h_out[t] = newton_while(c, pos, n, jlist[t], h_guess_j[t], &it, &code, &oob);
        iters_out[t] = it;
        code_out[t] = code;
        if (oob_out) oob_out[t] = oob;
    }
}

/* pyfluid.py:123 bisection_h(j, position_arr) */
void oracle_bisection_list(const oracle_consts *c, const double *pos, int64_t n, const int64_t *jlist, int64_t nj,
                           double *h_out, int32_t *code_out, int32_t *oob_out)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < nj; ++t) {
        int32_t oob = 0;
        double h = NAN;
        code_out[t] = bisection_h(c, pos, n, jlist[t], &h, &oob) ? 1 : 2;
        h_out[t] = h;
        if (oob_out) oob_out[t] = oob;
    }
}

/* Neighbour-set oracle: NB(j) = { i : distance(pos[j], pos[i]) / h_j < 2 },
 * the complement of M4's `q >= 2` branch (pyfluid.py:57-59).  Writes ascending indices;
 * returns the count (which may exceed cap; only cap entries are written). */
int64_t oracle_neighbours(const double *pos, int64_t n, int64_t j, double h_j, int64_t *out, int64_t cap)
{
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        double q = oracle_distance(pos + 3 * j, pos + 3 * i) / h_j;
        if (q < 2.0) {
            if (cnt < cap) out[cnt] = i;
            cnt++;
        }
    }
    return cnt;
}

/* Force-support oracle: i in NB(j) or j in NB(i) (pair term non-zero iff q_j<2 or q_i<2, a17). */
int64_t oracle_neighbours_sym(const double *pos, const double *h, int64_t n, int64_t j, int64_t *out, int64_t cap)
{
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        double d = oracle_distance(pos + 3 * j, pos + 3 * i);
        if (d / h[j] < 2.0 || d / h[i] < 2.0) {
            if (cnt < cap) out[cnt] = i;
            cnt++;
        }
    }
    return cnt;
}

/* pyfluid.py:261-285 Pi() */
static double Pi_ji(const oracle_consts *c, int64_t j, int64_t i, const double *pos, const double *vel,
                    const double *P, const double *rho, const double *h, int32_t *status)
{
    double dr[3] = { pos[3 * j] - pos[3 * i], pos[3 * j + 1] - pos[3 * i + 1], pos[3 * j + 2] - pos[3 * i + 2] };
    double dv[3] = { vel[3 * j] - vel[3 * i], vel[3 * j + 1] - vel[3 * i + 1], vel[3 * j + 2] - vel[3 * i + 2] };
    double v_dot_r = dot3(dr, dv);
    if (v_dot_r < 0.0) {
        double mean_h = 0.5 * (h[j] + h[i]);
        double dist = oracle_distance(pos + 3 * j, pos + 3 * i);
        double mu = mean_h * v_dot_r / (pow(dist, 2.0) + c->epsilon * pow(mean_h, 2.0));
        double csj = pow(c->gamma * P[j] / rho[j], 0.5);
        double csi = pow(c->gamma * P[i] / rho[i], 0.5);
        double cs = 0.5 * (csi + csj);
        double mean_rho = 0.5 * (rho[j] + rho[i]);
        double val = (-c->alpha * cs * mu + c->beta * pow(mu, 2.0)) / mean_rho;
        if (val < 0.0) *status |= ORC_NEG_PI;
        return val;
    }
    return 0.0;
}

/* pyfluid.py:316-353 energy_rate() */
static double energy_rate_j(const oracle_consts *c, int64_t j, const double *pos, const double *vel, const double *P,
                            const double *rho, const double *h, const double *omega, int64_t n, int32_t *status)
{
    double density_change = 0.0, viscosity_sum = 0.0;
    int32_t stl = 0;
#pragma omp parallel for reduction(+ : density_change, viscosity_sum) reduction(| : stl) schedule(static) if (g_wide)
    for (int64_t i = 0; i < n; ++i) {
        double vji[3] = { vel[3 * j] - vel[3 * i], vel[3 * j + 1] - vel[3 * i + 1], vel[3 * j + 2] - vel[3 * i + 2] };
        double gj[3], gi[3], mean[3];
        dellM4(pos + 3 * j, pos + 3 * i, h[j], gj);
        dellM4(pos + 3 * j, pos + 3 * i, h[i], gi);
        density_change += MASS_OF(c, i) * dot3(vji, gj);
        for (int k = 0; k < 3; ++k) mean[k] = 0.5 * (gj[k] + gi[k]);
        viscosity_sum += (MASS_OF(c, i) * Pi_ji(c, j, i, pos, vel, P, rho, h, &stl) * dot3(vji, mean));
    }
    *status |= stl;
    viscosity_sum *= 0.5;
    if (viscosity_sum < 0.0) *status |= ORC_NEG_VISC;
    return (P[j] / (pow(rho[j], 2.0) * omega[j]) * density_change + viscosity_sum);
}

/* pyfluid.py:378-389 grav_force() -- selected branch only (the others have no side effects) */
static double grav_force(double dist, double h)
{
    double x = dist / h;
    if (x < 1.0)
        return 1.0 / pow(h, 2.0) * (4.0 / 3.0 * x - 6.0 / 5.0 * pow(x, 3.0) + 1.0 / 2.0 * pow(x, 4.0));
    if (x >= 1.0 && x < 2.0)
        return 1.0 / pow(h, 2.0)
               * (8.0 / 3.0 * x - 3.0 * pow(x, 2.0) + 6.0 / 5.0 * pow(x, 3.0) - 1.0 / 6.0 * pow(x, 4.0)
                  - 1.0 / (15.0 * pow(x, 2.0)));
    if (x >= 2.0) return 1.0 / pow(dist, 2.0);
    return 0.0;
}

/* pyfluid.py:391-399 grav_potential_smoothlength_derivative() */
static double grav_dphi_dh(double dist, double h)
{
    double x = dist / h;
    if (x < 1.0)
        return -1.0 / pow(h, 2.0)
               * (2.0 * pow(x, 2.0) - 12.0 / 10.0 * pow(x, 3.0) + 6.0 / 10.0 * pow(x, 5.0) - 7.0 / 5.0);
    if (x >= 1.0 && x < 2.0)
        return -1.0 / pow(h, 2.0)
               * (4.0 * pow(x, 2.0) - 4.0 * pow(x, 3.0) + 3.0 / 2.0 * pow(x, 4.0) - 1.0 / 5.0 * pow(x, 5.0)
                  - 8.0 / 5.0);
    return 0.0;
}

/* pyfluid.py:401-410 xi() */
static double xi_j(const oracle_consts *c, const double *pos, int64_t n, int64_t j, double rho, double h)
{
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i)
        sum += c->particle_mass * grav_dphi_dh(oracle_distance(pos + 3 * j, pos + 3 * i), h);
    return -h / (3.0 * rho) * sum;
}

/* pyfluid.py:477-485 acceleration(): fluid_accel_viscosity_comp :459-474 plus, when G != 0,
 * smoothed_gravity_acceleration_comp :420-432 and smoothed_gravity_correction_comp :434-437.
 * With G == 0 the two gravity terms are +-0.0 and adding them does not change the sum. */
static void acceleration_j(const oracle_consts *c, int64_t j, const double *pos, const double *rho, const double *vel,
                           const double *P, const double *h, const double *omega, const double *xi, int64_t n,
                           double *acc, int32_t *status)
{
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    int32_t stl = 0;
#pragma omp parallel for reduction(+ : a0, a1, a2) reduction(| : stl) schedule(static) if (g_wide)
    for (int64_t i = 0; i < n; ++i) {
        const double m = MASS_OF(c, i);
        double term[3] = { 0.0, 0.0, 0.0 };
        const double *rj = pos + 3 * j, *ri = pos + 3 * i;
        int same = same_pos(rj, ri);
        double gj[3], gi[3];
        if (!same || c->G != 0.0) {
            dellM4(rj, ri, h[j], gj);
            dellM4(rj, ri, h[i], gi);
        }
        if (!same) {
            double aj = P[j] / (omega[j] * pow(rho[j], 2.0));
            double ai = P[i] / (omega[i] * pow(rho[i], 2.0));
            double pi_half = Pi_ji(c, j, i, pos, vel, P, rho, h, &stl) * 0.5;
            for (int k = 0; k < 3; ++k) {
                double accel = aj * gj[k] + ai * gi[k];
                double visc = pi_half * (gj[k] + gi[k]);
                term[k] = -m * (accel + visc);
            }
        }
        if (c->G != 0.0) {
            double g1[3] = { 0.0, 0.0, 0.0 };
            if (!same) {
                double dist = oracle_distance(rj, ri);
                double dir[3];
                direction_i_j(rj, ri, dir);
                double f = grav_force(dist, h[j]) + grav_force(dist, h[i]);
                for (int k = 0; k < 3; ++k) g1[k] = -c->G / 2.0 * m * (f * dir[k]);
            }
            double xoj = xi[j] / omega[j], xoi = xi[i] / omega[i];
            for (int k = 0; k < 3; ++k) {
                double g2 = -c->G / 2.0 * m * (xoj * gj[k] + xoi * gi[k]);
                term[k] = (term[k] + g1[k]) + g2;
            }
        }
        a0 += term[0];
        a1 += term[1];
        a2 += term[2];
    }
    acc[0] = a0;
    acc[1] = a1;
    acc[2] = a2;
    *status |= stl;
}

/* pyfluid.py:316 energy_rate(j, ...) for a list of j */
void oracle_energy_rate_list(const oracle_consts *c, const double *pos, const double *vel, const double *P,
                             const double *rho, const double *h, const double *omega, int64_t n,
                             const int64_t *jlist, int64_t nj, double *out, int32_t *status_out)
{
    int32_t status = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(| : status)
    for (int64_t t = 0; t < nj; ++t) {
        int32_t st = 0;
        out[t] = energy_rate_j(c, jlist[t], pos, vel, P, rho, h, omega, n, &st);
        status |= st;
    }
    if (status_out) *status_out = status;
}

/* pyfluid.py:477 acceleration(j, ...) for a list of j; out is (nj,3) */
void oracle_acceleration_list(const oracle_consts *c, const double *pos, const double *rho, const double *vel,
                              const double *P, const double *h, const double *omega, const double *xi, int64_t n,
                              const int64_t *jlist, int64_t nj, double *out, int32_t *status_out)
{
    int32_t status = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(| : status)
    for (int64_t t = 0; t < nj; ++t) {
        int32_t st = 0;
        acceleration_j(c, jlist[t], pos, rho, vel, P, h, omega, xi, n, out + 3 * t, &st);
        status |= st;
    }
    if (status_out) *status_out = status;
}

/* pyfluid.py:401 xi(j, position_arr, density_j, smoothlength_j) for a list of j */
void oracle_xi_list(const oracle_consts *c, const double *pos, int64_t n, const int64_t *jlist, int64_t nj,
                    const double *rho_j, const double *h_j, double *out)
{
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t t = 0; t < nj; ++t)
        out[t] = xi_j(c, pos, n, jlist[t], rho_j[t], h_j[t]);
}

/* One derivative evaluation of the step: the six calls at pyfluid.py:502-507 / :517-522.
 * Returns the status bits raised on the way (the reference raises at the first one). */
static int32_t stage(const oracle_consts *c, const double *pos, const double *vel, const double *e, const double *h,
                     int64_t n, double *rho, double *omega, double *xi, double *P, double *dudt, double *acc)
{
    int32_t status = 0;
    for (int64_t i = 0; i < n; ++i) rho[i] = var_density_m(c, MASS_OF(c, i), h[i]);        /* var_density_arr :240 */
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t j = 0; j < n; ++j) omega[j] = omega_j(c, pos, n, j, rho[j], h[j]); /* omega_arr :99 */
    if (c->G != 0.0) {
#pragma omp parallel for schedule(dynamic, 8)
        for (int64_t j = 0; j < n; ++j) xi[j] = xi_j(c, pos, n, j, rho[j], h[j]);  /* xi_arr :412 */
    } else {
        for (int64_t j = 0; j < n; ++j) xi[j] = 0.0; /* multiplied by G == 0 at :436 */
    }
    for (int64_t i = 0; i < n; ++i) {                                       /* pressure_arr :251 */
        P[i] = (c->gamma - 1.0) * e[i] * rho[i];
        if (P[i] < 0.0) status |= ORC_NEG_P;
    }
    if (status) return status;
#pragma omp parallel for schedule(dynamic, 4) reduction(| : status)
    for (int64_t j = 0; j < n; ++j) {
        int32_t st = 0;
        dudt[j] = energy_rate_j(c, j, pos, vel, P, rho, h, omega, n, &st); /* energy_rate_arr :355 */
        acceleration_j(c, j, pos, rho, vel, P, h, omega, xi, n, acc + 3 * j, &st); /* acceleration_arr :487 */
        status |= st;
    }
    return status;
}

/* pyfluid.py:203-211 newton_smoothlength_arr() */
int32_t oracle_newton_arr(const oracle_consts *c, const double *pos, int64_t n, const double *h_guess, double *h_out,
                          int32_t *iters_out, int32_t *code_out)
{
    int32_t status = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : status)
    for (int64_t j = 0; j < n; ++j) {
        int32_t it = 0, code = 0, oob = 0;
        h_out[j] = newton_while(c, pos, n, j, h_guess[j], &it, &code, &oob);
        if (iters_out) iters_out[j] = it;
        if (code_out) code_out[j] = code;
        if (code == 2) status |= ORC_BISECT_NONE;
        if (h_out[j] < 0.0) status |= ORC_NEG_H;
    }
    return status;
}

/* pyfluid.py:498-532 var_smoothlength_leapfrog().  Returns the status bits; outputs are
 * only meaningful when it returns 0 (the reference raises instead of returning). */
int32_t oracle_leapfrog(const oracle_consts *c, const double *pos0, const double *vel0, const double *e0,
                        const double *h0, double dt, int64_t n, double *pos1, double *vel1, double *e1, double *h1)
{
    size_t N = (size_t)n;
    double *buf = (double *)malloc(sizeof(double) * N * (1 + 1 + 1 + 1 + 1 + 3 + 3 + 3 + 1 + 1));
    if (!buf) return -1;
    double *rho = buf, *omega = rho + N, *xi = omega + N, *P = xi + N, *dudt = P + N, *acc = dudt + N;
    double *pos_mid = acc + 3 * N, *vel_mid = pos_mid + 3 * N, *e_mid = vel_mid + 3 * N, *h_mid = e_mid + N;
    int32_t status = stage(c, pos0, vel0, e0, h0, n, rho, omega, xi, P, dudt, acc);
    if (status) goto done;
    for (size_t k = 0; k < 3 * N; ++k) {
        pos_mid[k] = pos0[k] + vel0[k] * 0.5 * dt;  /* :511 */
        vel_mid[k] = vel0[k] + acc[k] * 0.5 * dt;   /* :512 */
    }
    for (size_t k = 0; k < N; ++k) {
        e_mid[k] = e0[k] + dudt[k] * 0.5 * dt;      /* :513 */
        if (e_mid[k] < 0.0) status |= ORC_NEG_E;    /* :514 */
    }
    if (status) goto done;
    status = oracle_newton_arr(c, pos_mid, n, h0, h_mid, NULL, NULL); /* :516 */
    if (status) goto done;
    status = stage(c, pos_mid, vel_mid, e_mid, h_mid, n, rho, omega, xi, P, dudt, acc);
    if (status) goto done;
    for (size_t k = 0; k < 3 * N; ++k) {
        vel1[k] = vel0[k] + acc[k] * dt;                     /* :524 */
        pos1[k] = pos0[k] + 0.5 * (vel0[k] + vel1[k]) * dt;  /* :525 */
    }
    for (size_t k = 0; k < N; ++k) {
        e1[k] = e0[k] + dudt[k] * dt;                        /* :526 */
        if (e1[k] < 0.0) status |= ORC_NEG_E;                /* :527 */
    }
    if (status) goto done;
    status = oracle_newton_arr(c, pos1, n, h_mid, h1, NULL, NULL);    /* :530 */
done:
    free(buf);
    return status;
}

/* The per-particle share of one step for a sample of particles, at frozen full-N state:
 * 2 x (omega_j, energy_rate, acceleration) + 2 x newton_smoothlength_while -- what
 * var_smoothlength_leapfrog costs the reference per particle (pyfluid.py:502-530).  Used by
 * bench.py's cpu_baseline on a bounded sample.  Outputs are written so nothing is elided. */
void oracle_step_work_sample(const oracle_consts *c, const double *pos, const double *vel, const double *e,
                             const double *h, int64_t n, const int64_t *jlist, int64_t nj, double *sink)
{
    size_t N = (size_t)n;
    double *rho = (double *)malloc(sizeof(double) * N * 4);
    double *P = rho + N, *omega = P + N, *xi = omega + N;
    for (size_t i = 0; i < N; ++i) {
        rho[i] = var_density_m(c, MASS_OF(c, i), h[i]);
        P[i] = (c->gamma - 1.0) * e[i] * rho[i];
        omega[i] = 1.0; /* frozen stand-in for the neighbours' omega; the cost does not depend on it */
        xi[i] = 0.0;
    }
#pragma omp parallel for schedule(dynamic, 1) if (!g_wide)
    for (int64_t t = 0; t < nj; ++t) {
        int64_t j = jlist[t];
        double acc[3], s = 0.0;
        int32_t st = 0, it = 0, code = 0, oob = 0;
        for (int rep = 0; rep < 2; ++rep) {
            double om = omega_j(c, pos, n, j, rho[j], h[j]);
            s += om;
            if (c->G != 0.0) s += xi_j(c, pos, n, j, rho[j], h[j]);
            s += energy_rate_j(c, j, pos, vel, P, rho, h, omega, n, &st);
            acceleration_j(c, j, pos, rho, vel, P, h, omega, xi, n, acc, &st);
            s += acc[0] + acc[1] + acc[2];
            s += newton_while(c, pos, n, j, h[j], &it, &code, &oob);
        }
        sink[t] = s;
    }
    free(rho);
}
