// sphb_common.cuh -- shared device helpers for libsphb200 (sm_100a only).
//
// Pair math follows pyfluid.py of TRU-astrophysics/python-sph; each helper cites the lines it
// restates.  Prefactors 1/(pi h^3), 1/(pi h^4) are hoisted out of the sums (SURVEY.md 8(a) note).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/sphb200.h"

#define SPHB_FULL 0xffffffffu
#define SPHB_PI 3.141592653589793

extern thread_local char g_sphb_err[256];
extern long long g_sphb_launches;

#define SPHB_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            snprintf(g_sphb_err, sizeof(g_sphb_err), "%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return SPHB_E_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

#define SPHB_LAUNCHED()                                                                              \
    do {                                                                                             \
        ++g_sphb_launches;                                                                           \
        SPHB_CUDA(cudaGetLastError());                                                               \
    } while (0)

static inline cudaStream_t sphb_stream(void *s) { return (cudaStream_t)s; }

// one NVTX range per C-ABI entry point: shows up in Nsight Systems / ncu --nvtx timelines,
// costs two empty calls when no tool is attached
struct SphbNvtx {
    explicit SphbNvtx(const char *name) { nvtxRangePushA(name); }
    ~SphbNvtx() { nvtxRangePop(); }
};
#define SPHB_RANGE() SphbNvtx sphb_nvtx_range_(__func__)

// x edge of a cell: cell / xdiv (IEEE division, the same on host and device)
__host__ __device__ __forceinline__ double sphb_cellx(const sphb_grid &g)
{
    return g.cell / (double)(g.xdiv > 0 ? g.xdiv : 1);
}

// ---------------------------------------------------------------------------------------------
// nested cell grids (sphb_cells.nlevels > 1, include/sphb200.h): level of a particle = octave of its h
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int sphb_level_of(double h, int exp0, int nlev)
{
    if (!(h > 0.0) || !(h < CUDART_INF)) return nlev - 1;               // unbounded support: last level
    const int e = ((__double2hiint(h) >> 20) & 0x7ff) - 1023;           // floor(log2 h) for normal h
    return min(max(e - exp0, 0), nlev - 1);
}

__device__ __forceinline__ sphb_level sphb_load_level(const sphb_level *lev, int b)
{
    sphb_level L;
    const double *d = reinterpret_cast<const double *>(lev + b);
    L.grid.ox = __ldg(d);
    L.grid.oy = __ldg(d + 1);
    L.grid.oz = __ldg(d + 2);
    L.grid.cell = __ldg(d + 3);
    const int2 q = __ldg(reinterpret_cast<const int2 *>(d + 4));   // the 56-byte entries are 8-byte aligned only
    const int2 q2 = __ldg(reinterpret_cast<const int2 *>(d + 5));
    L.grid.nx = q.x;
    L.grid.ny = q.y;
    L.grid.nz = q2.x;
    L.grid.xdiv = 1;
    const uint2 r = __ldg(reinterpret_cast<const uint2 *>(d + 6));
    L.cell_base = r.x;
    L.ncell = r.y;
    return L;
}

// level of a global cell id (levels are numbered by ascending cell_base)
__device__ __forceinline__ int sphb_level_of_cell(const sphb_level *lev, int nlev, uint32_t cell)
{
    int b = 0;
    for (int q = 1; q < nlev; ++q)
        if (cell >= __ldg(&lev[q].cell_base)) b = q;
    return b;
}

// ---------------------------------------------------------------------------------------------
// exact reference geometry (membership tests, pyfluid.py:43-47): d.dot(d) is
// fma(dz,dz,fma(dy,dy,dx*dx)) on the reference's numpy (tests/golden/make_golden.py self-test)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double sphb_d2(double dx, double dy, double dz)
{
    return __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
}

__device__ __forceinline__ double sphb_dot3(double ax, double ay, double az, double bx, double by, double bz)
{
    return __fma_rn(az, bz, __fma_rn(ay, by, __dmul_rn(ax, bx)));
}

// 1/sqrt(x) for normal positive x: MUFU.RSQ64H seed (~2^-20) + one third-order step -> ~1e-17 rel.
__device__ __forceinline__ double sphb_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = x * y;
    double e = fma(-t, y, 1.0);          // 1 - x y^2
    double p = fma(0.375, e, 0.5);       // 1/2 + 3/8 e
    double ye = y * e;
    return fma(ye, p, y);                // y (1 + e/2 + 3/8 e^2)
}

// 1/x for normal x: MUFU.RCP64H seed + one third-order step
__device__ __forceinline__ double sphb_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    double t = fma(e, e, e);
    return fma(y, t, y);
}

// x != 0 for a non-negative double, on the integer pipe (a DSETP would take an FP64 issue slot)
__device__ __forceinline__ bool sphb_nonzero(double x) { return (__double2hiint(x) | __double2loint(x)) != 0; }

// max(a, 0) for a double on the integer pipe: 3 instructions instead of the 9 of fmax()
__device__ __forceinline__ double sphb_clamp0(double a)
{
    const int hi = __double2hiint(a), lo = __double2loint(a);
    const bool pos = hi >= 0;   // sign bit clear (a NaN with a clear sign bit stays NaN)
    return __hiloint2double(pos ? hi : 0, pos ? lo : 0);
}

// one 32-byte record (4 doubles) with a single 256-bit read-only load (LDG.E.256, sm_100+): a gathered
// record costs one L1 wavefront per touched line instead of two
__device__ __forceinline__ void sphb_ld4(const double *rec, size_t i, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(rec + 4 * i));
}

// ---------------------------------------------------------------------------------------------
// M4 cubic spline, dimensionless parts (pyfluid.py:56-70): W = f(q)/(pi h^3), W' = fp(q)/(pi h^4)
// np.piecewise semantics: q >= 2 -> 0, 1 <= q < 2 -> middle, q < 1 -> inner (also negative q), NaN -> 0
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sphb_m4(double q, double &f, double &fp)
{
    double q2 = q * q;
    double fin = fma(q2, fma(0.75, q, -1.5), 1.0);   // 1 - 3/2 q^2 + 3/4 q^3
    double fpin = q * fma(2.25, q, -3.0);            // 9/4 q^2 - 3 q
    double t = 2.0 - q;
    double t2 = t * t;
    double fmid = (0.25 * t) * t2;                   // 1/4 (2-q)^3
    double fpmid = -0.75 * t2;                       // -3/4 (2-q)^2
    bool inner = q < 1.0;
    bool mid = (q >= 1.0) && (q < 2.0);
    f = inner ? fin : (mid ? fmid : 0.0);
    fp = inner ? fpin : (mid ? fpmid : 0.0);
}

// The same functions in clamp form, for the hot loops: with t2 = max(2-q, 0), t1 = max(1-q, 0)
//   f = t2^3/4 - t1^3,   f' = 3 t1^2 - 3/4 t2^2      (identical polynomials on every branch)
// two integer-pipe clamps instead of six 64-bit selects; the inner-branch cancellation costs ~1e-16
// absolute, i.e. nothing at the 1e-10 parity bar.
__device__ __forceinline__ void sphb_m4c(double q, double &f, double &fp)
{
    const double t2 = sphb_clamp0(2.0 - q), t1 = sphb_clamp0(1.0 - q);
    const double s2 = t2 * t2, s1 = t1 * t1;
    f = fma(-s1, t1, (0.25 * t2) * s2);
    fp = fma(3.0, s1, -0.75 * s2);
}
__device__ __forceinline__ double sphb_m4c_fp(double q)
{
    const double t2 = sphb_clamp0(2.0 - q), t1 = sphb_clamp0(1.0 - q);
    return fma(3.0, t1 * t1, -0.75 * (t2 * t2));
}

__device__ __forceinline__ double sphb_m4_fp(double q)
{
    double fpin = q * fma(2.25, q, -3.0);
    double t = 2.0 - q;
    double fpmid = -0.75 * (t * t);
    bool inner = q < 1.0;
    bool mid = (q >= 1.0) && (q < 2.0);
    return inner ? fpin : (mid ? fpmid : 0.0);
}

// pyfluid.py:237-238 var_density(): m (eta/h)^3
__device__ __forceinline__ double sphb_var_density_m(const sphb_consts &k, double m, double h)
{
    double t = k.coupling_const / h;
    return m * (t * t * t);
}
__device__ __forceinline__ double sphb_var_density(const sphb_consts &k, double h)
{
    return sphb_var_density_m(k, k.particle_mass, h);
}

// per-particle masses (sphb_consts.mass): the mass of particle i; the weight a pair term carries inside a sum (m_i, or
// exactly 1 when the uniform mass is applied once after the sum, as the kernels always did); and that final factor
__device__ __forceinline__ double sphb_mass_of(const sphb_consts &k, size_t i) { return k.mass ? __ldg(k.mass + i) : k.particle_mass; }
__device__ __forceinline__ double sphb_mass_w(const sphb_consts &k, size_t i) { return k.mass ? __ldg(k.mass + i) : 1.0; }
__device__ __forceinline__ double sphb_mass_after(const sphb_consts &k) { return k.mass ? 1.0 : k.particle_mass; }

// ---------------------------------------------------------------------------------------------
// warp reductions (all 32 lanes participate)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int sphb_f2ord(float f)
{
    int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float sphb_ord2f(int i)
{
    return __int_as_float(i ^ ((i >> 31) & 0x7fffffff));
}
__device__ __forceinline__ float sphb_warp_min(float v)
{
    return sphb_ord2f(__reduce_min_sync(SPHB_FULL, sphb_f2ord(v)));
}
__device__ __forceinline__ float sphb_warp_max(float v)
{
    return sphb_ord2f(__reduce_max_sync(SPHB_FULL, sphb_f2ord(v)));
}
__device__ __forceinline__ double sphb_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SPHB_FULL, v, o);
    return v;
}

// float upper bounds
__device__ __forceinline__ float sphb_f_up(double x) { return __double2float_ru(x); }

// fp32 prefilter threshold for a support radius 2h: ((2h)(1+2^-20) + delta)^2 (1 + 1e-5), rounded up.
// !(h > 0) (negative, zero, NaN) means "unbounded support" in the reference's formulas (q < 1 for
// every distance when h < 0, pyfluid.py:57-60): +inf.
__device__ __forceinline__ float sphb_thr_of_h(double h, float delta)
{
    if (!(h > 0.0)) return CUDART_INF_F;
    float R = __fadd_ru(__fmul_ru(sphb_f_up(2.0 * h), 1.000001f), delta);
    float t = __fmul_ru(__fmul_ru(R, R), 1.00001f);
    return t;
}
__device__ __forceinline__ float sphb_radius_of_h(double h, float delta)
{
    if (!(h > 0.0)) return CUDART_INF_F;
    return __fmul_ru(__fadd_ru(__fmul_ru(sphb_f_up(2.0 * h), 1.000001f), delta), 1.00001f);
}
