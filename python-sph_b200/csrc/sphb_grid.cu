// sphb_grid.cu -- K1: cell-list neighbour build (cell keys, counting sort, cell start table) plus the
// permutation helpers.  The reference has no neighbour structure (pyfluid.py loops over all N,
// :93,:221,:325,:405,:480); the cell assignment rule is this repo's own and is pinned by a numpy
// restatement in tests/ (SURVEY.md 8(c) "self-oracle").
//
// Pipeline (all stream-ordered, no host sync):
//   k_cell_count    cell id per particle (fp64: floor((x-o)/cell), clamped) + histogram (atomicAdd)
//   k_scan_*        exclusive scan of the histogram -> cell_start[ncell+1]
//   k_cell_scatter  slot = cell_start[id] + atomicAdd(cursor[id]) -> unordered-within-cell permutation
//   k_cell_finish   one warp per cell: order the cell's entries by input index (=> the permutation is the
//                   STABLE sort by cell id, independent of atomic ordering), gather the fp64 record,
//                   emit the fp32 prefilter record, the per-slot cell id and the cell's h_max.
#include <stdio.h>

#include "sphb_common.cuh"

thread_local char g_sphb_err[256] = "";
long long g_sphb_launches = 0;

extern "C" const char *sphb_last_cuda_error(void) { return g_sphb_err; }
extern "C" int sphb_abi_version(void) { return SPHB_ABI_VERSION; }
extern "C" int64_t sphb_launch_count(void) { return g_sphb_launches; }

extern "C" int sphb_device_ok(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

extern "C" int sphb_status_reset(sphb_status *status, void *stream)
{
    if (!status) return SPHB_E_ARG;
    SPHB_CUDA(cudaMemsetAsync(status, 0, sizeof(sphb_status), sphb_stream(stream)));
    return SPHB_OK;
}

extern "C" int64_t sphb_grid_ncell(const sphb_grid *g)
{
    if (!g || g->nx <= 0 || g->ny <= 0 || g->nz <= 0) return -1;
    return (int64_t)g->nx * g->ny * g->nz;
}

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

static inline int64_t scan_blocks(int64_t ncell) { return (ncell + SCAN_TILE - 1) / SCAN_TILE; }
static inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

// workspace layout: counts[ncell] | cursor[ncell] | block_sums[nblk+1] | perm_tmp[n] | cell_tmp[n]
extern "C" int64_t sphb_levels_workspace_bytes(int64_t n, int64_t ncell);
extern "C" int64_t sphb_grid_workspace_bytes(int64_t n, const sphb_grid *g)
{
    return sphb_levels_workspace_bytes(n, sphb_grid_ncell(g));
}

__device__ __forceinline__ uint32_t sphb_cell_of(const sphb_grid &g, double x, double y, double z)
{
    // c = clamp(floor((x - origin)/cell), 0, dim-1); IEEE subtract and divide, no contraction
    double cx = floor(__ddiv_rn(__dsub_rn(x, g.ox), sphb_cellx(g)));
    double cy = floor(__ddiv_rn(__dsub_rn(y, g.oy), g.cell));
    double cz = floor(__ddiv_rn(__dsub_rn(z, g.oz), g.cell));
    cx = fmin(fmax(cx, 0.0), (double)(g.nx - 1));   // NaN -> 0
    cy = fmin(fmax(cy, 0.0), (double)(g.ny - 1));
    cz = fmin(fmax(cz, 0.0), (double)(g.nz - 1));
    return (uint32_t)(((long long)cz * g.ny + (long long)cy) * g.nx + (long long)cx);
}

// what the build kernels need to place a particle: one grid (nlev <= 1), or the device table of nested grids
struct SphbSpec {
    sphb_grid g0;            // the one grid / the origin of the fp32 frame
    int nlev, exp0;
    const sphb_level *lev;   // device [nlev]
    float *level_hmax;       // device [nlev] (written)
};

static inline SphbSpec sphb_spec_single(const sphb_grid &g)
{
    SphbSpec s;
    s.g0 = g;
    s.nlev = 1;
    s.exp0 = 0;
    s.lev = nullptr;
    s.level_hmax = nullptr;
    return s;
}

// global cell id of a particle
__device__ __forceinline__ uint32_t sphb_cell_of_spec(const SphbSpec &sp, double x, double y, double z, double h)
{
    if (sp.nlev <= 1) return sphb_cell_of(sp.g0, x, y, z);
    const sphb_level L = sphb_load_level(sp.lev, sphb_level_of(h, sp.exp0, sp.nlev));
    return L.cell_base + sphb_cell_of(L.grid, x, y, z);
}

// raise level_hmax[b] to hm (non-negative floats order like their bit patterns; look before the atomic)
__device__ __forceinline__ void sphb_raise_level_hmax(float *level_hmax, int b, float hm)
{
    int *p = reinterpret_cast<int *>(level_hmax) + b;
    const int v = __float_as_int(hm);
    if (v > __ldcg(p)) atomicMax(p, v);
}

__global__ void k_pack_posh(const double *__restrict__ pos, const double *__restrict__ h, long long n,
                            double *__restrict__ posh)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 a = make_double2(pos[3 * i], pos[3 * i + 1]);
    double2 b = make_double2(pos[3 * i + 2], h[i]);
    reinterpret_cast<double2 *>(posh)[2 * i] = a;
    reinterpret_cast<double2 *>(posh)[2 * i + 1] = b;
}

__global__ void k_cell_count(const double *__restrict__ posh, long long n, SphbSpec sp, uint32_t *__restrict__ cell_tmp,
                             uint32_t *__restrict__ counts)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 a = reinterpret_cast<const double2 *>(posh)[2 * i];
    const double2 b = reinterpret_cast<const double2 *>(posh)[2 * i + 1];
    const uint32_t c = sphb_cell_of_spec(sp, a.x, a.y, b.x, b.y);
    cell_tmp[i] = c;
    atomicAdd(counts + c, 1u);
}

// --- exclusive scan, three passes ----------------------------------------------------------------
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *sh /*[SCAN_THREADS/32]*/, uint32_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(SPHB_FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const uint32_t s = sh[w];
        if (w < warp) woff += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return woff + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
    k_scan_reduce(const uint32_t *__restrict__ counts, long long ncell, uint32_t *__restrict__ block_sums)
{
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < ncell) v += counts[base + k];
    uint32_t total;
    block_excl_scan(v, sh, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocks(uint32_t *__restrict__ block_sums, long long nblk)
{
    // single block: exclusive scan of block_sums[0..nblk) in place, total -> block_sums[nblk]
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    uint32_t carry = 0;
    for (long long base = 0; base < nblk; base += SCAN_THREADS) {
        const long long i = base + threadIdx.x;
        const uint32_t v = i < nblk ? block_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_excl_scan(v, sh, total);
        if (i < nblk) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) block_sums[nblk] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
    k_scan_apply(const uint32_t *__restrict__ counts, long long ncell, const uint32_t *__restrict__ block_sums,
                 uint32_t *__restrict__ cell_start)
{
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    uint32_t c[SCAN_ITEMS];
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        c[k] = (base + k < ncell) ? counts[base + k] : 0u;
        v += c[k];
    }
    uint32_t total;
    uint32_t ex = block_excl_scan(v, sh, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < ncell) cell_start[base + k] = ex;
        ex += c[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) cell_start[ncell] = block_sums[gridDim.x];
}

__global__ void k_cell_scatter(const uint32_t *__restrict__ cell_tmp, long long n, const uint32_t *__restrict__ cell_start,
                               uint32_t *__restrict__ cursor, uint32_t *__restrict__ perm_tmp)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cell_tmp[i];
    const uint32_t slot = cell_start[c] + atomicAdd(cursor + c, 1u);
    perm_tmp[slot] = (uint32_t)i;
}

#define FIN_WARPS 8
// cells larger than this keep the (valid, but atomics-ordered) scatter order
#define FIN_SORT_MAX 8192

__device__ __forceinline__ void sphb_emit_records(const double *__restrict__ posh_in, uint32_t src, uint32_t slot,
                                                  const sphb_grid &g, float delta, uint32_t cell,
                                                  uint32_t *__restrict__ perm, uint32_t *__restrict__ cell_id,
                                                  double *__restrict__ posh, float *__restrict__ frec, float &hm)
{
    const double2 a = reinterpret_cast<const double2 *>(posh_in)[2 * (size_t)src];
    const double2 b = reinterpret_cast<const double2 *>(posh_in)[2 * (size_t)src + 1];
    reinterpret_cast<double2 *>(posh)[2 * (size_t)slot] = a;
    reinterpret_cast<double2 *>(posh)[2 * (size_t)slot + 1] = b;
    float4 f;
    f.x = (float)(a.x - g.ox);
    f.y = (float)(a.y - g.oy);
    f.z = (float)(b.x - g.oz);
    f.w = sphb_thr_of_h(b.y, delta);
    reinterpret_cast<float4 *>(frec)[slot] = f;
    perm[slot] = src;
    cell_id[slot] = cell;
    if (b.y > 0.0) hm = fmaxf(hm, sphb_f_up(b.y));
    else if (!(b.y > 0.0)) hm = CUDART_INF_F;   // unbounded support
}

__global__ void __launch_bounds__(FIN_WARPS * 32)
    k_cell_finish(const double *__restrict__ posh_in, const long long *__restrict__ tie_key, SphbSpec sp, long long ncell,
                  float delta,
                  const uint32_t *__restrict__ cell_start, const uint32_t *__restrict__ perm_tmp,
                  uint32_t *__restrict__ perm, uint32_t *__restrict__ cell_id, float *__restrict__ cell_hmax,
                  double *__restrict__ posh, float *__restrict__ frec)
{
    const int lane = threadIdx.x & 31;
    const long long cell = (long long)blockIdx.x * FIN_WARPS + (threadIdx.x >> 5);
    if (cell >= ncell) return;
    const uint32_t s = cell_start[cell], e = cell_start[cell + 1];
    const uint32_t cnt = e - s;
    float hm = 0.f;
    if (cnt == 0) {
        if (lane == 0) cell_hmax[cell] = 0.f;
        return;
    }
    const sphb_grid &g = sp.g0;   // only its origin is used below (the fp32 frame)
    if (cnt <= 32) {
        const uint32_t v = lane < cnt ? perm_tmp[s + lane] : 0xffffffffu;
        const long long key = (lane < cnt) ? (tie_key ? tie_key[v] : (long long)v) : 0x7fffffffffffffffll;
        uint32_t rank = 0;
        for (uint32_t k = 0; k < cnt; ++k) {
            const long long o = __shfl_sync(SPHB_FULL, key, k);
            rank += (o < key) ? 1u : 0u;
        }
        if (lane < cnt) sphb_emit_records(posh_in, v, s + rank, g, delta, (uint32_t)cell, perm, cell_id, posh, frec, hm);
    } else if (cnt <= FIN_SORT_MAX) {
        for (uint32_t a = lane; a < cnt; a += 32) {
            const uint32_t v = perm_tmp[s + a];
            const long long key = tie_key ? tie_key[v] : (long long)v;
            uint32_t rank = 0;
            for (uint32_t k = 0; k < cnt; ++k) {
                const uint32_t o = perm_tmp[s + k];
                rank += ((tie_key ? tie_key[o] : (long long)o) < key) ? 1u : 0u;
            }
            sphb_emit_records(posh_in, v, s + rank, g, delta, (uint32_t)cell, perm, cell_id, posh, frec, hm);
        }
    } else {
        for (uint32_t a = lane; a < cnt; a += 32)
            sphb_emit_records(posh_in, perm_tmp[s + a], s + a, g, delta, (uint32_t)cell, perm, cell_id, posh, frec, hm);
    }
    hm = sphb_warp_max(hm);
    if (lane == 0) {
        cell_hmax[cell] = hm;
        if (sp.nlev > 1) sphb_raise_level_hmax(sp.level_hmax, sphb_level_of_cell(sp.lev, sp.nlev, (uint32_t)cell), hm);
    }
}

__global__ void __launch_bounds__(FIN_WARPS * 32)
    k_set_h(const double *__restrict__ h_sorted, SphbSpec sp, long long ncell, float delta,
            const uint32_t *__restrict__ cell_start, float *__restrict__ cell_hmax, double *__restrict__ posh,
            float *__restrict__ frec)
{
    const int lane = threadIdx.x & 31;
    const long long cell = (long long)blockIdx.x * FIN_WARPS + (threadIdx.x >> 5);
    if (cell >= ncell) return;
    const uint32_t s = cell_start[cell], e = cell_start[cell + 1];
    float hm = 0.f;
    for (uint32_t i = s + lane; i < e; i += 32) {
        const double h = h_sorted[i];
        posh[4 * (size_t)i + 3] = h;
        frec[4 * (size_t)i + 3] = sphb_thr_of_h(h, delta);
        if (h > 0.0) hm = fmaxf(hm, sphb_f_up(h));
        else hm = CUDART_INF_F;
    }
    hm = sphb_warp_max(hm);
    if (lane == 0) {
        cell_hmax[cell] = hm;
        if (sp.nlev > 1 && e > s) sphb_raise_level_hmax(sp.level_hmax, sphb_level_of_cell(sp.lev, sp.nlev, (uint32_t)cell), hm);
    }
}

template <typename T>
__global__ void k_gather(const T *__restrict__ src, const uint32_t *__restrict__ perm, long long n, int ncomp,
                         T *__restrict__ dst)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * ncomp) return;
    const long long s = t / ncomp;
    const int k = (int)(t - s * ncomp);
    dst[t] = src[(size_t)perm[s] * ncomp + k];
}

template <typename T>
__global__ void k_scatter(const T *__restrict__ src, const uint32_t *__restrict__ perm, long long n, int ncomp,
                          T *__restrict__ dst)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * ncomp) return;
    const long long s = t / ncomp;
    const int k = (int)(t - s * ncomp);
    dst[(size_t)perm[s] * ncomp + k] = src[t];
}

static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

extern "C" int sphb_pack_posh(const double *pos, const double *h, int64_t n, double *posh, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || (n > 0 && (!pos || !h || !posh))) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_pack_posh<<<nblk(n, 256), 256, 0, sphb_stream(stream)>>>(pos, h, n, posh);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

// the build pipeline for one grid or for nested grids (`sp`): counting sort by global cell id
static int sphb_build_impl(const double *posh_in, const int64_t *tie_key, int64_t n, const SphbSpec &sp, int64_t ncell,
                           float delta, uint32_t *perm, uint32_t *cell_start, float *cell_hmax, uint32_t *cell_id,
                           double *posh_sorted, float *frec_sorted, void *workspace, int64_t workspace_bytes, cudaStream_t st)
{
    if (ncell <= 0 || n < 0 || !cell_start || !cell_hmax || !workspace) return SPHB_E_ARG;
    if (n > 0 && (!posh_in || !perm || !cell_id || !posh_sorted || !frec_sorted)) return SPHB_E_ARG;
    if (n >= (int64_t)0xfffffff0ll || ncell >= (int64_t)0xfffffff0ll) return SPHB_E_ARG;
    if (workspace_bytes < sphb_levels_workspace_bytes(n, ncell)) return SPHB_E_ARG;
    const int64_t nb = scan_blocks(ncell);
    char *w = (char *)workspace;
    uint32_t *counts = (uint32_t *)w;
    w += align256(4 * ncell);
    uint32_t *cursor = (uint32_t *)w;
    w += align256(4 * ncell);
    uint32_t *block_sums = (uint32_t *)w;
    w += align256(4 * (nb + 1));
    uint32_t *perm_tmp = (uint32_t *)w;
    w += align256(4 * n);
    uint32_t *cell_tmp = (uint32_t *)w;

    SPHB_CUDA(cudaMemsetAsync(counts, 0, (size_t)align256(4 * ncell) * 2, st));   // counts + cursor
    if (sp.nlev > 1) SPHB_CUDA(cudaMemsetAsync(sp.level_hmax, 0, sizeof(float) * sp.nlev, st));
    if (n > 0) {
        k_cell_count<<<nblk(n, 256), 256, 0, st>>>(posh_in, n, sp, cell_tmp, counts);
        SPHB_LAUNCHED();
    }
    k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(counts, ncell, block_sums);
    SPHB_LAUNCHED();
    k_scan_blocks<<<1, SCAN_THREADS, 0, st>>>(block_sums, nb);
    SPHB_LAUNCHED();
    k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(counts, ncell, block_sums, cell_start);
    SPHB_LAUNCHED();
    if (n > 0) {
        k_cell_scatter<<<nblk(n, 256), 256, 0, st>>>(cell_tmp, n, cell_start, cursor, perm_tmp);
        SPHB_LAUNCHED();
    }
    k_cell_finish<<<nblk(ncell, FIN_WARPS), FIN_WARPS * 32, 0, st>>>(posh_in, (const long long *)tie_key, sp, ncell, delta,
                                                                     cell_start, perm_tmp, perm, cell_id, cell_hmax,
                                                                     posh_sorted, frec_sorted);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int64_t sphb_levels_workspace_bytes(int64_t n, int64_t ncell)
{
    if (ncell < 0 || n < 0) return -1;
    return align256(4 * ncell) * 2 + align256(4 * (scan_blocks(ncell) + 1)) + align256(4 * n) * 2;
}

extern "C" int sphb_grid_build(const double *posh_in, const int64_t *tie_key, int64_t n, const sphb_grid *g, float delta,
                               uint32_t *perm,
                               uint32_t *cell_start, float *cell_hmax, uint32_t *cell_id, double *posh_sorted,
                               float *frec_sorted, void *workspace, int64_t workspace_bytes, void *stream)
{
    SPHB_RANGE();
    const int64_t ncell = sphb_grid_ncell(g);
    if (ncell <= 0 || !(g->cell > 0.0)) return SPHB_E_ARG;
    return sphb_build_impl(posh_in, tie_key, n, sphb_spec_single(*g), ncell, delta, perm, cell_start, cell_hmax, cell_id,
                           posh_sorted, frec_sorted, workspace, workspace_bytes, sphb_stream(stream));
}

static inline bool sphb_spec_of_cells(const sphb_cells *c, SphbSpec &sp)
{
    if (!c || c->nlevels < 2 || c->nlevels > SPHB_MAX_LEVELS || !c->levels || !c->level_hmax) return false;
    sp.g0 = c->grid;
    sp.nlev = c->nlevels;
    sp.exp0 = c->exp0;
    sp.lev = c->levels;
    sp.level_hmax = const_cast<float *>(c->level_hmax);
    return true;
}

extern "C" int sphb_grid_build_levels(const double *posh_in, const int64_t *tie_key, int64_t n, const sphb_cells *cells,
                                      int64_t ncell, uint32_t *perm, void *workspace, int64_t workspace_bytes, void *stream)
{
    SPHB_RANGE();
    SphbSpec sp;
    if (!sphb_spec_of_cells(cells, sp) || cells->n != n) return SPHB_E_ARG;
    return sphb_build_impl(posh_in, tie_key, n, sp, ncell, cells->delta, perm, const_cast<uint32_t *>(cells->cell_start),
                           const_cast<float *>(cells->cell_hmax), const_cast<uint32_t *>(cells->cell_id),
                           const_cast<double *>(cells->posh), const_cast<float *>(cells->frec), workspace, workspace_bytes,
                           sphb_stream(stream));
}

// per-octave statistics (sphb_level_stats): lanes of a warp mostly share an octave, so each octave present in the warp
// is reduced by shuffles and lands in shared memory with one atomic per quantity; one set of global atomics per block
#define LS_BINS 16
#define LS_THREADS 256
#define LS_BLOCKS 1184
__device__ __forceinline__ int sphb_f2ord_down(double v) { return sphb_f2ord(__double2float_rd(v)); }
__device__ __forceinline__ int sphb_f2ord_up(double v) { return sphb_f2ord(__double2float_ru(v)); }

__global__ void __launch_bounds__(LS_THREADS)
    k_level_stats(const double *__restrict__ posh, long long n, int exp_top, double *__restrict__ out)
{
    __shared__ unsigned int s_cnt[LS_BINS];
    __shared__ double s_sum[LS_BINS];
    __shared__ int s_lo[LS_BINS][3], s_hi[LS_BINS][3];
    for (int q = threadIdx.x; q < LS_BINS; q += blockDim.x) {
        s_cnt[q] = 0;
        s_sum[q] = 0.0;
        for (int c = 0; c < 3; ++c) {
            s_lo[q][c] = 0x7fffffff;
            s_hi[q][c] = (int)0x80000000;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
        const long long i = i0 + threadIdx.x;
        int bin = -1;
        double x = 0.0, y = 0.0, z = 0.0, h = 0.0;
        if (i < n) {
            sphb_ld4(posh, (size_t)i, x, y, z, h);
            bin = sphb_level_of(h, exp_top - (LS_BINS - 1), LS_BINS);
        }
        unsigned rem = __ballot_sync(SPHB_FULL, bin >= 0);
        while (rem) {
            const int b = __shfl_sync(SPHB_FULL, bin, __ffs(rem) - 1);
            const bool in = bin == b;
            const unsigned m = __ballot_sync(SPHB_FULL, in);
            rem &= ~m;
            const double sh = sphb_warp_sum(in && h > 0.0 && h < CUDART_INF ? h : 0.0);
            const int lx = __reduce_min_sync(SPHB_FULL, in ? sphb_f2ord_down(x) : 0x7fffffff);
            const int ly = __reduce_min_sync(SPHB_FULL, in ? sphb_f2ord_down(y) : 0x7fffffff);
            const int lz = __reduce_min_sync(SPHB_FULL, in ? sphb_f2ord_down(z) : 0x7fffffff);
            const int hx = __reduce_max_sync(SPHB_FULL, in ? sphb_f2ord_up(x) : (int)0x80000000);
            const int hy = __reduce_max_sync(SPHB_FULL, in ? sphb_f2ord_up(y) : (int)0x80000000);
            const int hz = __reduce_max_sync(SPHB_FULL, in ? sphb_f2ord_up(z) : (int)0x80000000);
            if (lane == 0) {
                atomicAdd(&s_cnt[b], (unsigned)__popc(m));
                atomicAdd(&s_sum[b], sh);
                atomicMin(&s_lo[b][0], lx);
                atomicMin(&s_lo[b][1], ly);
                atomicMin(&s_lo[b][2], lz);
                atomicMax(&s_hi[b][0], hx);
                atomicMax(&s_hi[b][1], hy);
                atomicMax(&s_hi[b][2], hz);
            }
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < LS_BINS; q += blockDim.x) {
        if (s_cnt[q] == 0) continue;
        atomicAdd(reinterpret_cast<unsigned long long *>(out) + 8 * q, (unsigned long long)s_cnt[q]);
        atomicAdd(out + 8 * q + 1, s_sum[q]);
        for (int c = 0; c < 3; ++c) {
            atomicMin(reinterpret_cast<int *>(out + 8 * q + 2 + c), s_lo[q][c]);
            atomicMax(reinterpret_cast<int *>(out + 8 * q + 5 + c), s_hi[q][c]);
        }
    }
}

__global__ void k_level_stats_out(double *__restrict__ out)
{
    const int q = threadIdx.x;
    if (q >= LS_BINS) return;
    const unsigned long long cnt = reinterpret_cast<unsigned long long *>(out)[8 * q];
    out[8 * q] = (double)cnt;
    for (int c = 2; c < 8; ++c) {
        const int key = reinterpret_cast<int *>(out + 8 * q + c)[0];
        out[8 * q + c] = cnt ? (double)sphb_ord2f(key) : 0.0;
    }
}

// the 16 x 8 output doubles double as the reduction scratch: row b = {count (u64), sum h, then the six bounds as
// ordered-int keys in the low words}; k_level_stats_out converts the row in place
__global__ void k_level_stats_init(double *out)
{
    const int q = threadIdx.x;
    if (q >= LS_BINS) return;
    reinterpret_cast<unsigned long long *>(out)[8 * q] = 0ull;
    out[8 * q + 1] = 0.0;
    for (int c = 0; c < 3; ++c) {
        reinterpret_cast<int *>(out + 8 * q + 2 + c)[0] = 0x7fffffff;
        reinterpret_cast<int *>(out + 8 * q + 5 + c)[0] = (int)0x80000000;
    }
}

extern "C" int sphb_level_stats(const double *posh, int64_t n, int32_t exp_top, double *out, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || !out || (n > 0 && !posh)) return SPHB_E_ARG;
    cudaStream_t st = sphb_stream(stream);
    k_level_stats_init<<<1, 32, 0, st>>>(out);
    SPHB_LAUNCHED();
    if (n > 0) {
        long long nb = (n + LS_THREADS - 1) / LS_THREADS;
        if (nb > LS_BLOCKS) nb = LS_BLOCKS;
        k_level_stats<<<(unsigned)nb, LS_THREADS, 0, st>>>(posh, (long long)n, exp_top, out);
        SPHB_LAUNCHED();
    }
    k_level_stats_out<<<1, 32, 0, st>>>(out);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_grid_set_h(const double *h_sorted, int64_t n, const sphb_grid *g, float delta,
                               const uint32_t *cell_start, float *cell_hmax, double *posh_sorted, float *frec_sorted,
                               void *stream)
{
    SPHB_RANGE();
    const int64_t ncell = sphb_grid_ncell(g);
    if (ncell <= 0 || n < 0 || !cell_start || !cell_hmax) return SPHB_E_ARG;
    if (n > 0 && (!h_sorted || !posh_sorted || !frec_sorted)) return SPHB_E_ARG;
    k_set_h<<<nblk(ncell, FIN_WARPS), FIN_WARPS * 32, 0, sphb_stream(stream)>>>(h_sorted, sphb_spec_single(*g), ncell, delta,
                                                                                cell_start, cell_hmax, posh_sorted, frec_sorted);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_grid_set_h_levels(const double *h_sorted, const sphb_cells *cells, int64_t ncell, void *stream)
{
    SPHB_RANGE();
    SphbSpec sp;
    if (!sphb_spec_of_cells(cells, sp) || ncell <= 0 || !cells->cell_start || !cells->cell_hmax) return SPHB_E_ARG;
    if (cells->n > 0 && (!h_sorted || !cells->posh || !cells->frec)) return SPHB_E_ARG;
    cudaStream_t st = sphb_stream(stream);
    SPHB_CUDA(cudaMemsetAsync(sp.level_hmax, 0, sizeof(float) * sp.nlev, st));
    k_set_h<<<nblk(ncell, FIN_WARPS), FIN_WARPS * 32, 0, st>>>(h_sorted, sp, ncell, cells->delta, cells->cell_start,
                                                               const_cast<float *>(cells->cell_hmax),
                                                               const_cast<double *>(cells->posh),
                                                               const_cast<float *>(cells->frec));
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_gather_f64(const double *src, const uint32_t *perm, int64_t n, int32_t ncomp, double *dst,
                               void *stream)
{
    SPHB_RANGE();
    if (n < 0 || ncomp <= 0 || (n > 0 && (!src || !perm || !dst))) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_gather<double><<<nblk(n * ncomp, 256), 256, 0, sphb_stream(stream)>>>(src, perm, n, ncomp, dst);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_scatter_f64(const double *src, const uint32_t *perm, int64_t n, int32_t ncomp, double *dst,
                                void *stream)
{
    SPHB_RANGE();
    if (n < 0 || ncomp <= 0 || (n > 0 && (!src || !perm || !dst))) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_scatter<double><<<nblk(n * ncomp, 256), 256, 0, sphb_stream(stream)>>>(src, perm, n, ncomp, dst);
    SPHB_LAUNCHED();
    return SPHB_OK;
}

extern "C" int sphb_gather_u32(const uint32_t *src, const uint32_t *perm, int64_t n, uint32_t *dst, void *stream)
{
    SPHB_RANGE();
    if (n < 0 || (n > 0 && (!src || !perm || !dst))) return SPHB_E_ARG;
    if (n == 0) return SPHB_OK;
    k_gather<uint32_t><<<nblk(n, 256), 256, 0, sphb_stream(stream)>>>(src, perm, n, 1, dst);
    SPHB_LAUNCHED();
    return SPHB_OK;
}
