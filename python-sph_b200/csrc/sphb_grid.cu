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
extern "C" int64_t sphb_grid_workspace_bytes(int64_t n, const sphb_grid *g)
{
    const int64_t ncell = sphb_grid_ncell(g);
    if (ncell < 0 || n < 0) return -1;
    return align256(4 * ncell) * 2 + align256(4 * (scan_blocks(ncell) + 1)) + align256(4 * n) * 2;
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

__global__ void k_cell_count(const double *__restrict__ posh, long long n, sphb_grid g, uint32_t *__restrict__ cell_tmp,
                             uint32_t *__restrict__ counts)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 a = reinterpret_cast<const double2 *>(posh)[2 * i];
    const double2 b = reinterpret_cast<const double2 *>(posh)[2 * i + 1];
    const uint32_t c = sphb_cell_of(g, a.x, a.y, b.x);
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
    k_cell_finish(const double *__restrict__ posh_in, const long long *__restrict__ tie_key, sphb_grid g, long long ncell,
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
    if (lane == 0) cell_hmax[cell] = hm;
}

__global__ void __launch_bounds__(FIN_WARPS * 32)
    k_set_h(const double *__restrict__ h_sorted, sphb_grid g, long long ncell, float delta,
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
    if (lane == 0) cell_hmax[cell] = hm;
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

extern "C" int sphb_grid_build(const double *posh_in, const int64_t *tie_key, int64_t n, const sphb_grid *g, float delta,
                               uint32_t *perm,
                               uint32_t *cell_start, float *cell_hmax, uint32_t *cell_id, double *posh_sorted,
                               float *frec_sorted, void *workspace, int64_t workspace_bytes, void *stream)
{
    SPHB_RANGE();
    const int64_t ncell = sphb_grid_ncell(g);
    if (ncell <= 0 || n < 0 || !(g->cell > 0.0) || !cell_start || !cell_hmax || !workspace) return SPHB_E_ARG;
    if (n > 0 && (!posh_in || !perm || !cell_id || !posh_sorted || !frec_sorted)) return SPHB_E_ARG;
    if (n >= (int64_t)0xfffffff0ll || ncell >= (int64_t)0xfffffff0ll) return SPHB_E_ARG;
    if (workspace_bytes < sphb_grid_workspace_bytes(n, g)) return SPHB_E_ARG;
    cudaStream_t st = sphb_stream(stream);
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
    if (n > 0) {
        k_cell_count<<<nblk(n, 256), 256, 0, st>>>(posh_in, n, *g, cell_tmp, counts);
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
    k_cell_finish<<<nblk(ncell, FIN_WARPS), FIN_WARPS * 32, 0, st>>>(posh_in, (const long long *)tie_key, *g, ncell, delta,
                                                                     cell_start, perm_tmp,
                                                                     perm, cell_id, cell_hmax, posh_sorted, frec_sorted);
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
    k_set_h<<<nblk(ncell, FIN_WARPS), FIN_WARPS * 32, 0, sphb_stream(stream)>>>(h_sorted, *g, ncell, delta, cell_start,
                                                                                cell_hmax, posh_sorted, frec_sorted);
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
