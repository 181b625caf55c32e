/*
 * sphb200.h -- C-ABI of libsphb200.so: the B200 (sm_100a) implementation of python-sph's
 * per-step particle loop (TRU-astrophysics/python-sph, pyfluid.py).
 *
 * The reference is pure Python; its "FFI" for this path is the set of module-level functions of
 * pyfluid.py.  Each entry point below names the reference function(s) it replaces (file:line).  The
 * Python drop-in module (python-sph_b200/pyfluid.py) binds these with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - all sizes are int64_t; all entry points are stream-ordered on `stream` (a cudaStream_t passed
 *     as void*), never synchronise, and return 0 on success or a negative SPHB_E_* code;
 *   - workspace is caller-owned (sphb_*_workspace_bytes);
 *   - particles are handled in "grid order": sphb_grid_build() sorts them by cell and every later
 *     call takes and returns arrays in that order; `perm[s]` is the caller index of sorted slot s;
 *   - soft failures and the reference's RuntimeError conditions are reported through a device
 *     status block (sphb_status), read by the host once per stage.
 */
#ifndef SPHB200_H
#define SPHB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB_ABI_VERSION 1

/* error codes */
#define SPHB_OK 0
#define SPHB_E_ARG (-1)      /* bad argument (null pointer, n < 0, grid too large ...) */
#define SPHB_E_CUDA (-2)     /* a CUDA runtime call failed; sphb_last_cuda_error() has the text */
#define SPHB_E_NODEVICE (-3) /* no sm_100 device / kernel image not loadable */

/* status bits (pyfluid.py RuntimeError sites); identical to oracle/sph_oracle.c ORC_* */
#define SPHB_ST_NEG_H 1        /* "Obtained negative smoothing length."             pyfluid.py:209-210 */
#define SPHB_ST_NEG_P 2        /* "Obtained negative pressure."                     pyfluid.py:257-258 */
#define SPHB_ST_NEG_PI 4       /* "Obtained negative value for Pi"                  pyfluid.py:281-282 */
#define SPHB_ST_NEG_VISC 8     /* "Obtained negative energy rate due to viscosity." pyfluid.py:349-350 */
#define SPHB_ST_NEG_E 16       /* "Obtained negative energy."                       pyfluid.py:514-515,527-528 */
#define SPHB_ST_BISECT_NONE 32 /* bisection_h fell through and returned None        pyfluid.py:153 */

/* pyfluid.py:5-41 -- module constants, passed per call (the reference reads them at call time) */
typedef struct sphb_consts {
    double particle_mass;  /* PARTICLE_MASS                     :5  */
    double coupling_const; /* COUPLING_CONST (eta)              :9  */
    double h_tol;          /* SMOOTHLENGTH_VARIATION_TOLERANCE  :13 */
    int32_t newton_limit;  /* NEWTON_ITERATION_LIMIT            :14 */
    int32_t bisect_limit;  /* BISECTION_ITERATION_LIMIT         :23 */
    double alpha;          /* ALPHA_SPH                         :18 */
    double beta;           /* BETA_SPH                          :19 */
    double epsilon;        /* EPSILON                           :20 */
    double bisect_tol;     /* BISECTION_TOLERANCE               :24 */
    double bisect_a;       /* INITIAL_BISECTION_SMOOTHLENGTH_A  :25 */
    double bisect_b;       /* INITIAL_BISECTION_SMOOTHLENGTH_B  :26 */
    double G;              /* G (0: hydro only; else sphb_force needs bgrav and sphb_gravity_direct adds the pair sum) :35 */
    double gamma;          /* ADIABATIC_INDEX                   :41 */
    const double *mass;    /* optional DEVICE array of PER-PARTICLE masses (the reference authors' own
                              NOTE at pyfluid.py:3-4), indexed like the particle arrays of the call (sorted slots for the
                              gather kernels); NULL: every particle has particle_mass, which is the reference.  With it
                              every "PARTICLE_MASS" inside a sum over i becomes m_i (:95,216,327,340,474) and the one of
                              var_density / zetaprime becomes the target's m_j (:115,238) */
} sphb_consts;

/* Cell grid: edge `cell` in y and z, edge cell/xdiv in x (x is the fastest axis: finer x cells make the row
 * spans of a search tighter).  Cell of a point: cx = clamp(floor((x - ox)/(cell/xdiv)), 0, nx-1),
 * cy = clamp(floor((y - oy)/cell), 0, ny-1), cz likewise; linear id = (cz*ny + cy)*nx + cx.
 * (The reference has no cell list; SURVEY.md 8(c) "self-oracle".) */
typedef struct sphb_grid {
    double ox, oy, oz; /* origin */
    double cell;       /* cell edge in y and z */
    int32_t nx, ny, nz;
    int32_t xdiv;      /* x edge = cell / xdiv; values < 1 mean 1 */
} sphb_grid;

/* device status block, zeroed by sphb_status_reset */
typedef struct sphb_status {
    int32_t flags;          /* OR of SPHB_ST_* */
    int32_t newton_warn;    /* # "WARNING: Failure to converge in newton_new_h."   pyfluid.py:178 */
    int32_t bisect_oob;     /* # "ERROR: Root out of bisection bounds."            pyfluid.py:136 */
    int32_t bisect_none;    /* # "ERROR: Failure to converge in Bisection_new_h."  pyfluid.py:153 */
    int32_t nfail;          /* entries in the Newton-failure list (internal)        */
    int32_t reserved[3];
} sphb_status;

/* Nested cell grids for distributions whose h spans decades (BASELINE config 5, the Plummer sphere): one uniform grid
 * would put 10^3..10^5 particles of the dense core into a cell sized for the mean h.  Particles are binned by the
 * OCTAVE of their h at build time -- level l holds 2^(exp0+l) <= h < 2^(exp0+l+1), clamped to [0, nlevels-1]; h <= 0 or
 * non-finite (unbounded support, pyfluid.py:57-60) goes to the last level -- and every level has its own grid: its own
 * cell edge (~ one support radius of ITS particles), its own bounding box.  Cells are numbered level-major (global cell id
 * = cell_base + local id), so the sorted order, cell_start, cell_hmax and cell_id stay single arrays and a search just
 * walks the levels one after the other: candidates are still visited in ascending sorted slot.  A level's grid must
 * contain all of its particles (no clamped outliers): searches skip a level whose box they do not touch. */
#define SPHB_MAX_LEVELS 16
typedef struct sphb_level {
    sphb_grid grid;     /* absolute origin; xdiv is ignored (1) */
    uint32_t cell_base; /* first global cell id of this level */
    uint32_t ncell;     /* nx*ny*nz */
} sphb_level;

/* sorted-order particle views built by sphb_grid_build and consumed by every gather kernel */
typedef struct sphb_cells {
    sphb_grid grid;             /* the one grid, or (nlevels > 1) only the origin of the fp32 frame of `frec` */
    int64_t n;                  /* particles in the grid (targets + ghosts) */
    const uint32_t *cell_start; /* [ncell+1] first sorted slot of each cell */
    const float *cell_hmax;     /* [ncell]   max h in the cell, rounded up to float */
    const uint32_t *cell_id;    /* [n]       linear cell id of each sorted slot */
    const double *posh;         /* [n][4]    x, y, z, h   (fp64, exact) */
    const float *frec;          /* [n][4]    x-ox, y-oy, z-oz, (2h+delta)^2 upper bound (fp32 prefilter record) */
    float delta;                /* absolute error bound of the fp32 positions */
    float pad_;
    int32_t nlevels;            /* 0 or 1: one grid (`grid`); 2..SPHB_MAX_LEVELS: nested grids */
    int32_t exp0;               /* octave of level 0 */
    const sphb_level *levels;   /* DEVICE [nlevels] */
    const float *level_hmax;    /* DEVICE [nlevels] max h per level, rounded up (kept current by build / set_h) */
} sphb_cells;

const char *sphb_last_cuda_error(void);
int sphb_abi_version(void);
/* number of CUDA devices with compute capability 10.x visible to the library (0 = none) */
int sphb_device_ok(void);

int sphb_status_reset(sphb_status *status, void *stream);

/* ---- K1: cell-list build (no reference counterpart; enables O(N n_ngb) sums) ------------------ */
int64_t sphb_grid_ncell(const sphb_grid *g);
int64_t sphb_grid_workspace_bytes(int64_t n, const sphb_grid *g);
/* pack caller-order AoS (pos (n,3), h (n)) into posh (n,4) */
int sphb_pack_posh(const double *pos, const double *h, int64_t n, double *posh, void *stream);
/* Sort `posh_in` (any order) into grid order.  Outputs: perm[s] = index into posh_in of sorted
 * slot s, cell_start, cell_hmax, cell_id, posh_sorted, frec_sorted.  Ties in cell id are ordered by
 * ascending tie_key[input index] (unique keys, e.g. global particle ids), or by ascending input index when
 * tie_key is NULL (stable sort).  delta = fp32 position error bound to bake into frec.w. */
int sphb_grid_build(const double *posh_in, const int64_t *tie_key, int64_t n, const sphb_grid *g, float delta, uint32_t *perm,
                    uint32_t *cell_start, float *cell_hmax, uint32_t *cell_id, double *posh_sorted,
                    float *frec_sorted, void *workspace, int64_t workspace_bytes, void *stream);
/* The same for nested grids (cells->nlevels > 1): cells->grid gives the fp32 frame, cells->levels / exp0 the level table
 * (device), cells->level_hmax is WRITTEN (device [nlevels]); the output arrays are the ones `cells` points to plus
 * perm.  ncell = the sum of the levels' cell counts. */
int64_t sphb_levels_workspace_bytes(int64_t n, int64_t ncell);
int sphb_grid_build_levels(const double *posh_in, const int64_t *tie_key, int64_t n, const sphb_cells *cells, int64_t ncell,
                           uint32_t *perm, void *workspace, int64_t workspace_bytes, void *stream);
/* per-octave statistics of a particle set, to choose the levels from: for the 16 octaves [2^(exp_top-15+b), 2^(exp_top-14+b))
 * (b = 0: everything below as well; b = 15: everything above, and h <= 0 / non-finite), out[b] = {count, sum h,
 * min x, y, z, max x, y, z} (bounds rounded outward to float; empty octaves: count 0).  out: DEVICE [16][8] doubles. */
int sphb_level_stats(const double *posh, int64_t n, int32_t exp_top, double *out, void *stream);
/* refresh h in posh/frec/cell_hmax of an existing grid (positions unchanged) */
int sphb_grid_set_h(const double *h_sorted, int64_t n, const sphb_grid *g, float delta, const uint32_t *cell_start,
                    float *cell_hmax, double *posh_sorted, float *frec_sorted, void *stream);
/* nested grids: also refreshes cells->level_hmax */
int sphb_grid_set_h_levels(const double *h_sorted, const sphb_cells *cells, int64_t ncell, void *stream);
/* dst[s*ncomp + k] = src[perm[s]*ncomp + k]  (apply a build's permutation to any fp64 array) */
int sphb_gather_f64(const double *src, const uint32_t *perm, int64_t n, int32_t ncomp, double *dst, void *stream);
/* dst[perm[s]*ncomp + k] = src[s*ncomp + k]  (undo it) */
int sphb_scatter_f64(const double *src, const uint32_t *perm, int64_t n, int32_t ncomp, double *dst, void *stream);
int sphb_gather_u32(const uint32_t *src, const uint32_t *perm, int64_t n, uint32_t *dst, void *stream);

/* ---- neighbour lists kept between kernels (optional) -----------------------------------------------
 * Every gather kernel first scans the cells around its 32 targets; within one step several kernels scan the SAME
 * snapshot (pyfluid.py:502-507: omega_arr then energy_rate_arr/acceleration_arr at (pos0, h0); :516-522: the Newton
 * solve then the same two at pos_mid).  With a sphb_lists the first of them (sphb_density_omega or sphb_newton_h, the
 * PRODUCER) scans with the symmetric support max(h_j, h_i) and parks every warp's lists in `store`; sphb_force (the
 * CONSUMER) drains them instead of scanning again.  Sums are unchanged: the lists are supersets in the same slot order
 * and membership is still decided pair by pair in fp64.
 *   store        sphb_lists_bytes(n) bytes of device memory
 *   broken       device flag, zeroed by the caller before the producer runs; the Newton producer sets it when some
 *                h ends above cover * (h the cell list holds), i.e. when the parked lists may be incomplete for the new
 *                h -- the consumer then scans for itself.  A warp whose lists overflowed also scans for itself.
 *   hmax_global  device float: max h of the snapshot (sphb_hmax), bounds the symmetric search of the producer
 *   cover        >= 1: how much every h may grow between producer and consumer (1 when h does not change)
 * The producer and the consumer must be called with the same `target` flags, and a Newton producer with h_guess = the
 * h the cell list holds (the cover is relative to it).  NULL = every kernel scans. */
typedef struct sphb_lists {
    void *store;
    int32_t *broken;
    const float *hmax_global;
    double cover;
} sphb_lists;
size_t sphb_lists_bytes(int64_t n);

/* ---- neighbour lists kept across kernels AND steps (no reference counterpart: the reference sums over all N) ----------
 * sphb_nlist_build scans the cell list ONCE -- every pair with |r_j - r_i| <= 2 max(c_j h_j, c_i h_i) + s_j h_j + s_i h_i,
 * c = cover_h and s = skin (or the per-particle `margins`): the symmetric support plus what both records may move --
 * and stores each warp's candidates (32 consecutive sorted slots = one warp) as rows of slot indices in `pool`.  The
 * sphb_nlist_* gather kernels only drain those rows: they read the CURRENT sorted records (positions and h rewritten in
 * place, slot order frozen) and decide membership pair by pair in fp64, exactly like the scanning kernels; entries
 * outside the current support add exact zeros and rows are in ascending slot order, so the sums are the same bits.  The
 * lists stay complete while, for every record, h_now <= cover_h * h_build and |x_now - x_build| <= skin * h_build;
 * sphb_nlist_check and sphb_nlist_newton_h set SPHB_NL_BROKEN in *flags otherwise (SPHB_NL_AGING at half of a bound),
 * and the caller rebuilds (and redoes the step that saw BROKEN). */
#define SPHB_NL_NONE 0xffffffffu /* padding entry of a row */
#define SPHB_NL_POOL 1       /* pool too small for this build: enlarge (sphb_nlist_pool_rows is always enough) and rebuild */
#define SPHB_NL_BROKEN 2     /* a record left what the lists cover, or a Newton solve needs the bisection fallback */
#define SPHB_NL_AGING 4      /* half of a bound is used up: rebuild before the next step */
#define SPHB_NL_UNLISTABLE 8 /* some target has more candidates than a warp's reserve holds (unbounded h): scanning path only */
typedef struct sphb_nlist {
    int64_t n;                 /* sorted slots covered (targets + ghosts) */
    uint32_t *warp_off;        /* [ceil(n/32)] first pool row of each warp's lists */
    uint32_t *warp_rows;       /* [ceil(n/32)] rows in use: row k holds entry k of each of the 32 lanes */
    uint32_t *pool;            /* [pool_rows][32] sorted-slot indices, SPHB_NL_NONE where a lane has fewer entries */
    int64_t pool_rows;
    unsigned long long *cursor; /* device: rows handed out by the build */
    int32_t *flags;            /* device: OR of SPHB_NL_* */
    const double *posh_build;  /* [n][4] the sorted records the lists were built from (must stay alive and unchanged) */
    double cover_h;            /* >= 1 */
    double skin;               /* >= 0, in units of each record's h at build time */
    int32_t *maxima;           /* optional device [2], zeroed by the build: float bit patterns of the largest FRACTIONS of
                                  their allowances any record has used up in any check since: |x_now - x_build| / (s_i h_build)
                                  and (h_now / h_build - 1) / (c_i - 1) */
    const float *margins;      /* optional device [n][2] = {c_i, s_i}: PER-PARTICLE cover (>= 1) and skin (>= 0) in place of
                                  cover_h / skin (which must then be their maxima).  A pair is listed iff
                                  |r_j - r_i| <= 2 max(c_j h_j, c_i h_i) + s_j h_j + s_i h_i: a fast particle pays for its own
                                  allowance, its partners keep tight lists; all records are listed for the whole life of the
                                  lists as long as every one of them stays within its own bounds */
    const float *level_dmax;   /* with margins: device [max(nlevels, 1)] upper bounds of s_i h_i over the records of each
                                  level of the cell list (one entry for a single grid): bounds the cells searched */
    const float *cell_dmax;    /* optional device [ncell] upper bounds of s_i h_i over the records of each cell (0 for an empty
                                  cell): cells beyond the targets' own reach are then searched only where their records'
                                  allowances bridge the gap -- a fast record does not widen the search of a quiet region */
} sphb_nlist;
int64_t sphb_nlist_pool_rows(int64_t n);
int64_t sphb_sizeof_nlist(void); /* sizeof(sphb_nlist), for bindings to check their own layout against */
/* cells->posh must be nl->posh_build.  Zeroes *cursor and *flags first.
 * omega (optional, needs k): Omega_j = 1 + h_j/(3 rho_j) sum_i m dW/dh with rho_j = m (eta/h_j)^3 at the build-time records
 * (what sphb_nlist_density_omega would return for them), formed while the lists are still in the staging area. */
int sphb_nlist_build(const sphb_cells *cells, const uint8_t *target, const float *hmax_global, const sphb_nlist *nl,
                     const sphb_consts *k, double *omega, void *stream);
int sphb_nlist_check(const sphb_nlist *nl, const double *posh_now, void *stream);
/* Decomposition by ranges of sorted slots (every rank sorts the same global cell list and lists only its own range of
 * targets): mark[s] = 1 for every slot the lists name -- the rank's exact ghost set --, and rename the slots in place,
 * entry v -> map[v] (map increasing on the named slots: rows stay in ascending order). */
int sphb_nlist_mark(const sphb_nlist *nl, uint8_t *mark, void *stream);
int sphb_nlist_remap(const sphb_nlist *nl, const int32_t *map, void *stream);
/* sphb_density_omega on kept lists (h_j = posh[j][3]) */
int sphb_nlist_density_omega(const sphb_nlist *nl, const sphb_consts *k, const uint8_t *target, const double *posh,
                             const double *rho_in, double *rho_sum, double *dwdh_sum, double *omega, int32_t *ngb_count,
                             void *stream);
/* sphb_newton_h on kept lists, guess = h_guess[j] (NULL: posh[j][3]).  posh_out (optional, != posh): the records with the solved h.
 * Non-convergence within the limit is NOT bisected here: h_out = NaN and SPHB_NL_BROKEN (redo on the scanning path). */
int sphb_nlist_newton_h(const sphb_nlist *nl, const sphb_consts *k, const uint8_t *target, const double *posh,
                        const double *h_guess, double *h_out, double *omega_out, double *posh_out, int32_t *iters, sphb_status *status,
                        void *stream);
/* sphb_force on kept lists (G must be 0); posh carries the current h.  Records: recB + recC from sphb_eos (recD NULL),
 * or recB + recD from sphb_nlist_eos (recC NULL): recD is a dense [n][2] array of {rho, cs} -- what the viscosity
 * branch gathers per neighbour -- so that a gathered entry shares its cache line with seven others; same bits. */
int sphb_nlist_eos(const double *posh, const double *vel, const double *u, const double *omega, int64_t n,
                   const sphb_consts *k, double *recB, double *recD, sphb_status *status, void *stream);
int sphb_nlist_force(const sphb_nlist *nl, const sphb_consts *k, const uint8_t *target, const double *posh,
                     const double *recB, const double *recC, const double *recD, double *acc, double *dudt,
                     sphb_status *status, void *stream);
/* ghost refresh plumbing: rows idx[0..n_idx) of up to four fp64 arrays (row widths w0..w3; unused: NULL, 0) gathered into
 * the packed buffer buf[n_idx][w0+w1+w2+w3] (unpack = 0), or scattered back from it (unpack != 0) */
int sphb_rows_pack(double *s0, int32_t w0, double *s1, int32_t w1, double *s2, int32_t w2, double *s3, int32_t w3,
                   const int64_t *idx, int64_t n_idx, double *buf, int32_t unpack, void *stream);
/* caller order <-> slot order for the whole state in one pass (any pointer group may be NULL) */
int sphb_gather_state(const double *pos, const double *h, const double *vel, const double *u, const uint32_t *perm,
                      int64_t n, double *posh_s, double *vel_s, double *u_s, void *stream);
int sphb_scatter_state(const double *posh_s, const double *vel_s, const double *u_s, const uint32_t *perm, int64_t n,
                       double *pos, double *h, double *vel, double *u, void *stream);

/* ---- K2: M4 density summation + grad-h sums ----------------------------------------------------
 * replaces density/density_arr (pyfluid.py:218-232) and omega_j/omega_arr (pyfluid.py:91-105).
 * For every slot j whose `target` flag is set (NULL = all slots):
 *   rho_sum[j]  = sum_i m W(|r_j - r_i|, h_j)                      (incl. i = j)
 *   dwdh_sum[j] = sum_i m dW/dh(|r_j - r_i|, h_j)
 *   omega[j]    = 1 + h_j/(3 rho_j) dwdh_sum[j]                    (if omega non-NULL; rho_j = rho_in[j], or
 *                                                                   var_density(h_j) = m (eta/h_j)^3 when rho_in is NULL,
 *                                                                   which is what the step uses, pyfluid.py:502-503)
 *   xi[j]       = -h_j/(3 rho_j) sum_i m dphi/dh(|r_j - r_i|, h_j)    (optional; xi/xi_arr, pyfluid.py:401-418 -- the
 *                                                                   grad-h term of the softened gravity, compact support)
 * h_j is cells->posh[j][3] unless h_target != NULL.  ngb_count (optional) = #{i : q < 2}. */
int sphb_density_omega(const sphb_cells *cells, const sphb_consts *k, const uint8_t *target, const double *h_target,
                       const double *rho_in, double *rho_sum, double *dwdh_sum, double *omega, double *xi,
                       int32_t *ngb_count, const sphb_lists *lists, void *stream);

/* ---- K3: Newton-Raphson smoothing-length solve ---------------------------------------------------
 * replaces newton_smoothlength_iteration/_while/_arr (pyfluid.py:155-179, 203-211) and
 * bisection_h (pyfluid.py:123-153).  h_out[j]: converged h, or the bisection result, or NaN when
 * bisection returned None.  iters[j]: Newton iterations performed.  code[j]: 0 Newton converged,
 * 1 bisection converged, 2 bisection returned None.  Counters/flags land in *status.
 * omega_out (optional): Omega_j = 1 + h/(3 rho) sum_i m dW/dh at the RETURNED h with rho = m (eta/h)^3 -- what the step
 * computes next (pyfluid.py:517-518); it comes from one more pass over the converged iteration's neighbour list. */
int sphb_newton_h(const sphb_cells *cells, const sphb_consts *k, const uint8_t *target, const double *h_guess,
                  double *h_out, double *omega_out, int32_t *iters, int32_t *code, uint32_t *fail_list,
                  sphb_status *status, const sphb_lists *lists, void *stream);
/* bisection for an explicit list of target slots (also used for bisection_h(j, pos) itself);
 * n_list < 0: take the list length from status->nfail (device side).  oob (optional, per list entry) */
int sphb_bisect_h(const sphb_cells *cells, const sphb_consts *k, const uint32_t *list, int64_t n_list,
                  double *h_out, double *omega_out, int32_t *code, int32_t *oob, int32_t indexed_by_slot,
                  sphb_status *status, void *stream);

/* ---- neighbour sets (SURVEY.md 8(c) oracle: NB(j) = {i : distance(r_j, r_i)/h_j < 2}) -------------
 * exact reference arithmetic (fma-ordered dot, IEEE sqrt and divide).  Writes, for each of the
 * n_query slots in `query`, up to `cap` neighbour slots (ascending slot order) and the true count.
 * symmetric != 0: the force support {i : q_j < 2 or q_i < 2}; needs hmax_global (sphb_hmax). */
int sphb_neighbours(const sphb_cells *cells, const uint32_t *query, int64_t n_query, int32_t symmetric,
                    const float *hmax_global, uint32_t *out, int64_t cap, int32_t *count, void *stream);

/* ---- K4: EOS prologue + momentum/energy sums ------------------------------------------------------
 * sphb_eos: replaces var_density_arr (:240), pressure_arr (:251) and builds the per-particle force
 * records: rho = m (eta/h)^3, P = (gamma-1) u rho, A = P/(omega rho^2), cs = sqrt(gamma P/rho).
 * recB[j] = {vx, vy, vz, A}; recC[j] = {rho, cs, 1/h, 1/(pi h^4)}.
 * rho_in / P_in (optional): take density / pressure from the caller instead (energy_rate_arr and
 * acceleration_arr receive them as arguments, pyfluid.py:355,487); u may then be NULL.
 * recB/recC may be NULL (var_density_arr / pressure_arr only).
 * bgrav (optional, needs omega): (G/2) xi/omega per particle -- the factor of the compact grad-h gravity correction
 * (smoothed_gravity_correction_comp, pyfluid.py:434-437); zeros when xi is NULL. */
int sphb_eos(const double *posh, const double *vel, const double *u, const double *omega, const double *rho_in,
             const double *P_in, const double *xi, int64_t n, const sphb_consts *k, double *rho, double *P, double *recB,
             double *recC, double *bgrav, sphb_status *status, void *stream);
/* sphb_force: replaces energy_rate/_arr (pyfluid.py:316-361), Pi (:261-285),
 * fluid_accel_viscosity_comp (:459-474) and the compact terms of acceleration/_arr (:477-493): with bgrav != NULL
 * the grad-h gravity correction (:434-437) is included; the long-range gravity sum is sphb_gravity_direct.
 * acc (n,3) and dudt (n) for target slots. */
int sphb_force(const sphb_cells *cells, const sphb_consts *k, const uint8_t *target, const double *recB,
               const double *recC, const double *bgrav, const float *hmax_global, double *acc, double *dudt,
               sphb_status *status, const sphb_lists *lists, void *stream);
/* Softened self-gravity, the long-range part (smoothed_gravity_acceleration_comp, pyfluid.py:420-432): exact all-pairs
 * fp64 sum, acc (n,3) (+)= sum_i -G/2 m (grav_force(d,h_j) + grav_force(d,h_i)) (r_j - r_i)/d.  O(n^2).  posh in any
 * order; acc in the same order. */
int sphb_gravity_direct(const double *posh, int64_t n, const sphb_consts *k, int32_t accumulate, double *acc,
                        void *stream);
/* The same sum for a SUBSET of targets against all sources: what each rank of a decomposed run evaluates for its own
 * particles once the sources (x, y, z, h of every rank) have been all-gathered.  k->mass, if set, indexes the sources. */
int sphb_gravity_direct_targets(const double *posh_targets, int64_t n_targets, const double *posh_sources, int64_t n_sources,
                                const sphb_consts *k, int32_t accumulate, double *acc, void *stream);
/* grav_potential, grav_force, grav_potential_smoothlength_derivative elementwise (pyfluid.py:365-399) */
int sphb_grav_eval(const double *dist, const double *h, int64_t n, double *phi, double *force, double *dphidh,
                   void *stream);

/* Per-pair terms for explicit (j, i) slot pairs (pairs[2*p], pairs[2*p+1]), same records as sphb_force:
 * pi_out[p] = Pi(j, i, ...) (pyfluid.py:261-285); acc_out[3*p..] = fluid_accel_viscosity_comp(j, i, ...)
 * (pyfluid.py:459-474; zero for identical positions); grav_out[3*p..] = smoothed_gravity_acceleration_comp(j, i, ...)
 * (:420-432); corr_out[3*p..] = smoothed_gravity_correction_comp(j, i, ...) (:434-437, needs bgrav from sphb_eos).
 * Any output may be NULL. */
int sphb_pair_terms(const double *posh, const double *recB, const double *recC, const uint32_t *pairs, int64_t n_pairs,
                    const sphb_consts *k, double *pi_out, double *acc_out, const double *bgrav, double *grav_out,
                    double *corr_out, sphb_status *status, void *stream);

/* ---- K5: integrator updates (pyfluid.py:511-515 predictor, :524-528 corrector) ------------------- */
int sphb_predict(const double *posh0, const double *vel0, const double *u0, const double *acc0, const double *dudt0,
                 double dt, int64_t n, double *posh_mid, double *vel_mid, double *u_mid, sphb_status *status,
                 void *stream);
int sphb_correct(const double *posh0, const double *vel0, const double *u0, const double *acc_mid,
                 const double *dudt_mid, double dt, int64_t n, double *posh1, double *vel1, double *u1,
                 sphb_status *status, void *stream);

/* ---- helpers ------------------------------------------------------------------------------------ */
/* out[0..5] = min x,y,z, max x,y,z ; out[6] = max h ; out[7] = sum h   over posh (n,4) */
int sphb_bounds(const double *posh, int64_t n, double *out8, void *workspace, int64_t workspace_bytes, void *stream);
int64_t sphb_bounds_workspace_bytes(int64_t n);
/* hmax_out[0] = (float, rounded up) max h over posh */
int sphb_hmax(const double *posh, int64_t n, float *hmax_out, void *stream);
/* distance(rj, ri) = sqrt((rj-ri).dot(rj-ri)) for n pairs of 3-vectors (pyfluid.py:43-47) */
int sphb_distance(const double *rj, const double *ri, int64_t n, double *out, void *stream);
/* M4(dist,h), M4_d1(dist,h), M4_smoothlength_derivative(dist,h) elementwise (pyfluid.py:56-70,85-87);
 * any output may be NULL */
int sphb_m4_eval(const double *dist, const double *h, int64_t n, double *W, double *dW, double *dWdh, void *stream);
/* OR `flag` into status if any h_out < 0 (newton_smoothlength_arr's final check, pyfluid.py:209) */
int sphb_check_neg_h(const double *h, int64_t n, sphb_status *status, void *stream);
/* DFMA-throughput microbenchmark: runs `iters` dependent-chain rounds on blocks x 256 threads and
 * writes elapsed ms (CUDA events, synchronises) and flops done */
int sphb_fp64_peak(int32_t blocks, int32_t iters, double *ms_host, double *flops_host);
/* launch counter: number of kernels this library has launched since load (for bench.py gpu_launches) */
int64_t sphb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SPHB200_H */
