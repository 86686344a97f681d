/*
 * wost.h — C ABI of libwost.so, the B200 (sm_100a) Walk-on-Stars engine.
 *
 * The reference (Tsuchijo/DCRMonteCarlo) is pure Python and has no FFI; its boundary for the walk
 * loop is the Python object API (solvers/WoStSolver.py:22,141-157,319-353; geometry/Polylines.py:8-63;
 * geometry/PolylinesSimple.py:199-307).  This header is what a binding for that path binds instead of
 * the Python loops; INTEGRATION.md shows the ctypes stub.  Each entry point cites the reference
 * interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *  - plain pointers and sizes only; every data pointer may be a HOST pointer (pageable or pinned) or a
 *    DEVICE pointer on the scene's device — the library detects which (cudaPointerGetAttributes) and
 *    stages host buffers itself.  Calls with host outputs return after the results are in place;
 *    calls whose buffers are all on the device are stream-ordered and return immediately.
 *  - `stream` is a cudaStream_t (NULL = legacy default stream).
 *  - the caller owns every in/out buffer; the library owns scene and field handles, which are
 *    immutable after creation and may be used concurrently from several streams.
 *  - every function returns 0 on success or a negative wost_status; wost_last_error() gives the
 *    thread-local message.  There is no CPU fallback: without a CUDA device every compute call fails
 *    with WOST_ERR_CUDA.
 *  - all reals are fp32 unless stated (the reference computes in torch's default float32).
 */
#ifndef WOST_H
#define WOST_H
#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define WOST_VERSION 201   /* major*10000 + minor*100 + patch */

typedef enum {
    WOST_OK = 0,
    WOST_ERR_INVALID = -1,     /* bad argument */
    WOST_ERR_CUDA = -2,        /* CUDA runtime error / no device */
    WOST_ERR_ALLOC = -3,
    WOST_ERR_UNSUPPORTED = -4
} wost_status;

typedef struct wost_scene wost_scene_t;   /* boundary polylines on one device                      */
typedef struct wost_field wost_field_t;   /* one device-resident coefficient / data field          */

/* ---- field descriptors: the device stand-ins for the reference's Python callables --------------
 * g, f, alpha, sigma of WostSolver_2D.__init__ (solvers/WoStSolver.py:22) are evaluated at a (2,)
 * tensor inside the walk loop (:253-256,277-283,295).  Here a field is a constant plus a sum of
 * analytic terms, or a bilinear table, optionally masked. */
enum { WOST_TERM_PRODUCT = 0, WOST_TERM_SIGMOID_CIRCLE = 1 };
enum { WOST_TRIG_NONE = 0, WOST_TRIG_SIN = 1, WOST_TRIG_COS = 2 };
enum { WOST_FIELD_TERMS = 0, WOST_FIELD_GRID = 1 };
enum { WOST_MASK_NONE = 0, WOST_MASK_BOX = 1, WOST_MASK_DISC = 2 };

typedef struct {
    int32_t kind;        /* WOST_TERM_*                                                         */
    int32_t px, py;      /* PRODUCT: monomial x^px y^py                                          */
    int32_t t1, t2;      /* PRODUCT: trig factor kinds WOST_TRIG_*                               */
    float A;             /* amplitude                                                            */
    float q, cx, cy;     /* PRODUCT: exp(-q((x-cx)^2+(y-cy)^2)) if q != 0.  SIGMOID: k = q       */
    float R;             /* SIGMOID_CIRCLE: A*sigmoid(-k(|x-c|-R))  (reference utils.py:123-129) */
    float w1x, w1y, p1;  /* trig1(w1x*x + w1y*y + p1)                                            */
    float w2x, w2y, p2;  /* trig2(w2x*x + w2y*y + p2)                                            */
} wost_term_t;           /* 64 bytes */

typedef struct {
    int32_t kind;        /* WOST_FIELD_*                                                         */
    int32_t n_terms;
    float c0;            /* constant offset                                                      */
    int32_t mask_kind;   /* WOST_MASK_*                                                          */
    float mask[4];       /* BOX: xmin,xmax,ymin,ymax (outside if x<xmin|x>xmax|y<ymin|y>ymax);    */
                         /* DISC: cx,cy,R^2,- (outside if |x-c|^2 > R^2)                          */
    float outside;       /* value outside the mask                                               */
    int32_t nx, ny;      /* GRID: node (i,j) at (x0+i*dx, y0+j*dy), value grid[i*ny+j]; clamped  */
    float x0, y0, dx, dy;
    const wost_term_t* terms;   /* HOST pointer, n_terms entries (copied by wost_field_create)    */
    const float* grid;          /* HOST pointer, nx*ny entries (copied)                            */
} wost_field_desc_t;

/* how sigma' of the delta-tracking branch is formed (solvers/WoStSolver.py:88-127) */
enum {
    WOST_SP_FULL = 0,    /* sigma/alpha + 0.5(lap(alpha)/alpha - |grad ln alpha|^2/2), closed form (:102-121) */
    WOST_SP_RATIO = 1,   /* sigma/alpha — the reference's fallback when autograd fails (:123-127)            */
    WOST_SP_FIELD = 2    /* fields.sigma_prime is sigma' itself (tabulated)                                  */
};

/* which estimator wost_solve runs */
enum {
    WOST_COMPAT_REFERENCE = 0,  /* the reference's walk, quirks included — the parity mode                   */
    WOST_COMPAT_PHYSICAL = 1    /* textbook Walk on Stars: hits by true ray distance, reflection into the hemisphere facing */
                                /* the domain, closing vertex of closed loops is a silhouette candidate, termination        */
                                /* projects onto the Dirichlet boundary, source radius from the disc Green's function with   */
                                /* an independent direction, visibility tested.  With delta_tracking: variable coefficients  */
                                /* through U = sqrt(alpha) u and the screened ball kernel's own weights; sigma_bar must be a */
                                /* true majorant of |sigma'|, steps are capped at 1/sqrt(sigma_bar), screened_icdf is not    */
                                /* used, Neumann walls need d(alpha)/dn = 0.  Not in the reference; validated against        */
                                /* analytic solutions.                                                                      */
};

typedef struct {
    const wost_field_t* g;            /* Dirichlet data (boundaryDirichlet, :45-48,295); NULL = 0  */
    const wost_field_t* f;            /* source (:50,242-258); NULL = no source                     */
    const wost_field_t* alpha;        /* diffusion (:57-60); NULL = 1                               */
    const wost_field_t* sigma;        /* absorption (:55-56,61); NULL = 0                           */
    const wost_field_t* sigma_prime;  /* only for WOST_SP_FIELD                                     */
} wost_fields_t;

typedef struct {
    int64_t n_walks;          /* nWalks of solve() (:319) handled by THIS call                     */
    int32_t max_steps;        /* maxSteps                                                          */
    float eps;                /* eps; rmin = eps/2 (:167)                                          */
    int32_t delta_tracking;   /* use_delta_tracking (:51,64)                                       */
    int32_t sp_mode;          /* WOST_SP_*                                                         */
    float sigma_bar;          /* (:130-136)                                                        */
    const float* screened_icdf;   /* delta tracking: inverse CDF of the screened radius sampler    */
    int32_t icdf_len;             /*   (solvers/utils.py:181-195 as a table; HOST or DEVICE ptr)   */
    uint64_t seed;            /* Philox4x32-10 key                                                 */
    int64_t point_index_base; /* global index of pts[0]  } Philox counter = (point, walk, step, 0) */
    int64_t walk_offset;      /* global index of walk 0  }  => results independent of sharding     */
    int32_t compat_mode;      /* WOST_COMPAT_*                                                     */
    /* physical mode with variable coefficients: optional spatially varying majorant.  A max-pyramid of
     * |sigma'| over the bounding box: level 0 has n x n cells (n = 2^(levels-1), cell (i,j) covers
     * [x0+i dx, x0+(i+1) dx) x [y0+j dy, ...), value at [i*n+j]); level l+1 holds the maxima of 2x2 blocks of
     * level l and follows it in memory; the last level is one cell.  Each step then uses the maximum over the cells
     * its ball touches instead of sigma_bar, and steps are only capped where sigma' is large.  levels = 0: one
     * majorant (sigma_bar) for the whole domain.  HOST or DEVICE pointer. */
    int32_t majorant_levels;
    const float* majorant;
    float majorant_x0, majorant_y0, majorant_dx, majorant_dy;
    /* Per-solver specialised kernel: the fields of this solve compiled into the walk kernel with NVRTC (cached per field
     * set), instead of the interpreter over field descriptors.  Results are bit-identical either way.
     * 0 = auto (jobs of >= 2^18 walks, the 8th solve with the same fields, or whenever the kernel is compiled already), 1 = always (error if NVRTC is
     * unavailable), 2 = never.  The environment variable WOST_JIT (0 / 1) overrides; WOST_JIT_CACHE=<dir> keeps compiled
     * kernels on disk. */
    int32_t jit;
    /* pts[k] has the global index point_index_base + k * point_index_stride (0 is read as 1): a rank of a multi-GPU job
     * that owns every world-th evaluation point passes base = rank, stride = world. */
    int64_t point_index_stride;
} wost_solve_params_t;

#define WOST_WALK_BLOCK 1024   /* walks per deterministic reduction block */

#ifndef __CUDACC_RTC__   /* the entry points are host functions: hidden from the run-time (NVRTC) compile of the kernels */
/* ---- library ------------------------------------------------------------------------------- */
int wost_version(void);
const char* wost_last_error(void);
int wost_device_count(void);                 /* number of CUDA devices, 0 if none / no driver */

/* ---- scene: replaces constructing PolyLinesSimple(points) objects (geometry/PolylinesSimple.py:205-212)
 * for the Dirichlet and (optional) Neumann boundaries of WostSolver_2D.__init__ (:34-35).
 * xy arrays are row-major (N,2).  Zero-length segments are rejected (the reference yields NaN, Q15). */
int wost_scene_create(const float* dirichlet_xy, int32_t n_dirichlet_vtx,
                      const float* neumann_xy, int32_t n_neumann_vtx,
                      int32_t device, wost_scene_t** out);
int wost_scene_destroy(wost_scene_t* scene);
/* A scene keeps the device scratch of its calls (per stream: staged host buffers, per-walk totals, statistics) so that
 * steady-state calls allocate nothing; this returns that memory to the driver (waits for the device first). */
int wost_scene_trim(wost_scene_t* scene, int64_t* out_released_bytes);

int wost_field_create(const wost_field_desc_t* desc, int32_t device, wost_field_t** out);
int wost_field_destroy(wost_field_t* field);
/* evaluate a field (and optionally gradient / Laplacian) at B points — parity entry for the callables */
int wost_field_eval(const wost_field_t* field, const float* p_xy, int64_t B,
                    float* out_v, float* out_gx, float* out_gy, float* out_lap, void* stream);
/* sigma' at B points with the given fields/mode (solvers/WoStSolver.py:88-127) */
int wost_sigma_prime_eval(const wost_fields_t* fields, int32_t sp_mode, const float* p_xy, int64_t B,
                          float* out_sp, void* stream);

/* ---- the walk: replaces WostSolver_2D.solve / _solveUnified (solvers/WoStSolver.py:162-353).
 * pts_xy (n_pts,2).  Outputs (each may be NULL):
 *   out_mean[n_pts]  fp64 mean of the per-walk totals (the reference returns total/nWalks, :311)
 *   out_m2[n_pts]    fp64 sum of squared deviations (variance = m2/(n-1)) — not in the reference
 *   out_block_stats[n_pts * nblk * 2] fp64 (mean, M2) per block of WOST_WALK_BLOCK walks,
 *                    nblk = ceil(n_walks/WOST_WALK_BLOCK): inputs of wost_merge_block_stats for
 *                    walk-sharded multi-GPU runs
 *   out_walk_vals[n_pts * n_walks] fp32 per-walk totals
 *   out_steps[1]     total walk steps taken (one step = one pass of the loop at :206)
 *   trace (return_history, :198-223,261-267,301-309): for the first n_trace walks (point-major flat index)
 *          out_trace[n_trace][trace_cap + 1][8] and out_trace_len[n_trace].  Row k < len is step k:
 *          (x, y, dDirichlet, dNeumann, sample_x, sample_y, source_contribution, 0) — the last four NaN without a
 *          source; row `len` is the terminal record (x_g, y_g, boundary_contribution, walk_total, steps, 0, 0, 1)
 *          where (x_g, y_g) is the point the Dirichlet data was read at.  Unused entries are NaN. */
int wost_solve(const wost_scene_t* scene, const wost_fields_t* fields, const wost_solve_params_t* params,
               const float* pts_xy, int64_t n_pts,
               double* out_mean, double* out_m2, double* out_block_stats, float* out_walk_vals,
               uint64_t* out_steps,
               int64_t n_trace, int32_t trace_cap, float* out_trace, int32_t* out_trace_len,
               void* stream);

/* Shared-walk solve for MANY source terms (DC-resistivity surveys: one source per current-electrode pair).  The walk of
 * solvers/WoStSolver.py:206-291 does not depend on the source f — f only enters the contributions (:253-258) — so one
 * set of walks serves every source: per step each source adds its own contribution, and every source's per-walk total
 * is exactly what wost_solve with that source (fields->f is ignored here) and the same seed would produce.
 * Outputs are source-major: out_mean[n_sources * n_pts], out_m2 likewise, out_block_stats[n_sources * n_pts * nblk * 2].
 * The reference has no counterpart: it re-walks everything per source (tests/testGeophysicalScenario.py:137-151). */
int wost_solve_multi_source(const wost_scene_t* scene, const wost_fields_t* fields, const wost_field_t* const* sources,
                            int32_t n_sources, const wost_solve_params_t* params, const float* pts_xy, int64_t n_pts,
                            double* out_mean, double* out_m2, double* out_block_stats, uint64_t* out_steps, void* stream);

/* Fixed-order Chan merge of per-block (mean, M2) into per-point (mean, M2): the same device code
 * wost_solve runs internally, exposed so that gathered shards merge bit-identically.
 * block_stats[n_pts * nblk * 2]; block b holds min(WOST_WALK_BLOCK, n_walks - b*WOST_WALK_BLOCK) walks. */
int wost_merge_block_stats(const double* block_stats, int64_t n_pts, int64_t n_walks, int32_t device,
                           double* out_mean, double* out_m2, void* stream);

/* ---- geometry primitives, batched over B queries (parity entry points) ------------------------
 * which: 0 = Dirichlet polyline of the scene, 1 = Neumann polyline. */
/* PolyLinesSimple.distance (geometry/PolylinesSimple.py:214-224 -> :26-49); out_seg = arg-min segment (first) */
int wost_geom_distance(const wost_scene_t* scene, int32_t which, const float* p_xy, int64_t B,
                       float* out_d, int32_t* out_seg, void* stream);
/* PolyLinesSimple.silhouetteDistance / isSilhouette (:242-265 -> :52-102); out_mask (B, V-2) optional */
int wost_geom_silhouette(const wost_scene_t* scene, int32_t which, const float* p_xy, int64_t B,
                         float* out_d, uint8_t* out_mask, void* stream);
/* PolyLinesSimple.rayIntersection (:281-292 -> :105-132); out_s (B, S): segment parameter s or +inf */
int wost_geom_ray(const wost_scene_t* scene, int32_t which, const float* p_xy, const float* dir_xy, int64_t B,
                  float* out_s, void* stream);
/* PolyLinesSimple.intersectPolylines (:294-307 -> :135-197) */
int wost_geom_intersect(const wost_scene_t* scene, int32_t which, const float* p_xy, const float* dir_xy,
                        const float* r, int64_t B,
                        float* out_pt, float* out_nrm, uint8_t* out_found, int32_t* out_seg, void* stream);

/* specialised-kernel bookkeeping: kernels compiled by NVRTC so far, cache hits, solves that ran a specialised kernel;
 * wost_jit_last_note() says why the last solve did not ("" if it did) */
int wost_jit_stats(int64_t* out_compiled, int64_t* out_cache_hits, int64_t* out_launches);
const char* wost_jit_last_note(void);
/* developer diagnostic, no device needed: generate + compile the specialised kernel for five field descriptors
 * (g, f, alpha, sigma, sigma_prime; NULL = absent) and write <prefix>.cu / <prefix>.cubin for cuobjdump */
int wost_jit_offline(const wost_field_desc_t* const descs[5], int32_t neu, int32_t src, int32_t delta, int32_t trace,
                     int32_t phys, int32_t big, int32_t multi, int32_t sp_mode, int32_t min_blocks,
                     int32_t n_dirichlet_seg /* -1: run-time sizes */, int32_t n_neumann_seg, const char* arch, const char* prefix);

/* Self-test (tests): n random operand sets through the library's two hand-written division sequences (csrc/wost_device.cuh:
 * div2_by_near_one; the reciprocal form of dirichlet_distance<RCP> for the given divisors, each verified on the host first)
 * and square root (sqrt_in_range) against the compiler's IEEE operations.  out_mismatches = {quotients of unit directions,
 * random operands near one, reciprocal form, square roots}. */
int wost_selftest_division(int32_t device, int64_t n, uint64_t seed, const float* divisors, int32_t n_divisors,
                           int64_t out_mismatches[4]);

/* FP32 FMA-chain microbenchmark: measured non-tensor fp32 TFLOP/s of the device (roofline denominator) */
int wost_fp32_peak(int32_t device, double* out_tflops, double* out_sm_mhz_effective);
#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* WOST_H */
