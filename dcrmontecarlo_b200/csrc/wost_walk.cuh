// wost_walk.cuh — the walk kernel body (sm_100a): one thread per walk, persistent warps with walk regeneration.
//
// Included by wost_lib.cu (static instantiations; fields evaluated by the interpreter of wost_device.cuh) and compiled
// at run time by NVRTC for a solver's own fields (wost_jit in wost_lib.cu: the field provider FP is then generated code
// in which every field is a compile-time constant expression, so the interpreter folds to straight-line arithmetic).
#pragma once
#include "wost_device.cuh"

namespace wost {


struct WalkArgs {
    const float4* dseg; int n_dseg;
    const float4* nseg; int n_nseg;
    int stage_smem;                        // bit 0 / 1: Dirichlet / Neumann segment table is staged in shared memory
    DevFields F;
    const float* pts; long long n_pts; long long n_walks;
    const float* alpha0;                   // delta tracking: alpha at every evaluation point (all walks of a point start there)
    int max_steps; float eps, rmin;
    int sp_mode; float sigma_bar, inv_sigma_bar, sqrt_sigma_bar;
    const float* icdf; int icdf_len;
    const float* iprob;                    // 1 - 1/I0(z) on [0, 21]: the table of include/wost_math.h (delta tracking)
    uint32_t key0, key1; uint32_t ks[20];  // Philox key and its ten round keys
    long long point_index_base, point_index_stride, walk_offset;   // global index of point p: base + p * stride
    float* walk_vals;
    unsigned long long* counter;           // next unassigned flat walk index
    unsigned long long* steps_total;
    int chunk;                             // walks a warp reserves per atomic
    uint32_t walks_magic;                  // floor(2^32 / n_walks) (0xffffffff for n_walks = 1): flat walk index -> (point, walk)
    int lanes;                             // lanes of a warp that take walks (32; fewer for small jobs: walks that do not share
                                           // a warp do not wait for each other's divergent branches, see solve_impl)
    float ndisc_x, ndisc_y, ndisc_r, ndisc_r2;   // disc enclosing the Neumann polyline (inflated), for culling
    int sil_coop_max, ray_coop_max;        // answer a query cooperatively when at most this many lanes need it
    Bvh dbvh, nbvh; float bvh_slack;       // hierarchies for large polylines (nodes == nullptr: brute force)
    WideBvh nwide; int wide_coop_max;      // 32-wide Neumann hierarchy: cooperative queries when few lanes need one
    WideBvh dwide;                         // 32-wide Dirichlet hierarchy: the distance query, one warp per query
    int neu_closed; float phys_nudge;      // physical mode: closed Neumann loop?  pull-back of a reflected walker
    float phys_rcap;                       // physical mode with variable coefficients: step radius cap 1/sqrt(sigma_bar)
    MajorantPyramid maj;                   // ... or a spatially varying majorant (data == nullptr: sigma_bar everywhere)
    const float4* src_support;             // per source (cx, cy, R^2): exactly zero outside
    SourceGrid sgrid;                      // which sources can be non-zero where (masks == nullptr: test every source's disc)
    int n_src; const DevField* srcs;       // shared-walk multi-source solve: n_src > 0 source fields (device array); per-walk
                                           // totals then form rows walk_vals[walk][n_src]
    long long n_trace; int trace_cap; float* trace; int* trace_len;
};


// Shared-walk solves evaluate MANY source terms per step, nearly all of them exactly zero at any one point (a DC-resistivity
// source is a pair of narrow Gaussian blobs at two electrodes; each blob takes its exact-zero underflow shortcut beyond
// sqrt(110 / q)).  The host bins the blobs' discs into a regular grid over their bounding box: per cell one bit per source that
// can be non-zero anywhere in the cell (conservative: discs inflated by a cell), `outside` for points off the grid.  A step
// then visits the set bits of ONE mask instead of testing every source.  Skipped sources contribute exactly nothing either
// way, so every source's per-walk total keeps the bits of a single-source solve.

// calls body(k, f_k(x, y)) for every source k that may be non-zero at (x, y), in increasing k
template <class FP, class Body>
__device__ __forceinline__ void for_each_source(const WalkArgs& a, float x, float y, Body&& body) {
    if (a.sgrid.masks) {
        const float fx = (x - a.sgrid.x0) * a.sgrid.inv_dx, fy = (y - a.sgrid.y0) * a.sgrid.inv_dy;
        long long cell = (long long)a.sgrid.nx * a.sgrid.ny;                           // the `outside` mask
        if (fx >= 0.0f && fy >= 0.0f && fx < (float)a.sgrid.nx && fy < (float)a.sgrid.ny) cell = (long long)(int)fy * a.sgrid.nx + (int)fx;
        const unsigned long long* m = a.sgrid.masks + cell * a.sgrid.words;
        if (a.sgrid.blobs) {
            // Blob mode.  f_k = ((0 + t_0) + t_1) + ... in term order (field_eval_inl); a blob that is not listed here is
            // beyond its exact-zero radius and would add +-0, which changes nothing unless the whole sum is zero -- and a zero
            // sum is skipped by the caller either way.  Each listed blob: the arithmetic of term_value for a bare Gaussian.
            int cur = -1; float acc = 0.0f;
            for (int w = 0; w < a.sgrid.words; ++w) {
                unsigned long long bits = __ldg(m + w);
                while (bits) {
                    const int b = w * 64 + __ffsll((long long)bits) - 1;
                    bits &= bits - 1ull;
                    const int k = __ldg(a.sgrid.blob_src + b);
                    if (k != cur) { if (cur >= 0) body(cur, acc); cur = k; acc = 0.0f; }
                    const float4 g = __ldg(a.sgrid.blobs + b);                          // A, q, cx, cy
                    const float ddx = x - g.z, ddy = y - g.w, e = -g.y * (ddx * ddx + ddy * ddy);
                    acc += e < -110.0f ? g.x * 0.0f * 1.0f : g.x * wm_expf(e);
                }
            }
            if (cur >= 0) body(cur, acc);
            return;
        }
        for (int w = 0; w < a.sgrid.words; ++w) {
            unsigned long long bits = __ldg(m + w);
            while (bits) {
                const int k = w * 64 + __ffsll((long long)bits) - 1;
                bits &= bits - 1ull;
                body(k, FP::source(a, k, x, y));
            }
        }
    } else {
        for (int k = 0; k < a.n_src; ++k) {
            const float4 sup = __ldg(a.src_support + k);
            if ((x - sup.x) * (x - sup.x) + (y - sup.y) * (y - sup.y) > sup.z) continue;   // exactly zero there
            body(k, FP::source(a, k, x, y));
        }
    }
}

// ---- field providers ---------------------------------------------------------------------------------------------------
// The walk reads its fields through a provider FP: g, f, alpha (value and jet), sigma, sigma' table, sources of a
// shared-walk solve.  InterpFP: the interpreter over field descriptors (delta-tracking and Dirichlet-only Laplace kernels
// read headers and term tables from shared memory, the others from the kernel parameters -- measured: +24 % cfg 1b,
// +5 % cfg 1a, -2 % on the mixed-boundary Laplace kernel).  The specialised kernels define their own provider.
template <bool NEU, bool SRC, bool DELTA>
struct InterpFP {
    static constexpr bool SMF = DELTA || (!NEU && !SRC);
    static constexpr bool STAGE_FIELDS = true;         // copy field headers / terms to shared memory at CTA start
    static constexpr bool MULTI = true;                // shared-walk multi-source code compiled in
    static constexpr int N_DSEG = -1, N_NSEG = -1, SIL_COOP_MAX = 0, RAY_COOP_MAX = 0, STAGE = 0;   // scene sizes: run-time values (WalkArgs)
    static constexpr bool DIR_RCP = false;             // Dirichlet distance: generic IEEE division (dirichlet_distance, wost_device.cuh)
    static constexpr bool ALPHA_IN_RANGE = false;      // alpha's bounds are not known at compile time
    __device__ __forceinline__ static bool has_g(const WalkArgs& a) { return a.F.g.present != 0; }
    __device__ __forceinline__ static bool has_alpha(const WalkArgs& a) { return a.F.alpha.present != 0; }
    __device__ __forceinline__ static bool has_sigma(const WalkArgs& a) { return a.F.sigma.present != 0; }
    __device__ __forceinline__ static int sp_mode(const WalkArgs& a) { return a.sp_mode; }
    __device__ __forceinline__ static float g(const WalkArgs& a, float x, float y) {
        return DELTA ? field_eval_s(FIELD_G, x, y) : (SMF ? field_eval_inl<true>(shared_field(FIELD_G), x, y) : field_eval_inl<false>(a.F.g, x, y));
    }
    __device__ __forceinline__ static float f(const WalkArgs& a, float x, float y) {
        return DELTA ? field_eval_s(FIELD_F, x, y) : field_eval_inl<false>(a.F.f, x, y);
    }
    __device__ __forceinline__ static float source(const WalkArgs& a, int k, float x, float y) {
        return SMF ? field_eval_s(FIELD_SOURCE0 + k, x, y) : field_eval(a.srcs[k], x, y);
    }
    __device__ __forceinline__ static float alpha(const WalkArgs& a, float x, float y) {
        return a.F.alpha.present ? field_eval_s(FIELD_ALPHA, x, y) : 1.0f;
    }
    __device__ __forceinline__ static Jet alpha_jet(const WalkArgs&, float x, float y) { return field_jet_s(FIELD_ALPHA, x, y); }
    __device__ __forceinline__ static float sigma(const WalkArgs&, float x, float y) { return field_eval_s(FIELD_SIGMA, x, y); }
    __device__ __forceinline__ static float sigma_prime_field(const WalkArgs&, float x, float y) { return field_eval_s(FIELD_SIGMA_PRIME, x, y); }
};

// sigma' (solvers/WoStSolver.py:88-127) in closed form through a provider: the arithmetic of sigma_prime_at (wost_device.cuh).
// `alpha_xy`: alpha(x, y) if the caller has it already, a negative value otherwise.
// `jet`: value / gradient / Laplacian of alpha at (x, y) if the caller has evaluated them already (WOST_SP_FULL), else nullptr.
template <class FP>
__device__ __forceinline__ float sigma_prime_fp(const WalkArgs& a, float x, float y, float alpha_xy, const Jet* jet = nullptr) {
    const int mode = FP::sp_mode(a);
    if (mode == WOST_SP_FIELD) return FP::sigma_prime_field(a, x, y);
    const float sg = FP::has_sigma(a) ? FP::sigma(a, x, y) : 0.0f;
    if (mode == WOST_SP_RATIO) return div_z(sg, fmaxf(alpha_xy >= 0.0f ? alpha_xy : FP::alpha(a, x, y), 1e-8f));
    Jet j; j.v = 1.0f; j.gx = j.gy = j.l = 0.0f;
    if (FP::has_alpha(a)) j = jet ? *jet : FP::alpha_jet(a, x, y);
    if (j.v < 1e-8f) { j.v = 1e-8f; j.gx = j.gy = j.l = 0.0f; }
    const float ratio = div_z(sg, j.v);
    const float la = j.v + 1e-8f, lgx = div_z(j.gx, la), lgy = div_z(j.gy, la);
    const float corr = 0.5f * ((j.l + 1e-8f) / j.v - (lgx * lgx + lgy * lgy) / 2.0f);
    return ratio + corr;
}

// One thread per walk, persistent warps with walk regeneration: a lane whose walk has terminated is handed the
// next walk index in the same loop iteration, so all 32 lanes keep stepping until the job runs dry (walk lengths
// are geometric-tailed; without regeneration a warp idles until its longest walk ends).  Warps reserve `chunk`
// consecutive walk indices per global atomic and deal them out with ballot/popc.
//
// The loop body restates solvers/WoStSolver.py:206-298 of the reference, quirks included.
// BIG: the scene has hierarchies (large polylines); small scenes get a kernel without any traversal code.
template <bool NEU, bool SRC, bool DELTA, bool TRACE, bool PHYS, bool BIG, class FP>
__device__ __forceinline__ void walk_body(const WalkArgs& a) {
    extern __shared__ float4 smem[];
    const float4* dseg = a.dseg; const float4* nseg = a.nseg;
    // polyline sizes and query strategy: run-time values, or compile-time constants in a kernel specialised for the scene
    // (then the unused query strategies below are not even compiled)
    const int n_dseg = FP::N_DSEG >= 0 ? FP::N_DSEG : a.n_dseg, n_nseg = FP::N_NSEG >= 0 ? FP::N_NSEG : a.n_nseg;
    const int sil_coop_max = FP::N_NSEG >= 0 ? FP::SIL_COOP_MAX : a.sil_coop_max, ray_coop_max = FP::N_NSEG >= 0 ? FP::RAY_COOP_MAX : a.ray_coop_max;
    // Shared memory: [field headers: g, f, alpha, sigma, sigma', then the sources of a shared-walk solve][segment tables]
    // [the fields' term tables].  The delta-tracking kernels' out-of-line field interpreter reads headers and terms from
    // here (wost_device.cuh, SM = true); the other kernels inline the interpreter and read the kernel parameters directly.
    constexpr bool SMF = FP::SMF;                      // the interpreter reads field tables from shared memory
    constexpr bool TINY_NEU = NEU && !BIG && FP::N_NSEG >= 1 && FP::N_NSEG <= 2;   // scene-specialised kernels only
    if (SMF && FP::STAGE_FIELDS) {
        const int nf = FIELD_SOURCE0 + a.n_src;
        uint32_t* hdr = reinterpret_cast<uint32_t*>(smem);
        for (int i = threadIdx.x; i < nf * DEVFIELD_F4 * 4; i += blockDim.x) {
            const int f = i / (DEVFIELD_F4 * 4), w = i - f * (DEVFIELD_F4 * 4);
            const uint32_t* src = f < FIELD_SOURCE0 ? reinterpret_cast<const uint32_t*>(&a.F.g + f) : reinterpret_cast<const uint32_t*>(a.srcs + (f - FIELD_SOURCE0));
            hdr[i] = src[w];
        }
        __syncthreads();
        const DevField* sF = reinterpret_cast<const DevField*>(smem);
        for (int f = 0; f < nf; ++f) {
            const int n4 = 4 * sF[f].n_terms, off = sF[f].term_off;
            const float4* src = reinterpret_cast<const float4*>(sF[f].terms);
            for (int i = threadIdx.x; i < n4; i += blockDim.x) smem[off + i] = __ldg(src + i);
        }
        __syncthreads();
    }
    // segment tables without a hierarchy are staged into shared memory (all lanes read the same segment: broadcast);
    // polylines with a hierarchy are read through L1 by the traversals.  stage_smem: bit 0 Dirichlet, bit 1 Neumann.
    {
        float4* sp = smem + DEVFIELD_F4 * (FIELD_SOURCE0 + a.n_src);
        const int stage_smem = FP::N_DSEG >= 0 ? FP::STAGE : a.stage_smem;   // compile-time in scene-specialised kernels: LDS, not generic loads
        if (stage_smem & 1) {
            for (int i = threadIdx.x; i < 2 * n_dseg; i += blockDim.x) sp[i] = a.dseg[i];
            dseg = sp; sp += 2 * n_dseg;
        }
        if (NEU && (stage_smem & 2)) {
            for (int i = threadIdx.x; i < 2 * n_nseg; i += blockDim.x) sp[i] = a.nseg[i];
            nseg = sp;
        }
        if (stage_smem) __syncthreads();
    }
    // Terminated walks are parked here (where g is read, the walk's weight and running total, its index) and their
    // boundary term is evaluated later for many lanes at once: evaluated on the spot, g would run with the one or two
    // lanes that happen to terminate in an iteration (ncu, cfg 4: 18 % of all instructions at 2 of 32 lanes).
    __shared__ float4 parked[2 * 256];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long total = (unsigned long long)a.n_pts * (unsigned long long)a.n_walks;

    // warp-uniform reservation [next, end)
    unsigned long long next = 0, end = 0;
    bool exhausted = false;
    unsigned parked_mask = 0u;                 // warp-uniform: lanes with a parked walk
    bool flush_now = false;                    // warp-uniform: a lane terminated again while still parked

    // per-lane walk state
    bool active = false, retired = lane >= a.lanes;
    unsigned long long id = 0; uint32_t pidx = 0, widx = 0;
    float x = 0.f, y = 0.f, dD = 1.0f, atten = 1.0f, total_v = 0.0f, phi_n = 0.0f;
    float alpha_x = 1.0f;                      // alpha at the walker's position (delta tracking), carried from step to step
    bool onB = false; int steps = 0;
    unsigned long long steps_acc = 0;
    uint32_t o[4] = {0u, 0u, 0u, 0u};          // Philox block (kept across steps: Laplace walks use one word per step)
    const float4 nseg0 = (NEU && lane < n_nseg) ? a.nseg[2 * lane] : make_float4(0.f, 0.f, 0.f, 0.f);   // segment `lane`, register-resident

    while (true) {
        // ---- regeneration ------------------------------------------------------------------------
        unsigned need = __ballot_sync(FULL, !active && !retired);
        while (need) {
            const bool mine = (need >> lane) & 1u;
            if (next >= end && !exhausted) {                           // reserve the next chunk of walk indices
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.counter, (unsigned long long)a.chunk);
                base = __shfl_sync(FULL, base, 0);
                next = base < total ? base : total;
                end = base + (unsigned long long)a.chunk < total ? base + (unsigned long long)a.chunk : total;
                if (next >= end) exhausted = true;
            }
            if (next >= end) { if (mine) retired = true; break; }      // job ran dry
            const unsigned long long avail = end - next;
            const int rank = __popc(need & lt_mask);
            if (mine && (unsigned long long)rank < avail) {
                id = next + (unsigned long long)rank;
                unsigned long long p, w;
                if (total <= 0xffffffffull) {                          // the usual case: 32 bits, division by multiplication
                    // walks_magic = floor(2^32 / n_walks): the estimate is the quotient or one less (id < 2^32)
                    uint32_t p32 = __umulhi((uint32_t)id, a.walks_magic);
                    uint32_t w32 = (uint32_t)id - p32 * (uint32_t)a.n_walks;
                    if (w32 >= (uint32_t)a.n_walks) { ++p32; w32 -= (uint32_t)a.n_walks; }
                    p = p32; w = w32;
                } else { p = id / (unsigned long long)a.n_walks; w = id - p * (unsigned long long)a.n_walks; }
                pidx = (uint32_t)(a.point_index_base + (long long)p * a.point_index_stride); widx = (uint32_t)(a.walk_offset + (long long)w);
                x = __ldg(a.pts + 2 * p); y = __ldg(a.pts + 2 * p + 1);
                dD = 1.0f;                                             // :190 sentinel (Q6)
                atten = 1.0f; total_v = 0.0f; onB = false; phi_n = 0.0f; steps = 0;   // :188-195
                if (DELTA) alpha_x = __ldg(a.alpha0 + p);                // = alpha_at<true>(a.F, x, y), evaluated once per point
                if (PHYS && DELTA) atten = 1.0f / sqrtf(alpha_x);          // u = U / sqrt(alpha): the walk estimates U
                active = true;
            }
            const unsigned long long cnt = (unsigned long long)__popc(need);
            next += cnt < avail ? cnt : avail;
            need = __ballot_sync(FULL, !active && !retired);
        }
        // Dirichlet-only delta-tracking kernels: one CTA barrier per iteration keeps the warps in the same code region.  With the
        // interpreter (static kernels, instruction-fetch bound) that is +5 %; with the specialised kernels it is within noise
        // (1.49 vs 1.43 / 1.52 vs 1.56 ms on cfg 1b).  For the kernels with cooperative Neumann loops it is a loss
        // (iteration times differ per warp), so only here.
        bool none;
        if (!NEU && DELTA && !PHYS) none = !__syncthreads_or(active ? 1 : 0);
        else none = __ballot_sync(FULL, active) == 0u;

        // ---- boundary terms of the parked walks (:295-298), all parked lanes together ----------------------------------
        if ((flush_now || none) && parked_mask) {
            if ((parked_mask >> lane) & 1u) {
                const float4 p0 = parked[threadIdx.x], p1 = parked[256 + threadIdx.x];
                const float gx_ = p0.x, gy_ = p0.y, w_ = p0.z, tot_ = p0.w;
                const unsigned long long pid = ((unsigned long long)__float_as_uint(p1.y) << 32) | __float_as_uint(p1.x);
                float bc = 0.0f;
                if (FP::has_g(a)) bc = FP::g(a, gx_, gy_);
                if (PHYS && DELTA) bc = w_ * (bc * sqrtf(FP::alpha(a, gx_, gy_)));          // U = sqrt(alpha) g on the boundary
                else if (DELTA) bc = bc * w_;
                if (SRC && FP::MULTI && a.n_src > 0) {                                     // one total per source: same walk, same boundary term
                    float* row = a.walk_vals + (size_t)pid * a.n_src;
                    for (int k = 0; k < a.n_src; ++k) row[k] = row[k] + bc;
                } else a.walk_vals[pid] = tot_ + bc;
                if (TRACE) {
                    if ((long long)pid < a.n_trace) {                         // terminal row: where g was read, what it contributed
                        const int nst = __float_as_int(p1.z), len = min(nst, a.trace_cap);
                        a.trace_len[pid] = len;
                        float4* t = reinterpret_cast<float4*>(a.trace) + ((size_t)pid * (a.trace_cap + 1) + len) * 2;
                        t[0] = make_float4(gx_, gy_, bc, tot_ + bc); t[1] = make_float4((float)nst, 0.0f, 0.0f, 1.0f);
                    }
                }
            }
            parked_mask = 0u;
        }
        flush_now = false;
        if (none) break;

        // ---- this iteration: every active lane either takes one step of the reference's loop or terminates ----
        // reference: the loop condition tests the PREVIOUS step's dDirichlet (:206, Q5).
        // physical:  the distance at the current position decides, and g is read at the closest boundary point.
        int dir_arg = -1;
        if (PHYS && active)
            dD = (BIG && a.dbvh.nodes) ? bvh_dirichlet_distance(a.dseg, n_dseg, a.dbvh, x, y, &dir_arg) : dirichlet_distance<FP::DIR_RCP>(dseg, n_dseg, x, y, &dir_arg);
        const bool stepping = active && steps < a.max_steps && dD > a.eps && !(PHYS && DELTA && atten == 0.0f);   // weight 0: absorbed
        {
            // terminal: the boundary contribution is read at the un-projected point (:295-298, Q5/Q7); park the walk.
            // A lane whose previous walk is still parked waits one iteration: the parked walks are resolved first.
            const bool term = active && !stepping;
            const unsigned tmask = __ballot_sync(FULL, term);
            if (!TRACE && !FP::has_g(a) && !(SRC && FP::MULTI && a.n_src > 0)) {        // g = 0: nothing to evaluate, nothing to park
                if (term) {
                    a.walk_vals[id] = total_v + (DELTA ? 0.0f * atten : 0.0f);
                    steps_acc += (unsigned long long)steps;
                    active = false;
                }
            } else if (tmask) {
                const unsigned conflict = tmask & parked_mask;
                if (term && !((parked_mask >> lane) & 1u)) {
                    float gx_ = x, gy_ = y;
                    if (PHYS && dir_arg >= 0) segment_closest_point(a.dseg[2 * dir_arg], a.dseg[2 * dir_arg + 1], x, y, gx_, gy_);
                    parked[threadIdx.x] = make_float4(gx_, gy_, atten, total_v);
                    parked[256 + threadIdx.x] = make_float4(__uint_as_float((unsigned)id), __uint_as_float((unsigned)(id >> 32)), __int_as_float(steps), 0.0f);
                    steps_acc += (unsigned long long)steps;
                    active = false;
                }
                parked_mask |= tmask & ~conflict;
                flush_now = conflict != 0u;
            }
        }

        // very large Dirichlet polylines: the distance query of every stepping lane, one warp-cooperative descent each
        if (BIG && !PHYS && a.dwide.boxes) {
            unsigned need = __ballot_sync(FULL, stepping);
            while (need) {
                const int src = __ffs(need) - 1; need &= need - 1u;
                const float q = wide_dirichlet_distance_sq(a.dseg, n_dseg, a.dwide, __shfl_sync(FULL, x, src), __shfl_sync(FULL, y, src), lane, nullptr);
                if (lane == src) dD = sqrtf(q);
            }
        }

        // ---- phase A: Dirichlet distance, direction --------------------------------------------------------------
        float dN = CUDART_INF_F, r = 0.f, dx = 0.f, dy = 0.f, ex = 0.f, ey = 0.f, ox = 0.f, oy = 0.f;
        float pd_sb = a.sigma_bar;                 // physical delta tracking: the majorant this step uses
        bool want_ray = false, want_sil = false;
        float gap = 0.0f;
        if (stepping) {
            if (!PHYS && !(BIG && a.dwide.boxes))
                dD = (BIG && a.dbvh.nodes) ? bvh_dirichlet_distance(a.dseg, n_dseg, a.dbvh, x, y, nullptr)
                                  : dirichlet_distance<FP::DIR_RCP>(dseg, n_dseg, x, y, nullptr);      // :208
            uint32_t w0;
            if (PHYS) {
                philox4x32_10_ks(pidx, widx, (uint32_t)steps, 2u, a.ks, o);      // stream tag 2: physical mode
                w0 = o[0];
            } else if (!SRC && !DELTA) {
                // Laplace walks use one 32-bit word per step: one Philox block (stream tag 1) serves four steps
                const int sel = steps & 3;
                if (sel == 0) philox4x32_10_ks(pidx, widx, (uint32_t)steps >> 2, 1u, a.ks, o);
                w0 = sel == 0 ? o[0] : (sel == 1 ? o[1] : (sel == 2 ? o[2] : o[3]));
            } else {
                philox4x32_10_ks(pidx, widx, (uint32_t)steps, 0u, a.ks, o);
                w0 = o[0];
            }
            float theta;
            if (PHYS) {
                // uniform direction; on a reflecting wall, uniform in the hemisphere around the inward normal
                theta = (NEU && onB) ? phi_n + (u24(w0) - 0.5f) * 3.14159274101257324f : (u24(w0) * 2.0f) * 3.14159274101257324f;
            } else {
                theta = (u24(w0) * 2.0f) * 3.14159274101257324f;                        // :226
                if (NEU && onB) theta = theta / 2.0f + phi_n;                           // :227-228 (Q2)
            }
            if (PHYS) sincosf(theta, &dy, &dx);
            else wm_sincosf_small(theta, &dy, &dx);                                     // :230-232 (|theta| < 10)
            if (NEU) {
                if (PHYS) { ex = dx; ey = dy; ox = x; oy = y; }
                else {
                    // intersect_polylines_jit prologue (:149-159): normalise, offset the origin by 1e-6
                    const float dn = sqrt_in_range(norm2_sq(dx, dy));                   // torch.norm (:151); 1 to within a few ulp: (dx, dy) = (cos, sin)
                    div2_by_near_one(dx, dy, dn, ex, ey);                               // ex = dx / dn, ey = dy / dn (:152)
                    ox = x + 1e-6f * ex; oy = y + 1e-6f * ey;
                }
                // the silhouette distance only matters if it can be smaller than dDirichlet (:212): every Neumann
                // vertex is at least (|p - c| - R) away, so outside that margin min(dD, dN) = dD without looking.
                if (TINY_NEU) {
                    // a Neumann polyline of one or two segments (the DCR scenes' ground surface): testing every ray against it
                    // with the exact arithmetic costs less than the cull + prefilter + the divergence they cause; one segment
                    // has no interior vertex, hence no silhouette vertex (:64-81)
                    want_sil = TRACE || n_nseg > 1;
                    want_ray = true;
                    if (PHYS) { const float gx = x - a.ndisc_x, gy = y - a.ndisc_y; gap = sqrtf(gx * gx + gy * gy) - a.ndisc_r; }
                } else {
                const float gx = x - a.ndisc_x, gy = y - a.ndisc_y;
                const float g2 = gx * gx + gy * gy, lim = dD * 1.001f + a.ndisc_r;   // gap = sqrt(g2) - R > 1.001 dD, without the root
                if (PHYS) gap = sqrtf(g2) - a.ndisc_r;
                want_sil = TRACE || !(lim * lim < g2);
                want_ray = ray_may_hit_disc(ox, oy, ex, ey, a.ndisc_x, a.ndisc_y, a.ndisc_r2);
                }
                // physical hits are limited to the star radius r <= max(dD, rmin): farther polylines cannot be hit
                if (PHYS && gap > fmaxf(dD, a.rmin) + a.phys_nudge) want_ray = false;
            }
        }

        // ---- phase B: queries against the Neumann polyline -------------------------------------------------------------
        // Each query is answered per lane (a loop over all segments) when most lanes of the warp need it, or
        // warp-cooperatively (32 segments per instruction for one query) when only a few do.
        float best_s = CUDART_INF_F; int best_k = -1;
        if (NEU) {
            const bool small = n_nseg <= 32;                                          // warp-uniform
            float dN2 = CUDART_INF_F;                                                   // squared; rooted once below
            unsigned need = __ballot_sync(FULL, want_sil);                              // silhouette distance (:211)
            if (BIG && a.nbvh.nodes) {
                // large polyline; only vertices closer than dDirichlet can change r (:212).  Few lanes: one warp-cooperative
                // descent of the 32-wide tree per query; many lanes: every lane descends the binary tree itself.
                const float bound = TRACE ? CUDART_INF_F : dD * dD * 1.000001f;
                if (a.nwide.boxes && __popc(need) <= a.wide_coop_max) {
                    while (need) {
                        const int src = __ffs(need) - 1; need &= need - 1u;
                        const float q = wide_silhouette_distance_sq(a.nseg, n_nseg, a.nwide, __shfl_sync(FULL, x, src), __shfl_sync(FULL, y, src),
                                                                    __shfl_sync(FULL, bound, src), lane);
                        dN2 = lane == src ? q : dN2;
                    }
                } else if (want_sil) dN2 = bvh_silhouette_distance_sq(a.nseg, n_nseg, a.nbvh, x, y, bound);
            } else if (sil_coop_max == 0 || __popc(need) > sil_coop_max) {
                if (want_sil) dN2 = silhouette_distance_sq(nseg, n_nseg, x, y);
            } else {
                while (need) {
                    const int src = __ffs(need) - 1; need &= need - 1u;
                    const float qx_ = __shfl_sync(FULL, x, src), qy_ = __shfl_sync(FULL, y, src);
                    const float q = small ? silhouette_distance_sq_coop<true>(nseg, n_nseg, nseg0, qx_, qy_, lane)
                                          : silhouette_distance_sq_coop<false>(nseg, n_nseg, nseg0, qx_, qy_, lane);
                    dN2 = lane == src ? q : dN2;
                }
            }
            if (PHYS && a.neu_closed && want_sil) dN2 = fminf(dN2, closing_vertex_silhouette_sq(a.nseg, n_nseg, x, y));
            dN = dN2 == CUDART_INF_F ? CUDART_INF_F : sqrtf(dN2);           // no silhouette vertex: spare sqrt its special-value path
            need = __ballot_sync(FULL, want_ray);                                       // ray vs polyline (:162-178)
            if (BIG && a.nbvh.nodes) {
                if (a.nwide.boxes && __popc(need) <= a.wide_coop_max) {
                    while (need) {
                        const int src = __ffs(need) - 1; need &= need - 1u;
                        float cs; int ck;
                        wide_ray_cast<PHYS>(a.nseg, n_nseg, a.nwide, a.bvh_slack, __shfl_sync(FULL, ox, src), __shfl_sync(FULL, oy, src),
                                            __shfl_sync(FULL, ex, src), __shfl_sync(FULL, ey, src), lane, cs, ck);
                        best_s = lane == src ? cs : best_s; best_k = lane == src ? ck : best_k;
                    }
                } else if (want_ray) bvh_ray_cast<PHYS>(a.nseg, n_nseg, a.nbvh, a.bvh_slack, ox, oy, ex, ey, best_s, best_k);
            } else if (TINY_NEU) {
                if (want_ray) {
#pragma unroll
                    for (int k = 0; k < (FP::N_NSEG > 0 ? FP::N_NSEG : 1); ++k) {
                        const float s = ray_segment_exact<PHYS>(nseg[2 * k], ox, oy, ex, ey);
                        if (s < best_s) { best_s = s; best_k = k; }
                    }
                }
            } else if (ray_coop_max == 0 || __popc(need) > ray_coop_max) {
                if (want_ray) ray_cast<PHYS>(nseg, n_nseg, ox, oy, ex, ey, best_s, best_k);
            } else if (small) {
                // one segment per lane: the warp runs the division-free prefilter for each ray in turn and hands the
                // ray's owner the mask of candidate segments; afterwards all owners resolve their (one or two)
                // candidates at the same time with the reference's exact arithmetic, in index order (first index wins ties)
                unsigned cand = 0u;
                while (need) {
                    const int src = __ffs(need) - 1; need &= need - 1u;
                    const float rox = __shfl_sync(FULL, ox, src), roy = __shfl_sync(FULL, oy, src);
                    const float rex = __shfl_sync(FULL, ex, src), rey = __shfl_sync(FULL, ey, src);
                    const unsigned b = __ballot_sync(FULL, lane < n_nseg && ray_segment_candidate(nseg0, rox, roy, rex, rey));
                    cand = lane == src ? b : cand;
                }
                while (cand) {
                    const int k = __ffs(cand) - 1; cand &= cand - 1u;
                    const float s = ray_segment_exact<PHYS>(nseg[2 * k], ox, oy, ex, ey);
                    if (s < best_s) { best_s = s; best_k = k; }
                }
            } else {
                while (need) {
                    const int src = __ffs(need) - 1; need &= need - 1u;
                    const float rox = __shfl_sync(FULL, ox, src), roy = __shfl_sync(FULL, oy, src);
                    const float rex = __shfl_sync(FULL, ex, src), rey = __shfl_sync(FULL, ey, src);
                    float cs; int ck;
                    ray_cast_coop<false, PHYS>(nseg, n_nseg, nseg0, rox, roy, rex, rey, lane, cs, ck);
                    best_s = lane == src ? cs : best_s; best_k = lane == src ? ck : best_k;
                }
            }
        }
        if (stepping) {
            float m = NEU ? (dN < dD ? dN : dD) : dD;                                   // :212 / :215
            if (PHYS && DELTA) {
                if (a.maj.data) {                                                       // majorant of this step's ball
                    float M;
                    r = majorant_radius(a.maj, x, y, m, a.rmin, M);
                    pd_sb = fmaxf(M, 1e-8f / (r * r));                                  // sigma' = 0 around here: plain WoSt step
                } else {
                    m = m < a.phys_rcap ? m : a.phys_rcap;                              // keeps r sqrt(sigma_bar) <= 1
                    r = (m > a.rmin) ? m : a.rmin;
                }
            } else r = (m > a.rmin) ? m : a.rmin;
        }

        // ---- phase C: move, source sample, delta tracking ---------------------------------------------------------
        if (stepping) {
            float qx, qy;
            // physical + variable coefficients: volume sample y, its visibility and kernel ratio, Bessel terms of this ball
            float pd_yx = 0.f, pd_yy = 0.f, pd_ratio = 1.0f, pd_i0c = 1.0f, pd_k0c = 0.0f, pd_qc = 0.0f, pd_m1 = 0.0f;
            bool pd_vis = true;
            if (PHYS && (SRC || DELTA)) {
                // source sample (physical): independent direction, rho^2/r^2 ~ -ln  <=>  rho ~ 4 rho ln(r/rho)/r^2 (the 2D
                // disc Green's function), counted only if visible from x inside the star-shaped region
                const float th2 = (NEU && onB) ? phi_n + (u24(o[1]) - 0.5f) * 3.14159274101257324f : (u24(o[1]) * 2.0f) * 3.14159274101257324f;
                float s2, c2; sincosf(th2, &s2, &c2);
                const float u23 = u24p(o[2]) * u24p(o[3]);
                const float rho = r * sqrtf(u23);
                bool vis = true;
                if (NEU && gap <= rho && ray_may_hit_disc(x, y, c2, s2, a.ndisc_x, a.ndisc_y, a.ndisc_r2)) {
                    float vs; int vk;
                    if (BIG && a.nbvh.nodes) bvh_ray_cast<true>(a.nseg, n_nseg, a.nbvh, a.bvh_slack, x, y, c2, s2, vs, vk);
                    else ray_cast<true>(nseg, n_nseg, x, y, c2, s2, vs, vk);
                    vis = vk < 0 || vs > rho;
                }
                const float yx = x + rho * c2, yy = y + rho * s2;
                float wsrc = r * r / 4.0f;                                              // |G| of the Laplace ball kernel
                if (DELTA) {
                    // screened ball kernel G = ratio(rho) G_laplace; the source of the transformed equation is f / sqrt(alpha)
                    const float c = r * sqrtf(pd_sb);
                    float sc;
                    pd_qc = 0.25f * (c * c);
                    bessel_i0m1_s(pd_qc, pd_m1, sc);
                    pd_i0c = 1.0f + pd_m1;
                    pd_k0c = sc - (0.5f * logf(pd_qc) + EULER_GAMMA) * pd_i0c;
                    pd_ratio = phys_green_ratio(pd_i0c, sc, pd_qc * u23, -0.5f * logf(u23));
                    pd_yx = yx; pd_yy = yy; pd_vis = vis;
                    if (SRC && vis) wsrc = atten * ((pd_ratio * wsrc) / sqrtf(FP::alpha(a, yx, yy)));
                }
                float pc = 0.0f;
                if (SRC && FP::MULTI && a.n_src > 0) {
                    if (vis) {
                        float* row = a.walk_vals + (size_t)id * a.n_src;
                        for_each_source<FP>(a, yx, yy, [&](int k, float fk) {
                            const float ck = fk * wsrc;
                            if (ck != 0.0f) row[k] = row[k] + ck;
                        });
                    }
                } else if (SRC) {
                    pc = vis ? FP::f(a, yx, yy) * wsrc : 0.0f;
                    total_v += pc;
                }
                if (TRACE && SRC) {
                    if ((long long)id < a.n_trace && steps < a.trace_cap)
                        reinterpret_cast<float4*>(a.trace)[((size_t)id * (a.trace_cap + 1) + steps) * 2 + 1] = make_float4(yx, yy, pc, 0.0f);
                }
            }
            // physical + variable coefficients: null-collision inside the star with probability 1 - 1/I0(c), else walk on
            bool pd_vol = false;
            if (PHYS && DELTA) {
                uint32_t o2[4];
                philox4x32_10_ks(pidx, widx, (uint32_t)steps, 3u, a.ks, o2);     // stream tag 3: branch choice
                pd_vol = u24(o2[0]) < pd_m1 / pd_i0c;
            }
            if (NEU) {
                if (PHYS) {
                    // a wall within r + nudge counts as hit, so a free step ends at least `nudge` short of every wall
                    if (best_k < 0 || best_s > r + a.phys_nudge) { qx = x + r * ex; qy = y + r * ey; onB = false; }
                    else {
                        if (DELTA && !pd_vol) {                                         // flux of the screened kernel through the wall
                            const float ct = fminf(best_s, r) * sqrtf(pd_sb);
                            atten = atten * phys_wall_weight(pd_i0c, pd_k0c, 0.25f * (ct * ct));
                        }
                        // reflect: sit `nudge` off the wall on the side we came from, remember that side's normal
                        const float4 s1 = nseg[2 * best_k + 1];
                        float nx = s1.x, ny = s1.y;
                        if (nx * ex + ny * ey > 0.0f) { nx = -nx; ny = -ny; }
                        qx = (x + best_s * ex) + a.phys_nudge * nx; qy = (y + best_s * ey) + a.phys_nudge * ny; onB = true;
                        phi_n = atan2f(ny, nx);
                    }
                } else if (best_k < 0 || best_s > r || best_s <= 0.0f) {                // :166-174 miss
                    qx = x + r * ex; qy = y + r * ey; onB = false;
                } else {                                                                // :176-197 hit
                    qx = ox + best_s * ex; qy = oy + best_s * ey; onB = true;
                    phi_n = nseg[2 * best_k + 1].z;                                     // atan2 of the left normal (Q3)
                }
            } else {
                qx = x + r * dx; qy = y + r * dy; onB = false;                          // :238-239
            }
            if (TRACE) {
                if ((long long)id < a.n_trace && steps < a.trace_cap) {
                    float4* t = reinterpret_cast<float4*>(a.trace) + ((size_t)id * (a.trace_cap + 1) + steps) * 2;
                    t[0] = make_float4(x, y, dD, dN);
                }
            }

            if (PHYS && DELTA && pd_vol) {
                if (!pd_vis) { atten = 0.0f; qx = x; qy = y; }                          // the sample fell behind a wall
                else {
                    const float sp = sigma_prime_fp<FP>(a, pd_yx, pd_yy, -1.0f);
                    atten = (atten * (pd_ratio * (pd_qc * pd_i0c / pd_m1))) * (1.0f - sp / pd_sb);
                    qx = pd_yx; qy = pd_yy;
                }
                onB = false;
            }

            float sx = qx, sy = qy, gn = 0.0f, sbgn = 0.0f, alpha_s = 1.0f;
            bool have_alpha_s = false, edge = false;
            Jet jet_s; jet_s.v = 1.0f; jet_s.gx = jet_s.gy = jet_s.l = 0.0f;
            bool have_jet_s = false;
            if (DELTA && !PHYS) {
                sbgn = interior_probability(a.iprob, r * a.sqrt_sigma_bar);                      // sigma_bar * |G^sb|(r)
                gn = sbgn * a.inv_sigma_bar;                                            // screenedGreensNorm2D (utils.py:29-44)
                edge = u24(o[1]) > sbgn;                                                // :271-272 (the draw does not depend on the sample)
            }
            if (!PHYS && (SRC || DELTA)) {                                              // :242 (Q10: also without a source)
                float rho;
                if (DELTA) {                                                            // screened radius: inverse-CDF table (Q9)
                    const float pos = u24(o[2]) * (float)(a.icdf_len - 1);
                    int i = min((int)pos, a.icdf_len - 2);
                    const float fr = pos - (float)i;
                    const float t0 = __ldg(a.icdf + i), t1 = __ldg(a.icdf + i + 1);
                    rho = t0 + fr * (t1 - t0);
                } else {                                                                // pdf -ln(rho): product of two uniforms (Q8)
                    rho = fmaxf(u24p(o[2]) * u24p(o[3]), 1e-6f);
                }
                const float rs = rho * r;                                               // utils.py:117
                sx = x + rs * dx; sy = y + rs * dy;                                     // :245
                float contrib = 0.0f;
                // :248-250 compares |sample - x| > |next - x| (two rounded norms).  sqrt is monotone, so a smaller-or-equal
                // square decides "not greater", and a square larger by more than 2^-20 relative decides "greater" (the roots then
                // differ by more than an ulp); only the sliver in between needs the roots themselves.
                const float n2s = norm2_sq(sx - x, sy - y), n2q = norm2_sq(qx - x, qy - y);
                bool beyond = n2s > n2q;
                if (beyond && !(n2s > n2q * 1.000001f)) beyond = sqrtf(n2s) > sqrtf(n2q);
                if (beyond) { sx = qx; sy = qy; }
                // Delta tracking with the closed-form sigma' (WOST_SP_FULL): the walker's destination is the sample point on
                // interior steps (:281-284, where sigma' needs value, gradient and Laplacian of alpha) and next_point on edge
                // steps (:277, value only).  ONE evaluation site for all lanes: the jet at the destination.  Its value also serves
                // the source term of interior lanes; the few edge lanes evaluate alpha at their sample point separately.
                if (DELTA && FP::sp_mode(a) == WOST_SP_FULL && FP::has_alpha(a)) {
                    jet_s = FP::alpha_jet(a, edge ? qx : sx, edge ? qy : sy); have_jet_s = true;
                    if (!edge) { alpha_s = jet_s.v; have_alpha_s = true; }
                }
                if (!beyond && SRC) {
                    if (FP::MULTI && a.n_src > 0) {
                        // shared walks: the path does not depend on f, so one walk serves every source; each source
                        // accumulates exactly the sum a single-source solve would (same expressions, same order)
                        float* row = a.walk_vals + (size_t)id * a.n_src;
                        float den = 1.0f;
                        if (DELTA) {
                            if (!have_alpha_s) { alpha_s = FP::alpha(a, sx, sy); have_alpha_s = true; }
                            den = sqrtf(alpha_s * alpha_x);
                        }
                        const float w4 = r * r / 4.0f;
                        for_each_source<FP>(a, sx, sy, [&](int k, float fk) {
                            if (fk != 0.0f) row[k] = row[k] + (DELTA ? (fk * gn / den) * atten : fk * w4);   // fk != 0: plain division
                        });
                    } else if (DELTA) {                                                 // :252-254
                        if (!have_alpha_s) { alpha_s = FP::alpha(a, sx, sy); have_alpha_s = true; }
                        contrib = div_z(FP::f(a, sx, sy) * gn, FP::ALPHA_IN_RANGE ? sqrt_in_range(alpha_s * alpha_x) : sqrtf(alpha_s * alpha_x)) * atten;
                    } else
                        contrib = FP::f(a, sx, sy) * (r * r / 4.0f);       // :256
                }
                if (SRC) total_v += contrib;                                            // :258
                if (TRACE && SRC) {                                                     // :261-267 history: the source sample
                    if ((long long)id < a.n_trace && steps < a.trace_cap)
                        reinterpret_cast<float4*>(a.trace)[((size_t)id * (a.trace_cap + 1) + steps) * 2 + 1] = make_float4(sx, sy, contrib, 0.0f);
                }
            }
            if (DELTA && !PHYS) {                                                       // :271-284
                // alpha(current_point) is the value computed when the walker arrived here (same function, same point).
                // Both branches need alpha at their destination: evaluate it at ONE call site for all lanes (the edge
                // branch at next_point, the interior branch at sample_point unless the source term already did).
                const float tx = edge ? qx : sx, ty = edge ? qy : sy;
                const float alpha_t = have_jet_s ? jet_s.v : ((!edge && have_alpha_s) ? alpha_s : FP::alpha(a, tx, ty));
                // specialised kernels whose alpha is bounded away from 0 and infinity (2^-30 ... 2^30, known when the kernel is
                // generated): quotient and root by the compiler's own fast-path sequences without their range checks
                const float ratio = FP::ALPHA_IN_RANGE ? sqrt_in_range(div_in_range(alpha_t, alpha_x)) : sqrtf(alpha_t / alpha_x);
                if (edge) {
                    atten = atten * ratio;                                              // :277
                } else {
                    const float sp = sigma_prime_fp<FP>(a, sx, sy, alpha_t, have_jet_s ? &jet_s : nullptr);   // :281 (alpha(sample) is alpha_t here)
                    const float sc = fmaxf(1.0f - sp / a.sigma_bar, 0.0f);              // :282
                    atten = (atten * ratio) * sc;                                       // :283
                }
                x = tx; y = ty; alpha_x = alpha_t;                                      // :278,284
            } else { x = qx; y = qy; }                                                  // :287
            ++steps;                                                                    // :291
        }
    }
    // total step count: warp reduce, one atomic per warp
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) steps_acc += __shfl_xor_sync(FULL, steps_acc, off);
    if (lane == 0 && steps_acc) atomicAdd(a.steps_total, steps_acc);
}


}  // namespace wost
