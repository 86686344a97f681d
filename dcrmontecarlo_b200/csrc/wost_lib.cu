// wost_lib.cu — kernels and the C ABI (include/wost.h) of the B200 Walk-on-Stars engine.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
// (see dcrmontecarlo_b200/build.py).  -fmad=false is part of the numerical contract, see wost_device.cuh.
#include "wost_device.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

using namespace wost;

// =================================================================================================
// errors
// =================================================================================================
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(WOST_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
    } while (0)

// =================================================================================================
// scratch arena: device memory a scene keeps per stream for the temporaries of its calls (staged host buffers, per-walk
// totals, block statistics, counters).  Calls on one stream are ordered, so the next call may reuse the bytes of the
// previous one; steady-state calls allocate nothing.  (Round 1 used cudaMallocAsync / cudaFreeAsync per call and raised
// the release threshold of the device's DEFAULT memory pool -- a process-wide side effect, and ~10 driver calls per solve.)
// =================================================================================================
struct Arena {
    struct Block { char* p; size_t cap, off; };
    std::vector<Block> blocks;
    void reset() { for (auto& b : blocks) b.off = 0; }   // start of a call: everything is free again
    // Bytes come from the newest block only.  A call that outgrows it gets a block sized for the whole call next time (twice
    // what is held so far, at most 256 MiB extra); the outgrown blocks stay allocated until wost_scene_trim / _destroy --
    // freeing or merging them here would synchronise with the device (cudaFree) in the middle of a stream of solves: round 1's
    // merge on the NEXT call cost that call ~1 ms, which is what a sweep over growing job sizes then measured.
    void* take(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        if (!blocks.empty()) { Block& b = blocks.back(); if (b.off + bytes <= b.cap) { void* r = b.p + b.off; b.off += bytes; return r; } }
        size_t held = 0;
        for (auto& b : blocks) held += b.cap;
        const size_t cap = std::max(std::max(bytes + (bytes >> 2), std::min(2 * held, held + ((size_t)256 << 20))), (size_t)1 << 20);
        char* p = nullptr;
        if (cudaMalloc((void**)&p, cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        blocks.push_back({p, cap, bytes});
        return p;
    }
    size_t bytes() const { size_t t = 0; for (auto& b : blocks) t += b.cap; return t; }
    void release() { for (auto& b : blocks) cudaFree(b.p); blocks.clear(); }
};

// =================================================================================================
// handles
// =================================================================================================
struct wost_scene {
    int device = 0;
    int n_dseg = 0, n_nseg = 0;           // segments
    int n_dvtx = 0, n_nvtx = 0;
    float4* dseg = nullptr;               // 2 float4 per Dirichlet segment
    float4* nseg = nullptr;               // 2 float4 per Neumann segment
    int sm_count = 0;
    size_t smem_optin = 0;
    float ndisc_x = 0.f, ndisc_y = 0.f, ndisc_r = 0.f, ndisc_r2 = 0.f;   // inflated disc enclosing the Neumann polyline
    float4* dbvh = nullptr; int dbvh_leaves = 0;      // implicit BVH over the Dirichlet segments (large polylines only)
    float4* nbvh = nullptr; int nbvh_leaves = 0;      // same for the Neumann segments
    float4* ncones = nullptr;                          // silhouette cones of the Neumann hierarchy
    float4* nwide_boxes = nullptr; float4* nwide_cones = nullptr; WideBvh nwide{};   // 32-wide hierarchy (cooperative queries)
    float4* dwide_boxes = nullptr; WideBvh dwide{};                                   // same for the Dirichlet polyline (boxes only)
    float bvh_slack = 0.f;                             // ray/box slack (1e-4 of the scene scale)
    int neu_closed = 0;                                // first Neumann vertex == last
    int dir_rcp = 0;                                   // Dirichlet table carries verified reciprocals of u.u (dirichlet_distance<RCP>)
    float phys_nudge = 0.f;                            // 1e-5 of the scene scale
    float4* dseg_as_neu = nullptr;                     // the Dirichlet polyline in the Neumann layout and vice versa, for the
    float4* nseg_as_dir = nullptr;                     //   primitive entry points (intersect on `which = 0`, distance on `which = 1`)
    mutable std::mutex arena_mu;
    mutable std::map<cudaStream_t, Arena> arenas;      // per stream (a scene may be used from several streams at once)
    // shared-walk solves: the source grid (wost_walk.cuh, SourceGrid) of each source set used with this scene, by field ids
    mutable std::map<std::vector<uint64_t>, SourceGrid> source_grids;
    Arena* arena_for(cudaStream_t st) const {
        std::lock_guard<std::mutex> lk(arena_mu);
        Arena& a = arenas[st];
        a.reset();
        return &a;
    }
};

struct wost_field {
    int device = 0;
    float4 support = make_float4(0.f, 0.f, 3.0e38f, 0.f);   // (cx, cy, R^2): the field is EXACTLY zero outside this disc
    DevField d{};
    wost_term_t* terms = nullptr;
    float* grid = nullptr;
    std::vector<wost_term_t> h_terms;     // host copy of the device-form terms (code generation for the specialised kernels)
    uint64_t uid = 0;                     // unique per created field (fields are immutable): cache key of specialised kernels
};
static std::atomic<uint64_t> g_field_uid{1};

// =================================================================================================
// kernels: the walk (body in wost_walk.cuh; these are the statically compiled instantiations, fields by interpreter)
// =================================================================================================
#include "wost_walk.cuh"

template <bool NEU, bool SRC, bool DELTA, bool TRACE, bool PHYS, bool BIG>
__global__ void __launch_bounds__(256, 4) walk_kernel(const WalkArgs a) {
    walk_body<NEU, SRC, DELTA, TRACE, PHYS, BIG, InterpFP<NEU, SRC, DELTA>>(a);
}

// =================================================================================================
// kernels: deterministic statistics
// =================================================================================================
// One warp per (point, block of WOST_WALK_BLOCK walks): two-pass mean / M2 in fp64, fixed summation order.
// vals[(p * n_walks + w) * stride + blockIdx.y] (stride = number of sources of a shared-walk solve, else 1);
// stats[((blockIdx.y * n_pts_total + p_off + p) * nblk + b) * 2].
__global__ void __launch_bounds__(256) block_stats_kernel(const float* __restrict__ vals, long long n_pts, long long n_walks,
                                                          long long nblk, double* __restrict__ stats, int stride = 1,
                                                          long long n_pts_total = 0, long long p_off = 0) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_pts * nblk) return;
    const long long p = warp / nblk, b = warp - p * nblk;
    const long long w0 = b * WOST_WALK_BLOCK;
    const int n = (int)min((long long)WOST_WALK_BLOCK, n_walks - w0);
    const float* v0 = vals + (p * n_walks + w0) * stride + blockIdx.y;
    stats += 2 * ((long long)blockIdx.y * n_pts_total + p_off) * nblk;
    struct Strided { const float* p; int st; __device__ float operator[](int i) const { return p[(long long)i * st]; } } v{v0, stride};
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += (double)v[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const double mean = s / (double)n;
    double q = 0.0;
    for (int i = lane; i < n; i += 32) { const double d = (double)v[i] - mean; q += d * d; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    if (lane == 0) { stats[2 * warp] = mean; stats[2 * warp + 1] = q; }
}

// One thread per point: Chan merge of its blocks in block order.
__global__ void merge_stats_kernel(const double* __restrict__ stats, long long n_pts, long long n_walks, long long nblk,
                                   double* __restrict__ out_mean, double* __restrict__ out_m2) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts) return;
    double na = 0.0, ma = 0.0, qa = 0.0;
    for (long long b = 0; b < nblk; ++b) {
        const double nb = (double)min((long long)WOST_WALK_BLOCK, n_walks - b * WOST_WALK_BLOCK);
        const double mb = stats[2 * (p * nblk + b)], qb = stats[2 * (p * nblk + b) + 1];
        const double n = na + nb, delta = mb - ma;
        ma = ma + delta * (nb / n);
        qa = qa + qb + (delta * delta) * (na * nb / n);
        na = n;
    }
    if (out_mean) out_mean[p] = ma;
    if (out_m2) out_m2[p] = qa;
}

// =================================================================================================
// kernels: batched primitives (parity entry points) — the same device functions the walk calls
// =================================================================================================
__global__ void geom_distance_kernel(const float4* seg, int n, Bvh bvh, const float* p, long long B, float* out_d, int* out_seg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    int arg;
    const float d = bvh.nodes ? bvh_dirichlet_distance(seg, n, bvh, p[2 * i], p[2 * i + 1], &arg) : dirichlet_distance(seg, n, p[2 * i], p[2 * i + 1], &arg);
    if (out_d) out_d[i] = d;
    if (out_seg) out_seg[i] = arg;
}

// Dirichlet-layout table -> the generic (ax, ay, ux, uy) view used by the Neumann-style queries
struct SegView { const float4* seg; int n; int dirichlet_layout; };
__device__ __forceinline__ float4 seg_au(const SegView& s, int k) {
    if (!s.dirichlet_layout) return s.seg[2 * k];
    const float4 a = s.seg[2 * k], u = s.seg[2 * k + 1];
    return make_float4(a.x, a.y, u.x, u.y);
}

__global__ void geom_silhouette_kernel(SegView sv, Bvh bvh, const float* p, long long B, float* out_d, uint8_t* out_mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float px = p[2 * i], py = p[2 * i + 1];
    if (bvh.nodes && !out_mask && !sv.dirichlet_layout) {                // distance only: through the hierarchy
        if (out_d) out_d[i] = sqrtf(bvh_silhouette_distance_sq(sv.seg, sv.n, bvh, px, py, CUDART_INF_F));
        return;
    }
    float sil = CUDART_INF_F, prev_c = 0.0f;
    for (int k = 0; k < sv.n; ++k) {
        const float4 s0 = seg_au(sv, k);
        const float vx = px - s0.x, vy = py - s0.y;
        const float c = s0.z * vy - s0.w * vx;
        if (k > 0) {
            const bool is_sil = prev_c * c < 0.0f;
            if (is_sil) sil = fminf(sil, norm2_sq(vx, vy));
            if (out_mask) out_mask[i * (long long)(sv.n - 1) + (k - 1)] = is_sil ? 1 : 0;
        }
        prev_c = c;
    }
    if (out_d) out_d[i] = sqrtf(sil);
}

__global__ void geom_ray_kernel(SegView sv, const float* p, const float* dir, long long B, float* out_s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float ox = p[2 * i], oy = p[2 * i + 1], ex = dir[2 * i], ey = dir[2 * i + 1];
    for (int k = 0; k < sv.n; ++k) {
        out_s[i * (long long)sv.n + k] = ray_segment_s(seg_au(sv, k), ox, oy, ex, ey);
    }
}

__global__ void geom_intersect_kernel(const float4* nseg, int n, Bvh bvh, float slack, const float* p, const float* dir, const float* rr, long long B,
                                      float* out_pt, float* out_nrm, uint8_t* out_found, int* out_seg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float x = p[2 * i], y = p[2 * i + 1], dx = dir[2 * i], dy = dir[2 * i + 1], r = rr[i];
    float qx, qy, nx, ny; int found = 0, segk = -1;
    const float dn = norm2(dx, dy);
    if (dn < 1e-10f) { qx = x; qy = y; nx = 1.0f; ny = 0.0f; }                           // :150-154
    else {
        const float ex = dx / dn, ey = dy / dn;
        const float ox = x + 1e-6f * ex, oy = y + 1e-6f * ey;
        float best_s; int best_k;
        if (bvh.nodes) bvh_ray_cast(nseg, n, bvh, slack, ox, oy, ex, ey, best_s, best_k);
        else ray_cast(nseg, n, ox, oy, ex, ey, best_s, best_k);
        if (best_k < 0 || best_s > r || best_s <= 0.0f) { qx = x + r * ex; qy = y + r * ey; nx = ny = 0.0f; }
        else {
            qx = ox + best_s * ex; qy = oy + best_s * ey;
            const float4 s1 = nseg[2 * best_k + 1]; nx = s1.x; ny = s1.y; found = 1; segk = best_k;
        }
    }
    if (out_pt) { out_pt[2 * i] = qx; out_pt[2 * i + 1] = qy; }
    if (out_nrm) { out_nrm[2 * i] = nx; out_nrm[2 * i + 1] = ny; }
    if (out_found) out_found[i] = (uint8_t)found;
    if (out_seg) out_seg[i] = segk;
}

__global__ void field_eval_kernel(DevField F, const float* p, long long B, float* v, float* gx, float* gy, float* lap) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float x = p[2 * i], y = p[2 * i + 1];
    if (v) v[i] = field_eval(F, x, y);
    if (gx || gy || lap) {
        const Jet j = field_jet(F, x, y);
        if (gx) gx[i] = j.gx;
        if (gy) gy[i] = j.gy;
        if (lap) lap[i] = j.l;
    }
}

// alpha at the evaluation points, for the walk kernel's regeneration (there only a lane or two start a walk per iteration)
__global__ void alpha0_kernel(DevFields F, const float* p, long long B, float* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) out[i] = alpha_at(F, p[2 * i], p[2 * i + 1]);
}

__global__ void sigma_prime_kernel(DevFields F, int sp_mode, const float* p, long long B, float* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    out[i] = sigma_prime_at(F, sp_mode, p[2 * i], p[2 * i + 1]);
}

// FP32 FMA-chain microbenchmark (8 independent chains per thread)
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 12345.678f) out[0] = s;
}

// =================================================================================================
// host helpers
// =================================================================================================
static bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// A device view of a caller buffer: used in place if it already lives on the device, otherwise staged through a
// stream-ordered temporary (copied in for inputs, copied back for outputs on finish()).
template <typename T>
struct Staged {
    T* dev = nullptr; T* host = nullptr; size_t count = 0; bool temp = false; bool is_out = false; cudaStream_t st = nullptr;
    bool pooled = false;                                                // temporary from cudaMallocAsync (no arena given)
    int init(const T* p, size_t n, bool out, cudaStream_t s, Arena* ar = nullptr) {
        count = n; is_out = out; st = s;
        if (!p || n == 0) { dev = nullptr; return 0; }
        if (is_device_ptr(p)) { dev = const_cast<T*>(p); return 0; }
        host = const_cast<T*>(p); temp = true;
        if (ar) { dev = (T*)ar->take(n * sizeof(T)); if (!dev) return fail(WOST_ERR_ALLOC, "device scratch allocation failed"); }
        else { CU(cudaMallocAsync((void**)&dev, n * sizeof(T), s)); pooled = true; }
        if (!out) CU(cudaMemcpyAsync(dev, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
        return 0;
    }
    int finish() {
        if (temp && dev) {
            if (is_out) CU(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
            if (pooled) CU(cudaFreeAsync(dev, st));
            dev = nullptr;
        }
        return 0;
    }
    bool host_out() const { return temp && is_out; }
    ~Staged() { if (pooled && dev) cudaFreeAsync(dev, st); }           // error paths: finish() was not reached
    Staged() = default;
    Staged(const Staged&) = delete;
    Staged& operator=(const Staged&) = delete;
};

// stream-ordered scratch buffer, released on every exit path
template <typename T>
struct Scratch {
    T* p = nullptr; cudaStream_t st = nullptr; bool pooled = false;
    int alloc(size_t n, cudaStream_t s, Arena* ar = nullptr) {
        st = s;
        if (ar) { p = (T*)ar->take(n * sizeof(T)); if (!p) return fail(WOST_ERR_ALLOC, "device scratch allocation failed"); return 0; }
        CU(cudaMallocAsync((void**)&p, n * sizeof(T), s)); pooled = true; return 0;
    }
    int release() { if (p) { T* q = p; p = nullptr; if (pooled) CU(cudaFreeAsync(q, st)); } return 0; }
    ~Scratch() { if (p && pooled) cudaFreeAsync(p, st); }               // error paths
    Scratch() = default;
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
};

struct DeviceGuard {
    int prev = -1; bool ok = false;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Implicit BVH over the index order of a polyline (see wost_device.cuh).  `xy` are the nvtx vertices.
static std::vector<float4> build_bvh(const float* xy, int nvtx, float inflate, int* n_leaves_out) {
    const int nseg = nvtx - 1;
    int leaves = 1;
    while (leaves * WOST_BVH_LEAF < nseg) leaves *= 2;
    const float4 empty = make_float4(3e18f, 3e18f, 3e18f, 3e18f);       // a far-away point: never near, never hit
    std::vector<float4> nodes(2 * (size_t)leaves, empty);
    for (int j = 0; j < leaves; ++j) {
        const int s0 = j * WOST_BVH_LEAF, s1 = std::min(s0 + WOST_BVH_LEAF, nseg);
        if (s0 >= nseg) continue;
        float xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY;
        for (int v = s0; v <= s1; ++v) {                                 // segments s0..s1-1 touch vertices s0..s1
            xmin = std::fmin(xmin, xy[2 * v]); xmax = std::fmax(xmax, xy[2 * v]);
            ymin = std::fmin(ymin, xy[2 * v + 1]); ymax = std::fmax(ymax, xy[2 * v + 1]);
        }
        nodes[leaves + j] = make_float4(xmin - inflate, ymin - inflate, xmax + inflate, ymax + inflate);
    }
    for (int i = leaves - 1; i >= 1; --i) {
        const float4 a = nodes[2 * i], b = nodes[2 * i + 1];
        const bool a_empty = a.x == 3e18f, b_empty = b.x == 3e18f;
        if (a_empty && b_empty) nodes[i] = empty;
        else if (b_empty) nodes[i] = a;
        else if (a_empty) nodes[i] = b;
        else nodes[i] = make_float4(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmax(a.z, b.z), std::fmax(a.w, b.w));
    }
    *n_leaves_out = leaves;
    return nodes;
}

// Direction cones for the silhouette cull (see cone_excludes_silhouette): per node (axis x, axis y, sin h, R).
// Leaf j answers for vertices j*LEAF .. (j+1)*LEAF-1, whose tests involve segments j*LEAF-1 .. (j+1)*LEAF-1.
static std::vector<float4> build_cones(const float* xy, int nvtx, const std::vector<float4>& nodes, int leaves) {
    const int nseg = nvtx - 1;
    struct Cone { double ax, ay, h; bool empty; };
    std::vector<Cone> c(2 * (size_t)leaves, Cone{1.0, 0.0, 0.0, true});
    auto merge = [](const Cone& a, const Cone& b) {
        if (a.empty) return b;
        if (b.empty) return a;
        double sx = a.ax + b.ax, sy = a.ay + b.ay, n = std::sqrt(sx * sx + sy * sy);
        Cone r{1.0, 0.0, M_PI, false};
        if (n < 1e-9) return r;                                          // opposite axes: everything
        r.ax = sx / n; r.ay = sy / n;
        const double da = std::acos(std::fmax(-1.0, std::fmin(1.0, r.ax * a.ax + r.ay * a.ay)));
        const double db = std::acos(std::fmax(-1.0, std::fmin(1.0, r.ax * b.ax + r.ay * b.ay)));
        r.h = std::fmin(M_PI, std::fmax(da + a.h, db + b.h));
        return r;
    };
    for (int j = 0; j < leaves; ++j) {
        const int s0 = std::max(j * WOST_BVH_LEAF - 1, 0), s1 = std::min((j + 1) * WOST_BVH_LEAF, nseg);   // segments [s0, s1)
        Cone acc{1.0, 0.0, 0.0, true};
        for (int k = s0; k < s1; ++k) {
            const double ux = (double)xy[2 * k + 2] - xy[2 * k], uy = (double)xy[2 * k + 3] - xy[2 * k + 1], n = std::sqrt(ux * ux + uy * uy);
            acc = merge(acc, Cone{ux / n, uy / n, 0.0, false});
        }
        c[leaves + j] = acc;
    }
    for (int i = leaves - 1; i >= 1; --i) c[i] = merge(c[2 * i], c[2 * i + 1]);
    std::vector<float4> out(2 * (size_t)leaves);
    for (size_t i = 1; i < out.size(); ++i) {
        const float4 b = nodes[i];
        const double hx = 0.5 * ((double)b.z - b.x), hy = 0.5 * ((double)b.w - b.y);
        const double R = std::sqrt(hx * hx + hy * hy) * 1.0001;          // boxes are already inflated
        const bool usable = !c[i].empty && c[i].h < 0.5 * M_PI - 1e-3;
        out[i] = make_float4((float)c[i].ax, (float)c[i].ay, usable ? (float)std::sin(c[i].h + 1e-4) : 2.0f, (float)R);
    }
    return out;
}

// 32-wide hierarchy over the index order (see wost_device.cuh): level 0 = blocks of 32 segments, level l+1 groups 32
// level-l nodes, until at most 32 nodes remain.  Returns boxes and cones of all levels concatenated.
static void build_wide(const float* xy, int nvtx, float inflate, std::vector<float4>& boxes, std::vector<float4>& cones, WideBvh& w) {
    const int nseg = nvtx - 1;
    struct Cone { double ax, ay, h; bool empty; };
    auto merge = [](const Cone& a, const Cone& b) {
        if (a.empty) return b;
        if (b.empty) return a;
        double sx = a.ax + b.ax, sy = a.ay + b.ay, n = std::sqrt(sx * sx + sy * sy);
        Cone r{1.0, 0.0, M_PI, false};
        if (n < 1e-9) return r;
        r.ax = sx / n; r.ay = sy / n;
        const double da = std::acos(std::fmax(-1.0, std::fmin(1.0, r.ax * a.ax + r.ay * a.ay)));
        const double db = std::acos(std::fmax(-1.0, std::fmin(1.0, r.ax * b.ax + r.ay * b.ay)));
        r.h = std::fmin(M_PI, std::fmax(da + a.h, db + b.h));
        return r;
    };
    struct Node { double xmin, ymin, xmax, ymax; Cone c; };
    std::vector<std::vector<Node>> levels;
    {   // level 0
        const int nb = (nseg + 31) / 32;
        std::vector<Node> L(nb);
        for (int b = 0; b < nb; ++b) {
            const int s0 = b * 32, s1 = std::min(s0 + 32, nseg);
            Node nd{1e300, 1e300, -1e300, -1e300, Cone{1, 0, 0, true}};
            for (int v = s0; v <= s1; ++v) {
                nd.xmin = std::fmin(nd.xmin, xy[2 * v]); nd.xmax = std::fmax(nd.xmax, xy[2 * v]);
                nd.ymin = std::fmin(nd.ymin, xy[2 * v + 1]); nd.ymax = std::fmax(nd.ymax, xy[2 * v + 1]);
            }
            for (int k = std::max(s0 - 1, 0); k < s1; ++k) {             // vertex j's test involves segments j-1 and j
                const double ux = (double)xy[2 * k + 2] - xy[2 * k], uy = (double)xy[2 * k + 3] - xy[2 * k + 1], n = std::sqrt(ux * ux + uy * uy);
                nd.c = merge(nd.c, Cone{ux / n, uy / n, 0.0, false});
            }
            L[b] = nd;
        }
        levels.push_back(L);
    }
    while ((int)levels.back().size() > 32 && (int)levels.size() < WOST_WIDE_MAX_LEVELS) {
        const std::vector<Node>& P = levels.back();
        std::vector<Node> L((P.size() + 31) / 32);
        for (size_t g = 0; g < L.size(); ++g) {
            Node nd = P[g * 32];
            for (size_t k = g * 32 + 1; k < std::min(P.size(), g * 32 + 32); ++k) {
                nd.xmin = std::fmin(nd.xmin, P[k].xmin); nd.xmax = std::fmax(nd.xmax, P[k].xmax);
                nd.ymin = std::fmin(nd.ymin, P[k].ymin); nd.ymax = std::fmax(nd.ymax, P[k].ymax);
                nd.c = merge(nd.c, P[k].c);
            }
            L[g] = nd;
        }
        levels.push_back(L);
    }
    boxes.clear(); cones.clear();
    w.n_levels = (int)levels.size();
    for (int l = 0; l < w.n_levels; ++l) {
        w.off[l] = (int)boxes.size(); w.cnt[l] = (int)levels[l].size();
        for (const Node& nd : levels[l]) {
            const float4 b = make_float4((float)nd.xmin - inflate, (float)nd.ymin - inflate, (float)nd.xmax + inflate, (float)nd.ymax + inflate);
            const double hx = 0.5 * ((double)b.z - b.x), hy = 0.5 * ((double)b.w - b.y);
            const double R = std::sqrt(hx * hx + hy * hy) * 1.0001;
            const bool usable = !nd.c.empty && nd.c.h < 0.5 * M_PI - 1e-3;
            boxes.push_back(b);
            cones.push_back(make_float4((float)nd.c.ax, (float)nd.c.ay, usable ? (float)std::sin(nd.c.h + 1e-4) : 2.0f, (float)R));
        }
    }
}

// 1 - 1/I0 table of include/wost_math.h on a device (128 KB, built once per device, kept for the life of the process)
static const float* iprob_table(int device) {
    static std::mutex mu; static float* tab[64] = {nullptr};
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!tab[device]) {
        std::vector<float> h(WM_IPROB_N);
        wm_interior_probability_table(h.data());
        float* d = nullptr;
        if (cudaMalloc((void**)&d, sizeof(float) * WM_IPROB_N) != cudaSuccess ||
            cudaMemcpy(d, h.data(), sizeof(float) * WM_IPROB_N, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); cudaFree(d); return nullptr; }
        tab[device] = d;
    }
    return tab[device];
}

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

static inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// When to answer a Neumann query warp-cooperatively (at most this many lanes need it) instead of per lane.
// Instruction-count model: per-lane loop ~12 (silhouette) / ~20 (ray) per segment for the whole warp; cooperative
// ~14 / ~16 per 32-segment chunk plus ~10 per query.
static void coop_thresholds(int n, int* sil, int* ray) {
    const int chunks = (n + 31) / 32;
    *sil = n >= 8 ? (12 * n) / (14 * chunks + 10) : 0;
    *ray = n >= 8 ? (20 * n) / (16 * chunks + 14) : 0;
    if (*sil > 32) *sil = 32;
    if (*ray > 32) *ray = 32;
    // one segment per lane (8 ... 32 segments): always cooperative -- measured with the specialised kernels, +1 ... +1.6 % on
    // cfg 2 / cfg 4 over the model's 16 / 21, and the per-lane loops drop out of the generated kernel
    if (n >= 8 && n <= 32) { *sil = 32; *ray = 32; }
    *sil = env_int("WOST_SIL_COOP_MAX", *sil);
    *ray = env_int("WOST_RAY_COOP_MAX", *ray);
}

typedef void (*walk_kernel_t)(const WalkArgs);
template <bool TRACE, bool BIG>
static walk_kernel_t pick_kernel(bool neu, bool src, bool delta, bool phys) {
    const int m = (neu ? 4 : 0) | (src ? 2 : 0) | (delta ? 1 : 0);
    if (phys) {                                        // physical mode; delta = variable coefficients
        switch (m) {
            case 0: return walk_kernel<false, false, false, TRACE, true, BIG>;
            case 1: return walk_kernel<false, false, true, TRACE, true, BIG>;
            case 2: return walk_kernel<false, true, false, TRACE, true, BIG>;
            case 3: return walk_kernel<false, true, true, TRACE, true, BIG>;
            case 4: return walk_kernel<true, false, false, TRACE, true, BIG>;
            case 5: return walk_kernel<true, false, true, TRACE, true, BIG>;
            case 6: return walk_kernel<true, true, false, TRACE, true, BIG>;
            default: return walk_kernel<true, true, true, TRACE, true, BIG>;
        }
    }
    switch (m) {
        case 0: return walk_kernel<false, false, false, TRACE, false, BIG>;
        case 1: return walk_kernel<false, false, true, TRACE, false, BIG>;
        case 2: return walk_kernel<false, true, false, TRACE, false, BIG>;
        case 3: return walk_kernel<false, true, true, TRACE, false, BIG>;
        case 4: return walk_kernel<true, false, false, TRACE, false, BIG>;
        case 5: return walk_kernel<true, false, true, TRACE, false, BIG>;
        case 6: return walk_kernel<true, true, false, TRACE, false, BIG>;
        default: return walk_kernel<true, true, true, TRACE, false, BIG>;
    }
}

// the device copy of a smooth-circle term carries the squared radii beyond which the step is exactly 0 / 1 in its
// (otherwise unused) w1x / w1y slots: same fp32 expressions as the oracle evaluates per call
// Correctly rounded reciprocal y of b, and the proof by exhaustion that  q0 = a y; r = fma(-b, q0, a); q = fma(y, r, q0)
// equals the IEEE quotient a / b for every numerator (all 2^23 mantissas of one binade: scaling by powers of two is exact).
// Used for the divisions by the Dirichlet segments' u.u (dirichlet_distance<RCP>, wost_device.cuh).  ~30 ms per divisor.
static bool verified_reciprocal(float b, float* y_out) {
    static std::mutex mu;
    static std::map<uint32_t, std::pair<bool, float>> memo;
    uint32_t key; std::memcpy(&key, &b, 4);
    std::lock_guard<std::mutex> lk(mu);
    auto it = memo.find(key);
    if (it == memo.end()) {
        bool ok = b >= 9.094947e-13f && b <= 1.0995116e12f;               // [2^-40, 2^40]: nothing under- or overflows on the way
        float y = 0.0f;
        if (ok) {
            y = (float)(1.0 / (double)b);
            double err = std::fabs(1.0 - (double)b * (double)y);           // b y is exact in double: pick the nearest of the neighbours
            for (float c : {std::nextafterf(y, 0.0f), std::nextafterf(y, INFINITY)}) {
                const double e = std::fabs(1.0 - (double)b * (double)c);
                if (e < err) { err = e; y = c; }
            }
            for (uint32_t m = 0; m < (1u << 23) && ok; ++m) {
                const uint32_t bits = 0x3f800000u | m;
                float a; std::memcpy(&a, &bits, 4);
                volatile float q0 = a * y;
                const float r = std::fmaf(-b, q0, a);
                const float q = std::fmaf(y, r, q0);
                volatile float ref = a / b;
                ok = q == ref;
            }
        }
        it = memo.emplace(key, std::make_pair(ok, y)).first;
    }
    *y_out = it->second.second;
    return it->second.first;
}

// Brute-force check of the two hand-written division sequences against the compiler's IEEE division (tests, -m gpu):
// [0] div2_by_near_one on unit directions (cos t, sin t) / |(cos t, sin t)|, [1] the same on random a in [-2, 2], b in
// [1/2, 2), [2] the reciprocal sequence of dirichlet_distance<RCP> on the given divisors with random in-range numerators.
__global__ void division_selftest_kernel(long long n, uint32_t k0, uint32_t k1, const float* __restrict__ by, int n_div, unsigned long long* mism) {
    unsigned long long bad0 = 0, bad1 = 0, bad2 = 0, bad3 = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint32_t o[4];
        philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0u, 7u, k0, k1, o);
        {   // the walk's own use: a direction from the walk's sincos, normalised
            const float theta = (u24(o[0]) * 2.0f) * 3.14159274101257324f * ((o[1] & 1u) ? 0.5f : 1.0f) + ((o[1] & 2u) ? u24(o[2]) * 6.5f - 3.2f : 0.0f);
            float sn, cs; wm_sincosf_small(theta, &sn, &cs);
            const float dn = sqrt_in_range(norm2_sq(cs, sn));
            bad3 += __float_as_uint(dn) != __float_as_uint(norm2(cs, sn));
            float q0, q1; div2_by_near_one(cs, sn, dn, q0, q1);
            volatile float r0 = cs / dn, r1 = sn / dn;
            bad0 += (__float_as_uint(q0) != __float_as_uint(r0)) + (__float_as_uint(q1) != __float_as_uint(r1));
        }
        {   // random operands over the whole admitted range
            const float b = __uint_as_float(0x3f000000u | (o[3] & 0x00ffffffu));                       // [1/2, 2)
            const uint32_t ea = 67u + (o[1] >> 8) % 62u;                                               // 2^-60 ... 2^1
            const float a0 = __uint_as_float((o[0] & 0x80000000u) | (ea << 23) | (o[0] & 0x007fffffu));
            const float a1 = (o[2] & 0xffu) == 0u ? 0.0f : __uint_as_float((o[2] & 0x80000000u) | (((67u + (o[2] >> 8) % 62u)) << 23) | (o[1] & 0x007fffffu));
            float q0, q1; div2_by_near_one(a0, a1, b, q0, q1);
            volatile float r0 = a0 / b, r1 = a1 / b;
            bad1 += (__float_as_uint(q0) != __float_as_uint(r0)) + (__float_as_uint(q1 + 0.0f) != __float_as_uint(r1 + 0.0f));
        }
        {   // square roots over [2^-100, 2^127]
            const float x = __uint_as_float(((27u + (o[3] >> 8) % 228u) << 23) | (o[0] & 0x007fffffu));
            volatile float r = sqrtf(x);
            bad3 += __float_as_uint(sqrt_in_range(x)) != __float_as_uint(r);
        }
        if (n_div > 0) {
            const int k = (int)(o[3] >> 24) % n_div;
            const float b = by[k], y = by[n_div + k];
            const uint32_t ea = 67u + (o[2] >> 8) % 120u;                                              // 2^-60 ... 2^59
            const float a = __uint_as_float((o[1] & 0x80000000u) | (ea << 23) | (o[2] & 0x007fffffu));
            const float p = a * y;
            const float q = fmaf(y, fmaf(-b, p, a), p);
            volatile float r = a / b;
            bad2 += __float_as_uint(q) != __float_as_uint(r);
        }
    }
    if (bad0) atomicAdd(mism, bad0);
    if (bad1) atomicAdd(mism + 1, bad1);
    if (bad2) atomicAdd(mism + 2, bad2);
    if (bad3) atomicAdd(mism + 3, bad3);
}

static std::vector<wost_term_t> device_form_terms(const wost_field_desc_t* d) {
    std::vector<wost_term_t> terms(d->terms, d->terms + (d->kind == WOST_FIELD_TERMS ? d->n_terms : 0));
    for (auto& t : terms)
        if (t.kind == WOST_TERM_SIGMOID_CIRCLE) {
            const float ro = t.R + 88.0f / t.q, ri = t.R - 18.0f / t.q;
            const float ro2 = ro * ro, ri2 = ri * ri;
            t.w1x = ro2 * 1.0001f;
            t.w1y = ri > 0.0f ? ri2 * 0.9999f : -1.0f;
        }
    return terms;
}

#include "wost_jit.inc"

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int wost_version(void) { return WOST_VERSION; }
const char* wost_last_error(void) { return g_err.c_str(); }

int wost_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int wost_scene_create(const float* dxy, int32_t nd, const float* nxy, int32_t nn, int32_t device, wost_scene_t** out) {
    if (!out) return fail(WOST_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!dxy || nd < 2) return fail(WOST_ERR_INVALID, "Dirichlet polyline needs at least 2 vertices");
    if (nn == 1 || nn < 0 || (nn > 0 && !nxy)) return fail(WOST_ERR_INVALID, "Neumann polyline needs 0 or >= 2 vertices");
    if (is_device_ptr(dxy) || is_device_ptr(nxy)) return fail(WOST_ERR_INVALID, "scene vertices must be host pointers");
    if (wost_device_count() <= 0) return fail(WOST_ERR_CUDA, "no CUDA device available (libwost has no CPU fallback)");
    DeviceGuard g(device);
    if (!g.ok) return fail(WOST_ERR_CUDA, "cannot select CUDA device " + std::to_string(device));
    // volatile: keep the host compiler from contracting these into FMAs — the tables must hold exactly the
    // fp32 values torch computes for u = b - a and u.u (geometry/PolylinesSimple.py:37,42)
    std::vector<float4> ds(2 * (size_t)(nd - 1)), ns(nn ? 2 * (size_t)(nn - 1) : 0);
    for (int k = 0; k + 1 < nd; ++k) {
        volatile float ax = dxy[2 * k], ay = dxy[2 * k + 1], bx = dxy[2 * k + 2], by = dxy[2 * k + 3];
        volatile float ux = bx - ax, uy = by - ay;
        volatile float xx = ux * ux, yy = uy * uy;
        volatile float uu = xx + yy;
        if (!(uu > 0.0f) || !std::isfinite(uu))
            return fail(WOST_ERR_INVALID, "Dirichlet segment " + std::to_string(k) + " has zero or non-finite length");
        ds[2 * k] = make_float4(ax, ay, bx, by);
        ds[2 * k + 1] = make_float4(ux, uy, uu, 0.0f);
    }
    for (int k = 0; k + 1 < nn; ++k) {
        volatile float ax = nxy[2 * k], ay = nxy[2 * k + 1], bx = nxy[2 * k + 2], by = nxy[2 * k + 3];
        volatile float ux = bx - ax, uy = by - ay;
        volatile float xx = ux * ux;
        const float len = sqrtf(fmaf(uy, uy, xx));                       // torch.norm (:184)
        if (!(len > 0.0f) || !std::isfinite(len))
            return fail(WOST_ERR_INVALID, "Neumann segment " + std::to_string(k) + " has zero or non-finite length");
        float nx, ny;
        if (len < 1e-10f) { nx = 0.0f; ny = 1.0f; }                      // :186-189
        else { volatile float tx = ux / len, ty = uy / len; nx = -ty; ny = tx; }   // :191-194
        ns[2 * k] = make_float4(ax, ay, ux, uy);
        ns[2 * k + 1] = make_float4(nx, ny, atan2f(ny, nx), 0.0f);       // solvers/WoStSolver.py:228
    }
    // reciprocals of u.u for the scene-specialised kernels' Dirichlet distance -- only if every segment's divisor verifies
    // (polylines of at most 64 segments with at most 8 distinct lengths: squares, regular polygons, surface lines)
    int dir_rcp = 0;
    if (nd - 1 <= 64 && env_int("WOST_DIRICHLET_RCP", 1)) {
        std::map<uint32_t, float> ys;
        bool ok = true;
        for (int k = 0; k + 1 < nd && ok; ++k) {
            uint32_t key; std::memcpy(&key, &ds[2 * k + 1].z, 4);
            if (!ys.count(key)) {
                float y = 0.0f;
                ok = ys.size() < 8 && verified_reciprocal(ds[2 * k + 1].z, &y);
                ys[key] = y;
            }
        }
        if (ok) {
            for (int k = 0; k + 1 < nd; ++k) { uint32_t key; std::memcpy(&key, &ds[2 * k + 1].z, 4); ds[2 * k + 1].w = ys[key]; }
            dir_rcp = 1;
        }
    }
    // the other layout of each polyline, for the primitive entry points (PolyLinesSimple methods work on either boundary)
    std::vector<float4> ds_n(ds.size()), ns_d(ns.size());
    for (int k = 0; k + 1 < nd; ++k) {
        const float ax = ds[2 * k].x, ay = ds[2 * k].y; volatile float ux = ds[2 * k + 1].x, uy = ds[2 * k + 1].y;
        volatile float xx = ux * ux;
        const float len = sqrtf(fmaf(uy, uy, xx));
        volatile float tx = ux / len, ty = uy / len;
        ds_n[2 * k] = make_float4(ax, ay, ux, uy); ds_n[2 * k + 1] = make_float4(-ty, tx, atan2f(tx, -ty), 0.0f);
    }
    for (int k = 0; k + 1 < nn; ++k) {
        volatile float ax = ns[2 * k].x, ay = ns[2 * k].y, ux = ns[2 * k].z, uy = ns[2 * k].w;
        volatile float xx = ux * ux, yy = uy * uy; volatile float uu = xx + yy;
        ns_d[2 * k] = make_float4(ax, ay, nxy[2 * k + 2], nxy[2 * k + 3]); ns_d[2 * k + 1] = make_float4(ux, uy, uu, 0.0f);
    }
    auto* s = new wost_scene();
    s->device = device; s->n_dvtx = nd; s->n_nvtx = nn; s->n_dseg = nd - 1; s->n_nseg = nn ? nn - 1 : 0;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete s; return fail(WOST_ERR_CUDA, "cudaGetDeviceProperties failed"); }
    s->sm_count = prop.multiProcessorCount; s->smem_optin = prop.sharedMemPerBlockOptin;
    cudaError_t e = cudaMalloc((void**)&s->dseg, ds.size() * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemcpy(s->dseg, ds.data(), ds.size() * sizeof(float4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !ns.empty()) {
        e = cudaMalloc((void**)&s->nseg, ns.size() * sizeof(float4));
        if (e == cudaSuccess) e = cudaMemcpy(s->nseg, ns.data(), ns.size() * sizeof(float4), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->dseg_as_neu, ds_n.size() * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemcpy(s->dseg_as_neu, ds_n.data(), ds_n.size() * sizeof(float4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !ns_d.empty()) {
        e = cudaMalloc((void**)&s->nseg_as_dir, ns_d.size() * sizeof(float4));
        if (e == cudaSuccess) e = cudaMemcpy(s->nseg_as_dir, ns_d.data(), ns_d.size() * sizeof(float4), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        cudaFree(s->dseg); cudaFree(s->nseg); cudaFree(s->dseg_as_neu); cudaFree(s->nseg_as_dir); delete s;
        return fail(WOST_ERR_CUDA, std::string("scene upload: ") + cudaGetErrorString(e));
    }
    if (nn > 0) {
        double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
        for (int k = 0; k < nn; ++k) {
            xmin = std::fmin(xmin, nxy[2 * k]); xmax = std::fmax(xmax, nxy[2 * k]);
            ymin = std::fmin(ymin, nxy[2 * k + 1]); ymax = std::fmax(ymax, nxy[2 * k + 1]);
        }
        const double cx = 0.5 * (xmin + xmax), cy = 0.5 * (ymin + ymax);
        double r2 = 0.0;
        for (int k = 0; k < nn; ++k) r2 = std::fmax(r2, (nxy[2 * k] - cx) * (nxy[2 * k] - cx) + (nxy[2 * k + 1] - cy) * (nxy[2 * k + 1] - cy));
        const double scale = std::fmax(std::fmax(std::fabs(xmin), std::fabs(xmax)), std::fmax(std::fabs(ymin), std::fabs(ymax)));
        const double R = std::sqrt(r2) * 1.001 + 1e-4 * scale + 1e-30;   // slack: fp32 rounding of s, t and of the cull itself
        s->ndisc_x = (float)cx; s->ndisc_y = (float)cy; s->ndisc_r = (float)R; s->ndisc_r2 = (float)(R * R);
    }
    {   // hierarchies for large polylines (thresholds overridable for experiments)
        double scale = 0.0;
        for (int k = 0; k < 2 * nd; ++k) scale = std::fmax(scale, std::fabs((double)dxy[k]));
        for (int k = 0; k < 2 * nn; ++k) scale = std::fmax(scale, std::fabs((double)nxy[k]));
        const float inflate = (float)(1e-5 * scale + 1e-30);
        s->bvh_slack = (float)(1e-4 * scale + 1e-30);
        s->phys_nudge = (float)(1e-5 * scale);
        s->neu_closed = (nn >= 4 && nxy[0] == nxy[2 * nn - 2] && nxy[1] == nxy[2 * nn - 1]) ? 1 : 0;
        s->dir_rcp = dir_rcp;
        cudaError_t be = cudaSuccess;
        if (s->n_dseg >= env_int("WOST_BVH_MIN_DIRICHLET", 48)) {
            const std::vector<float4> nodes = build_bvh(dxy, nd, inflate, &s->dbvh_leaves);
            be = cudaMalloc((void**)&s->dbvh, nodes.size() * sizeof(float4));
            if (be == cudaSuccess) be = cudaMemcpy(s->dbvh, nodes.data(), nodes.size() * sizeof(float4), cudaMemcpyHostToDevice);
            std::vector<float4> wb, wc;
            build_wide(dxy, nd, inflate, wb, wc, s->dwide);
            if (s->n_dseg >= env_int("WOST_WIDE_MIN_DIRICHLET", 512) && s->dwide.cnt[s->dwide.n_levels - 1] <= 32) {
                if (be == cudaSuccess) be = cudaMalloc((void**)&s->dwide_boxes, wb.size() * sizeof(float4));
                if (be == cudaSuccess) be = cudaMemcpy(s->dwide_boxes, wb.data(), wb.size() * sizeof(float4), cudaMemcpyHostToDevice);
                s->dwide.boxes = s->dwide_boxes;
            }
        }
        if (be == cudaSuccess && s->n_nseg >= env_int("WOST_BVH_MIN_NEUMANN", 192)) {
            const std::vector<float4> nodes = build_bvh(nxy, nn, inflate, &s->nbvh_leaves);
            be = cudaMalloc((void**)&s->nbvh, nodes.size() * sizeof(float4));
            if (be == cudaSuccess) be = cudaMemcpy(s->nbvh, nodes.data(), nodes.size() * sizeof(float4), cudaMemcpyHostToDevice);
            const std::vector<float4> cones = build_cones(nxy, nn, nodes, s->nbvh_leaves);
            if (be == cudaSuccess) be = cudaMalloc((void**)&s->ncones, cones.size() * sizeof(float4));
            if (be == cudaSuccess) be = cudaMemcpy(s->ncones, cones.data(), cones.size() * sizeof(float4), cudaMemcpyHostToDevice);
            std::vector<float4> wb, wc;
            build_wide(nxy, nn, inflate, wb, wc, s->nwide);
            if ((int)wb.size() > 0 && s->nwide.cnt[s->nwide.n_levels - 1] <= 32) {
                if (be == cudaSuccess) be = cudaMalloc((void**)&s->nwide_boxes, wb.size() * sizeof(float4));
                if (be == cudaSuccess) be = cudaMemcpy(s->nwide_boxes, wb.data(), wb.size() * sizeof(float4), cudaMemcpyHostToDevice);
                if (be == cudaSuccess) be = cudaMalloc((void**)&s->nwide_cones, wc.size() * sizeof(float4));
                if (be == cudaSuccess) be = cudaMemcpy(s->nwide_cones, wc.data(), wc.size() * sizeof(float4), cudaMemcpyHostToDevice);
                s->nwide.boxes = s->nwide_boxes; s->nwide.cones = s->nwide_cones;
            }
        }
        if (be != cudaSuccess) {
            cudaFree(s->dseg); cudaFree(s->nseg); cudaFree(s->dbvh); cudaFree(s->nbvh); cudaFree(s->ncones); cudaFree(s->nwide_boxes); cudaFree(s->nwide_cones); cudaFree(s->dwide_boxes); delete s;
            return fail(WOST_ERR_CUDA, std::string("BVH upload: ") + cudaGetErrorString(be));
        }
    }
    *out = s;
    return WOST_OK;
}

int wost_scene_destroy(wost_scene_t* s) {
    if (!s) return WOST_OK;
    DeviceGuard g(s->device);
    cudaFree(s->dseg); cudaFree(s->nseg); cudaFree(s->dbvh); cudaFree(s->nbvh); cudaFree(s->ncones);
    cudaFree(s->nwide_boxes); cudaFree(s->nwide_cones); cudaFree(s->dwide_boxes);
    cudaFree(s->dseg_as_neu); cudaFree(s->nseg_as_dir);
    for (auto& kv : s->arenas) kv.second.release();
    for (auto& kv : s->source_grids) cudaFree((void*)kv.second.masks);
    delete s;
    return WOST_OK;
}

int wost_scene_trim(wost_scene_t* s, int64_t* out_released_bytes) {
    if (!s) return fail(WOST_ERR_INVALID, "scene is NULL");
    DeviceGuard g(s->device);
    std::lock_guard<std::mutex> lk(s->arena_mu);
    size_t tot = 0;
    CU(cudaDeviceSynchronize());                                         // nothing may still be using the scratch
    for (auto& kv : s->arenas) { tot += kv.second.bytes(); kv.second.release(); }
    s->arenas.clear();
    if (out_released_bytes) *out_released_bytes = (int64_t)tot;
    return WOST_OK;
}

int wost_field_create(const wost_field_desc_t* d, int32_t device, wost_field_t** out) {
    if (!out) return fail(WOST_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!d) return fail(WOST_ERR_INVALID, "desc is NULL");
    if (d->kind != WOST_FIELD_TERMS && d->kind != WOST_FIELD_GRID) return fail(WOST_ERR_INVALID, "unknown field kind");
    if (d->kind == WOST_FIELD_TERMS && (d->n_terms < 0 || (d->n_terms > 0 && !d->terms))) return fail(WOST_ERR_INVALID, "bad term list");
    if (d->kind == WOST_FIELD_GRID && (d->nx < 2 || d->ny < 2 || !d->grid || !(d->dx > 0.0f) || !(d->dy > 0.0f)))
        return fail(WOST_ERR_INVALID, "grid field needs nx,ny >= 2, positive spacing and data");
    if (d->mask_kind < WOST_MASK_NONE || d->mask_kind > WOST_MASK_DISC) return fail(WOST_ERR_INVALID, "unknown mask kind");
    for (int k = 0; d->kind == WOST_FIELD_TERMS && k < d->n_terms; ++k) {
        const wost_term_t& t = d->terms[k];
        if (t.kind != WOST_TERM_PRODUCT && t.kind != WOST_TERM_SIGMOID_CIRCLE) return fail(WOST_ERR_INVALID, "unknown term kind");
        if (t.px < 0 || t.py < 0 || t.px > 16 || t.py > 16) return fail(WOST_ERR_INVALID, "monomial power out of range [0,16]");
        if (t.t1 < 0 || t.t1 > 2 || t.t2 < 0 || t.t2 > 2) return fail(WOST_ERR_INVALID, "unknown trig kind");
    }
    if (wost_device_count() <= 0) return fail(WOST_ERR_CUDA, "no CUDA device available (libwost has no CPU fallback)");
    DeviceGuard g(device);
    if (!g.ok) return fail(WOST_ERR_CUDA, "cannot select CUDA device " + std::to_string(device));
    auto* f = new wost_field();
    f->device = device; f->uid = g_field_uid.fetch_add(1);
    DevField& D = f->d;
    D.present = 1; D.kind = d->kind; D.n_terms = d->kind == WOST_FIELD_TERMS ? d->n_terms : 0; D.mask_kind = d->mask_kind;
    D.c0 = d->c0; D.m0 = d->mask[0]; D.m1 = d->mask[1]; D.m2 = d->mask[2]; D.m3 = d->mask[3]; D.outside = d->outside;
    D.nx = d->nx; D.ny = d->ny; D.x0 = d->x0; D.y0 = d->y0; D.dx = d->dx; D.dy = d->dy;
    // compact support: a sum of Gaussian terms (no constant, no value outside a mask) evaluates to exactly zero where every
    // term takes its underflow shortcut (exp argument < -110, wost_device.cuh); used to skip far sources in shared-walk solves
    if (d->kind == WOST_FIELD_TERMS && d->c0 == 0.0f && D.n_terms > 0 && (d->mask_kind == WOST_MASK_NONE || d->outside == 0.0f)) {
        bool compact = true; double mx = 0.0, my = 0.0;
        for (int k = 0; k < D.n_terms; ++k) {
            const wost_term_t& t = d->terms[k];
            if (t.kind != WOST_TERM_PRODUCT || !(t.q > 0.0f)) { compact = false; break; }
            mx += t.cx; my += t.cy;
        }
        if (compact) {
            mx /= D.n_terms; my /= D.n_terms;
            double R = 0.0;
            for (int k = 0; k < D.n_terms; ++k) {
                const wost_term_t& t = d->terms[k];
                R = std::fmax(R, std::hypot(t.cx - mx, t.cy - my) + std::sqrt(112.0 / t.q));   // a little beyond e = -110
            }
            R = R * 1.001 + 1e-6;
            f->support = make_float4((float)mx, (float)my, (float)(R * R), 0.0f);
        }
    }
    cudaError_t e = cudaSuccess;
    if (D.n_terms > 0) {
        e = cudaMalloc((void**)&f->terms, sizeof(wost_term_t) * D.n_terms);
        const std::vector<wost_term_t> terms = device_form_terms(d);
        if (e == cudaSuccess) e = cudaMemcpy(f->terms, terms.data(), sizeof(wost_term_t) * D.n_terms, cudaMemcpyHostToDevice);
        f->h_terms = terms;
    }
    if (e == cudaSuccess && d->kind == WOST_FIELD_GRID) {
        const size_t n = (size_t)d->nx * d->ny;
        e = cudaMalloc((void**)&f->grid, sizeof(float) * n);
        if (e == cudaSuccess) e = cudaMemcpy(f->grid, d->grid, sizeof(float) * n, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        cudaFree(f->terms); cudaFree(f->grid); delete f;
        return fail(WOST_ERR_CUDA, std::string("field upload: ") + cudaGetErrorString(e));
    }
    D.terms = f->terms; D.grid = f->grid;
    *out = f;
    return WOST_OK;
}

int wost_field_destroy(wost_field_t* f) {
    if (!f) return WOST_OK;
    DeviceGuard g(f->device);
    cudaFree(f->terms); cudaFree(f->grid);
    delete f;
    return WOST_OK;
}

static DevField dev_field_of(const wost_field_t* f) { DevField z{}; return f ? f->d : z; }
static DevFields dev_fields_of(const wost_fields_t* F) {
    DevFields D{};
    if (F) { D.g = dev_field_of(F->g); D.f = dev_field_of(F->f); D.alpha = dev_field_of(F->alpha); D.sigma = dev_field_of(F->sigma); D.sigma_prime = dev_field_of(F->sigma_prime); }
    return D;
}

int wost_field_eval(const wost_field_t* f, const float* p, int64_t B, float* v, float* gx, float* gy, float* lap, void* stream) {
    if (!f || !p || B < 0) return fail(WOST_ERR_INVALID, "bad arguments");
    if (B == 0) return WOST_OK;
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    Staged<float> sp, sv, sgx, sgy, sl;
    int rc;
    if ((rc = sp.init(p, 2 * B, false, st)) || (rc = sv.init(v, B, true, st)) || (rc = sgx.init(gx, B, true, st)) ||
        (rc = sgy.init(gy, B, true, st)) || (rc = sl.init(lap, B, true, st))) return rc;
    field_eval_kernel<<<blocks_for(B, 256), 256, 0, st>>>(f->d, sp.dev, B, sv.dev, sgx.dev, sgy.dev, sl.dev);
    CU(cudaGetLastError());
    const bool sync = sv.host_out() || sgx.host_out() || sgy.host_out() || sl.host_out();
    if ((rc = sp.finish()) || (rc = sv.finish()) || (rc = sgx.finish()) || (rc = sgy.finish()) || (rc = sl.finish())) return rc;
    if (sync) CU(cudaStreamSynchronize(st));
    return WOST_OK;
}

int wost_sigma_prime_eval(const wost_fields_t* F, int32_t sp_mode, const float* p, int64_t B, float* out, void* stream) {
    if (!F || !p || !out || B < 0) return fail(WOST_ERR_INVALID, "bad arguments");
    if (sp_mode == WOST_SP_FIELD && !F->sigma_prime) return fail(WOST_ERR_INVALID, "WOST_SP_FIELD needs fields.sigma_prime");
    const wost_field_t* any = F->alpha ? F->alpha : (F->sigma ? F->sigma : F->sigma_prime);
    if (!any) return fail(WOST_ERR_INVALID, "no coefficient field given");
    if (B == 0) return WOST_OK;
    DeviceGuard g(any->device);
    cudaStream_t st = (cudaStream_t)stream;
    Staged<float> sp, so; int rc;
    if ((rc = sp.init(p, 2 * B, false, st)) || (rc = so.init(out, B, true, st))) return rc;
    sigma_prime_kernel<<<blocks_for(B, 256), 256, 0, st>>>(dev_fields_of(F), sp_mode, sp.dev, B, so.dev);
    CU(cudaGetLastError());
    const bool sync = so.host_out();
    if ((rc = sp.finish()) || (rc = so.finish())) return rc;
    if (sync) CU(cudaStreamSynchronize(st));
    return WOST_OK;
}

}  // extern "C"

// Shared implementation of wost_solve (n_sources == 0: the source is fields->f) and wost_solve_multi_source.
// Source grid of a shared-walk solve (SourceGrid, wost_device.cuh / for_each_source in wost_walk.cuh): built once per
// (scene, source set) and kept with the scene.  Returns a grid with masks == nullptr when binning cannot help (no source
// with compact support).
static int source_grid_for(const wost_scene_t* scene, const wost_field_t* const* sources, int n, SourceGrid* out) {
    SourceGrid none{}; *out = none;
    std::vector<uint64_t> key(n);
    for (int k = 0; k < n; ++k) key[k] = sources[k]->uid;
    std::lock_guard<std::mutex> lk(scene->arena_mu);
    auto it = scene->source_grids.find(key);
    if (it != scene->source_grids.end()) { *out = it->second; return WOST_OK; }
    if (scene->source_grids.size() >= 64) return WOST_OK;               // bounded (grids live as long as the scene): test every disc instead
    // discs outside which each term is exactly zero (exp argument below -110: wost_device.cuh, term_value); a source without
    // compact support is listed everywhere.  Blob mode: every source is a plain sum of bare Gaussians.
    struct Disc { int bit; double x, y, r; };
    std::vector<Disc> discs; std::vector<char> everywhere(n, 0);
    std::vector<float4> blobs; std::vector<int> blob_src;
    bool blob_mode = env_int("WOST_SOURCE_BLOBS", 1) != 0;
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int k = 0; k < n; ++k) {
        const wost_field* f = sources[k];
        if (!(f->support.z < 1.0e38f)) { everywhere[k] = 1; blob_mode = false; continue; }
        if (f->d.mask_kind != WOST_MASK_NONE || f->d.c0 != 0.0f) blob_mode = false;
        for (const wost_term_t& t : f->h_terms) {
            if (t.kind != WOST_TERM_PRODUCT || t.px || t.py || t.t1 != WOST_TRIG_NONE || t.t2 != WOST_TRIG_NONE) blob_mode = false;
            const double r = std::sqrt(112.0 / (double)t.q) * 1.001 + 1e-6;
            discs.push_back({k, t.cx, t.cy, r});
            blobs.push_back(make_float4(t.A, t.q, t.cx, t.cy)); blob_src.push_back(k);
            x0 = std::min(x0, t.cx - r); x1 = std::max(x1, t.cx + r); y0 = std::min(y0, t.cy - r); y1 = std::max(y1, t.cy + r);
        }
    }
    if (discs.empty()) { scene->source_grids[key] = none; return WOST_OK; }
    if (blob_mode) for (size_t b = 0; b < discs.size(); ++b) discs[b].bit = (int)b;   // one bit per blob instead of per source
    const int n_bits = blob_mode ? (int)discs.size() : n;
    SourceGrid g{};
    g.nx = 64; g.ny = 64; g.words = (n_bits + 63) / 64;
    const double dx = std::max((x1 - x0) / g.nx, 1e-30), dy = std::max((y1 - y0) / g.ny, 1e-30);
    g.x0 = (float)x0; g.y0 = (float)y0; g.inv_dx = (float)(1.0 / dx); g.inv_dy = (float)(1.0 / dy);
    std::vector<unsigned long long> m(((size_t)g.nx * g.ny + 1) * g.words, 0ull);
    auto set = [&](size_t cell, int bit) { m[cell * g.words + bit / 64] |= 1ull << (bit % 64); };
    if (!blob_mode) for (int k = 0; k < n; ++k) if (everywhere[k]) for (size_t c = 0; c <= (size_t)g.nx * g.ny; ++c) set(c, k);
    for (const Disc& d : discs) {                                       // the disc's box plus one cell all round (fp32 cell lookup)
        const int i0 = std::max(0, (int)std::floor((d.x - d.r - x0) / dx) - 1), i1 = std::min(g.nx - 1, (int)std::floor((d.x + d.r - x0) / dx) + 1);
        const int j0 = std::max(0, (int)std::floor((d.y - d.r - y0) / dy) - 1), j1 = std::min(g.ny - 1, (int)std::floor((d.y + d.r - y0) / dy) + 1);
        for (int j = j0; j <= j1; ++j)
            for (int i = i0; i <= i1; ++i) {
                // cell rectangle grown by one cell on every side against the disc itself (not its box: a fifth fewer entries)
                const double cx0 = x0 + (i - 1) * dx, cx1 = x0 + (i + 2) * dx, cy0 = y0 + (j - 1) * dy, cy1 = y0 + (j + 2) * dy;
                const double qx = std::min(std::max(d.x, cx0), cx1) - d.x, qy = std::min(std::max(d.y, cy0), cy1) - d.y;
                if (qx * qx + qy * qy <= d.r * d.r) set((size_t)j * g.nx + i, d.bit);
            }
    }
    // one allocation: masks | blobs | blob -> source
    const size_t mask_bytes = (m.size() * sizeof(unsigned long long) + 15) / 16 * 16, blob_bytes = blob_mode ? blobs.size() * sizeof(float4) : 0,
                 src_bytes = blob_mode ? blob_src.size() * sizeof(int) : 0;
    char* dm = nullptr;
    cudaError_t e = cudaMalloc((void**)&dm, mask_bytes + blob_bytes + src_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(dm, m.data(), m.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && blob_mode) e = cudaMemcpy(dm + mask_bytes, blobs.data(), blob_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && blob_mode) e = cudaMemcpy(dm + mask_bytes + blob_bytes, blob_src.data(), src_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaGetLastError(); cudaFree(dm);
        return fail(WOST_ERR_ALLOC, "source grid: device allocation failed");
    }
    g.masks = reinterpret_cast<const unsigned long long*>(dm);
    if (blob_mode) { g.blobs = reinterpret_cast<const float4*>(dm + mask_bytes); g.blob_src = reinterpret_cast<const int*>(dm + mask_bytes + blob_bytes); }
    scene->source_grids[key] = g;
    *out = g;
    return WOST_OK;
}

static int solve_impl(const wost_scene_t* scene, const wost_fields_t* fields, const wost_field_t* const* sources, int n_sources,
                      const wost_solve_params_t* P, const float* pts_xy, int64_t n_pts,
                      double* out_mean, double* out_m2, double* out_block_stats, float* out_walk_vals, uint64_t* out_steps,
                      int64_t n_trace, int32_t trace_cap, float* out_trace, int32_t* out_trace_len, void* stream) {
    if (!scene || !P || !pts_xy) return fail(WOST_ERR_INVALID, "scene, params and pts_xy are required");
    if (n_pts < 0 || P->n_walks <= 0 || P->max_steps < 0) return fail(WOST_ERR_INVALID, "n_pts >= 0, n_walks > 0, max_steps >= 0 required");
    if (!(P->eps >= 0.0f)) return fail(WOST_ERR_INVALID, "eps must be >= 0");
    const long long pstride = P->point_index_stride > 0 ? P->point_index_stride : 1;
    if (P->point_index_stride < 0) return fail(WOST_ERR_INVALID, "point_index_stride must be >= 0");
    if (n_pts >= (1ll << 32) || P->n_walks + P->walk_offset >= (1ll << 32) || (n_pts - 1) * pstride + P->point_index_base >= (1ll << 32))
        return fail(WOST_ERR_INVALID, "point and walk indices must fit 32 bits (Philox counter words)");
    const bool delta = P->delta_tracking != 0;
    if (P->compat_mode != WOST_COMPAT_REFERENCE && P->compat_mode != WOST_COMPAT_PHYSICAL) return fail(WOST_ERR_INVALID, "unknown compat_mode");
    const bool phys_delta = P->compat_mode == WOST_COMPAT_PHYSICAL && delta;
    if (delta) {
        if (!(P->sigma_bar > 0.0f)) return fail(WOST_ERR_INVALID, "delta tracking needs sigma_bar > 0");
        if (!phys_delta && (!P->screened_icdf || P->icdf_len < 2)) return fail(WOST_ERR_INVALID, "delta tracking needs the screened radius table");
        // physical mode: the weights are power series in (r sqrt(sigma_bar))^2 / 4, r <= max(1/sqrt(sigma_bar), eps/2)
        if (phys_delta && !((double)P->eps / 2.0 * sqrt((double)P->sigma_bar) <= 2.0))
            return fail(WOST_ERR_INVALID, "physical mode: eps/2 * sqrt(sigma_bar) must be <= 2 (the smallest step must resolve the absorption length)");
        if (P->sp_mode < WOST_SP_FULL || P->sp_mode > WOST_SP_FIELD) return fail(WOST_ERR_INVALID, "unknown sp_mode");
        if (P->sp_mode == WOST_SP_FIELD && !(fields && fields->sigma_prime)) return fail(WOST_ERR_INVALID, "WOST_SP_FIELD needs fields.sigma_prime");
    }
    if (n_trace > 0 && (!out_trace || !out_trace_len || trace_cap <= 0)) return fail(WOST_ERR_INVALID, "trace buffers missing");
    if (fields) {
        const wost_field_t* fs[5] = {fields->g, fields->f, fields->alpha, fields->sigma, fields->sigma_prime};
        for (auto* f : fs) if (f && f->device != scene->device) return fail(WOST_ERR_INVALID, "field and scene live on different devices");
    }
    if (n_pts == 0) { if (out_steps && !is_device_ptr(out_steps)) *out_steps = 0; return WOST_OK; }

    DeviceGuard g(scene->device);
    if (!g.ok) return fail(WOST_ERR_CUDA, "cannot select the scene's device");
    cudaStream_t st = (cudaStream_t)stream;
    Arena* ar = scene->arena_for(st);                                  // this call's temporaries (reused by the next call on `st`)
    const long long W = P->n_walks;
    const long long nblk = (W + WOST_WALK_BLOCK - 1) / WOST_WALK_BLOCK;
    const bool trace = n_trace > 0;
    const int S = n_sources > 0 ? n_sources : 1;                        // totals per walk
    const bool neu = scene->n_nseg > 0, src = n_sources > 0 || (fields && fields->f);
    for (int k = 0; k < n_sources; ++k)
        if (!sources[k] || sources[k]->device != scene->device) return fail(WOST_ERR_INVALID, "source field missing or on another device");
    if (n_sources > 0 && (n_trace > 0 || out_walk_vals)) return fail(WOST_ERR_UNSUPPORTED, "trace / per-walk output is not available for multi-source solves");

    Staged<float> s_pts, s_trace, s_icdf, s_maj; Staged<double> s_mean, s_m2, s_blk; Staged<uint64_t> s_steps; Staged<int32_t> s_tlen;
    int rc;
    if ((rc = s_pts.init(pts_xy, 2 * n_pts, false, st, ar))) return rc;
    const bool use_icdf = delta && !phys_delta;
    if ((rc = s_icdf.init(use_icdf ? P->screened_icdf : nullptr, use_icdf ? P->icdf_len : 0, false, st, ar))) return rc;
    size_t maj_len = 0;
    if (phys_delta && P->majorant_levels > 0) {
        if (P->majorant_levels > 13 || !P->majorant || !(P->majorant_dx > 0.0f) || !(P->majorant_dy > 0.0f))
            return fail(WOST_ERR_INVALID, "majorant pyramid: 1..13 levels, data and positive cell sizes needed");
        for (int l = 0, n = 1 << (P->majorant_levels - 1); l < P->majorant_levels; ++l, n >>= 1) maj_len += (size_t)n * n;
    }
    if ((rc = s_maj.init(maj_len ? P->majorant : nullptr, maj_len, false, st, ar))) return rc;
    if ((rc = s_mean.init(out_mean, (size_t)S * n_pts, true, st, ar)) || (rc = s_m2.init(out_m2, (size_t)S * n_pts, true, st, ar)) ||
        (rc = s_blk.init(out_block_stats, (size_t)S * 2 * n_pts * nblk, true, st, ar)) || (rc = s_steps.init(out_steps, 1, true, st, ar)) ||
        (rc = s_trace.init(out_trace, trace ? (size_t)n_trace * (trace_cap + 1) * 8 : 0, true, st, ar)) ||
        (rc = s_tlen.init(out_trace_len, trace ? n_trace : 0, true, st, ar))) return rc;

    // scratch: per-walk totals (unless the caller wants them), counters, block statistics
    // The per-walk buffer is bounded: evaluation points are processed in passes of at most WOST_MAX_WALK_VALS walks
    // (default 2^29 = 2 GiB of fp32), so arbitrarily large jobs run in constant scratch memory.
    long long max_vals = 1ll << 29;
    if (const char* e = std::getenv("WOST_MAX_WALK_VALS")) max_vals = std::max(1ll, std::atoll(e));
    if (W > max_vals) max_vals = W;                                     // at least one point per pass
    const long long pts_per_pass = std::max(1ll, std::min((long long)n_pts, max_vals / (W * S)));
    Scratch<float> vals_s, alpha0_s; Scratch<DevField> srcs_s; Scratch<float4> sup_s; Scratch<unsigned long long> ctrs_s; Scratch<double> blk_s;
    float* vals = nullptr; bool vals_temp = false;
    const bool vals_dev_out = out_walk_vals && is_device_ptr(out_walk_vals);
    if (!vals_dev_out) { if ((rc = vals_s.alloc((size_t)pts_per_pass * W * S, st, ar))) return rc; vals = vals_s.p; vals_temp = true; }
    // shared-memory layout of the walk kernel (float4 units): field headers | staged segment tables | term tables
    const size_t d_bytes = scene->dbvh ? 0 : sizeof(float4) * 2 * (size_t)scene->n_dseg;
    const size_t n_bytes = scene->nbvh ? 0 : sizeof(float4) * 2 * (size_t)scene->n_nseg;
    const int stage_smem = (d_bytes + n_bytes <= 96 * 1024) ? ((d_bytes ? 1 : 0) | (n_bytes ? 2 : 0)) : 0;
    DevFields DF = dev_fields_of(fields);
    size_t smem_f4 = (size_t)DEVFIELD_F4 * (FIELD_SOURCE0 + n_sources) + (stage_smem ? (d_bytes + n_bytes) / sizeof(float4) : 0);
    for (DevField* f : {&DF.g, &DF.f, &DF.alpha, &DF.sigma, &DF.sigma_prime}) { f->term_off = (int32_t)smem_f4; smem_f4 += 4 * (size_t)f->n_terms; }
    DevField* d_srcs = nullptr; float4* d_sup = nullptr;
    std::vector<DevField> h(n_sources); std::vector<float4> hs(n_sources);
    for (int k = 0; k < n_sources; ++k) {
        h[k] = sources[k]->d; hs[k] = sources[k]->support;
        h[k].term_off = (int32_t)smem_f4; smem_f4 += 4 * (size_t)h[k].n_terms;
    }
    const size_t smem = smem_f4 * sizeof(float4);
    if (smem + 16 * 1024 > scene->smem_optin)
        return fail(WOST_ERR_UNSUPPORTED, "field term tables and segment tables do not fit shared memory (" + std::to_string(smem) + " bytes)");
    if (n_sources > 0) {
        if ((rc = srcs_s.alloc(n_sources, st, ar)) || (rc = sup_s.alloc(n_sources, st, ar))) return rc;
        d_srcs = srcs_s.p; d_sup = sup_s.p;
        CU(cudaMemcpyAsync(d_srcs, h.data(), sizeof(DevField) * n_sources, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_sup, hs.data(), sizeof(float4) * n_sources, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));                                   // the host vectors go out of scope
    }
    if ((rc = ctrs_s.alloc(2, st, ar))) return rc;
    unsigned long long* ctrs = ctrs_s.p;
    if (delta && (rc = alpha0_s.alloc((size_t)n_pts, st, ar))) return rc;
    float* alpha0 = alpha0_s.p;
    CU(cudaMemsetAsync(ctrs, 0, 2 * sizeof(unsigned long long), st));
    double* blk = s_blk.dev; bool blk_temp = false;
    if (!blk) { if ((rc = blk_s.alloc((size_t)2 * n_pts * nblk * S, st, ar))) return rc; blk = blk_s.p; blk_temp = true; }
    if (trace) {
        CU(cudaMemsetAsync(s_trace.dev, 0xff, sizeof(float) * (size_t)n_trace * (trace_cap + 1) * 8, st));   // NaN fill
        CU(cudaMemsetAsync(s_tlen.dev, 0, sizeof(int32_t) * n_trace, st));
    }

    WalkArgs a{};
    a.dseg = scene->dseg; a.n_dseg = scene->n_dseg; a.nseg = scene->nseg; a.n_nseg = scene->n_nseg;
    a.F = DF;
    a.n_src = n_sources; a.srcs = d_srcs; a.src_support = d_sup;
    if (n_sources > 0 && env_int("WOST_SOURCE_GRID", 1) && (rc = source_grid_for(scene, sources, n_sources, &a.sgrid))) return rc;
    if (alpha0) { alpha0_kernel<<<blocks_for(n_pts, 256), 256, 0, st>>>(a.F, s_pts.dev, n_pts, alpha0); CU(cudaGetLastError()); }
    a.pts = s_pts.dev; a.n_pts = n_pts; a.n_walks = W;
    a.walks_magic = W <= 1 ? 0xffffffffu : (uint32_t)((1ull << 32) / (unsigned long long)W);
    a.max_steps = P->max_steps; a.eps = P->eps; a.rmin = (float)((double)P->eps / 2.0);   // :167
    a.sp_mode = P->sp_mode; a.sigma_bar = P->sigma_bar;
    a.inv_sigma_bar = delta ? (float)(1.0 / (double)P->sigma_bar) : 0.0f;
    a.sqrt_sigma_bar = delta ? (float)std::sqrt((double)P->sigma_bar) : 0.0f;
    a.icdf = s_icdf.dev; a.icdf_len = P->icdf_len;
    if (delta && !phys_delta && !(a.iprob = iprob_table(scene->device))) return fail(WOST_ERR_ALLOC, "interior-probability table: device allocation failed");
    a.key0 = (uint32_t)P->seed; a.key1 = (uint32_t)(P->seed >> 32);
    for (int r = 0; r < 10; ++r) { a.ks[2 * r] = a.key0 + (uint32_t)r * 0x9E3779B9u; a.ks[2 * r + 1] = a.key1 + (uint32_t)r * 0xBB67AE85u; }
    a.point_index_base = P->point_index_base; a.point_index_stride = pstride; a.walk_offset = P->walk_offset;
    a.walk_vals = vals; a.counter = ctrs; a.steps_total = ctrs + 1;
    a.ndisc_x = scene->ndisc_x; a.ndisc_y = scene->ndisc_y; a.ndisc_r = scene->ndisc_r; a.ndisc_r2 = scene->ndisc_r2;
    coop_thresholds(scene->n_nseg, &a.sil_coop_max, &a.ray_coop_max);
    a.dbvh.nodes = scene->dbvh; a.dbvh.n_leaves = scene->dbvh_leaves; a.nbvh.nodes = scene->nbvh; a.nbvh.cones = scene->ncones; a.nbvh.n_leaves = scene->nbvh_leaves;
    a.dwide = scene->dwide; a.nwide = scene->nwide; a.wide_coop_max = env_int("WOST_WIDE_COOP_MAX", 20);
    a.bvh_slack = scene->bvh_slack; a.neu_closed = scene->neu_closed; a.phys_nudge = scene->phys_nudge;
    a.phys_rcap = phys_delta ? 1.0f / sqrtf(P->sigma_bar) : 0.0f;
    a.maj.data = maj_len ? s_maj.dev : nullptr; a.maj.levels = P->majorant_levels;
    a.maj.x0 = P->majorant_x0; a.maj.y0 = P->majorant_y0; a.maj.dx = P->majorant_dx; a.maj.dy = P->majorant_dy;
    a.n_trace = trace ? n_trace : 0; a.trace_cap = trace_cap; a.trace = s_trace.dev; a.trace_len = s_tlen.dev;

    a.stage_smem = stage_smem;
    const bool phys = P->compat_mode == WOST_COMPAT_PHYSICAL;
    const bool big = scene->dbvh != nullptr || scene->nbvh != nullptr;
    const void* kern = (const void*)(big ? (trace ? pick_kernel<true, true>(neu, src, delta, phys) : pick_kernel<false, true>(neu, src, delta, phys))
                                         : (trace ? pick_kernel<true, false>(neu, src, delta, phys) : pick_kernel<false, false>(neu, src, delta, phys)));
    const bool dbg_t = std::getenv("WOST_TIMING") != nullptr;
    auto t_now = [] { return std::chrono::steady_clock::now(); };
    auto t_ms = [](std::chrono::steady_clock::time_point a_, std::chrono::steady_clock::time_point b_) { return std::chrono::duration<double, std::milli>(b_ - a_).count(); };
    const auto T0 = t_now();
    {   // per-solver specialised kernel (wost_jit.inc): params->jit 0 = auto (jobs of >= 2^18 walks, the 8th solve with the
        // same fields, or already compiled), 1 = always, 2 = never; the environment variable WOST_JIT (0 / 1) overrides
        int mode = P->jit;
        if (const char* e = std::getenv("WOST_JIT")) mode = std::atoi(e) ? 1 : 2;
        std::string why = "specialisation switched off";
        const jit::Kernel* K = nullptr;
        if (mode != 2) {
            jit::Flags fl{neu, src, delta, trace, phys, big, n_sources > 0, delta ? P->sp_mode : 0, env_int("WOST_JIT_MIN_BLOCKS", 4)};
            if (!big && scene->n_dseg <= 64 && scene->n_nseg <= 64 && stage_smem == ((d_bytes ? 1 : 0) | (n_bytes ? 2 : 0))) {
                // small scene: its sizes and query strategy become compile-time constants too
                fl.n_dseg = scene->n_dseg; fl.n_nseg = scene->n_nseg; fl.sil_coop_max = a.sil_coop_max; fl.ray_coop_max = a.ray_coop_max;
                fl.stage = stage_smem;
                fl.dir_rcp = scene->dir_rcp != 0;
            }
            wost_fields_t none{};
            K = jit::get(fields ? fields : &none, fl, scene->device, mode == 1 || (long long)n_pts * W >= (1ll << 18),
                         env_int("WOST_JIT_AUTO_AFTER", 8), &why);
        }
        if (K) { kern = (const void*)K->fn; why.clear(); }
        else if (mode == 1 && !std::getenv("WOST_JIT_SOFT"))
            return fail(WOST_ERR_UNSUPPORTED, "specialised kernel requested (jit = 1) but unavailable: " + why);
        std::lock_guard<std::mutex> lk(jit::g_mu);
        jit::g_last_note = why;
        if (K) ++jit::g_launches;
    }
    const auto T1 = t_now();
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0, occ_small = 0;                                         // resident CTAs per SM: 256 threads / one-warp CTAs (small jobs)
    const int small_wps = env_int("WOST_SMALL_WARPS_PER_SCHEDULER", 2), forced_lanes = env_int("WOST_LANES_PER_WARP", 0);
    {
        static std::mutex mu;                                           // the query costs microseconds: once per (kernel, shared memory)
        static std::map<std::tuple<const void*, size_t, int>, std::pair<int, int>> cache;
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find({kern, smem, scene->device});
        if (it == cache.end()) {
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_small, kern, 32, smem));
            it = cache.emplace(std::make_tuple(kern, smem, scene->device), std::make_pair(occ, occ_small)).first;
        }
        occ = it->second.first; occ_small = it->second.second;
    }
    const auto T2 = t_now();
    if (occ < 1) return fail(WOST_ERR_CUDA, "walk kernel does not fit on an SM");
    { const int cap = env_int("WOST_MAX_CTAS_PER_SM", 0); if (cap > 0 && occ > cap) occ = cap; }   // experiments: fewer resident warps
    for (long long p0 = 0; p0 < n_pts; p0 += pts_per_pass) {
        const long long np = std::min(pts_per_pass, (long long)n_pts - p0);
        const long long total = np * W;
        // Small jobs are latency-bound: the solve takes as long as its longest walk, and walks that share a warp wait for
        // each other's divergent branches.  A job that cannot fill the machine deals its walks to fewer lanes per warp
        // (about `small_wps` warps per SM scheduler) in one-warp CTAs, which the block scheduler spreads over all SMs.
        int lanes = 32, threads = 256, occ_l = occ;
        {
            const long long slots = (long long)scene->sm_count * 4 * small_wps;            // warps the small-job layout uses
            if (total < slots * 32) {
                lanes = 1;
                while (lanes < 32 && total > slots * lanes) lanes *= 2;
            }
            if (forced_lanes >= 1 && forced_lanes <= 32) lanes = forced_lanes;
            if (lanes < 32) { threads = 32; occ_l = occ_small; }
        }
        a.lanes = lanes;
        long long grid = (long long)scene->sm_count * occ_l;           // persistent: one wave, a multiple of the SM count
        const long long per_cta = (long long)lanes * (threads / 32);
        const long long want = (total + per_cta - 1) / per_cta;
        if (grid > want) grid = want;
        const long long nwarps = grid * (threads / 32);
        // walks a warp reserves per atomic: one per lane.  Larger reservations (round 1: up to 1 024) save atomics nobody
        // misses and leave some warps with a private backlog at the end of the job while others have run dry: 32 instead of
        // the old total / (8 x warps) is +1.5 % on the headline job and +3 ... +8 % on the short-walk scenes
        // (profiles/r2_chunk_sweep.txt)
        a.chunk = 32;
        if (total < nwarps * 32) a.chunk = (int)((total + nwarps - 1) / nwarps);
        if (a.chunk < 1) a.chunk = 1;
        { const int forced = env_int("WOST_CHUNK", 0); if (forced > 0) a.chunk = forced; }   // measurement knob
        a.pts = s_pts.dev + 2 * p0; a.n_pts = np; a.point_index_base = P->point_index_base + p0 * pstride;
        a.alpha0 = alpha0 ? alpha0 + p0 : nullptr;
        a.walk_vals = vals_dev_out ? out_walk_vals + (size_t)p0 * W : vals;
        if (trace) {                                                    // the first n_trace walks in point-major order
            const long long first = p0 * W;
            a.n_trace = std::max(0ll, std::min((long long)n_trace - first, total));
            a.trace = s_trace.dev + (size_t)first * (trace_cap + 1) * 8; a.trace_len = s_tlen.dev + first;
        }
        if (p0 > 0) CU(cudaMemsetAsync(ctrs, 0, sizeof(unsigned long long), st));   // walk counter only; steps accumulate
        if (n_sources > 0) CU(cudaMemsetAsync(vals, 0, sizeof(float) * (size_t)np * W * S, st));   // rows accumulate
        const auto T3 = t_now();
        { void* kargs[] = {(void*)&a}; CU(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(threads), kargs, smem, st)); }
        if (dbg_t) std::fprintf(stderr, "[wost timing] jit::get %.3f ms, occupancy %.3f ms (occ %d), to launch %.3f ms, launch %.3f ms\n", t_ms(T0, T1), t_ms(T1, T2), occ, t_ms(T2, T3), t_ms(T3, t_now()));
        if (n_sources > 0)
            block_stats_kernel<<<dim3(blocks_for(np * nblk * 32, 256), S), 256, 0, st>>>(a.walk_vals, np, W, nblk, blk, S, n_pts, p0);
        else
            block_stats_kernel<<<blocks_for(np * nblk * 32, 256), 256, 0, st>>>(a.walk_vals, np, W, nblk, blk + 2 * p0 * nblk);
        CU(cudaGetLastError());
        if (out_walk_vals && vals_temp)
            CU(cudaMemcpyAsync(out_walk_vals + (size_t)p0 * W, vals, sizeof(float) * (size_t)np * W, cudaMemcpyDeviceToHost, st));
    }
    if (s_mean.dev || s_m2.dev) {
        merge_stats_kernel<<<blocks_for((long long)S * n_pts, 128), 128, 0, st>>>(blk, (long long)S * n_pts, W, nblk, s_mean.dev, s_m2.dev);
        CU(cudaGetLastError());
    }
    if (s_steps.dev) CU(cudaMemcpyAsync(s_steps.dev, ctrs + 1, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    bool sync = s_mean.host_out() || s_m2.host_out() || s_blk.host_out() || s_steps.host_out() || s_trace.host_out() || s_tlen.host_out();
    if (out_walk_vals && vals_temp) sync = true;
    if ((rc = s_pts.finish()) || (rc = s_icdf.finish()) || (rc = s_maj.finish()) || (rc = s_mean.finish()) || (rc = s_m2.finish()) || (rc = s_blk.finish()) ||
        (rc = s_steps.finish()) || (rc = s_trace.finish()) || (rc = s_tlen.finish())) return rc;
    // Release the scratch BEFORE the synchronisation (the destructors only cover the error paths): with the frees queued
    // after it, host-side callers (a sync per solve) saw their calls jitter between 10 and 40 ms; this order is steady.
    if ((rc = vals_s.release()) || (rc = srcs_s.release()) || (rc = sup_s.release()) || (rc = blk_s.release()) ||
        (rc = ctrs_s.release()) || (rc = alpha0_s.release())) return rc;
    if (sync) CU(cudaStreamSynchronize(st));
    return WOST_OK;
}

extern "C" {

int wost_solve(const wost_scene_t* scene, const wost_fields_t* fields, const wost_solve_params_t* P,
               const float* pts_xy, int64_t n_pts,
               double* out_mean, double* out_m2, double* out_block_stats, float* out_walk_vals, uint64_t* out_steps,
               int64_t n_trace, int32_t trace_cap, float* out_trace, int32_t* out_trace_len, void* stream) {
    return solve_impl(scene, fields, nullptr, 0, P, pts_xy, n_pts, out_mean, out_m2, out_block_stats, out_walk_vals, out_steps,
                      n_trace, trace_cap, out_trace, out_trace_len, stream);
}

int wost_solve_multi_source(const wost_scene_t* scene, const wost_fields_t* fields, const wost_field_t* const* sources,
                            int32_t n_sources, const wost_solve_params_t* P, const float* pts_xy, int64_t n_pts,
                            double* out_mean, double* out_m2, double* out_block_stats, uint64_t* out_steps, void* stream) {
    if (n_sources < 1 || !sources) return fail(WOST_ERR_INVALID, "at least one source field is required");
    if (n_sources > 4096) return fail(WOST_ERR_INVALID, "at most 4096 sources per call");
    return solve_impl(scene, fields, sources, n_sources, P, pts_xy, n_pts, out_mean, out_m2, out_block_stats, nullptr, out_steps,
                      0, 0, nullptr, nullptr, stream);
}

int wost_merge_block_stats(const double* block_stats, int64_t n_pts, int64_t n_walks, int32_t device,
                           double* out_mean, double* out_m2, void* stream) {
    if (!block_stats || n_pts < 0 || n_walks <= 0) return fail(WOST_ERR_INVALID, "bad arguments");
    if (n_pts == 0) return WOST_OK;
    if (wost_device_count() <= 0) return fail(WOST_ERR_CUDA, "no CUDA device available (libwost has no CPU fallback)");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    const long long nblk = (n_walks + WOST_WALK_BLOCK - 1) / WOST_WALK_BLOCK;
    Staged<double> sb, sm, sq; int rc;
    if ((rc = sb.init(block_stats, 2 * n_pts * nblk, false, st)) || (rc = sm.init(out_mean, n_pts, true, st)) || (rc = sq.init(out_m2, n_pts, true, st))) return rc;
    merge_stats_kernel<<<blocks_for(n_pts, 128), 128, 0, st>>>(sb.dev, n_pts, n_walks, nblk, sm.dev, sq.dev);
    CU(cudaGetLastError());
    const bool sync = sm.host_out() || sq.host_out();
    if ((rc = sb.finish()) || (rc = sm.finish()) || (rc = sq.finish())) return rc;
    if (sync) CU(cudaStreamSynchronize(st));
    return WOST_OK;
}

// ---- geometry parity entry points ------------------------------------------------------------------
static int seg_view(const wost_scene_t* s, int which, SegView* v) {
    if (!s) return fail(WOST_ERR_INVALID, "scene is NULL");
    if (which == 0) { v->seg = s->dseg; v->n = s->n_dseg; v->dirichlet_layout = 1; return 0; }
    if (which == 1) {
        if (!s->n_nseg) return fail(WOST_ERR_INVALID, "scene has no Neumann polyline");
        v->seg = s->nseg; v->n = s->n_nseg; v->dirichlet_layout = 0; return 0;
    }
    return fail(WOST_ERR_INVALID, "which must be 0 (Dirichlet) or 1 (Neumann)");
}

int wost_geom_distance(const wost_scene_t* s, int32_t which, const float* p, int64_t B, float* out_d, int32_t* out_seg, void* stream) {
    SegView v; int rc;
    if ((rc = seg_view(s, which, &v))) return rc;
    if (!p || B < 0) return fail(WOST_ERR_INVALID, "bad arguments");
    if (B == 0) return WOST_OK;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    Arena* ar = s->arena_for(st);
    // the distance kernel wants the Dirichlet layout: the Neumann polyline has one too, built by wost_scene_create
    const float4* seg = v.dirichlet_layout ? v.seg : s->nseg_as_dir;
    Staged<float> sp, sd; Staged<int32_t> ss;
    if ((rc = sp.init(p, 2 * B, false, st, ar)) || (rc = sd.init(out_d, B, true, st, ar)) || (rc = ss.init(out_seg, B, true, st, ar))) return rc;
    Bvh bvh{}; if (v.dirichlet_layout) { bvh.nodes = s->dbvh; bvh.n_leaves = s->dbvh_leaves; }
    geom_distance_kernel<<<blocks_for(B, 256), 256, 0, st>>>(seg, v.n, bvh, sp.dev, B, sd.dev, ss.dev);
    CU(cudaGetLastError());
    const bool sync = sd.host_out() || ss.host_out();
    if ((rc = sp.finish()) || (rc = sd.finish()) || (rc = ss.finish())) return rc;
    if (sync) CU(cudaStreamSynchronize(st));
    return WOST_OK;
}

int wost_geom_silhouette(const wost_scene_t* s, int32_t which, const float* p, int64_t B, float* out_d, uint8_t* out_mask, void* stream) {
    SegView v; int rc;
    if ((rc = seg_view(s, which, &v))) return rc;
    if (!p || B < 0) return fail(WOST_ERR_INVALID, "bad arguments");
    if (B == 0) return WOST_OK;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    Arena* ar = s->arena_for(st);
    Staged<float> sp, sd; Staged<uint8_t> sm;
    if ((rc = sp.init(p, 2 * B, false, st, ar)) || (rc = sd.init(out_d, B, true, st, ar)) || (rc = sm.init(out_mask, (size_t)B * (v.n - 1), true, st, ar))) return rc;
    Bvh bvh{}; if (!v.dirichlet_layout) { bvh.nodes = s->nbvh; bvh.cones = s->ncones; bvh.n_leaves = s->nbvh_leaves; }
    geom_silhouette_kernel<<<blocks_for(B, 256), 256, 0, st>>>(v, bvh, sp.dev, B, sd.dev, sm.dev);
    CU(cudaGetLastError());
    const bool sync = sd.host_out() || sm.host_out();
    if ((rc = sp.finish()) || (rc = sd.finish()) || (rc = sm.finish())) return rc;
    if (sync) CU(cudaStreamSynchronize(st));
    return WOST_OK;
}

int wost_geom_ray(const wost_scene_t* s, int32_t which, const float* p, const float* dir, int64_t B, float* out_s, void* stream) {
    SegView v; int rc;
    if ((rc = seg_view(s, which, &v))) return rc;
    if (!p || !dir || !out_s || B < 0) return fail(WOST_ERR_INVALID, "bad arguments");
    if (B == 0) return WOST_OK;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    Arena* ar = s->arena_for(st);
    Staged<float> sp, sdir, so;
    if ((rc = sp.init(p, 2 * B, false, st, ar)) || (rc = sdir.init(dir, 2 * B, false, st, ar)) || (rc = so.init(out_s, (size_t)B * v.n, true, st, ar))) return rc;
    geom_ray_kernel<<<blocks_for(B, 256), 256, 0, st>>>(v, sp.dev, sdir.dev, B, so.dev);
    CU(cudaGetLastError());
    const bool sync = so.host_out();
    if ((rc = sp.finish()) || (rc = sdir.finish()) || (rc = so.finish())) return rc;
    if (sync) CU(cudaStreamSynchronize(st));
    return WOST_OK;
}

int wost_geom_intersect(const wost_scene_t* s, int32_t which, const float* p, const float* dir, const float* r, int64_t B,
                        float* out_pt, float* out_nrm, uint8_t* out_found, int32_t* out_seg, void* stream) {
    SegView v; int rc;
    if ((rc = seg_view(s, which, &v))) return rc;
    if (!p || !dir || !r || B < 0) return fail(WOST_ERR_INVALID, "bad arguments");
    if (B == 0) return WOST_OK;
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    Arena* ar = s->arena_for(st);
    // intersect runs on the Neumann layout: the Dirichlet polyline has one too, built by wost_scene_create
    const float4* seg = v.dirichlet_layout ? s->dseg_as_neu : v.seg;
    Staged<float> sp, sdir, sr, spt, snr; Staged<uint8_t> sf; Staged<int32_t> ss;
    if ((rc = sp.init(p, 2 * B, false, st, ar)) || (rc = sdir.init(dir, 2 * B, false, st, ar)) || (rc = sr.init(r, B, false, st, ar)) ||
        (rc = spt.init(out_pt, 2 * B, true, st, ar)) || (rc = snr.init(out_nrm, 2 * B, true, st, ar)) || (rc = sf.init(out_found, B, true, st, ar)) ||
        (rc = ss.init(out_seg, B, true, st, ar))) return rc;
    Bvh bvh{}; if (!v.dirichlet_layout) { bvh.nodes = s->nbvh; bvh.n_leaves = s->nbvh_leaves; }
    geom_intersect_kernel<<<blocks_for(B, 256), 256, 0, st>>>(seg, v.n, bvh, s->bvh_slack, sp.dev, sdir.dev, sr.dev, B, spt.dev, snr.dev, sf.dev, ss.dev);
    CU(cudaGetLastError());
    const bool sync = spt.host_out() || snr.host_out() || sf.host_out() || ss.host_out();
    if ((rc = sp.finish()) || (rc = sdir.finish()) || (rc = sr.finish()) || (rc = spt.finish()) || (rc = snr.finish()) || (rc = sf.finish()) || (rc = ss.finish())) return rc;
    if (sync) CU(cudaStreamSynchronize(st));
    return WOST_OK;
}

// Developer diagnostic (no device needed): generate and compile the specialised kernel for the given field descriptors and
// write <prefix>.cu / <prefix>.cubin, so that registers, spills and SASS can be inspected with cuobjdump on a CPU-only box.
int wost_jit_offline(const wost_field_desc_t* const descs[5] /* g, f, alpha, sigma, sigma_prime; NULL = absent */,
                     int32_t neu, int32_t src, int32_t delta, int32_t trace, int32_t phys, int32_t big, int32_t multi,
                     int32_t sp_mode, int32_t min_blocks, int32_t n_dseg, int32_t n_nseg, const char* arch, const char* prefix) {
    wost_field tmp[5]; wost_fields_t F{};
    const wost_field_t** slots[5] = {&F.g, &F.f, &F.alpha, &F.sigma, &F.sigma_prime};
    for (int i = 0; i < 5; ++i) {
        const wost_field_desc_t* d = descs[i];
        if (!d) continue;
        DevField& D = tmp[i].d;
        D.present = 1; D.kind = d->kind; D.n_terms = d->kind == WOST_FIELD_TERMS ? d->n_terms : 0; D.mask_kind = d->mask_kind;
        D.c0 = d->c0; D.m0 = d->mask[0]; D.m1 = d->mask[1]; D.m2 = d->mask[2]; D.m3 = d->mask[3]; D.outside = d->outside;
        tmp[i].h_terms = device_form_terms(d);
        *slots[i] = &tmp[i];
    }
    jit::Flags fl{neu != 0, src != 0, delta != 0, trace != 0, phys != 0, big != 0, multi != 0, sp_mode, min_blocks};
    if (n_dseg >= 0) { fl.n_dseg = n_dseg; fl.n_nseg = n_nseg; coop_thresholds(n_nseg, &fl.sil_coop_max, &fl.ray_coop_max); fl.stage = 1 | (n_nseg ? 2 : 0);
                       fl.dir_rcp = env_int("WOST_DIRICHLET_RCP", 1) != 0; }
    const std::string source = jit::generate(&F, fl);
    const std::string pre = prefix ? prefix : "wost_walk_jit";
    if (FILE* fp = std::fopen((pre + ".cu").c_str(), "w")) { std::fwrite(source.data(), 1, source.size(), fp); std::fclose(fp); }
    std::vector<char> cubin; std::string err; double ms = 0.0;
    if (!jit::compile(source, arch && *arch ? std::string("--gpu-architecture=") + arch : std::string("--gpu-architecture=sm_100a"), &cubin, &err, &ms))
        return fail(WOST_ERR_UNSUPPORTED, err);
    if (FILE* fp = std::fopen((pre + ".cubin").c_str(), "wb")) { std::fwrite(cubin.data(), 1, cubin.size(), fp); std::fclose(fp); }
    g_err = "compiled in " + std::to_string(ms) + " ms";
    return WOST_OK;
}

int wost_jit_stats(int64_t* out_compiled, int64_t* out_cache_hits, int64_t* out_launches) {
    std::lock_guard<std::mutex> lk(jit::g_mu);
    if (out_compiled) *out_compiled = jit::g_compiled;
    if (out_cache_hits) *out_cache_hits = jit::g_hits;
    if (out_launches) *out_launches = jit::g_launches;
    return WOST_OK;
}

const char* wost_jit_last_note(void) {
    static thread_local std::string note;
    std::lock_guard<std::mutex> lk(jit::g_mu);
    note = jit::g_last_note;
    return note.c_str();
}

int wost_selftest_division(int32_t device, int64_t n, uint64_t seed, const float* divisors, int32_t n_divisors, int64_t out_mismatches[4]) {
    if (wost_device_count() <= 0) return fail(WOST_ERR_CUDA, "no CUDA device available");
    if (n <= 0 || !out_mismatches || (n_divisors > 0 && !divisors) || n_divisors > 64) return fail(WOST_ERR_INVALID, "bad arguments");
    DeviceGuard g(device);
    float ys[64]; float bs[64];
    for (int k = 0; k < n_divisors; ++k) {
        bs[k] = divisors[k];
        if (!verified_reciprocal(bs[k], &ys[k])) return fail(WOST_ERR_INVALID, "divisor " + std::to_string(k) + " does not verify (out of [2^-40, 2^40]?)");
    }
    unsigned long long* d = nullptr; float* dby = nullptr;
    CU(cudaMalloc((void**)&d, 4 * sizeof(unsigned long long)));
    CU(cudaMemset(d, 0, 4 * sizeof(unsigned long long)));
    if (n_divisors > 0) {
        CU(cudaMalloc((void**)&dby, 2 * sizeof(float) * n_divisors));
        CU(cudaMemcpy(dby, bs, sizeof(float) * n_divisors, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dby + n_divisors, ys, sizeof(float) * n_divisors, cudaMemcpyHostToDevice));
    }
    const int threads = 256; const long long blocks = std::min<long long>((n + threads - 1) / threads, 148 * 32);
    division_selftest_kernel<<<(unsigned)blocks, threads>>>(n, (uint32_t)seed, (uint32_t)(seed >> 32), dby, n_divisors, d);
    CU(cudaGetLastError());
    unsigned long long h[4];
    CU(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(dby);
    for (int i = 0; i < 4; ++i) out_mismatches[i] = (int64_t)h[i];
    return WOST_OK;
}

int wost_fp32_peak(int32_t device, double* out_tflops, double* out_sm_mhz_effective) {
    if (wost_device_count() <= 0) return fail(WOST_ERR_CUDA, "no CUDA device available");
    DeviceGuard g(device);
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, device));
    float* d = nullptr;
    CU(cudaMalloc((void**)&d, sizeof(float)));
    const int iters = 1 << 15, threads = 256, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) fma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0));
        fma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0.0f; CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * threads * (double)blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (out_tflops) *out_tflops = best;
    if (out_sm_mhz_effective) *out_sm_mhz_effective = best * 1e12 / (2.0 * 128.0 * prop.multiProcessorCount) / 1e6;
    return WOST_OK;
}

}  // extern "C"
