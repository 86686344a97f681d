/*
 * wost_oracle.c — CPU ORACLE (test infrastructure only; see wost_oracle.h).
 *
 * Plain-C restatement of the reference's Walk-on-Stars hot path.  Every function cites the
 * reference lines it follows (paths relative to the reference repository root).
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include "wost_oracle.h"
#include "../include/wost_math.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Elementary functions.  Default: include/wost_math.h — explicit fp32 arithmetic that the CUDA kernel evaluates
 * identically, so Philox-mode walks can be compared bit for bit.  g_libm = 1 (set for the duration of an ORC_RNG_MT
 * solve, or by orc_set_libm): glibc's expf/sinf/cosf, as close as C gets to the torch calls of the reference; that is
 * the configuration pinned against the reference fixtures.
 * ---------------------------------------------------------------------------------------- */
static int g_libm = 0;
void orc_set_libm(int on) { g_libm = on; }
static float g_iprob[WM_IPROB_N]; static int g_iprob_ready = 0;   /* 1 - 1/I0 table of include/wost_math.h (filled before the parallel region) */
static void iprob_init(void) { if (!g_iprob_ready) { wm_interior_probability_table(g_iprob); g_iprob_ready = 1; } }
static inline float o_expf(float x) { return g_libm ? expf(x) : wm_expf(x); }
static inline float o_smooth_step(float a) { return g_libm ? 1.0f / (1.0f + expf(a)) : wm_smooth_step(a); }
static inline void o_sincosf(float a, float* sn, float* cs) {
    if (g_libm) { *sn = sinf(a); *cs = cosf(a); } else wm_sincosf(a, sn, cs);
}

/* ------------------------------------------------------------------------------------------
 * RNG streams
 * ---------------------------------------------------------------------------------------- */

/* mt19937 — the engine behind torch's CPU generator and numpy's legacy RandomState.
 * torch.manual_seed(s) / np.random.seed(s) both run init_genrand(s). */
typedef struct { uint32_t mt[624]; int idx; } mt_t;

static void mt_seed(mt_t* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}
static uint32_t mt_u32(mt_t* g) {
    if (g->idx >= 624) {
        uint32_t* mt = g->mt;
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
/* torch.rand(1) on CPU, float32: 24 random bits * 2^-24  (solvers/WoStSolver.py:226,272) */
static float mt_torch_rand(mt_t* g) { return (float)(mt_u32(g) & 0xffffffu) * (1.0f / 16777216.0f); }
/* numpy legacy random_sample(): 53-bit double from two draws (solvers/utils.py:145,148,187,193) */
static double mt_numpy_double(mt_t* g) {
    uint32_t a = mt_u32(g) >> 5, b = mt_u32(g) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
static double mt_numpy_uniform(mt_t* g, double lo, double hi) { return lo + (hi - lo) * mt_numpy_double(g); }

/* Philox4x32-10 (Salmon et al. 2011), the counter-based stream the CUDA kernel uses. */
void orc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ------------------------------------------------------------------------------------------
 * Geometry primitives — geometry/PolylinesSimple.py
 * ---------------------------------------------------------------------------------------- */

/* torch.norm over 2 elements: ATen's fp32 reduction evaluates sqrt(fma(y, y, x*x)) — verified bit-for-bit
 * against torch 2.11 CPU on 2e4 random vectors (plain x*x+y*y matches only 91 %). */
static inline float norm2f(float a, float b) { return sqrtf(fmaf(b, b, a * a)); }

/* distance_to_polyline_jit, PolylinesSimple.py:26-49 */
float orc_distance(const float* pts, int n, float px, float py) {
    float best = INFINITY;
    for (int k = 0; k + 1 < n; ++k) {
        float ax = pts[2 * k], ay = pts[2 * k + 1], bx = pts[2 * k + 2], by = pts[2 * k + 3];
        float ux = bx - ax, uy = by - ay;             /* :37 */
        float vx = px - ax, vy = py - ay;             /* :38 */
        float dot_uv = vx * ux + vy * uy;             /* :41 */
        float dot_uu = ux * ux + uy * uy;             /* :42 */
        float t = dot_uv / dot_uu;                    /* :43 */
        t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
        float cx = (1.0f - t) * ax + t * bx;          /* :46 */
        float cy = (1.0f - t) * ay + t * by;
        float d = norm2f(cx - px, cy - py);           /* :47 */
        if (d < best) best = d;                       /* :49 */
    }
    return best;
}

/* cross_product_2d_jit, PolylinesSimple.py:14-23: a_x b_y - a_y b_x (mul, mul, sub) */
static inline float cross2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }

/* is_silhouette_jit, PolylinesSimple.py:52-81. Interior vertices 1..n-2 only (Q4). Returns count. */
int orc_is_silhouette(const float* pts, int n, float px, float py, uint8_t* mask) {
    int cnt = 0;
    for (int i = 1; i + 1 < n; ++i) {
        float ax = pts[2 * i - 2], ay = pts[2 * i - 1], bx = pts[2 * i], by = pts[2 * i + 1];
        float cx = pts[2 * i + 2], cy = pts[2 * i + 3];
        float c1 = cross2(bx - ax, by - ay, px - ax, py - ay);   /* :77 */
        float c2 = cross2(cx - bx, cy - by, px - bx, py - by);   /* :78 */
        int s = (c1 * c2 < 0.0f);                                 /* :81 */
        if (mask) mask[i - 1] = (uint8_t)s;
        cnt += s;
    }
    return cnt;
}

/* silhouette_distance_jit, PolylinesSimple.py:84-102 */
float orc_silhouette_distance(const float* pts, int n, float px, float py) {
    float best = INFINITY;                                        /* :98-99 inf if none */
    for (int i = 1; i + 1 < n; ++i) {
        float ax = pts[2 * i - 2], ay = pts[2 * i - 1], bx = pts[2 * i], by = pts[2 * i + 1];
        float cx = pts[2 * i + 2], cy = pts[2 * i + 3];
        float c1 = cross2(bx - ax, by - ay, px - ax, py - ay);
        float c2 = cross2(cx - bx, cy - by, px - bx, py - by);
        if (c1 * c2 < 0.0f) {
            float d = norm2f(bx - px, by - py);                   /* :101 */
            if (d < best) best = d;
        }
    }
    return best;
}

/* ray_intersection_jit, PolylinesSimple.py:105-132. Returns the SEGMENT parameter s (Q1). */
void orc_ray_intersection(const float* pts, int n, float px, float py, float dx, float dy, float* out_s) {
    for (int k = 0; k + 1 < n; ++k) {
        float ax = pts[2 * k], ay = pts[2 * k + 1], bx = pts[2 * k + 2], by = pts[2 * k + 3];
        float ux = bx - ax, uy = by - ay;             /* :119 */
        float wx = px - ax, wy = py - ay;             /* :120 */
        float d = cross2(dx, dy, ux, uy);             /* :123 */
        float s = cross2(dx, dy, wx, wy) / d;         /* :124 */
        float t = cross2(ux, uy, wx, wy) / d;         /* :125 */
        int valid = (s >= 0.0f) && (s <= 1.0f) && (t > 0.0f);   /* :128 */
        out_s[k] = valid ? s : INFINITY;              /* :130 */
    }
}

/* intersect_polylines_jit, PolylinesSimple.py:135-197. Returns found flag. */
int orc_intersect_polylines(const float* pts, int n, float px, float py, float dx, float dy, float r,
                            float* out_pt, float* out_nrm, int32_t* out_seg) {
    float dn = norm2f(dx, dy);                        /* :149 */
    if (out_seg) *out_seg = -1;
    if (dn < 1e-10f) {                                /* :150-154 */
        out_pt[0] = px; out_pt[1] = py; out_nrm[0] = 1.0f; out_nrm[1] = 0.0f;
        return 0;
    }
    float ex = dx / dn, ey = dy / dn;                 /* :156 */
    float ox = px + 1e-6f * ex, oy = py + 1e-6f * ey; /* :159 */
    float best = INFINITY; int idx = -1;
    for (int k = 0; k + 1 < n; ++k) {                 /* :162 via ray_intersection_jit */
        float ax = pts[2 * k], ay = pts[2 * k + 1], bx = pts[2 * k + 2], by = pts[2 * k + 3];
        float ux = bx - ax, uy = by - ay, wx = ox - ax, wy = oy - ay;
        float d = cross2(ex, ey, ux, uy);
        float s = cross2(ex, ey, wx, wy) / d;
        float t = cross2(ux, uy, wx, wy) / d;
        if ((s >= 0.0f) && (s <= 1.0f) && (t > 0.0f) && s < best) { best = s; idx = k; }  /* :165-178 first index on ties */
    }
    if (idx < 0 || best > r || best <= 0.0f) {        /* :166-174 */
        out_pt[0] = px + r * ex; out_pt[1] = py + r * ey; out_nrm[0] = 0.0f; out_nrm[1] = 0.0f;
        return 0;
    }
    float sx = pts[2 * idx + 2] - pts[2 * idx], sy = pts[2 * idx + 3] - pts[2 * idx + 1];   /* :181-183 */
    float sl = norm2f(sx, sy);                        /* :184 */
    if (sl < 1e-10f) { out_nrm[0] = 0.0f; out_nrm[1] = 1.0f; }   /* :186-189 */
    else { float tx = sx / sl, ty = sy / sl; out_nrm[0] = -ty; out_nrm[1] = tx; }   /* :191-194 left normal (Q3) */
    out_pt[0] = ox + best * ex; out_pt[1] = oy + best * ey;      /* :196 */
    if (out_seg) *out_seg = idx;
    return 1;
}

void orc_distance_batch(const float* pts, int n, const float* q, int64_t B, float* out) {
    for (int64_t i = 0; i < B; ++i) out[i] = orc_distance(pts, n, q[2 * i], q[2 * i + 1]);
}
void orc_silhouette_distance_batch(const float* pts, int n, const float* q, int64_t B, float* out) {
    for (int64_t i = 0; i < B; ++i) out[i] = orc_silhouette_distance(pts, n, q[2 * i], q[2 * i + 1]);
}
void orc_intersect_batch(const float* pts, int n, const float* q, const float* d, const float* r, int64_t B,
                         float* out_pt, float* out_nrm, uint8_t* out_found, int32_t* out_seg) {
    for (int64_t i = 0; i < B; ++i)
        out_found[i] = (uint8_t)orc_intersect_polylines(pts, n, q[2 * i], q[2 * i + 1], d[2 * i], d[2 * i + 1], r[i],
                                                        out_pt + 2 * i, out_nrm + 2 * i, out_seg + i);
}

/* ------------------------------------------------------------------------------------------
 * Fields: f, alpha, sigma, g as sums of analytic terms or a bilinear table (fp32).
 * These stand in for the reference's user callables (solvers/WoStSolver.py:22,253-256,277-283,295).
 * ---------------------------------------------------------------------------------------- */

static inline float ipowf(float x, int p) { float r = 1.0f; for (int i = 0; i < p; ++i) r *= x; return r; }

typedef struct { float v, gx, gy, l; } jet_t;   /* value, gradient, laplacian */

static inline jet_t jet_mul(jet_t a, jet_t b) {
    jet_t r;
    r.l = a.v * b.l + 2.0f * (a.gx * b.gx + a.gy * b.gy) + b.v * a.l;
    r.gx = a.v * b.gx + b.v * a.gx;
    r.gy = a.v * b.gy + b.v * a.gy;
    r.v = a.v * b.v;
    return r;
}

static jet_t term_jet(const orc_term_t* t, float x, float y) {
    jet_t r;
    if (t->kind == ORC_TERM_SIGMOID_CIRCLE) {
        /* utils.py:123-129 torch_smooth_circle: sigmoid(-k (|x-c| - R)) */
        float ddx = x - t->cx, ddy = y - t->cy;
        float rho = norm2f(ddx, ddy);
        float s = o_smooth_step(t->q * (rho - t->R));
        float s1 = -t->q * s * (1.0f - s);
        float s2 = t->q * t->q * s * (1.0f - s) * (1.0f - 2.0f * s);
        float inv = rho > 0.0f ? 1.0f / rho : 0.0f;
        r.v = t->A * s; r.gx = t->A * s1 * ddx * inv; r.gy = t->A * s1 * ddy * inv;
        r.l = t->A * (s2 + s1 * inv);
        return r;
    }
    r.v = t->A; r.gx = r.gy = r.l = 0.0f;
    if (t->px | t->py) {
        jet_t m;
        float mx = ipowf(x, t->px), my = ipowf(y, t->py);
        float mx1 = t->px ? t->px * ipowf(x, t->px - 1) : 0.0f, my1 = t->py ? t->py * ipowf(y, t->py - 1) : 0.0f;
        float mx2 = t->px > 1 ? t->px * (t->px - 1) * ipowf(x, t->px - 2) : 0.0f;
        float my2 = t->py > 1 ? t->py * (t->py - 1) * ipowf(y, t->py - 2) : 0.0f;
        m.v = mx * my; m.gx = mx1 * my; m.gy = mx * my1; m.l = mx2 * my + mx * my2;
        r = jet_mul(r, m);
    }
    if (t->q != 0.0f) {
        jet_t e; float ddx = x - t->cx, ddy = y - t->cy, d2 = ddx * ddx + ddy * ddy;
        e.v = o_expf(-t->q * d2); e.gx = -2.0f * t->q * ddx * e.v; e.gy = -2.0f * t->q * ddy * e.v;
        e.l = e.v * (4.0f * t->q * t->q * d2 - 4.0f * t->q);
        r = jet_mul(r, e);
    }
    for (int k = 0; k < 2; ++k) {
        int kind = k ? t->t2 : t->t1;
        if (kind == ORC_TRIG_NONE) continue;
        float wx = k ? t->w2x : t->w1x, wy = k ? t->w2y : t->w1y, p = k ? t->p2 : t->p1;
        float a = wx * x + wy * y + p, sn, cs; o_sincosf(a, &sn, &cs);
        jet_t g; float w2 = wx * wx + wy * wy;
        if (kind == ORC_TRIG_SIN) { g.v = sn; g.gx = cs * wx; g.gy = cs * wy; g.l = -sn * w2; }
        else { g.v = cs; g.gx = -sn * wx; g.gy = -sn * wy; g.l = -cs * w2; }
        r = jet_mul(r, g);
    }
    return r;
}

static float term_value(const orc_term_t* t, float x, float y) {
    if (t->kind == ORC_TERM_SIGMOID_CIRCLE) {
        float rho = norm2f(x - t->cx, y - t->cy);
        return t->A * o_smooth_step(t->q * (rho - t->R));
    }
    float v = t->A;
    if (t->px | t->py) v *= ipowf(x, t->px) * ipowf(y, t->py);
    if (t->q != 0.0f) { float ddx = x - t->cx, ddy = y - t->cy; v *= o_expf(-t->q * (ddx * ddx + ddy * ddy)); }
    if (t->t1 != ORC_TRIG_NONE) { float sn, cs; o_sincosf(t->w1x * x + t->w1y * y + t->p1, &sn, &cs); v *= t->t1 == ORC_TRIG_SIN ? sn : cs; }
    if (t->t2 != ORC_TRIG_NONE) { float sn, cs; o_sincosf(t->w2x * x + t->w2y * y + t->p2, &sn, &cs); v *= t->t2 == ORC_TRIG_SIN ? sn : cs; }
    return v;
}

static int field_masked_out(const orc_field_t* f, float x, float y) {
    if (f->mask_kind == ORC_MASK_BOX) return x < f->mask[0] || x > f->mask[1] || y < f->mask[2] || y > f->mask[3];
    if (f->mask_kind == ORC_MASK_DISC) { float ddx = x - f->mask[0], ddy = y - f->mask[1]; return ddx * ddx + ddy * ddy > f->mask[2]; }
    return 0;
}

static void grid_cell(const orc_field_t* f, float x, float y, int* i, int* j, float* tx, float* ty) {
    float fx = (x - f->x0) / f->dx, fy = (y - f->y0) / f->dy;
    fx = fminf(fmaxf(fx, 0.0f), (float)(f->nx - 1));    /* fmaxf/fminf drop NaN: a diverged walker reads the table edge */
    fy = fminf(fmaxf(fy, 0.0f), (float)(f->ny - 1));
    int ii = (int)fx, jj = (int)fy;
    if (ii > f->nx - 2) ii = f->nx - 2;
    if (jj > f->ny - 2) jj = f->ny - 2;
    *i = ii; *j = jj; *tx = fx - (float)ii; *ty = fy - (float)jj;
}

float orc_field_eval(const orc_field_t* f, float x, float y) {
    if (!f) return 0.0f;
    if (field_masked_out(f, x, y)) return f->outside;
    if (f->kind == ORC_FIELD_GRID) {
        int i, j; float tx, ty; grid_cell(f, x, y, &i, &j, &tx, &ty);
        const float* g = f->grid + (int64_t)i * f->ny + j;
        float v00 = g[0], v01 = g[1], v10 = g[f->ny], v11 = g[f->ny + 1];
        float a = v00 + ty * (v01 - v00), b = v10 + ty * (v11 - v10);
        return a + tx * (b - a);
    }
    float v = f->c0;
    for (int k = 0; k < f->n_terms; ++k) v += term_value(&f->terms[k], x, y);
    return v;
}

void orc_field_eval_d(const orc_field_t* f, float x, float y, float* v, float* gx, float* gy, float* lap) {
    *v = *gx = *gy = *lap = 0.0f;
    if (!f) return;
    if (field_masked_out(f, x, y)) { *v = f->outside; return; }
    if (f->kind == ORC_FIELD_GRID) {
        int i, j; float tx, ty; grid_cell(f, x, y, &i, &j, &tx, &ty);
        const float* g = f->grid + (int64_t)i * f->ny + j;
        float v00 = g[0], v01 = g[1], v10 = g[f->ny], v11 = g[f->ny + 1];
        float a = v00 + ty * (v01 - v00), b = v10 + ty * (v11 - v10);
        *v = a + tx * (b - a);
        *gx = (b - a) / f->dx;
        *gy = ((v01 - v00) + tx * ((v11 - v10) - (v01 - v00))) / f->dy;
        return;
    }
    float sv = f->c0, sx = 0.0f, sy = 0.0f, sl = 0.0f;
    for (int k = 0; k < f->n_terms; ++k) { jet_t t = term_jet(&f->terms[k], x, y); sv += t.v; sx += t.gx; sy += t.gy; sl += t.l; }
    *v = sv; *gx = sx; *gy = sy; *lap = sl;
}

void orc_field_eval_batch(const orc_field_t* f, const float* q, int64_t B, float* v, float* gx, float* gy, float* lap) {
    for (int64_t i = 0; i < B; ++i) {
        float a, b, c, d; orc_field_eval_d(f, q[2 * i], q[2 * i + 1], &a, &b, &c, &d);
        /* the value always comes from the value-only path (the one the walk uses) */
        v[i] = orc_field_eval(f, q[2 * i], q[2 * i + 1]);
        if (gx) gx[i] = b;
        if (gy) gy[i] = c;
        if (lap) lap[i] = d;
    }
}

static inline float alpha_at(const orc_params_t* p, float x, float y) { return p->alpha ? orc_field_eval(p->alpha, x, y) : 1.0f; }

/* sigma_prime closure, solvers/WoStSolver.py:88-127 */
float orc_sigma_prime(const orc_params_t* p, float x, float y) {
    if (p->sp_mode == ORC_SP_FIELD) return orc_field_eval(p->sigma_prime, x, y);
    float sg = p->sigma ? orc_field_eval(p->sigma, x, y) : 0.0f;
    if (p->sp_mode == ORC_SP_RATIO) {
        float a = alpha_at(p, x, y); if (a < 1e-8f) a = 1e-8f;       /* :86 clamp */
        return sg / a;                                              /* :102,127 */
    }
    float a = 1.0f, gx = 0.0f, gy = 0.0f, lap = 0.0f;
    if (p->alpha) orc_field_eval_d(p->alpha, x, y, &a, &gx, &gy, &lap);
    if (a < 1e-8f) { a = 1e-8f; gx = gy = lap = 0.0f; }              /* clamp kills the gradient */
    float ratio = sg / a;                                            /* :102 */
    float lapl = lap + 1e-8f;                                        /* utils.py:54 */
    float la = a + 1e-8f;                                            /* :112 log(alpha + 1e-8) */
    float lgx = gx / la, lgy = gy / la;                              /* :114 */
    float n2 = lgx * lgx + lgy * lgy;                                /* :115 */
    float corr = 0.5f * (lapl / a - n2 / 2.0f);                      /* :119 */
    return ratio + corr;                                             /* :121 */
}

/* ------------------------------------------------------------------------------------------
 * Green's-function helpers — solvers/utils.py:5-61 (scipy.special.i0/k0 restated in double)
 * ---------------------------------------------------------------------------------------- */

/* I0 by its power series sum (z^2/4)^k/(k!)^2 (all terms positive; z <= ~40 on this path) */
double orc_i0(double z) {
    double q = 0.25 * z * z, term = 1.0, sum = 1.0;
    for (int k = 1; k < 500; ++k) { term *= q / ((double)k * (double)k); sum += term; if (term < 1e-17 * sum) break; }
    return sum;
}
static double i0_minus_1(double z) {
    double q = 0.25 * z * z, term = 1.0, sum = 0.0;
    for (int k = 1; k < 500; ++k) { term *= q / ((double)k * (double)k); sum += term; if (term < 1e-17 * sum) break; }
    return sum;
}
/* K0(z) = int_0^inf exp(-z cosh t) dt, trapezoid rule (exponentially convergent, h = 1/8) */
double orc_k0(double z) {
    if (z <= 0.0) return INFINITY;
    const double h = 0.125; double sum = 0.5 * exp(-z);
    for (int k = 1; k < 4000; ++k) { double a = z * cosh(k * h); if (a > 745.0) break; sum += exp(-a); }
    return h * sum;
}
/* screenedGreensNorm2D, utils.py:29-44: (1 - 1/I0(R sqrt(sb))) / sb */
double orc_screened_greens_norm(double R, double sb) { return (1.0 / sb) * (1.0 - 1.0 / orc_i0(R * sqrt(sb))); }
/* screenedGreens2D, utils.py:5-26 */
double orc_screened_greens(double r, double R, double sb) {
    double s = sqrt(sb);
    return 1.0 / (2.0 * M_PI) * (orc_k0(r * s) - (orc_k0(R * s) / orc_i0(R * s)) * orc_i0(r * s));
}

/* GreensDistribution2D._refill_cache, utils.py:138-151 */
static void greens_refill(mt_t* np_rng, int n, double* out) {
    const double small_val = 1e-6, max_log = -log(small_val);
    int cnt = 0;
    while (cnt < n) {
        double cand = mt_numpy_uniform(np_rng, small_val, 1.0);      /* :145 */
        double dens = -log(cand);                                    /* :146 */
        if (mt_numpy_uniform(np_rng, 0.0, max_log) < dens) out[cnt++] = cand;   /* :148-149 */
    }
}
/* ScreenedGreensDistribution2D._refill_cache, utils.py:181-195.  The radius passes through a float32
 * tensor (torch.tensor([r,0]).norm()), and scipy's i0/k0 run their float32 loops on it. */
static void screened_refill(mt_t* np_rng, double sb, int n, double* out) {
    const double max_density = orc_screened_greens_norm(1.0, sb);   /* :184 */
    const double s = sqrt(sb);
    const double K0R = orc_k0(1.0 * s), I0R = orc_i0(1.0 * s);      /* :21,23 with R = 1.0 (python float) */
    int cnt = 0;
    while (cnt < n) {
        double cand = mt_numpy_uniform(np_rng, 1e-6, 1.0);           /* :187 */
        float rf = (float)cand; rf = sqrtf(rf * rf + 0.0f * 0.0f);   /* :189-190, utils.py:20 norm in fp32 */
        float zf = rf * (float)s;                                    /* :22 tensor * np.float64 -> fp32 */
        double I0r = (double)(float)orc_i0((double)zf);              /* f->f ufunc loop */
        double K0r = (double)(float)orc_k0((double)zf);
        double dens = fabs(1.0 / (2.0 * M_PI) * (K0r - (K0R / I0R) * I0r));   /* :26,191 */
        if (mt_numpy_uniform(np_rng, 0.0, max_density) < dens) out[cnt++] = cand;   /* :193-194 */
    }
}
void orc_greens_cache_fill(uint64_t seed_numpy, int n, double* out) { mt_t g; mt_seed(&g, (uint32_t)seed_numpy); greens_refill(&g, n, out); }
void orc_screened_cache_fill(uint64_t seed_numpy, double sb, int n, double* out) { mt_t g; mt_seed(&g, (uint32_t)seed_numpy); screened_refill(&g, sb, n, out); }

/* Inverse CDF of the density the screened rejection sampler actually realises:
 * p(rho) ∝ min(|G^sb(rho; R=1)|, envelope) on [1e-6, 1], envelope = screenedGreensNorm2D(1, sb).
 * table[i] = rho at u = i/(n-1). */
void orc_screened_icdf(double sb, int n, float* table) {
    const int M = 16384; const double lo = 1e-6, hi = 1.0;
    double* cdf = (double*)malloc(sizeof(double) * (M + 1));
    double* xs = (double*)malloc(sizeof(double) * (M + 1));
    const double env = orc_screened_greens_norm(1.0, sb);
    /* nodes graded towards 0 (log-singular density): x = lo + (hi-lo) * s^2 */
    double prev = 0.0; cdf[0] = 0.0;
    for (int i = 0; i <= M; ++i) {
        double s = (double)i / M; xs[i] = lo + (hi - lo) * s * s;
        double d = fabs(orc_screened_greens(xs[i], 1.0, sb)); if (d > env) d = env;
        if (i > 0) cdf[i] = cdf[i - 1] + 0.5 * (d + prev) * (xs[i] - xs[i - 1]);
        prev = d;
    }
    double tot = cdf[M]; int j = 0;
    for (int i = 0; i < n; ++i) {
        double u = tot * (double)i / (double)(n - 1);
        while (j < M - 1 && cdf[j + 1] < u) ++j;
        double w = cdf[j + 1] > cdf[j] ? (u - cdf[j]) / (cdf[j + 1] - cdf[j]) : 0.0;
        if (w < 0) w = 0; if (w > 1) w = 1;
        table[i] = (float)(xs[j] + w * (xs[j + 1] - xs[j]));
    }
    free(cdf); free(xs);
}

/* ------------------------------------------------------------------------------------------
 * The walk — solvers/WoStSolver.py:162-316 (_solveUnified)
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    int mode;
    /* MT mode: shared sequential streams + the cycled cache (utils.py:109-117) */
    mt_t* torch_rng; mt_t* np_rng;
    double* cache; int cache_n, cache_idx, cache_filled;
    /* PHILOX mode: per-step outputs */
    uint32_t k0, k1, o[4];
} walk_rng_t;

#define ORC_CACHE_SIZE 10000

static double cache_next(walk_rng_t* g, const orc_params_t* p) {    /* _get_cached_sample :109-117 */
    if (!g->cache_filled || g->cache_idx >= g->cache_n) {
        if (p->delta) screened_refill(g->np_rng, (double)p->sigma_bar, g->cache_n, g->cache);
        else greens_refill(g->np_rng, g->cache_n, g->cache);
        g->cache_idx = 0; g->cache_filled = 1;
    }
    return g->cache[g->cache_idx++];
}

static inline float u24(uint32_t o) { return (float)(o >> 8) * (1.0f / 16777216.0f); }           /* [0,1) */
static inline float u24p(uint32_t o) { return (float)((o >> 8) + 1u) * (1.0f / 16777216.0f); }   /* (0,1] */

/* sigma_bar * screenedGreensNorm2D(r) = 1 - 1/I0(z) and the norm itself */
static void greens_norm_pair(const orc_params_t* p, float r, int r_is_rmin, double rmin_d, int mode, float* gn, float* sbgn) {
    double sb = (double)p->sigma_bar;
    if (mode == ORC_RNG_MT) {
        /* fidelity to the reference's mixed precision (utils.py:43-44 under NEP-50 scalar rules) */
        if (r_is_rmin) {
            double g = orc_screened_greens_norm(rmin_d, sb);        /* python float path: all double */
            *gn = (float)g; *sbgn = (float)(sb * g);
        } else {
            float zf = r * (float)sqrt(sb);
            float I0f = (float)orc_i0((double)zf);
            float b = 1.0f - 1.0f / I0f;
            *gn = (float)(1.0 / sb) * b;
            *sbgn = (float)sb * *gn;
        }
    } else {
        /* the kernel's arithmetic: fp32 throughout, 1 - 1/I0 from the table of include/wost_math.h */
        *sbgn = wm_interior_probability_lookup(g_iprob, r * (float)sqrt(sb));
        *gn = *sbgn * (float)(1.0 / sb);
    }
}

static float run_walk(const orc_params_t* p, walk_rng_t* g, float x0, float y0, uint32_t pidx, uint32_t widx,
                      float* ref_total, int32_t* n_steps, int32_t trace_cap, float* trace, int32_t* trace_len) {
    const int has_neu = p->neu_pts && p->n_neu > 0;
    const int has_src = p->f != NULL;
    const double rmin_d = (double)p->eps / 2.0;                      /* :167 */
    const float rmin = (float)rmin_d, eps = p->eps;
    float x = x0, y = y0;
    float dD = 1.0f;                                                 /* :190 sentinel (Q6) */
    int onB = 0; float nx = 0.0f, ny = 1.0f;                         /* :193-194 */
    float atten = 1.0f, total = 0.0f;
    int steps = 0;
    while (steps < p->max_steps && dD > eps) {                       /* :206 */
        dD = orc_distance(p->dir_pts, p->n_dir, x, y);               /* :208 */
        float dN = INFINITY, r; int r_is_rmin;
        if (has_neu) {
            dN = orc_silhouette_distance(p->neu_pts, p->n_neu, x, y);   /* :211 */
            float m = dN < dD ? dN : dD;                             /* :212 python min/max */
            r_is_rmin = !(m > rmin); r = r_is_rmin ? rmin : m;
        } else { r_is_rmin = !(dD > rmin); r = r_is_rmin ? rmin : dD; }   /* :215 */
        if (trace && steps < trace_cap) { float* t = trace + 4 * steps; t[0] = x; t[1] = y; t[2] = dD; t[3] = dN; }

        float u_theta, u_mu = 0.0f;
        if (g->mode == ORC_RNG_MT) u_theta = mt_torch_rand(g->torch_rng);   /* :226 */
        else if (!has_src && !p->delta) {
            /* Laplace walks need one 32-bit word per step: one Philox block (stream tag 1) serves four steps */
            orc_philox4x32_10(pidx, widx, (uint32_t)steps >> 2, 1u, g->k0, g->k1, g->o); u_theta = u24(g->o[steps & 3]);
        } else { orc_philox4x32_10(pidx, widx, (uint32_t)steps, 0u, g->k0, g->k1, g->o); u_theta = u24(g->o[0]); u_mu = u24(g->o[1]); }
        float theta = (u_theta * 2.0f) * 3.14159274101257324f;       /* :226 fp32 */
        if (onB && has_neu) theta = theta / 2.0f + (p->atan2_fn ? p->atan2_fn(ny, nx) : atan2f(ny, nx));   /* :227-228 (Q2) */
        float dx, dy;                                                /* :230-232 */
        if (p->sincos_fn) p->sincos_fn(theta, &dx, &dy);
        else if (g->mode == ORC_RNG_MT) { dx = cosf(theta); dy = sinf(theta); }
        else wm_sincosf_small(theta, &dy, &dx);

        float qx, qy;                                                /* next_point */
        if (has_neu) {
            float pt[2], nr[2];
            onB = orc_intersect_polylines(p->neu_pts, p->n_neu, x, y, dx, dy, r, pt, nr, NULL);   /* :236 */
            qx = pt[0]; qy = pt[1]; nx = nr[0]; ny = nr[1];
        } else { qx = x + r * dx; qy = y + r * dy; onB = 0; }        /* :238-239 */

        float sx = qx, sy = qy;                                      /* sample_point */
        if (has_src || p->delta) {                                   /* :242 (Q10: sampled also without a source) */
            float rs;
            if (g->mode == ORC_RNG_MT) {
                double ns = cache_next(g, p);                        /* :244 */
                rs = r_is_rmin ? (float)(ns * rmin_d) : (float)ns * r;   /* utils.py:117 */
            } else {
                float rho;
                if (p->delta) {
                    float pos = u24(g->o[2]) * (float)(p->icdf_len - 1); int i = (int)pos;
                    if (i > p->icdf_len - 2) i = p->icdf_len - 2;
                    float fr = pos - (float)i;
                    rho = p->icdf[i] + fr * (p->icdf[i + 1] - p->icdf[i]);
                } else {
                    rho = u24p(g->o[2]) * u24p(g->o[3]);             /* product of two uniforms has pdf -ln(rho) (Q8) */
                    if (rho < 1e-6f) rho = 1e-6f;
                }
                rs = rho * r;
            }
            sx = x + rs * dx; sy = y + rs * dy;                      /* :245 */
            float contrib;
            if (norm2f(sx - x, sy - y) > norm2f(qx - x, qy - y)) {   /* :248-250 */
                sx = qx; sy = qy; contrib = 0.0f;
            } else if (!has_src) {
                contrib = 0.0f;
            } else if (p->delta) {                                   /* :252-254 */
                float gn, sbgn; greens_norm_pair(p, r, r_is_rmin, rmin_d, g->mode, &gn, &sbgn);
                contrib = (orc_field_eval(p->f, sx, sy) * gn / sqrtf(alpha_at(p, sx, sy) * alpha_at(p, x, y))) * atten;
            } else {
                contrib = orc_field_eval(p->f, sx, sy) * (r * r / 4.0f);   /* :256, utils.py:61 */
            }
            if (has_src) { total += contrib; if (ref_total) *ref_total += contrib; }   /* :258 */
        }

        if (p->delta) {                                              /* :271 */
            if (g->mode == ORC_RNG_MT) u_mu = mt_torch_rand(g->torch_rng);   /* :272 */
            float gn, sbgn; greens_norm_pair(p, r, r_is_rmin, rmin_d, g->mode, &gn, &sbgn);   /* :273 */
            if (u_mu > sbgn) {                                       /* :275 */
                atten = atten * sqrtf(alpha_at(p, qx, qy) / alpha_at(p, x, y));   /* :277 */
                x = qx; y = qy;                                      /* :278 */
            } else {
                float sp = orc_sigma_prime(p, sx, sy);               /* :281 */
                float sc = 1.0f - sp / p->sigma_bar; if (0.0f > sc) sc = 0.0f;   /* :282 */
                atten = (atten * sqrtf(alpha_at(p, sx, sy) / alpha_at(p, x, y))) * sc;   /* :283 */
                x = sx; y = sy;                                      /* :284 */
            }
        } else { x = qx; y = qy; }                                   /* :287 */
        ++steps;                                                     /* :291 */
    }
    float bc = p->g ? orc_field_eval(p->g, x, y) : 0.0f;             /* :295 (Q5, Q7) */
    if (p->delta) bc = bc * atten;                                   /* :296-297 */
    total += bc; if (ref_total) *ref_total += bc;                    /* :298 */
    *n_steps = steps;
    if (trace_len) *trace_len = steps < trace_cap ? steps : trace_cap;
    return total;
}

/* ------------------------------------------------------------------------------------------
 * "Physical" mode: textbook Walk on Stars (Sawhney, Miller, Gkioulekas, Crane 2023) for constant coefficients.
 * Not part of the reference (whose mixed-boundary walks leak); shares its primitives.
 * ---------------------------------------------------------------------------------------- */
static float phys_distance(const float* pts, int n, float px, float py, float* cxo, float* cyo) {
    float best = INFINITY; *cxo = px; *cyo = py;
    for (int k = 0; k + 1 < n; ++k) {
        float ax = pts[2 * k], ay = pts[2 * k + 1], bx = pts[2 * k + 2], by = pts[2 * k + 3];
        float ux = bx - ax, uy = by - ay, vx = px - ax, vy = py - ay;
        float t = (vx * ux + vy * uy) / (ux * ux + uy * uy);
        t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
        float cx = (1.0f - t) * ax + t * bx, cy = (1.0f - t) * ay + t * by;
        float q = fmaf(cy - py, cy - py, (cx - px) * (cx - px));
        if (q < best) { best = q; *cxo = cx; *cyo = cy; }
    }
    return sqrtf(best);
}
/* silhouette distance including the closing vertex of a closed loop */
static float phys_silhouette_distance(const float* pts, int n, float px, float py) {
    float best = orc_silhouette_distance(pts, n, px, py);
    if (n >= 4 && pts[0] == pts[2 * n - 2] && pts[1] == pts[2 * n - 1]) {
        float ax = pts[2 * n - 4], ay = pts[2 * n - 3], bx = pts[0], by = pts[1], cx = pts[2], cy = pts[3];
        float c1 = cross2(bx - ax, by - ay, px - ax, py - ay), c2 = cross2(cx - bx, cy - by, px - bx, py - by);
        if (c1 * c2 < 0.0f) { float d = norm2f(bx - px, by - py); if (d < best) best = d; }
    }
    return best;
}
/* first hit by true ray distance t within tmax; returns segment or -1 */
static int phys_ray(const float* pts, int n, float ox, float oy, float ex, float ey, float tmax, float* t_hit) {
    float best = INFINITY; int idx = -1;
    for (int k = 0; k + 1 < n; ++k) {
        float ax = pts[2 * k], ay = pts[2 * k + 1], ux = pts[2 * k + 2] - ax, uy = pts[2 * k + 3] - ay, wx = ox - ax, wy = oy - ay;
        float d = cross2(ex, ey, ux, uy);
        float s = cross2(ex, ey, wx, wy) / d, t = cross2(ux, uy, wx, wy) / d;
        if ((s >= 0.0f) && (s <= 1.0f) && (t > 0.0f) && t < best) { best = t; idx = k; }
    }
    if (idx < 0 || best > tmax) return -1;
    *t_hit = best;
    return idx;
}

static float run_walk_physical(const orc_params_t* p, walk_rng_t* g, float x0, float y0, uint32_t pidx, uint32_t widx,
                               int32_t* n_steps, int32_t trace_cap, float* trace, int32_t* trace_len) {
    const int has_neu = p->neu_pts && p->n_neu > 0, has_src = p->f != NULL;
    const float rmin = (float)((double)p->eps / 2.0), eps = p->eps;
    float x = x0, y = y0, total = 0.0f, phi_in = 0.0f, cx = x0, cy = y0, dD;
    int onB = 0, steps = 0;
    for (;;) {
        dD = phys_distance(p->dir_pts, p->n_dir, x, y, &cx, &cy);
        if (!(steps < p->max_steps && dD > eps)) break;
        float dN = has_neu ? phys_silhouette_distance(p->neu_pts, p->n_neu, x, y) : INFINITY;
        float m = dN < dD ? dN : dD, r = m > rmin ? m : rmin;
        if (trace && steps < trace_cap) { float* t = trace + 4 * steps; t[0] = x; t[1] = y; t[2] = dD; t[3] = dN; }
        orc_philox4x32_10(pidx, widx, (uint32_t)steps, 2u, g->k0, g->k1, g->o);
        /* direction: uniform, or uniform in the hemisphere around the inward normal when sitting on the wall */
        float theta = onB ? phi_in + (u24(g->o[0]) - 0.5f) * 3.14159274101257324f : (u24(g->o[0]) * 2.0f) * 3.14159274101257324f;
        float ex = cosf(theta), ey = sinf(theta);
        if (has_src) {
            float th2 = onB ? phi_in + (u24(g->o[1]) - 0.5f) * 3.14159274101257324f : (u24(g->o[1]) * 2.0f) * 3.14159274101257324f;
            float sx_ = cosf(th2), sy_ = sinf(th2);
            float rho = r * sqrtf(u24p(g->o[2]) * u24p(g->o[3]));        /* rho^2/r^2 has pdf -ln: rho pdf 4 rho ln(r/rho)/r^2 */
            float th; int vis = 1;
            if (has_neu) vis = phys_ray(p->neu_pts, p->n_neu, x, y, sx_, sy_, rho, &th) < 0;
            if (vis) total += orc_field_eval(p->f, x + rho * sx_, y + rho * sy_) * (r * r / 4.0f);
        }
        /* a wall within r + nudge counts as hit, so a free step always ends at least `nudge` short of every wall
         * (fp32 rounding of x + r e is ~1e-7 of the scene scale, the nudge 1e-5) */
        float t_hit; int k = has_neu ? phys_ray(p->neu_pts, p->n_neu, x, y, ex, ey, r + p->phys_nudge, &t_hit) : -1;
        if (k >= 0) {
            float ux = p->neu_pts[2 * k + 2] - p->neu_pts[2 * k], uy = p->neu_pts[2 * k + 3] - p->neu_pts[2 * k + 1];
            float len = norm2f(ux, uy), nx = -uy / len, ny = ux / len;
            if (nx * ex + ny * ey > 0.0f) { nx = -nx; ny = -ny; }        /* face the side the walker came from */
            x = (x + t_hit * ex) + p->phys_nudge * nx; y = (y + t_hit * ey) + p->phys_nudge * ny;   /* sit `nudge` off the wall */
            phi_in = atan2f(ny, nx); onB = 1;
        } else { x = x + r * ex; y = y + r * ey; onB = 0; }
        ++steps;
    }
    total += p->g ? orc_field_eval(p->g, cx, cy) : 0.0f;                /* g at the closest Dirichlet point */
    *n_steps = steps;
    if (trace_len) *trace_len = steps < trace_cap ? steps : trace_cap;
    return total;
}

/* ---- physical mode with variable coefficients (delta tracking done by the book; not in the reference) -------------
 * -div(alpha grad u) + sigma u = f.  With U = sqrt(alpha) u:  lap U - sigma' U = -f / sqrt(alpha)  (sigma' as in
 * WoStSolver.py:88-121), rewritten with a majorant sigma_bar:  lap U - sigma_bar U = -[f / sqrt(alpha) + (sigma_bar - sigma') U].
 * On the star-shaped region St(x, r) with the ball's screened Green's function G (zero on the sphere, zero-Neumann walls):
 *     U(x) = int_{dSt} P U + int_{St} G [f / sqrt(alpha) + (sigma_bar - sigma') U],
 * estimated with one sample per step: a direction e (first hit at distance t <= r: sphere or wall) and a volume point
 * y = x + rho e2 with rho ~ 4 rho ln(r / rho) / r^2 (the Laplace Green's density; the screened one is G = ratio(rho) * G0).
 *   - source:   w * ratio(rho) * r^2/4 * f(y) / sqrt(alpha(y))                      if y is visible from x
 *   - with probability p_v = 1 - 1/I0(c), c = r sqrt(sigma_bar): continue from y with
 *                w *= ratio(rho) * (c^2/4) / p_v * (1 - sigma'(y) / sigma_bar)      (w = 0 if y is not visible)
 *   - else continue from the hit point with  w *= 2 pi Q(t) I0(c)  (= 1 on the sphere, in [1, I0(c)] on a wall), where
 *                2 pi Q(t) = c_t [K1(c_t) + K0(c) I1(c_t) / I0(c)],  c_t = t sqrt(sigma_bar)  is the flux of G through dSt.
 * The radius is capped at 1/sqrt(sigma_bar) so c <= 1 and all weights stay near 1.  Walls must have d(alpha)/dn = 0
 * (otherwise the transformed problem has a Robin condition there).  Bessel functions in double, by quadrature / series. */
static double orc_i1(double z) {
    double q = 0.25 * z * z, term = 1.0, sum = 1.0;
    for (int k = 1; k < 500; ++k) { term *= q / ((double)k * (double)(k + 1)); sum += term; if (term < 1e-17 * sum) break; }
    return 0.5 * z * sum;
}
/* K1(z) = int_0^inf exp(-z cosh t) cosh t dt, trapezoid rule like orc_k0 */
double orc_k1(double z) {
    if (z <= 0.0) return INFINITY;
    const double h = 0.0625; double sum = 0.5 * exp(-z);
    for (int k = 1; k < 8000; ++k) { double ch = cosh(k * h), a = z * ch; if (a > 745.0) break; sum += exp(-a) * ch; }
    return h * sum;
}
/* G_screened(rho; r) / G_laplace(rho; r) */
double orc_phys_green_ratio(double rho, double r, double sb) {
    double s = sqrt(sb), L = log(r / rho);
    if (!(L > 1e-7)) return 1.0 / orc_i0(r * s);
    return (orc_k0(rho * s) - (orc_k0(r * s) / orc_i0(r * s)) * orc_i0(rho * s)) / L;
}
/* 2 pi Q(t) I0(c): weight of a wall hit at distance t inside a ball of radius r */
double orc_phys_wall_weight(double t, double r, double sb) {
    double s = sqrt(sb), ct = t * s, c = r * s;
    if (ct <= 0.0) return orc_i0(c);
    if (ct > c) ct = c;
    return ct * (orc_k1(ct) * orc_i0(c) + orc_k0(c) * orc_i1(ct));
}

/* spatially varying majorant: maximum of the pyramid over the cells the ball touches, at the level whose cells are
 * at least 2r wide */
float orc_majorant_over_ball(const orc_params_t* p, float x, float y, float r) {
    int l = 0, n = 1 << (p->maj_levels - 1), off = 0;
    float cx = p->maj_dx, cy = p->maj_dy;
    while ((cx < 2.0f * r || cy < 2.0f * r) && l < p->maj_levels - 1) { off += n * n; n >>= 1; cx *= 2.0f; cy *= 2.0f; ++l; }
    int i0 = (int)floorf((x - r - p->maj_x0) / cx), i1 = (int)floorf((x + r - p->maj_x0) / cx);
    int j0 = (int)floorf((y - r - p->maj_y0) / cy), j1 = (int)floorf((y + r - p->maj_y0) / cy);
    i0 = i0 < 0 ? 0 : (i0 > n - 1 ? n - 1 : i0); i1 = i1 < 0 ? 0 : (i1 > n - 1 ? n - 1 : i1);
    j0 = j0 < 0 ? 0 : (j0 > n - 1 ? n - 1 : j0); j1 = j1 < 0 ? 0 : (j1 > n - 1 ? n - 1 : j1);
    const float* L = p->majorant + off;
    float m = -1.0f;
    for (int i = i0; i <= i1; ++i) for (int j = j0; j <= j1; ++j) { float v = L[i * n + j]; if (v > m) m = v; }
    return m;
}
/* the largest r <= r0 (by halving, never below rmin) with r^2 M(ball(x, r)) <= 1 */
float orc_majorant_radius(const orc_params_t* p, float x, float y, float r0, float rmin, float* M) {
    float r = r0 > rmin ? r0 : rmin;
    for (int it = 0; it < 48; ++it) {
        *M = orc_majorant_over_ball(p, x, y, r);
        if (r * r * *M <= 1.0f || r <= rmin) break;
        float lo = 1.0f / sqrtf(*M), half = 0.5f * r;
        r = half > lo ? half : lo; r = r > rmin ? r : rmin;
    }
    return r;
}

static float run_walk_physical_delta(const orc_params_t* p, walk_rng_t* g, float x0, float y0, uint32_t pidx, uint32_t widx,
                                     int32_t* n_steps, int32_t trace_cap, float* trace, int32_t* trace_len) {
    const int has_neu = p->neu_pts && p->n_neu > 0, has_src = p->f != NULL;
    const float rmin = (float)((double)p->eps / 2.0), eps = p->eps, rcap = 1.0f / sqrtf(p->sigma_bar);
    float x = x0, y = y0, total = 0.0f, phi_in = 0.0f, cx = x0, cy = y0, dD;
    float w = 1.0f / sqrtf(alpha_at(p, x0, y0));
    int onB = 0, steps = 0;
    uint32_t o2[4];
    for (;;) {
        dD = phys_distance(p->dir_pts, p->n_dir, x, y, &cx, &cy);
        if (!(steps < p->max_steps && dD > eps && w != 0.0f)) break;
        float dN = has_neu ? phys_silhouette_distance(p->neu_pts, p->n_neu, x, y) : INFINITY;
        float m = dN < dD ? dN : dD, r, sbf = p->sigma_bar;
        if (p->maj_levels > 0) {
            float M; r = orc_majorant_radius(p, x, y, m, rmin, &M);
            float floor_ = 1e-8f / (r * r); sbf = M > floor_ ? M : floor_;
        } else { m = m < rcap ? m : rcap; r = m > rmin ? m : rmin; }
        const double sb = (double)sbf;
        if (trace && steps < trace_cap) { float* t = trace + 4 * steps; t[0] = x; t[1] = y; t[2] = dD; t[3] = dN; }
        orc_philox4x32_10(pidx, widx, (uint32_t)steps, 2u, g->k0, g->k1, g->o);
        orc_philox4x32_10(pidx, widx, (uint32_t)steps, 3u, g->k0, g->k1, o2);
        float theta = onB ? phi_in + (u24(g->o[0]) - 0.5f) * 3.14159274101257324f : (u24(g->o[0]) * 2.0f) * 3.14159274101257324f;
        float ex = cosf(theta), ey = sinf(theta);
        float th2 = onB ? phi_in + (u24(g->o[1]) - 0.5f) * 3.14159274101257324f : (u24(g->o[1]) * 2.0f) * 3.14159274101257324f;
        float sx_ = cosf(th2), sy_ = sinf(th2);
        float rho = r * sqrtf(u24p(g->o[2]) * u24p(g->o[3]));
        float th; int vis = 1;
        if (has_neu) vis = phys_ray(p->neu_pts, p->n_neu, x, y, sx_, sy_, rho, &th) < 0;
        const float yx = x + rho * sx_, yy = y + rho * sy_;
        const float ratio = (float)orc_phys_green_ratio((double)rho, (double)r, sb);
        if (has_src && vis) total += w * (orc_field_eval(p->f, yx, yy) * (ratio * (r * r / 4.0f)) / sqrtf(alpha_at(p, yx, yy)));
        const double c = (double)r * sqrt(sb), i0c = orc_i0(c), pv = i0_minus_1(c) / i0c;
        if ((double)u24(o2[0]) < pv) {                                       /* null-collision inside the star */
            if (!vis) w = 0.0f;
            else {
                const float sp = orc_sigma_prime(p, yx, yy);
                w = (w * (ratio * (float)(0.25 * c * c / pv))) * (1.0f - sp / sbf);
                x = yx; y = yy; onB = 0;
            }
        } else {
            float t_hit; int k = has_neu ? phys_ray(p->neu_pts, p->n_neu, x, y, ex, ey, r + p->phys_nudge, &t_hit) : -1;
            if (k >= 0) {
                float ux = p->neu_pts[2 * k + 2] - p->neu_pts[2 * k], uy = p->neu_pts[2 * k + 3] - p->neu_pts[2 * k + 1];
                float len = norm2f(ux, uy), nx = -uy / len, ny = ux / len;
                if (nx * ex + ny * ey > 0.0f) { nx = -nx; ny = -ny; }
                w = w * (float)orc_phys_wall_weight((double)t_hit, (double)r, sb);
                x = (x + t_hit * ex) + p->phys_nudge * nx; y = (y + t_hit * ey) + p->phys_nudge * ny;
                phi_in = atan2f(ny, nx); onB = 1;
            } else { x = x + r * ex; y = y + r * ey; onB = 0; }
        }
        ++steps;
    }
    if (p->g) total += w * (orc_field_eval(p->g, cx, cy) * sqrtf(alpha_at(p, cx, cy)));
    *n_steps = steps;
    if (trace_len) *trace_len = steps < trace_cap ? steps : trace_cap;
    return total;
}

int orc_solve(const orc_params_t* p, const float* pts, int64_t n_pts,
              double* mean, double* m2, float* walk_vals, int64_t* steps_total, int32_t* walk_steps,
              int64_t n_trace, int32_t trace_cap, float* trace, int32_t* trace_len) {
    const int64_t W = p->n_walks;
    int64_t steps_sum = 0;
    iprob_init();
    if (p->compat_mode == 1 && (p->rng_mode != ORC_RNG_PHILOX || (p->delta && !(p->sigma_bar > 0.0f)))) return -1;   /* physical: Philox only */
    if (p->rng_mode == ORC_RNG_MT) {
        /* one sequential stream over all points and walks, like the reference; libm / torch elementary functions */
        const int libm_before = g_libm; g_libm = 1;
        mt_t torch_rng, np_rng; mt_seed(&torch_rng, (uint32_t)p->seed); mt_seed(&np_rng, (uint32_t)p->seed_numpy);
        walk_rng_t g; memset(&g, 0, sizeof g);
        g.mode = ORC_RNG_MT; g.torch_rng = &torch_rng; g.np_rng = &np_rng;
        g.cache = (double*)malloc(sizeof(double) * ORC_CACHE_SIZE); g.cache_n = ORC_CACHE_SIZE;   /* :168,173 */
        for (int64_t pi = 0; pi < n_pts; ++pi) {
            float ref_total = 0.0f;                                  /* :183 point_total (fp32, across walks) */
            double s = 0.0, s2 = 0.0;
            for (int64_t w = 0; w < W; ++w) {
                int64_t flat = pi * W + w; int32_t ns;
                float* tr = (trace && flat < n_trace) ? trace + (size_t)flat * trace_cap * 4 : NULL;
                float v = run_walk(p, &g, pts[2 * pi], pts[2 * pi + 1], 0, 0, &ref_total, &ns, trace_cap, tr,
                                   (trace_len && flat < n_trace) ? trace_len + flat : NULL);
                if (walk_vals) walk_vals[flat] = v;
                if (walk_steps) walk_steps[flat] = ns;
                steps_sum += ns; s += v; s2 += (double)v * v;
            }
            mean[pi] = (double)(ref_total / (float)W);               /* :311 */
            if (m2) { double mu = s / (double)W; m2[pi] = s2 - (double)W * mu * mu; if (m2[pi] < 0) m2[pi] = 0; }
        }
        free(g.cache);
        g_libm = libm_before;
    } else {
        int nt = p->n_threads;
#ifdef _OPENMP
        if (nt <= 0) nt = omp_get_max_threads();
#else
        nt = 1;
#endif
        (void)nt;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt) reduction(+ : steps_sum)
        for (int64_t pi = 0; pi < n_pts; ++pi) {
            walk_rng_t g; memset(&g, 0, sizeof g);
            g.mode = ORC_RNG_PHILOX; g.k0 = (uint32_t)p->seed; g.k1 = (uint32_t)(p->seed >> 32);
            double s = 0.0;
            float* vals = walk_vals ? walk_vals + pi * W : (float*)malloc(sizeof(float) * (size_t)W);
            for (int64_t w = 0; w < W; ++w) {
                int64_t flat = pi * W + w; int32_t ns;
                float* tr = (trace && flat < n_trace) ? trace + (size_t)flat * trace_cap * 4 : NULL;
                int32_t* tl = (trace_len && flat < n_trace) ? trace_len + flat : NULL;
                float v = p->compat_mode == 1 && p->delta
                    ? run_walk_physical_delta(p, &g, pts[2 * pi], pts[2 * pi + 1], (uint32_t)(p->point_index_base + pi), (uint32_t)(p->walk_offset + w), &ns, trace_cap, tr, tl)
                    : p->compat_mode == 1
                    ? run_walk_physical(p, &g, pts[2 * pi], pts[2 * pi + 1], (uint32_t)(p->point_index_base + pi), (uint32_t)(p->walk_offset + w), &ns, trace_cap, tr, tl)
                    : run_walk(p, &g, pts[2 * pi], pts[2 * pi + 1], (uint32_t)(p->point_index_base + pi), (uint32_t)(p->walk_offset + w), NULL, &ns, trace_cap, tr, tl);
                vals[w] = v; if (walk_steps) walk_steps[flat] = ns;
                steps_sum += ns; s += v;
            }
            double mu = s / (double)W, q = 0.0;
            for (int64_t w = 0; w < W; ++w) { double d = (double)vals[w] - mu; q += d * d; }
            mean[pi] = mu; if (m2) m2[pi] = q;
            if (!walk_vals) free(vals);
        }
    }
    if (steps_total) *steps_total = steps_sum;
    return 0;
}
