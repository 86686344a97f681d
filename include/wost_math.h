/*
 * wost_math.h — the elementary functions of the walk (sin/cos, exp, 1 - 1/I0) as explicit fp32 arithmetic.
 *
 * Why: the walk kernel (CUDA, -fmad=false) and the CPU oracle (gcc, -ffp-contract=off) must take the SAME walks for
 * the same Philox key, bit for bit.  A walk amplifies a one-ulp difference in a direction cosine ~2x per step, so
 * "the same libm to within an ulp" is not enough: CUDA's sincosf/expf/cyl_bessel_i0f and glibc's cosf/sinf/expf
 * round differently now and then.  Everything here is built from IEEE-754 operations that are correctly rounded on
 * both sides (+, -, *, /, sqrt, fma, round-to-nearest-even, int<->float conversion), written out in one fixed order, so
 * both compilers produce identical results.  This header plays the role of libm for both; it is not part of the
 * reference's algorithm (the reference calls torch.cos/sin/exp and scipy.special.i0, solvers/WoStSolver.py:230-231,
 * utils.py:123-129, solvers/utils.py:43) and the oracle's RNG-replay pinning mode keeps calling libm / torch instead.
 *
 * Accuracy (checked in tests/test_oracle_pinned.py against libm in double): sin/cos <= 1.5 ulp for |a| <= 1e5,
 * exp <= 1.5 ulp, interior probability <= 4e-7 absolute.
 *
 * Usable from C99, C++ host code, nvcc device code and NVRTC (no includes needed there).
 */
#ifndef WOST_MATH_H
#define WOST_MATH_H

#if !defined(__CUDACC_RTC__)
#include <math.h>
#include <string.h>
#endif
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define WM_FN __host__ __device__ __forceinline__
#else
#define WM_FN static inline
#endif

#if defined(__CUDA_ARCH__)
#define WM_AS_FLOAT(i) __int_as_float(i)
#else
WM_FN float wm_as_float_(int i) { float f; memcpy(&f, &i, sizeof f); return f; }
#define WM_AS_FLOAT(i) wm_as_float_(i)
#endif

/* sin and cos of a (radians).  Cody-Waite reduction by pi/2 in three fused steps, minimax polynomials on
 * [-pi/4, pi/4] (least-squares/Lawson fit of (sin r / r - 1)/r^2 and (cos r - 1)/r^2, tools/gen_math_coeffs.py).
 * |a| > 1e5 (never reached by walk directions; possible for user trig fields): the argument is first reduced with
 * the exact double-precision remainder by the double nearest 2 pi, then treated like a small one. */
WM_FN void wm_sincosf_reduced_(float a, float* sn, float* cs) {
    const float j = rintf(a * 0.636619747f);                     /* nearest multiple of pi/2 */
    const int q = (int)j;
    float r = fmaf(j, -1.57079601e+00f, a);
    r = fmaf(j, -3.13916473e-07f, r);
    r = fmaf(j, -5.39030253e-15f, r);
    const float t = r * r;
    float ps = fmaf(t, -1.95829773e-04f, 8.33272561e-03f);
    ps = fmaf(t, ps, -1.66666642e-01f);
    const float s = fmaf(r * t, ps, r);                          /* r + r^3 S(r^2) */
    float pc = fmaf(t, 2.44577168e-05f, -1.38875434e-03f);
    pc = fmaf(t, pc, 4.16666493e-02f);
    pc = fmaf(t, pc, -0.5f);
    const float c = fmaf(t, pc, 1.0f);                           /* 1 + r^2 C(r^2) */
    float so = (q & 1) ? c : s, co = (q & 1) ? s : c;
    if (q & 2) so = -so;
    if ((q + 1) & 2) co = -co;
    *sn = so; *cs = co;
}
/* any argument (user trig fields) */
WM_FN void wm_sincosf(float a, float* sn, float* cs) {
    if (!(fabsf(a) <= 1.0e5f)) a = (float)fmod((double)a, 6.283185307179586);   /* NaN / inf -> NaN */
    wm_sincosf_reduced_(a, sn, cs);
}
/* |a| <= 1e5 guaranteed by the caller (walk directions: |theta| < 10) */
WM_FN void wm_sincosf_small(float a, float* sn, float* cs) { wm_sincosf_reduced_(a, sn, cs); }

/* e^x.  x = j ln 2 + r, |r| <= ln2/2; e^r = 1 + r + r^2 E(r) (degree-4 fit); scaled by 2^j in two exact-power
 * multiplications so results in the denormal range are rounded once.  Saturates to +inf above 88.7228 and to 0
 * below -103.98 (smallest denormal is e^-103.28): callers' "far field is exactly zero" shortcuts rely on that. */
WM_FN float wm_expf(float x) {
    if (x > 88.7228f) return WM_AS_FLOAT(0x7f800000);
    if (x < -103.98f) return 0.0f;
    const float j = rintf(x * 1.44269502f);
    float r = fmaf(j, -6.93145752e-01f, x);
    r = fmaf(j, -1.42860677e-06f, r);
    float p = fmaf(r, 1.39421492e-03f, 8.36377591e-03f);
    p = fmaf(r, p, 4.16663624e-02f);
    p = fmaf(r, p, 1.66665733e-01f);
    p = fmaf(r, p, 0.5f);
    p = fmaf(r * r, p, r) + 1.0f;
    const int ji = (int)j, j1 = ji / 2, j2 = ji - j1;
    return (p * WM_AS_FLOAT((j1 + 127) << 23)) * WM_AS_FLOAT((j2 + 127) << 23);
}

/* sigmoid(-a) = 1/(1 + e^a), the smooth step of utils.py:123-129.  Beyond a = 87 the quotient is below 2^-125 and is
 * taken as exactly 0 (keeps inf and denormals out of the reciprocal). */
WM_FN float wm_smooth_step(float a) { return a > 87.0f ? 0.0f : 1.0f / (1.0f + wm_expf(a)); }

/* sigma_bar * screenedGreensNorm2D(r, sigma_bar) = 1 - 1/I0(z), z = r sqrt(sigma_bar) (solvers/utils.py:29-44).
 * z < 3: (I0 - 1)/I0 from eight terms of I0 - 1 = sum_{k>=1} (z^2/4)^k / (k!)^2 (truncation 3e-9 relative at z = 3), so
 * nothing cancels for small z.  3 <= z <= 21: 1/I0(z) = e^-z sqrt(z) h(1/z) with a degree-7 fit of the slowly varying
 * h(w) = 1 / (e^-z I0(z) sqrt(z)) -> sqrt(2 pi) (tools/gen_math_coeffs.py; 1e-7 absolute on the probability).
 * z > 21: 1/I0 < 2^-25, the difference rounds to 1. */
WM_FN float wm_interior_probability(float z) {
    if (z > 21.0f) return 1.0f;
    if (z < 3.0f) {
        const float q = 0.25f * z * z;
        float p = fmaf(q, 6.15118755e-10f, 3.93680022e-08f);         /* 1/(8!)^2, 1/(7!)^2 */
        p = fmaf(q, p, 1.92901234e-06f);                             /* 1/(6!)^2 */
        p = fmaf(q, p, 6.94444461e-05f);                             /* 1/(5!)^2 */
        p = fmaf(q, p, 1.73611112e-03f);                             /* 1/(4!)^2 */
        p = fmaf(q, p, 2.77777780e-02f);                             /* 1/(3!)^2 */
        p = fmaf(q, p, 0.25f);
        p = fmaf(q, p, 1.0f);
        const float m1 = q * p;
        return m1 / (1.0f + m1);
    }
    const float w = 1.0f / z;
    float h = fmaf(w, -131.514343f, 169.010208f);
    h = fmaf(w, h, -79.2524719f);
    h = fmaf(w, h, 17.0182533f);
    h = fmaf(w, h, -2.13383055f);
    h = fmaf(w, h, -0.0156128379f);
    h = fmaf(w, h, -0.31692341f);
    h = fmaf(w, h, 2.5066669f);
    return 1.0f - (wm_expf(-z) * sqrtf(z)) * h;
}

/* The same probability from a table: the walk evaluates it every step, and the three branches above diverge within a warp
 * (measured on the DCR scene: ~12 % of the kernel's instructions).  table[i] = wm_interior_probability(i h), h =
 * ZMAX / (N - 1); linear interpolation is within 3e-8 of the closed form (h^2 max|P''| / 8), below its own 1e-7.  Kernel and
 * oracle build the table with the function below (pure fp32, deterministic) and read it with the same lookup. */
#define WM_IPROB_N 32768
#define WM_IPROB_ZMAX 21.0f
#if !defined(__CUDACC_RTC__)
static inline void wm_interior_probability_table(float* table) {
    for (int i = 0; i < WM_IPROB_N; ++i) table[i] = wm_interior_probability((float)i * (WM_IPROB_ZMAX / (float)(WM_IPROB_N - 1)));
}
#endif
WM_FN float wm_interior_probability_lookup(const float* table, float z) {
    if (!(z < WM_IPROB_ZMAX)) return 1.0f;                       /* 1/I0 < 2^-25 (and NaN / inf) */
    const float pos = z * ((float)(WM_IPROB_N - 1) / WM_IPROB_ZMAX);
    int i = (int)pos;
    i = i < 0 ? 0 : (i > WM_IPROB_N - 2 ? WM_IPROB_N - 2 : i);
    const float fr = pos - (float)i;
    const float t0 = table[i], t1 = table[i + 1];
    return t0 + fr * (t1 - t0);
}

#endif /* WOST_MATH_H */
