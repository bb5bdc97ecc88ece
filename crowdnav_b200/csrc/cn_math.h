/*
 * cn_math.h -- the numeric specification of the crowd-navigation step.
 *
 * Every floating-point primitive the env-step uses is defined here ONCE, in
 * plain C, out of operations that IEEE-754 rounds identically on an x86 host
 * and on an sm_100a device: + - * / sqrtf fmaf, int<->float conversions and
 * 32/64-bit integer arithmetic.  No libm / libdevice transcendental is ever
 * called, so `gcc -ffp-contract=off` and `nvcc -fmad=false` give the same
 * bits; that is what makes "poses bit-exact on the integer grid" (and in
 * practice the whole observation) true by construction, not by tolerance.
 *
 * The header is compiled three ways:
 *   - nvcc (device)  : the CUDA step kernel (crowdnav_b200/csrc/cn_step.cu)
 *   - gcc  (host)    : the CPU oracle (oracle/cn_oracle.c) -- test-only
 *   - g++  (host)    : CPU unit tests of the primitives against libm
 *
 * Only PRIMITIVES live here.  The algorithms (LiDAR cast, waypoint logic,
 * risk block, reward) are written independently in the kernel and in the
 * oracle; the oracle restates the reference line by line, the kernel is a
 * warp-parallel design.
 *
 * Accuracy (checked in tests/test_math_primitives.py against float64 libm):
 *   cn_sincos_bin  <= 1.2e-7 abs      cn_atan2 <= 4e-7 abs
 *   cn_exp         <= 3e-7 rel on [-8, 8]
 */
#ifndef CN_MATH_H
#define CN_MATH_H

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define CN_HD __host__ __device__ __forceinline__
/* the larger primitives: inlined by default.  Building with -DCN_NOINLINE_BIG turns them into real calls
 * (smaller code); measured on B200 that is slower (c2 24.0 vs 20.8 us, c3 66.9 vs 56.3 us), so it stays off. */
#if defined(CN_NOINLINE_BIG)
#define CN_HD_BIG static __host__ __device__ __noinline__
#else
#define CN_HD_BIG __host__ __device__ __forceinline__
#endif
#else
#define CN_HD static inline
#define CN_HD_BIG static inline
#endif

/* ---- fixed-point grid ---------------------------------------------------
 * positions: int32 in units of 2^-24 m  (range +-128 m, resolution 6e-8 m)
 * angles   : uint32 binary angle, 2*pi / 2^32 rad per unit (wrap is free)
 */
#define CN_GRID        5.9604644775390625e-08f /* 2^-24 */
#define CN_INV_GRID    16777216.0f             /* 2^24  */
#define CN_BIN2RAD     1.4629180792671596e-09f /* 2*pi / 2^32 */
#define CN_RAD2BIN     683565275.57643158f     /* 2^32 / (2*pi) */
#define CN_PI          3.14159265358979323846f
#define CN_PIO2        1.57079632679489661923f
#define CN_PIO4        0.78539816339744830962f
#define CN_TWO_PI      6.28318530717958647692f
#define CN_INV_TWO_PI  0.15915494309189533577f

/* ---- bit casts and float->int with one definition of rounding ---------- */
CN_HD uint32_t cn_f2bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
CN_HD float cn_bits2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
/* round-to-nearest-even float -> int32 (caller guarantees |x| < 2^31) */
CN_HD int32_t cn_f2i(float x) {
#if defined(__CUDA_ARCH__)
    return __float2int_rn(x);
#else
    return (int32_t)lrintf(x);
#endif
}
/* round-to-nearest-even float -> int64 */
CN_HD int64_t cn_f2ll(float x) {
#if defined(__CUDA_ARCH__)
    return __float2ll_rn(x);
#else
    return (int64_t)llrintf(x);
#endif
}

/* ---- rounding helpers ---------------------------------------------------
 * The reference rounds in float64: np.around(x, 3) = rint(x * 1000) / 1000
 * (half to even) and Python-2 round(x, n) (half away from zero).  In fp32 the
 * product x * 1000 is itself rounded, which would create ties (and 1e-3 flips
 * against the reference) that the exact product does not have.  So the nearest
 * integer is taken of the EXACT product: k = rint(fl(x*c)), then the residual
 * e = x*c - k from one fma (exact to 2^-24 relative) moves k by one when the
 * rounded product landed on the wrong side of a half.
 */
CN_HD float cn_round_ha(float x) {          /* half away from zero */
    float t = truncf(x);
    float d = x - t;                       /* exact */
    if (fabsf(d) >= 0.5f) t += (x < 0.0f) ? -1.0f : 1.0f;
    return t;
}
CN_HD float cn_fix_scaled(float x, float c, float k) {
    float e = fmaf(x, c, -k);
    if (e > 0.5f) k += 1.0f;
    else if (e < -0.5f) k -= 1.0f;
    return k;
}
CN_HD float cn_rint_scaled(float x, float c)  { return cn_fix_scaled(x, c, rintf(x * c)); }
CN_HD float cn_round_scaled(float x, float c) { return cn_fix_scaled(x, c, cn_round_ha(x * c)); }
/* k / 1000 and k / 100 for integer-valued |k| <= 2^24.  q0 = k * RN(1/c) is
 * within 1 ulp; Markstein's fma correction then yields exactly RN(k / c), i.e.
 * the same bits as an IEEE division (checked exhaustively in
 * tests/test_math_primitives.py) at 3 instructions instead of ~12. */
CN_HD float cn_div1000(float k) { float q = k * 0.001f; return fmaf(fmaf(-q, 1000.0f, k), 0.001f, q); }
CN_HD float cn_div100(float k)  { float q = k * 0.01f;  return fmaf(fmaf(-q, 100.0f, k), 0.01f, q); }
/* np.around(x, 3) */
CN_HD_BIG float cn_np_round3(float x) { return cn_div1000(cn_rint_scaled(x, 1000.0f)); }
/* Python-2 round(x, 3) / round(x, 2) */
CN_HD_BIG float cn_py_round3(float x) { return cn_div1000(cn_round_scaled(x, 1000.0f)); }
CN_HD_BIG float cn_py_round2(float x) { return cn_div100(cn_round_scaled(x, 100.0f)); }

/* ---- trigonometry on binary angles --------------------------------------
 * Range reduction is exact integer arithmetic (quadrant = top 2 bits after a
 * 45-degree bias); the polynomials are the classic single-precision minimax
 * kernels on [-pi/4, pi/4].
 */
CN_HD void cn_sincos_bin_body(uint32_t a, float* s_out, float* c_out) {
    uint32_t q = (a + 0x20000000u) >> 30;
    int32_t  r = (int32_t)(a - (q << 30));          /* [-2^29, 2^29) */
    float x = (float)r * CN_BIN2RAD;
    float z = x * x;
    float ps = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    float s = fmaf(ps * z, x, x);
    float pc = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    float c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
    switch (q & 3u) {
        case 0:  *s_out =  s; *c_out =  c; break;
        case 1:  *s_out =  c; *c_out = -s; break;
        case 2:  *s_out = -s; *c_out = -c; break;
        default: *s_out = -c; *c_out =  s; break;
    }
}

#if defined(__CUDACC__) && defined(CN_NOINLINE_BIG)
/* as a real call the two results travel in registers (a struct by value), not through the stack */
typedef struct { float s, c; } cn_sincos_t;
static __host__ __device__ __noinline__ cn_sincos_t cn_sincos_bin_call(uint32_t a) {
    cn_sincos_t r; cn_sincos_bin_body(a, &r.s, &r.c); return r;
}
CN_HD void cn_sincos_bin(uint32_t a, float* s_out, float* c_out) {
    cn_sincos_t r = cn_sincos_bin_call(a); *s_out = r.s; *c_out = r.c;
}
#else
CN_HD_BIG void cn_sincos_bin(uint32_t a, float* s_out, float* c_out) { cn_sincos_bin_body(a, s_out, c_out); }
#endif

/* radians -> binary angle, any finite |x| < 2^31 / RAD2BIN handled by int64 wrap */
CN_HD uint32_t cn_rad2bin(float x) {
    return (uint32_t)(uint64_t)cn_f2ll(x * CN_RAD2BIN);
}
/* binary angle -> radians in [-pi, pi) */
CN_HD float cn_bin2rad(uint32_t a) { return (float)(int32_t)a * CN_BIN2RAD; }

/* sin/cos of a float radian argument (|x| up to a few hundred) */
CN_HD void cn_sincos_rad(float x, float* s_out, float* c_out) {
    float k = rintf(x * CN_INV_TWO_PI);
    float r = fmaf(-k, 6.28318548202514648f, x);        /* hi part of 2*pi (float) */
    r = fmaf(-k, -1.74845553146951715e-07f, r);         /* 2*pi - hi */
    cn_sincos_bin(cn_rad2bin(r), s_out, c_out);
}

/* atan2(y, x) in [-pi, pi]; atan2(0, 0) = 0.  One division: t = min/max in
 * [0, 1], degree-15 odd minimax polynomial (|err| <= 1.3e-7), octant fix-up. */
CN_HD_BIG float cn_atan2(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    if (mx == 0.0f) return 0.0f;
    float t = mn / mx;
    float z = t * t;
    float p = fmaf(-4.0545530866e-03f, z, 2.1862907262e-02f);
    p = fmaf(p, z, -5.5912254479e-02f);
    p = fmaf(p, z, 9.6421920913e-02f);
    p = fmaf(p, z, -1.3908627529e-01f);
    p = fmaf(p, z, 1.9946565254e-01f);
    p = fmaf(p, z, -3.3329860750e-01f);
    p = fmaf(p, z, 9.9999933557e-01f);
    float r = p * t;
    if (ay > ax) r = CN_PIO2 - r;
    if (x < 0.0f) r = CN_PI - r;
    return (y < 0.0f) ? -r : r;
}

/* exp(x) for |x| <= 80: Cody-Waite reduction + degree-6 Taylor on |r| <= ln2/2 */
CN_HD_BIG float cn_exp(float x) {
    float n = rintf(x * 1.44269504088896341f);
    float r = fmaf(-n, 0.693145751953125f, x);
    r = fmaf(-n, 1.42860682030941723e-06f, r);
    float p = fmaf(1.3888889e-3f, r, 8.3333333e-3f);
    p = fmaf(p, r, 4.1666668e-2f);
    p = fmaf(p, r, 1.6666667e-1f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int32_t e = cn_f2i(n);
    return cn_bits2f(cn_f2bits(p) + ((uint32_t)e << 23));
}

/* ---- Philox2x32-10 counter-based RNG (Salmon et al. 2011) ----------------
 * Stateless: the stream position is (global env id, episode, step, pedestrian),
 * so a world's random numbers do not depend on how envs are sharded over GPUs.
 * Two 32-bit outputs per call = one (vx, vy) draw.
 */
typedef struct { uint32_t v[2]; } cn_u32x2;

CN_HD_BIG cn_u32x2 cn_philox2x32(uint32_t c0, uint32_t c1, uint32_t k0) {
    for (int i = 0; i < 10; ++i) {
        uint64_t p = (uint64_t)0xD256D193u * c0;
        uint32_t n0 = (uint32_t)(p >> 32) ^ k0 ^ c1;
        c1 = (uint32_t)p;
        c0 = n0;
        k0 += 0x9E3779B9u;
    }
    cn_u32x2 r; r.v[0] = c0; r.v[1] = c1;
    return r;
}
/* counter/key packing used by the env: one draw per (env, episode, step, pedestrian, purpose) */
CN_HD cn_u32x2 cn_env_rand(uint32_t seed_lo, uint32_t seed_hi, uint32_t gid, uint32_t episode,
                           uint32_t step, uint32_t ped, uint32_t purpose) {
    uint32_t c1 = (step & 0xFFFFu) | ((ped & 0xFFu) << 16) | (purpose << 24);
    uint32_t key = seed_lo ^ (seed_hi * 0x85EBCA6Bu) ^ (episode * 0xC2B2AE35u);
    return cn_philox2x32(gid, c1, key);
}
/* uniform in [0, 1) from the top 24 bits */
CN_HD float cn_u01(uint32_t bits) { return (float)(bits >> 8) * 5.9604644775390625e-08f; }
/* uniform in [-amp, amp) */
CN_HD float cn_usym(uint32_t bits, float amp) { return fmaf(cn_u01(bits), 2.0f, -1.0f) * amp; }

#endif /* CN_MATH_H */
