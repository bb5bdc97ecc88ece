/*
 * cn_math64.h -- float64 primitives of the `risk_faithful` perception block
 * (the reference's own LiDAR segmentation / tracker / collision cone,
 * environment_stage_1_nobonus.py:270-907).  That block is a chain of DISCRETE
 * decisions on float64 values (gradient == 0, IoU rounded to 3 dp > 0, round(x, 3)
 * of hit points), so it is computed in float64 exactly like CPython does:
 * + - * / sqrt fma are IEEE on x86 and on sm_100a, the two transcendentals
 * (sin / cos of the hit-point angle, UTL:119-121) are defined here once.
 * Compiled by nvcc (device + host) and gcc (oracle); `-fmad=false` /
 * `-ffp-contract=off` keep the operation order as written.
 *
 * Only PRIMITIVES live here (same rule as cn_math.h).
 */
#ifndef CN_MATH64_H
#define CN_MATH64_H

#include "cn_math.h"

#if defined(__CUDACC__)
#define CN64_TABLE static __device__ __constant__ const
#define CN64_HOST_TABLE static const
#else
#define CN64_TABLE static const
#endif

/* sin and cos of a double radian argument, |a| < 1e5.  Cody-Waite reduction by pi/2 in two parts
 * (fdlibm's pio2_1 / pio2_1t split), then the classic minimax kernels on [-pi/4, pi/4].
 * |error| <= 2e-16 (tests/test_faithful.py checks against libm). */
CN_HD_BIG void cn_sincos64(double a, double* s_out, double* c_out) {
    double k = rint(a * 6.36619772367581382433e-01);
    double r = a - k * 1.57079632673412561417e+00;
    r = r - k * 6.07710050650619224932e-11;
    double z = r * r;
    double ps = -2.50507602534068634195e-08 + z * 1.58969099521155010221e-10;
    ps = 2.75573137070700676789e-06 + z * ps;
    ps = -1.98412698298579493134e-04 + z * ps;
    ps = 8.33333333332248946124e-03 + z * ps;
    ps = -1.66666666666666324348e-01 + z * ps;
    double s = r + (r * z) * ps;
    double pc = 2.08757232129817482790e-09 + z * -1.13596475577881948265e-11;
    pc = -2.75573143513906633035e-07 + z * pc;
    pc = 2.48015872894767294178e-05 + z * pc;
    pc = -1.38888888888741095749e-03 + z * pc;
    pc = 4.16666666666666019037e-02 + z * pc;
    double c = (1.0 - 0.5 * z) + (z * z) * pc;
    long long q = (long long)k;
    switch ((int)(q & 3)) {
        case 0:  *s_out =  s; *c_out =  c; break;
        case 1:  *s_out =  c; *c_out = -s; break;
        case 2:  *s_out = -s; *c_out = -c; break;
        default: *s_out = -c; *c_out =  s; break;
    }
}

/* Python round(x, 3) as an integer number of thousandths: the decimal nearest to the EXACT binary value of x
 * (CPython rounds the exact value: floatobject.c double_round), ties away from zero (Python 2).  The exact
 * product x * 1000 = r + e with r the rounded product and e one fma residual. */
CN_HD long long cn_py_round3_k64(double x) {
    double r = x * 1000.0;
    double e = fma(x, 1000.0, -r);
    double n = floor(r);
    double d = r - n;                       /* exact */
    int up;
    if (d > 0.5) up = 1;
    else if (d < 0.5) up = 0;
    else if (e > 0.0) up = 1;
    else if (e < 0.0) up = 0;
    else up = (x > 0.0);                    /* true tie: away from zero */
    return (long long)n + up;
}
/* the double CPython returns for that decimal: the correctly rounded k / 1000 */
CN_HD double cn_milli64(long long k) {
    const double kd = (double)k;
    if (k > 33554432ll || k < -33554432ll) return kd / 1000.0;
    /* Markstein: q = k * RN(1/1000) is within an ulp, one fma correction gives RN(k / 1000) -- the same bits as the
     * IEEE division for every |k| <= 2^25 (exhaustive check: tests/test_faithful.py) at 3 instructions instead of ~20 */
    const double q = kd * 0.001;
    return fma(fma(-q, 1000.0, kd), 0.001, q);
}
/* the same for a caller that guarantees |k| <= 2^25 */
CN_HD double cn_milli64_small(int32_t k) {
    const double kd = (double)k;
    const double q = kd * 0.001;
    return fma(fma(-q, 1000.0, kd), 0.001, q);
}
CN_HD double cn_py_round3_64(double x) { return cn_milli64(cn_py_round3_k64(x)); }
/* np.around(x, 3) in float64: rint(x * 1000) / 1000 (numpy multiplies, rints, divides) */
CN_HD double cn_np_round3_64(double x) { return rint(x * 1000.0) / 1000.0; }
/* math.hypot as sqrt(x^2 + y^2) (differs from libm's by at most an ulp; used where 1e-16 cannot flip a decision) */
CN_HD double cn_hypot64(double x, double y) { return sqrt(x * x + y * y); }
/* a config float that was typed as a short decimal (0.6f, 0.12f, 0.15f ...) -> the double of that decimal */
CN_HD double cn_dec64(float f) { return rint((double)f * 1.0e6) / 1.0e6; }

/* shapely's Point.buffer(r) ring: 64 vertices at angle -k*pi/32 (UTL:256, GEOS default resolution 16);
 * cos / sin of those angles as data */
CN64_TABLE double CNF_RING_COS[64] = {0x1.0000000000000p+0, 0x1.fd88da3d12526p-1, 0x1.f6297cff75cb0p-1, 0x1.e9f4156c62ddap-1, 0x1.d906bcf328d46p-1, 0x1.c38b2f180bdb1p-1, 0x1.a9b66290ea1a3p-1, 0x1.8bc806b151741p-1, 0x1.6a09e667f3bcdp-1, 0x1.44cf325091dd6p-1, 0x1.1c73b39ae68c9p-1, 0x1.e2b5d3806f63ep-2, 0x1.87de2a6aea964p-2, 0x1.294062ed59f05p-2, 0x1.8f8b83c69a60dp-3, 0x1.917a6bc29b438p-4, 0x1.1a62633145c07p-54, -0x1.917a6bc29b42fp-4, -0x1.8f8b83c69a608p-3, -0x1.294062ed59f02p-2, -0x1.87de2a6aea962p-2, -0x1.e2b5d3806f63cp-2, -0x1.1c73b39ae68c6p-1, -0x1.44cf325091dd5p-1, -0x1.6a09e667f3bccp-1, -0x1.8bc806b151741p-1, -0x1.a9b66290ea1a4p-1, -0x1.c38b2f180bdb0p-1, -0x1.d906bcf328d46p-1, -0x1.e9f4156c62ddap-1, -0x1.f6297cff75cb0p-1, -0x1.fd88da3d12525p-1, -0x1.0000000000000p+0, -0x1.fd88da3d12526p-1, -0x1.f6297cff75cb0p-1, -0x1.e9f4156c62ddbp-1, -0x1.d906bcf328d47p-1, -0x1.c38b2f180bdb1p-1, -0x1.a9b66290ea1a5p-1, -0x1.8bc806b151742p-1, -0x1.6a09e667f3bcep-1, -0x1.44cf325091ddap-1, -0x1.1c73b39ae68c8p-1, -0x1.e2b5d3806f63fp-2, -0x1.87de2a6aea96dp-2, -0x1.294062ed59f07p-2, -0x1.8f8b83c69a619p-3, -0x1.917a6bc29b421p-4, -0x1.a79394c9e8a0ap-53, 0x1.917a6bc29b407p-4, 0x1.8f8b83c69a60cp-3, 0x1.294062ed59f00p-2, 0x1.87de2a6aea967p-2, 0x1.e2b5d3806f63ap-2, 0x1.1c73b39ae68c5p-1, 0x1.44cf325091dd7p-1, 0x1.6a09e667f3bcbp-1, 0x1.8bc806b15173ep-1, 0x1.a9b66290ea1a3p-1, 0x1.c38b2f180bdafp-1, 0x1.d906bcf328d44p-1, 0x1.e9f4156c62ddap-1, 0x1.f6297cff75cafp-1, 0x1.fd88da3d12526p-1};
CN64_TABLE double CNF_RING_SIN[64] = {0x0.0p+0, -0x1.917a6bc29b42cp-4, -0x1.8f8b83c69a60ap-3, -0x1.294062ed59f05p-2, -0x1.87de2a6aea963p-2, -0x1.e2b5d3806f63bp-2, -0x1.1c73b39ae68c8p-1, -0x1.44cf325091dd6p-1, -0x1.6a09e667f3bccp-1, -0x1.8bc806b151741p-1, -0x1.a9b66290ea1a3p-1, -0x1.c38b2f180bdb0p-1, -0x1.d906bcf328d46p-1, -0x1.e9f4156c62ddbp-1, -0x1.f6297cff75cb0p-1, -0x1.fd88da3d12525p-1, -0x1.0000000000000p+0, -0x1.fd88da3d12526p-1, -0x1.f6297cff75cb0p-1, -0x1.e9f4156c62ddbp-1, -0x1.d906bcf328d46p-1, -0x1.c38b2f180bdb1p-1, -0x1.a9b66290ea1a5p-1, -0x1.8bc806b151742p-1, -0x1.6a09e667f3bcdp-1, -0x1.44cf325091dd6p-1, -0x1.1c73b39ae68c8p-1, -0x1.e2b5d3806f63fp-2, -0x1.87de2a6aea965p-2, -0x1.294062ed59f06p-2, -0x1.8f8b83c69a617p-3, -0x1.917a6bc29b43cp-4, -0x1.1a62633145c07p-53, 0x1.917a6bc29b42bp-4, 0x1.8f8b83c69a60ep-3, 0x1.294062ed59f01p-2, 0x1.87de2a6aea961p-2, 0x1.e2b5d3806f63bp-2, 0x1.1c73b39ae68c6p-1, 0x1.44cf325091dd4p-1, 0x1.6a09e667f3bccp-1, 0x1.8bc806b15173ep-1, 0x1.a9b66290ea1a3p-1, 0x1.c38b2f180bdb0p-1, 0x1.d906bcf328d44p-1, 0x1.e9f4156c62ddap-1, 0x1.f6297cff75cafp-1, 0x1.fd88da3d12526p-1, 0x1.0000000000000p+0, 0x1.fd88da3d12526p-1, 0x1.f6297cff75cb0p-1, 0x1.e9f4156c62ddbp-1, 0x1.d906bcf328d45p-1, 0x1.c38b2f180bdb1p-1, 0x1.a9b66290ea1a5p-1, 0x1.8bc806b151740p-1, 0x1.6a09e667f3bcep-1, 0x1.44cf325091ddap-1, 0x1.1c73b39ae68c8p-1, 0x1.e2b5d3806f640p-2, 0x1.87de2a6aea96ep-2, 0x1.294062ed59f08p-2, 0x1.8f8b83c69a61bp-3, 0x1.917a6bc29b425p-4};

#define CN64_DEG2RAD 0x1.1df46a2529d39p-6   /* math.radians: pi / 180 */

#endif /* CN_MATH64_H */
