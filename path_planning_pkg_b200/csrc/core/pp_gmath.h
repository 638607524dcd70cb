// Binary32 sinf / cosf / atanf / atan2f / acosf with the results of glibc 2.39 on x86-64 (the libm the reference
// links against on the benchmark box), restated as PP_HD code so the device returns the SAME BITS as the stock
// reference build: the reference's float transcendentals (Dubins.cpp:23-33, :185, :218-244, :259-285, :299 and
// Grid3D.cpp:213) feed g and f, and one ulp there can reorder two open nodes.
//
// glibc is not part of the reference tree; what is restated here is its published algorithm, pinned by exhaustive
// comparison against the installed libm.so.6 (tests/cpp/gmath_check.c: all 2^32 arguments of sinf, cosf, atanf, acosf,
// and >= 10^9 random + structured pairs of atan2f, zero mismatches):
//   * sinf / cosf / sincosf : sysdeps/ieee754/flt-32/s_sincosf.h (the ARM optimized-routines kernel): reduction by
//     pi/2 in double, degree-7 / degree-8 polynomials in double, one rounding to float.  x86-64 glibc selects the
//     __sinf_fma / __cosf_fma ifunc variants on every FMA + AVX2 CPU; those are the same C compiled with -mfma -mavx2,
//     i.e. every `a + b*c` of the kernel is ONE fused multiply-add.  The fma() calls below are exactly the
//     vfmadd / vfnmadd instructions of those variants; on a CPU without FMA glibc would round twice and differ.
//   * atanf (sysdeps/ieee754/flt-32/s_atanf.c), atan2f (e_atan2f.c), acosf (e_acosf.c): the fdlibm float kernels,
//     plain float + - * / sqrt with no contraction (these have no ifunc variants).
// Coefficients are the IEEE bit patterns found in the library's .rodata.
#ifndef PP_GMATH_H
#define PP_GMATH_H

#include "pp_defs.h"

#ifdef __CUDA_ARCH__
#define PP_G_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PP_G_F2U(x) __float_as_uint(x)
#define PP_G_U2F(u) __uint_as_float(u)
#else
#define PP_G_FMA(a, b, c) fma((a), (b), (c))
static inline uint32_t pp_g_f2u(float x) { union { float f; uint32_t u; } v; v.f = x; return v.u; }
static inline float pp_g_u2f(uint32_t u) { union { float f; uint32_t u; } v; v.u = u; return v.f; }
#define PP_G_F2U(x) pp_g_f2u(x)
#define PP_G_U2F(u) pp_g_u2f(u)
#endif

// ---- sinf / cosf ---------------------------------------------------------------------------------
#define PP_G_HPI_INV 10680707.430881744   /* 2/pi * 2^24 */
#define PP_G_HPI     1.5707963267948966
#define PP_G_PI63    3.4061215800865545e-19
#define PP_G_C0 1.0
#define PP_G_C1 -0.49999999725108224
#define PP_G_C2 0.041666623324344516
#define PP_G_C3 -0.001388676379437604
#define PP_G_C4 2.4390450703564542e-05
#define PP_G_S1 -0.16666654943701084
#define PP_G_S2 0.008332178146138854
#define PP_G_S3 -0.00019517298981385725

// sinf_poly (s_sincosf.h) with table p = __sincosf_table[neg]: table[1] is table[0] with the cosine coefficients negated
PP_HD float pp_g_sin_poly(double x, double x2)
{
    double x3 = x * x2;
    double s1 = PP_G_FMA(x2, PP_G_S3, PP_G_S2);
    double x7 = x3 * x2;
    double s = PP_G_FMA(x3, PP_G_S1, x);
    return (float)PP_G_FMA(x7, s1, s);
}

PP_HD float pp_g_cos_poly(double x2, bool neg)
{
    const double c0 = neg ? -PP_G_C0 : PP_G_C0, c1 = neg ? -PP_G_C1 : PP_G_C1, c2 = neg ? -PP_G_C2 : PP_G_C2;
    const double c3 = neg ? -PP_G_C3 : PP_G_C3, c4 = neg ? -PP_G_C4 : PP_G_C4;
    double x4 = x2 * x2;
    double d1 = PP_G_FMA(x2, c1, c0);
    double d2 = PP_G_FMA(x2, c4, c3);
    double x6 = x4 * x2;
    double c = PP_G_FMA(x4, c2, d1);
    return (float)PP_G_FMA(x6, d2, c);
}

// reduce_large: |x| >= 120, x finite.  4/pi as overlapping 32-bit words (__inv_pio4 of s_sincosf.c).  Out of line: never
// reached by planner arguments (headings and tangent angles of a few pi), so its table stays off the hot path's frame.
PP_HD_NOINLINE_FN double pp_g_reduce_large(uint32_t xi, int* np)
{
    const uint32_t pp_g_inv_pio4[24] = {
        0xa2u, 0xa2f9u, 0xa2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u,
        0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u,
        0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u };
    const uint32_t* arr = &pp_g_inv_pio4[(xi >> 26) & 15];
    const int shift = (int)((xi >> 23) & 7);
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    uint64_t res0 = (uint64_t)(uint32_t)(xi * arr[0]);
    uint64_t res1 = (uint64_t)xi * arr[4];
    uint64_t res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    uint64_t n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    double x = (double)(int64_t)res0;
    *np = (int)n;
    return x * PP_G_PI63;
}

// sign[n & 3] of the table: +1, -1, -1, +1
PP_HD double pp_g_quadrant_sign(int n) { return (((n & 3) == 1) || ((n & 3) == 2)) ? -1.0 : 1.0; }

// want_cos = false: sinf(y); true: cosf(y)
PP_HD float pp_g_sincos(float y, bool want_cos)
{
    double x = (double)y;
    const uint32_t iy = PP_G_F2U(y);
    const uint32_t top = (iy >> 20) & 0x7ffu;          // abstop12
    if (top < 0x3f4u)                                  // |y| < pi/4
    {
        double x2 = x * x;
        if (top < 0x398u)                              // |y| < 2^-12
            return want_cos ? 1.0f : y;
        return want_cos ? pp_g_cos_poly(x2, false) : pp_g_sin_poly(x, x2);
    }
    int n;
    int sign_off = 0;
    if (top < 0x42fu)                                  // |y| < 120: reduce_fast
    {
        double r = x * PP_G_HPI_INV;
        n = ((int32_t)r + 0x800000) >> 24;
        x = PP_G_FMA(-(double)n, PP_G_HPI, x);
    }
    else if (top < 0x7f8u)
    {
        x = pp_g_reduce_large(iy, &n);
        sign_off = (int)(iy >> 31);
    }
    else return y - y;                                 // inf / NaN -> NaN (__math_invalidf)
    const int q = n + sign_off;
    const double s = pp_g_quadrant_sign(q);
    const bool neg = (q & 2) != 0;
    const double x2 = x * x;
    const bool odd = ((n & 1) != 0) != want_cos;       // sinf: sinf_poly(.., n); cosf: sinf_poly(.., n ^ 1)
    return odd ? pp_g_cos_poly(x2, neg) : pp_g_sin_poly(x * s, x2);
}

PP_HD_NOINLINE_FN float pp_g_sinf(float x) { return pp_g_sincos(x, false); }
PP_HD_NOINLINE_FN float pp_g_cosf(float x) { return pp_g_sincos(x, true); }

// (sinf(y), cosf(y)) with ONE argument reduction: both results are the two polynomials of the same reduced argument, swapped by
// the quadrant (what glibc's own sincosf does; the reference's `sin(h); cos(h)` pairs compile to it, Dubins.cpp:23-33).  Same bits
// as pp_g_sinf / pp_g_cosf by construction -- checked over all arguments by tests/cpp/gmath_check.cpp -- at half the instructions
// and, since the search kernel is bound by instruction fetch, one hot routine instead of two.
PP_HD_NOINLINE_FN void pp_g_sincosf(float y, float* sn, float* cs)
{
    double x = (double)y;
    const uint32_t iy = PP_G_F2U(y);
    const uint32_t top = (iy >> 20) & 0x7ffu;
    if (top < 0x3f4u)
    {
        if (top < 0x398u) { *sn = y; *cs = 1.0f; return; }
        const double x2 = x * x;
        *sn = pp_g_sin_poly(x, x2); *cs = pp_g_cos_poly(x2, false);
        return;
    }
    int n;
    int sign_off = 0;
    if (top < 0x42fu)
    {
        double r = x * PP_G_HPI_INV;
        n = ((int32_t)r + 0x800000) >> 24;
        x = PP_G_FMA(-(double)n, PP_G_HPI, x);
    }
    else if (top < 0x7f8u)
    {
        x = pp_g_reduce_large(iy, &n);
        sign_off = (int)(iy >> 31);
    }
    else { *sn = y - y; *cs = y - y; return; }
    const int q = n + sign_off;
    const double s = pp_g_quadrant_sign(q);
    const bool neg = (q & 2) != 0;
    const double x2 = x * x;
    const float a = pp_g_cos_poly(x2, neg), b = pp_g_sin_poly(x * s, x2);
    const bool odd = (n & 1) != 0;
    *sn = odd ? a : b;
    *cs = odd ? b : a;
}

// ---- atanf (s_atanf.c) ------------------------------------------------------------------------------
PP_HD float pp_g_atanf_core(float x)
{
    const float atanhi[4] = { 4.636476e-01f, 7.853981e-01f, 9.827937e-01f, 1.5707963e+00f };
    const float atanlo[4] = { 5.0121582e-09f, 3.7748947e-08f, 3.4473217e-08f, 7.5497894e-08f };
    const float aT0 = 3.3333334e-01f, aT1 = -2.e-01f, aT2 = 1.4285715e-01f, aT3 = -1.11111104e-01f,
                aT4 = 9.090887e-02f, aT5 = -7.691876e-02f, aT6 = 6.661073e-02f, aT7 = -5.8335703e-02f,
                aT8 = 4.976878e-02f, aT9 = -3.653157e-02f, aT10 = 1.628582e-02f;
    const uint32_t hx = PP_G_F2U(x);
    const uint32_t ix = hx & 0x7fffffffu;
    int id;
    if (ix >= 0x4c000000u)                             // |x| >= 2^25
    {
        if (ix > 0x7f800000u) return x + x;            // NaN
        if ((int32_t)hx > 0) return atanhi[3] + atanlo[3];
        return -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000u)                              // |x| < 0.4375
    {
        if (ix < 0x31000000u) return x;                // |x| < 2^-29
        id = -1;
    }
    else
    {
        // the four argument reductions of s_atanf.c, written as numerator / denominator so that the IEEE division (a long
        // instruction sequence on the device) is emitted once; the operands are the ones of the original expressions
        x = fabsf(x);
        float num, den;
        if (ix < 0x3f980000u)                          // |x| < 1.1875
        {
            if (ix < 0x3f300000u) { id = 0; num = 2.0f * x - 1.0f; den = 2.0f + x; }
            else { id = 1; num = x - 1.0f; den = x + 1.0f; }
        }
        else
        {
            if (ix < 0x401c0000u) { id = 2; num = x - 1.5f; den = 1.0f + 1.5f * x; }
            else { id = 3; num = -1.0f; den = x; }
        }
        x = num / den;
    }
    float z = x * x;
    float w = z * z;
    float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return ((int32_t)hx < 0) ? -z : z;
}

PP_HD_NOINLINE_FN float pp_g_atanf(float x) { return pp_g_atanf_core(x); }

// ---- atan2f (e_atan2f.c) ----------------------------------------------------------------------------
PP_HD_NOINLINE_FN float pp_g_atan2f(float y, float x)
{
    const float tiny = 1.e-30f, pi_o_4 = 7.853982e-01f, pi_o_2 = 1.5707964e+00f, pi = 3.1415927e+00f,
                pi_lo = -8.742278e-08f;
    const uint32_t hx = PP_G_F2U(x), hy = PP_G_F2U(y);
    const uint32_t ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;            // NaN
    if (hx == 0x3f800000u) return pp_g_atanf_core(y);                  // x == 1.0
    const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);     // 2*sign(x) + sign(y)
    if (iy == 0u)
    {
        switch (m)
        {
            case 0: case 1: return y;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (ix == 0u) return ((int32_t)hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000u)
    {
        if (iy == 0x7f800000u)
        {
            switch (m)
            {
                case 0: return pi_o_4 + tiny;
                case 1: return -pi_o_4 - tiny;
                case 2: return 3.0f * pi_o_4 + tiny;
                default: return -3.0f * pi_o_4 - tiny;
            }
        }
        switch (m)
        {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (iy == 0x7f800000u) return ((int32_t)hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = ((int)iy - (int)ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if ((int32_t)hx < 0 && k < -60) z = 0.0f;
    else z = pp_g_atanf_core(fabsf(y / x));
    switch (m)
    {
        case 0: return z;
        case 1: return PP_G_U2F(PP_G_F2U(z) ^ 0x80000000u);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

// ---- acosf (e_acosf.c) -------------------------------------------------------------------------------
PP_HD_NOINLINE_FN float pp_g_acosf(float x)
{
    const float pi = 3.1415925e+00f, pio2_hi = 1.5707963e+00f, pio2_lo = 7.5497894e-08f;
    const float pS0 = 1.6666667e-01f, pS1 = -3.2556581e-01f, pS2 = 2.0121253e-01f, pS3 = -4.0055536e-02f,
                pS4 = 7.91535e-04f, pS5 = 3.479331e-05f;
    const float qS1 = -2.403395e+00f, qS2 = 2.0209458e+00f, qS3 = -6.88284e-01f, qS4 = 7.7038154e-02f;
    const uint32_t hx = PP_G_F2U(x);
    const uint32_t ix = hx & 0x7fffffffu;
    if (ix == 0x3f800000u)
    {
        if ((int32_t)hx > 0) return 0.0f;
        return pi + 2.0f * pio2_lo;
    }
    if (ix > 0x3f800000u) return (x - x) / (x - x);    // |x| > 1 or NaN -> NaN
    if (ix < 0x3f000000u)                              // |x| < 0.5
    {
        if (ix <= 0x32800000u) return pio2_hi + pio2_lo;
        float z = x * x;
        float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        float r = p / q;
        return pio2_hi - (x - (pio2_lo - r * x));
    }
    if ((int32_t)hx < 0)                               // x < -0.5
    {
        float z = (1.0f + x) * 0.5f;
        float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        float s = sqrtf(z);
        float r = p / q;
        float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    }
    float z = (1.0f - x) * 0.5f;                       // x > 0.5
    float s = sqrtf(z);
    float df = PP_G_U2F(PP_G_F2U(s) & 0xfffff000u);
    float c = (z - df * df) / (s + df);
    float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    float r = p / q;
    float w = r * s + c;
    return 2.0f * (df + w);
}

// math policy of the Dubins code (pp_dubins.h) and the APF angle: bit-for-bit the stock glibc 2.39 x86-64 libm
struct PPMathGlibc
{
    PP_HD static float sin(float x) { return pp_g_sinf(x); }
    PP_HD static float cos(float x) { return pp_g_cosf(x); }
    PP_HD static void  sincos(float x, float& s, float& c) { pp_g_sincosf(x, &s, &c); }
    PP_HD static float atan2(float y, float x) { return pp_g_atan2f(y, x); }
    PP_HD static float acos(float x) { return pp_g_acosf(x); }
};

#endif
