// FP32 SIMT elementary functions for the K-POP heuristic path (north_star (d): "the Dubins non-holonomic heuristic,
// in FP32 SIMT").  Plain single-precision polynomial kernels (Cody-Waite reduction + minimax polynomials of the
// classic single-precision libm literature) written with IEEE +, -, *, /, sqrt and rint only and evaluated in a fixed
// order, so -- compiled without FMA contraction -- they give the same bits on the device and in the CPU restatement
// (oracle/port/fmath.inc).  Accuracy ~1e-7 absolute on the ranges that occur (|x| < 100 for sin / cos), which keeps
// the Dubins lengths within 1e-6 relative of the reference's glibc-float evaluation (tests/test_gpu_kpop.py).
#ifndef PP_FMATH_H
#define PP_FMATH_H

#include "pp_defs.h"

#define PP_FM_PI    3.14159274101257324219f   /* float(M_PI)   */
#define PP_FM_PI_2  1.57079637050628662109f   /* float(M_PI_2) */
#define PP_FM_PI_4  0.78539818525314331055f   /* float(M_PI_4) */

PP_HD float pp_fm_nan()
{
    union { unsigned u; float f; } c; c.u = 0x7fc00000u; return c.f;
}

// sin and cos of x: n = rint(x * 2/pi), r = x - n*pi/2 in three exact-product steps, polynomials on [-pi/4, pi/4]
PP_HD void pp_fm_sincos(float x, float& s, float& c)
{
    const float q = rintf(x * 0.636619772367581343f);
    const int n = (int)q;
    float r = x - q * 1.5703125f;
    r = r - q * 4.837512969970703125e-4f;
    r = r - q * 7.54978995489188216e-8f;
    const float z = r * r;
    const float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    const float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    switch (n & 3)
    {
        case 0:  s = sp;  c = cp;  break;
        case 1:  s = cp;  c = -sp; break;
        case 2:  s = -sp; c = -cp; break;
        default: s = -cp; c = sp;  break;
    }
}

// atan of a non-negative argument
PP_HD float pp_fm_atan_pos(float x)
{
    float y = 0.0f;
    if (x > 2.414213562373095f) { y = PP_FM_PI_2; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y = PP_FM_PI_4; x = (x - 1.0f) / (x + 1.0f); }
    const float z = x * x;
    return y + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x);
}

PP_HD float pp_fm_atan2(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax == 0.0f) a = (ay == 0.0f) ? 0.0f : PP_FM_PI_2;
    else a = pp_fm_atan_pos(ay / ax);
    if (x < 0.0f) a = PP_FM_PI - a;
    return (y < 0.0f) ? -a : a;
}

PP_HD float pp_fm_asin_core(float a)      // |a| <= 0.5
{
    const float z = a * a;
    return ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * a + a;
}

// acos; NaN outside [-1, 1] (the RSL / LSR candidates of circles closer than 2r never win the fold)
PP_HD float pp_fm_acos(float x)
{
    if (!(fabsf(x) <= 1.0f)) return pp_fm_nan();
    if (x > 0.5f) return 2.0f * pp_fm_asin_core(sqrtf(0.5f * (1.0f - x)));
    if (x < -0.5f) return PP_FM_PI - 2.0f * pp_fm_asin_core(sqrtf(0.5f * (1.0f + x)));
    return PP_FM_PI_2 - pp_fm_asin_core(x);
}

// math policy of the Dubins code (pp_dubins.h): the K-POP flavour
struct PPMathFp32
{
    PP_HD static float sin(float x) { float s, c; pp_fm_sincos(x, s, c); return s; }
    PP_HD static float cos(float x) { float s, c; pp_fm_sincos(x, s, c); return c; }
    PP_HD static float atan2(float y, float x) { return pp_fm_atan2(y, x); }
    PP_HD static float acos(float x) { return pp_fm_acos(x); }
};

#endif
