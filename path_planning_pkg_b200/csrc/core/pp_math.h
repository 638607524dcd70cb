// Scalar arithmetic of the hot path with the reference's float<->double promotions reproduced
// operation by operation.  Compile WITHOUT fused multiply-add contraction
// (nvcc -fmad=false, g++ -ffp-contract=off): the reference is built for baseline x86-64, which has
// no FMA, so every a*b+c below must round twice.
//
// Float transcendentals: the reference calls glibc's sinf/cosf/atan2f/acosf, which are not correctly
// rounded and are libm-version dependent.  Two policies:
//   PPMathGlibc  (default)  pp_gmath.h: the algorithms of glibc 2.39 / x86-64 restated operation by operation;
//                           bit-identical to the stock reference build on this image (oracle/_ref/libref_oracle.so)
//   PPMathPinned (-DPP_MATH_PINNED builds lib/libpp_b200_pinned.so)  evaluate in double, round once to float:
//                           libm-version independent; oracle/_ref/libref_oracle_crm.so is the reference built
//                           against the same definition
#ifndef PP_MATH_H
#define PP_MATH_H

#include "pp_defs.h"
#include "pp_gmath.h"

#define PP_PI    3.14159265358979323846   /* M_PI   */
#define PP_PI_2  1.57079632679489661923   /* M_PI_2 */

PP_HD_NOINLINE_FN float pp_pin_sinf(float x) { return (float)sin((double)x); }
PP_HD_NOINLINE_FN float pp_pin_cosf(float x) { return (float)cos((double)x); }
PP_HD_NOINLINE_FN float pp_pin_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
PP_HD_NOINLINE_FN float pp_pin_acosf(float x) { return (float)acos((double)x); }

struct PPMathPinned
{
    PP_HD static float sin(float x) { return pp_pin_sinf(x); }
    PP_HD static float cos(float x) { return pp_pin_cosf(x); }
    PP_HD static void  sincos(float x, float& s, float& c) { s = pp_pin_sinf(x); c = pp_pin_cosf(x); }
    PP_HD static float atan2(float y, float x) { return pp_pin_atan2f(y, x); }
    PP_HD static float acos(float x) { return pp_pin_acosf(x); }
};

// the policy of the reference-parity (EXACT) code paths
#ifdef PP_MATH_PINNED
typedef PPMathPinned PPMathExact;
#else
typedef PPMathGlibc PPMathExact;
#endif
PP_HD float pp_sinf(float x) { return PPMathExact::sin(x); }
PP_HD float pp_cosf(float x) { return PPMathExact::cos(x); }
PP_HD void  pp_sincosf(float x, float& s, float& c) { PPMathExact::sincos(x, s, c); }
PP_HD float pp_atan2f(float y, float x) { return PPMathExact::atan2(y, x); }
PP_HD float pp_acosf(float x) { return PPMathExact::acos(x); }

// glibc hypotf == (float)sqrt((double)x*x + (double)y*y) (checked on 2e8 random pairs)
PP_HD float pp_hypotf(float x, float y)
{
    double dx = (double)x, dy = (double)y;
    return (float)sqrt(dx * dx + dy * dy);
}

// wrap_pi<float> (common.h:15-29): fmod in double rounded to float, comparisons against M_PI in
// double, correction in double rounded to float.
// fmod(a, 2*pi) bit-exactly, without the generic (slow, iterative) fmod for the arguments that occur:
// |a| < 2pi -> a; 2pi <= |a| < 4pi -> a -/+ 2pi, exact by Sterbenz' lemma; otherwise the library fmod.
PP_HD_NOINLINE_FN double pp_fmod_2pi_slow(double a) { return fmod(a, 2 * PP_PI); }
PP_HD double pp_fmod_2pi(double a)
{
    double m = fabs(a);
    if (m < 2 * PP_PI) return a;
    if (m < 4 * PP_PI) return (a < 0) ? a + 2 * PP_PI : a - 2 * PP_PI;
    return pp_fmod_2pi_slow(a);
}

PP_HD float pp_wrap_pi(float angle)
{
    float w = (float)pp_fmod_2pi((double)angle);
    if ((double)w > PP_PI) return (float)((double)w - 2 * PP_PI);
    if ((double)w < -PP_PI) return (float)((double)w + 2 * PP_PI);
    return w;
}

// wrap_pi<double> (common.h:15-29): what `wrap_pi(theta - M_PI_2)` instantiates when theta is
// float (Dubins.cpp:354, :376, :385 ...), result then rounded to float by the assignment.
PP_HD double pp_wrap_pi_d(double angle)
{
    double w = pp_fmod_2pi(angle);
    if (w > PP_PI) return w - 2 * PP_PI;
    if (w < -PP_PI) return w + 2 * PP_PI;
    return w;
}

// get_heading_index<float> (common.h:9-12, :32-36): roundf(h/p)*p in float, (+M_PI)/p in double,
// truncation.  Returns `bins` (one past the end) for headings >= pi - p/2.
PP_HD int pp_heading_index(float heading, float precision)
{
    float rounded = roundf(heading / precision) * precision;
    return (int)(((double)rounded + PP_PI) / (double)precision);
}

#endif
