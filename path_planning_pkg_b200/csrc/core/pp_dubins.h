// Dubins CSC paths (RSR, RSL, LSR, LSL): shortest length = non-holonomic heuristic, and the sampled
// path of the analytic "Dubins shot".  Behaviour follows lib/Dubins.cpp of the reference:
//   length  : Dubins.cpp:19-69 (candidate order RSR,RSL,LSR,LSL, strict '<', NaN never wins)
//   params  : Dubins.cpp:180-323 (one +-2*pi correction only; M_PI_2 sums in double, stored as float)
//   sampling: Dubins.cpp:326-563 (theta accumulated in float; headings through wrap_pi<double>)
// The four candidates are independent, so a warp evaluates (successor, candidate) pairs on separate
// lanes and folds them in candidate order; this header provides the per-candidate pieces.
#ifndef PP_DUBINS_H
#define PP_DUBINS_H

#include "pp_math.h"

enum { PP_RSR = 0, PP_RSL = 1, PP_LSR = 2, PP_LSL = 3 };

struct PPDubinsCenters
{
    float srx, sry, slx, sly;   // start right / left circle centres
    float grx, gry, glx, gly;   // goal right / left circle centres
};

// Dubins.cpp:23-34.  M = math policy (PPMathExact: reference parity, pp_math.h; PPMathFp32: K-POP mode)
template <class M = PPMathExact>
PP_HD void pp_dubins_centers(float r, float sx, float sy, float sh, float gx, float gy, float gh,
                             PPDubinsCenters& c)
{
    float ss = M::sin(sh), cs = M::cos(sh);
    float sg = M::sin(gh), cg = M::cos(gh);
    c.srx = sx + r * ss; c.sry = sy - r * cs;
    c.slx = sx - r * ss; c.sly = sy + r * cs;
    c.grx = gx + r * sg; c.gry = gy - r * cg;
    c.glx = gx - r * sg; c.gly = gy + r * cg;
}

// centre pair used by candidate `type` (Dubins.cpp:37, :42, :51, :60)
PP_HD void pp_dubins_pick(const PPDubinsCenters& c, int type, float& csx, float& csy, float& cgx, float& cgy)
{
    bool s_right = (type == PP_RSR) || (type == PP_RSL);
    bool g_right = (type == PP_RSR) || (type == PP_LSR);
    csx = s_right ? c.srx : c.slx; csy = s_right ? c.sry : c.sly;
    cgx = g_right ? c.grx : c.glx; cgy = g_right ? c.gry : c.gly;
}

// Pure-float tail of get_params_{rsr,rsl,lsr,lsl} (Dubins.cpp:180-323) once the transcendentals are known:
//   theta = atan2f(dc) for every type; for RSL / LSR additionally ac = acosf(2r/dist) and the sin / cos of
//   theta_t1 (= +-ac + theta) and of p[2] (= theta_t1 -+ pi).  Splitting it this way lets a warp evaluate the
//   transcendentals of all (successor, type) pairs in three parallel stages instead of ten in sequence.
// p[4] = {start angle, delta on the start circle, start angle on the goal circle, delta on the goal circle}.
PP_HD float pp_dubins_theta_t1(int type, float ac, float theta) { return (type == PP_RSL) ? ac + theta : -ac + theta; }
PP_HD float pp_dubins_p2(int type, float theta_t1)
{
    return (type == PP_RSL) ? (float)((double)theta_t1 - PP_PI) : (float)((double)theta_t1 + PP_PI);
}

// One body for the four types (the kernel is bound by instruction fetch: four specialised copies were 5.8 KB of SASS).  With
// sr / gr = "the start / goal circle is the right-hand one", Dubins.cpp:180-323 reads, for every type:
//   p0 = +-pi/2 + sh (+ iff sr), theta_g = +-pi/2 + gh (+ iff gr)                      [double sum, stored as float]
//   theta_t1 = +-pi/2 + theta (RSR +, LSL -) or +-ac + theta (RSL +, LSR -); p2 = theta_t1 (RSR, LSL) or theta_t1 -+ pi
//   p1 = theta_t1 - p0, brought to <= 0 (sr) or >= 0 (!sr) by ONE -+2pi step; p3 = theta_g - p2 likewise with gr
//   straight segment between the centres (RSR, LSL) or between the tangent points (RSL, LSR)
//   length = dist + r * (-+p1 -+ p3) (- iff right-hand circle); -(a + b) == (-a) + (-b) in IEEE arithmetic
PP_HD float pp_dubins_finish(int type, float r, float sh, float gh, float csx, float csy, float cgx, float cgy,
                             float theta, float ac, float cos_t1, float sin_t1, float cos_p2, float sin_p2, float p[4])
{
    const bool sr = (type == PP_RSR) || (type == PP_RSL), gr = (type == PP_RSR) || (type == PP_LSR);
    const bool same = (sr == gr);                    // RSR, LSL
    const double hs = sr ? PP_PI_2 : -PP_PI_2, hg = gr ? PP_PI_2 : -PP_PI_2;
    const float p0 = (float)(hs + (double)sh);
    const float theta_g = (float)(hg + (double)gh);
    float theta_t1, p2;
    if (same) { theta_t1 = (float)(hs + (double)theta); p2 = theta_t1; }
    else { theta_t1 = pp_dubins_theta_t1(type, ac, theta); p2 = pp_dubins_p2(type, theta_t1); }
    float p1 = theta_t1 - p0;
    if (sr ? (p1 > 0) : (p1 < 0)) p1 = (float)((double)p1 + (sr ? -2 * PP_PI : 2 * PP_PI));
    float p3 = theta_g - p2;
    if (gr ? (p3 > 0) : (p3 < 0)) p3 = (float)((double)p3 + (gr ? -2 * PP_PI : 2 * PP_PI));
    p[0] = p0; p[1] = p1; p[2] = p2; p[3] = p3;
    float dx = cgx - csx, dy = cgy - csy;
    if (!same)
    {
        float ssx = csx + r * cos_t1;
        float ssy = csy + r * sin_t1;
        float esx = cgx + r * cos_p2;
        float esy = cgy + r * sin_p2;
        dx = esx - ssx; dy = esy - ssy;
    }
    const float dist_st = sqrtf(dx * dx + dy * dy);
    return dist_st + r * ((sr ? -p1 : p1) + (gr ? -p3 : p3));
}

// acosf(2 r / dist) of the RSL / LSR candidates (NaN when the circle centres are closer than 2r)
PP_HD float pp_dubins_acos_arg(float r, float csx, float csy, float cgx, float cgy)
{
    float dcx = cgx - csx, dcy = cgy - csy;
    float dist = sqrtf(dcx * dcx + dcy * dcy);
    return 2 * r / dist;
}

// One candidate, all of it on the calling thread (stateless kernels, the Dubins shot).  Returns the path length
// (NaN for RSL/LSR when the centres are closer than 2r).
template <class M = PPMathExact>
PP_HD_NOINLINE_FN float pp_dubins_candidate(int type, float r, float sh, float gh,
                                         float csx, float csy, float cgx, float cgy, float p[4])
{
    float theta = M::atan2(cgy - csy, cgx - csx);
    float ac = 0.0f, c1 = 0.0f, s1 = 0.0f, c2 = 0.0f, s2 = 0.0f;
    if (type == PP_RSL || type == PP_LSR)
    {
        ac = M::acos(pp_dubins_acos_arg(r, csx, csy, cgx, cgy));
        float t1 = pp_dubins_theta_t1(type, ac, theta);
        float p2 = pp_dubins_p2(type, t1);
        c1 = M::cos(t1); s1 = M::sin(t1); c2 = M::cos(p2); s2 = M::sin(p2);
    }
    return pp_dubins_finish(type, r, sh, gh, csx, csy, cgx, cgy, theta, ac, c1, s1, c2, s2, p);
}

// Sequential fold of the four candidates, Dubins.cpp:36-68.
template <class M = PPMathExact>
PP_HD_NOINLINE_FN float pp_dubins_shortest(float r, float sx, float sy, float sh, float gx, float gy, float gh,
                               int& best_type, float best_p[4], PPDubinsCenters& c)
{
    pp_dubins_centers<M>(r, sx, sy, sh, gx, gy, gh, c);
    float best = 0.0f;
    best_type = PP_RSR;
    for (int type = 0; type < 4; type++)
    {
        float csx, csy, cgx, cgy, p[4];
        pp_dubins_pick(c, type, csx, csy, cgx, cgy);
        float len = pp_dubins_candidate<M>(type, r, sh, gh, csx, csy, cgx, cgy, p);
        if (type == 0 || len < best)
        {
            best = len; best_type = type;
            best_p[0] = p[0]; best_p[1] = p[1]; best_p[2] = p[2]; best_p[3] = p[3];
        }
    }
    return best;
}

// Geometry of the sampled path, shared by all samples (sample_path_*, Dubins.cpp:326-563).
struct PPDubinsPlan
{
    int   type;
    float p[4];
    float csx, csy, cgx, cgy;   // centres of the chosen start / goal circles
    float ssx, ssy;             // start of the straight segment
    float st_theta, st_cos, st_sin;
    int   size_1, size_2, size_3;   // cumulative sample counts; total samples = size_3 + 1
    float s1, s2;               // +1 for a left arc, -1 for a right arc
    float curvature;            // 1 / r_min
};

template <class M = PPMathExact>
PP_HD_NOINLINE_FN void pp_dubins_plan(float r, float step, float ang_step, int type, const float p[4],
                          const PPDubinsCenters& c, PPDubinsPlan& pl)
{
    pl.type = type;
    pl.p[0] = p[0]; pl.p[1] = p[1]; pl.p[2] = p[2]; pl.p[3] = p[3];
    pp_dubins_pick(c, type, pl.csx, pl.csy, pl.cgx, pl.cgy);
    pl.s1 = ((type == PP_RSR) || (type == PP_RSL)) ? -1.0f : 1.0f;
    pl.s2 = ((type == PP_RSR) || (type == PP_LSR)) ? -1.0f : 1.0f;
    pl.ssx = pl.csx + r * M::cos(p[0] + p[1]);
    pl.ssy = pl.csy + r * M::sin(p[0] + p[1]);
    float esx = pl.cgx + r * M::cos(p[2]);
    float esy = pl.cgy + r * M::sin(p[2]);
    float dx = esx - pl.ssx, dy = esy - pl.ssy;
    float length_st = sqrtf(dx * dx + dy * dy);
    // floor(-p1/as) for a right first arc, floor(p1/as) for a left one (Dubins.cpp:342, :407, :467, :527)
    float a1 = (pl.s1 < 0) ? -p[1] : p[1];
    float a3 = (pl.s2 < 0) ? -p[3] : p[3];
    pl.size_1 = (int)floorf(a1 / ang_step);
    pl.size_2 = pl.size_1 + (int)floorf(length_st / step);
    pl.size_3 = pl.size_2 + (int)floorf(a3 / ang_step);
    pl.st_theta = M::atan2(dy, dx);
    pl.st_cos = M::cos(pl.st_theta);
    pl.st_sin = M::sin(pl.st_theta);
    pl.curvature = 1 / r;
}

// Sample k of the plan.  `acc` is the float accumulator of the segment the sample lies on:
// theta (arcs; p0 -/+ k*ang_step accumulated one step at a time) or dist (straight; k*step likewise).
template <class M = PPMathExact>
PP_HD_NOINLINE_FN void pp_dubins_sample(const PPDubinsPlan& pl, float r, int k, float acc,
                            float& x, float& y, float& heading, float& curvature)
{
    if (k < pl.size_1)
    {
        x = pl.csx + r * M::cos(acc);
        y = pl.csy + r * M::sin(acc);
        heading = (float)pp_wrap_pi_d((double)acc + (double)pl.s1 * PP_PI_2);
        curvature = pl.curvature;
    }
    else if (k < pl.size_2)
    {
        x = pl.ssx + acc * pl.st_cos;
        y = pl.ssy + acc * pl.st_sin;
        heading = pl.st_theta;
        curvature = 0.0f;
    }
    else if (k < pl.size_3)
    {
        x = pl.cgx + r * M::cos(acc);
        y = pl.cgy + r * M::sin(acc);
        heading = (float)pp_wrap_pi_d((double)acc + (double)pl.s2 * PP_PI_2);
        curvature = pl.curvature;
    }
    else
    {
        float a = pl.p[2] + pl.p[3];
        x = pl.cgx + r * M::cos(a);
        y = pl.cgy + r * M::sin(a);
        heading = (float)pp_wrap_pi_d((double)a + (double)pl.s2 * PP_PI_2);
        curvature = 0.0f;
    }
}

#endif
