// Velocity profile and trajectory assembly: the step the caller runs right after find_path.
//
//   pp_velocity_profile    VelocityGenerator<float>::generate_velocity_profile (lib/VelocityGenerator.cpp:19-85): three
//                          sequential passes over v^2 along a path stored goal -> start -- curvature / speed cap with the
//                          braking budget, forward pass with the acceleration budget, backward pass with the braking budget.
//   pp_trajectory_assemble HybridAStar::reconstruct_path (lib/HybridAStar.cpp:208-262: reversed Dubins samples, parent chain,
//                          grid -> world, curvature shifted by one point) + the profile + the message layout of
//                          LocalPlanner::publish_trajectory (src/local_planner.cpp:346-372: x.., y.., heading.. in start -> goal
//                          order, then the velocities), straight from the search's device-resident path records.
//
// Arithmetic follows the reference operation by operation (float, with its double promotions: `sqrt(1.0 - ...)` and the
// product with the acceleration limit are double, rounded to float on assignment; std::hypot(float, float) == pp_hypotf;
// std::min / std::max as the exact ternaries of libstdc++, which decides what a NaN does).  Compile without FMA contraction.
// NaN results (v^2 * curvature above the lateral limit makes the reference take the root of a negative number and publish
// NaN velocities) are NaN here too; the NaN's sign / payload bits are the platform's.
#ifndef PP_VELOCITY_H
#define PP_VELOCITY_H

#include "pp_defs.h"
#include "pp_math.h"

struct PPVelLimits      // the five constructor arguments of VelocityGenerator (VelocityGenerator.cpp:6-14)
{
    float max_velocity, coast_velocity, max_lat_acc, max_long_acc, max_long_dec;
};

PP_HD float pp_std_min(float a, float b) { return (b < a) ? b : a; }       // std::min: bits/stl_algobase.h
PP_HD float pp_std_max(float a, float b) { return (a < b) ? b : a; }       // std::max

// `acc * std::sqrt(1.0 - (lat * lat) / lat_sqr)` assigned to a float (VelocityGenerator.cpp:42, :59, :71)
PP_HD float pp_long_acc_rem(float acc, float lat, float lat_sqr)
{
    return (float)((double)acc * sqrt(1.0 - (double)((lat * lat) / lat_sqr)));
}

// Path point p (0 = goal end, n - 1 = start end) is (px[p * pstride], py[p * pstride]) with curvature curv[p * cstride].
// v2 = scratch of n floats, vel = n floats out (index 0 = the start of the path).  Returns the reference's feasibility flag.
// n >= 1 (the reference indexes path[n - 1] unconditionally).
PP_HD bool pp_velocity_profile(const PPVelLimits& L, float vel_init, float max_velocity_curr, const float* px, const float* py,
                               int pstride, const float* curv, int cstride, int n, float* v2, float* vel, bool coast, bool stop)
{
    const float lat_sqr = L.max_lat_acc * L.max_lat_acc;
    float vmax = coast ? L.coast_velocity : L.max_velocity;
    vmax = pp_std_min(vmax, max_velocity_curr);
    const float vmax_sqr = vmax * vmax;
    v2[0] = vel_init * vel_init;
    float cap = v2[0];
    for (int i = 0; i < n - 1; i++)                                        // VelocityGenerator.cpp:35-49
    {
        const int p = n - i - 1;
        float step = pp_hypotf(px[(p - 1) * pstride] - px[p * pstride], py[(p - 1) * pstride] - py[p * pstride]);
        float lat = v2[i] * curv[p * cstride];
        float rem = pp_long_acc_rem(L.max_long_dec, lat, lat_sqr);
        cap = pp_std_max(cap - 2 * rem * step, vmax_sqr);
        float k1 = curv[(p - 1) * cstride];
        v2[i + 1] = (k1 != 0) ? pp_std_min(L.max_lat_acc / k1, cap) : cap;
    }
    if (stop) v2[n - 1] = 0.0f;                                            // :52
    for (int i = 0; i < n - 1; i++)                                        // forward pass, :55-63
    {
        const int p = n - i - 1;
        float step = pp_hypotf(px[(p - 1) * pstride] - px[p * pstride], py[(p - 1) * pstride] - py[p * pstride]);
        float lat = v2[i] * curv[p * cstride];
        float rem = pp_long_acc_rem(L.max_long_acc, lat, lat_sqr);
        v2[i + 1] = pp_std_min(v2[i] + 2 * rem * step, v2[i + 1]);
    }
    for (int i = n - 1; i > 0; i--)                                        // backward pass, :66-76
    {
        const int p = n - i - 1;
        float step = pp_hypotf(px[(p + 1) * pstride] - px[p * pstride], py[(p + 1) * pstride] - py[p * pstride]);
        float lat = v2[i] * curv[p * cstride];
        float rem = pp_long_acc_rem(L.max_long_dec, lat, lat_sqr);
        v2[i - 1] = pp_std_min(v2[i] + 2 * rem * step, v2[i - 1]);
        vel[i - 1] = sqrtf(v2[i - 1]);
    }
    vel[n - 1] = sqrtf(v2[n - 1]);                                         // :79
    return vel_init < (vel[0] + 0.25f);                                    // :80-81
}

// Grid-frame record of the search -> world pose, HybridAStar.cpp:224-233 / :243-251: (p - goal_grid) rotated by -grid_heading
// (cos / sin of that angle come from the host, evaluated once per frame with the platform libm like the reference does per
// point), heading wrapped, goal offset added.
struct PPWorldFrame
{
    float goal_gx, goal_gy;      // goal in the grid frame
    float goal_wx, goal_wy;      // goal in the world frame
    float angle;                 // -grid_heading
    float c, s;                  // cosf(angle), sinf(angle)
};

PP_HD void pp_to_world(const PPWorldFrame& F, float x, float y, float h, float& wx, float& wy, float& wh)
{
    float rx = x - F.goal_gx, ry = y - F.goal_gy;
    wx = rx * F.c + ry * F.s;
    wy = -rx * F.s + ry * F.c;
    wh = pp_wrap_pi(h - F.angle);
    wx += F.goal_wx;
    wy += F.goal_wy;
}

struct PPTrajPt { float x, y, heading, curvature; };      // = PPPathPt of pp_search.h (kept separate: no search include here)

// One query's trajectory, cooperatively by the lanes of `w` (all lanes call with identical arguments).
//   src            the search's records: [0, n_dubins) Dubins samples in forward order, then the parent chain terminal -> start
//   traj           4 * n floats out: x[0..n), y[0..n), heading[0..n) in start -> goal order, then velocity[0..n)
//   curv_tmp, v2_tmp  scratch of n floats each
// Returns n (0 when the query failed) and the feasibility flag through *feasible.
template <class W>
PP_HD int pp_trajectory_assemble(const W& w, const PPWorldFrame& F, const PPVelLimits& L, const PPTrajPt* src, int n_dubins,
                                 int n_chain, int cap, float vel_init, float max_velocity_curr, bool stop, float* traj,
                                 float* curv_tmp, float* v2_tmp, int* feasible)
{
    const int nd = n_dubins < cap ? n_dubins : cap;
    int nc = cap - nd; if (nc < 0) nc = 0; if (n_chain < nc) nc = n_chain;
    const int n = nd + nc;
    if (n <= 0) { if (w.lane() == 0) *feasible = 0; return 0; }
    // path order t = 0 (goal end) .. n - 1 (start end): reversed Dubins samples, then the chain (HybridAStar.cpp:215-257)
    for (int t = w.lane(); t < n; t += W::LANES)
    {
        const PPTrajPt& p = (t < nd) ? src[nd - 1 - t] : src[t];
        float wx, wy, wh;
        pp_to_world(F, p.x, p.y, p.heading, wx, wy, wh);
        const int i = n - 1 - t;                                           // publish_trajectory walks the path in reverse
        traj[i] = wx; traj[n + i] = wy; traj[2 * n + i] = wh;
        // curvature[t] = curvature of the record before it; the first entry is 0 (HybridAStar.cpp:212, :236, :256, :261)
        float k = 0.0f;
        if (t > 0) { const PPTrajPt& q = (t - 1 < nd) ? src[nd - 1 - (t - 1)] : src[t - 1]; k = q.curvature; }
        curv_tmp[t] = k;
    }
    w.sync();
    if (w.lane() == 0)
    {
        // path point p is trajectory sample n - 1 - p: negative stride over the x and y blocks
        bool ok = pp_velocity_profile(L, vel_init, max_velocity_curr, traj + (n - 1), traj + n + (n - 1), -1, curv_tmp, 1, n, v2_tmp,
                                      traj + 3 * n, false, stop);
        *feasible = ok ? 1 : 0;
    }
    w.sync();
    return n;
}

#endif
