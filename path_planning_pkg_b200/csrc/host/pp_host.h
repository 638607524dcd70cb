// Host-side (CPU) part of the product: everything the reference computes once per constructor /
// goal change / obstacle message in a handful of scalar operations -- derived constants, the
// VehicleModel displacement table, the world->grid frame, start-node construction, the APF obstacle
// list, per-box rasterisation descriptors, and the final grid->world path transform.  These use the
// platform libm exactly like the reference (so they agree bit-for-bit with a reference built on the
// same host) and their outputs are uploaded to the device; all per-cell / per-state / per-query work
// is in the CUDA kernels.  Compile with -ffp-contract=off (the reference is baseline x86-64: no FMA).
#ifndef PP_HOST_H
#define PP_HOST_H

#include <cmath>
#include <vector>
#include <string>
#include <limits>
#include <algorithm>
#include <utility>

#include "../core/pp_defs.h"
#include "../core/pp_map.h"
#include "../../../include/pp_b200.h"

struct PPHostModel
{
    PPConsts C;
    std::vector<float> off_xy;   // [S][bins+1][2]; column `bins` duplicates column 0
};

struct PPHostFrame
{
    PPFrame F;
    float grid_heading;
    float goal_world[3];
};

// wrap_pi<float>, common.h:15-29 (host copy; the device copy is pp_math.h)
inline float pp_host_wrap_pi(float angle)
{
    float w = std::fmod(angle, 2 * M_PI);
    if (w > M_PI) return w - 2 * M_PI;
    if (w < -M_PI) return w + 2 * M_PI;
    return w;
}

// get_heading_index<float>, common.h:9-12, :32-36
inline int pp_host_heading_index(float heading, float precision)
{
    float rounded = std::round(heading / precision) * precision;
    return static_cast<int>((rounded + M_PI) / precision);
}

// Vector2D::get_rotated_vector (common.h:55-61): rotation by -angle
inline void pp_host_rotate(float x, float y, float angle, float& rx, float& ry)
{
    float c = std::cos(angle), s = std::sin(angle);
    rx = x * c + y * s;
    ry = -x * s + y * c;
}

// Constructors of Grid2D (Grid2D.cpp:7-62), VehicleModel (VehicleModel.cpp:7-47, :147-164),
// Dubins (Dubins.cpp:7-16) and HybridAStar (HybridAStar.cpp:7-24).
inline bool pp_host_build_model(const pp_params& p, PPHostModel& m, std::string& err)
{
    if (p.grid_size < 2 || p.grid_size > 8192) { err = "grid_size out of range"; return false; }
    if (p.num_steering < 1 || p.num_steering > PP_MAX_STEER) { err = "num_steering out of range"; return false; }
    if (p.num_angle_bins < 1 || p.num_angle_bins + 1 > PP_MAX_BINS) { err = "num_angle_bins out of range"; return false; }
    if (p.num_actions < 0 || 2 * p.num_actions + 1 > 16) { err = "num_actions out of range"; return false; }
    PPConsts& C = m.C;
    C.N = p.grid_size;
    C.n2 = static_cast<int>(std::round(p.grid_size * 0.5));
    C.n45 = static_cast<int>(std::round(p.grid_size * 0.8));
    C.res = p.resolution;
    C.log_thr = std::log(p.obstacle_threshold / (1.0 - p.obstacle_threshold));
    C.log_min = std::log(p.prob_min / (1.0 - p.prob_min));
    C.log_max = std::log(p.prob_max / (1.0 - p.prob_max));
    C.log_free = std::log(p.prob_free / (1.0 - p.prob_free));
    static const int d8[8][2] = {{0, -1}, {1, -1}, {1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}};
    static const int d4[4][2] = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}};
    C.n_act2d = p.allow_diag ? 8 : 4;
    for (int k = 0; k < 8; k++) { C.act_di[k] = 0; C.act_dj[k] = 0; C.act_cost[k] = 0.0f; }
    for (int k = 0; k < C.n_act2d; k++)
    {
        C.act_di[k] = p.allow_diag ? d8[k][0] : d4[k][0];
        C.act_dj[k] = p.allow_diag ? d8[k][1] : d4[k][1];
        C.act_cost[k] = C.res * std::sqrt(static_cast<float>(C.act_di[k] * C.act_di[k] + C.act_dj[k] * C.act_dj[k]));
    }

    C.S = p.num_steering;
    C.A = p.num_actions;
    C.bins = p.num_angle_bins;
    C.ts = p.step_size;
    C.max_lat_acc = p.max_lat_acc;
    C.max_lat_acc_sqr = p.max_lat_acc * p.max_lat_acc;
    C.precision = 2 * M_PI / p.num_angle_bins;
    std::vector<float> beta(C.S), curv(C.S);
    for (int i = 0; i < C.S; i++)
    {
        beta[i] = std::atan2(p.rear_to_cg * std::tan(p.steering[i]), p.wheelbase);
        curv[i] = std::cos(beta[i]) * std::tan(p.steering[i]) / p.wheelbase;
    }
    for (int i = 0; i < PP_MAX_STEER; i++) { C.abs_curv[i] = 0; C.act_cost3d[i] = 0; C.off_heading[i] = 0; }
    m.off_xy.assign((size_t)C.S * (C.bins + 1) * 2, 0.0f);
    for (int i = 0; i < C.S; i++)
    {
        C.off_heading[i] = C.ts * curv[i];
        C.act_cost3d[i] = C.ts + p.curvature_weights[i] * std::abs(curv[i]);
        for (int j = 0; j < C.bins; j++)
        {
            // VehicleModel::calculate_offset(beta, curvature, heading), VehicleModel.cpp:147-164
            const float heading = -M_PI + j * C.precision;
            const float dt = static_cast<float>(0.001);
            float ox = 0.0f, oy = 0.0f, cur = heading;
            int num_updates = static_cast<int>(C.ts / dt);
            for (int u = 0; u < num_updates; u++)
            {
                ox += dt * std::cos(beta[i] + cur);
                oy += dt * std::sin(beta[i] + cur);
                cur += dt * curv[i];
            }
            m.off_xy[((size_t)i * (C.bins + 1) + j) * 2] = ox;
            m.off_xy[((size_t)i * (C.bins + 1) + j) * 2 + 1] = oy;
        }
        // padding column: heading bin == bins is +pi == -pi, i.e. column 0 (the reference reads out of bounds there)
        m.off_xy[((size_t)i * (C.bins + 1) + C.bins) * 2] = m.off_xy[((size_t)i * (C.bins + 1)) * 2];
        m.off_xy[((size_t)i * (C.bins + 1) + C.bins) * 2 + 1] = m.off_xy[((size_t)i * (C.bins + 1)) * 2 + 1];
    }
    for (int i = 0; i < C.S; i++) C.abs_curv[i] = std::abs(curv[i]);

    C.apf_k = p.apf_rep_constant;
    C.apf_alpha = p.apf_active_angle;

    // tan_max(steering), HybridAStar.h:21-25; r_min, HybridAStar.cpp:23-24
    float smax = *std::max_element(p.steering, p.steering + p.num_steering);
    float tmax = std::tan(smax);
    C.r_min = p.wheelbase / (std::cos(std::atan2(p.rear_to_cg * tmax, p.wheelbase)) * tmax);
    C.step = p.step_size;
    C.ang_step = C.step / C.r_min;
    C.shot_interval = p.shot_interval;
    C.shot_decay = p.shot_decay;
    return true;
}

// Grid2D::update_goal_heading (Grid2D.cpp:260-266) + goal node of Grid3D::update_goal_heading (Grid3D.cpp:116-123)
inline void pp_host_update_goal(const PPConsts& C, const float* goal3, const float* start3, PPHostFrame& fr)
{
    fr.goal_world[0] = goal3[0]; fr.goal_world[1] = goal3[1]; fr.goal_world[2] = goal3[2];
    fr.grid_heading = std::atan2(goal3[1] - start3[1], goal3[0] - start3[0]);
    fr.F.goal_x = C.n45 * C.res;
    fr.F.goal_y = C.n2 * C.res;
    fr.F.goal_h = pp_host_wrap_pi(goal3[2] - fr.grid_heading);
    fr.F.goal_bin = pp_host_heading_index(fr.F.goal_h, C.precision);
    fr.F.goal_ci = C.n45;
    fr.F.goal_cj = C.n2;
}

// Grid3D::set_start_node (Grid3D.cpp:127-160) + HybridAStar::find_path (HybridAStar.cpp:72-74)
inline PPState pp_host_set_start(const PPConsts& C, const PPHostFrame& fr, float sx, float sy, float sh, float vel)
{
    PPState s;
    float rx, ry;
    pp_host_rotate(sx - fr.goal_world[0], sy - fr.goal_world[1], fr.grid_heading, rx, ry);
    float rh = pp_host_wrap_pi(sh - fr.grid_heading);
    float px = rx + C.n45 * C.res;
    float py = ry + C.n2 * C.res;
    int i = static_cast<int>(px / C.res);
    int j = static_cast<int>(py / C.res);
    if ((i > -1) && (i < C.N) && (j > -1) && (j < C.N))
    {
        s.x = px; s.y = py; s.heading = rh; s.ci = i; s.cj = j;
    }
    else
    {
        s.x = 0.0f; s.y = 0.0f; s.heading = 0.0f; s.ci = 0; s.cj = 0;   // "Default to (0, 0, 0) node"
    }
    s.bin = pp_host_heading_index(s.heading, C.precision);
    s.curv = C.S / 2;                     // get_default_action_index, VehicleModel.cpp:56-60
    s.g = 0.0f;
    s.vmin_sqr = vel * vel;
    s.f = std::numeric_limits<float>::max();
    return s;
}

// APF obstacle list, Grid3D::update_obstacles (Grid3D.cpp:26-40): out = n x (x, y, radius), grid-frame metres
inline void pp_host_apf_list(const PPConsts& C, const PPHostFrame& fr, const float* boxes, int n, float added_radius,
                             std::vector<float>& out)
{
    out.resize((size_t)n * 3);
    for (int k = 0; k < n; k++)
    {
        float rx, ry;
        pp_host_rotate(boxes[4 * k] - fr.goal_world[0], boxes[4 * k + 1] - fr.goal_world[1], fr.grid_heading, rx, ry);
        rx += C.n45 * C.res;
        ry += C.n2 * C.res;
        out[3 * k] = rx; out[3 * k + 1] = ry;
        out[3 * k + 2] = std::max(boxes[4 * k + 2], boxes[4 * k + 3]) / 2 + added_radius;
    }
}

// Conservative spatial index of the APF list (not in the reference, which scans every obstacle for every successor,
// Grid3D.cpp:203-226): the grid is cut into square bins of (1 << shift) cells; a bin lists, in ascending obstacle
// order, every obstacle whose influence disc (radius + one cell of slack for the float cell rounding) touches the bin.
// An obstacle outside a pose's bin list is at least its radius away, i.e. its term is exactly zero, so summing the
// listed obstacles in order reproduces the full std::accumulate.
inline void pp_host_apf_bins(const PPConsts& C, const std::vector<float>& apf, int n, int shift, int& bin_n,
                             std::vector<int>& off, std::vector<int>& idx)
{
    const int B = 1 << shift;
    bin_n = (C.N + B - 1) >> shift;
    const double w = (double)B * (double)C.res;
    off.assign((size_t)bin_n * bin_n + 1, 0);
    idx.clear();
    for (int pass = 0; pass < 2; pass++)
    {
        std::vector<int> fill;
        if (pass == 1)
        {
            long long acc = 0;
            for (size_t b = 0; b + 1 < off.size(); b++) { long long c = off[b]; off[b] = (int)acc; acc += c; }
            off[off.size() - 1] = (int)acc;
            idx.assign((size_t)acc, 0);
            fill.assign(off.begin(), off.end() - 1);
        }
        for (int k = 0; k < n; k++)
        {
            const double ox = apf[3 * k], oy = apf[3 * k + 1];
            const double reach = (double)apf[3 * k + 2] * 1.001 + 1.5 * (double)C.res + 1e-3;
            int i0 = (int)std::floor((ox - reach) / w), i1 = (int)std::floor((ox + reach) / w);
            int j0 = (int)std::floor((oy - reach) / w), j1 = (int)std::floor((oy + reach) / w);
            i0 = std::max(i0, 0); j0 = std::max(j0, 0); i1 = std::min(i1, bin_n - 1); j1 = std::min(j1, bin_n - 1);
            for (int bi = i0; bi <= i1; bi++)
                for (int bj = j0; bj <= j1; bj++)
                {
                    // distance from the obstacle centre to the bin rectangle
                    double dx = std::max(std::max(bi * w - ox, ox - (bi + 1) * w), 0.0);
                    double dy = std::max(std::max(bj * w - oy, oy - (bj + 1) * w), 0.0);
                    if (dx * dx + dy * dy > reach * reach) continue;
                    size_t b = (size_t)bi * bin_n + bj;
                    if (pass == 0) off[b]++; else idx[fill[b]++] = k;
                }
        }
    }
}

// Per-box rasterisation descriptor, the scalar prologue of Grid2D::update_obstacles (Grid2D.cpp:102-122).
struct PPBoxDesc
{
    int   start_i, start_j;   // bottom-left corner cell
    int   ni, nj;             // 2*ceil(dx/res), 2*ceil(dy/res) half-cell samples per axis
    float delta;              // log(c/(1-c)) - log_free
    int   lo_i, hi_i, lo_j, hi_j;   // inclusive cell bounding box of all samples (clipped to the grid)
};

inline void pp_host_box_descs(const PPConsts& C, const PPHostFrame& fr, const float* boxes, const float* conf, int n,
                              float& cos_h, float& sin_h, std::vector<PPBoxDesc>& out)
{
    cos_h = std::cos(fr.grid_heading);
    sin_h = std::sin(fr.grid_heading);
    out.resize(n);
    for (int k = 0; k < n; k++)
    {
        float px = boxes[4 * k] - boxes[4 * k + 2] / 2, py = boxes[4 * k + 1] - boxes[4 * k + 3] / 2;
        float x = px - fr.goal_world[0], y = py - fr.goal_world[1];
        float xo = x;
        x = xo * cos_h + y * sin_h;
        y = -xo * sin_h + y * cos_h;
        PPBoxDesc d;
        d.start_i = static_cast<int>(std::round(x / C.res) + C.n45);
        d.start_j = static_cast<int>(std::round(y / C.res) + C.n2);
        d.ni = 2 * static_cast<int>(std::ceil(boxes[4 * k + 2] / C.res));
        d.nj = 2 * static_cast<int>(std::ceil(boxes[4 * k + 3] / C.res));
        float log_conf = std::log(conf[k] / (1.0 - conf[k]));
        d.delta = log_conf - C.log_free;
        // bounding box of round(rotate(0.5 i, 0.5 j)) over the sample lattice: extremes are at the corners
        float ex = 0.5f * (d.ni > 0 ? d.ni - 1 : 0), ey = 0.5f * (d.nj > 0 ? d.nj - 1 : 0);
        float cx[4] = {0.0f, ex, 0.0f, ex}, cy[4] = {0.0f, 0.0f, ey, ey};
        float lo_u = 0, hi_u = 0, lo_v = 0, hi_v = 0;
        for (int q = 0; q < 4; q++)
        {
            float u = cx[q] * cos_h + cy[q] * sin_h, v = -cx[q] * sin_h + cy[q] * cos_h;
            lo_u = std::min(lo_u, u); hi_u = std::max(hi_u, u);
            lo_v = std::min(lo_v, v); hi_v = std::max(hi_v, v);
        }
        d.lo_i = std::max(0, d.start_i + static_cast<int>(std::floor(lo_u)) - 1);
        d.hi_i = std::min(C.N - 1, d.start_i + static_cast<int>(std::ceil(hi_u)) + 1);
        d.lo_j = std::max(0, d.start_j + static_cast<int>(std::floor(lo_v)) - 1);
        d.hi_j = std::min(C.N - 1, d.start_j + static_cast<int>(std::ceil(hi_v)) + 1);
        out[k] = d;
    }
}

// Per-line prologue of Grid2D::update_obstacles(lines, conf, width), Grid2D.cpp:147-156
inline void pp_host_line_descs(const PPConsts& C, const PPHostFrame& fr, const float* lines, const float* conf, int n,
                               std::vector<PPLineDesc>& out)
{
    out.resize(n);
    for (int k = 0; k < n; k++)
    {
        float sx, sy, ex, ey;
        pp_host_rotate(lines[4 * k] - fr.goal_world[0], lines[4 * k + 1] - fr.goal_world[1], fr.grid_heading, sx, sy);
        pp_host_rotate(lines[4 * k + 2] - fr.goal_world[0], lines[4 * k + 3] - fr.goal_world[1], fr.grid_heading, ex, ey);
        float dx = ex - sx, dy = ey - sy;
        float len = std::hypot(dx, dy);
        PPLineDesc d;
        d.sx = sx; d.sy = sy;
        d.nx = -dy / len; d.ny = dx / len;
        d.ux = dx / len; d.uy = dy / len;
        d.length = len;
        float log_conf = std::log(conf[k] / (1.0 - conf[k]));
        d.delta = log_conf - C.log_free;
        out[k] = d;
    }
}

// Prologue of Grid3D::relocate_obstacles (Grid3D.cpp:171-181)
inline void pp_host_reloc_desc(const PPConsts& C, float heading_new, float heading_prev, const float* goal_new,
                               const float* goal_prev_world, PPRelocDesc& d)
{
    float dh = heading_new - heading_prev;
    d.cos_d = std::cos(dh); d.sin_d = std::sin(dh);
    float gpx, gpy;       // goal_prev: (n45, n2) rotated by dh
    pp_host_rotate(static_cast<float>(C.n45), static_cast<float>(C.n2), dh, gpx, gpy);
    float nox, noy;       // goal_new_to_old rotated by the new heading
    pp_host_rotate(goal_prev_world[0] - goal_new[0], goal_prev_world[1] - goal_new[1], heading_new, nox, noy);
    d.ox = static_cast<float>(C.n45) + (nox / C.res) - gpx;
    d.oy = static_cast<float>(C.n2) + (noy / C.res) - gpy;
}

// Grid->world transform of HybridAStar::reconstruct_path (HybridAStar.cpp:224-233, :243-251):
// (p - goal_grid).get_rotated_vector(-grid_heading) + goal
inline void pp_host_to_world(const PPHostFrame& fr, float x, float y, float h, float& wx, float& wy, float& wh)
{
    float rx = x - fr.F.goal_x, ry = y - fr.F.goal_y;
    float a = -fr.grid_heading;
    float c = std::cos(a), s = std::sin(a);
    wx = rx * c + ry * s;
    wy = -rx * s + ry * c;
    wh = pp_host_wrap_pi(h - a);
    wx += fr.goal_world[0];
    wy += fr.goal_world[1];
}

#endif
