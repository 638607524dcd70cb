// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// CPU restatement ("port") of the reference's local-planner search hot path, in plain C++ with flat
// records and the C++ standard containers, each function citing the reference file:line it follows
// (paths relative to the reference repository).  It exists so the parity tests have an oracle even
// where the compiled reference (oracle/_ref) is absent, and it is itself pinned against oracle/_ref
// and the reference's golden vectors by tests/test_cpu_oracle.py.  Arithmetic follows SURVEY.md
// Appendix A: float state, double where the reference promotes, platform libm, no FMA contraction
// (compiled with -ffp-contract=off).
//
// Deliberately NOT shared with the product: the product emulates libstdc++'s red-black tree by hand on
// the device (csrc/core/pp_rbtree.h); this restatement simply uses std::set / std::unordered_map with the
// reference's comparators, so it inherits libstdc++'s behaviour by construction.
#include <cmath>
#include <cstring>
#include <cstdio>
#include <vector>
#include <set>
#include <unordered_map>
#include <algorithm>
#include <limits>
#include <numeric>
#include <tuple>

#include "../oracle_api.h"

namespace
{
// ---- common.h -------------------------------------------------------------------------------------
float wrap_pi_f(float angle)                     // common.h:15-29, T = float
{
    float w = std::fmod(angle, 2 * M_PI);
    if (w > M_PI) return w - 2 * M_PI;
    if (w < -M_PI) return w + 2 * M_PI;
    return w;
}
double wrap_pi_d(double angle)                   // common.h:15-29, T = double
{
    double w = std::fmod(angle, 2 * M_PI);
    if (w > M_PI) return w - 2 * M_PI;
    if (w < -M_PI) return w + 2 * M_PI;
    return w;
}
int heading_index(float heading, float precision)   // common.h:9-12, :32-36
{
    float rounded = std::round(heading / precision) * precision;
    return static_cast<int>((rounded + M_PI) / precision);
}
void rotate(float x, float y, float angle, float& rx, float& ry)   // common.h:55-61
{
    float c = std::cos(angle), s = std::sin(angle);
    rx = x * c + y * s;
    ry = -x * s + y * c;
}

struct N2 { int i, j; float g, h, f; int prev; };            // Node2D.h:15-20 (prev = closed index)
struct N2Less                                                  // Node2D.h:37-41
{
    bool operator()(const N2& a, const N2& b) const { return (a.i != b.i || a.j != b.j) && (a.f < b.f); }
};
struct N3 { float x, y, h, g, f, v2; int curv, bin, ci, cj, prev; };   // Node3D.h:17-25
struct N3Less                                                            // Node3D.h:45-54 (uses operator!=: cell or bin differ)
{
    bool operator()(const N3& a, const N3& b) const
    { return (a.ci != b.ci || a.cj != b.cj || a.bin != b.bin) && (a.f < b.f); }
};

struct Port
{
    orc_params p;
    // Grid2D (Grid2D.cpp:7-62)
    int N, n2, n45;
    float res, log_thr, log_min, log_max, log_free, grid_heading;
    float goal_loc[3];
    std::vector<float> map;                 // [i*N + j]
    std::vector<float> nm_g, nm_f, nm_h;    // _node_map costs
    std::vector<std::pair<int, int>> actions; std::vector<float> actions_cost;
    // VehicleModel (VehicleModel.cpp:7-47)
    int S, A, bins;
    float ts, max_lat_acc, max_lat_acc_sqr, precision;
    std::vector<float> abs_curv, act_cost, off_h, off_xy;   // off_xy [S][bins][2]
    // Grid3D
    float apf_k, apf_alpha;
    std::vector<float> apf;                 // K x (x, y, r)
    // Dubins (Dubins.cpp:7-16)
    float r_min, step, ang_step;
    // HybridAStar
    N3 goal_node;
    // AStar
    std::vector<unsigned char> visited;
    long n_pops, n_oob;
};


void build(Port& P, const orc_params& p)
{
    P.p = p;
    P.N = p.grid_size; P.res = p.resolution;
    P.n2 = static_cast<int>(std::round(p.grid_size * 0.5));          // Grid2D.cpp:16
    P.n45 = static_cast<int>(std::round(p.grid_size * 0.8));         // Grid2D.cpp:17
    P.log_thr = std::log(p.obstacle_threshold / (1.0 - p.obstacle_threshold));   // Grid2D.cpp:11-14
    P.log_min = std::log(p.prob_min / (1.0 - p.prob_min));
    P.log_max = std::log(p.prob_max / (1.0 - p.prob_max));
    P.log_free = std::log(p.prob_free / (1.0 - p.prob_free));
    P.grid_heading = std::atan2(0.0f, 0.0f);
    P.goal_loc[0] = P.goal_loc[1] = P.goal_loc[2] = 0.0f;
    size_t nn = (size_t)P.N * P.N;
    P.map.assign(nn, 0.0f); P.nm_g.assign(nn, 0.0f); P.nm_f.assign(nn, 0.0f); P.nm_h.assign(nn, 0.0f);
    P.visited.assign(nn, 0);
    if (p.allow_diag) P.actions = {{0, -1}, {1, -1}, {1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}};   // Grid2D.cpp:34-41
    else P.actions = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}};
    for (auto& a : P.actions)
        P.actions_cost.push_back(P.res * std::sqrt(static_cast<float>(a.first * a.first + a.second * a.second)));   // Grid2D.cpp:55
    for (int i = 0; i < P.N; i++)                                    // compute_heuristic, Grid2D.cpp:303-316
    {
        float dx = (P.n45 - i) * P.res, dx2 = dx * dx;
        for (int j = 0; j < P.N; j++)
        {
            float dy = (P.n2 - j) * P.res, dy2 = dy * dy;
            float h = std::sqrt(dx2 + dy2);
            P.nm_h[(size_t)i * P.N + j] = h; P.nm_f[(size_t)i * P.N + j] = h;   // set_heuristic_cost: f = g + h
        }
    }
    // VehicleModel.cpp:7-47
    P.S = p.num_steering; P.A = p.num_actions; P.bins = p.num_angle_bins;
    P.ts = p.step_size; P.max_lat_acc = p.max_lat_acc; P.max_lat_acc_sqr = p.max_lat_acc * p.max_lat_acc;
    P.precision = 2 * M_PI / p.num_angle_bins;
    std::vector<float> beta(P.S), curv(P.S);
    for (int i = 0; i < P.S; i++)
    {
        beta[i] = std::atan2(p.rear_to_cg * std::tan(p.steering[i]), p.wheelbase);
        curv[i] = std::cos(beta[i]) * std::tan(p.steering[i]) / p.wheelbase;
    }
    P.abs_curv.resize(P.S); P.act_cost.resize(P.S); P.off_h.resize(P.S); P.off_xy.assign((size_t)P.S * P.bins * 2, 0.0f);
    for (int i = 0; i < P.S; i++)
    {
        P.off_h[i] = P.ts * curv[i];
        P.act_cost[i] = P.ts + p.curvature_weights[i] * std::abs(curv[i]);
        for (int j = 0; j < P.bins; j++)
        {
            const float heading = -M_PI + j * P.precision;           // VehicleModel.cpp:37
            const float dt = static_cast<float>(0.001);              // calculate_offset, VehicleModel.cpp:147-164
            float ox = 0, oy = 0, cur = heading;
            int num = static_cast<int>(P.ts / dt);
            for (int u = 0; u < num; u++)
            {
                ox += dt * std::cos(beta[i] + cur);
                oy += dt * std::sin(beta[i] + cur);
                cur += dt * curv[i];
            }
            P.off_xy[((size_t)i * P.bins + j) * 2] = ox; P.off_xy[((size_t)i * P.bins + j) * 2 + 1] = oy;
        }
    }
    for (int i = 0; i < P.S; i++) P.abs_curv[i] = std::abs(curv[i]);
    P.apf_k = p.apf_rep_constant; P.apf_alpha = p.apf_active_angle;
    float smax = *std::max_element(p.steering, p.steering + p.num_steering);   // tan_max, HybridAStar.h:21-25
    float tmax = std::tan(smax);
    P.r_min = p.wheelbase / (std::cos(std::atan2(p.rear_to_cg * tmax, p.wheelbase)) * tmax);   // HybridAStar.cpp:23-24
    P.step = p.step_size; P.ang_step = P.step / P.r_min;
    std::memset(&P.goal_node, 0, sizeof(P.goal_node));
    P.goal_node.ci = P.goal_node.cj = -1;
}

float clampv(const Port& P, float v) { return std::max(std::min(v, P.log_max), P.log_min); }

// libm flavour of the functions the B200 build executes on the device (Dubins.cpp, atan2f of Grid3D::get_field_intensity):
// 0 = platform libm like the reference; 1 = "pinned": evaluate in double, round once to float (see oracle/cr_math.c).
// With flavour 1 this restatement equals oracle/_ref/libref_oracle_crm.so.
// K-POP mode only (new semantics): flavour 2 = the FP32 functions of fmath.inc for the heuristic Dubins length and the APF term.
#include "fmath.inc"
static bool g_pinned_libm = false;
static bool g_fast_libm = false;
struct FastLibmScope { bool saved; FastLibmScope() : saved(g_fast_libm) { g_fast_libm = true; } ~FastLibmScope() { g_fast_libm = saved; } };
static inline float m_sin(float x) { return g_fast_libm ? fm::sin(x) : g_pinned_libm ? (float)std::sin((double)x) : std::sin(x); }
static inline float m_cos(float x) { return g_fast_libm ? fm::cos(x) : g_pinned_libm ? (float)std::cos((double)x) : std::cos(x); }
static inline float m_atan2(float y, float x) { return g_fast_libm ? fm::atan2(y, x) : g_pinned_libm ? (float)std::atan2((double)y, (double)x) : std::atan2(y, x); }
static inline float m_acos(float x) { return g_fast_libm ? fm::acos(x) : g_pinned_libm ? (float)std::acos((double)x) : std::acos(x); }

// ---- Dubins ------------------------------------------------------------------------------------------
struct DubRes { float len; int type; float p[4]; float c[8]; };   // c = srx sry slx sly grx gry glx gly

float dub_cand(const Port& P, int type, float sh, float gh, float csx, float csy, float cgx, float cgy, float* p)
{
    const float r = P.r_min;
    float dcx = cgx - csx, dcy = cgy - csy;
    float theta = m_atan2(dcy, dcx);
    if (type == 0)        // get_params_rsr, Dubins.cpp:180-210
    {
        p[0] = M_PI_2 + sh; float t1 = M_PI_2 + theta; p[2] = t1; float tg = M_PI_2 + gh;
        p[1] = t1 - p[0]; if (p[1] > 0) p[1] -= 2 * M_PI;
        p[3] = tg - p[2]; if (p[3] > 0) p[3] -= 2 * M_PI;
        float d = std::sqrt(dcx * dcx + dcy * dcy);
        return d + r * -(p[1] + p[3]);
    }
    if (type == 3)        // get_params_lsl, Dubins.cpp:293-323
    {
        p[0] = -M_PI_2 + sh; float t1 = -M_PI_2 + theta; p[2] = t1; float tg = -M_PI_2 + gh;
        p[1] = t1 - p[0]; if (p[1] < 0) p[1] += 2 * M_PI;
        p[3] = tg - p[2]; if (p[3] < 0) p[3] += 2 * M_PI;
        float d = std::sqrt(dcx * dcx + dcy * dcy);
        return d + r * (p[1] + p[3]);
    }
    float dist = std::sqrt(dcx * dcx + dcy * dcy);
    float t1;
    if (type == 1)        // get_params_rsl, Dubins.cpp:212-250
    {
        p[0] = M_PI_2 + sh; t1 = m_acos(2 * r / dist) + theta; p[2] = t1 - M_PI; float tg = -M_PI_2 + gh;
        p[1] = t1 - p[0]; if (p[1] > 0) p[1] -= 2 * M_PI;
        p[3] = tg - p[2]; if (p[3] < 0) p[3] += 2 * M_PI;
    }
    else                  // get_params_lsr, Dubins.cpp:252-291
    {
        p[0] = -M_PI_2 + sh; t1 = -m_acos(2 * r / dist) + theta; p[2] = t1 + M_PI; float tg = M_PI_2 + gh;
        p[1] = t1 - p[0]; if (p[1] < 0) p[1] += 2 * M_PI;
        p[3] = tg - p[2]; if (p[3] > 0) p[3] -= 2 * M_PI;
    }
    float ssx = csx + r * m_cos(t1), ssy = csy + r * m_sin(t1);
    float esx = cgx + r * m_cos(p[2]), esy = cgy + r * m_sin(p[2]);
    float dx = esx - ssx, dy = esy - ssy;
    float d = std::sqrt(dx * dx + dy * dy);
    return (type == 1) ? d + r * (-p[1] + p[3]) : d + r * (p[1] - p[3]);
}

DubRes dub_shortest(const Port& P, const float* s, const float* g)    // Dubins.cpp:19-69
{
    DubRes R;
    const float r = P.r_min;
    R.c[0] = s[0] + r * m_sin(s[2]); R.c[1] = s[1] - r * m_cos(s[2]);
    R.c[2] = s[0] - r * m_sin(s[2]); R.c[3] = s[1] + r * m_cos(s[2]);
    R.c[4] = g[0] + r * m_sin(g[2]); R.c[5] = g[1] - r * m_cos(g[2]);
    R.c[6] = g[0] - r * m_sin(g[2]); R.c[7] = g[1] + r * m_cos(g[2]);
    const int cs[4] = {0, 0, 2, 2}, cg[4] = {4, 6, 4, 6};       // RSR, RSL, LSR, LSL centre pairs
    for (int t = 0; t < 4; t++)
    {
        float p[4];
        float len = dub_cand(P, t, s[2], g[2], R.c[cs[t]], R.c[cs[t] + 1], R.c[cg[t]], R.c[cg[t] + 1], p);
        if (t == 0 || len < R.len) { R.len = len; R.type = t; std::memcpy(R.p, p, sizeof(p)); }
    }
    return R;
}

// sample_path_{rsr,rsl,lsr,lsl}, Dubins.cpp:326-563
void dub_sample(const Port& P, const DubRes& R, std::vector<float>& xyh, std::vector<float>& curv)
{
    const float r = P.r_min;
    const int cs[4] = {0, 0, 2, 2}, cg[4] = {4, 6, 4, 6};
    float csx = R.c[cs[R.type]], csy = R.c[cs[R.type] + 1], cgx = R.c[cg[R.type]], cgy = R.c[cg[R.type] + 1];
    bool r1 = (R.type == 0 || R.type == 1), r2 = (R.type == 0 || R.type == 2);   // right first / second arc
    float ssx = csx + r * m_cos(R.p[0] + R.p[1]), ssy = csy + r * m_sin(R.p[0] + R.p[1]);
    float esx = cgx + r * m_cos(R.p[2]), esy = cgy + r * m_sin(R.p[2]);
    float dx = esx - ssx, dy = esy - ssy;
    float len_st = std::sqrt(dx * dx + dy * dy);
    int size_1 = static_cast<int>(std::floor((r1 ? -R.p[1] : R.p[1]) / P.ang_step));
    int size_2 = size_1 + static_cast<int>(std::floor(len_st / P.step));
    int size_3 = size_2 + static_cast<int>(std::floor((r2 ? -R.p[3] : R.p[3]) / P.ang_step));
    xyh.assign((size_t)(size_3 + 1) * 3, 0.0f); curv.assign(size_3 + 1, 0.0f);
    float theta = R.p[0], kappa = 1 / r;
    for (int i = 0; i < size_1; i++)
    {
        xyh[3 * i] = csx + r * m_cos(theta); xyh[3 * i + 1] = csy + r * m_sin(theta);
        xyh[3 * i + 2] = r1 ? wrap_pi_d(theta - M_PI_2) : wrap_pi_d(theta + M_PI_2);
        curv[i] = kappa;
        if (r1) theta -= P.ang_step; else theta += P.ang_step;
    }
    theta = m_atan2(dy, dx);
    float ct = m_cos(theta), st = m_sin(theta), dist = 0.0f;
    for (int i = size_1; i < size_2; i++)
    {
        xyh[3 * i] = ssx + dist * ct; xyh[3 * i + 1] = ssy + dist * st; xyh[3 * i + 2] = theta; curv[i] = 0.0f;
        dist += P.step;
    }
    theta = R.p[2];
    for (int i = size_2; i < size_3; i++)
    {
        xyh[3 * i] = cgx + r * m_cos(theta); xyh[3 * i + 1] = cgy + r * m_sin(theta);
        xyh[3 * i + 2] = r2 ? wrap_pi_d(theta - M_PI_2) : wrap_pi_d(theta + M_PI_2);
        curv[i] = kappa;
        if (r2) theta -= P.ang_step; else theta += P.ang_step;
    }
    xyh[3 * size_3] = cgx + r * m_cos(R.p[2] + R.p[3]); xyh[3 * size_3 + 1] = cgy + r * m_sin(R.p[2] + R.p[3]);
    xyh[3 * size_3 + 2] = r2 ? wrap_pi_d(R.p[2] + R.p[3] - M_PI_2) : wrap_pi_d(R.p[2] + R.p[3] + M_PI_2);
    curv[size_3] = 0.0f;
}

// ---- successors --------------------------------------------------------------------------------------
// VehicleModel::get_neighbors, VehicleModel.cpp:63-105
bool rollout(const Port& P, const N3& n, std::vector<N3>& out)
{
    int start = n.curv - P.A; if (start < 0) start = 0;
    int num = P.A * 2 + 1;
    out.clear();
    bool neglect = (n.v2 < 1.0f);
    int bin = std::min(n.bin, P.bins - 1 + 1);   // bin == bins is out of bounds in the reference
    for (int i = start; (i < start + num) && (i < P.S); i++)
    {
        float v2 = 0;
        if (!neglect)
        {
            float lat = n.v2 * P.abs_curv[i];
            if (lat > P.max_lat_acc) continue;
            float acc_long = std::sqrt(1.0 - ((lat * lat) / P.max_lat_acc_sqr));
            v2 = n.v2 - 2 * acc_long * P.ts;
        }
        int col = (bin >= P.bins) ? 0 : bin;      // defined here as column 0 (heading -pi == +pi)
        N3 o;
        o.x = n.x + P.off_xy[((size_t)i * P.bins + col) * 2];
        o.y = n.y + P.off_xy[((size_t)i * P.bins + col) * 2 + 1];
        o.h = wrap_pi_f(n.h + P.off_h[i]);
        o.g = n.g + P.act_cost[i]; o.f = o.g; o.v2 = v2; o.curv = i;
        o.bin = heading_index(o.h, P.precision); o.ci = -1; o.cj = -1; o.prev = -1;
        out.push_back(o);
    }
    return neglect;
}

// Grid3D::get_field_intensity, Grid3D.cpp:206-227
float field(const Port& P, float x, float y, float h)
{
    float acc = 0.0f;
    for (size_t k = 0; k < P.apf.size() / 3; k++)
    {
        float ox = P.apf[3 * k], oy = P.apf[3 * k + 1], rad = P.apf[3 * k + 2];
        if (g_fast_libm)                          // K-POP: the same term in FP32 throughout
        {
            float dx = ox - x, dy = oy - y;
            float dist = std::sqrt(dx * dx + dy * dy);
            if (dist < rad)
            {
                float ang = std::abs(wrap_pi_f(h - fm::atan2(dy, dx)));
                ang = std::max(P.apf_alpha - ang, 0.0f);
                float d = 1.0f / dist - 1.0f / rad;
                float fp = P.apf_k * (d * d);
                fp = fp * ang / P.apf_alpha;
                acc = acc + fp;
            }
            continue;
        }
        float distance = std::hypot(ox - x, oy - y);
        float angle = std::abs(wrap_pi_f(h - m_atan2(oy - y, ox - x)));
        angle = std::max(P.apf_alpha - angle, 0.0f);
        float fp = 0;
        if (distance < rad)
        {
            double d = 1.0 / distance - 1.0 / rad;
            fp = P.apf_k * (d * d);
            fp = fp * angle / P.apf_alpha;
        }
        acc = acc + fp;
    }
    return acc;
}

// Grid3D::get_neighbors, Grid3D.cpp:47-74
bool expand(const Port& P, const N3& n, std::vector<N3>& out)
{
    std::vector<N3> all;
    bool safe = rollout(P, n, all);
    out.clear();
    for (auto& o : all)
    {
        int i = static_cast<int>(o.x / P.res), j = static_cast<int>(o.y / P.res);
        if ((i > -1) && (i < P.N) && (j > -1) && (j < P.N) && (P.map[(size_t)i * P.N + j] < P.log_thr))
        {
            float fc = field(P, o.x, o.y, o.h);
            o.g += fc; o.f += fc; o.ci = i; o.cj = j;
            out.push_back(o);
        }
    }
    return safe;
}

// Grid3D::check_path, Grid3D.cpp:78-93
bool check_path(const Port& P, const std::vector<float>& xyh)
{
    for (size_t k = 0; k < xyh.size() / 3; k++)
    {
        int i1 = static_cast<int>(std::round(xyh[3 * k] / P.res)), j1 = static_cast<int>(std::round(xyh[3 * k + 1] / P.res));
        if ((i1 < 0) || (i1 >= P.N) || (j1 < 0) || (j1 >= P.N) || (P.map[(size_t)i1 * P.N + j1] >= P.log_thr)) return false;
    }
    return true;
}

// ---- lazy cached 2D A*: AStar::find_path(i,j) + a_star_search, AStar.cpp:100-113, :118-186 -----------------
float astar_lazy(Port& P, int si, int sj)
{
    const int N = P.N;
    if (P.visited[(size_t)si * N + sj]) return P.nm_f[(size_t)si * N + sj];
    size_t sc = (size_t)si * N + sj;
    P.nm_g[sc] = 0.0f; P.nm_f[sc] = P.nm_h[sc];                       // set_start_node_grid -> soft_reset
    std::set<N2, N2Less> open;
    std::vector<N2> closed;                                            // stable storage of the closed copies
    std::unordered_map<int, int> closed_at;                            // cell -> index in `closed`
    N2 s; s.i = si; s.j = sj; s.g = 0.0f; s.h = P.nm_h[sc]; s.f = s.h; s.prev = -1;
    open.insert(s);
    auto update_visited = [&](float total, int last)                   // AStar.cpp:209-218 + Grid2D.cpp:219-227
    {
        for (int c = last; c >= 0; c = closed[c].prev)
        {
            size_t cell = (size_t)closed[c].i * N + closed[c].j;
            P.visited[cell] = 1;
            P.nm_f[cell] = total - closed[c].g;
        }
    };
    const int gi = P.n45, gj = P.n2;
    while (!open.empty())
    {
        auto it = open.begin();
        int cell = it->i * N + it->j;
        int first;
        auto f = closed_at.find(cell);
        if (f == closed_at.end()) { closed.push_back(*it); first = (int)closed.size() - 1; closed_at[cell] = first; }
        else first = f->second;                                        // unordered_set::insert returns the old copy
        open.erase(it);
        const N2 cur = closed[first];
        if (cur.i == gi && cur.j == gj)
        {
            update_visited(cur.f, first);
            return cur.f;
        }
        float g_first = cur.g;
        struct Nb { int i, j; float w; };
        std::vector<Nb> nbs;                                           // Grid2D::get_neighbors, Grid2D.cpp:72-96
        for (size_t k = 0; k < P.actions.size(); k++)
        {
            int i = cur.i + P.actions[k].first, j = cur.j + P.actions[k].second;
            if ((i > -1) && (i < N) && (j > -1) && (j < N) && P.map[(size_t)i * N + j] < P.log_thr)
                nbs.push_back({i, j, P.actions_cost[k]});
        }
        for (auto& nb : nbs)
        {
            size_t c = (size_t)nb.i * N + nb.j;
            if (P.visited[c])
            {
                float total = P.nm_f[c] + g_first + nb.w;
                update_visited(total, first);
                return total;
            }
            if (closed_at.find((int)c) != closed_at.end()) continue;
            N2 key; key.i = nb.i; key.j = nb.j; key.g = P.nm_g[c]; key.h = P.nm_h[c]; key.f = P.nm_f[c]; key.prev = -1;
            auto it_node = open.find(key);
            if (it_node == open.end())
            {
                P.nm_g[c] = g_first + nb.w; P.nm_f[c] = P.nm_g[c] + P.nm_h[c];
                key.g = P.nm_g[c]; key.f = P.nm_f[c]; key.prev = first;
                open.insert(key);
            }
            else if ((g_first + nb.w) < it_node->g)
            {
                open.erase(it_node);
                P.nm_g[c] = g_first + nb.w; P.nm_f[c] = P.nm_g[c] + P.nm_h[c];
                key.g = P.nm_g[c]; key.f = P.nm_f[c]; key.prev = first;
                open.insert(key);
            }
        }
    }
    return std::numeric_limits<float>::max();
}

// Grid3D::set_start_node, Grid3D.cpp:127-160
N3 set_start(const Port& P, const float* s)
{
    N3 n; std::memset(&n, 0, sizeof(n));
    float rx, ry;
    rotate(s[0] - P.goal_loc[0], s[1] - P.goal_loc[1], P.grid_heading, rx, ry);
    float rh = wrap_pi_f(s[2] - P.grid_heading);
    float px = rx + P.n45 * P.res, py = ry + P.n2 * P.res;
    int i = static_cast<int>(px / P.res), j = static_cast<int>(py / P.res);
    if ((i > -1) && (i < P.N) && (j > -1) && (j < P.N)) { n.x = px; n.y = py; n.h = rh; n.ci = i; n.cj = j; }
    else { n.x = 0; n.y = 0; n.h = 0; n.ci = 0; n.cj = 0; }
    n.bin = heading_index(n.h, P.precision);
    n.curv = P.S / 2; n.prev = -1;
    return n;
}

void scrub(Port& P)
{
    std::fill(P.visited.begin(), P.visited.end(), 0);
    std::fill(P.nm_g.begin(), P.nm_g.end(), 0.0f);
    P.nm_f = P.nm_h;
}
}   // namespace

extern "C"
{
void* port_create(const orc_params* p) { Port* P = new Port(); build(*P, *p); return P; }
void port_destroy(void* h) { delete static_cast<Port*>(h); }

// HybridAStar::update_goal -> Grid3D::update_goal_heading (Grid3D.cpp:102-124) incl. relocate_obstacles (:169-203)
void port_update_goal(void* hv, const float* goal, const float* start)
{
    Port& P = *static_cast<Port*>(hv);
    float heading_prev = P.grid_heading, goal_prev[2] = {P.goal_loc[0], P.goal_loc[1]};
    P.goal_loc[0] = goal[0]; P.goal_loc[1] = goal[1]; P.goal_loc[2] = goal[2];
    P.grid_heading = std::atan2(goal[1] - start[1], goal[0] - start[0]);          // Grid2D.cpp:263
    float dh = P.grid_heading - heading_prev;
    float gpx, gpy, nox, noy;
    rotate(static_cast<float>(P.n45), static_cast<float>(P.n2), dh, gpx, gpy);
    rotate(goal_prev[0] - goal[0], goal_prev[1] - goal[1], P.grid_heading, nox, noy);
    float ox = static_cast<float>(P.n45) + (nox / P.res) - gpx, oy = static_cast<float>(P.n2) + (noy / P.res) - gpy;
    std::vector<float> nm((size_t)P.N * P.N, 0.0f);
    for (int i = 0; i < P.N; i++)
        for (int j = 0; j < P.N; j++)
        {
            float rx, ry;
            rotate(static_cast<float>(i), static_cast<float>(j), dh, rx, ry);
            int in = static_cast<int>(std::round(rx + ox)), jn = static_cast<int>(std::round(ry + oy));
            if ((in > -1) && (in < P.N) && (jn > -1) && (jn < P.N)) nm[(size_t)in * P.N + jn] = P.map[(size_t)i * P.N + j];
        }
    P.map.swap(nm);
    N3& g = P.goal_node;
    g.x = P.n45 * P.res; g.y = P.n2 * P.res; g.h = wrap_pi_f(goal[2] - P.grid_heading);
    g.bin = heading_index(g.h, P.precision); g.ci = P.n45; g.cj = P.n2; g.g = g.f = g.v2 = 0; g.curv = 0; g.prev = -1;
}

void port_reset(void* hv) { Port& P = *static_cast<Port*>(hv); std::fill(P.visited.begin(), P.visited.end(), 0); }   // AStar.cpp:56-60
void port_scrub(void* hv) { scrub(*static_cast<Port*>(hv)); }

// Grid2D::update_obstacles(boxes, conf), Grid2D.cpp:99-139 (scatter form, as in the reference)
void port_update_boxes_2d(void* hv, const float* b, const float* conf, int n)
{
    Port& P = *static_cast<Port*>(hv);
    for (int k = 0; k < n; k++)
    {
        float px = b[4 * k] - b[4 * k + 2] / 2 - P.goal_loc[0], py = b[4 * k + 1] - b[4 * k + 3] / 2 - P.goal_loc[1];
        float c = std::cos(P.grid_heading), s = std::sin(P.grid_heading);
        float x = px * c + py * s, y = -px * s + py * c;
        int start_i = static_cast<int>(std::round(x / P.res) + P.n45), start_j = static_cast<int>(std::round(y / P.res) + P.n2);
        int end_i = static_cast<int>(std::ceil(b[4 * k + 2] / P.res)), end_j = static_cast<int>(std::ceil(b[4 * k + 3] / P.res));
        float log_conf = std::log(conf[k] / (1.0 - conf[k]));
        for (int i = 0; i < 2 * end_i; i++)
            for (int j = 0; j < 2 * end_j; j++)
            {
                float ox = i * 0.5, oy = j * 0.5;
                float rx = ox * c + oy * s, ry = -ox * s + oy * c;
                int ip = start_i + static_cast<int>(std::round(rx)), jp = start_j + static_cast<int>(std::round(ry));
                if ((ip > -1) && (ip < P.N) && (jp > -1) && (jp < P.N))
                {
                    float& v = P.map[(size_t)ip * P.N + jp];
                    v += log_conf - P.log_free;
                    v = clampv(P, v);
                }
            }
    }
}

// Grid3D::update_obstacles, Grid3D.cpp:22-44
void port_update_boxes(void* hv, const float* b, const float* conf, int n, float added)
{
    Port& P = *static_cast<Port*>(hv);
    P.apf.resize((size_t)n * 3);
    for (int k = 0; k < n; k++)
    {
        float rx, ry;
        rotate(b[4 * k] - P.goal_loc[0], b[4 * k + 1] - P.goal_loc[1], P.grid_heading, rx, ry);
        rx += P.n45 * P.res; ry += P.n2 * P.res;
        P.apf[3 * k] = rx; P.apf[3 * k + 1] = ry; P.apf[3 * k + 2] = std::max(b[4 * k + 2], b[4 * k + 3]) / 2 + added;
    }
    port_update_boxes_2d(hv, b, conf, n);
}

// Grid2D::update_obstacles(lines, conf, width), Grid2D.cpp:142-194
void port_update_lines(void* hv, const float* l, const float* conf, int n, float width)
{
    Port& P = *static_cast<Port*>(hv);
    for (int k = 0; k < n; k++)
    {
        float sx, sy, ex, ey;
        rotate(l[4 * k] - P.goal_loc[0], l[4 * k + 1] - P.goal_loc[1], P.grid_heading, sx, sy);
        rotate(l[4 * k + 2] - P.goal_loc[0], l[4 * k + 3] - P.goal_loc[1], P.grid_heading, ex, ey);
        float dx = ex - sx, dy = ey - sy;
        float len = std::hypot(dx, dy);
        float nx = -dy / len, ny = dx / len;
        dx = dx / len; dy = dy / len;
        float prog = 0;
        float log_conf = std::log(conf[k] / (1.0 - conf[k]));
        size_t iter = 0;
        while ((prog <= len) && (iter < 100))
        {
            float pw = 0;
            float ix = sx + dx * prog, iy = sy + dy * prog;
            while (pw <= width)
            {
                float p1x = ix + nx * pw, p1y = iy + ny * pw, p2x = ix - nx * pw, p2y = iy - ny * pw;
                int i1 = static_cast<int>(std::round(p1x / P.res)) + P.n45, i2 = static_cast<int>(std::round(p2x / P.res)) + P.n45;
                int j1 = static_cast<int>(std::round(p1y / P.res)) + P.n2, j2 = static_cast<int>(std::round(p2y / P.res)) + P.n2;
                if ((i1 > -1) && (i1 < P.N) && (j1 > -1) && (j1 < P.N))
                { float& v = P.map[(size_t)i1 * P.N + j1]; v += log_conf - P.log_free; v = clampv(P, v); }
                if ((i2 > -1) && (i2 < P.N) && (j2 > -1) && (j2 < P.N))
                { float& v = P.map[(size_t)i2 * P.N + j2]; v += log_conf - P.log_free; v = clampv(P, v); }
                pw += P.res;
            }
            prog += P.res;
            iter++;
        }
    }
}

// Grid2D::update_obstacles(), Grid2D.cpp:197-208
void port_decay(void* hv)
{
    Port& P = *static_cast<Port*>(hv);
    for (auto& v : P.map) { v += P.log_free; v = clampv(P, v); }
}

void port_get_map(void* hv, float* out) { Port& P = *static_cast<Port*>(hv); std::memcpy(out, P.map.data(), P.map.size() * 4); }
void port_set_map(void* hv, const float* in) { Port& P = *static_cast<Port*>(hv); std::memcpy(P.map.data(), in, P.map.size() * 4); }

void port_get_consts(void* hv, orc_consts* c)
{
    Port& P = *static_cast<Port*>(hv);
    c->log_threshold = P.log_thr; c->log_min = P.log_min; c->log_max = P.log_max; c->log_free = P.log_free;
    c->grid_heading = P.grid_heading;
    for (int q = 0; q < 3; q++) c->goal_world[q] = P.goal_loc[q];
    c->goal_grid[0] = P.goal_node.x; c->goal_grid[1] = P.goal_node.y; c->goal_grid[2] = P.goal_node.h;
    c->goal_bin = P.goal_node.bin; c->goal_ci = P.n45; c->goal_cj = P.n2;
    c->precision = P.precision; c->r_min = P.r_min; c->ang_step = P.ang_step; c->num_apf = (int)(P.apf.size() / 3);
}
void port_get_apf(void* hv, float* out) { Port& P = *static_cast<Port*>(hv); std::memcpy(out, P.apf.data(), P.apf.size() * 4); }
void port_get_tables(void* hv, float* oxy, float* oh, float* ac, float* cu)
{
    Port& P = *static_cast<Port*>(hv);
    std::memcpy(oxy, P.off_xy.data(), P.off_xy.size() * 4);
    for (int i = 0; i < P.S; i++) { oh[i] = P.off_h[i]; ac[i] = P.act_cost[i]; cu[i] = P.abs_curv[i]; }
}

static void to_state(const N3& n, orc_state& s)
{
    s.x = n.x; s.y = n.y; s.heading = n.h; s.g = n.g; s.f = n.f; s.vmin_sqr = n.v2; s.curvature_index = n.curv;
    s.angle_bin = n.bin; s.ci = n.ci; s.cj = n.cj;
}
static N3 from_state(const orc_state& s)
{
    N3 n; n.x = s.x; n.y = s.y; n.h = s.heading; n.g = s.g; n.f = s.f; n.v2 = s.vmin_sqr; n.curv = s.curvature_index;
    n.bin = s.angle_bin; n.ci = s.ci; n.cj = s.cj; n.prev = -1; return n;
}

void port_set_start(void* hv, const float* s, orc_state* out)
{
    N3 n = set_start(*static_cast<Port*>(hv), s);
    to_state(n, *out);
}

void port_rollout_batch(void* hv, const orc_state* in, int n, orc_state* out, int* n_out, int* flags)
{
    Port& P = *static_cast<Port*>(hv);
    int stride = 2 * P.A + 1;
    std::vector<N3> nb;
    for (int k = 0; k < n; k++)
    {
        flags[k] = rollout(P, from_state(in[k]), nb) ? 1 : 0;
        n_out[k] = (int)nb.size();
        for (size_t q = 0; q < nb.size(); q++) to_state(nb[q], out[(size_t)k * stride + q]);
    }
}
void port_expand_batch(void* hv, const orc_state* in, int n, orc_state* out, int* n_out, int* flags)
{
    Port& P = *static_cast<Port*>(hv);
    int stride = 2 * P.A + 1;
    std::vector<N3> nb;
    for (int k = 0; k < n; k++)
    {
        flags[k] = expand(P, from_state(in[k]), nb) ? 1 : 0;
        n_out[k] = (int)nb.size();
        for (size_t q = 0; q < nb.size(); q++) to_state(nb[q], out[(size_t)k * stride + q]);
    }
}
void port_apf_batch(void* hv, const float* xyh, int n, float* out)
{
    Port& P = *static_cast<Port*>(hv);
    for (int k = 0; k < n; k++) out[k] = field(P, xyh[3 * k], xyh[3 * k + 1], xyh[3 * k + 2]);
}
int port_check_path(void* hv, const float* xyh, int n)
{
    return check_path(*static_cast<Port*>(hv), std::vector<float>(xyh, xyh + 3 * (size_t)n)) ? 1 : 0;
}
void port_dubins_length_batch(void* hv, const float* starts, int n, const float* goal, float* len, int* type, float* params4)
{
    Port& P = *static_cast<Port*>(hv);
    for (int k = 0; k < n; k++)
    {
        DubRes R = dub_shortest(P, starts + 3 * k, goal);
        len[k] = R.len;
        if (type) type[k] = R.type;
        if (params4) std::memcpy(params4 + 4 * k, R.p, 16);
    }
}
int port_dubins_path(void* hv, const float* s, const float* g, float* xyh, float* curv, int cap, float* length, int* flag)
{
    Port& P = *static_cast<Port*>(hv);
    DubRes R = dub_shortest(P, s, g);
    std::vector<float> pts, cv;
    dub_sample(P, R, pts, cv);
    *length = R.len; *flag = (std::abs(R.p[1]) > static_cast<float>(M_PI_2)) ? 1 : 0;   // Dubins.cpp:152
    int n = (int)cv.size();
    for (int k = 0; k < n && k < cap; k++) { std::memcpy(xyh + 3 * k, &pts[3 * k], 12); curv[k] = cv[k]; }
    return n;
}
void port_astar_lazy_batch(void* hv, const int* ij, int n, float* out)
{
    Port& P = *static_cast<Port*>(hv);
    for (int k = 0; k < n; k++) out[k] = astar_lazy(P, ij[2 * k], ij[2 * k + 1]);
}
void port_astar_dump(void* hv, unsigned char* visited, float* g, float* f)
{
    Port& P = *static_cast<Port*>(hv);
    size_t nn = (size_t)P.N * P.N;
    if (visited) std::memcpy(visited, P.visited.data(), nn);
    if (g) std::memcpy(g, P.nm_g.data(), nn * 4);
    if (f) std::memcpy(f, P.nm_f.data(), nn * 4);
}

// Exact 2D distance-to-goal field in double precision (Dijkstra) over the metric of Grid2D (Grid2D.cpp:32-58, :72-96):
// oracle for the device wavefront field (north_star (e)).  NOT the value AStar::find_path returns.
// out[i*N + j] = distance, or a negative value when unreachable.
void port_field2d(void* hv, double* out)
{
    Port& P = *static_cast<Port*>(hv);
    const int N = P.N;
    std::vector<double> d((size_t)N * N, -1.0);
    typedef std::pair<double, int> QE;
    std::vector<QE> heap;
    auto cmp = [](const QE& a, const QE& b) { return a.first > b.first; };
    int goal = P.n45 * N + P.n2;
    d[goal] = 0.0;
    heap.push_back(QE(0.0, goal));
    while (!heap.empty())
    {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        QE e = heap.back(); heap.pop_back();
        if (e.first > d[e.second]) continue;
        int ui = e.second / N, uj = e.second % N;
        for (size_t k = 0; k < P.actions.size(); k++)
        {
            int i = ui + P.actions[k].first, j = uj + P.actions[k].second;
            if (i < 0 || i >= N || j < 0 || j >= N) continue;
            if (!(P.map[(size_t)i * N + j] < P.log_thr)) continue;
            double nd = e.first + (double)P.actions_cost[k];
            int v = i * N + j;
            if (d[v] < 0.0 || nd < d[v]) { d[v] = nd; heap.push_back(QE(nd, v)); std::push_heap(heap.begin(), heap.end(), cmp); }
        }
    }
    std::memcpy(out, d.data(), sizeof(double) * d.size());
}

// HybridAStar::find_path + hybrid_a_star_search + reconstruct_path, HybridAStar.cpp:68-88, :93-199, :208-262
void port_find_path(void* hv, float vel, const float* s, orc_result* res, float* path_xyh, float* curv, int path_cap,
                    orc_pop* pops, int pop_cap)
{
    Port& P = *static_cast<Port*>(hv);
    N3 start = set_start(P, s);
    start.v2 = vel * vel; start.f = std::numeric_limits<float>::max(); start.g = 0;
    const N3 goal = P.goal_node;
    const float goal3[3] = {goal.x, goal.y, goal.h};
    int counter = 0, interval = P.p.shot_interval;
    bool allowed = false, shot_ok = false, success = false;
    float cost = std::numeric_limits<float>::max();
    std::set<N3, N3Less> open;
    std::vector<N3> closed;
    std::unordered_map<long long, int> closed_at;                      // (cell, bin) -> closed index
    auto key_of = [&](const N3& n) { return ((long long)n.ci * P.N + n.cj) * (P.bins + 1) + n.bin; };
    open.insert(start);
    std::vector<N3> nbrs;
    std::vector<float> dub_xyh, dub_curv;
    int terminal = -1;
    long n_pops = 0, n_oob = 0;
    while (!open.empty())
    {
        auto it = open.begin();
        long long key = key_of(*it);
        int first;
        auto f = closed_at.find(key);
        if (f == closed_at.end()) { closed.push_back(*it); first = (int)closed.size() - 1; closed_at[key] = first; }
        else first = f->second;
        open.erase(it);
        const N3 cur = closed[first];
        if (cur.ci == goal.ci && cur.cj == goal.cj)                    // Node3D operator== : cell only
        {
            terminal = first; success = true; cost = cur.g;
            break;
        }
        else if (allowed)
        {
            counter++;
            if (counter == interval)
            {
                const float st[3] = {cur.x, cur.y, cur.h};
                DubRes R = dub_shortest(P, st, goal3);
                dub_sample(P, R, dub_xyh, dub_curv);
                bool long_turn = std::abs(R.p[1]) > static_cast<float>(M_PI_2);
                if (!long_turn && check_path(P, dub_xyh))
                {
                    terminal = cur.prev; shot_ok = true; success = true; cost = cur.g + R.len;
                    break;
                }
                counter = 0;
                interval = std::max(interval - P.p.shot_decay, 50);
            }
        }
        if (pops && n_pops < pop_cap)
        {
            orc_pop& r = pops[n_pops];
            r.ci = cur.ci; r.cj = cur.cj; r.bin = cur.bin; r.x = cur.x; r.y = cur.y; r.heading = cur.h; r.g = cur.g; r.f = cur.f;
        }
        n_pops++;
        if (cur.bin >= P.bins) n_oob++;
        allowed = expand(P, cur, nbrs);
        for (auto& node : nbrs)
        {
            if (closed_at.find(key_of(node)) != closed_at.end()) continue;
            auto it_node = open.find(node);
            if (it_node == open.end())
            {
                float h1 = astar_lazy(P, node.ci, node.cj);
                const float st[3] = {node.x, node.y, node.h};
                float h2 = dub_shortest(P, st, goal3).len;
                node.f += std::max(h1, h2);
                node.prev = first;
                open.insert(node);
            }
            else if (node.g < it_node->g)
            {
                open.erase(it_node);
                float h1 = astar_lazy(P, node.ci, node.cj);
                const float st[3] = {node.x, node.y, node.h};
                float h2 = dub_shortest(P, st, goal3).len;
                node.f += std::max(h1, h2);
                node.prev = first;
                open.insert(node);
            }
        }
    }
    res->success = success ? 1 : 0; res->cost = cost; res->n_pops = (int)n_pops; res->n_pops_bin_oob = (int)n_oob;
    int n = 0;
    if (success)
    {
        // reconstruct_path, HybridAStar.cpp:208-262
        std::vector<float> pts, cv;
        cv.push_back(0.0f);
        auto to_world = [&](float x, float y, float h)
        {
            float rx = x - goal.x, ry = y - goal.y, a = -P.grid_heading;
            float c = std::cos(a), sn = std::sin(a);
            float wx = rx * c + ry * sn, wy = -rx * sn + ry * c, wh = wrap_pi_f(h - a);
            pts.push_back(wx + P.goal_loc[0]); pts.push_back(wy + P.goal_loc[1]); pts.push_back(wh);
        };
        if (shot_ok)
            for (int k = (int)dub_curv.size() - 1; k >= 0; k--)
            {
                to_world(dub_xyh[3 * k], dub_xyh[3 * k + 1], dub_xyh[3 * k + 2]);
                cv.push_back(dub_curv[k]);
            }
        for (int c = terminal; c >= 0; c = closed[c].prev)
        {
            to_world(closed[c].x, closed[c].y, closed[c].h);
            cv.push_back(P.abs_curv[closed[c].curv]);
        }
        cv.pop_back();
        n = (int)cv.size();
        for (int k = 0; k < n && k < path_cap; k++) { std::memcpy(path_xyh + 3 * k, &pts[3 * k], 12); curv[k] = cv[k]; }
    }
    res->n_path = n;
}

#include "footprint.inc"
}   // extern "C"

#include "kpop.inc"
