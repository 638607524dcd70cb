// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// C-ABI driver around the UNMODIFIED reference library (/root/reference/lib/*.cpp, compiled where
// they lie by oracle/Makefile into oracle/_ref/libref_oracle.so).  Nothing of the reference is
// copied: this file only *calls* the reference classes.
//
// Two tricks keep the reference sources untouched:
//   * `#define private public` around the reference headers so the driver can read internals
//     (_grid, _astar, _dubins, _node_map ...).  Only access control changes, not layout.
//   * link-time `-Wl,--wrap=` on Grid3D<float>::get_neighbors: HybridAStar.o calls it across a
//     translation-unit boundary once per expanded node (HybridAStar.cpp:157), so the wrapper
//     sees the exact pop sequence of the real search loop without replicating the loop.
//
// Oracle hygiene: AStar::reset() does not clear _node_map costs, so ref_scrub()
// restores the freshly-constructed state (g=0, f=h, prev=null, _visted cleared).

#include <vector>
#include <set>
#include <unordered_set>
#include <utility>
#include <limits>
#include <algorithm>
#include <numeric>
#include <array>
#include <string>
#include <cmath>
#include <cstring>
#include <iostream>
#include <functional>
#include <thread>
#include <chrono>
#include <atomic>
#include <memory>

#define private public
#define protected public
#include "HybridAStar.h"
#include "VelocityGenerator.h"
#undef private
#undef protected

#include "oracle_api.h"

using namespace planning;

namespace
{
    struct TraceCtx
    {
        orc_pop* buf   = nullptr;
        int      cap   = 0;
        long     n     = 0;
        long     n_oob = 0;
        int      bins  = 0;
        bool     active = false;
    };
    thread_local TraceCtx g_trace;

    struct Handle
    {
        orc_params p;
        std::unique_ptr<HybridAStar<float>> planner;
    };

    std::unique_ptr<HybridAStar<float>> make_planner(const orc_params& p)
    {
        std::vector<float> steering(p.steering, p.steering + p.num_steering);
        // the reference accepts a weights vector longer than steering (utils/hybrid_astar/test_hybrid_astar.cpp:33)
        std::vector<float> weights(p.curvature_weights, p.curvature_weights + p.num_steering);
        return std::unique_ptr<HybridAStar<float>>(new HybridAStar<float>(
            p.shot_interval, p.shot_decay, p.resolution, p.obstacle_threshold, p.prob_min, p.prob_max,
            p.prob_free, p.grid_size, p.allow_diag != 0, p.step_size, p.max_lat_acc, p.max_long_dec,
            p.wheelbase, p.rear_to_cg, p.apf_rep_constant, p.apf_active_angle, p.num_angle_bins,
            p.num_actions, steering, weights));
    }

    void to_state(const Node3D<float>& n, orc_state& s)
    {
        s.x = n._pose2D._x; s.y = n._pose2D._y; s.heading = n._pose2D._heading;
        s.g = n._cost_g; s.f = n._cost_f; s.vmin_sqr = n._vmin_sqr;
        s.curvature_index = n._curvature_index; s.angle_bin = n._angle_bin;
        if (n._base_node) { s.ci = n._base_node->_posd._x; s.cj = n._base_node->_posd._y; }
        else              { s.ci = -1; s.cj = -1; }
    }

    Node3D<float> from_state(Handle* h, const orc_state& s)
    {
        Vector3D<float> pose(s.x, s.y, s.heading);
        Node3D<float> n(pose, s.g, s.vmin_sqr, s.curvature_index, s.angle_bin, nullptr, nullptr);
        n._cost_f = s.f;
        int N = h->p.grid_size;
        if (s.ci >= 0 && s.ci < N && s.cj >= 0 && s.cj < N)
            n._base_node = &h->planner->_grid._node_map[s.ci][s.cj];
        return n;
    }

    std::vector<Obstacle<float>> to_boxes(const float* b, int n)
    {
        std::vector<Obstacle<float>> v;
        v.reserve(n);
        for (int k = 0; k < n; k++) v.emplace_back(b[4 * k], b[4 * k + 1], b[4 * k + 2], b[4 * k + 3]);
        return v;
    }
}

// ---- link-time hook: one call per expanded node of the real search loop -------------------------
extern "C" bool __real__ZNK8planning6Grid3DIfE13get_neighborsERKNS_6Node3DIfEERSt6vectorIS3_SaIS3_EE(
    const Grid3D<float>* self, const Node3D<float>& node, std::vector<Node3D<float>>& nbrs);

extern "C" bool __wrap__ZNK8planning6Grid3DIfE13get_neighborsERKNS_6Node3DIfEERSt6vectorIS3_SaIS3_EE(
    const Grid3D<float>* self, const Node3D<float>& node, std::vector<Node3D<float>>& nbrs)
{
    TraceCtx& t = g_trace;
    if (t.active)
    {
        if (t.buf && t.n < t.cap)
        {
            orc_pop& r = t.buf[t.n];
            r.ci = node._base_node ? node._base_node->_posd._x : -1;
            r.cj = node._base_node ? node._base_node->_posd._y : -1;
            r.bin = node._angle_bin;
            r.x = node._pose2D._x; r.y = node._pose2D._y; r.heading = node._pose2D._heading;
            r.g = node._cost_g; r.f = node._cost_f;
        }
        t.n++;
        if (node._angle_bin >= t.bins) t.n_oob++;
    }
    return __real__ZNK8planning6Grid3DIfE13get_neighborsERKNS_6Node3DIfEERSt6vectorIS3_SaIS3_EE(self, node, nbrs);
}

extern "C"
{

void* ref_create(const orc_params* p)
{
    Handle* h = new Handle();
    h->p = *p;
    h->planner = make_planner(*p);
    return h;
}

void ref_destroy(void* hv) { delete static_cast<Handle*>(hv); }

void ref_update_goal(void* hv, const float* goal3, const float* start3)
{
    Handle* h = static_cast<Handle*>(hv);
    h->planner->update_goal(Vector3D<float>(goal3[0], goal3[1], goal3[2]),
                            Vector3D<float>(start3[0], start3[1], start3[2]));
}

void ref_reset(void* hv) { static_cast<Handle*>(hv)->planner->reset(); }

// restore the freshly constructed A* cache state
void ref_scrub(void* hv)
{
    Handle* h = static_cast<Handle*>(hv);
    auto& nm = h->planner->_grid._node_map;
    for (auto& row : nm)
        for (auto& n : row) { n._cost_g = 0.0f; n._cost_f = n._cost_h; n._prev = nullptr; }
    h->planner->reset();
}

void ref_update_boxes(void* hv, const float* boxes, const float* conf, int n, float apf_added_radius)
{
    Handle* h = static_cast<Handle*>(hv);
    h->planner->update_obstacles(to_boxes(boxes, n), std::vector<float>(conf, conf + n), apf_added_radius);
}

// Grid2D-only rasteriser (no APF list rebuild), Grid2D.cpp:99-139
void ref_update_boxes_2d(void* hv, const float* boxes, const float* conf, int n)
{
    Handle* h = static_cast<Handle*>(hv);
    static_cast<Grid2D<float>&>(h->planner->_grid).update_obstacles(to_boxes(boxes, n), std::vector<float>(conf, conf + n));
}

void ref_update_lines(void* hv, const float* l, const float* conf, int n, float width)
{
    Handle* h = static_cast<Handle*>(hv);
    std::vector<std::pair<Vector2D<float>, Vector2D<float>>> lines;
    for (int k = 0; k < n; k++)
        lines.emplace_back(Vector2D<float>(l[4 * k], l[4 * k + 1]), Vector2D<float>(l[4 * k + 2], l[4 * k + 3]));
    h->planner->update_obstacles(lines, std::vector<float>(conf, conf + n), width);
}

void ref_decay(void* hv) { static_cast<Handle*>(hv)->planner->update_obstacles(); }

void ref_get_map(void* hv, float* out)
{
    Handle* h = static_cast<Handle*>(hv);
    const auto& m = h->planner->get_obstacles();
    int N = h->p.grid_size;
    for (int i = 0; i < N; i++) std::memcpy(out + (size_t)i * N, m[i].data(), sizeof(float) * N);
}

void ref_set_map(void* hv, const float* in)
{
    Handle* h = static_cast<Handle*>(hv);
    auto& m = h->planner->_grid._obstacle_map;
    int N = h->p.grid_size;
    for (int i = 0; i < N; i++) std::memcpy(m[i].data(), in + (size_t)i * N, sizeof(float) * N);
}

void ref_get_consts(void* hv, orc_consts* c)
{
    Handle* h = static_cast<Handle*>(hv);
    auto& g = h->planner->_grid;
    c->log_threshold = g._obstacle_log_threshold;
    c->log_min = g._obstacle_log_prob_min;
    c->log_max = g._obstacle_log_prob_max;
    c->log_free = g._obstacle_log_prob_free;
    c->grid_heading = g._grid_heading;
    c->goal_world[0] = g._goal_location3D._x; c->goal_world[1] = g._goal_location3D._y; c->goal_world[2] = g._goal_location3D._heading;
    const Node3D<float>& gn = h->planner->_goal_node;
    c->goal_grid[0] = gn._pose2D._x; c->goal_grid[1] = gn._pose2D._y; c->goal_grid[2] = gn._pose2D._heading;
    c->goal_bin = gn._angle_bin;
    c->goal_ci = g._grid_size_4_5; c->goal_cj = g._grid_size_2;
    c->precision = g._model._precision;
    c->r_min = h->planner->_dubins._r_min;
    c->ang_step = h->planner->_dubins._ang_step_size;
    c->num_apf = (int)g._apf_obstacles.size();
}

void ref_get_apf(void* hv, float* out)
{
    Handle* h = static_cast<Handle*>(hv);
    auto& v = h->planner->_grid._apf_obstacles;
    for (size_t k = 0; k < v.size(); k++) { out[3 * k] = v[k].first._x; out[3 * k + 1] = v[k].first._y; out[3 * k + 2] = v[k].second; }
}

void ref_get_tables(void* hv, float* offset_xy, float* offset_heading, float* actions_cost, float* abs_curv)
{
    Handle* h = static_cast<Handle*>(hv);
    auto& m = h->planner->_grid._model;
    int S = h->p.num_steering, B = h->p.num_angle_bins;
    for (int i = 0; i < S; i++)
    {
        offset_heading[i] = m._offset_heading[i];
        actions_cost[i] = m._actions_cost[i];
        abs_curv[i] = m._abs_curvatures[i];
        for (int j = 0; j < B; j++)
        {
            offset_xy[((size_t)i * B + j) * 2] = m._offset_xy[i][j]._x;
            offset_xy[((size_t)i * B + j) * 2 + 1] = m._offset_xy[i][j]._y;
        }
    }
}

void ref_set_start(void* hv, const float* s, orc_state* out)
{
    Handle* h = static_cast<Handle*>(hv);
    Node3D<float> n = h->planner->_grid.set_start_node(Vector3D<float>(s[0], s[1], s[2]));
    to_state(n, *out);
}

// VehicleModel::get_neighbors (VehicleModel.cpp:63-105) for n states; out has room for n*(2A+1)
// successors, n_out[k] = count for state k, flags[k] = neglect_acceleration
void ref_rollout_batch(void* hv, const orc_state* in, int n, orc_state* out, int* n_out, int* flags)
{
    Handle* h = static_cast<Handle*>(hv);
    int stride = 2 * h->p.num_actions + 1;
    std::vector<Node3D<float>> nbrs;
    for (int k = 0; k < n; k++)
    {
        Node3D<float> node = from_state(h, in[k]);
        bool fl = h->planner->_grid._model.get_neighbors(node, nbrs);
        flags[k] = fl ? 1 : 0;
        n_out[k] = (int)nbrs.size();
        for (size_t s = 0; s < nbrs.size(); s++) to_state(nbrs[s], out[(size_t)k * stride + s]);
    }
}

// Grid3D::get_neighbors (Grid3D.cpp:47-74): roll-out + bounds + collision lookup + APF
void ref_expand_batch(void* hv, const orc_state* in, int n, orc_state* out, int* n_out, int* flags)
{
    Handle* h = static_cast<Handle*>(hv);
    int stride = 2 * h->p.num_actions + 1;
    std::vector<Node3D<float>> nbrs;
    for (int k = 0; k < n; k++)
    {
        Node3D<float> node = from_state(h, in[k]);
        bool fl = h->planner->_grid.get_neighbors(node, nbrs);
        flags[k] = fl ? 1 : 0;
        n_out[k] = (int)nbrs.size();
        for (size_t s = 0; s < nbrs.size(); s++) to_state(nbrs[s], out[(size_t)k * stride + s]);
    }
}

// Grid3D::get_field_intensity (Grid3D.cpp:206-227)
void ref_apf_batch(void* hv, const float* xyh, int n, float* out)
{
    Handle* h = static_cast<Handle*>(hv);
    for (int k = 0; k < n; k++)
    {
        Vector3D<float> pose(xyh[3 * k], xyh[3 * k + 1], xyh[3 * k + 2]);
        Node3D<float> node(pose, 0.0f, 0.0f, 0, 0, nullptr, nullptr);
        out[k] = h->planner->_grid.get_field_intensity(node);
    }
}

// Grid3D::check_path (Grid3D.cpp:78-93), 1 = collision free
int ref_check_path(void* hv, const float* xyh, int n)
{
    Handle* h = static_cast<Handle*>(hv);
    std::vector<Vector3D<float>> path;
    for (int k = 0; k < n; k++) path.emplace_back(xyh[3 * k], xyh[3 * k + 1], xyh[3 * k + 2]);
    return h->planner->_grid.check_path(path) ? 1 : 0;
}

// Dubins::get_shortest_path_length (Dubins.cpp:19-69) for n starts and one goal
void ref_dubins_length_batch(void* hv, const float* starts, int n, const float* goal3, float* len, int* type, float* params4)
{
    Handle* h = static_cast<Handle*>(hv);
    Vector3D<float> goal(goal3[0], goal3[1], goal3[2]);
    auto& d = h->planner->_dubins;
    for (int k = 0; k < n; k++)
    {
        Vector3D<float> s(starts[3 * k], starts[3 * k + 1], starts[3 * k + 2]);
        len[k] = d.get_shortest_path_length(s, goal);
        if (type) type[k] = (int)d._path_type;
        if (params4) for (int q = 0; q < 4; q++) params4[4 * k + q] = d._params[q];
    }
}

// Dubins::get_shortest_path (Dubins.cpp:125-153): returns number of samples (may exceed cap; only cap written)
int ref_dubins_path(void* hv, const float* start3, const float* goal3, float* xyh, float* curv, int cap, float* length, int* flag)
{
    Handle* h = static_cast<Handle*>(hv);
    std::vector<Vector3D<float>> path;
    std::vector<float> c;
    auto r = h->planner->_dubins.get_shortest_path(Vector3D<float>(start3[0], start3[1], start3[2]),
                                                   Vector3D<float>(goal3[0], goal3[1], goal3[2]), path, c);
    *length = r.first;
    *flag = r.second ? 1 : 0;
    int n = (int)path.size();
    for (int k = 0; k < n && k < cap; k++)
    {
        xyh[3 * k] = path[k]._x; xyh[3 * k + 1] = path[k]._y; xyh[3 * k + 2] = path[k]._heading;
        curv[k] = c[k];
    }
    return n;
}

// AStar::find_path(i, j) (AStar.cpp:100-113) called in sequence: stateful lazy cache
void ref_astar_lazy_batch(void* hv, const int* ij, int n, float* out)
{
    Handle* h = static_cast<Handle*>(hv);
    for (int k = 0; k < n; k++) out[k] = h->planner->_astar.find_path(ij[2 * k], ij[2 * k + 1]);
}

// dump of the lazy-A* cache: visited flags and node_map g/f
void ref_astar_dump(void* hv, unsigned char* visited, float* g, float* f)
{
    Handle* h = static_cast<Handle*>(hv);
    int N = h->p.grid_size;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
        {
            size_t c = (size_t)i * N + j;
            if (visited) visited[c] = h->planner->_astar._visted[i][j] ? 1 : 0;
            if (g) g[c] = h->planner->_grid._node_map[i][j]._cost_g;
            if (f) f[c] = h->planner->_grid._node_map[i][j]._cost_f;
        }
}

// HybridAStar::find_path (HybridAStar.cpp:68-88) on the planner as it stands (caller scrubs first)
void ref_find_path(void* hv, float vel, const float* s, orc_result* res, float* path_xyh, float* curv,
                   int path_cap, orc_pop* pops, int pop_cap)
{
    Handle* h = static_cast<Handle*>(hv);
    std::vector<Vector3D<float>> path;
    std::vector<float> c;
    g_trace.buf = pops; g_trace.cap = pop_cap; g_trace.n = 0; g_trace.n_oob = 0;
    g_trace.bins = h->p.num_angle_bins; g_trace.active = true;
    auto r = h->planner->find_path(vel, Vector3D<float>(s[0], s[1], s[2]), path, c);
    g_trace.active = false;
    res->success = r.second ? 1 : 0;
    res->cost = r.first;
    res->n_path = (int)path.size();
    res->n_pops = (int)g_trace.n;
    res->n_pops_bin_oob = (int)g_trace.n_oob;
    for (int k = 0; k < (int)path.size() && k < path_cap; k++)
    {
        path_xyh[3 * k] = path[k]._x; path_xyh[3 * k + 1] = path[k]._y; path_xyh[3 * k + 2] = path[k]._heading;
        curv[k] = c[k];
    }
}

// VelocityGenerator<float>::generate_velocity_profile (lib/VelocityGenerator.cpp:19-85) on a caller-supplied path
// (goal -> start order, as find_path returns it).  lim5 = the five constructor arguments.  Returns the feasibility flag.
int ref_velocity_profile(const float* lim5, float vel_init, float max_velocity_curr, const float* path_xyh, const float* curv, int n,
                         int coast, int stop, float* vel_out)
{
    VelocityGenerator<float> vg(lim5[0], lim5[1], lim5[2], lim5[3], lim5[4]);
    std::vector<Vector3D<float>> path;
    for (int k = 0; k < n; k++) path.emplace_back(path_xyh[3 * k], path_xyh[3 * k + 1], path_xyh[3 * k + 2]);
    std::vector<float> c(curv, curv + n), v;
    bool ok = vg.generate_velocity_profile(vel_init, max_velocity_curr, path, c, v, coast != 0, stop != 0);
    for (int k = 0; k < n; k++) vel_out[k] = v[k];
    return ok ? 1 : 0;
}

// ---- CPU baseline: many queries on one map, fresh-state planner per thread, scrub per query ------
// queries: n x (x, y, heading, vel); group maps are applied by the caller through `maps` (n_groups x N*N),
// goal/start frames through `frames` (n_groups x 6: goal3, start3) and APF lists rebuilt from `boxes`.
// Returns wall seconds; fills per-query cost/success/pops.
// Polynomial hash (mod 2^64) over the IEEE words of the returned path (x, y, heading per point) followed by its curvature:
// h = sum_k (w_k + 1) * P^(k+1).  The per-query path identity the parity checks compare; tests/orc.py::path_hash computes the
// same, vectorised, over the device's output.
static unsigned long long path_hash(const std::vector<Vector3D<float>>& path, const std::vector<float>& c)
{
    const unsigned long long P = 1099511628211ull;
    unsigned long long h = 0, pw = 1;
    auto mix = [&](float v) { unsigned u; std::memcpy(&u, &v, 4); pw *= P; h += ((unsigned long long)u + 1ull) * pw; };
    for (size_t k = 0; k < path.size(); k++) { mix(path[k]._x); mix(path[k]._y); mix(path[k]._heading); }
    for (size_t k = 0; k < c.size(); k++) mix(c[k]);
    return h;
}

// n_queries independent find_path calls on n_threads reference planners (one per thread, scrubbed per query).
// Optional per-query outputs: cost, success, pops, pops in heading bin == num_angle_bins, path points,
// path hash, seconds spent inside find_path.  Returns the wall-clock seconds of the whole batch.
double ref_bench_queries_ex(const orc_params* p, const float* frames6, const float* maps, const float* boxes,
                            const float* conf, int n_boxes, float apf_added_radius, int n_groups,
                            const float* queries4, const int* group_of, int n_queries, int n_threads,
                            float* cost, int* success, int* pops, int* pops_oob, int* n_path, unsigned long long* hash,
                            double* busy_s)
{
    std::atomic<int> next(0);
    auto worker = [&]()
    {
        std::unique_ptr<HybridAStar<float>> pl = make_planner(*p);
        int cur_group = -1;
        int N = p->grid_size;
        for (;;)
        {
            int q = next.fetch_add(1);
            if (q >= n_queries) break;
            int g = group_of[q];
            if (g != cur_group)
            {
                const float* fr = frames6 + 6 * (size_t)g;
                pl->update_goal(Vector3D<float>(fr[0], fr[1], fr[2]), Vector3D<float>(fr[3], fr[4], fr[5]));
                // rebuild the APF list in this frame, then overwrite the map with the group's final map
                pl->update_obstacles(to_boxes(boxes + 4 * (size_t)g * n_boxes, n_boxes),
                                     std::vector<float>(conf + (size_t)g * n_boxes, conf + (size_t)(g + 1) * n_boxes),
                                     apf_added_radius);
                auto& m = pl->_grid._obstacle_map;
                for (int i = 0; i < N; i++)
                    std::memcpy(m[i].data(), maps + ((size_t)g * N + i) * N, sizeof(float) * N);
                cur_group = g;
            }
            // scrub (F12)
            for (auto& row : pl->_grid._node_map)
                for (auto& n : row) { n._cost_g = 0.0f; n._cost_f = n._cost_h; n._prev = nullptr; }
            pl->reset();
            std::vector<Vector3D<float>> path;
            std::vector<float> c;
            g_trace.buf = nullptr; g_trace.cap = 0; g_trace.n = 0; g_trace.n_oob = 0;
            g_trace.bins = p->num_angle_bins; g_trace.active = true;
            const float* qq = queries4 + 4 * (size_t)q;
            auto q0 = std::chrono::steady_clock::now();
            auto r = pl->find_path(qq[3], Vector3D<float>(qq[0], qq[1], qq[2]), path, c);
            auto q1 = std::chrono::steady_clock::now();
            g_trace.active = false;
            if (cost) cost[q] = r.first;
            if (success) success[q] = r.second ? 1 : 0;
            if (pops) pops[q] = (int)g_trace.n;
            if (pops_oob) pops_oob[q] = (int)g_trace.n_oob;
            if (n_path) n_path[q] = (int)path.size();
            if (hash) hash[q] = path_hash(path, c);
            if (busy_s) busy_s[q] = std::chrono::duration<double>(q1 - q0).count();
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

double ref_bench_queries(const orc_params* p, const float* frames6, const float* maps, const float* boxes,
                         const float* conf, int n_boxes, float apf_added_radius, int n_groups,
                         const float* queries4, const int* group_of, int n_queries, int n_threads,
                         float* cost, int* success, int* pops)
{
    return ref_bench_queries_ex(p, frames6, maps, boxes, conf, n_boxes, apf_added_radius, n_groups, queries4, group_of, n_queries,
                                n_threads, cost, success, pops, nullptr, nullptr, nullptr, nullptr);
}

} // extern "C"
