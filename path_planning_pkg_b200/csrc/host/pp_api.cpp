// Host-side handle classes with the reference's C++ API (include/path_planning_pkg/*.h of this repo) on top of the
// C ABI of libpp_b200.so.  Built into lib/libpath_planning_b200.so -- the drop-in for the reference's static library
// `local_planner_lib` (reference CMakeLists.txt:134-146).  Each class owns / shares a pp_context; arguments are
// converted to the float arrays the ABI takes (T = double computes in FP32 on the device) and every computation
// runs in the CUDA kernels.  A CUDA failure is fatal here, as the reference API has no error channel
// (reference HybridAStar.cpp reports only {max, false} for "no path").
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>

#include "pp_b200.h"
#include "HybridAStar.h"
#include "VelocityGenerator.h"
#include "PedestrianHandler.h"

namespace planning
{
namespace detail
{
    static void check(int rc, const char* what)
    {
        if (rc != 0)
        {
            std::fprintf(stderr, "path_planning_b200: %s failed (%d): %s\n", what, rc, pp_last_error());
            std::abort();
        }
    }

    struct Backend
    {
        pp_context* ctx = nullptr;
        pp_params params;
        bool map_dirty = true;
        ~Backend() { if (ctx) pp_destroy(ctx); }
    };

    template <typename T>
    static pp_params make_params(int shot_interval, int shot_decay, T res, T thr, T pmin, T pmax, T pfree, int grid_size, bool diag,
                                 T step, T max_lat_acc, T max_long_dec, T wheelbase, T rear_to_cg, T apf_k, T apf_angle, int bins,
                                 int num_actions, const std::vector<T>& steering, const std::vector<T>& weights)
    {
        pp_params p;
        std::memset(&p, 0, sizeof(p));
        p.shot_interval = shot_interval; p.shot_decay = shot_decay; p.resolution = (float)res; p.obstacle_threshold = (float)thr;
        p.prob_min = (float)pmin; p.prob_max = (float)pmax; p.prob_free = (float)pfree; p.grid_size = grid_size;
        p.allow_diag = diag ? 1 : 0; p.step_size = (float)step; p.max_lat_acc = (float)max_lat_acc; p.max_long_dec = (float)max_long_dec;
        p.wheelbase = (float)wheelbase; p.rear_to_cg = (float)rear_to_cg; p.apf_rep_constant = (float)apf_k;
        p.apf_active_angle = (float)apf_angle; p.num_angle_bins = bins; p.num_actions = num_actions;
        p.num_steering = (int)steering.size();
        if (p.num_steering > PP_API_MAX_STEER) { std::fprintf(stderr, "path_planning_b200: more than %d steering primitives\n", PP_API_MAX_STEER); std::abort(); }
        for (int i = 0; i < p.num_steering; i++)
        {
            p.steering[i] = (float)steering[i];
            p.curvature_weights[i] = i < (int)weights.size() ? (float)weights[i] : 0.0f;
        }
        return p;
    }

    static std::shared_ptr<Backend> make_backend(const pp_params& p)
    {
        auto b = std::make_shared<Backend>();
        b->params = p;
        check(pp_create(&p, 0, 1, &b->ctx), "pp_create");
        return b;
    }

    // a context whose grid / map parameters are irrelevant (stand-alone Dubins / VehicleModel handles)
    static pp_params neutral_params()
    {
        std::vector<float> st{-0.5f, 0.0f, 0.5f}, w{0.0f, 0.0f, 0.0f};
        return make_params<float>(100, 10, 0.5f, 0.7f, 0.05f, 0.95f, 0.45f, 8, true, 0.5f, 2.0f, 2.0f, 2.0f, 1.0f, 1.0f, 1.0f, 72, 1, st, w);
    }

    template <typename T> static std::vector<float> boxes_of(const std::vector<Obstacle<T>>& o)
    {
        std::vector<float> b(o.size() * 4);
        for (size_t k = 0; k < o.size(); k++)
        {
            b[4 * k] = (float)o[k]._pose2D._x; b[4 * k + 1] = (float)o[k]._pose2D._y;
            b[4 * k + 2] = (float)o[k]._dimensions._x; b[4 * k + 3] = (float)o[k]._dimensions._y;
        }
        return b;
    }
    template <typename T> static std::vector<float> lines_of(const std::vector<std::pair<Vector2D<T>, Vector2D<T>>>& l)
    {
        std::vector<float> b(l.size() * 4);
        for (size_t k = 0; k < l.size(); k++)
        {
            b[4 * k] = (float)l[k].first._x; b[4 * k + 1] = (float)l[k].first._y;
            b[4 * k + 2] = (float)l[k].second._x; b[4 * k + 3] = (float)l[k].second._y;
        }
        return b;
    }
    template <typename T> static std::vector<float> floats_of(const std::vector<T>& v) { return std::vector<float>(v.begin(), v.end()); }

    template <typename T> static pp_state state_of(const Node3D<T>& n)
    {
        pp_state s;
        s.x = (float)n._pose2D._x; s.y = (float)n._pose2D._y; s.heading = (float)n._pose2D._heading;
        s.g = (float)n._cost_g; s.f = (float)n._cost_f; s.vmin_sqr = (float)n._vmin_sqr;
        s.curvature_index = n._curvature_index; s.angle_bin = n._angle_bin;
        s.ci = n._base_node ? n._base_node->_posd._x : -1; s.cj = n._base_node ? n._base_node->_posd._y : -1;
        return s;
    }
}   // namespace detail

using detail::Backend;
using detail::check;

// ===================================================== HybridAStar =====================================================
template <typename T> struct HybridAStar<T>::Impl
{
    std::shared_ptr<Backend> b;
    mutable std::vector<std::vector<T>> map_mirror;
    int last_expansions = 0;
};

template <typename T>
HybridAStar<T>::HybridAStar(int dubins_shot_interval, int dubins_shot_interval_decay, T grid_resolution, T obstacle_threshold,
                            T obstacle_prob_min, T obstacle_prob_max, T obstacle_prob_free, int grid_size,
                            bool grid_2d_allow_diag_moves, T step_size, T max_lat_acc, T max_long_dec, T wheelbase, T rear_to_cg,
                            T apf_rep_constant, T apf_active_angle, int num_angle_bins, int num_actions,
                            const std::vector<T>& steering, const std::vector<T>& curvature_weights)
    : _impl(new Impl())
{
    _impl->b = detail::make_backend(detail::make_params<T>(dubins_shot_interval, dubins_shot_interval_decay, grid_resolution,
        obstacle_threshold, obstacle_prob_min, obstacle_prob_max, obstacle_prob_free, grid_size, grid_2d_allow_diag_moves, step_size,
        max_lat_acc, max_long_dec, wheelbase, rear_to_cg, apf_rep_constant, apf_active_angle, num_angle_bins, num_actions, steering,
        curvature_weights));
    _impl->map_mirror.assign(grid_size, std::vector<T>(grid_size, T(0)));
    // One reference planner object carries its 2D heuristic cache from one find_path to the next; this
    // object does the same, so a whole session of update / find_path / reset calls returns what the reference returns.
    // PP_B200_HISTORY=0 gives every find_path a fresh cache instead.
    const char* e = std::getenv("PP_B200_HISTORY");
    if (!(e && std::strcmp(e, "0") == 0)) check(pp_set_history(_impl->b->ctx, 0, 1), "pp_set_history");
}

template <typename T> HybridAStar<T>::~HybridAStar() {}

template <typename T>
void HybridAStar<T>::update_obstacles(const std::vector<Obstacle<T>>& obstacles, const std::vector<T>& confidence, const T apf_added_radius)
{
    auto b = detail::boxes_of(obstacles); auto c = detail::floats_of(confidence);
    check(pp_update_obstacles_boxes(_impl->b->ctx, 0, b.data(), c.data(), (int)obstacles.size(), (float)apf_added_radius), "update_obstacles(boxes)");
    _impl->b->map_dirty = true;
}

template <typename T>
void HybridAStar<T>::update_obstacles(const std::vector<std::pair<Vector2D<T>, Vector2D<T>>>& lines, const std::vector<T>& confidence,
                                      const T line_width)
{
    auto l = detail::lines_of(lines); auto c = detail::floats_of(confidence);
    check(pp_update_obstacles_lines(_impl->b->ctx, 0, l.data(), c.data(), (int)lines.size(), (float)line_width), "update_obstacles(lines)");
    _impl->b->map_dirty = true;
}

template <typename T> void HybridAStar<T>::update_obstacles()
{
    check(pp_update_obstacles_decay(_impl->b->ctx, 0), "update_obstacles()");
    _impl->b->map_dirty = true;
}

template <typename T> void HybridAStar<T>::reset() { check(pp_reset(_impl->b->ctx, 0), "reset"); }

template <typename T> void HybridAStar<T>::update_goal(const Vector3D<T>& goal, const Vector3D<T>& start)
{
    float g[3] = {(float)goal._x, (float)goal._y, (float)goal._heading}, s[3] = {(float)start._x, (float)start._y, (float)start._heading};
    check(pp_update_goal(_impl->b->ctx, 0, g, s), "update_goal");
    _impl->b->map_dirty = true;
}

template <typename T> const std::vector<std::vector<T>>& HybridAStar<T>::get_obstacles() const
{
    if (_impl->b->map_dirty)
    {
        int N = _impl->b->params.grid_size;
        std::vector<float> m((size_t)N * N);
        check(pp_map_download(_impl->b->ctx, 0, m.data()), "get_obstacles");
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) _impl->map_mirror[i][j] = (T)m[(size_t)i * N + j];
        _impl->b->map_dirty = false;
    }
    return _impl->map_mirror;
}

template <typename T>
std::pair<T, bool> HybridAStar<T>::find_path(const T vel_init, const Vector3D<T>& start, std::vector<Vector3D<T>>& path,
                                             std::vector<T>& curvature)
{
    pp_query q;
    q.x = (float)start._x; q.y = (float)start._y; q.heading = (float)start._heading; q.vel = (float)vel_init; q.group = 0;
    pp_search_opts o;
    std::memset(&o, 0, sizeof(o));
    o.path_cap = 8192; o.max_slots = 1;
    // The class API returns what the reference returns (EXACT mode).  A deployment that prefers latency over
    // reference-identical plans can opt into the K-POP mode without touching the caller: PP_B200_SEARCH_MODE=kpop
    // (own semantics, ~20x lower single-query latency).
    static const int env_mode = [] {
        const char* e = std::getenv("PP_B200_SEARCH_MODE");
        return (e && (std::strcmp(e, "kpop") == 0 || std::strcmp(e, "KPOP") == 0)) ? PP_MODE_KPOP : PP_MODE_EXACT;
    }();
    o.mode = env_mode; o.kpop = 32;
    pp_result r;
    std::vector<float> xyh, curv;
    for (;;)
    {
        xyh.resize((size_t)o.path_cap * 3); curv.resize(o.path_cap);
        check(pp_find_path_batch(_impl->b->ctx, &q, 1, &o, &r, xyh.data(), curv.data(), nullptr), "find_path");
        // the reference's path vector is unbounded: a truncated path is not an answer, ask again with room for it
        if ((r.status & PP_STATUS_PATH_OVERFLOW) && o.path_cap < (1 << 20)) { o.path_cap *= 8; continue; }
        break;
    }
    _impl->last_expansions = r.n_pops;
    const int capacity_bits = PP_STATUS_OPEN_OVERFLOW | PP_STATUS_CLOSED_OVERFLOW | PP_STATUS_OPEN2D_OVERFLOW | PP_STATUS_ARENA_EXHAUSTED |
                              PP_STATUS_PATH_OVERFLOW;
    if (r.status & capacity_bits)
    {
        // The reference has no such bound (its containers are unbounded), so this is NOT its "no path" answer: say so loudly.
        std::fprintf(stderr, "path_planning_b200: HybridAStar::find_path gave up after the automatic capacity escalations "
                             "(pp_result.status = %d, %d expansions): reported as failure, the reference would have kept searching\n",
                     r.status, r.n_pops);
        return std::pair<T, bool>(std::numeric_limits<T>::max(), false);
    }
    if (r.status & PP_STATUS_NULL_TERMINAL)
        std::fprintf(stderr, "path_planning_b200: Dubins shot accepted from the start node: the reference dereferences a null _prev here "
                             "(lib/HybridAStar.cpp:137); returning the shot alone\n");
    if (!r.success) return std::pair<T, bool>(std::numeric_limits<T>::max(), false);
    // reconstruct_path (lib/HybridAStar.cpp:208-262) on caller vectors that may not be empty: after a Dubins shot it RESIZES path to the
    // shot's length and overwrites it (and curvature[1..]); without a shot it appends
    if (r.n_dubins > 0)
    {
        const T c0 = curvature.empty() ? T(0) : curvature[0];
        path.clear(); curvature.clear();
        for (int k = 0; k < r.n_path; k++)
        {
            path.push_back(Vector3D<T>((T)xyh[3 * k], (T)xyh[3 * k + 1], (T)xyh[3 * k + 2]));
            curvature.push_back((T)curv[k]);
        }
        if (!curvature.empty()) curvature[0] = c0;
    }
    else
        for (int k = 0; k < r.n_path; k++)
        {
            path.push_back(Vector3D<T>((T)xyh[3 * k], (T)xyh[3 * k + 1], (T)xyh[3 * k + 2]));
            curvature.push_back((T)curv[k]);
        }
    return std::pair<T, bool>((T)r.cost, true);
}

template <typename T> int HybridAStar<T>::last_expansions() const { return _impl->last_expansions; }
template <typename T> void* HybridAStar<T>::native_context() const { return _impl->b->ctx; }

// ======================================================== Dubins ========================================================
template <typename T> struct Dubins<T>::Impl { std::shared_ptr<Backend> b; float step; };

template <typename T> Dubins<T>::Dubins(T r_min, T step_size) : _impl(new Impl()), _path_type(Path::RSR), _r_min(r_min)
{
    _impl->b = detail::make_backend(detail::neutral_params());
    _impl->step = (float)step_size;
    check(pp_override_dubins(_impl->b->ctx, (float)r_min, (float)step_size), "Dubins");
    _params.fill(T(0));
}
template <typename T> Dubins<T>::~Dubins() {}

template <typename T> T Dubins<T>::get_shortest_path_length(const Vector3D<T>& start, const Vector3D<T>& goal)
{
    float s[3] = {(float)start._x, (float)start._y, (float)start._heading}, g[3] = {(float)goal._x, (float)goal._y, (float)goal._heading};
    float len = 0, p4[4]; int type = 0;
    check(pp_dubins_length_batch(_impl->b->ctx, s, 1, g, &len, &type, p4), "get_shortest_path_length");
    for (int k = 0; k < 4; k++) _params[k] = (T)p4[k];
    _path_type = static_cast<Path>(type);
    return (T)len;
}

template <typename T>
T Dubins<T>::get_shortest_path_length(const Vector3D<T>& start, const Vector3D<T>& goal, Vector2D<T>& center_s_r, Vector2D<T>& center_s_l,
                                      Vector2D<T>& center_g_r, Vector2D<T>& center_g_l)
{
    // circle centres: scalar glue, Dubins.cpp:73-85
    center_s_r = {start._x + _r_min * std::sin(start._heading), start._y - _r_min * std::cos(start._heading)};
    center_s_l = {start._x - _r_min * std::sin(start._heading), start._y + _r_min * std::cos(start._heading)};
    center_g_r = {goal._x + _r_min * std::sin(goal._heading), goal._y - _r_min * std::cos(goal._heading)};
    center_g_l = {goal._x - _r_min * std::sin(goal._heading), goal._y + _r_min * std::cos(goal._heading)};
    return get_shortest_path_length(start, goal);
}

template <typename T>
std::pair<T, bool> Dubins<T>::get_shortest_path(const Vector3D<T>& start, const Vector3D<T>& goal, std::vector<Vector3D<T>>& path,
                                                std::vector<T>& path_curvature)
{
    get_shortest_path_length(start, goal);       // also records _params / _path_type like the reference
    float s[3] = {(float)start._x, (float)start._y, (float)start._heading}, g[3] = {(float)goal._x, (float)goal._y, (float)goal._heading};
    const int cap = 1 << 16;
    std::vector<float> xyh((size_t)cap * 3), curv(cap);
    int n = 0, flag = 0; float len = 0;
    check(pp_dubins_path(_impl->b->ctx, s, g, xyh.data(), curv.data(), cap, &n, &len, &flag), "get_shortest_path");
    n = std::min(n, cap);
    path.resize(n); path_curvature.resize(n);
    for (int k = 0; k < n; k++) { path[k] = Vector3D<T>((T)xyh[3 * k], (T)xyh[3 * k + 1], (T)xyh[3 * k + 2]); path_curvature[k] = (T)curv[k]; }
    return std::pair<T, bool>((T)len, flag != 0);
}

template <typename T> std::string Dubins<T>::get_path_type() const
{
    static const char* names[4] = {"RSR", "RSL", "LSR", "LSL"};
    return names[static_cast<int>(_path_type)];
}

// ===================================================== VehicleModel =====================================================
template <typename T> struct VehicleModel<T>::Impl
{
    std::shared_ptr<Backend> b;
    int S, bins, A;
    T ts, max_lat_acc;
    std::vector<float> off_xy, off_h, cost, curv;
    std::vector<T> abs_curv;
    T precision;
};

template <typename T>
VehicleModel<T>::VehicleModel(T ts, T max_lat_acc, T max_long_dec, T wheelbase, T rear_to_cg, int num_angle_bins, int num_actions,
                              const std::vector<T>& steering, const std::vector<T>& curvature_weights)
    : _impl(new Impl())
{
    pp_params p = detail::make_params<T>(100, 10, T(0.5), T(0.7), T(0.05), T(0.95), T(0.45), 8, true, ts, max_lat_acc, max_long_dec,
                                         wheelbase, rear_to_cg, T(1), T(1), num_angle_bins, num_actions, steering, curvature_weights);
    _impl->b = detail::make_backend(p);
    _impl->S = (int)steering.size(); _impl->bins = num_angle_bins; _impl->A = num_actions; _impl->ts = ts; _impl->max_lat_acc = max_lat_acc;
    _impl->off_xy.resize((size_t)_impl->S * num_angle_bins * 2); _impl->off_h.resize(_impl->S); _impl->cost.resize(_impl->S); _impl->curv.resize(_impl->S);
    check(pp_get_tables(_impl->b->ctx, _impl->off_xy.data(), _impl->off_h.data(), _impl->cost.data(), _impl->curv.data()), "VehicleModel tables");
    _impl->abs_curv.assign(_impl->curv.begin(), _impl->curv.end());
    pp_consts_info ci;
    check(pp_get_consts(_impl->b->ctx, &ci), "VehicleModel consts");
    _impl->precision = (T)ci.precision;
}
template <typename T> VehicleModel<T>::~VehicleModel() {}
template <typename T> T VehicleModel<T>::get_precision() const { return _impl->precision; }
template <typename T> int VehicleModel<T>::get_default_action_index() const { return _impl->S / 2; }
template <typename T> const std::vector<T>& VehicleModel<T>::get_abs_curvatures() const { return _impl->abs_curv; }

template <typename T> static Node3D<T> node_of(const pp_state& s, const Node3D<T>* prev)
{
    Vector3D<T> pose((T)s.x, (T)s.y, (T)s.heading);
    Node3D<T> n(pose, (T)s.g, (T)s.vmin_sqr, s.curvature_index, s.angle_bin, prev);
    n._cost_f = (T)s.f;
    return n;
}

template <typename T> bool VehicleModel<T>::get_neighbors(const Node3D<T>& node, std::vector<Node3D<T>>& neighbors) const
{
    pp_state in = detail::state_of(node);
    int stride = 2 * _impl->A + 1, n_out = 0, flag = 0;
    std::vector<pp_state> out(stride);
    check(pp_rollout_batch(_impl->b->ctx, &in, 1, out.data(), &n_out, &flag), "VehicleModel::get_neighbors");
    neighbors.clear();
    for (int k = 0; k < n_out; k++) neighbors.push_back(node_of<T>(out[k], &node));
    return flag != 0;
}

template <typename T> std::pair<bool, Node3D<T>> VehicleModel<T>::simulate_action(const Node3D<T>& node, const int action_index) const
{
    // one primitive regardless of the previous action: roll out from a copy whose window is centred on action_index
    Node3D<T> probe = node;
    probe._curvature_index = action_index;
    std::vector<Node3D<T>> nb;
    get_neighbors(probe, nb);
    for (auto& n : nb)
        if (n._curvature_index == action_index) { n._prev = &node; return std::pair<bool, Node3D<T>>(true, n); }
    return std::pair<bool, Node3D<T>>(false, node);
}

// ======================================================== Grid2D ========================================================
template <typename T> Grid2D<T>::Grid2D(std::shared_ptr<Backend> backend) : _backend(backend) { init_host_tables(); }

template <typename T>
Grid2D<T>::Grid2D(T resolution, T obstacle_threshold, T obstacle_prob_min, T obstacle_prob_max, T obstacle_prob_free, int grid_size,
                  Vector2D<T> goal, Vector2D<T> start, bool allow_diag_moves)
{
    std::vector<T> st{T(-0.5), T(0), T(0.5)}, w{T(0), T(0), T(0)};
    _backend = detail::make_backend(detail::make_params<T>(100, 10, resolution, obstacle_threshold, obstacle_prob_min, obstacle_prob_max,
        obstacle_prob_free, grid_size, allow_diag_moves, T(0.5), T(2), T(2), T(2), T(1), T(1), T(1), 72, 1, st, w));
    init_host_tables();
    if (goal._x != T(0) || goal._y != T(0) || start._x != T(0) || start._y != T(0)) update_goal_heading(goal, start);
}

template <typename T>
Grid2D<T>::Grid2D(T resolution, T obstacle_threshold, T obstacle_prob_min, T obstacle_prob_max, T obstacle_prob_free, int grid_size,
                  bool allow_diag_moves)
    : Grid2D(resolution, obstacle_threshold, obstacle_prob_min, obstacle_prob_max, obstacle_prob_free, grid_size, Vector2D<T>(),
             Vector2D<T>(), allow_diag_moves) {}

template <typename T> Grid2D<T>::~Grid2D() {}

template <typename T> void Grid2D<T>::init_host_tables()
{
    const pp_params& p = _backend->params;
    const int N = p.grid_size;
    pp_consts_info ci;
    check(pp_get_consts(_backend->ctx, &ci), "Grid2D consts");
    _node_map.assign(N, std::vector<Node2D<T>>(N, Node2D<T>(0, 0)));
    const T res = (T)p.resolution;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
        {
            _node_map[i][j] = Node2D<T>(i, j);
            T dx = (ci.n45 - i) * res, dy = (ci.n2 - j) * res;
            _node_map[i][j].set_heuristic_cost(std::sqrt(dx * dx + dy * dy));
        }
    _obstacle_map.assign(N, std::vector<T>(N, T(0)));
    if (p.allow_diag) _actions = {{0, -1}, {1, -1}, {1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}};
    else _actions = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}};
    _actions_cost.clear();
    for (auto& a : _actions) _actions_cost.push_back(res * std::sqrt(static_cast<T>(a.first * a.first + a.second * a.second)));
}

template <typename T> void Grid2D<T>::refresh_mirror() const
{
    if (!_backend->map_dirty) return;
    const int N = _backend->params.grid_size;
    std::vector<float> m((size_t)N * N);
    check(pp_map_download(_backend->ctx, 0, m.data()), "map download");
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) _obstacle_map[i][j] = (T)m[(size_t)i * N + j];
    _backend->map_dirty = false;
}

template <typename T> void Grid2D<T>::get_neighbors(const int xd, const int yd, std::vector<std::pair<Node2D<T>*, T>>& neighbors)
{
    refresh_mirror();
    pp_consts_info ci;
    check(pp_get_consts(_backend->ctx, &ci), "Grid2D consts");
    const int N = _backend->params.grid_size;
    neighbors.clear();
    for (size_t k = 0; k < _actions.size(); k++)
    {
        int i = xd + _actions[k].first, j = yd + _actions[k].second;
        if (i > -1 && i < N && j > -1 && j < N && _obstacle_map[i][j] < (T)ci.log_threshold)
            neighbors.emplace_back(&_node_map[i][j], _actions_cost[k]);
    }
}

template <typename T> void Grid2D<T>::update_obstacles(const std::vector<Obstacle<T>>& obstacles, const std::vector<T>& confidence)
{
    auto b = detail::boxes_of(obstacles); auto c = detail::floats_of(confidence);
    check(pp_update_obstacles_boxes_2d(_backend->ctx, 0, b.data(), c.data(), (int)obstacles.size()), "Grid2D::update_obstacles(boxes)");
    _backend->map_dirty = true;
}
template <typename T>
void Grid2D<T>::update_obstacles(const std::vector<std::pair<Vector2D<T>, Vector2D<T>>>& lines, const std::vector<T>& confidence, const T line_width)
{
    auto l = detail::lines_of(lines); auto c = detail::floats_of(confidence);
    check(pp_update_obstacles_lines(_backend->ctx, 0, l.data(), c.data(), (int)lines.size(), (float)line_width), "Grid2D::update_obstacles(lines)");
    _backend->map_dirty = true;
}
template <typename T> void Grid2D<T>::update_obstacles() { check(pp_update_obstacles_decay(_backend->ctx, 0), "Grid2D::update_obstacles()"); _backend->map_dirty = true; }
template <typename T> void Grid2D<T>::clear_obstacles() { check(pp_clear_obstacles(_backend->ctx, 0), "clear_obstacles"); _backend->map_dirty = true; }

template <typename T> void Grid2D<T>::update_costs(const T total_cost, const Node2D<T>& last_node)
{
    for (const Node2D<T>* n = &last_node; n != nullptr; n = n->_prev) _node_map[n->_posd._x][n->_posd._y]._cost_f = total_cost - n->_cost_g;
}
template <typename T> T Grid2D<T>::get_node_total_cost(const int i, const int j) const { return _node_map[i][j]._cost_f; }
template <typename T> T Grid2D<T>::get_grid_heading() const
{
    pp_frame_info f; check(pp_get_frame(_backend->ctx, 0, &f), "frame"); return (T)f.grid_heading;
}
template <typename T> T Grid2D<T>::get_grid_resolution() const { return (T)_backend->params.resolution; }
template <typename T> int Grid2D<T>::get_grid_size() const { return _backend->params.grid_size; }
template <typename T> const std::vector<std::vector<T>>& Grid2D<T>::get_obstacle_map() const { refresh_mirror(); return _obstacle_map; }

template <typename T> Node2D<T> Grid2D<T>::update_goal_heading(const Vector2D<T>& goal, const Vector2D<T>& start)
{
    // Grid2D.cpp:260-266: goal + heading only; relocating the map is Grid3D's job
    float g[3] = {(float)goal._x, (float)goal._y, 0.0f}, s[3] = {(float)start._x, (float)start._y, 0.0f};
    check(pp_update_goal_frame(_backend->ctx, 0, g, s), "update_goal_heading");
    pp_frame_info f; check(pp_get_frame(_backend->ctx, 0, &f), "frame");
    return _node_map[f.goal_ci][f.goal_cj];
}

template <typename T> Node2D<T> Grid2D<T>::set_start_node(const Vector2D<T>& start)
{
    // Grid2D.cpp:270-291: int(rel / res) + offset (Grid3D's variant adds the offset BEFORE dividing, which truncates differently
    // for negative coordinates); an invalid start becomes node (0, 0)
    pp_frame_info f; check(pp_get_frame(_backend->ctx, 0, &f), "frame");
    pp_consts_info k; check(pp_get_consts(_backend->ctx, &k), "consts");
    const T dx = start._x - (T)f.goal_world[0], dy = start._y - (T)f.goal_world[1];
    const T h = (T)f.grid_heading, ch = std::cos(h), sh = std::sin(h);
    const T rx = dx * ch + dy * sh, ry = -dx * sh + dy * ch;
    const T res = (T)_backend->params.resolution;
    const int N = _backend->params.grid_size;
    int i = static_cast<int>(rx / res) + k.n45, j = static_cast<int>(ry / res) + k.n2;
    if (!((i > -1) && (i < N) && (j > -1) && (j < N))) { i = 0; j = 0; }
    return set_start_node_grid(i, j);
}
template <typename T> Node2D<T> Grid2D<T>::set_start_node_grid(const int i, const int j)
{
    _node_map[i][j].soft_reset();
    return _node_map[i][j];
}

// ======================================================== Grid3D ========================================================
template <typename T>
Grid3D<T>::Grid3D(T resolution, T obstacle_threshold, T obstacle_prob_min, T obstacle_prob_max, T obstacle_prob_free, int grid_size,
                  bool allow_diag_moves, T step_size, T max_lat_acc, T max_long_dec, T wheelbase, T rear_to_cg, T apf_rep_constant,
                  T apf_active_angle, int num_angle_bins, int num_actions, const std::vector<T>& steering,
                  const std::vector<T>& curvature_weights)
    : Grid2D<T>(detail::make_backend(detail::make_params<T>(100, 10, resolution, obstacle_threshold, obstacle_prob_min, obstacle_prob_max,
          obstacle_prob_free, grid_size, allow_diag_moves, step_size, max_lat_acc, max_long_dec, wheelbase, rear_to_cg, apf_rep_constant,
          apf_active_angle, num_angle_bins, num_actions, steering, curvature_weights)))
{
    int S = (int)steering.size();
    std::vector<float> oxy((size_t)S * num_angle_bins * 2), oh(S), ac(S), cu(S);
    check(pp_get_tables(this->_backend->ctx, oxy.data(), oh.data(), ac.data(), cu.data()), "Grid3D tables");
    _abs_curvatures.assign(cu.begin(), cu.end());
}

template <typename T>
void Grid3D<T>::update_obstacles(const std::vector<Obstacle<T>>& obstacles, const std::vector<T>& confidence, const T apf_added_radius)
{
    auto b = detail::boxes_of(obstacles); auto c = detail::floats_of(confidence);
    check(pp_update_obstacles_boxes(this->_backend->ctx, 0, b.data(), c.data(), (int)obstacles.size(), (float)apf_added_radius), "Grid3D::update_obstacles");
    this->_backend->map_dirty = true;
}

template <typename T> bool Grid3D<T>::get_neighbors(const Node3D<T>& node, std::vector<Node3D<T>>& neighbors) const
{
    pp_state in = detail::state_of(node);
    int stride = 2 * this->_backend->params.num_actions + 1, n_out = 0, flag = 0;
    std::vector<pp_state> out(stride);
    check(pp_expand_batch(this->_backend->ctx, 0, &in, 1, out.data(), &n_out, &flag), "Grid3D::get_neighbors");
    neighbors.clear();
    for (int k = 0; k < n_out; k++)
    {
        Node3D<T> n = node_of<T>(out[k], &node);
        n._base_node = &(this->_node_map[out[k].ci][out[k].cj]);
        neighbors.push_back(n);
    }
    return flag != 0;
}

template <typename T> bool Grid3D<T>::check_path(const std::vector<Vector3D<T>>& path) const
{
    std::vector<float> xyh(path.size() * 3);
    for (size_t k = 0; k < path.size(); k++) { xyh[3 * k] = (float)path[k]._x; xyh[3 * k + 1] = (float)path[k]._y; xyh[3 * k + 2] = (float)path[k]._heading; }
    int free_flag = 1;
    check(pp_check_path(this->_backend->ctx, 0, xyh.data(), (int)path.size(), &free_flag), "check_path");
    return free_flag != 0;
}

template <typename T> Vector3D<T> Grid3D<T>::get_goal_location() const
{
    pp_frame_info f; check(pp_get_frame(this->_backend->ctx, 0, &f), "frame");
    return Vector3D<T>((T)f.goal_world[0], (T)f.goal_world[1], (T)f.goal_world[2]);
}

template <typename T> Node3D<T> Grid3D<T>::update_goal_heading(const Vector3D<T>& goal, const Vector3D<T>& start)
{
    float g[3] = {(float)goal._x, (float)goal._y, (float)goal._heading}, s[3] = {(float)start._x, (float)start._y, (float)start._heading};
    check(pp_update_goal(this->_backend->ctx, 0, g, s), "Grid3D::update_goal_heading");
    this->_backend->map_dirty = true;
    pp_frame_info f; check(pp_get_frame(this->_backend->ctx, 0, &f), "frame");
    Vector3D<T> pose((T)f.goal_grid[0], (T)f.goal_grid[1], (T)f.goal_grid[2]);
    return Node3D<T>(pose, T(0), T(0), 0, f.goal_bin, &(this->_node_map[f.goal_ci][f.goal_cj]), nullptr);
}

template <typename T> Node3D<T> Grid3D<T>::set_start_node(const Vector3D<T>& start)
{
    pp_query q; q.x = (float)start._x; q.y = (float)start._y; q.heading = (float)start._heading; q.vel = 0.0f; q.group = 0;
    pp_state s; check(pp_set_start_batch(this->_backend->ctx, &q, 1, &s), "Grid3D::set_start_node");
    this->_node_map[s.ci][s.cj].soft_reset();
    Vector3D<T> pose((T)s.x, (T)s.y, (T)s.heading);
    return Node3D<T>(pose, T(0), T(0), s.curvature_index, s.angle_bin, &(this->_node_map[s.ci][s.cj]), nullptr);
}

template <typename T> const std::vector<T>& Grid3D<T>::get_abs_curvatures() const { return _abs_curvatures; }

// ======================================================== AStar ========================================================
#ifndef STORE_GRID_AS_REFERENCE
template <typename T>
AStar<T>::AStar(T grid_resolution, T obstacle_threshold, T obstacle_prob_min, T obstacle_prob_max, T obstacle_prob_free, int grid_size,
                bool grid_allow_diag_moves)
    : _owned(new Grid2D<T>(grid_resolution, obstacle_threshold, obstacle_prob_min, obstacle_prob_max, obstacle_prob_free, grid_size,
                           grid_allow_diag_moves)), _grid(_owned.get()), _fresh(true) {}
#else
template <typename T> AStar<T>::AStar(Grid2D<T>& grid) : _grid(&grid), _fresh(true) {}
#endif
template <typename T> AStar<T>::~AStar() {}
template <typename T> void AStar<T>::update_goal_node(const Node2D<T>&) {}       // the goal cell is fixed at (0.8 N, 0.5 N)
template <typename T> void AStar<T>::update_goal_start(const Vector2D<T>& goal, const Vector2D<T>& start, Node2D<T>& start_node)
{
    _grid->update_goal_heading(goal, start);
    start_node = _grid->set_start_node(start);
}
template <typename T> void AStar<T>::update_obstacles(const std::vector<Obstacle<T>>& o, const std::vector<T>& c) { _grid->update_obstacles(o, c); }
template <typename T>
void AStar<T>::update_obstacles(const std::vector<std::pair<Vector2D<T>, Vector2D<T>>>& l, const std::vector<T>& c, const T w) { _grid->update_obstacles(l, c, w); }
template <typename T> void AStar<T>::update_obstacles() { _grid->update_obstacles(); }
// AStar.cpp:56-60: only the visited flags are cleared; the node costs earlier searches left stay
template <typename T> void AStar<T>::reset() { if (!_fresh) check(pp_astar_lazy_reset(_grid->backend()->ctx, 0), "AStar::reset"); }
template <typename T> const std::vector<std::vector<T>>& AStar<T>::get_obstacles() const { return _grid->get_obstacle_map(); }

template <typename T> T AStar<T>::find_path(const int start_i, const int start_j)
{
    int ij[2] = {start_i, start_j};
    float out = 0;
    pp_context* ctx = _grid->backend()->ctx;
    check(_fresh ? pp_astar_lazy_batch(ctx, 0, ij, 1, &out) : pp_astar_lazy_continue(ctx, 0, ij, 1, &out), "AStar::find_path");
    _fresh = false;
    return out >= std::numeric_limits<float>::max() ? std::numeric_limits<T>::max() : (T)out;
}

template <typename T> T AStar<T>::find_path(const Vector2D<T>& goal, const Vector2D<T>& start, bool)
{
    _grid->update_goal_heading(goal, start);
    Node2D<T> s = _grid->set_start_node(start);
    return find_path(s._posd._x, s._posd._y);
}

// ================================================== VelocityGenerator ==================================================
template <typename T>
VelocityGenerator<T>::VelocityGenerator(T max_velocity, T coast_velocity, T max_lat_acc, T max_long_acc, T max_long_dec)
    : _max_velocity(max_velocity), _coast_velocity(coast_velocity), _max_lat_acc(max_lat_acc), _max_lat_acc_sqr(max_lat_acc * max_lat_acc),
      _max_long_acc(max_long_acc), _max_long_dec(max_long_dec) {}

// Three passes over v^2 along the path walked start -> goal (the vectors are stored goal -> start): curvature / speed
// limit pass with the braking budget, forward pass with the acceleration budget, backward pass with the braking
// budget (reference lib/VelocityGenerator.cpp:19-85).
template <typename T>
bool VelocityGenerator<T>::generate_velocity_profile(const T vel_init, const T max_velocity_curr, const std::vector<Vector3D<T>>& path,
                                                     const std::vector<T>& curvature, std::vector<T>& velocity, bool coast_to_goal,
                                                     bool stop_at_goal) const
{
    const std::size_t n = path.size();
    T vmax = std::min(coast_to_goal ? _coast_velocity : _max_velocity, max_velocity_curr);
    const T vmax_sqr = vmax * vmax;
    velocity.resize(n);
    std::vector<T> v2(n);
    auto seg = [&](std::size_t a, std::size_t b) { return std::hypot(path[a]._x - path[b]._x, path[a]._y - path[b]._y); };
    auto budget = [&](T acc, T v2_here, T kappa) -> T
    {
        T lat = v2_here * kappa;
        return acc * std::sqrt(1.0 - (lat * lat) / _max_lat_acc_sqr);
    };
    v2[0] = vel_init * vel_init;
    T cap = v2[0];
    for (std::size_t i = 0; i + 1 < n; i++)
    {
        std::size_t p = n - i - 1;
        T ds = seg(p - 1, p);
        cap = std::max(cap - 2 * budget(_max_long_dec, v2[i], curvature[p]) * ds, vmax_sqr);
        v2[i + 1] = (curvature[p - 1] != 0) ? std::min(_max_lat_acc / curvature[p - 1], cap) : cap;
    }
    if (stop_at_goal) v2[n - 1] = 0;
    for (std::size_t i = 0; i + 1 < n; i++)
    {
        std::size_t p = n - i - 1;
        v2[i + 1] = std::min(v2[i] + 2 * budget(_max_long_acc, v2[i], curvature[p]) * seg(p - 1, p), v2[i + 1]);
    }
    for (std::size_t i = n - 1; i > 0; i--)
    {
        std::size_t p = n - i - 1;
        v2[i - 1] = std::min(v2[i] + 2 * budget(_max_long_dec, v2[i], curvature[p]) * seg(p + 1, p), v2[i - 1]);
        velocity[i - 1] = std::sqrt(v2[i - 1]);
    }
    velocity[n - 1] = std::sqrt(v2[n - 1]);
    return vel_init < (velocity[0] + static_cast<T>(0.25));
}

// ================================================== PedestrianHandler ==================================================
template <typename T>
PedestrianHandler<T>::PedestrianHandler(T detection_arc_angle, T min_stop_dist, T min_allowable_ttc, T max_long_dec, T min_vel)
    : _half_arc(detection_arc_angle / 2), _min_stop_dist(min_stop_dist), _min_allowable_ttc(min_allowable_ttc),
      _max_long_dec(max_long_dec), _min_vel(min_vel) {}

template <typename T>
void PedestrianHandler<T>::bearing_and_range(const Vector3D<T>& pose, const Obstacle<T>& ped, T& rel_angle, T& long_dist) const
{
    const T dx = ped._pose2D._x - pose._x, dy = ped._pose2D._y - pose._y;
    rel_angle = wrap_pi(std::atan2(dy, dx) - pose._heading);
    long_dist = std::hypot(dx, dy) * std::cos(rel_angle);
}

template <typename T>
T PedestrianHandler<T>::time_to_collision(const T vel, const Vector3D<T>& pose, const Obstacle<T>& ped) const
{
    T rel_angle, long_dist;
    bearing_and_range(pose, ped, rel_angle, long_dist);
    const T inf = std::numeric_limits<T>::max();
    if (std::abs(rel_angle) > _half_arc) return inf;                       // outside the detection arc
    if ((2 * _max_long_dec * long_dist) > (vel * vel)) return inf;         // can stop before reaching
    return (-vel + std::sqrt(-2 * _max_long_dec * long_dist + vel * vel)) / (-_max_long_dec);
}

// smallest time-to-collision over the pedestrians decides: stop, cap the speed, or no cap (reference
// lib/PedestrianHandler.cpp:17-56)
template <typename T>
T PedestrianHandler<T>::calc_max_velocity(const T vel_curr, const Vector3D<T>& pose_curr, const std::vector<Obstacle<T>>& pedestrians) const
{
    const T inf = std::numeric_limits<T>::max();
    T best = inf;
    const Obstacle<T>* closest = nullptr;
    for (const auto& p : pedestrians)
    {
        T ttc = time_to_collision(vel_curr, pose_curr, p);
        if (ttc < best) { best = ttc; closest = &p; }
    }
    if (closest != nullptr)
    {
        T rel_angle, long_dist;
        bearing_and_range(pose_curr, *closest, rel_angle, long_dist);
        if (long_dist < _min_stop_dist) return 0;
        if (best < _min_allowable_ttc)
            return std::max((2 * long_dist - _max_long_dec * _min_allowable_ttc * _min_allowable_ttc) / (2 * _min_allowable_ttc), _min_vel);
    }
    return inf;
}

// explicit instantiations: the same two scalar types as the reference library
#define PP_INSTANTIATE(T)                     \
    template class HybridAStar<T>;            \
    template class Dubins<T>;                 \
    template class VehicleModel<T>;           \
    template class Grid2D<T>;                 \
    template class Grid3D<T>;                 \
    template class AStar<T>;                  \
    template class VelocityGenerator<T>;      \
    template class PedestrianHandler<T>;
PP_INSTANTIATE(float)
PP_INSTANTIATE(double)

}   // namespace planning
