// TEST INFRASTRUCTURE ONLY -- not part of the product, never loaded by path_planning_pkg_b200.
//
// Compiles the PRODUCT's PP_HD core (csrc/core/*.h: the code the CUDA kernels execute) with g++ for a
// single "lane" (PPWarpSerial) and exports it behind the oracle ABI with the prefix `emu_`, so the
// control logic of the kernels (libstdc++-exact red-black tree, lazy 2D A*, search loop, Dubins,
// APF, rasteriser arithmetic) can be checked against the unmodified reference on a machine without a
// GPU (`pytest -m "not gpu"`).  The GPU tests check the real kernels through libpp_b200.so.
#include <vector>
#include <cstring>
#include <cstdio>
#include <memory>
#include <cstdlib>

#include "../../path_planning_pkg_b200/csrc/core/pp_search.h"
#include "../../path_planning_pkg_b200/csrc/core/pp_kpop.h"
#include "../../path_planning_pkg_b200/csrc/core/pp_map.h"
#include "../../path_planning_pkg_b200/csrc/core/pp_velocity.h"
#include "../../path_planning_pkg_b200/csrc/host/pp_host.h"
#include "../../path_planning_pkg_b200/csrc/host/pp_footprint_host.h"
#include "../../oracle/oracle_api.h"

namespace
{
    struct Emu
    {
        pp_params p;
        PPHostModel m;
        PPHostFrame fr;
        std::vector<float> map;     // N*N
        std::vector<float> apf;     // K*3
        // search scratch
        std::vector<PPNode3> open3;
        std::vector<PPClosed3> closed;
        std::vector<PPHashSlot> chash;
        std::vector<unsigned> cell_state;
        std::vector<float> nm_g, nm_f, cl_g;
        std::vector<int> cl_prev;
        std::vector<PPNode2> open2;
        std::vector<PPPathPt> path;
        std::vector<int> bin_off, bin_idx; int bin_n = 0;
        PPLazy lazy;               // persistent lazy-A* state for emu_astar_lazy_batch
        bool lazy_init = false;
        PPResult last; float last_vel = 0.0f; bool have_last = false;   // the last emu_find_path, for emu_trajectory
        bool hist_on = false;      // planner-object history (pp_set_history): cell_state / nm_g / nm_f carried between queries
        unsigned hist_sid = 0;
        // growable containers (pp_arena.h): host memory stands in for the context's device arena
        std::vector<unsigned long long> arena_mem; PPArena arena;
    };

    // emu_find_path runs on tiny fixed pools + the arena (so every CPU search test exercises container growth) unless
    // PP_EMU_FIXED=1 asks for the round-1 fixed pools
    void attach_arena(Emu* e, PPWork& wk, int closed_max, int open_max, int open2_max)
    {
        if (e->arena_mem.empty()) e->arena_mem.resize((size_t)(768u << 20) / 8);
        std::memset(&e->arena, 0, sizeof(PPArena));
        e->arena.base = (unsigned long long)e->arena_mem.data(); e->arena.size = e->arena_mem.size() * 8ull;
        wk.arena = &e->arena; wk.closed_max = closed_max; wk.open3_max = open_max; wk.open2_max = open2_max;
    }

    PPGroup group_of(Emu* e)
    {
        PPGroup g;
        g.map = e->map.data();
        g.apf = e->apf.empty() ? nullptr : e->apf.data();
        g.K = (int)(e->apf.size() / 3);
        g.pad = 0;
        g.bin_shift = 4; g.bin_n = e->bin_n;
        g.bin_off = e->bin_off.empty() ? nullptr : e->bin_off.data();
        g.bin_idx = e->bin_idx.data();
        g.frame = e->fr.F;
        return g;
    }

    void setup_work(Emu* e, PPWork& wk, int closed_cap, int open_cap, int open2_cap, int path_cap)
    {
        int N = e->m.C.N;
        e->open3.resize(open_cap); e->closed.resize(closed_cap);
        int hc = 1; while (hc < 2 * closed_cap) hc <<= 1;
        e->chash.resize(hc);
        e->cell_state.resize((size_t)N * N); e->nm_g.resize((size_t)N * N); e->nm_f.resize((size_t)N * N);
        e->cl_g.resize((size_t)N * N); e->cl_prev.resize((size_t)N * N);
        e->open2.resize(open2_cap); e->path.resize(path_cap);
        wk.open3 = e->open3.data(); wk.open3_cap = open_cap;
        wk.closed = e->closed.data(); wk.closed_cap = closed_cap;
        wk.chash = e->chash.data(); wk.chash_cap = hc;
        wk.cell_state = e->cell_state.data(); wk.nm_g = e->nm_g.data(); wk.nm_f = e->nm_f.data();
        wk.cl_g = e->cl_g.data(); wk.cl_prev = e->cl_prev.data();
        wk.open2 = e->open2.data(); wk.open2_cap = open2_cap;
        wk.path = e->path.data(); wk.path_cap = path_cap;
        wk.trace = nullptr; wk.trace_cap = 0;
    }

    PPState to_pp(const orc_state& s)
    {
        PPState o; o.x = s.x; o.y = s.y; o.heading = s.heading; o.g = s.g; o.f = s.f; o.vmin_sqr = s.vmin_sqr;
        o.curv = s.curvature_index; o.bin = s.angle_bin; o.ci = s.ci; o.cj = s.cj; return o;
    }
}

extern "C"
{

void* emu_create(const orc_params* p)
{
    static_assert(sizeof(orc_params) == sizeof(pp_params), "param structs must match");
    Emu* e = new Emu();
    std::memcpy(&e->p, p, sizeof(pp_params));
    std::string err;
    if (!pp_host_build_model(e->p, e->m, err)) { std::fprintf(stderr, "emu_create: %s\n", err.c_str()); delete e; return nullptr; }
    e->map.assign((size_t)e->m.C.N * e->m.C.N, 0.0f);
    float z[3] = {0, 0, 0};
    pp_host_update_goal(e->m.C, z, z, e->fr);
    e->fr.grid_heading = 0.0f;   // Grid2D ctor: atan2(0, 0) = 0
    return e;
}

void emu_destroy(void* h) { delete static_cast<Emu*>(h); }

void emu_update_goal(void* h, const float* goal3, const float* start3)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    PPHostFrame prev = e->fr;
    pp_host_update_goal(C, goal3, start3, e->fr);
    // relocate_obstacles exactly as pp_update_goal does it (pp_map_reloc_{fill,scatter,gather}_kernel): forward scatter of the
    // source index, the largest source index (= last writer in raster order) wins, then gather
    PPRelocDesc d;
    pp_host_reloc_desc(C, e->fr.grid_heading, prev.grid_heading, e->fr.goal_world, prev.goal_world, d);
    const int N = C.N;
    std::vector<int> idx((size_t)N * N, -1);
    for (int s = 0; s < N * N; s++)
    {
        int in, jn;
        pp_reloc_target(d, s / N, s % N, in, jn);
        if (in > -1 && in < N && jn > -1 && jn < N && idx[(size_t)in * N + jn] < s) idx[(size_t)in * N + jn] = s;
    }
    std::vector<float> moved((size_t)N * N);
    for (int t = 0; t < N * N; t++) moved[t] = idx[t] >= 0 ? e->map[idx[t]] : 0.0f;
    e->map.swap(moved);
}

// AStar::reset() on the carried cache = pp_reset with history enabled (pp_hist_reset_kernel)
void emu_reset(void* h)
{
    Emu* e = static_cast<Emu*>(h);
    if (e->hist_on) for (auto& st : e->cell_state) st &= ~PP_CS_VISITED;
}
// = pp_set_history: the freshly constructed cache, carried from now on
void emu_set_history(void* h, int enable)
{
    Emu* e = static_cast<Emu*>(h);
    size_t nn = (size_t)e->m.C.N * e->m.C.N;
    e->hist_on = enable != 0; e->hist_sid = 0;
    e->cell_state.assign(nn, 0u); e->nm_g.assign(nn, 0.0f); e->nm_f.assign(nn, 0.0f);
}
// test hook: force the running lazy-search id (exercises the stamp wrap-around of pp_search_exact)
void emu_set_history_sid(void* h, unsigned sid) { static_cast<Emu*>(h)->hist_sid = sid; }
void emu_scrub(void* h) { static_cast<Emu*>(h)->lazy_init = false; }

void emu_update_boxes_2d(void* h, const float* boxes, const float* conf, int n)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    float ch, sh;
    std::vector<PPBoxDesc> d;
    pp_host_box_descs(C, e->fr, boxes, conf, n, ch, sh, d);
    // gather form, one "thread" per cell of each box's bounding rectangle, boxes in order
    for (int k = 0; k < n; k++)
        for (int i = d[k].lo_i; i <= d[k].hi_i; i++)
            for (int j = d[k].lo_j; j <= d[k].hi_j; j++)
            {
                int cnt = pp_box_count(d[k].ni, d[k].nj, ch, sh, i - d[k].start_i, j - d[k].start_j);
                if (cnt) e->map[(size_t)i * C.N + j] = pp_box_apply(e->map[(size_t)i * C.N + j], cnt, d[k].delta, C.log_min, C.log_max);
            }
}

void emu_update_boxes(void* h, const float* boxes, const float* conf, int n, float apf_added_radius)
{
    Emu* e = static_cast<Emu*>(h);
    pp_host_apf_list(e->m.C, e->fr, boxes, n, apf_added_radius, e->apf);
    pp_host_apf_bins(e->m.C, e->apf, n, 4, e->bin_n, e->bin_off, e->bin_idx);
    emu_update_boxes_2d(h, boxes, conf, n);
}

void emu_update_lines(void* h, const float* l, const float* conf, int n, float width)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    std::vector<PPLineDesc> d;
    pp_host_line_descs(C, e->fr, l, conf, n, d);
    for (int k = 0; k < n; k++)
    {
        float pl = 0.0f;
        for (int t = 0; t < 100 && pl <= d[k].length; t++)
        {
            for (float pw = 0.0f; pw <= width; pw += C.res)
            {
                int i1, j1, i2, j2;
                pp_line_cells(d[k], C.res, C.n45, C.n2, pl, pw, i1, j1, i2, j2);
                if (i1 > -1 && i1 < C.N && j1 > -1 && j1 < C.N)
                    e->map[(size_t)i1 * C.N + j1] = pp_box_apply(e->map[(size_t)i1 * C.N + j1], 1, d[k].delta, C.log_min, C.log_max);
                if (i2 > -1 && i2 < C.N && j2 > -1 && j2 < C.N)
                    e->map[(size_t)i2 * C.N + j2] = pp_box_apply(e->map[(size_t)i2 * C.N + j2], 1, d[k].delta, C.log_min, C.log_max);
            }
            pl += C.res;
        }
    }
}

void emu_decay(void* h)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    for (auto& v : e->map) v = pp_map_decay_cell(v, C.log_free, C.log_min, C.log_max);
}

void emu_get_map(void* h, float* out) { Emu* e = static_cast<Emu*>(h); std::memcpy(out, e->map.data(), e->map.size() * 4); }
void emu_set_map(void* h, const float* in) { Emu* e = static_cast<Emu*>(h); std::memcpy(e->map.data(), in, e->map.size() * 4); }

void emu_get_consts(void* h, orc_consts* c)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    c->log_threshold = C.log_thr; c->log_min = C.log_min; c->log_max = C.log_max; c->log_free = C.log_free;
    c->grid_heading = e->fr.grid_heading;
    for (int q = 0; q < 3; q++) c->goal_world[q] = e->fr.goal_world[q];
    c->goal_grid[0] = e->fr.F.goal_x; c->goal_grid[1] = e->fr.F.goal_y; c->goal_grid[2] = e->fr.F.goal_h;
    c->goal_bin = e->fr.F.goal_bin; c->goal_ci = e->fr.F.goal_ci; c->goal_cj = e->fr.F.goal_cj;
    c->precision = C.precision; c->r_min = C.r_min; c->ang_step = C.ang_step;
    c->num_apf = (int)(e->apf.size() / 3);
}

void emu_get_apf(void* h, float* out) { Emu* e = static_cast<Emu*>(h); std::memcpy(out, e->apf.data(), e->apf.size() * 4); }

void emu_get_tables(void* h, float* offset_xy, float* offset_heading, float* actions_cost, float* abs_curv)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    for (int i = 0; i < C.S; i++)
    {
        offset_heading[i] = C.off_heading[i]; actions_cost[i] = C.act_cost3d[i]; abs_curv[i] = C.abs_curv[i];
        for (int j = 0; j < C.bins; j++)
        {
            offset_xy[((size_t)i * C.bins + j) * 2] = e->m.off_xy[((size_t)i * (C.bins + 1) + j) * 2];
            offset_xy[((size_t)i * C.bins + j) * 2 + 1] = e->m.off_xy[((size_t)i * (C.bins + 1) + j) * 2 + 1];
        }
    }
}

void emu_set_start(void* h, const float* s, orc_state* out)
{
    Emu* e = static_cast<Emu*>(h);
    PPState st = pp_host_set_start(e->m.C, e->fr, s[0], s[1], s[2], 0.0f);
    out->x = st.x; out->y = st.y; out->heading = st.heading; out->g = 0.0f; out->f = 0.0f; out->vmin_sqr = 0.0f;
    out->curvature_index = st.curv; out->angle_bin = st.bin; out->ci = st.ci; out->cj = st.cj;
}

static void emu_succ(Emu* e, const orc_state* in, int n, orc_state* out, int* n_out, int* flags, bool expand)
{
    const PPConsts& C = e->m.C;
    PPWarpSerial w;
    int stride = 2 * C.A + 1;
    PPGroup G = group_of(e);
    for (int k = 0; k < n; k++)
    {
        PPState s = to_pp(in[k]);
        flags[k] = (s.vmin_sqr < 1.0f) ? 1 : 0;
        int start_index = s.curv - C.A; if (start_index < 0) start_index = 0;
        int cnt = 0;
        for (int q = 0; q < stride; q++)
        {
            int i = start_index + q;
            if (i >= C.S) break;
            PPSucc o;
            int bin = s.bin > C.bins ? C.bins : s.bin;
            if (!pp_rollout_one(C, e->m.off_xy.data(), s.x, s.y, s.heading, s.g, s.vmin_sqr, bin, i, o)) continue;
            o.ci = -1; o.cj = -1;
            float f = o.g;
            if (expand)
            {
                if (!pp_collision_free(C, G.map, o.x, o.y, o.ci, o.cj)) continue;
                float field = pp_apf_sum(w, C, G.apf, (const int*)0, G.K, o.x, o.y, o.heading);
                o.g = o.g + field; f = f + field;
            }
            orc_state& r = out[(size_t)k * stride + cnt];
            r.x = o.x; r.y = o.y; r.heading = o.heading; r.g = o.g; r.f = f; r.vmin_sqr = o.vmin_sqr;
            r.curvature_index = o.curv; r.angle_bin = o.bin; r.ci = o.ci; r.cj = o.cj;
            cnt++;
        }
        n_out[k] = cnt;
    }
}

void emu_rollout_batch(void* h, const orc_state* in, int n, orc_state* out, int* n_out, int* flags)
{ emu_succ(static_cast<Emu*>(h), in, n, out, n_out, flags, false); }

void emu_expand_batch(void* h, const orc_state* in, int n, orc_state* out, int* n_out, int* flags)
{ emu_succ(static_cast<Emu*>(h), in, n, out, n_out, flags, true); }

void emu_apf_batch(void* h, const float* xyh, int n, float* out)
{
    Emu* e = static_cast<Emu*>(h);
    PPWarpSerial w;
    PPGroup G = group_of(e);
    for (int k = 0; k < n; k++)
        out[k] = pp_apf_sum(w, e->m.C, G.apf, (const int*)0, G.K, xyh[3 * k], xyh[3 * k + 1], xyh[3 * k + 2]);
}

int emu_check_path(void* h, const float* xyh, int n)
{
    Emu* e = static_cast<Emu*>(h);
    for (int k = 0; k < n; k++)
        if (pp_path_point_blocked(e->m.C, e->map.data(), xyh[3 * k], xyh[3 * k + 1])) return 0;
    return 1;
}

void emu_dubins_length_batch(void* h, const float* starts, int n, const float* goal3, float* len, int* type, float* params4)
{
    Emu* e = static_cast<Emu*>(h);
    for (int k = 0; k < n; k++)
    {
        int t; float p[4]; PPDubinsCenters c;
        len[k] = pp_dubins_shortest(e->m.C.r_min, starts[3 * k], starts[3 * k + 1], starts[3 * k + 2], goal3[0], goal3[1], goal3[2], t, p, c);
        if (type) type[k] = t;
        if (params4) for (int q = 0; q < 4; q++) params4[4 * k + q] = p[q];
    }
}

int emu_dubins_path(void* h, const float* s, const float* g, float* xyh, float* curv, int cap, float* length, int* flag)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    int t; float p[4]; PPDubinsCenters c; PPDubinsPlan pl;
    *length = pp_dubins_shortest(C.r_min, s[0], s[1], s[2], g[0], g[1], g[2], t, p, c);
    *flag = (fabsf(p[1]) > (float)PP_PI_2) ? 1 : 0;
    pp_dubins_plan(C.r_min, C.step, C.ang_step, t, p, c, pl);
    int total = pl.size_3 + 1;
    float acc = p[0];
    for (int k = 0; k < total; k++)
    {
        if (k == pl.size_1) acc = 0.0f;
        if (k == pl.size_2) acc = p[2];
        float x, y, hh, kappa;
        pp_dubins_sample(pl, C.r_min, k, acc, x, y, hh, kappa);
        if (k < cap) { xyh[3 * k] = x; xyh[3 * k + 1] = y; xyh[3 * k + 2] = hh; curv[k] = kappa; }
        if (k < pl.size_1) acc = (pl.s1 < 0) ? acc - C.ang_step : acc + C.ang_step;
        else if (k < pl.size_2) acc = acc + C.step;
        else if (k < pl.size_3) acc = (pl.s2 < 0) ? acc - C.ang_step : acc + C.ang_step;
    }
    return total;
}

// velocity profile / trajectory: the PP_HD code the kernels run, on one host lane
void emu_velocity_profile_batch(const float* lim5, const float* paths_xy, const float* curv, const int* counts, int n, int cap,
                                const float* vel_init, const float* vcap, const int* flags, float* velocity, int* feasible)
{
    PPVelLimits L; L.max_velocity = lim5[0]; L.coast_velocity = lim5[1]; L.max_lat_acc = lim5[2]; L.max_long_acc = lim5[3]; L.max_long_dec = lim5[4];
    std::vector<float> v2(cap);
    for (int k = 0; k < n; k++)
    {
        int m = std::min(counts[k], cap);
        if (m < 1) { feasible[k] = 0; continue; }
        const float* xy = paths_xy + (size_t)k * cap * 2;
        int fl = flags ? flags[k] : 0;
        feasible[k] = pp_velocity_profile(L, vel_init[k], vcap ? vcap[k] : FLT_MAX, xy, xy + 1, 2, curv + (size_t)k * cap, 1, m, v2.data(),
                                          velocity + (size_t)k * cap, (fl & 1) != 0, (fl & 2) != 0) ? 1 : 0;
    }
}

// trajectory of the last emu_find_path (= pp_trajectory_batch for that query); traj has room for 4 * 4096 floats
int emu_trajectory(void* h, const float* lim5, float vcap, int stop, float* traj, int* feasible)
{
    Emu* e = static_cast<Emu*>(h);
    if (!e->have_last || !e->last.success) { *feasible = 0; return 0; }
    PPVelLimits L; L.max_velocity = lim5[0]; L.coast_velocity = lim5[1]; L.max_lat_acc = lim5[2]; L.max_long_acc = lim5[3]; L.max_long_dec = lim5[4];
    PPWorldFrame F;
    F.goal_gx = e->fr.F.goal_x; F.goal_gy = e->fr.F.goal_y; F.goal_wx = e->fr.goal_world[0]; F.goal_wy = e->fr.goal_world[1];
    F.angle = -e->fr.grid_heading; F.c = std::cos(F.angle); F.s = std::sin(F.angle);
    const int cap = (int)e->path.size();
    std::vector<float> tmp(2 * (size_t)cap);
    PPWarpSerial w;
    static_assert(sizeof(PPTrajPt) == sizeof(PPPathPt), "PPTrajPt mirrors PPPathPt");
    return pp_trajectory_assemble(w, F, L, reinterpret_cast<const PPTrajPt*>(e->path.data()), e->last.n_dubins, e->last.n_chain, cap,
                                  e->last_vel, vcap, stop != 0, traj, tmp.data(), tmp.data() + cap, feasible);
}

// generic footprint check: the product's table builder (host/pp_footprint_host.h) and the predicates the kernel executes
int emu_footprint_table(void* h, int bin, float length, float width, float rear, short* offs_ij, int cap)
{
    Emu* e = static_cast<Emu*>(h);
    std::vector<PPFootBin> bins; std::vector<PPCellOff> offs;
    pp_footprint_build(e->m.C, length, width, rear, bins, offs);
    const PPFootBin& b = bins[bin];
    for (int k = 0; k < b.count && k < cap; k++) { offs_ij[2 * k] = offs[b.first + k].di; offs_ij[2 * k + 1] = offs[b.first + k].dj; }
    return b.count;
}

void emu_footprint_check(void* h, const float* xyh, int n, float length, float width, float rear, int* free_out, int* cells_ij,
                         int* hits_out)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    std::vector<PPFootBin> bins; std::vector<PPCellOff> offs;
    pp_footprint_build(C, length, width, rear, bins, offs);
    for (int k = 0; k < n; k++)
    {
        float x = xyh[3 * k], y = xyh[3 * k + 1];
        int ci = (int)(x / C.res), cj = (int)(y / C.res);
        const PPFootBin& B = bins[pp_foot_bin(xyh[3 * k + 2], C.precision, C.bins)];
        int hits = 0;
        for (int t = 0; t < B.count; t++)
            if (pp_foot_cell_blocked(C, e->map.data(), ci + offs[B.first + t].di, cj + offs[B.first + t].dj)) hits++;
        free_out[k] = hits == 0; if (cells_ij) { cells_ij[2 * k] = ci; cells_ij[2 * k + 1] = cj; } if (hits_out) hits_out[k] = hits;
    }
}

void emu_astar_lazy_batch(void* h, const int* ij, int n, float* out)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    PPWork wk;
    if (!e->lazy_init)
    {
        setup_work(e, wk, 16, 16, 1 << 16, 16);
        std::fill(e->cell_state.begin(), e->cell_state.end(), 0u);
        e->lazy.open.init(wk.open2, wk.open2_cap);
        e->lazy.search_id = 0; e->lazy.status = 0; e->lazy.n_searches = 0; e->lazy.n_pops = 0;
        e->lazy_init = true;
    }
    else setup_work(e, wk, 16, 16, 1 << 16, 16);
    PPLazyNb nbs[8];
    for (int k = 0; k < n; k++) out[k] = pp_lazy_astar(C, e->map.data(), e->fr.F, wk, e->lazy, ij[2 * k], ij[2 * k + 1], nbs);
}

void emu_astar_dump(void* h, unsigned char* visited, float* g, float* f)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    for (int c = 0; c < C.N * C.N; c++)
    {
        unsigned st = e->cell_state.empty() ? 0u : e->cell_state[c];
        if (visited) visited[c] = (st & PP_CS_VISITED) ? 1 : 0;
        bool touched = (st & PP_CS_TOUCHED) != 0;
        if (g) g[c] = touched ? e->nm_g[c] : 0.0f;
        if (f) f[c] = touched ? e->nm_f[c] : pp_h2d(C, c / C.N, c % C.N);
    }
}

void emu_find_path(void* h, float vel, const float* s, orc_result* res, float* path_xyh, float* curv,
                   int path_cap, orc_pop* pops, int pop_cap)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    PPWork wk;
    const char* fixed = std::getenv("PP_EMU_FIXED");
    if (fixed && fixed[0] == '1') setup_work(e, wk, 1 << 20, 1 << 19, 1 << 16, 4096);
    else { setup_work(e, wk, 64, 48, 32, 4096); attach_arena(e, wk, 1 << 20, 1 << 19, 1 << 16); }
    static_assert(sizeof(PPPop) == sizeof(orc_pop), "pop structs must match");
    wk.trace = reinterpret_cast<PPPop*>(pops); wk.trace_cap = pops ? pop_cap : 0;
    PPState st = pp_host_set_start(C, e->fr, s[0], s[1], s[2], vel);
    PPGroup G = group_of(e);
    PPSmem sm;
    PPResult r;
    PPWarpSerial w;
    wk.lazy_sid = e->hist_on ? &e->hist_sid : nullptr;
    pp_search_exact(w, C, e->m.off_xy.data(), G, st, wk, sm, r);
    e->lazy_init = false;
    e->last = r; e->last_vel = vel; e->have_last = true;
    res->success = r.success; res->cost = r.cost; res->n_pops = r.n_pops; res->n_pops_bin_oob = r.n_pops_bin_oob;
    if (r.status) std::fprintf(stderr, "emu_find_path: status %d\n", r.status);
    // assemble in the reference's order (HybridAStar.cpp:208-262): reversed Dubins samples, then the chain
    int n = 0;
    if (r.success)
    {
        std::vector<PPPathPt> seq;
        for (int k = r.n_dubins - 1; k >= 0; k--) seq.push_back(wk.path[k]);
        for (int k = 0; k < r.n_chain; k++) seq.push_back(wk.path[r.n_dubins + k]);
        n = (int)seq.size();
        for (int k = 0; k < n && k < path_cap; k++)
        {
            float wx, wy, wh;
            pp_host_to_world(e->fr, seq[k].x, seq[k].y, seq[k].heading, wx, wy, wh);
            path_xyh[3 * k] = wx; path_xyh[3 * k + 1] = wy; path_xyh[3 * k + 2] = wh;
            curv[k] = (k == 0) ? 0.0f : seq[k - 1].curvature;   // curvature shifted by one (HybridAStar.cpp:212, :261)
        }
    }
    res->n_path = n;
}

// K-POP mode on one host lane (same signature as port_find_path_kpop); h1_field = 2D distance field, N*N floats
void emu_find_path_kpop(void* h, float vel, const float* s, int k, const float* h1_field, int max_nodes, orc_result* res,
                        float* path_xyh, float* curv, int path_cap, orc_pop* pops, int pop_cap)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    PPKWork wk;
    std::vector<PPKNode> nodes(max_nodes);
    int tc = 1; while (tc < 2 * max_nodes) tc <<= 1;
    std::vector<PPKSlot> table(tc);
    std::memset(table.data(), 0xFF, sizeof(PPKSlot) * table.size());     // all-ones = empty, as the C ABI hands it to the kernel
    int levels = 1; while (((size_t)PP_K_RUN0 << (levels - 1)) < (size_t)max_nodes + PP_K_RUN0 && levels < PP_K_LEVELS) levels++;
    std::vector<PPKEntry> arena((size_t)PP_K_RUN0 * ((1 << levels) - 1)), ta(max_nodes + 2 * PP_K_RUN0), tb(max_nodes + 2 * PP_K_RUN0);
    std::vector<PPPathPt> path(4096);
    wk.nodes = nodes.data(); wk.nodes_cap = max_nodes; wk.table = table.data(); wk.table_cap = tc;
    wk.arena = arena.data(); wk.tmp_a = ta.data(); wk.tmp_b = tb.data(); wk.lsm_levels = levels; wk.h1 = h1_field;
    wk.path = path.data(); wk.path_cap = (int)path.size();
    wk.trace = reinterpret_cast<PPPop*>(pops); wk.trace_cap = pops ? pop_cap : 0;
    PPState st = pp_host_set_start(C, e->fr, s[0], s[1], s[2], vel);
    PPGroup G = group_of(e);
    std::unique_ptr<PPKSmem> sm(new PPKSmem());
    PPResult r;
    PPWarpSerial w;
    pp_search_kpop(w, C, e->m.off_xy.data(), G, st, k, wk, *sm, r);
    res->success = r.success; res->cost = r.cost; res->n_pops = r.n_pops; res->n_pops_bin_oob = r.n_pops_bin_oob;
    if (r.status) std::fprintf(stderr, "emu_find_path_kpop: status %d\n", r.status);
    {   // the query must leave its hash table empty (all-ones) for the slot's next query
        const unsigned char* b = reinterpret_cast<const unsigned char*>(table.data());
        size_t dirty = 0;
        for (size_t t = 0; t < sizeof(PPKSlot) * table.size(); t++) dirty += (b[t] != 0xFF);
        if (dirty) { std::fprintf(stderr, "emu_find_path_kpop: %zu table bytes left dirty\n", dirty); res->success = -1; }
    }
    if (std::getenv("PP_EMU_VERBOSE")) std::fprintf(stderr, "kpop k=%d: pops %d iterations %d entries taken %d nodes %d\n", k, r.n_pops, r.n_lazy_searches, r.n_lazy_pops, r.n_closed);
    int n = 0;
    if (r.success)
    {
        std::vector<PPPathPt> seq;
        for (int q = r.n_dubins - 1; q >= 0; q--) seq.push_back(wk.path[q]);
        for (int q = 0; q < r.n_chain; q++) seq.push_back(wk.path[r.n_dubins + q]);
        n = (int)seq.size();
        for (int q = 0; q < n && q < path_cap; q++)
        {
            float wx, wy, wh;
            pp_host_to_world(e->fr, seq[q].x, seq[q].y, seq[q].heading, wx, wy, wh);
            path_xyh[3 * q] = wx; path_xyh[3 * q + 1] = wy; path_xyh[3 * q + 2] = wh;
            curv[q] = (q == 0) ? 0.0f : seq[q - 1].curvature;
        }
    }
    res->n_path = n;
}

// pp_fmath.h on n inputs: kind 0 sin, 1 cos, 2 atan2(y = a, x = b), 3 acos
void emu_fmath_batch(int kind, const float* a, const float* b, float* out, int n)
{
    for (int k = 0; k < n; k++)
    {
        float s, c;
        switch (kind)
        {
            case 0: pp_fm_sincos(a[k], s, c); out[k] = s; break;
            case 1: pp_fm_sincos(a[k], s, c); out[k] = c; break;
            case 2: out[k] = pp_fm_atan2(a[k], b[k]); break;
            default: out[k] = pp_fm_acos(a[k]); break;
        }
    }
}

// K-POP APF sum of n poses (x, y, heading): through the spatial index (as the kernel does) and by scanning every obstacle
void emu_kapf_batch(void* h, const float* xyh, int n, float* via_bins, float* via_scan)
{
    Emu* e = static_cast<Emu*>(h);
    const PPConsts& C = e->m.C;
    PPGroup G = group_of(e);
    for (int k = 0; k < n; k++)
    {
        const float x = xyh[3 * k], y = xyh[3 * k + 1], hd = xyh[3 * k + 2];
        float a = 0.0f, b = 0.0f;
        for (int q = 0; q < G.K; q++)
        {
            float t = pp_kapf_term(C, G.apf[3 * q], G.apf[3 * q + 1], G.apf[3 * q + 2], x, y, hd);
            if (t != 0.0f) b = b + t;
        }
        int ci, cj;
        if (pp_collision_free(C, G.map, x, y, ci, cj) || (ci > -1 && ci < C.N && cj > -1 && cj < C.N))
        {
            if (G.bin_off)
            {
                const int bb = (ci >> G.bin_shift) * G.bin_n + (cj >> G.bin_shift);
                for (int qq = G.bin_off[bb]; qq < G.bin_off[bb + 1]; qq++)
                {
                    const int q = G.bin_idx[qq];
                    float t = pp_kapf_term(C, G.apf[3 * q], G.apf[3 * q + 1], G.apf[3 * q + 2], x, y, hd);
                    if (t != 0.0f) a = a + t;
                }
            }
            else a = b;
        }
        else a = b;                                        // outside the grid: the search never evaluates the APF there
        via_bins[k] = a; via_scan[k] = b;
    }
}

} // extern "C"
