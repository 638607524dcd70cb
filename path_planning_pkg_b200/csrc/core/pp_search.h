// EXACT (single-pop) Hybrid A* search of one query, executed by one warp.
//
// Parity target: the expanded-node sequence, cost and path of the unmodified reference
// (lib/HybridAStar.cpp:93-199) on a freshly constructed planner.  To get there the
// two history-dependent containers of the reference are emulated, not "cleaned up":
//   * 3D open list  : std::set<Node3D>   -> PPRbTree<PPNode3> under `(a != b) && a.f < b.f` (F5)
//   * 3D closed set : unordered_set      -> exact (cell, bin) hash set + append-only log (F6)
//   * holonomic h1  : AStar::find_path(i,j), the lazily evaluated, cached, early-terminating 2D A*
//                     with its own std::set / unordered_set (F4) -> pp_lazy_astar() below.
// Work split inside the warp (template parameter W = lane policy):
//   control lane 0 : container walks (sequential by nature: every comparison depends on the last)
//   all lanes      : successor roll-out (one lane per steering primitive), collision lookup,
//                    APF (lanes over obstacles, order-preserving sum), Dubins candidates (lanes over
//                    (successor, CSC type) pairs), Dubins-shot sampling + collision check, scratch init.
// With W::LANES == 1 the same source runs on one host thread inside tests/cpp/host_emul.cpp.
#ifndef PP_SEARCH_H
#define PP_SEARCH_H

#include <stddef.h>
#include "pp_defs.h"
#include "pp_math.h"
#include "pp_dubins.h"
#include "pp_rbtree.h"
#include "pp_arena.h"

#define PP_MAX_SUCC 16
#define PP_NEAR_CAP 64
// Speculative parallel walks (phase 4a / 4b below): measured on B200 they shorten nothing -- a lone warp is bound by the
// dependent-issue latency of the commit code, a crowded SM by instruction fetch, and the extra code and shared memory cost more
// than the walks save -- so the default build commits sequentially.  -DPP_EXACT_SPEC=1 builds the variant;
// the CPU race harness (tests/cpp/search_mt.cpp) compiles it in both forms.
#ifndef PP_EXACT_SPEC
#define PP_EXACT_SPEC 0
#endif
#define PP_SPEC_WALKS 16      /* speculative tree walks per batch: (successor, find | insert) pairs, or the 8 neighbours of a 2D pop x 2 */
#define PP_SPEC_PATH 40       /* nodes recorded per walk: header + the height bound of a red-black tree of 2^19 nodes */
#define PP_MLOG_CAP 48        /* mutated nodes remembered between two speculation batches */

// Optional per-phase cycle accounting (library variant built with -DPP_PROFILE; never in the shipped .so).
// Phases: 0 scratch init, 1 pop + closed insert + erase, 2 roll-out/collision/APF, 3 Dubins candidates,
// 4 closed.find, 5 open.find (+erase), 6 lazy 2D A*, 7 open.insert.
#if defined(PP_PROFILE) && defined(__CUDA_ARCH__)
#define PP_PROF_DECL long long pp_prof_t = clock64(); long long pp_prof_a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PP_PROF_MARK(k) { long long t__ = clock64(); pp_prof_a[k] += t__ - pp_prof_t; pp_prof_t = t__; }
#define PP_PROF_FLUSH if (lane == 0) { for (int q__ = 0; q__ < 8; q__++) atomicAdd(&pp_prof_acc[q__], (unsigned long long)pp_prof_a[q__]); \
                                        for (int q__ = 0; q__ < 8; q__++) atomicAdd(&pp_prof_acc[8 + q__], (unsigned long long)pp_prof_c[q__]); }
// event counters: 0 find walks used as speculated, 1 find walks redone, 2 insert walks speculated and used, 3 speculated but redone,
// 4 insert walks without speculation (h1 not cached), 5 successors committed, 6 inserts attached, 7 mutation-log entries
#define PP_PROF_COUNT(k, v) { pp_prof_c[k] += (v); }
#define PP_PROF_CDECL long long pp_prof_c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#else
#define PP_PROF_COUNT(k, v)
#define PP_PROF_CDECL
#define PP_PROF_DECL
#define PP_PROF_MARK(k)
#define PP_PROF_FLUSH
#endif

// cell_state word of the lazy 2D A*: bit31 = _visted, bit30 = node_map entry touched this query,
// bits 0..29 = id of the lazy search that closed the cell.
#define PP_CS_VISITED 0x80000000u
#define PP_CS_TOUCHED 0x40000000u
#define PP_CS_STAMP   0x3fffffffu

struct PPNode3   // 3D open-list entry, 64 B; w = everything a tree walk reads (one 128-bit load)
{
    PPWalk   w;          // left, right, f, key = (ci*N + cj)*(bins+1) + bin
    int      parent, color;
    float    g, x;
    float    y, heading, vmin_sqr;
    int      curv;
    int      prev, bin;
    int      pad0, pad1;
};

struct PPNode2   // 2D open-list entry (copy of a Node2D at insertion time), 32 B
{
    PPWalk   w;          // left, right, f, key = cell
    int      parent, color;
    float    g;
    int      prev;       // cell of the closed parent, -1 = none
};

static_assert(sizeof(PPNode3) == 64 && sizeof(PPNode2) == 32, "node sizes");
static_assert(offsetof(PPNode3, parent) == offsetof(PPRbHead, parent) && offsetof(PPNode3, color) == offsetof(PPRbHead, color) &&
              offsetof(PPNode2, parent) == offsetof(PPRbHead, parent) && offsetof(PPNode2, color) == offsetof(PPRbHead, color),
              "both node types must start with the PPRbHead prefix (pp_rbtree.h)");

struct PPClosed3
{
    float    x, y, heading, g, f, vmin_sqr;
    int      curv, bin;
    unsigned key;
    int      prev;       // index of the parent in the closed log, -1 = none
};

struct PPHashSlot { unsigned key; int idx; };   // closed-set hash slot: key + index into the closed log (-1 = empty)

struct PPPathPt { float x, y, heading, curvature; };

struct PPSucc
{
    float x, y, heading, g, vmin_sqr;
    int   curv, bin, ci, cj, ok;
};

// per (map, goal) group
struct PPGroup
{
    const float* map;    // N*N log-odds, map[i*N + j]  (i along grid x)
    const float* apf;    // K x (x, y, radius) in grid-frame metres (Grid3D.cpp:22-44)
    int          K;
    int          bin_shift;   // spatial index of the APF list: square bins of (1 << bin_shift) cells, bin_n per side
    int          bin_n;
    int          pad;
    const int*   bin_off;     // bin_n*bin_n + 1 offsets into bin_idx, nullptr = no index (scan all K)
    const int*   bin_idx;     // per bin: ascending indices of the obstacles whose disc can reach the bin
    PPFrame      frame;
};

// per-query scratch (one slot per resident warp)
struct PPWork
{
    PPNode3*   open3;      int open3_cap;
    PPClosed3* closed;     int closed_cap;
    PPHashSlot* chash;     int chash_cap;     // power of two
    unsigned*  cell_state;                    // N*N
    float*     nm_g;                          // N*N  node_map[i][j]._cost_g
    float*     nm_f;                          // N*N  node_map[i][j]._cost_f
    float*     cl_g;                          // N*N  g of the closed copy (valid iff stamp == search id)
    int*       cl_prev;                       // N*N  parent cell of the closed copy
    PPNode2*   open2;      int open2_cap;
    PPPathPt*  path;       int path_cap;      // [dubins samples (forward) | parent chain (terminal -> start)]
    PPPop*     trace;      int trace_cap;     // optional
    // Planner-object history: nullptr = every query starts on the freshly constructed 2D cache
    // (cell_state cleared here).  Otherwise cell_state / nm_g / nm_f are the carried `_visted` + `_node_map` of ONE
    // reference planner object and *lazy_sid its running lazy-search id: nothing is cleared, the query continues on
    // whatever the object's earlier find_path calls left behind (AStar::reset() only drops the visited flags).
    unsigned*  lazy_sid = nullptr;
    // Growable containers (pp_arena.h): the pools above are the slot's small fixed ones; with an arena a container that
    // fills up moves into a block twice the size, up to the hard caps below (the caller's max_expansions / max_open /
    // max_open2d).  arena == nullptr: fixed pools, the caps are the pool sizes.
    PPArena*   arena = nullptr;
    int        closed_max = 0, open3_max = 0, open2_max = 0;
};

// one neighbour of the 2D node being expanded, as fetched up front (pp_lazy_astar); 16 B = one shared-memory word quad
struct
#if defined(__CUDACC__) || defined(__GNUC__)
__attribute__((aligned(16)))
#endif
PPLazyNb { int cell; float map; unsigned state; float nmf; };

struct PPSmem   // per-warp staging area (shared memory on the device)
{
    PPSucc succ[PP_MAX_SUCC];
    float  cand[PP_MAX_SUCC * 4];        // Dubins candidate lengths [successor][type] (also APF accumulators)
    int    near_idx[PP_NEAR_CAP];
    PPLazyNb lazy_nb[8];                 // neighbour staging of the control lane's lazy 2D A* (pp_lazy_astar)
    // staged Dubins evaluation (pp_dubins_h2_warp)
    float  d_centre[PP_MAX_SUCC * 4];    // start circle centres per successor: right x, y, left x, y
    float  d_theta[PP_MAX_SUCC * 4];     // atan2f of the centre offset per (successor, type)
    float  d_acos[PP_MAX_SUCC * 2];      // acosf(2r/dist) per (successor, RSL | LSR)
    float  d_sin[PP_MAX_SUCC * 4];       // sin / cos of theta_t1 and p2 per (successor, RSL | LSR)
    float  d_cos[PP_MAX_SUCC * 4];
#if PP_EXACT_SPEC
    // speculative walks (pp_spec_*): every walk of one expansion is done up front by its own lane on the tree as it stands;
    // the control lane then commits the successors in the reference's order and re-walks only what an earlier commit touched
    int    spec_path[PP_SPEC_WALKS][PP_SPEC_PATH];
    int    spec_np[PP_SPEC_WALKS];       // nodes on the path, -1 = no speculation for this walk
    int    spec_a[PP_SPEC_WALKS];        // find: the node found (or NIL); insert: the parent
    int    spec_b[PP_SPEC_WALKS];        // insert: 1 = left child, 0 = right child, -1 = equivalent element exists (dropped)
    float  spec_f[PP_SPEC_WALKS];        // insert: the f the position was searched for
    int    mlog[PP_MLOG_CAP];            // PPRbTree::mut log of the 3D open list
    int    mcount;
#endif
};

#if PP_EXACT_SPEC
// Would the recorded walk still visit the same nodes?  True iff none of them had a child pointer changed since the walk
// (all lanes; the same answer on every lane).
template <class W>
PP_HD bool pp_spec_valid(const W& w, const int* path, int np, const int* mlog, int mcount)
{
    // every lane leaves through the ballot: it is also the barrier between these reads of the log and the control lane's next
    // writes to it
    bool hit = (np < 0 || np > PP_SPEC_PATH || mcount > PP_MLOG_CAP);
    if (!hit)
        for (int base = 0; base < np; base += W::LANES)
        {
            const int q = base + w.lane();
            const int mine = (q < np) ? path[q] : -2;
            for (int j = 0; j < mcount; j++) hit = hit || (mine == mlog[j]);
        }
    return w.ballot(hit) == 0u;
}
#endif

// goal-side constants of the Dubins heuristic, evaluated once per query
struct PPDubinsGoal { float grx, gry, glx, gly; };

// ---- lane policy for a single host thread (tests only); the device policy lives in pp_kernels.cu ----
struct PPWarpSerial
{
    enum { LANES = 1, BW = 1 };                      // BW: width of one ballot group (a hardware warp)
    PP_HD int lane() const { return 0; }
    PP_HD int wlane() const { return 0; }            // lane inside its ballot group
    PP_HD int warp() const { return 0; }             // index of the ballot group
    PP_HD void sync() const {}
    PP_HD void wsync() const {}
    PP_HD unsigned ballot(bool p) const { return p ? 1u : 0u; }
    PP_HD unsigned lanemask_lt() const { return 0u; }
    template <class T> PP_HD T shfl(T v, int) const { return v; }
    PP_HD bool any(bool p, int*) const { return p; }
    // exclusive count of `p` over the lanes below this one, `total` over all lanes
    PP_HD int scan_count(bool p, int*, int& total) const { total = p ? 1 : 0; return 0; }
};

// ---------------------------------------------------------------------------------------------------
// (the comparators of the reference, Node3D.h:39-54 / Node2D.h:27-41, live in pp_rbtree.h: pp_lt)

// Euclidean 2D heuristic, Grid2D.cpp:303-316 (recomputed instead of stored: 2 muls, 1 add, 1 sqrt)
PP_HD float pp_h2d(const PPConsts& C, int i, int j)
{
    float dx = (C.n45 - i) * C.res;
    float dy = (C.n2 - j) * C.res;
    return sqrtf(dx * dx + dy * dy);
}

// out-of-line copy for the sites that run once per lazy search (the per-neighbour one stays inline)
PP_HD_NOINLINE_FN float pp_h2d_call(const PPConsts& C, int i, int j) { return pp_h2d(C, i, j); }

struct PPLazy
{
    PPRbTree<PPNode2> open;
    unsigned search_id;
    int status;
    int n_searches, n_pops;
    PPArena* arena = nullptr;   // nullptr = fixed pool
    int max_cap = 0;            // hard cap of the 2D open-list pool
    int blk = -1;               // arena class of the current pool, -1 = the slot's fixed pool
};

// The 2D open list outgrew its pool: move it into a block twice the size (single lane: this happens in the middle of the
// control lane's sequential work; the pool is small and the event rare).  Indices stay valid, so the tree is untouched.
PP_HD_NOINLINE_FN bool pp_lazy_grow(PPLazy& L)
{
    if (!L.arena || L.open.cap >= L.max_cap) return false;
    long long want = 2ll * L.open.cap;
    if (want > L.max_cap) want = L.max_cap;
    int k = pp_arena_class((unsigned long long)want * sizeof(PPNode2));
    PPNode2* nb = (PPNode2*)pp_arena_alloc_patient(L.arena, k, k);
    if (!nb) return false;
    const unsigned long long* src = (const unsigned long long*)L.open.n;
    unsigned long long* dst = (unsigned long long*)nb;
    const size_t words = (size_t)L.open.next * (sizeof(PPNode2) / 8);
    for (size_t q = 0; q < words; q++) dst[q] = src[q];
    if (L.blk >= 0) pp_arena_free(L.arena, L.open.n, L.blk);
    L.open.n = nb; L.open.cap = (int)want; L.blk = k;
    return true;
}

PP_HD void pp_lazy_touch(const PPConsts& C, PPWork& wk, int cell)
{
    unsigned st = wk.cell_state[cell];
    if (!(st & PP_CS_TOUCHED))
    {
        wk.nm_g[cell] = 0.0f;
        wk.nm_f[cell] = pp_h2d_call(C, cell / C.N, cell % C.N);
        wk.cell_state[cell] = st | PP_CS_TOUCHED;
    }
}

// AStar::update_visted + Grid2D::update_costs (AStar.cpp:209-218, Grid2D.cpp:219-227)
PP_HD_NOINLINE_FN void pp_lazy_update_visited(const PPConsts& C, PPWork& wk, float total, int last_cell)
{
    int c = last_cell;
    while (c >= 0)
    {
        pp_lazy_touch(C, wk, c);
        wk.cell_state[c] |= PP_CS_VISITED;
        wk.nm_f[c] = total - wk.cl_g[c];
        c = wk.cl_prev[c];
    }
}

PP_HD_NOINLINE_FN bool pp_lazy_insert(PPLazy& L, int cell, float g, float f, int prev)
{
    PPKey k; k.key = (unsigned)cell; k.f = f;
    int p; bool left;
    if (!L.open.insert_pos(k, p, left)) return true;   // silently dropped (F5)
    int s = L.open.alloc();
    if (s == PP_RB_NIL && pp_lazy_grow(L)) s = L.open.alloc();
    if (s == PP_RB_NIL) { L.status |= PP_STATUS_OPEN2D_OVERFLOW; return false; }
    PPNode2& n = L.open.n[s];
    n.w.f = f; n.w.key = (unsigned)cell; n.g = g; n.prev = prev;
    L.open.insert_and_rebalance(left, s, p);
    return true;
}

// AStar::find_path(i, j) + a_star_search (AStar.cpp:100-113, :118-186) on per-query scratch that
// starts in the freshly-constructed state (g = 0, f = Euclidean h, nothing visited).
PP_HD_NOINLINE_FN float pp_lazy_astar(const PPConsts& C, const float* map, const PPFrame& F, PPWork& wk,
                                   PPLazy& L, int ci, int cj, PPLazyNb* nbs)
{
    const int N = C.N;
    int cell = ci * N + cj;
    // the per-cell arrays, held in registers across the tree calls below (wk lives in the caller's stack frame: every use
    // through the reference would be a local-memory load first)
    unsigned* const cell_state = wk.cell_state;
    float* const nm_g = wk.nm_g;
    float* const nm_f = wk.nm_f;
    float* const cl_g = wk.cl_g;
    int* const cl_prev = wk.cl_prev;
    unsigned st = cell_state[cell];
    if (st & PP_CS_VISITED) return nm_f[cell];

    // Grid2D::set_start_node_grid -> soft_reset (Node2D.cpp:34-39)
    float h0 = pp_h2d_call(C, ci, cj);
    nm_g[cell] = 0.0f;
    nm_f[cell] = h0;
    cell_state[cell] = st | PP_CS_TOUCHED;

    unsigned sid = ++L.search_id;
    L.n_searches++;
    L.open.clear();
    if (!pp_lazy_insert(L, cell, 0.0f, h0, -1)) return FLT_MAX;
    const int goal_cell = F.goal_ci * N + F.goal_cj;

    while (!L.open.empty())
    {
        int it = L.open.begin();
        int c = (int)L.open.n[it].w.key;
        float cg;
        unsigned cs = cell_state[c];
        if ((cs & PP_CS_STAMP) == sid) cg = cl_g[c];   // re-pop: unordered_set::insert returns the old copy
        else
        {
            cg = L.open.n[it].g;
            cl_g[c] = cg;
            cl_prev[c] = L.open.n[it].prev;
            cell_state[c] = (cs & ~PP_CS_STAMP) | sid;
        }
        L.open.erase(it);
        L.n_pops++;

        int pi = c / N, pj = c - pi * N;
        if (c == goal_cell)
        {
            float total = cg + pp_h2d_call(C, pi, pj);   // copy's _cost_f = g + h
            pp_lazy_update_visited(C, wk, total, c);
            return total;
        }

        // Grid2D::get_neighbors (Grid2D.cpp:72-96): all valid neighbours first, in action order.
        // The 8 neighbours are distinct cells whose per-cell words nobody else touches while this node is expanded, so the
        // map value, cell_state and node_map f of all of them are fetched up front as 24 independent loads (one memory round
        // trip instead of 24 dependent ones) into the staging records `nbs`; they are consumed in action order by ONE rolled
        // loop.  (Until round 2 the consuming loop ran over register arrays, which made the compiler unroll it 8 times: 46 KB
        // of SASS for this function, more than the SM's 32 KB instruction cache, in a kernel that is bound by instruction
        // fetch at bench occupancy -- DESIGN.md section 7.)
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            const int i = pi + C.act_di[k], j = pj + C.act_dj[k];
            const bool in = (k < C.n_act2d && i > -1 && i < N && j > -1 && j < N);
            const int a = in ? i * N + j : c;                // harmless in-bounds address for the masked-out slots
            PPLazyNb e;
            e.cell = in ? a : -1; e.map = map[a]; e.state = cell_state[a]; e.nmf = nm_f[a];
            nbs[k] = e;
        }
#pragma unroll 1
        for (int k = 0; k < 8; k++)
        {
            const PPLazyNb e = nbs[k];
            if (e.cell < 0 || !(e.map < C.log_thr)) continue;
            const int nb = e.cell;
            const float w = C.act_cost[k];
            const unsigned ns = e.state;
            if (ns & PP_CS_VISITED)
            {
                float total = e.nmf + cg + w;
                pp_lazy_update_visited(C, wk, total, c);
                return total;
            }
            if ((ns & PP_CS_STAMP) == sid) continue;    // in the closed set of this search
            const float hn = pp_h2d(C, pi + C.act_di[k], pj + C.act_dj[k]);
            float cur_f = e.nmf;
            if (!(ns & PP_CS_TOUCHED))                  // pp_lazy_touch: first use this query -> g = 0, f = h
            {
                cur_f = hn;
                nm_g[nb] = 0.0f; nm_f[nb] = cur_f;
                cell_state[nb] = ns | PP_CS_TOUCHED;
            }
            PPKey key; key.key = (unsigned)nb; key.f = cur_f;    // node_map's current (possibly stale) f
            const int it_node = L.open.find(key);
            const float newg = cg + w;
            bool ins = (it_node == PP_RB_NIL);
            if (!ins && newg < L.open.n[it_node].g) { L.open.erase(it_node); ins = true; }
            if (ins)
            {
                const float nf = newg + hn;
                nm_g[nb] = newg; nm_f[nb] = nf;
                if (!pp_lazy_insert(L, nb, newg, nf, c)) return FLT_MAX;
            }
        }
    }
    return FLT_MAX;
}

// ---------------------------------------------------------------------------------------------------
// stateless pieces (also exported one by one through the C ABI for parity tests)

// One successor of VehicleModel::get_neighbors (VehicleModel.cpp:63-105) for steering index i.
// off_xy is [S][bins+1][2]; column `bins` duplicates column 0.  Returns false when the
// primitive is pruned by the lateral-acceleration limit.
PP_HD bool pp_rollout_one(const PPConsts& C, const float* off_xy, float x, float y, float heading, float g,
                          float vmin_sqr, int bin, int i, PPSucc& o)
{
    float v2 = 0.0f;
    if (!(vmin_sqr < 1.0f))
    {
        float lat = vmin_sqr * C.abs_curv[i];
        if (lat > C.max_lat_acc) return false;
        float acc_long = (float)sqrt(1.0 - (double)((lat * lat) / C.max_lat_acc_sqr));
        v2 = vmin_sqr - 2 * acc_long * C.ts;
    }
    const float* off = off_xy + ((size_t)i * (C.bins + 1) + bin) * 2;
    o.x = x + off[0];
    o.y = y + off[1];
    o.heading = pp_wrap_pi(heading + C.off_heading[i]);
    o.g = g + C.act_cost3d[i];
    o.vmin_sqr = v2;
    o.curv = i;
    o.bin = pp_heading_index(o.heading, C.precision);
    return true;
}

// bounds + collision lookup of Grid3D::get_neighbors (Grid3D.cpp:56-59)
PP_HD bool pp_collision_free(const PPConsts& C, const float* map, float x, float y, int& ci, int& cj)
{
    ci = pp_f2i_x86(x / C.res);
    cj = pp_f2i_x86(y / C.res);
    return (ci > -1) && (ci < C.N) && (cj > -1) && (cj < C.N) && (map[ci * C.N + cj] < C.log_thr);
}

// rounded-index lookup of Grid3D::check_path (Grid3D.cpp:83-90); true = blocked
PP_HD bool pp_path_point_blocked(const PPConsts& C, const float* map, float x, float y)
{
    int i1 = pp_f2i_x86(roundf(x / C.res));
    int j1 = pp_f2i_x86(roundf(y / C.res));
    return (i1 < 0) || (i1 >= C.N) || (j1 < 0) || (j1 >= C.N) || (map[i1 * C.N + j1] >= C.log_thr);
}

// one obstacle's term of Grid3D::get_field_intensity (Grid3D.cpp:209-223)
PP_HD_NOINLINE_FN float pp_apf_term(const PPConsts& C, float ox, float oy, float radius, float x, float y, float heading)
{
    float dx = ox - x, dy = oy - y;
    float distance = pp_hypotf(dx, dy);
    if (!(distance < radius)) return 0.0f;
    float angle = fabsf(pp_wrap_pi(heading - pp_atan2f(dy, dx)));
    float a = C.apf_alpha - angle;
    angle = (a < 0.0f) ? 0.0f : a;
    double d = 1.0 / (double)distance - 1.0 / (double)radius;
    float fp = (float)((double)C.apf_k * (d * d));
    fp = fp * angle / C.apf_alpha;
    return fp;
}

// Order-preserving warp sum of the APF terms over the obstacle index list idx[0..n) (or 0..n-1 when
// idx == nullptr): std::accumulate from T(0) in obstacle order (Grid3D.cpp:226).  Zero terms are skipped
// (x + 0 == x), non-zero terms are added one by one in index order.
template <class W>
PP_HD_NOINLINE_FN float pp_apf_sum(const W& w, const PPConsts& C, const float* apf, const int* idx, int n,
                       float x, float y, float heading)
{
    float acc = 0.0f;
#pragma unroll 1
    for (int base = 0; base < n; base += W::LANES)
    {
        int q = base + w.lane();
        float term = 0.0f;
        if (q < n)
        {
            int k = idx ? idx[q] : q;
            float ox = apf[3 * k], oy = apf[3 * k + 1], r = apf[3 * k + 2];
            // cheap conservative reject before the double-precision path
            float dx = ox - x, dy = oy - y, lim = r * 1.001f + 1e-3f;
            if (dx * dx + dy * dy <= lim * lim) term = pp_apf_term(C, ox, oy, r, x, y, heading);
        }
        unsigned m = w.ballot(term != 0.0f);
        while (m)
        {
            acc = acc + w.shfl(term, pp_ctz(m));
            m &= m - 1;
        }
    }
    return acc;
}

// Successors of one popped state, Grid3D::get_neighbors (Grid3D.cpp:47-74) = VehicleModel roll-out
// (one lane per steering primitive) + bounds/collision lookup + APF cost (lanes over obstacles).
// Results in sm.succ[0 .. 2A] (ok = 0 for pruned / colliding primitives); g includes the field cost.
template <class W>
PP_HD_NOINLINE_FN void pp_expand_warp(const W& w, const PPConsts& C, const float* off_xy, const PPGroup& G,
                          float px, float py, float ph, float pg, float pv2, int pcurv, int pbin, PPSmem& sm)
{
    const int lane = w.lane();
    const int n_succ_max = 2 * C.A + 1;
    int start_index = pcurv - C.A;
    if (start_index < 0) start_index = 0;
#pragma unroll 1
    for (int s = lane; s < n_succ_max; s += W::LANES)
    {
        PPSucc o; o.ok = 0;
        int i = start_index + s;
        int bin = (pbin > C.bins) ? C.bins : pbin;
        if (i < C.S && pp_rollout_one(C, off_xy, px, py, ph, pg, pv2, bin, i, o))
            o.ok = pp_collision_free(C, G.map, o.x, o.y, o.ci, o.cj) ? 1 : 0;
        sm.succ[s] = o;
    }
    // obstacles that can reach any successor (successors lie within ts of the parent), order kept
    int n_near = 0;
#pragma unroll 1
    for (int base = 0; base < G.K; base += W::LANES)
    {
        int k = base + lane;
        bool nearp = false;
        if (k < G.K)
        {
            float dx = G.apf[3 * k] - px, dy = G.apf[3 * k + 1] - py;
            float lim = G.apf[3 * k + 2] + C.ts * 1.5f + 0.05f;
            nearp = (dx * dx + dy * dy) <= lim * lim * 1.001f;
        }
        unsigned m = w.ballot(nearp);
        if (nearp)
        {
            const int pos = n_near + pp_popc(m & w.lanemask_lt());
            if (pos < PP_NEAR_CAP) sm.near_idx[pos] = k;
        }
        n_near += pp_popc(m);
    }
    const bool near_overflow = (n_near > PP_NEAR_CAP);
    w.sync();
    // APF of the surviving successors, added to g (and f), Grid3D.cpp:61-63
    if (near_overflow)
    {
        // rare: more obstacles in reach than the staging list holds -> one successor at a time over the full list
#pragma unroll 1
        for (int s = 0; s < n_succ_max; s++)
        {
            if (!sm.succ[s].ok) continue;
            float field = pp_apf_sum(w, C, G.apf, (const int*)0, G.K, sm.succ[s].x, sm.succ[s].y, sm.succ[s].heading);
            w.sync();
            if (lane == 0) sm.succ[s].g = sm.succ[s].g + field;
        }
    }
    else if (n_near > 0)
    {
        // one lane per (successor, near obstacle) pair, successor-major so that a successor's terms are met in
        // obstacle order; sm.cand[s] accumulates successor s's field (std::accumulate order, Grid3D.cpp:226)
#pragma unroll 1
        for (int s = lane; s < n_succ_max; s += W::LANES) sm.cand[s] = 0.0f;
        w.sync();
        const int n_pairs = n_succ_max * n_near;
#pragma unroll 1
        for (int base = 0; base < n_pairs; base += W::LANES)
        {
            int q = base + lane;
            float term = 0.0f;
            int s = 0;
            if (q < n_pairs)
            {
                s = q / n_near;
                if (sm.succ[s].ok)
                {
                    int k = sm.near_idx[q - s * n_near];
                    float ox = G.apf[3 * k], oy = G.apf[3 * k + 1], r = G.apf[3 * k + 2];
                    float sx = sm.succ[s].x, sy = sm.succ[s].y;
                    float dx = ox - sx, dy = oy - sy, lim = r * 1.001f + 1e-3f;   // cheap conservative reject
                    if (dx * dx + dy * dy <= lim * lim) term = pp_apf_term(C, ox, oy, r, sx, sy, sm.succ[s].heading);
                }
            }
            unsigned m = w.ballot(term != 0.0f);
            while (m)       // non-zero terms one by one, in pair order; zero terms are skipped (x + 0 == x)
            {
                const int src = pp_ctz(m);
                float tv = w.shfl(term, src);
                int ts = w.shfl(s, src);
                if (lane == 0) sm.cand[ts] = sm.cand[ts] + tv;
                m &= m - 1;
            }
        }
        w.sync();
#pragma unroll 1
        for (int s = lane; s < n_succ_max; s += W::LANES)
            if (sm.succ[s].ok) sm.succ[s].g = sm.succ[s].g + sm.cand[s];
    }
    w.sync();
}

// Dubins lengths of all (successor, CSC type) pairs of one expansion -> sm.cand[4*s + type]
// (Dubins::get_shortest_path_length, Dubins.cpp:19-69, evaluated for every surviving successor).  The ~10
// dependent float transcendentals of one candidate are spread over the warp in three parallel stages:
//   A  lanes = successors           : sin, cos of the successor heading -> start circle centres
//   B  lanes = (succ, type) + (succ, RSL|LSR) : atan2f of the centre offset  ||  acosf(2r/dist)
//   C  lanes = (succ, RSL|LSR, theta_t1|p2)   : sin, cos of the two tangent angles
//   D  lanes = (succ, type)         : float-only tail (pp_dubins_finish)
template <class W>
PP_HD_NOINLINE_FN void pp_dubins_h2_warp(const W& w, const PPConsts& C, const PPFrame& F, const PPDubinsGoal& gc, PPSmem& sm)
{
    const int lane = w.lane();
    const int n = 2 * C.A + 1;
    const float r = C.r_min;
#pragma unroll 1
    for (int s = lane; s < n; s += W::LANES)                                   // stage A
        if (sm.succ[s].ok)
        {
            float sn, cs;
            pp_sincosf(sm.succ[s].heading, sn, cs);
            sm.d_centre[4 * s] = sm.succ[s].x + r * sn;     sm.d_centre[4 * s + 1] = sm.succ[s].y - r * cs;
            sm.d_centre[4 * s + 2] = sm.succ[s].x - r * sn; sm.d_centre[4 * s + 3] = sm.succ[s].y + r * cs;
        }
    w.sync();
#pragma unroll 1
    for (int q = lane; q < 6 * n; q += W::LANES)                               // stage B
    {
        int s, type;
        if (q < 4 * n) { s = q >> 2; type = q & 3; } else { int e = q - 4 * n; s = e >> 1; type = 1 + (e & 1); }
        if (!sm.succ[s].ok) continue;
        bool s_right = (type == PP_RSR) || (type == PP_RSL), g_right = (type == PP_RSR) || (type == PP_LSR);
        float csx = s_right ? sm.d_centre[4 * s] : sm.d_centre[4 * s + 2], csy = s_right ? sm.d_centre[4 * s + 1] : sm.d_centre[4 * s + 3];
        float cgx = g_right ? gc.grx : gc.glx, cgy = g_right ? gc.gry : gc.gly;
        if (q < 4 * n) sm.d_theta[q] = pp_atan2f(cgy - csy, cgx - csx);
        else sm.d_acos[q - 4 * n] = pp_acosf(pp_dubins_acos_arg(r, csx, csy, cgx, cgy));
    }
    w.sync();
#pragma unroll 1
    for (int q = lane; q < 4 * n; q += W::LANES)                               // stage C
    {
        int s = q >> 2, type = 1 + ((q >> 1) & 1), which = q & 1;
        if (!sm.succ[s].ok) continue;
        float t1 = pp_dubins_theta_t1(type, sm.d_acos[2 * s + (type - 1)], sm.d_theta[4 * s + type]);
        float ang = which ? pp_dubins_p2(type, t1) : t1;
        float sa, ca;
        pp_sincosf(ang, sa, ca);
        sm.d_sin[q] = sa;
        sm.d_cos[q] = ca;
    }
    w.sync();
#pragma unroll 1
    for (int q = lane; q < 4 * n; q += W::LANES)                               // stage D
    {
        int s = q >> 2, type = q & 3;
        if (!sm.succ[s].ok) continue;
        bool s_right = (type == PP_RSR) || (type == PP_RSL), g_right = (type == PP_RSR) || (type == PP_LSR);
        float csx = s_right ? sm.d_centre[4 * s] : sm.d_centre[4 * s + 2], csy = s_right ? sm.d_centre[4 * s + 1] : sm.d_centre[4 * s + 3];
        float cgx = g_right ? gc.grx : gc.glx, cgy = g_right ? gc.gry : gc.gly;
        float ac = 0.0f, c1 = 0.0f, s1 = 0.0f, c2 = 0.0f, s2 = 0.0f, p[4];
        if (type == PP_RSL || type == PP_LSR)
        {
            int e = 4 * s + 2 * (type - 1);
            ac = sm.d_acos[2 * s + (type - 1)];
            c1 = sm.d_cos[e]; s1 = sm.d_sin[e]; c2 = sm.d_cos[e + 1]; s2 = sm.d_sin[e + 1];
        }
        sm.cand[q] = pp_dubins_finish(type, r, sm.succ[s].heading, F.goal_h, csx, csy, cgx, cgy, sm.d_theta[q], ac, c1, s1, c2, s2, p);
    }
    w.sync();
}

// ---------------------------------------------------------------------------------------------------
PP_HD unsigned pp_hash_key(unsigned key) { key *= 2654435761u; return key ^ (key >> 15); }

struct PPSearchState
{
    PPRbTree<PPNode3> open;
    PPLazy lazy;
    int n_closed;
    int status;
    int max_open;
    int open_blk;        // arena class of the 3D open-list pool, -1 = the slot's fixed pool
};

// ---- cooperative growth of the per-query containers (all lanes, at a converged point of the search loop) ----------
template <class W>
PP_HD void pp_coop_copy8(const W& w, void* dst, const void* src, size_t bytes)
{
    const unsigned long long* s = (const unsigned long long*)src;
    unsigned long long* d = (unsigned long long*)dst;
    const size_t words = bytes / 8;
#pragma unroll 1
    for (size_t q = w.lane(); q < words; q += W::LANES) d[q] = s[q];
}

template <class W, class T>
PP_HD T* pp_bcast_ptr(const W& w, T* p)
{
    unsigned long long v = (unsigned long long)p;
    unsigned lo = w.shfl((unsigned)(v & 0xffffffffull), 0), hi = w.shfl((unsigned)(v >> 32), 0);
    return (T*)(((unsigned long long)hi << 32) | (unsigned long long)lo);
}

PP_HD bool pp_hash_slot_claim(PPHashSlot* slot, unsigned key, int idx)
{
    // the table is filled by all lanes at once when the closed set moves to a larger table: claim with one 64-bit CAS
    union { PPHashSlot s; unsigned long long u; } empty, want;
    empty.s.key = 0xffffffffu; empty.s.idx = -1;
    want.s.key = key; want.s.idx = idx;
#ifdef __CUDA_ARCH__
    return atomicCAS((unsigned long long*)slot, empty.u, want.u) == empty.u;
#else
    return __sync_bool_compare_and_swap((unsigned long long*)slot, empty.u, want.u);
#endif
}

// The 3D open-list pool is (nearly) full: move it into a block twice the size.  Node indices stay valid.
template <class W>
PP_HD_NOINLINE_FN void pp_grow_open3(const W& w, PPWork& wk, PPSearchState& S)
{
    const int lane = w.lane();
    PPNode3* nb = nullptr; PPNode3* old = nullptr;
    int used = 0, k = -1, old_blk = -1;
    long long want = 0;
    if (lane == 0)
    {
        want = 2ll * S.open.cap;
        if (want > wk.open3_max) want = wk.open3_max;
        k = pp_arena_class((unsigned long long)want * sizeof(PPNode3));
        nb = (PPNode3*)pp_arena_alloc_patient(wk.arena, k, k);
        old = S.open.n; used = S.open.next; old_blk = S.open_blk;
        if (!nb) S.status |= PP_STATUS_ARENA_EXHAUSTED;
    }
    nb = pp_bcast_ptr(w, nb);
    if (!nb) return;
    old = pp_bcast_ptr(w, old);
    used = w.shfl(used, 0);
    pp_coop_copy8(w, nb, old, (size_t)used * sizeof(PPNode3));
    w.sync();
    if (lane == 0)
    {
        if (old_blk >= 0) pp_arena_free(wk.arena, old, old_blk);
        S.open.n = nb; S.open.cap = (int)want; S.open_blk = k;
    }
    w.sync();
}

// The closed log is full: move it into a block twice the size and rebuild the (cell, bin) hash set in a table twice the
// size (placement inside an open-addressing table carries no meaning, so the lanes re-insert concurrently).
// wk.closed / wk.chash of EVERY lane are updated; closed_blk / chash_blk track the arena classes (-1 = fixed pool).
template <class W>
PP_HD_NOINLINE_FN bool pp_grow_closed(const W& w, PPWork& wk, int n_closed, int& closed_blk, int& chash_blk, int& status)
{
    const int lane = w.lane();
    long long want = 2ll * wk.closed_cap;
    if (want > wk.closed_max) want = wk.closed_max;
    int hc = 1; while (hc < 2 * want) hc <<= 1;
    int k1 = pp_arena_class((unsigned long long)want * sizeof(PPClosed3));
    int k2 = pp_arena_class((unsigned long long)hc * sizeof(PPHashSlot));
    PPClosed3* nc = nullptr; PPHashSlot* nh = nullptr;
    if (lane == 0)
    {
        nc = (PPClosed3*)pp_arena_alloc_patient(wk.arena, k1, k1);
        nh = nc ? (PPHashSlot*)pp_arena_alloc_patient(wk.arena, k2, k2) : nullptr;
        if (nc && !nh) { pp_arena_free(wk.arena, nc, k1); nc = nullptr; }
        if (!nc) status |= PP_STATUS_ARENA_EXHAUSTED;
    }
    nc = pp_bcast_ptr(w, nc);
    nh = pp_bcast_ptr(w, nh);
    if (!nc) return false;
    pp_coop_copy8(w, nc, wk.closed, (size_t)n_closed * sizeof(PPClosed3));
    for (int c = lane; c < hc; c += W::LANES) { PPHashSlot e; e.key = 0xffffffffu; e.idx = -1; nh[c] = e; }
    w.sync();
    const unsigned mask = (unsigned)hc - 1u;
    for (int c = lane; c < n_closed; c += W::LANES)
    {
        const unsigned key = nc[c].key;
        unsigned h = pp_hash_key(key) & mask;
        while (!pp_hash_slot_claim(&nh[h], key, c)) h = (h + 1) & mask;
    }
    w.sync();
    if (lane == 0)
    {
        if (closed_blk >= 0) pp_arena_free(wk.arena, wk.closed, closed_blk);
        if (chash_blk >= 0) pp_arena_free(wk.arena, wk.chash, chash_blk);
    }
    closed_blk = w.shfl(k1, 0); chash_blk = w.shfl(k2, 0);        // the classes actually handed out (lane 0 knows)
    wk.closed = nc; wk.closed_cap = (int)want; wk.chash = nh; wk.chash_cap = hc;
    w.sync();
    return true;
}

// closed-set lookup: index into the closed log or -1
PP_HD int pp_closed_find(const PPWork& wk, unsigned key)
{
    unsigned mask = (unsigned)wk.chash_cap - 1u;
    unsigned h = pp_hash_key(key) & mask;
    for (;;)
    {
        const PPHashSlot s = wk.chash[h];      // one 8-byte load answers both "empty?" and "same key?"
        if (s.idx < 0) return -1;
        if (s.key == key) return s.idx;
        h = (h + 1) & mask;
    }
}

PP_HD void pp_closed_link(PPWork& wk, unsigned key, int idx)
{
    unsigned mask = (unsigned)wk.chash_cap - 1u;
    unsigned h = pp_hash_key(key) & mask;
    while (wk.chash[h].idx >= 0) h = (h + 1) & mask;
    PPHashSlot s; s.key = key; s.idx = idx;
    wk.chash[h] = s;
}

// links a new node under parent p; returns false when the pool is exhausted
PP_HD_NOINLINE_FN bool pp_open3_attach(PPSearchState& S, const PPSucc& s, unsigned key, float f, int prev, int p, bool left)
{
    int slot = S.open.alloc();
    if (slot == PP_RB_NIL) { S.status |= PP_STATUS_OPEN_OVERFLOW; return false; }
    PPNode3& n = S.open.n[slot];
    n.w.f = f; n.w.key = key; n.g = s.g; n.x = s.x; n.y = s.y; n.heading = s.heading;
    n.vmin_sqr = s.vmin_sqr; n.curv = s.curv; n.prev = prev; n.bin = s.bin;
    S.open.insert_and_rebalance(left, slot, p);
    if (S.open.count > S.max_open) S.max_open = S.open.count;
    return true;
}

PP_HD_NOINLINE_FN bool pp_open3_insert(PPSearchState& S, const PPSucc& s, unsigned key, float f, int prev)
{
    PPKey k; k.key = key; k.f = f;
    int p; bool left;
    if (!S.open.insert_pos(k, p, left)) return true;   // equal-f drop (F5)
    return pp_open3_attach(S, s, key, f, prev, p, left);
}

// Rarely executed pieces of the search loop, out of line so that the loop body stays small (the kernel is bound by
// instruction fetch at bench occupancy).

// parent chain of closed state `c` (-1 = none) into path[at ...]: reconstruct_path's second half (HybridAStar.cpp:238-256);
// returns the number of chain points (counted even when they no longer fit)
PP_HD_NOINLINE_FN int pp_chain_to_path(const PPConsts& C, PPWork& wk, int c, int at, int& status)
{
    int n_chain = 0;
    while (c >= 0)
    {
        const int pos = at + n_chain;
        if (pos < wk.path_cap)
        {
            PPPathPt& p = wk.path[pos];
            p.x = wk.closed[c].x; p.y = wk.closed[c].y; p.heading = wk.closed[c].heading;
            p.curvature = C.abs_curv[wk.closed[c].curv];
        }
        else status |= PP_STATUS_PATH_OVERFLOW;
        n_chain++;
        c = wk.closed[c].prev;
    }
    return n_chain;
}

// The analytic expansion from closed state cn (HybridAStar.cpp:129-149): Dubins::get_shortest_path (Dubins.cpp:125-153) sampled by
// the lanes into path[0 .. total) + Grid3D::check_path (Grid3D.cpp:78-93).  True = shot accepted (all lanes agree).
template <class W>
PP_HD_NOINLINE_FN bool pp_dubins_shot(const W& w, const PPConsts& C, const PPGroup& G, PPWork& wk, const PPClosed3& cn,
                                      int& total, float& len, int& status)
{
    const int lane = w.lane();
    const PPFrame& F = G.frame;
    int type; float p[4]; PPDubinsCenters cen; PPDubinsPlan pl;
    len = pp_dubins_shortest(C.r_min, cn.x, cn.y, cn.heading, F.goal_x, F.goal_y, F.goal_h, type, p, cen);
    if (fabsf(p[1]) > (float)PP_PI_2) return false;      // first arc longer than 90 degrees: Dubins.cpp:152, HybridAStar.cpp:135
    pp_dubins_plan(C.r_min, C.step, C.ang_step, type, p, cen, pl);
    total = pl.size_3 + 1;
    bool blocked = false, overflow = false;
    float acc = p[0];
    for (int k = 0; k < total; k++)
    {
        if (k == pl.size_1) acc = 0.0f;       // straight segment: dist accumulator
        if (k == pl.size_2) acc = p[2];       // goal arc: theta accumulator
        if ((k % W::LANES) == lane)
        {
            float x, y, h, kappa;
            pp_dubins_sample(pl, C.r_min, k, acc, x, y, h, kappa);
            if (pp_path_point_blocked(C, G.map, x, y)) blocked = true;
            if (k < wk.path_cap) { PPPathPt& q = wk.path[k]; q.x = x; q.y = y; q.heading = h; q.curvature = kappa; }
            else overflow = true;
        }
        if (k < pl.size_1) acc = (pl.s1 < 0) ? acc - C.ang_step : acc + C.ang_step;
        else if (k < pl.size_2) acc = acc + C.step;
        else if (k < pl.size_3) acc = (pl.s2 < 0) ? acc - C.ang_step : acc + C.ang_step;
    }
    const bool ok = (w.ballot(blocked) == 0u);
    if (ok && w.ballot(overflow) != 0u) status |= PP_STATUS_PATH_OVERFLOW;
    return ok;
}

// The search.  All lanes of the warp call it with identical arguments.
template <class W>
PP_HD_NOINLINE_FN void pp_search_exact(const W& w, const PPConsts& C, const float* off_xy, const PPGroup& G,
                                    const PPState& start, PPWork& wk, PPSmem& sm, PPResult& res)
{
    const int lane = w.lane();
    const int N = C.N;
    const PPFrame& F = G.frame;
    const unsigned kb = (unsigned)(C.bins + 1);

    PP_PROF_DECL
    PP_PROF_CDECL
    // the slot's fixed closed log / hash: wk follows these two containers as they grow and gets them back at the end
    PPClosed3* const closed0 = wk.closed; PPHashSlot* const chash0 = wk.chash;
    const int closed_cap0 = wk.closed_cap, chash_cap0 = wk.chash_cap;
    int closed_blk = -1, chash_blk = -1;
    // ---- scratch init (all lanes) ----
    const bool carry = (wk.lazy_sid != nullptr);
    unsigned sid0 = 0u;
    if (!carry) { for (int c = lane; c < N * N; c += W::LANES) wk.cell_state[c] = 0u; }
    else
    {
        sid0 = *wk.lazy_sid;
        if (sid0 > (PP_CS_STAMP >> 1))      // the 30-bit search id is half used up: drop all closed stamps, restart at 0
        {
            for (int c = lane; c < N * N; c += W::LANES) wk.cell_state[c] &= ~PP_CS_STAMP;
            sid0 = 0u;
        }
        w.sync();
        if (lane == 0)
        {
            // Grid3D::set_start_node (Grid3D.cpp:146, :154): soft_reset of the start cell's map node -> g = 0, f = h;
            // its visited flag is NOT touched
            int c0 = start.ci * N + start.cj;
            wk.nm_g[c0] = 0.0f; wk.nm_f[c0] = pp_h2d(C, start.ci, start.cj);
            wk.cell_state[c0] |= PP_CS_TOUCHED;
        }
    }
    for (int c = lane; c < wk.chash_cap; c += W::LANES) { PPHashSlot e; e.key = 0xffffffffu; e.idx = -1; wk.chash[c] = e; }
    w.sync();
    PP_PROF_MARK(0)

    PPSearchState S;
    int shot_counter = 0, shot_interval = C.shot_interval;
    bool shot_allowed = false;
    int n_pops = 0, n_oob = 0;
    int n_chain = 0, n_dubins = 0, success = 0;
    float cost = FLT_MAX;

    if (lane == 0)
    {
        S.open.init(wk.open3, wk.open3_cap);
#if PP_EXACT_SPEC
        S.open.mlog = sm.mlog; S.open.mcnt = &sm.mcount; S.open.mcap = PP_MLOG_CAP;
        sm.mcount = 0;
#endif
        S.lazy.open.init(wk.open2, wk.open2_cap);
        S.lazy.search_id = sid0; S.lazy.status = 0; S.lazy.n_searches = 0; S.lazy.n_pops = 0;
        S.lazy.arena = wk.arena; S.lazy.max_cap = wk.arena ? wk.open2_max : wk.open2_cap; S.lazy.blk = -1;
        S.n_closed = 0; S.status = 0; S.max_open = 0; S.open_blk = -1;
        // _open_set.insert(start_node), HybridAStar.cpp:103
        PPSucc s0;
        s0.x = start.x; s0.y = start.y; s0.heading = start.heading; s0.g = start.g;
        s0.vmin_sqr = start.vmin_sqr; s0.curv = start.curv; s0.bin = start.bin; s0.ci = start.ci; s0.cj = start.cj; s0.ok = 1;
        unsigned key0 = (unsigned)(start.ci * N + start.cj) * kb + (unsigned)start.bin;
        pp_open3_insert(S, s0, key0, start.f, -1);
    }

    const int n_succ_max = 2 * C.A + 1;
    enum { ACT_EXPAND = 0, ACT_FAIL = 1, ACT_GOAL = 2, ACT_SHOT = 3, ACT_ABORT = 4 };

    // goal circle centres (Dubins.cpp:29-33): the same for every heuristic evaluation of this query
    PPDubinsGoal gc;
    {
        float sg, cg;
        pp_sincosf(F.goal_h, sg, cg);
        gc.grx = F.goal_x + C.r_min * sg; gc.gry = F.goal_y - C.r_min * cg;
        gc.glx = F.goal_x - C.r_min * sg; gc.gly = F.goal_y + C.r_min * cg;
    }

    for (;;)
    {
        // ---------------- phase 0: make room (all lanes; only with an arena) ----------------
        if (wk.arena)
        {
            int grow = 0, ncl = 0;
            if (lane == 0)
            {
                if (S.open.next + n_succ_max + 2 > S.open.cap && S.open.cap < wk.open3_max) grow |= 1;
                if (S.n_closed + 1 > wk.closed_cap && wk.closed_cap < wk.closed_max) grow |= 2;
                ncl = S.n_closed;
            }
            grow = w.shfl(grow, 0);
            if (grow & 1) pp_grow_open3(w, wk, S);
            if (grow & 2)
            {
                ncl = w.shfl(ncl, 0);
                int st = 0;
                pp_grow_closed(w, wk, ncl, closed_blk, chash_blk, st);
                if (lane == 0) S.status |= st;
            }
        }
        // ---------------- phase 1: pop (control lane) ----------------
        int action = ACT_EXPAND, cur = -1;
        if (lane == 0)
        {
            if (S.open.empty()) action = ACT_FAIL;
            else
            {
                int it = S.open.begin();
                const PPNode3& n = S.open.n[it];
                cur = pp_closed_find(wk, n.w.key);     // _closed_set.insert(*it).first
                if (cur < 0)
                {
                    if (S.n_closed >= wk.closed_cap) { S.status |= PP_STATUS_CLOSED_OVERFLOW; action = ACT_ABORT; }
                    else
                    {
                        cur = S.n_closed++;
                        PPClosed3& c = wk.closed[cur];
                        c.x = n.x; c.y = n.y; c.heading = n.heading; c.g = n.g; c.f = n.w.f; c.vmin_sqr = n.vmin_sqr;
                        c.curv = n.curv; c.bin = n.bin; c.key = n.w.key; c.prev = n.prev;
                        pp_closed_link(wk, n.w.key, cur);
                    }
                }
                if (action != ACT_ABORT)
                {
                    S.open.erase(it);
                    unsigned cell = wk.closed[cur].key / kb;
                    if (cell == (unsigned)(F.goal_ci * N + F.goal_cj)) action = ACT_GOAL;   // cell equality only (F6)
                    else if (shot_allowed)
                    {
                        shot_counter++;
                        if (shot_counter == shot_interval) action = ACT_SHOT;
                    }
                }
            }
        }
        action = w.shfl(action, 0);
        cur = w.shfl(cur, 0);
        w.sync();
        PP_PROF_MARK(1)
        if (action == ACT_FAIL || action == ACT_ABORT) break;

        // popped node, read by every lane
        const PPClosed3 cn = wk.closed[cur];

        if (action == ACT_GOAL)
        {
            success = 1; cost = cn.g;
            // reconstruct_path from _terminal_node = *it_first (HybridAStar.cpp:118, :238-256)
            if (lane == 0) { int st = 0; n_chain = pp_chain_to_path(C, wk, cur, 0, st); S.status |= st; }
            break;
        }

        if (action == ACT_SHOT)
        {
            // ---------------- Dubins shot (all lanes), HybridAStar.cpp:129-149 ----------------
            int st = 0, total = 0;
            float len = 0.0f;
            if (pp_dubins_shot(w, C, G, wk, cn, total, len, st))
            {
                success = 1; cost = cn.g + len; n_dubins = total;
                if (lane == 0)
                {
                    // _terminal_node = *(it_first->_prev), HybridAStar.cpp:137
                    if (cn.prev < 0) st |= PP_STATUS_NULL_TERMINAL;
                    n_chain = pp_chain_to_path(C, wk, cn.prev, n_dubins, st);
                    S.status |= st;
                }
                break;
            }
            shot_counter = 0;
            int ni = shot_interval - C.shot_decay;
            shot_interval = (ni < 50) ? 50 : ni;
        }

        // ---------------- phase 3: expansion (all lanes), Grid3D.cpp:47-74 ----------------
        n_pops++;
        if (cn.bin >= C.bins) n_oob++;
        if (lane == 0 && wk.trace && (n_pops - 1) < wk.trace_cap)
        {
            PPPop& t = wk.trace[n_pops - 1];
            unsigned cell = cn.key / kb;
            t.ci = (int)(cell / (unsigned)N); t.cj = (int)(cell % (unsigned)N); t.bin = cn.bin;
            t.x = cn.x; t.y = cn.y; t.heading = cn.heading; t.g = cn.g; t.f = cn.f;
        }
        shot_allowed = (cn.vmin_sqr < 1.0f);     // neglect_acceleration, VehicleModel.cpp:76
        pp_expand_warp(w, C, off_xy, G, cn.x, cn.y, cn.heading, cn.g, cn.vmin_sqr, cn.curv, cn.bin, sm);
        PP_PROF_MARK(2)
        // closed-set membership of all successors at once (one probe per lane; the closed set only changes at a
        // pop, so this equals the per-successor `_closed_set.find(node)` of HybridAStar.cpp:162); members drop out
#pragma unroll 1
        for (int s = lane; s < n_succ_max; s += W::LANES)
            if (sm.succ[s].ok)
            {
                unsigned key = (unsigned)(sm.succ[s].ci * N + sm.succ[s].cj) * kb + (unsigned)sm.succ[s].bin;
                if (pp_closed_find(wk, key) >= 0) sm.succ[s].ok = 0;
            }
        w.sync();
        PP_PROF_MARK(4)
        // Dubins candidates of every surviving successor, transcendentals spread over the warp in stages
        pp_dubins_h2_warp(w, C, F, gc, sm);
        PP_PROF_MARK(3)

#if PP_EXACT_SPEC
        // ---------------- phase 4a: every tree walk of this expansion, speculatively and in parallel (all lanes) ----------------
        // walk 2s = open.find of successor s (probe f = g + field), walk 2s + 1 = its insert position (f = g + max(h1, h2)), the
        // latter only when h1 is already cached (the cell is _visted: its cost can no longer change, AStar.cpp:100-105)
        PPRbPool pool; pool.base = (char*)pp_bcast_ptr(w, (lane == 0) ? S.open.n : (PPNode3*)0); pool.stride = (int)sizeof(PPNode3);
        if (lane == 0) sm.mcount = 0;
        for (int q = lane; q < 2 * n_succ_max; q += W::LANES)
        {
            const int s = q >> 1;
            sm.spec_np[q] = -1;
            if (!sm.succ[s].ok) continue;
            const PPSucc& sc = sm.succ[s];
            PPKey k; k.key = (unsigned)(sc.ci * N + sc.cj) * kb + (unsigned)sc.bin;
            int np = 0;
            if ((q & 1) == 0)
            {
                k.f = sc.g;
                sm.spec_a[q] = pp_rb_find_walk(pool, k, sm.spec_path[q], PP_SPEC_PATH, np);
                sm.spec_np[q] = np;
            }
            else
            {
                const int cell = sc.ci * N + sc.cj;
                if (wk.cell_state[cell] & PP_CS_VISITED)
                {
                    const float h1 = wk.nm_f[cell];
                    float h2 = sm.cand[4 * s];
                    for (int t = 1; t < 4; t++) if (sm.cand[4 * s + t] < h2) h2 = sm.cand[4 * s + t];
                    k.f = sc.g + ((h1 < h2) ? h2 : h1);
                    int p; bool left;
                    const bool ins = pp_rb_insert_pos_walk(pool, k, p, left, sm.spec_path[q], PP_SPEC_PATH, np);
                    sm.spec_a[q] = p; sm.spec_b[q] = ins ? (left ? 1 : 0) : -1; sm.spec_f[q] = k.f;
                    sm.spec_np[q] = np;
                }
            }
        }
        w.sync();
        PP_PROF_MARK(5)

        // ---------------- phase 4b: successors into the containers, in the reference's order (HybridAStar.cpp:159-193) ----------------
        // The control lane commits; before each use of a speculative walk all lanes check it against the nodes the earlier
        // commits of this expansion touched (pp_spec_valid) and the control lane walks again only when they meet.
        int abort = 0;
        for (int s = 0; s < n_succ_max; s++)
        {
            if (!sm.succ[s].ok) continue;            // uniform: shared memory
            const PPSucc& sc = sm.succ[s];
            const unsigned key = (unsigned)(sc.ci * N + sc.cj) * kb + (unsigned)sc.bin;
            const bool v_find = pp_spec_valid(w, sm.spec_path[2 * s], sm.spec_np[2 * s], sm.mlog, sm.mcount);
            int todo = 0;                            // bit 0: insert, bit 1: h1 still has to come from the lazy A*
            if (lane == 0)
            {
                PPKey k; k.key = key; k.f = sc.g;                           // f == g + field at find time
                int it_node = v_find ? sm.spec_a[2 * s] : S.open.find(k);
                PP_PROF_COUNT(v_find ? 0 : 1, 1) PP_PROF_COUNT(5, 1)
                if (it_node == PP_RB_NIL) todo = 1;
                else if (sc.g < S.open.n[it_node].g) { S.open.erase(it_node); todo = 1; }
                if (todo && sm.spec_np[2 * s + 1] < 0) todo |= 2;
            }
            todo = w.shfl(todo, 0);
            PP_PROF_MARK(5)
            if (!todo) continue;
            float f = 0.0f;
            if (todo & 2)
            {
                if (lane == 0)
                {
                    float h1 = pp_lazy_astar(C, G.map, F, wk, S.lazy, sc.ci, sc.cj, sm.lazy_nb);
                    float h2 = sm.cand[4 * s];
                    for (int t = 1; t < 4; t++) if (sm.cand[4 * s + t] < h2) h2 = sm.cand[4 * s + t];
                    f = sc.g + ((h1 < h2) ? h2 : h1);
                }
                PP_PROF_MARK(6)
            }
            w.sync();                                // sm.mcount after a possible erase above
            const bool v_ins = pp_spec_valid(w, sm.spec_path[2 * s + 1], (todo & 2) ? -1 : sm.spec_np[2 * s + 1], sm.mlog, sm.mcount);
            if (lane == 0)
            {
                int p = 0; bool left = false, ins;
                if (!(todo & 2)) f = sm.spec_f[2 * s + 1];
                if (v_ins) { p = sm.spec_a[2 * s + 1]; ins = sm.spec_b[2 * s + 1] >= 0; left = sm.spec_b[2 * s + 1] == 1; }
                else { PPKey k; k.key = key; k.f = f; ins = S.open.insert_pos(k, p, left); }
                PP_PROF_COUNT(v_ins ? 2 : ((todo & 2) ? 4 : 3), 1) PP_PROF_COUNT(6, ins ? 1 : 0) PP_PROF_COUNT(7, sm.mcount)
                if (ins && !pp_open3_attach(S, sc, key, f, cur, p, left)) abort = 1;     // !ins: equal-f drop (F5)
                if (S.lazy.status) { S.status |= S.lazy.status; abort = 1; }
            }
            abort = w.shfl(abort, 0);
            w.sync();                                // sm.mlog / sm.mcount of this commit
            PP_PROF_MARK(7)
            if (abort) break;
        }
#else
        // ---------------- phase 4: successors into the containers (control lane), HybridAStar.cpp:159-193 ----------------
        int abort = 0;
        if (lane == 0)
        {
#pragma unroll 1
            for (int s = 0; s < n_succ_max && !abort; s++)
            {
                const PPSucc& sc = sm.succ[s];
                if (!sc.ok) continue;
                unsigned key = (unsigned)(sc.ci * N + sc.cj) * kb + (unsigned)sc.bin;
                PPKey k; k.key = key; k.f = sc.g;                           // f == g + field at find time
                int it_node = S.open.find(k);
                bool do_insert = false;
                if (it_node == PP_RB_NIL) do_insert = true;
                else if (sc.g < S.open.n[it_node].g) { S.open.erase(it_node); do_insert = true; }
                PP_PROF_MARK(5)
                if (do_insert)
                {
                    float h1 = pp_lazy_astar(C, G.map, F, wk, S.lazy, sc.ci, sc.cj, sm.lazy_nb);
                    PP_PROF_MARK(6)
                    float h2 = sm.cand[4 * s];
                    for (int t = 1; t < 4; t++) if (sm.cand[4 * s + t] < h2) h2 = sm.cand[4 * s + t];
                    float f = sc.g + ((h1 < h2) ? h2 : h1);
                    if (!pp_open3_insert(S, sc, key, f, cur)) abort = 1;
                    if (S.lazy.status) { S.status |= S.lazy.status; abort = 1; }
                    PP_PROF_MARK(7)
                }
            }
        }
        abort = w.shfl(abort, 0);
        w.sync();
#endif
        if (abort) break;
    }
    PP_PROF_FLUSH

    if (lane == 0)
    {
        if (wk.arena)
        {
            if (S.open_blk >= 0) pp_arena_free(wk.arena, S.open.n, S.open_blk);
            if (S.lazy.blk >= 0) pp_arena_free(wk.arena, S.lazy.open.n, S.lazy.blk);
            if (closed_blk >= 0) pp_arena_free(wk.arena, wk.closed, closed_blk);
            if (chash_blk >= 0) pp_arena_free(wk.arena, wk.chash, chash_blk);
        }
        if (carry) *wk.lazy_sid = S.lazy.search_id;
        res.success = success;
        res.status = S.status;
        res.cost = cost;
        res.n_pops = n_pops;
        res.n_pops_bin_oob = n_oob;
        res.n_chain = n_chain;
        res.n_dubins = n_dubins;
        res.n_lazy_searches = S.lazy.n_searches;
        res.n_lazy_pops = S.lazy.n_pops;
        res.max_open = S.max_open;
        res.n_closed = S.n_closed;
        res.pad = 0;
    }
    wk.closed = closed0; wk.closed_cap = closed_cap0; wk.chash = chash0; wk.chash_cap = chash_cap0;
}

#endif
