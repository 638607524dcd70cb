// K-POP search mode: up to k <= 32 nodes popped and expanded per iteration by one CTA (W::LANES cooperating lanes: 4 warps
// on the device, 1 lane in the host emulation of tests/cpp), on a CTA-parallel priority queue, with the exact 2D
// distance field as the holonomic heuristic.
//
// This is NEW semantics (north_star: "optionally k pops per iteration within one query"; SURVEY.md §7 mode
// definitions): the reference's equal-f drops (F5) and its order-dependent lazy 2D A* (F4) are replaced by clean rules
// that parallelise, so results differ from the reference (typically a few % lower cost with fewer expansions).
// The rules are stated in oracle/port/kpop.inc (CPU restatement); this file must reproduce that restatement bit
// for bit -- pop sequence, cost, path.  Summary:
//   open list = total order (f, key, idx), lazy deletion, best-g-wins per key; an iteration takes the min(k, |open|)
//   smallest entries, the valid ones are the pops (rank order), all closed at once; goal / Dubins shot decided in rank
//   order; every pop expands every steering primitive of its window -> candidates c = rank*(2A+1) + a; per key the
//   smallest pack(g, order = iteration*1024 + c) wins, also against the key's recorded best; winners get node indices in
//   c order and f = g + max(h1[cell], Dubins).  The heuristic Dubins length and the APF term are evaluated with the
//   fixed-order FP32 functions of pp_fmath.h, and so is the Dubins shot (centres, candidates, plan, samples).
//
// Data structures (per query slot, global memory), all driven by the CTA's lanes together:
//   nodes[]   append-only log of generated nodes (parent links are log indices; every node remembers its key's hash slot)
//   table[]   open-addressing hash (key -> best pack, best node, closed bit), all-ones = empty, cleaned by the query that
//             filled it (cost proportional to the nodes generated, not to the capacity); candidates race with atomicMin(pack):
//             the winner is the minimum, independent of thread order => deterministic
//   LSM queue sorted runs in levels of capacity 256 << level (log-structured merge): a batch of new entries is
//             rank-sorted in shared memory and merged down the levels with merge-path merges (the carry settles in the
//             first level that can hold it together with that level's run); the k smallest are found by ranking the first
//             k entries of every non-empty run against each other (binary searches in shared memory)
// Decisions that need a ballot over the <= 32 popped entries (validity, ranks, shot counter) are taken by ballot group 0
// and broadcast through shared memory; prefix counts over all lanes go through W::scan_count.
#ifndef PP_KPOP_H
#define PP_KPOP_H

#include "pp_search.h"
#include "pp_fmath.h"

#define PP_K_MAXPOP 32
#define PP_K_MAXSUCC 8                                  // (2A+1) supported by this mode
#define PP_K_MAXCAND (PP_K_MAXPOP * PP_K_MAXSUCC)        // 256
#define PP_K_LEVELS 16                                   // LSM levels: capacity 256 << level
#define PP_K_RUN0 256
#define PP_K_NONE 0xffffffffu                           // node word of a key without a node yet (= the all-ones empty state)
#define PP_K_OPEN 0x80000000u                           // node word = index | PP_K_OPEN while the key is open, index alone once closed
#define PP_K_EMPTY 0xffffffffu

struct PPKEntry { float f; unsigned key; unsigned idx; unsigned pad; };               // 16 B queue entry
struct PPKNode { float x, y, heading, g, v2, f; int curv, bin, cell, parent; unsigned key; int slot; };   // 48 B; slot = its key's hash slot
struct PPKSlot { unsigned key; unsigned node; unsigned long long pack; };             // 16 B hash slot
struct PPKCand { float x, y, heading, g, v2; int curv_bin_ok; int cell; int slot; unsigned long long pack; };   // 40 B

struct PPKWork
{
    PPKNode*  nodes;   int nodes_cap;
    PPKSlot*  table;   int table_cap;          // power of two
    PPKEntry* arena;                           // level l at offset PP_K_RUN0 * ((1 << l) - 1), PP_K_LEVELS levels used up to lsm_levels
    PPKEntry* l0;                              // level 0 of the queue: PPKSmem::l0 (set by pp_search_kpop)
    PPKEntry* tmp_a;   PPKEntry* tmp_b;        // merge scratch, each nodes_cap + PP_K_RUN0 entries
    int       lsm_levels;
    const float* h1;                           // N*N exact 2D distance field of the query's group
    PPPathPt* path;    int path_cap;
    PPPop*    trace;   int trace_cap;
};

struct PPKSmem
{
    union
    {
        PPKCand  cand[PP_K_MAXCAND];                       // expansion
        PPKEntry sel[PP_K_LEVELS * PP_K_MAXPOP];           // pop selection: first k entries of every run
        struct { PPKEntry batch[PP_K_MAXCAND]; PPKEntry sorted[PP_K_MAXCAND]; } q;   // new queue entries of this iteration
    } u;
    PPKEntry l0[PP_K_RUN0];                                // level 0 of the LSM queue
    PPKEntry popped[PP_K_MAXPOP];
    PPKNode  parents[PP_K_MAXPOP];                         // copies of this iteration's pops, by rank
    int      pop_idx[PP_K_MAXPOP];                         // their node indices
    int      head[PP_K_LEVELS], size[PP_K_LEVELS];         // live range [head, size) of every level's run
    int      taken[PP_K_LEVELS];
    unsigned short wlist[PP_K_MAXCAND];                    // candidate index of the t-th winner
    float    shot[4][5];                                   // Dubins shot: (length, p[0..3]) of the four candidates
    int      bc[8];                                        // values decided by ballot group 0, broadcast to the CTA
    int      scan[32];                                     // per-warp counts of W::scan_count
};

PP_HD bool pp_kless(const PPKEntry& a, const PPKEntry& b)
{
    if (a.f != b.f) return a.f < b.f;
    if (a.key != b.key) return a.key < b.key;
    return a.idx < b.idx;
}
PP_HD unsigned pp_fbits(float v)
{
    union { float f; unsigned u; } c; c.f = v; return c.u;
}
PP_HD PPKEntry pp_kinf()
{
    PPKEntry e; e.f = INFINITY; e.key = 0xffffffffu; e.idx = 0xffffffffu; e.pad = 0; return e;
}

// ---- atomics: device / plain (single host lane) / GCC builtins (PP_HOST_ATOMICS: the multi-threaded host emulation of
// tests/cpp/search_mt.cpp, which runs this file under ThreadSanitizer) ----------------------------------------------------
PP_HD unsigned pp_atomic_cas_u32(unsigned* p, unsigned expect, unsigned val)
{
#ifdef __CUDA_ARCH__
    return atomicCAS(p, expect, val);
#elif defined(PP_HOST_ATOMICS)
    __atomic_compare_exchange_n(p, &expect, val, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return expect;                                   // the old value in either case
#else
    unsigned old = *p; if (old == expect) *p = val; return old;
#endif
}
PP_HD void pp_atomic_min_u64(unsigned long long* p, unsigned long long val)
{
#ifdef __CUDA_ARCH__
    atomicMin(p, val);
#elif defined(PP_HOST_ATOMICS)
    unsigned long long cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (val < cur && !__atomic_compare_exchange_n(p, &cur, val, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
#else
    if (val < *p) *p = val;
#endif
}
PP_HD void pp_atomic_inc_i32(int* p)
{
#ifdef __CUDA_ARCH__
    atomicAdd(p, 1);
#elif defined(PP_HOST_ATOMICS)
    __atomic_fetch_add(p, 1, __ATOMIC_RELAXED);
#else
    (*p)++;
#endif
}
PP_HD void pp_fence()
{
#ifdef __CUDA_ARCH__
    __threadfence_block();
#endif
}

// ---- hash table -----------------------------------------------------------------------------------------------------
// (key, node) of a slot with one 8-byte load: they share the slot's first half
PP_HD void pp_kslot_load(const PPKSlot* s, unsigned& key, unsigned& node)
{
#ifdef __CUDA_ARCH__
    const uint2 v = *reinterpret_cast<const uint2*>(s);
    key = v.x; node = v.y;
#elif defined(PP_HOST_ATOMICS)
    key = __atomic_load_n(&s->key, __ATOMIC_RELAXED); node = s->node;   // the key may be CAS-ed by another lane right now
#else
    key = s->key; node = s->node;
#endif
}
PP_HD void pp_kslot_clear(PPKSlot* s)
{
#if !defined(__CUDA_ARCH__) && defined(PP_HOST_ATOMICS)
    __atomic_store_n(&s->key, PP_K_EMPTY, __ATOMIC_RELAXED); __atomic_store_n(&s->node, PP_K_NONE, __ATOMIC_RELAXED);
    __atomic_store_n(&s->pack, ~0ull, __ATOMIC_RELAXED);
#else
    PPKSlot e; e.key = PP_K_EMPTY; e.node = PP_K_NONE; e.pack = ~0ull;
    *s = e;
#endif
}
PP_HD int pp_ktable_find(const PPKWork& wk, unsigned key, unsigned& node)
{
    unsigned mask = (unsigned)wk.table_cap - 1u, h = pp_hash_key(key) & mask;
    for (;;)
    {
        unsigned k;
        pp_kslot_load(wk.table + h, k, node);
        if (k == key) return (int)h;
        if (k == PP_K_EMPTY) return -1;
        h = (h + 1) & mask;
    }
}
// `node` = the slot's node word; a key inserted by this call (or concurrently, in the same expansion phase) has none yet
PP_HD int pp_ktable_find_or_insert(PPKWork& wk, unsigned key, unsigned& node)
{
    unsigned mask = (unsigned)wk.table_cap - 1u, h = pp_hash_key(key) & mask;
    for (;;)
    {
        unsigned k;
        pp_kslot_load(wk.table + h, k, node);
        if (k == key) return (int)h;
        if (k == PP_K_EMPTY)
        {
            unsigned old = pp_atomic_cas_u32(&wk.table[h].key, PP_K_EMPTY, key);
            if (old == PP_K_EMPTY || old == key) { node = PP_K_NONE; return (int)h; }
        }
        h = (h + 1) & mask;
    }
}

// ---- sort of n <= PP_K_MAXCAND entries, shared memory src -> dst: every lane ranks its entries against all others (the
// order is total, so the ranks are a permutation); no barrier inside, unlike a sorting network ---------------------------
template <class W>
PP_HD void pp_kranksort(const W& w, const PPKEntry* src, PPKEntry* dst, int n)
{
    for (int t = w.lane(); t < n; t += W::LANES)
    {
        const PPKEntry e = src[t];
        int rank = 0;
        for (int j = 0; j < n; j++) rank += pp_kless(src[j], e) ? 1 : 0;
        dst[rank] = e;
    }
    w.sync();
}

// ---- warp merge of two sorted runs (merge path): dst[0 .. na+nb) ------------------------------------------------------
template <class W>
PP_HD void pp_kmerge(const W& w, const PPKEntry* A, int na, const PPKEntry* B, int nb, PPKEntry* dst)
{
    const int total = na + nb;
    const int seg = (total + W::LANES - 1) / W::LANES;
    int d0 = w.lane() * seg; if (d0 > total) d0 = total;
    int d1 = d0 + seg; if (d1 > total) d1 = total;
    // number of A elements among the first d0 outputs
    int lo = d0 - nb; if (lo < 0) lo = 0;
    int hi = d0 < na ? d0 : na;
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (pp_kless(B[d0 - 1 - mid], A[mid])) hi = mid; else lo = mid + 1;
    }
    int i = lo, j = d0 - lo;
    for (int o = d0; o < d1; o++)
    {
        bool take_a = (j >= nb) || (i < na && !pp_kless(B[j], A[i]));
        dst[o] = take_a ? A[i] : B[j];
        if (take_a) i++; else j++;
    }
    w.sync();
}

template <class W>
PP_HD void pp_kcopy(const W& w, const PPKEntry* src, PPKEntry* dst, int n)
{
    for (int t = w.lane(); t < n; t += W::LANES) dst[t] = src[t];
    w.sync();
}

// run of level l: level 0 (<= 256 entries, touched by almost every iteration) lives in shared memory, the rest in the arena
PP_HD PPKEntry* pp_klevel(const PPKWork& wk, int l) { return l == 0 ? wk.l0 : wk.arena + (size_t)PP_K_RUN0 * (((size_t)1 << l) - 1); }

// insert a sorted batch (shared memory, m <= PP_K_MAXCAND entries); false = queue capacity exhausted.
// The carry walks down the levels: it settles in the first level that can hold it together with that level's live run
// (merged), otherwise it absorbs the run and moves on.  carry <= capacity(l) holds at every level, so the queue only
// overflows when the live entries exceed the top level's capacity (sized >= nodes_cap by the host).
template <class W>
PP_HD bool pp_klsm_insert(const W& w, PPKWork& wk, PPKSmem& sm, const PPKEntry* batch, int m)
{
    const PPKEntry* carry = batch;
    int carry_n = m;
    PPKEntry* t0 = wk.tmp_a; PPKEntry* t1 = wk.tmp_b;
    for (int l = 0; l < wk.lsm_levels; l++)
    {
        const int cnt = sm.size[l] - sm.head[l];
        const int cap = PP_K_RUN0 << l;
        PPKEntry* L = pp_klevel(wk, l);
        if (cnt == 0)
        {
            pp_kcopy(w, carry, L, carry_n);
            if (w.lane() == 0) { sm.head[l] = 0; sm.size[l] = carry_n; }
            w.sync();
            return true;
        }
        // level 0 settles without leaving shared memory: the merge target is the (now free) unsorted-batch area
        PPKEntry* dst = (l == 0 && cnt + carry_n <= cap && carry != sm.u.q.batch) ? sm.u.q.batch : t0;
        pp_kmerge(w, L + sm.head[l], cnt, carry, carry_n, dst);
        carry_n += cnt;
        if (carry_n <= cap)
        {
            pp_kcopy(w, dst, L, carry_n);
            if (w.lane() == 0) { sm.head[l] = 0; sm.size[l] = carry_n; }
            w.sync();
            return true;
        }
        if (w.lane() == 0) { sm.head[l] = 0; sm.size[l] = 0; }
        w.sync();
        carry = t0;
        PPKEntry* t = t0; t0 = t1; t1 = t;
    }
    return false;
}

// the min(k, total) smallest entries -> sm.popped[0 .. n) in ascending order; returns n
template <class W>
PP_HD int pp_klsm_pop(const W& w, PPKWork& wk, PPKSmem& sm, int k)
{
    const int lane = w.lane();
    const int nl = wk.lsm_levels;
    // compact list of the non-empty runs (identical on every lane): usually 3-5 of the levels hold a run
    int lv[PP_K_LEVELS];
    int nr = 0;
    for (int l = 0; l < nl; l++) if (sm.size[l] > sm.head[l]) lv[nr++] = l;
    for (int t = lane; t < PP_K_LEVELS; t += W::LANES) sm.taken[t] = 0;
    for (int t = lane; t < PP_K_MAXPOP; t += W::LANES) sm.popped[t] = pp_kinf();
    if (nr == 0) { w.sync(); return 0; }
    // stage the first k entries of every non-empty run (padded with +inf)
    for (int t = lane; t < nr * PP_K_MAXPOP; t += W::LANES)
    {
        int c = t / PP_K_MAXPOP, p = t - c * PP_K_MAXPOP, l = lv[c];
        int at = sm.head[l] + p;
        sm.u.sel[t] = (p < k && at < sm.size[l]) ? pp_klevel(wk, l)[at] : pp_kinf();
    }
    w.sync();
    // rank every staged entry among all staged entries; ranks < k are the result.  The entries taken from a run form
    // a prefix of it, so counting them per run gives the new heads.
    for (int t = lane; t < nr * PP_K_MAXPOP; t += W::LANES)
    {
        const PPKEntry e = sm.u.sel[t];
        if (e.idx == 0xffffffffu) continue;
        int c = t / PP_K_MAXPOP, p = t - c * PP_K_MAXPOP;
        int rank = p;
        for (int c2 = 0; c2 < nr && rank < k; c2++)
        {
            if (c2 == c) continue;
            const PPKEntry* r = sm.u.sel + c2 * PP_K_MAXPOP;
            int lo = 0, hi = PP_K_MAXPOP;                       // first position whose entry is not less than e
            while (lo < hi) { int mid = (lo + hi) >> 1; if (pp_kless(r[mid], e)) lo = mid + 1; else hi = mid; }
            rank += lo;
        }
        if (rank < k)
        {
            sm.popped[rank] = e;
            pp_atomic_inc_i32(&sm.taken[lv[c]]);
        }
    }
    w.sync();
    int n = 0;
    for (int c = 0; c < nr; c++) n += sm.taken[lv[c]];
    for (int t = lane; t < nl; t += W::LANES) sm.head[t] += sm.taken[t];
    w.sync();
    return n;
}

// Heuristic Dubins length: the sequential candidate fold of Dubins.cpp:19-69 with the float tail of the reference
// (pp_dubins_finish) and the transcendentals in FP32 (pp_fmath.h); goal circle centres precomputed the same way.
PP_HD float pp_kdubins_cand(int type, float r, float sh, float gh, float csx, float csy, float cgx, float cgy)
{
    float p[4];
    const float theta = pp_fm_atan2(cgy - csy, cgx - csx);
    float ac = 0.0f, c1 = 0.0f, s1 = 0.0f, c2 = 0.0f, s2 = 0.0f;
    if (type == PP_RSL || type == PP_LSR)
    {
        ac = pp_fm_acos(pp_dubins_acos_arg(r, csx, csy, cgx, cgy));
        const float t1 = pp_dubins_theta_t1(type, ac, theta);
        const float p2 = pp_dubins_p2(type, t1);
        pp_fm_sincos(t1, s1, c1);
        pp_fm_sincos(p2, s2, c2);
    }
    return pp_dubins_finish(type, r, sh, gh, csx, csy, cgx, cgy, theta, ac, c1, s1, c2, s2, p);
}
PP_HD float pp_kdubins(const PPConsts& C, const PPFrame& F, const PPDubinsGoal& gc, float x, float y, float h)
{
    const float r = C.r_min;
    float sn, cs;
    pp_fm_sincos(h, sn, cs);
    float srx = x + r * sn, sry = y - r * cs, slx = x - r * sn, sly = y + r * cs;
    float best = 0.0f;
    for (int type = 0; type < 4; type++)
    {
        bool s_right = (type == PP_RSR) || (type == PP_RSL), g_right = (type == PP_RSR) || (type == PP_LSR);
        float len = pp_kdubins_cand(type, r, h, F.goal_h, s_right ? srx : slx, s_right ? sry : sly,
                                    g_right ? gc.grx : gc.glx, g_right ? gc.gry : gc.gly);
        if (type == 0 || len < best) best = len;
    }
    return best;
}

// One obstacle's APF term (Grid3D.cpp:209-223) in FP32 throughout -- the K-POP flavour of pp_apf_term
PP_HD float pp_kapf_term(const PPConsts& C, float ox, float oy, float radius, float x, float y, float heading)
{
    const float dx = ox - x, dy = oy - y;
    const float dist = sqrtf(dx * dx + dy * dy);
    if (!(dist < radius)) return 0.0f;
    float ang = fabsf(pp_wrap_pi(heading - pp_fm_atan2(dy, dx)));
    const float a = C.apf_alpha - ang;
    ang = (a < 0.0f) ? 0.0f : a;
    const float d = 1.0f / dist - 1.0f / radius;
    float fp = C.apf_k * (d * d);
    fp = fp * ang / C.apf_alpha;
    return fp;
}

// The Dubins shot (HybridAStar.cpp:129-149): all lanes; returns true when accepted; samples in path[0 .. n_dubins).
// The sample positions come from float accumulators advanced one step at a time (Dubins.cpp:351-386): that chain is
// inherently sequential, so one lane runs it as three tight loops into `accb` (shared memory) and the 32 lanes then
// evaluate the samples (sin / cos, collision lookup) in parallel.
template <class W>
PP_HD bool pp_try_shot(const W& w, const PPConsts& C, const float* map, const PPFrame& F, float x, float y, float h,
                       PPPathPt* path, int path_cap, float* accb, int acc_cap, int* scratch, float (*shot)[5], float& len_out, int& n_dubins,
                       int& overflow)
{
    int type = PP_RSR; float p[4] = {0.0f, 0.0f, 0.0f, 0.0f}; PPDubinsCenters cen; PPDubinsPlan pl;
    // the four candidates on four lanes, then the sequential fold of Dubins.cpp:36-68 on every lane
    pp_dubins_centers<PPMathFp32>(C.r_min, x, y, h, F.goal_x, F.goal_y, F.goal_h, cen);
    for (int t = w.lane(); t < 4; t += W::LANES)
    {
        float csx, csy, cgx, cgy, pc[4];
        pp_dubins_pick(cen, t, csx, csy, cgx, cgy);
        shot[t][0] = pp_dubins_candidate<PPMathFp32>(t, C.r_min, h, F.goal_h, csx, csy, cgx, cgy, pc);
        shot[t][1] = pc[0]; shot[t][2] = pc[1]; shot[t][3] = pc[2]; shot[t][4] = pc[3];
    }
    w.sync();
    float len = 0.0f;
    for (int t = 0; t < 4; t++)
        if (t == 0 || shot[t][0] < len) { len = shot[t][0]; type = t; p[0] = shot[t][1]; p[1] = shot[t][2]; p[2] = shot[t][3]; p[3] = shot[t][4]; }
    w.sync();
    if (fabsf(p[1]) > (float)PP_PI_2) return false;                       // Dubins.cpp:152
    pp_dubins_plan<PPMathFp32>(C.r_min, C.step, C.ang_step, type, p, cen, pl);
    const int total = pl.size_3 + 1;
    bool blocked = false, over = false;
    if (total <= acc_cap)
    {
        if (w.lane() == 0)
        {
            float a = p[0];
            const float d1 = (pl.s1 < 0) ? -C.ang_step : C.ang_step, d3 = (pl.s2 < 0) ? -C.ang_step : C.ang_step;
            for (int k = 0; k < pl.size_1; k++) { accb[k] = a; a = a + d1; }          // a - s == a + (-s) exactly
            a = 0.0f;
            for (int k = pl.size_1; k < pl.size_2; k++) { accb[k] = a; a = a + C.step; }
            a = p[2];
            for (int k = pl.size_2; k < pl.size_3; k++) { accb[k] = a; a = a + d3; }
            accb[pl.size_3] = 0.0f;
        }
        w.sync();
        // a blocked sample anywhere rejects the shot: stop after the first round of samples that saw one
        // lane l takes samples l*R, l*R + 1, ...: every round looks at points spread over the whole path, so a blocked path
        // is usually rejected in the first round
        const int R = (total + W::LANES - 1) / W::LANES;
        for (int rr = 0; rr < R; rr++)
        {
            const int k = w.lane() * R + rr;
            if (k < total)
            {
                float sx, sy, sh, kappa;
                pp_dubins_sample<PPMathFp32>(pl, C.r_min, k, accb[k], sx, sy, sh, kappa);
                if (pp_path_point_blocked(C, map, sx, sy)) blocked = true;
                if (k < path_cap) { PPPathPt& q = path[k]; q.x = sx; q.y = sy; q.heading = sh; q.curvature = kappa; }
                else over = true;
            }
            if (w.any(blocked, scratch)) return false;
        }
    }
    else
    {
        float acc = p[0];
        for (int k = 0; k < total; k++)
        {
            if (k == pl.size_1) acc = 0.0f;
            if (k == pl.size_2) acc = p[2];
            if ((k % W::LANES) == w.lane())
            {
                float sx, sy, sh, kappa;
                pp_dubins_sample<PPMathFp32>(pl, C.r_min, k, acc, sx, sy, sh, kappa);
                if (pp_path_point_blocked(C, map, sx, sy)) blocked = true;
                if (k < path_cap) { PPPathPt& q = path[k]; q.x = sx; q.y = sy; q.heading = sh; q.curvature = kappa; }
                else over = true;
            }
            if (k < pl.size_1) acc = (pl.s1 < 0) ? acc - C.ang_step : acc + C.ang_step;
            else if (k < pl.size_2) acc = acc + C.step;
            else if (k < pl.size_3) acc = (pl.s2 < 0) ? acc - C.ang_step : acc + C.ang_step;
        }
    }
    if (w.any(blocked, scratch)) return false;
    overflow = w.any(over, scratch) ? 1 : 0;
    len_out = len; n_dubins = total;
    return true;
}

// ---- the search ---------------------------------------------------------------------------------------------------------
template <class W>
PP_HD_NOINLINE void pp_search_kpop(const W& w, const PPConsts& C, const float* off_xy, const PPGroup& G, const PPState& start,
                                   int kpop, PPKWork& wk, PPKSmem& sm, PPResult& res)
{
    const int lane = w.lane();
    const int N = C.N;
    const PPFrame& F = G.frame;
    const unsigned kb = (unsigned)(C.bins + 1);
    const int n_succ = 2 * C.A + 1;
    const unsigned goal_cell = (unsigned)(F.goal_ci * N + F.goal_cj);

    wk.l0 = sm.l0;
    // ---- init ----  (the hash table arrives all-ones = empty: the host clears it once, every query cleans up after itself)
    for (int t = lane; t < PP_K_LEVELS; t += W::LANES) { sm.head[t] = 0; sm.size[t] = 0; sm.taken[t] = 0; }
    w.sync();
    PPDubinsGoal gc;
    {
        float sg, cg;
        pp_fm_sincos(F.goal_h, sg, cg);
        gc.grx = F.goal_x + C.r_min * sg; gc.gry = F.goal_y - C.r_min * cg;
        gc.glx = F.goal_x - C.r_min * sg; gc.gly = F.goal_y + C.r_min * cg;
    }
    int n_nodes = 1;
    if (lane == 0)
    {
        PPKNode n0;
        n0.x = start.x; n0.y = start.y; n0.heading = start.heading; n0.g = 0.0f; n0.v2 = start.vmin_sqr; n0.f = start.f;
        n0.curv = start.curv; n0.bin = start.bin; n0.cell = start.ci * N + start.cj; n0.parent = -1;
        n0.key = (unsigned)n0.cell * kb + (unsigned)start.bin;
        unsigned nw0;
        int slot = pp_ktable_find_or_insert(wk, n0.key, nw0);
        n0.slot = slot;
        wk.nodes[0] = n0;
        wk.table[slot].node = 0u | PP_K_OPEN;
        wk.table[slot].pack = ((unsigned long long)pp_fbits(0.0f) << 32);
        PPKEntry e; e.f = n0.f; e.key = n0.key; e.idx = 0u; e.pad = 0;
        pp_klevel(wk, 0)[0] = e;
        sm.head[0] = 0; sm.size[0] = 1;
    }
    w.sync();

    int counter = 0, interval = C.shot_interval;
    PP_PROF_DECL
    PP_PROF_MARK(0)
    int n_pops = 0, n_oob = 0, status = 0, success = 0, n_chain = 0, n_dubins = 0, terminal = -1, n_iter = 0, n_popped = 0;
    float cost = FLT_MAX;

    for (unsigned it = 1;; it++)
    {
        // ---- pop the k smallest entries ----
        const int nb = pp_klsm_pop(w, wk, sm, kpop);
        PP_PROF_MARK(1)
        if (nb == 0) break;                                                     // open list exhausted: failure
        n_iter++; n_popped += nb;
        // validity (lazy deletion), ranks, closing, goal / shot candidates -- in rank (= queue) order
        int n_valid = 0, r_g = -1, r_s = -1;
        if (w.warp() == 0)                                                       // ballot group 0 decides, then broadcasts
        {
            const int wl = w.wlane();
            for (int base = 0; base < nb; base += W::BW)
            {
                const int b = base + wl;
                bool valid = false, is_goal = false, slow = false, oob = false;
                PPKEntry me = pp_kinf();
                int slot = -1;
                if (b < nb)
                {
                    me = sm.popped[b];
                    unsigned nw = 0u;
                    slot = pp_ktable_find(wk, me.key, nw);
                    valid = (slot >= 0) && (nw == (me.idx | PP_K_OPEN));         // closed or superseded => mismatch
                }
                const unsigned vm = w.ballot(valid);
                int rank = n_valid;
                { unsigned below = vm & w.lanemask_lt(); while (below) { rank++; below &= below - 1; } }
                if (valid)
                {
                    wk.table[slot].node = me.idx;                                // closed: open bit cleared
                    const PPKNode nd = wk.nodes[me.idx];
                    sm.parents[rank] = nd;
                    sm.pop_idx[rank] = (int)me.idx;
                    is_goal = ((unsigned)nd.cell == goal_cell);
                    slow = (nd.v2 < 1.0f);
                    oob = (nd.bin >= C.bins);
                    if (wk.trace && (n_pops + rank) < wk.trace_cap)
                    {
                        PPPop& t = wk.trace[n_pops + rank];
                        t.ci = nd.cell / N; t.cj = nd.cell % N; t.bin = nd.bin; t.x = nd.x; t.y = nd.y; t.heading = nd.heading; t.g = nd.g; t.f = nd.f;
                    }
                }
                const unsigned gm = w.ballot(is_goal), sl = w.ballot(slow), om = w.ballot(oob);
                if (r_g < 0 && gm)                                               // first goal pop
                {
                    int src = 0; { unsigned t = gm; while (!(t & 1u)) { t >>= 1; src++; } }
                    r_g = n_valid; { unsigned bl = vm & ((src == 0) ? 0u : (0xffffffffu >> (32 - src))); while (bl) { r_g++; bl &= bl - 1; } }
                }
                if (r_s < 0)                                                     // shot counter over the slow pops
                {
                    int ns = 0; { unsigned t = sl; while (t) { ns++; t &= t - 1; } }
                    const int need = interval - counter;
                    if (need >= 1 && need <= ns)
                    {
                        unsigned t = sl; int src = 0;
                        for (int q = 1; q < need; q++) t &= t - 1;               // drop the first need-1 slow pops
                        while (!(t & 1u)) { t >>= 1; src++; }
                        r_s = n_valid; { unsigned bl = vm & ((src == 0) ? 0u : (0xffffffffu >> (32 - src))); while (bl) { r_s++; bl &= bl - 1; } }
                        counter = interval;
                    }
                    else counter += ns;
                }
                { unsigned t = om; while (t) { n_oob++; t &= t - 1; } }
                { unsigned t = vm; while (t) { n_valid++; t &= t - 1; } }
                w.wsync();
            }
            if (wl == 0) { sm.bc[0] = n_valid; sm.bc[1] = r_g; sm.bc[2] = r_s; sm.bc[3] = counter; sm.bc[4] = n_oob; }
        }
        w.sync();
        n_valid = sm.bc[0]; r_g = sm.bc[1]; r_s = sm.bc[2]; counter = sm.bc[3]; n_oob = sm.bc[4];
        w.sync();
        if (n_valid == 0) continue;
        n_pops += n_valid;

        PP_PROF_MARK(2)
        // ---- goal / shot, rank ordered ----
        if (r_g >= 0 && (r_s < 0 || r_g < r_s)) { success = 1; terminal = sm.pop_idx[r_g]; cost = sm.parents[r_g].g; break; }
        if (r_s >= 0)
        {
            const PPKNode nd = sm.parents[r_s];
            float len = 0.0f; int ovf = 0;
            if (pp_try_shot(w, C, G.map, F, nd.x, nd.y, nd.heading, wk.path, wk.path_cap, reinterpret_cast<float*>(&sm.u),
                            (int)(sizeof(sm.u) / sizeof(float)), sm.scan, sm.shot, len, n_dubins, ovf))
            {
                if (ovf) status |= PP_STATUS_PATH_OVERFLOW;
                success = 1; cost = nd.g + len; terminal = nd.parent;
                if (terminal < 0) status |= PP_STATUS_NULL_TERMINAL;
                break;
            }
            n_dubins = 0;
            counter = 0;
            int ni = interval - C.shot_decay;
            interval = (ni < 50) ? 50 : ni;
            if (r_g >= 0) { success = 1; terminal = sm.pop_idx[r_g]; cost = sm.parents[r_g].g; break; }
        }
        if (n_nodes + n_valid * n_succ > wk.nodes_cap) { status |= PP_STATUS_CLOSED_OVERFLOW; break; }

        PP_PROF_MARK(6)
        // ---- expansion: candidate c = rank * n_succ + a ----
        const int n_cand = n_valid * n_succ;
        for (int c = lane; c < n_cand; c += W::LANES)
        {
            const int r = c / n_succ, a = c - r * n_succ;
            const PPKNode& pn = sm.parents[r];
            PPKCand cd; cd.curv_bin_ok = 0; cd.slot = -1; cd.pack = ~0ull; cd.cell = 0;
            cd.x = cd.y = cd.heading = cd.g = cd.v2 = 0.0f;
            int start_index = pn.curv - C.A; if (start_index < 0) start_index = 0;
            const int i = start_index + a;
            PPSucc o; o.ok = 0;
            const int pbin = (pn.bin > C.bins) ? C.bins : pn.bin;
            if (i < C.S && pp_rollout_one(C, off_xy, pn.x, pn.y, pn.heading, pn.g, pn.v2, pbin, i, o) &&
                pp_collision_free(C, G.map, o.x, o.y, o.ci, o.cj))
            {
                // APF: std::accumulate over the obstacles in order (Grid3D.cpp:226)
                float field = 0.0f;
                int q_lo = 0, q_hi = G.K;                                   // the successor's bin of the APF index, if any
                if (G.bin_off)
                {
                    const int bb = (o.ci >> G.bin_shift) * G.bin_n + (o.cj >> G.bin_shift);
                    q_lo = G.bin_off[bb]; q_hi = G.bin_off[bb + 1];
                }
                for (int qq = q_lo; qq < q_hi; qq++)
                {
                    const int q = G.bin_off ? G.bin_idx[qq] : qq;
                    float ox = G.apf[3 * q], oy = G.apf[3 * q + 1], rad = G.apf[3 * q + 2];
                    float dx = ox - o.x, dy = oy - o.y, lim = rad * 1.001f + 1e-3f;
                    if (dx * dx + dy * dy <= lim * lim)
                    {
                        float term = pp_kapf_term(C, ox, oy, rad, o.x, o.y, o.heading);
                        if (term != 0.0f) field = field + term;
                    }
                }
                const float g = o.g + field;
                const int cell = o.ci * N + o.cj;
                const unsigned key = (unsigned)cell * kb + (unsigned)o.bin;
                unsigned nw = 0u;
                const int sl2 = pp_ktable_find_or_insert(wk, key, nw);
                if (nw & PP_K_OPEN)                                          // not closed (includes: no node yet)
                {
                    cd.x = o.x; cd.y = o.y; cd.heading = o.heading; cd.g = g; cd.v2 = o.vmin_sqr;
                    cd.curv_bin_ok = (o.curv & 0xff) | ((o.bin & 0xff) << 8) | (1 << 16);
                    cd.cell = cell; cd.slot = sl2;
                    cd.pack = ((unsigned long long)pp_fbits(g) << 32) | (unsigned long long)(it * 1024u + (unsigned)c);
                    pp_atomic_min_u64(&wk.table[sl2].pack, cd.pack);
                }
            }
            sm.u.cand[c] = cd;
        }
        pp_fence();
        w.sync();

        PP_PROF_MARK(3)
        // ---- winners in c order -> nodes, queue entries ----
        int n_new = 0;
        for (int base = 0; base < n_cand; base += W::LANES)
        {
            const int c = base + lane;
            bool win = false;
            if (c < n_cand)
            {
                const PPKCand& cd = sm.u.cand[c];
                win = ((cd.curv_bin_ok >> 16) & 1) && (wk.table[cd.slot].pack == cd.pack);
            }
            int n_win = 0;
            const int pos = n_new + w.scan_count(win, sm.scan, n_win);
            if (win) sm.wlist[pos] = (unsigned short)c;
            n_new += n_win;
        }
        w.sync();
        // the t-th winner becomes node n_nodes + t; the work (Dubins length) is spread evenly over the lanes.  The
        // candidate staging area is reused for the batch: winner t reads cand[c >= t] before batch[t] is written, and
        // the winners of later rounds sit at higher addresses than anything written so far (16 t < 40 c).
        for (int base = 0; base < n_new; base += W::LANES)
        {
            const int t = base + lane;
            PPKEntry e = pp_kinf();
            if (t < n_new)
            {
                const int c = sm.wlist[t];
                const PPKCand cd = sm.u.cand[c];
                const int r = c / n_succ;
                const int idx = n_nodes + t;
                const float h2 = pp_kdubins(C, F, gc, cd.x, cd.y, cd.heading);
                const float h1 = wk.h1[cd.cell];
                PPKNode nd;
                nd.x = cd.x; nd.y = cd.y; nd.heading = cd.heading; nd.g = cd.g; nd.v2 = cd.v2;
                nd.f = cd.g + ((h1 < h2) ? h2 : h1);
                nd.curv = cd.curv_bin_ok & 0xff; nd.bin = (cd.curv_bin_ok >> 8) & 0xff; nd.cell = cd.cell;
                nd.parent = sm.pop_idx[r];
                nd.key = (unsigned)cd.cell * kb + (unsigned)nd.bin; nd.slot = cd.slot;
                wk.nodes[idx] = nd;
                wk.table[cd.slot].node = (unsigned)idx | PP_K_OPEN;
                e.f = nd.f; e.key = nd.key; e.idx = (unsigned)idx; e.pad = 0;
            }
            w.sync();
            if (t < n_new) sm.u.q.batch[t] = e;
            w.sync();
        }
        n_nodes += n_new;
        PP_PROF_MARK(4)
        if (n_new > 0)
        {
            pp_kranksort(w, sm.u.q.batch, sm.u.q.sorted, n_new);
            if (!pp_klsm_insert(w, wk, sm, sm.u.q.sorted, n_new)) { status |= PP_STATUS_OPEN_OVERFLOW; break; }
        }
        PP_PROF_MARK(5)
    }

    PP_PROF_FLUSH
    // ---- path: [dubins samples (already in wk.path[0 .. n_dubins)) | parent chain terminal -> start] ----
    if (lane == 0)
    {
        if (success)
            for (int c = terminal; c >= 0; c = wk.nodes[c].parent)
            {
                int at = n_dubins + n_chain;
                if (at < wk.path_cap)
                {
                    PPPathPt& q = wk.path[at];
                    q.x = wk.nodes[c].x; q.y = wk.nodes[c].y; q.heading = wk.nodes[c].heading;
                    q.curvature = C.abs_curv[wk.nodes[c].curv];
                }
                else status |= PP_STATUS_PATH_OVERFLOW;
                n_chain++;
            }
        res.success = success; res.status = status; res.cost = cost; res.n_pops = n_pops; res.n_pops_bin_oob = n_oob;
        res.n_chain = n_chain; res.n_dubins = n_dubins;
        res.n_lazy_searches = n_iter;      // K-POP: iterations
        res.n_lazy_pops = n_popped;        // K-POP: queue entries taken (valid + stale)
        res.max_open = 0;
        res.n_closed = n_nodes; res.pad = 0;
    }
    // ---- leave the hash table empty for the slot's next query: every inserted key belongs to at least one logged node ----
    w.sync();
    // (superseded nodes of one key share a slot: several lanes may store the same all-ones record to it -- harmless)
    for (int t = lane; t < n_nodes; t += W::LANES) pp_kslot_clear(wk.table + wk.nodes[t].slot);
    w.sync();
}

#endif
