// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Race check of the search cores on the CPU (K-POP: multi-warp CTA; EXACT: one warp).  The product's pp_search_kpop
// (csrc/core/pp_kpop.h) and pp_search_exact (csrc/core/pp_search.h) are written against a lane policy W (lane ids, barriers, ballots); the CUDA kernel instantiates it with a CTA of 4 warps.  Here the
// same code runs on NL host threads, one per lane, in ballot groups ("warps") of BW lanes, with every barrier a
// pthread barrier and every device atomic a GCC builtin (PP_HOST_ATOMICS).  Built with -fsanitize=thread, any pair of
// conflicting shared-memory / pool accesses that is not ordered by a barrier is reported by ThreadSanitizer, and the
// result must be bit-identical to the single-lane run of the same code.
//
//   search_mt <scenario.bin>     (written by tests/test_cpu_search_mt.py; layout above main below)
#define PP_HOST_ATOMICS 1
#include <pthread.h>
#include <thread>

#include "host_emul.cpp"

namespace
{
template <int NL, int BWW>
struct MTShared
{
    pthread_barrier_t block;
    pthread_barrier_t group[NL / BWW];
    unsigned bal[NL / BWW][BWW];
    int flag[NL];
    MTShared()
    {
        pthread_barrier_init(&block, nullptr, NL);
        for (int g = 0; g < NL / BWW; g++) pthread_barrier_init(&group[g], nullptr, BWW);
    }
};

template <int NL, int BWW>
struct PPBlockHost
{
    enum { LANES = NL, BW = BWW };
    int id;
    MTShared<NL, BWW>* sh;
    int lane() const { return id; }
    int wlane() const { return id % BWW; }
    int warp() const { return id / BWW; }
    void sync() const { pthread_barrier_wait(&sh->block); }
    void wsync() const { pthread_barrier_wait(&sh->group[warp()]); }
    unsigned ballot(bool p) const
    {
        sh->bal[warp()][wlane()] = p ? 1u : 0u;
        wsync();
        unsigned m = 0;
        for (int t = 0; t < BWW; t++) m |= sh->bal[warp()][t] << t;
        wsync();
        return m;
    }
    unsigned lanemask_lt() const { return (1u << wlane()) - 1u; }
    bool any(bool p, int*) const
    {
        sh->flag[id] = p ? 1 : 0;
        sync();
        int r = 0;
        for (int t = 0; t < NL; t++) r |= sh->flag[t];
        sync();
        return r != 0;
    }
    int scan_count(bool p, int*, int& total) const
    {
        sh->flag[id] = p ? 1 : 0;
        sync();
        int pos = 0, tot = 0;
        for (int t = 0; t < NL; t++) { if (t < id) pos += sh->flag[t]; tot += sh->flag[t]; }
        sync();
        total = tot;
        return pos;
    }
};

// one warp of NL lanes for the EXACT mode (pp_search_exact): sync = the warp barrier, ballots / shuffles over all NL lanes
template <int NL>
struct MTWarpShared
{
    pthread_barrier_t bar;
    unsigned bal[NL];
    unsigned char xch[NL][16];
    MTWarpShared() { pthread_barrier_init(&bar, nullptr, NL); }
};

template <int NL>
struct PPWarpHostMT
{
    enum { LANES = NL };
    int id;
    MTWarpShared<NL>* sh;
    int lane() const { return id; }
    void sync() const { pthread_barrier_wait(&sh->bar); }
    unsigned ballot(bool p) const
    {
        sh->bal[id] = p ? 1u : 0u;
        sync();
        unsigned m = 0;
        for (int t = 0; t < NL; t++) m |= sh->bal[t] << t;
        sync();
        return m;
    }
    unsigned lanemask_lt() const { return (1u << id) - 1u; }
    template <class T> T shfl(T v, int src) const
    {
        static_assert(sizeof(T) <= 16, "shuffle payload");
        std::memcpy(sh->xch[id], &v, sizeof(T));
        sync();
        T r; std::memcpy(&r, sh->xch[src], sizeof(T));
        sync();
        return r;
    }
};

struct ExactPools
{
    std::vector<PPNode3> open3; std::vector<PPClosed3> closed; std::vector<PPHashSlot> chash; std::vector<unsigned> cell_state;
    std::vector<float> nm_g, nm_f, cl_g; std::vector<int> cl_prev; std::vector<PPNode2> open2; std::vector<PPPathPt> path;
    std::vector<PPPop> trace;
    unsigned hist_sid = 0;      // planner-object history (PPWork::lazy_sid), used by the carried-cache pass only
    std::vector<unsigned long long> arena_mem; PPArena arena;   // growable containers (pp_arena.h) on host memory
    PPWork wk;
    explicit ExactPools(int N)
    {
        // tiny fixed pools + arena: the cooperative growth paths (copy by all lanes, concurrent re-hash) run under TSan too
        const int closed_cap = 96, open_cap = 64, open2_cap = 32;
        open3.resize(open_cap); closed.resize(closed_cap);
        int hc = 1; while (hc < 2 * closed_cap) hc <<= 1;
        chash.resize(hc);
        cell_state.resize((size_t)N * N); nm_g.resize((size_t)N * N); nm_f.resize((size_t)N * N); cl_g.resize((size_t)N * N);
        cl_prev.resize((size_t)N * N); open2.resize(open2_cap); path.resize(4096); trace.resize(1 << 16);
        wk.open3 = open3.data(); wk.open3_cap = open_cap; wk.closed = closed.data(); wk.closed_cap = closed_cap;
        wk.chash = chash.data(); wk.chash_cap = hc; wk.cell_state = cell_state.data(); wk.nm_g = nm_g.data(); wk.nm_f = nm_f.data();
        wk.cl_g = cl_g.data(); wk.cl_prev = cl_prev.data(); wk.open2 = open2.data(); wk.open2_cap = open2_cap;
        wk.path = path.data(); wk.path_cap = (int)path.size(); wk.trace = trace.data(); wk.trace_cap = (int)trace.size();
        arena_mem.resize((size_t)(256u << 20) / 8);
        std::memset(&arena, 0, sizeof(PPArena));
        arena.base = (unsigned long long)arena_mem.data(); arena.size = arena_mem.size() * 8ull;
        wk.arena = &arena; wk.closed_max = 1 << 16; wk.open3_max = 1 << 15; wk.open2_max = 1 << 14;
    }
};

template <int NL>
PPResult run_exact_mt(Emu* e, const PPState& st, ExactPools& P)
{
    MTWarpShared<NL> sh;
    std::unique_ptr<PPSmem> sm(new PPSmem());
    PPResult res[NL];
    PPGroup G = group_of(e);
    std::vector<std::thread> th;
    for (int t = 0; t < NL; t++)
        th.emplace_back([&, t] {
            PPWarpHostMT<NL> w; w.id = t; w.sh = &sh;
            PPWork wk = P.wk;
            pp_search_exact(w, e->m.C, e->m.off_xy.data(), G, st, wk, *sm, res[t]);
        });
    for (auto& t : th) t.join();
    return res[0];
}

bool same_exact(const PPResult& a, const ExactPools& A, const PPResult& b, const ExactPools& B)
{
    if (a.success != b.success || a.status != b.status || a.n_pops != b.n_pops || a.n_chain != b.n_chain || a.n_dubins != b.n_dubins ||
        a.n_closed != b.n_closed || a.n_lazy_pops != b.n_lazy_pops || std::memcmp(&a.cost, &b.cost, 4) != 0) return false;
    int np = std::min(a.n_pops, (int)A.trace.size());
    if (std::memcmp(A.trace.data(), B.trace.data(), sizeof(PPPop) * np) != 0) return false;
    int n = std::min(a.n_chain + a.n_dubins, (int)A.path.size());
    return std::memcmp(A.path.data(), B.path.data(), sizeof(PPPathPt) * n) == 0;
}

struct Pools
{
    std::vector<PPKNode> nodes; std::vector<PPKSlot> table; std::vector<PPKEntry> arena, ta, tb; std::vector<PPPathPt> path;
    std::vector<PPPop> trace;
    PPKWork wk;
    Pools(int max_nodes, const float* h1)
    {
        nodes.resize(max_nodes);
        int tc = 1; while (tc < 2 * max_nodes) tc <<= 1;
        table.resize(tc);
        std::memset(table.data(), 0xFF, sizeof(PPKSlot) * table.size());
        int levels = 1; while (((size_t)PP_K_RUN0 << (levels - 1)) < (size_t)max_nodes + PP_K_RUN0 && levels < PP_K_LEVELS) levels++;
        arena.resize((size_t)PP_K_RUN0 * ((1 << levels) - 1)); ta.resize(max_nodes + 2 * PP_K_RUN0); tb.resize(max_nodes + 2 * PP_K_RUN0);
        path.resize(4096); trace.resize(1 << 16);
        wk.nodes = nodes.data(); wk.nodes_cap = max_nodes; wk.table = table.data(); wk.table_cap = tc;
        wk.arena = arena.data(); wk.tmp_a = ta.data(); wk.tmp_b = tb.data(); wk.lsm_levels = levels; wk.h1 = h1; wk.l0 = nullptr;
        wk.path = path.data(); wk.path_cap = (int)path.size(); wk.trace = trace.data(); wk.trace_cap = (int)trace.size();
    }
    bool table_clean() const
    {
        const unsigned char* b = reinterpret_cast<const unsigned char*>(table.data());
        for (size_t t = 0; t < sizeof(PPKSlot) * table.size(); t++) if (b[t] != 0xFF) return false;
        return true;
    }
};

template <int NL, int BWW>
PPResult run_mt(Emu* e, const PPState& st, int k, Pools& P)
{
    MTShared<NL, BWW> sh;
    std::unique_ptr<PPKSmem> sm(new PPKSmem());
    PPResult res[NL];
    PPGroup G = group_of(e);
    std::vector<std::thread> th;
    for (int t = 0; t < NL; t++)
        th.emplace_back([&, t] {
            PPBlockHost<NL, BWW> w; w.id = t; w.sh = &sh;
            PPKWork wk = P.wk;                       // every lane has its own copy of the pointer record, like registers
            pp_search_kpop(w, e->m.C, e->m.off_xy.data(), G, st, k, wk, *sm, res[t]);
        });
    for (auto& t : th) t.join();
    return res[0];
}

bool same(const PPResult& a, const Pools& A, const PPResult& b, const Pools& B)
{
    if (a.success != b.success || a.status != b.status || a.n_pops != b.n_pops || a.n_chain != b.n_chain || a.n_dubins != b.n_dubins ||
        a.n_closed != b.n_closed || std::memcmp(&a.cost, &b.cost, 4) != 0) return false;
    int np = std::min(a.n_pops, (int)A.trace.size());
    if (std::memcmp(A.trace.data(), B.trace.data(), sizeof(PPPop) * np) != 0) return false;
    int n = std::min(a.n_chain + a.n_dubins, (int)A.path.size());
    return std::memcmp(A.path.data(), B.path.data(), sizeof(PPPathPt) * n) == 0;
}
}

// scenario.bin: orc_params | goal3 | start3 | int n_boxes | boxes[n][4] | conf[n] | float apf_added_radius | map[N*N] |
//               h1[N*N] | int n_queries | queries[n][4] (x, y, heading, vel) | int k
int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: search_mt scenario.bin\n"); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror("scenario"); return 2; }
    auto rd = [&](void* p, size_t n) { if (std::fread(p, 1, n, f) != n) { std::fprintf(stderr, "short scenario file\n"); std::exit(2); } };
    orc_params prm; float goal[3], start[3]; int nb;
    rd(&prm, sizeof(prm)); rd(goal, 12); rd(start, 12); rd(&nb, 4);
    std::vector<float> boxes((size_t)nb * 4), conf(nb); float radius;
    rd(boxes.data(), boxes.size() * 4); rd(conf.data(), conf.size() * 4); rd(&radius, 4);
    Emu* e = static_cast<Emu*>(emu_create(&prm));
    if (!e) return 2;
    const int N = e->m.C.N;
    emu_update_goal(e, goal, start);
    emu_update_boxes(e, boxes.data(), conf.data(), nb, radius);       // APF list + its spatial index
    std::vector<float> map((size_t)N * N), h1((size_t)N * N);
    rd(map.data(), map.size() * 4); rd(h1.data(), h1.size() * 4);
    emu_set_map(e, map.data());
    int nq; rd(&nq, 4);
    std::vector<float> q((size_t)nq * 4); rd(q.data(), q.size() * 4);
    int k; rd(&k, 4);
    std::fclose(f);
    int bad = 0;
    for (int i = 0; i < nq; i++)
    {
        PPState st = pp_host_set_start(e->m.C, e->fr, q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
        const int max_nodes = 1 << 16;
        Pools P1(max_nodes, h1.data()), P4(max_nodes, h1.data()), P8(max_nodes, h1.data());
        std::unique_ptr<PPKSmem> sm(new PPKSmem());
        PPResult r1; PPWarpSerial w1;
        { PPGroup G = group_of(e); PPKWork wk = P1.wk; pp_search_kpop(w1, e->m.C, e->m.off_xy.data(), G, st, k, wk, *sm, r1); }
        PPResult r4 = run_mt<4, 2>(e, st, k, P4);
        PPResult r8 = run_mt<8, 4>(e, st, k, P8);
        bool ok = same(r1, P1, r4, P4) && same(r1, P1, r8, P8) && P1.table_clean() && P4.table_clean() && P8.table_clean() && r1.status == 0;
        std::printf("query %d: success %d pops %d cost %.6f nodes %d | 4 lanes (2 x 2) and 8 lanes (2 x 4): %s\n", i, r1.success, r1.n_pops,
                    r1.cost, r1.n_closed, ok ? "identical" : "MISMATCH");
        if (!ok) bad++;
        // EXACT mode: one lane vs a warp of 8 and of 32 lanes
        ExactPools E1(N), E8(N), E32(N);
        std::unique_ptr<PPSmem> xsm(new PPSmem());
        PPResult x1;
        { PPGroup G = group_of(e); PPWork wk = E1.wk; pp_search_exact(w1, e->m.C, e->m.off_xy.data(), G, st, wk, *xsm, x1); }
        PPResult x8 = run_exact_mt<8>(e, st, E8);
        PPResult x32 = run_exact_mt<32>(e, st, E32);
        bool xok = same_exact(x1, E1, x8, E8) && same_exact(x1, E1, x32, E32) && x1.status == 0;
        std::printf("query %d EXACT: success %d pops %d cost %.6f lazy pops %d | 8 and 32 lanes: %s\n", i, x1.success, x1.n_pops, x1.cost,
                    x1.n_lazy_pops, xok ? "identical" : "MISMATCH");
        if (!xok) bad++;
    }
    // EXACT mode with planner-object history (pp_set_history): all queries in sequence on ONE carried 2D cache, one lane vs a
    // warp of 32, the second half of the sequence started just below the stamp wrap-around so that the all-lane stamp drop
    // of pp_search_exact runs under the race detector too.  Results and the carried cache must stay identical.
    {
        ExactPools H1(N), H32(N);
        H1.wk.lazy_sid = &H1.hist_sid; H32.wk.lazy_sid = &H32.hist_sid;
        std::unique_ptr<PPSmem> xsm(new PPSmem());
        PPWarpSerial w1;
        for (int rep = 0; rep < 2; rep++)
            for (int i = 0; i < nq; i++)
            {
                if (rep == 1 && i == 0) { H1.hist_sid = (PP_CS_STAMP >> 1) - 3; H32.hist_sid = (PP_CS_STAMP >> 1) - 3; }
                PPState st = pp_host_set_start(e->m.C, e->fr, q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
                PPResult x1;
                { PPGroup G = group_of(e); PPWork wk = H1.wk; pp_search_exact(w1, e->m.C, e->m.off_xy.data(), G, st, wk, *xsm, x1); }
                PPResult x32 = run_exact_mt<32>(e, st, H32);
                bool cache_ok = H1.hist_sid == H32.hist_sid;
                for (size_t c = 0; c < H1.cell_state.size() && cache_ok; c++)
                {
                    cache_ok = H1.cell_state[c] == H32.cell_state[c];
                    if (cache_ok && (H1.cell_state[c] & PP_CS_TOUCHED))
                        cache_ok = std::memcmp(&H1.nm_g[c], &H32.nm_g[c], 4) == 0 && std::memcmp(&H1.nm_f[c], &H32.nm_f[c], 4) == 0;
                }
                bool hok = same_exact(x1, H1, x32, H32) && cache_ok && x1.status == 0;
                std::printf("query %d.%d EXACT on the carried cache: pops %d cost %.6f lazy searches %d sid %u | 32 lanes: %s\n", rep, i,
                            x1.n_pops, x1.cost, x1.n_lazy_searches, H1.hist_sid, hok ? "identical" : "MISMATCH");
                if (!hok) bad++;
            }
    }
    emu_destroy(e);
    return bad ? 1 : 0;
}
