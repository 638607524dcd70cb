// sm_100a kernels of the local-planner hot path.  Compiled with -fmad=false (bit-exact parity with the
// reference's x86-64 arithmetic needs every a*b+c rounded twice) and full-precision div/sqrt.
//
//   pp_search_kernel        persistent warps, one query per warp at a time (dynamic fetch), EXACT mode
//   pp_successor_kernel     one warp per popped state: roll-out lanes = steering primitives,
//                           collision lookup, APF lanes = obstacles           (north_star (b), (c))
//   pp_apf_kernel           one warp per pose, order-preserving warp sum
//   pp_collision_kernel     one thread per point (the reference's check is a single-cell lookup, F3)
//   pp_footprint_kernel     one warp per pose: footprint window staged in shared memory, lanes = footprint cells, ballot
//   pp_velocity_profile_kernel  one thread per path: the three sequential v^2 passes of VelocityGenerator
//   pp_trajectory_kernel    one warp per query of the last batch: world transform (lanes = points) + profile + message layout
//   pp_dubins_length_kernel 4 lanes per state (one per CSC candidate), FP64-evaluated/rounded ("pinned libm")
//   pp_dubins_path_kernel   one warp: plan + parallel sampling
//   pp_lazy_astar_kernel    one control lane: the lazy cached 2D A* on a fresh cache (parity tests)
//   pp_map_decay_kernel     float4 grid-stride streaming, HBM/L2 bound
//   pp_map_update_kernel    gather-form log-odds rasterisation (+ optional fused decay), one CTA per 32x32 tile, device-side binning
//   pp_map_lines_kernel     lane-line rasteriser, one CTA per 32x32 tile, lines in order, shared-memory sample counters
//   pp_map_reloc_*          forward-scatter resample with atomicMax(source index) = "last writer wins"
#ifndef PP_KERNELS_CUH
#define PP_KERNELS_CUH

#include <cuda_runtime.h>
#ifdef PP_PROFILE
__device__ unsigned long long pp_prof_acc[16];
#endif
#include "../core/pp_search.h"
#include "../core/pp_kpop.h"
#include "../core/pp_map.h"
#include "../core/pp_footprint.h"
#include "../core/pp_velocity.h"

// ONE warp per CTA: a query that runs on long after its neighbours have finished (expansion counts span three orders of
// magnitude) then holds one warp's worth of registers, not a whole CTA's.  With 4-warp CTAs the tail CTAs of the batches in flight
// kept a third of the SMs' warp slots idle (streamed C4 batches: 14.2 M expansions/s with 4 warps per CTA, 23.9 M with 1).
#ifndef PP_SEARCH_WARPS
#define PP_SEARCH_WARPS 1
#endif
#ifndef PP_SEARCH_MIN_BLOCKS
#define PP_SEARCH_MIN_BLOCKS 16    // resident CTAs per SM the register allocation is tuned for (128 registers per thread)
#endif
#define PP_TILE 32

struct PPWarpDev
{
    enum { LANES = 32 };
    __host__ __device__ __forceinline__ int lane() const
    {
#ifdef __CUDA_ARCH__
        return threadIdx.x & 31;
#else
        return 0;
#endif
    }
    __host__ __device__ __forceinline__ void sync() const
    {
#ifdef __CUDA_ARCH__
        __syncwarp();
#endif
    }
    __host__ __device__ __forceinline__ unsigned ballot(bool p) const
    {
#ifdef __CUDA_ARCH__
        return __ballot_sync(0xffffffffu, p);
#else
        return p ? 1u : 0u;
#endif
    }
    __host__ __device__ __forceinline__ unsigned lanemask_lt() const
    {
#ifdef __CUDA_ARCH__
        return (1u << (threadIdx.x & 31)) - 1u;
#else
        return 0u;
#endif
    }
    template <class T> __host__ __device__ __forceinline__ T shfl(T v, int src) const
    {
#ifdef __CUDA_ARCH__
        return __shfl_sync(0xffffffffu, v, src);
#else
        return v;
#endif
    }
};

// The K-POP mode runs one query on a CTA of NW warps: LANES = 32*NW cooperating lanes, ballots per hardware warp.
template <int NW>
struct PPBlockDev
{
    enum { LANES = 32 * NW, BW = 32 };
    __device__ __forceinline__ int lane() const { return threadIdx.x; }
    __device__ __forceinline__ int wlane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int warp() const { return threadIdx.x >> 5; }
    __device__ __forceinline__ void sync() const { if (NW == 1) __syncwarp(); else __syncthreads(); }
    __device__ __forceinline__ void wsync() const { __syncwarp(); }
    __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ unsigned lanemask_lt() const { return (1u << (threadIdx.x & 31)) - 1u; }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ bool any(bool p, int* scratch) const
    {
        if (NW == 1) return __ballot_sync(0xffffffffu, p) != 0u;
        return __syncthreads_or(p ? 1 : 0) != 0;
    }
    __device__ __forceinline__ int scan_count(bool p, int* scratch, int& total) const
    {
        const unsigned m = __ballot_sync(0xffffffffu, p);
        int pos = __popc(m & lanemask_lt());
        if (NW == 1) { total = __popc(m); return pos; }
        if (wlane() == 0) scratch[warp()] = __popc(m);
        __syncthreads();
        int t = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) { int v = scratch[k]; if (k < warp()) pos += v; t += v; }
        __syncthreads();
        total = t;
        return pos;
    }
};

// ---------------------------------------------------------------------------------------------------
struct PPBatchArgs
{
    PPConsts        C;
    const float*    off_xy;
    const PPGroup*  groups;
    const PPQuery*  queries;
    const int*      qmap;       // optional indirection (retry pass): work item -> query index
    const int*      order;      // optional fetch order of the work items (longest expected first), nullptr = as given
    int             n_queries;  // number of work items
    int             n_slots;
    int*            counter;
    PPResult*       results;
    PPPathPt*       paths;      int path_cap;
    PPPop*          trace;      int trace_cap;
    // per-slot scratch pools
    PPNode3*        open3;      int open3_cap;
    PPClosed3*      closed;     int closed_cap;
    PPHashSlot*     chash;      int chash_cap;
    unsigned*       cell_state;
    float*          nm_g;
    float*          nm_f;
    float*          cl_g;
    int*            cl_prev;
    PPNode2*        open2;      int open2_cap;
    // planner-object history (pp_set_history): the carried 2D cache of ONE planner; only valid with one slot and one query
    unsigned*       hist_cell_state;
    float*          hist_nm_g;
    float*          hist_nm_f;
    unsigned*       hist_sid;   // nullptr = fresh cache per query
    // growable containers (csrc/core/pp_arena.h): nullptr = the fixed pools above are all a query gets
    PPArena*        arena;
    int             closed_max, open3_max, open2_max;   // hard caps (the caller's max_expansions / max_open / max_open2d)
};

__device__ __forceinline__ void pp_slot_work(const PPBatchArgs& a, int slot, PPWork& wk)
{
    size_t nn = (size_t)a.C.N * a.C.N;
    wk.open3 = a.open3 + (size_t)slot * a.open3_cap;   wk.open3_cap = a.open3_cap;
    wk.closed = a.closed + (size_t)slot * a.closed_cap; wk.closed_cap = a.closed_cap;
    wk.chash = a.chash + (size_t)slot * a.chash_cap;   wk.chash_cap = a.chash_cap;
    wk.cell_state = a.cell_state + slot * nn;
    wk.nm_g = a.nm_g + slot * nn;
    wk.nm_f = a.nm_f + slot * nn;
    wk.cl_g = a.cl_g + slot * nn;
    wk.cl_prev = a.cl_prev + slot * nn;
    wk.open2 = a.open2 + (size_t)slot * a.open2_cap;   wk.open2_cap = a.open2_cap;
    wk.path = nullptr; wk.path_cap = 0; wk.trace = nullptr; wk.trace_cap = 0;
    wk.lazy_sid = nullptr;
    wk.arena = a.arena; wk.closed_max = a.closed_max; wk.open3_max = a.open3_max; wk.open2_max = a.open2_max;
    if (a.hist_sid)
    {
        wk.cell_state = a.hist_cell_state; wk.nm_g = a.hist_nm_g; wk.nm_f = a.hist_nm_f; wk.lazy_sid = a.hist_sid;
    }
}

// AStar::reset() (AStar.cpp:56-60) on a carried cache: only the visited flags go, node costs stay
__global__ void __launch_bounds__(256) pp_hist_reset_kernel(unsigned* __restrict__ cell_state, int n)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) cell_state[c] &= ~PP_CS_VISITED;
}

__global__ void __launch_bounds__(PP_SEARCH_WARPS * 32, PP_SEARCH_MIN_BLOCKS)
pp_search_kernel(const __grid_constant__ PPBatchArgs a)
{
    __shared__ PPSmem sm[PP_SEARCH_WARPS];
    const int warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * PP_SEARCH_WARPS + warp;
    if (slot >= a.n_slots) return;
    PPWarpDev w;
    PPWork wk;
    pp_slot_work(a, slot, wk);
    for (;;)
    {
        int q = 0;
        if (w.lane() == 0) q = atomicAdd(a.counter, 1);
        q = w.shfl(q, 0);
        if (q >= a.n_queries) break;
        if (a.order) q = a.order[q];
        if (a.qmap) q = a.qmap[q];
        wk.path = a.paths + (size_t)q * a.path_cap; wk.path_cap = a.path_cap;
        wk.trace = a.trace ? a.trace + (size_t)q * a.trace_cap : nullptr;
        wk.trace_cap = a.trace ? a.trace_cap : 0;
        const PPQuery Q = a.queries[q];
        const PPGroup G = a.groups[Q.group];
        PPResult res;
        pp_search_exact(w, a.C, a.off_xy, G, Q.start, wk, sm[warp], res);
        if (w.lane() == 0) a.results[q] = res;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// K-POP mode: one warp (= one CTA of 32 threads) per query at a time; see csrc/core/pp_kpop.h
struct PPKpopArgs
{
    PPConsts        C;
    const float*    off_xy;
    const PPGroup*  groups;
    const float*    field2d;    // num_groups x N*N exact 2D distance fields
    const PPQuery*  queries;
    const int*      qmap;
    const int*      order;      // optional launch order of the work items (pp_kpop_order_kernel)
    int             n_queries;
    int             n_slots;
    int             kpop;
    int*            counter;
    PPResult*       results;
    PPPathPt*       paths;      int path_cap;
    PPPop*          trace;      int trace_cap;
    // per-slot pools
    PPKNode*        nodes;      int nodes_cap;
    PPKSlot*        table;      int table_cap;
    PPKEntry*       arena;      size_t arena_cap;
    PPKEntry*       tmp_a;      PPKEntry* tmp_b;   size_t tmp_cap;
    int             lsm_levels;
};

#ifndef PP_KPOP_MIN_BLOCKS
#define PP_KPOP_MIN_BLOCKS 6      // 80 registers: 6 CTAs x 4 warps per SM measured best on large batches
#endif
template <int NW>
__global__ void __launch_bounds__(32 * NW, PP_KPOP_MIN_BLOCKS) pp_kpop_kernel(const __grid_constant__ PPKpopArgs a)
{
    __shared__ PPKSmem sm;
    __shared__ int s_q;
    const int slot = blockIdx.x;
    if (slot >= a.n_slots) return;
    PPBlockDev<NW> w;
    PPKWork wk;
    wk.nodes = a.nodes + (size_t)slot * a.nodes_cap;   wk.nodes_cap = a.nodes_cap;
    wk.table = a.table + (size_t)slot * a.table_cap;   wk.table_cap = a.table_cap;
    wk.arena = a.arena + (size_t)slot * a.arena_cap;
    wk.tmp_a = a.tmp_a + (size_t)slot * a.tmp_cap;     wk.tmp_b = a.tmp_b + (size_t)slot * a.tmp_cap;
    wk.lsm_levels = a.lsm_levels;
    const size_t nn = (size_t)a.C.N * a.C.N;
    for (;;)
    {
        if (w.lane() == 0) s_q = atomicAdd(a.counter, 1);
        w.sync();
        int q = s_q;
        w.sync();
        if (q >= a.n_queries) break;
        if (a.order) q = a.order[q];
        if (a.qmap) q = a.qmap[q];
        const PPQuery Q = a.queries[q];
        const PPGroup G = a.groups[Q.group];
        wk.h1 = a.field2d + nn * Q.group;
        wk.path = a.paths + (size_t)q * a.path_cap; wk.path_cap = a.path_cap;
        wk.trace = a.trace ? a.trace + (size_t)q * a.trace_cap : nullptr;
        wk.trace_cap = a.trace ? a.trace_cap : 0;
        PPResult res;
#ifdef PP_PROFILE
        unsigned long long t_begin; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
#endif
        pp_search_kpop(w, a.C, a.off_xy, G, Q.start, a.kpop, wk, sm, res);
#ifdef PP_PROFILE
        unsigned long long t_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        res.max_open = (int)((t_begin / 1000ull) & 0x7fffffffull); res.n_pops_bin_oob = (int)((t_end / 1000ull) & 0x7fffffffull);   // us
#endif
        if (w.lane() == 0) a.results[q] = res;
        w.sync();
    }
}

// Longest-expected-first launch order of a K-POP batch: rank the work items by (exact 2D distance of the start cell) x
// (how many iterations per query the item's group needed in the previous batch, when known), descending, ties by index.
// Only the order in which the resident slots fetch queries changes -- every query is independent, results are
// unaffected -- but the batch no longer ends on a long query that was fetched last.
// group_cost == nullptr selects the EXACT-mode key: a start the exact 2D field cannot connect to the goal makes the reference's
// search exhaust the whole reachable state space (its longest queries by far), so those go first; the rest by distance.
__device__ __forceinline__ float pp_kpop_order_key(const PPQuery& Q, const float* field2d, int N, const float* group_cost, float unknown_cost)
{
    const bool in = Q.start.ci >= 0 && Q.start.ci < N && Q.start.cj >= 0 && Q.start.cj < N;
    float k = in ? field2d[(size_t)N * N * Q.group + (size_t)Q.start.ci * N + Q.start.cj] : 0.0f;
    if (!group_cost) return (k < 3.0e38f) ? k : 3.4e38f;
    if (!(k < 3.0e38f)) return 0.0f;                                  // K-POP: unreachable ends at once
    const float gc = group_cost[Q.group];
    return k * (gc > 0.0f ? gc : unknown_cost);
}

// Any block size up to 256 threads.  The EXACT-mode upload launches it with ONE warp per CTA: while other lanes' search kernels
// fill the GPU a freed warp slot is all there is, and a 256-thread CTA would wait for eight of them on one SM.
__global__ void __launch_bounds__(256) pp_kpop_order_kernel(const PPQuery* queries, const int* qmap, int n, const float* field2d,
                                                            int N, const float* group_cost, float unknown_cost, int* order)
{
    __shared__ float s_key[256];
    const int B = (int)blockDim.x;
    const int i = blockIdx.x * B + threadIdx.x;
    const float ki = (i < n) ? pp_kpop_order_key(queries[qmap ? qmap[i] : i], field2d, N, group_cost, unknown_cost) : 0.0f;
    int rank = 0;
    for (int base = 0; base < n; base += B)                            // tiles of B keys staged in shared memory
    {
        const int j = base + threadIdx.x;
        __syncthreads();
        s_key[threadIdx.x] = (j < n) ? pp_kpop_order_key(queries[qmap ? qmap[j] : j], field2d, N, group_cost, unknown_cost) : -1.0f;
        __syncthreads();
        const int m = min(B, n - base);
        for (int t = 0; t < m; t++)
        {
            const float kj = s_key[t];
            rank += (kj > ki || (kj == ki && base + t < i)) ? 1 : 0;
        }
    }
    if (i < n) order[rank] = i;
}

// ---------------------------------------------------------------------------------------------------
// stateless batches
struct PPSuccArgs
{
    PPConsts       C;
    const float*   off_xy;
    PPGroup        G;
    const PPState* in;
    int            n;
    int            expand;    // 0: VehicleModel::get_neighbors only, 1: Grid3D::get_neighbors
    PPState*       out;       // n x (2A+1), compacted
    int*           n_out;
    int*           flags;
};

__global__ void __launch_bounds__(128) pp_successor_kernel(const __grid_constant__ PPSuccArgs a)
{
    __shared__ PPSmem sm[4];
    const int warp = threadIdx.x >> 5;
    const int k = blockIdx.x * 4 + warp;
    if (k >= a.n) return;
    PPWarpDev w;
    const int lane = w.lane();
    const PPState s = a.in[k];
    const int stride = 2 * a.C.A + 1;
    PPSmem& m = sm[warp];
    if (a.expand)
        pp_expand_warp(w, a.C, a.off_xy, a.G, s.x, s.y, s.heading, s.g, s.vmin_sqr, s.curv, s.bin, m);
    else
    {
        int start_index = s.curv - a.C.A;
        if (start_index < 0) start_index = 0;
        if (lane < stride)
        {
            PPSucc o; o.ok = 0; o.ci = -1; o.cj = -1;
            int i = start_index + lane;
            int bin = (s.bin > a.C.bins) ? a.C.bins : s.bin;
            if (i < a.C.S && pp_rollout_one(a.C, a.off_xy, s.x, s.y, s.heading, s.g, s.vmin_sqr, bin, i, o)) o.ok = 1;
            m.succ[lane] = o;
        }
        __syncwarp();
    }
    if (lane == 0)
    {
        int cnt = 0;
        for (int q = 0; q < stride; q++)
        {
            const PPSucc& o = m.succ[q];
            if (!o.ok) continue;
            PPState r;
            r.x = o.x; r.y = o.y; r.heading = o.heading; r.g = o.g; r.f = o.g; r.vmin_sqr = o.vmin_sqr;
            r.curv = o.curv; r.bin = o.bin; r.ci = a.expand ? o.ci : -1; r.cj = a.expand ? o.cj : -1;
            a.out[(size_t)k * stride + cnt] = r;
            cnt++;
        }
        a.n_out[k] = cnt;
        a.flags[k] = (s.vmin_sqr < 1.0f) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(128) pp_apf_kernel(const __grid_constant__ PPConsts C, PPGroup G,
                                                      const float* xyh, int n, float* out)
{
    const int k = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (k >= n) return;
    PPWarpDev w;
    float v = pp_apf_sum(w, C, G.apf, (const int*)0, G.K, xyh[3 * k], xyh[3 * k + 1], xyh[3 * k + 2]);
    if (w.lane() == 0) out[k] = v;
}

// mode 0: successor lookup (truncated index, Grid3D.cpp:56-59) -> free flag + cell;
// mode 1: path-point lookup (rounded index, Grid3D.cpp:83-90) -> free flag
__global__ void pp_collision_kernel(const __grid_constant__ PPConsts C, const float* map, const float* pts, int stride,
                                    int n, int mode, int* free_out, int* cells)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float x = pts[(size_t)k * stride], y = pts[(size_t)k * stride + 1];
    if (mode == 0)
    {
        int ci, cj;
        bool ok = pp_collision_free(C, map, x, y, ci, cj);
        free_out[k] = ok ? 1 : 0;
        if (cells) { cells[2 * k] = ci; cells[2 * k + 1] = cj; }
    }
    else free_out[k] = pp_path_point_blocked(C, map, x, y) ? 0 : 1;
}

// Generic vehicle-footprint collision check (core/pp_footprint.h; north_star (c)).  One warp per pose, warps stride over the
// batch.  Per pose: base cell by truncation like the reference's successor check (Grid3D.cpp:53-54), heading bin -> offset
// list, lanes = footprint cells, verdict = `no lane saw a blocked cell` (ballot; the blocked count by a warp reduce on request).
// free_out[k] = 1 / 0, cells[2k..] = base cell, hits[k] = blocked footprint cells.
//   STAGED = false (default): every lane gathers its cells straight from the map.  The offset list is sorted (di, dj), so
//     32 consecutive entries cover 2-3 rows of the rectangle = a handful of 32-byte sectors per warp load, and L1 plays the
//     staging buffer.  No map cell is read twice for one pose, so there is nothing for shared memory to reuse.
//   STAGED = true (PP_B200_FOOT_STAGED=1, kept for the A/B record in profiles/): the bounding window of the list is first
//     staged into the warp's shared-memory tile with row-contiguous loads (cells outside the grid staged as +inf), then tested
//     from the tile.  Measured 4x more instructions per pose (index arithmetic of the staging loop; the window holds ~1.7x the
//     cells the footprint needs) and issue-bound -- see DESIGN.md section 12.
#define PP_FOOT_WARPS 4
struct PPFootArgs
{
    PPConsts         C;
    const float*     map;
    const float*     xyh;      // n x (x, y, heading), grid frame
    int              n;
    const PPFootBin* bins;     // C.bins + 1
    const PPCellOff* offs;
    const int*       lin;      // per offset: di * N + dj (the interior fast path adds it to the pose's cell address)
    int              win;      // tile side (>= every bin's bounding box), STAGED only
    int*             free_out;
    int*             cells;    // optional
    int*             hits;     // optional
};

template <bool STAGED>
__global__ void __launch_bounds__(PP_FOOT_WARPS * 32) pp_footprint_kernel(const __grid_constant__ PPFootArgs a)
{
    extern __shared__ float pp_foot_tiles[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = a.C.N;
    const float inf = __int_as_float(0x7f800000);
    const float thr = a.C.log_thr;
    // a warp takes 32 poses at a time: lane l does the exact index arithmetic of pose base + l once (IEEE division,
    // double-precision heading index -- the expensive, bit-exact part), then the warp walks the 32 footprints together
    for (int base = (blockIdx.x * PP_FOOT_WARPS + warp) * 32; base < a.n; base += gridDim.x * PP_FOOT_WARPS * 32)
    {
        const int k = base + lane;
        int ci = 0, cj = 0, bin = 0;
        if (k < a.n)
        {
            const float x = a.xyh[3 * (size_t)k], y = a.xyh[3 * (size_t)k + 1], h = a.xyh[3 * (size_t)k + 2];
            ci = (int)(x / a.C.res); cj = (int)(y / a.C.res);
            bin = pp_foot_bin(h, a.C.precision, a.C.bins);
        }
        const int cnt = min(32, a.n - base);
        int my_hits = 0;
        for (int p = 0; p < cnt; p++)
        {
            const int pci = __shfl_sync(0xffffffffu, ci, p), pcj = __shfl_sync(0xffffffffu, cj, p);
            const PPFootBin B = a.bins[__shfl_sync(0xffffffffu, bin, p)];
            int blocked = 0;
            if (STAGED)
            {
                float* tile = pp_foot_tiles + (size_t)warp * a.win * a.win;
                const int rows = B.imax - B.imin + 1, cols = B.jmax - B.jmin + 1;
                const int i0 = pci + B.imin, j0 = pcj + B.jmin;
                for (int t = lane; t < rows * cols; t += 32)
                {
                    int r = t / cols, c = t - r * cols;
                    int gi = i0 + r, gj = j0 + c;
                    bool inside = (gi > -1) && (gi < N) && (gj > -1) && (gj < N);
                    tile[r * a.win + c] = inside ? a.map[(size_t)gi * N + gj] : inf;
                }
                __syncwarp();
                for (int t = lane; t < B.count; t += 32)
                {
                    const PPCellOff o = a.offs[B.first + t];
                    float v = tile[(o.di - B.imin) * a.win + (o.dj - B.jmin)];
                    if (!(v < thr)) blocked++;
                }
                __syncwarp();       // the tile is restaged for the next pose
            }
            else
            {
                // the rectangle's bounding box lies inside the grid for all but border poses: no per-cell bounds test then,
                // one precomputed linear offset per cell
                const bool interior = (pci + B.imin > -1) && (pci + B.imax < N) && (pcj + B.jmin > -1) && (pcj + B.jmax < N);
                if (interior)
                {
                    // 32-bit cell indices (N * N < 2^31): one IMAD.WIDE per address instead of 64-bit pointer arithmetic
                    const int cell0 = pci * N + pcj;
                    const float* __restrict__ map = a.map;
                    const int* __restrict__ lin = a.lin;
                    // warp-uniform trip count (a per-lane bound makes the unrolled loop and its remainders diverge)
                    for (int t0 = B.first; t0 < B.first + B.count; t0 += 32)
                    {
                        const int t = t0 + lane;
                        if (t < B.first + B.count && !(__ldg(map + (cell0 + lin[t])) < thr)) blocked++;
                    }
                }
                else
                {
                    const PPCellOff* __restrict__ offs = a.offs + B.first;
                    for (int t = lane; t < B.count; t += 32)
                    {
                        const PPCellOff o = offs[t];
                        const int gi = pci + o.di, gj = pcj + o.dj;
                        const bool inside = ((unsigned)gi < (unsigned)N) && ((unsigned)gj < (unsigned)N);
                        const float v = inside ? __ldg(a.map + (size_t)gi * N + gj) : inf;
                        if (!(v < thr)) blocked++;
                    }
                }
            }
            // verdict by ballot; the count only when the caller asked for it (one REDUX)
            int total = a.hits ? __reduce_add_sync(0xffffffffu, blocked) : (__ballot_sync(0xffffffffu, blocked != 0) ? 1 : 0);
            if (lane == p) my_hits = total;
        }
        if (k < a.n)
        {
            a.free_out[k] = my_hits ? 0 : 1;
            if (a.cells) { a.cells[2 * (size_t)k] = ci; a.cells[2 * (size_t)k + 1] = cj; }
            if (a.hits) a.hits[k] = my_hits;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// velocity profile / trajectory (core/pp_velocity.h)
static_assert(sizeof(PPTrajPt) == sizeof(PPPathPt), "PPTrajPt mirrors PPPathPt");

// VelocityGenerator::generate_velocity_profile for n independent paths, one thread each (the passes are sequential scans).
// paths_xy [n][cap][2] and curvature [n][cap] in the reference's order (goal -> start); flags bit 0 = coast_to_goal, bit 1 =
// stop_at_goal; velocity [n][cap] out, v2 [n][cap] scratch.
__global__ void __launch_bounds__(128) pp_velocity_profile_kernel(PPVelLimits L, const float* paths_xy, const float* curvature,
                                                                   const int* counts, int n, int cap, const float* vel_init,
                                                                   const float* vcap, const int* flags, float* velocity, float* v2,
                                                                   int* feasible)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int m = counts[k];
    if (m < 1) { feasible[k] = 0; return; }
    if (m > cap) m = cap;
    const float* xy = paths_xy + (size_t)k * cap * 2;
    int fl = flags ? flags[k] : 0;
    bool ok = pp_velocity_profile(L, vel_init[k], vcap ? vcap[k] : FLT_MAX, xy, xy + 1, 2, curvature + (size_t)k * cap, 1, m,
                                  v2 + (size_t)k * cap, velocity + (size_t)k * cap, (fl & 1) != 0, (fl & 2) != 0);
    feasible[k] = ok ? 1 : 0;
}

struct PPTrajArgs
{
    PPVelLimits         L;
    const PPResult*     results;
    const PPPathPt*     paths;      int path_cap;
    const PPQuery*      queries;
    const PPWorldFrame* frames;     // per group
    const float*        vel_init;   // per query
    const float*        vcap;       // per query or nullptr (no cap)
    const int*          stop;       // per query or nullptr
    int                 n;
    float*              traj;       // [n][4 * path_cap]
    float*              tmp;        // [n][2 * path_cap]
    int*                n_samples;
    int*                feasible;
};

__global__ void __launch_bounds__(128) pp_trajectory_kernel(const __grid_constant__ PPTrajArgs a)
{
    PPWarpDev w;
    const int warps = blockDim.x >> 5;
    for (int q = blockIdx.x * warps + (threadIdx.x >> 5); q < a.n; q += gridDim.x * warps)
    {
        const PPResult r = a.results[q];
        int n = 0;
        if (r.success)
            n = pp_trajectory_assemble(w, a.frames[a.queries[q].group], a.L, reinterpret_cast<const PPTrajPt*>(a.paths + (size_t)q * a.path_cap),
                                       r.n_dubins, r.n_chain, a.path_cap, a.vel_init[q], a.vcap ? a.vcap[q] : FLT_MAX,
                                       a.stop ? a.stop[q] != 0 : false, a.traj + (size_t)q * 4 * a.path_cap,
                                       a.tmp + (size_t)q * 2 * a.path_cap, a.tmp + (size_t)q * 2 * a.path_cap + a.path_cap, a.feasible + q);
        else if (w.lane() == 0) a.feasible[q] = 0;
        if (w.lane() == 0) a.n_samples[q] = n;
    }
}

// 4 lanes per state: one CSC candidate each, folded in candidate order (Dubins.cpp:36-68)
// FP32 SIMT Dubins heuristic (north_star (d)), one thread per start: the flavour the K-POP mode evaluates per node
__global__ void __launch_bounds__(256) pp_dubins_length_fp32_kernel(const __grid_constant__ PPConsts C, const float* starts, int n,
                                                                     float gx, float gy, float gh, float* length)
{
    PPFrame F; F.goal_x = gx; F.goal_y = gy; F.goal_h = gh; F.goal_ci = F.goal_cj = F.goal_bin = 0;
    PPDubinsGoal gc;
    float sg, cg;
    pp_fm_sincos(gh, sg, cg);
    gc.grx = gx + C.r_min * sg; gc.gry = gy - C.r_min * cg;
    gc.glx = gx - C.r_min * sg; gc.gly = gy + C.r_min * cg;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        length[k] = pp_kdubins(C, F, gc, starts[3 * k], starts[3 * k + 1], starts[3 * k + 2]);
}

__global__ void __launch_bounds__(128) pp_dubins_length_kernel(float r_min, const float* starts, int n, float gx, float gy,
                                                                float gh, float* length, int* type, float* params4)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int k = t >> 2, cand = t & 3;
    bool active = k < n;
    float sx = 0, sy = 0, sh = 0;
    if (active) { sx = starts[3 * k]; sy = starts[3 * k + 1]; sh = starts[3 * k + 2]; }
    PPDubinsCenters c;
    float p[4] = {0, 0, 0, 0};
    float len = 0.0f;
    if (active)
    {
        pp_dubins_centers(r_min, sx, sy, sh, gx, gy, gh, c);
        float csx, csy, cgx, cgy;
        pp_dubins_pick(c, cand, csx, csy, cgx, cgy);
        len = pp_dubins_candidate(cand, r_min, sh, gh, csx, csy, cgx, cgy, p);
    }
    // fold across the 4 lanes of the group
    int base = (threadIdx.x & 31) & ~3;
    float best = __shfl_sync(0xffffffffu, len, base);
    int best_t = 0;
    for (int q = 1; q < 4; q++)
    {
        float l = __shfl_sync(0xffffffffu, len, base + q);
        if (l < best) { best = l; best_t = q; }
    }
    float bp[4];
    for (int q = 0; q < 4; q++) bp[q] = __shfl_sync(0xffffffffu, p[q], base + best_t);
    if (active && cand == 0)
    {
        length[k] = best;
        if (type) type[k] = best_t;
        if (params4) for (int q = 0; q < 4; q++) params4[4 * k + q] = bp[q];
    }
}

// one warp: Dubins::get_shortest_path (Dubins.cpp:125-153)
__global__ void __launch_bounds__(32) pp_dubins_path_kernel(const __grid_constant__ PPConsts C, float sx, float sy, float sh,
                                                             float gx, float gy, float gh, PPPathPt* out, int cap, int* n_out,
                                                             float* length, int* flag)
{
    const int lane = threadIdx.x & 31;
    int type; float p[4]; PPDubinsCenters cen; PPDubinsPlan pl;
    float len = pp_dubins_shortest(C.r_min, sx, sy, sh, gx, gy, gh, type, p, cen);
    pp_dubins_plan(C.r_min, C.step, C.ang_step, type, p, cen, pl);
    int total = pl.size_3 + 1;
    float acc = p[0];
    for (int k = 0; k < total; k++)
    {
        if (k == pl.size_1) acc = 0.0f;
        if (k == pl.size_2) acc = p[2];
        if ((k & 31) == lane && k < cap)
        {
            PPPathPt q;
            pp_dubins_sample(pl, C.r_min, k, acc, q.x, q.y, q.heading, q.curvature);
            out[k] = q;
        }
        if (k < pl.size_1) acc = (pl.s1 < 0) ? acc - C.ang_step : acc + C.ang_step;
        else if (k < pl.size_2) acc = acc + C.step;
        else if (k < pl.size_3) acc = (pl.s2 < 0) ? acc - C.ang_step : acc + C.ang_step;
    }
    if (lane == 0) { *n_out = total; *length = len; *flag = (fabsf(p[1]) > (float)PP_PI_2) ? 1 : 0; }
}

// lazy cached 2D A* on a fresh cache, cells queried in sequence (AStar.cpp:100-113); slot 0 scratch
__global__ void __launch_bounds__(32) pp_lazy_astar_kernel(const __grid_constant__ PPBatchArgs a, PPGroup G, const int* ij,
                                                            int n, float* out, int* status, int restart, unsigned* sid_io)
{
    PPWork wk;
    pp_slot_work(a, 0, wk);
    const int lane = threadIdx.x & 31;
    if (restart)      // AStar::reset() / fresh planner: nothing visited, node costs back to g = 0, f = h
        for (int c = lane; c < a.C.N * a.C.N; c += 32) wk.cell_state[c] = 0u;
    __syncwarp();
    if (lane == 0)
    {
        PPLazy L;
        PPLazyNb nbs[8];
        L.open.init(wk.open2, wk.open2_cap);
        L.search_id = restart ? 0u : *sid_io; L.status = 0; L.n_searches = 0; L.n_pops = 0;
        for (int k = 0; k < n; k++) out[k] = pp_lazy_astar(a.C, G.map, G.frame, wk, L, ij[2 * k], ij[2 * k + 1], nbs);
        *status = L.status;
        *sid_io = L.search_id;
    }
}

// ---------------------------------------------------------------------------------------------------
// map update

// Grid2D::update_obstacles(), Grid2D.cpp:197-208.  n4 = N*N/4 float4 elements (+ scalar tail).
__global__ void __launch_bounds__(256) pp_map_decay_kernel(float* __restrict__ map, size_t n, float log_free, float lo, float hi)
{
    size_t n4 = n >> 2;
    float4* m4 = reinterpret_cast<float4*>(map);
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    {
        float4 v = m4[i];
        v.x = pp_map_decay_cell(v.x, log_free, lo, hi);
        v.y = pp_map_decay_cell(v.y, log_free, lo, hi);
        v.z = pp_map_decay_cell(v.z, log_free, lo, hi);
        v.w = pp_map_decay_cell(v.w, log_free, lo, hi);
        m4[i] = v;
    }
    for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        map[i] = pp_map_decay_cell(map[i], log_free, lo, hi);
}

// one box as the device sees it: Grid2D.cpp:104-123's per-box prologue (done on the host with the host libm, like the reference)
// plus the bounding rectangle of its samples' cells, clipped to the grid (lo > hi: nothing to draw)
struct PPBoxDescDev { int start_i, start_j, ni, nj; float delta; int pad; short lo_i, hi_i, lo_j, hi_j; };
static_assert(sizeof(PPBoxDescDev) == 32, "box descriptor layout");

// Grid2D::update_obstacles(boxes, conf) (Grid2D.cpp:99-139), optionally followed by Grid2D::update_obstacles() (the whole-map
// decay, Grid2D.cpp:197-208), in ONE pass over the map: gather form, one CTA per 32x32 tile, 4 cells per thread.
// Every CTA bins the boxes itself -- it scans the n descriptors' bounding rectangles (index order, 256 per step, ballot
// compaction keeps the order the reference applies them in) -- so the host sends the descriptors and nothing else.
// A tile that no box touches streams through the decay (or exits untouched when there is none).
__global__ void __launch_bounds__(256) pp_map_update_kernel(float* __restrict__ map, int N, const PPBoxDescDev* __restrict__ descs,
                                                             int n_boxes, float cos_h, float sin_h, float lo, float hi,
                                                             int do_decay, float log_free)
{
    __shared__ int s_list[256];
    __shared__ int s_warp[8];
    __shared__ PPBoxDescDev sd[64];
    const int ti = blockIdx.y, tj = blockIdx.x;
    const int t_lo_i = ti * PP_TILE, t_hi_i = t_lo_i + PP_TILE - 1, t_lo_j = tj * PP_TILE, t_hi_j = t_lo_j + PP_TILE - 1;
    const int lj = threadIdx.x & 31, li0 = threadIdx.x >> 5;   // 8 rows per pass, 4 passes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cj = t_lo_j + lj;
    float v[4];
    int   ci[4];
    bool  loaded = false;
#pragma unroll
    for (int r = 0; r < 4; r++) { ci[r] = t_lo_i + li0 + 8 * r; v[r] = 0.0f; }
    for (int base = 0; base < n_boxes; base += 256)
    {
        const int k = base + threadIdx.x;
        bool hit = false;
        if (k < n_boxes)
        {
            // bounding rectangles only (8 bytes of the 32-byte record)
            const short4 bb = *reinterpret_cast<const short4*>(&descs[k].lo_i);
            hit = bb.x <= bb.y && bb.z <= bb.w && bb.x <= t_hi_i && bb.y >= t_lo_i && bb.z <= t_hi_j && bb.w >= t_lo_j;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { const int c = s_warp[q]; if (q < warp) off += c; total += c; }
        if (hit) s_list[off + __popc(m & ((1u << lane) - 1u))] = k;
        __syncthreads();
        if (total > 0 && !loaded)
        {
            loaded = true;
#pragma unroll
            for (int r = 0; r < 4; r++) v[r] = (ci[r] < N && cj < N) ? map[(size_t)ci[r] * N + cj] : 0.0f;
        }
        for (int b0 = 0; b0 < total; b0 += 64)
        {
            const int nb = min(64, total - b0);
            if (threadIdx.x < nb) sd[threadIdx.x] = descs[s_list[b0 + threadIdx.x]];
            __syncthreads();
            for (int b = 0; b < nb; b++)
            {
                const PPBoxDescDev d = sd[b];
                if (cj < d.lo_j || cj > d.hi_j) continue;
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    if (ci[r] < d.lo_i || ci[r] > d.hi_i) continue;
                    const int cnt = pp_box_count(d.ni, d.nj, cos_h, sin_h, ci[r] - d.start_i, cj - d.start_j);
                    if (cnt) v[r] = pp_box_apply(v[r], cnt, d.delta, lo, hi);
                }
            }
            __syncthreads();
        }
    }
    if (!loaded && !do_decay) return;
    if (!loaded)
    {
#pragma unroll
        for (int r = 0; r < 4; r++) v[r] = (ci[r] < N && cj < N) ? map[(size_t)ci[r] * N + cj] : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
        if (ci[r] < N && cj < N)
            map[(size_t)ci[r] * N + cj] = do_decay ? pp_map_decay_cell(v[r], log_free, lo, hi) : v[r];
}

// Grid2D::update_obstacles(lines, conf, width), Grid2D.cpp:142-194, one CTA per 32x32 tile.  Lines are applied one after the
// other (clamping is order dependent across lines); inside a line every sample adds the same delta, so a tile counts the
// samples of the line that fall on each of its cells (shared-memory counters) and applies clamp(v + delta) `count` times.
// Every CTA walks the line list itself and skips the lines whose sample rectangle misses its tile.
__global__ void __launch_bounds__(256) pp_map_lines_kernel(float* __restrict__ map, int N, int n45, int n2, float res,
                                                            const PPLineDesc* __restrict__ lines, int n_lines, float width,
                                                            float lo, float hi)
{
    __shared__ int   s_cnt[PP_TILE * PP_TILE];
    __shared__ int   s_any;
    const int ti = blockIdx.y, tj = blockIdx.x;
    const int t_lo_i = ti * PP_TILE, t_lo_j = tj * PP_TILE;
    const int lj = threadIdx.x & 31, li0 = threadIdx.x >> 5;
    const int cj = t_lo_j + lj;
    float v[4];
    bool loaded = false, dirty = false;
#pragma unroll
    for (int r = 0; r < 4; r++) v[r] = 0.0f;
    for (int k = 0; k < n_lines; k++)
    {
        const PPLineDesc d = lines[k];
        // conservative rectangle of the line's samples in cells (NaN anywhere: every comparison fails, the line is skipped -- it
        // draws nothing in the reference either, see pp_line_cells)
        const float reach = fminf(d.length, 100.0f * res) + res;
        const float ex = d.sx + d.ux * reach, ey = d.sy + d.uy * reach;
        const float wx = fabsf(d.nx) * (width + res), wy = fabsf(d.ny) * (width + res);
        const float x_lo = fminf(d.sx, ex) - wx, x_hi = fmaxf(d.sx, ex) + wx, y_lo = fminf(d.sy, ey) - wy, y_hi = fmaxf(d.sy, ey) + wy;
        const float ci_lo = x_lo / res + (float)n45 - 2.0f, ci_hi = x_hi / res + (float)n45 + 2.0f;
        const float cj_lo = y_lo / res + (float)n2 - 2.0f, cj_hi = y_hi / res + (float)n2 + 2.0f;
        const bool overlap = ci_lo <= (float)(t_lo_i + PP_TILE - 1) && ci_hi >= (float)t_lo_i &&
                             cj_lo <= (float)(t_lo_j + PP_TILE - 1) && cj_hi >= (float)t_lo_j;
        if (!overlap) continue;                       // uniform over the CTA
        for (int q = threadIdx.x; q < PP_TILE * PP_TILE; q += 256) s_cnt[q] = 0;
        if (threadIdx.x == 0) s_any = 0;
        __syncthreads();
        const int t = threadIdx.x;
        if (t < 100)
        {
            const float pl = pp_accumulate_steps(res, t);
            if (pl <= d.length)
                for (float pw = 0.0f; pw <= width; pw += res)
                {
                    int i1, j1, i2, j2;
                    pp_line_cells(d, res, n45, n2, pl, pw, i1, j1, i2, j2);
                    i1 -= t_lo_i; j1 -= t_lo_j; i2 -= t_lo_i; j2 -= t_lo_j;
                    if (i1 > -1 && i1 < PP_TILE && j1 > -1 && j1 < PP_TILE && i1 + t_lo_i < N && j1 + t_lo_j < N)
                    { atomicAdd(&s_cnt[i1 * PP_TILE + j1], 1); s_any = 1; }
                    if (i2 > -1 && i2 < PP_TILE && j2 > -1 && j2 < PP_TILE && i2 + t_lo_i < N && j2 + t_lo_j < N)
                    { atomicAdd(&s_cnt[i2 * PP_TILE + j2], 1); s_any = 1; }
                }
        }
        __syncthreads();
        if (s_any)
        {
            if (!loaded)
            {
                loaded = true;
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    const int ci = t_lo_i + li0 + 8 * r;
                    v[r] = (ci < N && cj < N) ? map[(size_t)ci * N + cj] : 0.0f;
                }
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                const int q = s_cnt[(li0 + 8 * r) * PP_TILE + lj];
                if (q) { v[r] = pp_box_apply(v[r], q, d.delta, lo, hi); dirty = true; }
            }
        }
        __syncthreads();
    }
    if (dirty)
    {
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            const int ci = t_lo_i + li0 + 8 * r;
            if (ci < N && cj < N) map[(size_t)ci * N + cj] = v[r];
        }
    }
}

// Grid3D::relocate_obstacles, Grid3D.cpp:169-203
__global__ void pp_map_reloc_fill_kernel(int* idx, size_t n)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) idx[i] = -1;
}

__global__ void pp_map_reloc_scatter_kernel(int* idx, int N, PPRelocDesc d)
{
    size_t n = (size_t)N * N, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += stride)
    {
        int i = (int)(s / N), j = (int)(s - (size_t)i * N);
        int in, jn;
        pp_reloc_target(d, i, j, in, jn);
        if (in > -1 && in < N && jn > -1 && jn < N) atomicMax(&idx[(size_t)in * N + jn], (int)s);
    }
}

__global__ void pp_map_reloc_gather_kernel(const float* __restrict__ old_map, float* __restrict__ new_map, int* idx, size_t n)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
    {
        int s = idx[t];
        new_map[t] = (s >= 0) ? old_map[s] : 0.0f;
        idx[t] = 0;     // leave the scratch zeroed for the line rasteriser
    }
}

#endif
