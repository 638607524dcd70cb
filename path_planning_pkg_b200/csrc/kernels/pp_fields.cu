// Heuristic FIELD kernels (north_star (d), (e); BASELINE config C3): whole-grid fields for one goal, used by the
// throughput modes instead of the reference's lazily evaluated per-query heuristics.
//
//   pp_field2d_*          2D holonomic-with-obstacles distance field: block-tiled Bellman relaxation (fast iterative
//                         method) in ONE cooperative launch of persistent CTAs.  A CTA pulls a 32x32 tile + halo of the
//                         field into shared memory, relaxes it to a local fixed point, writes it back; tiles whose
//                         neighbourhood did not change are skipped; global sweeps (separated by a grid barrier) repeat
//                         until nothing changes.  A distance is stored as the exact pair
//                         (#straight steps, #diagonal steps) of its path, so there is no accumulation error: the
//                         result equals a double-precision Dijkstra to one rounding (it is NOT the reference's
//                         AStar::find_path value, which is order dependent and inadmissible -- SURVEY F4).
//   pp_dubins_field_kernel Dubins length for every (i, j, heading bin) state of the grid, FP32 SIMT, fused with
//                         max(h2d, .) and written once (N*N*bins floats).
//
// This translation unit is compiled WITH fused multiply-add (default -fmad=true): its outputs carry a tolerance
// (1e-5 relative, north_star), not bit-exactness, and FFMA doubles the FP32 pipe rate.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <math.h>
#include <float.h>

#include "../core/pp_defs.h"
#include "../core/pp_dubins.h"
#include "pp_fields.h"

namespace cg = cooperative_groups;

#define F2D_TILE 32
#define F2D_UNREACHED 0xffffffffu
#define F2D_MAX_OWN 1024          /* tiles one persistent CTA may own (the launcher keeps T*T / grid below it) */

// value of a packed (a << 16 | b) path descriptor: a straight steps, b diagonal steps
__device__ __forceinline__ double f2d_value(unsigned ab, double c1, double c2)
{
    return (double)(ab >> 16) * c1 + (double)(ab & 0xffffu) * c2;
}

// One tile of one global sweep (all F2D_TILE x F2D_TILE threads of the CTA).  Returns (to every thread) whether the tile changed.
__device__ __forceinline__ int f2d_relax_tile(const float* __restrict__ map, unsigned* __restrict__ field, int ti, int tj, int N,
                                              float log_thr, double c1, double c2, int allow_diag,
                                              unsigned (*s)[F2D_TILE + 2])
{
    const int li = threadIdx.y, lj = threadIdx.x;           // lj fastest = grid j (contiguous in memory)
    const int gi = ti * F2D_TILE + li, gj = tj * F2D_TILE + lj;
    // halo reads may race with the neighbouring tiles' write-back of the same sweep: values only ever decrease towards the same
    // fixed point, so a stale read costs at most another sweep (volatile: always from L2, never a stale L1 line)
    auto load = [&](int i, int j) -> unsigned
    { return (i >= 0 && i < N && j >= 0 && j < N) ? *(volatile const unsigned*)&field[(size_t)i * N + j] : F2D_UNREACHED; };
    s[li + 1][lj + 1] = load(gi, gj);
    if (li == 0) s[0][lj + 1] = load(gi - 1, gj);
    if (li == F2D_TILE - 1) s[F2D_TILE + 1][lj + 1] = load(gi + 1, gj);
    if (lj == 0) s[li + 1][0] = load(gi, gj - 1);
    if (lj == F2D_TILE - 1) s[li + 1][F2D_TILE + 1] = load(gi, gj + 1);
    if (li == 0 && lj == 0) s[0][0] = load(gi - 1, gj - 1);
    if (li == 0 && lj == F2D_TILE - 1) s[0][F2D_TILE + 1] = load(gi - 1, gj + 1);
    if (li == F2D_TILE - 1 && lj == 0) s[F2D_TILE + 1][0] = load(gi + 1, gj - 1);
    if (li == F2D_TILE - 1 && lj == F2D_TILE - 1) s[F2D_TILE + 1][F2D_TILE + 1] = load(gi + 1, gj + 1);
    const bool inb = gi < N && gj < N;
    // a cell takes part when it is free; the source keeps its 0 regardless
    const bool free_cell = inb && (map[(size_t)gi * N + gj] < log_thr);
    unsigned cur = s[li + 1][lj + 1];
    const unsigned start_val = cur;
    __syncthreads();

    for (int it = 0; it < 4 * F2D_TILE; it++)
    {
        int changed = 0;
        if (free_cell)
        {
            unsigned best = cur;
            double bv = (best == F2D_UNREACHED) ? 1e300 : f2d_value(best, c1, c2);
#pragma unroll
            for (int di = -1; di <= 1; di++)
#pragma unroll
                for (int dj = -1; dj <= 1; dj++)
                {
                    if (di == 0 && dj == 0) continue;
                    const bool diag = (di != 0) && (dj != 0);
                    if (diag && !allow_diag) continue;
                    unsigned nb = s[li + 1 + di][lj + 1 + dj];
                    if (nb == F2D_UNREACHED) continue;
                    unsigned cand = nb + (diag ? 1u : 0x10000u);
                    double cv = f2d_value(cand, c1, c2);
                    if (cv < bv) { bv = cv; best = cand; }
                }
            if (best != cur) { cur = best; changed = 1; }
        }
        __syncthreads();                       // everyone has read the old neighbourhood
        if (changed) s[li + 1][lj + 1] = cur;
        if (!__syncthreads_or(changed)) break;
    }
    const int tile_changed = __syncthreads_or(cur != start_val);
    if (cur != start_val) field[(size_t)gi * N + gj] = cur;
    return tile_changed;
}

// The whole field in ONE cooperative launch: persistent CTAs (one per SM slot, all co-resident) sweep the active tiles, meet at a
// grid barrier, and go on until a sweep changes nothing -- no kernel launch and no host round trip per sweep (the first version
// needed 68 launches and 17 host read-backs for a 2048^2 map).
//   flags: 2 x T*T bytes (tiles that changed in the previous / this sweep); ctl: [0..2] change flags of three consecutive sweeps,
//   [3] = number of sweeps done (read back by the host afterwards).
__global__ void __launch_bounds__(F2D_TILE * F2D_TILE)
pp_field2d_persistent_kernel(const float* __restrict__ map, unsigned* __restrict__ field, unsigned char* __restrict__ flags,
                             int* __restrict__ ctl, float* __restrict__ out, int N, int T, float log_thr, double c1, double c2,
                             int allow_diag, int goal_i, int goal_j, int max_sweeps)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned s[F2D_TILE + 2][F2D_TILE + 2];
    __shared__ unsigned char s_run[F2D_MAX_OWN];
    const int tid = threadIdx.y * F2D_TILE + threadIdx.x;
    const size_t n = (size_t)N * N, gstride = (size_t)gridDim.x * (F2D_TILE * F2D_TILE);
    // the goal cell is the source even when it is marked occupied (the reference never tests the start cell of a search)
    const size_t goal = (size_t)goal_i * N + goal_j;
    const int goal_tile = (goal_i / F2D_TILE) * T + goal_j / F2D_TILE;
    for (size_t c = (size_t)blockIdx.x * (F2D_TILE * F2D_TILE) + tid; c < n; c += gstride) field[c] = (c == goal) ? 0u : F2D_UNREACHED;
    for (size_t t = (size_t)blockIdx.x * (F2D_TILE * F2D_TILE) + tid; t < (size_t)2 * T * T; t += gstride) flags[t] = (t == (size_t)goal_tile) ? 1 : 0;
    if (blockIdx.x == 0 && tid < 4) ctl[tid] = 0;
    grid.sync();

    int sweep = 0;
    for (; sweep < max_sweeps; sweep++)
    {
        const unsigned char* fa = flags + (size_t)(sweep & 1) * T * T;
        unsigned char* fb = flags + (size_t)((sweep + 1) & 1) * T * T;
        if (blockIdx.x == 0 && tid == 0) ctl[(sweep + 1) % 3] = 0;           // written from the next sweep on, last read two barriers ago
        // which of this CTA's tiles (blockIdx, blockIdx + grid, ...) run in this sweep: a tile runs only when it or one of its 8
        // neighbours changed in the previous sweep.  One thread per tile decides (and clears the flag of the tiles that sit out).
        for (int q = tid; q < F2D_MAX_OWN; q += F2D_TILE * F2D_TILE)
        {
            const int t = blockIdx.x + q * gridDim.x;
            int r = 0;
            if (t < T * T)
            {
                const int ti = t / T, tj = t - ti * T;
                for (int di = -1; di <= 1; di++)
                    for (int dj = -1; dj <= 1; dj++)
                    {
                        int a = ti + di, b = tj + dj;
                        if (a >= 0 && a < T && b >= 0 && b < T) r |= fa[a * T + b];
                    }
                if (!r) fb[t] = 0;
            }
            s_run[q] = (unsigned char)r;
        }
        __syncthreads();
        for (int q = 0, t = blockIdx.x; t < T * T; q++, t += gridDim.x)
        {
            if (!s_run[q]) continue;                  // uniform: shared memory
            const int ti = t / T, tj = t - ti * T;
            const int changed = f2d_relax_tile(map, field, ti, tj, N, log_thr, c1, c2, allow_diag, s);
            if (tid == 0)
            {
                fb[t] = changed ? 1 : 0;       // the tile's own CTA is the only writer of its flag
                if (changed) ctl[sweep % 3] = 1;
            }
            __syncthreads();
        }
        __threadfence();
        grid.sync();
        if (*(volatile int*)&ctl[sweep % 3] == 0) { sweep++; break; }
    }
    if (blockIdx.x == 0 && tid == 0) ctl[3] = sweep;
    for (size_t c = (size_t)blockIdx.x * (F2D_TILE * F2D_TILE) + tid; c < n; c += gstride)
    {
        unsigned ab = field[c];
        out[c] = (ab == F2D_UNREACHED) ? FLT_MAX : (float)f2d_value(ab, c1, c2);
    }
}

int pp_launch_field2d(cudaStream_t stream, const float* map, int N, float log_thr, float cost_straight, float cost_diag, int allow_diag,
                      int goal_i, int goal_j, unsigned* work, unsigned char* tile_flags, int* d_ctl, float* out, int sm_count,
                      int* sweeps_out, unsigned long long* launches)
{
    int T = (N + F2D_TILE - 1) / F2D_TILE;
    double c1 = (double)cost_straight, c2 = (double)cost_diag;
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp_field2d_persistent_kernel, F2D_TILE * F2D_TILE, 0);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return (int)cudaErrorLaunchOutOfResources;
    int blocks = occ * sm_count;                    // all CTAs co-resident: the grid barrier needs that
    if (blocks > T * T) blocks = T * T;
    if ((T * T + blocks - 1) / blocks > F2D_MAX_OWN) return (int)cudaErrorInvalidValue;      // grids beyond ~12 000 cells per side
    int max_sweeps = 16 * T + 64;
    void* args[] = {(void*)&map, (void*)&work, (void*)&tile_flags, (void*)&d_ctl, (void*)&out, (void*)&N, (void*)&T, (void*)&log_thr,
                    (void*)&c1, (void*)&c2, (void*)&allow_diag, (void*)&goal_i, (void*)&goal_j, (void*)&max_sweeps};
    e = cudaLaunchCooperativeKernel((const void*)pp_field2d_persistent_kernel, dim3(blocks), dim3(F2D_TILE, F2D_TILE), args, 0, stream);
    if (e != cudaSuccess) return (int)e;
    *launches += 1;
    int h_sweeps = 0;
    e = cudaMemcpyAsync(&h_sweeps, d_ctl + 3, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return (int)e;
    *sweeps_out = h_sweeps;
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Dubins field sweep, FP32 SIMT (north_star (d)).  Everything is single precision and branch-light: the lengths carry a
// 1e-5 relative tolerance (tests/test_gpu_fields.py), not bit-exactness, so this translation unit may use FFMA, the
// approximate divide, and the closed form of the tangent segment.
//   atan2 : one divide (the octant picks the quotient directly) + a degree-4 odd polynomial in t^2
//   acos  : sqrt reduction + degree-5 polynomial (|x| > 0.5), NaN outside [-1, 1] (circles closer than 2r: the RSL / LSR
//           candidate never wins the fold, Dubins.cpp:42-66)
//   RSL / LSR straight segment: sqrt(D^2 - 4 r^2) instead of the distance between the two tangent points the reference
//           builds with four sin / cos calls (Dubins.cpp:232-236, :281-285) -- the same number to ~1e-7 relative
#define FLD_PI    3.14159274101257324219f
#define FLD_PI_2  1.57079637050628662109f
#define FLD_PI_4  0.78539818525314331055f
#define FLD_2PI   6.28318548202514648438f

__device__ __forceinline__ float fld_atan2(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    float y0, num, den;
    if (ay > 2.414213562373095f * ax) { y0 = FLD_PI_2; num = -ax; den = ay; }
    else if (ay > 0.4142135623730950f * ax) { y0 = FLD_PI_4; num = ay - ax; den = ay + ax; }
    else { y0 = 0.0f; num = ay; den = ax; }
    const float t = (den == 0.0f) ? 0.0f : __fdividef(num, den);
    const float z = t * t;
    float a = y0 + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * t + t);
    if (x < 0.0f) a = FLD_PI - a;
    return (y < 0.0f) ? -a : a;
}

__device__ __forceinline__ float fld_asin_core(float a)
{
    const float z = a * a;
    return ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * a + a;
}

__device__ __forceinline__ float fld_acos(float x)       // NaN for |x| > 1 through the sqrt of a negative number
{
    if (x > 0.5f) return 2.0f * fld_asin_core(sqrtf(0.5f * (1.0f - x)));
    if (x < -0.5f) return FLD_PI - 2.0f * fld_asin_core(sqrtf(0.5f * (1.0f + x)));
    return FLD_PI_2 - fld_asin_core(x);
}

// shortest of RSR, RSL, LSR, LSL in the reference's fold order; (ssn, scs) = sin / cos of the start heading
__device__ __forceinline__ float fld_dubins(float r, float sx, float sy, float sh, float ssn, float scs, float gh,
                                            float grx, float gry, float glx, float gly)
{
    const float srx = sx + r * ssn, sry = sy - r * scs, slx = sx - r * ssn, sly = sy + r * scs;
    const float four_r2 = 4.0f * r * r, two_r = 2.0f * r;
    // RSR (Dubins.cpp:180-206): both arcs clockwise
    float dx = grx - srx, dy = gry - sry;
    float d2 = dx * dx + dy * dy, th = fld_atan2(dy, dx);
    float p1 = th - sh; if (p1 > 0.0f) p1 -= FLD_2PI;
    float p3 = gh - th; if (p3 > 0.0f) p3 -= FLD_2PI;
    float best = sqrtf(d2) - r * (p1 + p3);
    // RSL (Dubins.cpp:209-255)
    dx = glx - srx; dy = gly - sry;
    d2 = dx * dx + dy * dy; th = fld_atan2(dy, dx);
    {
        const float ac = fld_acos(two_r / sqrtf(d2));          // IEEE divide + sqrt: acos amplifies argument error near 1
        const float t1 = ac + th, p2 = t1 - FLD_PI;
        p1 = t1 - (FLD_PI_2 + sh); if (p1 > 0.0f) p1 -= FLD_2PI;
        p3 = (gh - FLD_PI_2) - p2; if (p3 < 0.0f) p3 += FLD_2PI;
        const float len = sqrtf(d2 - four_r2) + r * (p3 - p1);
        if (len < best) best = len;                       // NaN (circles overlap) never wins
    }
    // LSR (Dubins.cpp:258-304)
    dx = grx - slx; dy = gry - sly;
    d2 = dx * dx + dy * dy; th = fld_atan2(dy, dx);
    {
        const float ac = fld_acos(two_r / sqrtf(d2));          // IEEE divide + sqrt: acos amplifies argument error near 1
        const float t1 = th - ac, p2 = t1 + FLD_PI;
        p1 = t1 - (sh - FLD_PI_2); if (p1 < 0.0f) p1 += FLD_2PI;
        p3 = (gh + FLD_PI_2) - p2; if (p3 > 0.0f) p3 -= FLD_2PI;
        const float len = sqrtf(d2 - four_r2) + r * (p1 - p3);
        if (len < best) best = len;
    }
    // LSL (Dubins.cpp:307-323): both arcs counter-clockwise
    dx = glx - slx; dy = gly - sly;
    d2 = dx * dx + dy * dy; th = fld_atan2(dy, dx);
    p1 = th - sh; if (p1 < 0.0f) p1 += FLD_2PI;
    p3 = gh - th; if (p3 < 0.0f) p3 += FLD_2PI;
    {
        const float len = sqrtf(d2) + r * (p1 + p3);
        if (len < best) best = len;
    }
    return best;
}

// out[(i*N + j)*bins + b] = max(h2d[i*N + j], dubins((i*res, j*res, -pi + b*precision) -> goal)).
// One thread per (cell-in-group, bin); a CTA covers cpb = blockDim.x / bins consecutive cells per step, so consecutive
// threads write consecutive floats; (i, j) advance incrementally (no division in the loop).
__global__ void __launch_bounds__(320)
pp_dubins_field_kernel(const float* __restrict__ h2d, float* __restrict__ out, int N, int bins, int cpb, float res, float precision,
                       float r_min, float goal_x, float goal_y, float goal_h)
{
    __shared__ float s_sin[PP_MAX_BINS], s_cos[PP_MAX_BINS], s_head[PP_MAX_BINS];
    for (int b = threadIdx.x; b < bins; b += blockDim.x)
    {
        float h = (float)(-PP_PI + (double)(b * precision));
        s_head[b] = h;
        sincosf(h, &s_sin[b], &s_cos[b]);
    }
    float gs, gcs;
    sincosf(goal_h, &gs, &gcs);
    const float grx = goal_x + r_min * gs, gry = goal_y - r_min * gcs, glx = goal_x - r_min * gs, gly = goal_y + r_min * gcs;
    __syncthreads();
    const int c = threadIdx.x / bins, b = threadIdx.x - c * bins;
    if (c >= cpb) return;
    const float sh = s_head[b], ssn = s_sin[b], scs = s_cos[b];
    const long long n_cells = (long long)N * N, stride = (long long)gridDim.x * cpb;
    const int stride_i = (int)(stride / N), stride_j = (int)(stride % N);
    long long cell = (long long)blockIdx.x * cpb + c;
    int i = (int)(cell / N), j = (int)(cell % N);
    for (; cell < n_cells; cell += stride)
    {
        const float len = fld_dubins(r_min, i * res, j * res, sh, ssn, scs, goal_h, grx, gry, glx, gly);
        const float h1 = h2d ? h2d[cell] : 0.0f;
        out[cell * bins + b] = (h1 < len) ? len : h1;
        i += stride_i; j += stride_j;
        if (j >= N) { j -= N; i++; }
    }
}

int pp_launch_dubins_field(cudaStream_t stream, const float* h2d, float* out, int N, int bins, float res, float precision, float r_min,
                           float goal_x, float goal_y, float goal_h, int sm_count, unsigned long long* launches)
{
    int cpb = 288 / bins; if (cpb < 1) cpb = 1;                       // 72 bins: 4 cells = 288 threads = 9 warps per step
    int threads = ((cpb * bins + 31) / 32) * 32;
    if (threads > 320) { cpb = 320 / bins; if (cpb < 1) cpb = 1; threads = ((cpb * bins + 31) / 32) * 32; }
    long long groups = ((long long)N * N + cpb - 1) / cpb;
    long long blocks = (long long)sm_count * 6;
    if (blocks > groups) blocks = groups;
    pp_dubins_field_kernel<<<(int)blocks, threads, 0, stream>>>(h2d, out, N, bins, cpb, res, precision, r_min, goal_x, goal_y, goal_h);
    *launches += 1;
    return (int)cudaGetLastError();
}
