// Heuristic FIELD kernels (north_star (d), (e); BASELINE config C3): whole-grid fields for one goal, used by the
// throughput modes instead of the reference's lazily evaluated per-query heuristics.
//
//   pp_field2d_*          2D holonomic-with-obstacles distance field: block-tiled Bellman relaxation (fast iterative
//                         method).  Each CTA pulls a 32x32 tile + halo of the field into shared memory, relaxes it to
//                         a local fixed point, writes it back; tiles whose neighbourhood did not change are skipped;
//                         global sweeps repeat until nothing changes.  A distance is stored as the exact pair
//                         (#straight steps, #diagonal steps) of its path, so there is no accumulation error: the
//                         result equals a double-precision Dijkstra to one rounding (it is NOT the reference's
//                         AStar::find_path value, which is order dependent and inadmissible -- SURVEY F4).
//   pp_dubins_field_kernel Dubins length for every (i, j, heading bin) state of the grid, FP32 SIMT, fused with
//                         max(h2d, .) and written once (N*N*bins floats).
//
// This translation unit is compiled WITH fused multiply-add (default -fmad=true): its outputs carry a tolerance
// (1e-5 relative, north_star), not bit-exactness, and FFMA doubles the FP32 pipe rate.
#include <cuda_runtime.h>
#include <math.h>
#include <float.h>

#include "../core/pp_defs.h"
#include "../core/pp_dubins.h"
#include "pp_fields.h"

#define F2D_TILE 32
#define F2D_UNREACHED 0xffffffffu

// value of a packed (a << 16 | b) path descriptor: a straight steps, b diagonal steps
__device__ __forceinline__ double f2d_value(unsigned ab, double c1, double c2)
{
    return (double)(ab >> 16) * c1 + (double)(ab & 0xffffu) * c2;
}

__global__ void pp_field2d_init_kernel(unsigned* __restrict__ field, unsigned char* __restrict__ tile_active, int N, int T,
                                       int goal_i, int goal_j)
{
    // the goal cell is the source even when it is marked occupied (the reference never tests the start cell of a search)
    size_t n = (size_t)N * N, stride = (size_t)gridDim.x * blockDim.x;
    const size_t goal = (size_t)goal_i * N + goal_j, goal_tile = (size_t)(goal_i / F2D_TILE) * T + goal_j / F2D_TILE;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) field[c] = (c == goal) ? 0u : F2D_UNREACHED;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < (size_t)2 * T * T; t += stride)
        tile_active[t] = (t == goal_tile) ? 1 : 0;
}

// One global sweep.  active_in / active_out: per-tile flags of the previous / next sweep.
__global__ void __launch_bounds__(F2D_TILE * F2D_TILE)
pp_field2d_sweep_kernel(const float* __restrict__ map, unsigned* __restrict__ field, const unsigned char* __restrict__ active_in,
                        unsigned char* __restrict__ active_out, int* __restrict__ any_change, int N, int T, float log_thr,
                        double c1, double c2, int allow_diag)
{
    const int ti = blockIdx.y, tj = blockIdx.x;
    // run only when this tile or one of its 8 neighbours changed in the previous sweep
    __shared__ int run;
    if (threadIdx.x == 0 && threadIdx.y == 0)
    {
        int r = 0;
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++)
            {
                int a = ti + di, b = tj + dj;
                if (a >= 0 && a < T && b >= 0 && b < T) r |= active_in[a * T + b];
            }
        run = r;
    }
    __syncthreads();
    if (!run) return;

    __shared__ unsigned s[F2D_TILE + 2][F2D_TILE + 2];
    const int li = threadIdx.y, lj = threadIdx.x;           // lj fastest = grid j (contiguous in memory)
    const int gi = ti * F2D_TILE + li, gj = tj * F2D_TILE + lj;
    auto load = [&](int i, int j) -> unsigned { return (i >= 0 && i < N && j >= 0 && j < N) ? field[(size_t)i * N + j] : F2D_UNREACHED; };
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
    int tile_changed = __syncthreads_or(cur != start_val);
    if (cur != start_val) field[(size_t)gi * N + gj] = cur;
    if (tile_changed && threadIdx.x == 0 && threadIdx.y == 0)
    {
        active_out[ti * T + tj] = 1;
        *any_change = 1;
    }
}

__global__ void pp_field2d_finish_kernel(const unsigned* __restrict__ field, float* __restrict__ out, size_t n, double c1, double c2)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride)
    {
        unsigned ab = field[c];
        out[c] = (ab == F2D_UNREACHED) ? FLT_MAX : (float)f2d_value(ab, c1, c2);
    }
}

int pp_launch_field2d(cudaStream_t stream, const float* map, int N, float log_thr, float cost_straight, float cost_diag, int allow_diag,
                      int goal_i, int goal_j, unsigned* work, unsigned char* tile_flags, int* d_flag, float* out, int sm_count,
                      int* sweeps_out, unsigned long long* launches)
{
    const int T = (N + F2D_TILE - 1) / F2D_TILE;
    const double c1 = (double)cost_straight, c2 = (double)cost_diag;
    pp_field2d_init_kernel<<<sm_count * 4, 256, 0, stream>>>(work, tile_flags, N, T, goal_i, goal_j);
    *launches += 1;
    dim3 grid(T, T), block(F2D_TILE, F2D_TILE);
    int sweeps = 0, h_flag = 1;
    unsigned char* fa = tile_flags;
    unsigned char* fb = tile_flags + (size_t)T * T;
    while (h_flag && sweeps < 16 * T + 64)
    {
        cudaMemsetAsync(d_flag, 0, sizeof(int), stream);
        // a few sweeps per host round trip
        for (int k = 0; k < 4; k++)
        {
            cudaMemsetAsync(fb, 0, (size_t)T * T, stream);
            pp_field2d_sweep_kernel<<<grid, block, 0, stream>>>(map, work, fa, fb, d_flag, N, T, log_thr, c1, c2, allow_diag);
            *launches += 1;
            unsigned char* t = fa; fa = fb; fb = t;
            sweeps++;
        }
        cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream);
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return (int)e;
    }
    pp_field2d_finish_kernel<<<sm_count * 4, 256, 0, stream>>>(work, out, (size_t)N * N, c1, c2);
    *launches += 1;
    *sweeps_out = sweeps;
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// FP32 transcendentals of the field sweep: CUDA's single-precision functions (<= 2 ulp), not the FP64 "pinned libm"
__device__ __forceinline__ float fld_dubins(float r, float sx, float sy, float sh, float ssn, float scs, float gh,
                                            float grx, float gry, float glx, float gly)
{
    const float srx = sx + r * ssn, sry = sy - r * scs, slx = sx - r * ssn, sly = sy + r * scs;
    float best = 0.0f;
#pragma unroll
    for (int type = 0; type < 4; type++)
    {
        const bool s_right = (type == PP_RSR) || (type == PP_RSL), g_right = (type == PP_RSR) || (type == PP_LSR);
        const float csx = s_right ? srx : slx, csy = s_right ? sry : sly;
        const float cgx = g_right ? grx : glx, cgy = g_right ? gry : gly;
        const float theta = atan2f(cgy - csy, cgx - csx);
        float ac = 0.0f, c1 = 0.0f, s1 = 0.0f, c2 = 0.0f, s2 = 0.0f, p[4];
        if (type == PP_RSL || type == PP_LSR)
        {
            ac = acosf(pp_dubins_acos_arg(r, csx, csy, cgx, cgy));
            const float t1 = pp_dubins_theta_t1(type, ac, theta);
            const float p2 = pp_dubins_p2(type, t1);
            sincosf(t1, &s1, &c1);
            sincosf(p2, &s2, &c2);
        }
        const float len = pp_dubins_finish(type, r, sh, gh, csx, csy, cgx, cgy, theta, ac, c1, s1, c2, s2, p);
        if (type == 0 || len < best) best = len;      // NaN never wins (Dubins.cpp:42-66)
    }
    return best;
}

// out[(i*N + j)*bins + b] = max(h2d[i*N + j], dubins((i*res, j*res, -pi + b*precision) -> goal))
__global__ void __launch_bounds__(256)
pp_dubins_field_kernel(const float* __restrict__ h2d, float* __restrict__ out, int N, int bins, float res, float precision, float r_min,
                       float goal_x, float goal_y, float goal_h)
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
    const size_t total = (size_t)N * N * bins, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
    {
        const size_t cell = t / bins;
        const int b = (int)(t - cell * bins);
        const int i = (int)(cell / N), j = (int)(cell - (size_t)i * N);
        const float len = fld_dubins(r_min, i * res, j * res, s_head[b], s_sin[b], s_cos[b], goal_h, grx, gry, glx, gly);
        const float h1 = h2d ? h2d[cell] : 0.0f;
        out[t] = (h1 < len) ? len : h1;
    }
}

int pp_launch_dubins_field(cudaStream_t stream, const float* h2d, float* out, int N, int bins, float res, float precision, float r_min,
                           float goal_x, float goal_y, float goal_h, int sm_count, unsigned long long* launches)
{
    pp_dubins_field_kernel<<<sm_count * 8, 256, 0, stream>>>(h2d, out, N, bins, res, precision, r_min, goal_x, goal_y, goal_h);
    *launches += 1;
    return (int)cudaGetLastError();
}
