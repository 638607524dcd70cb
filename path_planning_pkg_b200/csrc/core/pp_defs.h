// Shared definitions of the B200 local-planner hot path (device-side records and constants).
//
// Everything in csrc/core/ is written once as `PP_HD` (= __host__ __device__ under nvcc) code so
// that the per-query control logic (containers, lazy 2D A*, search loop) can ALSO be compiled by
// g++ inside the test-only harness tests/cpp/host_emul.cpp and debugged on a machine without a GPU.
// The shipped library (csrc/cabi/pp_cabi.cu -> lib/libpp_b200.so) only ever runs it inside CUDA
// kernels; there is no CPU execution path in the product.
#ifndef PP_DEFS_H
#define PP_DEFS_H

#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define PP_HD __host__ __device__ __forceinline__
#define PP_HD_NOINLINE __host__ __device__ __noinline__
#define PP_HD_NOINLINE_FN static __host__ __device__ __noinline__   /* free functions: one copy per translation unit */
#else
#define PP_HD inline
#define PP_HD_NOINLINE inline
#define PP_HD_NOINLINE_FN inline
#endif

// float -> int as the reference's x86-64 build does it (cvttss2si): NaN and values outside the int range give INT_MIN
// ("integer indefinite"), which every `> -1` bounds test of the reference then rejects.  CUDA's own conversion gives 0 for
// NaN and saturates, which would turn a NaN coordinate into a valid cell.
PP_HD int pp_f2i_x86(float v)
{
    return (v >= -2147483648.0f && v < 2147483648.0f) ? (int)v : (int)0x80000000;
}

// set bits / index of the lowest set bit (m != 0) of a ballot mask
PP_HD int pp_popc(unsigned m)
{
#ifdef __CUDA_ARCH__
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}
PP_HD int pp_ctz(unsigned m)
{
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

#define PP_MAX_STEER 16
#define PP_MAX_BINS 128   /* heading bins + 1 padding column must fit */

// query status bits
#define PP_OK 0
#ifndef PP_STATUS_OPEN_OVERFLOW       /* same values in include/pp_b200.h */
#define PP_STATUS_OPEN_OVERFLOW 1     /* 3D open-list node pool exhausted */
#define PP_STATUS_CLOSED_OVERFLOW 2   /* expansion cap (closed log) reached */
#define PP_STATUS_OPEN2D_OVERFLOW 4   /* 2D open-list node pool exhausted */
#define PP_STATUS_PATH_OVERFLOW 8     /* path / dubins sample buffer too small */
#define PP_STATUS_NULL_TERMINAL 16    /* reference would dereference a null _prev */
#define PP_STATUS_ARENA_EXHAUSTED 32  /* a container had to grow and the context's arena had no block left */
#endif

// Constants derived on the host exactly as the reference constructors derive them
// (Grid2D.cpp:7-20, VehicleModel.cpp:7-47, HybridAStar.cpp:7-24) and uploaded once per context.
struct PPConsts
{
    // Grid2D
    int   N;               // _grid_size
    int   n2;              // _grid_size_2   = round(N*0.5)
    int   n45;             // _grid_size_4_5 = round(N*0.8)
    float res;             // _resolution
    float log_thr, log_min, log_max, log_free;
    int   n_act2d;         // 8 or 4
    int   act_di[8], act_dj[8];
    float act_cost[8];     // res*sqrt(di^2+dj^2)
    // VehicleModel
    int   S;               // number of steering primitives
    int   A;               // _num_actions
    int   bins;            // num_angle_bins
    float ts, max_lat_acc, max_lat_acc_sqr;
    float precision;       // float(2*M_PI/bins)
    float abs_curv[PP_MAX_STEER];
    float act_cost3d[PP_MAX_STEER];
    float off_heading[PP_MAX_STEER];
    // Grid3D
    float apf_k, apf_alpha;
    // Dubins
    float r_min, step, ang_step;
    // HybridAStar
    int   shot_interval, shot_decay;
};

// Per (map, goal) frame: set by update_goal (Grid3D.cpp:102-124) on the host.
struct PPFrame
{
    float goal_x, goal_y, goal_h;   // goal pose in the grid frame (_goal_node._pose2D)
    int   goal_ci, goal_cj, goal_bin;
};

// One Hybrid A* state (Node3D.h:17-25) without the two raw pointers.
struct PPState
{
    float x, y, heading;
    float g, f;
    float vmin_sqr;
    int   curv;     // _curvature_index
    int   bin;      // _angle_bin (may equal `bins`)
    int   ci, cj;   // base node cell, -1 = none
};

// One expanded node, written to the optional pop trace.
struct PPPop
{
    int   ci, cj, bin;
    float x, y, heading;
    float g, f;
};

// One query of a batch.
struct PPQuery
{
    PPState start;      // start node in the grid frame (host: set_start_node, Grid3D.cpp:127-160)
    int     group;      // index of the (map, goal) group the query runs on
    int     pad;
};

// Fixed-size result record of one query.
struct PPResult
{
    int   success;
    int   status;
    float cost;
    int   n_pops;          // expansions (get_neighbors calls), HybridAStar.cpp:157
    int   n_pops_bin_oob;  // expansions whose bin == bins
    int   n_chain;         // parent-chain points (terminal node back to the start)
    int   n_dubins;        // Dubins-shot samples (0 when the goal cell was popped)
    int   n_lazy_searches; // 2D A* searches actually run (cache misses)
    int   n_lazy_pops;     // 2D A* pops over all lazy searches
    int   max_open;        // high-water mark of the 3D open list
    int   n_closed;        // distinct closed states
    int   pad;
};

#endif
