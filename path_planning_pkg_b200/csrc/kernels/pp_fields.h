// Launchers of the heuristic-field kernels (pp_fields.cu, compiled with FMA) for the C ABI (pp_cabi.cu).
#ifndef PP_FIELDS_H
#define PP_FIELDS_H

#include <cuda_runtime.h>

// 2D distance field from the goal cell, one cooperative launch.  work: N*N uint32; tile_flags: 2*T*T bytes (T = ceil(N/32));
// d_ctl: 4 ints.  Synchronises the stream.  Returns a cudaError_t as int; *sweeps_out = number of global relaxation sweeps.
int pp_launch_field2d(cudaStream_t stream, const float* map, int N, float log_thr, float cost_straight, float cost_diag, int allow_diag,
                      int goal_i, int goal_j, unsigned* work, unsigned char* tile_flags, int* d_ctl, float* out, int sm_count,
                      int* sweeps_out, unsigned long long* launches);

// out[(i*N + j)*bins + b] = max(h2d[cell] (may be null), Dubins length from (i*res, j*res, -pi + b*precision) to the goal)
int pp_launch_dubins_field(cudaStream_t stream, const float* h2d, float* out, int N, int bins, float res, float precision, float r_min,
                           float goal_x, float goal_y, float goal_h, int sm_count, unsigned long long* launches);

#endif
