/* pp_b200.h -- C ABI of the B200-native local-planner hot path (libpp_b200.so).
 *
 * This is the drop-in boundary below the reference's C++ class API.  The reference has no FFI: its
 * callers (src/local_planner.cpp:158-166, :204-205, :241, :287-288, :316 and utils/\*\/test_*.cpp)
 * use the C++ classes of include/path_planning_pkg/ directly.  The replacement keeps those classes
 * (include/path_planning_pkg/\*.h in this repo, implemented in path_planning_pkg_b200/csrc/host/)
 * as thin host-side handles and routes every per-cell / per-state / per-query computation through
 * the entry points below into hand-written sm_100a kernels.  Plain pointers and sizes only; every
 * function returns 0 on success or a negative pp_status, pp_last_error() gives the text.
 * There is NO CPU implementation behind this ABI: without a CUDA device pp_create() fails.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference repo).
 */
#ifndef PP_B200_H
#define PP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PP_API_MAX_STEER 16

enum pp_status_code
{
    PP_SUCCESS = 0,
    PP_ERR_INVALID = -1,      /* bad argument */
    PP_ERR_CUDA = -2,         /* CUDA runtime error, see pp_last_error() */
    PP_ERR_NO_DEVICE = -3,    /* no CUDA device: the library has no CPU path */
    PP_ERR_CAPACITY = -4      /* a per-query capacity was exceeded (reported per query in pp_result.status) */
};

/* The 20 constructor arguments of HybridAStar<T> (include/path_planning_pkg/HybridAStar.h:33-38). */
typedef struct pp_params
{
    int   shot_interval;            /* dubins_shot_interval */
    int   shot_decay;               /* dubins_shot_interval_decay */
    float resolution;               /* grid_resolution */
    float obstacle_threshold;
    float prob_min;
    float prob_max;
    float prob_free;
    int   grid_size;
    int   allow_diag;               /* grid_2d_allow_diag_moves */
    float step_size;
    float max_lat_acc;
    float max_long_dec;
    float wheelbase;
    float rear_to_cg;
    float apf_rep_constant;
    float apf_active_angle;
    int   num_angle_bins;
    int   num_actions;
    int   num_steering;
    float steering[PP_API_MAX_STEER];
    float curvature_weights[PP_API_MAX_STEER];
} pp_params;

/* Derived constants (what the reference constructors compute), for inspection and tests. */
typedef struct pp_consts_info
{
    float log_threshold, log_min, log_max, log_free;
    float precision, r_min, ang_step;
    int   n2, n45;
} pp_consts_info;

/* Frame of one (map, goal) group after pp_update_goal. */
typedef struct pp_frame_info
{
    float grid_heading;
    float goal_world[3];
    float goal_grid[3];
    int   goal_bin, goal_ci, goal_cj;
    int   num_apf;
} pp_frame_info;

/* Node3D (include/path_planning_pkg/Node3D.h:17-25) without the raw pointers. */
typedef struct pp_state
{
    float x, y, heading;
    float g, f;
    float vmin_sqr;
    int   curvature_index;
    int   angle_bin;
    int   ci, cj;
} pp_state;

typedef struct pp_pop
{
    int   ci, cj, bin;
    float x, y, heading;
    float g, f;
} pp_pop;

/* One query of a batch: HybridAStar::find_path(vel_init, start, ...) on planner `group`. */
typedef struct pp_query
{
    float x, y, heading;   /* start pose, world frame */
    float vel;             /* vel_init */
    int   group;
} pp_query;

/* pp_result.status bits: which per-query capacity was still exhausted after the automatic retries (0 = none) */
#ifndef PP_STATUS_OPEN_OVERFLOW
#define PP_STATUS_OPEN_OVERFLOW 1     /* open-list pool (EXACT) / queue (K-POP) exhausted */
#define PP_STATUS_CLOSED_OVERFLOW 2   /* expansion cap (EXACT closed log) / node log (K-POP) reached */
#define PP_STATUS_OPEN2D_OVERFLOW 4   /* 2D open-list pool of the lazy heuristic exhausted */
#define PP_STATUS_PATH_OVERFLOW 8     /* path_cap too small: the returned path is truncated */
#define PP_STATUS_NULL_TERMINAL 16    /* the reference would dereference a null _prev here (Dubins shot from the start node) */
#define PP_STATUS_ARENA_EXHAUSTED 32  /* a container had to grow and the context's memory arena had no block left */
#endif

typedef struct pp_result
{
    int   success;
    int   status;            /* 0 or PP_STATUS_* bits */
    float cost;
    int   n_pops;            /* node expansions */
    int   n_pops_bin_oob;    /* expansions in heading bin == num_angle_bins (undefined in the reference) */
    int   n_chain;
    int   n_dubins;
    int   n_lazy_searches;
    int   n_lazy_pops;
    int   max_open;
    int   n_closed;
    int   n_path;            /* points written to the path output (= n_dubins + n_chain) */
} pp_result;

typedef struct pp_search_opts
{
    /* PP_MODE_EXACT: hard caps of the per-query containers (the reference's are unbounded).  Every resident query starts on small
     * pools and a container that fills up moves into a block twice the size from the context's arena, up to these caps (0 = defaults:
     * 2^24 closed states, 2^23 open nodes, 2^20 2D open nodes).  A query that meets a cap is re-run with caps 8x larger, up to three
     * times, before pp_result.status reports it.  PP_MODE_KPOP: max_expansions sizes the fixed per-slot node log (0 = 2^17). */
    int max_expansions;      /* closed states per query */
    int max_open;            /* 3D open-list nodes per query */
    int max_open2d;          /* 2D open-list nodes per query (lazy heuristic) */
    int path_cap;            /* path points per query in the output arrays */
    int trace_cap;           /* pops recorded per query (0 = no trace) */
    int max_slots;           /* resident query slots (0 = auto from free memory) */
    int mode;                /* PP_MODE_EXACT (0, reference-identical single pop) or PP_MODE_KPOP (1) */
    int kpop;                /* pops per iteration in PP_MODE_KPOP, 1..32 (0 = 32) */
} pp_search_opts;

#define PP_MODE_EXACT 0
#define PP_MODE_KPOP 1

typedef struct pp_context pp_context;

const char* pp_last_error(void);
int  pp_device_count(void);

/* HybridAStar<T>::HybridAStar(...) (lib/HybridAStar.cpp:7-24): one context = shared parameters and
 * `num_groups` independent planner instances (map + goal frame + APF list each). */
int  pp_create(const pp_params* params, int device, int num_groups, pp_context** out);
void pp_destroy(pp_context* ctx);
/* A lane: a context that shares `parent`'s parameters, maps, goal frames and APF lists (read-only view) and owns its stream, search
 * scratch and batch buffers.  Batches on different lanes run concurrently on the device, so the drain of one batch (its few longest
 * queries, one warp each) overlaps the bulk of the next: continuous batching of HybridAStar::find_path calls
 * (lib/HybridAStar.cpp:68-88) that the single-threaded reference has no counterpart for.  Map updates go through the parent; a lane
 * sees them at its next pp_batch_upload.  Destroy lanes before their parent. */
int  pp_create_lane(pp_context* parent, pp_context** out);
/* Bytes of device memory the search scratch of this context may take (per-query pools + the arena they grow into);
 * 0 = 75 % of the memory free at first use.  Applies from the next (re)allocation. */
int  pp_set_memory_budget(pp_context* ctx, unsigned long long bytes);
int  pp_get_consts(pp_context* ctx, pp_consts_info* out);
int  pp_get_frame(pp_context* ctx, int group, pp_frame_info* out);
/* VehicleModel tables (lib/VehicleModel.cpp:31-40): offset_xy[S][bins][2], heading offsets, costs, |curvature| */
int  pp_get_tables(pp_context* ctx, float* offset_xy, float* offset_heading, float* actions_cost, float* abs_curv);

/* HybridAStar::update_goal (lib/HybridAStar.cpp:55-59 -> Grid3D::update_goal_heading, Grid3D.cpp:102-124,
 * including relocate_obstacles, Grid3D.cpp:169-203). */
int  pp_update_goal(pp_context* ctx, int group, const float* goal3, const float* start3);
/* Grid2D::update_goal_heading (lib/Grid2D.cpp:260-266): goal location + grid heading only, the map stays where it is (the
 * relocating form above is Grid3D's override). */
int  pp_update_goal_frame(pp_context* ctx, int group, const float* goal3, const float* start3);
/* HybridAStar::reset (lib/HybridAStar.cpp:49-52 -> AStar::reset, lib/AStar.cpp:56-60).  Without history (below) every query
 * starts on a fresh 2D heuristic cache and this is a no-op; with history it drops the carried cache's visited flags and, like
 * the reference, keeps the node costs. */
int  pp_reset(pp_context* ctx, int group);
/* Planner-object history.  One reference HybridAStar object keeps the `_visted` flags and `_node_map` costs of
 * its 2D heuristic (lib/AStar.cpp:100-113, lib/Grid2D.cpp:219-227) from one find_path call to the next: only reset() clears
 * the flags, nothing clears the costs, map updates invalidate neither.  enable = 1 gives `group` that object semantics: a
 * pp_find_path_batch of ONE query in PP_MODE_EXACT on this group continues on the cache the group's earlier single queries
 * left (starting from the freshly constructed state), so a sequence of update / find_path / reset calls returns what the same
 * sequence returns on one reference object.  Batches of several queries always run every query on a fresh cache.
 * enable = 0 drops the cache.  The C++ class HybridAStar<T> enables it for its planner. */
int  pp_set_history(pp_context* ctx, int group, int enable);
/* HybridAStar::update_obstacles(boxes, confidence, apf_added_radius) (lib/HybridAStar.cpp:29-33 ->
 * Grid3D.cpp:22-44 -> Grid2D.cpp:99-139).  boxes = n x (x, y, dx, dy), world frame. */
int  pp_update_obstacles_boxes(pp_context* ctx, int group, const float* boxes_xydxdy, const float* confidence,
                               int n, float apf_added_radius);
/* Grid2D::update_obstacles(boxes, confidence) only (no APF list rebuild), lib/Grid2D.cpp:99-139 */
int  pp_update_obstacles_boxes_2d(pp_context* ctx, int group, const float* boxes_xydxdy, const float* confidence, int n);
/* One round of the Grid2D map update in a single pass over the map: Grid2D::update_obstacles(boxes, confidence) immediately followed
 * by Grid2D::update_obstacles() (lib/Grid2D.cpp:99-139 then :197-208) -- what src/local_planner.cpp:241 and :288 do one after the
 * other; every cell is read and written once.  Asynchronous on the context's stream (pp_sync waits). */
int  pp_update_obstacles_boxes_2d_decay(pp_context* ctx, int group, const float* boxes_xydxdy, const float* confidence, int n);
/* HybridAStar::update_obstacles(lines, confidence, line_width) (lib/HybridAStar.cpp:36-40 -> Grid2D.cpp:142-194) */
int  pp_update_obstacles_lines(pp_context* ctx, int group, const float* lines_x1y1x2y2, const float* confidence,
                               int n, float line_width);
/* HybridAStar::update_obstacles() (lib/HybridAStar.cpp:43-46 -> Grid2D.cpp:197-208) */
int  pp_update_obstacles_decay(pp_context* ctx, int group);
/* HybridAStar::get_obstacles() (lib/HybridAStar.cpp:62-65): N*N floats, row = i (grid x) */
int  pp_map_download(pp_context* ctx, int group, float* out_nn);
int  pp_map_upload(pp_context* ctx, int group, const float* in_nn);
/* Grid3D::update_obstacles without the Grid2D rasterisation (lib/Grid3D.cpp:22-44): rebuilds the group's APF obstacle list only.
 * A rank that receives the map itself by pp_broadcast_maps calls this instead of pp_update_obstacles_boxes. */
int  pp_update_obstacles_apf(pp_context* ctx, int group, const float* boxes_xydxdy, int n, float apf_added_radius);
/* device pointer of a group's map; state derived from the map is dropped.  A caller that writes through it (its own collective)
 * calls pp_map_mark_dirty once the write has completed. */
void* pp_map_device_ptr(pp_context* ctx, int group);
int  pp_map_mark_dirty(pp_context* ctx, int group);
/* ---- map replication across the GPUs of one box (BASELINE north_star: "the map replicated by an NCCL broadcast over NVLink after
 * each update"; the reference is single-process: src/local_planner.cpp:241, :287-288 update the one map in place) ----
 * One communicator per context, one rank per GPU.  pp_comm_unique_id (rank 0) fills 128 bytes (ncclUniqueId) the caller hands to
 * every rank (torch.distributed / MPI / a file); pp_comm_init joins; pp_broadcast_maps replaces groups [first, first + n) on every
 * rank by root's maps, in place, on the context's stream (asynchronous: pp_sync waits).  NCCL is resolved with dlopen. */
int  pp_comm_unique_id(void* id128);
int  pp_comm_init(pp_context* ctx, int nranks, int rank, const void* id128);
int  pp_broadcast_maps(pp_context* ctx, int first_group, int n_groups, int root);
int  pp_comm_destroy(pp_context* ctx);
int  pp_sync(pp_context* ctx);

/* Grid3D::set_start_node (lib/Grid3D.cpp:127-160) + HybridAStar.cpp:73-74 for n queries (host side) */
int  pp_set_start_batch(pp_context* ctx, const pp_query* queries, int n, pp_state* out);

/* ---- stateless batched kernels (each directly comparable with one reference function) ---- */
/* VehicleModel::get_neighbors (lib/VehicleModel.cpp:63-105): out[n][2A+1], count per state, neglect flag */
int  pp_rollout_batch(pp_context* ctx, const pp_state* in, int n, pp_state* out, int* n_out, int* flags);
/* Grid3D::get_neighbors (lib/Grid3D.cpp:47-74): roll-out + bounds/collision lookup + APF cost */
int  pp_expand_batch(pp_context* ctx, int group, const pp_state* in, int n, pp_state* out, int* n_out, int* flags);
/* successor collision lookup alone (lib/Grid3D.cpp:56-59): free[k] = 1 when in bounds and below threshold; cells out */
int  pp_collision_batch(pp_context* ctx, int group, const float* xy, int n, int* free_out, int* cells_ij);
/* ---- generic vehicle-footprint collision check (north_star (c)) ----
 * The reference's check is the one-cell footprint: a pose is free iff map[int(x/res)][int(y/res)] is inside the grid and below
 * the threshold (lib/Grid3D.cpp:53-59); the vehicle's size is folded into the obstacles by the caller
 * (src/local_planner.cpp:230-231, :287).  pp_set_footprint builds, per heading bin, the cell offsets an oriented rectangle
 * [-rear_overhang, length - rear_overhang] x [-width/2, +width/2] (vehicle frame, x forward, origin = the pose) covers, sampled
 * every half cell like the reference's box rasteriser (lib/Grid2D.cpp:110-133); length = width = 0 gives {(0, 0)}, i.e. exactly
 * the reference's check.  pp_footprint_batch tests n grid-frame poses (x, y, heading): free_out[k] = 1 iff every footprint cell
 * around the pose's cell is inside the grid and below the threshold; cells_ij (optional) = the pose's cell as the reference
 * indexes it; hits_out (optional) = number of blocked footprint cells; kernel_ms (optional) = device time of the kernel alone. */
int  pp_set_footprint(pp_context* ctx, float length, float width, float rear_overhang);
int  pp_get_footprint(pp_context* ctx, int bin, int* count, short* offs_ij, int cap);
int  pp_footprint_batch(pp_context* ctx, int group, const float* xyh, int n, int* free_out, int* cells_ij, int* hits_out,
                        float* kernel_ms);
/* Grid3D::get_field_intensity (lib/Grid3D.cpp:206-227) for n poses (x, y, heading) */
int  pp_apf_batch(pp_context* ctx, int group, const float* xyh, int n, float* out);
/* Grid3D::check_path (lib/Grid3D.cpp:78-93): blocked[k] per point, returns collision-free flag in *free_out */
int  pp_check_path(pp_context* ctx, int group, const float* xyh, int n, int* free_out);
/* Dubins::get_shortest_path_length (lib/Dubins.cpp:19-69) for n starts and one goal */
int  pp_dubins_length_batch(pp_context* ctx, const float* starts_xyh, int n, const float* goal3, float* length,
                            int* type, float* params4);
/* the same length, every operation in FP32 (one thread per start; the heuristic flavour the K-POP mode uses): within
   1e-5 relative of Dubins::get_shortest_path_length (lib/Dubins.cpp:19-69) */
int  pp_dubins_length_fp32_batch(pp_context* ctx, const float* starts_xyh, int n, const float* goal3, float* length);
/* Dubins::get_shortest_path (lib/Dubins.cpp:125-153): returns sample count in *n_out */
int  pp_dubins_path(pp_context* ctx, const float* start3, const float* goal3, float* xyh, float* curvature, int cap,
                    int* n_out, float* length, int* long_turn_flag);
/* AStar::find_path(i, j) (lib/AStar.cpp:100-113) called in sequence on a fresh cache for n cells */
int  pp_astar_lazy_batch(pp_context* ctx, int group, const int* ij, int n, float* out);
/* same, continuing on the cache left by the previous call (AStar::find_path between two AStar::reset(), lib/AStar.cpp:56-60) */
int  pp_astar_lazy_continue(pp_context* ctx, int group, const int* ij, int n, float* out);
/* AStar::reset() (lib/AStar.cpp:56-60) on that cache: drops the visited flags, keeps the node costs */
int  pp_astar_lazy_reset(pp_context* ctx, int group);
/* Dubins::Dubins(r_min, step_size) (lib/Dubins.cpp:7-16) for a stand-alone Dubins<T> handle */
int  pp_override_dubins(pp_context* ctx, float r_min, float step_size);
/* Grid2D::clear_obstacles (lib/Grid2D.cpp:211-216) */
int  pp_clear_obstacles(pp_context* ctx, int group);

/* ---- heuristic fields (north_star (d), (e); BASELINE config C3) ---- */
/* 2D holonomic-with-obstacles distance-to-goal field over Grid2D (same metric as lib/Grid2D.cpp:32-58, :72-96: 8- or 4-connected,
 * step costs res*sqrt(di^2+dj^2), cells >= threshold blocked) by block-tiled Bellman relaxation; FLT_MAX = unreachable.  This is the
 * exact distance field, NOT the order-dependent value of AStar::find_path (lib/AStar.cpp:100-186).
 * out_nn may be NULL (result stays on the device for the throughput search modes). */
int  pp_heuristic_field_2d(pp_context* ctx, int group, float* out_nn, int* sweeps, float* kernel_ms);
/* h(i, j, b) = max(h2d(i, j), Dubins length from pose (i*res, j*res, -pi + b*precision) to the goal) for every state of the grid
 * (Dubins::get_shortest_path_length, lib/Dubins.cpp:19-69, FP32 SIMT); out_nnb = N*N*bins floats or NULL. */
int  pp_heuristic_field_3d(pp_context* ctx, int group, int use_h2d, float* out_nnb, float* kernel_ms);

/* ---- the search: HybridAStar::find_path (lib/HybridAStar.cpp:68-88) for a batch of queries ---- */
/* paths: n x path_cap x (x, y, heading) world frame in the reference's order (goal -> start);
 * curvature: n x path_cap; trace: n x trace_cap pops (may be NULL). */
int  pp_find_path_batch(pp_context* ctx, const pp_query* queries, int n, const pp_search_opts* opts,
                        pp_result* results, float* paths_xyh, float* curvature, pp_pop* trace);
/* Split form used by bench.py: upload once, run (timed on the device), fetch. */
int  pp_batch_upload(pp_context* ctx, const pp_query* queries, int n, const pp_search_opts* opts);
int  pp_batch_run(pp_context* ctx, float* kernel_ms);
/* pp_batch_run in two halves: enqueue the search on the context's stream and return; wait for it (and re-run the queries that met
 * a capacity, as pp_batch_run does).  kernel_ms = device time of all passes (CUDA events on the context's stream). */
int  pp_batch_run_async(pp_context* ctx);
int  pp_batch_wait(pp_context* ctx, float* kernel_ms);
int  pp_batch_fetch(pp_context* ctx, pp_result* results, float* paths_xyh, float* curvature, pp_pop* trace);
/* ---- the step after the path: velocity profile and trajectory message ---- */
/* The five constructor arguments of VelocityGenerator<T> (lib/VelocityGenerator.cpp:6-14). */
typedef struct pp_velocity_limits
{
    float max_velocity, coast_velocity, max_lat_acc, max_long_acc, max_long_dec;
} pp_velocity_limits;
/* VelocityGenerator::generate_velocity_profile (lib/VelocityGenerator.cpp:19-85) for n independent paths: paths_xy [n][cap][2] and
 * curvature [n][cap] in the reference's order (goal -> start), counts[k] points each (>= 1); vel_init[k]; max_velocity_curr may be
 * NULL (no cap); flags[k] bit 0 = coast_to_goal, bit 1 = stop_at_goal (NULL = 0).  velocity [n][cap] (index 0 = the path's start)
 * and the feasibility flag out.  Bit-identical to the reference except for the sign / payload of NaN results. */
int  pp_velocity_profile_batch(pp_context* ctx, const pp_velocity_limits* limits, const float* paths_xy, const float* curvature,
                               const int* counts, int n, int cap, const float* vel_init, const float* max_velocity_curr,
                               const int* flags, float* velocity, int* feasible);
/* For every query of the last pp_batch_run / pp_find_path_batch, on the device and straight from the device-resident path records:
 * HybridAStar::reconstruct_path (lib/HybridAStar.cpp:208-262) + generate_velocity_profile (vel_init = the query's, coast_to_goal =
 * false) + the layout LocalPlanner::publish_trajectory gives /local_planner/trajectory (src/local_planner.cpp:346-372):
 * traj[k] = x[0..m), y[0..m), heading[0..m) from start to goal, then velocity[0..m), m = n_samples[k] (0 for a failed query),
 * one row of 4 * path_cap floats per query.  max_velocity_curr / stop_at_goal per query may be NULL. */
int  pp_trajectory_batch(pp_context* ctx, const pp_velocity_limits* limits, const float* max_velocity_curr, const int* stop_at_goal,
                         float* traj, int* n_samples, int* feasible, float* kernel_ms);
/* number of CUDA kernels this context has launched so far (bench.py's gpu_launches) */
unsigned long long pp_kernel_launches(pp_context* ctx);
/* queries of the last pp_batch_run / pp_find_path_batch that exhausted a pool and were re-run with larger pools */
int  pp_batch_retried(pp_context* ctx);
/* CUDA-event bracket on the context's stream (for measuring the asynchronous map / field calls in between) */
int  pp_timer_begin(pp_context* ctx);
int  pp_timer_end(pp_context* ctx, float* ms);

#ifdef __cplusplus
}
#endif
#endif
