/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * Plain-C ABI shared by the two CPU oracles:
 *   oracle/_ref/libref_oracle*.so  ("ref_" prefix): the UNMODIFIED reference sources under
 *       /root/reference/lib, compiled where they lie (oracle/Makefile) and driven by
 *       oracle/ref_driver.cpp through link-time --wrap hooks.
 *   oracle/libpp_oracle_port.so    ("port_" prefix): the CPU restatement in oracle/port/.
 * Both export the same function set so tests/ can run one check against either.
 */
#ifndef PP_ORACLE_API_H
#define PP_ORACLE_API_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_STEER 16

/* The 20 constructor arguments of HybridAStar<T> (HybridAStar.h:33-38), T = float. */
typedef struct orc_params
{
    int   shot_interval;
    int   shot_decay;
    float resolution;
    float obstacle_threshold;
    float prob_min;
    float prob_max;
    float prob_free;
    int   grid_size;
    int   allow_diag;
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
    float steering[ORC_MAX_STEER];
    float curvature_weights[ORC_MAX_STEER];
} orc_params;

/* One Node3D (Node3D.h:17-25) flattened: cell = base node indices, -1 when no base node. */
typedef struct orc_state
{
    float x, y, heading;
    float g, f;
    float vmin_sqr;
    int   curvature_index;
    int   angle_bin;
    int   ci, cj;
} orc_state;

/* One expanded node (one call of Grid3D::get_neighbors from the search loop, HybridAStar.cpp:157). */
typedef struct orc_pop
{
    int   ci, cj, bin;
    float x, y, heading;
    float g, f;
} orc_pop;

/* Derived constants, for cross-checking host code. */
typedef struct orc_consts
{
    float log_threshold, log_min, log_max, log_free;
    float grid_heading;
    float goal_world[3];
    float goal_grid[3];
    int   goal_bin;
    int   goal_ci, goal_cj;
    float precision;
    float r_min;
    float ang_step;
    int   num_apf;
} orc_consts;

typedef struct orc_result
{
    int   success;
    float cost;
    int   n_path;       /* number of path points written (reference order: goal -> start) */
    int   n_pops;       /* expansions = get_neighbors calls inside the search */
    int   n_pops_bin_oob; /* expansions whose angle_bin == num_angle_bins */
} orc_result;

#define ORC_DECL(prefix)                                                                              \
    void* prefix##_create(const orc_params* p);                                                        \
    void  prefix##_destroy(void* h);                                                                   \
    void  prefix##_update_goal(void* h, const float* goal3, const float* start3);                      \
    void  prefix##_reset(void* h);                                                                     \
    void  prefix##_scrub(void* h);                                                                     \
    void  prefix##_update_boxes(void* h, const float* boxes_xydxdy, const float* conf, int n,          \
                                float apf_added_radius);                                               \
    void  prefix##_update_boxes_2d(void* h, const float* boxes_xydxdy, const float* conf, int n);      \
    void  prefix##_update_lines(void* h, const float* lines_x1y1x2y2, const float* conf, int n,        \
                                float width);                                                          \
    void  prefix##_decay(void* h);                                                                     \
    void  prefix##_get_map(void* h, float* out_nn);                                                    \
    void  prefix##_set_map(void* h, const float* in_nn);                                               \
    void  prefix##_get_consts(void* h, orc_consts* out);                                               \
    void  prefix##_get_apf(void* h, float* out_xyr);                                                   \
    void  prefix##_get_tables(void* h, float* offset_xy, float* offset_heading, float* actions_cost,   \
                              float* abs_curv);                                                        \
    void  prefix##_set_start(void* h, const float* start3, orc_state* out);                            \
    int   prefix##_rollout(void* h, const orc_state* in, orc_state* out, int* n_out);                  \
    int   prefix##_expand(void* h, const orc_state* in, orc_state* out, int* n_out);                   \
    float prefix##_apf(void* h, float x, float y, float heading);                                      \
    int   prefix##_check_path(void* h, const float* xyh, int n);                                       \
    float prefix##_dubins_length(void* h, const float* start3, const float* goal3, int* type,          \
                                 float* params4);                                                      \
    int   prefix##_dubins_path(void* h, const float* start3, const float* goal3, float* xyh,           \
                               float* curv, int cap, float* length, int* flag);                        \
    float prefix##_astar_lazy(void* h, int i, int j);                                                  \
    void  prefix##_find_path(void* h, float vel, const float* start3, orc_result* res, float* path_xyh,\
                             float* curv, int path_cap, orc_pop* pops, int pop_cap);

ORC_DECL(ref)
ORC_DECL(port)

/* New semantics, stated by the port only (oracle/port/footprint.inc): generic vehicle-footprint collision check. */
int  port_footprint_table(void* h, int bin, float length, float width, float rear, short* offs_ij, int cap);
void port_footprint_check(void* h, const float* xyh, int n, float length, float width, float rear, int* free_out,
                          int* cells_ij, int* hits_out);

#ifdef __cplusplus
}
#endif
#endif
