// path_planning_pkg API surface, B200 build: the planner facade src/local_planner.cpp talks to.
// Same constructor and public methods as the reference class (reference: include/path_planning_pkg/HybridAStar.h:27-75,
// lib/HybridAStar.cpp:7-88); the object is a thin host handle over one pp_context (include/pp_b200.h): the log-odds
// map, the APF list and all search state live in B200 HBM and every call below runs CUDA kernels.  T = float is the
// production type (src/local_planner.cpp:509) and is bit-compatible with the C ABI; T = double converts at the
// boundary and computes in FP32 on the device.  Not thread-safe, like the reference.
#ifndef PP_B200_API_HYBRID_ASTAR_H
#define PP_B200_API_HYBRID_ASTAR_H

#include <algorithm>
#include <memory>
#include <utility>
#include <vector>
#include "Grid3D.h"
#include "Dubins.h"
#include "AStar.h"
#include "common.h"

namespace planning
{
    // tan of the largest steering angle (reference HybridAStar.h:21-25)
    template <typename T> T tan_max(const std::vector<T>& vect) { return std::tan(*std::max_element(vect.begin(), vect.end())); }

    template <typename T> class HybridAStar
    {
    public:
        HybridAStar(int dubins_shot_interval, int dubins_shot_interval_decay, T grid_resolution, T obstacle_threshold,
                    T obstacle_prob_min, T obstacle_prob_max, T obstacle_prob_free, int grid_size, bool grid_2d_allow_diag_moves,
                    T step_size, T max_lat_acc, T max_long_dec, T wheelbase, T rear_to_cg, T apf_rep_constant,
                    T apf_active_angle, int num_angle_bins, int num_actions, const std::vector<T>& steering,
                    const std::vector<T>& curvature_weights);
        ~HybridAStar();
        HybridAStar(const HybridAStar&) = delete;
        HybridAStar& operator=(const HybridAStar&) = delete;

        // boxes -> log-odds map + APF list (reference lib/HybridAStar.cpp:29-33)
        void update_obstacles(const std::vector<Obstacle<T>>& obstacles, const std::vector<T>& confidence, const T apf_added_radius);
        // lane lines -> log-odds map (reference lib/HybridAStar.cpp:36-40)
        void update_obstacles(const std::vector<std::pair<Vector2D<T>, Vector2D<T>>>& lines, const std::vector<T>& confidence,
                              const T line_width);
        // whole-map decay towards "free" (reference lib/HybridAStar.cpp:43-46)
        void update_obstacles();
        // drops the visited flags of the planner's 2D heuristic cache and, like the reference, keeps the node costs (reference
        // lib/HybridAStar.cpp:49-52 -> lib/AStar.cpp:56-60; C ABI pp_reset on the carried cache, pp_set_history)
        void reset();
        // new goal / grid heading, obstacles relocated into the new frame (reference lib/HybridAStar.cpp:55-59; pp_update_goal)
        void update_goal(const Vector3D<T>& goal, const Vector3D<T>& start);
        const std::vector<std::vector<T>>& get_obstacles() const;
        // {cost, success}; on success appends the path (goal -> start order, world frame) and its curvature; on failure
        // returns {numeric_limits<T>::max(), false} and leaves the vectors untouched (reference lib/HybridAStar.cpp:68-88).
        // Successive calls share the 2D heuristic cache exactly like successive calls on one reference object:
        // whole sessions return the reference's results, not only the first query (PP_B200_HISTORY=0 turns that off).
        std::pair<T, bool> find_path(const T vel_init, const Vector3D<T>& start, std::vector<Vector3D<T>>& path,
                                     std::vector<T>& curvature);

        // B200 extensions (not in the reference): expansions of the last find_path, and the raw C-ABI context
        int last_expansions() const;
        void* native_context() const;

    private:
        struct Impl;
        std::unique_ptr<Impl> _impl;
    };
}

#endif
