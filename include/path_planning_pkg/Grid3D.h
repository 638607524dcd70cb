// path_planning_pkg API surface, B200 build: Grid2D plus the heading dimension -- successor generation with collision
// lookup and APF cost, Dubins-path collision check, goal / start node construction (reference:
// include/path_planning_pkg/Grid3D.h:13-48, lib/Grid3D.cpp).  get_neighbors / check_path run pp_successor_kernel /
// pp_collision_kernel on the device map.
#ifndef PP_B200_API_GRID3D_H
#define PP_B200_API_GRID3D_H

#include <utility>
#include <vector>
#include "Grid2D.h"
#include "VehicleModel.h"
#include "common.h"

namespace planning
{
    template <typename T> class Grid3D : public Grid2D<T>
    {
    public:
        Grid3D(T resolution, T obstacle_threshold, T obstacle_prob_min, T obstacle_prob_max, T obstacle_prob_free, int grid_size,
               bool allow_diag_moves, T step_size, T max_lat_acc, T max_long_dec, T wheelbase, T rear_to_cg, T apf_rep_constant,
               T apf_active_angle, int num_angle_bins, int num_actions, const std::vector<T>& steering,
               const std::vector<T>& curvature_weights);

        using Grid2D<T>::update_obstacles;

        // Boxes: APF list rebuild on the host (Grid3D.cpp:22-44) + pp_update_obstacles_boxes -> pp_map_boxes_kernel.
        void update_obstacles(const std::vector<Obstacle<T>>& obstacles, const std::vector<T>& confidence, const T apf_added_radius);

        // Goal / start in the goal-centred frame.  update_goal_heading relocates the device map when the goal moves
        // (pp_update_goal -> pp_map_reloc_* kernels, Grid3D.cpp:102-124, :169-203); set_start_node maps a start outside the
        // grid silently to cell (0, 0) like the reference (Grid3D.cpp:127-160).
        Node3D<T> update_goal_heading(const Vector3D<T>& goal, const Vector3D<T>& start);
        Node3D<T> set_start_node(const Vector3D<T>& start);
        Vector3D<T> get_goal_location() const;

        // One popped state -> its successors: pp_expand_batch (n = 1) -> pp_successor_kernel, i.e. VehicleModel roll-out,
        // single-cell collision lookup and APF cost (Grid3D.cpp:47-74).  The return value is the roll-out's
        // "neglect acceleration" flag, as in the reference.
        bool get_neighbors(const Node3D<T>& node, std::vector<Node3D<T>>& neighbors) const;

        // Dubins-shot collision check, rounded-index lookups (Grid3D.cpp:78-93): pp_check_path -> pp_collision_kernel.
        bool check_path(const std::vector<Vector3D<T>>& path) const;

        const std::vector<T>& get_abs_curvatures() const;

    private:
        std::vector<T> _abs_curvatures;
    };
}

#endif
