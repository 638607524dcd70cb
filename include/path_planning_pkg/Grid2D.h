// path_planning_pkg API surface, B200 build: the goal-centred log-odds occupancy grid (reference:
// include/path_planning_pkg/Grid2D.h:14-61, lib/Grid2D.cpp).  The map lives in HBM inside a pp_context; the three
// update_obstacles overloads run the rasteriser / decay kernels.  get_obstacle_map() downloads a host mirror.  The
// Node2D accessors (get_neighbors, update_costs, set_start_node*) exist for API compatibility with code that drives a
// 2D search by hand; they work on a host-side node table and the mirrored map and are not used by the device search.
#ifndef PP_B200_API_GRID2D_H
#define PP_B200_API_GRID2D_H

#include <memory>
#include <utility>
#include <vector>
#include "Node2D.h"
#include "Obstacle.h"
#include "common.h"

namespace planning
{
    namespace detail { struct Backend; }

    template <typename T> class Grid2D
    {
    public:
        Grid2D(T resolution, T obstacle_threshold, T obstacle_prob_min, T obstacle_prob_max, T obstacle_prob_free, int grid_size,
               Vector2D<T> goal, Vector2D<T> start, bool allow_diag_moves);
        Grid2D(T resolution, T obstacle_threshold, T obstacle_prob_min, T obstacle_prob_max, T obstacle_prob_free, int grid_size,
               bool allow_diag_moves);
        virtual ~Grid2D();

        // ---- map updates: all three run on the device map (pp_update_obstacles_* of include/pp_b200.h) ----
        // boxes -> pp_map_boxes_kernel (Grid2D.cpp:99-139), lane lines -> pp_map_lines_kernel (Grid2D.cpp:142-194),
        // no argument -> whole-map decay, pp_map_decay_kernel (Grid2D.cpp:197-208); clear_obstacles zeroes the map.
        void update_obstacles(const std::vector<Obstacle<T>>& obstacles, const std::vector<T>& confidence);
        void update_obstacles(const std::vector<std::pair<Vector2D<T>, Vector2D<T>>>& lines, const std::vector<T>& confidence,
                              const T line_width);
        void update_obstacles();
        void clear_obstacles();

        // ---- frame ----
        Node2D<T> update_goal_heading(const Vector2D<T>& goal, const Vector2D<T>& start);   // Grid2D.cpp:260-266
        T get_grid_heading() const;
        T get_grid_resolution() const;
        int get_grid_size() const;
        const std::vector<std::vector<T>>& get_obstacle_map() const;                       // downloads the host mirror when stale

        // ---- host-side Node2D table, for callers that drive a 2D search by hand (AStar does not need it here) ----
        void get_neighbors(const int xd, const int yd, std::vector<std::pair<Node2D<T>*, T>>& neighbors);
        void update_costs(const T total_cost, const Node2D<T>& last_node);
        T get_node_total_cost(const int i, const int j) const;
        Node2D<T> set_start_node(const Vector2D<T>& start);
        Node2D<T> set_start_node_grid(const int i, const int j);

        // B200 extension: the shared native backend (pp_context wrapper)
        std::shared_ptr<detail::Backend> backend() const { return _backend; }

    protected:
        explicit Grid2D(std::shared_ptr<detail::Backend> backend);
        void init_host_tables();
        void refresh_mirror() const;

        std::shared_ptr<detail::Backend> _backend;
        std::vector<std::vector<Node2D<T>>> _node_map;               // host-side node table (API compatibility only)
        mutable std::vector<std::vector<T>> _obstacle_map;           // host mirror of the device map
        std::vector<std::pair<int, int>> _actions;
        std::vector<T> _actions_cost;
    };
}

#endif
