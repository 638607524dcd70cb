// path_planning_pkg API surface, B200 build: the 2D grid A* used as the holonomic-with-obstacles heuristic (reference:
// include/path_planning_pkg/AStar.h:15-58, lib/AStar.cpp).  find_path(i, j) is the lazily evaluated, cached,
// early-terminating search; here it runs in pp_lazy_astar_kernel on a cache that lives in HBM and
// persists across calls until reset().  Both constructor variants of the reference exist; the library is built like
// the reference's (CMakeLists.txt:128) with -DSTORE_GRID_AS_REFERENCE, i.e. AStar shares the caller's grid.
#ifndef PP_B200_API_ASTAR_H
#define PP_B200_API_ASTAR_H

#include <limits>
#include <memory>
#include <utility>
#include <vector>
#include "Grid2D.h"
#include "common.h"

namespace planning
{
    template <typename T> class AStar
    {
    public:
#ifndef STORE_GRID_AS_REFERENCE
        AStar(T grid_resolution, T obstacle_threshold, T obstacle_prob_min, T obstacle_prob_max, T obstacle_prob_free,
              int grid_size, bool grid_allow_diag_moves = true);
#else
        AStar(Grid2D<T>& grid);
#endif
        ~AStar();

        void update_goal_node(const Node2D<T>& goal_node);
        void update_goal_start(const Vector2D<T>& goal, const Vector2D<T>& start, Node2D<T>& start_node);
        void update_obstacles(const std::vector<Obstacle<T>>& obstacles, const std::vector<T>& confidence);
        void update_obstacles(const std::vector<std::pair<Vector2D<T>, Vector2D<T>>>& lines, const std::vector<T>& confidence,
                              const T line_width);
        void update_obstacles();
        void reset();
        const std::vector<std::vector<T>>& get_obstacles() const;
        // cost of the 8-connected path from the cell of `start` to the goal cell; max() when unreachable
        T find_path(const Vector2D<T>& goal, const Vector2D<T>& start, bool get_cost_only = true);
        T find_path(const int start_i, const int start_j);

    private:
        std::unique_ptr<Grid2D<T>> _owned;     // only in the by-value variant
        Grid2D<T>* _grid;
        bool _fresh;                           // the device cache must be restarted on the next query
    };
}

#endif
