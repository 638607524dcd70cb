// TEST DRIVER for the C++ drop-in API (include/path_planning_pkg/*.h + libpath_planning_b200.so), run on the GPU box
// by tests/test_gpu_cpp_api.py.  It replays the scenario of the reference's own smoke driver
// (utils/hybrid_astar/test_hybrid_astar.cpp:13-91: 60x60 @0.5 m, 4 lane lines + 3 boxes applied 5x with decay,
// start (18,18,pi/2) -> goal (26,36,0), v = 2) through the class API exactly as a caller of the reference would, and
// prints machine-readable results.  It also touches the other public classes once.
#include <cmath>
#include <cstdio>
#include <utility>
#include <vector>

#include "HybridAStar.h"
#include "VelocityGenerator.h"
#include "PedestrianHandler.h"

using namespace planning;

template <typename T> static int run_scenario(const char* tag)
{
    std::vector<T> steering{-30, -15, 0, 15, 30};
    for (auto& a : steering) a = a * M_PI / 180.0f;
    std::vector<T> weights{0, 0, 0, 0, 0, 0, 0};
    HybridAStar<T> planner(300, 10, T(0.5), T(0.75), T(0.1), T(0.95), T(0.4), 60, true, T(0.75), T(4.0), T(2.0), T(2.269), T(1.1), T(1.0),
                           static_cast<T>(M_PI / 4), 72, 1, steering, weights);
    std::vector<std::pair<Vector2D<T>, Vector2D<T>>> lines{{{T(21.9), T(4.5)}, {T(21.9), T(31.5)}},
                                                            {{T(20.4), T(33.0)}, {T(38.4), T(33.0)}},
                                                            {{T(10.5), T(4.5)}, {T(10.5), T(40.5)}},
                                                            {{T(9.0), T(42.0)}, {T(39.0), T(42.0)}}};
    std::vector<Obstacle<T>> boxes{Obstacle<T>(T(18.0), T(22.8), T(3.5), T(2.9)), Obstacle<T>(T(14.25), T(28.5), T(2.0), T(5.3)),
                                   Obstacle<T>(T(18.0), T(34.8), T(3.5), T(2.9))};
    Vector3D<T> start(T(18.0), T(18.0), static_cast<T>(M_PI_2)), goal(T(26.0), T(36.0), T(0));
    planner.update_goal(goal, start);
    for (int k = 0; k < 5; k++)
    {
        planner.update_obstacles();
        planner.update_obstacles(lines, std::vector<T>(lines.size(), T(0.6)), T(1.25));
        planner.update_obstacles(boxes, std::vector<T>(boxes.size(), T(0.75)), T(2.5));
    }
    const auto& map = planner.get_obstacles();
    int occupied = 0;
    for (auto& row : map) for (auto v : row) occupied += (v >= std::log(0.75 / 0.25)) ? 1 : 0;
    std::vector<Vector3D<T>> path;
    std::vector<T> curvature;
    std::pair<T, bool> r = planner.find_path(T(2.0), start, path, curvature);
    std::printf("%s success %d cost %.6f points %zu curvature %zu expansions %d occupied %d\n", tag, r.second ? 1 : 0, (double)r.first,
                path.size(), curvature.size(), planner.last_expansions(), occupied);
    for (auto it = path.rbegin(); it != path.rend(); ++it) std::printf("%s pt %.6f %.6f %.6f\n", tag, (double)it->_x, (double)it->_y, (double)it->_heading);
    // velocity profile on the returned path (host-side post-processing class)
    VelocityGenerator<T> vg(T(5), T(1.5), T(2.0), T(1.5), T(2.5));
    std::vector<T> vel;
    bool feasible = vg.generate_velocity_profile(T(2.0), T(10), path, curvature, vel, false, true);
    std::printf("%s velocity feasible %d n %zu v0 %.4f vend %.4f\n", tag, feasible ? 1 : 0, vel.size(), (double)vel.front(), (double)vel.back());
    // failure contract: a start boxed in by an obstacle -> {max, false}, vectors untouched
    std::vector<Vector3D<T>> p2; std::vector<T> c2;
    HybridAStar<T> blocked(300, 10, T(0.5), T(0.75), T(0.1), T(0.95), T(0.4), 60, true, T(0.75), T(4.0), T(2.0), T(2.269), T(1.1), T(1.0),
                           static_cast<T>(M_PI / 4), 72, 1, steering, weights);
    blocked.update_goal(goal, start);
    std::vector<Obstacle<T>> wall{Obstacle<T>(T(18.0), T(18.0), T(6.0), T(6.0))};
    for (int k = 0; k < 3; k++) blocked.update_obstacles(wall, std::vector<T>(1, T(0.9)), T(1.0));
    auto rb = blocked.find_path(T(2.0), start, p2, c2);
    std::printf("%s blocked success %d cost_is_max %d points %zu\n", tag, rb.second ? 1 : 0,
                rb.first == std::numeric_limits<T>::max() ? 1 : 0, p2.size());
    return r.second ? 0 : 1;
}

int main()
{
    int rc = run_scenario<float>("f32");
    rc |= run_scenario<double>("f64");

    // Dubins: the reference's utils/vehicle_dubins scenario, (0,0,0) -> (20,-20,pi/2)
    float rmin = 2.269f / (std::tan(30.0f * (float)M_PI / 180.0f) * std::cos(std::atan2(1.1f * std::tan(30.0f * (float)M_PI / 180.0f), 2.269f)));
    Dubins<float> dubins(rmin, 0.5f);
    Vector3D<float> ds(0, 0, 0), dg(20, -20, (float)M_PI_2);
    float len = dubins.get_shortest_path_length(ds, dg);
    std::vector<Vector3D<float>> dpath; std::vector<float> dcurv;
    auto dr = dubins.get_shortest_path(ds, dg, dpath, dcurv);
    std::printf("dubins rmin %.5f length %.6f type %s samples %zu long_turn %d\n", rmin, len, dubins.get_path_type().c_str(), dpath.size(), dr.second ? 1 : 0);

    // VehicleModel: successors of one state
    std::vector<float> st7{-30, -20, -10, 0, 10, 20, 30};
    for (auto& a : st7) a = a * (float)M_PI / 180.0f;
    VehicleModel<float> model(0.5f, 4.0f, 2.0f, 2.269f, 1.1f, 72, 1, st7, std::vector<float>(7, 0.0f));
    Vector3D<float> pose(0, 0, 0);
    Node3D<float> node(pose, 0.0f, 16.0f, 3, get_heading_index(0.0f, model.get_precision()), nullptr);
    std::vector<Node3D<float>> nb;
    bool neglect = model.get_neighbors(node, nb);
    auto sim = model.simulate_action(node, 6);
    std::printf("vehicle neighbors %zu neglect %d default_action %d sim_ok %d sim_x %.6f sim_y %.6f\n", nb.size(), neglect ? 1 : 0,
                model.get_default_action_index(), sim.first ? 1 : 0, sim.second._pose2D._x, sim.second._pose2D._y);

    // Grid3D + AStar sharing one grid (STORE_GRID_AS_REFERENCE layout, as in HybridAStar.cpp:22)
    std::vector<float> st5{-30, -15, 0, 15, 30};
    for (auto& a : st5) a = a * (float)M_PI / 180.0f;
    Grid3D<float> grid(0.5f, 0.75f, 0.1f, 0.95f, 0.4f, 60, true, 0.75f, 4.0f, 2.0f, 2.269f, 1.1f, 1.0f, (float)(M_PI / 4), 72, 1, st5,
                       std::vector<float>(5, 0.0f));
    AStar<float> astar(grid);
    Node3D<float> goal_node = grid.update_goal_heading(Vector3D<float>(26, 36, 0), Vector3D<float>(18, 18, (float)M_PI_2));
    std::vector<Obstacle<float>> boxes{Obstacle<float>(18.0f, 22.8f, 3.5f, 2.9f)};
    grid.update_obstacles(boxes, std::vector<float>(1, 0.9f), 2.5f);
    Node3D<float> start_node = grid.set_start_node(Vector3D<float>(18, 18, (float)M_PI_2));
    std::vector<Node3D<float>> succ;
    bool allowed = grid.get_neighbors(start_node, succ);
    float h_first = astar.find_path(start_node._base_node->_posd._x, start_node._base_node->_posd._y);
    float h_again = astar.find_path(start_node._base_node->_posd._x, start_node._base_node->_posd._y);   // cached now
    std::printf("grid3d goal_bin %d start_cell %d %d successors %zu shots_allowed %d h2d %.5f cached %.5f\n", goal_node._angle_bin,
                start_node._base_node->_posd._x, start_node._base_node->_posd._y, succ.size(), allowed ? 1 : 0, h_first, h_again);

    PedestrianHandler<float> ped(2.0f, 3.0f, 2.0f, 2.5f, 0.5f);
    std::vector<Obstacle<float>> walkers{Obstacle<float>(6.0f, 0.2f, 0.5f, 0.5f)};
    std::printf("pedestrian vmax %.4f\n", ped.calc_max_velocity(5.0f, Vector3D<float>(0, 0, 0), walkers));
    return rc;
}
