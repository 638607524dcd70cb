// path_planning_pkg API surface, B200 build: kinematic-bicycle motion primitives (reference:
// include/path_planning_pkg/VehicleModel.h:13-46, lib/VehicleModel.cpp:7-136).  The displacement table is integrated on
// the host exactly as the reference does; successor roll-out runs in pp_successor_kernel (one warp per state).
#ifndef PP_B200_API_VEHICLE_MODEL_H
#define PP_B200_API_VEHICLE_MODEL_H

#include <memory>
#include <utility>
#include <vector>
#include "Node3D.h"
#include "common.h"

namespace planning
{
    template <typename T> class VehicleModel
    {
    public:
        VehicleModel(T ts, T max_lat_acc, T max_long_dec, T wheelbase, T rear_to_cg, int num_angle_bins, int num_actions,
                     const std::vector<T>& steering, const std::vector<T>& curvature_weights);
        ~VehicleModel();
        VehicleModel(const VehicleModel&) = delete;
        VehicleModel& operator=(const VehicleModel&) = delete;

        T get_precision() const;
        int get_default_action_index() const;
        // successors of `node`; returns whether accelerations were neglected (= Dubins shots allowed)
        bool get_neighbors(const Node3D<T>& node, std::vector<Node3D<T>>& neighbors) const;
        std::pair<bool, Node3D<T>> simulate_action(const Node3D<T>& node, const int action_index) const;
        const std::vector<T>& get_abs_curvatures() const;

    private:
        struct Impl;
        std::unique_ptr<Impl> _impl;
    };
}

#endif
