// path_planning_pkg API surface, B200 build: velocity profile along a returned path (reference:
// include/path_planning_pkg/VelocityGenerator.h:9-29, lib/VelocityGenerator.cpp:19-85).  Runs after the search on ~60
// points, sequentially dependent: kept on the host ("next" row N3 in SURVEY.md §8f).
#ifndef PP_B200_API_VELOCITY_GENERATOR_H
#define PP_B200_API_VELOCITY_GENERATOR_H

#include <vector>
#include "common.h"

namespace planning
{
    template <typename T> class VelocityGenerator
    {
    public:
        VelocityGenerator(T max_velocity, T coast_velocity, T max_lat_acc, T max_long_acc, T max_long_dec);
        // path / curvature in goal -> start order; returns whether vel_init is feasible for the profile
        bool generate_velocity_profile(const T vel_init, const T max_velocity_curr, const std::vector<Vector3D<T>>& path,
                                       const std::vector<T>& curvature, std::vector<T>& velocity, bool coast_to_goal,
                                       bool stop_at_goal = false) const;

    private:
        const T _max_velocity, _coast_velocity, _max_lat_acc, _max_lat_acc_sqr, _max_long_acc, _max_long_dec;
    };
}

#endif
