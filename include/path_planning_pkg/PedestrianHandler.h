// path_planning_pkg API surface, B200 build: velocity cap from time-to-collision with pedestrians (reference:
// include/path_planning_pkg/PedestrianHandler.h:11-36, lib/PedestrianHandler.cpp:17-90).  A handful of scalars per
// frame: host side ("next" row N4 in SURVEY.md §8f).
#ifndef PP_B200_API_PEDESTRIAN_HANDLER_H
#define PP_B200_API_PEDESTRIAN_HANDLER_H

#include <limits>
#include <vector>
#include "Obstacle.h"
#include "common.h"

namespace planning
{
    template <typename T> class PedestrianHandler
    {
    public:
        PedestrianHandler(T detection_arc_angle, T min_stop_dist, T min_allowable_ttc, T max_long_dec, T min_vel);
        T calc_max_velocity(const T vel_curr, const Vector3D<T>& pose_curr, const std::vector<Obstacle<T>>& pedestrians) const;

    private:
        T time_to_collision(const T vel_curr, const Vector3D<T>& pose_curr, const Obstacle<T>& pedestrian) const;
        void bearing_and_range(const Vector3D<T>& pose_curr, const Obstacle<T>& pedestrian, T& rel_angle, T& long_dist) const;
        const T _half_arc, _min_stop_dist, _min_allowable_ttc, _max_long_dec, _min_vel;
    };
}

#endif
