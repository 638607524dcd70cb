// path_planning_pkg API surface, B200 build: axis-aligned box obstacle (reference: include/path_planning_pkg/Obstacle.h:10-21,
// lib/Obstacle.cpp:7-19).  Heading is atan2 of the velocity and is not used by the rasteriser.
#ifndef PP_B200_API_OBSTACLE_H
#define PP_B200_API_OBSTACLE_H

#include <cmath>
#include "common.h"

namespace planning
{
    template <typename T> struct Obstacle
    {
        Vector3D<T> _pose2D;       // centre and heading
        Vector2D<T> _velocity;
        Vector2D<T> _dimensions;   // extents along world x and y

        Obstacle(T position_x, T position_y, T velocity_x, T velocity_y, T dimension_x, T dimension_y)
            : _pose2D(position_x, position_y, T(0)), _velocity(velocity_x, velocity_y), _dimensions(dimension_x, dimension_y)
        {
            _pose2D._heading = std::atan2(_velocity._y, _velocity._x);
        }
        Obstacle(T position_x, T position_y, T dimension_x, T dimension_y)
            : Obstacle(position_x, position_y, T(0), T(0), dimension_x, dimension_y) {}
    };
}

#endif
