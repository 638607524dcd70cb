#ifndef PP_STUB_GEOMETRY_POSE_H
#define PP_STUB_GEOMETRY_POSE_H
#include "pp_replay_event.h"
namespace geometry_msgs
{
    struct Point { double x = 0, y = 0, z = 0; };
    struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
    struct Vector3 { double x = 0, y = 0, z = 0; };
    struct Pose { Point position; Quaternion orientation; };
    struct Twist { Vector3 linear, angular; };
    // replay harness: planar pose from (x, y, yaw)
    inline void pp_replay_set_pose(Pose& p, double x, double y, double yaw)
    {
        p.position.x = x; p.position.y = y; p.position.z = 0;
        p.orientation.x = 0; p.orientation.y = 0; p.orientation.z = std::sin(yaw / 2); p.orientation.w = std::cos(yaw / 2);
    }
}
#endif
