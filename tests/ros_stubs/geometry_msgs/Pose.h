#ifndef PP_STUB_GEOMETRY_POSE_H
#define PP_STUB_GEOMETRY_POSE_H
namespace geometry_msgs
{
    struct Point { double x = 0, y = 0, z = 0; };
    struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
    struct Vector3 { double x = 0, y = 0, z = 0; };
    struct Pose { Point position; Quaternion orientation; };
    struct Twist { Vector3 linear, angular; };
}
#endif
