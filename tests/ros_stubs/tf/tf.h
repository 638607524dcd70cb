// TEST INFRASTRUCTURE stub of the two tf classes pose_to_vector3d uses (yaw extraction from a quaternion).
#ifndef PP_STUB_TF_H
#define PP_STUB_TF_H
#include <cmath>
namespace tf
{
    struct Quaternion { double x, y, z, w; Quaternion(double x_, double y_, double z_, double w_) : x(x_), y(y_), z(z_), w(w_) {} };
    struct Matrix3x3
    {
        Quaternion q;
        explicit Matrix3x3(const Quaternion& q_) : q(q_) {}
        void getRPY(double& roll, double& pitch, double& yaw) const
        {
            roll = std::atan2(2 * (q.w * q.x + q.y * q.z), 1 - 2 * (q.x * q.x + q.y * q.y));
            pitch = std::asin(2 * (q.w * q.y - q.z * q.x));
            yaw = std::atan2(2 * (q.w * q.z + q.x * q.y), 1 - 2 * (q.y * q.y + q.z * q.z));
        }
    };
}
#endif
