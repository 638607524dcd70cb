#ifndef PP_STUB_ODOMETRY_H
#define PP_STUB_ODOMETRY_H
#include "geometry_msgs/Pose.h"
#include "boost/shared_ptr.hpp"
namespace nav_msgs
{
    struct Odometry
    {
        struct PoseWithCov { geometry_msgs::Pose pose; } pose;
        struct TwistWithCov { geometry_msgs::Twist twist; } twist;
        typedef boost::shared_ptr<const Odometry> ConstPtr;
    };
}
#endif
