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
    // replay harness (ros/ros.h): script line "odom x y yaw vx vy"
    inline const char* pp_replay_kind(const Odometry*) { return "odom"; }
    inline void pp_replay_fill(Odometry& m, const pp_replay::Event& e)
    {
        const std::vector<std::string>& w = pp_replay::words_of(e);
        geometry_msgs::pp_replay_set_pose(m.pose.pose, pp_replay::num(w, 0), pp_replay::num(w, 1), pp_replay::num(w, 2));
        m.twist.twist.linear.x = pp_replay::num(w, 3); m.twist.twist.linear.y = pp_replay::num(w, 4);
    }
}
#endif
