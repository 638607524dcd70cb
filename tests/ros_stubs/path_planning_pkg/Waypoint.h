// stub of the generated message header for msg/Waypoint.msg (Header, bool stop_at_waypoint, geometry_msgs/Pose pose)
#ifndef PP_STUB_WAYPOINT_H
#define PP_STUB_WAYPOINT_H
#include "geometry_msgs/Pose.h"
#include "boost/shared_ptr.hpp"
namespace path_planning_pkg
{
    struct Waypoint { bool stop_at_waypoint = false; geometry_msgs::Pose pose; typedef boost::shared_ptr<const Waypoint> ConstPtr; };
    // replay harness (ros/ros.h): script line "waypoint x y yaw stop"
    inline const char* pp_replay_kind(const Waypoint*) { return "waypoint"; }
    inline void pp_replay_fill(Waypoint& m, const pp_replay::Event& e)
    {
        const std::vector<std::string>& w = pp_replay::words_of(e);
        geometry_msgs::pp_replay_set_pose(m.pose, pp_replay::num(w, 0), pp_replay::num(w, 1), pp_replay::num(w, 2));
        m.stop_at_waypoint = pp_replay::num(w, 3) != 0.0;
    }
}
#endif
