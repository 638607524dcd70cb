// stub of the generated message header for msg/Waypoint.msg (Header, bool stop_at_waypoint, geometry_msgs/Pose pose)
#ifndef PP_STUB_WAYPOINT_H
#define PP_STUB_WAYPOINT_H
#include "geometry_msgs/Pose.h"
#include "boost/shared_ptr.hpp"
namespace path_planning_pkg
{
    struct Waypoint { bool stop_at_waypoint = false; geometry_msgs::Pose pose; typedef boost::shared_ptr<const Waypoint> ConstPtr; };
}
#endif
