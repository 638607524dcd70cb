#ifndef PP_STUB_BBOX_H
#define PP_STUB_BBOX_H
#include <string>
#include "geometry_msgs/Pose.h"
namespace perception_pkg
{
    struct bounding_box { std::string class_name; double length = 0, width = 0, height = 0, confidence = 0; geometry_msgs::Point centroid; };
}
#endif
