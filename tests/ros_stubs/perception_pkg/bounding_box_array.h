#ifndef PP_STUB_BBOX_ARRAY_H
#define PP_STUB_BBOX_ARRAY_H
#include <vector>
#include "bounding_box.h"
#include "boost/shared_ptr.hpp"
namespace perception_pkg
{
    struct bounding_box_array { std::vector<bounding_box> bbs_array; typedef boost::shared_ptr<const bounding_box_array> ConstPtr; };
}
#endif
