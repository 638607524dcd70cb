#ifndef PP_STUB_BBOX_ARRAY_H
#define PP_STUB_BBOX_ARRAY_H
#include <algorithm>
#include <vector>
#include "bounding_box.h"
#include "boost/shared_ptr.hpp"
namespace perception_pkg
{
    struct bounding_box_array { std::vector<bounding_box> bbs_array; typedef boost::shared_ptr<const bounding_box_array> ConstPtr; };
    // replay harness (ros/ros.h): script line "objects n { class_with_underscores length width confidence cx cy } x n"
    inline const char* pp_replay_kind(const bounding_box_array*) { return "objects"; }
    inline void pp_replay_fill(bounding_box_array& m, const pp_replay::Event& e)
    {
        const std::vector<std::string>& w = pp_replay::words_of(e);
        size_t n = (size_t)pp_replay::num(w, 0);
        for (size_t k = 0; k < n && 1 + 6 * k + 5 < w.size(); k++)
        {
            bounding_box b;
            b.class_name = w[1 + 6 * k];
            std::replace(b.class_name.begin(), b.class_name.end(), '_', ' ');
            b.length = pp_replay::num(w, 2 + 6 * k); b.width = pp_replay::num(w, 3 + 6 * k); b.confidence = pp_replay::num(w, 4 + 6 * k);
            b.centroid.x = pp_replay::num(w, 5 + 6 * k); b.centroid.y = pp_replay::num(w, 6 + 6 * k);
            m.bbs_array.push_back(b);
        }
    }
}
#endif
