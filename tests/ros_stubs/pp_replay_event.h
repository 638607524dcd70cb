// TEST INFRASTRUCTURE: forward declaration shared by the stub message headers (see ros/ros.h for the replay harness)
#ifndef PP_REPLAY_EVENT_H
#define PP_REPLAY_EVENT_H
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
namespace pp_replay
{
    struct Event;
    // accessors implemented in ros/ros.h terms would create an include cycle; the message headers only need these two
    const std::vector<std::string>& words_of(const Event& e);
    inline double num(const std::vector<std::string>& w, size_t k) { return k < w.size() ? std::strtod(w[k].c_str(), nullptr) : 0.0; }
}
#endif
