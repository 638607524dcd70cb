// TEST INFRASTRUCTURE: minimal stand-in for ROS 1 so that the reference's UNMODIFIED src/local_planner.cpp can be
// compiled and linked against this repo's headers and libpath_planning_b200.so (link check of the drop-in boundary,
// SURVEY.md §8b).  Only what that one translation unit touches is declared; nothing here talks to a ROS master.
#ifndef PP_STUB_ROS_H
#define PP_STUB_ROS_H
#include <cstdio>
#include <iterator>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include "boost/shared_ptr.hpp"

#define ROS_INFO(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_INFO_STREAM(x) do { std::ostringstream oss__; oss__ << x; std::printf("%s\n", oss__.str().c_str()); } while (0)

namespace ros
{
    inline void init(int&, char**, const std::string&) {}
    inline bool ok() { return false; }        // the stub node leaves its loop immediately
    inline void spinOnce() {}
    struct Rate { explicit Rate(double) {} void sleep() {} };
    struct Subscriber {};
    struct Publisher { template <class M> void publish(const M&) const {} };
    struct NodeHandle
    {
        template <class T> bool param(const std::string&, T& out, const T& def) const { out = def; return false; }
        bool param(const std::string&, std::string& out, const std::string& def) const { out = def; return false; }
        // list parameters: the launch-file values (launch/local_planner.launch: steering in degrees, curvature weights)
        template <class T> bool getParam(const std::string& name, std::vector<T>& out) const
        {
            if (name.find("steering") != std::string::npos) out = {T(-40), T(-20), T(0), T(20), T(40)};
            else out = {T(1), T(0.5), T(0), T(0.5), T(1)};
            return true;
        }
        template <class M, class C> Subscriber subscribe(const std::string&, unsigned, void (C::*)(const typename M::ConstPtr&), C*) { return Subscriber(); }
        template <class P, class C> Subscriber subscribe(const std::string&, unsigned, void (C::*)(P), C*) { return Subscriber(); }
        template <class M> Publisher advertise(const std::string&, unsigned, bool = false) { return Publisher(); }
    };
}
#endif
