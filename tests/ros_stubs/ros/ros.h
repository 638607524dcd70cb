// TEST INFRASTRUCTURE: minimal in-process stand-in for ROS 1 so that the reference's UNMODIFIED src/local_planner.cpp can be
// compiled, linked AND RUN against this repo's headers and libpath_planning_b200.so (drop-in boundary, SURVEY.md §8b;
// replay harness, §8(f) N4).  Only what that one translation unit touches exists; nothing here talks to a ROS master.
//
// Two modes:
//   * no script (default): ros::ok() is false, the node constructs, leaves its loop at once and exits (link check).
//   * replay: PP_REPLAY_SCRIPT=<file> holds parameter overrides and a timeline of messages; every ros::spinOnce() delivers
//     the messages of the next tick to the node's subscribed callbacks, ros::ok() turns false after the last tick, and
//     every message the node publishes is appended to PP_REPLAY_OUT=<file> as raw IEEE bits.  PP_REPLAY_TIMES=<file>
//     receives the wall time of each loop iteration (spinOnce return -> Rate::sleep), i.e. of update_trajectory().
//
// Script lines (blank-separated, '#' comments):
//   param <name> <value...>                      parameter-server entry (scalars, strings, lists)
//   odom <x> <y> <yaw> <vx> <vy>                 nav_msgs/Odometry
//   waypoint <x> <y> <yaw> <stop 0|1>            path_planning_pkg/Waypoint
//   objects <n> { <class_with_underscores> <length> <width> <confidence> <cx> <cy> } x n
//   lanes <n> { <x1> <y1> <x2> <y2> } x n        std_msgs/Float{32,64}MultiArray, dim[0] = n lines, dim[1] = 4
//   tick                                         end of one spinOnce batch
#ifndef PP_STUB_ROS_H
#define PP_STUB_ROS_H
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iterator>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>
#include "boost/shared_ptr.hpp"
#include "pp_replay_event.h"

#define ROS_INFO(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_INFO_STREAM(x) do { std::ostringstream oss__; oss__ << x; std::printf("%s\n", oss__.str().c_str()); } while (0)

namespace pp_replay
{
    struct Event { std::string kind; std::vector<std::string> words; };      // one script line: kind + remaining tokens

    struct State
    {
        bool active = false;
        std::map<std::string, std::vector<std::string>> params;
        std::vector<std::vector<Event>> ticks;
        size_t next_tick = 0;
        std::multimap<std::string, std::function<void(const Event&)>> subscribers;     // by message kind
        std::FILE* out = nullptr;
        std::FILE* times = nullptr;
        std::chrono::steady_clock::time_point t_spin;
    };
    inline State& state() { static State s; return s; }
    inline const std::vector<std::string>& words_of(const Event& e) { return e.words; }

    inline void load()
    {
        State& s = state();
        const char* path = std::getenv("PP_REPLAY_SCRIPT");
        if (!path) return;
        std::ifstream in(path);
        if (!in) { std::fprintf(stderr, "pp_replay: cannot read %s\n", path); std::exit(2); }
        std::string line;
        std::vector<Event> cur;
        while (std::getline(in, line))
        {
            std::istringstream ls(line);
            Event e;
            if (!(ls >> e.kind) || e.kind[0] == '#') continue;
            for (std::string w; ls >> w;) e.words.push_back(w);
            if (e.kind == "param") { if (!e.words.empty()) s.params[e.words[0]] = std::vector<std::string>(e.words.begin() + 1, e.words.end()); }
            else if (e.kind == "tick") { s.ticks.push_back(cur); cur.clear(); }
            else cur.push_back(e);
        }
        if (!cur.empty()) s.ticks.push_back(cur);
        if (const char* o = std::getenv("PP_REPLAY_OUT")) s.out = std::fopen(o, "w");
        if (const char* t = std::getenv("PP_REPLAY_TIMES")) s.times = std::fopen(t, "w");
        s.active = true;
    }

    template <class T> inline bool parse(const std::string& w, T& out) { std::istringstream is(w); return bool(is >> out); }
    inline bool parse(const std::string& w, bool& out) { out = (w == "1" || w == "true" || w == "True"); return true; }
    inline bool parse(const std::string& w, std::string& out) { out = w; return true; }

    // message type behind a callback parameter `const boost::shared_ptr<const M>&`
    template <class P> struct msg_of
    {
        typedef typename std::remove_cv<typename std::remove_reference<P>::type>::type ptr_type;
        typedef typename std::remove_cv<typename ptr_type::element_type>::type type;
    };
}

namespace ros
{
    inline void init(int&, char**, const std::string&) { pp_replay::load(); }
    inline bool ok()
    {
        pp_replay::State& s = pp_replay::state();
        bool more = s.active && s.next_tick < s.ticks.size();
        if (s.active && !more)
        {
            if (s.out) { std::fclose(s.out); s.out = nullptr; }
            if (s.times) { std::fclose(s.times); s.times = nullptr; }
        }
        return more;
    }
    inline void spinOnce()
    {
        pp_replay::State& s = pp_replay::state();
        if (s.active && s.next_tick < s.ticks.size())
        {
            for (const pp_replay::Event& e : s.ticks[s.next_tick])
            {
                auto range = s.subscribers.equal_range(e.kind);
                for (auto it = range.first; it != range.second; ++it) it->second(e);
            }
            if (s.out) std::fprintf(s.out, "tick %zu\n", s.next_tick);
            s.next_tick++;
        }
        s.t_spin = std::chrono::steady_clock::now();
    }
    struct Rate
    {
        explicit Rate(double) {}
        void sleep()
        {
            pp_replay::State& s = pp_replay::state();
            if (s.times)
                std::fprintf(s.times, "%zu %.6f\n", s.next_tick - 1,
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s.t_spin).count());
        }
    };
    struct Subscriber {};
    struct Publisher
    {
        std::string topic;
        // messages define pp_replay_dump(FILE*, const M&) next to their type (found by argument-dependent lookup)
        template <class M> void publish(const M& m) const
        {
            pp_replay::State& s = pp_replay::state();
            if (s.out) { std::fprintf(s.out, "pub %s", topic.c_str()); pp_replay_dump(s.out, m); std::fprintf(s.out, "\n"); }
        }
    };
    struct NodeHandle
    {
        template <class T> bool param(const std::string& name, T& out, const T& def) const
        {
            auto& p = pp_replay::state().params;
            auto it = p.find(name);
            if (it != p.end() && !it->second.empty() && pp_replay::parse(it->second[0], out)) return true;
            out = def; return false;
        }
        // list parameters: script override, else the launch-file values (launch/local_planner.launch: steering in degrees,
        // curvature weights)
        template <class T> bool getParam(const std::string& name, std::vector<T>& out) const
        {
            auto& p = pp_replay::state().params;
            auto it = p.find(name);
            if (it != p.end())
            {
                out.clear();
                for (const std::string& w : it->second) { T v; if (pp_replay::parse(w, v)) out.push_back(v); }
                return true;
            }
            if (name.find("steering") != std::string::npos) out = {T(-40), T(-20), T(0), T(20), T(40)};
            else out = {T(1), T(0.5), T(0), T(0.5), T(1)};
            return true;
        }
        // messages define pp_replay_kind(const M*) and pp_replay_fill(M&, const Event&) next to their type
        template <class P, class C> Subscriber subscribe(const std::string&, unsigned, void (C::*cb)(P), C* obj)
        {
            typedef typename pp_replay::msg_of<P>::type M;
            pp_replay::state().subscribers.emplace(pp_replay_kind(static_cast<const M*>(nullptr)), [cb, obj](const pp_replay::Event& e)
            {
                std::shared_ptr<M> m = std::make_shared<M>();
                pp_replay_fill(*m, e);
                boost::shared_ptr<const M> cm = m;
                (obj->*cb)(cm);
            });
            return Subscriber();
        }
        template <class M> Publisher advertise(const std::string& topic, unsigned, bool = false) { Publisher p; p.topic = topic; return p; }
    };
}
#endif
