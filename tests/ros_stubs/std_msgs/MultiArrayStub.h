#ifndef PP_STUB_MULTIARRAY_H
#define PP_STUB_MULTIARRAY_H
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "boost/shared_ptr.hpp"
#include "pp_replay_event.h"
namespace std_msgs
{
    struct MultiArrayDimension { std::string label; unsigned size = 0, stride = 0; };
    struct MultiArrayLayout { std::vector<MultiArrayDimension> dim; unsigned data_offset = 0; };
    template <class T> struct MultiArrayT
    {
        MultiArrayLayout layout; std::vector<T> data;
        typedef boost::shared_ptr<const MultiArrayT<T>> ConstPtr;
    };
    // replay harness (ros/ros.h): script line "lanes n { x1 y1 x2 y2 } x n" -> dim[0] = n lines, dim[1] = 4 values
    template <class T> inline const char* pp_replay_kind(const MultiArrayT<T>*) { return "lanes"; }
    template <class T> inline void pp_replay_fill(MultiArrayT<T>& m, const pp_replay::Event& e)
    {
        const std::vector<std::string>& w = pp_replay::words_of(e);
        size_t n = (size_t)pp_replay::num(w, 0);
        m.layout.dim.resize(2);
        m.layout.dim[0].label = "lines"; m.layout.dim[0].size = (unsigned)n; m.layout.dim[0].stride = (unsigned)(4 * n);
        m.layout.dim[1].label = "x1,y1,x2,y2"; m.layout.dim[1].size = 4; m.layout.dim[1].stride = 4;
        for (size_t k = 0; k < 4 * n; k++) m.data.push_back((T)pp_replay::num(w, 1 + k));
    }
    // a published message as raw IEEE bits: layout sizes, then every data word
    inline void pp_replay_put(std::FILE* f, float v) { uint32_t b; std::memcpy(&b, &v, 4); std::fprintf(f, " %08x", b); }
    inline void pp_replay_put(std::FILE* f, double v) { uint64_t b; std::memcpy(&b, &v, 8); std::fprintf(f, " %016llx", (unsigned long long)b); }
    template <class T> inline void pp_replay_dump(std::FILE* f, const MultiArrayT<T>& m)
    {
        std::fprintf(f, " dims");
        for (const MultiArrayDimension& d : m.layout.dim) std::fprintf(f, " %s:%u:%u", d.label.c_str(), d.size, d.stride);
        std::fprintf(f, " data %zu :", m.data.size());
        for (T v : m.data) pp_replay_put(f, v);
    }
}
#endif
