// path_planning_pkg API surface, B200 build: CSC Dubins paths (reference: include/path_planning_pkg/Dubins.h:12-63,
// lib/Dubins.cpp:19-153).  Lengths and sampled paths are evaluated by the CUDA kernels pp_dubins_length_kernel /
// pp_dubins_path_kernel through the C ABI.
#ifndef PP_B200_API_DUBINS_H
#define PP_B200_API_DUBINS_H

#include <array>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "common.h"

namespace planning
{
    enum class Path { RSR, RSL, LSR, LSL };

    template <typename T> class Dubins
    {
    public:
        Dubins(T r_min, T step_size);
        ~Dubins();
        Dubins(const Dubins&) = delete;
        Dubins& operator=(const Dubins&) = delete;

        T get_shortest_path_length(const Vector3D<T>& start, const Vector3D<T>& goal);
        T get_shortest_path_length(const Vector3D<T>& start, const Vector3D<T>& goal, Vector2D<T>& center_s_r,
                                   Vector2D<T>& center_s_l, Vector2D<T>& center_g_r, Vector2D<T>& center_g_l);
        // {length, first arc longer than 90 deg}; fills the sampled poses and their curvature
        std::pair<T, bool> get_shortest_path(const Vector3D<T>& start, const Vector3D<T>& goal, std::vector<Vector3D<T>>& path,
                                             std::vector<T>& path_curvature);
        std::string get_path_type() const;

    private:
        struct Impl;
        std::unique_ptr<Impl> _impl;
        std::array<T, 4> _params;
        Path _path_type;
        T _r_min;
    };
}

#endif
