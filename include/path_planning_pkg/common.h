// path_planning_pkg API surface, B200 build: math helpers and the two small vector value types.
// Source-compatible with the reference header of the same name (reference: include/path_planning_pkg/common.h:7-221):
// same namespace, names, member names and operator set, so callers such as src/local_planner.cpp and
// utils/*/test_*.cpp compile unchanged.  These stay header-inline on the host (they are scalar glue, not hot path).
#ifndef PP_B200_API_COMMON_H
#define PP_B200_API_COMMON_H

#include <cmath>

namespace planning
{
    // round(value / precision) * precision                                  (reference common.h:8-12)
    template <typename T> T round_to_nearest(const T value, const T precision) { return std::round(value / precision) * precision; }

    // angle folded into [-pi, pi] with the reference's promotion pattern     (reference common.h:14-29)
    template <typename T> T wrap_pi(const T angle)
    {
        T folded = std::fmod(angle, 2 * M_PI);
        if (folded > M_PI) return folded - 2 * M_PI;
        if (folded < -M_PI) return folded + 2 * M_PI;
        return folded;
    }

    // heading bin; may return num_bins for headings within half a bin of +pi  (reference common.h:31-36)
    template <typename T> int get_heading_index(const T heading, const T precision)
    {
        return static_cast<int>((round_to_nearest(heading, precision) + M_PI) / precision);
    }

    template <typename T> struct Vector2D
    {
        T _x, _y;

        Vector2D() : _x(T(0)), _y(T(0)) {}
        Vector2D(T x, T y) : _x(x), _y(y) {}
        Vector2D(const Vector2D<T>& o) : _x(o._x), _y(o._y) {}
        template <typename U> Vector2D(const Vector2D<U>& o) : _x(static_cast<T>(o._x)), _y(static_cast<T>(o._y)) {}
        ~Vector2D() {}

        // rotation by -angle (the grid frame convention)                      (reference common.h:55-70)
        Vector2D<T> get_rotated_vector(const T angle) const
        {
            const T c = std::cos(angle), s = std::sin(angle);
            return {_x * c + _y * s, -_x * s + _y * c};
        }
        void rotate_vector(const T angle) { *this = get_rotated_vector(angle); }

        Vector2D& operator=(const Vector2D<T>& o) { _x = o._x; _y = o._y; return *this; }
        template <typename U> Vector2D& operator=(const Vector2D<U>& o) { _x = static_cast<T>(o._x); _y = static_cast<T>(o._y); return *this; }

#define PP_V2_OP(op)                                                                                       \
        Vector2D<T> operator op(const Vector2D<T>& o) const { return {_x op o._x, _y op o._y}; }           \
        Vector2D<T> operator op(const T o) const { return {_x op o, _y op o}; }
        PP_V2_OP(+) PP_V2_OP(-) PP_V2_OP(*) PP_V2_OP(/)
#undef PP_V2_OP
    };

    template <typename T> struct Vector3D
    {
        T _x, _y, _heading;

        Vector3D() : _x(T(0)), _y(T(0)), _heading(T(0)) {}
        Vector3D(T x, T y, T heading) : _x(x), _y(y), _heading(heading) {}
        Vector3D(const Vector3D<T>& o) : _x(o._x), _y(o._y), _heading(o._heading) {}
        template <typename U> Vector3D(const Vector3D<U>& o)
            : _x(static_cast<T>(o._x)), _y(static_cast<T>(o._y)), _heading(static_cast<T>(o._heading)) {}
        ~Vector3D() {}

        // position rotated by -angle, heading reduced by angle and wrapped      (reference common.h:162-169)
        Vector3D<T> get_rotated_vector(const T angle) const
        {
            const T c = std::cos(angle), s = std::sin(angle);
            return {_x * c + _y * s, -_x * s + _y * c, wrap_pi<T>(_heading - angle)};
        }

        Vector3D& operator=(const Vector3D<T>& o) { _x = o._x; _y = o._y; _heading = o._heading; return *this; }
        template <typename U> Vector3D& operator=(const Vector3D<U>& o)
        {
            _x = static_cast<T>(o._x); _y = static_cast<T>(o._y); _heading = static_cast<T>(o._heading);
            return *this;
        }

#define PP_V3_OP(op) \
        Vector3D<T> operator op(const Vector3D<T>& o) const { return {_x op o._x, _y op o._y, _heading op o._heading}; }
        PP_V3_OP(+) PP_V3_OP(-) PP_V3_OP(*) PP_V3_OP(/)
#undef PP_V3_OP
    };
}

#endif
