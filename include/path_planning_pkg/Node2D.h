// path_planning_pkg API surface, B200 build: the 2D grid node value type.
// Mirrors the public members and operators of the reference struct (reference: include/path_planning_pkg/Node2D.h:12-85,
// lib/Node2D.cpp:7-40), including its deliberately odd ordering -- `a < b` is false for equal cells and otherwise
// compares f -- because std::set behaviour under that ordering is observable.  No Boost dependency:
// the hash combiner is written out.
#ifndef PP_B200_API_NODE2D_H
#define PP_B200_API_NODE2D_H

#include <cstddef>
#include <functional>
#include <iostream>
#include "common.h"

namespace planning
{
    template <typename T> struct Node2D
    {
        Vector2D<int> _posd;       // grid indices
        T _cost_g, _cost_h, _cost_f;
        const Node2D<T>* _prev;

        Node2D(int xd, int yd, T cost_g, T cost_h, const Node2D<T>* prev)
            : _posd(xd, yd), _cost_g(cost_g), _cost_h(cost_h), _cost_f(cost_g + cost_h), _prev(prev) {}
        Node2D(int xd, int yd) : Node2D(xd, yd, T(0), T(0), nullptr) {}

        void set_accumulated_cost(const T cost_g) { _cost_g = cost_g; _cost_f = cost_g + _cost_h; }
        void set_heuristic_cost(const T cost_h) { _cost_h = cost_h; _cost_f = _cost_g + cost_h; }
        void soft_reset() { _cost_g = T(0); _cost_f = _cost_h; _prev = nullptr; }   // keeps indices and h

        bool same_cell(const Node2D<T>& o) const { return _posd._x == o._posd._x && _posd._y == o._posd._y; }
        friend bool operator==(const Node2D<T>& a, const Node2D<T>& b) { return a.same_cell(b); }
        friend bool operator!=(const Node2D<T>& a, const Node2D<T>& b) { return !a.same_cell(b); }
        // all four relations are false for the same cell (reference Node2D.h:37-59)
        friend bool operator<(const Node2D<T>& a, const Node2D<T>& b) { return !a.same_cell(b) && a._cost_f < b._cost_f; }
        friend bool operator<=(const Node2D<T>& a, const Node2D<T>& b) { return !a.same_cell(b) && a._cost_f <= b._cost_f; }
        friend bool operator>(const Node2D<T>& a, const Node2D<T>& b) { return !a.same_cell(b) && a._cost_f > b._cost_f; }
        friend bool operator>=(const Node2D<T>& a, const Node2D<T>& b) { return !a.same_cell(b) && a._cost_f >= b._cost_f; }

        friend std::ostream& operator<<(std::ostream& os, const Node2D<T>& n)
        {
            return os << "xd = " << n._posd._x << " yd = " << n._posd._y << "\n"
                      << "cost_g = " << n._cost_g << " cost_h = " << n._cost_h << " cost_f = " << n._cost_f << "\n" << std::endl;
        }

        // golden-ratio combiner over (x, y), then std::hash<size_t>
        static void hash_mix(std::size_t& seed, int v) { seed ^= std::hash<int>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2); }
        struct HashFunction
        {
            std::size_t operator()(const Node2D<T>& n) const
            {
                std::size_t seed = 0;
                hash_mix(seed, n._posd._x); hash_mix(seed, n._posd._y);
                return std::hash<std::size_t>()(seed);
            }
        };
    };
}

#endif
