// path_planning_pkg API surface, B200 build: the Hybrid A* state value type.
// Mirrors the public members and operators of the reference struct (reference: include/path_planning_pkg/Node3D.h:14-100,
// lib/Node3D.cpp:7-62).  Two observable quirks are kept on purpose: equality looks at the 2D cell
// only (the reference compares a node's heading bin with itself), while inequality, ordering and the hash do use the bin.
#ifndef PP_B200_API_NODE3D_H
#define PP_B200_API_NODE3D_H

#include <algorithm>
#include <functional>
#include <vector>
#include "Node2D.h"
#include "common.h"

namespace planning
{
    template <typename T> struct Node3D
    {
        // Field for field the device record PPState (csrc/core/pp_defs.h) / pp_state (include/pp_b200.h) carries the same
        // data with 32-bit indices in place of the two pointers:
        Vector3D<T> _pose2D;                 // continuous pose in the goal-centred grid frame      -> x, y, heading
        T _cost_g, _cost_f, _vmin_sqr;       // path cost, g + h, squared minimum speed             -> g, f, vmin_sqr
        int _curvature_index, _angle_bin;    // steering primitive that led here, heading bin (may equal num_angle_bins)
        const Node2D<T>* _base_node;         // the 2D cell the pose falls into                      -> ci, cj
        const Node3D<T>* _prev;              // parent in the closed set                             -> index into the closed log

        Node3D(Vector3D<T>& pose2D, T cost_g, T vmin_sqr, int curvature_index, int angle_bin, const Node2D<T>* base_node,
               const Node3D<T>* prev)
            : _pose2D(pose2D), _cost_g(cost_g), _cost_f(cost_g), _vmin_sqr(vmin_sqr), _curvature_index(curvature_index),
              _angle_bin(angle_bin), _base_node(base_node), _prev(prev) {}
        Node3D(Vector3D<T>& pose2D, T cost_g, T vmin_sqr, int curvature_index, int angle_bin, const Node3D<T>* prev)
            : Node3D(pose2D, cost_g, vmin_sqr, curvature_index, angle_bin, nullptr, prev) {}
        Node3D() : _pose2D(), _cost_g(T(0)), _cost_f(T(0)), _vmin_sqr(T(0)), _curvature_index(0), _angle_bin(0),
                   _base_node(nullptr), _prev(nullptr) {}

        // Cost bookkeeping of lib/Node3D.cpp:27-42: f accumulates g and then max(h, base node's cached 2D cost).
        void set_accumulated_cost(const T cost_g) { _cost_g = cost_g; _cost_f += cost_g; }
        void set_heuristic_cost(const T cost_h) { _cost_f += (_base_node != nullptr) ? std::max(cost_h, _base_node->_cost_f) : cost_h; }
        void soft_reset() { _cost_g = T(0); _cost_f = T(0); _prev = nullptr; }

        // The ordering the reference's std::set runs on (not a strict weak ordering): two nodes compare by f only
        // when they differ in (cell, bin); the device open list reproduces it in csrc/core/pp_rbtree.h::pp_lt.
        bool differs(const Node3D<T>& o) const { return (*_base_node != *o._base_node) || (_angle_bin != o._angle_bin); }
        friend bool operator==(const Node3D<T>& a, const Node3D<T>& b) { return *a._base_node == *b._base_node; }   // cell only (F6)
        friend bool operator!=(const Node3D<T>& a, const Node3D<T>& b) { return a.differs(b); }
        friend bool operator<(const Node3D<T>& a, const Node3D<T>& b) { return a.differs(b) && a._cost_f < b._cost_f; }
        friend bool operator<=(const Node3D<T>& a, const Node3D<T>& b) { return a.differs(b) && a._cost_f <= b._cost_f; }
        friend bool operator>(const Node3D<T>& a, const Node3D<T>& b) { return a.differs(b) && a._cost_f > b._cost_f; }
        friend bool operator>=(const Node3D<T>& a, const Node3D<T>& b) { return a.differs(b) && a._cost_f >= b._cost_f; }

        friend std::ostream& operator<<(std::ostream& os, const Node3D<T>& n)
        {
            if (n._base_node != nullptr) os << "xd = " << n._base_node->_posd._x << " yd = " << n._base_node->_posd._y << "\n";
            return os << "x = " << n._pose2D._x << " y = " << n._pose2D._y << " heading = " << n._pose2D._heading << "\n"
                      << "cost_g = " << n._cost_g << " cost_h = " << (n._cost_f - n._cost_g) << " cost_f = " << n._cost_f << "\n"
                      << "vmin_sqr = " << n._vmin_sqr << " curvature_index = " << n._curvature_index
                      << " angle_bin = " << n._angle_bin << "\n" << std::endl;
        }

        struct HashFunction
        {
            std::size_t operator()(const Node3D<T>& n) const
            {
                std::size_t seed = 0;
                Node2D<T>::hash_mix(seed, n._base_node->_posd._x);
                Node2D<T>::hash_mix(seed, n._base_node->_posd._y);
                Node2D<T>::hash_mix(seed, n._angle_bin);
                return std::hash<std::size_t>()(seed);
            }
        };
    };
}

#endif
