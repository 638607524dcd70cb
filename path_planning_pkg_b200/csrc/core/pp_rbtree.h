// Index-based red-black tree that reproduces libstdc++'s std::set *walks and rebalancing* exactly.
//
// Why: the reference keeps its open lists in std::set with a comparator that is not a strict weak
// ordering (Node3D.h:45-54, Node2D.h:37-41: `(a != b) && (a.f < b.f)`).  Which element a find()
// returns and whether an insert() is silently dropped then depends on the SHAPE of the tree and on
// the exact comparison sequence of libstdc++, so "identical expansion sequence"
// needs the same tree, not just any priority queue.  The algorithms below restate, with 32-bit
// indices into a per-query node pool instead of pointers:
//   find             = _M_lower_bound walk + final check     (bits/stl_tree.h, _Rb_tree::find)
//   insert_unique    = _M_get_insert_unique_pos + _M_insert_ (bits/stl_tree.h)
//   insert rebalance = _Rb_tree_insert_and_rebalance         (libstdc++ src/c++98/tree.cc)
//   erase            = _Rb_tree_rebalance_for_erase          (libstdc++ src/c++98/tree.cc)
//   decrement        = _Rb_tree_decrement
// Node 0 of the pool is the header (parent = root, left = leftmost, right = rightmost, red).
//
// `Node` must provide int fields parent, left, right, color.  `Less` is a functor
// bool(const Key&, const Node&) / bool(const Node&, const Key&) supplied by the caller.
#ifndef PP_RBTREE_H
#define PP_RBTREE_H

#include "pp_defs.h"

#define PP_RB_NIL (-1)
#define PP_RB_RED 0
#define PP_RB_BLACK 1
#define PP_RB_HEADER 0

template <class Node>
struct PPRbTree
{
    Node* n;        // pool; n[0] is the header
    int   cap;      // pool capacity (including header)
    int   next;     // first never-used slot
    int   free_head;  // singly linked (through .parent) list of recycled slots
    int   count;

    PP_HD void init(Node* pool, int capacity)
    {
        n = pool; cap = capacity;
        clear();
    }

    PP_HD void clear()
    {
        n[PP_RB_HEADER].parent = PP_RB_NIL;
        n[PP_RB_HEADER].left = PP_RB_HEADER;
        n[PP_RB_HEADER].right = PP_RB_HEADER;
        n[PP_RB_HEADER].color = PP_RB_RED;
        next = 1; free_head = PP_RB_NIL; count = 0;
    }

    PP_HD bool empty() const { return count == 0; }
    PP_HD int  begin() const { return n[PP_RB_HEADER].left; }
    PP_HD int  root() const { return n[PP_RB_HEADER].parent; }

    // returns a free slot or PP_RB_NIL when the pool is exhausted
    PP_HD int alloc()
    {
        if (free_head != PP_RB_NIL) { int s = free_head; free_head = n[s].parent; return s; }
        if (next < cap) return next++;
        return PP_RB_NIL;
    }

    PP_HD void release(int s) { n[s].parent = free_head; free_head = s; }

    PP_HD void rotate_left(int x)
    {
        int y = n[x].right;
        n[x].right = n[y].left;
        if (n[y].left != PP_RB_NIL) n[n[y].left].parent = x;
        n[y].parent = n[x].parent;
        if (x == n[PP_RB_HEADER].parent) n[PP_RB_HEADER].parent = y;
        else if (x == n[n[x].parent].left) n[n[x].parent].left = y;
        else n[n[x].parent].right = y;
        n[y].left = x;
        n[x].parent = y;
    }

    PP_HD void rotate_right(int x)
    {
        int y = n[x].left;
        n[x].left = n[y].right;
        if (n[y].right != PP_RB_NIL) n[n[y].right].parent = x;
        n[y].parent = n[x].parent;
        if (x == n[PP_RB_HEADER].parent) n[PP_RB_HEADER].parent = y;
        else if (x == n[n[x].parent].right) n[n[x].parent].right = y;
        else n[n[x].parent].left = y;
        n[y].right = x;
        n[x].parent = y;
    }

    // _Rb_tree_decrement
    PP_HD int decrement(int x) const
    {
        if (x == PP_RB_HEADER) return n[x].right;   // end() -> rightmost
        if (n[x].left != PP_RB_NIL)
        {
            int y = n[x].left;
            while (n[y].right != PP_RB_NIL) y = n[y].right;
            return y;
        }
        int y = n[x].parent;
        while (x == n[y].left) { x = y; y = n[y].parent; }
        return y;
    }

    // _Rb_tree_insert_and_rebalance
    PP_HD void insert_and_rebalance(bool insert_left, int x, int p)
    {
        n[x].parent = p; n[x].left = PP_RB_NIL; n[x].right = PP_RB_NIL; n[x].color = PP_RB_RED;
        if (insert_left)
        {
            n[p].left = x;   // also sets leftmost = x when p is the header
            if (p == PP_RB_HEADER) { n[PP_RB_HEADER].parent = x; n[PP_RB_HEADER].right = x; }
            else if (p == n[PP_RB_HEADER].left) n[PP_RB_HEADER].left = x;
        }
        else
        {
            n[p].right = x;
            if (p == n[PP_RB_HEADER].right) n[PP_RB_HEADER].right = x;
        }
        while (x != n[PP_RB_HEADER].parent && n[n[x].parent].color == PP_RB_RED)
        {
            int xp = n[x].parent;
            int xpp = n[xp].parent;
            if (xp == n[xpp].left)
            {
                int y = n[xpp].right;
                if (y != PP_RB_NIL && n[y].color == PP_RB_RED)
                {
                    n[xp].color = PP_RB_BLACK; n[y].color = PP_RB_BLACK; n[xpp].color = PP_RB_RED;
                    x = xpp;
                }
                else
                {
                    if (x == n[xp].right) { x = xp; rotate_left(x); }
                    n[n[x].parent].color = PP_RB_BLACK;
                    n[xpp].color = PP_RB_RED;
                    rotate_right(xpp);
                }
            }
            else
            {
                int y = n[xpp].left;
                if (y != PP_RB_NIL && n[y].color == PP_RB_RED)
                {
                    n[xp].color = PP_RB_BLACK; n[y].color = PP_RB_BLACK; n[xpp].color = PP_RB_RED;
                    x = xpp;
                }
                else
                {
                    if (x == n[xp].left) { x = xp; rotate_right(x); }
                    n[n[x].parent].color = PP_RB_BLACK;
                    n[xpp].color = PP_RB_RED;
                    rotate_left(xpp);
                }
            }
        }
        n[n[PP_RB_HEADER].parent].color = PP_RB_BLACK;
        count++;
    }

    // _Rb_tree_rebalance_for_erase; recycles slot z
    PP_HD void erase(int z)
    {
        int y = z, x = PP_RB_NIL, x_parent = PP_RB_NIL;
        if (n[y].left == PP_RB_NIL) x = n[y].right;
        else if (n[y].right == PP_RB_NIL) x = n[y].left;
        else
        {
            y = n[y].right;
            while (n[y].left != PP_RB_NIL) y = n[y].left;
            x = n[y].right;
        }
        if (y != z)
        {
            // relink y in place of z
            n[n[z].left].parent = y;
            n[y].left = n[z].left;
            if (y != n[z].right)
            {
                x_parent = n[y].parent;
                if (x != PP_RB_NIL) n[x].parent = n[y].parent;
                n[n[y].parent].left = x;
                n[y].right = n[z].right;
                n[n[z].right].parent = y;
            }
            else x_parent = y;
            if (n[PP_RB_HEADER].parent == z) n[PP_RB_HEADER].parent = y;
            else if (n[n[z].parent].left == z) n[n[z].parent].left = y;
            else n[n[z].parent].right = y;
            n[y].parent = n[z].parent;
            int c = n[y].color; n[y].color = n[z].color; n[z].color = c;
            y = z;
        }
        else
        {
            x_parent = n[y].parent;
            if (x != PP_RB_NIL) n[x].parent = n[y].parent;
            if (n[PP_RB_HEADER].parent == z) n[PP_RB_HEADER].parent = x;
            else if (n[n[z].parent].left == z) n[n[z].parent].left = x;
            else n[n[z].parent].right = x;
            if (n[PP_RB_HEADER].left == z)
            {
                if (n[z].right == PP_RB_NIL) n[PP_RB_HEADER].left = n[z].parent;
                else { int m = x; while (n[m].left != PP_RB_NIL) m = n[m].left; n[PP_RB_HEADER].left = m; }
            }
            if (n[PP_RB_HEADER].right == z)
            {
                if (n[z].left == PP_RB_NIL) n[PP_RB_HEADER].right = n[z].parent;
                else { int m = x; while (n[m].right != PP_RB_NIL) m = n[m].right; n[PP_RB_HEADER].right = m; }
            }
        }
        if (n[y].color != PP_RB_RED)
        {
            while (x != n[PP_RB_HEADER].parent && (x == PP_RB_NIL || n[x].color == PP_RB_BLACK))
            {
                if (x == n[x_parent].left)
                {
                    int w = n[x_parent].right;
                    if (n[w].color == PP_RB_RED)
                    {
                        n[w].color = PP_RB_BLACK; n[x_parent].color = PP_RB_RED;
                        rotate_left(x_parent);
                        w = n[x_parent].right;
                    }
                    if ((n[w].left == PP_RB_NIL || n[n[w].left].color == PP_RB_BLACK) &&
                        (n[w].right == PP_RB_NIL || n[n[w].right].color == PP_RB_BLACK))
                    {
                        n[w].color = PP_RB_RED;
                        x = x_parent;
                        x_parent = n[x_parent].parent;
                    }
                    else
                    {
                        if (n[w].right == PP_RB_NIL || n[n[w].right].color == PP_RB_BLACK)
                        {
                            n[n[w].left].color = PP_RB_BLACK;
                            n[w].color = PP_RB_RED;
                            rotate_right(w);
                            w = n[x_parent].right;
                        }
                        n[w].color = n[x_parent].color;
                        n[x_parent].color = PP_RB_BLACK;
                        if (n[w].right != PP_RB_NIL) n[n[w].right].color = PP_RB_BLACK;
                        rotate_left(x_parent);
                        break;
                    }
                }
                else
                {
                    int w = n[x_parent].left;
                    if (n[w].color == PP_RB_RED)
                    {
                        n[w].color = PP_RB_BLACK; n[x_parent].color = PP_RB_RED;
                        rotate_right(x_parent);
                        w = n[x_parent].left;
                    }
                    if ((n[w].right == PP_RB_NIL || n[n[w].right].color == PP_RB_BLACK) &&
                        (n[w].left == PP_RB_NIL || n[n[w].left].color == PP_RB_BLACK))
                    {
                        n[w].color = PP_RB_RED;
                        x = x_parent;
                        x_parent = n[x_parent].parent;
                    }
                    else
                    {
                        if (n[w].left == PP_RB_NIL || n[n[w].left].color == PP_RB_BLACK)
                        {
                            n[n[w].right].color = PP_RB_BLACK;
                            n[w].color = PP_RB_RED;
                            rotate_left(w);
                            w = n[x_parent].left;
                        }
                        n[w].color = n[x_parent].color;
                        n[x_parent].color = PP_RB_BLACK;
                        if (n[w].left != PP_RB_NIL) n[n[w].left].color = PP_RB_BLACK;
                        rotate_right(x_parent);
                        break;
                    }
                }
            }
            if (x != PP_RB_NIL) n[x].color = PP_RB_BLACK;
        }
        count--;
        if (count == 0)
        {
            // libstdc++ leaves header.left/right == header when the tree becomes empty
            n[PP_RB_HEADER].parent = PP_RB_NIL;
            n[PP_RB_HEADER].left = PP_RB_HEADER;
            n[PP_RB_HEADER].right = PP_RB_HEADER;
        }
        release(z);
    }

    // std::set::find(k): lower-bound walk, then reject when k < *j.  lt_nk(node, key), lt_kn(key, node).
    template <class Key, class LtNK, class LtKN>
    PP_HD int find(const Key& k, LtNK lt_nk, LtKN lt_kn) const
    {
        int x = n[PP_RB_HEADER].parent, y = PP_RB_HEADER;
        while (x != PP_RB_NIL)
        {
            if (!lt_nk(n[x], k)) { y = x; x = n[x].left; }
            else x = n[x].right;
        }
        if (y == PP_RB_HEADER || lt_kn(k, n[y])) return PP_RB_NIL;
        return y;
    }

    // std::set::insert(v) position search (_M_get_insert_unique_pos).  Returns true when the key must be
    // inserted under parent `p` (left child iff `left`); false when an equivalent element exists.
    template <class Key, class LtNK, class LtKN>
    PP_HD bool insert_pos(const Key& k, LtNK lt_nk, LtKN lt_kn, int& p, bool& left) const
    {
        int x = n[PP_RB_HEADER].parent, y = PP_RB_HEADER;
        bool comp = true;
        while (x != PP_RB_NIL)
        {
            y = x;
            comp = lt_kn(k, n[x]);
            x = comp ? n[x].left : n[x].right;
        }
        int j = y;
        if (comp)
        {
            if (j == n[PP_RB_HEADER].left)   // j == begin()
            {
                p = y; left = true;          // _M_insert_: p == header or k < p
                left = (y == PP_RB_HEADER) || lt_kn(k, n[y]);
                return true;
            }
            j = decrement(j);
        }
        if (lt_nk(n[j], k))
        {
            p = y;
            left = (y == PP_RB_HEADER) || lt_kn(k, n[y]);
            return true;
        }
        return false;
    }
};

#endif
