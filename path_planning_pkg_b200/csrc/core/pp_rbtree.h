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
// `Node` must start with a 16-byte aligned PPWalk record `w` = {left, right, f, key} -- everything a
// walk step needs, fetched with ONE 128-bit load per level -- followed by int fields parent, color and the
// payload.  The ordering is the reference's for both of its sets (Node3D.h:45-54, Node2D.h:37-41):
//     a < b  <=>  (a.key != b.key) && (a.f < b.f)       key = cell (2D) or (cell, heading bin) (3D)
#ifndef PP_RBTREE_H
#define PP_RBTREE_H

#include "pp_defs.h"

#ifndef PP_EXACT_SPEC      /* speculative parallel walks (pp_search.h): a build variant, off by default */
#define PP_EXACT_SPEC 0
#endif

#define PP_RB_NIL (-1)
#define PP_RB_RED 0
#define PP_RB_BLACK 1
#define PP_RB_HEADER 0

struct
#if defined(__CUDACC__) || defined(__GNUC__)
__attribute__((aligned(16)))
#endif
PPWalk
{
    int      left, right;
    float    f;
    unsigned key;
};

struct PPKey { unsigned key; float f; };

// the node pools live in global memory: saying so lets the compiler emit LDG / STG instead of generic loads and stores
#ifdef __CUDA_ARCH__
#define PP_ASSUME_GLOBAL(ptr) __builtin_assume(__isGlobal(ptr))
#else
#define PP_ASSUME_GLOBAL(ptr)
#endif

// The walk record of a node with ONE 128-bit load.  Left to itself the compiler loads (f, key) first and the child index
// it needs afterwards -- two dependent loads per tree level on the latency-critical path of every find.
PP_HD PPWalk pp_walk_load(const PPWalk* p)
{
#ifdef __CUDA_ARCH__
    uint4 v = *reinterpret_cast<const uint4*>(p);
    asm volatile("" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w));     // all four words are live here: no narrowing of the load
    PPWalk r; r.left = (int)v.x; r.right = (int)v.y; r.f = __uint_as_float(v.z); r.key = v.w;
    return r;
#else
    return *p;
#endif
}

// the reference's non-strict-weak ordering
PP_HD bool pp_lt(unsigned ka, float fa, unsigned kb, float fb) { return (ka != kb) && (fa < fb); }

// Everything the tree algorithms touch of a node: both node types (PPNode3, PPNode2 in pp_search.h) start with it, so ONE copy
// of the walks and of the rebalancing code serves both trees (pool addressed as bytes + stride).  The search kernel is bound
// by instruction fetch when many warps are resident (ncu: `no_instruction` is the top stall at bench occupancy), so code that
// two containers can share is code that is fetched once.
struct PPRbHead
{
    PPWalk w;
    int    parent, color;
};

// container state without the pool pointer (the typed wrapper below owns that)
struct PPRbState
{
    int   cap;        // pool capacity (including header)
    int   next;       // first never-used slot
    int   free_head;  // singly linked (through .parent) list of recycled slots
    int   count;
    // Optional mutation log (speculative walks, pp_search.h): indices of the nodes whose child pointers changed since the
    // caller last reset *mcnt (the header stands for the root pointer).  *mcnt > mcap = too many / an erase: everything changed.
    int*  mlog = nullptr;
    int*  mcnt = nullptr;
    int   mcap = 0;

#if PP_EXACT_SPEC
    PP_HD void mut(int x) { if (mlog) { int c = *mcnt; if (c < mcap) mlog[c] = x; *mcnt = c + 1; } }
    PP_HD void mut_all() { if (mlog) *mcnt = mcap + 1; }
#else       // no speculative walks in this build: no log, and no test for one in every rotation
    PP_HD void mut(int) {}
    PP_HD void mut_all() {}
#endif
};

struct PPRbPool
{
    char* base; int stride;
    // unsigned 32 x 32 -> 64-bit product: ONE multiply-add per node address (a signed index costs a sign extension and a second
    // multiply on the dependent chain of every tree level); PP_RB_NIL is never dereferenced
    PP_HD PPRbHead& operator[](int i) const
    {
        return *reinterpret_cast<PPRbHead*>(base + (unsigned long long)(unsigned)i * (unsigned long long)(unsigned)stride);
    }
};

// child `side` of node x: 0 = left, 1 = right (PPWalk starts with {left, right})
PP_HD int& pp_rb_child(const PPRbPool n, int x, int side) { return (&n[x].w.left)[side]; }

// local_Rb_tree_rotate_left (up = 1: x's RIGHT child moves up) and local_Rb_tree_rotate_right (up = 0) of tree.cc as one body:
// the two are mirror images, and the kernel that runs them is bound by instruction fetch, so every mirrored pair below
// (rotations, the two halves of the insert and of the erase rebalancing loops) is written once over a side index.
PP_HD_NOINLINE_FN void pp_rb_rotate(PPRbState& t, const PPRbPool n, int x, int up)
{
    const int dn = up ^ 1;
    const int y = pp_rb_child(n, x, up);
    const int xp = n[x].parent;
    const bool x_is_root = (x == n[PP_RB_HEADER].parent);
    t.mut(x); t.mut(y); t.mut(x_is_root ? PP_RB_HEADER : xp);
    const int yd = pp_rb_child(n, y, dn);
    pp_rb_child(n, x, up) = yd;
    if (yd != PP_RB_NIL) n[yd].parent = x;
    n[y].parent = xp;
    if (x_is_root) n[PP_RB_HEADER].parent = y;
    else if (x == pp_rb_child(n, xp, dn)) pp_rb_child(n, xp, dn) = y;
    else pp_rb_child(n, xp, up) = y;
    pp_rb_child(n, y, dn) = x;
    n[x].parent = y;
}

// _Rb_tree_insert_and_rebalance
PP_HD_NOINLINE_FN void pp_rb_insert_and_rebalance(PPRbState& t, const PPRbPool n, bool insert_left, int x, int p)
{
    PP_ASSUME_GLOBAL(n.base);
    n[x].parent = p; n[x].w.left = PP_RB_NIL; n[x].w.right = PP_RB_NIL; n[x].color = PP_RB_RED;
    t.mut(p);
    if (insert_left)
    {
        n[p].w.left = x;   // also sets leftmost = x when p is the header
        if (p == PP_RB_HEADER) { n[PP_RB_HEADER].parent = x; n[PP_RB_HEADER].w.right = x; }
        else if (p == n[PP_RB_HEADER].w.left) n[PP_RB_HEADER].w.left = x;
    }
    else
    {
        n[p].w.right = x;
        if (p == n[PP_RB_HEADER].w.right) n[PP_RB_HEADER].w.right = x;
    }
    while (x != n[PP_RB_HEADER].parent && n[n[x].parent].color == PP_RB_RED)
    {
        const int xp = n[x].parent;
        const int xpp = n[xp].parent;
        const int side = (xp == n[xpp].w.left) ? 0 : 1;       // which child of the grandparent the parent is
        const int y = pp_rb_child(n, xpp, side ^ 1);          // the uncle
        if (y != PP_RB_NIL && n[y].color == PP_RB_RED)
        {
            n[xp].color = PP_RB_BLACK; n[y].color = PP_RB_BLACK; n[xpp].color = PP_RB_RED;
            x = xpp;
        }
        else
        {
            if (x == pp_rb_child(n, xp, side ^ 1)) { x = xp; pp_rb_rotate(t, n, x, side ^ 1); }
            n[n[x].parent].color = PP_RB_BLACK;
            n[xpp].color = PP_RB_RED;
            pp_rb_rotate(t, n, xpp, side);
        }
    }
    n[n[PP_RB_HEADER].parent].color = PP_RB_BLACK;
    t.count++;
}

// _Rb_tree_rebalance_for_erase; recycles slot z
PP_HD_NOINLINE_FN void pp_rb_erase(PPRbState& t, const PPRbPool n, int z)
{
    PP_ASSUME_GLOBAL(n.base);
    t.mut_all();
    int y = z, x = PP_RB_NIL, x_parent = PP_RB_NIL;
    if (n[y].w.left == PP_RB_NIL) x = n[y].w.right;
    else if (n[y].w.right == PP_RB_NIL) x = n[y].w.left;
    else
    {
        y = n[y].w.right;
        while (n[y].w.left != PP_RB_NIL) y = n[y].w.left;
        x = n[y].w.right;
    }
    if (y != z)
    {
        // relink y in place of z
        n[n[z].w.left].parent = y;
        n[y].w.left = n[z].w.left;
        if (y != n[z].w.right)
        {
            x_parent = n[y].parent;
            if (x != PP_RB_NIL) n[x].parent = n[y].parent;
            n[n[y].parent].w.left = x;
            n[y].w.right = n[z].w.right;
            n[n[z].w.right].parent = y;
        }
        else x_parent = y;
        if (n[PP_RB_HEADER].parent == z) n[PP_RB_HEADER].parent = y;
        else if (n[n[z].parent].w.left == z) n[n[z].parent].w.left = y;
        else n[n[z].parent].w.right = y;
        n[y].parent = n[z].parent;
        int c = n[y].color; n[y].color = n[z].color; n[z].color = c;
        y = z;
    }
    else
    {
        x_parent = n[y].parent;
        if (x != PP_RB_NIL) n[x].parent = n[y].parent;
        if (n[PP_RB_HEADER].parent == z) n[PP_RB_HEADER].parent = x;
        else if (n[n[z].parent].w.left == z) n[n[z].parent].w.left = x;
        else n[n[z].parent].w.right = x;
        if (n[PP_RB_HEADER].w.left == z)
        {
            if (n[z].w.right == PP_RB_NIL) n[PP_RB_HEADER].w.left = n[z].parent;
            else { int m = x; while (n[m].w.left != PP_RB_NIL) m = n[m].w.left; n[PP_RB_HEADER].w.left = m; }
        }
        if (n[PP_RB_HEADER].w.right == z)
        {
            if (n[z].w.left == PP_RB_NIL) n[PP_RB_HEADER].w.right = n[z].parent;
            else { int m = x; while (n[m].w.right != PP_RB_NIL) m = n[m].w.right; n[PP_RB_HEADER].w.right = m; }
        }
    }
    if (n[y].color != PP_RB_RED)
    {
        while (x != n[PP_RB_HEADER].parent && (x == PP_RB_NIL || n[x].color == PP_RB_BLACK))
        {
            const int side = (x == n[x_parent].w.left) ? 0 : 1;      // which child of x_parent x is; far = the other side
            const int far = side ^ 1;
            int w = pp_rb_child(n, x_parent, far);                   // the sibling
            if (n[w].color == PP_RB_RED)
            {
                n[w].color = PP_RB_BLACK; n[x_parent].color = PP_RB_RED;
                pp_rb_rotate(t, n, x_parent, far);
                w = pp_rb_child(n, x_parent, far);
            }
            const int wn = pp_rb_child(n, w, side), wf = pp_rb_child(n, w, far);      // the sibling's near and far child
            const bool wf_black = (wf == PP_RB_NIL || n[wf].color == PP_RB_BLACK);
            if (wf_black && (wn == PP_RB_NIL || n[wn].color == PP_RB_BLACK))
            {
                n[w].color = PP_RB_RED;
                x = x_parent;
                x_parent = n[x_parent].parent;
            }
            else
            {
                if (wf_black)
                {
                    n[wn].color = PP_RB_BLACK;
                    n[w].color = PP_RB_RED;
                    pp_rb_rotate(t, n, w, side);
                    w = pp_rb_child(n, x_parent, far);
                }
                n[w].color = n[x_parent].color;
                n[x_parent].color = PP_RB_BLACK;
                const int wf2 = pp_rb_child(n, w, far);
                if (wf2 != PP_RB_NIL) n[wf2].color = PP_RB_BLACK;
                pp_rb_rotate(t, n, x_parent, far);
                break;
            }
        }
        if (x != PP_RB_NIL) n[x].color = PP_RB_BLACK;
    }
    t.count--;
    if (t.count == 0)
    {
        // libstdc++ leaves header.left/right == header when the tree becomes empty
        n[PP_RB_HEADER].parent = PP_RB_NIL;
        n[PP_RB_HEADER].w.left = PP_RB_HEADER;
        n[PP_RB_HEADER].w.right = PP_RB_HEADER;
    }
    n[z].parent = t.free_head; t.free_head = z;        // release(z)
}

// ---- the two walks, optionally recording the nodes they visit ----------------------------------------------------------------
// A recorded path lets the caller decide later whether the walk would still go the same way: node keys and costs never change
// while a node is in the tree, so a walk is unchanged as long as no node on its path had a child pointer changed (PPRbState::mut).
// path[0] is always the header (= the root pointer).  n_path > cap means the path did not fit (treat as "cannot tell").

// std::set::find(k): lower-bound walk, then reject when k < *j.
PP_HD int pp_rb_find_walk(const PPRbPool n, const PPKey& k, int* path, int cap, int& n_path)
{
    PP_ASSUME_GLOBAL(n.base);
    int x = n[PP_RB_HEADER].parent, y = PP_RB_HEADER;
    PPWalk yw; yw.left = 0; yw.right = 0; yw.f = 0.0f; yw.key = 0u;
    int c = 0;
    if (c < cap) path[c] = PP_RB_HEADER;
    c++;
    while (x != PP_RB_NIL)
    {
        if (c < cap) path[c] = x;
        c++;
        const PPWalk r = pp_walk_load(&n[x].w);        // one 128-bit load per level
        if (!pp_lt(r.key, r.f, k.key, k.f)) { y = x; yw = r; x = r.left; }
        else x = r.right;
    }
    n_path = c;
    if (y == PP_RB_HEADER || pp_lt(k.key, k.f, yw.key, yw.f)) return PP_RB_NIL;
    return y;
}

// _M_get_insert_unique_pos.  libstdc++ steps to the in-order predecessor of the leaf position with _Rb_tree_decrement when the
// last turn was to the left; that predecessor is the last node at which the walk turned RIGHT (none: the position is left of
// begin()), so it is tracked during the walk instead of being looked up by a second pointer chase.
PP_HD bool pp_rb_insert_pos_walk(const PPRbPool n, const PPKey& k, int& p, bool& left, int* path, int cap, int& n_path)
{
    PP_ASSUME_GLOBAL(n.base);
    int x = n[PP_RB_HEADER].parent, y = PP_RB_HEADER;
    bool comp = true;
    int j = PP_RB_NIL;
    PPWalk jw; jw.left = 0; jw.right = 0; jw.f = 0.0f; jw.key = 0u;
    int c = 0;
    if (c < cap) path[c] = PP_RB_HEADER;
    c++;
    while (x != PP_RB_NIL)
    {
        if (c < cap) path[c] = x;
        c++;
        const PPWalk r = pp_walk_load(&n[x].w);
        y = x;
        comp = pp_lt(k.key, k.f, r.key, r.f);
        if (comp) x = r.left;
        else { j = x; jw = r; x = r.right; }
    }
    n_path = c;
    if (j == PP_RB_NIL)                    // never turned right: left of the leftmost node, or an empty tree
    {
        p = y; left = true;                // _M_insert_: p == header, or k < p (== comp, which is true here)
        return true;
    }
    if (pp_lt(jw.key, jw.f, k.key, k.f))
    {
        p = y;
        left = (y == PP_RB_HEADER) || comp;   // _M_insert_ re-evaluates k < p, which is `comp` of the last step
        return true;
    }
    return false;
}

// Out-of-line entry points of the two walks.  Key in, position out, all in registers: a by-reference key and by-reference results
// would make each call a round trip through the caller's stack frame (local memory) on the latency-critical path of the control lane.
PP_HD_NOINLINE_FN int pp_rb_find(const PPRbPool n, unsigned key, float f)
{
    PPKey k; k.key = key; k.f = f;
    int np;
    return pp_rb_find_walk(n, k, (int*)0, 0, np);
}

// returns the parent in the low word and, in the high word, 1 = attach as left child, 0 = as right child, -1 = an equivalent
// element exists (the insert is dropped)
PP_HD_NOINLINE_FN long long pp_rb_insert_pos(const PPRbPool n, unsigned key, float f)
{
    PPKey k; k.key = key; k.f = f;
    int np, p = PP_RB_NIL; bool left = false;
    const bool ins = pp_rb_insert_pos_walk(n, k, p, left, (int*)0, 0, np);
    const int how = ins ? (left ? 1 : 0) : -1;
    return (long long)(((unsigned long long)(unsigned)how << 32) | (unsigned long long)(unsigned)p);
}

// Typed view: `Node` starts with a PPRbHead-compatible prefix (PPWalk w; int parent, color;) followed by its payload.
template <class Node>
struct PPRbTree : PPRbState
{
    Node* n;        // pool; n[0] is the header

    PP_HD PPRbPool pool() const { PPRbPool p; p.base = (char*)n; p.stride = (int)sizeof(Node); return p; }

    PP_HD void init(Node* pool_, int capacity)
    {
        n = pool_; cap = capacity;
        clear();
    }

    PP_HD void clear()
    {
        n[PP_RB_HEADER].parent = PP_RB_NIL;
        n[PP_RB_HEADER].w.left = PP_RB_HEADER;
        n[PP_RB_HEADER].w.right = PP_RB_HEADER;
        n[PP_RB_HEADER].color = PP_RB_RED;
        next = 1; free_head = PP_RB_NIL; count = 0;
    }

    PP_HD bool empty() const { return count == 0; }
    PP_HD int  begin() const { return n[PP_RB_HEADER].w.left; }
    PP_HD int  root() const { return n[PP_RB_HEADER].parent; }

    // returns a free slot or PP_RB_NIL when the pool is exhausted
    PP_HD int alloc()
    {
        if (free_head != PP_RB_NIL) { int s = free_head; free_head = n[s].parent; return s; }
        if (next < cap) return next++;
        return PP_RB_NIL;
    }

    PP_HD void insert_and_rebalance(bool insert_left, int x, int p) { pp_rb_insert_and_rebalance(*this, pool(), insert_left, x, p); }
    PP_HD void erase(int z) { pp_rb_erase(*this, pool(), z); }
    PP_HD int  find(const PPKey& k) const { return pp_rb_find(pool(), k.key, k.f); }
    PP_HD bool insert_pos(const PPKey& k, int& p, bool& left) const
    {
        const long long r = pp_rb_insert_pos(pool(), k.key, k.f);
        const int how = (int)(r >> 32);
        p = (int)(unsigned)(r & 0xffffffffll); left = (how == 1);
        return how >= 0;
    }
};

#endif
