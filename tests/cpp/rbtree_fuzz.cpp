// TEST DRIVER: the product's index-based red-black tree (path_planning_pkg_b200/csrc/core/pp_rbtree.h, the EXACT mode's
// open lists) against libstdc++'s std::set driven by the reference's comparator `(a != b) && (a.f < b.f)` (Node3D.h:45-54,
// Node2D.h:37-41) -- NOT a strict weak ordering, so which insert is dropped and which element find() returns depends on the
// tree's shape and on libstdc++'s exact walks.  Random sequences of insert / find / erase / pop-min over
// small key and cost domains (many equal costs, many repeated keys); after EVERY operation the two structures must agree on
// the operation's outcome, on the size, and node for node on shape, colours, keys and costs (std::set's nodes are walked through
// the public _M_node member of its iterator and _Rb_tree_node_base).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>

#include "../../path_planning_pkg_b200/csrc/core/pp_search.h"     // PPNode2 (a node type of the product) + pp_rbtree.h

struct Item { unsigned key; float f; };
struct RefLess { bool operator()(const Item& a, const Item& b) const { return (a.key != b.key) && (a.f < b.f); } };
typedef std::set<Item, RefLess> RefSet;

static bool same_subtree(const std::_Rb_tree_node_base* a, const PPRbTree<PPNode2>& T, int b, long& visited)
{
    if (a == nullptr || b == PP_RB_NIL) return a == nullptr && b == PP_RB_NIL;
    const Item& ia = *static_cast<const std::_Rb_tree_node<Item>*>(a)->_M_valptr();
    const PPNode2& nb = T.n[b];
    if (ia.key != nb.w.key || ia.f != nb.w.f) return false;
    if ((a->_M_color == std::_S_red) != (nb.color == PP_RB_RED)) return false;
    visited++;
    return same_subtree(a->_M_left, T, nb.w.left, visited) && same_subtree(a->_M_right, T, nb.w.right, visited);
}

static bool same_tree(const RefSet& S, const PPRbTree<PPNode2>& T)
{
    if ((int)S.size() != T.count) return false;
    if (S.empty()) return T.root() == PP_RB_NIL;
    // root of a std::set: walk up from begin()
    const std::_Rb_tree_node_base* r = S.begin()._M_node;
    while (r->_M_parent->_M_parent != r || r->_M_color == std::_S_red) r = r->_M_parent;     // header is red and header.parent.parent == header
    if (r->_M_parent->_M_parent != r) return false;
    long visited = 0;
    if (!same_subtree(r, T, T.root(), visited) || visited != (long)S.size()) return false;
    const Item& lo = *S.begin();
    return T.n[T.begin()].w.key == lo.key && T.n[T.begin()].w.f == lo.f;
}

int main(int argc, char** argv)
{
    const int rounds = argc > 1 ? std::atoi(argv[1]) : 40;
    long ops = 0, drops = 0, hits = 0;
    int largest = 0;
    for (int round = 0; round < rounds; round++)
    {
        std::mt19937 rng(1000 + round);
        const unsigned key_space = 8u << (round % 10);           // 8 .. 4096 distinct keys
        const unsigned f_space = 4u << (round % 12);             // 4 .. 8192 distinct costs (equal costs are "equivalent": dropped)
        const int cap = 8192;
        std::vector<PPNode2> pool(cap);
        PPRbTree<PPNode2> T; T.init(pool.data(), cap);
        RefSet S;
        for (int step = 0; step < 6000; step++, ops++)
        {
            unsigned what = rng() % 16;
            Item it; it.key = rng() % key_space; it.f = 0.25f * (float)(rng() % f_space);
            PPKey k; k.key = it.key; k.f = it.f;
            if (what < 9 && (int)S.size() < cap - 2)              // insert_unique
            {
                bool ref_ins = S.insert(it).second;
                int p; bool left;
                bool ins = T.insert_pos(k, p, left);
                if (ins)
                {
                    int s = T.alloc();
                    T.n[s].w.key = it.key; T.n[s].w.f = it.f; T.n[s].g = 0.0f; T.n[s].prev = -1;
                    T.insert_and_rebalance(left, s, p);
                }
                else drops++;
                if (ins != ref_ins) { std::printf("round %d step %d: insert outcome differs\n", round, step); return 1; }
            }
            else if (what < 12)                                   // find, erase what was found (HybridAStar.cpp:165-191)
            {
                auto rf = S.find(it);
                int f = T.find(k);
                if ((rf == S.end()) != (f == PP_RB_NIL)) { std::printf("round %d step %d: find hit/miss differs\n", round, step); return 1; }
                if (f != PP_RB_NIL)
                {
                    hits++;
                    if (rf->key != T.n[f].w.key || rf->f != T.n[f].w.f) { std::printf("round %d step %d: find returns another element\n", round, step); return 1; }
                    if (what == 11) { S.erase(rf); T.erase(f); }
                }
            }
            else if (!S.empty())                                  // pop the minimum (HybridAStar.cpp:110-116)
            {
                int b = T.begin();
                if (S.begin()->key != T.n[b].w.key || S.begin()->f != T.n[b].w.f) { std::printf("round %d step %d: begin differs\n", round, step); return 1; }
                S.erase(S.begin()); T.erase(b);
            }
            if (T.count > largest) largest = T.count;
            if (!same_tree(S, T)) { std::printf("round %d step %d: trees differ (size %zu vs %d)\n", round, step, S.size(), T.count); return 1; }
        }
    }
    std::printf("rbtree_fuzz ok: %ld operations, %ld dropped inserts, %ld find hits, largest tree %d nodes, trees identical after every operation\n", ops, drops, hits, largest);
    return 0;
}
