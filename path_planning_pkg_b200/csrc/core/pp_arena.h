// Device-side block allocator behind the growable per-query containers of the EXACT search.
//
// The reference's containers (std::set open list, unordered_set closed set, the lazy A*'s own set) are unbounded
// (lib/HybridAStar.cpp:103-193, lib/AStar.cpp:118-186); per-query expansion counts on the C4 workload span 3 orders of
// magnitude (p50 30 k, max 1.3 M).  Fixed worst-case pools per resident query would either cap the search or spread
// 2 368 resident queries over tens of GB (TLB / L2 thrash).  So every resident query starts on small fixed pools and
// moves a container that fills up into a block twice the size taken from this arena (power-of-two size classes, free
// lists per class, bump allocation for fresh blocks); blocks go back to their class list when the query ends.
// Allocation is rare (a handful per query) and done by one control lane, so a spin lock is enough.
#ifndef PP_ARENA_H
#define PP_ARENA_H

#include "pp_defs.h"

#define PP_ARENA_MIN_SHIFT 12      /* smallest block: 4 KB */
#define PP_ARENA_CLASSES 28        /* ... up to 2^39 bytes */

struct PPArena
{
    unsigned long long base;       // address of the arena memory
    unsigned long long size;       // bytes
    unsigned long long bump;       // first never-used byte offset
    int                lock;
    int                pad;
    unsigned long long free_head[PP_ARENA_CLASSES];   // per class: (byte offset + 1) of the first free block, 0 = none;
                                                      // a free block's first 8 bytes hold the next link
    unsigned long long n_alloc, n_fail, peak;         // statistics
};

// smallest class whose block holds `bytes`
PP_HD int pp_arena_class(unsigned long long bytes)
{
    int k = 0;
    while (k < PP_ARENA_CLASSES - 1 && (1ull << (k + PP_ARENA_MIN_SHIFT)) < bytes) k++;
    return k;
}

PP_HD unsigned long long pp_arena_block_bytes(int k) { return 1ull << (k + PP_ARENA_MIN_SHIFT); }

PP_HD void pp_arena_lock(PPArena* A)
{
#ifdef __CUDA_ARCH__
    while (atomicCAS(&A->lock, 0, 1) != 0) { __nanosleep(64); }
    __threadfence();
#else
    while (__sync_lock_test_and_set(&A->lock, 1)) {}
#endif
}

PP_HD void pp_arena_unlock(PPArena* A)
{
#ifdef __CUDA_ARCH__
    __threadfence();
    atomicExch(&A->lock, 0);
#else
    __sync_lock_release(&A->lock);
#endif
}

// One block of class k or, when neither its free list nor the never-used part of the arena has one, a free block of a larger
// class (blocks are never split or merged); nullptr when the arena is exhausted.  k_got = the class of the block handed out:
// the caller frees it with that class.  Call from ONE lane.
PP_HD_NOINLINE_FN void* pp_arena_alloc(PPArena* A, int k, int& k_got)
{
    k_got = k;
    if (!A || k >= PP_ARENA_CLASSES) return nullptr;
    const unsigned long long bytes = pp_arena_block_bytes(k);
    void* out = nullptr;
    pp_arena_lock(A);
    volatile unsigned long long* fh = A->free_head;
    unsigned long long h = fh[k];
    if (h != 0ull)
    {
        out = (void*)(A->base + (h - 1ull));
        fh[k] = *(volatile unsigned long long*)out;
    }
    else
    {
        volatile unsigned long long* bump = &A->bump;
        unsigned long long off = *bump;
        if (off + bytes <= A->size)
        {
            *bump = off + bytes;
            out = (void*)(A->base + off);
            if (off + bytes > A->peak) A->peak = off + bytes;
        }
        else
            for (int q = k + 1; q < PP_ARENA_CLASSES && !out; q++)
                if (fh[q] != 0ull)
                {
                    out = (void*)(A->base + (fh[q] - 1ull));
                    fh[q] = *(volatile unsigned long long*)out;
                    k_got = q;
                }
    }
    if (out) A->n_alloc++; else A->n_fail++;
    pp_arena_unlock(A);
    return out;
}

// pp_arena_alloc that waits for other queries to give blocks back when the arena is momentarily empty: resident queries finish
// (and free) all the time, so a short bounded wait often succeeds; only then does the caller give up (and the query is re-run
// later with fewer neighbours, pp_batch_wait).  A host lane never waits.
PP_HD_NOINLINE_FN void* pp_arena_alloc_patient(PPArena* A, int k, int& k_got)
{
    void* p = pp_arena_alloc(A, k, k_got);
#ifdef __CUDA_ARCH__
    for (int tries = 0; !p && tries < 2000; tries++)      // up to ~0.2 s
    {
        __nanosleep(100000);
        p = pp_arena_alloc(A, k, k_got);
    }
#endif
    return p;
}

PP_HD_NOINLINE_FN void pp_arena_free(PPArena* A, void* p, int k)
{
    if (!A || !p) return;
    pp_arena_lock(A);
    volatile unsigned long long* fh = A->free_head;
    *(volatile unsigned long long*)p = fh[k];
    fh[k] = ((unsigned long long)p - A->base) + 1ull;
    pp_arena_unlock(A);
}

#endif
