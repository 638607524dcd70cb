// Per-cell arithmetic of the probabilistic occupancy-map update (log-odds Grid2D).
//
// Reference: Grid2D::update_obstacles(boxes, conf) Grid2D.cpp:99-139 is a SCATTER: for every box, for
// every half-cell sample (i, j) of the box, rotate the offset (0.5 i, 0.5 j) by the grid heading,
// round, and do `v = clamp(v + delta)` on the hit cell -- a cell hit n times is incremented and
// clamped n times, boxes in input order.  The device kernel is the equivalent GATHER: each cell, for
// each box overlapping its tile (in input order), counts the samples that land on it and applies
// clamp(v + delta) that many times.  The count is found by testing the few lattice points around
// the inverse-rotated cell offset with the reference's exact forward arithmetic, so the result is
// bit-identical to the scatter.
#ifndef PP_MAP_H
#define PP_MAP_H

#include "pp_defs.h"

// std::max(std::min(v, hi), lo), Grid2D.cpp:134, :204
PP_HD float pp_clamp_logodds(float v, float lo, float hi)
{
    float m = (hi < v) ? hi : v;
    return (m < lo) ? lo : m;
}

// whole-map decay, Grid2D::update_obstacles() Grid2D.cpp:197-208
PP_HD float pp_map_decay_cell(float v, float log_free, float lo, float hi)
{
    return pp_clamp_logodds(v + log_free, lo, hi);
}

// forward map of one half-cell sample, Grid2D.cpp:127-130: offset (0.5 i, 0.5 j) rotated by the grid heading, rounded
PP_HD void pp_box_sample_offset(int i, int j, float cos_h, float sin_h, int& a, int& b)
{
    float ox = (float)(i * 0.5), oy = (float)(j * 0.5);
    float rx = ox * cos_h + oy * sin_h;
    float ry = -ox * sin_h + oy * cos_h;
    a = (int)roundf(rx);
    b = (int)roundf(ry);
}

// number of samples (i, j) in [0, ni) x [0, nj) whose rounded rotated offset equals (a, b)
PP_HD int pp_box_count(int ni, int nj, float cos_h, float sin_h, int a, int b)
{
    // Inverse rotation of the cell offset gives the lattice neighbourhood to test.  The pre-image of the unit rounding square
    // around (a, b) is a rotated unit square; in (i, j) lattice units (half cells) its bounding box reaches at most
    // |cos| + |sin| <= 1.4143 from the centre (ic, jc), so i in [ic - 1.4143, ic + 1.4143] is a subset of {i0 - 1, i0, i0 + 1}
    // with i0 = round(ic) (|ic - i0| <= 0.5, margin 0.085 -- the float error of ic, jc is below 1e-3 for offsets up to 2^13).
    // Every candidate is then checked with the reference's exact forward arithmetic, so a superset is all that is needed.
    float ic = 2.0f * ((float)a * cos_h - (float)b * sin_h);
    float jc = 2.0f * ((float)a * sin_h + (float)b * cos_h);
    int i0 = (int)roundf(ic), j0 = (int)roundf(jc);
    int count = 0;
    for (int di = -1; di <= 1; di++)
    {
        int i = i0 + di;
        if (i < 0 || i >= ni) continue;
        for (int dj = -1; dj <= 1; dj++)
        {
            int j = j0 + dj;
            if (j < 0 || j >= nj) continue;
            int sa, sb;
            pp_box_sample_offset(i, j, cos_h, sin_h, sa, sb);
            if (sa == a && sb == b) count++;
        }
    }
    return count;
}

// `count` applications of v = clamp(v + delta), Grid2D.cpp:133-134 (count is ~4 for half-cell sampling)
PP_HD float pp_box_apply(float v, int count, float delta, float lo, float hi)
{
    for (int c = 0; c < count; c++) v = pp_clamp_logodds(v + delta, lo, hi);
    return v;
}

// ---- lane-line rasteriser, Grid2D::update_obstacles(lines, conf, width) Grid2D.cpp:142-194 ----
// Host prologue per line (Grid2D.cpp:147-156): rotated end points, unit direction / normal, length.
struct PPLineDesc
{
    float sx, sy;        // start_point (grid frame, metres relative to the goal)
    float ux, uy;        // delta / line_length
    float nx, ny;        // delta_normal
    float length;        // line_length
    float delta;         // log(c/(1-c)) - log_free
};

// prog_length after `t` iterations of `prog_length += res` from 0 (float accumulation, Grid2D.cpp:188)
PP_HD float pp_accumulate_steps(float step, int t)
{
    float acc = 0.0f;
    for (int q = 0; q < t; q++) acc += step;
    return acc;
}

// The two sample cells of length-step value `pl` and width-step value `pw` (Grid2D.cpp:163-170).
PP_HD void pp_line_cells(const PPLineDesc& d, float res, int n45, int n2, float pl, float pw,
                         int& i1, int& j1, int& i2, int& j2)
{
    float ix = d.sx + d.ux * pl, iy = d.sy + d.uy * pl;         // intercept
    float p1x = ix + d.nx * pw, p1y = iy + d.ny * pw;
    float p2x = ix - d.nx * pw, p2y = iy - d.ny * pw;
    // x86 float -> int (pp_f2i_x86): a degenerate line (start == end: 0/0 direction) or a NaN point gives INT_MIN + n, which
    // fails the caller's `> -1` test exactly like the reference; the unsigned add keeps that wrap-around defined
    i1 = (int)((unsigned)pp_f2i_x86(roundf(p1x / res)) + (unsigned)n45);
    i2 = (int)((unsigned)pp_f2i_x86(roundf(p2x / res)) + (unsigned)n45);
    j1 = (int)((unsigned)pp_f2i_x86(roundf(p1y / res)) + (unsigned)n2);
    j2 = (int)((unsigned)pp_f2i_x86(roundf(p2y / res)) + (unsigned)n2);
}

// ---- map relocation on goal change, Grid3D::relocate_obstacles Grid3D.cpp:169-203 ----
// Host prologue: cos/sin of (new heading - old heading) and origin_new_to_old.  Forward scatter
// "last writer in raster order wins" is reproduced on the device by an atomicMax of the source index.
struct PPRelocDesc
{
    float cos_d, sin_d;
    float ox, oy;        // origin_new_to_old
};

PP_HD void pp_reloc_target(const PPRelocDesc& d, int i, int j, int& i_new, int& j_new)
{
    float x = (float)i, y = (float)j;
    float rx = x * d.cos_d + y * d.sin_d;
    float ry = -x * d.sin_d + y * d.cos_d;
    float lx = rx + d.ox, ly = ry + d.oy;
    i_new = (int)roundf(lx);
    j_new = (int)roundf(ly);
}

#endif
