// Generic vehicle-footprint collision check (north_star (c)).
//
// The reference's collision check is a single-cell lookup: a successor is kept iff the cell under its position,
// (int(x/res), int(y/res)), is inside the grid and below the occupancy threshold (lib/Grid3D.cpp:53-59); the vehicle's
// size is handled upstream by inflating obstacles (src/local_planner.cpp:230-231, :287).  That check is the footprint
// {(0, 0)} of the general form built here: per heading bin a list of cell offsets (di, dj) around the pose's cell, a pose
// is free iff EVERY cell (ci + di, cj + dj) is inside the grid and below the threshold.  With the table of a zero-size
// vehicle the result equals the reference's check bit for bit (booleans and cells); with a real length x width it checks
// the un-inflated map against the oriented rectangle.
//
// Table (host, pp_footprint_build in host/pp_footprint_host.h): for bin b the rectangle [-rear, length - rear] x [-width/2, +width/2] in the vehicle
// frame (x forward, origin = the pose's reference point) is sampled every half cell along both axes, far edges included,
// like the reference's box rasteriser samples obstacle boxes (lib/Grid2D.cpp:110-133); every sample is rotated by the bin's
// heading -pi + b*precision and rounded to a cell offset; duplicates are dropped, offsets sorted (di, dj).  Column `bins`
// (the reference's out-of-range bin) repeats bin 0.
// Kernel (pp_footprint_kernel): one warp per pose; the bounding window of the bin's offsets is staged from the map into
// shared memory with row-contiguous loads (out-of-grid cells staged as +inf = blocked), lanes test the footprint cells
// against the staged tile, the verdict is a warp ballot.
#ifndef PP_FOOTPRINT_H
#define PP_FOOTPRINT_H

#include "pp_defs.h"
#include "pp_math.h"

#define PP_FOOT_MAX_WIN 96      // largest supported window side (cells): 96*96*4 B = 36 KB of shared memory per warp

struct alignas(4) PPCellOff { short di, dj; };      // one 32-bit load

struct alignas(16) PPFootBin      // one 128-bit load
{
    short imin, imax, jmin, jmax;    // bounding box of the bin's offsets
    int   first, count;              // slice of the offset list
};

// heading bin of a pose for the table: the reference's index (common.h:31-36), its out-of-range value folded onto bin 0
PP_HD int pp_foot_bin(float heading, float precision, int bins)
{
    int b = pp_heading_index(heading, precision);
    if (b < 0) b = 0;
    if (b >= bins) b = 0;
    return b;
}

// the per-cell predicate of Grid3D.cpp:56-59 on an explicit cell; true = blocked (outside the grid counts as blocked)
PP_HD bool pp_foot_cell_blocked(const PPConsts& C, const float* map, int i, int j)
{
    if (!((i > -1) && (i < C.N) && (j > -1) && (j < C.N))) return true;
    return !(map[i * C.N + j] < C.log_thr);
}

#endif
