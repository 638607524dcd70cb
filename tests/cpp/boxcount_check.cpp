// TEST ONLY.  The gather form of the box rasteriser (pp_box_count, csrc/core/pp_map.h: per cell, test the 3 x 3 lattice points around the
// inverse-rotated cell offset with the exact forward arithmetic of Grid2D.cpp:127-130) against the scatter the reference performs, for
// random grid headings and box sizes: every cell of a 121 x 121 window must get the same sample count.
#include <cstdio>
#include <cmath>
#include <vector>
#include <random>
#include "../../path_planning_pkg_b200/csrc/core/pp_map.h"
int main()
{
    std::mt19937 rng(7);
    long long checked = 0, bad = 0;
    for (int trial = 0; trial < 1500; trial++)
    {
        float h = (float)((rng() >> 8) * (1.0 / 16777216.0) * 6.283185307179586 - 3.141592653589793);
        if (trial < 8) h = (float)(trial * 0.7853981633974483);           // axis-aligned and diagonal frames
        float c = std::cos(h), s = std::sin(h);
        int ni = 2 + rng() % 70, nj = 2 + rng() % 70;
        const int R = 60;
        std::vector<int> cnt((2 * R + 1) * (2 * R + 1), 0);
        for (int i = 0; i < ni; i++)
            for (int j = 0; j < nj; j++)
            {
                int a, b; pp_box_sample_offset(i, j, c, s, a, b);
                if (a >= -R && a <= R && b >= -R && b <= R) cnt[(a + R) * (2 * R + 1) + (b + R)]++;
            }
        for (int a = -R; a <= R; a++)
            for (int b = -R; b <= R; b++)
            {
                int g = pp_box_count(ni, nj, c, s, a, b);
                checked++;
                if (g != cnt[(a + R) * (2 * R + 1) + (b + R)]) { if (bad++ < 5) std::printf("mismatch h=%g ni=%d nj=%d a=%d b=%d gather %d scatter %d\n", h, ni, nj, a, b, g, cnt[(a + R) * (2 * R + 1) + (b + R)]); }
            }
    }
    std::printf("checked %lld cells, mismatches %lld\n", checked, bad);
    return bad != 0;
}
