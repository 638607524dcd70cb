// Host-side table builder of the generic footprint collision check (see core/pp_footprint.h for the definition).
#ifndef PP_FOOTPRINT_HOST_H
#define PP_FOOTPRINT_HOST_H

#include "../core/pp_footprint.h"
#include <algorithm>
#include <cmath>
#include <vector>

// Builds the table described above; `bins_out` gets C.bins + 1 entries.  Returns the largest window side.
inline int pp_footprint_build(const PPConsts& C, float length, float width, float rear, std::vector<PPFootBin>& bins_out,
                              std::vector<PPCellOff>& offs_out)
{
    bins_out.clear(); offs_out.clear();
    const double half = 0.5 * (double)C.res;
    const int na = (length > 0.0f) ? (int)std::ceil((double)length / half) + 1 : 1;
    const int nb = (width > 0.0f) ? (int)std::ceil((double)width / half) + 1 : 1;
    int win = 1;
    for (int b = 0; b < C.bins; b++)
    {
        const double theta = -PP_PI + (double)b * (double)C.precision;
        const double c = std::cos(theta), s = std::sin(theta);
        std::vector<std::pair<int, int>> cells;
        for (int a = 0; a < na; a++)
        {
            double u = std::min(-(double)rear + a * half, (double)length - (double)rear);
            for (int q = 0; q < nb; q++)
            {
                double v = std::min(-0.5 * (double)width + q * half, 0.5 * (double)width);
                int di = (int)std::lround((u * c - v * s) / (double)C.res);
                int dj = (int)std::lround((u * s + v * c) / (double)C.res);
                cells.emplace_back(di, dj);
            }
        }
        std::sort(cells.begin(), cells.end());
        cells.erase(std::unique(cells.begin(), cells.end()), cells.end());
        PPFootBin fb;
        fb.first = (int)offs_out.size(); fb.count = (int)cells.size();
        int imin = cells[0].first, imax = imin, jmin = cells[0].second, jmax = jmin;
        for (auto& p : cells)
        {
            imin = std::min(imin, p.first); imax = std::max(imax, p.first);
            jmin = std::min(jmin, p.second); jmax = std::max(jmax, p.second);
            PPCellOff o; o.di = (short)p.first; o.dj = (short)p.second;
            offs_out.push_back(o);
        }
        fb.imin = (short)imin; fb.imax = (short)imax; fb.jmin = (short)jmin; fb.jmax = (short)jmax;
        win = std::max(win, std::max(imax - imin + 1, jmax - jmin + 1));
        bins_out.push_back(fb);
    }
    bins_out.push_back(bins_out[0]);     // column `bins` == column 0
    return win;
}

#endif
