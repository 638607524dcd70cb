// TEST DRIVER for SURVEY.md §8(f) rows N3 / N4: VelocityGenerator::generate_velocity_profile
// (reference lib/VelocityGenerator.cpp:19-85) and PedestrianHandler::calc_max_velocity
// (reference lib/PedestrianHandler.cpp:17-56), float and double.
//
// One source, two builds (both classes live in namespace planning, so they cannot share a binary):
//   oracle/_ref/velped_ref    reference headers + the reference's own objects   (oracle/Makefile)
//   tests/cpp/bin/velped_b200 this repo's headers + libpath_planning_b200.so    (build.py)
// Both print every result as raw IEEE bits; tests/test_cpu_velped.py demands identical output and compares both with
// the committed tests/golden/velped_ref.txt.  Inputs come from std::mt19937 raw bits (no distribution classes, which
// are implementation-defined).  Neither class touches the GPU.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "VelocityGenerator.h"
#include "PedestrianHandler.h"

using namespace planning;

static std::mt19937 rng;
static double uni(double lo, double hi) { return lo + (hi - lo) * ((rng() >> 8) * (1.0 / 16777216.0)); }

template <typename T> static void put(T v);
template <> void put<float>(float v) { uint32_t b; std::memcpy(&b, &v, 4); std::printf(" %08x", b); }
template <> void put<double>(double v) { uint64_t b; std::memcpy(&b, &v, 8); std::printf(" %016llx", (unsigned long long)b); }

// A path in the layout find_path returns (goal -> start order, HybridAStar.cpp:208-262): arcs of piecewise-constant
// curvature walked start -> goal, then reversed; the curvature vector holds ABSOLUTE values like the caller's
// (HybridAStar.cpp:236, :256).  `kappa_scale` 0 gives a straight path (the `curvature != 0` branch); `signed_curv`
// keeps the sign, which drives v^2 negative and the profile to NaN - the bits of that must agree as well.
template <typename T> static void make_path(int n, double ds, double kappa_scale, bool signed_curv, std::vector<Vector3D<T>>& path,
                                            std::vector<T>& curv)
{
    std::vector<Vector3D<T>> fwd; std::vector<T> cf;
    double x = uni(-5, 5), y = uni(-5, 5), h = uni(-3, 3), k = 0;
    for (int i = 0; i < n; i++)
    {
        if (i % 7 == 0) k = kappa_scale * uni(-0.24, 0.24);
        if (i % 11 == 5) k = 0;
        fwd.emplace_back(T(x), T(y), T(h)); cf.push_back(T(signed_curv ? k : std::fabs(k)));
        x += ds * std::cos(h); y += ds * std::sin(h); h += ds * k;
    }
    path.assign(fwd.rbegin(), fwd.rend()); curv.assign(cf.rbegin(), cf.rend());
}

template <typename T> static void run_velocity(const char* tag)
{
    rng.seed(1234);
    int id = 0;
    for (int rep = 0; rep < 64; rep++)
    {
        // limits around the launch defaults (launch/local_planner.launch: 5 / 1.5 / 2.0 / 1.5 / 2.5)
        VelocityGenerator<T> vg(T(uni(3, 8)), T(uni(1, 2)), T(uni(1.5, 3.5)), T(uni(1, 2)), T(uni(2, 3)));
        int n = (rep < 4) ? 1 + rep : 2 + (int)(rng() % 120);            // includes the 1-point path of callback_odom (:216-222)
        std::vector<Vector3D<T>> path; std::vector<T> curv;
        make_path<T>(n, uni(0.2, 0.6), (rep % 5 == 0) ? 0.0 : 1.0, rep % 16 == 7, path, curv);
        for (int flags = 0; flags < 4; flags++)
        {
            T v0 = T(uni(0, 2.4)), vcap = (rep % 3 == 0) ? std::numeric_limits<T>::max() : T(uni(0.5, 6));
            std::vector<T> vel;
            bool ok = vg.generate_velocity_profile(v0, vcap, path, curv, vel, (flags & 1) != 0, (flags & 2) != 0);
            std::printf("%s vel %d n %zu ok %d :", tag, id++, vel.size(), ok ? 1 : 0);
            for (T v : vel) put<T>(v);
            std::printf("\n");
        }
    }
}

template <typename T> static void run_pedestrians(const char* tag)
{
    rng.seed(4321);
    for (int rep = 0; rep < 400; rep++)
    {
        // launch defaults: arc 40 deg, stop 4 m, ttc 4 s, dec 1.5, min vel 0.5 (launch/local_planner.launch:40-45), varied
        PedestrianHandler<T> ph(T(uni(20, 120) * M_PI / 180.0), T(uni(1, 5)), T(uni(2, 6)), T(uni(0.8, 2.5)), T(uni(0.3, 1.0)));
        Vector3D<T> pose(T(uni(-20, 20)), T(uni(-20, 20)), T(uni(-3.14, 3.14)));
        std::vector<Obstacle<T>> peds;
        int np = rep % 6;                                                  // 0 pedestrians -> no cap
        for (int k = 0; k < np; k++)
        {
            double r = uni(0.3, 12), a = pose._heading + uni(-0.6, 0.6);
            peds.emplace_back(T(pose._x + r * std::cos(a)), T(pose._y + r * std::sin(a)), T(0.5), T(0.5));
        }
        T v = T(uni(2, 9));
        std::printf("%s ped %d n %d :", tag, rep, np);
        put<T>(ph.calc_max_velocity(v, pose, peds));
        std::printf("\n");
    }
}

int main()
{
    run_velocity<float>("f32"); run_velocity<double>("f64");
    run_pedestrians<float>("f32"); run_pedestrians<double>("f64");
    return 0;
}
