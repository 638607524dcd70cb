// TEST ONLY.  Pins path_planning_pkg_b200/csrc/core/pp_gmath.h (the PP_HD restatement of glibc 2.39's binary32
// sinf / cosf / atanf / atan2f / acosf) against the libm this process is linked with -- the one the stock reference
// build (oracle/_ref/libref_oracle.so) calls.
//
//   gmath_check exhaustive [stride]   every stride-th of the 2^32 float bit patterns through sinf, cosf, atanf, acosf
//                                     (+ sincosf == (sinf, cosf), which GCC substitutes in the reference's Dubins.o)
//   gmath_check atan2 <n> [seed]      n random (y, x) pairs: uniform bit patterns, and pairs drawn the way the planner
//                                     produces them (metre-scale offsets), plus the special-value grid
// Prints one line per function: "<name> checked <count> mismatches <count>"; exit code 1 on any mismatch.
// NaN results compare equal to NaN results (payload / sign of a NaN is not part of the claim).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../../path_planning_pkg_b200/csrc/core/pp_gmath.h"

static inline bool same(float a, float b)
{
    if (a != a && b != b) return true;
    return pp_g_f2u(a) == pp_g_f2u(b);
}

struct Job { uint64_t lo, hi, stride; uint64_t bad[6]; uint64_t n; uint32_t first_bad[6]; };

static void* run_exhaustive(void* p)
{
    Job* j = (Job*)p;
    for (uint64_t u = j->lo; u < j->hi; u += j->stride)
    {
        const float x = pp_g_u2f((uint32_t)u);
        volatile float xv = x;     // keep the compiler from folding the libm calls
        float rs = sinf(xv), rc = cosf(xv), ra = atanf(xv), rk = acosf(xv), ss, cc;
        sincosf(xv, &ss, &cc);
        if (!same(rs, pp_g_sinf(x))) { if (!j->bad[0]++) j->first_bad[0] = (uint32_t)u; }
        if (!same(rc, pp_g_cosf(x))) { if (!j->bad[1]++) j->first_bad[1] = (uint32_t)u; }
        if (!same(ra, pp_g_atanf(x))) { if (!j->bad[2]++) j->first_bad[2] = (uint32_t)u; }
        if (!same(rk, pp_g_acosf(x))) { if (!j->bad[3]++) j->first_bad[3] = (uint32_t)u; }
        if (!same(ss, rs) || !same(cc, rc)) { if (!j->bad[4]++) j->first_bad[4] = (uint32_t)u; }
        float ps, pc;
        pp_g_sincosf(x, &ps, &pc);             // the shared-reduction form the search kernel calls
        if (!same(ps, rs) || !same(pc, rc)) { if (!j->bad[5]++) j->first_bad[5] = (uint32_t)u; }
        j->n++;
    }
    return 0;
}

static uint64_t rng_state;
static inline uint64_t rng() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: gmath_check exhaustive [stride] | atan2 <n> [seed]\n"); return 2; }
    int rc = 0;
    if (!strcmp(argv[1], "exhaustive"))
    {
        const uint64_t stride = argc > 2 ? strtoull(argv[2], 0, 10) : 1;
        const int T = 16;
        pthread_t th[T]; Job jobs[T];
        const uint64_t span = (1ULL << 32) / T;
        for (int t = 0; t < T; t++)
        {
            memset(&jobs[t], 0, sizeof(Job));
            jobs[t].lo = span * t; jobs[t].hi = span * (t + 1); jobs[t].stride = stride;
            pthread_create(&th[t], 0, run_exhaustive, &jobs[t]);
        }
        uint64_t bad[6] = {0, 0, 0, 0, 0, 0}, n = 0; uint32_t fb[6] = {0, 0, 0, 0, 0, 0};
        for (int t = 0; t < T; t++)
        {
            pthread_join(th[t], 0);
            for (int k = 0; k < 6; k++) { if (jobs[t].bad[k] && !bad[k]) fb[k] = jobs[t].first_bad[k]; bad[k] += jobs[t].bad[k]; }
            n += jobs[t].n;
        }
        const char* names[6] = {"sinf", "cosf", "atanf", "acosf", "sincosf_vs_sinf_cosf", "pp_g_sincosf"};
        for (int k = 0; k < 6; k++)
        {
            printf("%s checked %llu mismatches %llu", names[k], (unsigned long long)n, (unsigned long long)bad[k]);
            if (bad[k]) { printf(" first 0x%08x", fb[k]); rc = 1; }
            printf("\n");
        }
        return rc;
    }
    if (!strcmp(argv[1], "atan2"))
    {
        const uint64_t n = argc > 2 ? strtoull(argv[2], 0, 10) : 1000000;
        rng_state = argc > 3 ? strtoull(argv[3], 0, 10) * 2654435761ULL + 88172645463325252ULL : 88172645463325252ULL;
        uint64_t bad = 0, cnt = 0; float by = 0, bx = 0;
        // special-value grid
        const float sp[] = {0.0f, -0.0f, 1.0f, -1.0f, INFINITY, -INFINITY, NAN, 1e-45f, -1e-45f, 3.4e38f, -3.4e38f, 1.1754944e-38f,
                            0.5f, 2.0f, 1e30f, 1e-30f, -1e30f, -1e-30f};
        const int ns = (int)(sizeof(sp) / sizeof(sp[0]));
        for (int a = 0; a < ns; a++)
            for (int b = 0; b < ns; b++)
            {
                volatile float y = sp[a], x = sp[b];
                if (!same(atan2f(y, x), pp_g_atan2f(sp[a], sp[b]))) { if (!bad++) { by = sp[a]; bx = sp[b]; } }
                cnt++;
            }
        for (uint64_t k = 0; k < n; k++)
        {
            float y, x;
            const uint64_t r = rng();
            if (k & 1) { y = pp_g_u2f((uint32_t)r); x = pp_g_u2f((uint32_t)(r >> 32)); }
            else
            {
                // planner-like: offsets of a few hundred metres with 24-bit mantissas
                y = ((float)(int32_t)(uint32_t)r) * (1.0f / 8388608.0f);
                x = ((float)(int32_t)(uint32_t)(r >> 32)) * (1.0f / 8388608.0f);
                if ((k & 6) == 2) x *= 1.0f / 1024.0f;
                if ((k & 6) == 4) y *= 1.0f / 1024.0f;
            }
            volatile float yv = y, xv = x;
            if (!same(atan2f(yv, xv), pp_g_atan2f(y, x))) { if (!bad++) { by = y; bx = x; } }
            cnt++;
        }
        printf("atan2f checked %llu mismatches %llu", (unsigned long long)cnt, (unsigned long long)bad);
        if (bad) { printf(" first y=%a x=%a", by, bx); rc = 1; }
        printf("\n");
        return rc;
    }
    return 2;
}
