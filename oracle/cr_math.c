/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * "Pinned libm" for the libref_oracle_crm.so flavour of the reference build (oracle/Makefile).
 * The reference calls glibc's float transcendentals, whose results are NOT correctly rounded and
 * differ between libm versions (measured here, glibc 2.39: sinf 1.4 %, cosf 1.2 %, acosf 5.6 %,
 * atan2f 16 % of random arguments differ from the correctly rounded float).  SURVEY.md H2 / §8(c)
 * call this "libm version unpinned".  For the functions that the B200 build executes ON THE DEVICE
 * (Dubins.cpp, and atan2f inside Grid3D::get_field_intensity) the crm flavour redirects the float
 * calls -- by `objcopy --redefine-sym` on the compiled, unmodified objects -- to the versions
 * below: evaluate in double, round once to float.  The device code does the same with CUDA's
 * double-precision functions, so both sides agree bit-for-bit except when the double result lies
 * within ~2 double-ulp of a float rounding boundary (~1e-8 of calls).
 */
#include <math.h>

float pp_cr_sinf(float x) { return (float)sin((double)x); }
float pp_cr_cosf(float x) { return (float)cos((double)x); }
void  pp_cr_sincosf(float x, float* s, float* c) { *s = (float)sin((double)x); *c = (float)cos((double)x); }
float pp_cr_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
float pp_cr_acosf(float x) { return (float)acos((double)x); }
