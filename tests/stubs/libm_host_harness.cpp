// Host harness for 3dvision_b200/csrc/b3d_libm.cuh (test infrastructure, CPU only): the header's sin_libm / cos_libm
// against the installed libm's sinf / cosf over float bit patterns first, first + stride, ... < last.
#include "../../3dvision_b200/csrc/b3d_libm.cuh"
#include <math.h>

static bool same(float a, float b) {
    uint32_t x, y; memcpy(&x, &a, 4); memcpy(&y, &b, 4);
    return x == y || (a != a && b != b);
}
// returns the number of mismatches; *first_bad = the first mismatching bit pattern (if any)
extern "C" long libm_compare(unsigned long long first, unsigned long long last, unsigned long long stride, int which, unsigned* first_bad) {
    long bad = 0;
    for (unsigned long long b = first; b < last; b += stride) {
        const uint32_t u = (uint32_t)b;
        float x; memcpy(&x, &u, 4);
        const float got = which ? b3d::libm::cos_libm(x) : b3d::libm::sin_libm(x);
        const float want = which ? cosf(x) : sinf(x);
        if (!same(got, want)) { if (!bad && first_bad) *first_bad = u; ++bad; }
    }
    return bad;
}
extern "C" float libm_sin(float x) { return b3d::libm::sin_libm(x); }
extern "C" float libm_cos(float x) { return b3d::libm::cos_libm(x); }
