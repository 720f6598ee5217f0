// b3d_libm.cuh — glibc's sinf / cosf, restated (host + device).
//
// The reference turns the ICP update into a rotation through Eigen's AngleAxisf -> Quaternionf conversion
// (src/registration.cpp:369-371; Eigen/src/Geometry/Quaternion.h: w = cos(angle / 2), vec = sin(angle / 2) * axis with
// std::sin / std::cos on float, i.e. libm's sinf / cosf).  CUDA's sinf / cosf are accurate to about 1 ulp but not the
// same function: they differ from glibc's in a few per cent of arguments, which a randomised soak of b3d_icp against the
// oracle showed as 1-ulp pose differences whenever an ill-conditioned solve returned angles that are not tiny.  For
// bit-parity the device therefore evaluates glibc's own algorithm:
//
//   glibc >= 2.28, sysdeps/ieee754/flt-32/{s_sinf.c, s_cosf.c, s_sincosf.h, s_sincosf_data.c} (from ARM's optimized
//   routines): the argument is widened to double; |x| < pi/4 goes straight to an odd / even polynomial in double;
//   |x| < 120 is reduced with n = round(x * 2/pi) (a 2^24-scaled double -> int32 conversion) and x - n * pi/2;
//   larger arguments are reduced against a 192-bit table of 4/pi; the polynomial result is rounded to float once.
//
// The constants below were read out of the __sincosf_table / __inv_pio4 objects of this image's libm.so.6 (glibc
// 2.39); tests/test_libm_host.py compiles this header with g++ and compares it with the installed sinf / cosf over the
// whole float range (strided in the test suite; exhaustive once, recorded in DESIGN.md).  x86-64 libm selects an
// FMA-compiled variant of the same source at run time (ifunc) on every CPU that has FMA — which is what the reference and
// the oracle run — and one contraction in it is visible: x - n * pi/2 of the |x| < 120 reduction.  With that one fused
// (explicit fma() below) this header equals the installed sinf AND cosf on all 2^32 arguments; un-fused, 12 + 22
// arguments with 17 < |x| < 120 differ by one ulp and none below pi/4, where every half-angle of a converging ICP lies.
//
// Translation units including this header must be compiled with --fmad=false (as all of libb3d is).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define B3D_LIBM_HD __host__ __device__ __forceinline__
#else
#define B3D_LIBM_HD inline
#endif

namespace b3d {
namespace libm {

struct SinCosTable {
    double sign[4];
    double hpi_inv, hpi;
    double c0, c1, c2, c3, c4;
    double s1, s2, s3;
};

B3D_LIBM_HD uint32_t as_u32(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
B3D_LIBM_HD uint32_t abstop12(float x) { return (as_u32(x) >> 20) & 0x7ffu; }

// which = 0: the table for even quadrants' sign pattern, 1: the negated cosine polynomial (quadrants 2, 3)
B3D_LIBM_HD SinCosTable table(int which) {
    SinCosTable t;
    t.sign[0] = 1.0; t.sign[1] = -1.0; t.sign[2] = -1.0; t.sign[3] = 1.0;
    t.hpi_inv = 0x1.45F306DC9C883p+23;                          // 2/pi * 2^24
    t.hpi = 0x1.921FB54442D18p0;                                // pi/2
    const double sg = which ? -1.0 : 1.0;
    t.c0 = sg * 0x1p0;
    t.c1 = sg * -0x1.ffffffd0c621cp-2;
    t.c2 = sg * 0x1.55553e1068f19p-5;
    t.c3 = sg * -0x1.6c087e89a359dp-10;
    t.c4 = sg * 0x1.99343027bf8c3p-16;
    t.s1 = -0x1.555545995a603p-3;
    t.s2 = 0x1.1107605230bc4p-7;
    t.s3 = -0x1.994eb3774cf24p-13;
    return t;
}

// sin (n even) or cos (n odd) of the reduced argument, evaluated in double and rounded to float once
B3D_LIBM_HD float sinf_poly(double x, double x2, const SinCosTable& p, int n) {
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double s1 = p.s2 + x2 * p.s3;
        const double x7 = x3 * x2;
        const double s = x + x3 * p.s1;
        return (float)(s + x7 * s1);
    }
    const double x4 = x2 * x2;
    const double c2 = p.c3 + x2 * p.c4;
    const double c1 = p.c0 + x2 * p.c1;
    const double x6 = x4 * x2;
    const double c = c1 + x4 * p.c2;
    return (float)(c + x6 * c2);
}

B3D_LIBM_HD double reduce_fast(double x, const SinCosTable& p, int& n_out) {
    const double r = x * p.hpi_inv;
    const int n = ((int32_t)r + 0x800000) >> 24;                // quadrant in bits 24..31 of the scaled value, rounded
    n_out = n;
    return fma(-(double)n, p.hpi, x);                           // x - n * hpi, fused: see the header comment
}

B3D_LIBM_HD uint32_t inv_pio4(unsigned i) {                      // 4/pi, 32 bits per step of 8 bits
    const uint32_t t[24] = {0xa2u, 0xa2f9u, 0xa2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u,
                            0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u,
                            0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u};
    return t[i];
}
B3D_LIBM_HD double reduce_large(uint32_t xi, int& n_out) {
    const unsigned a = (xi >> 26) & 15u;
    const int shift = (int)((xi >> 23) & 7u);
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    uint64_t res0 = (uint32_t)(xi * inv_pio4(a));                // 32-bit product, as in the source
    const uint64_t res1 = (uint64_t)xi * inv_pio4(a + 4u);
    const uint64_t res2 = (uint64_t)xi * inv_pio4(a + 8u);
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    const uint64_t n = (res0 + (1ull << 61)) >> 62;
    res0 -= n << 62;
    const double x = (double)(int64_t)res0;
    n_out = (int)n;
    return x * 0x1.921FB54442D18p-62;                           // pi / 2^63... scaled: pi63
}

B3D_LIBM_HD float sin_libm(float y) {
    double x = (double)y;
    int n;
    if (abstop12(y) < 0x3f4u) {                                  // |y| < pi/4 (top 12 bits of 0x1.921FB6p-1f)
        const double s = x * x;
        if (abstop12(y) < 0x398u) return y;                      // |y| < 2^-12
        return sinf_poly(x, s, table(0), 0);
    }
    if (abstop12(y) < 0x42fu) {                                  // |y| < 120
        const SinCosTable p0 = table(0);
        x = reduce_fast(x, p0, n);
        const double s = p0.sign[n & 3];
        return sinf_poly(x * s, x * x, table((n & 2) ? 1 : 0), n);
    }
    if (abstop12(y) < 0x7f8u) {                                  // finite
        const uint32_t xi = as_u32(y);
        const int sign = (int)(xi >> 31);
        x = reduce_large(xi, n);
        const SinCosTable p0 = table(0);
        const double s = p0.sign[(n + sign) & 3];
        return sinf_poly(x * s, x * x, table(((n + sign) & 2) ? 1 : 0), n);
    }
    return (y - y) / (y - y);                                    // inf / nan -> nan
}

B3D_LIBM_HD float cos_libm(float y) {
    double x = (double)y;
    int n;
    if (abstop12(y) < 0x3f4u) {
        const double x2 = x * x;
        if (abstop12(y) < 0x398u) return 1.0f;
        return sinf_poly(x, x2, table(0), 1);
    }
    if (abstop12(y) < 0x42fu) {
        const SinCosTable p0 = table(0);
        x = reduce_fast(x, p0, n);
        const double s = p0.sign[n & 3];
        return sinf_poly(x * s, x * x, table((n & 2) ? 1 : 0), n ^ 1);
    }
    if (abstop12(y) < 0x7f8u) {
        const uint32_t xi = as_u32(y);
        const int sign = (int)(xi >> 31);
        x = reduce_large(xi, n);
        const SinCosTable p0 = table(0);
        const double s = p0.sign[(n + sign) & 3];
        return sinf_poly(x * s, x * x, table(((n + sign) & 2) ? 1 : 0), n ^ 1);
    }
    return (y - y) / (y - y);
}

}  // namespace libm
}  // namespace b3d
