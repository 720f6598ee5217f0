// b3d_featmath.cuh — scalar numerics of the feature stages that must reproduce the CPU path bit for bit.
//
//  * sym_eig3(): Eigen 3.4 SelfAdjointEigenSolver<Matrix3f>::compute() — the *iterative* solver the
//    reference calls at src/registration.cpp:122 (not computeDirect): scale by the largest |coefficient|
//    of the lower triangle, closed-form Householder tridiagonalisation of the 3x3 case, implicit
//    symmetric QR steps with a Wilkinson shift (Givens rotations accumulated into Q), ascending
//    selection sort of the eigenvalues with their vectors.
//  * atan2_libm(): glibc 2.39's atan2f / atanf (the Sun fdlibm float algorithm; verified here against
//    libm on 2e8 inputs, 0 mismatches).  src/registration.cpp:175 bins theta = atan2(w.n, u.n); CUDA's
//    atan2f is 2-ulp accurate but not the same function, and a last-bit difference at a bin edge moves a
//    histogram count.
// Compile with --fmad=false (every a*b+c below is two roundings, as in the CPU build).
#pragma once
#include "b3d_linalg.cuh"

namespace b3d {

// Eigen::JacobiRotation<float>::makeGivens(p, q) (real case); Givens is b3d_linalg.cuh's
B3D_HD Givens make_givens(float p, float q) {
    Givens g;
    if (q == 0.0f) { g.c = p < 0.0f ? -1.0f : 1.0f; g.s = 0.0f; }
    else if (p == 0.0f) { g.c = 0.0f; g.s = q < 0.0f ? 1.0f : -1.0f; }
    else if (fabsf(p) > fabsf(q)) {
        const float t = q / p;
        float u = sqrtf(1.0f + t * t);
        if (p < 0.0f) u = -u;
        g.c = 1.0f / u; g.s = -t * g.c;
    } else {
        const float t = p / q;
        float u = sqrtf(1.0f + t * t);
        if (q < 0.0f) u = -u;
        g.s = -1.0f / u; g.c = -t * g.s;
    }
    return g;
}

// Eigen::numext::hypot (positive_real_hypot)
B3D_HD float eigen_hypot(float x, float y) {
    x = fabsf(x); y = fabsf(y);
    const float p = fmaxf(x, y);
    if (p == 0.0f) return 0.0f;
    const float qp = fminf(y, x) / p;
    return p * sqrtf(1.0f + qp * qp);
}

// lower triangle in: a00 a10 a11 a20 a21 a22.  Out: eigenvalues ascending, Q columns = eigenvectors (Q(r, c)).
B3D_HD void sym_eig3(float a00, float a10, float a11, float a20, float a21, float a22, float evals[3], Mat3& Q) {
    float scale = fmaxf(fmaxf(fmaxf(fabsf(a00), fabsf(a10)), fmaxf(fabsf(a11), fabsf(a20))), fmaxf(fabsf(a21), fabsf(a22)));
    if (scale == 0.0f) scale = 1.0f;
    a00 /= scale; a10 /= scale; a11 /= scale; a20 /= scale; a21 /= scale; a22 /= scale;
    const float kMin = 1.17549435e-38f;                  // numeric_limits<float>::min()
    float d0 = a00, d1, d2, e0, e1;
    const float v1norm2 = a20 * a20;
    if (v1norm2 <= kMin) {
        d1 = a11; d2 = a22; e0 = a10; e1 = a21;
        mat3_identity(Q);
    } else {
        const float beta = sqrtf(a10 * a10 + v1norm2);
        const float inv_beta = 1.0f / beta;
        const float m01 = a10 * inv_beta, m02 = a20 * inv_beta;
        const float q = 2.0f * m01 * a21 + m02 * (a22 - a11);
        d1 = a11 + m02 * q; d2 = a22 - m02 * q;
        e0 = beta; e1 = a21 - m01 * q;
        Q(0, 0) = 1.0f; Q(0, 1) = 0.0f; Q(0, 2) = 0.0f;
        Q(1, 0) = 0.0f; Q(1, 1) = m01;  Q(1, 2) = m02;
        Q(2, 0) = 0.0f; Q(2, 1) = m02;  Q(2, 2) = -m01;
    }
    float diag[3] = {d0, d1, d2}, sub[2] = {e0, e1};
    const float precision_inv = 1.0f / 1.1920929e-07f;   // 1 / epsilon
    int end = 2, start = 0, iter = 0;
    while (end > 0) {
        for (int i = start; i < end; ++i) {
            if (fabsf(sub[i]) < kMin) sub[i] = 0.0f;
            else {
                const float ss = precision_inv * sub[i];
                if (ss * ss <= (fabsf(diag[i]) + fabsf(diag[i + 1]))) sub[i] = 0.0f;
            }
        }
        while (end > 0 && sub[end - 1] == 0.0f) --end;
        if (end <= 0) break;
        if (++iter > 90) break;                          // m_maxIterations * n
        start = end - 1;
        while (start > 0 && sub[start - 1] != 0.0f) --start;
        // one implicit QR step on the unreduced block [start, end]
        const float td = (diag[end - 1] - diag[end]) * 0.5f;
        const float e = sub[end - 1];
        float mu = diag[end];
        if (td == 0.0f) mu -= fabsf(e);
        else if (e != 0.0f) {
            const float e2 = e * e;
            const float h = eigen_hypot(td, e);
            if (e2 == 0.0f) mu -= e / ((td + (td > 0.0f ? h : -h)) / e);
            else            mu -= e2 / (td + (td > 0.0f ? h : -h));
        }
        float x = diag[start] - mu, z = sub[start];
        for (int k = start; k < end && z != 0.0f; ++k) {
            const Givens r = make_givens(x, z);
            const float sdk = r.s * diag[k] + r.c * sub[k];
            const float dkp1 = r.s * sub[k] + r.c * diag[k + 1];
            diag[k] = r.c * (r.c * diag[k] - r.s * sub[k]) - r.s * (r.c * sub[k] - r.s * diag[k + 1]);
            diag[k + 1] = r.s * sdk + r.c * dkp1;
            sub[k] = r.c * sdk - r.s * dkp1;
            if (k > start) sub[k - 1] = r.c * sub[k - 1] - r.s * z;
            x = sub[k];
            if (k < end - 1) { z = -r.s * sub[k + 1]; sub[k + 1] = r.c * sub[k + 1]; }
            if (!(r.c == 1.0f && r.s == 0.0f)) {        // Q.applyOnTheRight(k, k+1, rot)
#pragma unroll
                for (int row = 0; row < 3; ++row) {
                    const float xi = Q(row, k), yi = Q(row, k + 1);
                    Q(row, k) = r.c * xi + (-r.s) * yi;
                    Q(row, k + 1) = r.s * xi + r.c * yi;
                }
            }
        }
    }
    for (int i = 0; i < 2; ++i) {                        // ascending selection sort, vectors follow
        int k = 0; float mn = diag[i];
        for (int j = i + 1; j < 3; ++j) if (diag[j] < mn) { mn = diag[j]; k = j - i; }
        if (k > 0) {
            const float tmp = diag[i]; diag[i] = diag[k + i]; diag[k + i] = tmp;
#pragma unroll
            for (int row = 0; row < 3; ++row) { const float t2 = Q(row, i); Q(row, i) = Q(row, k + i); Q(row, k + i) = t2; }
        }
    }
    evals[0] = diag[0] * scale; evals[1] = diag[1] * scale; evals[2] = diag[2] * scale;
}

// ---- glibc atanf / atan2f (fdlibm float), finite inputs and the zero / infinity cases ----
__device__ __forceinline__ float atan_libm(float x) {
    const float hi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float lo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const int hx = __float_as_int(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {                              // |x| >= 2^25
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? hi[3] + lo[3] : -hi[3] - lo[3];
    }
    if (ix < 0x3ee00000) {                               // |x| < 0.4375
        if (ix < 0x31000000) return x;                   // |x| < 2^-29
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {                           // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else                 { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else                 { id = 3; x = -1.0f / x; }
        }
    }
    const float z = x * x, w = z * z;
    const float s1 = z * (3.3333334327e-01f + w * (1.4285714924e-01f + w * (9.0908870101e-02f + w * (6.6610731184e-02f +
                     w * (4.9768779427e-02f + w * 1.6285819933e-02f)))));
    const float s2 = w * (-2.0000000298e-01f + w * (-1.1111110449e-01f + w * (-7.6918758452e-02f + w * (-5.8335702866e-02f +
                     w * -3.6531571299e-02f))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = hi[id] - ((x * (s1 + s2) - lo[id]) - x);
    return hx < 0 ? -r : r;
}

__device__ __forceinline__ float atan2_libm(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
                pi_lo = -8.7422776573e-08f;
    const int hx = __float_as_int(x), hy = __float_as_int(y), ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atan_libm(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        if (m < 2) return y;
        return m == 2 ? pi + tiny : -pi - tiny;
    }
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) { case 0: return pi_o_4 + tiny; case 1: return -pi_o_4 - tiny; case 2: return 3.0f * pi_o_4 + tiny; default: return -3.0f * pi_o_4 - tiny; }
        }
        switch (m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return pi + tiny; default: return -pi - tiny; }
    }
    if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = atan_libm(fabsf(y / x));
    switch (m) {
        case 0: return z;
        case 1: return __int_as_float(__float_as_int(z) ^ (int)0x80000000);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

}  // namespace b3d
