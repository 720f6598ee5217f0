// b3d_linalg.cuh — small dense kernels used on-device by the registration path.
//
// Everything here reproduces, operation for operation, the fp32 arithmetic the
// reference obtains from Eigen 3.4 at these call sites (reference repo paths):
//   JacobiSVD<Matrix3f>(H, FullU|FullV)            src/registration.cpp:255, :388
//   V * U^T, determinant() sign fix                src/registration.cpp:256-262, :389-394
//   rowwise().mean(), centring, H = Sc * Tc^T      src/registration.cpp:248-254
//   Matrix<float,6,6>::ldlt().solve(-ATb)          src/registration.cpp:366
//   AngleAxis(x)*AngleAxis(y)*AngleAxis(z)         src/registration.cpp:369-371
// The translation unit including this header MUST be compiled with --fmad=false:
// the reference is built without FMA contraction, and a single fused multiply-add
// can flip an inlier at the RANSAC threshold.  3-element reductions use Eigen's
// unrolled order  a0 + (a1 + a2).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>

#include "b3d_libm.cuh"

#define B3D_HD __host__ __device__ __forceinline__

namespace b3d {

B3D_HD float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }

// 3x3 matrices are stored column-major in a flat array: m[c*3 + r]  (Eigen::Matrix3f storage).
struct Mat3 {
    float m[9];
    B3D_HD float& operator()(int r, int c) { return m[c * 3 + r]; }
    B3D_HD float operator()(int r, int c) const { return m[c * 3 + r]; }
};

B3D_HD void mat3_identity(Mat3& A) {
#pragma unroll
    for (int i = 0; i < 9; ++i) A.m[i] = (i % 4 == 0) ? 1.0f : 0.0f;
}

// out = A * B^T, coefficient-wise lazy product: out(i,j) = A(i,0)B(j,0) + (A(i,1)B(j,1) + A(i,2)B(j,2))
B3D_HD void mat3_mul_bt(const Mat3& A, const Mat3& B, Mat3& out) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i)
            out(i, j) = sum3(A(i, 0) * B(j, 0), A(i, 1) * B(j, 1), A(i, 2) * B(j, 2));
}

B3D_HD float mat3_det(const Mat3& A) {
    float c0 = A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1));
    float c1 = A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0));
    float c2 = A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
    return c0 - c1 + c2;
}

// y = A x (+ nothing): y_r = A(r,0)x0 + (A(r,1)x1 + A(r,2)x2)
B3D_HD void mat3_vec(const Mat3& A, float x0, float x1, float x2, float& y0, float& y1, float& y2) {
    y0 = sum3(A(0, 0) * x0, A(0, 1) * x1, A(0, 2) * x2);
    y1 = sum3(A(1, 0) * x0, A(1, 1) * x1, A(1, 2) * x2);
    y2 = sum3(A(2, 0) * x0, A(2, 1) * x1, A(2, 2) * x2);
}

// ---- plane rotations (Eigen::JacobiRotation) -----------------------------------------
struct Givens { float c, s; };

// In-plane rotation of two 3-element strided vectors: x' = c x + s y ; y' = -s x + c y.
B3D_HD void rotate_pair(float* x, float* y, int stride, Givens g) {
    if (g.c == 1.0f && g.s == 0.0f) return;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float xi = x[i * stride], yi = y[i * stride];
        x[i * stride] = g.c * xi + g.s * yi;
        y[i * stride] = -g.s * xi + g.c * yi;
    }
}
// rows p,q of a column-major Mat3 have stride 3; columns have stride 1.
B3D_HD void rot_rows(Mat3& A, int p, int q, Givens g) { rotate_pair(&A.m[p], &A.m[q], 3, g); }
B3D_HD void rot_cols(Mat3& A, int p, int q, Givens g) { rotate_pair(&A.m[3 * p], &A.m[3 * q], 1, Givens{g.c, -g.s}); }

B3D_HD Givens jacobi_from_sym2(float x, float y, float z) {   // makeJacobi(x, y, z)
    Givens g;
    float deno = 2.0f * fabsf(y);
    if (deno < FLT_MIN) { g.c = 1.0f; g.s = 0.0f; return g; }
    float tau = (x - z) / deno;
    float w = sqrtf(tau * tau + 1.0f);
    float t = (tau > 0.0f) ? (1.0f / (tau + w)) : (1.0f / (tau - w));
    float sgn = (t > 0.0f) ? 1.0f : -1.0f;
    float n = 1.0f / sqrtf(t * t + 1.0f);
    g.s = -sgn * (y / fabsf(y)) * fabsf(t) * n;
    g.c = n;
    return g;
}

// 2x2 real SVD step on the (p,q) sub-block of W (real_2x2_jacobi_svd).
B3D_HD void svd2x2(const Mat3& W, int p, int q, Givens& left, Givens& right) {
    float a = W(p, p), b = W(p, q), c = W(q, p), d = W(q, q);
    Givens r1;
    float tr = a + d, df = c - b;
    if (fabsf(df) < FLT_MIN) { r1.c = 1.0f; r1.s = 0.0f; }
    else {
        float u = tr / df;
        float h = sqrtf(1.0f + u * u);
        r1.s = 1.0f / h;
        r1.c = u / h;
    }
    if (!(r1.c == 1.0f && r1.s == 0.0f)) {
        float na = r1.c * a + r1.s * c, nb = r1.c * b + r1.s * d;
        float nc = -r1.s * a + r1.c * c, nd = -r1.s * b + r1.c * d;
        a = na; b = nb; c = nc; d = nd;
    }
    right = jacobi_from_sym2(a, b, d);
    float tc = right.c, ts = -right.s;                // right.transpose()
    left.c = r1.c * tc - r1.s * ts;
    left.s = r1.c * ts + r1.s * tc;
}

// Two-sided Jacobi SVD of a 3x3 (JacobiSVD, square => no preconditioner). Singular
// values sorted descending; U, V full.
B3D_HD void svd3(const Mat3& A, Mat3& U, Mat3& V, float sv[3]) {
    float scale = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(A.m[i]));
    mat3_identity(U); mat3_identity(V);
    if (!isfinite(scale)) { sv[0] = sv[1] = sv[2] = 0.0f; return; }
    if (scale == 0.0f) scale = 1.0f;
    Mat3 W;
#pragma unroll
    for (int i = 0; i < 9; ++i) W.m[i] = A.m[i] / scale;
    float maxd = fmaxf(fabsf(W(0, 0)), fmaxf(fabsf(W(1, 1)), fabsf(W(2, 2))));
    const float prec = 2.0f * FLT_EPSILON;
    for (int sweep = 0; sweep < 64; ++sweep) {          // Eigen loops until a sweep is clean
        bool clean = true;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                float thr = fmaxf(FLT_MIN, prec * maxd);
                if (fabsf(W(p, q)) > thr || fabsf(W(q, p)) > thr) {
                    clean = false;
                    Givens gl, gr;
                    svd2x2(W, p, q, gl, gr);
                    rot_rows(W, p, q, gl);
                    rot_cols(U, p, q, Givens{gl.c, -gl.s});
                    rot_cols(W, p, q, gr);
                    rot_cols(V, p, q, gr);
                    maxd = fmaxf(maxd, fmaxf(fabsf(W(p, p)), fabsf(W(q, q))));
                }
            }
        if (clean) break;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float d = W(i, i);
        sv[i] = fabsf(d);
        if (d < 0.0f) { U(0, i) = -U(0, i); U(1, i) = -U(1, i); U(2, i) = -U(2, i); }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) sv[i] *= scale;
    for (int i = 0; i < 3; ++i) {
        int at = i; float mx = sv[i];
        for (int k = i + 1; k < 3; ++k) if (sv[k] > mx) { mx = sv[k]; at = k; }
        if (mx == 0.0f) break;
        if (at != i) {
            float t = sv[i]; sv[i] = sv[at]; sv[at] = t;
            for (int r = 0; r < 3; ++r) {
                t = U(r, i); U(r, i) = U(r, at); U(r, at) = t;
                t = V(r, i); V(r, i) = V(r, at); V(r, at) = t;
            }
        }
    }
}

// R = V U^T, reflected through V.col(2) when det(R) < 0.
B3D_HD void rotation_from_cross_covariance(const Mat3& H, Mat3& R) {
    Mat3 U, V; float sv[3];
    svd3(H, U, V, sv);
    mat3_mul_bt(V, U, R);
    if (mat3_det(R) < 0.0f) {
        V(0, 2) *= -1.0f; V(1, 2) *= -1.0f; V(2, 2) *= -1.0f;
        mat3_mul_bt(V, U, R);
    }
}

// 3-point Kabsch (registration.cpp:242-264). s[k], q[k] are the k-th source / target point.
B3D_HD void kabsch_three_points(const float s[3][3], const float q[3][3], Mat3& R, float t[3]) {
    float cs[3], cq[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        cs[r] = sum3(s[0][r], s[1][r], s[2][r]) / 3.0f;
        cq[r] = sum3(q[0][r], q[1][r], q[2][r]) / 3.0f;
    }
    Mat3 Sc, Qc;                                   // column k = centred point k
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int r = 0; r < 3; ++r) { Sc(r, k) = s[k][r] - cs[r]; Qc(r, k) = q[k][r] - cq[r]; }
    Mat3 H;
    mat3_mul_bt(Sc, Qc, H);
    rotation_from_cross_covariance(H, R);
    float r0, r1, r2;
    mat3_vec(R, cs[0], cs[1], cs[2], r0, r1, r2);
    t[0] = cq[0] - r0; t[1] = cq[1] - r1; t[2] = cq[2] - r2;
}

// ---- 6x6 LDLT with diagonal pivoting + solve (Eigen::LDLT<Matrix<float,6,6>,Lower>) ----
B3D_HD float halving_sum(const float* c, int n) {
    if (n == 1) return c[0];
    if (n == 2) return c[0] + c[1];
    if (n == 3) return c[0] + (c[1] + c[2]);
    if (n == 4) return (c[0] + c[1]) + (c[2] + c[3]);
    return (c[0] + c[1]) + (c[2] + (c[3] + c[4]));
}

// A: symmetric 6x6, row-major a[r*6+c] (only the lower triangle is read). Solves A x = b.
// Every array index below is a compile-time constant once the loops are unrolled (the pivot row is matched against
// each candidate in a static loop instead of being used as an index), so on the device the whole factorisation lives
// in registers; with run-time indices it sat in local memory and the one-thread solve cost ~13k cycles per ICP iteration.
B3D_HD void ldlt6_solve(const float* A, const float* b, float* x) {
    float L[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) L[i] = A[i];
    int perm[6] = {0, 1, 2, 3, 4, 5};
    float tmp[6];
    bool stop = false;
#define LL(r, c) L[(r) * 6 + (c)]
#define B3D_SWAP(a, b) { const float t_ = (a); (a) = (b); (b) = t_; }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (stop) continue;
        int piv = k; float pmax = fabsf(LL(k, k));
#pragma unroll
        for (int i = k + 1; i < 6; ++i) { const float v = fabsf(LL(i, i)); if (v > pmax) { pmax = v; piv = i; } }
        perm[k] = piv;
#pragma unroll
        for (int p = k + 1; p < 6; ++p) {
            if (piv != p) continue;
#pragma unroll
            for (int j = 0; j < k; ++j) B3D_SWAP(LL(k, j), LL(p, j));
#pragma unroll
            for (int i = p + 1; i < 6; ++i) B3D_SWAP(LL(i, k), LL(i, p));
            B3D_SWAP(LL(k, k), LL(p, p));
#pragma unroll
            for (int i = k + 1; i < p; ++i) B3D_SWAP(LL(i, k), LL(p, i));
        }
        if (k > 0) {
#pragma unroll
            for (int j = 0; j < k; ++j) tmp[j] = LL(j, j) * LL(k, j);
            float acc = LL(k, 0) * tmp[0];
#pragma unroll
            for (int j = 1; j < k; ++j) acc = acc + LL(k, j) * tmp[j];
            LL(k, k) -= acc;
#pragma unroll
            for (int i = k + 1; i < 6; ++i) {
                float c = LL(i, 0) * tmp[0];
#pragma unroll
                for (int j = 1; j < k; ++j) c = c + LL(i, j) * tmp[j];
                LL(i, k) -= c;
            }
        }
        const float d = LL(k, k);
        const bool ok = fabsf(d) > 0.0f;
        if (k == 0 && !ok) {
#pragma unroll
            for (int j = 0; j < 6; ++j) perm[j] = j;
            stop = true;
            continue;
        }
        if (ok) {
#pragma unroll
            for (int i = k + 1; i < 6; ++i) LL(i, k) /= d;
        }
    }
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = b[i];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int p = k + 1; p < 6; ++p) if (perm[k] == p) B3D_SWAP(y[k], y[p]);
    }
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        float c[5];
#pragma unroll
        for (int j = 0; j < i; ++j) c[j] = LL(i, j) * y[j];
        y[i] -= halving_sum(c, i);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = (fabsf(LL(i, i)) > FLT_MIN) ? (y[i] / LL(i, i)) : 0.0f;
#pragma unroll
    for (int len = 1; len < 6; ++len) {
        const int row = 5 - len, first = row + 1;
        float c[5];
#pragma unroll
        for (int j = 0; j < len; ++j) c[j] = LL(first + j, row) * y[first + j];
        float s;
        if (len >= 4) { s = (c[0] + c[2]) + (c[1] + c[3]); if (len == 5) s = s + c[4]; }
        else s = halving_sum(c, len);
        y[row] -= s;
    }
#pragma unroll
    for (int k = 5; k >= 0; --k) {
#pragma unroll
        for (int p = k + 1; p < 6; ++p) if (perm[k] == p) B3D_SWAP(y[k], y[p]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = y[i];
#undef B3D_SWAP
#undef LL
}

// ---- Rx(a) Ry(b) Rz(g) through unit quaternions (AngleAxis products are Quaternion
// products in Eigen; .matrix() is Quaternion::toRotationMatrix()) ---------------------
struct Quat { float w, x, y, z; };
// Coefficient order of Eigen 3.4's SSE quat_product (Geometry/arch/Geometry_SIMD.h), which is what a default x86-64 build
// of the reference runs for Quaternionf * Quaternionf: (a * b.wwww - a.zxyx * b.yzxx) + (a.yzxz * b.zxyz + a.wwwy * b.xyzy),
// the w lane with its second parenthesis negated.
B3D_HD Quat quat_mul(Quat a, Quat b) {
    Quat r;
    r.x = (a.x * b.w - a.z * b.y) + (a.y * b.z + a.w * b.x);
    r.y = (a.y * b.w - a.x * b.z) + (a.z * b.x + a.w * b.y);
    r.z = (a.z * b.w - a.y * b.x) + (a.x * b.y + a.w * b.z);
    r.w = (a.w * b.w - a.x * b.x) + (-(a.z * b.z + a.y * b.y));
    return r;
}
B3D_HD void euler_xyz_to_matrix(float ax, float ay, float az, Mat3& R) {
    float hx = 0.5f * ax, hy = 0.5f * ay, hz = 0.5f * az;
    // std::sin / std::cos of the reference are glibc's sinf / cosf, not CUDA's (b3d_libm.cuh)
    float sx = libm::sin_libm(hx), sy = libm::sin_libm(hy), sz = libm::sin_libm(hz);
    Quat qx{libm::cos_libm(hx), sx * 1.0f, sx * 0.0f, sx * 0.0f};
    Quat qy{libm::cos_libm(hy), sy * 0.0f, sy * 1.0f, sy * 0.0f};
    Quat qz{libm::cos_libm(hz), sz * 0.0f, sz * 0.0f, sz * 1.0f};
    Quat q = quat_mul(quat_mul(qx, qy), qz);
    float tx = 2.0f * q.x, ty = 2.0f * q.y, tz = 2.0f * q.z;
    float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R(0, 0) = 1.0f - (tyy + tzz); R(0, 1) = txy - twz;          R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;          R(1, 1) = 1.0f - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;          R(2, 1) = tyz + twx;          R(2, 2) = 1.0f - (txx + tyy);
}

// 4x4 column-major product C = A*B, accumulation order of Eigen's packet kernel:
// ((a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j) + a_i3 b_3j
B3D_HD void mat4_mul(const float* A, const float* B, float* C) {
    float o[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            float r = A[0 * 4 + i] * B[j * 4 + 0];
            for (int k = 1; k < 4; ++k) r = A[k * 4 + i] * B[j * 4 + k] + r;
            o[j * 4 + i] = r;
        }
    for (int i = 0; i < 16; ++i) C[i] = o[i];
}

}  // namespace b3d
