// =============================================================================
// oracle/registration_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement of the registration hot path of stojicnnnn/3DVision
// (reference: src/registration.cpp:204-414, types include/registration.hpp:10-30)
// written from the reference's *behaviour*, Eigen-free, so that it compiles
// in an image that has no Eigen.  It is the parity oracle for the CUDA path:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.  The product (3dvision_b200/) never
// links, imports or calls anything in this directory.
//
// PARITY STATUS: "parity unpinned".  The reference ships no tests, golden
// vectors or fixtures for this path and cannot be compiled
// here (needs Eigen, absent; no network).  The arithmetic that lives inside
// un-vendored Eigen 3.4 (JacobiSVD<Matrix3f>, LDLT<6x6>, 3-element redux order,
// AngleAxis→Quaternion products, SelfAdjointEigenSolver) is restated below from
// the published Eigen 3.4.0 algorithms.  What *is* pinned: libstdc++'s
// mt19937 + uniform_int_distribution (real std:: objects are used, and a
// hand-rolled Lemire mapping is KAT-checked against them), and analytic
// known-answer tests in tests/.
//
// Floating-point environment of the reference (README.md:13, CMakeLists.txt):
// -O3, baseline x86-64 (SSE2), no -march, no -ffast-math  =>  IEEE fp32, one
// rounding per operation, no FMA contraction.  Build this file with
//   g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math
// =============================================================================
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <random>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// Small fixed-size helpers that mimic Eigen 3.4 evaluation order.
// ---------------------------------------------------------------------------
struct V3 { float v[3]; float& operator[](int i) { return v[i]; } float operator[](int i) const { return v[i]; } };
struct M3 { float a[3][3]; };   // a[row][col]

// Eigen's fully unrolled, non-vectorised redux of 3 coefficients splits the
// range in halves: redux(0,3) = f(redux(0,1), redux(1,2)) = a0 + (a1 + a2).
// Variant builds for the sensitivity sweep of oracle/sensitivity.py (Makefile target `variants`): each ORC_* macro swaps
// ONE of the evaluation orders this restatement assumes of Eigen 3.4 for the other plausible
// reading, so the sweep can report how many bit-exact outputs depend on it.
#if defined(ORC_REDUX3_LEFT)
static inline float red3(float a0, float a1, float a2) { return (a0 + a1) + a2; }                  // left-to-right redux
#else
static inline float red3(float a0, float a1, float a2) { return a0 + (a1 + a2); }
#endif

static inline V3 load3(const float* p) { return V3{{p[0], p[1], p[2]}}; }

// (M * v)[r] = red3(M(r,0)v0, M(r,1)v1, M(r,2)v2)  — lazy coeff-based product.
static inline V3 matvec(const M3& m, const V3& s) {
    V3 o;
    for (int r = 0; r < 3; ++r) o.v[r] = red3(m.a[r][0] * s.v[0], m.a[r][1] * s.v[1], m.a[r][2] * s.v[2]);
    return o;
}
// C = A * B^T : C(i,j) = red3(A(i,0)B(j,0), A(i,1)B(j,1), A(i,2)B(j,2))
static inline M3 mul_abt(const M3& A, const M3& B) {
    M3 c;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c.a[i][j] = red3(A.a[i][0] * B.a[j][0], A.a[i][1] * B.a[j][1], A.a[i][2] * B.a[j][2]);
    return c;
}
// Eigen determinant_impl<3>: bruteforce_det3_helper(m,0,1,2) - (m,1,0,2) + (m,2,0,1)
static inline float det3(const M3& m) {
    auto h = [&](int a, int b, int c) {
        return m.a[0][a] * (m.a[1][b] * m.a[2][c] - m.a[1][c] * m.a[2][b]);
    };
    return h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1);
}

// ---------------------------------------------------------------------------
// Eigen 3.4 JacobiSVD<Matrix3f>(M, ComputeFullU|ComputeFullV), square => no QR
// preconditioner (registration.cpp:255, :388).  Two-sided Jacobi.
// ---------------------------------------------------------------------------
struct Rot { float c, s; };

// JacobiRotation::makeJacobi(x, y, z)
static inline Rot make_jacobi(float x, float y, float z) {
    Rot r;
    float deno = 2.0f * std::fabs(y);
    if (deno < std::numeric_limits<float>::min()) { r.c = 1.0f; r.s = 0.0f; return r; }
    float tau = (x - z) / deno;
    float w = std::sqrt(tau * tau + 1.0f);
    float t = (tau > 0.0f) ? 1.0f / (tau + w) : 1.0f / (tau - w);
    float sign_t = t > 0.0f ? 1.0f : -1.0f;
    float n = 1.0f / std::sqrt(t * t + 1.0f);
    r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
    r.c = n;
    return r;
}

// apply_rotation_in_the_plane on two strided 3-vectors:
//   x' = c*x + s*y ; y' = -s*x + c*y   (skipped when c==1 && s==0)
static inline void rot_plane(float* x, int incx, float* y, int incy, int n, Rot j) {
    if (j.c == 1.0f && j.s == 0.0f) return;
    for (int i = 0; i < n; ++i) {
        float xi = x[i * incx], yi = y[i * incy];
        x[i * incx] = j.c * xi + j.s * yi;
        y[i * incy] = -j.s * xi + j.c * yi;
    }
}
static inline void apply_left(M3& m, int p, int q, Rot j)  { rot_plane(&m.a[p][0], 1, &m.a[q][0], 1, 3, j); }
static inline void apply_right(M3& m, int p, int q, Rot j) { Rot jt{j.c, -j.s}; rot_plane(&m.a[0][p], 3, &m.a[0][q], 3, 3, jt); }

// internal::real_2x2_jacobi_svd
static inline void real_2x2_jacobi_svd(const M3& W, int p, int q, Rot* j_left, Rot* j_right) {
    float m00 = W.a[p][p], m01 = W.a[p][q], m10 = W.a[q][p], m11 = W.a[q][q];
    Rot rot1;
    float t = m00 + m11;
    float d = m10 - m01;
    if (std::fabs(d) < std::numeric_limits<float>::min()) { rot1.s = 0.0f; rot1.c = 1.0f; }
    else {
        float u = t / d;
        float tmp = std::sqrt(1.0f + u * u);
        rot1.s = 1.0f / tmp;
        rot1.c = u / tmp;
    }
    // m.applyOnTheLeft(0,1,rot1)
    if (!(rot1.c == 1.0f && rot1.s == 0.0f)) {
        float a0 = m00, a1 = m01, b0 = m10, b1 = m11;
        m00 = rot1.c * a0 + rot1.s * b0;  m01 = rot1.c * a1 + rot1.s * b1;
        m10 = -rot1.s * a0 + rot1.c * b0; m11 = -rot1.s * a1 + rot1.c * b1;
    }
    *j_right = make_jacobi(m00, m01, m11);
    // *j_left = rot1 * j_right->transpose();  (c1,s1)*(c2,s2) = (c1c2 - s1s2, c1s2 + s1c2)
    Rot jt{j_right->c, -j_right->s};
    j_left->c = rot1.c * jt.c - rot1.s * jt.s;
    j_left->s = rot1.c * jt.s + rot1.s * jt.c;
}

static void jacobi_svd3(const M3& Min, M3& U, M3& V, float S[3]) {
    const float precision = 2.0f * std::numeric_limits<float>::epsilon();
    const float considerAsZero = std::numeric_limits<float>::min();
    float scale = 0.0f;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(Min.a[i][j]));
    if (!(std::isfinite(scale))) {           // m_info = InvalidInput; U,V,S left unspecified by Eigen
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { U.a[i][j] = (i == j); V.a[i][j] = (i == j); }
        S[0] = S[1] = S[2] = 0.0f; return;
    }
    if (scale == 0.0f) scale = 1.0f;
    M3 W;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        W.a[i][j] = Min.a[i][j] / scale;
        U.a[i][j] = (i == j) ? 1.0f : 0.0f;
        V.a[i][j] = (i == j) ? 1.0f : 0.0f;
    }
    float maxDiag = std::max(std::fabs(W.a[0][0]), std::max(std::fabs(W.a[1][1]), std::fabs(W.a[2][2])));
    bool finished = false;
    while (!finished) {
        finished = true;
        for (int p = 1; p < 3; ++p) {
            for (int q = 0; q < p; ++q) {
                float threshold = std::max(considerAsZero, precision * maxDiag);
                if (std::fabs(W.a[p][q]) > threshold || std::fabs(W.a[q][p]) > threshold) {
                    finished = false;
                    Rot jl, jr;
                    real_2x2_jacobi_svd(W, p, q, &jl, &jr);
                    apply_left(W, p, q, jl);
                    apply_right(U, p, q, Rot{jl.c, -jl.s});   // U.applyOnTheRight(p,q,j_left.transpose())
                    apply_right(W, p, q, jr);
                    apply_right(V, p, q, jr);
                    maxDiag = std::max(maxDiag, std::max(std::fabs(W.a[p][p]), std::fabs(W.a[q][q])));
                }
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        float a = W.a[i][i];
        S[i] = std::fabs(a);
        if (a < 0.0f) for (int r = 0; r < 3; ++r) U.a[r][i] = -U.a[r][i];
    }
    for (int i = 0; i < 3; ++i) S[i] *= scale;
    for (int i = 0; i < 3; ++i) {
        int pos = 0; float mx = S[i];
        for (int k = i + 1; k < 3; ++k) if (S[k] > mx) { mx = S[k]; pos = k - i; }
        if (mx == 0.0f) break;
        if (pos) {
            pos += i;
            std::swap(S[i], S[pos]);
            for (int r = 0; r < 3; ++r) { std::swap(U.a[r][i], U.a[r][pos]); std::swap(V.a[r][i], V.a[r][pos]); }
        }
    }
}

// R = V U^T with the reflection fix of registration.cpp:256-262 / :389-394.
static inline M3 rotation_from_svd(const M3& H) {
    M3 U, V; float S[3];
    jacobi_svd3(H, U, V, S);
    M3 R = mul_abt(V, U);
    if (det3(R) < 0.0f) {
        for (int r = 0; r < 3; ++r) V.a[r][2] *= -1.0f;
        R = mul_abt(V, U);
    }
    return R;
}

// ---------------------------------------------------------------------------
// Eigen 3.4 LDLT<Matrix<float,6,6>, Lower>: compute() + solve()
// (registration.cpp:366).  In-place, diagonal pivoting, pseudo-inverse of D.
// ---------------------------------------------------------------------------
static inline float redn_halving(const float* c, int n) {     // redux_novec_unroller
#if defined(ORC_LDLT_LINEAR)
    float acc = c[0];                                          // variant: plain left-to-right sum
    for (int i = 1; i < n; ++i) acc = acc + c[i];
    return acc;
#endif
    if (n == 1) return c[0];
    int h = n / 2;
    return redn_halving(c, h) + redn_halving(c + h, n - h);
}

static void ldlt6_solve(const float Ain[6][6], const float bin[6], float x[6]) {
    const int N = 6;
    float m[6][6];
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) m[i][j] = Ain[i][j];
    int tr[6];
    float temp[6];
    bool all_zero_diag = false;
    for (int k = 0; k < N; ++k) {
        int big = k; float best = std::fabs(m[k][k]);
        for (int i = k + 1; i < N; ++i) if (std::fabs(m[i][i]) > best) { best = std::fabs(m[i][i]); big = i; }
        tr[k] = big;
        if (k != big) {
            int s = N - big - 1;
            for (int j = 0; j < k; ++j) std::swap(m[k][j], m[big][j]);
            for (int i = 0; i < s; ++i) std::swap(m[big + 1 + i][k], m[big + 1 + i][big]);
            std::swap(m[k][k], m[big][big]);
            for (int i = k + 1; i < big; ++i) { float tmp = m[i][k]; m[i][k] = m[big][i]; m[big][i] = tmp; }
        }
        int rs = N - k - 1;
        if (k > 0) {
            for (int j = 0; j < k; ++j) temp[j] = m[j][j] * m[k][j];
            float acc = m[k][0] * temp[0];
            for (int j = 1; j < k; ++j) acc = acc + m[k][j] * temp[j];
            m[k][k] -= acc;
            for (int i = 0; i < rs; ++i) {
                float c = m[k + 1 + i][0] * temp[0];
                for (int j = 1; j < k; ++j) c = c + m[k + 1 + i][j] * temp[j];
                m[k + 1 + i][k] -= c;
            }
        }
        float akk = m[k][k];
        bool pivot_ok = std::fabs(akk) > 0.0f;
        if (k == 0 && !pivot_ok) {
            for (int j = 0; j < N; ++j) tr[j] = j;
            all_zero_diag = true;
            break;
        }
        if (rs > 0 && pivot_ok) for (int i = 0; i < rs; ++i) m[k + 1 + i][k] /= akk;
    }
    (void)all_zero_diag;
    // solve: dst = P b
    float d[6];
    for (int i = 0; i < N; ++i) d[i] = bin[i];
    for (int k = 0; k < N; ++k) if (tr[k] != k) std::swap(d[k], d[tr[k]]);
    // L^-1 (unit lower; fully unrolled, row-strided => non-vectorised halving redux)
    for (int i = 1; i < N; ++i) {
        float c[6];
        for (int j = 0; j < i; ++j) c[j] = m[i][j] * d[j];
        d[i] -= redn_halving(c, i);
    }
    // pseudo-inverse of D
    for (int i = 0; i < N; ++i) {
        if (std::fabs(m[i][i]) > std::numeric_limits<float>::min()) d[i] /= m[i][i];
        else d[i] = 0.0f;
    }
    // L^-T (unit upper; rows of L^T are contiguous columns of L => Packet4f redux when len >= 4)
    for (int li = 1; li < N; ++li) {
        int di = N - li - 1, st = di + 1;
        float c[6];
        for (int j = 0; j < li; ++j) c[j] = m[st + j][di] * d[st + j];
        float sum;
#if defined(ORC_LDLT_LINEAR)
        sum = redn_halving(c, li);
#else
        if (li >= 4) {
            sum = (c[0] + c[2]) + (c[1] + c[3]);            // SSE predux: (a0+a2)+(a1+a3)
            if (li == 5) sum = sum + c[4];
        } else sum = redn_halving(c, li);
#endif
        d[di] -= sum;
    }
    // P^T
    for (int k = N - 1; k >= 0; --k) if (tr[k] != k) std::swap(d[k], d[tr[k]]);
    for (int i = 0; i < N; ++i) x[i] = d[i];
}

// ---------------------------------------------------------------------------
// (AngleAxisf(a,X) * AngleAxisf(b,Y) * AngleAxisf(g,Z)).matrix()
// registration.cpp:369-371.  In Eigen, AngleAxis*AngleAxis returns a
// Quaternion product, and .matrix() is Quaternion::toRotationMatrix().
// ---------------------------------------------------------------------------
struct Quat { float w, x, y, z; };
static inline Quat quat_from_aa(float angle, int axis) {
    float ha = 0.5f * angle;
    float s = std::sin(ha), c = std::cos(ha);
    Quat q{c, 0.0f, 0.0f, 0.0f};
    // sin(ha) * axis  (axis is a unit basis vector: s*1 or s*0)
    float v[3] = {s * (axis == 0 ? 1.0f : 0.0f), s * (axis == 1 ? 1.0f : 0.0f), s * (axis == 2 ? 1.0f : 0.0f)};
    q.x = v[0]; q.y = v[1]; q.z = v[2];
    return q;
}
// Quaternionf * Quaternionf.  On x86-64 (SSE2 is baseline, EIGEN_VECTORIZE_SSE is on by default) Eigen 3.4 does not
// run the generic scalar product but internal::quat_product<Architecture::Target, ..., float> (Geometry/arch/
// Geometry_SIMD.h): with a = (x,y,z,w) packets, res = (a * b.wwww - a.zxyx * b.yzxx) + (sign-flipped in w)(a.yzxz * b.zxyz
// + a.wwwy * b.xyzy), i.e. per coefficient (t1 - t2) + (s1 + s2), the w lane negating its second parenthesis.
// ORC_QUAT_SCALAR: the generic template's left-to-right order instead (round 1's reading; moves ICP results at 1e-7).
static inline Quat quat_mul(const Quat& a, const Quat& b) {
    Quat r;
#if defined(ORC_QUAT_SCALAR)
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
#else
    r.x = (a.x * b.w - a.z * b.y) + (a.y * b.z + a.w * b.x);
    r.y = (a.y * b.w - a.x * b.z) + (a.z * b.x + a.w * b.y);
    r.z = (a.z * b.w - a.y * b.x) + (a.x * b.y + a.w * b.z);
    r.w = (a.w * b.w - a.x * b.x) + (-(a.z * b.z + a.y * b.y));
#endif
    return r;
}
static inline M3 quat_to_matrix(const Quat& q) {
    M3 r;
    const float tx = 2.0f * q.x, ty = 2.0f * q.y, tz = 2.0f * q.z;
    const float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    r.a[0][0] = 1.0f - (tyy + tzz); r.a[0][1] = txy - twz;          r.a[0][2] = txz + twy;
    r.a[1][0] = txy + twz;          r.a[1][1] = 1.0f - (txx + tzz); r.a[1][2] = tyz - twx;
    r.a[2][0] = txz - twy;          r.a[2][1] = tyz + twx;          r.a[2][2] = 1.0f - (txx + tyy);
    return r;
}

// 4x4 column-major helpers (Eigen::Matrix4f storage): T[c*4 + r]
static inline float& T_at(float* T, int r, int c) { return T[c * 4 + r]; }
static inline float T_at(const float* T, int r, int c) { return T[c * 4 + r]; }
static inline void T_identity(float* T) { for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0f : 0.0f; }
// Matrix4f * Matrix4f: column-major lhs, Packet4f rows => sequential pmadd over k.
static inline void T_mul(const float* A, const float* B, float* C) {
    float out[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
#if defined(ORC_MAT4_PAIRWISE)
            float r = (T_at(A, i, 0) * T_at(B, 0, j) + T_at(A, i, 1) * T_at(B, 1, j)) + (T_at(A, i, 2) * T_at(B, 2, j) + T_at(A, i, 3) * T_at(B, 3, j));
#else
            float r = T_at(A, i, 0) * T_at(B, 0, j);
            for (int k = 1; k < 4; ++k) r = T_at(A, i, k) * T_at(B, k, j) + r;
#endif
            out[j * 4 + i] = r;
        }
    std::memcpy(C, out, sizeof(out));
}

// ---------------------------------------------------------------------------
// Feature matching, registration.cpp:216-232
// ---------------------------------------------------------------------------
static void match_rows(const float* sd, size_t row0, size_t row1, const float* td, size_t nt, uint32_t* corr) {
    for (size_t i = row0; i < row1; ++i) {
        float best_dist = std::numeric_limits<float>::max();
        size_t best_idx = 0;
        const float* a = sd + i * 33;
        for (size_t j = 0; j < nt; ++j) {
            const float* b = td + j * 33;
            float dist = 0;
            for (int d = 0; d < 33; ++d) {
                float diff = a[d] - b[d];
                dist += diff * diff;
            }
            if (dist < best_dist) { best_dist = dist; best_idx = j; }
        }
        corr[i] = (uint32_t)best_idx;
    }
}

// 3-point Kabsch, registration.cpp:242-268.  Returns R (row/col) and t.
static inline void kabsch3(const V3 s[3], const V3 q[3], M3& R, V3& t) {
    V3 sc, tc;
    for (int r = 0; r < 3; ++r) {
        sc.v[r] = red3(s[0].v[r], s[1].v[r], s[2].v[r]) / 3.0f;    // rowwise().mean()
        tc.v[r] = red3(q[0].v[r], q[1].v[r], q[2].v[r]) / 3.0f;
    }
    M3 Sc, Tc;   // points as columns, centred
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) { Sc.a[r][k] = s[k].v[r] - sc.v[r]; Tc.a[r][k] = q[k].v[r] - tc.v[r]; }
    M3 H = mul_abt(Sc, Tc);
    R = rotation_from_svd(H);
    V3 Rs = matvec(R, sc);
    for (int r = 0; r < 3; ++r) t.v[r] = tc.v[r] - Rs.v[r];
}

}  // namespace

// =============================================================================
// extern "C" surface (ctypes-loaded by oracle/oracle.py)
// =============================================================================
extern "C" {

// -- RNG known-answer taps --------------------------------
void orc_mt19937_raw(uint32_t seed, size_t count, uint32_t* out) {
    std::mt19937 rng(seed);
    for (size_t i = 0; i < count; ++i) out[i] = (uint32_t)rng();
}
// the real libstdc++ distribution, as registration.cpp:235-239 uses it
void orc_uniform_indices_std(uint32_t seed, uint64_t n, size_t count, uint64_t* out) {
    std::mt19937 rng(seed);
    std::uniform_int_distribution<size_t> dist(0, (size_t)n - 1);
    for (size_t i = 0; i < count; ++i) out[i] = dist(rng);
}
// hand-rolled Lemire mapping (what the CUDA path implements); returns raw draws consumed
size_t orc_uniform_indices_lemire(uint32_t seed, uint64_t n, size_t count, uint64_t* out) {
    std::mt19937 rng(seed);
    size_t consumed = 0;
    const uint32_t R = (uint32_t)n;
    for (size_t i = 0; i < count; ++i) {
        uint64_t p = (uint64_t)(uint32_t)rng() * R; ++consumed;
        uint32_t lo = (uint32_t)p;
        if (lo < R) {
            uint32_t thr = (uint32_t)(-R) % R;
            while (lo < thr) { p = (uint64_t)(uint32_t)rng() * R; ++consumed; lo = (uint32_t)p; }
        }
        out[i] = p >> 32;
    }
    return consumed;
}

// -- linear-algebra taps --------------------------------------------------------
// M, U, V row-major 3x3
void orc_svd3(const float* M, float* U, float* V, float* S) {
    M3 m, u, v; std::memcpy(m.a, M, 36);
    jacobi_svd3(m, u, v, S);
    std::memcpy(U, u.a, 36); std::memcpy(V, v.a, 36);
}
void orc_ldlt6_solve(const float* A_rowmajor, const float* b, float* x) {
    float A[6][6]; std::memcpy(A, A_rowmajor, sizeof(A));
    ldlt6_solve(A, b, x);
}
// s,q: 3 points each (row i = point i), R row-major, t
void orc_kabsch3(const float* s, const float* q, float* R, float* t) {
    V3 sv[3] = {load3(s), load3(s + 3), load3(s + 6)};
    V3 qv[3] = {load3(q), load3(q + 3), load3(q + 6)};
    M3 r; V3 tt; kabsch3(sv, qv, r, tt);
    std::memcpy(R, r.a, 36); std::memcpy(t, tt.v, 12);
}
void orc_euler_xyz(float a, float b, float g, float* R_rowmajor) {
    Quat q = quat_mul(quat_mul(quat_from_aa(a, 0), quat_from_aa(b, 1)), quat_from_aa(g, 2));
    M3 r = quat_to_matrix(q); std::memcpy(R_rowmajor, r.a, 36);
}

// -- feature matching (registration.cpp:216-232) --------------------------------
void orc_match_features(const float* src_desc, size_t row0, size_t row1,
                        const float* tgt_desc, size_t n_tgt, uint32_t* corr /* [n_src], rows row0..row1 written */) {
    match_rows(src_desc, row0, row1, tgt_desc, n_tgt, corr);
}

// -- RANSAC (registration.cpp:234-295) given correspondences --------------------
// counts (optional, [max_iterations]): inliers per iteration; -1 = degenerate triple
// (`continue`, :240); -2 = not executed (after the early-exit break, :290).
// iter_lo/iter_hi: only iterations in [iter_lo, iter_hi) are *scored* (the RNG is
// still advanced for all earlier ones); used for bounded-slice timing and for
// sharded-hypothesis tests.  Pass 0, max_iterations for the reference behaviour.
int orc_ransac(const float* src, size_t n_src, const float* tgt, size_t n_tgt,
               const uint32_t* corr, float voxel_size, int max_iterations, float confidence,
               int iter_lo, int iter_hi,
               float* out_T, float* out_fitness, float* out_rmse,
               int32_t* counts, int32_t* best_iter, int32_t* iters_run) {
    (void)n_tgt;
    float distance_threshold = voxel_size * 1.5f;
    float bestT[16]; T_identity(bestT);
    float best_fitness = 0.0f, best_rmse = 0.0f;
    int best_id = -1, run = 0;
    if (counts) for (int i = 0; i < max_iterations; ++i) counts[i] = -2;
    std::mt19937 rng(42);
    std::uniform_int_distribution<size_t> dist(0, n_src - 1);
    for (int iter = 0; iter < max_iterations; ++iter) {
        size_t i0 = dist(rng), i1 = dist(rng), i2 = dist(rng);
        if (iter < iter_lo) continue;
        if (iter >= iter_hi) break;
        run = iter + 1;
        if (i0 == i1 || i1 == i2 || i0 == i2) { if (counts) counts[iter] = -1; continue; }
        V3 s[3] = {load3(src + 3 * i0), load3(src + 3 * i1), load3(src + 3 * i2)};
        V3 q[3] = {load3(tgt + 3 * (size_t)corr[i0]), load3(tgt + 3 * (size_t)corr[i1]), load3(tgt + 3 * (size_t)corr[i2])};
        M3 R; V3 t; kabsch3(s, q, R, t);
        int inliers = 0; float total_error = 0;
        for (size_t i = 0; i < n_src; ++i) {
            V3 p = matvec(R, load3(src + 3 * i));
            const float* qq = tgt + 3 * (size_t)corr[i];
            float d0 = (p.v[0] + t.v[0]) - qq[0], d1 = (p.v[1] + t.v[1]) - qq[1], d2 = (p.v[2] + t.v[2]) - qq[2];
            float err = std::sqrt(red3(d0 * d0, d1 * d1, d2 * d2));
            if (err < distance_threshold) { ++inliers; total_error += err * err; }
        }
        if (counts) counts[iter] = inliers;
        float fitness = static_cast<float>(inliers) / n_src;
        float rmse = inliers > 0 ? std::sqrt(total_error / inliers) : 999.0f;
        if (fitness > best_fitness) {
            T_identity(bestT);
            for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T_at(bestT, r, c) = R.a[r][c]; T_at(bestT, r, 3) = t.v[r]; }
            best_fitness = fitness; best_rmse = rmse; best_id = iter;
        }
        if (fitness > confidence) break;
    }
    std::memcpy(out_T, bestT, sizeof(bestT));
    *out_fitness = best_fitness; *out_rmse = best_rmse;
    if (best_iter) *best_iter = best_id;
    if (iters_run) *iters_run = run;
    return 0;
}

// hypothesis tap: (R row-major, t) of iteration `iter`; returns 0 if degenerate
int orc_ransac_hypothesis(const float* src, size_t n_src, const float* tgt, const uint32_t* corr,
                          int iter, float* R, float* t, uint64_t* triple) {
    std::mt19937 rng(42);
    std::uniform_int_distribution<size_t> dist(0, n_src - 1);
    size_t i0 = 0, i1 = 0, i2 = 0;
    for (int k = 0; k <= iter; ++k) { i0 = dist(rng); i1 = dist(rng); i2 = dist(rng); }
    if (triple) { triple[0] = i0; triple[1] = i1; triple[2] = i2; }
    if (i0 == i1 || i1 == i2 || i0 == i2) return 0;
    V3 s[3] = {load3(src + 3 * i0), load3(src + 3 * i1), load3(src + 3 * i2)};
    V3 q[3] = {load3(tgt + 3 * (size_t)corr[i0]), load3(tgt + 3 * (size_t)corr[i1]), load3(tgt + 3 * (size_t)corr[i2])};
    M3 Rm; V3 tv; kabsch3(s, q, Rm, tv);
    std::memcpy(R, Rm.a, 36); std::memcpy(t, tv.v, 12);
    return 1;
}

// full Registration::ransacRegistration (registration.cpp:204-295)
int orc_ransac_registration(const float* src, size_t n_src, const float* tgt, size_t n_tgt,
                            const float* src_desc, const float* tgt_desc,
                            float voxel_size, int max_iterations, float confidence,
                            float* out_T, float* out_fitness, float* out_rmse) {
    std::vector<uint32_t> corr(n_src);
    match_rows(src_desc, 0, n_src, tgt_desc, n_tgt, corr.data());
    return orc_ransac(src, n_src, tgt, n_tgt, corr.data(), voxel_size, max_iterations, confidence,
                      0, max_iterations, out_T, out_fitness, out_rmse, nullptr, nullptr, nullptr);
}

// -- ICP (registration.cpp:297-414) ------------------------------------------
// nn_idx0 / nn_d2_0 (optional, [n_src]): brute-force NN index and squared distance
// of iteration 0 (before thresholding) — the bit-exact tap.
// stop_on_convergence=0 disables the :406 break (used for iters/s timing only).
int orc_icp(const float* src, size_t n_src, const float* tgt, const float* tgt_normals, size_t n_tgt,
            const float* T0, float distance_threshold, int max_iterations, int point_to_plane,
            int stop_on_convergence,
            float* out_T, float* out_fitness, float* out_rmse, int32_t* iters_run,
            uint32_t* nn_idx0, float* nn_d2_0, int32_t* ncorr_per_iter) {
    float T[16]; std::memcpy(T, T0, sizeof(T));
    float resT[16]; std::memcpy(resT, T0, sizeof(resT));
    float res_fitness = 0.0f, res_rmse = 0.0f;
    const bool plane = point_to_plane && tgt_normals != nullptr;
    int it_done = 0;
    std::vector<V3> src_corr, tgt_corr;
    for (int iter = 0; iter < max_iterations; ++iter) {
        M3 R; V3 t;
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R.a[r][c] = T_at(T, r, c); t.v[r] = T_at(T, r, 3); }
        int n_corr = 0; float total_error = 0;
        float ATA[6][6]; float ATb[6];
        for (int i = 0; i < 6; ++i) { ATb[i] = 0; for (int j = 0; j < 6; ++j) ATA[i][j] = 0; }
        src_corr.clear(); tgt_corr.clear();
        for (size_t i = 0; i < n_src; ++i) {
            V3 p = matvec(R, load3(src + 3 * i));
            for (int r = 0; r < 3; ++r) p.v[r] = p.v[r] + t.v[r];
            float best_dist2 = std::numeric_limits<float>::max();
            size_t best_idx = 0;
            for (size_t j = 0; j < n_tgt; ++j) {
                const float* q = tgt + 3 * j;
                float e0 = p.v[0] - q[0], e1 = p.v[1] - q[1], e2 = p.v[2] - q[2];
                float d2 = red3(e0 * e0, e1 * e1, e2 * e2);
                if (d2 < best_dist2) { best_dist2 = d2; best_idx = j; }
            }
            if (iter == 0) { if (nn_idx0) nn_idx0[i] = (uint32_t)best_idx; if (nn_d2_0) nn_d2_0[i] = best_dist2; }
            float d = std::sqrt(best_dist2);
            if (d > distance_threshold) continue;
            ++n_corr;
            total_error += best_dist2;
            const float* q = tgt + 3 * best_idx;
            if (plane) {
                const float* n = tgt_normals + 3 * best_idx;
                float J[6];
                J[0] = p.v[1] * n[2] - p.v[2] * n[1];
                J[1] = p.v[2] * n[0] - p.v[0] * n[2];
                J[2] = p.v[0] * n[1] - p.v[1] * n[0];
                J[3] = n[0]; J[4] = n[1]; J[5] = n[2];
                float residual = red3((p.v[0] - q[0]) * n[0], (p.v[1] - q[1]) * n[1], (p.v[2] - q[2]) * n[2]);
                for (int a = 0; a < 6; ++a) { for (int b = 0; b < 6; ++b) ATA[a][b] += J[a] * J[b]; ATb[a] += J[a] * residual; }
            } else {
                src_corr.push_back(p);
                tgt_corr.push_back(load3(q));
            }
        }
        if (ncorr_per_iter) ncorr_per_iter[iter] = n_corr;
        if (n_corr < 3) break;
        float delta[16]; T_identity(delta);
        if (plane) {
            float nb[6], x[6];
            for (int i = 0; i < 6; ++i) nb[i] = -ATb[i];
            ldlt6_solve(ATA, nb, x);
            Quat qd = quat_mul(quat_mul(quat_from_aa(x[0], 0), quat_from_aa(x[1], 1)), quat_from_aa(x[2], 2));
            M3 dR = quat_to_matrix(qd);
            for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T_at(delta, r, c) = dR.a[r][c]; T_at(delta, r, 3) = x[3 + r]; }
        } else {
            V3 sm{{0, 0, 0}}, tm{{0, 0, 0}};
            for (size_t i = 0; i < src_corr.size(); ++i) for (int r = 0; r < 3; ++r) { sm.v[r] += src_corr[i].v[r]; tm.v[r] += tgt_corr[i].v[r]; }
            float nf = static_cast<float>(src_corr.size());
            for (int r = 0; r < 3; ++r) { sm.v[r] /= nf; tm.v[r] /= nf; }
            M3 H; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) H.a[i][j] = 0;
            for (size_t i = 0; i < src_corr.size(); ++i) {
                float a[3], b[3];
                for (int r = 0; r < 3; ++r) { a[r] = src_corr[i].v[r] - sm.v[r]; b[r] = tgt_corr[i].v[r] - tm.v[r]; }
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H.a[r][c] += a[r] * b[c];
            }
            M3 dR = rotation_from_svd(H);
            V3 Rs = matvec(dR, sm);
            for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T_at(delta, r, c) = dR.a[r][c]; T_at(delta, r, 3) = tm.v[r] - Rs.v[r]; }
        }
        T_mul(delta, T, T);
        float prev_rmse = res_rmse;
        res_rmse = std::sqrt(total_error / n_corr);
        res_fitness = static_cast<float>(n_corr) / n_src;
        std::memcpy(resT, T, sizeof(T));
        it_done = iter + 1;
        if (stop_on_convergence && iter > 0 && std::fabs(prev_rmse - res_rmse) < 1e-6f) break;
    }
    std::memcpy(out_T, resT, sizeof(resT));
    *out_fitness = res_fitness; *out_rmse = res_rmse;
    if (iters_run) *iters_run = it_done;
    return 0;
}

}  // extern "C"
