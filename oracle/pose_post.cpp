// =============================================================================
// oracle/pose_post.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the pose post-processing and mask resize around the registration hot path of
// stojicnnnn/3DVision:
//   nearest-neighbour mask resize      src/pipeline.cpp:38-41   (cv::resize, INTER_NEAREST)
//   T_world_object                     src/pipeline.cpp:136-137 (extrinsics * refined.transformation.inverse())
//   Pipeline::filterDuplicates         src/pipeline.cpp:153-180
// Third-party arithmetic restated from the published algorithms ("parity unpinned", like the rest of oracle/):
//   * Eigen 3.4 Matrix4f::inverse() on SSE = internal::compute_inverse_size4<Architecture::Target, float, ...>
//     (LU/arch/InverseSize4.h): Intel's 2x2-block cofactor scheme; every packet operation below is one
//     4-lane SSE instruction in Eigen, written out lane by lane in the same order.
//   * Matrix4f * Matrix4f: column-major packet product (sequential pmadd over k), as in registration_oracle.cpp.
//   * OpenCV 4.x resizeNN: sx = min(cvFloor(x * ifx), src_w - 1) with ifx = 1 / (dst_w / (double)src_w) in double.
// Build with the Makefile's flags (-ffp-contract=off).
// =============================================================================
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct P4 { float v[4]; };
inline P4 mul(P4 a, P4 b) { return {{a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2], a.v[3] * b.v[3]}}; }
inline P4 add(P4 a, P4 b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2], a.v[3] + b.v[3]}}; }
inline P4 sub(P4 a, P4 b) { return {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2], a.v[3] - b.v[3]}}; }
inline P4 sw2(P4 a, P4 b, int p, int q, int r, int s) { return {{a.v[p], a.v[q], b.v[r], b.v[s]}}; }    // vec4f_swizzle2 = _mm_shuffle_ps
inline P4 movelh(P4 a, P4 b) { return {{a.v[0], a.v[1], b.v[0], b.v[1]}}; }
inline P4 movehl(P4 a, P4 b) { return {{b.v[2], b.v[3], a.v[2], a.v[3]}}; }                               // _mm_movehl_ps(a, b)
inline P4 dup(P4 a, int p) { return {{a.v[p], a.v[p], a.v[p], a.v[p]}}; }

// M, out: column-major 4x4 (Eigen::Matrix4f storage); input and result storage orders match
void mat4_inverse(const float* M, float* out) {
    P4 L1, L2, L3, L4;
    std::memcpy(L1.v, M, 16); std::memcpy(L2.v, M + 4, 16); std::memcpy(L3.v, M + 8, 16); std::memcpy(L4.v, M + 12, 16);
    // four 2x2 sub-matrices of the input: [[A, B], [C, D]]
    const P4 A = movelh(L1, L2), B = movehl(L2, L1), C = movelh(L3, L4), D = movehl(L4, L3);
    // AB = A# * B, DC = D# * C  (# = adjugate)
    P4 AB = mul(sw2(A, A, 3, 3, 0, 0), B);
    AB = sub(AB, mul(sw2(A, A, 1, 1, 2, 2), sw2(B, B, 2, 3, 0, 1)));
    P4 DC = mul(sw2(D, D, 3, 3, 0, 0), C);
    DC = sub(DC, mul(sw2(D, D, 1, 1, 2, 2), sw2(C, C, 2, 3, 0, 1)));
    // determinants of the sub-matrices
    P4 dA = mul(sw2(A, A, 3, 3, 1, 1), A); dA = sub(dA, movehl(dA, dA));
    P4 dB = mul(sw2(B, B, 3, 3, 1, 1), B); dB = sub(dB, movehl(dB, dB));
    P4 dC = mul(sw2(C, C, 3, 3, 1, 1), C); dC = sub(dC, movehl(dC, dC));
    P4 dD = mul(sw2(D, D, 3, 3, 1, 1), D); dD = sub(dD, movehl(dD, dD));
    P4 d = mul(sw2(DC, DC, 0, 2, 1, 3), AB);
    d = add(d, movehl(d, d));
    d = add(d, sw2(d, d, 1, 0, 0, 0));
    const P4 d1 = mul(dA, dD), d2 = mul(dB, dC);
    // det = |A||D| + |B||C| - trace(A# B D# C)
    const P4 det = dup(sub(add(d1, d2), d), 0);
    P4 rd = {{1.0f / det.v[0], 1.0f / det.v[1], 1.0f / det.v[2], 1.0f / det.v[3]}};
    // iD = D |A| - C (A# B)
    P4 iD = mul(sw2(C, C, 0, 0, 2, 2), movelh(AB, AB));
    iD = add(iD, mul(sw2(C, C, 1, 1, 3, 3), movehl(AB, AB)));
    iD = sub(mul(D, dup(dA, 0)), iD);
    // iA = A |D| - B (D# C)
    P4 iA = mul(sw2(B, B, 0, 0, 2, 2), movelh(DC, DC));
    iA = add(iA, mul(sw2(B, B, 1, 1, 3, 3), movehl(DC, DC)));
    iA = sub(mul(A, dup(dD, 0)), iA);
    // iB = C |B| - D (A# B)#
    P4 iB = mul(D, sw2(AB, AB, 3, 0, 3, 0));
    iB = sub(iB, mul(sw2(D, D, 1, 0, 3, 2), sw2(AB, AB, 2, 1, 2, 1)));
    iB = sub(mul(C, dup(dB, 0)), iB);
    // iC = B |C| - A (D# C)#
    P4 iC = mul(A, sw2(DC, DC, 3, 0, 3, 0));
    iC = sub(iC, mul(sw2(A, A, 1, 0, 3, 2), sw2(DC, DC, 2, 1, 2, 1)));
    iC = sub(mul(B, dup(dC, 0)), iC);
    rd.v[1] = -rd.v[1]; rd.v[2] = -rd.v[2];                     // pxor with the sign mask (+, -, -, +)
    iA = mul(iA, rd); iB = mul(iB, rd); iC = mul(iC, rd); iD = mul(iD, rd);
    const P4 r0 = sw2(iA, iB, 3, 1, 3, 1), r1 = sw2(iA, iB, 2, 0, 2, 0), r2 = sw2(iC, iD, 3, 1, 3, 1), r3 = sw2(iC, iD, 2, 0, 2, 0);
    std::memcpy(out, r0.v, 16); std::memcpy(out + 4, r1.v, 16); std::memcpy(out + 8, r2.v, 16); std::memcpy(out + 12, r3.v, 16);
}

// column-major C = A * B, Eigen's packet kernel order: r = a_i0 b_0j; r = a_ik b_kj + r
void mat4_mul(const float* A, const float* B, float* C) {
    float o[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            float r = A[0 * 4 + i] * B[j * 4 + 0];
            for (int k = 1; k < 4; ++k) r = A[k * 4 + i] * B[j * 4 + k] + r;
            o[j * 4 + i] = r;
        }
    std::memcpy(C, o, sizeof(o));
}

inline float norm3(float a0, float a1, float a2) { return std::sqrt(a0 * a0 + (a1 * a1 + a2 * a2)); }    // Eigen 3-term redux a0 + (a1 + a2)

}  // namespace

extern "C" {

void orc_mat4_inverse(const float* M_colmajor, float* out_colmajor) { mat4_inverse(M_colmajor, out_colmajor); }

// pipeline.cpp:136-137.  extrinsics_or_null == NULL returns T_camera_object alone.
void orc_world_pose(const float* extrinsics_or_null, const float* refined_T, float* out) {
    float inv[16];
    mat4_inverse(refined_T, inv);
    if (extrinsics_or_null) mat4_mul(extrinsics_or_null, inv, out); else std::memcpy(out, inv, sizeof(inv));
}

// Pipeline::filterDuplicates, pipeline.cpp:153-180.  poses: n x 16 column-major.  Returns the number kept (out holds n at most).
size_t orc_filter_duplicates(const float* poses, size_t n, float min_distance, float* out) {
    std::vector<const float*> kept;
    for (size_t w = 0; w < n; ++w) {
        const float* wp = poses + 16 * w;
        bool is_dup = false;
        for (size_t i = 0; i < kept.size(); ++i) {
            const float* f = kept[i];
            const float dist = norm3(wp[12] - f[12], wp[13] - f[13], wp[14] - f[14]);
            if (dist < min_distance) {
                is_dup = true;
                const float existing = norm3(f[12], f[13], f[14]), current = norm3(wp[12], wp[13], wp[14]);
                if (current < existing) kept[i] = wp;
                break;
            }
        }
        if (!is_dup) kept.push_back(wp);
    }
    for (size_t i = 0; i < kept.size(); ++i) std::memcpy(out + 16 * i, kept[i], 16 * sizeof(float));
    return kept.size();
}

// cv::resize(mask, resized, depth.size(), 0, 0, INTER_NEAREST), pipeline.cpp:39-41 (single-channel 8-bit)
void orc_resize_mask_nearest(const uint8_t* src, int sw, int sh, int dw, int dh, uint8_t* dst) {
    const double ifx = 1.0 / ((double)dw / sw), ify = 1.0 / ((double)dh / sh);
    for (int y = 0; y < dh; ++y) {
        int sy = (int)std::floor(y * ify); if (sy > sh - 1) sy = sh - 1;
        for (int x = 0; x < dw; ++x) {
            int sx = (int)std::floor(x * ifx); if (sx > sw - 1) sx = sw - 1;
            dst[(size_t)y * dw + x] = src[(size_t)sy * sw + sx];
        }
    }
}

}  // extern "C"
