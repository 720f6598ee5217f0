// =============================================================================
// oracle/pipeline_inputs.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the stages that *feed* the registration hot path in
// stojicnnnn/3DVision, used only to build the inputs of BASELINE.json
// configs[0] (the demo procedural box scene) for golden fixtures:
//   procedural scene / mask / model   src/pipeline.cpp:211-257, 275-282
//   mask + scale + deprojection       src/pipeline.cpp:42-84
//   voxelDownsample                   src/registration.cpp:15-60
//   estimateNormals (+findKNN)        src/registration.cpp:63-81, 105-130
//   computeFPFH (+findRadiusNN)       src/registration.cpp:83-102, 133-201
// These stages are OUT of the round-1 hot-path scope (SURVEY.md §8f rows
// f-1..f-4); every function here is a parity target of the widened path.  Eigen's
// SelfAdjointEigenSolver<Matrix3f> is restated from the Eigen 3.4.0 algorithm
// (scaled tridiagonalisation + implicit symmetric QR) — "parity unpinned".
// Build: g++ -O2 -std=c++17 -ffp-contract=off
// =============================================================================
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

struct P3 { float x, y, z; };
#if defined(ORC_REDUX3_LEFT)                                   // sensitivity-sweep variant, see registration_oracle.cpp
static inline float red3(float a0, float a1, float a2) { return (a0 + a1) + a2; }
#else
static inline float red3(float a0, float a1, float a2) { return a0 + (a1 + a2); }
#endif
static inline float sqnorm(const P3& a, const P3& b) {   // (a - b).squaredNorm()
    float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z;
    return red3(d0 * d0, d1 * d1, d2 * d2);
}

// registration.cpp:15-27
struct VoxelKey { int x, y, z; bool operator==(const VoxelKey& o) const { return x == o.x && y == o.y && z == o.z; } };
struct VoxelKeyHash {
    size_t operator()(const VoxelKey& k) const {
        size_t h = std::hash<int>()(k.x);
        h ^= std::hash<int>()(k.y) + 0x9e3779b9 + (h << 6) + (h >> 2);
        h ^= std::hash<int>()(k.z) + 0x9e3779b9 + (h << 6) + (h >> 2);
        return h;
    }
};

// ---- Eigen 3.4 SelfAdjointEigenSolver<Matrix3f>::compute (iterative path) ----
struct Giv { float c, s; };
static inline Giv make_givens(float p, float q) {
    Giv g;
    if (q == 0.0f) { g.c = p < 0.0f ? -1.0f : 1.0f; g.s = 0.0f; }
    else if (p == 0.0f) { g.c = 0.0f; g.s = q < 0.0f ? 1.0f : -1.0f; }
    else if (std::fabs(p) > std::fabs(q)) {
        float t = q / p; float u = std::sqrt(1.0f + t * t); if (p < 0.0f) u = -u;
        g.c = 1.0f / u; g.s = -t * g.c;
    } else {
        float t = p / q; float u = std::sqrt(1.0f + t * t); if (q < 0.0f) u = -u;
        g.s = -1.0f / u; g.c = -t * g.s;
    }
    return g;
}
static inline float eig_hypot(float x, float y) {
    x = std::fabs(x); y = std::fabs(y);
    float p = std::max(x, y);
    if (p == 0.0f) return 0.0f;
    float qp = std::min(y, x) / p;
    return p * std::sqrt(1.0f + qp * qp);
}
// A symmetric (row-major a[r][c], lower triangle read). evals ascending, evecs columns.
static void self_adjoint_eig3(const float A[3][3], float evals[3], float Q[3][3]) {
    float m[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = (j <= i) ? A[i][j] : 0.0f;
    float scale = 0.0f;
    for (int i = 0; i < 3; ++i) for (int j = 0; j <= i; ++j) scale = std::max(scale, std::fabs(m[i][j]));
    if (scale == 0.0f) scale = 1.0f;
    for (int i = 0; i < 3; ++i) for (int j = 0; j <= i; ++j) m[i][j] /= scale;
    float diag[3], sub[2];
    const float tol = std::numeric_limits<float>::min();
    diag[0] = m[0][0];
    float v1norm2 = m[2][0] * m[2][0];
    if (v1norm2 <= tol) {
        diag[1] = m[1][1]; diag[2] = m[2][2]; sub[0] = m[1][0]; sub[1] = m[2][1];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Q[i][j] = (i == j) ? 1.0f : 0.0f;
    } else {
        float beta = std::sqrt(m[1][0] * m[1][0] + v1norm2);
        float invBeta = 1.0f / beta;
        float m01 = m[1][0] * invBeta, m02 = m[2][0] * invBeta;
        float q = 2.0f * m01 * m[2][1] + m02 * (m[2][2] - m[1][1]);
        diag[1] = m[1][1] + m02 * q; diag[2] = m[2][2] - m02 * q;
        sub[0] = beta; sub[1] = m[2][1] - m01 * q;
        float QQ[3][3] = {{1, 0, 0}, {0, m01, m02}, {0, m02, -m01}};
        std::memcpy(Q, QQ, sizeof(QQ));
    }
    const int n = 3; int end = n - 1, start = 0, iter = 0; const int maxIt = 30;
    const float considerAsZero = std::numeric_limits<float>::min();
    const float precision_inv = 1.0f / std::numeric_limits<float>::epsilon();
    while (end > 0) {
        for (int i = start; i < end; ++i) {
            if (std::fabs(sub[i]) < considerAsZero) sub[i] = 0.0f;
            else {
                const float ss = precision_inv * sub[i];
                if (ss * ss <= (std::fabs(diag[i]) + std::fabs(diag[i + 1]))) sub[i] = 0.0f;
            }
        }
        while (end > 0 && sub[end - 1] == 0.0f) end--;
        if (end <= 0) break;
        iter++; if (iter > maxIt * n) break;
        start = end - 1;
        while (start > 0 && sub[start - 1] != 0.0f) start--;
        // tridiagonal_qr_step
        float td = (diag[end - 1] - diag[end]) * 0.5f;
        float e = sub[end - 1];
        float mu = diag[end];
        if (td == 0.0f) mu -= std::fabs(e);
        else if (e != 0.0f) {
            const float e2 = e * e; const float h = eig_hypot(td, e);
            if (e2 == 0.0f) mu -= e / ((td + (td > 0.0f ? h : -h)) / e);
            else mu -= e2 / (td + (td > 0.0f ? h : -h));
        }
        float x = diag[start] - mu, z = sub[start];
        for (int k = start; k < end && z != 0.0f; ++k) {
            Giv rot = make_givens(x, z);
            float sdk = rot.s * diag[k] + rot.c * sub[k];
            float dkp1 = rot.s * sub[k] + rot.c * diag[k + 1];
            diag[k] = rot.c * (rot.c * diag[k] - rot.s * sub[k]) - rot.s * (rot.c * sub[k] - rot.s * diag[k + 1]);
            diag[k + 1] = rot.s * sdk + rot.c * dkp1;
            sub[k] = rot.c * sdk - rot.s * dkp1;
            if (k > start) sub[k - 1] = rot.c * sub[k - 1] - rot.s * z;
            x = sub[k];
            if (k < end - 1) { z = -rot.s * sub[k + 1]; sub[k + 1] = rot.c * sub[k + 1]; }
            // q.applyOnTheRight(k,k+1,rot): cols k,k+1 with rot^T = (c,-s)
            if (!(rot.c == 1.0f && rot.s == 0.0f)) {
                for (int r = 0; r < 3; ++r) {
                    float xi = Q[r][k], yi = Q[r][k + 1];
                    Q[r][k] = rot.c * xi + (-rot.s) * yi;
                    Q[r][k + 1] = rot.s * xi + rot.c * yi;
                }
            }
        }
    }
    for (int i = 0; i < n - 1; ++i) {
        int k = 0; float mn = diag[i];
        for (int j = i + 1; j < n; ++j) if (diag[j] < mn) { mn = diag[j]; k = j - i; }
        if (k > 0) { std::swap(diag[i], diag[k + i]); for (int r = 0; r < 3; ++r) std::swap(Q[r][i], Q[r][k + i]); }
    }
    for (int i = 0; i < 3; ++i) evals[i] = diag[i] * scale;
}

}  // namespace

extern "C" {

// Procedural scene + dummy mask + deprojection (pipeline.cpp:211-257, 42-84), CPU branch.
// Returns the number of points written (xyz packed); capacity must be >= w*h.
size_t orc_demo_scene_points(int w, int h, float scale_to_meters, float clipping_max, float* xyz_out) {
    float fx = 900, fy = 900, cx = w / 2.0f, cy = h / 2.0f;
    const float floor_z = 1.0f, box_z = 0.8f;
    int mcx = w / 2, mcy = h / 2;
    size_t n = 0;
    for (int v = 0; v < h; ++v) {
        for (int u = 0; u < w; ++u) {
            float z = floor_z;
            if (std::abs(u - cx) < 100 && std::abs(v - cy) < 100) z = box_z;
            unsigned short d_val = static_cast<unsigned short>(z * scale_to_meters);
            // convertTo(CV_32FC1, 1.0/scale): OpenCV's 16u -> 32f scaling works in float (cvtScale16u32f casts alpha to float and
            // evaluates src*a + b per element, one rounding) — third-party behaviour, restated
            float fz = (float)d_val * (float)(1.0 / (double)scale_to_meters);
            bool in_mask = (u >= mcx - 100 && u <= mcx + 100 && v >= mcy - 100 && v <= mcy + 100);  // cv::rectangle, inclusive
            if (!in_mask) fz = 0.0f;
            if (fz <= 0 || fz > clipping_max) continue;
            float x = (u - cx) * fz / fx;
            float y = (v - cy) * fz / fy;
            xyz_out[3 * n + 0] = x; xyz_out[3 * n + 1] = y; xyz_out[3 * n + 2] = fz; ++n;
        }
    }
    return n;
}

// Mask + scale + deprojection of one instance, CPU branch of Pipeline::processInstance (pipeline.cpp:38-84), for any
// depth image: convertTo(CV_32FC1, 1/scale) [float arithmetic, as OpenCV's 16u->32f scaling]; threshold(mask, 10, 255, BINARY) and setTo(0, mask_bool == 0) when a mask is
// given (apply_mask); skip z <= 0 or z > clipping_max; x = (u - cx) * z / fx; colours bgr -> rgb / 255.  Raster order.
size_t orc_depth_to_cloud(const uint16_t* depth, int w, int h, const uint8_t* mask_or_null, float scale_to_meters, float clipping_max,
                          float fx, float fy, float cx, float cy, const uint8_t* bgr_or_null, float* xyz_out, float* rgb_out) {
    size_t n = 0;
    for (int v = 0; v < h; ++v) {
        for (int u = 0; u < w; ++u) {
            const size_t px = (size_t)v * w + u;
            float z = (float)depth[px] * (float)(1.0 / (double)scale_to_meters);   // OpenCV cvtScale16u32f: float arithmetic, see above
            if (mask_or_null && !(mask_or_null[px] > 10)) z = 0.0f;
            if (z <= 0 || z > clipping_max) continue;
            xyz_out[3 * n] = (u - cx) * z / fx; xyz_out[3 * n + 1] = (v - cy) * z / fy; xyz_out[3 * n + 2] = z;
            if (bgr_or_null && rgb_out) {
                rgb_out[3 * n] = bgr_or_null[3 * px + 2] / 255.0f; rgb_out[3 * n + 1] = bgr_or_null[3 * px + 1] / 255.0f;
                rgb_out[3 * n + 2] = bgr_or_null[3 * px] / 255.0f;
            }
            ++n;
        }
    }
    return n;
}

// Dummy model (pipeline.cpp:275-282). Returns count; capacity >= 41*41.
size_t orc_demo_model_points(float* xyz_out) {
    size_t n = 0;
    for (float x = -0.1f; x <= 0.1f; x += 0.005f)
        for (float y = -0.1f; y <= 0.1f; y += 0.005f) { xyz_out[3 * n] = x; xyz_out[3 * n + 1] = y; xyz_out[3 * n + 2] = 0.0f; ++n; }
    return n;
}

// voxelDownsample (registration.cpp:29-60): output order = libstdc++ unordered_map iteration order.
size_t orc_voxel_downsample(const float* xyz, size_t n, float voxel_size, float* out_xyz) {
    std::unordered_map<VoxelKey, std::vector<size_t>, VoxelKeyHash> grid;
    float inv = 1.0f / voxel_size;
    for (size_t i = 0; i < n; ++i) {
        VoxelKey key{static_cast<int>(std::floor(xyz[3 * i] * inv)), static_cast<int>(std::floor(xyz[3 * i + 1] * inv)),
                     static_cast<int>(std::floor(xyz[3 * i + 2] * inv))};
        grid[key].push_back(i);
    }
    size_t m = 0;
    for (auto& kv : grid) {
        float a0 = 0, a1 = 0, a2 = 0;
        for (size_t idx : kv.second) { a0 += xyz[3 * idx]; a1 += xyz[3 * idx + 1]; a2 += xyz[3 * idx + 2]; }
        float cnt = static_cast<float>(kv.second.size());
        out_xyz[3 * m] = a0 / cnt; out_xyz[3 * m + 1] = a1 / cnt; out_xyz[3 * m + 2] = a2 / cnt; ++m;
    }
    return m;
}

// estimateNormals (registration.cpp:105-130), k nearest incl. self, (d2, idx) pair ordering.
void orc_estimate_normals(const float* xyz, size_t n, int k, float* normals_out) {
    const P3* pts = reinterpret_cast<const P3*>(xyz);
    std::vector<std::pair<float, size_t>> dists;
    for (size_t i = 0; i < n; ++i) {
        dists.clear(); dists.reserve(n);
        for (size_t j = 0; j < n; ++j) dists.emplace_back(sqnorm(pts[j], pts[i]), j);
        int kk = std::min(k, (int)dists.size());
        std::partial_sort(dists.begin(), dists.begin() + kk, dists.end());
        float c0 = 0, c1 = 0, c2 = 0;
        for (int a = 0; a < kk; ++a) { const P3& p = pts[dists[a].second]; c0 += p.x; c1 += p.y; c2 += p.z; }
        float nf = static_cast<float>(kk);
        c0 /= nf; c1 /= nf; c2 /= nf;
        float cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int a = 0; a < kk; ++a) {
            const P3& p = pts[dists[a].second];
            float d[3] = {p.x - c0, p.y - c1, p.z - c2};
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[r][c] += d[r] * d[c];
        }
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[r][c] /= nf;
        float ev[3], Q[3][3];
        self_adjoint_eig3(cov, ev, Q);
        float nx = Q[0][0], ny = Q[1][0], nz = Q[2][0];
        // normals[i].dot(-points[i]) < 0  => flip
        if (red3(nx * (-pts[i].x), ny * (-pts[i].y), nz * (-pts[i].z)) < 0) { nx = -nx; ny = -ny; nz = -nz; }
        normals_out[3 * i] = nx; normals_out[3 * i + 1] = ny; normals_out[3 * i + 2] = nz;
    }
}

// computeFPFH (registration.cpp:133-201)
void orc_compute_fpfh(const float* xyz, const float* normals, size_t n, float radius, float* desc_out /* n x 33 */) {
    const P3* pts = reinterpret_cast<const P3*>(xyz);
    const P3* nrm = reinterpret_cast<const P3*>(normals);
    const float r2 = radius * radius;
    std::vector<std::vector<uint32_t>> nbrs(n);
    std::vector<std::pair<float, size_t>> dists;
    for (size_t i = 0; i < n; ++i) {
        dists.clear();
        for (size_t j = 0; j < n; ++j) { float d2 = sqnorm(pts[j], pts[i]); if (d2 <= r2) dists.emplace_back(d2, j); }
        std::sort(dists.begin(), dists.end());
        int cnt = std::min(100, (int)dists.size());
        nbrs[i].resize(cnt);
        for (int a = 0; a < cnt; ++a) nbrs[i][a] = (uint32_t)dists[a].second;
    }
    std::vector<std::array<float, 33>> spfh(n);
    for (size_t idx = 0; idx < n; ++idx) {
        std::array<float, 33> hist{};
        for (uint32_t ni : nbrs[idx]) {
            if (ni == idx) continue;
            float df[3] = {pts[ni].x - pts[idx].x, pts[ni].y - pts[idx].y, pts[ni].z - pts[idx].z};
            float dist = std::sqrt(red3(df[0] * df[0], df[1] * df[1], df[2] * df[2]));
            if (dist < 1e-8f) continue;
            float u[3] = {nrm[idx].x, nrm[idx].y, nrm[idx].z};
            float dn[3] = {df[0] / dist, df[1] / dist, df[2] / dist};
            float v[3] = {u[1] * dn[2] - u[2] * dn[1], u[2] * dn[0] - u[0] * dn[2], u[0] * dn[1] - u[1] * dn[0]};
            float w[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
            float nn[3] = {nrm[ni].x, nrm[ni].y, nrm[ni].z};
            float alpha = red3(v[0] * nn[0], v[1] * nn[1], v[2] * nn[2]);
            float phi = red3(u[0] * dn[0], u[1] * dn[1], u[2] * dn[2]);
            float theta = std::atan2(red3(w[0] * nn[0], w[1] * nn[1], w[2] * nn[2]), red3(u[0] * nn[0], u[1] * nn[1], u[2] * nn[2]));
            int bin_a = std::clamp(static_cast<int>((alpha + 1.0f) * 5.5f), 0, 10);
            int bin_p = std::clamp(static_cast<int>((phi + 1.0f) * 5.5f), 0, 10);
            int bin_t = std::clamp(static_cast<int>((theta / M_PI + 1.0f) * 5.5f), 0, 10);
            hist[bin_a] += 1.0f; hist[11 + bin_p] += 1.0f; hist[22 + bin_t] += 1.0f;
        }
        float sum = 0; for (float x : hist) sum += x;
        if (sum > 0) for (float& x : hist) x /= sum;
        spfh[idx] = hist;
    }
    for (size_t i = 0; i < n; ++i) {
        std::array<float, 33> f = spfh[i];
        for (uint32_t ni : nbrs[i]) {
            if (ni == i) continue;
            float dist = std::sqrt(sqnorm(pts[ni], pts[i]));
            if (dist < 1e-8f) continue;
            float weight = 1.0f / dist;
            for (int d = 0; d < 33; ++d) f[d] += weight * spfh[ni][d];
        }
        float sum = 0; for (float x : f) sum += x;
        if (sum > 0) for (float& x : f) x /= sum;
        std::memcpy(desc_out + 33 * i, f.data(), 33 * sizeof(float));
    }
}

}  // extern "C"
