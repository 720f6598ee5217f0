// oracle/ref_shim.cpp — C surface over the UNMODIFIED reference (test infrastructure only).
// Built only by `make -C oracle ref EIGEN_DIR=...`, together with /root/reference/src/registration.cpp compiled where it
// lies (no reference source is copied into this repo).  Exposes the reference's own Registration:: functions
// (include/registration.hpp:32-59) under the same argument conventions as the restated oracle's entry points
// (registration_oracle.cpp, pipeline_inputs.cpp), so oracle/diff_ref.py can run both on the same inputs and turn
// "parity unpinned" into a measured statement.  Needs Eigen >= 3.3 headers, which this image does not have.
#include "registration.hpp"

#include <cstdint>
#include <cstring>

using namespace industry_picking;

static PointCloud cloud_from(const float* xyz, size_t n, const float* normals = nullptr) {
    PointCloud c;
    c.points.resize(n);
    if (n) std::memcpy(c.points.data(), xyz, sizeof(float) * 3 * n);           // vector<Vector3f> is float[3n]
    if (normals) { c.normals.resize(n); if (n) std::memcpy(c.normals.data(), normals, sizeof(float) * 3 * n); }
    return c;
}
static FPFHFeatures feats_from(const float* desc, size_t n) {
    FPFHFeatures f;
    f.descriptors.resize(n);
    if (n) std::memcpy(f.descriptors.data(), desc, sizeof(float) * 33 * n);     // vector<array<float,33>> is float[33n]
    return f;
}
static void result_out(const RegistrationResult& r, float* T_colmajor, float* fitness, float* rmse) {
    std::memcpy(T_colmajor, r.transformation.data(), sizeof(float) * 16);      // Matrix4f is column-major
    *fitness = r.fitness; *rmse = r.rmse;
}

extern "C" {

// Registration::ransacRegistration, src/registration.cpp:204-295
int ref_ransac_registration(const float* src, size_t n_src, const float* tgt, size_t n_tgt, const float* src_desc, const float* tgt_desc,
                            float voxel_size, int max_iterations, float confidence, float* out_T, float* out_fitness, float* out_rmse) {
    const RegistrationResult r = Registration::ransacRegistration(cloud_from(src, n_src), cloud_from(tgt, n_tgt), feats_from(src_desc, n_src),
                                                                  feats_from(tgt_desc, n_tgt), voxel_size, max_iterations, confidence);
    result_out(r, out_T, out_fitness, out_rmse);
    return 0;
}

// Registration::icpRefine, src/registration.cpp:297-414 (tgt_normals may be null: target.hasNormals() == false)
int ref_icp(const float* src, size_t n_src, const float* tgt, const float* tgt_normals, size_t n_tgt, const float* T0_colmajor,
            float distance_threshold, int max_iterations, int point_to_plane, float* out_T, float* out_fitness, float* out_rmse) {
    Eigen::Matrix4f T0;
    std::memcpy(T0.data(), T0_colmajor, sizeof(float) * 16);
    const RegistrationResult r = Registration::icpRefine(cloud_from(src, n_src), cloud_from(tgt, n_tgt, tgt_normals), T0, distance_threshold,
                                                         max_iterations, point_to_plane != 0);
    result_out(r, out_T, out_fitness, out_rmse);
    return 0;
}

// Registration::voxelDownsample, src/registration.cpp:15-60; returns the number of output points (out_xyz holds n at most)
size_t ref_voxel_downsample(const float* xyz, size_t n, float voxel_size, float* out_xyz) {
    const PointCloud d = Registration::voxelDownsample(cloud_from(xyz, n), voxel_size);
    if (!d.points.empty()) std::memcpy(out_xyz, d.points.data(), sizeof(float) * 3 * d.points.size());
    return d.points.size();
}

// Registration::estimateNormals, src/registration.cpp:105-130
void ref_estimate_normals(const float* xyz, size_t n, int k, float* normals_out) {
    PointCloud c = cloud_from(xyz, n);
    Registration::estimateNormals(c, k);
    if (n) std::memcpy(normals_out, c.normals.data(), sizeof(float) * 3 * n);
}

// Registration::computeFPFH, src/registration.cpp:133-201
void ref_compute_fpfh(const float* xyz, const float* normals, size_t n, float radius, float* desc_out) {
    const FPFHFeatures f = Registration::computeFPFH(cloud_from(xyz, n, normals), radius);
    if (n) std::memcpy(desc_out, f.descriptors.data(), sizeof(float) * 33 * n);
}

}  // extern "C"
