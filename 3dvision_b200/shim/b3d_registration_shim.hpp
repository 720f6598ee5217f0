// b3d_registration_shim.hpp — C++ drop-in for the registration hot path of stojicnnnn/3DVision.
//
// Include this AFTER the reference's own headers (include/registration.hpp,
// include/gpu_registration.hpp); it only needs their types:
//   industry_picking::PointCloud          (registration.hpp:10-19)   three std::vector<Eigen::Vector3f>
//   industry_picking::FPFHFeatures        (registration.hpp:21-24)   std::vector<std::array<float,33>>
//   industry_picking::RegistrationResult  (registration.hpp:26-30)   Matrix4f + fitness + rmse
// and provides, in namespace industry_picking::b3d_shim, functions with exactly the signatures of
//   Registration::ransacRegistration      (registration.hpp:40-48)
//   Registration::icpRefine               (registration.hpp:50-57)
//   GPURegistration::icpRefine            (gpu_registration.hpp:10-16)
//   GPURegistration::isCudaAvailable      (gpu_registration.hpp:18)
//   Registration::voxelDownsample / estimateNormals / computeFPFH   (registration.hpp:34-38)
// implemented over the C-ABI of b3d.h (libb3d.so, sm_100a CUDA kernels).  b3d_registration_impl.cpp
// turns them into the out-of-line definitions of the reference's static member functions.
//
// Contract kept from the reference:
//   * inputs are const& to caller-owned host memory, result by value, no state visible to the caller;
//   * any failure (no CUDA device, CUDA error) throws std::runtime_error, so the orchestrator's
//     catch(...) / catch(std::exception) blocks (src/pipeline.cpp:114-121, 146-149) keep working;
//   * callable concurrently from the orchestrator's pool threads (src/pipeline.cpp:321-327): each
//     host thread gets its own b3d_ctx (stream + workspace) through a thread_local.
// Raw pointers are valid because std::vector<Eigen::Vector3f> is float[3n] and
// std::vector<std::array<float,33>> is float[33n] in memory (static_asserts below).
#pragma once

#include <array>
#include <atomic>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "b3d.h"

namespace industry_picking {
namespace b3d_shim {

static_assert(sizeof(Eigen::Vector3f) == 3 * sizeof(float), "Eigen::Vector3f must be 3 packed floats");
static_assert(sizeof(std::array<float, B3D_DESC_DIM>) == B3D_DESC_DIM * sizeof(float), "descriptor rows must be 33 packed floats");
static_assert(sizeof(Eigen::Matrix4f) == 16 * sizeof(float), "Eigen::Matrix4f must be 16 floats (column-major)");

// Which GPU a host thread's context lives on.  The reference's unit of parallelism is its worker pool
// (src/pipeline.cpp:16, 321-327; include/thread_pool.hpp): every pool thread that first calls into the shim is dealt
// the next usable sm_100 device round-robin, so a pool of num_threads workers spreads over all the GPUs of the box
// (instances are independent: no collective is needed, SURVEY.md 8e "batched multi-object").  device_index() >= 0 pins
// every new thread context to that device instead (set it before the first call, or export B3D_DEVICE).
inline int& device_index() {
    static int dev = [] { const char* e = std::getenv("B3D_DEVICE"); return e ? std::atoi(e) : -1; }();
    return dev;
}
inline int next_device() {
    if (device_index() >= 0) return device_index();
    static std::atomic<unsigned> dealt{0};
    const int n = b3d_device_count();
    return n > 0 ? (int)(dealt.fetch_add(1u) % (unsigned)n) : 0;
}

struct ThreadContext {
    b3d_ctx* ctx = nullptr;
    int device = -1;
    ~ThreadContext() { if (ctx) b3d_ctx_destroy(ctx); }
};
inline ThreadContext& thread_context() { thread_local ThreadContext tc; return tc; }

inline b3d_ctx* context() {
    ThreadContext& tc = thread_context();
    if (!tc.ctx) {
        tc.device = next_device();
        int rc = b3d_ctx_create(tc.device, &tc.ctx);
        if (rc != B3D_OK) throw std::runtime_error(std::string("b3d: ") + b3d_strerror(rc));
    }
    return tc.ctx;
}
inline int context_device() { context(); return thread_context().device; }

inline void check(b3d_ctx* ctx, int rc) {
    if (rc != B3D_OK) {
        std::string msg = b3d_last_error(ctx);
        throw std::runtime_error("b3d: " + (msg.empty() ? std::string(b3d_strerror(rc)) : msg));
    }
}

inline const float* xyz(const std::vector<Eigen::Vector3f>& v) { return v.empty() ? nullptr : reinterpret_cast<const float*>(v.data()); }

inline RegistrationResult ransacRegistration(const PointCloud& source, const PointCloud& target,
                                             const FPFHFeatures& source_features, const FPFHFeatures& target_features,
                                             float voxel_size, int max_iterations = 100000, float confidence = 0.999f) {
    if (source_features.descriptors.size() != source.points.size() || target_features.descriptors.size() != target.points.size())
        throw std::runtime_error("b3d: descriptor count does not match point count");
    b3d_ctx* ctx = context();
    RegistrationResult result;
    const float* sd = source_features.descriptors.empty() ? nullptr : source_features.descriptors.data()->data();
    const float* td = target_features.descriptors.empty() ? nullptr : target_features.descriptors.data()->data();
    check(ctx, b3d_ransac(ctx, xyz(source.points), source.points.size(), xyz(target.points), target.points.size(), sd, td,
                          voxel_size, max_iterations, confidence, result.transformation.data(), &result.fitness, &result.rmse, nullptr));
    return result;
}

inline RegistrationResult icpRefine(const PointCloud& source, const PointCloud& target, const Eigen::Matrix4f& initial_transform,
                                    float distance_threshold, int max_iterations = 200, bool point_to_plane = true) {
    b3d_ctx* ctx = context();
    RegistrationResult result;
    const bool has_normals = target.normals.size() == target.points.size() && !target.points.empty();   // PointCloud::hasNormals()
    check(ctx, b3d_icp(ctx, xyz(source.points), source.points.size(), xyz(target.points), has_normals ? xyz(target.normals) : nullptr,
                       target.points.size(), initial_transform.data(), distance_threshold, max_iterations, point_to_plane ? 1 : 0,
                       result.transformation.data(), &result.fitness, &result.rmse, nullptr));
    return result;
}

inline bool isCudaAvailable() { return b3d_cuda_available() != 0; }

// ---- stages that feed the hot path (registration.hpp:34-38; src/registration.cpp:29-60, 105-130, 133-201) ----
inline float* xyz(std::vector<Eigen::Vector3f>& v) { return v.empty() ? nullptr : reinterpret_cast<float*>(v.data()); }

inline PointCloud voxelDownsample(const PointCloud& cloud, float voxel_size) {
    b3d_ctx* ctx = context();
    PointCloud result;
    const size_t n = cloud.points.size();
    const bool colors = cloud.colors.size() == n && n > 0;                       // PointCloud::hasColors()
    result.points.resize(n);
    if (colors) result.colors.resize(n);
    size_t m = 0;
    check(ctx, b3d_voxel_downsample(ctx, xyz(cloud.points), n, colors ? xyz(cloud.colors) : nullptr, voxel_size,
                                    xyz(result.points), colors ? xyz(result.colors) : nullptr, n, &m));
    result.points.resize(m);
    if (colors) result.colors.resize(m);
    return result;                                                              // normals are dropped, as in the reference
}

inline void estimateNormals(PointCloud& cloud, int k = 30) {
    b3d_ctx* ctx = context();
    cloud.normals.resize(cloud.points.size());
    check(ctx, b3d_estimate_normals(ctx, xyz(cloud.points), cloud.points.size(), k, xyz(cloud.normals)));
}

inline FPFHFeatures computeFPFH(const PointCloud& cloud, float radius) {
    if (cloud.normals.size() != cloud.points.size()) throw std::runtime_error("b3d: computeFPFH needs one normal per point");
    b3d_ctx* ctx = context();
    FPFHFeatures features;
    features.descriptors.resize(cloud.points.size());
    check(ctx, b3d_compute_fpfh(ctx, xyz(cloud.points), xyz(cloud.normals), cloud.points.size(), radius,
                                features.descriptors.empty() ? nullptr : features.descriptors.data()->data()));
    return features;
}

// GPURegistration::icpRefine has no point_to_plane argument; the reference GPU path is always
// point-to-plane (src/gpu_impl.cpp:141-260).  Parity target is the CPU semantics.
inline RegistrationResult gpuIcpRefine(const PointCloud& source, const PointCloud& target, const Eigen::Matrix4f& initial_transform,
                                       float distance_threshold, int max_iterations = 200) {
    if (!isCudaAvailable()) throw std::runtime_error("CUDA not available");      // src/gpu_impl.cpp:258
    return icpRefine(source, target, initial_transform, distance_threshold, max_iterations, true);
}

// ---- pose post-processing of Pipeline::processInstance / Pipeline::filterDuplicates (src/pipeline.cpp:136-137, 153-180) ----
// T_world_object = camera_extrinsics * refined.transformation.inverse()
inline Eigen::Matrix4f worldPose(const Eigen::Matrix4f& camera_extrinsics, const Eigen::Matrix4f& refined_transformation) {
    b3d_ctx* ctx = context();
    Eigen::Matrix4f out;
    check(ctx, b3d_world_poses(ctx, refined_transformation.data(), 1, camera_extrinsics.data(), out.data()));
    return out;
}
inline std::vector<Eigen::Matrix4f> filterDuplicates(const std::vector<Eigen::Matrix4f>& waypoints, float min_distance) {
    b3d_ctx* ctx = context();
    std::vector<Eigen::Matrix4f> filtered(waypoints.size());
    size_t kept = 0;
    check(ctx, b3d_filter_duplicates(ctx, waypoints.empty() ? nullptr : waypoints.front().data(), waypoints.size(), min_distance,
                                     filtered.empty() ? nullptr : filtered.front().data(), &kept));
    filtered.resize(kept);
    return filtered;
}

}  // namespace b3d_shim
}  // namespace industry_picking
