// b3d_registration_impl.cpp — out-of-line definitions of the reference's hot-path entry points on
// top of libb3d.so.  Build it INSTEAD of the bodies it replaces:
//   Registration::ransacRegistration   src/registration.cpp:204-295
//   Registration::icpRefine            src/registration.cpp:297-414
//   GPURegistration::icpRefine         src/gpu_impl.cpp:141-260
//   GPURegistration::isCudaAvailable   src/gpu_impl.cpp
//   (with -DB3D_SHIM_FEATURE_STAGES) Registration::voxelDownsample / estimateNormals / computeFPFH
// (see INTEGRATION.md for the three-line patch to the reference's CMakeLists.txt / sources).
// B3D_REFERENCE_HEADERS lets a test substitute declarations for the reference's headers.
#ifdef B3D_REFERENCE_HEADERS
#include B3D_REFERENCE_HEADERS
#else
#include "registration.hpp"
#include "gpu_registration.hpp"
#endif
#include "b3d_registration_shim.hpp"

namespace industry_picking {

RegistrationResult Registration::ransacRegistration(const PointCloud& source, const PointCloud& target,
                                                    const FPFHFeatures& source_features, const FPFHFeatures& target_features,
                                                    float voxel_size, int max_iterations, float confidence) {
    return b3d_shim::ransacRegistration(source, target, source_features, target_features, voxel_size, max_iterations, confidence);
}

RegistrationResult Registration::icpRefine(const PointCloud& source, const PointCloud& target, const Eigen::Matrix4f& initial_transform,
                                           float distance_threshold, int max_iterations, bool point_to_plane) {
    return b3d_shim::icpRefine(source, target, initial_transform, distance_threshold, max_iterations, point_to_plane);
}

RegistrationResult GPURegistration::icpRefine(const PointCloud& source, const PointCloud& target, const Eigen::Matrix4f& initial_transform,
                                              float distance_threshold, int max_iterations) {
    return b3d_shim::gpuIcpRefine(source, target, initial_transform, distance_threshold, max_iterations);
}

bool GPURegistration::isCudaAvailable() { return b3d_shim::isCudaAvailable(); }

// The stages that feed the hot path.  Define B3D_SHIM_FEATURE_STAGES when the reference's own bodies
// (src/registration.cpp:29-60, 105-130, 133-201) are removed as well; otherwise they stay the reference's.
#ifdef B3D_SHIM_FEATURE_STAGES
PointCloud Registration::voxelDownsample(const PointCloud& cloud, float voxel_size) { return b3d_shim::voxelDownsample(cloud, voxel_size); }
void Registration::estimateNormals(PointCloud& cloud, int k) { b3d_shim::estimateNormals(cloud, k); }
FPFHFeatures Registration::computeFPFH(const PointCloud& cloud, float radius) { return b3d_shim::computeFPFH(cloud, radius); }
#endif

}  // namespace industry_picking
