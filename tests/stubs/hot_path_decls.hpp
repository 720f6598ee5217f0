// Test-only declarations of the boundary the shim implements, written from SURVEY.md §8(b): the
// three container types and the four static entry points of the reference's registration.hpp /
// gpu_registration.hpp (hot-path members only).  Used where /root/reference is not available
// (the GPU box); where it is, tests compile against the reference's own headers instead.
#pragma once
#include <Eigen/Core>
#include <array>
#include <vector>
namespace industry_picking {
struct PointCloud {
    std::vector<Eigen::Vector3f> points, normals, colors;
    size_t size() const { return points.size(); }
    bool hasNormals() const { return normals.size() == points.size(); }
};
struct FPFHFeatures { std::vector<std::array<float, 33>> descriptors; };
struct RegistrationResult { Eigen::Matrix4f transformation = Eigen::Matrix4f::Identity(); float fitness = 0.0f; float rmse = 0.0f; };
class Registration {
public:
    static PointCloud voxelDownsample(const PointCloud& cloud, float voxel_size);
    static void estimateNormals(PointCloud& cloud, int k = 30);
    static FPFHFeatures computeFPFH(const PointCloud& cloud, float radius);
    static RegistrationResult ransacRegistration(const PointCloud&, const PointCloud&, const FPFHFeatures&, const FPFHFeatures&,
                                                 float voxel_size, int max_iterations = 100000, float confidence = 0.999f);
    static RegistrationResult icpRefine(const PointCloud&, const PointCloud&, const Eigen::Matrix4f&, float distance_threshold,
                                        int max_iterations = 200, bool point_to_plane = true);
};
class GPURegistration {
public:
    static RegistrationResult icpRefine(const PointCloud&, const PointCloud&, const Eigen::Matrix4f&, float distance_threshold,
                                        int max_iterations = 200);
    static bool isCudaAvailable();
};
}  // namespace industry_picking
