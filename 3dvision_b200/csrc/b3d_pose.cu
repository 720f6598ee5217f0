// b3d_pose.cu — pose post-processing of the orchestrator around the registration hot path:
//   T_world_object = camera_extrinsics * refined.transformation.inverse()      src/pipeline.cpp:136-137
//   Pipeline::filterDuplicates                                                  src/pipeline.cpp:153-180
// Tiny, latency-only work (<= a few hundred poses per run); it lives on the device so that a batch of refined poses
// can be turned into the orchestrator's waypoints without a separate host implementation of the reference's 4x4
// arithmetic: Matrix4f::inverse() follows Eigen 3.4's SSE kernel (2x2-block cofactor scheme, LU/arch/InverseSize4.h)
// lane for lane, the 4x4 product its column-major packet order (b3d_linalg.cuh).  Compile with --fmad=false.
#include "b3d_common.cuh"
#include "b3d_linalg.cuh"

namespace b3d {

// One 4-lane SSE operation of Eigen's kernel per statement; shuffles are index maps.
struct Lanes {
    float v[4];
    __device__ __forceinline__ float operator[](int i) const { return v[i]; }
};
__device__ __forceinline__ Lanes shuf(const Lanes& a, const Lanes& b, int p, int q, int r, int s) { return Lanes{{a[p], a[q], b[r], b[s]}}; }
__device__ __forceinline__ Lanes lo_pair(const Lanes& a, const Lanes& b) { return Lanes{{a[0], a[1], b[0], b[1]}}; }     // movelh
__device__ __forceinline__ Lanes hi_pair(const Lanes& a, const Lanes& b) { return Lanes{{b[2], b[3], a[2], a[3]}}; }     // movehl(a, b)
__device__ __forceinline__ Lanes splat(const Lanes& a, int p) { return Lanes{{a[p], a[p], a[p], a[p]}}; }
__device__ __forceinline__ Lanes operator*(const Lanes& a, const Lanes& b) { return Lanes{{a[0] * b[0], a[1] * b[1], a[2] * b[2], a[3] * b[3]}}; }
__device__ __forceinline__ Lanes operator+(const Lanes& a, const Lanes& b) { return Lanes{{a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]}}; }
__device__ __forceinline__ Lanes operator-(const Lanes& a, const Lanes& b) { return Lanes{{a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]}}; }

// Eigen::Matrix4f::inverse(), column-major in and out
__device__ void mat4_inverse_eigen(const float* __restrict__ M, float* __restrict__ out) {
    const Lanes c0{{M[0], M[1], M[2], M[3]}}, c1{{M[4], M[5], M[6], M[7]}}, c2{{M[8], M[9], M[10], M[11]}}, c3{{M[12], M[13], M[14], M[15]}};
    const Lanes A = lo_pair(c0, c1), B = hi_pair(c1, c0), C = lo_pair(c2, c3), D = hi_pair(c3, c2);
    const Lanes AB = shuf(A, A, 3, 3, 0, 0) * B - shuf(A, A, 1, 1, 2, 2) * shuf(B, B, 2, 3, 0, 1);          // A# B
    const Lanes DC = shuf(D, D, 3, 3, 0, 0) * C - shuf(D, D, 1, 1, 2, 2) * shuf(C, C, 2, 3, 0, 1);          // D# C
    Lanes dA = shuf(A, A, 3, 3, 1, 1) * A; dA = dA - hi_pair(dA, dA);
    Lanes dB = shuf(B, B, 3, 3, 1, 1) * B; dB = dB - hi_pair(dB, dB);
    Lanes dC = shuf(C, C, 3, 3, 1, 1) * C; dC = dC - hi_pair(dC, dC);
    Lanes dD = shuf(D, D, 3, 3, 1, 1) * D; dD = dD - hi_pair(dD, dD);
    Lanes d = shuf(DC, DC, 0, 2, 1, 3) * AB;
    d = d + hi_pair(d, d);
    d = d + shuf(d, d, 1, 0, 0, 0);
    const Lanes det = splat((dA * dD + dB * dC) - d, 0);                  // |A||D| + |B||C| - trace(A# B D# C)
    const float r = 1.0f / det[0];
    const Lanes rd{{r, -r, -r, r}};
    Lanes iD = shuf(C, C, 0, 0, 2, 2) * lo_pair(AB, AB);
    iD = iD + shuf(C, C, 1, 1, 3, 3) * hi_pair(AB, AB);
    iD = D * splat(dA, 0) - iD;
    Lanes iA = shuf(B, B, 0, 0, 2, 2) * lo_pair(DC, DC);
    iA = iA + shuf(B, B, 1, 1, 3, 3) * hi_pair(DC, DC);
    iA = A * splat(dD, 0) - iA;
    Lanes iB = D * shuf(AB, AB, 3, 0, 3, 0);
    iB = iB - shuf(D, D, 1, 0, 3, 2) * shuf(AB, AB, 2, 1, 2, 1);
    iB = C * splat(dB, 0) - iB;
    Lanes iC = A * shuf(DC, DC, 3, 0, 3, 0);
    iC = iC - shuf(A, A, 1, 0, 3, 2) * shuf(DC, DC, 2, 1, 2, 1);
    iC = B * splat(dC, 0) - iC;
    iA = iA * rd; iB = iB * rd; iC = iC * rd; iD = iD * rd;
    const Lanes o0 = shuf(iA, iB, 3, 1, 3, 1), o1 = shuf(iA, iB, 2, 0, 2, 0), o2 = shuf(iC, iD, 3, 1, 3, 1), o3 = shuf(iC, iD, 2, 0, 2, 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) { out[i] = o0[i]; out[4 + i] = o1[i]; out[8 + i] = o2[i]; out[12 + i] = o3[i]; }
}

__global__ void world_pose_kernel(const float* __restrict__ refined, unsigned n, const float* __restrict__ extrinsics_or_null, float* __restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float inv[16];
    mat4_inverse_eigen(refined + 16 * (size_t)i, inv);                    // T_camera_object, pipeline.cpp:136
    if (extrinsics_or_null) {
        float e[16], w[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = extrinsics_or_null[k];
        mat4_mul(e, inv, w);                                              // T_world_object, pipeline.cpp:137
#pragma unroll
        for (int k = 0; k < 16; ++k) out[16 * (size_t)i + k] = w[k];
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) out[16 * (size_t)i + k] = inv[k];
    }
}

// pipeline.cpp:153-180 is a sequential, order-dependent scan (a pose is compared with the poses kept so far and may replace
// one of them in place), so one thread walks the waypoints; `slot` holds, per kept pose, the index of the waypoint in it.
__global__ void filter_duplicates_kernel(const float* __restrict__ poses, unsigned n, float min_distance, unsigned* __restrict__ slot,
                                         float* __restrict__ out, unsigned* __restrict__ out_n) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    auto norm3 = [](float a0, float a1, float a2) { return sqrtf(a0 * a0 + (a1 * a1 + a2 * a2)); };       // Eigen's 3-term redux order
    unsigned kept = 0;
    for (unsigned w = 0; w < n; ++w) {
        const float* wp = poses + 16 * (size_t)w;
        bool dup = false;
        for (unsigned i = 0; i < kept; ++i) {
            const float* f = poses + 16 * (size_t)slot[i];
            const float dist = norm3(wp[12] - f[12], wp[13] - f[13], wp[14] - f[14]);
            if (dist < min_distance) {
                dup = true;
                if (norm3(wp[12], wp[13], wp[14]) < norm3(f[12], f[13], f[14])) slot[i] = w;               // replace, :168-170
                break;
            }
        }
        if (!dup) slot[kept++] = w;
    }
    for (unsigned i = 0; i < kept; ++i)
        for (int k = 0; k < 16; ++k) out[16 * (size_t)i + k] = poses[16 * (size_t)slot[i] + k];
    *out_n = kept;
}

int world_poses_impl(b3d_ctx* c, const float* refined, size_t n, const float* extrinsics_or_null, float* out) {
    if (n == 0) return B3D_OK;
    if (n > (1u << 24)) return fail(c, B3D_ERR_INVALID, "world_poses: too many poses");
    B3D_CUDA(c, c->stage_a.ensure(sizeof(float) * 16 * (n + 1)));
    B3D_CUDA(c, c->stage_b.ensure(sizeof(float) * 16 * n));
    float* d_in = c->stage_a.as<float>();
    float* d_ext = d_in + 16 * n;
    B3D_CUDA(c, cudaMemcpyAsync(d_in, refined, sizeof(float) * 16 * n, cudaMemcpyHostToDevice, c->stream));
    if (extrinsics_or_null) B3D_CUDA(c, cudaMemcpyAsync(d_ext, extrinsics_or_null, sizeof(float) * 16, cudaMemcpyHostToDevice, c->stream));
    world_pose_kernel<<<div_up((long long)n, 64), 64, 0, c->stream>>>(d_in, (unsigned)n, extrinsics_or_null ? d_ext : nullptr, c->stage_b.as<float>());
    B3D_LAUNCHED(c);
    B3D_CUDA(c, cudaMemcpyAsync(out, c->stage_b.p, sizeof(float) * 16 * n, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    return B3D_OK;
}

// Diagnostic (b3d_euler_rotations): the ICP update's rotation, AngleAxis(x, X) * AngleAxis(y, Y) * AngleAxis(z, Z) -> matrix
// (src/registration.cpp:369-371), for caller-supplied angles - the device build of b3d_libm.cuh's sinf / cosf and of the
// quaternion product, checkable against the oracle on every argument range.
__global__ void euler_rotations_kernel(const float* __restrict__ angles, unsigned n, float* __restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Mat3 R;
    euler_xyz_to_matrix(angles[3 * (size_t)i], angles[3 * (size_t)i + 1], angles[3 * (size_t)i + 2], R);
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) out[9 * (size_t)i + 3 * r + cc] = R(r, cc);
}
int euler_rotations_impl(b3d_ctx* c, const float* angles, size_t n, float* out) {
    if (n == 0) return B3D_OK;
    if (n > (1u << 26)) return fail(c, B3D_ERR_INVALID, "euler_rotations: too many triples");
    B3D_CUDA(c, c->stage_a.ensure(sizeof(float) * 3 * n));
    B3D_CUDA(c, c->stage_b.ensure(sizeof(float) * 9 * n));
    B3D_CUDA(c, cudaMemcpyAsync(c->stage_a.p, angles, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    euler_rotations_kernel<<<div_up((long long)n, 128), 128, 0, c->stream>>>(c->stage_a.as<float>(), (unsigned)n, c->stage_b.as<float>());
    B3D_LAUNCHED(c);
    B3D_CUDA(c, cudaMemcpyAsync(out, c->stage_b.p, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    return B3D_OK;
}

int filter_duplicates_impl(b3d_ctx* c, const float* poses, size_t n, float min_distance, float* out, size_t* out_n) {
    *out_n = 0;
    if (n == 0) return B3D_OK;
    if (n > (1u << 20)) return fail(c, B3D_ERR_INVALID, "filter_duplicates: too many poses");
    B3D_CUDA(c, c->stage_a.ensure(sizeof(float) * 16 * n));
    B3D_CUDA(c, c->stage_b.ensure(sizeof(float) * 16 * n));
    B3D_CUDA(c, c->stage_c.ensure(sizeof(unsigned) * (n + 1)));
    B3D_CUDA(c, cudaMemcpyAsync(c->stage_a.p, poses, sizeof(float) * 16 * n, cudaMemcpyHostToDevice, c->stream));
    unsigned* slot = c->stage_c.as<unsigned>();
    filter_duplicates_kernel<<<1, 32, 0, c->stream>>>(c->stage_a.as<float>(), (unsigned)n, min_distance, slot, c->stage_b.as<float>(), slot + n);
    B3D_LAUNCHED(c);
    unsigned kept = 0;
    B3D_CUDA(c, cudaMemcpyAsync(&kept, slot + n, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    if (kept) B3D_CUDA(c, cudaMemcpyAsync(out, c->stage_b.p, sizeof(float) * 16 * kept, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    *out_n = kept;
    return B3D_OK;
}

}  // namespace b3d
