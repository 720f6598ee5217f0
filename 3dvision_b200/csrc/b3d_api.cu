// b3d_api.cu — the extern "C" surface declared in include/b3d.h.
// Thin: argument checks, H2D staging, stage sequencing.  All arithmetic of the path
// runs in the CUDA kernels of b3d_match.cu / b3d_ransac.cu / b3d_icp.cu.
#include "b3d_common.cuh"
#include <atomic>
#include <string.h>
#include <new>

namespace b3d {

__global__ void xyz_to_float4_kernel(const float* __restrict__ xyz, unsigned n, float4* __restrict__ out) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], 0.0f);
}

int xyz_to_float4(b3d_ctx* c, const float* xyz_dev, size_t n, float4* out) {
    if (n == 0) return B3D_OK;
    xyz_to_float4_kernel<<<grid_for((long long)n, 256, 8), 256, 0, c->stream>>>(xyz_dev, (unsigned)n, out);
    B3D_LAUNCHED(c);
    return B3D_OK;
}

// Roofline denominator for the scoring kernel: sustained rate of *separate* FMUL and FADD
// instructions (the reference arithmetic is un-fused), 8 independent chains per thread.
__global__ void __launch_bounds__(256) fp32_issue_rate_kernel(float* out, float a, float b, int iters) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { x[k] = x[k] * a; x[k] = x[k] + b; }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    if (s == 123.456f) out[0] = s;          // keeps the chains alive; practically never true
}

static int upload_cloud(b3d_ctx* c, const float* xyz, size_t n, int on_device, DevBuf& stage, DevBuf& dst) {
    B3D_CUDA(c, dst.ensure(sizeof(float4) * (n ? n : 1)));
    if (n == 0) return B3D_OK;
    const float* dev = xyz;
    if (!on_device) {
        B3D_CUDA(c, stage.ensure(sizeof(float) * 3 * n));
        B3D_CUDA(c, cudaMemcpyAsync(stage.p, xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
        dev = stage.as<float>();
    }
    return xyz_to_float4(c, dev, n, dst.as<float4>());
}

static bool device_usable() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return false; }
    return true;
}

}  // namespace b3d

using namespace b3d;

extern "C" {

static bool device_is_sm100(int dev) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return false; }
    return major == 10;                     // the library carries sm_100a code only
}

int b3d_device_count(void) {
    // leading run of sm_100 devices: contexts address devices by CUDA ordinal, so a foreign device in between ends the run
    static std::atomic<int> cached{-1};
    const int have = cached.load(std::memory_order_relaxed);
    if (have > 0) return have;
    if (!device_usable()) return 0;
    int n = 0, usable = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    while (usable < n && device_is_sm100(usable)) ++usable;
    if (usable > 0) cached.store(usable, std::memory_order_relaxed);
    return usable;
}

int b3d_cuda_available(void) {
    // The orchestrator asks before every instance (src/pipeline.cpp:107) and cudaGetDeviceProperties costs milliseconds,
    // so a positive answer is remembered for the process; a negative one is re-probed (a device may appear later).
    return b3d_device_count() > 0 ? 1 : 0;
}

const char* b3d_strerror(int status) {
    switch (status) {
        case B3D_OK: return "ok";
        case B3D_ERR_NO_DEVICE: return "no usable CUDA device (sm_100a required; there is no CPU fallback)";
        case B3D_ERR_CUDA: return "CUDA error";
        case B3D_ERR_INVALID: return "invalid argument";
        case B3D_ERR_ALLOC: return "allocation failed";
        case B3D_ERR_STATE: return "call out of order";
        case B3D_ERR_RNG_WINDOW: return "RNG acceptance window exhausted";
        default: return "unknown status";
    }
}

const char* b3d_last_error(const b3d_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int b3d_ctx_create(int device, b3d_ctx** out) {
    if (!out) return B3D_ERR_INVALID;
    *out = nullptr;
    if (!device_usable()) return B3D_ERR_NO_DEVICE;
    if (device < 0 || !device_is_sm100(device)) return B3D_ERR_NO_DEVICE;       // no such device, or not a Blackwell one
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return B3D_ERR_NO_DEVICE; }
    b3d_ctx* c = new (std::nothrow) b3d_ctx();
    if (!c) return B3D_ERR_ALLOC;
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return B3D_ERR_CUDA; }
    c->stream = c->own_stream;
    {   // keep freed workspace in the device's default pool instead of returning it to the driver at every sync
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else cudaGetLastError();
    }
    alloc_stream_valid() = false;
    if (c->state.ensure(sizeof(DeviceState)) != cudaSuccess ||
        cudaMallocHost(&c->h_state, sizeof(DeviceState)) != cudaSuccess) { b3d_ctx_destroy(c); return B3D_ERR_ALLOC; }
    cudaMemsetAsync(c->state.p, 0, sizeof(DeviceState), c->stream);
    memset(c->h_state, 0, sizeof(DeviceState));
    for (int s = 0; s < kStages; ++s) { cudaEventCreate(&c->ev_start[s]); cudaEventCreate(&c->ev_stop[s]); }
    *out = c;
    return B3D_OK;
}

void b3d_ctx_destroy(b3d_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    alloc_stream_valid() = false;
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->own_stream && c->own_stream != c->stream) cudaStreamSynchronize(c->own_stream);
    comm_destroy_impl(c);
    DevBuf* bufs[] = {&c->stage_a, &c->stage_b, &c->stage_c, &c->src4, &c->tgt4, &c->nrm4, &c->sdesc, &c->tdesc, &c->corr, &c->raw,
                      &c->draws, &c->scan_tmp, &c->hyp, &c->counts, &c->pairs, &c->seqsum, &c->grid_slots, &c->grid_cursor,
                      &c->grid_pts, &c->grid_nrm, &c->pt_slot, &c->pt_rank, &c->partials, &c->nn_idx, &c->nn_d2, &c->state,
                      &c->seq_rec, &c->seq_match, &c->seq_P, &c->seq_Q, &c->seq_N, &c->ess_terms, &c->ess_bsum, &c->ess_guess, &c->ess_summ, &c->fine_slots, &c->fine_pts, &c->nbh_slot27, &c->nbh_cursor, &c->icp_cache, &c->icp_cache_idx, &c->bail_list_a, &c->bail_list_b, &c->bail_state, &c->src_slots, &c->src_sorted, &c->src_slot, &c->src_rank,
                      &c->tc_a_tiles, &c->tc_b_tiles, &c->tc_norm2, &c->tc_best, &c->tc_aux, &c->dist_keys};
    for (DevBuf* b : bufs) b->release();
    for (DevBuf& b : c->fbuf) b.release();
    if (c->h_state) cudaFreeHost(c->h_state);
    for (int s = 0; s < kStages; ++s) { if (c->ev_start[s]) cudaEventDestroy(c->ev_start[s]); if (c->ev_stop[s]) cudaEventDestroy(c->ev_stop[s]); }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int b3d_ctx_set_stream(b3d_ctx* c, void* cuda_stream) {
    if (!c) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return B3D_OK;
}

uint64_t b3d_kernel_launches(const b3d_ctx* c) { return c ? c->launches : 0; }

float b3d_stage_ms(const b3d_ctx* c, int stage) {
    if (!c || stage < 0 || stage >= kStages || !c->ev_valid[stage]) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(c->ev_stop[stage]) != cudaSuccess) { cudaGetLastError(); return -1.0f; }
    if (cudaEventElapsedTime(&ms, c->ev_start[stage], c->ev_stop[stage]) != cudaSuccess) { cudaGetLastError(); return -1.0f; }
    return ms;
}

int b3d_measure_fp32_rate(b3d_ctx* c, double* out_ops_per_second) {
    if (!c || !out_ops_per_second) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    B3D_CUDA(c, c->seqsum.ensure(64));
    const int iters = 1 << 14, blocks = kNumSMs * 8, threads = 256;
    cudaEvent_t e0, e1;
    B3D_CUDA(c, cudaEventCreate(&e0)); B3D_CUDA(c, cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, c->stream);
        fp32_issue_rate_kernel<<<blocks, threads, 0, c->stream>>>(c->seqsum.as<float>(), 0.999999f, 1e-7f, iters);
        c->launches++;
        cudaEventRecord(e1, c->stream);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return fail_cuda(c, e, "fp32 rate kernel", __FILE__, __LINE__); }
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *out_ops_per_second = (double)blocks * threads * (double)iters * 16.0 / ((double)best_ms * 1e-3);
    return B3D_OK;
}

// ---- staged API -------------------------------------------------------------------------

int b3d_set_clouds(b3d_ctx* c, const float* src_xyz, size_t n_src, const float* tgt_xyz, const float* tgt_normals, size_t n_tgt, int on_device) {
    if (!c) return B3D_ERR_INVALID;
    if ((n_src && !src_xyz) || (n_tgt && !tgt_xyz)) return fail(c, B3D_ERR_INVALID, "set_clouds: null cloud pointer");
    if (n_src >= 0xFFFFFFFFull || n_tgt >= 0xFFFFFFFFull) return fail(c, B3D_ERR_INVALID, "set_clouds: more than 2^32-2 points");
    B3D_CUDA(c, enter(c));
    // new clouds invalidate everything sized by the old ones: descriptors too (they may be a caller's device pointer of the
    // old n_src / n_tgt rows), so a stale set_features cannot be matched against the new clouds
    c->have_clouds = false; c->have_feats = false; c->have_corr = false; c->prepared = false; c->scored = false; c->model_ready = false;
    c->sdesc_p = nullptr; c->tdesc_p = nullptr;
    int rc = upload_cloud(c, src_xyz, n_src, on_device, c->stage_a, c->src4); if (rc) return rc;
    rc = upload_cloud(c, tgt_xyz, n_tgt, on_device, c->stage_b, c->tgt4); if (rc) return rc;
    c->has_normals = tgt_normals != nullptr;
    if (tgt_normals) { rc = upload_cloud(c, tgt_normals, n_tgt, on_device, c->stage_c, c->nrm4); if (rc) return rc; }
    c->n_src = n_src; c->n_tgt = n_tgt;
    c->have_clouds = true;
    return B3D_OK;
}

int b3d_set_features(b3d_ctx* c, const float* src_desc, const float* tgt_desc, int on_device) {
    if (!c) return B3D_ERR_INVALID;
    if (!c->have_clouds) return fail(c, B3D_ERR_STATE, "set_features: call set_clouds first");
    if ((c->n_src && !src_desc) || (c->n_tgt && !tgt_desc)) return fail(c, B3D_ERR_INVALID, "set_features: null descriptor pointer");
    B3D_CUDA(c, enter(c));
    c->have_feats = false;
    c->model_ready = false;                  // c->tdesc is about to be overwritten: the resident model's descriptors are gone
    if (on_device) { c->sdesc_p = src_desc; c->tdesc_p = tgt_desc; }
    else {
        const size_t sb = sizeof(float) * kDescDim * c->n_src, tb = sizeof(float) * kDescDim * c->n_tgt;
        B3D_CUDA(c, c->sdesc.ensure(sb ? sb : 4)); B3D_CUDA(c, c->tdesc.ensure(tb ? tb : 4));
        if (sb) B3D_CUDA(c, cudaMemcpyAsync(c->sdesc.p, src_desc, sb, cudaMemcpyHostToDevice, c->stream));
        if (tb) B3D_CUDA(c, cudaMemcpyAsync(c->tdesc.p, tgt_desc, tb, cudaMemcpyHostToDevice, c->stream));
        c->sdesc_p = c->sdesc.as<float>(); c->tdesc_p = c->tdesc.as<float>();
    }
    c->have_feats = true;
    return B3D_OK;
}

int b3d_set_match_mode(b3d_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 2) return B3D_ERR_INVALID;
    c->match_mode = mode;
    return B3D_OK;
}

int b3d_score_recounts(b3d_ctx* c, uint64_t* out) {
    if (!c || !out) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    unsigned long long v = 0;
    B3D_CUDA(c, cudaMemcpyAsync(&v, &c->state.as<DeviceState>()->score_recounts, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = v;
    return B3D_OK;
}

int b3d_icp_exact_sum_stats(b3d_ctx* c, uint32_t out[128]) {
    if (!c || !out) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    B3D_CUDA(c, cudaMemcpyAsync(out, c->state.as<DeviceState>()->ess_stats, sizeof(uint32_t) * 128, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    return B3D_OK;
}

int b3d_icp_exact_sum_dump(b3d_ctx* c, float* terms_out, size_t capacity_floats, size_t* out_stride, float sums_out[32]) {
    if (!c || !out_stride || !sums_out) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    const size_t stride = (c->n_src + 4095) / 4096 * 4096;             // ess::padded_terms: whole 4096-term chunks
    *out_stride = stride;
    B3D_CUDA(c, cudaMemcpyAsync(sums_out, c->state.as<DeviceState>()->ess_sums, sizeof(float) * 32, cudaMemcpyDeviceToHost, c->stream));
    if (terms_out) {
        const size_t have = c->ess_terms.cap / sizeof(float);
        const size_t n = capacity_floats < have ? capacity_floats : have;
        B3D_CUDA(c, cudaMemcpyAsync(terms_out, c->ess_terms.as<float>(), sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    }
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    return B3D_OK;
}

int b3d_euler_rotations(b3d_ctx* c, const float* angles_xyz, size_t n, float* out_R_rowmajor) {
    if (!c || (n && (!angles_xyz || !out_R_rowmajor))) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return euler_rotations_impl(c, angles_xyz, n, out_R_rowmajor);
}

int b3d_sequential_sum(b3d_ctx* c, const float* terms, size_t n, float* out_sum, uint32_t out_stats[3]) {
    if (!c || !out_sum || (n && !terms)) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return sequential_sum_impl(c, terms, n, out_sum, out_stats);
}

int b3d_voxel_downsample(b3d_ctx* c, const float* xyz, size_t n, const float* colors_or_null, float voxel_size,
                         float* out_xyz, float* out_colors_or_null, size_t capacity, size_t* out_n) {
    if (!c || !out_n || (n && (!xyz || !out_xyz))) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return voxel_downsample_impl(c, xyz, n, colors_or_null, voxel_size, out_xyz, out_colors_or_null, capacity, out_n);
}

int b3d_prepare_model(b3d_ctx* c, const float* model_xyz, size_t n, float voxel_size, int normals_k, float fpfh_radius, size_t* out_n_points) {
    if (!c || (n && !model_xyz)) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return prepare_model_impl(c, model_xyz, n, voxel_size, normals_k, fpfh_radius, out_n_points);
}

int b3d_register_scene(b3d_ctx* c, const float* scene_xyz, size_t n, float voxel_size, int normals_k, float fpfh_radius,
                       int ransac_max_iterations, float ransac_confidence, float icp_distance_threshold, int icp_max_iterations,
                       int point_to_plane, b3d_scene_result* out) {
    if (!c || !out || (n && !scene_xyz)) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return register_scene_impl(c, scene_xyz, n, voxel_size, normals_k, fpfh_radius, ransac_max_iterations, ransac_confidence,
                               icp_distance_threshold, icp_max_iterations, point_to_plane, out);
}

int b3d_register_scene_device(b3d_ctx* c, const float* scene_xyz_dev, size_t n, float voxel_size, int normals_k, float fpfh_radius,
                              int ransac_max_iterations, float ransac_confidence, float icp_distance_threshold, int icp_max_iterations,
                              int point_to_plane, b3d_scene_result* out) {
    if (!c || !out || (n && !scene_xyz_dev)) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return register_scene_device_impl(c, scene_xyz_dev, n, voxel_size, normals_k, fpfh_radius, ransac_max_iterations, ransac_confidence,
                                      icp_distance_threshold, icp_max_iterations, point_to_plane, out);
}

int b3d_depth_to_cloud(b3d_ctx* c, const uint16_t* depth, int width, int height, const uint8_t* mask_or_null, int mask_width, int mask_height,
                       float scale_to_meters, float clipping_max, float fx, float fy, float cx, float cy, const uint8_t* bgr_or_null,
                       float* out_xyz, float* out_rgb_or_null, size_t capacity, size_t* out_n) {
    if (!c || !depth || !out_xyz || !out_n) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return depth_to_cloud_impl(c, depth, width, height, mask_or_null, mask_width, mask_height, scale_to_meters, clipping_max, fx, fy, cx, cy,
                               bgr_or_null, out_xyz, out_rgb_or_null, capacity, out_n);
}

int b3d_register_depth(b3d_ctx* c, const uint16_t* depth, int width, int height, const uint8_t* mask_or_null, int mask_width, int mask_height,
                       float scale_to_meters, float clipping_max, float fx, float fy, float cx, float cy, float voxel_size, int normals_k,
                       float fpfh_radius, int ransac_max_iterations, float ransac_confidence, float icp_distance_threshold,
                       int icp_max_iterations, int point_to_plane, b3d_scene_result* out) {
    if (!c || !depth || !out) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return register_depth_impl(c, depth, width, height, mask_or_null, mask_width, mask_height, scale_to_meters, clipping_max, fx, fy, cx, cy,
                               voxel_size, normals_k, fpfh_radius, ransac_max_iterations, ransac_confidence, icp_distance_threshold,
                               icp_max_iterations, point_to_plane, out);
}

int b3d_world_poses(b3d_ctx* c, const float* refined_T, size_t n, const float extrinsics_or_null[16], float* out_T) {
    if (!c || (n && (!refined_T || !out_T))) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return world_poses_impl(c, refined_T, n, extrinsics_or_null, out_T);
}

int b3d_filter_duplicates(b3d_ctx* c, const float* poses, size_t n, float min_distance, float* out_poses, size_t* out_n) {
    if (!c || !out_n || (n && (!poses || !out_poses))) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return filter_duplicates_impl(c, poses, n, min_distance, out_poses, out_n);
}

int b3d_estimate_normals(b3d_ctx* c, const float* xyz, size_t n, int k, float* out_normals) {
    if (!c || (n && (!xyz || !out_normals))) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return estimate_normals_impl(c, xyz, n, k, out_normals);
}

int b3d_compute_fpfh(b3d_ctx* c, const float* xyz, const float* normals, size_t n, float radius, float* out_desc) {
    if (!c || (n && (!xyz || !normals || !out_desc))) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return compute_fpfh_impl(c, xyz, normals, n, radius, out_desc);
}

int b3d_set_icp_mode(b3d_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 3) return B3D_ERR_INVALID;
    c->icp_mode = mode;
    return B3D_OK;
}

int b3d_set_finish_mode(b3d_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 1) return B3D_ERR_INVALID;
    c->finish_mode = mode;
    return B3D_OK;
}

int b3d_set_score_mode(b3d_ctx* c, int mode) {
    if (!c || mode < 0 || mode > 4) return B3D_ERR_INVALID;
    c->score_mode = mode;
    return B3D_OK;
}

int b3d_match_features(b3d_ctx* c, size_t row0, size_t row1) {
    if (!c) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    int rc = match_features_impl(c, row0, row1);
    if (rc == B3D_OK && row0 == 0 && row1 == c->n_src) c->have_corr = true;
    return rc;
}

int b3d_get_correspondences(b3d_ctx* c, uint32_t* out_host) {
    if (!c || !out_host) return B3D_ERR_INVALID;
    if (!c->have_clouds || c->corr.cap < sizeof(uint32_t) * c->n_src) return fail(c, B3D_ERR_STATE, "get_correspondences: none computed");
    B3D_CUDA(c, enter(c));
    if (c->n_src) B3D_CUDA(c, cudaMemcpyAsync(out_host, c->corr.p, sizeof(uint32_t) * c->n_src, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    return B3D_OK;
}

int b3d_get_correspondences_dev(b3d_ctx* c, uint32_t* out_dev) {
    if (!c || !out_dev) return B3D_ERR_INVALID;
    if (!c->have_clouds || c->corr.cap < sizeof(uint32_t) * c->n_src) return fail(c, B3D_ERR_STATE, "get_correspondences_dev: none computed");
    B3D_CUDA(c, enter(c));
    if (c->n_src) B3D_CUDA(c, cudaMemcpyAsync(out_dev, c->corr.p, sizeof(uint32_t) * c->n_src, cudaMemcpyDeviceToDevice, c->stream));
    return B3D_OK;
}

int b3d_set_correspondences(b3d_ctx* c, const uint32_t* corr, int on_device) {
    if (!c) return B3D_ERR_INVALID;
    if (!c->have_clouds) return fail(c, B3D_ERR_STATE, "set_correspondences: call set_clouds first");
    B3D_CUDA(c, enter(c));
    B3D_CUDA(c, c->corr.ensure(sizeof(uint32_t) * (c->n_src ? c->n_src : 1)));
    if (corr && c->n_src && corr != c->corr.as<uint32_t>())
        B3D_CUDA(c, cudaMemcpyAsync(c->corr.p, corr, sizeof(uint32_t) * c->n_src, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    c->have_corr = true; c->prepared = false;
    return B3D_OK;
}

int b3d_correspondences_devptr(b3d_ctx* c, void** out) {
    if (!c || !out) return B3D_ERR_INVALID;
    if (!c->have_clouds) return fail(c, B3D_ERR_STATE, "correspondences_devptr: call set_clouds first");
    B3D_CUDA(c, enter(c));
    B3D_CUDA(c, c->corr.ensure(sizeof(uint32_t) * (c->n_src ? c->n_src : 1)));
    *out = c->corr.p;
    return B3D_OK;
}

int b3d_ransac_prepare(b3d_ctx* c, float voxel, int max_iterations, float confidence) {
    if (!c) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return ransac_prepare_impl(c, voxel, max_iterations, confidence);
}
int b3d_ransac_score(b3d_ctx* c, int h0, int h1) {
    if (!c) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return ransac_score_impl(c, h0, h1);
}
int b3d_ransac_reduce(b3d_ctx* c, int h0, int h1, const int64_t* limit_key_dev, int64_t* keys_dev) {
    if (!c) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return ransac_reduce_impl(c, h0, h1, limit_key_dev, keys_dev);
}
int b3d_ransac_reduce3(b3d_ctx* c, int h0, int h1, int64_t* keys3_dev) {
    if (!c) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return ransac_reduce3_impl(c, h0, h1, reinterpret_cast<unsigned long long*>(keys3_dev));
}

int b3d_ransac_finish(b3d_ctx* c, const int64_t* keys_dev, float* T, float* fitness, float* rmse, int32_t* best) {
    if (!c || !T || !fitness || !rmse) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return ransac_finish_impl(c, keys_dev, T, fitness, rmse, best);
}

int b3d_ransac_counts(b3d_ctx* c, int h0, int h1, int32_t* out_host) {
    if (!c || !out_host) return B3D_ERR_INVALID;
    if (!c->prepared) return fail(c, B3D_ERR_STATE, "ransac_counts: not prepared");
    if (h0 < 0 || h1 > c->H || h0 > h1) return fail(c, B3D_ERR_INVALID, "ransac_counts: bad range");
    B3D_CUDA(c, enter(c));
    { int rc = ransac_generate_impl(c, h0, h1); if (rc != B3D_OK) return rc; }   // degenerate-triple flags need the hypotheses
    if (h1 > h0 && c->n_src) B3D_CUDA(c, cudaMemcpyAsync(out_host, c->counts.as<int>() + h0, sizeof(int) * (size_t)(h1 - h0), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int h = h0; h < h1; ++h) {
        const bool was_scored = c->scored && h >= c->scored_lo && h < c->scored_hi;
        if (c->n_src == 0) out_host[h - h0] = -2;
        else if (!was_scored && out_host[h - h0] >= 0) out_host[h - h0] = -2;
    }
    return B3D_OK;
}

int b3d_ransac_hypotheses(b3d_ctx* c, int h0, int h1, float* out_host) {
    if (!c || !out_host) return B3D_ERR_INVALID;
    if (!c->prepared) return fail(c, B3D_ERR_STATE, "ransac_hypotheses: not prepared");
    if (h0 < 0 || h1 > c->H || h0 > h1) return fail(c, B3D_ERR_INVALID, "ransac_hypotheses: bad range");
    B3D_CUDA(c, enter(c));
    if (h1 == h0 || c->n_src == 0) return B3D_OK;
    { int rc = ransac_generate_impl(c, h0, h1); if (rc != B3D_OK) return rc; }
    const size_t n = (size_t)(h1 - h0);
    float* tmp = nullptr;
    B3D_CUDA(c, cudaMallocHost(&tmp, sizeof(float) * 12 * n));
    cudaError_t e = cudaMemcpy2DAsync(tmp, sizeof(float) * n, c->hyp.as<float>() + h0, sizeof(float) * (size_t)c->H, sizeof(float) * n, 12,
                                      cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) for (size_t h = 0; h < n; ++h) for (int k = 0; k < 12; ++k) out_host[h * 12 + k] = tmp[(size_t)k * n + h];
    cudaFreeHost(tmp);
    if (e != cudaSuccess) return fail_cuda(c, e, "ransac_hypotheses copy", __FILE__, __LINE__);
    return B3D_OK;
}

int b3d_icp_run(b3d_ctx* c, const float* T0, float thr, int max_iter, int p2plane, int stop_on_conv,
                float* T, float* fitness, float* rmse, int32_t* iters) {
    if (!c || !T0 || !T || !fitness || !rmse) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return icp_run_impl(c, T0, thr, max_iter, p2plane, stop_on_conv, T, fitness, rmse, iters);
}

int b3d_icp_nearest(b3d_ctx* c, const float* T, float thr, uint32_t* idx_host, float* d2_host) {
    if (!c || !T || !idx_host || !d2_host) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return icp_nearest_impl(c, T, thr, idx_host, d2_host);
}

// ---- whole-path entry points -------------------------------------------------------------

int b3d_ransac(b3d_ctx* c, const float* src_xyz, size_t n_src, const float* tgt_xyz, size_t n_tgt,
               const float* src_desc, const float* tgt_desc, float voxel_size, int max_iterations, float confidence,
               float* out_T, float* out_fitness, float* out_rmse, int32_t* out_best) {
    if (!c || !out_T || !out_fitness || !out_rmse) return B3D_ERR_INVALID;
    int rc = b3d_set_clouds(c, src_xyz, n_src, tgt_xyz, nullptr, n_tgt, 0); if (rc) return rc;
    rc = b3d_set_features(c, src_desc, tgt_desc, 0); if (rc) return rc;
    rc = b3d_match_features(c, 0, n_src); if (rc) return rc;
    rc = b3d_ransac_prepare(c, voxel_size, max_iterations, confidence); if (rc) return rc;
    rc = b3d_ransac_score(c, 0, max_iterations); if (rc) return rc;
    int64_t* keys = reinterpret_cast<int64_t*>(&c->state.as<DeviceState>()->best_key);
    rc = b3d_ransac_reduce(c, 0, max_iterations, nullptr, keys); if (rc) return rc;
    return b3d_ransac_finish(c, keys, out_T, out_fitness, out_rmse, out_best);
}

int b3d_icp(b3d_ctx* c, const float* src_xyz, size_t n_src, const float* tgt_xyz, const float* tgt_normals, size_t n_tgt,
            const float* T0, float thr, int max_iterations, int point_to_plane,
            float* out_T, float* out_fitness, float* out_rmse, int32_t* out_iters) {
    if (!c || !T0 || !out_T || !out_fitness || !out_rmse) return B3D_ERR_INVALID;
    int rc = b3d_set_clouds(c, src_xyz, n_src, tgt_xyz, tgt_normals, n_tgt, 0); if (rc) return rc;
    return b3d_icp_run(c, T0, thr, max_iterations, point_to_plane, 1, out_T, out_fitness, out_rmse, out_iters);
}

}  // extern "C"
