// b3d_common.cuh — context, workspace and launch plumbing shared by the kernels.
// Host side of the C-ABI in include/b3d.h.  No Eigen, no torch, no CPU compute path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <random>
#include <string>

#include "../../include/b3d.h"

namespace b3d {

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr int kDescDim = B3D_DESC_DIM;
constexpr int kStages = 10;

// Workspace growth goes through the stream-ordered allocator on the calling context's stream (set by enter() at every
// C-ABI entry): cudaFree / cudaMalloc synchronise the whole device, which stalls every other worker thread's context
// whenever one of them meets a larger instance (measured: batched throughput swinging 3x between runs).
inline cudaStream_t& alloc_stream() { static thread_local cudaStream_t s = nullptr; return s; }
inline bool& alloc_stream_valid() { static thread_local bool v = false; return v; }

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool ordered_alloc = false;               // p came from cudaMallocAsync
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    // Grow-only: steady-state calls with non-growing sizes never touch the allocator.
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        const bool ordered = alloc_stream_valid();
        if (p) {
            cudaError_t e = (ordered && ordered_alloc) ? cudaFreeAsync(p, alloc_stream()) : cudaFree(p);
            p = nullptr; cap = 0;
            if (e != cudaSuccess) return e;
        }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = ordered ? cudaMallocAsync(&p, want, alloc_stream()) : cudaMalloc(&p, want);
        if (e == cudaSuccess) { cap = want; ordered_alloc = ordered; }
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Device-resident scalars shared by the staged kernels (one 256 B block per context).
struct DeviceState {
    // RANSAC
    unsigned long long best_key;      // (fitness_bits << 32) | (0xFFFFFFFF - id); 0 = none
    unsigned long long exit_key;      // 0xFFFFFFFF - first id with fitness > confidence; 0 = none
    unsigned int accepted_total;      // accepted RNG draws inside the raw window
    unsigned int pad0;
    unsigned int pair_smax_bits;      // max |coordinate| over source / matched-target points of the pair array
    unsigned int pair_qmax_bits;
    unsigned long long score_recounts; // 32-pair groups re-counted with the reference arithmetic by the last score call
    // ICP
    float T[16];                      // current transform, column-major
    float res_T[16];                  // RegistrationResult.transformation
    float res_fitness, res_rmse;
    int iterations;                   // iterations applied
    int done;                         // 1 once converged / broke out
    int n_corr_last;
    unsigned int seq_count;           // matched pairs of the current reference-order ICP iteration
    unsigned int ess_arrive;          // CTAs of the current exact-sum chain launch that have stored their sum
    float ess_sums[32];               // the exact sequential sums of the current iteration (b3d_ess.cuh)
    float ess_means[8];               // point-to-point: src_mean(3), tgt_mean(3) between the two passes
    unsigned int ess_stats[32][4];    // per sum, accumulated over the call: walk rounds, blocks added term by term, SM cycles / 16 (diagnostic)
    float out18[20];                  // RANSAC result: T(16), fitness, rmse, best id (as int bits), unused
    float fin_Rt[12];                 // RANSAC finish: the winner's (R row-major, t), rebuilt from its index triple
    int fin_id, fin_none;             // the winner's id; 1 if no hypothesis had an inlier
    int rng_starved;                  // a hypothesis needed index draws beyond the accepted-draw window (count -3)
    int score_exit, score_exit_id;    // chunked scoring: an id with fitness > confidence has been met; the first such id
};

}  // namespace b3d

struct b3d_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;

    // resident clouds / features
    size_t n_src = 0, n_tgt = 0;
    bool has_normals = false, have_clouds = false, have_feats = false, have_corr = false;
    b3d::DevBuf stage_a, stage_b, stage_c;   // raw xyz / descriptor staging for H2D
    b3d::DevBuf src4, tgt4, nrm4;            // float4 SoA-friendly point arrays
    b3d::DevBuf sdesc, tdesc;                // fp32 descriptors, row-major [n][33]
    const float* sdesc_p = nullptr;          // where the descriptors actually live (may be caller's device memory)
    const float* tdesc_p = nullptr;
    b3d::DevBuf corr;                        // uint32 [n_src]
    int match_mode = 0;                      // 0 auto, 1 exact CUDA-core kernel, 2 tcgen05 screen + exact re-score
    b3d::DevBuf tc_a_tiles, tc_b_tiles, tc_norm2, tc_best, tc_aux;

    // RANSAC
    std::mt19937 host_rng{42};               // src/registration.cpp:235 — seed is the literal 42
    size_t raw_have = 0;                     // raw 32-bit outputs cached on device
    b3d::DevBuf raw;                         // uint32 [raw_have...]
    b3d::DevBuf draws;                       // uint32 [3*H] accepted index draws
    b3d::DevBuf scan_tmp;                    // block sums for scans
    b3d::DevBuf hyp;                         // float [12][H] (SoA): R row-major 9, t 3
    b3d::DevBuf counts;                      // int32 [H]
    b3d::DevBuf pairs;                       // float4 [2][n_src]: (s_i, 0), (q_corr[i], 0)
    b3d::DevBuf bail_list_a, bail_list_b, bail_state;   // bail-out scoring: survivor id lists + device plan
    b3d::DevBuf seqsum;                      // float scratch for the exact sequential rmse sum
    int H = 0;
    int hyp_lo = 0, hyp_hi = 0;              // id range whose (R,t) have been generated since the last prepare
    unsigned pair_stride = 0;                // pair array stride (n_src rounded up to the pair tile)
    int score_mode = 0;                      // 0 packed FMA screen, 1 un-fused everywhere, 2 scalar FMA screen, 3 bail-out (exact argmax)
    float ransac_thr = 0.f, ransac_cut = 0.f, confidence = 0.f;
    bool prepared = false, scored = false;
    bool counts_pruned = false;              // the last scoring was a bail-out run: counts hold -4 marks
    int scored_lo = 0, scored_hi = 0;

    // ICP
    b3d::DevBuf grid_slots, grid_cursor, grid_pts, grid_nrm, pt_slot, pt_rank, partials;
    b3d::DevBuf nn_idx, nn_d2;
    b3d::DevBuf fine_slots, fine_pts;                        // second ICP level: neighbourhood table + per-cell 27-cell point lists
    b3d::DevBuf nbh_slot27, nbh_cursor;
    b3d::DevBuf icp_cache, icp_cache_idx;                    // per query slot: reference position + hold radius^2, and the match it certifies
    b3d::DevBuf seq_rec, seq_match, seq_P, seq_Q, seq_N;            // reference-order sums: per-query records (legacy chain: compacted pairs)
    b3d::DevBuf ess_terms, ess_bsum, ess_guess, ess_summ;           // exact sequential sums (b3d_ess.cuh): terms, fp64 block sums, guesses, summaries
    bool fin_smem_opt_in = false, tc_smem_opt_in = false, seq_smem_opt_in = false;
    int finish_mode = 0;                                     // RANSAC rmse: 0 parallel exact sum (b3d_ess.cuh), 1 one dependent add chain
    bool ess_smem_opt_in = false;                            // cudaFuncSetAttribute done on this context's device
    int icp_mode = 0;                                        // 0 / 2: sums in the reference's order (exact, parallel); 1: fp64 tree sums; 3: legacy one-chain replay
    b3d::DevBuf src_slots, src_sorted, src_slot, src_rank;   // source reordered by target cell (coherent warps)

    // feature stages (b3d_features.cu): scratch pool, slots named there
    static constexpr int kFeatureBufs = 48;
    b3d::DevBuf fbuf[kFeatureBufs];
    bool model_ready = false;                // b3d_prepare_model has left the target cloud / normals / FPFH resident

    // multi-GPU (b3d_dist.cu): communicator of the group this context is a rank of
    void* comm = nullptr;                    // ncclComm_t
    bool comm_owned = false;                 // created by b3d_comm_init (else attached)
    int comm_rank = 0, comm_world = 1;
    b3d::DevBuf dist_keys;                   // u64 [world + 1][3]: the gathered selection keys

    // device scalars + pinned host mirror
    b3d::DevBuf state;                       // b3d::DeviceState
    b3d::DeviceState* h_state = nullptr;     // pinned

    cudaEvent_t ev_start[b3d::kStages] = {}, ev_stop[b3d::kStages] = {};
    bool ev_valid[b3d::kStages] = {};
};

namespace b3d {

inline int fail_cuda(b3d_ctx* c, cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    if (c) c->err = buf;
    return (e == cudaErrorMemoryAllocation) ? B3D_ERR_ALLOC : B3D_ERR_CUDA;
}
inline int fail(b3d_ctx* c, int code, const char* msg) { if (c) c->err = msg; return code; }

#define B3D_CUDA(ctx, expr)                                                               \
    do { cudaError_t _e = (expr); if (_e != cudaSuccess) return b3d::fail_cuda((ctx), _e, #expr, __FILE__, __LINE__); } while (0)
// kernel launch bookkeeping: count it, then surface launch-configuration errors immediately
#define B3D_LAUNCHED(ctx)                                                                 \
    do { (ctx)->launches++; cudaError_t _e = cudaGetLastError();                         \
         if (_e != cudaSuccess) return b3d::fail_cuda((ctx), _e, "kernel launch", __FILE__, __LINE__); } while (0)

// every C-ABI entry: select the device and route workspace growth to this context's stream
inline cudaError_t enter(b3d_ctx* c) {
    cudaError_t e = cudaSetDevice(c->device);
    alloc_stream() = c->stream; alloc_stream_valid() = (e == cudaSuccess);
    return e;
}

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
inline int grid_for(long long work_items, int per_block, int max_waves = 8) {
    long long blocks = (work_items + per_block - 1) / per_block;
    if (blocks < 1) blocks = 1;
    long long cap = (long long)kNumSMs * max_waves;
    return (int)(blocks < cap ? blocks : cap);
}

struct StageTimer {
    b3d_ctx* c; int s;
    StageTimer(b3d_ctx* ctx, int stage) : c(ctx), s(stage) { cudaEventRecord(c->ev_start[s], c->stream); }
    ~StageTimer() { cudaEventRecord(c->ev_stop[s], c->stream); c->ev_valid[s] = true; }
};

// stages implemented in the .cu files
int match_features_impl(b3d_ctx* c, size_t row0, size_t row1);
int match_features_tc_impl(b3d_ctx* c, size_t row0, size_t row1);
int ransac_prepare_impl(b3d_ctx* c, float voxel, int max_iterations, float confidence);
int ransac_generate_impl(b3d_ctx* c, int h0, int h1);
int ransac_score_impl(b3d_ctx* c, int h0, int h1);
int ransac_reduce_impl(b3d_ctx* c, int h0, int h1, const int64_t* limit_key_dev, int64_t* keys_dev);
int ransac_finish_impl(b3d_ctx* c, const int64_t* keys_dev, float* T, float* fitness, float* rmse, int32_t* best);
int sequential_sum_impl(b3d_ctx* c, const float* terms_host, size_t n_terms, float* out_sum, uint32_t* out_stats);
int ransac_reduce3_impl(b3d_ctx* c, int h0, int h1, unsigned long long* keys3_dev);
int comm_destroy_impl(b3d_ctx* c);
int ransac_sharded_resident_impl(b3d_ctx* c, float voxel, int H, float confidence, int match, float* T, float* fitness, float* rmse, int32_t* best);
int register_scene_sharded_impl(b3d_ctx* c, const float* xyz, size_t n, float voxel, int k, float radius, int ransac_iterations, float confidence,
                                float icp_threshold, int icp_iterations, int point_to_plane, b3d_scene_result* out);
int icp_run_impl(b3d_ctx* c, const float* T0, float thr, int max_iter, int p2plane, int stop_on_conv,
                 float* T, float* fitness, float* rmse, int32_t* iters);
int icp_nearest_impl(b3d_ctx* c, const float* T, float thr, uint32_t* idx_host, float* d2_host);
int xyz_to_float4(b3d_ctx* c, const float* xyz_dev, size_t n, float4* out);
int voxel_downsample_impl(b3d_ctx* c, const float* xyz, size_t n, const float* colors, float voxel, float* out_xyz, float* out_colors,
                          size_t capacity, size_t* out_n);
int estimate_normals_impl(b3d_ctx* c, const float* xyz, size_t n, int k, float* out_normals);
int compute_fpfh_impl(b3d_ctx* c, const float* xyz, const float* normals, size_t n, float radius, float* out_desc);
int world_poses_impl(b3d_ctx* c, const float* refined, size_t n, const float* extrinsics_or_null, float* out);
int filter_duplicates_impl(b3d_ctx* c, const float* poses, size_t n, float min_distance, float* out, size_t* out_n);
int euler_rotations_impl(b3d_ctx* c, const float* angles, size_t n, float* out);
int depth_to_cloud_impl(b3d_ctx* c, const uint16_t* depth, int w, int h, const uint8_t* mask, int mask_w, int mask_h, float scale, float clip,
                        float fx, float fy, float cx, float cy, const uint8_t* bgr, float* out_xyz, float* out_rgb, size_t capacity, size_t* out_n);
int register_depth_impl(b3d_ctx* c, const uint16_t* depth, int w, int h, const uint8_t* mask, int mask_w, int mask_h, float scale, float clip, float fx, float fy,
                        float cx, float cy, float voxel, int k, float radius, int ransac_iterations, float confidence, float icp_threshold,
                        int icp_iterations, int point_to_plane, b3d_scene_result* out);
int register_scene_device_impl(b3d_ctx* c, const float* xyz_dev, size_t n, float voxel, int k, float radius, int ransac_iterations, float confidence,
                               float icp_threshold, int icp_iterations, int point_to_plane, b3d_scene_result* out);
int prepare_model_impl(b3d_ctx* c, const float* xyz, size_t n, float voxel, int k, float radius, size_t* out_n);
int register_scene_impl(b3d_ctx* c, const float* xyz, size_t n, float voxel, int k, float radius, int ransac_iterations, float confidence,
                        float icp_threshold, int icp_iterations, int point_to_plane, b3d_scene_result* out);

}  // namespace b3d
