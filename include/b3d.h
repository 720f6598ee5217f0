/* =============================================================================
 * b3d.h — C-ABI of the B200-native registration hot path.
 *
 * This is the drop-in boundary for ONE path of stojicnnnn/3DVision: FPFH
 * descriptor matching + RANSAC + ICP.  Every entry point cites the reference
 * interface it replaces (paths relative to the reference repo).  Plain pointers
 * and sizes only; no C++/torch types; functions never throw — they return 0 on
 * success or a negative b3d_status.  There is NO CPU fallback: without a usable
 * CUDA device every compute entry point returns B3D_ERR_NO_DEVICE.
 *
 * Layout conventions (identical to the reference's in-memory types):
 *   points / normals : packed float xyz, 12 B per point
 *                      (std::vector<Eigen::Vector3f>, include/registration.hpp:10-19)
 *   descriptors      : row-major float[n][33], 132 B stride
 *                      (std::vector<std::array<float,33>>, include/registration.hpp:21-24)
 *   transforms       : float[16] column-major (Eigen::Matrix4f storage,
 *                      include/registration.hpp:26-30)
 * Threading: a b3d_ctx is single-threaded; use one context per host thread
 * (the orchestrator calls from a pool, src/pipeline.cpp:321-327).  Contexts are
 * independent (own stream + workspace) and may run concurrently.
 * ============================================================================= */
#ifndef B3D_H_
#define B3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b3d_ctx b3d_ctx;

typedef enum b3d_status {
    B3D_OK             = 0,
    B3D_ERR_NO_DEVICE  = -1,   /* no CUDA device / driver: caller's catch(...) path (src/pipeline.cpp:114) */
    B3D_ERR_CUDA       = -2,   /* a CUDA runtime call or kernel failed; see b3d_last_error() */
    B3D_ERR_INVALID    = -3,   /* bad argument (null pointer, size 0 where not allowed, ...) */
    B3D_ERR_ALLOC      = -4,   /* device or host allocation failed */
    B3D_ERR_STATE      = -5,   /* staged call issued out of order */
    B3D_ERR_RNG_WINDOW = -6    /* internal: raw RNG window too small (retried automatically) */
} b3d_status;

#define B3D_DESC_DIM 33
#define B3D_NO_MATCH 0xFFFFFFFFu

/* ---- lifetime ---------------------------------------------------------------- */

/* GPURegistration::isCudaAvailable(), include/gpu_registration.hpp:18 (src/gpu_impl.cpp).
 * 1 if a CUDA device with compute capability 10.x is usable, else 0. Never fails. */
int b3d_cuda_available(void);

/* Number of usable devices: CUDA devices 0 .. n-1 of this process with compute capability 10.x (the library carries
 * sm_100a code only).  0 if there is none.  Device ordinals are CUDA's (CUDA_VISIBLE_DEVICES applies). */
int b3d_device_count(void);

/* Creates a context on `device` (stream + persistent workspace; no cudaMalloc on
 * the steady-state call path).  Replaces the per-call cudaMalloc/cudaFree block of
 * GPURegistration::icpRefine, src/gpu_impl.cpp:155-186, 246-255. */
int  b3d_ctx_create(int device, b3d_ctx** out);
void b3d_ctx_destroy(b3d_ctx* ctx);

/* Run all work of this context on an externally owned cudaStream_t (e.g. the
 * current torch stream) instead of the context's own stream. NULL restores the own
 * stream; to select the legacy default stream pass cudaStreamLegacy ((void*)0x1). */
int b3d_ctx_set_stream(b3d_ctx* ctx, void* cuda_stream);

const char* b3d_strerror(int status);
const char* b3d_last_error(const b3d_ctx* ctx);

/* ---- whole-path entry points (host buffers in, 18 floats out) ------------------- */

/* Registration::ransacRegistration(source, target, source_features, target_features,
 *   voxel_size, max_iterations = 100000, confidence = 0.999f)
 * include/registration.hpp:40-48, src/registration.cpp:204-295.
 * Feature matching (strict-< argmin, lowest index on ties), mt19937(42) hypotheses,
 * inlier scoring at 1.5*voxel_size, strict-> best, early exit on fitness > confidence.
 * out_best_iteration (optional): id of the winning hypothesis, -1 if none. */
int b3d_ransac(b3d_ctx* ctx,
               const float* src_xyz, size_t n_src,
               const float* tgt_xyz, size_t n_tgt,
               const float* src_desc, const float* tgt_desc,
               float voxel_size, int max_iterations, float confidence,
               float out_T_colmajor[16], float* out_fitness, float* out_rmse,
               int32_t* out_best_iteration);

/* Registration::icpRefine(source, target, initial_transform, distance_threshold,
 *   max_iterations = 200, point_to_plane = true)
 * include/registration.hpp:50-57, src/registration.cpp:297-414; and
 * GPURegistration::icpRefine(...), include/gpu_registration.hpp:10-16
 * (src/gpu_impl.cpp:141-260) which is the same call with point_to_plane = 1.
 * tgt_normals may be NULL (target.hasNormals() == false => point-to-point).
 * out_iterations (optional): number of iterations whose update was applied. */
int b3d_icp(b3d_ctx* ctx,
            const float* src_xyz, size_t n_src,
            const float* tgt_xyz, const float* tgt_normals_or_null, size_t n_tgt,
            const float T0_colmajor[16], float distance_threshold,
            int max_iterations, int point_to_plane,
            float out_T_colmajor[16], float* out_fitness, float* out_rmse,
            int32_t* out_iterations);

/* ---- staged API: resident inputs, hypothesis sharding, parity taps ---------------
 * Call order: set_clouds -> [set_features -> match_features | set_correspondences]
 *             -> ransac_prepare -> ransac_score -> ransac_reduce -> ransac_finish
 *             set_clouds -> icp_run
 * `on_device` != 0 means the pointer is a device pointer valid on the context's
 * device and stream. */

int b3d_set_clouds(b3d_ctx* ctx, const float* src_xyz, size_t n_src,
                   const float* tgt_xyz, const float* tgt_normals_or_null, size_t n_tgt, int on_device);
int b3d_set_features(b3d_ctx* ctx, const float* src_desc, const float* tgt_desc, int on_device);

/* Kernel used by descriptor matching: 0 = automatic, 1 = exact CUDA-core brute force,
 * 2 = tcgen05 tensor-core screen + exact fp32 re-score. All three return identical indices. */
int b3d_set_match_mode(b3d_ctx* ctx, int mode);
/* src/registration.cpp:216-232 for source rows [row0,row1). */
int b3d_match_features(b3d_ctx* ctx, size_t row0, size_t row1);
int b3d_get_correspondences(b3d_ctx* ctx, uint32_t* out_host /* [n_src] */);
/* Stream-ordered device-to-device copy of correspondences[n_src] into caller-owned device memory
 * (e.g. a torch tensor that is then all-reduced between ranks). */
int b3d_get_correspondences_dev(b3d_ctx* ctx, uint32_t* out_dev /* [n_src] */);
int b3d_set_correspondences(b3d_ctx* ctx, const uint32_t* corr /* [n_src] */, int on_device);
/* Device pointer to correspondences[n_src] (uint32), for an all-gather between ranks. */
int b3d_correspondences_devptr(b3d_ctx* ctx, void** out_devptr);

/* src/registration.cpp:213, 235-268: threshold, RNG stream -> index triples,
 * 3-point Kabsch (R,t) for hypotheses [0,max_iterations). */
int b3d_ransac_prepare(b3d_ctx* ctx, float voxel_size, int max_iterations, float confidence);
/* Inlier-scoring kernel: 0 = packed FMA screen with exact re-count of pairs inside the error band
 * (default), 1 = the reference's un-fused arithmetic for every pair, 2 = scalar FMA screen; all
 * three give identical counts for every hypothesis.  3 = bail-out: hypotheses that provably cannot
 * reach the best full count found so far are dropped part-way (their count reads -4); the winner,
 * its transform, fitness and rmse are identical to modes 0-2, per-hypothesis counts are not all
 * available.  Not used by default.  4 = mode 0 restricted to two hypotheses per thread (mode 0 takes four per
 * thread on long hypothesis ranges; kept for A/B timing). */
int b3d_set_score_mode(b3d_ctx* ctx, int mode);
/* src/registration.cpp:270-279 for hypothesis ids [h0,h1) (this rank's shard). */
int b3d_ransac_score(b3d_ctx* ctx, int h0, int h1);
/* src/registration.cpp:281-290 over ids [h0,h1): writes two int64 keys to keys_dev
 *   keys[0] = (fitness_bits << 32) | (0xFFFFFFFF - id)   of the local best (0 if none)
 *   keys[1] = 0xFFFFFFFF - (first id with fitness > confidence)   (0 if none)
 * Both combine across ranks with MAX.  If limit_key_dev != NULL it points to a
 * reduced keys[1]; only ids <= that first-exit id are considered for keys[0]. */
int b3d_ransac_reduce(b3d_ctx* ctx, int h0, int h1, const int64_t* limit_key_dev, int64_t* keys_dev);
/* The sharded form of the selection (what b3d_ransac_sharded gathers): three keys for ids [h0,h1) in one call,
 *   keys3[0] = best key over the ids up to this range's own first exit (all of them if it has none)
 *   keys3[1] = 0xFFFFFFFF - (first id of the range with fitness > confidence)   (0 if none)
 *   keys3[2] = best key over all ids of the range.
 * With the ranges of all ranks in id order the global result is: scan ranks, max of keys3[2] until the first rank with
 * keys3[1] != 0, whose keys3[0] enters the max last. */
int b3d_ransac_reduce3(b3d_ctx* ctx, int h0, int h1, int64_t* keys3_dev);
/* Recomputes the winner's transform/fitness/rmse from the globally reduced keys[0]. */
int b3d_ransac_finish(b3d_ctx* ctx, const int64_t* keys_dev,
                      float out_T_colmajor[16], float* out_fitness, float* out_rmse, int32_t* out_best_iteration);
/* How b3d_ransac_finish forms the winner's rmse, i.e. the reference's sequential fp32 sum of err^2 over the inliers
 * (src/registration.cpp:277): 0 (default) = the parallel exact-summation scheme (csrc/b3d_ess.cuh), 1 = one dependent
 * add per inlier (the older kernel, kept to cross-check).  Both return the same bits. */
int b3d_set_finish_mode(b3d_ctx* ctx, int mode);
/* Parity taps: per-hypothesis inlier counts (-1 degenerate triple, -2 not scored /
 * after early exit) and (R row-major 9, t 3) per hypothesis. */
int b3d_ransac_counts(b3d_ctx* ctx, int h0, int h1, int32_t* out_host);
int b3d_ransac_hypotheses(b3d_ctx* ctx, int h0, int h1, float* out_host /* 12 per hypothesis */);

/* ---- stages that feed the hot path; host buffers in and out ------------
 * Replaces Registration::voxelDownsample (include/registration.hpp:34, src/registration.cpp:29-60).
 * One output point per occupied voxel floor(p / voxel): the mean of its members added in input order,
 * emitted in the iteration order of the reference's std::unordered_map (the container's insertion / rehash rules are
 * replayed on the device with per-rehash radix sorts, its growth schedule taken from libstdc++'s own _Prime_rehash_policy),
 * because later stages index into this order.  colors_or_null / out_colors_or_null: averaged the
 * same way when given (normals are dropped, as in the reference).  *out_n = number of voxels; if it
 * exceeds `capacity` nothing is written and B3D_ERR_INVALID is returned (n is always enough). */
int b3d_voxel_downsample(b3d_ctx* ctx, const float* xyz, size_t n, const float* colors_or_null, float voxel_size,
                         float* out_xyz, float* out_colors_or_null, size_t capacity, size_t* out_n);
/* Replaces Registration::estimateNormals (registration.hpp:36, src/registration.cpp:63-81, 105-130):
 * k nearest (self included) ordered by (d2, index), centroid and covariance summed in that order,
 * eigenvector of the smallest eigenvalue (Eigen SelfAdjointEigenSolver, iterative), flipped towards the
 * origin.  1 <= k <= 128. */
int b3d_estimate_normals(b3d_ctx* ctx, const float* xyz, size_t n, int k, float* out_normals);
/* Replaces Registration::computeFPFH (registration.hpp:38, src/registration.cpp:83-102, 133-201):
 * neighbours with d2 <= radius^2, the 100 first by (d2, index); SPFH bins; 1/dist-weighted sum in list
 * order; L1 normalisation.  out_desc is n x 33. */
int b3d_compute_fpfh(b3d_ctx* ctx, const float* xyz, const float* normals, size_t n, float radius, float* out_desc);

/* ---- one call per registration, everything resident ------------------------------------------------
 * The per-instance body of Pipeline::processInstance (src/pipeline.cpp:86-129: voxelDownsample ->
 * estimateNormals -> computeFPFH -> ransacRegistration -> icpRefine) as ONE call, with the model side
 * prepared once per run as Pipeline::run does with the reference model (src/pipeline.cpp:275-294).
 * Identical results to the five separate calls (tests/test_gpu_features.py); no host containers and no
 * PCIe round trips between the stages.  b3d_set_clouds / b3d_ransac / b3d_icp replace the resident model. */
typedef struct b3d_scene_result {
    float coarse_T[16];             /* ransacRegistration result, column-major */
    float coarse_fitness, coarse_rmse;
    int32_t coarse_best_iteration;  /* -1 if no hypothesis had an inlier */
    float T[16];                    /* icpRefine result */
    float fitness, rmse;
    int32_t icp_iterations;
    uint32_t n_source_points;       /* scene points after down-sampling */
} b3d_scene_result;
int b3d_prepare_model(b3d_ctx* ctx, const float* model_xyz, size_t n, float voxel_size, int normals_k, float fpfh_radius,
                      size_t* out_n_points);
int b3d_register_scene(b3d_ctx* ctx, const float* scene_xyz, size_t n, float voxel_size, int normals_k, float fpfh_radius,
                       int ransac_max_iterations, float ransac_confidence, float icp_distance_threshold, int icp_max_iterations,
                       int point_to_plane, b3d_scene_result* out);

/* b3d_register_scene for a cloud that already lives on this context's device (packed xyz floats, e.g. the output of a
 * camera pipeline): no host hop at all.  The memory is only read, on the context's stream. */
int b3d_register_scene_device(b3d_ctx* ctx, const float* scene_xyz_dev, size_t n, float voxel_size, int normals_k, float fpfh_radius,
                              int ransac_max_iterations, float ransac_confidence, float icp_distance_threshold, int icp_max_iterations,
                              int point_to_plane, b3d_scene_result* out);

/* Depth image of one instance -> cloud: the CPU branch of Pipeline::processInstance, src/pipeline.cpp:38-84 (what
 * GPUDepth::preprocess + GPUPointCloud::generate, include/gpu_depth.hpp:9-22, do on the reference's GPU branch, but in
 * the CPU branch's raster order): z = float(depth) * float(1 / scale_to_meters) (OpenCV's 16u -> 32f convertTo); zero where mask <= 10 (mask_or_null == NULL: no
 * masking); keep 0 < z <= clipping_max; x = (u - cx) z / fx, y = (v - cy) z / fy; rgb = bgr reversed / 255.
 * mask_width x mask_height is the mask's own size (0 x 0: the depth image's size); when it differs the mask is read
 * through OpenCV's nearest-neighbour resize map, cv::resize(mask, ..., depth.size(), 0, 0, INTER_NEAREST) of
 * src/pipeline.cpp:39-41 (source pixel min(floor(x / (width / (double)mask_width)), mask_width - 1)).  The colour image
 * must have the depth image's size.  capacity / out_n as in b3d_voxel_downsample (width*height always suffices). */
int b3d_depth_to_cloud(b3d_ctx* ctx, const uint16_t* depth, int width, int height, const uint8_t* mask_or_null, int mask_width, int mask_height,
                       float scale_to_meters, float clipping_max, float fx, float fy, float cx, float cy, const uint8_t* bgr_or_null,
                       float* out_xyz, float* out_rgb_or_null, size_t capacity, size_t* out_n);
/* b3d_depth_to_cloud followed by b3d_register_scene without the cloud leaving the device: src/pipeline.cpp:38-129. */
int b3d_register_depth(b3d_ctx* ctx, const uint16_t* depth, int width, int height, const uint8_t* mask_or_null, int mask_width, int mask_height,
                       float scale_to_meters, float clipping_max, float fx, float fy, float cx, float cy, float voxel_size, int normals_k,
                       float fpfh_radius, int ransac_max_iterations, float ransac_confidence, float icp_distance_threshold,
                       int icp_max_iterations, int point_to_plane, b3d_scene_result* out);

/* ---- pose post-processing of the orchestrator ---------------------------------------------
 * src/pipeline.cpp:136-137 for n refined poses at once: out[i] = camera_extrinsics * refined_T[i].inverse()
 * (extrinsics_or_null == NULL: the inverse, T_camera_object, alone).  All matrices float[16] column-major, packed.
 * Matrix4f::inverse() follows Eigen 3.4's SSE kernel operation for operation, the product its packet order. */
int b3d_world_poses(b3d_ctx* ctx, const float* refined_T, size_t n, const float extrinsics_or_null[16], float* out_T);
/* Pipeline::filterDuplicates(waypoints, min_distance), src/pipeline.cpp:153-180: walks the n poses in order; a pose whose
 * translation is closer than min_distance to a pose kept so far is dropped, or takes that pose's slot if it is nearer
 * the origin.  out holds n poses at most; *out_n = number kept. */
int b3d_filter_duplicates(b3d_ctx* ctx, const float* poses, size_t n, float min_distance, float* out_poses, size_t* out_n);

/* How the ICP sums (ATA / ATb / total_error, src/registration.cpp:343-354; centroids and cross-covariance, :374-386) are
 * accumulated.  0 (default; 2 is an alias): in the reference's order — one matched point at a time, source order, fp32 —
 * reproduced bit for bit by a parallel exact-summation scheme (csrc/b3d_ess.cuh), for both error metrics.  This is what
 * b3d_icp, b3d_register_scene and the C++ shim run: it holds the 1e-5 / 1e-6 m bar (in fact equality with the CPU
 * path) even at the orchestrator's default threshold 0.4 * voxel (src/pipeline.cpp:104), which sits at the noise floor
 * where the matched set flips with the last bit of the pose.  1 = deterministic fp64 tree sums: opt-in fast mode,
 * order-free; within 1e-5 / 1e-6 m on well-conditioned thresholds, but up to ~1e-4 away at the noise floor and for
 * point-to-point (the reference's own rounding noise).  3 = the reference order through one dependent add chain per sum
 * (slow; an independent implementation kept to cross-check mode 0 at sizes the CPU oracle cannot reach). */
int b3d_set_icp_mode(b3d_ctx* ctx, int mode);
/* ICP on resident clouds. stop_on_convergence = 0 disables the |d rmse| < 1e-6 break
 * (src/registration.cpp:406) for fixed-iteration throughput runs. */
int b3d_icp_run(b3d_ctx* ctx, const float T0_colmajor[16], float distance_threshold,
                int max_iterations, int point_to_plane, int stop_on_convergence,
                float out_T_colmajor[16], float* out_fitness, float* out_rmse, int32_t* out_iterations);
/* Parity tap for src/registration.cpp:325-338: nearest target of every transformed
 * source point under T within distance_threshold.  idx = B3D_NO_MATCH where none. */
int b3d_icp_nearest(b3d_ctx* ctx, const float T_colmajor[16], float distance_threshold,
                    uint32_t* out_idx_host, float* out_d2_host);

/* ---- batched multi-object registration: the orchestrator's worker pool ---------------------------------
 * Pipeline::run pushes one processInstance task per object into a pool of num_threads workers (src/pipeline.cpp:16,
 * 321-327; include/thread_pool.hpp); b3d_pool is that pool behind the C-ABI: n_workers persistent host threads, each
 * with its own context on devices[w % n_devices] (n_devices == 0: device 0), so one pool spreads over all the GPUs of a
 * box.  b3d_pool_register runs ransacRegistration + icpRefine (src/pipeline.cpp:97-129) for every instance and blocks
 * until all are done; results[i] belongs to items[i] and equals what b3d_ransac + b3d_icp return for it.  Instances are
 * independent: no collective.  The first non-zero per-instance status is returned (each result carries its own). */
typedef struct b3d_instance {
    const float* src_xyz; size_t n_src;                         /* scene instance cloud (source_down) */
    const float* tgt_xyz; const float* tgt_normals_or_null; size_t n_tgt;   /* model cloud (+ normals) */
    const float* src_desc; const float* tgt_desc;               /* FPFH, n x 33 */
    float voxel_size; int ransac_max_iterations; float ransac_confidence;
    float icp_distance_threshold; int icp_max_iterations; int point_to_plane;
} b3d_instance;
typedef struct b3d_instance_result {
    float coarse_T[16]; float coarse_fitness, coarse_rmse;      /* ransacRegistration */
    float T[16]; float fitness, rmse; int32_t icp_iterations;   /* icpRefine */
    int32_t status;                                             /* b3d_status of this instance */
} b3d_instance_result;
typedef struct b3d_pool b3d_pool;
int  b3d_pool_create(int n_workers, const int* devices, int n_devices, b3d_pool** out);
int  b3d_pool_register(b3d_pool* pool, const b3d_instance* items, size_t n, b3d_instance_result* results);
void b3d_pool_destroy(b3d_pool* pool);

/* ---- multi-GPU: one process (or host thread) per GPU, NCCL over NVLink --------------------------------
 * The two parts of the path that shard: feature-matching rows (one in-place ncclAllGather of the index slices) and RANSAC
 * hypothesis ids (contiguous ranges; one ncclAllGather of three 64-bit keys per rank, from which every rank resolves the
 * reference's sequential selection — strict-> best, break at the first id with fitness > confidence,
 * src/registration.cpp:281-290 — and rebuilds the winner from its index triple).  Results are identical on every rank
 * and identical to the single-GPU call.  Batched multi-object work needs no collective (one context per instance, any
 * device); single-cloud ICP stays on one GPU.  NCCL is bound at run time (libnccl.so.2), there is no link-time dependency.
 *
 * b3d_comm_unique_id: ncclGetUniqueId into 128 caller bytes (rank 0 calls it and hands the bytes to the other ranks by
 * whatever the host program uses: MPI, a file, torch.distributed, shared memory between pool threads).
 * b3d_comm_init: ncclCommInitRank on the context's device; collective over all `world` ranks.  world == 1 is a no-op group.
 * b3d_comm_attach: use a communicator the host program already owns (an ncclComm_t) instead; never destroyed by us. */
int b3d_comm_unique_id(void* out_id128);
int b3d_comm_init(b3d_ctx* ctx, const void* id128, int rank, int world);
int b3d_comm_attach(b3d_ctx* ctx, void* nccl_comm, int rank, int world);
int b3d_comm_destroy(b3d_ctx* ctx);
/* Registration::ransacRegistration (include/registration.hpp:40-48) called collectively by every rank with the SAME host
 * inputs; each rank uploads the clouds, the target descriptors and only its own rows of the source descriptors. */
int b3d_ransac_sharded(b3d_ctx* ctx, const float* src_xyz, size_t n_src, const float* tgt_xyz, size_t n_tgt,
                       const float* src_desc, const float* tgt_desc, float voxel_size, int max_iterations, float confidence,
                       float out_T_colmajor[16], float* out_fitness, float* out_rmse, int32_t* out_best_iteration);
/* The same on inputs already resident (b3d_set_clouds + b3d_set_features on every rank); match_features == 0 uses the
 * correspondences set by b3d_set_correspondences instead of matching. */
int b3d_ransac_sharded_resident(b3d_ctx* ctx, float voxel_size, int max_iterations, float confidence, int match_features,
                                float out_T_colmajor[16], float* out_fitness, float* out_rmse, int32_t* out_best_iteration);
/* b3d_register_scene called collectively: every rank runs the front end on the whole scene, matching and scoring are
 * sharded as above, the refinement runs on every rank from the common coarse pose. */
int b3d_register_scene_sharded(b3d_ctx* ctx, const float* scene_xyz, size_t n, float voxel_size, int normals_k, float fpfh_radius,
                               int ransac_max_iterations, float ransac_confidence, float icp_distance_threshold, int icp_max_iterations,
                               int point_to_plane, b3d_scene_result* out);

/* ---- instrumentation ---------------------------------------------------------------- */
/* Number of kernels this context has launched since creation (bench.py: gpu_launches). */
uint64_t b3d_kernel_launches(const b3d_ctx* ctx);
/* Device time in ms of the last call of each stage, measured with CUDA events on the
 * context's stream: 0 match, 1 ransac_prepare, 2 ransac_score, 3 ransac_reduce+finish,
 * 4 icp grid build, 5 icp iterations (incl. the lazily built second level), 6 icp source binning, 7 voxel down-sampling,
 * 8 normals, 9 FPFH. Returns -1 for an unknown stage or one that has not run. */
float b3d_stage_ms(const b3d_ctx* ctx, int stage);

/* Number of 32-pair groups the last b3d_ransac_score call had to re-count with the reference
 * arithmetic because a pair fell inside the screening error band (diagnostic). */
int b3d_score_recounts(b3d_ctx* ctx, uint64_t* out_groups);

/* Exact-sum diagnostics of the last b3d_icp / b3d_icp_run call in a reference-order mode (3dvision_b200/csrc/b3d_ess.cuh):
 * per running sum v < 32, out[4v] = walk rounds, out[4v+1] = 32-term blocks that had to be added term by term,
 * out[4v+2] = SM cycles / 16 its chain took, accumulated over the call's iterations (sums 0-27: point-to-plane;
 * 0-6 and 16-24: the two point-to-point passes). */
int b3d_icp_exact_sum_stats(b3d_ctx* ctx, uint32_t out[128]);

/* Diagnostic: R = AngleAxis(x, UnitX) * AngleAxis(y, UnitY) * AngleAxis(z, UnitZ) as a matrix, the rotation of the ICP update
 * (src/registration.cpp:369-371), for n caller-supplied angle triples (host, [n][3]) -> out_R_rowmajor (host, [n][9]). Exercises the
 * device build of glibc's sinf / cosf (csrc/b3d_libm.cuh) and of Eigen's quaternion product on arbitrary arguments. */
int b3d_euler_rotations(b3d_ctx* ctx, const float* angles_xyz, size_t n, float* out_R_rowmajor);

/* Diagnostic: the term arrays terms[v][stride] (v < 28 point-to-plane; stride = n_src rounded up to 4096) the last iteration of the
 * last reference-order b3d_icp / b3d_icp_run call summed, and the 32 sums it obtained — lets a test re-add them in order on the host.
 * terms_out may be NULL (sums only); at most capacity_floats floats are copied. */
int b3d_icp_exact_sum_dump(b3d_ctx* ctx, float* terms_out, size_t capacity_floats, size_t* out_stride, float sums_out[32]);

/* Diagnostic: the fp32 value of `float s = 0; for (i < n) s += terms[i];` (the accumulation loops of
 * src/registration.cpp:277 `total_error += err * err` and :351-357 / :377-391, whose order the default modes keep), computed by
 * the same parallel exact-sum passes those paths use (b3d_ess.cuh) — so they can be checked on arbitrary, adversarial term
 * sequences. terms: host, n floats. out_stats (may be NULL): walk rounds, 32-term blocks added term by term, SM cycles / 16. */
int b3d_sequential_sum(b3d_ctx* ctx, const float* terms, size_t n, float* out_sum, uint32_t out_stats[3]);

/* Measures the sustained issue rate of separate (un-fused) FMUL + FADD instructions on this
 * device, in lane-operations per second: the roofline denominator of the scoring kernel, whose
 * arithmetic must not be contracted into FMAs (reference build: no FMA, README.md:13). */
int b3d_measure_fp32_rate(b3d_ctx* ctx, double* out_ops_per_second);

#ifdef __cplusplus
}
#endif
#endif /* B3D_H_ */
