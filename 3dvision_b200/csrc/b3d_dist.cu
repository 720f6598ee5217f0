// b3d_dist.cu — the multi-GPU form of the hot path behind the C-ABI: one process (or host thread) per GPU, an NCCL
// communicator per context, and the two shardings SURVEY.md §8(e) names —
//   feature matching: source rows split across ranks, one in-place ncclAllGather of the index slices;
//   RANSAC:           hypothesis ids split in contiguous ranges, ONE ncclAllGather of three 64-bit keys per rank
//                     (best over all its ids, its first id with fitness > confidence, best over ids up to that one);
//                     ranges are contiguous and ordered by rank, so every rank resolves the reference's sequential rule
//                     (strict-> best, break at the first exit, src/registration.cpp:281-290) from the gathered keys
//                     locally, then rebuilds the winner from its index triple — nothing else crosses NVLink.
// Everything is enqueued on the context's stream: no host synchronisation between match, gather, score, select,
// gather, resolve and finish.  NCCL is bound at run time (dlopen of libnccl.so.2 — the copy the process already
// loaded, e.g. torch's, else the system one), so libb3d.so has no link-time dependency on it and single-GPU users
// never touch it.  Batched multi-object work and single-cloud ICP need no collective and are not here.
#include "b3d_common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>

namespace b3d {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
    });
    return api;
}

static int fail_nccl(b3d_ctx* c, ncclResult_t r, const char* what) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
    return fail(c, B3D_ERR_CUDA, buf);
}
#define B3D_NCCL(ctx, expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) return fail_nccl((ctx), _r, #expr); } while (0)

// keys per rank, see header comment.  all[r * 3 + {0: best up to own exit, 1: own first exit, 2: best over all}]
__global__ void resolve_keys_kernel(const unsigned long long* __restrict__ all, int world, unsigned long long* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long best = 0ull, exit_key = 0ull;
    for (int r = 0; r < world; ++r) {
        const unsigned long long up_to_exit = all[3 * r], ex = all[3 * r + 1], overall = all[3 * r + 2];
        if (ex != 0ull) {                                   // the reference breaks here: later ids (and ranks) never ran
            best = up_to_exit > best ? up_to_exit : best;
            exit_key = ex;
            break;
        }
        best = overall > best ? overall : best;
    }
    out[0] = best; out[1] = exit_key;
}

int comm_init_impl(b3d_ctx* c, const void* id128, int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world) return fail(c, B3D_ERR_INVALID, "comm_init: bad rank / world");
    if (c->comm) return fail(c, B3D_ERR_STATE, "comm_init: the context already has a communicator");
    c->comm_rank = rank; c->comm_world = world; c->comm_owned = false;
    if (world == 1) return B3D_OK;                          // a one-rank group needs no NCCL at all
    if (!id128) return fail(c, B3D_ERR_INVALID, "comm_init: null unique id");
    if (!nccl().ok) return fail(c, B3D_ERR_NO_DEVICE, "comm_init: libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    B3D_NCCL(c, nccl().CommInitRank(&comm, world, id, rank));
    c->comm = comm; c->comm_owned = true;
    return B3D_OK;
}

int comm_destroy_impl(b3d_ctx* c) {
    if (c->comm && c->comm_owned && nccl().ok) {
        cudaStreamSynchronize(c->stream);
        nccl().CommDestroy(reinterpret_cast<ncclComm_t>(c->comm));
    }
    c->comm = nullptr; c->comm_owned = false; c->comm_rank = 0; c->comm_world = 1;
    return B3D_OK;
}

static inline void shard(long long total, int rank, int world, long long& lo, long long& hi) {     // contiguous, ordered by rank
    const long long chunk = (total + world - 1) / world;
    lo = chunk * rank < total ? chunk * rank : total;
    hi = lo + chunk < total ? lo + chunk : total;
}

// Registration::ransacRegistration on resident clouds + features (or correspondences when match == 0), collective.
int ransac_sharded_resident_impl(b3d_ctx* c, float voxel, int H, float confidence, int match,
                                 float* T, float* fitness, float* rmse, int32_t* best) {
    if (!c->have_clouds) return fail(c, B3D_ERR_STATE, "ransac_sharded: clouds not set");
    const int world = c->comm_world, rank = c->comm_rank;
    if (world > 1 && !c->comm) return fail(c, B3D_ERR_STATE, "ransac_sharded: no communicator (b3d_comm_init)");
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(c->comm);
    const size_t n_src = c->n_src;
    if (match) {
        long long r0, r1; shard((long long)n_src, rank, world, r0, r1);
        const size_t chunk = (n_src + world - 1) / world;
        B3D_CUDA(c, c->corr.ensure(sizeof(uint32_t) * (chunk * world ? chunk * world : 1)));       // room for the padded gather
        if (r1 > r0) { int rc = b3d_match_features(c, (size_t)r0, (size_t)r1); if (rc != B3D_OK) return rc; }
        else if (!c->have_feats) return fail(c, B3D_ERR_STATE, "ransac_sharded: features not set");
        if (world > 1 && chunk)
            B3D_NCCL(c, nccl().AllGather(c->corr.as<uint32_t>() + chunk * rank, c->corr.p, chunk, ncclUint32, comm, c->stream));
        c->have_corr = true;
    }
    int rc = ransac_prepare_impl(c, voxel, H, confidence); if (rc != B3D_OK) return rc;
    long long h0, h1; shard(H, rank, world, h0, h1);
    rc = b3d_ransac_score(c, (int)h0, (int)h1); if (rc != B3D_OK) return rc;
    DeviceState* st = c->state.as<DeviceState>();
    B3D_CUDA(c, c->dist_keys.ensure(sizeof(unsigned long long) * 3 * (size_t)(world + 1)));
    unsigned long long* all = c->dist_keys.as<unsigned long long>();
    unsigned long long* mine = all + 3 * (size_t)rank;
    rc = ransac_reduce3_impl(c, (int)h0, (int)h1, mine); if (rc != B3D_OK) return rc;
    if (world > 1) B3D_NCCL(c, nccl().AllGather(mine, all, 3, ncclUint64, comm, c->stream));
    {
        StageTimer timer(c, 3);
        resolve_keys_kernel<<<1, 32, 0, c->stream>>>(all, world, &st->best_key);
        B3D_LAUNCHED(c);
    }
    return ransac_finish_impl(c, reinterpret_cast<const int64_t*>(&st->best_key), T, fitness, rmse, best);
}

// Host buffers in: every rank uploads the clouds and the target descriptors, but only ITS rows of the source descriptors.
int ransac_sharded_impl(b3d_ctx* c, const float* src_xyz, size_t n_src, const float* tgt_xyz, size_t n_tgt, const float* src_desc,
                        const float* tgt_desc, float voxel, int H, float confidence, float* T, float* fitness, float* rmse, int32_t* best) {
    int rc = b3d_set_clouds(c, src_xyz, n_src, tgt_xyz, nullptr, n_tgt, 0); if (rc != B3D_OK) return rc;
    if ((n_src && !src_desc) || (n_tgt && !tgt_desc)) return fail(c, B3D_ERR_INVALID, "ransac_sharded: null descriptor pointer");
    long long r0, r1; shard((long long)n_src, c->comm_rank, c->comm_world, r0, r1);
    const size_t sb = sizeof(float) * kDescDim * n_src, tb = sizeof(float) * kDescDim * n_tgt;
    B3D_CUDA(c, c->sdesc.ensure(sb ? sb : 4)); B3D_CUDA(c, c->tdesc.ensure(tb ? tb : 4));
    if (r1 > r0)
        B3D_CUDA(c, cudaMemcpyAsync(c->sdesc.as<float>() + (size_t)kDescDim * (size_t)r0, src_desc + (size_t)kDescDim * (size_t)r0,
                                    sizeof(float) * kDescDim * (size_t)(r1 - r0), cudaMemcpyHostToDevice, c->stream));
    if (tb) B3D_CUDA(c, cudaMemcpyAsync(c->tdesc.p, tgt_desc, tb, cudaMemcpyHostToDevice, c->stream));
    c->sdesc_p = c->sdesc.as<float>(); c->tdesc_p = c->tdesc.as<float>();
    c->have_feats = true;                                   // rows outside [r0, r1) of sdesc are never read by this rank
    return ransac_sharded_resident_impl(c, voxel, H, confidence, 1, T, fitness, rmse, best);
}

}  // namespace b3d

using namespace b3d;

extern "C" {

int b3d_comm_unique_id(void* out_id128) {
    if (!out_id128) return B3D_ERR_INVALID;
    if (!nccl().ok) return B3D_ERR_NO_DEVICE;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return B3D_ERR_CUDA;
    memcpy(out_id128, &id, sizeof(id));
    return B3D_OK;
}

int b3d_comm_init(b3d_ctx* c, const void* id128, int rank, int world) {
    if (!c) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return comm_init_impl(c, id128, rank, world);
}

int b3d_comm_attach(b3d_ctx* c, void* nccl_comm, int rank, int world) {
    if (!c || world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_comm)) return B3D_ERR_INVALID;
    if (c->comm) return fail(c, B3D_ERR_STATE, "comm_attach: the context already has a communicator");
    if (world > 1 && !nccl().ok) return fail(c, B3D_ERR_NO_DEVICE, "comm_attach: libnccl.so.2 could not be loaded");
    c->comm = world > 1 ? nccl_comm : nullptr; c->comm_owned = false; c->comm_rank = rank; c->comm_world = world;
    return B3D_OK;
}

int b3d_comm_destroy(b3d_ctx* c) {
    if (!c) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return comm_destroy_impl(c);
}

int b3d_ransac_sharded(b3d_ctx* c, const float* src_xyz, size_t n_src, const float* tgt_xyz, size_t n_tgt, const float* src_desc,
                       const float* tgt_desc, float voxel_size, int max_iterations, float confidence,
                       float* out_T, float* out_fitness, float* out_rmse, int32_t* out_best) {
    if (!c || !out_T || !out_fitness || !out_rmse) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return ransac_sharded_impl(c, src_xyz, n_src, tgt_xyz, n_tgt, src_desc, tgt_desc, voxel_size, max_iterations, confidence,
                               out_T, out_fitness, out_rmse, out_best);
}

int b3d_register_scene_sharded(b3d_ctx* c, const float* scene_xyz, size_t n, float voxel_size, int normals_k, float fpfh_radius,
                               int ransac_max_iterations, float ransac_confidence, float icp_distance_threshold, int icp_max_iterations,
                               int point_to_plane, b3d_scene_result* out) {
    if (!c || !out || (n && !scene_xyz)) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    if (c->comm_world > 1 && !c->comm) return fail(c, B3D_ERR_STATE, "register_scene_sharded: no communicator (b3d_comm_init)");
    return register_scene_sharded_impl(c, scene_xyz, n, voxel_size, normals_k, fpfh_radius, ransac_max_iterations, ransac_confidence,
                                       icp_distance_threshold, icp_max_iterations, point_to_plane, out);
}

int b3d_ransac_sharded_resident(b3d_ctx* c, float voxel_size, int max_iterations, float confidence, int match_features,
                                float* out_T, float* out_fitness, float* out_rmse, int32_t* out_best) {
    if (!c || !out_T || !out_fitness || !out_rmse) return B3D_ERR_INVALID;
    B3D_CUDA(c, enter(c));
    return ransac_sharded_resident_impl(c, voxel_size, max_iterations, confidence, match_features, out_T, out_fitness, out_rmse, out_best);
}

}  // extern "C"
