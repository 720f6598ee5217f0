// b3d_pool.cu — the orchestrator's worker pool over independent instances (SURVEY.md §8e "batched multi-object").
// The reference pushes one processInstance task per detected object into a pool of num_threads workers
// (src/pipeline.cpp:16, 321-327; include/thread_pool.hpp:14-80); each task runs ransacRegistration and then icpRefine on
// its own clouds (src/pipeline.cpp:97-129).  This is that pool behind the C-ABI: persistent host threads, each owning one
// b3d_ctx (stream + workspace) on one of the given devices (worker w -> devices[w % n_devices], so one pool can span all
// the GPUs of a box), pulling instances off a shared counter.  Instances are independent: no collective, no ordering
// between them; results land in the caller's array at the instance's index.  Host-side plumbing only — every
// computation is the same kernels b3d_ransac / b3d_icp launch.
#include "b3d_common.cuh"
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

struct b3d_pool {
    std::vector<std::thread> threads;
    std::vector<b3d_ctx*> ctxs;
    std::mutex m;
    std::condition_variable wake, done_cv;
    const b3d_instance* items = nullptr;
    b3d_instance_result* results = nullptr;
    size_t n = 0;
    std::atomic<size_t> next{0};
    size_t finished_workers = 0;
    unsigned long long generation = 0;
    bool stop = false;
};

namespace b3d {

static void run_instance(b3d_ctx* c, const b3d_instance& in, b3d_instance_result& out) {
    for (int i = 0; i < 16; ++i) out.coarse_T[i] = out.T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    out.coarse_fitness = out.coarse_rmse = out.fitness = out.rmse = 0.0f;
    out.icp_iterations = 0;
    int rc = b3d_ransac(c, in.src_xyz, in.n_src, in.tgt_xyz, in.n_tgt, in.src_desc, in.tgt_desc, in.voxel_size, in.ransac_max_iterations,
                        in.ransac_confidence, out.coarse_T, &out.coarse_fitness, &out.coarse_rmse, nullptr);                  // pipeline.cpp:97-102
    if (rc == B3D_OK)
        rc = b3d_icp(c, in.src_xyz, in.n_src, in.tgt_xyz, in.tgt_normals_or_null, in.n_tgt, out.coarse_T, in.icp_distance_threshold,
                     in.icp_max_iterations, in.point_to_plane, out.T, &out.fitness, &out.rmse, &out.icp_iterations);        // pipeline.cpp:104-128
    out.status = rc;
}

static void worker_main(b3d_pool* p, size_t w) {
    unsigned long long seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(p->m);
            p->wake.wait(lk, [&] { return p->stop || p->generation != seen; });
            if (p->stop) return;
            seen = p->generation;
        }
        for (size_t i = p->next.fetch_add(1); i < p->n; i = p->next.fetch_add(1)) run_instance(p->ctxs[w], p->items[i], p->results[i]);
        {
            std::lock_guard<std::mutex> lk(p->m);
            if (++p->finished_workers == p->threads.size()) p->done_cv.notify_all();
        }
    }
}

}  // namespace b3d

extern "C" {

int b3d_pool_create(int n_workers, const int* devices, int n_devices, b3d_pool** out) {
    if (!out || n_workers < 1 || n_workers > 256 || (n_devices > 0 && !devices)) return B3D_ERR_INVALID;
    *out = nullptr;
    b3d_pool* p = new (std::nothrow) b3d_pool();
    if (!p) return B3D_ERR_ALLOC;
    for (int w = 0; w < n_workers; ++w) {
        b3d_ctx* c = nullptr;
        const int dev = n_devices > 0 ? devices[w % n_devices] : 0;
        const int rc = b3d_ctx_create(dev, &c);
        if (rc != B3D_OK) { for (b3d_ctx* x : p->ctxs) b3d_ctx_destroy(x); delete p; return rc; }
        p->ctxs.push_back(c);
    }
    for (int w = 0; w < n_workers; ++w) p->threads.emplace_back(b3d::worker_main, p, (size_t)w);
    *out = p;
    return B3D_OK;
}

int b3d_pool_register(b3d_pool* p, const b3d_instance* items, size_t n, b3d_instance_result* results) {
    if (!p || (n && (!items || !results))) return B3D_ERR_INVALID;
    if (n == 0) return B3D_OK;
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->items = items; p->results = results; p->n = n;
        p->next.store(0); p->finished_workers = 0;
        ++p->generation;
    }
    p->wake.notify_all();
    {
        std::unique_lock<std::mutex> lk(p->m);
        p->done_cv.wait(lk, [&] { return p->finished_workers == p->threads.size(); });
    }
    for (size_t i = 0; i < n; ++i) if (results[i].status != B3D_OK) return results[i].status;
    return B3D_OK;
}

void b3d_pool_destroy(b3d_pool* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->stop = true;
    }
    p->wake.notify_all();
    for (std::thread& t : p->threads) t.join();
    for (b3d_ctx* c : p->ctxs) b3d_ctx_destroy(c);
    delete p;
}

}  // extern "C"
