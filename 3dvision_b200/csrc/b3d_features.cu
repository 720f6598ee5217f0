// b3d_features.cu — the stages that feed the registration hot path:
//   Registration::voxelDownsample   src/registration.cpp:29-60
//   Registration::estimateNormals   src/registration.cpp:63-81, 105-130   (k nearest by (d2, index), covariance, eigenvector)
//   Registration::computeFPFH       src/registration.cpp:83-102, 133-201  (radius neighbours capped at 100, SPFH, weighted FPFH)
// The reference does all three with O(N^2) scans on one thread; here neighbour queries run on the voxel-hash grid of
// b3d_grid.cuh, one warp per query, and every floating-point reduction whose order the result depends on is replayed
// in the reference's order, so the outputs are bit-identical to the CPU path (tests/test_gpu_features.py).
//
// Neighbour selection.  The reference orders candidates by the pair (d2, index) (std::pair operator<).  d2 >= 0, so
// the 64-bit key (float bits of d2) << 32 | index has the same order.  A warp streams candidate cells, keeps keys below
// the current cut in a shared-memory buffer and bitonic-sorts the buffer whenever it fills; what survives is the exact
// prefix of the reference's sorted list.  A block of cells of Chebyshev radius rho around the query's cell provably
// contains every point closer than cover(rho) = rho*cell - 2*slack (slack bounds the rounding of the cell
// assignment), so a k-NN list is final once its last d2 is below cover^2 (shrunk by 2^-18 for the rounding of d2
// itself), and a radius query is final after rho = 1 when radius^2 is.  Queries the grid cannot settle (isolated
// points, k >= n, a radius tiny against the coordinates) fall through to an exact scan of all points by the same warp.
//
// Voxel down-sampling.  Keys floor(p * (1/voxel)) are packed to 63 bits and radix-sorted with the point index
// (cub::DeviceRadixSort — stable, so each voxel's members stay in input order and the running sum per voxel adds
// them exactly as the reference's `for idx in indices` does).  The OUTPUT ORDER of the reference is the iteration
// order of its std::unordered_map<VoxelKey,...> (registration.cpp:44); RANSAC later samples by index, so the order
// is part of the result.  It depends only on the sequence of first insertions, so the distinct keys are brought to
// the host in first-appearance order and pushed through the same libstdc++ container with the same hash.
#include "b3d_common.cuh"
#include "b3d_grid.cuh"
#include "b3d_featmath.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <unordered_map>
#include <vector>

namespace b3d {

// fbuf slots
enum { F_PTS4 = 0, F_NRM4, F_SLOTS, F_GP, F_SORTED, F_PT_SLOT, F_PT_RANK, F_COV, F_OUT, F_NBR, F_NBR_CNT, F_SPFH,
       F_KEYS_A, F_KEYS_B, F_IDX_A, F_IDX_B, F_CUB, F_SEG, F_VOX_MEAN, F_VOX_COL, F_VOX_KEY, F_VOX_FIRST, F_VOX_ORDER, F_PERM, F_RAW, F_RAW2, F_ORD_OPEN, F_ORD_KEYS, F_ORD_SEQ, F_PIPE_NRM, F_IMG_DEPTH, F_IMG_MASK, F_IMG_BGR, F_IMG_XYZ, F_IMG_RGB, F_COUNT };
static_assert(F_COUNT <= b3d_ctx::kFeatureBufs, "grow b3d_ctx::fbuf");

constexpr int kFeatWarps = 8;                 // warps (= queries in flight) per block
constexpr int kKeyBuf = 256;                  // per-warp key buffer: sorted prefix + staged candidates
constexpr int kMaxList = 128;                 // longest neighbour list kept (k <= 128; FPFH keeps 100)
constexpr unsigned long long kNoKey = ~0ull;
constexpr unsigned kFpfhMaxNn = 100;          // computeFPFH keeps at most 100 neighbours (registration.cpp:139)

// ---------------------------------------------------------------------------------
// warp-level exact top-K by (d2, index)
// ---------------------------------------------------------------------------------
struct WarpList {
    unsigned long long* keys;                 // shared, kKeyBuf entries
    unsigned n_sorted, n_staged, K;
    unsigned long long tau;                   // keys >= tau can no longer enter the list
};

__device__ __forceinline__ void warp_bitonic_sort(unsigned long long* keys, unsigned P, unsigned lane) {
    for (unsigned k = 2; k <= P; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned t = lane; t < (P >> 1); t += 32) {
                const unsigned i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));
                const unsigned p = i | j;
                const unsigned long long a = keys[i], b = keys[p];
                const bool ascending = (i & k) == 0u;
                if ((a > b) == ascending) { keys[i] = b; keys[p] = a; }
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ void list_flush(WarpList& L, unsigned lane) {
    if (L.n_staged == 0u) return;
    const unsigned total = L.n_sorted + L.n_staged;
    unsigned P = 32u;
    while (P < total) P <<= 1;
    for (unsigned i = total + lane; i < P; i += 32) L.keys[i] = kNoKey;
    __syncwarp();
    warp_bitonic_sort(L.keys, P, lane);
    L.n_sorted = min(total, L.K);
    L.n_staged = 0u;
    if (L.n_sorted == L.K) { const unsigned long long last = L.keys[L.K - 1u]; if (last < L.tau) L.tau = last; }
    __syncwarp();
}

__device__ __forceinline__ void list_push(WarpList& L, bool valid, unsigned long long key, unsigned lane) {
    const bool pass = valid && key < L.tau;
    const unsigned b = __ballot_sync(0xffffffffu, pass);
    if (b) {
        if (pass) L.keys[L.n_sorted + L.n_staged + __popc(b & ((1u << lane) - 1u))] = key;
        L.n_staged += __popc(b);
        if (L.n_sorted + L.n_staged > (unsigned)kKeyBuf - 32u) { __syncwarp(); list_flush(L, lane); }
    }
}

__device__ __forceinline__ unsigned long long make_key(float4 p, float qx, float qy, float qz) {
    const float e0 = p.x - qx, e1 = p.y - qy, e2 = p.z - qz;              // (points[j] - query).squaredNorm()
    const float d2 = e0 * e0 + (e1 * e1 + e2 * e2);
    return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(p.w);
}

// all points of the cells at Chebyshev distance exactly `rho` (rho == 1: the whole 3x3x3 block) from (cx, cy, cz)
__device__ __forceinline__ void visit_shell(WarpList& L, const GridView& g, int cx, int cy, int cz, int rho,
                                            float qx, float qy, float qz, unsigned lane) {
    const int side = 2 * rho + 1, cells = side * side * side;
    for (int base = 0; base < cells; base += 32) {
        const int t = base + (int)lane;
        unsigned start = 0, count = 0;
        if (t < cells) {
            const int dx = t % side - rho, dy = (t / side) % side - rho, dz = t / (side * side) - rho;
            const int m = max(abs(dx), max(abs(dy), abs(dz)));
            if (m == rho || rho == 1) {
                const unsigned long long key = pack_cell(cx + dx, cy + dy, cz + dz);
                unsigned slot = hash_cell(key) & g.mask;
                while (true) {
                    const CellSlot s = g.slots[slot];
                    if (s.key == key) { start = s.start; count = s.count; break; }
                    if (s.key == kEmptyKey) break;
                    slot = (slot + 1u) & g.mask;
                }
            }
        }
        unsigned live = __ballot_sync(0xffffffffu, count != 0u);
        while (live) {
            const int src = __ffs(live) - 1;
            live &= live - 1u;
            const unsigned st = __shfl_sync(0xffffffffu, start, src), cnt = __shfl_sync(0xffffffffu, count, src);
            for (unsigned off = 0; off < cnt; off += 32) {
                const bool valid = off + lane < cnt;
                unsigned long long key = kNoKey;
                if (valid) key = make_key(g.pts[st + off + lane], qx, qy, qz);
                list_push(L, valid, key, lane);
            }
        }
    }
}

__device__ __forceinline__ void visit_all(WarpList& L, const float4* __restrict__ pts, unsigned n, float qx, float qy, float qz,
                                          unsigned lane) {
    for (unsigned off = 0; off < n; off += 32) {
        const bool valid = off + lane < n;
        unsigned long long key = kNoKey;
        if (valid) key = make_key(pts[off + lane], qx, qy, qz);
        list_push(L, valid, key, lane);
    }
}

// squared radius below which every point is guaranteed to lie inside the visited block of Chebyshev radius rho
__device__ __forceinline__ float covered2(const GridView& g, int rho) {
    const float cover = (float)rho * g.cell - 2.0f * g.slack;
    return cover > 0.0f ? cover * cover * (1.0f - 3.8146973e-6f) : 0.0f;          // 2^-18
}

constexpr int kMaxRing = 3;

// sorted (d2, index) keys of the K nearest points of query i (self included) end up in L.keys[0 .. n_sorted)
__device__ __forceinline__ void knn_query(WarpList& L, const GridView& g, unsigned n, float4 q, unsigned K, unsigned lane) {
    L.n_sorted = 0u; L.n_staged = 0u; L.K = K; L.tau = kNoKey;
    const int cx = cell_coord(q.x, g.inv), cy = cell_coord(q.y, g.inv), cz = cell_coord(q.z, g.inv);
    bool settled = false;
    if (K < n) {
        for (int rho = 1; rho <= kMaxRing && !settled; ++rho) {
            visit_shell(L, g, cx, cy, cz, rho, q.x, q.y, q.z, lane);
            list_flush(L, lane);
            if (L.n_sorted == K) {
                const float dk = __uint_as_float((unsigned)(L.keys[K - 1u] >> 32));
                settled = dk < covered2(g, rho);
            }
        }
    }
    if (!settled) {                                                               // exact scan of every point
        __syncwarp();
        L.n_sorted = 0u; L.n_staged = 0u; L.tau = kNoKey;
        visit_all(L, g.pts, n, q.x, q.y, q.z, lane);
        list_flush(L, lane);
    }
}

// ---------------------------------------------------------------------------------
// estimateNormals
// ---------------------------------------------------------------------------------
// centroid (3 running sums) and covariance (9 running sums) over the neighbours in list order, one lane per sum
__device__ __forceinline__ void covariance_in_list_order(const float (*nb)[kMaxList], unsigned kk, unsigned lane, float* __restrict__ cov_out) {
    const float nf = (float)kk;
    float c = 0.0f;
    if (lane < 3) { const float* v = nb[lane]; for (unsigned a = 0; a < kk; ++a) c += v[a]; c /= nf; }   // registration.cpp:111-113
    const float c0 = __shfl_sync(0xffffffffu, c, 0), c1 = __shfl_sync(0xffffffffu, c, 1), c2 = __shfl_sync(0xffffffffu, c, 2);
    if (lane < 9) {                                                                // registration.cpp:115-120
        const unsigned r = lane / 3u, cc = lane % 3u;
        const float mr = r == 0 ? c0 : (r == 1 ? c1 : c2), mc = cc == 0 ? c0 : (cc == 1 ? c1 : c2);
        const float* vr = nb[r]; const float* vc = nb[cc];
        float s = 0.0f;
        for (unsigned a = 0; a < kk; ++a) { const float dr = vr[a] - mr, dc = vc[a] - mc; s += dr * dc; }
        cov_out[lane] = s / nf;
    }
}

__global__ void __launch_bounds__(kFeatWarps * 32)
knn_covariance_kernel(const float4* __restrict__ pts, unsigned n, unsigned K, const CellSlot* __restrict__ slots,
                      const float4* __restrict__ sorted, const GridParams* __restrict__ gp, float* __restrict__ cov9) {
    __shared__ unsigned long long s_keys[kFeatWarps][kKeyBuf];
    __shared__ float s_nb[kFeatWarps][3][kMaxList];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const GridView g = make_view(slots, sorted, gp);
    WarpList L; L.keys = s_keys[warp];
    for (unsigned i = blockIdx.x * kFeatWarps + warp; i < n; i += gridDim.x * kFeatWarps) {
        const float4 q = pts[i];
        knn_query(L, g, n, q, K, lane);
        const unsigned kk = L.n_sorted;                                            // min(k, n)
        for (unsigned a = lane; a < kk; a += 32) {
            const float4 p = pts[(unsigned)L.keys[a]];
            s_nb[warp][0][a] = p.x; s_nb[warp][1][a] = p.y; s_nb[warp][2][a] = p.z;
        }
        __syncwarp();
        covariance_in_list_order(s_nb[warp], kk, lane, cov9 + (size_t)i * 9);
        __syncwarp();
    }
}

// Fused path: the radius lists computeFPFH needs anyway are sorted by (d2, index) and hold every point within the
// radius (up to 100), so whenever a list has at least k entries its first k ARE the k nearest neighbours; only points
// with fewer than k neighbours inside the radius run their own k-NN query (on the same grid).
__global__ void __launch_bounds__(kFeatWarps * 32)
list_covariance_kernel(const float4* __restrict__ pts, unsigned n, unsigned K, const unsigned* __restrict__ nbr,
                       const unsigned* __restrict__ nbr_cnt, const CellSlot* __restrict__ slots, const float4* __restrict__ sorted,
                       const GridParams* __restrict__ gp, float* __restrict__ cov9) {
    __shared__ unsigned long long s_keys[kFeatWarps][kKeyBuf];
    __shared__ float s_nb[kFeatWarps][3][kMaxList];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const GridView g = make_view(slots, sorted, gp);
    WarpList L; L.keys = s_keys[warp];
    for (unsigned i = blockIdx.x * kFeatWarps + warp; i < n; i += gridDim.x * kFeatWarps) {
        const unsigned cnt = nbr_cnt[i];
        unsigned kk;
        if (cnt >= K) {
            kk = K;
            for (unsigned a = lane; a < kk; a += 32) {
                const float4 p = pts[nbr[(size_t)i * kFpfhMaxNn + a]];
                s_nb[warp][0][a] = p.x; s_nb[warp][1][a] = p.y; s_nb[warp][2][a] = p.z;
            }
        } else {
            knn_query(L, g, n, pts[i], K, lane);
            kk = L.n_sorted;
            for (unsigned a = lane; a < kk; a += 32) {
                const float4 p = pts[(unsigned)L.keys[a]];
                s_nb[warp][0][a] = p.x; s_nb[warp][1][a] = p.y; s_nb[warp][2][a] = p.z;
            }
        }
        __syncwarp();
        covariance_in_list_order(s_nb[warp], kk, lane, cov9 + (size_t)i * 9);
        __syncwarp();
    }
}

__global__ void normal_from_covariance_kernel(const float4* __restrict__ pts, const float* __restrict__ cov9, unsigned n,
                                              float* __restrict__ out_xyz) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* c = cov9 + (size_t)i * 9;
    float ev[3]; Mat3 Q;
    sym_eig3(c[0], c[3], c[4], c[6], c[7], c[8], ev, Q);                           // registration.cpp:122-123
    float nx = Q(0, 0), ny = Q(1, 0), nz = Q(2, 0);
    const float4 p = pts[i];
    const float d = nx * (-p.x) + (ny * (-p.y) + nz * (-p.z));                     // normals[i].dot(-points[i])
    if (d < 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
    out_xyz[3 * (size_t)i] = nx; out_xyz[3 * (size_t)i + 1] = ny; out_xyz[3 * (size_t)i + 2] = nz;
}

// ---------------------------------------------------------------------------------
// computeFPFH
// ---------------------------------------------------------------------------------

__global__ void __launch_bounds__(kFeatWarps * 32)
radius_neighbors_kernel(const float4* __restrict__ pts, unsigned n, float r2, const CellSlot* __restrict__ slots,
                        const float4* __restrict__ sorted, const GridParams* __restrict__ gp,
                        unsigned* __restrict__ nbr, unsigned* __restrict__ nbr_cnt) {
    __shared__ unsigned long long s_keys[kFeatWarps][kKeyBuf];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const GridView g = make_view(slots, sorted, gp);
    const bool grid_ok = r2 < covered2(g, 1);
    const unsigned long long tau0 = (unsigned long long)(__float_as_uint(r2) + 1u) << 32;    // d2 <= r2  (registration.cpp:93)
    WarpList L; L.keys = s_keys[warp];
    for (unsigned i = blockIdx.x * kFeatWarps + warp; i < n; i += gridDim.x * kFeatWarps) {
        const float4 q = pts[i];
        L.n_sorted = 0u; L.n_staged = 0u; L.K = kFpfhMaxNn; L.tau = tau0;
        if (grid_ok) visit_shell(L, g, cell_coord(q.x, g.inv), cell_coord(q.y, g.inv), cell_coord(q.z, g.inv), 1, q.x, q.y, q.z, lane);
        else         visit_all(L, sorted, n, q.x, q.y, q.z, lane);
        list_flush(L, lane);
        for (unsigned a = lane; a < L.n_sorted; a += 32) nbr[(size_t)i * kFpfhMaxNn + a] = (unsigned)L.keys[a];
        if (lane == 0) nbr_cnt[i] = L.n_sorted;
        __syncwarp();
    }
}

// registration.cpp:144-196: simplified point feature histogram of every point over its neighbour list.
// The bins are counters (+= 1.0f), so their values do not depend on the order of the neighbours.
__global__ void __launch_bounds__(kFeatWarps * 32)
spfh_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm, unsigned n, const unsigned* __restrict__ nbr,
            const unsigned* __restrict__ nbr_cnt, float* __restrict__ spfh) {
    __shared__ int s_hist[kFeatWarps][33];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (unsigned i = blockIdx.x * kFeatWarps + warp; i < n; i += gridDim.x * kFeatWarps) {
        s_hist[warp][lane] = 0;
        if (lane == 0) s_hist[warp][32] = 0;
        __syncwarp();
        const float4 p = pts[i], u = nrm[i];
        const unsigned cnt = nbr_cnt[i];
        for (unsigned a = lane; a < cnt; a += 32) {
            const unsigned ni = nbr[(size_t)i * kFpfhMaxNn + a];
            if (ni == i) continue;
            const float4 pn = pts[ni], nn = nrm[ni];
            const float d0 = pn.x - p.x, d1 = pn.y - p.y, d2 = pn.z - p.z;
            const float dist = sqrtf(d0 * d0 + (d1 * d1 + d2 * d2));                 // diff.norm()
            if (dist < 1e-8f) continue;
            const float e0 = d0 / dist, e1 = d1 / dist, e2 = d2 / dist;            // dn = diff / dist
            const float v0 = u.y * e2 - u.z * e1, v1 = u.z * e0 - u.x * e2, v2 = u.x * e1 - u.y * e0;      // v = u x dn
            const float w0 = u.y * v2 - u.z * v1, w1 = u.z * v0 - u.x * v2, w2 = u.x * v1 - u.y * v0;      // w = u x v
            const float alpha = v0 * nn.x + (v1 * nn.y + v2 * nn.z);
            const float phi = u.x * e0 + (u.y * e1 + u.z * e2);
            const float theta = atan2_libm(w0 * nn.x + (w1 * nn.y + w2 * nn.z), u.x * nn.x + (u.y * nn.y + u.z * nn.z));
            const int ba = min(max((int)((alpha + 1.0f) * 5.5f), 0), 10);
            const int bp = min(max((int)((phi + 1.0f) * 5.5f), 0), 10);
            const int bt = min(max((int)(((double)theta / 3.14159265358979323846 + 1.0) * 5.5), 0), 10);   // M_PI is a double
            atomicAdd(&s_hist[warp][ba], 1); atomicAdd(&s_hist[warp][11 + bp], 1); atomicAdd(&s_hist[warp][22 + bt], 1);
        }
        __syncwarp();
        int tot = s_hist[warp][lane];
        if (lane == 0) tot += s_hist[warp][32];
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
        const float sum = (float)tot;                                               // integers: any summation order gives this
        float* out = spfh + (size_t)i * 33;
        const float h = (float)s_hist[warp][lane];
        out[lane] = tot > 0 ? h / sum : h;
        if (lane == 0) { const float h32 = (float)s_hist[warp][32]; out[32] = tot > 0 ? h32 / sum : h32; }
        __syncwarp();
    }
}

// registration.cpp:185-199: f = spfh[i] + sum over neighbours (in list order) of (1/dist) * spfh[neighbour]; L1-normalise.
// Lane d owns bin d (lane 0 also bin 32) and adds left to right, as the reference's inner loop does.
__global__ void __launch_bounds__(kFeatWarps * 32)
fpfh_kernel(const float4* __restrict__ pts, unsigned n, const unsigned* __restrict__ nbr, const unsigned* __restrict__ nbr_cnt,
            const float* __restrict__ spfh, float* __restrict__ desc) {
    __shared__ float s_w[kFeatWarps][kMaxList];
    __shared__ unsigned s_ni[kFeatWarps][kMaxList];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (unsigned i = blockIdx.x * kFeatWarps + warp; i < n; i += gridDim.x * kFeatWarps) {
        const float4 p = pts[i];
        const unsigned cnt = nbr_cnt[i];
        for (unsigned a = lane; a < cnt; a += 32) {
            unsigned ni = nbr[(size_t)i * kFpfhMaxNn + a];
            float w = 0.0f;
            if (ni == i) ni = 0xFFFFFFFFu;
            else {
                const float4 pn = pts[ni];
                const float d0 = pn.x - p.x, d1 = pn.y - p.y, d2 = pn.z - p.z;
                const float dist = sqrtf(d0 * d0 + (d1 * d1 + d2 * d2));
                if (dist < 1e-8f) ni = 0xFFFFFFFFu; else w = 1.0f / dist;
            }
            s_w[warp][a] = w; s_ni[warp][a] = ni;
        }
        __syncwarp();
        float f = spfh[(size_t)i * 33 + lane];
        float f32 = lane == 0 ? spfh[(size_t)i * 33 + 32] : 0.0f;
        for (unsigned a = 0; a < cnt; ++a) {
            const unsigned ni = s_ni[warp][a];
            if (ni == 0xFFFFFFFFu) continue;
            const float w = s_w[warp][a];
            const float* row = spfh + (size_t)ni * 33;
            const float t = w * row[lane];
            f += t;
            if (lane == 0) { const float t2 = w * row[32]; f32 += t2; }
        }
        float sum = 0.0f;
#pragma unroll
        for (int d = 0; d < 32; ++d) sum += __shfl_sync(0xffffffffu, f, d);
        sum += __shfl_sync(0xffffffffu, f32, 0);
        float* out = desc + (size_t)i * 33;
        out[lane] = sum > 0.0f ? f / sum : f;
        if (lane == 0) out[32] = sum > 0.0f ? f32 / sum : f32;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// voxelDownsample
// ---------------------------------------------------------------------------------
constexpr int kVoxBias = 1 << 20;

__global__ void voxel_key_kernel(const float* __restrict__ xyz, unsigned n, float inv, unsigned long long* __restrict__ keys,
                                 unsigned* __restrict__ idx, unsigned* __restrict__ out_of_range) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float fx = floorf(xyz[3 * (size_t)i] * inv), fy = floorf(xyz[3 * (size_t)i + 1] * inv), fz = floorf(xyz[3 * (size_t)i + 2] * inv);
    const float lim = (float)(kVoxBias - 1);
    if (!(fabsf(fx) < lim && fabsf(fy) < lim && fabsf(fz) < lim)) { atomicExch(out_of_range, 1u); keys[i] = 0ull; idx[i] = i; return; }
    const int kx = (int)fx, ky = (int)fy, kz = (int)fz;
    keys[i] = ((unsigned long long)(unsigned)(kx + kVoxBias) << 42) | ((unsigned long long)(unsigned)(ky + kVoxBias) << 21) |
              (unsigned long long)(unsigned)(kz + kVoxBias);
    idx[i] = i;
}

struct SegHead { const unsigned long long* k; __device__ unsigned operator()(unsigned i) const { return (i == 0u || k[i] != k[i - 1u]) ? 1u : 0u; } };
struct SegStart { unsigned* start; __device__ void operator()(unsigned i, unsigned prefix, unsigned flag) const { if (flag) start[prefix] = i; } };

// one thread per voxel: running sums over its members in input order (registration.cpp:48-55)
__global__ void voxel_mean_kernel(const float* __restrict__ xyz, const float* __restrict__ colors, unsigned n,
                                  const unsigned long long* __restrict__ keys, const unsigned* __restrict__ idx,
                                  const unsigned* __restrict__ seg_start, const unsigned* __restrict__ n_vox_ptr,
                                  float* __restrict__ mean, float* __restrict__ mean_col, int* __restrict__ key3,
                                  unsigned* __restrict__ first, unsigned* __restrict__ vox_id) {
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned m = *n_vox_ptr;
    if (v >= m) return;
    const unsigned s = seg_start[v], e = (v + 1u < m) ? seg_start[v + 1u] : n;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
    for (unsigned t = s; t < e; ++t) {
        const size_t j = idx[t];
        a0 += xyz[3 * j]; a1 += xyz[3 * j + 1]; a2 += xyz[3 * j + 2];
        if (colors) { c0 += colors[3 * j]; c1 += colors[3 * j + 1]; c2 += colors[3 * j + 2]; }
    }
    const float cnt = (float)(e - s);
    mean[3 * (size_t)v] = a0 / cnt; mean[3 * (size_t)v + 1] = a1 / cnt; mean[3 * (size_t)v + 2] = a2 / cnt;
    if (colors) { mean_col[3 * (size_t)v] = c0 / cnt; mean_col[3 * (size_t)v + 1] = c1 / cnt; mean_col[3 * (size_t)v + 2] = c2 / cnt; }
    const unsigned long long k = keys[s];
    key3[3 * (size_t)v] = (int)((k >> 42) & 0x1FFFFFull) - kVoxBias;
    key3[3 * (size_t)v + 1] = (int)((k >> 21) & 0x1FFFFFull) - kVoxBias;
    key3[3 * (size_t)v + 2] = (int)(k & 0x1FFFFFull) - kVoxBias;
    first[v] = idx[s];                              // smallest member index = first appearance (stable sort)
    vox_id[v] = v;
}

__global__ void gather_keys_kernel(const int* __restrict__ key3, const unsigned* __restrict__ order, unsigned m, int* __restrict__ out) {
    const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const size_t v = order[r];
    out[3 * (size_t)r] = key3[3 * v]; out[3 * (size_t)r + 1] = key3[3 * v + 1]; out[3 * (size_t)r + 2] = key3[3 * v + 2];
}

__global__ void gather_voxels_kernel(const float* __restrict__ mean, const float* __restrict__ mean_col, const unsigned* __restrict__ order,
                                     const unsigned* __restrict__ perm, unsigned m, float* __restrict__ out, float* __restrict__ out_col) {
    const unsigned pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= m) return;
    const size_t v = order[perm[pos]];
    out[3 * (size_t)pos] = mean[3 * v]; out[3 * (size_t)pos + 1] = mean[3 * v + 1]; out[3 * (size_t)pos + 2] = mean[3 * v + 2];
    if (mean_col) { out_col[3 * (size_t)pos] = mean_col[3 * v]; out_col[3 * (size_t)pos + 1] = mean_col[3 * v + 1]; out_col[3 * (size_t)pos + 2] = mean_col[3 * v + 2]; }
}

// ---- iteration order of libstdc++'s unordered_map, computed on the device --------------------------------------
// A node enters the container's singly linked list either right behind the "before" node of its bucket (bucket
// already in use: it becomes the first node of that bucket's run) or at the global head (bucket empty).  A rehash
// re-inserts every node, in current list order, by the same rule into the new bucket array.  So after a rehash to
// B buckets at element count n_p, the list is what inserting S = (list order before the rehash) ++ (the elements that
// arrive until the next rehash) into an empty B-bucket table gives, and that is: runs ordered by the time their
// bucket was first used, latest first; inside a run, latest first — i.e. S sorted descending by (first-use time of
// the element's bucket, own time).  The rehash points and bucket counts come from libstdc++'s own
// _Prime_rehash_policy object (the one unordered_map uses), so growth matches whatever libstdc++ is linked.
struct RehashStep { unsigned at; unsigned buckets; };

static void rehash_schedule(unsigned m, std::vector<RehashStep>& out) {
    std::__detail::_Prime_rehash_policy policy;                                     // max_load_factor 1.0, like the reference's map
    size_t buckets = 1, count = 0;                                                  // a default-constructed map has one bucket
    while (count < m) {
        const std::pair<bool, size_t> r = policy._M_need_rehash(buckets, count, 1);
        if (r.first) { buckets = r.second; out.push_back({(unsigned)count, (unsigned)buckets}); }
        const size_t quiet_until = policy._M_next_resize;                           // no decision changes while count + 1 <= this
        count = quiet_until > count + 1 ? quiet_until : count + 1;
    }
}

__device__ __forceinline__ unsigned long long voxel_hash(int x, int y, int z) {     // registration.cpp:20-26; std::hash<int> is the identity
    unsigned long long h = (unsigned long long)(long long)x;
    h ^= (unsigned long long)(long long)y + 0x9e3779b9ull + (h << 6) + (h >> 2);
    h ^= (unsigned long long)(long long)z + 0x9e3779b9ull + (h << 6) + (h >> 2);
    return h;
}

__global__ void order_open_kernel(const unsigned* __restrict__ seq_prev, unsigned n_prev, unsigned n_now, const int* __restrict__ key3,
                                  unsigned buckets, unsigned* __restrict__ open, unsigned* __restrict__ bucket_of, unsigned* __restrict__ elem) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_now) return;
    const unsigned r = t < n_prev ? seq_prev[t] : t;                                // new arrivals come in first-appearance order
    const unsigned b = (unsigned)(voxel_hash(key3[3 * (size_t)r], key3[3 * (size_t)r + 1], key3[3 * (size_t)r + 2]) % buckets);
    atomicMin(&open[b], t);
    bucket_of[t] = b; elem[t] = r;
}

// Sort input in DESCENDING own time: a stable descending sort on the bucket's first-use time alone then leaves equal keys in
// that order, i.e. the (first-use time, own time) descending order the container has — with half the key bits to sort.
__global__ void order_key_kernel(const unsigned* __restrict__ open, const unsigned* __restrict__ bucket_of, const unsigned* __restrict__ elem,
                                 unsigned n_now, unsigned* __restrict__ keys, unsigned* __restrict__ vals) {
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_now) return;
    const unsigned t = n_now - 1u - p;
    keys[p] = open[bucket_of[t]];
    vals[p] = elem[t];
}

// perm[pos] = first-appearance rank of the voxel the reference emits at position pos
static int container_order_device(b3d_ctx* c, const int* key3_ordered, unsigned m, unsigned* perm) {
    std::vector<RehashStep> steps;
    rehash_schedule(m, steps);
    const unsigned max_buckets = steps.empty() ? 1u : steps.back().buckets;
    B3D_CUDA(c, c->fbuf[F_ORD_OPEN].ensure(sizeof(unsigned) * ((size_t)max_buckets + 2 * (size_t)m)));
    B3D_CUDA(c, c->fbuf[F_ORD_KEYS].ensure(sizeof(unsigned) * 3 * (size_t)m));
    B3D_CUDA(c, c->fbuf[F_ORD_SEQ].ensure(sizeof(unsigned) * 2 * (size_t)m));
    unsigned* open = c->fbuf[F_ORD_OPEN].as<unsigned>();
    unsigned* bucket_of = open + max_buckets; unsigned* elem = bucket_of + m;
    unsigned* keys_in = c->fbuf[F_ORD_KEYS].as<unsigned>(); unsigned* keys_out = keys_in + m; unsigned* vals_in = keys_out + m;
    unsigned* seq[2] = {c->fbuf[F_ORD_SEQ].as<unsigned>(), c->fbuf[F_ORD_SEQ].as<unsigned>() + m};
    int cur = 0;
    unsigned n_prev = 0;
    for (size_t p = 0; p < steps.size(); ++p) {
        const unsigned n_now = p + 1 < steps.size() ? steps[p + 1].at : m;          // elements present when the next rehash (or the end) comes
        const unsigned buckets = steps[p].buckets;
        int bits = 1; while ((1ull << bits) < (unsigned long long)n_now + 1ull) ++bits;
        B3D_CUDA(c, cudaMemsetAsync(open, 0xFF, sizeof(unsigned) * buckets, c->stream));
        order_open_kernel<<<div_up(n_now, 256), 256, 0, c->stream>>>(seq[cur], n_prev, n_now, key3_ordered, buckets, open, bucket_of, elem);
        B3D_LAUNCHED(c);
        order_key_kernel<<<div_up(n_now, 256), 256, 0, c->stream>>>(open, bucket_of, elem, n_now, keys_in, vals_in);
        B3D_LAUNCHED(c);
        unsigned* dst = (p + 1 == steps.size()) ? perm : seq[cur ^ 1];
        size_t tmp_bytes = 0;
        B3D_CUDA(c, cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, keys_in, keys_out, vals_in, dst, (int)n_now, 0, bits, c->stream));
        B3D_CUDA(c, c->fbuf[F_CUB].ensure(tmp_bytes + 16));
        B3D_CUDA(c, cub::DeviceRadixSort::SortPairsDescending(c->fbuf[F_CUB].p, tmp_bytes, keys_in, keys_out, vals_in, dst, (int)n_now, 0, bits, c->stream));
        c->launches += 3;
        cur ^= 1;
        n_prev = n_now;
    }
    return B3D_OK;
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
struct PointGrid { GridParams* gp; CellSlot* slots; float4* sorted; unsigned capacity; };

static unsigned pow2_capacity(size_t v) { unsigned c = 1024; while ((size_t)c < v) c <<= 1; return c; }

static int build_point_grid(b3d_ctx* c, const float4* pts, unsigned n, float cell, PointGrid* out) {
    const unsigned capacity = pow2_capacity(2 * (size_t)n);
    B3D_CUDA(c, c->fbuf[F_SLOTS].ensure(sizeof(CellSlot) * capacity));
    B3D_CUDA(c, c->fbuf[F_GP].ensure(sizeof(GridParams)));
    B3D_CUDA(c, c->fbuf[F_SORTED].ensure(sizeof(float4) * n));
    B3D_CUDA(c, c->fbuf[F_PT_SLOT].ensure(sizeof(unsigned) * n));
    B3D_CUDA(c, c->fbuf[F_PT_RANK].ensure(sizeof(unsigned) * n));
    GridParams* gp = c->fbuf[F_GP].as<GridParams>();
    CellSlot* slots = c->fbuf[F_SLOTS].as<CellSlot>();
    B3D_CUDA(c, cudaMemsetAsync(gp, 0, sizeof(GridParams), c->stream));
    grid_bounds_kernel<<<grid_for(n, 256, 2), 256, 0, c->stream>>>(pts, n, gp);
    B3D_LAUNCHED(c);
    grid_init_kernel<<<grid_for(capacity, 256, 4), 256, 0, c->stream>>>(slots, capacity, gp, cell / 1.02f, n);
    B3D_LAUNCHED(c);
    grid_insert_kernel<<<grid_for(n, 256, 8), 256, 0, c->stream>>>(pts, n, slots, gp, capacity - 1u, nullptr, c->fbuf[F_PT_SLOT].as<unsigned>(),
                                                                   c->fbuf[F_PT_RANK].as<unsigned>(), &gp->occupied);
    B3D_LAUNCHED(c);
    const unsigned tiles = (unsigned)div_up(capacity, kScanTile);
    B3D_CUDA(c, c->scan_tmp.ensure(sizeof(unsigned) * (tiles + 1)));
    SlotCount cnt{slots}; SlotStart st{slots};
    scan_tile_sums_kernel<<<tiles, kScanThreads, 0, c->stream>>>(cnt, capacity, c->scan_tmp.as<unsigned>());
    B3D_LAUNCHED(c);
    scan_tile_offsets_kernel<<<1, kScanThreads, 0, c->stream>>>(c->scan_tmp.as<unsigned>(), tiles, (unsigned*)nullptr);
    B3D_LAUNCHED(c);
    scan_emit_kernel<<<tiles, kScanThreads, 0, c->stream>>>(cnt, st, capacity, c->scan_tmp.as<unsigned>());
    B3D_LAUNCHED(c);
    grid_scatter_kernel<<<grid_for(n, 256, 8), 256, 0, c->stream>>>(pts, nullptr, n, slots, c->fbuf[F_PT_SLOT].as<unsigned>(),
                                                                    c->fbuf[F_PT_RANK].as<unsigned>(), c->fbuf[F_SORTED].as<float4>(), nullptr, nullptr);
    B3D_LAUNCHED(c);
    out->gp = gp; out->slots = slots; out->sorted = c->fbuf[F_SORTED].as<float4>(); out->capacity = capacity;
    return B3D_OK;
}

static int cloud_max_abs(b3d_ctx* c, const float4* pts, unsigned n, float* max_abs) {
    B3D_CUDA(c, c->fbuf[F_GP].ensure(sizeof(GridParams)));
    GridParams* gp = c->fbuf[F_GP].as<GridParams>();
    B3D_CUDA(c, cudaMemsetAsync(gp, 0, sizeof(GridParams), c->stream));
    grid_bounds_kernel<<<grid_for(n, 256, 2), 256, 0, c->stream>>>(pts, n, gp);
    B3D_LAUNCHED(c);
    unsigned bits = 0;
    B3D_CUDA(c, cudaMemcpyAsync(&bits, &gp->max_abs_bits, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    union { unsigned u; float f; } cv; cv.u = bits;
    *max_abs = cv.f;
    return B3D_OK;
}

static int upload_points(b3d_ctx* c, const float* xyz, size_t n, int raw_slot, int dst_slot) {
    B3D_CUDA(c, c->fbuf[raw_slot].ensure(sizeof(float) * 3 * n));
    B3D_CUDA(c, c->fbuf[dst_slot].ensure(sizeof(float4) * n));
    B3D_CUDA(c, cudaMemcpyAsync(c->fbuf[raw_slot].p, xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    return xyz_to_float4(c, c->fbuf[raw_slot].as<float>(), n, c->fbuf[dst_slot].as<float4>());
}

// device core: pts (float4, n) -> d_out (packed xyz normals, n)
int estimate_normals_dev(b3d_ctx* c, const float4* pts, unsigned n, int k, float* d_out) {
    if (k < 1 || k > kMaxList) return fail(c, B3D_ERR_INVALID, "estimate_normals: k must be in [1, 128]");
    StageTimer timer(c, 8);
    int rc;
    const unsigned K = (unsigned)k < n ? (unsigned)k : n;
    // Cell edge: aim at ~16 points per occupied cell, for which one 3x3x3 block almost always settles a 30-NN query.
    // Density is only known after a build, so start from a bounding-cube guess and correct it (surface scaling,
    // occupancy ~ cell^2) at most twice.  Any edge gives the same answer; this is a speed knob only.
    PointGrid g;
    float max_abs = 0.0f;
    rc = cloud_max_abs(c, pts, n, &max_abs);
    if (rc != B3D_OK) return rc;
    const float target = 0.55f * (float)K + 1.0f;
    float cell = 2.0f * (max_abs > 0.0f ? max_abs : 1.0f) * cbrtf(target / (float)n);
    for (int attempt = 0; attempt < 3; ++attempt) {
        rc = build_point_grid(c, pts, n, cell, &g);
        if (rc != B3D_OK) return rc;
        if (attempt == 2) break;
        GridParams h;
        B3D_CUDA(c, cudaMemcpyAsync(&h, g.gp, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        B3D_CUDA(c, cudaStreamSynchronize(c->stream));
        const float per_cell = (float)n / (float)(h.occupied ? h.occupied : 1u);
        if (per_cell > 0.6f * target && per_cell < 1.7f * target) break;
        const float next = h.cell * sqrtf(target / per_cell);
        if (!(next > h.cell * 1.05f || next < h.cell * 0.95f)) break;               // pinned at the coordinate-range floor
        cell = next;
    }
    B3D_CUDA(c, c->fbuf[F_COV].ensure(sizeof(float) * 9 * n));
    knn_covariance_kernel<<<grid_for(n, kFeatWarps, 16), kFeatWarps * 32, 0, c->stream>>>(pts, n, K, g.slots, g.sorted, g.gp, c->fbuf[F_COV].as<float>());
    B3D_LAUNCHED(c);
    normal_from_covariance_kernel<<<div_up(n, 128), 128, 0, c->stream>>>(pts, c->fbuf[F_COV].as<float>(), n, d_out);
    B3D_LAUNCHED(c);
    return B3D_OK;
}

int estimate_normals_impl(b3d_ctx* c, const float* xyz, size_t n_, int k, float* out_normals) {
    if (n_ == 0) return B3D_OK;
    if (n_ > 0x7FFFFFFFu) return fail(c, B3D_ERR_INVALID, "estimate_normals: too many points");
    const unsigned n = (unsigned)n_;
    int rc = upload_points(c, xyz, n, F_RAW, F_PTS4);
    if (rc != B3D_OK) return rc;
    B3D_CUDA(c, c->fbuf[F_OUT].ensure(sizeof(float) * 3 * n));
    rc = estimate_normals_dev(c, c->fbuf[F_PTS4].as<float4>(), n, k, c->fbuf[F_OUT].as<float>());
    if (rc != B3D_OK) return rc;
    B3D_CUDA(c, cudaMemcpyAsync(out_normals, c->fbuf[F_OUT].p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    return B3D_OK;
}

// radius lists (shared by the fused normals and by FPFH): grid of cell 1.02 r, then one warp per query
static int radius_lists_dev(b3d_ctx* c, const float4* pts, unsigned n, float radius, PointGrid* g) {
    if (n > 0x7FFFFFFFu / kFpfhMaxNn) return fail(c, B3D_ERR_INVALID, "compute_fpfh: too many points");
    const float r2 = radius * radius;                                               // registration.cpp:90
    int rc = build_point_grid(c, pts, n, radius * 1.02f, g);
    if (rc != B3D_OK) return rc;
    B3D_CUDA(c, c->fbuf[F_NBR].ensure(sizeof(unsigned) * kFpfhMaxNn * (size_t)n));
    B3D_CUDA(c, c->fbuf[F_NBR_CNT].ensure(sizeof(unsigned) * n));
    radius_neighbors_kernel<<<grid_for(n, kFeatWarps, 16), kFeatWarps * 32, 0, c->stream>>>(pts, n, r2, g->slots, g->sorted, g->gp,
                                                                                          c->fbuf[F_NBR].as<unsigned>(), c->fbuf[F_NBR_CNT].as<unsigned>());
    B3D_LAUNCHED(c);
    return B3D_OK;
}

static int fpfh_from_lists_dev(b3d_ctx* c, const float4* pts, const float4* nrm, unsigned n, float* d_out) {
    B3D_CUDA(c, c->fbuf[F_SPFH].ensure(sizeof(float) * 33 * (size_t)n));
    unsigned* nbr = c->fbuf[F_NBR].as<unsigned>(); unsigned* cnt = c->fbuf[F_NBR_CNT].as<unsigned>();
    const int blocks = grid_for(n, kFeatWarps, 16);
    spfh_kernel<<<blocks, kFeatWarps * 32, 0, c->stream>>>(pts, nrm, n, nbr, cnt, c->fbuf[F_SPFH].as<float>());
    B3D_LAUNCHED(c);
    fpfh_kernel<<<blocks, kFeatWarps * 32, 0, c->stream>>>(pts, n, nbr, cnt, c->fbuf[F_SPFH].as<float>(), d_out);
    B3D_LAUNCHED(c);
    return B3D_OK;
}

// device core: pts / nrm (float4, n) -> d_out (n x 33)
int compute_fpfh_dev(b3d_ctx* c, const float4* pts, const float4* nrm, unsigned n, float radius, float* d_out) {
    StageTimer timer(c, 9);
    PointGrid g;
    int rc = radius_lists_dev(c, pts, n, radius, &g);
    if (rc != B3D_OK) return rc;
    return fpfh_from_lists_dev(c, pts, nrm, n, d_out);
}

// fused: normals (k-NN taken from the radius lists) and FPFH from one neighbour search
static int normals_and_fpfh_dev(b3d_ctx* c, const float4* pts, unsigned n, int k, float radius, float* d_nrm_xyz, float4* nrm4, float* d_desc) {
    if (k < 1 || k > kMaxList) return fail(c, B3D_ERR_INVALID, "estimate_normals: k must be in [1, 128]");
    const unsigned K = (unsigned)k < n ? (unsigned)k : n;
    PointGrid g;
    {
        StageTimer timer(c, 8);
        int rc = radius_lists_dev(c, pts, n, radius, &g);
        if (rc != B3D_OK) return rc;
        B3D_CUDA(c, c->fbuf[F_COV].ensure(sizeof(float) * 9 * n));
        list_covariance_kernel<<<grid_for(n, kFeatWarps, 16), kFeatWarps * 32, 0, c->stream>>>(pts, n, K, c->fbuf[F_NBR].as<unsigned>(),
            c->fbuf[F_NBR_CNT].as<unsigned>(), g.slots, g.sorted, g.gp, c->fbuf[F_COV].as<float>());
        B3D_LAUNCHED(c);
        normal_from_covariance_kernel<<<div_up(n, 128), 128, 0, c->stream>>>(pts, c->fbuf[F_COV].as<float>(), n, d_nrm_xyz);
        B3D_LAUNCHED(c);
        rc = xyz_to_float4(c, d_nrm_xyz, n, nrm4);
        if (rc != B3D_OK) return rc;
    }
    StageTimer timer(c, 9);
    return fpfh_from_lists_dev(c, pts, nrm4, n, d_desc);
}

int compute_fpfh_impl(b3d_ctx* c, const float* xyz, const float* normals, size_t n_, float radius, float* out_desc) {
    if (n_ == 0) return B3D_OK;
    if (!normals) return fail(c, B3D_ERR_INVALID, "compute_fpfh: normals required");
    if (n_ > 0x7FFFFFFFu / kFpfhMaxNn) return fail(c, B3D_ERR_INVALID, "compute_fpfh: too many points");
    const unsigned n = (unsigned)n_;
    int rc = upload_points(c, xyz, n, F_RAW, F_PTS4);
    if (rc != B3D_OK) return rc;
    rc = upload_points(c, normals, n, F_RAW2, F_NRM4);
    if (rc != B3D_OK) return rc;
    B3D_CUDA(c, c->fbuf[F_OUT].ensure(sizeof(float) * 33 * (size_t)n));
    rc = compute_fpfh_dev(c, c->fbuf[F_PTS4].as<float4>(), c->fbuf[F_NRM4].as<float4>(), n, radius, c->fbuf[F_OUT].as<float>());
    if (rc != B3D_OK) return rc;
    B3D_CUDA(c, cudaMemcpyAsync(out_desc, c->fbuf[F_OUT].p, sizeof(float) * 33 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    return B3D_OK;
}

// device core: d_xyz / d_col (packed, n) -> packed means in fbuf[F_OUT] (*d_out_xyz, *d_out_col), *m_out voxels, reference order
int voxel_downsample_dev(b3d_ctx* c, const float* d_xyz, unsigned n, const float* d_col, float voxel, float** d_out_xyz, float** d_out_col_p,
                         unsigned* m_out) {
    if (!(voxel > 0.0f)) return fail(c, B3D_ERR_INVALID, "voxel_downsample: voxel_size must be positive");
    StageTimer timer(c, 7);
    const float inv = 1.0f / voxel;                                                 // registration.cpp:32
    B3D_CUDA(c, c->fbuf[F_KEYS_A].ensure(sizeof(unsigned long long) * n)); B3D_CUDA(c, c->fbuf[F_KEYS_B].ensure(sizeof(unsigned long long) * n));
    B3D_CUDA(c, c->fbuf[F_IDX_A].ensure(sizeof(unsigned) * n));            B3D_CUDA(c, c->fbuf[F_IDX_B].ensure(sizeof(unsigned) * n));
    B3D_CUDA(c, c->fbuf[F_SEG].ensure(sizeof(unsigned) * (n + 4)));
    unsigned* flags = c->fbuf[F_SEG].as<unsigned>() + n;                            // [0] out-of-range flag, [1] voxel count
    B3D_CUDA(c, cudaMemsetAsync(flags, 0, 2 * sizeof(unsigned), c->stream));
    auto* keys_a = c->fbuf[F_KEYS_A].as<unsigned long long>(); auto* keys_b = c->fbuf[F_KEYS_B].as<unsigned long long>();
    auto* idx_a = c->fbuf[F_IDX_A].as<unsigned>(); auto* idx_b = c->fbuf[F_IDX_B].as<unsigned>();
    voxel_key_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(d_xyz, n, inv, keys_a, idx_a, flags);
    B3D_LAUNCHED(c);
    size_t tmp_bytes = 0;
    B3D_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a, keys_b, idx_a, idx_b, (int)n, 0, 63, c->stream));
    B3D_CUDA(c, c->fbuf[F_CUB].ensure(tmp_bytes + 16));
    B3D_CUDA(c, cub::DeviceRadixSort::SortPairs(c->fbuf[F_CUB].p, tmp_bytes, keys_a, keys_b, idx_a, idx_b, (int)n, 0, 63, c->stream));
    c->launches += 8;                                                               // library launches (one upsweep/scan/downsweep set per digit)
    // segment heads -> voxel starts
    const unsigned tiles = (unsigned)div_up(n, kScanTile);
    B3D_CUDA(c, c->scan_tmp.ensure(sizeof(unsigned) * (tiles + 1)));
    SegHead head{keys_b}; SegStart emit{c->fbuf[F_SEG].as<unsigned>()};
    scan_tile_sums_kernel<<<tiles, kScanThreads, 0, c->stream>>>(head, n, c->scan_tmp.as<unsigned>());
    B3D_LAUNCHED(c);
    scan_tile_offsets_kernel<<<1, kScanThreads, 0, c->stream>>>(c->scan_tmp.as<unsigned>(), tiles, flags + 1);
    B3D_LAUNCHED(c);
    scan_emit_kernel<<<tiles, kScanThreads, 0, c->stream>>>(head, emit, n, c->scan_tmp.as<unsigned>());
    B3D_LAUNCHED(c);
    unsigned h_flags[2];
    B3D_CUDA(c, cudaMemcpyAsync(h_flags, flags, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    if (h_flags[0]) return fail(c, B3D_ERR_INVALID, "voxel_downsample: |coordinate / voxel_size| must stay below 2^20");
    const unsigned m = h_flags[1];
    *m_out = m;
    B3D_CUDA(c, c->fbuf[F_VOX_MEAN].ensure(sizeof(float) * 3 * m));
    if (d_col) B3D_CUDA(c, c->fbuf[F_VOX_COL].ensure(sizeof(float) * 3 * m));
    B3D_CUDA(c, c->fbuf[F_VOX_KEY].ensure(sizeof(int) * 3 * m * 2));                // [0, 3m) by voxel, [3m, 6m) in first-appearance order
    B3D_CUDA(c, c->fbuf[F_VOX_FIRST].ensure(sizeof(unsigned) * m * 2));
    B3D_CUDA(c, c->fbuf[F_VOX_ORDER].ensure(sizeof(unsigned) * m * 2));
    B3D_CUDA(c, c->fbuf[F_PERM].ensure(sizeof(unsigned) * m));
    B3D_CUDA(c, c->fbuf[F_OUT].ensure(sizeof(float) * 3 * m * 2));
    float* mean = c->fbuf[F_VOX_MEAN].as<float>(); float* mean_col = d_col ? c->fbuf[F_VOX_COL].as<float>() : nullptr;
    int* key3 = c->fbuf[F_VOX_KEY].as<int>(); int* key3_ordered = key3 + 3 * (size_t)m;
    unsigned* first = c->fbuf[F_VOX_FIRST].as<unsigned>(); unsigned* first_sorted = first + m;
    unsigned* vox = c->fbuf[F_VOX_ORDER].as<unsigned>(); unsigned* order = vox + m;
    voxel_mean_kernel<<<div_up(m, 128), 128, 0, c->stream>>>(d_xyz, d_col, n, keys_b, idx_b, c->fbuf[F_SEG].as<unsigned>(), flags + 1,
                                                             mean, mean_col, key3, first, vox);
    B3D_LAUNCHED(c);
    int first_bits = 1; while ((1ull << first_bits) < (unsigned long long)n) ++first_bits;     // first-appearance indices are < n
    B3D_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, first, first_sorted, vox, order, (int)m, 0, first_bits, c->stream));
    B3D_CUDA(c, c->fbuf[F_CUB].ensure(tmp_bytes + 16));
    B3D_CUDA(c, cub::DeviceRadixSort::SortPairs(c->fbuf[F_CUB].p, tmp_bytes, first, first_sorted, vox, order, (int)m, 0, first_bits, c->stream));
    c->launches += 4;
    gather_keys_kernel<<<div_up(m, 256), 256, 0, c->stream>>>(key3, order, m, key3_ordered);
    B3D_LAUNCHED(c);
    {   // the container's iteration order, replayed on the device (cross-checked against a real std::unordered_map by the tests,
        // through the CPU oracle — there is no host implementation in this library)
        int rc = container_order_device(c, key3_ordered, m, c->fbuf[F_PERM].as<unsigned>());
        if (rc != B3D_OK) return rc;
    }
    float* d_out = c->fbuf[F_OUT].as<float>(); float* d_out_col = d_col ? d_out + 3 * (size_t)m : nullptr;
    gather_voxels_kernel<<<div_up(m, 256), 256, 0, c->stream>>>(mean, mean_col, order, c->fbuf[F_PERM].as<unsigned>(), m, d_out, d_out_col);
    B3D_LAUNCHED(c);
    *d_out_xyz = d_out; *d_out_col_p = d_out_col;
    return B3D_OK;
}

int voxel_downsample_impl(b3d_ctx* c, const float* xyz, size_t n_, const float* colors, float voxel, float* out_xyz, float* out_colors,
                          size_t capacity, size_t* out_n) {
    *out_n = 0;
    if (n_ == 0) return B3D_OK;
    if (n_ > 0x7FFFFFFFu) return fail(c, B3D_ERR_INVALID, "voxel_downsample: too many points");
    if (colors && !out_colors) return fail(c, B3D_ERR_INVALID, "voxel_downsample: colors given but no output for them");
    const unsigned n = (unsigned)n_;
    B3D_CUDA(c, c->fbuf[F_RAW].ensure(sizeof(float) * 3 * n));
    B3D_CUDA(c, cudaMemcpyAsync(c->fbuf[F_RAW].p, xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    const float* d_col = nullptr;
    if (colors) {
        B3D_CUDA(c, c->fbuf[F_RAW2].ensure(sizeof(float) * 3 * n));
        B3D_CUDA(c, cudaMemcpyAsync(c->fbuf[F_RAW2].p, colors, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
        d_col = c->fbuf[F_RAW2].as<float>();
    }
    float* d_out = nullptr; float* d_out_col = nullptr; unsigned m = 0;
    int rc = voxel_downsample_dev(c, c->fbuf[F_RAW].as<float>(), n, d_col, voxel, &d_out, &d_out_col, &m);
    if (rc != B3D_OK) return rc;
    *out_n = m;
    if (m > capacity) return fail(c, B3D_ERR_INVALID, "voxel_downsample: output capacity too small (out_n holds the size needed)");
    B3D_CUDA(c, cudaMemcpyAsync(out_xyz, d_out, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, c->stream));
    if (colors) B3D_CUDA(c, cudaMemcpyAsync(out_colors, d_out_col, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    return B3D_OK;
}

// ---------------------------------------------------------------------------------
// depth image -> instance cloud (pipeline.cpp:38-84, CPU-branch semantics)
// ---------------------------------------------------------------------------------
struct DepthImage {
    const unsigned short* depth; const unsigned char* mask; const unsigned char* bgr;
    int w; float inv_scale; float clip, fx, fy, cx, cy;
    int mask_w, mask_h;          // mask size; when it differs from the depth image the mask is read through OpenCV's
    double ifx, ify;             // nearest-neighbour resize map (pipeline.cpp:39-41): sx = min(floor(x * ifx), mask_w - 1)
    __device__ unsigned char mask_at(unsigned px) const {
        if (ifx == 1.0 && ify == 1.0) return mask[px];
        const int u = (int)(px % (unsigned)w), v = (int)(px / (unsigned)w);
        const int sx = min((int)floor((double)u * ifx), mask_w - 1), sy = min((int)floor((double)v * ify), mask_h - 1);
        return mask[(size_t)sy * (size_t)mask_w + (size_t)sx];
    }
    __device__ float z_at(unsigned px) const {
        float z = (float)depth[px] * inv_scale;                                     // convertTo(CV_32FC1, 1.0 / scale): OpenCV scales 16u -> 32f in float
        if (mask && !(mask_at(px) > 10)) z = 0.0f;                                  // threshold(mask, 10) ; setTo(0, mask == 0)
        return z;
    }
};
struct PixelKept {
    DepthImage im;
    __device__ unsigned operator()(unsigned px) const { const float z = im.z_at(px); return (z <= 0.0f || z > im.clip) ? 0u : 1u; }
};
struct PixelEmit {                       // stable compaction: output order is the reference's raster order
    DepthImage im; float* xyz; float* rgb;
    __device__ void operator()(unsigned px, unsigned prefix, unsigned flag) const {
        if (!flag) return;
        const float z = im.z_at(px);
        const int u = (int)(px % (unsigned)im.w), v = (int)(px / (unsigned)im.w);
        xyz[3 * (size_t)prefix] = ((float)u - im.cx) * z / im.fx;
        xyz[3 * (size_t)prefix + 1] = ((float)v - im.cy) * z / im.fy;
        xyz[3 * (size_t)prefix + 2] = z;
        if (rgb) {
            rgb[3 * (size_t)prefix] = (float)im.bgr[3 * (size_t)px + 2] / 255.0f;
            rgb[3 * (size_t)prefix + 1] = (float)im.bgr[3 * (size_t)px + 1] / 255.0f;
            rgb[3 * (size_t)prefix + 2] = (float)im.bgr[3 * (size_t)px] / 255.0f;
        }
    }
};

// device core: host images in, packed xyz (and rgb) left in fbuf[F_IMG_XYZ] / fbuf[F_IMG_RGB]; *n_out points
static int depth_to_cloud_dev(b3d_ctx* c, const uint16_t* depth, int w, int h, const uint8_t* mask, int mask_w, int mask_h, float scale, float clip,
                              float fx, float fy, float cx, float cy, const uint8_t* bgr, unsigned* n_out) {
    const size_t px = (size_t)w * (size_t)h;
    if (mask && (mask_w <= 0 || mask_h <= 0)) { mask_w = w; mask_h = h; }          // 0 x 0: the mask has the depth image's size
    const size_t mpx = mask ? (size_t)mask_w * (size_t)mask_h : 0;
    if (mpx > 0x7FFFFFFFu) return fail(c, B3D_ERR_INVALID, "depth_to_cloud: bad mask size");
    B3D_CUDA(c, c->fbuf[F_IMG_DEPTH].ensure(px * 2)); B3D_CUDA(c, c->fbuf[F_IMG_XYZ].ensure(px * 12));
    B3D_CUDA(c, cudaMemcpyAsync(c->fbuf[F_IMG_DEPTH].p, depth, px * 2, cudaMemcpyHostToDevice, c->stream));
    if (mask) { B3D_CUDA(c, c->fbuf[F_IMG_MASK].ensure(mpx)); B3D_CUDA(c, cudaMemcpyAsync(c->fbuf[F_IMG_MASK].p, mask, mpx, cudaMemcpyHostToDevice, c->stream)); }
    if (bgr) {
        B3D_CUDA(c, c->fbuf[F_IMG_BGR].ensure(px * 3)); B3D_CUDA(c, c->fbuf[F_IMG_RGB].ensure(px * 12));
        B3D_CUDA(c, cudaMemcpyAsync(c->fbuf[F_IMG_BGR].p, bgr, px * 3, cudaMemcpyHostToDevice, c->stream));
    }
    DepthImage im{c->fbuf[F_IMG_DEPTH].as<unsigned short>(), mask ? c->fbuf[F_IMG_MASK].as<unsigned char>() : nullptr,
                  bgr ? c->fbuf[F_IMG_BGR].as<unsigned char>() : nullptr, w, (float)(1.0 / (double)scale), clip, fx, fy, cx, cy,
                  mask ? mask_w : w, mask ? mask_h : h,
                  mask ? 1.0 / ((double)w / (double)mask_w) : 1.0, mask ? 1.0 / ((double)h / (double)mask_h) : 1.0};      // cv::resize: ifx = 1 / (dst_w / src_w)
    PixelKept kept{im}; PixelEmit emit{im, c->fbuf[F_IMG_XYZ].as<float>(), bgr ? c->fbuf[F_IMG_RGB].as<float>() : nullptr};
    const unsigned tiles = (unsigned)div_up((long long)px, kScanTile);
    B3D_CUDA(c, c->scan_tmp.ensure(sizeof(unsigned) * (tiles + 2)));
    unsigned* total = c->scan_tmp.as<unsigned>() + tiles + 1;
    scan_tile_sums_kernel<<<tiles, kScanThreads, 0, c->stream>>>(kept, (unsigned)px, c->scan_tmp.as<unsigned>());
    B3D_LAUNCHED(c);
    scan_tile_offsets_kernel<<<1, kScanThreads, 0, c->stream>>>(c->scan_tmp.as<unsigned>(), tiles, total);
    B3D_LAUNCHED(c);
    scan_emit_kernel<<<tiles, kScanThreads, 0, c->stream>>>(kept, emit, (unsigned)px, c->scan_tmp.as<unsigned>());
    B3D_LAUNCHED(c);
    B3D_CUDA(c, cudaMemcpyAsync(n_out, total, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    return B3D_OK;
}

int depth_to_cloud_impl(b3d_ctx* c, const uint16_t* depth, int w, int h, const uint8_t* mask, int mask_w, int mask_h, float scale, float clip,
                        float fx, float fy, float cx, float cy, const uint8_t* bgr, float* out_xyz, float* out_rgb, size_t capacity, size_t* out_n) {
    *out_n = 0;
    if (w <= 0 || h <= 0 || (size_t)w * (size_t)h > 0x7FFFFFFFu) return fail(c, B3D_ERR_INVALID, "depth_to_cloud: bad image size");
    if (!(scale > 0.0f)) return fail(c, B3D_ERR_INVALID, "depth_to_cloud: scale_to_meters must be positive");
    if (bgr && !out_rgb) return fail(c, B3D_ERR_INVALID, "depth_to_cloud: colour image given but no output for it");
    unsigned n = 0;
    int rc = depth_to_cloud_dev(c, depth, w, h, mask, mask_w, mask_h, scale, clip, fx, fy, cx, cy, bgr, &n);
    if (rc != B3D_OK) return rc;
    *out_n = n;
    if (n > capacity) return fail(c, B3D_ERR_INVALID, "depth_to_cloud: output capacity too small (out_n holds the size needed)");
    if (n) {
        B3D_CUDA(c, cudaMemcpyAsync(out_xyz, c->fbuf[F_IMG_XYZ].p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        if (bgr) B3D_CUDA(c, cudaMemcpyAsync(out_rgb, c->fbuf[F_IMG_RGB].p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return B3D_OK;
}

// ---------------------------------------------------------------------------------
// One call per side of a registration, everything resident: raw points in, (points, normals, FPFH) left on the device
// where the matching / RANSAC / ICP stages read them.  This is the per-instance body of Pipeline::processInstance
// (src/pipeline.cpp:86-129) without the host containers between the stages.
// ---------------------------------------------------------------------------------
static int prepare_cloud_resident(b3d_ctx* c, const float* xyz, bool on_device, unsigned n, float voxel, int k, float radius,
                                  DevBuf& pts4, DevBuf& nrm4, DevBuf& desc, unsigned* m_out) {
    const float* d_raw = xyz;
    if (!on_device) {
        B3D_CUDA(c, c->fbuf[F_RAW].ensure(sizeof(float) * 3 * (size_t)n));
        B3D_CUDA(c, cudaMemcpyAsync(c->fbuf[F_RAW].p, xyz, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
        d_raw = c->fbuf[F_RAW].as<float>();
    }
    float* d_down = nullptr; float* d_unused = nullptr; unsigned m = 0;
    int rc = voxel_downsample_dev(c, d_raw, n, nullptr, voxel, &d_down, &d_unused, &m);
    if (rc != B3D_OK) return rc;
    *m_out = m;
    if (m == 0) return B3D_OK;
    B3D_CUDA(c, pts4.ensure(sizeof(float4) * m)); B3D_CUDA(c, nrm4.ensure(sizeof(float4) * m));
    B3D_CUDA(c, desc.ensure(sizeof(float) * 33 * (size_t)m));
    B3D_CUDA(c, c->fbuf[F_PIPE_NRM].ensure(sizeof(float) * 3 * (size_t)m));
    rc = xyz_to_float4(c, d_down, m, pts4.as<float4>());
    if (rc != B3D_OK) return rc;
    return normals_and_fpfh_dev(c, pts4.as<float4>(), m, k, radius, c->fbuf[F_PIPE_NRM].as<float>(), nrm4.as<float4>(), desc.as<float>());
}

int prepare_model_impl(b3d_ctx* c, const float* xyz, size_t n, float voxel, int k, float radius, size_t* out_n) {
    if (n > 0x7FFFFFFFu) return fail(c, B3D_ERR_INVALID, "prepare_model: too many points");
    c->model_ready = false; c->have_clouds = false; c->have_feats = false; c->have_corr = false; c->prepared = false; c->scored = false;
    unsigned m = 0;
    if (n) { int rc = prepare_cloud_resident(c, xyz, false, (unsigned)n, voxel, k, radius, c->tgt4, c->nrm4, c->tdesc, &m); if (rc != B3D_OK) return rc; }
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    c->n_tgt = m; c->has_normals = m > 0; c->tdesc_p = c->tdesc.as<float>();
    c->model_ready = true;
    if (out_n) *out_n = m;
    return B3D_OK;
}

static int register_resident(b3d_ctx* c, const float* xyz, bool on_device, size_t n, float voxel, int k, float radius, int ransac_iterations,
                             float confidence, float icp_threshold, int icp_iterations, int point_to_plane, b3d_scene_result* out, bool sharded);

int register_scene_impl(b3d_ctx* c, const float* xyz, size_t n, float voxel, int k, float radius, int ransac_iterations, float confidence,
                        float icp_threshold, int icp_iterations, int point_to_plane, b3d_scene_result* out) {
    return register_resident(c, xyz, false, n, voxel, k, radius, ransac_iterations, confidence, icp_threshold, icp_iterations, point_to_plane, out, false);
}

int register_scene_sharded_impl(b3d_ctx* c, const float* xyz, size_t n, float voxel, int k, float radius, int ransac_iterations, float confidence,
                                float icp_threshold, int icp_iterations, int point_to_plane, b3d_scene_result* out) {
    return register_resident(c, xyz, false, n, voxel, k, radius, ransac_iterations, confidence, icp_threshold, icp_iterations, point_to_plane, out, true);
}

int register_scene_device_impl(b3d_ctx* c, const float* xyz_dev, size_t n, float voxel, int k, float radius, int ransac_iterations, float confidence,
                               float icp_threshold, int icp_iterations, int point_to_plane, b3d_scene_result* out) {
    return register_resident(c, xyz_dev, true, n, voxel, k, radius, ransac_iterations, confidence, icp_threshold, icp_iterations, point_to_plane, out, false);
}

// depth image + mask of one instance -> pose: Pipeline::processInstance (pipeline.cpp:38-129) up to the refined transform
int register_depth_impl(b3d_ctx* c, const uint16_t* depth, int w, int h, const uint8_t* mask, int mask_w, int mask_h, float scale, float clip, float fx, float fy,
                        float cx, float cy, float voxel, int k, float radius, int ransac_iterations, float confidence, float icp_threshold,
                        int icp_iterations, int point_to_plane, b3d_scene_result* out) {
    if (!c->model_ready) return fail(c, B3D_ERR_STATE, "register_depth: call b3d_prepare_model first");
    if (w <= 0 || h <= 0 || (size_t)w * (size_t)h > 0x7FFFFFFFu) return fail(c, B3D_ERR_INVALID, "register_depth: bad image size");
    if (!(scale > 0.0f)) return fail(c, B3D_ERR_INVALID, "register_depth: scale_to_meters must be positive");
    unsigned n = 0;
    int rc = depth_to_cloud_dev(c, depth, w, h, mask, mask_w, mask_h, scale, clip, fx, fy, cx, cy, nullptr, &n);
    if (rc != B3D_OK) return rc;
    return register_resident(c, c->fbuf[F_IMG_XYZ].as<float>(), true, n, voxel, k, radius, ransac_iterations, confidence, icp_threshold,
                             icp_iterations, point_to_plane, out, false);
}

static int register_resident(b3d_ctx* c, const float* xyz, bool on_device, size_t n, float voxel, int k, float radius, int ransac_iterations,
                             float confidence, float icp_threshold, int icp_iterations, int point_to_plane, b3d_scene_result* out, bool sharded) {
    if (!c->model_ready) return fail(c, B3D_ERR_STATE, "register_scene: call b3d_prepare_model first");
    if (n > 0x7FFFFFFFu) return fail(c, B3D_ERR_INVALID, "register_scene: too many points");
    for (int i = 0; i < 16; ++i) out->coarse_T[i] = out->T[i] = (i % 5 == 0) ? 1.0f : 0.0f;    // RegistrationResult defaults (registration.hpp:27-29)
    out->coarse_fitness = out->coarse_rmse = out->fitness = out->rmse = 0.0f;
    out->coarse_best_iteration = -1; out->icp_iterations = 0; out->n_source_points = 0;
    c->have_clouds = false; c->have_feats = false; c->have_corr = false; c->prepared = false; c->scored = false;
    unsigned m = 0;
    if (n) { int rc = prepare_cloud_resident(c, xyz, on_device, (unsigned)n, voxel, k, radius, c->src4, c->fbuf[F_NRM4], c->sdesc, &m); if (rc != B3D_OK) return rc; }
    out->n_source_points = m;
    c->n_src = m; c->sdesc_p = c->sdesc.as<float>();
    c->have_clouds = true; c->have_feats = true;
    if (sharded) {
        // every rank ran the (cheap, deterministic) front end on the whole scene; matching rows and hypothesis ids are split
        // over the communicator (b3d_dist.cu); the refinement runs on every rank from the same coarse pose ("single-cloud ICP
        // stays on one GPU": replicas, no collective)
        int rcs = ransac_sharded_resident_impl(c, voxel, ransac_iterations, confidence, 1, out->coarse_T, &out->coarse_fitness, &out->coarse_rmse,
                                               &out->coarse_best_iteration);
        if (rcs != B3D_OK) return rcs;
        return b3d_icp_run(c, out->coarse_T, icp_threshold, icp_iterations, point_to_plane, 1, out->T, &out->fitness, &out->rmse, &out->icp_iterations);
    }
    int rc = b3d_match_features(c, 0, m); if (rc != B3D_OK) return rc;
    rc = b3d_ransac_prepare(c, voxel, ransac_iterations, confidence); if (rc != B3D_OK) return rc;
    rc = b3d_ransac_score(c, 0, ransac_iterations); if (rc != B3D_OK) return rc;
    int64_t* keys = reinterpret_cast<int64_t*>(&c->state.as<DeviceState>()->best_key);
    rc = b3d_ransac_reduce(c, 0, ransac_iterations, nullptr, keys); if (rc != B3D_OK) return rc;
    rc = b3d_ransac_finish(c, keys, out->coarse_T, &out->coarse_fitness, &out->coarse_rmse, &out->coarse_best_iteration); if (rc != B3D_OK) return rc;
    return b3d_icp_run(c, out->coarse_T, icp_threshold, icp_iterations, point_to_plane, 1, out->T, &out->fitness, &out->rmse, &out->icp_iterations);
}

}  // namespace b3d
