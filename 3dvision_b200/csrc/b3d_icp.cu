// b3d_icp.cu — point-to-plane / point-to-point ICP on a voxel-hash uniform grid.
// Replaces Registration::icpRefine (src/registration.cpp:297-414) and the body of
// GPURegistration::icpRefine (src/gpu_impl.cpp:141-260) of the reference; it is not a
// port of cuda/icp.cu.  Compile with --fmad=false.
//
// Nearest neighbour (registration.cpp:325-338).  The reference scans every target
// and then discards matches with sqrt(d2) > thr.  A match that survives is therefore
// within thr of the query, so it is enough to search the 27 cells around the query in
// a grid whose cell edge is >= 1.02*thr: the result (index and d2 bits) is identical
// for every point the reference keeps, and points it drops are dropped here too.
// Ties resolve on (d2, original index), which reproduces the reference's strict '<'
// scan order independent of the order atomics placed points in their cell.
//
// Grid build: cells live in an open-addressing hash table (64-bit packed cell key);
// points claim a slot with atomicCAS and a rank inside the cell with a
// warp-aggregated atomicAdd (one atomic per distinct cell per warp); a prefix sum
// over slot counts gives cell starts; points are then scattered into cell order as
// float4 (x, y, z, original index) so a cell is one contiguous, vectorised read.
//
// Per iteration one fused kernel transforms, searches, thresholds and accumulates the
// normal equations (or centroid/cross-covariance sums) in fp64 per thread ->
// warp-shuffle -> block -> per-block partials; a one-block kernel adds the partials
// in a fixed order (deterministic), solves (6x6 pivoted LDLT or 3x3 Jacobi SVD from
// b3d_linalg.cuh), updates T and the convergence flag on the device.  Nothing but
// the final 18 floats crosses PCIe.
#include "b3d_common.cuh"
#include "b3d_linalg.cuh"
#include "b3d_scan.cuh"
#include "b3d_grid.cuh"
#include "b3d_ess.cuh"
#include <float.h>
#include <math.h>

namespace b3d {

__device__ __forceinline__ void visit_cell(const GridView& g, int cx, int cy, int cz, float px, float py, float pz,
                                           float& best_d2, unsigned& best_idx, unsigned& best_pos) {
    const unsigned long long key = pack_cell(cx, cy, cz);
    unsigned slot = hash_cell(key) & g.mask;
    unsigned start = 0, count = 0;
    while (true) {
        const CellSlot s = g.slots[slot];
        if (s.key == key) { start = s.start; count = s.count; break; }
        if (s.key == kEmptyKey) break;
        slot = (slot + 1u) & g.mask;
    }
    auto consider = [&](const float4 q, unsigned pos) {
        float e0 = px - q.x, e1 = py - q.y, e2 = pz - q.z;
        float d2 = e0 * e0 + (e1 * e1 + e2 * e2);                  // (p - q).squaredNorm()
        unsigned idx = __float_as_uint(q.w);
        if (d2 < best_d2 || (d2 == best_d2 && idx < best_idx)) { best_d2 = d2; best_idx = idx; best_pos = pos; }
    };
    unsigned k = 0;
    for (; k + 4 <= count; k += 4) {                               // four independent loads in flight
        const float4 q0 = g.pts[start + k], q1 = g.pts[start + k + 1], q2 = g.pts[start + k + 2], q3 = g.pts[start + k + 3];
        consider(q0, start + k); consider(q1, start + k + 1); consider(q2, start + k + 2); consider(q3, start + k + 3);
    }
    for (; k < count; ++k) consider(g.pts[start + k], start + k);
}

// Nearest target of p within the 27-cell neighbourhood: lexicographic (d2, index) minimum.
// The query's own cell is searched first; a neighbour cell is skipped when even its nearest face
// (shrunk by `slack` to cover rounding in the cell assignment) is farther than the best match so
// far, which cannot change the minimum or its tie-break.  The face tests are evaluated branch-free
// into a 27-bit mask and the surviving neighbours are then visited in compacted order, so the
// lanes of a warp stay converged for max(popcount) rounds instead of the union of their subsets.
__device__ __forceinline__ void grid_nearest(float px, float py, float pz, const GridView& g,
                                             float& best_d2, unsigned& best_idx, unsigned& best_pos) {
    best_d2 = FLT_MAX; best_idx = B3D_NO_MATCH; best_pos = 0u;
    const int cx = cell_coord(px, g.inv), cy = cell_coord(py, g.inv), cz = cell_coord(pz, g.inv);
    visit_cell(g, cx, cy, cz, px, py, pz, best_d2, best_idx, best_pos);
    // squared distance from p to the low / high faces of its cell along each axis, made conservative
    float f2[3][3];                                                // [axis][0: low side, 1: same, 2: high side]
    {
        const float lo0 = fmaxf(px - (float)cx * g.cell - g.slack, 0.0f), hi0 = fmaxf((float)(cx + 1) * g.cell - px - g.slack, 0.0f);
        const float lo1 = fmaxf(py - (float)cy * g.cell - g.slack, 0.0f), hi1 = fmaxf((float)(cy + 1) * g.cell - py - g.slack, 0.0f);
        const float lo2 = fmaxf(pz - (float)cz * g.cell - g.slack, 0.0f), hi2 = fmaxf((float)(cz + 1) * g.cell - pz - g.slack, 0.0f);
        f2[0][0] = __fmul_rd(lo0, lo0) * 0.999f; f2[0][1] = 0.0f; f2[0][2] = __fmul_rd(hi0, hi0) * 0.999f;
        f2[1][0] = __fmul_rd(lo1, lo1) * 0.999f; f2[1][1] = 0.0f; f2[1][2] = __fmul_rd(hi1, hi1) * 0.999f;
        f2[2][0] = __fmul_rd(lo2, lo2) * 0.999f; f2[2][1] = 0.0f; f2[2][2] = __fmul_rd(hi2, hi2) * 0.999f;
    }
    unsigned todo = 0u;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        if (k == 13) continue;
        const float gap2 = f2[0][k % 3] + (f2[1][(k / 3) % 3] + f2[2][k / 9]);      // indices fold at compile time
        todo |= (gap2 > best_d2 ? 0u : 1u) << k;
    }
    while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1u;
        const int dx = k % 3 - 1, dy = (k / 3) % 3 - 1, dz = k / 9 - 1;
        // the best may have improved since the mask was built: re-test before paying for the probe
        const float gx = dx < 0 ? f2[0][0] : (dx > 0 ? f2[0][2] : 0.0f);
        const float gy = dy < 0 ? f2[1][0] : (dy > 0 ? f2[1][2] : 0.0f);
        const float gz = dz < 0 ? f2[2][0] : (dz > 0 ? f2[2][2] : 0.0f);
        if (gx + (gy + gz) > best_d2) continue;
        visit_cell(g, cx + dx, cy + dy, cz + dz, px, py, pz, best_d2, best_idx, best_pos);
    }
}

// Fine-grid parameters from the density the coarse build observed.  For a surface sampled with
// `ppc` points per coarse cell the point spacing is about cell/sqrt(ppc); a fine cell of twice that
// holds a handful of points.  The fine cell is never finer than max|coord|/65536 (so fl(p*inv) stays
// within 2^-7 cell, as for the coarse grid) and the level is switched off when it would not be at
// least ~2x finer than the coarse one.
__global__ void fine_params_kernel(const GridParams* __restrict__ gc, GridParams* __restrict__ gf, unsigned capacity, unsigned n) {
    if (threadIdx.x != 0) return;
    const float max_abs = __uint_as_float(gc->max_abs_bits);
    const float ppc = (float)n / (float)max(gc->occupied, 1u);
    float cell = 2.0f * gc->cell / sqrtf(fmaxf(ppc, 1.0f));
    cell = fmaxf(cell, gc->cell * (1.0f / 16.0f));
    cell = fmaxf(cell, max_abs * (1.0001f / kCoordLimit));
    const bool usable = isfinite(gc->cell) && gc->cell > 0.0f && gc->inv_cell > 0.0f;
    const bool finer = usable && isfinite(cell) && cell > 0.0f && cell < 0.6f * gc->cell;
    // sparse target: lists over the coarse cells themselves.  They hold every point of the 27 coarse cells, i.e. every
    // match the reference can keep, so the lookup is complete and the coarse walk is never needed.
    gf->inv_cell = finer ? 1.0f / cell : (usable ? gc->inv_cell : 0.0f);
    gf->cell = finer ? 1.0f / gf->inv_cell : (usable ? gc->cell : 0.0f);
    gf->slack = gc->slack;
    gf->mask = capacity - 1u;
    gf->max_abs_bits = gc->max_abs_bits;
    gf->n_points = n;
    gf->enabled = 0u;                                          // set once the lists exist (build_neighbourhood_lists)
    gf->occupied = usable ? 1u : 0u;                           // "wanted": the level can be built for this target
    gf->complete = (usable && !finer) ? 1u : 0u;
    const float reach = gf->cell * 0.98f - gf->slack;          // every target closer than this lies in the 27 fine cells
    gf->accept2 = (usable && reach > 0.0f) ? __fmul_rd(reach, reach) * 0.999f : 0.0f;
}

// Second level = per-cell NEIGHBOURHOOD LISTS.  Walking 27 hash cells per query is 27 dependent L2 round trips and
// leaves ~10 of 32 lanes active; measured, a query that probes ONE cell and scans one contiguous run costs a third
// (configs[1]: 94 -> 34 us per iteration with an own-cell-only probe).  So the 27-cell union is materialised once per
// call: every target point is appended to the list of each of the 27 fine cells around it (count with atomics, scan,
// scatter), 27*n float4 in total; a query then probes the table with its own cell and scans that list.  The list of
// cell c holds exactly the points of the 27 cells around c, so the acceptance rule is unchanged: a lexicographic
// (d2, index) minimum closer than the fine reach is the global one.  A cell without a list has no point within the
// reach.  If the table fills up (very sparse targets) the level is flagged unusable and every query takes the coarse walk.
constexpr int kNeighbourhoodProbes = 256;
__device__ __forceinline__ unsigned neighbourhood_claim(CellSlot* __restrict__ slots, unsigned mask, unsigned long long key) {
    unsigned slot = hash_cell(key) & mask;
    for (int probe = 0; probe < kNeighbourhoodProbes / 2; ++probe) {
        const unsigned long long prev = atomicCAS(&slots[slot].key, kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) return slot;
        slot = (slot + 1u) & mask;
    }
    return 0xFFFFFFFFu;
}
__global__ void neighbourhood_count_kernel(const float4* __restrict__ pts, unsigned n, CellSlot* __restrict__ slots, GridParams* __restrict__ gf,
                                           unsigned* __restrict__ pt_slot27) {
    if (!gf->enabled) return;
    const float inv = gf->inv_cell;
    const unsigned mask = gf->mask;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < (size_t)n * 27u; w += (size_t)gridDim.x * blockDim.x) {
        const unsigned i = (unsigned)(w / 27u), k = (unsigned)(w % 27u);
        const float4 p = pts[i];
        const unsigned long long key = pack_cell(cell_coord(p.x, inv) + (int)(k % 3u) - 1, cell_coord(p.y, inv) + (int)((k / 3u) % 3u) - 1,
                                                 cell_coord(p.z, inv) + (int)(k / 9u) - 1);
        const unsigned slot = neighbourhood_claim(slots, mask, key);
        if (slot == 0xFFFFFFFFu) gf->overflow = 1u; else atomicAdd(&slots[slot].count, 1u);
        pt_slot27[w] = slot;
    }
}
__global__ void neighbourhood_fill_kernel(const float4* __restrict__ pts, unsigned n, const CellSlot* __restrict__ slots,
                                          const GridParams* __restrict__ gf, const unsigned* __restrict__ pt_slot27,
                                          unsigned* __restrict__ cursor, float4* __restrict__ lists) {
    if (!gf->enabled || gf->overflow) return;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < (size_t)n * 27u; w += (size_t)gridDim.x * blockDim.x) {
        const unsigned slot = pt_slot27[w];
        const unsigned i = (unsigned)(w / 27u);
        float4 p = pts[i]; p.w = __uint_as_float(i);
        lists[slots[slot].start + atomicAdd(&cursor[slot], 1u)] = p;
    }
}
// the source is binned by the cells the queries will probe: fine when that level is usable, else coarse
__global__ void binning_params_kernel(GridParams* gp) {
    if (threadIdx.x == 0) gp[2] = (gp[1].enabled && !gp[1].overflow) ? gp[1] : gp[0];
}

// Two-level nearest neighbour.  Dense targets put tens of points into a coarse cell (its edge is
// the inlier threshold, ~10x the point spacing in configs[1]) although the true neighbour is almost
// always within a fraction of it.  The fine level is searched first; if it returns a match closer than
// the fine reach, no point outside the 27 fine cells around the query can be closer or tie (they are all
// farther than the reach), so that match IS the global lexicographic (d2, index) minimum.  Otherwise the
// coarse search, which is exact for everything within the threshold, runs from scratch for that query.
// (A warp-pooled variant that deals the surviving neighbour visits out 32 at a time and merges through
// shared-memory atomicMin keys was measured and was not faster: 93 vs 95 us at 300k x 100k, slower on
// sparse targets; the per-thread form is kept.)
// One probe: the neighbourhood list of p's fine cell (see neighbourhood_* kernels).  Returns false if the cell has none.
// Keeps the kKeep lexicographically smallest (d2, index) candidates in order (bd[0], bi[0] is the match) and the smallest d2
// among all the others (`next_d2`), which is what the re-use certificate of icp_match needs.
constexpr int kKeep = 4;
__device__ __forceinline__ bool neighbourhood_nearest(float px, float py, float pz, const GridView& d, float (&bd)[kKeep], unsigned (&bi)[kKeep],
                                                      float& next_d2) {
#pragma unroll
    for (int k = 0; k < kKeep; ++k) { bd[k] = FLT_MAX; bi[k] = B3D_NO_MATCH; }
    next_d2 = FLT_MAX;
    const unsigned long long key = pack_cell(cell_coord(px, d.inv), cell_coord(py, d.inv), cell_coord(pz, d.inv));
    unsigned slot = hash_cell(key) & d.mask;
    unsigned start = 0, count = 0;
    for (int probe = 0; probe < kNeighbourhoodProbes; ++probe) {
        const CellSlot s = d.slots[slot];
        if (s.key == key) { start = s.start; count = s.count; break; }
        if (s.key == kEmptyKey) return false;
        slot = (slot + 1u) & d.mask;
    }
    auto consider = [&](const float4 q) {
        const float e0 = px - q.x, e1 = py - q.y, e2 = pz - q.z;
        float cd = e0 * e0 + (e1 * e1 + e2 * e2);                  // (p - q).squaredNorm()
        unsigned ci = __float_as_uint(q.w);
        if (!(cd < bd[kKeep - 1] || (cd == bd[kKeep - 1] && ci < bi[kKeep - 1]))) { next_d2 = fminf(next_d2, cd); return; }
        next_d2 = fminf(next_d2, bd[kKeep - 1]);                   // the displaced last entry joins "the others"
#pragma unroll
        for (int k = 0; k < kKeep; ++k) {                          // insertion: bubble the new key down to its place
            const bool before = cd < bd[k] || (cd == bd[k] && ci < bi[k]);
            const float td = before ? bd[k] : cd; const unsigned ti = before ? bi[k] : ci;
            bd[k] = before ? cd : bd[k]; bi[k] = before ? ci : bi[k];
            cd = td; ci = ti;
        }
    };
    unsigned k = 0;
    for (; k + 4 <= count; k += 4) {
        const float4 q0 = d.pts[start + k], q1 = d.pts[start + k + 1], q2 = d.pts[start + k + 2], q3 = d.pts[start + k + 3];
        consider(q0); consider(q1); consider(q2); consider(q3);
    }
    for (; k < count; ++k) consider(d.pts[start + k]);
    return count != 0u;
}

// `hold2`: squared distance the query may move while its match provably stays among the kKeep candidates kept (0 = no
// certificate).  The list holds every target within `reach`, so every target that is NOT kept is at least
// min(next_d2, reach) away; by the triangle inequality the nearest target stays one of the kept ones while the query has
// moved less than half the gap between that bound and the current best.
__device__ __forceinline__ void grid_nearest2(float px, float py, float pz, const GridView& coarse, const GridView& fine,
                                              bool fine_on, bool complete, float accept2, float& best_d2, unsigned& best_idx,
                                              unsigned (&keep)[kKeep], float& hold2) {
    hold2 = 0.0f;
#pragma unroll
    for (int k = 0; k < kKeep; ++k) keep[k] = B3D_NO_MATCH;
    if (fine_on) {
        float bd[kKeep], next_d2;
        const bool found = neighbourhood_nearest(px, py, pz, fine, bd, keep, next_d2);
        best_d2 = bd[0]; best_idx = keep[0];
        if (complete || (found && best_d2 < accept2)) {
            if (found) {
                const float other = fminf(sqrtf(next_d2), sqrtf(accept2)) * 0.9999f;
                const float gap = 0.5f * (other - sqrtf(best_d2)) - 2.0f * fine.slack;
                hold2 = gap > 0.0f ? __fmul_rd(gap, gap) : 0.0f;
            }
            return;
        }
    }
    unsigned pos;
    grid_nearest(px, py, pz, coarse, best_d2, best_idx, pos);
}

// One query of an iteration kernel.  `cache` keeps, per query slot, where the query stood when its match was last searched
// and how far it may move before the search has to be repeated (see hold2), `cache_idx` the kKeep candidates that search kept:
// once ICP has converged the pose changes by microns per iteration and almost every query only re-evaluates its kKeep
// candidates exactly (kKeep gathers instead of a list scan).  The lexicographic (d2, index) minimum over them is what a fresh
// search would return, so every sum downstream is unchanged.
__device__ __forceinline__ void icp_match(float x, float y, float z, unsigned slot_i, const GridView& g, const GridView& gfine,
                                          const GridParams* __restrict__ gp, const float4* __restrict__ tgt4,
                                          float4* __restrict__ cache, uint4* __restrict__ cache_idx, float& d2, unsigned& idx) {
    if (cache) {
        const float4 c = cache[slot_i];
        if (c.w > 0.0f) {
            const float m0 = x - c.x, m1 = y - c.y, m2 = z - c.z;
            if (m0 * m0 + (m1 * m1 + m2 * m2) < c.w) {
                const uint4 k4 = cache_idx[slot_i];
                const unsigned cand[kKeep] = {k4.x, k4.y, k4.z, k4.w};
                d2 = FLT_MAX; idx = B3D_NO_MATCH;
#pragma unroll
                for (int k = 0; k < kKeep; ++k) {
                    if (cand[k] == B3D_NO_MATCH) continue;
                    const float4 q = tgt4[cand[k]];
                    const float e0 = x - q.x, e1 = y - q.y, e2 = z - q.z;
                    const float cd = e0 * e0 + (e1 * e1 + e2 * e2);            // (p - q).squaredNorm()
                    if (cd < d2 || (cd == d2 && cand[k] < idx)) { d2 = cd; idx = cand[k]; }
                }
                return;
            }
        }
    }
    float hold2; unsigned keep[kKeep];
    grid_nearest2(x, y, z, g, gfine, gp[1].enabled != 0u && gp[1].overflow == 0u, gp[1].complete != 0u, gp[1].accept2, d2, idx, keep, hold2);
    if (cache) {
        cache[slot_i] = make_float4(x, y, z, idx != B3D_NO_MATCH ? hold2 : 0.0f);
        cache_idx[slot_i] = make_uint4(keep[0], keep[1], keep[2], keep[3]);
    }
}

__device__ __forceinline__ void load_Rt(const float* __restrict__ T, float R[9], float t[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r * 3 + c] = T[c * 4 + r];
        t[r] = T[12 + r];
    }
}
__device__ __forceinline__ void transform_point(const float R[9], const float t[3], float4 s, float& x, float& y, float& z) {
    x = (R[0] * s.x + (R[1] * s.y + R[2] * s.z)) + t[0];     // R*s + t, Eigen order (registration.cpp:326)
    y = (R[3] * s.x + (R[4] * s.y + R[5] * s.z)) + t[1];
    z = (R[6] * s.x + (R[7] * s.y + R[8] * s.z)) + t[2];
}

// parity tap: per-source nearest neighbour under T
__global__ void icp_nearest_kernel(const float4* __restrict__ src, unsigned n_src, const float* __restrict__ T, float thr,
                                   const CellSlot* __restrict__ slots, const float4* __restrict__ gpts,
                                   const GridParams* __restrict__ gp, uint32_t* __restrict__ out_idx, float* __restrict__ out_d2) {
    float R[9], t[3];
    load_Rt(T, R, t);
    const GridView g = make_view(slots, gpts, gp);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += gridDim.x * blockDim.x) {
        float x, y, z; transform_point(R, t, src[i], x, y, z);
        float d2; unsigned idx, pos;
        grid_nearest(x, y, z, g, d2, idx, pos);
        if (idx != B3D_NO_MATCH && sqrtf(d2) > thr) { idx = B3D_NO_MATCH; }
        out_idx[i] = idx;
        out_d2[i] = (idx == B3D_NO_MATCH) ? FLT_MAX : d2;
    }
}

// ---------------------------------------------------------------------------------
// fused iteration: transform -> NN -> threshold -> accumulate
// ---------------------------------------------------------------------------------
constexpr int kIcpThreads = 256;
constexpr int kAccPlane = 28;     // 21 (upper ATA) + 6 (ATb) + 1 (sum d2)
constexpr int kAccPoint = 16;     // 3 (sum p) + 3 (sum q) + 9 (sum p q^T) + 1 (sum d2)
constexpr int kAccMax = 28;
constexpr int kPartialStride = 32;   // doubles per block partial: kAccMax values + count

// Block-wide sum of one contribution vector per thread: fp64 from the first addition on.
template <int NV>
__device__ __forceinline__ void block_reduce_store(const float (&contrib)[NV], int n_corr, double* __restrict__ partial_out) {
    __shared__ double red[kIcpThreads / 32][kAccMax + 1];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double a = (double)contrib[v];
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
        if (lane == 0) red[warp][v] = a;
    }
    {
        int cnt = n_corr;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
        if (lane == 0) red[warp][kAccMax] = (double)cnt;
    }
    __syncthreads();
    if (threadIdx.x < kPartialStride) {
        const int v = threadIdx.x;
        double s = 0.0;
        if (v < NV || v == kAccMax) {
#pragma unroll
            for (int w = 0; w < kIcpThreads / 32; ++w) s += red[w][v];
        }
        partial_out[v] = s;
    }
}

// One query per thread: the search phase needs few registers, so occupancy (and with it the
// number of hash probes and cell reads in flight) is high; the per-point terms are formed in
// fp32 exactly as the reference forms them and enter fp64 at the first addition.
template <bool PLANE>
__global__ void __launch_bounds__(kIcpThreads, 4)
icp_accumulate_kernel(const float4* __restrict__ src, unsigned n_src, const DeviceState* __restrict__ st, float thr,
                      const CellSlot* __restrict__ slots, const float4* __restrict__ gpts,
                      const CellSlot* __restrict__ fslots, const float4* __restrict__ fpts,
                      const float4* __restrict__ tgt4, const float4* __restrict__ nrm4,
                      const GridParams* __restrict__ gp, double* __restrict__ partials,
                      float4* __restrict__ cache, uint4* __restrict__ cache_idx) {
    // programmatic dependent launch: let the solve kernel behind us get resident now, and wait here for the one before us
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
    if (st->done) return;
    constexpr int NV = PLANE ? kAccPlane : kAccPoint;
    // ---- search phase (few live registers) ----
    float x = 0.f, y = 0.f, z = 0.f, d2 = 0.f;
    unsigned pos = 0;
    int n_corr = 0;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_src) {
        float R[9], t[3];
        load_Rt(st->T, R, t);
        const GridView g = make_view(slots, gpts, gp);
        const GridView gfine = make_view(fslots, fpts, gp + 1);
        transform_point(R, t, src[i], x, y, z);
        unsigned idx;
        icp_match(x, y, z, i, g, gfine, gp, tgt4, cache, cache_idx, d2, idx);
        pos = idx;                                                           // original target index
        n_corr = (idx != B3D_NO_MATCH && !(sqrtf(d2) > thr)) ? 1 : 0;       // registration.cpp:337-338 (d == thr is kept)
    }
    // ---- contribution phase ----
    float c[NV];
    if (!n_corr) { x = 0.f; y = 0.f; z = 0.f; d2 = 0.f; }      // unmatched (or NaN) queries contribute exact zeros
    const float keep = n_corr ? 1.0f : 0.0f;
    const float4 q = n_corr ? tgt4[pos] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (PLANE) {
        const float4 n = n_corr ? nrm4[pos] : make_float4(0.f, 0.f, 0.f, 0.f);
        float J[6];
        J[0] = y * n.z - z * n.y;                              // p.cross(n), registration.cpp:346
        J[1] = z * n.x - x * n.z;
        J[2] = x * n.y - y * n.x;
        J[3] = n.x; J[4] = n.y; J[5] = n.z;
        float r = (x - q.x) * n.x + ((y - q.y) * n.y + (z - q.z) * n.z);       // (p - q).dot(n)
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) c[k++] = J[a] * J[b];                  // zero normal => zero terms when unmatched
#pragma unroll
        for (int a = 0; a < 6; ++a) c[21 + a] = J[a] * r;
        c[27] = d2 * keep;
    } else {
        c[0] = x * keep; c[1] = y * keep; c[2] = z * keep; c[3] = q.x; c[4] = q.y; c[5] = q.z;
#pragma unroll
        for (int v = 6; v < 15; ++v) c[v] = 0.0f;
        c[15] = d2 * keep;
    }
    if (PLANE) {
        block_reduce_store<NV>(c, n_corr, partials + (size_t)blockIdx.x * kPartialStride);
    } else {
        // point-to-point: sum p, sum q, sum p q^T (exact fp64 products), sum d2
        __shared__ double red[kIcpThreads / 32][kAccMax + 1];
        const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const double p[3] = {(double)c[0], (double)c[1], (double)c[2]}, q[3] = {(double)c[3], (double)c[4], (double)c[5]};
        double vals[kAccPoint];
#pragma unroll
        for (int a = 0; a < 3; ++a) { vals[a] = p[a]; vals[3 + a] = q[a]; }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) vals[6 + a * 3 + b] = p[a] * q[b];
        vals[15] = (double)c[15];
#pragma unroll
        for (int v = 0; v < kAccPoint; ++v) {
            double a = vals[v];
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
            if (lane == 0) red[warp][v] = a;
        }
        int cnt = n_corr;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
        if (lane == 0) red[warp][kAccMax] = (double)cnt;
        __syncthreads();
        if (threadIdx.x < kPartialStride) {
            const int v = threadIdx.x;
            double s = 0.0;
            if (v < kAccPoint || v == kAccMax) {
#pragma unroll
                for (int w = 0; w < kIcpThreads / 32; ++w) s += red[w][v];
            }
            partials[(size_t)blockIdx.x * kPartialStride + v] = s;
        }
    }
}

// one block: fixed-order sum of the block partials, solve, update T / result / flags
constexpr int kUpdateThreads = 1024;
template <bool PLANE>
__global__ void __launch_bounds__(kUpdateThreads)
icp_update_kernel(const double* __restrict__ partials, int n_blocks, int iter, float n_src_f, int stop_on_convergence,
                  DeviceState* __restrict__ st) {
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
    if (st->done) return;
    // fixed-order (hence run-to-run deterministic) sum of the block partials: 8 strided groups, then 8 -> 1
    __shared__ double grp[kUpdateThreads / kPartialStride][kPartialStride];
    __shared__ double tot[kPartialStride];
    {
        const int v = threadIdx.x % kPartialStride, gidx = threadIdx.x / kPartialStride;
        double s = 0.0;
#pragma unroll 8
        for (int b = gidx; b < n_blocks; b += kUpdateThreads / kPartialStride) s += partials[(size_t)b * kPartialStride + v];
        grp[gidx][v] = s;
    }
    __syncthreads();
    if (threadIdx.x < kPartialStride) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kUpdateThreads / kPartialStride; ++k) s += grp[k][threadIdx.x];
        tot[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int n_corr = (int)(tot[kAccMax] + 0.5);
    st->n_corr_last = n_corr;
    if (n_corr < 3) { st->done = 1; return; }              // registration.cpp:361
    float delta[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) delta[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    float total_error;
    if (PLANE) {
        float A[36], nb[6], x[6];
        int k = 0;
        for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b) { float val = (float)tot[k++]; A[a * 6 + b] = val; A[b * 6 + a] = val; }
        for (int a = 0; a < 6; ++a) nb[a] = -(float)tot[21 + a];
        total_error = (float)tot[27];
        ldlt6_solve(A, nb, x);                               // registration.cpp:366
        Mat3 dR; euler_xyz_to_matrix(x[0], x[1], x[2], dR);  // registration.cpp:369-371
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) delta[c * 4 + r] = dR(r, c); delta[12 + r] = x[3 + r]; }
    } else {
        const double n = (double)n_corr;
        double pm[3], qm[3];
        for (int a = 0; a < 3; ++a) { pm[a] = tot[a] / n; qm[a] = tot[3 + a] / n; }
        Mat3 H;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) H(a, b) = (float)(tot[6 + a * 3 + b] - n * pm[a] * qm[b]);   // sum (p-pm)(q-qm)^T
        total_error = (float)tot[15];
        Mat3 dR; rotation_from_cross_covariance(H, dR);      // registration.cpp:388-394
        float pmf[3] = {(float)pm[0], (float)pm[1], (float)pm[2]}, qmf[3] = {(float)qm[0], (float)qm[1], (float)qm[2]};
        float r0, r1, r2; mat3_vec(dR, pmf[0], pmf[1], pmf[2], r0, r1, r2);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) delta[c * 4 + r] = dR(r, c);
        delta[12] = qmf[0] - r0; delta[13] = qmf[1] - r1; delta[14] = qmf[2] - r2;     // registration.cpp:396
    }
    float Tn[16];
    mat4_mul(delta, st->T, Tn);                              // T = delta * T
    const float prev_rmse = st->res_rmse;
    const float rmse = sqrtf(total_error / (float)n_corr);
    for (int i = 0; i < 16; ++i) { st->T[i] = Tn[i]; st->res_T[i] = Tn[i]; }
    st->res_rmse = rmse;
    st->res_fitness = (float)n_corr / n_src_f;
    st->iterations = iter + 1;
    if (stop_on_convergence && iter > 0 && fabsf(prev_rmse - rmse) < 1e-6f) st->done = 1;    // registration.cpp:406
}

// ---------------------------------------------------------------------------------
// Reference-order point-to-point iteration (default for point-to-point).
// The reference forms the centroids, the cross-covariance and total_error by adding one point at a
// time in source order, in fp32 (registration.cpp:341, 374-386).  Point-to-point ICP converges
// slowly, so the ~1e-6 m rounding noise of those sums is amplified to ~1e-4 in the final pose; the
// only way to land inside the reference's own noise is to add in the same order.  So: the search
// kernel stores every query's result at its ORIGINAL index, a prefix sum compacts the matched
// (p, q) records in source order, and one block replays the reference's loops — 31 producer warps
// stage tiles of records in shared memory while the lanes of one consumer warp each own one
// running sum and add left to right (the dependent FADD chain is the floor).  Same solve as the
// fast path afterwards.  b3d_set_icp_mode(ctx, 1) selects the fp64 tree sums instead.
// ---------------------------------------------------------------------------------
template <bool BINNED>
__global__ void __launch_bounds__(kIcpThreads, 4)
icp_search_kernel(const float4* __restrict__ src, unsigned n_src, DeviceState* __restrict__ st, float thr,
                  const CellSlot* __restrict__ slots, const float4* __restrict__ gpts,
                  const CellSlot* __restrict__ fslots, const float4* __restrict__ fpts,
                  const GridParams* __restrict__ gp, float4* __restrict__ rec, uint32_t* __restrict__ match,
                  const float4* __restrict__ tgt4, float4* __restrict__ cache, uint4* __restrict__ cache_idx) {
    if (st->done) return;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_src) return;
    if (i == 0) st->seq_count = 0u;                                         // counted again by this iteration's terms pass
    float R[9], t[3];
    load_Rt(st->T, R, t);
    const GridView g = make_view(slots, gpts, gp);
    const GridView gfine = make_view(fslots, fpts, gp + 1);
    const float4 s = src[i];
    const unsigned orig = BINNED ? __float_as_uint(s.w) : i;
    float x, y, z, d2; unsigned idx;
    transform_point(R, t, s, x, y, z);
    icp_match(x, y, z, i, g, gfine, gp, tgt4, cache, cache_idx, d2, idx);
    const bool keep = idx != B3D_NO_MATCH && !(sqrtf(d2) > thr);            // registration.cpp:337-338
    rec[orig] = make_float4(x, y, z, d2);
    match[orig] = keep ? idx : B3D_NO_MATCH;
}

struct MatchFlag { const uint32_t* m; __device__ unsigned operator()(unsigned i) const { return m[i] != B3D_NO_MATCH ? 1u : 0u; } };
struct CompactPairs {
    const float4* rec; const uint32_t* m; const float4* tgt4; float4* outP; float4* outQ;
    const float4* nrm4; float4* outN;              // point-to-plane replay only (else null)
    __device__ void operator()(unsigned i, unsigned prefix, unsigned flag) const {
        if (flag) {
            const uint32_t j = m[i];
            outP[prefix] = rec[i]; outQ[prefix] = tgt4[j];
            if (outN) outN[prefix] = nrm4[j];
        }
    }
};

constexpr int kSeqThreads = 1024;
constexpr int kSeqTile = 512;
constexpr int kSeqStride = kSeqTile + 1;      // +1: the 7 component arrays land in different banks

__global__ void __launch_bounds__(kSeqThreads)
icp_seq_p2p_kernel(const float4* __restrict__ P, const float4* __restrict__ Q, const unsigned* __restrict__ n_ptr,
                   int iter, float n_src_f, int stop_on_convergence, DeviceState* __restrict__ st) {
    if (st->done) return;
    __shared__ float tile[2][7][kSeqStride];           // px py pz qx qy qz d2
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned n = *n_ptr;
    if (n < 3u) { if (tid == 0) { st->n_corr_last = (int)n; st->done = 1; } return; }      // registration.cpp:361
    const unsigned n_tiles = (n + kSeqTile - 1) / kSeqTile;
    const bool consumer = warp == 31;                  // highest warp id wins issue arbitration on its sub-partition
    float sums[2] = {0.0f, 0.0f};                      // lane's running sum in pass 0 / pass 1
    float pm[3] = {0, 0, 0}, qm[3] = {0, 0, 0};
    for (int pass = 0; pass < 2; ++pass) {
        float acc = 0.0f;
        const int r = lane / 3, cc = lane % 3;         // pass 1: lane < 9 owns H(r, cc)
        for (unsigned t = 0; t <= n_tiles; ++t) {
            if (!consumer && tid < kSeqTile && t < n_tiles) {
                const unsigned k = t * kSeqTile + tid;
                float4 p = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
                if (k < n) { p = P[k]; q = Q[k]; }
                float (*b)[kSeqStride] = tile[t & 1];
                b[0][tid] = p.x; b[1][tid] = p.y; b[2][tid] = p.z; b[3][tid] = q.x; b[4][tid] = q.y; b[5][tid] = q.z; b[6][tid] = p.w;
            } else if (consumer && t > 0) {
                const float (*b)[kSeqStride] = tile[(t - 1) & 1];
                const unsigned m = min((unsigned)kSeqTile, n - (t - 1) * kSeqTile);
                // register ping-pong: the next 8 operands are loaded (and, in pass 1, centred and multiplied)
                // under the current 8 dependent adds; only real records are ever added (no padding terms)
                if (pass == 0) {
                    if (lane < 7) {                    // src_mean (3), tgt_mean (3), total_error
                        const float* v = b[lane];
                        unsigned k = 0;
                        if (m >= 16) {
                            float a[8], c8[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) a[j] = v[j];
                            for (; k + 24 <= m; k += 16) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) c8[j] = v[k + 8 + j];
#pragma unroll
                                for (int j = 0; j < 8; ++j) acc += a[j];
#pragma unroll
                                for (int j = 0; j < 8; ++j) a[j] = v[k + 16 + j];
#pragma unroll
                                for (int j = 0; j < 8; ++j) acc += c8[j];
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc += a[j];
                            k += 8;
                        }
                        for (; k < m; ++k) acc += v[k];
                    }
                } else if (lane < 9) {                 // H(r,cc) += (p_r - pm_r) * (q_cc - qm_cc)
                    const float* vp = b[r]; const float* vq = b[3 + cc];
                    const float mp = pm[r], mq = qm[cc];
                    unsigned k = 0;
                    if (m >= 16) {
                        float a[8], c8[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) a[j] = (vp[j] - mp) * (vq[j] - mq);
                        for (; k + 24 <= m; k += 16) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) c8[j] = (vp[k + 8 + j] - mp) * (vq[k + 8 + j] - mq);
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc += a[j];
#pragma unroll
                            for (int j = 0; j < 8; ++j) a[j] = (vp[k + 16 + j] - mp) * (vq[k + 16 + j] - mq);
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc += c8[j];
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc += a[j];
                        k += 8;
                    }
                    for (; k < m; ++k) { const float a1 = vp[k] - mp, b1 = vq[k] - mq; acc += a1 * b1; }
                }
            }
            __syncthreads();
        }
        sums[pass] = acc;
        if (pass == 0) {                               // means: sum / float(n), registration.cpp:380-381
            const float nf = (float)n;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                pm[a] = __shfl_sync(0xffffffffu, acc, a) / nf;
                qm[a] = __shfl_sync(0xffffffffu, acc, 3 + a) / nf;
            }
        }
    }
    if (!consumer) return;
    Mat3 H;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) H(a, b) = __shfl_sync(0xffffffffu, sums[1], a * 3 + b);
    const float total_error = __shfl_sync(0xffffffffu, sums[0], 6);
    if (lane != 0) return;
    st->n_corr_last = (int)n;
    Mat3 dR; rotation_from_cross_covariance(H, dR);      // registration.cpp:388-394
    float delta[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) delta[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    float r0, r1, r2; mat3_vec(dR, pm[0], pm[1], pm[2], r0, r1, r2);
    for (int rr = 0; rr < 3; ++rr) for (int c2 = 0; c2 < 3; ++c2) delta[c2 * 4 + rr] = dR(rr, c2);
    delta[12] = qm[0] - r0; delta[13] = qm[1] - r1; delta[14] = qm[2] - r2;             // registration.cpp:396
    float Tn[16];
    mat4_mul(delta, st->T, Tn);
    const float prev_rmse = st->res_rmse;
    const float rmse = sqrtf(total_error / (float)(int)n);
    for (int i = 0; i < 16; ++i) { st->T[i] = Tn[i]; st->res_T[i] = Tn[i]; }
    st->res_rmse = rmse;
    st->res_fitness = (float)(int)n / n_src_f;
    st->iterations = iter + 1;
    if (stop_on_convergence && iter > 0 && fabsf(prev_rmse - rmse) < 1e-6f) st->done = 1;
}

// Reference-order point-to-plane iteration (b3d_set_icp_mode(ctx, 2)): same pipeline as icp_seq_p2p_kernel for
// registration.cpp:343-354.  Producers form J = [p x n | n] and r = (p - q).n per matched record with the
// reference's un-fused arithmetic; lanes 0..20 of the consumer warp own the upper triangle of ATA (J_a*J_b is
// commutative, so the lower triangle is bit-identical), lanes 21..26 own ATb, lane 27 owns total_error, and each adds
// its products left to right in source order.  The 6x6 solve and the update are the fast path's.
__global__ void __launch_bounds__(kSeqThreads)
icp_seq_plane_kernel(const float4* __restrict__ P, const float4* __restrict__ Q, const float4* __restrict__ N,
                     const unsigned* __restrict__ n_ptr, int iter, float n_src_f, int stop_on_convergence,
                     DeviceState* __restrict__ st) {
    if (st->done) return;
    __shared__ float tile[2][8][kSeqStride];           // J0..J5, residual, d2
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned n = *n_ptr;
    if (n < 3u) { if (tid == 0) { st->n_corr_last = (int)n; st->done = 1; } return; }      // registration.cpp:361
    const unsigned n_tiles = (n + kSeqTile - 1) / kSeqTile;
    const bool consumer = warp == 31;
    int ra = 7, rb = 7; bool unit = true;              // lane 27 (and the idle lanes above it): d2 * 1
    if (lane < 21) { int k = lane, a = 0; while (k >= 6 - a) { k -= 6 - a; ++a; } ra = a; rb = a + k; unit = false; }
    else if (lane < 27) { ra = lane - 21; rb = 6; unit = false; }
    float acc = 0.0f;
    for (unsigned t = 0; t <= n_tiles; ++t) {
        if (!consumer && tid < kSeqTile && t < n_tiles) {
            const unsigned k = t * kSeqTile + tid;
            float J0 = 0, J1 = 0, J2 = 0, J3 = 0, J4 = 0, J5 = 0, res = 0, d2 = 0;
            if (k < n) {
                const float4 p = P[k], q = Q[k], nn = N[k];
                J0 = p.y * nn.z - p.z * nn.y; J1 = p.z * nn.x - p.x * nn.z; J2 = p.x * nn.y - p.y * nn.x;   // p.cross(n)
                J3 = nn.x; J4 = nn.y; J5 = nn.z;
                const float a0 = (p.x - q.x) * nn.x, a1 = (p.y - q.y) * nn.y, a2 = (p.z - q.z) * nn.z;
                res = a0 + (a1 + a2);                                                                        // (p - q).dot(n)
                d2 = p.w;
            }
            float (*b)[kSeqStride] = tile[t & 1];
            b[0][tid] = J0; b[1][tid] = J1; b[2][tid] = J2; b[3][tid] = J3; b[4][tid] = J4; b[5][tid] = J5; b[6][tid] = res; b[7][tid] = d2;
        } else if (consumer && t > 0 && lane < 28) {
            const float (*b)[kSeqStride] = tile[(t - 1) & 1];
            const unsigned m = min((unsigned)kSeqTile, n - (t - 1) * kSeqTile);
            const float* va = b[ra]; const float* vb = b[rb];
            unsigned k = 0;
            if (m >= 16) {
                float a[8], c8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = va[j] * (unit ? 1.0f : vb[j]);
                for (; k + 24 <= m; k += 16) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) c8[j] = va[k + 8 + j] * (unit ? 1.0f : vb[k + 8 + j]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc += a[j];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = va[k + 16 + j] * (unit ? 1.0f : vb[k + 16 + j]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc += c8[j];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) acc += a[j];
                k += 8;
            }
            for (; k < m; ++k) { const float prod = va[k] * (unit ? 1.0f : vb[k]); acc += prod; }
        }
        __syncthreads();
    }
    if (!consumer) return;
    float A[36], nb[6], x[6];
    {
        int k = 0;
        for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b) { const float val = __shfl_sync(0xffffffffu, acc, k++); A[a * 6 + b] = val; A[b * 6 + a] = val; }
        for (int a = 0; a < 6; ++a) nb[a] = -__shfl_sync(0xffffffffu, acc, 21 + a);
    }
    const float total_error = __shfl_sync(0xffffffffu, acc, 27);
    if (lane != 0) return;
    st->n_corr_last = (int)n;
    float delta[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) delta[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    ldlt6_solve(A, nb, x);                               // registration.cpp:366
    Mat3 dR; euler_xyz_to_matrix(x[0], x[1], x[2], dR);  // registration.cpp:369-371
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) delta[c * 4 + r] = dR(r, c); delta[12 + r] = x[3 + r]; }
    float Tn[16];
    mat4_mul(delta, st->T, Tn);
    const float prev_rmse = st->res_rmse;
    const float rmse = sqrtf(total_error / (float)(int)n);
    for (int i = 0; i < 16; ++i) { st->T[i] = Tn[i]; st->res_T[i] = Tn[i]; }
    st->res_rmse = rmse;
    st->res_fitness = (float)(int)n / n_src_f;
    st->iterations = iter + 1;
    if (stop_on_convergence && iter > 0 && fabsf(prev_rmse - rmse) < 1e-6f) st->done = 1;
}

// ---------------------------------------------------------------------------------
// Reference-order iteration, parallel (default): the sums of registration.cpp:341-358, 374-386 bit for bit without the
// dependent add chain, see b3d_ess.cuh.  The search kernel leaves (p, d2) and the match of every query at its ORIGINAL
// index; a terms pass forms each record's products exactly as the reference forms them (+0 for queries the reference
// skips, which leaves an fp32 running sum untouched, so the records need no compaction); fp64 block sums -> guesses ->
// integer block summaries; then one CTA per sum walks its summaries and the last CTA to arrive solves and updates T.
// ---------------------------------------------------------------------------------
struct PlaneTerms {                                   // ATA upper triangle (21), ATb (6), total_error
    const float4* rec; const uint32_t* match; const float4* tgt4; const float4* nrm4;
    __device__ __forceinline__ bool operator()(unsigned k, float (&t)[kAccPlane]) const {
        const uint32_t j = match[k];
        if (j == B3D_NO_MATCH) {
#pragma unroll
            for (int v = 0; v < kAccPlane; ++v) t[v] = 0.0f;
            return false;
        }
        const float4 p = rec[k], q = tgt4[j], n = nrm4[j];
        float J[6];
        J[0] = p.y * n.z - p.z * n.y; J[1] = p.z * n.x - p.x * n.z; J[2] = p.x * n.y - p.y * n.x;      // p.cross(n), registration.cpp:346
        J[3] = n.x; J[4] = n.y; J[5] = n.z;
        const float a0 = (p.x - q.x) * n.x, a1 = (p.y - q.y) * n.y, a2 = (p.z - q.z) * n.z;
        const float r = a0 + (a1 + a2);                                                                   // (p - q).dot(n)
        int c = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) t[c++] = J[a] * J[b];
#pragma unroll
        for (int a = 0; a < 6; ++a) t[21 + a] = J[a] * r;
        t[27] = p.w;
        return true;
    }
};
constexpr int kAccPoint0 = 7, kAccPoint1 = 9;
struct PointTerms0 {                                  // src_mean sums (3), tgt_mean sums (3), total_error
    const float4* rec; const uint32_t* match; const float4* tgt4;
    __device__ __forceinline__ bool operator()(unsigned k, float (&t)[kAccPoint0]) const {
        const uint32_t j = match[k];
        if (j == B3D_NO_MATCH) {
#pragma unroll
            for (int v = 0; v < kAccPoint0; ++v) t[v] = 0.0f;
            return false;
        }
        const float4 p = rec[k], q = tgt4[j];
        t[0] = p.x; t[1] = p.y; t[2] = p.z; t[3] = q.x; t[4] = q.y; t[5] = q.z; t[6] = p.w;
        return true;
    }
};
struct PointTerms1 {                                  // H(r, c) += (p_r - src_mean_r) * (q_c - tgt_mean_c), registration.cpp:384-386
    const float4* rec; const uint32_t* match; const float4* tgt4; const DeviceState* st;
    __device__ __forceinline__ bool operator()(unsigned k, float (&t)[kAccPoint1]) const {
        const uint32_t j = match[k];
        if (j == B3D_NO_MATCH) {
#pragma unroll
            for (int v = 0; v < kAccPoint1; ++v) t[v] = 0.0f;
            return false;
        }
        const float4 p = rec[k], q = tgt4[j];
        const float pc[3] = {p.x - st->ess_means[0], p.y - st->ess_means[1], p.z - st->ess_means[2]};
        const float qc[3] = {q.x - st->ess_means[3], q.y - st->ess_means[4], q.z - st->ess_means[5]};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) t[a * 3 + b] = pc[a] * qc[b];
        return true;
    }
};

__device__ __forceinline__ void icp_commit(DeviceState* st, const float (&delta)[16], float total_error, int n, int iter, float n_src_f, int stop_on_convergence) {
    float Tn[16];
    mat4_mul(delta, st->T, Tn);                              // T = delta * T
    const float prev_rmse = st->res_rmse;
    const float rmse = sqrtf(total_error / (float)n);
    for (int i = 0; i < 16; ++i) { st->T[i] = Tn[i]; st->res_T[i] = Tn[i]; }
    st->res_rmse = rmse;
    st->res_fitness = (float)n / n_src_f;
    st->iterations = iter + 1;
    if (stop_on_convergence && iter > 0 && fabsf(prev_rmse - rmse) < 1e-6f) st->done = 1;    // registration.cpp:406
}

// PHASE 0: point-to-plane sums -> solve.  1: point-to-point first pass -> means.  2: point-to-point second pass -> solve.
template <int NV, int PHASE>
__global__ void __launch_bounds__(ess::kChainThreads)
icp_ess_chain_kernel(const float* __restrict__ terms, size_t stride, const ess::BlockSummary* __restrict__ summ, unsigned nb_stride, unsigned n,
                     int iter, float n_src_f, int stop_on_convergence, DeviceState* __restrict__ st) {
    if (st->done) return;
    extern __shared__ __align__(128) unsigned char ess_smem[];
    ess::ChainSmem& sm = *reinterpret_cast<ess::ChainSmem*>(ess_smem);
    const unsigned v = blockIdx.x;
    const float s = ess::chain(terms + (size_t)v * stride, summ + (size_t)v * nb_stride, n, sm, st->ess_stats[(PHASE == 2 ? 16 : 0) + v]);
    if (threadIdx.x != 0) return;
    st->ess_sums[(PHASE == 2 ? 16 : 0) + v] = s;
    __threadfence();
    if (atomicAdd(&st->ess_arrive, 1u) != (unsigned)(NV - 1)) return;      // the last CTA to arrive solves
    __threadfence();
    st->ess_arrive = 0u;
    const volatile float* sums = st->ess_sums;
    const int n_corr = (int)st->seq_count;
    st->n_corr_last = n_corr;
    if (n_corr < 3) { st->done = 1; return; }              // registration.cpp:361
    float delta[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) delta[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    if (PHASE == 0) {
        float A[36], nb[6], x[6];
        int k = 0;
        for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b) { const float val = sums[k++]; A[a * 6 + b] = val; A[b * 6 + a] = val; }
        for (int a = 0; a < 6; ++a) nb[a] = -sums[21 + a];
        ldlt6_solve(A, nb, x);                               // registration.cpp:366
        Mat3 dR; euler_xyz_to_matrix(x[0], x[1], x[2], dR);  // registration.cpp:369-371
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) delta[c * 4 + r] = dR(r, c); delta[12 + r] = x[3 + r]; }
        icp_commit(st, delta, sums[27], n_corr, iter, n_src_f, stop_on_convergence);
    } else if (PHASE == 1) {
        const float nf = (float)n_corr;                      // means: sum / float(n), registration.cpp:380-381
        for (int a = 0; a < 6; ++a) st->ess_means[a] = sums[a] / nf;
    } else {
        Mat3 H;
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) H(a, b) = sums[16 + a * 3 + b];
        Mat3 dR; rotation_from_cross_covariance(H, dR);      // registration.cpp:388-394
        const float pm0 = st->ess_means[0], pm1 = st->ess_means[1], pm2 = st->ess_means[2];
        float r0, r1, r2; mat3_vec(dR, pm0, pm1, pm2, r0, r1, r2);
        for (int rr = 0; rr < 3; ++rr) for (int c2 = 0; c2 < 3; ++c2) delta[c2 * 4 + rr] = dR(rr, c2);
        delta[12] = st->ess_means[3] - r0; delta[13] = st->ess_means[4] - r1; delta[14] = st->ess_means[5] - r2;     // registration.cpp:396
        icp_commit(st, delta, sums[6], n_corr, iter, n_src_f, stop_on_convergence);
    }
}

__global__ void icp_state_init_kernel(DeviceState* st, const float* __restrict__ T0) {
    int i = threadIdx.x;
    if (i < 16) { st->T[i] = T0[i]; st->res_T[i] = T0[i]; }
    if (i == 0) { st->res_fitness = 0.0f; st->res_rmse = 0.0f; st->iterations = 0; st->done = 0; st->n_corr_last = 0; st->ess_arrive = 0u; st->seq_count = 0u; }
    if (i < 32) { st->ess_stats[i][0] = 0u; st->ess_stats[i][1] = 0u; st->ess_stats[i][2] = 0u; st->ess_stats[i][3] = 0u; }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static unsigned pow2_at_least(size_t v) { unsigned c = 1024; while ((size_t)c < v) c <<= 1; return c; }

static int build_grid(b3d_ctx* c, float thr, GridParams** gp_out, unsigned* capacity_out) {
    StageTimer timer(c, 4);
    const unsigned n = (unsigned)c->n_tgt;
    const unsigned capacity = pow2_at_least(2 * (size_t)n);
    B3D_CUDA(c, c->grid_slots.ensure(sizeof(CellSlot) * capacity));
    B3D_CUDA(c, c->grid_cursor.ensure(3 * sizeof(GridParams)));          // [0] coarse (cell = 1.02 thr), [1] fine, [2] the one the source is binned by
    B3D_CUDA(c, c->fine_slots.ensure(sizeof(CellSlot) * 16));
    B3D_CUDA(c, c->fine_pts.ensure(sizeof(float4) * 16));
    B3D_CUDA(c, c->grid_pts.ensure(sizeof(float4) * (n ? n : 1)));
    B3D_CUDA(c, c->grid_nrm.ensure(sizeof(float4) * (n ? n : 1)));
    B3D_CUDA(c, c->pt_slot.ensure(sizeof(unsigned) * (n ? n : 1)));
    B3D_CUDA(c, c->pt_rank.ensure(sizeof(unsigned) * (n ? n : 1)));
    GridParams* gp = c->grid_cursor.as<GridParams>();
    CellSlot* slots = c->grid_slots.as<CellSlot>();
    B3D_CUDA(c, cudaMemsetAsync(gp, 0, 3 * sizeof(GridParams), c->stream));
    if (n) { grid_bounds_kernel<<<grid_for(n, 256, 2), 256, 0, c->stream>>>(c->tgt4.as<float4>(), n, gp); B3D_LAUNCHED(c); }
    grid_init_kernel<<<grid_for(capacity, 256, 4), 256, 0, c->stream>>>(slots, capacity, gp, thr, n);
    B3D_LAUNCHED(c);
    if (n) {
        grid_insert_kernel<<<grid_for(n, 256, 8), 256, 0, c->stream>>>(c->tgt4.as<float4>(), n, slots, gp, capacity - 1u, nullptr,
                                                                       c->pt_slot.as<unsigned>(), c->pt_rank.as<unsigned>(), &gp->occupied);
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
        grid_scatter_kernel<<<grid_for(n, 256, 8), 256, 0, c->stream>>>(c->tgt4.as<float4>(), c->has_normals ? c->nrm4.as<float4>() : nullptr, n, slots,
                                                                        c->pt_slot.as<unsigned>(), c->pt_rank.as<unsigned>(),
                                                                        c->grid_pts.as<float4>(), c->grid_nrm.as<float4>(), nullptr);
        B3D_LAUNCHED(c);
        fine_params_kernel<<<1, 32, 0, c->stream>>>(gp, gp + 1, 0u, n);     // decides the second level's cell; lists are built lazily
        B3D_LAUNCHED(c);
    }
    binning_params_kernel<<<1, 32, 0, c->stream>>>(gp);
    B3D_LAUNCHED(c);
    *gp_out = gp; *capacity_out = capacity;
    return B3D_OK;
}

// Launch with programmatic stream serialization: the kernel may become resident while its predecessor is still running and
// blocks in cudaGridDependencySynchronize() until that one has completed — the ~2-3 us of launch latency between the two
// dependent kernels of an ICP iteration disappear behind the predecessor.
template <class... KArgs, class... Args>
static cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Second level, built only for calls that are still iterating after kListsAfter iterations: refinements that converge in
// a handful of iterations never pay for it, long runs amortise it within a few iterations.
constexpr int kListsAfter = 8;
static int build_neighbourhood_lists(b3d_ctx* c, GridParams* gp, bool* built) {
    *built = false;
    const unsigned n = (unsigned)c->n_tgt;
    GridParams* gf = gp + 1;
    GridParams h[2];
    B3D_CUDA(c, cudaMemcpyAsync(h, gp, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!h[1].occupied || n == 0) return B3D_OK;
    // table size: in the sparse ("complete") case 27 * occupied coarse cells bounds the number of list cells exactly; in the
    // dense case a surface needs ~1.3 n and a solid ~7 n cells — 8 n slots, and a fill-up only disables the level
    const size_t want = h[1].complete ? 2 * 27 * (size_t)h[0].occupied : 8 * (size_t)n;
    if (want > (1ull << 28)) return B3D_OK;                    // would not fit a 32-bit slot index comfortably: keep the walk
    const unsigned ncap = pow2_at_least(want);
    // footprint: table + cursor per slot, list entry + slot id per (point, cell); a very large sparse target is not worth 4 GB
    if ((size_t)ncap * (sizeof(CellSlot) + sizeof(unsigned)) + 27 * (size_t)n * (sizeof(float4) + sizeof(unsigned)) > (4ull << 30)) return B3D_OK;
    B3D_CUDA(c, c->fine_slots.ensure(sizeof(CellSlot) * ncap));
    B3D_CUDA(c, c->fine_pts.ensure(sizeof(float4) * 27 * (size_t)n));
    B3D_CUDA(c, c->nbh_slot27.ensure(sizeof(unsigned) * 27 * (size_t)n));
    B3D_CUDA(c, c->nbh_cursor.ensure(sizeof(unsigned) * ncap));
    CellSlot* nslots = c->fine_slots.as<CellSlot>();
    const unsigned nmask = ncap - 1u, one = 1u;
    B3D_CUDA(c, cudaMemcpyAsync(&gf->mask, &nmask, sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
    B3D_CUDA(c, cudaMemcpyAsync(&gf->enabled, &one, sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
    slots_clear_kernel<<<grid_for(ncap, 256, 4), 256, 0, c->stream>>>(nslots, ncap);
    B3D_LAUNCHED(c);
    B3D_CUDA(c, cudaMemsetAsync(c->nbh_cursor.p, 0, sizeof(unsigned) * ncap, c->stream));
    neighbourhood_count_kernel<<<grid_for(27ll * n, 256, 16), 256, 0, c->stream>>>(c->tgt4.as<float4>(), n, nslots, gf, c->nbh_slot27.as<unsigned>());
    B3D_LAUNCHED(c);
    const unsigned ntiles = (unsigned)div_up(ncap, kScanTile);
    B3D_CUDA(c, c->scan_tmp.ensure(sizeof(unsigned) * (ntiles + 1)));
    SlotCount fcnt{nslots}; SlotStart fst{nslots};
    scan_tile_sums_kernel<<<ntiles, kScanThreads, 0, c->stream>>>(fcnt, ncap, c->scan_tmp.as<unsigned>());
    B3D_LAUNCHED(c);
    scan_tile_offsets_kernel<<<1, kScanThreads, 0, c->stream>>>(c->scan_tmp.as<unsigned>(), ntiles, (unsigned*)nullptr);
    B3D_LAUNCHED(c);
    scan_emit_kernel<<<ntiles, kScanThreads, 0, c->stream>>>(fcnt, fst, ncap, c->scan_tmp.as<unsigned>());
    B3D_LAUNCHED(c);
    neighbourhood_fill_kernel<<<grid_for(27ll * n, 256, 16), 256, 0, c->stream>>>(c->tgt4.as<float4>(), n, nslots, gf, c->nbh_slot27.as<unsigned>(),
                                                                                c->nbh_cursor.as<unsigned>(), c->fine_pts.as<float4>());
    B3D_LAUNCHED(c);
    binning_params_kernel<<<1, 32, 0, c->stream>>>(gp);
    B3D_LAUNCHED(c);
    *built = true;
    return B3D_OK;
}

// Reorder the source so that queries falling into the same target cell (under the initial transform)
// are adjacent: the lanes of a warp then probe the same hash slots and scan the same cell points
// (L1 hits / broadcast instead of scattered L2 reads) and run similar trip counts.  Pure scheduling:
// the sums are order-independent up to fp64 rounding and every per-point result is unchanged.
static int bin_source_by_cell(b3d_ctx* c, const GridParams* gp, const float* T_dev, const float4** src_out) {
    const unsigned n = (unsigned)c->n_src;
    const unsigned capacity = pow2_at_least(2 * (size_t)n);
    B3D_CUDA(c, c->src_slots.ensure(sizeof(CellSlot) * capacity));
    B3D_CUDA(c, c->src_sorted.ensure(sizeof(float4) * n));
    B3D_CUDA(c, c->src_slot.ensure(sizeof(unsigned) * n));
    B3D_CUDA(c, c->src_rank.ensure(sizeof(unsigned) * n));
    CellSlot* slots = c->src_slots.as<CellSlot>();
    slots_clear_kernel<<<grid_for(capacity, 256, 4), 256, 0, c->stream>>>(slots, capacity);
    B3D_LAUNCHED(c);
    grid_insert_kernel<<<grid_for(n, 256, 8), 256, 0, c->stream>>>(c->src4.as<float4>(), n, slots, gp, capacity - 1u, T_dev,
                                                                   c->src_slot.as<unsigned>(), c->src_rank.as<unsigned>(), nullptr);
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
    grid_scatter_kernel<<<grid_for(n, 256, 8), 256, 0, c->stream>>>(c->src4.as<float4>(), nullptr, n, slots, c->src_slot.as<unsigned>(),
                                                                    c->src_rank.as<unsigned>(), c->src_sorted.as<float4>(), nullptr, nullptr);
    B3D_LAUNCHED(c);
    *src_out = c->src_sorted.as<float4>();
    return B3D_OK;
}

int icp_run_impl(b3d_ctx* c, const float* T0, float thr, int max_iter, int p2plane, int stop_on_conv,
                 float* T, float* fitness, float* rmse, int32_t* iters) {
    if (!c->have_clouds) return fail(c, B3D_ERR_STATE, "icp_run: clouds not set");
    DeviceState* st = c->state.as<DeviceState>();
    // RegistrationResult starts as {initial_transform, 0, 0} (registration.cpp:309-311)
    for (int i = 0; i < 16; ++i) T[i] = T0[i];
    *fitness = 0.0f; *rmse = 0.0f; if (iters) *iters = 0;
    if (max_iter <= 0 || c->n_src == 0 || c->n_tgt == 0) return B3D_OK;    // n_corr < 3 at iteration 0 => break
    const bool plane = p2plane && c->has_normals;            // registration.cpp:343
    GridParams* gp; unsigned capacity;
    int rc = build_grid(c, thr, &gp, &capacity);
    if (rc != B3D_OK) return rc;

    for (int i = 0; i < 16; ++i) c->h_state->out18[i] = T0[i];
    B3D_CUDA(c, cudaMemcpyAsync(st->out18, c->h_state->out18, sizeof(float) * 16, cudaMemcpyHostToDevice, c->stream));
    icp_state_init_kernel<<<1, 32, 0, c->stream>>>(st, st->out18);
    B3D_LAUNCHED(c);

    const unsigned n_src = (unsigned)c->n_src;
    const float4* src = c->src4.as<float4>();
    const bool binned = n_src >= 16384;
    const unsigned seq_tiles = (unsigned)div_up(n_src, kScanTile);
    // icp_mode 0 / 2 (default): sums in the reference's order, parallel and exact (b3d_ess.cuh); 1: fp64 tree sums (fast,
    // order-free, tolerance-level); 3: the one-chain replay kernels (one dependent add per matched point; kept as an
    // independent cross-check of the parallel form at sizes the CPU oracle cannot reach)
    const bool tree = c->icp_mode == 1;
    const bool legacy = c->icp_mode == 3;
    const bool exact = !tree && !legacy;
    const bool replay = legacy;
    const size_t ess_stride = ess::padded_terms(n_src);
    const unsigned ess_nb_stride = (unsigned)(ess_stride / ess::kBlock);
    if (!tree) { B3D_CUDA(c, c->seq_rec.ensure(sizeof(float4) * n_src)); B3D_CUDA(c, c->seq_match.ensure(sizeof(uint32_t) * n_src)); }
    if (legacy) {
        if (plane) B3D_CUDA(c, c->seq_N.ensure(sizeof(float4) * n_src));
        B3D_CUDA(c, c->seq_P.ensure(sizeof(float4) * n_src));   B3D_CUDA(c, c->seq_Q.ensure(sizeof(float4) * n_src));
        B3D_CUDA(c, c->scan_tmp.ensure(sizeof(unsigned) * (seq_tiles + 1)));
    }
    if (exact) {
        if (!c->ess_smem_opt_in) {                                // per device: the chain kernels' TMA ring is larger than 48 KB
            B3D_CUDA(c, cudaFuncSetAttribute(icp_ess_chain_kernel<kAccPlane, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ess::ChainSmem)));
            B3D_CUDA(c, cudaFuncSetAttribute(icp_ess_chain_kernel<kAccPoint0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ess::ChainSmem)));
            B3D_CUDA(c, cudaFuncSetAttribute(icp_ess_chain_kernel<kAccPoint1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ess::ChainSmem)));
            c->ess_smem_opt_in = true;
        }
        const size_t nv = plane ? kAccPlane : kAccPoint1;
        B3D_CUDA(c, c->ess_terms.ensure(sizeof(float) * nv * ess_stride));
        B3D_CUDA(c, c->ess_bsum.ensure(sizeof(double) * nv * ess_nb_stride));
        B3D_CUDA(c, c->ess_guess.ensure(sizeof(double) * nv * (ess_stride / ess::kSuperTerms)));      // super-block sums
        B3D_CUDA(c, c->ess_summ.ensure(sizeof(ess::BlockSummary) * nv * ess_nb_stride));
    }
    if (binned) {                                            // below that the reorder costs more than it saves
        StageTimer timer(c, 6);
        rc = bin_source_by_cell(c, gp + 2, st->out18, &src);
        if (rc != B3D_OK) return rc;
    }
    const int blocks = div_up(n_src, kIcpThreads);           // one query per thread
    B3D_CUDA(c, c->partials.ensure(sizeof(double) * kPartialStride * (size_t)blocks));
    B3D_CUDA(c, c->icp_cache.ensure(sizeof(float4) * n_src)); B3D_CUDA(c, c->icp_cache_idx.ensure(sizeof(uint4) * n_src));
    float4* cache = c->icp_cache.as<float4>(); uint4* cache_idx = c->icp_cache_idx.as<uint4>();
    B3D_CUDA(c, cudaMemsetAsync(cache, 0, sizeof(float4) * n_src, c->stream));      // hold2 = 0: nothing to re-use yet
    {
        StageTimer timer(c, 5);
        const CellSlot* slots = c->grid_slots.as<CellSlot>();
        for (int iter = 0; iter < max_iter; ++iter) {
            if (plane && tree) {
                B3D_CUDA(c, launch_dependent(icp_accumulate_kernel<true>, dim3(blocks), dim3(kIcpThreads), c->stream, src, n_src, (const DeviceState*)st, thr,
                                             slots, (const float4*)c->grid_pts.as<float4>(), (const CellSlot*)c->fine_slots.as<CellSlot>(),
                                             (const float4*)c->fine_pts.as<float4>(), (const float4*)c->tgt4.as<float4>(), (const float4*)c->nrm4.as<float4>(),
                                             (const GridParams*)gp, c->partials.as<double>(), cache, cache_idx));
                B3D_LAUNCHED(c);
                B3D_CUDA(c, launch_dependent(icp_update_kernel<true>, dim3(1), dim3(kUpdateThreads), c->stream, (const double*)c->partials.as<double>(), blocks, iter,
                                             (float)c->n_src, stop_on_conv, st));
                B3D_LAUNCHED(c);
            } else if (exact) {
                // reference-order sums, parallel: search -> terms (+ fp64 block sums) -> guesses -> summaries -> chains + solve
                if (binned) icp_search_kernel<true><<<blocks, kIcpThreads, 0, c->stream>>>(src, n_src, st, thr, slots, c->grid_pts.as<float4>(), c->fine_slots.as<CellSlot>(),
                                                                                           c->fine_pts.as<float4>(), gp, c->seq_rec.as<float4>(), c->seq_match.as<uint32_t>(),
                                                                                           c->tgt4.as<float4>(), cache, cache_idx);
                else        icp_search_kernel<false><<<blocks, kIcpThreads, 0, c->stream>>>(src, n_src, st, thr, slots, c->grid_pts.as<float4>(), c->fine_slots.as<CellSlot>(),
                                                                                            c->fine_pts.as<float4>(), gp, c->seq_rec.as<float4>(), c->seq_match.as<uint32_t>(),
                                                                                           c->tgt4.as<float4>(), cache, cache_idx);
                B3D_LAUNCHED(c);
                float* terms = c->ess_terms.as<float>(); double* bsum = c->ess_bsum.as<double>(); double* ssum = c->ess_guess.as<double>();
                ess::BlockSummary* summ = c->ess_summ.as<ess::BlockSummary>();
                const int term_blocks = div_up(n_src, ess::kTermsThreads);
                const unsigned summ_blocks = (unsigned)div_up(term_blocks, ess::kSummaryWarps);
                if (plane) {
                    PlaneTerms fn{c->seq_rec.as<float4>(), c->seq_match.as<uint32_t>(), c->tgt4.as<float4>(), c->nrm4.as<float4>()};
                    ess::terms_kernel<kAccPlane><<<term_blocks, ess::kTermsThreads, 0, c->stream>>>(fn, n_src, &st->done, terms, ess_stride, bsum, ssum, &st->seq_count);
                    B3D_LAUNCHED(c);
                    ess::summary_kernel<<<dim3(summ_blocks, kAccPlane), ess::kSummaryWarps * 32, 0, c->stream>>>(terms, ess_stride, bsum, ssum, n_src, &st->done, summ);
                    B3D_LAUNCHED(c);
                    icp_ess_chain_kernel<kAccPlane, 0><<<kAccPlane, ess::kChainThreads, sizeof(ess::ChainSmem), c->stream>>>(terms, ess_stride, summ, ess_nb_stride, n_src, iter, (float)c->n_src, stop_on_conv, st);
                    B3D_LAUNCHED(c);
                } else {
                    PointTerms0 fn0{c->seq_rec.as<float4>(), c->seq_match.as<uint32_t>(), c->tgt4.as<float4>()};
                    ess::terms_kernel<kAccPoint0><<<term_blocks, ess::kTermsThreads, 0, c->stream>>>(fn0, n_src, &st->done, terms, ess_stride, bsum, ssum, &st->seq_count);
                    B3D_LAUNCHED(c);
                    ess::summary_kernel<<<dim3(summ_blocks, kAccPoint0), ess::kSummaryWarps * 32, 0, c->stream>>>(terms, ess_stride, bsum, ssum, n_src, &st->done, summ);
                    B3D_LAUNCHED(c);
                    icp_ess_chain_kernel<kAccPoint0, 1><<<kAccPoint0, ess::kChainThreads, sizeof(ess::ChainSmem), c->stream>>>(terms, ess_stride, summ, ess_nb_stride, n_src, iter, (float)c->n_src, stop_on_conv, st);
                    B3D_LAUNCHED(c);
                    PointTerms1 fn1{c->seq_rec.as<float4>(), c->seq_match.as<uint32_t>(), c->tgt4.as<float4>(), st};
                    ess::terms_kernel<kAccPoint1><<<term_blocks, ess::kTermsThreads, 0, c->stream>>>(fn1, n_src, &st->done, terms, ess_stride, bsum, ssum, (unsigned*)nullptr);
                    B3D_LAUNCHED(c);
                    ess::summary_kernel<<<dim3(summ_blocks, kAccPoint1), ess::kSummaryWarps * 32, 0, c->stream>>>(terms, ess_stride, bsum, ssum, n_src, &st->done, summ);
                    B3D_LAUNCHED(c);
                    icp_ess_chain_kernel<kAccPoint1, 2><<<kAccPoint1, ess::kChainThreads, sizeof(ess::ChainSmem), c->stream>>>(terms, ess_stride, summ, ess_nb_stride, n_src, iter, (float)c->n_src, stop_on_conv, st);
                    B3D_LAUNCHED(c);
                }
            } else if (replay) {
                // reference-order sums: search -> ordered compaction -> sequential replay (see icp_seq_p2p_kernel)
                if (binned) icp_search_kernel<true><<<blocks, kIcpThreads, 0, c->stream>>>(src, n_src, st, thr, slots, c->grid_pts.as<float4>(), c->fine_slots.as<CellSlot>(),
                                                                                           c->fine_pts.as<float4>(), gp, c->seq_rec.as<float4>(), c->seq_match.as<uint32_t>(),
                                                                                           c->tgt4.as<float4>(), cache, cache_idx);
                else        icp_search_kernel<false><<<blocks, kIcpThreads, 0, c->stream>>>(src, n_src, st, thr, slots, c->grid_pts.as<float4>(), c->fine_slots.as<CellSlot>(),
                                                                                            c->fine_pts.as<float4>(), gp, c->seq_rec.as<float4>(), c->seq_match.as<uint32_t>(),
                                                                                           c->tgt4.as<float4>(), cache, cache_idx);
                B3D_LAUNCHED(c);
                MatchFlag flag{c->seq_match.as<uint32_t>()};
                CompactPairs emit{c->seq_rec.as<float4>(), c->seq_match.as<uint32_t>(), c->tgt4.as<float4>(), c->seq_P.as<float4>(), c->seq_Q.as<float4>(),
                                  plane ? c->nrm4.as<float4>() : nullptr, plane ? c->seq_N.as<float4>() : nullptr};
                scan_tile_sums_kernel<<<seq_tiles, kScanThreads, 0, c->stream>>>(flag, n_src, c->scan_tmp.as<unsigned>());
                B3D_LAUNCHED(c);
                scan_tile_offsets_kernel<<<1, kScanThreads, 0, c->stream>>>(c->scan_tmp.as<unsigned>(), seq_tiles, &st->seq_count);
                B3D_LAUNCHED(c);
                scan_emit_kernel<<<seq_tiles, kScanThreads, 0, c->stream>>>(flag, emit, n_src, c->scan_tmp.as<unsigned>());
                B3D_LAUNCHED(c);
                if (plane) icp_seq_plane_kernel<<<1, kSeqThreads, 0, c->stream>>>(c->seq_P.as<float4>(), c->seq_Q.as<float4>(), c->seq_N.as<float4>(), &st->seq_count, iter, (float)c->n_src, stop_on_conv, st);
                else       icp_seq_p2p_kernel<<<1, kSeqThreads, 0, c->stream>>>(c->seq_P.as<float4>(), c->seq_Q.as<float4>(), &st->seq_count, iter, (float)c->n_src, stop_on_conv, st);
                B3D_LAUNCHED(c);
            } else {
                icp_accumulate_kernel<false><<<blocks, kIcpThreads, 0, c->stream>>>(src, n_src, st, thr, slots, c->grid_pts.as<float4>(),
                                                                                    c->fine_slots.as<CellSlot>(), c->fine_pts.as<float4>(),
                                                                                    c->tgt4.as<float4>(), c->nrm4.as<float4>(), gp, c->partials.as<double>(),
                                                                                   cache, cache_idx);
                B3D_LAUNCHED(c);
                icp_update_kernel<false><<<1, kUpdateThreads, 0, c->stream>>>(c->partials.as<double>(), blocks, iter, (float)c->n_src, stop_on_conv, st);
                B3D_LAUNCHED(c);
            }
            // poll the device-side done flag now and then so converged runs stop launching
            // the second level pays for itself within a few iterations when the queries outnumber the targets (its build is
            // O(27 n_tgt), a search O(n_src)): then build it after the first two iterations (before that the queries are too far
            // from the surface for the fine level to settle them); otherwise only for calls still iterating after kListsAfter
            const int lists_at = (max_iter >= 2 * kListsAfter && c->n_src >= 2 * c->n_tgt) ? 2 : kListsAfter;
            // polls: after iterations 2 and 4 (a refinement behind RANSAC converges in a handful of iterations; every launch behind
            // the convergence point would be an empty kernel), at the second level's build point, then every 16th iteration
            if (((iter & 15) == 15 || iter == lists_at - 1 || ((iter == 2 || iter == 4) && stop_on_conv)) && iter + 1 < max_iter) {
                B3D_CUDA(c, cudaMemcpyAsync(&c->h_state->done, &st->done, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                B3D_CUDA(c, cudaStreamSynchronize(c->stream));
                if (c->h_state->done) break;
                if (iter == lists_at - 1) {                         // a long run: build the second level now and re-bin by its cells
                    bool built = false;
                    rc = build_neighbourhood_lists(c, gp, &built);
                    if (rc != B3D_OK) return rc;
                    if (built && binned) {
                        rc = bin_source_by_cell(c, gp + 2, st->T, &src); if (rc != B3D_OK) return rc;
                        B3D_CUDA(c, cudaMemsetAsync(cache, 0, sizeof(float4) * n_src, c->stream));      // query slots were re-ordered
                    }
                }
            }
        }
    }
    B3D_CUDA(c, cudaMemcpyAsync(c->h_state, st, sizeof(DeviceState), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 16; ++i) T[i] = c->h_state->res_T[i];
    *fitness = c->h_state->res_fitness; *rmse = c->h_state->res_rmse;
    if (iters) *iters = c->h_state->iterations;
    return B3D_OK;
}

int icp_nearest_impl(b3d_ctx* c, const float* T, float thr, uint32_t* idx_host, float* d2_host) {
    if (!c->have_clouds) return fail(c, B3D_ERR_STATE, "icp_nearest: clouds not set");
    if (c->n_src == 0) return B3D_OK;
    DeviceState* st = c->state.as<DeviceState>();
    GridParams* gp; unsigned capacity;
    int rc = build_grid(c, thr, &gp, &capacity);
    if (rc != B3D_OK) return rc;
    const unsigned n_src = (unsigned)c->n_src;
    B3D_CUDA(c, c->nn_idx.ensure(sizeof(uint32_t) * n_src));
    B3D_CUDA(c, c->nn_d2.ensure(sizeof(float) * n_src));
    for (int i = 0; i < 16; ++i) c->h_state->out18[i] = T[i];
    B3D_CUDA(c, cudaMemcpyAsync(st->out18, c->h_state->out18, sizeof(float) * 16, cudaMemcpyHostToDevice, c->stream));
    icp_nearest_kernel<<<grid_for(n_src, 256, 8), 256, 0, c->stream>>>(c->src4.as<float4>(), n_src, st->out18, thr, c->grid_slots.as<CellSlot>(),
                                                                       c->grid_pts.as<float4>(), gp, c->nn_idx.as<uint32_t>(), c->nn_d2.as<float>());
    B3D_LAUNCHED(c);
    B3D_CUDA(c, cudaMemcpyAsync(idx_host, c->nn_idx.p, sizeof(uint32_t) * n_src, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaMemcpyAsync(d2_host, c->nn_d2.p, sizeof(float) * n_src, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    return B3D_OK;
}

}  // namespace b3d
