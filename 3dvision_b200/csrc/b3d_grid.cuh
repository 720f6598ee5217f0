// b3d_grid.cuh — voxel-hash uniform grid shared by ICP (b3d_icp.cu) and the feature stages (b3d_features.cu).
// Cells live in an open-addressing hash table (64-bit packed cell key); points claim a slot with atomicCAS and a
// rank inside the cell with a warp-aggregated atomicAdd (one atomic per distinct cell per warp); a prefix sum over
// the slot counts gives cell starts; points are then scattered into cell order as float4 (x, y, z, original index).
// Kernels are `static` so both translation units can include this file.
#pragma once
#include "b3d_common.cuh"
#include "b3d_scan.cuh"
#include <float.h>
#include <math.h>

namespace b3d {

// ---------------------------------------------------------------------------------
// voxel-hash grid
// ---------------------------------------------------------------------------------
struct __align__(16) CellSlot {
    unsigned long long key;     // packed cell coordinate, kEmptyKey if unused
    unsigned start;             // first point of the cell in grid_pts
    unsigned count;             // points in the cell
};
constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kCoordBias = 1 << 20;
constexpr float kCoordLimit = 65536.0f;     // |cell coordinate| bound that keeps fl(p*inv) within 2^-7 cell

struct GridParams {
    float inv_cell;
    float cell;                 // 1 / inv_cell
    float slack;                // conservative bound on how far a point can sit outside its nominal cell (rounding)
    unsigned mask;              // capacity - 1
    unsigned max_abs_bits;      // max |coordinate| over the target, as float bits
    unsigned n_points;
    unsigned occupied;          // cells in use (counted while inserting)
    unsigned enabled;           // fine grid only: 0 when it would not pay off
    float accept2;              // fine grid only: a match closer than sqrt(accept2) is provably the global nearest
    unsigned complete;          // fine level only: its cell is the coarse cell, so its lists hold every match the reference keeps
    unsigned overflow;          // fine level only: 1 if its neighbourhood table filled up (level unusable for this call)
};

__device__ __forceinline__ int cell_coord(float v, float inv_cell) {
    float f = floorf(v * inv_cell);
    f = fminf(fmaxf(f, -(float)(kCoordBias - 2)), (float)(kCoordBias - 2));
    return (int)f;
}
__device__ __forceinline__ unsigned long long pack_cell(int cx, int cy, int cz) {
    return ((unsigned long long)(unsigned)(cx + kCoordBias) << 42) | ((unsigned long long)(unsigned)(cy + kCoordBias) << 21) |
           (unsigned long long)(unsigned)(cz + kCoordBias);
}
__device__ __forceinline__ unsigned hash_cell(unsigned long long k) {
    // three 32-bit multiplies by large odd constants + a final fold; the full 64-bit key is what gets
    // compared in the table, so hash quality only affects probe lengths, never results
    const unsigned x = (unsigned)(k >> 42), y = (unsigned)(k >> 21) & 0x1FFFFFu, z = (unsigned)k & 0x1FFFFFu;
    unsigned h = x * 73856093u ^ y * 19349669u ^ z * 83492791u;
    return h ^ (h >> 15);
}

static __global__ void grid_bounds_kernel(const float4* __restrict__ pts, unsigned n, GridParams* gp) {
    float m = 0.0f;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pts[i];
        m = fmaxf(m, fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z))));
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0 && isfinite(m)) atomicMax(&gp->max_abs_bits, __float_as_uint(m));
}

static __global__ void slots_clear_kernel(CellSlot* slots, unsigned capacity) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += gridDim.x * blockDim.x) {
        slots[i].key = kEmptyKey; slots[i].start = 0u; slots[i].count = 0u;
    }
}

static __global__ void grid_init_kernel(CellSlot* slots, unsigned capacity, GridParams* gp, float thr, unsigned n_points) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += gridDim.x * blockDim.x) {
        slots[i].key = kEmptyKey; slots[i].start = 0u; slots[i].count = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float max_abs = __uint_as_float(gp->max_abs_bits);
        float cell = thr * 1.02f;
        float inv = (cell > 0.0f) ? 1.0f / cell : INFINITY;
        float inv_cap = (max_abs > 0.0f) ? kCoordLimit / max_abs : kCoordLimit;
        gp->inv_cell = fminf(inv, inv_cap);
        gp->cell = 1.0f / gp->inv_cell;
        gp->slack = max_abs * 4.76837158e-7f;            // 2^-21 * max|coordinate|: > 4 ulp of any coordinate
        gp->mask = capacity - 1u;
        gp->n_points = n_points;
    }
}

// claim a slot + a rank inside the cell; one atomicAdd per distinct cell per warp
static __global__ void grid_insert_kernel(const float4* __restrict__ pts, unsigned n, CellSlot* slots, const GridParams* __restrict__ gp,
                                   unsigned mask, const float* __restrict__ T_or_null,
                                   unsigned* __restrict__ pt_slot, unsigned* __restrict__ pt_rank, unsigned* occupied) {
    if (gp->inv_cell == 0.0f) return;                      // disabled (fine grid that would not pay off)
    const float inv = gp->inv_cell;
    float Tm[12];
    if (T_or_null) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) Tm[r * 4 + cc] = T_or_null[cc * 4 + r];
    }
    const unsigned lane = threadIdx.x & 31;
    const unsigned total = (n + 31u) & ~31u;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const bool live = i < n;
        unsigned slot = 0xFFFFFFFFu;
        if (live) {
            float4 p = pts[i];
            if (T_or_null) {
                float x = Tm[0] * p.x + Tm[1] * p.y + Tm[2] * p.z + Tm[3];
                float y = Tm[4] * p.x + Tm[5] * p.y + Tm[6] * p.z + Tm[7];
                float z = Tm[8] * p.x + Tm[9] * p.y + Tm[10] * p.z + Tm[11];
                p.x = x; p.y = y; p.z = z;
            }
            unsigned long long key = pack_cell(cell_coord(p.x, inv), cell_coord(p.y, inv), cell_coord(p.z, inv));
            slot = hash_cell(key) & mask;
            while (true) {
                unsigned long long prev = atomicCAS(&slots[slot].key, kEmptyKey, key);
                if (prev == kEmptyKey && occupied) atomicAdd(occupied, 1u);
                if (prev == kEmptyKey || prev == key) break;
                slot = (slot + 1u) & mask;
            }
        }
        unsigned peers = __match_any_sync(0xffffffffu, slot);
        unsigned leader = __ffs(peers) - 1u;
        unsigned base = 0;
        if (live && lane == leader) base = atomicAdd(&slots[slot].count, (unsigned)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (live) { pt_slot[i] = slot; pt_rank[i] = base + __popc(peers & ((1u << lane) - 1u)); }
    }
}

struct SlotCount { const CellSlot* s; __device__ unsigned operator()(unsigned i) const { return s[i].count; } };
struct SlotStart { CellSlot* s; __device__ void operator()(unsigned i, unsigned prefix, unsigned) const { s[i].start = prefix; } };

static __global__ void grid_scatter_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm, unsigned n,
                                    const CellSlot* __restrict__ slots, const unsigned* __restrict__ pt_slot,
                                    const unsigned* __restrict__ pt_rank, float4* __restrict__ out_pts, float4* __restrict__ out_nrm,
                                    const GridParams* __restrict__ gp_or_null) {
    if (gp_or_null && gp_or_null->inv_cell == 0.0f) return;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned pos = slots[pt_slot[i]].start + pt_rank[i];
        float4 p = pts[i]; p.w = __uint_as_float(i);
        out_pts[pos] = p;
        if (nrm) out_nrm[pos] = nrm[i];
    }
}

struct GridView {
    const CellSlot* __restrict__ slots;
    const float4* __restrict__ pts;
    float inv, cell, slack;
    unsigned mask;
};
__device__ __forceinline__ GridView make_view(const CellSlot* slots, const float4* pts, const GridParams* __restrict__ gp) {
    GridView g; g.slots = slots; g.pts = pts; g.inv = gp->inv_cell; g.cell = gp->cell; g.slack = gp->slack; g.mask = gp->mask;
    return g;
}

}  // namespace b3d
