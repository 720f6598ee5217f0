// b3d_match.cu — FPFH descriptor correspondence matching.
// Replaces the O(Ns*Nt*33) loop of src/registration.cpp:216-232 (reference repo):
//   for each source i: argmin_j sum_{d=0..32} (a_d - b_d)^2, summed sequentially in d,
//   strict '<' so ties go to the lowest j; best_idx starts at 0, best_dist at FLT_MAX.
// Compile with --fmad=false: every distance is bit-identical to the reference's.
//
// match_exact_kernel: CUDA-core register-tiled brute force.  64 source rows x 64
// target columns per shared-memory tile, 4x4 micro-tile per thread, descriptors
// staged d-major so one LDS.128 feeds four rows/columns.
#include "b3d_common.cuh"
#include <float.h>

namespace b3d {

constexpr int kMT = 64;            // source rows per block
constexpr int kNT = 64;            // target columns per smem tile
constexpr int kPad = 4;            // keeps 16 B alignment of every d-row
constexpr int kMatchThreads = 256; // 16 x 16 threads, 4 x 4 outputs each

__global__ void __launch_bounds__(kMatchThreads)
match_exact_kernel(const float* __restrict__ sdesc, const float* __restrict__ tdesc,
                   unsigned row0, unsigned row1, unsigned n_tgt, uint32_t* __restrict__ corr) {
    __shared__ __align__(16) float As[kDescDim][kMT + kPad];
    __shared__ __align__(16) float Bs[kDescDim][kNT + kPad];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const unsigned block_row = row0 + blockIdx.x * kMT;

    // A tile: transposing load, once per block
    for (int e = tid; e < kMT * kDescDim; e += kMatchThreads) {
        int r = e & (kMT - 1), d = e >> 6;
        unsigned row = block_row + r; if (row >= row1) row = row1 - 1;
        As[d][r] = sdesc[(size_t)row * kDescDim + d];
    }

    float best[4]; unsigned bidx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { best[i] = FLT_MAX; bidx[i] = 0u; }

    const unsigned n_tiles = (n_tgt + kNT - 1) / kNT;
    for (unsigned tile = 0; tile < n_tiles; ++tile) {
        __syncthreads();                                   // previous tile fully consumed (and As visible)
        const unsigned col0 = tile * kNT;
        const unsigned valid = min((unsigned)kNT, n_tgt - col0);
        const float* src = tdesc + (size_t)col0 * kDescDim;
        for (int e = tid; e < kNT * kDescDim; e += kMatchThreads) {
            int c = e / kDescDim, d = e - c * kDescDim;    // coalesced read of the [64][33] block
            Bs[d][c] = (c < (int)valid) ? src[e] : 0.0f;
        }
        __syncthreads();

        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 3
        for (int d = 0; d < kDescDim; ++d) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[d][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[d][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float diff = a[i] - b[j];
                    acc[i][j] = acc[i][j] + diff * diff;
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                unsigned col = col0 + tx * 4 + j;
                if (col < n_tgt && acc[i][j] < best[i]) { best[i] = acc[i][j]; bidx[i] = col; }
            }
    }

    // lexicographic (distance, index) minimum over the 16 threads that share a row
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float bd = best[i]; unsigned bi = bidx[i];
#pragma unroll
        for (int m = 8; m >= 1; m >>= 1) {
            float od = __shfl_xor_sync(0xffffffffu, bd, m);
            unsigned oi = __shfl_xor_sync(0xffffffffu, bi, m);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        unsigned row = block_row + ty * 4 + i;
        if (tx == 0 && row < row1) corr[row] = bi;
    }
}

int match_features_impl(b3d_ctx* c, size_t row0, size_t row1) {
    if (!c->have_clouds || !c->have_feats) return fail(c, B3D_ERR_STATE, "match_features: clouds/features not set");
    if (row1 > c->n_src || row0 > row1) return fail(c, B3D_ERR_INVALID, "match_features: bad row range");
    B3D_CUDA(c, c->corr.ensure(sizeof(uint32_t) * (c->n_src ? c->n_src : 1)));
    if (row0 == row1) return B3D_OK;
    StageTimer timer(c, 0);
    if (c->n_tgt == 0) {   // no target rows: the reference leaves best_idx at 0
        B3D_CUDA(c, cudaMemsetAsync(c->corr.as<uint32_t>() + row0, 0, sizeof(uint32_t) * (row1 - row0), c->stream));
        return B3D_OK;
    }
    // tensor-core screen pays off once there are enough pairs to amortise operand packing
    const bool use_tc = c->match_mode == 2 || (c->match_mode == 0 && (double)(row1 - row0) * (double)c->n_tgt >= 2.5e7);
    if (use_tc) return match_features_tc_impl(c, row0, row1);
    int blocks = div_up((long long)(row1 - row0), kMT);
    match_exact_kernel<<<blocks, kMatchThreads, 0, c->stream>>>(c->sdesc_p, c->tdesc_p, (unsigned)row0, (unsigned)row1,
                                                                (unsigned)c->n_tgt, c->corr.as<uint32_t>());
    B3D_LAUNCHED(c);
    return B3D_OK;
}

}  // namespace b3d
