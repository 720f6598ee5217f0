// b3d_match_tc.cu — FPFH descriptor matching on the 5th-gen tensor cores (tcgen05 + TMEM),
// with an exact fp32 re-scoring pass that keeps the selected indices bit-identical to
// src/registration.cpp:216-232 (strict '<', lowest index on ties, sequential-d fp32 sums).
//
// Idea.  d2(i,j) = |a_i|^2 + |b_j|^2 - 2 a_i.b_j.  The dot products are one GEMM with K = 33;
// tensor cores cannot reproduce the reference's fp32 rounding, so they are used only as a
// SCREEN: a pair is re-scored exactly when its screened distance could still beat (or tie) the
// row's best exact distance so far.  With |screen - exact| <= band for every pair, the true
// argmin (and every pair tied with it) always passes the screen, so the packed 64-bit
// atomicMin key (exact_d2_bits << 32 | j) ends at exactly the reference's answer.
//
//   operands : bf16 hi/lo split, folded into ONE K' = 112 GEMM
//                A' = [a_hi | a_hi | a_lo | 1 1 1 | 0...]      (128 source rows per tile)
//                B' = [b_hi | b_lo | b_hi | c0 c1 c2 | 0...]   (256 target rows per tile),
//              c0+c1+c2 = -|b_j|^2/2, so  acc = a.b - |b|^2/2  and the screen is one compare:
//                acc >= u_i,  u_i = (|a_i|^2 - best_i - band_i) / 2
//   layout   : operands are pre-packed (match_prep_kernel) into the canonical no-swizzle K-major
//              UMMA core-matrix image of each tile, so a tile is ONE contiguous
//              cp.async.bulk (TMA bulk copy, mbarrier complete_tx) and needs no tensor map
//   pipeline : warp 0 = bulk-copy producer (3-stage B ring + A tile), warp 1 = single-thread
//              tcgen05.mma issuer (M128 x N256 x K16, 7 k-steps, fp32 accumulators in TMEM,
//              double-buffered: 2 x 256 columns), warps 2..17 = epilogue (two tcgen05.ld 32x32b.x32 in
//              flight, one TMEM lane = one source row per thread, 3-input max tree + one compare)
//   schedule : persistent, one CTA per SM, contiguous slice of the (row block, column tile) space
//   error    : band_i = 2^-13 (|a_i| + max_j |b_j|)^2 covers the dropped a_lo.b_lo / residual
//              terms (<= 3*2^-18 |a||b|), fp32 accumulation inside the tensor core over 112
//              terms, and the reference's own sequential rounding (<= 35*2^-24 d2); derivation in
//              DESIGN.md.  Non-finite or huge descriptors switch the screen off (everything is
//              re-scored exactly) instead of trusting it.
#include "b3d_common.cuh"
#include <cuda_bf16.h>
#include <float.h>

namespace b3d {

constexpr int kTcM = 128;                  // source rows per tile (UMMA M)
constexpr int kTcN = 256;                  // target rows per tile (UMMA N)
constexpr int kTcK = 112;                  // packed K' (7 x UMMA_K 16)
constexpr int kTcChunks = kTcK / 8;        // 16-byte K chunks per row
constexpr int kTcStages = 3;
constexpr int kTcEpiWarps = 16;            // 4 TMEM lane quarters x 4 column parts of 64
constexpr int kTcThreads = 64 + 32 * kTcEpiWarps;   // producer + mma + epilogue warps
constexpr uint32_t kATileBytes = kTcM * kTcK * 2;     // 28 672
constexpr uint32_t kBTileBytes = kTcN * kTcK * 2;     // 57 344
constexpr uint32_t kTcSmemBytes = kATileBytes + kTcStages * kBTileBytes + 256 + 1024;
constexpr float kBandKappa = 1.220703125e-4f;          // 2^-13
constexpr int kSeeds = 8;

struct MatchAux {                 // device-resident scalars of one matching call
    unsigned max_bnorm_bits;      // max_j |b_j - mu| as float bits
    unsigned bad;                 // 1 if any descriptor value is non-finite or huge
    float mu[kDescDim];           // screen origin (see match_center_kernel)
};

// ---- exact reference arithmetic (registration.cpp:221-225) ----------------------------------
__device__ __forceinline__ float exact_dist(const float* __restrict__ a, const float* __restrict__ b) {
    float dist = 0.0f;
#pragma unroll
    for (int d = 0; d < kDescDim; ++d) { float diff = a[d] - b[d]; dist = dist + diff * diff; }
    return dist;
}
__device__ __forceinline__ unsigned long long pack_key(float d, unsigned j) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)j;
}

// ---- operand packing -------------------------------------------------------------------------
// Tile image (bf16 units): [(chunk * rows/8 + row/8) * 8 + row%8] * 8 + k%8   -> LBO = rows*16 B, SBO = 128 B
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Screen origin.  |a - b|^2 does not change when the same vector mu is subtracted from both sides, but the
// screen's error band scales with (|a - mu| + |b - mu|)^2.  Real FPFH rows of a smooth surface sit in a small ball
// far from the origin (their nearest-neighbour distances are ~1e-4 against norms of ~0.4), so screening raw values
// lets thousands of targets per row through to the exact re-score; around the cloud's mean the same band is 20-60x
// tighter.  mu is the mean of an evenly strided sample of the target rows (any vector is valid; the fl(v - mu)
// rounding adds <= 2^-23 (|a'| + |b'|)^2 to the screen error, inside the band's margin).  The exact re-score always
// uses the raw descriptors.  Fixed summation order, so mu — and with it the amount of re-scoring — is reproducible.
constexpr int kCenterGroups = 31, kCenterSample = 4096;
__global__ void __launch_bounds__(kDescDim * kCenterGroups)
match_center_kernel(const float* __restrict__ tdesc, unsigned n, MatchAux* __restrict__ aux) {
    __shared__ double part[kCenterGroups][kDescDim];
    const unsigned d = threadIdx.x % kDescDim, g = threadIdx.x / kDescDim;
    const unsigned stride = n > (unsigned)kCenterSample ? n / kCenterSample : 1u;
    const unsigned rows = n > (unsigned)kCenterSample ? (unsigned)kCenterSample : n;
    double s = 0.0;
    for (unsigned k = g; k < rows; k += kCenterGroups) {
        const float v = tdesc[(size_t)k * stride * kDescDim + d];
        if (fabsf(v) < 1e18f) s += (double)v;
    }
    part[g][d] = s;
    __syncthreads();
    if (g == 0) {
        double t = 0.0;
        for (int k = 0; k < kCenterGroups; ++k) t += part[k][d];
        aux->mu[d] = rows ? (float)(t / (double)rows) : 0.0f;
    }
}

template <bool IS_B>
__global__ void match_prep_kernel(const float* __restrict__ desc, unsigned n, unsigned first, unsigned n_padded,
                                  __nv_bfloat16* __restrict__ tiles, float* __restrict__ norm2, MatchAux* __restrict__ aux,
                                  unsigned* __restrict__ tile_bmax_bits) {
    constexpr int ROWS = IS_B ? kTcN : kTcM;
    const unsigned r = first + blockIdx.x * blockDim.x + threadIdx.x;          // rows [first, n_padded): a rank packs only the row blocks it matches
    if (r >= n_padded) return;
    __align__(16) __nv_bfloat16 row[kTcK];
#pragma unroll
    for (int k = 0; k < kTcK; ++k) row[k] = __float2bfloat16_rn(0.0f);
    if (r < n) {
        double s2 = 0.0; bool bad = false;
        const float* src = desc + (size_t)r * kDescDim;
        for (int d = 0; d < kDescDim; ++d) {
            float v = src[d];
            if (!(fabsf(v) < 1e18f)) bad = true;
            v = v - aux->mu[d];
            s2 += (double)v * (double)v;
            __nv_bfloat16 hi, lo; split_bf16(v, hi, lo);
            if (IS_B) { row[d] = hi; row[33 + d] = lo; row[66 + d] = hi; }
            else      { row[d] = hi; row[33 + d] = hi; row[66 + d] = lo; }
        }
        if (IS_B) {
            float h = (float)(-0.5 * s2);
            __nv_bfloat16 c0 = __float2bfloat16_rn(h);
            float r1 = h - __bfloat162float(c0);
            __nv_bfloat16 c1 = __float2bfloat16_rn(r1);
            float r2 = r1 - __bfloat162float(c1);
            row[99] = c0; row[100] = c1; row[101] = __float2bfloat16_rn(r2);
            float nb = (float)sqrt(s2) * 1.0000002f;
            if (isfinite(nb)) { atomicMax(&aux->max_bnorm_bits, __float_as_uint(nb)); atomicMax(&tile_bmax_bits[r / ROWS], __float_as_uint(nb)); }
        } else {
            row[99] = row[100] = row[101] = __float2bfloat16_rn(1.0f);
            norm2[r] = (float)s2;
        }
        if (bad) atomicOr(&aux->bad, 1u);
    } else if (IS_B) {
        row[99] = __float2bfloat16_rn(-3.0e38f);        // padded target rows can never pass the screen
    }
    const unsigned tile = r / ROWS, rt = r % ROWS;
    uint4* dst = reinterpret_cast<uint4*>(tiles + (size_t)tile * ROWS * kTcK);
    const uint4* srcv = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int c = 0; c < kTcChunks; ++c) dst[(c * (ROWS / 8) + rt / 8) * 8 + (rt % 8)] = srcv[c];
}

// initial best per row: the reference's (FLT_MAX, 0) start, tightened by a few exact evaluations
__global__ void match_seed_kernel(const float* __restrict__ sdesc, const float* __restrict__ tdesc, unsigned row0, unsigned row1,
                                  unsigned n_tgt, unsigned long long* __restrict__ best) {
    const unsigned i = row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row1) return;
    unsigned long long key = pack_key(FLT_MAX, 0u);
    const float* a = sdesc + (size_t)i * kDescDim;
    for (int s = 0; s < kSeeds; ++s) {
        unsigned j = (unsigned)(((unsigned long long)i * 2654435761ull + (unsigned long long)s * 40503ull + 12345ull) % n_tgt);
        float d = exact_dist(a, tdesc + (size_t)j * kDescDim);
        if (d < FLT_MAX || j == 0) { unsigned long long k2 = pack_key(d, j); key = k2 < key ? k2 : key; }
    }
    best[i] = key;
}

__global__ void match_finalize_kernel(const unsigned long long* __restrict__ best, unsigned row0, unsigned row1, uint32_t* __restrict__ corr) {
    const unsigned i = row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < row1) corr[i] = (uint32_t)(best[i] & 0xFFFFFFFFull);
}

// ---- PTX helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=256, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

// ---- the screen + re-score kernel -----------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
match_tc_kernel(const __nv_bfloat16* __restrict__ a_tiles, const __nv_bfloat16* __restrict__ b_tiles,
                const float* __restrict__ sdesc, const float* __restrict__ tdesc, const float* __restrict__ a_norm2,
                const MatchAux* __restrict__ aux, unsigned row0, unsigned row1, unsigned n_tgt,
                unsigned rb_first, unsigned n_rb, unsigned n_nt, unsigned long long* __restrict__ best,
                const unsigned* __restrict__ tile_bmax_bits) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_smem = smem_u32(smem);
    const uint32_t b_smem = a_smem + kATileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kATileBytes + kTcStages * kBTileBytes);
    // bars: [0..2] full, [3..5] empty, [6..7] tmem_full, [8..9] tmem_empty, [10] a_full, [11] a_free
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(BAR(s), 1); mbar_init(BAR(3 + s), 1); }
        mbar_init(BAR(6), 1); mbar_init(BAR(7), 1);
        mbar_init(BAR(8), kTcEpiWarps); mbar_init(BAR(9), kTcEpiWarps);
        mbar_init(BAR(10), 1); mbar_init(BAR(11), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const unsigned long long total = (unsigned long long)n_rb * n_nt;
    const unsigned t_begin = (unsigned)(total * blockIdx.x / gridDim.x);
    const unsigned t_end = (unsigned)(total * (blockIdx.x + 1) / gridDim.x);

    if (warp == 0) {
        // ===== producer: one elected lane issues the bulk copies =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, rb_count = 0;
            unsigned cur_rb = 0xFFFFFFFFu;
            for (unsigned t = t_begin; t < t_end; ++t) {
                const unsigned rb = t / n_nt, nt = t - rb * n_nt;
                if (rb != cur_rb) {
                    mbar_wait(BAR(11), (rb_count & 1u) ^ 1u);                      // previous row block's MMAs retired
                    mbar_arrive_expect_tx(BAR(10), kATileBytes);
                    bulk_g2s(a_smem, a_tiles + (size_t)(rb_first + rb) * kTcM * kTcK, kATileBytes, BAR(10));
                    cur_rb = rb; ++rb_count;
                }
                mbar_wait(BAR(3 + stage), phase ^ 1u);
                mbar_arrive_expect_tx(BAR(stage), kBTileBytes);
                bulk_g2s(b_smem + stage * kBTileBytes, b_tiles + (size_t)nt * kTcN * kTcK, kBTileBytes, BAR(stage));
                if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: a single thread drives the tensor core =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, rb_count = 0;
            unsigned cur_rb = 0xFFFFFFFFu;
            for (unsigned t = t_begin; t < t_end; ++t) {
                const unsigned rb = t / n_nt;
                if (rb != cur_rb) { mbar_wait(BAR(10), rb_count & 1u); cur_rb = rb; ++rb_count; }
                mbar_wait(BAR(stage), phase);                                     // B tile landed
                mbar_wait(BAR(8 + acc), acc_phase ^ 1u);                          // accumulator drained by the epilogue
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)kTcN;
                const uint32_t b_base = b_smem + stage * kBTileBytes;
#pragma unroll
                for (int ks = 0; ks < kTcK / 16; ++ks) {
                    const uint64_t ad = umma_desc(a_smem + ks * 2 * (kTcM * 16), kTcM * 16, 128);
                    const uint64_t bd = umma_desc(b_base + ks * 2 * (kTcN * 16), kTcN * 16, 128);
                    tc_mma_bf16(d_tmem, ad, bd, kIdesc, ks > 0 ? 1u : 0u);
                }
                tc_commit(BAR(3 + stage));                                        // smem slot reusable when these MMAs retire
                tc_commit(BAR(6 + acc));                                          // accumulator ready for the epilogue
                const bool last_of_rb = (t + 1 == t_end) || ((t + 1) / n_nt != rb);
                if (last_of_rb) tc_commit(BAR(11));
                if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
                acc ^= 1u; if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ===== epilogue: thread = one source row (TMEM lane), 64 of the tile's 256 columns =====
        // The epilogue has to read every fp32 accumulator once: 128 KB of TMEM per tile at ~64 B/clk/SM is
        // ~2000 cycles against 896 cycles of MMA, so this kernel is TMEM-read bound (measured: sharing each
        // B tile across 4 row blocks to cut L2 traffic 4x did not help).
        const int ew = warp - 2;
        const int quarter = warp & 3;                      // TMEM lanes a warp may touch: 32 * (warpid % 4)
        const int part = ew >> 2;                          // which 64-column slice
        const unsigned bad = aux->bad;
        const float bmax = __uint_as_float(aux->max_bnorm_bits);
        uint32_t acc = 0, acc_phase = 0;
        unsigned cur_rb = 0xFFFFFFFFu, since_refresh = 0;
        unsigned i = 0; bool row_live = false; float na = 0.0f, ra = 0.0f, band = 0.0f, best_d = FLT_MAX, u = INFINITY;
        auto threshold = [&]() {
            float t = bad ? -INFINITY : 0.5f * ((na - best_d) - band) - 1e-30f;
            return (t == t) ? t : -INFINITY;               // NaN norm => re-score everything
        };
        auto rescore = [&](unsigned mask, unsigned colbase) {
            while (mask) {
                const int k = __ffs(mask) - 1;
                mask &= mask - 1u;
                const unsigned j = colbase + k;
                if (j < n_tgt) {
                    const float d = exact_dist(sdesc + (size_t)i * kDescDim, tdesc + (size_t)j * kDescDim);
                    const unsigned long long key = pack_key(d, j);
                    const unsigned long long old = atomicMin(&best[i], key);
                    const float nd = __uint_as_float((unsigned)((old < key ? old : key) >> 32));
                    if (nd < best_d) { best_d = nd; u = threshold(); }
                }
            }
        };
        for (unsigned t = t_begin; t < t_end; ++t) {
            const unsigned rb = t / n_nt, nt = t - rb * n_nt;
            if (rb != cur_rb) {
                cur_rb = rb; since_refresh = 1000;
                i = (rb_first + rb) * kTcM + quarter * 32 + lane;
                row_live = (i >= row0 && i < row1);
                na = 0.0f; band = 0.0f;
                ra = 0.0f;
                if (row_live) {
                    na = a_norm2[i];
                    ra = sqrtf(na) * 1.0000002f;
                    float s = ra + bmax;
                    band = kBandKappa * s * s;
                }
            }
            {   // the band only has to cover THIS tile's targets: (|a'| + max over the tile of |b'|)^2 instead of the global maximum
                const float tb = fminf(__uint_as_float(tile_bmax_bits[nt]), bmax);
                const float s = ra + tb;
                const float tile_band = kBandKappa * s * s;
                if (tile_band != band) { band = tile_band; u = row_live ? threshold() : INFINITY; }
            }
            if (++since_refresh >= 4) {                    // other warps tighten the row's best too; a stale value is only looser
                since_refresh = 0;
                best_d = row_live ? __uint_as_float((unsigned)(*(volatile unsigned long long*)&best[i] >> 32)) : 0.0f;
                u = row_live ? threshold() : INFINITY;
            }
            mbar_wait(BAR(6 + acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (uint32_t)kTcN + part * 64;
            const unsigned col0 = nt * kTcN + part * 64;
            uint32_t r0[32], r1[32];
            __syncwarp();
            tc_ld32_issue(taddr0, r0);                     // both 32-column chunks in flight, one wait
            tc_ld32_issue(taddr0 + 32, r1);
            tc_ld_wait();
            float m0 = __uint_as_float(r0[0]), m1 = __uint_as_float(r1[0]);
#pragma unroll
            for (int k = 1; k + 1 < 32; k += 2) {
                m0 = fmaxf(m0, fmaxf(__uint_as_float(r0[k]), __uint_as_float(r0[k + 1])));
                m1 = fmaxf(m1, fmaxf(__uint_as_float(r1[k]), __uint_as_float(r1[k + 1])));
            }
            m0 = fmaxf(m0, __uint_as_float(r0[31])); m1 = fmaxf(m1, __uint_as_float(r1[31]));
            if (row_live && (bad || fmaxf(m0, m1) >= u)) {                         // rare: exact re-score
                unsigned mask0 = 0u, mask1 = 0u;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    mask0 |= (__uint_as_float(r0[k]) >= u ? 1u : 0u) << k;
                    mask1 |= (__uint_as_float(r1[k]) >= u ? 1u : 0u) << k;
                }
                if (bad) { mask0 = 0xFFFFFFFFu; mask1 = 0xFFFFFFFFu; }
                rescore(mask0, col0);
                rescore(mask1, col0 + 32);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(8 + acc));
            acc ^= 1u; if (acc == 0) acc_phase ^= 1u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
int match_features_tc_impl(b3d_ctx* c, size_t row0, size_t row1) {
    const unsigned n_src = (unsigned)c->n_src, n_tgt = (unsigned)c->n_tgt;
    const unsigned n_rb_all = (unsigned)div_up(n_src, kTcM), n_nt = (unsigned)div_up(n_tgt, kTcN);
    B3D_CUDA(c, c->tc_a_tiles.ensure((size_t)n_rb_all * kATileBytes));
    B3D_CUDA(c, c->tc_b_tiles.ensure((size_t)n_nt * kBTileBytes));
    B3D_CUDA(c, c->tc_norm2.ensure(sizeof(float) * (size_t)n_rb_all * kTcM));
    B3D_CUDA(c, c->tc_best.ensure(sizeof(unsigned long long) * (size_t)n_src));
    B3D_CUDA(c, c->tc_aux.ensure(sizeof(MatchAux)));
    B3D_CUDA(c, c->tc_aux.ensure(sizeof(MatchAux) + sizeof(unsigned) * n_nt));
    MatchAux* aux = c->tc_aux.as<MatchAux>();
    unsigned* tile_bmax = reinterpret_cast<unsigned*>(aux + 1);
    B3D_CUDA(c, cudaMemsetAsync(aux, 0, sizeof(MatchAux) + sizeof(unsigned) * n_nt, c->stream));
    match_center_kernel<<<1, kDescDim * kCenterGroups, 0, c->stream>>>(c->tdesc_p, n_tgt, aux);
    B3D_LAUNCHED(c);
    const unsigned rb_first = (unsigned)(row0 / kTcM);
    const unsigned n_rb = (unsigned)div_up((long long)row1, kTcM) - rb_first;
    match_prep_kernel<false><<<div_up(n_rb * kTcM, 128), 128, 0, c->stream>>>(c->sdesc_p, n_src, rb_first * kTcM, (rb_first + n_rb) * kTcM,
                                                                              c->tc_a_tiles.as<__nv_bfloat16>(), c->tc_norm2.as<float>(), aux, nullptr);
    B3D_LAUNCHED(c);
    match_prep_kernel<true><<<div_up(n_nt * kTcN, 128), 128, 0, c->stream>>>(c->tdesc_p, n_tgt, 0u, n_nt * kTcN, c->tc_b_tiles.as<__nv_bfloat16>(),
                                                                             nullptr, aux, tile_bmax);
    B3D_LAUNCHED(c);
    match_seed_kernel<<<div_up((long long)(row1 - row0), 128), 128, 0, c->stream>>>(c->sdesc_p, c->tdesc_p, (unsigned)row0, (unsigned)row1, n_tgt,
                                                                                    c->tc_best.as<unsigned long long>());
    B3D_LAUNCHED(c);
    if (!c->tc_smem_opt_in) {                                  // per device (a pool may drive several GPUs from one process)
        B3D_CUDA(c, cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
        c->tc_smem_opt_in = true;
    }
    unsigned long long total = (unsigned long long)n_rb * n_nt;
    int grid = (int)(total < (unsigned long long)kNumSMs ? total : (unsigned long long)kNumSMs);
    match_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, c->stream>>>(c->tc_a_tiles.as<__nv_bfloat16>(), c->tc_b_tiles.as<__nv_bfloat16>(),
                                                                   c->sdesc_p, c->tdesc_p, c->tc_norm2.as<float>(), aux,
                                                                   (unsigned)row0, (unsigned)row1, n_tgt, rb_first, n_rb, n_nt,
                                                                   c->tc_best.as<unsigned long long>(), tile_bmax);
    B3D_LAUNCHED(c);
    match_finalize_kernel<<<div_up((long long)(row1 - row0), 256), 256, 0, c->stream>>>(c->tc_best.as<unsigned long long>(), (unsigned)row0, (unsigned)row1,
                                                                                       c->corr.as<uint32_t>());
    B3D_LAUNCHED(c);
    return B3D_OK;
}

}  // namespace b3d
