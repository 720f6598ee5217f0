// b3d_ransac.cu — RANSAC hypothesis generation, inlier scoring and best selection.
// Replaces src/registration.cpp:234-295 of the reference.  Compile with --fmad=false.
//
//  * RNG: the reference draws from std::mt19937(42) through
//    std::uniform_int_distribution<size_t>(0, Ns-1) (registration.cpp:235-239).  The raw
//    32-bit stream depends only on the literal seed, so it is generated once per
//    context and kept resident; `raw[j]` is then a counter-based generator.  libstdc++
//    maps raw -> index with Lemire's multiply-shift + rejection; rejections shift all
//    later draws, so accepted draws are compacted with a device-wide prefix sum.
//  * Hypotheses: one thread per hypothesis runs the 3-point Kabsch + Jacobi SVD of
//    b3d_linalg.cuh and stores (R,t) in SoA form; degenerate triples are marked.
//  * Scoring (dominant cost, FP32-issue bound): one hypothesis per thread with the
//    (source, matched target) pair stream broadcast out of shared memory, so every
//    LDS feeds 27 x KH arithmetic instructions and no cross-lane reduction is needed.
//    The grid is (hypothesis tiles) x (pair ranges); partial counts meet in integer
//    atomics, which are exact.  sqrt is monotone, so  sqrt(d2) < thr  is evaluated as
//    d2 < cut  with cut = the smallest float whose correctly rounded sqrt is >= thr.
//  * Selection: strict '>' on fitness with earliest id winning, early exit on
//    fitness > confidence (registration.cpp:284-290) as packed 64-bit max-reductions.
#include "b3d_common.cuh"
#include "b3d_linalg.cuh"
#include "b3d_scan.cuh"
#include "b3d_ess.cuh"
#include <float.h>
#include <math.h>

namespace b3d {

// ---------------------------------------------------------------------------------
// RNG mapping
// ---------------------------------------------------------------------------------
struct LemireAccept {
    const uint32_t* raw; uint32_t range, reject_below;
    __device__ unsigned operator()(unsigned j) const {
        uint32_t lo = (uint32_t)((uint64_t)raw[j] * range);
        return lo >= reject_below ? 1u : 0u;
    }
};
struct DrawEmit {
    const uint32_t* raw; uint32_t range; uint32_t* draws; unsigned need;
    __device__ void operator()(unsigned j, unsigned prefix, unsigned accepted) const {
        if (accepted && prefix < need) draws[prefix] = (uint32_t)(((uint64_t)raw[j] * range) >> 32);
    }
};

// ---------------------------------------------------------------------------------
// pairs[i] = s_i ; pairs[stride + i] = q_{corr[i]}.  stride = n rounded up to the pair tile; the
// padding holds sentinel pairs that can never be inliers.  Also records max |coordinate| of
// both sides (float bits, atomicMax) for the screening error bound.
// ---------------------------------------------------------------------------------
__global__ void gather_pairs_kernel(const float4* __restrict__ src4, const float4* __restrict__ tgt4,
                                    const uint32_t* __restrict__ corr, unsigned n, unsigned n_tgt, unsigned stride,
                                    float4* __restrict__ pairs, DeviceState* __restrict__ st) {
    float smax = 0.0f, qmax = 0.0f;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < stride; i += gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(1e18f, 1e18f, 1e18f, 0.f);
        if (i < n) {
            s = src4[i];
            unsigned j = corr[i];
            q = (j < n_tgt) ? tgt4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            smax = fmaxf(smax, fmaxf(fabsf(s.x), fmaxf(fabsf(s.y), fabsf(s.z))));
            qmax = fmaxf(qmax, fmaxf(fabsf(q.x), fmaxf(fabsf(q.y), fabsf(q.z))));
        }
        pairs[i] = s;
        pairs[stride + i] = q;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, m));
        qmax = fmaxf(qmax, __shfl_xor_sync(0xffffffffu, qmax, m));
    }
    if ((threadIdx.x & 31) == 0) {      // NaN-free by construction of fmaxf; inf stays inf and disables the screen
        atomicMax(&st->pair_smax_bits, __float_as_uint(smax));
        atomicMax(&st->pair_qmax_bits, __float_as_uint(qmax));
    }
}

// ---------------------------------------------------------------------------------
// hypothesis generation: hyp is SoA float[12][H]
// ---------------------------------------------------------------------------------
__global__ void hypothesis_kernel(const uint32_t* __restrict__ draws, DeviceState* __restrict__ st,
                                  const float4* __restrict__ pairs, unsigned pair_stride, int H, int h_lo, int h_hi,
                                  float* __restrict__ hyp, int* __restrict__ counts) {
    int h = h_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= h_hi) return;
    if ((unsigned)(3 * (unsigned)h + 2u) >= st->accepted_total) { counts[h] = -3; st->rng_starved = 1; return; }   // raw window too small: surfaced by finish
    uint32_t i0 = draws[3 * (size_t)h], i1 = draws[3 * (size_t)h + 1], i2 = draws[3 * (size_t)h + 2];
    if (i0 == i1 || i1 == i2 || i0 == i2) {                // registration.cpp:240 `continue`
        counts[h] = -1;
#pragma unroll
        for (int k = 0; k < 12; ++k) hyp[(size_t)k * H + h] = 0.0f;
        return;
    }
    float s[3][3], q[3][3];
    const uint32_t id[3] = {i0, i1, i2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float4 a = pairs[id[k]], b = pairs[pair_stride + id[k]];
        s[k][0] = a.x; s[k][1] = a.y; s[k][2] = a.z;
        q[k][0] = b.x; q[k][1] = b.y; q[k][2] = b.z;
    }
    Mat3 R; float t[3];
    kabsch_three_points(s, q, R, t);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) hyp[(size_t)(r * 3 + cc) * H + h] = R(r, cc);
#pragma unroll
    for (int r = 0; r < 3; ++r) hyp[(size_t)(9 + r) * H + h] = t[r];
    counts[h] = 0;
}

// Early exit (registration.cpp:290): scoring runs in chunks of hypothesis ids; after each chunk this kernel looks for an id
// whose fitness exceeds the confidence.  Once one exists the remaining chunks' kernels return at once (score_exit), and
// mark_unscored_kernel gives every id behind the first such id the count -2 ("never ran"), as the reference's loop leaves it.
__global__ void exit_check_kernel(const int* __restrict__ counts, int a, int b, float n_src_f, float confidence, DeviceState* __restrict__ st) {
    const int h = a + blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= b) return;
    const int c = counts[h];
    if (c > 0 && (float)c / n_src_f > confidence) { atomicMin(&st->score_exit_id, h); st->score_exit = 1; }
}
__global__ void mark_unscored_kernel(int* __restrict__ counts, int h0, int h1, const DeviceState* __restrict__ st) {
    const int h = h0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (h < h1 && st->score_exit && h > st->score_exit_id) counts[h] = -2;
}
__global__ void exit_reset_kernel(DeviceState* st) { st->score_exit = 0; st->score_exit_id = 0x7FFFFFFF; }

__global__ void reset_counts_kernel(int* counts, int h0, int h1) {
    int h = h0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (h < h1 && (counts[h] > 0 || counts[h] == -4)) counts[h] = 0;      // -4: dropped by an earlier bail-out run; -1 (degenerate) stays
}

// ---------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------
constexpr int kScoreThreads = 256;
constexpr int kPairTile = 512;        // pairs per smem tile: 2 x 8 KB

template <int KH>
__global__ void __launch_bounds__(kScoreThreads)
score_exact_kernel(const float* __restrict__ hyp, int H, int h0, int h1,
                   const float4* __restrict__ pairs, unsigned n_pairs, unsigned pair_stride, unsigned pairs_per_y,
                   float cut, int* __restrict__ counts, const DeviceState* __restrict__ st) {
    if (st->score_exit) return;                        // an earlier chunk met fitness > confidence: the reference broke there
    __shared__ float4 sS[kPairTile];
    __shared__ float4 sQ[kPairTile];
    const int tid = threadIdx.x;
    float R[KH][9], t[KH][3];
    int cnt[KH], hid[KH];
#pragma unroll
    for (int k = 0; k < KH; ++k) {
        int h = h0 + (blockIdx.x * KH + k) * kScoreThreads + tid;
        hid[k] = h;
        int hc = h < h1 ? h : h1 - 1;
#pragma unroll
        for (int e = 0; e < 9; ++e) R[k][e] = hyp[(size_t)e * H + hc];
#pragma unroll
        for (int e = 0; e < 3; ++e) t[k][e] = hyp[(size_t)(9 + e) * H + hc];
        cnt[k] = 0;
    }
    const unsigned p_begin = blockIdx.y * pairs_per_y;
    const unsigned p_end = min(n_pairs, p_begin + pairs_per_y);
    for (unsigned base = p_begin; base < p_end; base += kPairTile) {
        const unsigned m = min((unsigned)kPairTile, p_end - base);
        __syncthreads();
        for (unsigned e = tid; e < m; e += kScoreThreads) { sS[e] = pairs[base + e]; sQ[e] = pairs[pair_stride + base + e]; }
        __syncthreads();
#pragma unroll 4
        for (unsigned j = 0; j < m; ++j) {
            const float4 s = sS[j], q = sQ[j];
#pragma unroll
            for (int k = 0; k < KH; ++k) {
                // (R s + t - q) in the reference's evaluation order, no contraction
                float x = (R[k][0] * s.x + (R[k][1] * s.y + R[k][2] * s.z)) + t[k][0];
                float y = (R[k][3] * s.x + (R[k][4] * s.y + R[k][5] * s.z)) + t[k][1];
                float z = (R[k][6] * s.x + (R[k][7] * s.y + R[k][8] * s.z)) + t[k][2];
                float dx = x - q.x, dy = y - q.y, dz = z - q.z;
                float d2 = dx * dx + (dy * dy + dz * dz);
                cnt[k] += (d2 < cut) ? 1 : 0;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KH; ++k) {
        int h = hid[k];
        if (h < h1 && counts[h] >= 0) {
            if (gridDim.y == 1) counts[h] = cnt[k];
            else if (cnt[k]) atomicAdd(&counts[h], cnt[k]);
        }
    }
}

// Screened scoring.  The reference's un-fused arithmetic costs 27 instructions per
// (hypothesis, pair); the same squared distance evaluated with FMAs costs 15.  The two values
// differ by at most `beta` (derived in DESIGN.md from the 11 roundings per coordinate), so
//   d2_fma <  cut - beta  => the reference counts the pair,
//   d2_fma >  cut + beta  => it does not,
// and only pairs inside the band need the reference arithmetic.  Each thread keeps two counters
// per hypothesis over a group of 32 pairs (below lo / below hi); when they agree the group had no
// pair in the band and the count is exact by construction, otherwise that group alone is
// re-counted with the un-fused arithmetic, cooperatively by the whole warp.  The inner loop stays
// branch-free.
constexpr int kGroup = 32;

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
        v = o > v ? o : v;
    }
    return v;
}

// One compare + one predicated add each (the compiler otherwise emits add-then-conditional-move
// pairs; ALU-pipe instructions cost two issue cycles on this part, so they are worth trimming).
__device__ __forceinline__ void count_below(int& cnt, float v, float bound) {       // cnt += (v < bound)
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt) : "f"(v), "f"(bound));
}
__device__ __forceinline__ void count_not_above(int& cnt, float v, float bound) {   // cnt += !(v > bound); NaN counts
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@!p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt) : "f"(v), "f"(bound));
}

// Warp-cooperative exact re-count of one 32-pair group for the hypothesis held by `src_lane`:
// (R,t) is broadcast with shuffles, lane l evaluates pair l with the reference arithmetic, and
// the count is a ballot.  ~45 warp instructions instead of 32 x 27 serial ones in one lane.
__device__ __forceinline__ int warp_exact_group_count(const float (&R)[9], const float (&t)[3], int src_lane,
                                                      const float4* sS, const float4* sQ, float cut) {
    float Rb[9], tb[3];
#pragma unroll
    for (int e = 0; e < 9; ++e) Rb[e] = __shfl_sync(0xffffffffu, R[e], src_lane);
#pragma unroll
    for (int e = 0; e < 3; ++e) tb[e] = __shfl_sync(0xffffffffu, t[e], src_lane);
    const int lane = threadIdx.x & 31;
    const float4 s = sS[lane], q = sQ[lane];
    float x = (Rb[0] * s.x + (Rb[1] * s.y + Rb[2] * s.z)) + tb[0];
    float y = (Rb[3] * s.x + (Rb[4] * s.y + Rb[5] * s.z)) + tb[1];
    float z = (Rb[6] * s.x + (Rb[7] * s.y + Rb[8] * s.z)) + tb[2];
    float dx = x - q.x, dy = y - q.y, dz = z - q.z;
    float d2 = dx * dx + (dy * dy + dz * dz);
    return __popc(__ballot_sync(0xffffffffu, d2 < cut));
}

template <int KH>
__global__ void __launch_bounds__(kScoreThreads)
score_screen_kernel(const float* __restrict__ hyp, int H, int h0, int h1,
                    const float4* __restrict__ pairs, unsigned pair_stride, unsigned pairs_per_y,
                    float cut, float thr, const DeviceState* __restrict__ st, int* __restrict__ counts) {
    if (st->score_exit) return;                        // an earlier chunk met fitness > confidence: the reference broke there
    __shared__ float4 sS[kPairTile];
    __shared__ float4 sQ[kPairTile];
    const int tid = threadIdx.x;
    float R[KH][9], t[KH][3], lo[KH], hi[KH];
    int total[KH], hid[KH];
    unsigned recounts = 0;
    const float smax = __uint_as_float(st->pair_smax_bits), qmax = __uint_as_float(st->pair_qmax_bits);
#pragma unroll
    for (int k = 0; k < KH; ++k) {
        int h = h0 + (blockIdx.x * KH + k) * kScoreThreads + tid;
        hid[k] = h;
        int hc = h < h1 ? h : h1 - 1;
#pragma unroll
        for (int e = 0; e < 9; ++e) R[k][e] = hyp[(size_t)e * H + hc];
#pragma unroll
        for (int e = 0; e < 3; ++e) t[k][e] = hyp[(size_t)(9 + e) * H + hc];
        total[k] = 0;
        // error bound of the fused evaluation against the reference's (see header comment)
        float rowsum = fmaxf(fabsf(R[k][0]) + fabsf(R[k][1]) + fabsf(R[k][2]),
                             fmaxf(fabsf(R[k][3]) + fabsf(R[k][4]) + fabsf(R[k][5]), fabsf(R[k][6]) + fabsf(R[k][7]) + fabsf(R[k][8])));
        float A = rowsum * smax + fmaxf(fabsf(t[k][0]), fmaxf(fabsf(t[k][1]), fabsf(t[k][2]))) + qmax;
        float e = A * 7.152557373046875e-7f;                    // 12 * 2^-24 * A  (11 roundings per coordinate + margin)
        float beta = 1.25f * (4.0f * (thr * 1.02f + e) * e + 2e-6f * cut);      // 25 % on top of the derived bound
        if (beta < 0.5f * cut) { lo[k] = cut - beta; hi[k] = cut + beta; }
        else { lo[k] = -INFINITY; hi[k] = INFINITY; }           // bound useless or NaN: every group is re-counted exactly
    }
    const unsigned p_begin = blockIdx.y * pairs_per_y;
    const unsigned p_end = min(pair_stride, p_begin + pairs_per_y);     // multiples of kPairTile
    for (unsigned base = p_begin; base < p_end; base += kPairTile) {
        __syncthreads();
        for (unsigned e = tid; e < kPairTile; e += kScoreThreads) { sS[e] = pairs[base + e]; sQ[e] = pairs[pair_stride + base + e]; }
        __syncthreads();
#pragma unroll 1
        for (int g = 0; g < kPairTile; g += kGroup) {
            int clo[KH], chi[KH];
#pragma unroll
            for (int k = 0; k < KH; ++k) { clo[k] = 0; chi[k] = 0; }
#pragma unroll 8
            for (int j = 0; j < kGroup; ++j) {
                const float4 s = sS[g + j], q = sQ[g + j];
#pragma unroll
                for (int k = 0; k < KH; ++k) {
                    float dx = __fmaf_rn(R[k][0], s.x, __fmaf_rn(R[k][1], s.y, __fmaf_rn(R[k][2], s.z, t[k][0] - q.x)));
                    float dy = __fmaf_rn(R[k][3], s.x, __fmaf_rn(R[k][4], s.y, __fmaf_rn(R[k][5], s.z, t[k][1] - q.y)));
                    float dz = __fmaf_rn(R[k][6], s.x, __fmaf_rn(R[k][7], s.y, __fmaf_rn(R[k][8], s.z, t[k][2] - q.z)));
                    float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
                    count_below(clo[k], d2, lo[k]);
                    count_not_above(chi[k], d2, hi[k]);         // NaN counts as "maybe"
                }
            }
#pragma unroll
            for (int k = 0; k < KH; ++k) {
                unsigned need = __ballot_sync(0xffffffffu, clo[k] != chi[k]);     // lanes whose group touched the band
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1u;
                    ++recounts;
                    const int exact = warp_exact_group_count(R[k], t[k], src, sS + g, sQ + g, cut);
                    if ((threadIdx.x & 31) == src) clo[k] = exact;
                }
                total[k] += clo[k];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KH; ++k) {
        int h = hid[k];
        if (h < h1 && counts[h] >= 0) {
            if (gridDim.y == 1) counts[h] = total[k];
            else if (total[k]) atomicAdd(&counts[h], total[k]);
        }
    }
    if ((tid & 31) == 0 && recounts) atomicAdd(const_cast<unsigned long long*>(&st->score_recounts), (unsigned long long)recounts);
}

// Packed variant of the screen: Blackwell's fma.rn.f32x2 / add.rn.f32x2 / mul.rn.f32x2 (FFMA2,
// FADD2, FMUL2) evaluate two hypotheses per instruction.  The math throughput per lane is the
// same as scalar FFMA, but the instruction count halves, which matters because the scalar
// kernels are bound by instruction issue (1 warp-instruction / clock / SM sub-partition), not by
// the FMA pipe.  Pair tiles are staged in shared memory pre-duplicated ((sx,sx),(sy,sy),...) and
// pre-negated (-q) so every operand of the packed chain is a 64-bit register pair.
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

__device__ __forceinline__ float screen_lo_hi(const float* R, const float* t, float smax, float qmax, float thr, float cut, float& lo, float& hi) {
    float rowsum = fmaxf(fabsf(R[0]) + fabsf(R[1]) + fabsf(R[2]), fmaxf(fabsf(R[3]) + fabsf(R[4]) + fabsf(R[5]), fabsf(R[6]) + fabsf(R[7]) + fabsf(R[8])));
    float A = rowsum * smax + fmaxf(fabsf(t[0]), fmaxf(fabsf(t[1]), fabsf(t[2]))) + qmax;
    float e = A * 7.152557373046875e-7f;                        // 12 * 2^-24 * A
    float beta = 1.25f * (4.0f * (thr * 1.02f + e) * e + 2e-6f * cut);      // 25 % on top of the derived bound
    if (beta < 0.5f * cut) { lo = cut - beta; hi = cut + beta; } else { lo = -INFINITY; hi = INFINITY; }
    return beta;
}

// Optional restriction of a scoring launch, decided ON THE DEVICE (bail-out phases): an id list instead
// of the contiguous range, and a pair sub-range.  Null members mean "everything".
struct ScoreSubset {
    const int* list;            // hypothesis ids, or null for h0 + index
    const int* n_list;          // number of ids in `list` (device scalar), or null for h1 - h0
    const unsigned* prange;     // [lo, hi) pair range (multiples of the pair tile), or null for all pairs
    int always_atomic;          // counts accumulate across launches
};

template <int KP>      // packed hypothesis pairs per thread (KH = 2 * KP hypotheses)
__global__ void __launch_bounds__(kScoreThreads)
score_screen2_kernel(const float* __restrict__ hyp, int H, int h0, int h1,
                     const float4* __restrict__ pairs, unsigned pair_stride, unsigned pairs_per_y,
                     float cut, float thr, const DeviceState* __restrict__ st, int* __restrict__ counts, ScoreSubset sub) {
    constexpr int KH = 2 * KP;
    if (st->score_exit) return;                        // an earlier chunk met fitness > confidence: the reference broke there
    const int n_items = sub.n_list ? *sub.n_list : (h1 - h0);
    if ((int)(blockIdx.x * KH * kScoreThreads) >= n_items) return;
    unsigned r_lo = 0u, r_hi = pair_stride;
    if (sub.prange) { r_lo = sub.prange[0]; r_hi = min(sub.prange[1], pair_stride); }
    {
        const unsigned yb = blockIdx.y * pairs_per_y;
        if (max(yb, r_lo) >= min(yb + pairs_per_y, r_hi)) return;
    }
    __shared__ float4 sP0[kPairTile];      // (sx, sx, sy, sy)
    __shared__ float4 sP1[kPairTile];      // (sz, sz, -qx, -qx)
    __shared__ float4 sP2[kPairTile];      // (-qy, -qy, -qz, -qz)
    const int tid = threadIdx.x, lane = tid & 31;
    float R[KH][9], t[KH][3], lo[KH], hi[KH];
    unsigned long long Rp[KP][9], tp[KP][3];
    int total[KH], hid[KH];
    unsigned recounts = 0;
    const float smax = __uint_as_float(st->pair_smax_bits), qmax = __uint_as_float(st->pair_qmax_bits);
#pragma unroll
    for (int k = 0; k < KH; ++k) {
        const int idx = (blockIdx.x * KH + k) * kScoreThreads + tid;
        const int ic = idx < n_items ? idx : n_items - 1;
        const int hc = sub.list ? sub.list[ic] : h0 + ic;
        hid[k] = idx < n_items ? hc : -1;
#pragma unroll
        for (int e = 0; e < 9; ++e) R[k][e] = hyp[(size_t)e * H + hc];
#pragma unroll
        for (int e = 0; e < 3; ++e) t[k][e] = hyp[(size_t)(9 + e) * H + hc];
        total[k] = 0;
        screen_lo_hi(R[k], t[k], smax, qmax, thr, cut, lo[k], hi[k]);
    }
#pragma unroll
    for (int p = 0; p < KP; ++p) {
#pragma unroll
        for (int e = 0; e < 9; ++e) Rp[p][e] = pk2(R[2 * p][e], R[2 * p + 1][e]);
#pragma unroll
        for (int e = 0; e < 3; ++e) tp[p][e] = pk2(t[2 * p][e], t[2 * p + 1][e]);
    }
    const unsigned p_begin = max(blockIdx.y * pairs_per_y, r_lo);
    const unsigned p_end = min(blockIdx.y * pairs_per_y + pairs_per_y, r_hi);
    for (unsigned base = p_begin; base < p_end; base += kPairTile) {
        __syncthreads();
        for (unsigned e = tid; e < kPairTile; e += kScoreThreads) {
            const float4 s = pairs[base + e], q = pairs[pair_stride + base + e];
            sP0[e] = make_float4(s.x, s.x, s.y, s.y);
            sP1[e] = make_float4(s.z, s.z, -q.x, -q.x);
            sP2[e] = make_float4(-q.y, -q.y, -q.z, -q.z);
        }
        __syncthreads();
#pragma unroll 1
        for (int g = 0; g < kPairTile; g += kGroup) {
            int clo[KH], chi[KH];
#pragma unroll
            for (int k = 0; k < KH; ++k) { clo[k] = 0; chi[k] = 0; }
#pragma unroll 8
            for (int j = 0; j < kGroup; ++j) {
                const ulonglong2 a = reinterpret_cast<const ulonglong2*>(sP0)[g + j];     // SX, SY
                const ulonglong2 b = reinterpret_cast<const ulonglong2*>(sP1)[g + j];     // SZ, -QX
                const ulonglong2 c = reinterpret_cast<const ulonglong2*>(sP2)[g + j];     // -QY, -QZ
#pragma unroll
                for (int p = 0; p < KP; ++p) {
                    unsigned long long dx = fma2(Rp[p][0], a.x, fma2(Rp[p][1], a.y, fma2(Rp[p][2], b.x, add2(tp[p][0], b.y))));
                    unsigned long long dy = fma2(Rp[p][3], a.x, fma2(Rp[p][4], a.y, fma2(Rp[p][5], b.x, add2(tp[p][1], c.x))));
                    unsigned long long dz = fma2(Rp[p][6], a.x, fma2(Rp[p][7], a.y, fma2(Rp[p][8], b.x, add2(tp[p][2], c.y))));
                    unsigned long long d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    float d2a, d2b; upk2(d2, d2a, d2b);
                    count_below(clo[2 * p], d2a, lo[2 * p]);         count_not_above(chi[2 * p], d2a, hi[2 * p]);
                    count_below(clo[2 * p + 1], d2b, lo[2 * p + 1]); count_not_above(chi[2 * p + 1], d2b, hi[2 * p + 1]);
                }
            }
#pragma unroll
            for (int k = 0; k < KH; ++k) {
                unsigned need = __ballot_sync(0xffffffffu, clo[k] != chi[k]);
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1u;
                    ++recounts;
                    // exact re-count of this 32-pair group for lane src's hypothesis k, one pair per lane
                    float Rb[9], tb[3];
#pragma unroll
                    for (int e = 0; e < 9; ++e) Rb[e] = __shfl_sync(0xffffffffu, R[k][e], src);
#pragma unroll
                    for (int e = 0; e < 3; ++e) tb[e] = __shfl_sync(0xffffffffu, t[k][e], src);
                    const float4 pa = sP0[g + lane], pb = sP1[g + lane], pc = sP2[g + lane];
                    const float sx = pa.x, sy = pa.z, sz = pb.x, qx = -pb.z, qy = -pc.x, qz = -pc.z;
                    float x = (Rb[0] * sx + (Rb[1] * sy + Rb[2] * sz)) + tb[0];
                    float y = (Rb[3] * sx + (Rb[4] * sy + Rb[5] * sz)) + tb[1];
                    float z = (Rb[6] * sx + (Rb[7] * sy + Rb[8] * sz)) + tb[2];
                    float ex = x - qx, ey = y - qy, ez = z - qz;
                    float d2 = ex * ex + (ey * ey + ez * ez);
                    const int exact = __popc(__ballot_sync(0xffffffffu, d2 < cut));
                    if (lane == src) clo[k] = exact;
                }
                total[k] += clo[k];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KH; ++k) {
        int h = hid[k];
        if (h >= 0 && counts[h] >= 0) {
            if (gridDim.y == 1 && !sub.always_atomic) counts[h] = total[k];
            else if (total[k]) atomicAdd(&counts[h], total[k]);
        }
    }
    if (lane == 0 && recounts) atomicAdd(const_cast<unsigned long long*>(&st->score_recounts), (unsigned long long)recounts);
}

// ---------------------------------------------------------------------------------
// Bail-out scoring (b3d_set_score_mode(ctx, 3)): the exact argmax without scoring every pair.
// A hypothesis whose count so far plus the pairs not yet visited cannot reach the best FULL count
// already known can neither win nor tie, so it is dropped (the classic RANSAC bail-out test, in
// its exact form).  Everything is decided on the device; the host only enqueues a fixed sequence.
//   phase 0: all ids on the first ~6 % of the pairs -> the id with the most inliers so far is scored
//            on all pairs -> B (a true lower bound on the maximum)
//   phase 1: all ids up to pair P1 = (1 - B/Nc + 2 %) Nc, the point where an id with no inlier at all
//            becomes hopeless -> prune
//   phase 2/3: survivors only, pruned once more in between.
// The bound is capped at the smallest count whose fitness exceeds `confidence`, so a pruned id can
// never be the one that would have triggered the reference's early exit (registration.cpp:290).
// Pruned ids end with count -4; survivors and the phase-0 candidate carry exact full counts.
// ---------------------------------------------------------------------------------
struct BailState {
    int B;                      // best full / partial count known (lower bound of the maximum)
    int bound;                  // min(B, first count whose fitness > confidence)
    int cand, cand_count;
    int n_list[2];
    unsigned prange[2];
    unsigned p1, p2;
    unsigned long long cand_key;
};

__global__ void bail_pick_kernel(const int* __restrict__ counts, int h0, int h1, BailState* bs) {
    unsigned long long best = 0ull;
    for (int h = h0 + blockIdx.x * blockDim.x + threadIdx.x; h < h1; h += gridDim.x * blockDim.x) {
        int c = counts[h];
        if (c < 0) continue;
        unsigned long long key = ((unsigned long long)(unsigned)c << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)h);
        best = key > best ? key : best;
    }
    best = warp_max_u64(best);
    if ((threadIdx.x & 31) == 0 && best) atomicMax(&bs->cand_key, best);
}

// step = 0: after phase 0 -> candidate list of one id, all remaining pairs
// step = 1: after the candidate's full score -> B, bound, P1; hide the candidate from later phases
// step = 2: after the first prune -> range [P1, P2)        step = 3: after the second prune -> [P2, end)
// step = 4: restore the candidate's exact count
__global__ void bail_plan_kernel(int step, BailState* bs, int* counts, int* list_a, unsigned p0, unsigned stride, unsigned n_pairs,
                                 float n_src_f, float confidence) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (step == 0) {
        bs->cand = bs->cand_key ? (int)(0xFFFFFFFFu - (unsigned)(bs->cand_key & 0xFFFFFFFFull)) : -1;
        list_a[0] = bs->cand < 0 ? 0 : bs->cand;
        bs->n_list[0] = bs->cand < 0 ? 0 : 1;
        bs->prange[0] = p0; bs->prange[1] = stride;
    } else if (step == 1) {
        bs->cand_count = bs->cand < 0 ? 0 : counts[bs->cand];
        bs->B = bs->cand_count;
        if (bs->cand >= 0) counts[bs->cand] = -5;                         // scored in full already: skip it in the phases
        // smallest count whose fitness exceeds the confidence (never pruned below it)
        int lo = 0, hi = (int)n_pairs + 1;                                 // fitness(hi) need not exceed: hi acts as "none"
        while (lo < hi) { int mid = lo + (hi - lo) / 2; if ((float)mid / n_src_f > confidence) hi = mid; else lo = mid + 1; }
        bs->bound = min(bs->B, lo);
        float f1 = 1.0f - (float)bs->bound / (float)n_pairs + 0.02f;
        unsigned p1 = (unsigned)(fminf(fmaxf(f1, 0.0f), 1.0f) * (float)n_pairs);
        p1 = ((p1 + kPairTile - 1) / kPairTile) * kPairTile;
        p1 = min(max(p1, p0), stride);
        bs->p1 = p1; bs->p2 = min(stride, ((p1 + (stride - p1) / 2 + kPairTile - 1) / kPairTile) * kPairTile);
        bs->prange[0] = p0; bs->prange[1] = p1;
    } else if (step == 2) {
        bs->prange[0] = bs->p1; bs->prange[1] = bs->p2;
    } else if (step == 3) {
        bs->prange[0] = bs->p2; bs->prange[1] = stride;
    } else if (step == 4) {
        if (bs->cand >= 0) counts[bs->cand] = bs->cand_count;
    }
}

// keep ids that can still reach the bound; src == null means the contiguous range [h0,h1)
__global__ void bail_prune_kernel(const int* __restrict__ src, const int* __restrict__ n_src_list, int h0, int h1, int* __restrict__ counts,
                                  BailState* bs, int which_done /* 1: pairs < p1 visited, 2: pairs < p2 visited */, unsigned n_pairs,
                                  int* __restrict__ dst, int* __restrict__ n_dst) {
    const int n_items = src ? *n_src_list : (h1 - h0);
    const unsigned done = min(which_done == 1 ? bs->p1 : bs->p2, n_pairs);
    const int remaining = (int)(n_pairs - done);
    const int bound = bs->bound;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false; int h = -1;
    if (idx < n_items) {
        h = src ? src[idx] : h0 + idx;
        const int c = counts[h];
        if (c >= 0) {
            keep = c + remaining >= bound;
            if (!keep) counts[h] = -4;
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && ballot) base = atomicAdd(n_dst, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) dst[base + __popc(ballot & ((1u << lane) - 1u))] = h;
}

// ---------------------------------------------------------------------------------
// selection
// ---------------------------------------------------------------------------------

// mode 0: exit key (first id with fitness > confidence).  mode 1: best key over ids <= limit.
// mode 2: exit key -> out_key[0] AND the best key over all ids of the range -> out_key[1] in one pass (sharded selection).
__global__ void select_kernel(const int* __restrict__ counts, int h0, int h1, float n_src_f, float confidence,
                              int mode, const long long* __restrict__ limit_key, unsigned long long* __restrict__ out_key) {
    if (mode == 2) {
        unsigned long long ex = 0ull, best = 0ull;
        for (int h = h0 + blockIdx.x * blockDim.x + threadIdx.x; h < h1; h += gridDim.x * blockDim.x) {
            const int c = counts[h];
            if (c <= 0) continue;
            const float fitness = (float)c / n_src_f;
            const unsigned long long idk = (unsigned long long)(0xFFFFFFFFu - (unsigned)h);
            if (fitness > confidence) ex = idk > ex ? idk : ex;
            if (fitness > 0.0f) { const unsigned long long key = ((unsigned long long)__float_as_uint(fitness) << 32) | idk; best = key > best ? key : best; }
        }
        ex = warp_max_u64(ex); best = warp_max_u64(best);
        __shared__ unsigned long long w2[2][32];
        if ((threadIdx.x & 31) == 0) { w2[0][threadIdx.x >> 5] = ex; w2[1][threadIdx.x >> 5] = best; }
        __syncthreads();
        if (threadIdx.x < 64) {
            const int which = threadIdx.x >> 5, l = threadIdx.x & 31;
            unsigned long long v = (l < (int)(blockDim.x >> 5)) ? w2[which][l] : 0ull;
            v = warp_max_u64(v);
            if (l == 0 && v) atomicMax(out_key + which, v);
        }
        return;
    }
    unsigned limit_id = 0xFFFFFFFFu;
    if (mode == 1 && limit_key) {
        unsigned long long lk = (unsigned long long)(*limit_key);
        if (lk != 0ull) limit_id = 0xFFFFFFFFu - (unsigned)(lk & 0xFFFFFFFFull);
    }
    unsigned long long best = 0ull;
    for (int h = h0 + blockIdx.x * blockDim.x + threadIdx.x; h < h1; h += gridDim.x * blockDim.x) {
        int c = counts[h];
        if (c <= 0) continue;
        float fitness = (float)c / n_src_f;                       // registration.cpp:281
        unsigned long long key;
        if (mode == 0) { if (!(fitness > confidence)) continue; key = (unsigned long long)(0xFFFFFFFFu - (unsigned)h); }
        else {
            if ((unsigned)h > limit_id) continue;
            if (!(fitness > 0.0f)) continue;                      // best_result.fitness starts at 0 (registration.hpp:28)
            key = ((unsigned long long)__float_as_uint(fitness) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)h);
        }
        best = key > best ? key : best;
    }
    best = warp_max_u64(best);
    __shared__ unsigned long long wbest[32];
    if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long v = (threadIdx.x < (blockDim.x >> 5)) ? wbest[threadIdx.x] : 0ull;
        v = warp_max_u64(v);
        if (threadIdx.x == 0 && v) atomicMax(out_key, v);
    }
}

// ---------------------------------------------------------------------------------
// winner: transform, fitness, and rmse with the reference's sequential fp32 sum
// (registration.cpp:270-282).  One block; inlier errors are compacted in order, then
// a single thread adds them left to right.
// ---------------------------------------------------------------------------------
constexpr int kFinishThreads = 1024;
constexpr int kFinishProducers = kFinishThreads - 32;      // warps 0..30 evaluate pairs, warp 31 lane 0 adds (highest-id warp
                                                           // wins issue arbitration on its sub-partition)
constexpr int kFinishPer = 4;                              // consecutive pairs per producer thread per tile
constexpr int kFinishTile = kFinishProducers * kFinishPer; // 3968 pairs: amortises the per-tile latency

// One block, software-pipelined: the 31 producer warps evaluate a tile of pairs with the reference
// arithmetic and compact the inliers' err^2 IN ORDER into one of two shared buffers while a single
// consumer thread adds the previous tile left to right — the reference's sequential fp32 sum
// (registration.cpp:277) cannot be re-associated without changing its bits, so the dependent FADD
// chain (one add per inlier) is the floor; everything else is hidden behind it.
__global__ void __launch_bounds__(kFinishThreads)
finish_kernel(const long long* __restrict__ key_ptr, const uint32_t* __restrict__ draws,
              const float4* __restrict__ pairs, unsigned n_pairs, unsigned pair_stride,
              float n_src_f, float thr, DeviceState* __restrict__ st) {
    __shared__ __align__(16) float vals[2][kFinishTile];
    __shared__ unsigned tile_count[2];
    __shared__ unsigned warp_cnt[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long key = (unsigned long long)(*key_ptr);
    float* out = st->out18;
    if (key == 0ull) {                                 // no hypothesis ever beat fitness 0
        if (tid < 16) out[tid] = (tid % 5 == 0) ? 1.0f : 0.0f;
        if (tid == 16) { out[16] = 0.0f; out[17] = 0.0f; out[18] = __int_as_float(-1); }
        return;
    }
    const int h = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    // The winner may have been generated and scored on another rank: rebuild its (R,t) from the
    // index triple with the same Kabsch code (bit-identical by construction).
    __shared__ float Rt[12];
    if (tid == 0) {
        float s3[3][3], q3[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t id = draws[3 * h + k];
            const float4 a = pairs[id], b = pairs[pair_stride + id];
            s3[k][0] = a.x; s3[k][1] = a.y; s3[k][2] = a.z;
            q3[k][0] = b.x; q3[k][1] = b.y; q3[k][2] = b.z;
        }
        Mat3 Rm; float tv[3];
        kabsch_three_points(s3, q3, Rm, tv);
        for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) Rt[r * 3 + cc] = Rm(r, cc); Rt[9 + r] = tv[r]; }
    }
    __syncthreads();
    float R[9], t[3];
#pragma unroll
    for (int e = 0; e < 9; ++e) R[e] = Rt[e];
#pragma unroll
    for (int e = 0; e < 3; ++e) t[e] = Rt[9 + e];
    const unsigned n_tiles = (n_pairs + kFinishTile - 1) / kFinishTile;
    float running = 0.0f;
    int inlier_total = 0;
    for (unsigned tile = 0; tile <= n_tiles; ++tile) {
        if (warp < 31 && tile < n_tiles) {
            // ---- producers: tile `tile` -> buffer tile & 1 ----
            const unsigned i0 = tile * kFinishTile + (unsigned)tid * kFinishPer;
            float e2[kFinishPer]; bool in[kFinishPer]; unsigned mine = 0;
#pragma unroll
            for (int j = 0; j < kFinishPer; ++j) {
                e2[j] = 0.0f; in[j] = false;
                if (i0 + j < n_pairs) {
                    float4 s = pairs[i0 + j], q = pairs[pair_stride + i0 + j];
                    float x = (R[0] * s.x + (R[1] * s.y + R[2] * s.z)) + t[0];
                    float y = (R[3] * s.x + (R[4] * s.y + R[5] * s.z)) + t[1];
                    float z = (R[6] * s.x + (R[7] * s.y + R[8] * s.z)) + t[2];
                    float dx = x - q.x, dy = y - q.y, dz = z - q.z;
                    float err = sqrtf(dx * dx + (dy * dy + dz * dz));
                    in[j] = err < thr;
                    e2[j] = err * err;
                    mine += in[j] ? 1u : 0u;
                }
            }
            const unsigned inc_lane = warp_inclusive_scan(mine);           // ordered: thread order == pair order
            if (lane == 31) warp_cnt[warp] = inc_lane;
            asm volatile("bar.sync 1, %0;" ::"n"(kFinishProducers) : "memory");
            if (warp == 0) {                           // exclusive scan of the 31 producer-warp counts
                unsigned c = (lane < 31) ? warp_cnt[lane] : 0u;      // lane l holds warp l's count (warp 31 is the consumer)
                unsigned inc = warp_inclusive_scan(c);
                warp_cnt[lane] = inc - c;
                if (lane == 31) tile_count[tile & 1] = inc;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kFinishProducers) : "memory");
            unsigned at = warp_cnt[warp] + inc_lane - mine;
#pragma unroll
            for (int j = 0; j < kFinishPer; ++j) if (in[j]) vals[tile & 1][at++] = e2[j];
        } else if (tid == kFinishProducers && tile > 0) {
            // ---- consumer: adds tile - 1 left to right ----
            const float* v = vals[(tile - 1) & 1];
            const unsigned m = tile_count[(tile - 1) & 1];
            float acc = running;
            unsigned k = 0;
            if (m >= 16) {                             // register ping-pong: the next 8 loads fly under the current 8 adds
                float4 a0 = *reinterpret_cast<const float4*>(v), a1 = *reinterpret_cast<const float4*>(v + 4), b0, b1;
                for (; k + 24 <= m; k += 16) {
                    b0 = *reinterpret_cast<const float4*>(v + k + 8);  b1 = *reinterpret_cast<const float4*>(v + k + 12);
                    acc += a0.x; acc += a0.y; acc += a0.z; acc += a0.w; acc += a1.x; acc += a1.y; acc += a1.z; acc += a1.w;
                    a0 = *reinterpret_cast<const float4*>(v + k + 16); a1 = *reinterpret_cast<const float4*>(v + k + 20);
                    acc += b0.x; acc += b0.y; acc += b0.z; acc += b0.w; acc += b1.x; acc += b1.y; acc += b1.z; acc += b1.w;
                }
                acc += a0.x; acc += a0.y; acc += a0.z; acc += a0.w; acc += a1.x; acc += a1.y; acc += a1.z; acc += a1.w;
                k += 8;
            }
            for (; k < m; ++k) acc += v[k];
            running = acc;
            inlier_total += (int)m;
        }
        __syncthreads();                               // tile produced / previous tile consumed
    }
    if (tid == kFinishProducers) {
        int inliers = inlier_total;                    // recounted here: the winner may have been scored on another rank
        float fitness = (float)inliers / n_src_f;
        float rmse = inliers > 0 ? sqrtf(running / (float)inliers) : 999.0f;
        // Matrix4f column-major: T(r,c) at c*4 + r
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) out[c * 4 + r] = R[r * 3 + c]; out[12 + r] = t[r]; out[r * 4 + 3] = 0.0f; }
        out[15] = 1.0f;
        out[16] = fitness; out[17] = rmse; out[18] = __int_as_float(h);
    }
}

// ---------------------------------------------------------------------------------
// winner, parallel form (default): the same reference-order sum of err^2 (registration.cpp:277) through the exact
// sequential-sum scheme of b3d_ess.cuh — terms (err^2 of the inliers, +0 elsewhere, which leaves an fp32 running sum
// untouched) -> block summaries -> one warp walks them.  Bit-identical to the one-chain kernel above, which is kept as
// b3d_set_finish_mode(ctx, 1) for cross-checks.
// ---------------------------------------------------------------------------------
__global__ void finish_head_kernel(const long long* __restrict__ key_ptr, const uint32_t* __restrict__ draws, const float4* __restrict__ pairs,
                                   unsigned pair_stride, DeviceState* __restrict__ st) {
    if (threadIdx.x != 0) return;
    const unsigned long long key = (unsigned long long)(*key_ptr);
    st->seq_count = 0u;
    if (key == 0ull) { st->fin_none = 1; return; }      // no hypothesis ever beat fitness 0
    st->fin_none = 0;
    const int h = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    // the winner may have been generated and scored on another rank: rebuild its (R,t) from the index triple
    float s3[3][3], q3[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t id = draws[3 * h + k];
        const float4 a = pairs[id], b = pairs[pair_stride + id];
        s3[k][0] = a.x; s3[k][1] = a.y; s3[k][2] = a.z;
        q3[k][0] = b.x; q3[k][1] = b.y; q3[k][2] = b.z;
    }
    Mat3 Rm; float tv[3];
    kabsch_three_points(s3, q3, Rm, tv);
    for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) st->fin_Rt[r * 3 + cc] = Rm(r, cc); st->fin_Rt[9 + r] = tv[r]; }
    st->fin_id = h;
}
struct FinishTerms {
    const float4* pairs; unsigned pair_stride; float thr; const DeviceState* st;
    __device__ __forceinline__ bool operator()(unsigned i, float (&t)[1]) const {
        const float* Rt = st->fin_Rt;
        const float4 s = pairs[i], q = pairs[pair_stride + i];
        const float x = (Rt[0] * s.x + (Rt[1] * s.y + Rt[2] * s.z)) + Rt[9];
        const float y = (Rt[3] * s.x + (Rt[4] * s.y + Rt[5] * s.z)) + Rt[10];
        const float z = (Rt[6] * s.x + (Rt[7] * s.y + Rt[8] * s.z)) + Rt[11];
        const float dx = x - q.x, dy = y - q.y, dz = z - q.z;
        const float err = sqrtf(dx * dx + (dy * dy + dz * dz));
        const bool in = err < thr;                            // registration.cpp:275
        t[0] = in ? err * err : 0.0f;                         // total_error += err * err, :277
        return in;
    }
};
__global__ void __launch_bounds__(ess::kChainThreads)
finish_tail_kernel(const float* __restrict__ terms, const ess::BlockSummary* __restrict__ summ, unsigned n, float n_src_f, DeviceState* __restrict__ st) {
    float* out = st->out18;
    if (st->fin_none) {
        if (threadIdx.x < 16) out[threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.0f : 0.0f;
        if (threadIdx.x == 16) { out[16] = 0.0f; out[17] = 0.0f; out[18] = __int_as_float(-1); }
        return;
    }
    extern __shared__ __align__(128) unsigned char ess_smem[];
    ess::ChainSmem& sm = *reinterpret_cast<ess::ChainSmem*>(ess_smem);
    const float total = ess::chain(terms, summ, n, sm, nullptr);
    if (threadIdx.x != 0) return;
    const int inliers = (int)st->seq_count;                   // recounted here: the winner may have been scored on another rank
    const float* Rt = st->fin_Rt;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) out[c * 4 + r] = Rt[r * 3 + c]; out[12 + r] = Rt[9 + r]; out[r * 4 + 3] = 0.0f; }
    out[15] = 1.0f;
    out[16] = (float)inliers / n_src_f;                       // registration.cpp:281
    out[17] = inliers > 0 ? sqrtf(total / (float)inliers) : 999.0f;
    out[18] = __int_as_float(st->fin_id);
}

// Diagnostic (b3d_sequential_sum): the exact-sum passes over caller-supplied terms, so the machinery the rmse and ICP sums
// ride on can be checked directly against `for (x : terms) s += x` on adversarial inputs (ties, cancellation, binade
// crossings, zeros, denormals, non-finite terms).
struct LoadTerms {
    const float* x;
    __device__ __forceinline__ bool operator()(unsigned i, float (&t)[1]) const { t[0] = x[i]; return true; }
};
__global__ void __launch_bounds__(ess::kChainThreads)
sequential_sum_tail_kernel(const float* __restrict__ terms, const ess::BlockSummary* __restrict__ summ, unsigned n, DeviceState* __restrict__ st) {
    extern __shared__ __align__(128) unsigned char ess_smem[];
    ess::ChainSmem& sm = *reinterpret_cast<ess::ChainSmem*>(ess_smem);
    if (threadIdx.x < 4) st->ess_stats[0][threadIdx.x] = 0u;
    __syncwarp();
    const float total = ess::chain(terms, summ, n, sm, st->ess_stats[0]);
    if (threadIdx.x == 0) st->ess_sums[0] = total;
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static float sqrt_cut(float thr) {
    // smallest float c with sqrtf(c) >= thr, so that (sqrtf(d2) < thr) <=> (d2 < c) for d2 >= 0
    if (!(thr > 0.0f)) return 0.0f;                    // err < thr is never true for thr <= 0 (err >= 0)
    float c = thr * thr;
    while (sqrtf(c) >= thr && c > 0.0f) c = nextafterf(c, 0.0f);
    while (sqrtf(c) < thr) c = nextafterf(c, INFINITY);
    return c;
}

static int ensure_raw(b3d_ctx* c, size_t need) {
    if (need <= c->raw_have) return B3D_OK;
    size_t target = need + need / 16 + 4096;
    DevBuf grown;
    B3D_CUDA(c, grown.ensure(sizeof(uint32_t) * target));
    if (c->raw_have) B3D_CUDA(c, cudaMemcpyAsync(grown.p, c->raw.p, sizeof(uint32_t) * c->raw_have, cudaMemcpyDeviceToDevice, c->stream));
    size_t add = target - c->raw_have;
    uint32_t* host = nullptr;
    B3D_CUDA(c, cudaMallocHost(&host, sizeof(uint32_t) * add));
    for (size_t i = 0; i < add; ++i) host[i] = (uint32_t)c->host_rng();
    cudaError_t e = cudaMemcpyAsync(grown.as<uint32_t>() + c->raw_have, host, sizeof(uint32_t) * add, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFreeHost(host);
    if (e != cudaSuccess) { grown.release(); return fail_cuda(c, e, "raw RNG upload", __FILE__, __LINE__); }
    c->raw.release();
    c->raw = grown;
    c->raw_have = target;
    return B3D_OK;
}

int ransac_prepare_impl(b3d_ctx* c, float voxel, int max_iterations, float confidence) {
    if (!c->have_clouds || !c->have_corr) return fail(c, B3D_ERR_STATE, "ransac_prepare: clouds/correspondences not set");
    if (max_iterations < 0) return fail(c, B3D_ERR_INVALID, "ransac_prepare: max_iterations < 0");
    if (max_iterations > (1 << 29)) return fail(c, B3D_ERR_INVALID, "ransac_prepare: max_iterations above 2^29 (3 draws per id must index 32-bit arrays)");
    if (c->n_src >= 0xFFFFFFFFull) return fail(c, B3D_ERR_INVALID, "ransac_prepare: n_src must be < 2^32 - 1");
    c->prepared = false; c->scored = false;
    c->H = max_iterations;
    c->ransac_thr = voxel * 1.5f;                       // registration.cpp:213
    c->ransac_cut = sqrt_cut(c->ransac_thr);
    c->confidence = confidence;
    const int H = max_iterations;
    const unsigned n = (unsigned)c->n_src;
    if (H == 0 || n == 0) { c->prepared = true; return B3D_OK; }
    const unsigned pair_stride = (unsigned)div_up(n, kPairTile) * kPairTile;
    c->pair_stride = pair_stride;
    B3D_CUDA(c, c->pairs.ensure(sizeof(float4) * 2 * (size_t)pair_stride));
    B3D_CUDA(c, c->draws.ensure(sizeof(uint32_t) * 3 * (size_t)H));
    B3D_CUDA(c, c->hyp.ensure(sizeof(float) * 12 * (size_t)H));
    B3D_CUDA(c, c->counts.ensure(sizeof(int) * (size_t)H));

    const uint32_t range = n;
    const uint32_t reject_below = (uint32_t)(0u - range) % range;       // 2^32 mod n
    const double p_rej = (double)reject_below / 4294967296.0;
    const size_t need = 3 * (size_t)H;
    size_t window = (size_t)((double)need / (1.0 - p_rej)) + (size_t)(8.0 * sqrt((double)need * p_rej) / (1.0 - p_rej)) + 64;

    for (int attempt = 0; attempt < 4; ++attempt) {
        int rc = ensure_raw(c, window);
        if (rc != B3D_OK) return rc;
        StageTimer timer(c, 1);
        const unsigned tiles = (unsigned)div_up((long long)window, kScanTile);
        B3D_CUDA(c, c->scan_tmp.ensure(sizeof(unsigned) * (tiles + 1)));
        LemireAccept accept{c->raw.as<uint32_t>(), range, reject_below};
        DrawEmit emit{c->raw.as<uint32_t>(), range, c->draws.as<uint32_t>(), (unsigned)need};
        DeviceState* st = c->state.as<DeviceState>();
        scan_tile_sums_kernel<<<tiles, kScanThreads, 0, c->stream>>>(accept, (unsigned)window, c->scan_tmp.as<unsigned>());
        B3D_LAUNCHED(c);
        scan_tile_offsets_kernel<<<1, kScanThreads, 0, c->stream>>>(c->scan_tmp.as<unsigned>(), tiles, &st->accepted_total);
        B3D_LAUNCHED(c);
        scan_emit_kernel<<<tiles, kScanThreads, 0, c->stream>>>(accept, emit, (unsigned)window, c->scan_tmp.as<unsigned>());
        B3D_LAUNCHED(c);
        B3D_CUDA(c, cudaMemsetAsync(&st->rng_starved, 0, sizeof(int), c->stream));
        if (attempt == 0) {
            B3D_CUDA(c, cudaMemsetAsync(&st->pair_smax_bits, 0, 2 * sizeof(unsigned), c->stream));
            gather_pairs_kernel<<<grid_for(pair_stride, 256), 256, 0, c->stream>>>(c->src4.as<float4>(), c->tgt4.as<float4>(),
                                                                                    c->corr.as<uint32_t>(), n, (unsigned)c->n_tgt, pair_stride,
                                                                                    c->pairs.as<float4>(), st);
            B3D_LAUNCHED(c);
        }
        c->hyp_lo = c->hyp_hi = 0;          // hypotheses are generated lazily, per scored / inspected id range
        // The acceptance window is sized 8 sigma above the expectation; verify it on the host only
        // when rejections are frequent enough for that to matter (huge clouds).
        if (p_rej < 1e-3) break;
        B3D_CUDA(c, cudaMemcpyAsync(&c->h_state->accepted_total, &st->accepted_total, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
        B3D_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->h_state->accepted_total >= need) break;
        if (attempt == 3) return fail(c, B3D_ERR_RNG_WINDOW, "ransac_prepare: RNG acceptance window exhausted");
        window *= 2;
    }
    c->prepared = true; c->counts_pruned = false;
    return B3D_OK;
}

// 3-point Kabsch for ids [h0,h1) (registration.cpp:239-268).  Lazy so that a rank which scores only
// its shard of the ids does not pay for everybody else's hypotheses.
int ransac_generate_impl(b3d_ctx* c, int h0, int h1) {
    if (h1 <= h0 || c->n_src == 0) return B3D_OK;
    auto launch = [&](int a, int b) -> int {
        if (b <= a) return B3D_OK;
        hypothesis_kernel<<<div_up(b - a, 128), 128, 0, c->stream>>>(c->draws.as<uint32_t>(), c->state.as<DeviceState>(), c->pairs.as<float4>(),
                                                                     c->pair_stride, c->H, a, b, c->hyp.as<float>(), c->counts.as<int>());
        B3D_LAUNCHED(c);
        return B3D_OK;
    };
    const bool have = c->hyp_hi > c->hyp_lo;
    if (have && h0 <= c->hyp_hi && h1 >= c->hyp_lo) {          // overlaps / touches: generate only what is missing
        int rc = launch(h0, min(h1, c->hyp_lo)); if (rc != B3D_OK) return rc;
        rc = launch(max(h0, c->hyp_hi), h1); if (rc != B3D_OK) return rc;
        c->hyp_lo = min(c->hyp_lo, h0); c->hyp_hi = max(c->hyp_hi, h1);
    } else {                                                   // disjoint: track the new range (the old data stays valid but untracked)
        int rc = launch(h0, h1); if (rc != B3D_OK) return rc;
        c->hyp_lo = h0; c->hyp_hi = h1;
    }
    return B3D_OK;
}

static int ransac_score_bailout(b3d_ctx* c, int h0, int h1) {
    c->counts_pruned = true;                      // ids it drops read -4 until the next full scoring or prepare
    const unsigned n = (unsigned)c->n_src, stride = c->pair_stride;
    const int nh = h1 - h0;
    B3D_CUDA(c, c->bail_list_a.ensure(sizeof(int) * (size_t)nh));
    B3D_CUDA(c, c->bail_list_b.ensure(sizeof(int) * (size_t)nh));
    B3D_CUDA(c, c->bail_state.ensure(sizeof(BailState)));
    BailState* bs = c->bail_state.as<BailState>();
    int* counts = c->counts.as<int>();
    int* la = c->bail_list_a.as<int>(); int* lb = c->bail_list_b.as<int>();
    const float4* pairs = c->pairs.as<float4>();
    const DeviceState* st = c->state.as<DeviceState>();
    B3D_CUDA(c, cudaMemsetAsync(bs, 0, sizeof(BailState), c->stream));
    B3D_CUDA(c, cudaMemsetAsync(&c->state.as<DeviceState>()->score_recounts, 0, sizeof(unsigned long long), c->stream));
    reset_counts_kernel<<<div_up(nh, 256), 256, 0, c->stream>>>(counts, h0, h1);
    B3D_LAUNCHED(c);
    // launch geometry: worst case (every id, every pair); blocks outside the device-side subset return at once
    const bool wide = nh >= 64 * kScoreThreads;                              // four hypotheses per thread, as in ransac_score_impl
    const int bx = div_up(nh, kScoreThreads * (wide ? 4 : 2));
    int by = 1;
    const int want_blocks = kNumSMs * 64;
    if (bx < want_blocks) by = min(div_up(want_blocks, bx), div_up((long long)n, kPairTile * 4));
    if (by < 1) by = 1;
    unsigned per_y = (unsigned)div_up((long long)n, by);
    per_y = (unsigned)div_up(per_y, kPairTile) * kPairTile;
    by = div_up((long long)n, per_y);
    auto score = [&](int gx, const int* list, const int* n_list) -> int {
        if (wide && gx > 1)
            score_screen2_kernel<2><<<dim3(gx, by), kScoreThreads, 0, c->stream>>>(c->hyp.as<float>(), c->H, h0, h1, pairs, stride, per_y, c->ransac_cut,
                                                                                   c->ransac_thr, st, counts, ScoreSubset{list, n_list, bs->prange, 1});
        else
            score_screen2_kernel<1><<<dim3(gx, by), kScoreThreads, 0, c->stream>>>(c->hyp.as<float>(), c->H, h0, h1, pairs, stride, per_y, c->ransac_cut,
                                                                                   c->ransac_thr, st, counts, ScoreSubset{list, n_list, bs->prange, 1});
        B3D_LAUNCHED(c);
        return B3D_OK;
    };
    auto plan = [&](int step, unsigned p0) -> int {
        bail_plan_kernel<<<1, 32, 0, c->stream>>>(step, bs, counts, la, p0, stride, n, (float)c->n_src, c->confidence);
        B3D_LAUNCHED(c);
        return B3D_OK;
    };
    unsigned p0 = (unsigned)div_up((long long)(0.06 * n), kPairTile) * kPairTile;
    if (p0 > stride) p0 = stride;
    int rc;
    // phase 0: everyone on [0, p0)
    const unsigned pr0[2] = {0u, p0};
    B3D_CUDA(c, cudaMemcpyAsync(bs->prange, pr0, sizeof(pr0), cudaMemcpyHostToDevice, c->stream));
    if ((rc = score(bx, nullptr, nullptr))) return rc;
    bail_pick_kernel<<<grid_for(nh, 256, 4), 256, 0, c->stream>>>(counts, h0, h1, bs);
    B3D_LAUNCHED(c);
    if ((rc = plan(0, p0))) return rc;
    if ((rc = score(1, la, &bs->n_list[0]))) return rc;                   // the candidate on every remaining pair
    if ((rc = plan(1, p0))) return rc;                                     // B, bound, P1, range [p0, P1)
    // phase 1: everyone on [p0, P1), then prune
    if ((rc = score(bx, nullptr, nullptr))) return rc;
    B3D_CUDA(c, cudaMemsetAsync(&bs->n_list[0], 0, 2 * sizeof(int), c->stream));
    bail_prune_kernel<<<div_up(nh, 256), 256, 0, c->stream>>>(nullptr, nullptr, h0, h1, counts, bs, 1, n, la, &bs->n_list[0]);
    B3D_LAUNCHED(c);
    if ((rc = plan(2, p0))) return rc;
    // phase 2: survivors on [P1, P2), prune again
    if ((rc = score(bx, la, &bs->n_list[0]))) return rc;
    bail_prune_kernel<<<div_up(nh, 256), 256, 0, c->stream>>>(la, &bs->n_list[0], h0, h1, counts, bs, 2, n, lb, &bs->n_list[1]);
    B3D_LAUNCHED(c);
    if ((rc = plan(3, p0))) return rc;
    // phase 3: survivors on [P2, end)
    if ((rc = score(bx, lb, &bs->n_list[1]))) return rc;
    return plan(4, p0);
}

static int ransac_score_range(b3d_ctx* c, int h0, int h1) {
    const unsigned n = (unsigned)c->n_src;
    const int nh = h1 - h0;
    // two hypotheses per thread (one packed FFMA2 lane pair) unless there are hardly any; the pair-range
    // split below supplies the parallelism when the id range is short (multi-GPU shards)
    // ... and four (two packed pairs: two independent FFMA2 chains per thread, 123 registers, 16 warps/SM) when the range is long:
    // the extra instruction-level parallelism keeps the FMA pipe busier than the occupancy it costs (63.7 -> 59.7 ms at 1M x 100k;
    // six per thread, 166 registers, is back at 63.1 ms; capping registers for a third block spills: 68.6 ms)
    const int KH = (nh >= 64 * kScoreThreads && c->score_mode != 4) ? 4 : (nh >= 2 * kScoreThreads) ? 2 : 1;
    if (c->score_mode == 3 && nh >= 8192 && n >= 16 * kPairTile) return ransac_score_bailout(c, h0, h1);
    const bool packed_mode = c->score_mode == 0 || c->score_mode == 3 || c->score_mode == 4;
    const int KHe = (KH == 4 && !packed_mode) ? 2 : KH;                     // the un-fused and scalar-screen kernels exist for 1 and 2 only
    const int bx = div_up(nh, kScoreThreads * KHe);
    // Split the pair stream so that the grid is many waves deep: blocks are long-running and
    // compute-bound, so a shallow grid loses up to a full wave to the tail.
    int by = 1;
    const int want_blocks = kNumSMs * 64;
    if (bx < want_blocks) by = min(div_up(want_blocks, bx), div_up((long long)n, kPairTile * 4));
    if (by < 1) by = 1;
    unsigned per_y = (unsigned)div_up((long long)n, by);
    per_y = (unsigned)div_up(per_y, kPairTile) * kPairTile;
    by = div_up((long long)n, per_y);
    if (by > 1 || c->counts_pruned) {
        reset_counts_kernel<<<div_up(nh, 256), 256, 0, c->stream>>>(c->counts.as<int>(), h0, h1);
        B3D_LAUNCHED(c);
        c->counts_pruned = false;
    }
    dim3 grid(bx, by);
    const float4* pairs = c->pairs.as<float4>();
    const DeviceState* st = c->state.as<DeviceState>();
    if (c->score_mode == 1) {               // reference arithmetic for every pair (verification / comparison)
        if (KHe == 2) score_exact_kernel<2><<<grid, kScoreThreads, 0, c->stream>>>(c->hyp.as<float>(), c->H, h0, h1, pairs, n, c->pair_stride, per_y, c->ransac_cut, c->counts.as<int>(), st);
        else         score_exact_kernel<1><<<grid, kScoreThreads, 0, c->stream>>>(c->hyp.as<float>(), c->H, h0, h1, pairs, n, c->pair_stride, per_y, c->ransac_cut, c->counts.as<int>(), st);
    } else if (c->score_mode == 0 || c->score_mode == 3 || c->score_mode == 4) {   // packed FFMA2 screen (two hypotheses per instruction)
        if (KH == 4) score_screen2_kernel<2><<<grid, kScoreThreads, 0, c->stream>>>(c->hyp.as<float>(), c->H, h0, h1, pairs, c->pair_stride, per_y, c->ransac_cut, c->ransac_thr, st, c->counts.as<int>(), ScoreSubset{nullptr, nullptr, nullptr, 0});
        else if (KH == 2) score_screen2_kernel<1><<<grid, kScoreThreads, 0, c->stream>>>(c->hyp.as<float>(), c->H, h0, h1, pairs, c->pair_stride, per_y, c->ransac_cut, c->ransac_thr, st, c->counts.as<int>(), ScoreSubset{nullptr, nullptr, nullptr, 0});
        else         score_screen_kernel<1><<<grid, kScoreThreads, 0, c->stream>>>(c->hyp.as<float>(), c->H, h0, h1, pairs, c->pair_stride, per_y, c->ransac_cut, c->ransac_thr, st, c->counts.as<int>());
    } else {                                // scalar FMA screen
        if (KHe == 2) score_screen_kernel<2><<<grid, kScoreThreads, 0, c->stream>>>(c->hyp.as<float>(), c->H, h0, h1, pairs, c->pair_stride, per_y, c->ransac_cut, c->ransac_thr, st, c->counts.as<int>());
        else         score_screen_kernel<1><<<grid, kScoreThreads, 0, c->stream>>>(c->hyp.as<float>(), c->H, h0, h1, pairs, c->pair_stride, per_y, c->ransac_cut, c->ransac_thr, st, c->counts.as<int>());
    }
    B3D_LAUNCHED(c);
    return B3D_OK;
}

int ransac_score_impl(b3d_ctx* c, int h0, int h1) {
    if (!c->prepared) return fail(c, B3D_ERR_STATE, "ransac_score: call ransac_prepare first");
    if (h0 < 0 || h1 > c->H || h0 > h1) return fail(c, B3D_ERR_INVALID, "ransac_score: bad hypothesis range");
    c->scored = true; c->scored_lo = h0; c->scored_hi = h1;
    if (h0 == h1 || c->n_src == 0) return B3D_OK;
    StageTimer timer(c, 2);
    { int rc = ransac_generate_impl(c, h0, h1); if (rc != B3D_OK) return rc; }
    const unsigned n = (unsigned)c->n_src;
    exit_reset_kernel<<<1, 1, 0, c->stream>>>(c->state.as<DeviceState>());
    B3D_LAUNCHED(c);
    B3D_CUDA(c, cudaMemsetAsync(&c->state.as<DeviceState>()->score_recounts, 0, sizeof(unsigned long long), c->stream));
    // fitness <= 1, so an exit is only possible below confidence 1: then score in chunks of ids (each >= ~2e9 pair evaluations,
    // a millisecond) with a device-side exit flag, so a clean scene stops where the reference's loop breaks
    const int nh_all = h1 - h0;
    long long chunk = nh_all;
    if (c->confidence < 1.0f && c->score_mode != 3) {
        long long want = (long long)(8.0e9 / (double)n);       // >= ~8e9 pair evaluations (four to five milliseconds) per chunk: shorter
        want = want < 32768 ? 32768 : want;                    // chunks cost (6.15 against 5.97 ms for 1e10 pairs in three chunks)
        const long long n_chunks = (nh_all + want - 1) / want;
        if (n_chunks > 1) chunk = ((nh_all + n_chunks - 1) / n_chunks + 1023) / 1024 * 1024;   // equal chunks: no inefficient short tail launch
    }
    for (long long a = h0; a < h1; a += chunk) {
        const int b = (int)((a + chunk < h1) ? a + chunk : h1);
        int rc = ransac_score_range(c, (int)a, b); if (rc != B3D_OK) return rc;
        if (chunk < nh_all) {
            exit_check_kernel<<<div_up(b - (int)a, 256), 256, 0, c->stream>>>(c->counts.as<int>(), (int)a, b, (float)c->n_src, c->confidence, c->state.as<DeviceState>());
            B3D_LAUNCHED(c);
        }
    }
    if (chunk < nh_all) {
        mark_unscored_kernel<<<div_up(nh_all, 256), 256, 0, c->stream>>>(c->counts.as<int>(), h0, h1, c->state.as<DeviceState>());
        B3D_LAUNCHED(c);
    }
    return B3D_OK;
}

int ransac_reduce_impl(b3d_ctx* c, int h0, int h1, const int64_t* limit_key_dev, int64_t* keys_dev) {
    if (!c->prepared) return fail(c, B3D_ERR_STATE, "ransac_reduce: call ransac_prepare first");
    if (h0 < 0 || h1 > c->H || h0 > h1 || !keys_dev) return fail(c, B3D_ERR_INVALID, "ransac_reduce: bad arguments");
    StageTimer timer(c, 3);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(keys_dev);
    const float n_src_f = (float)c->n_src;              // size_t -> float, as `inliers / source.size()` does
    const int blocks = grid_for(h1 - h0, 256, 4);
    if (!limit_key_dev) {
        B3D_CUDA(c, cudaMemsetAsync(keys, 0, 2 * sizeof(unsigned long long), c->stream));
        if (h1 > h0 && c->n_src) {
            select_kernel<<<blocks, 256, 0, c->stream>>>(c->counts.as<int>(), h0, h1, n_src_f, c->confidence, 0, nullptr, keys + 1);
            B3D_LAUNCHED(c);
            select_kernel<<<blocks, 256, 0, c->stream>>>(c->counts.as<int>(), h0, h1, n_src_f, c->confidence, 1,
                                                         reinterpret_cast<const long long*>(keys + 1), keys);
            B3D_LAUNCHED(c);
        }
    } else {
        B3D_CUDA(c, cudaMemsetAsync(keys, 0, sizeof(unsigned long long), c->stream));
        if (h1 > h0 && c->n_src) {
            select_kernel<<<blocks, 256, 0, c->stream>>>(c->counts.as<int>(), h0, h1, n_src_f, c->confidence, 1,
                                                         reinterpret_cast<const long long*>(limit_key_dev), keys);
            B3D_LAUNCHED(c);
        }
    }
    return B3D_OK;
}

// Sharded selection: keys3[0] = best over ids of [h0,h1) up to this range's own first exit, keys3[1] = that exit key (0: none),
// keys3[2] = best over all ids of the range.  Two launches, no host synchronisation.
int ransac_reduce3_impl(b3d_ctx* c, int h0, int h1, unsigned long long* keys3) {
    if (!c->prepared) return fail(c, B3D_ERR_STATE, "ransac_reduce: call ransac_prepare first");
    if (h0 < 0 || h1 > c->H || h0 > h1 || !keys3) return fail(c, B3D_ERR_INVALID, "ransac_reduce: bad arguments");
    StageTimer timer(c, 3);
    B3D_CUDA(c, cudaMemsetAsync(keys3, 0, 3 * sizeof(unsigned long long), c->stream));
    if (h1 > h0 && c->n_src) {
        const float n_src_f = (float)c->n_src;
        const int blocks = grid_for(h1 - h0, 256, 4);
        select_kernel<<<blocks, 256, 0, c->stream>>>(c->counts.as<int>(), h0, h1, n_src_f, c->confidence, 2, nullptr, keys3 + 1);
        B3D_LAUNCHED(c);
        select_kernel<<<blocks, 256, 0, c->stream>>>(c->counts.as<int>(), h0, h1, n_src_f, c->confidence, 1,
                                                     reinterpret_cast<const long long*>(keys3 + 1), keys3);
        B3D_LAUNCHED(c);
    }
    return B3D_OK;
}

int ransac_finish_impl(b3d_ctx* c, const int64_t* keys_dev, float* T, float* fitness, float* rmse, int32_t* best) {
    if (!c->prepared || !keys_dev) return fail(c, B3D_ERR_STATE, "ransac_finish: not prepared");
    DeviceState* st = c->state.as<DeviceState>();
    if (c->H == 0 || c->n_src == 0) {
        for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
        *fitness = 0.0f; *rmse = 0.0f; if (best) *best = -1;
        return B3D_OK;
    }
    if (c->finish_mode == 1) {
        StageTimer timer(c, 3);
        finish_kernel<<<1, kFinishThreads, 0, c->stream>>>(reinterpret_cast<const long long*>(keys_dev), c->draws.as<uint32_t>(),
                                                           c->pairs.as<float4>(), (unsigned)c->n_src, c->pair_stride,
                                                           (float)c->n_src, c->ransac_thr, st);
        B3D_LAUNCHED(c);
    } else {
        const unsigned n = (unsigned)c->n_src;
        const size_t stride = ess::padded_terms(n);
        B3D_CUDA(c, c->ess_terms.ensure(sizeof(float) * stride));
        B3D_CUDA(c, c->ess_bsum.ensure(sizeof(double) * (stride / ess::kBlock)));
        B3D_CUDA(c, c->ess_guess.ensure(sizeof(double) * (stride / ess::kSuperTerms)));
        B3D_CUDA(c, c->ess_summ.ensure(sizeof(ess::BlockSummary) * (stride / ess::kBlock)));
        if (!c->fin_smem_opt_in) {
            B3D_CUDA(c, cudaFuncSetAttribute(finish_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ess::ChainSmem)));
            c->fin_smem_opt_in = true;
        }
        StageTimer timer(c, 3);
        finish_head_kernel<<<1, 32, 0, c->stream>>>(reinterpret_cast<const long long*>(keys_dev), c->draws.as<uint32_t>(), c->pairs.as<float4>(), c->pair_stride, st);
        B3D_LAUNCHED(c);
        FinishTerms fn{c->pairs.as<float4>(), c->pair_stride, c->ransac_thr, st};
        const int term_blocks = div_up(n, ess::kTermsThreads);
        ess::terms_kernel<1><<<term_blocks, ess::kTermsThreads, 0, c->stream>>>(fn, n, &st->fin_none, c->ess_terms.as<float>(), stride, c->ess_bsum.as<double>(),
                                                                                 c->ess_guess.as<double>(), &st->seq_count);
        B3D_LAUNCHED(c);
        ess::summary_kernel<<<dim3((unsigned)div_up(term_blocks, ess::kSummaryWarps), 1), ess::kSummaryWarps * 32, 0, c->stream>>>(
            c->ess_terms.as<float>(), stride, c->ess_bsum.as<double>(), c->ess_guess.as<double>(), n, &st->fin_none, c->ess_summ.as<ess::BlockSummary>());
        B3D_LAUNCHED(c);
        finish_tail_kernel<<<1, ess::kChainThreads, sizeof(ess::ChainSmem), c->stream>>>(c->ess_terms.as<float>(), c->ess_summ.as<ess::BlockSummary>(), n, (float)c->n_src, st);
        B3D_LAUNCHED(c);
    }
    B3D_CUDA(c, cudaMemcpyAsync(c->h_state->out18, st->out18, sizeof(float) * 20, cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaMemcpyAsync(&c->h_state->rng_starved, &st->rng_starved, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    // hypotheses whose index draws fell outside the accepted-draw window were never generated (count -3): the window is sized 8 sigma
    // above its expectation, so this is practically unreachable, but it must not pass silently
    if (c->h_state->rng_starved) return fail(c, B3D_ERR_RNG_WINDOW, "ransac: accepted-draw window too small for this call (hypotheses were skipped)");
    for (int i = 0; i < 16; ++i) T[i] = c->h_state->out18[i];
    *fitness = c->h_state->out18[16];
    *rmse = c->h_state->out18[17];
    int32_t id; memcpy(&id, &c->h_state->out18[18], sizeof(id));
    if (best) *best = id;
    return B3D_OK;
}

int sequential_sum_impl(b3d_ctx* c, const float* terms_host, size_t n_terms, float* out_sum, uint32_t* out_stats) {
    if (n_terms > 0x7fffffffull) return fail(c, B3D_ERR_INVALID, "sequential_sum: too many terms");
    DeviceState* st = c->state.as<DeviceState>();
    const unsigned n = (unsigned)n_terms;
    if (n == 0) { *out_sum = 0.0f; if (out_stats) { out_stats[0] = out_stats[1] = out_stats[2] = 0u; } return B3D_OK; }
    const size_t stride = ess::padded_terms(n);
    B3D_CUDA(c, c->ess_terms.ensure(sizeof(float) * stride));
    B3D_CUDA(c, c->ess_bsum.ensure(sizeof(double) * (stride / ess::kBlock)));
    B3D_CUDA(c, c->ess_guess.ensure(sizeof(double) * (stride / ess::kSuperTerms)));
    B3D_CUDA(c, c->ess_summ.ensure(sizeof(ess::BlockSummary) * (stride / ess::kBlock)));
    if (!c->seq_smem_opt_in) {
        B3D_CUDA(c, cudaFuncSetAttribute(sequential_sum_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ess::ChainSmem)));
        c->seq_smem_opt_in = true;
    }
    float* terms = c->ess_terms.as<float>();
    B3D_CUDA(c, cudaMemcpyAsync(terms, terms_host, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    LoadTerms fn{terms};                                     // in place: every thread reads its own element before storing it
    const int term_blocks = div_up(n, ess::kTermsThreads);
    ess::terms_kernel<1><<<term_blocks, ess::kTermsThreads, 0, c->stream>>>(fn, n, nullptr, terms, stride, c->ess_bsum.as<double>(), c->ess_guess.as<double>(), nullptr);
    B3D_LAUNCHED(c);
    ess::summary_kernel<<<dim3((unsigned)div_up(term_blocks, ess::kSummaryWarps), 1), ess::kSummaryWarps * 32, 0, c->stream>>>(
        terms, stride, c->ess_bsum.as<double>(), c->ess_guess.as<double>(), n, nullptr, c->ess_summ.as<ess::BlockSummary>());
    B3D_LAUNCHED(c);
    sequential_sum_tail_kernel<<<1, ess::kChainThreads, sizeof(ess::ChainSmem), c->stream>>>(terms, c->ess_summ.as<ess::BlockSummary>(), n, st);
    B3D_LAUNCHED(c);
    uint32_t stats[4];
    B3D_CUDA(c, cudaMemcpyAsync(out_sum, &st->ess_sums[0], sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaMemcpyAsync(stats, st->ess_stats[0], sizeof(stats), cudaMemcpyDeviceToHost, c->stream));
    B3D_CUDA(c, cudaStreamSynchronize(c->stream));
    if (out_stats) { out_stats[0] = stats[0]; out_stats[1] = stats[1]; out_stats[2] = stats[2]; }
    return B3D_OK;
}

}  // namespace b3d
