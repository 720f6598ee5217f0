// b3d_ess.cuh — exact sequential fp32 sums computed in parallel ("ESS").
//
// The reference adds its ICP normal equations, centroids and error sums one matched point at a time in
// fp32 (src/registration.cpp:341-358, 374-386) and its RANSAC total_error likewise (:270-279).  The result
// depends on that order in the last bits, and ICP at a threshold near the noise floor amplifies the last
// bit, so the only way to land on the reference's numbers is to produce
//     s_n = fl(...fl(fl(0 + x_0) + x_1)... + x_{n-1})
// bit for bit.  Done naively that is one dependent FADD per element (4 cycles each: 0.6 ms for 300 000
// points however many SMs there are).  This header removes the dependency without changing a bit.
//
// Frame.  Fix a sign and two adjacent binades [2^E, 2^(E+2)) and measure in units of u = 2^(E-23), the
// lower binade's ulp: a float there is an integer V in [2^23, 2^25), any integer below 2^24 (grid 1), an
// even one above (grid 2).  IEEE round-to-nearest of V + y is "round to the grid of the side the real sum
// falls on, ties to the even grid point".  Adding a multiple of 4 to V moves every grid point of either
// grid to a grid point of the same parity, so as long as no partial sum changes side,
//     run(V + 4q; x_0..x_{m-1}) = run(V; x_0..x_{m-1}) + 4q          (translation invariance)
// where run() is the chain of rounded additions.
//
// Block summary.  For a block of 16 consecutive terms and a GUESS g of the running sum at the block's start
// (fp64 prefix sums of the terms, rounded to float), a thread runs the chain with NATIVE float additions
// from the four starts V(g) + r, r = 0..3, and records the four net advances o[r], together with the
// distance (margin) of the closest partial sum to the frame's three borders 2^E, 2^(E+1), 2^(E+2).
// If the TRUE running sum is V(g) + d with |d| < margin, its chain takes the same side at every step as the
// guessed one, hence ends at V(g) + d + o[d mod 4].  Binade changes inside the block, ties and cancellation
// are all inside the native additions; nothing is approximated.
//
// Walk.  The map d -> d + h[d mod 4] (h = o minus the change of guess to the next block) composes
// associatively, so a warp scan gives every block of a 512-term super-block its offset d from the offset at
// the start.  The sequential pass holds the true sum, and per round checks 32 blocks at once, one per
// lane (same frame? |d| < margin?); everything before the first block that fails is folded in at once, that
// block's 16 terms are added one by one, and the walk resumes behind it.  Correctness never depends on the
// guess — a wrong guess, a partial sum too close to a border, a third binade, a sign change, a non-finite
// term all just take the one-by-one path — only speed does.
//
// The arithmetic core below is `__host__ __device__` so tests/test_ess_host.py can run it against native
// sequential float addition on the CPU (g++, no GPU) on adversarial inputs; the product only ever calls
// it from kernels.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define B3D_HD __host__ __device__ __forceinline__
#else
#define B3D_HD inline
#endif

namespace b3d {
namespace ess {

constexpr int kBlock = 32;                    // terms per block summary
constexpr int kSuper = 32;                    // block summaries per super-block (one per lane)
constexpr int kSuperTerms = kBlock * kSuper;  // 512
constexpr unsigned kFail = 0x80000000u;       // BlockSummary::tag bit: this block must be added term by term
constexpr unsigned kZeroOnZero = 0x40000000u; // BlockSummary::tag value: every term of the block is +-0 and the guessed running sum is 0 —
                                              // a no-op while the true sum is +0 (leading skipped records, sums that never leave zero)
constexpr int kV23 = 1 << 23, kV24 = 1 << 24, kV25 = 1 << 25;
constexpr int kMarginSlack = 8;               // covers the start offsets r <= 3 and the rounding of the fp64 position estimate

struct alignas(16) BlockSummary {
    int Vg;           // the guess in frame units
    unsigned tag;     // frame: (sign << 8) | biased exponent field of the lower binade; kFail if the block is unusable
    int margin;       // the summary holds for true starts V(g) + d with |d| < margin
    int pad;
    int o[4];         // net advance of the chain started at V(g) + r (0 where that start is not a float).  The stored form
                      // (after the scan) holds W[k] = Vg + F[k] + o[(k + F[k]) & 3] here: what the walk adds to its base offset
    int F[4];         // exclusive prefix, within the super-block, of the maps d -> d + h[d & 3] (filled by the warp scan)
};
// W of the header comment: with the walk's base = d_a - F_a[r0], block j starts at offset base + F_j[r0] and ends at
// units base + W_j[r0].
B3D_HD void finish_summary(BlockSummary& b) {
    int w[4];
    for (int k = 0; k < 4; ++k) w[k] = (int)((unsigned)b.Vg + (unsigned)b.F[k] + (unsigned)b.o[((unsigned)k + (unsigned)b.F[k]) & 3u]);
    for (int k = 0; k < 4; ++k) b.o[k] = w[k];
}

B3D_HD unsigned f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    unsigned u; memcpy(&u, &f, 4); return u;
#endif
}
B3D_HD float u2f(unsigned u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
B3D_HD double u2d(unsigned long long u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}

// Frame of a guess: the binade border nearest to g (in ratio) becomes the frame's inner border.
B3D_HD unsigned frame_of(float g) {
    const unsigned gb = f2u(g);
    const unsigned ef = (gb >> 23) & 0xFFu;
    if (ef < 40u || ef > 250u) return kFail;                     // zero / subnormal / tiny / huge / non-finite sums: term by term
    const unsigned EF = (gb & 0x400000u) ? ef : ef - 1u;         // mantissa >= 1.5: border above, else border below
    return ((gb >> 31) << 8) | EF;
}
// x in frame units; false if x is outside the frame (other sign, other binades, zero, non-finite).
B3D_HD bool to_units(unsigned x_bits, unsigned tag, int& V) {
    const unsigned sh = ((x_bits >> 23) & 0xFFu) - (tag & 0xFFu);
    if ((x_bits >> 31) != ((tag >> 8) & 1u) || sh > 1u || (tag & kFail)) return false;
    V = (int)(((x_bits & 0x7FFFFFu) | 0x800000u) << sh);
    return true;
}
// V in [2^23, 2^25), even if >= 2^24  ->  the float
B3D_HD float from_units(int V, unsigned tag) {
    const unsigned up = V >= kV24 ? 1u : 0u;
    return u2f((((tag >> 8) & 1u) << 31) | (((tag & 0xFFu) + up) << 23) | (((unsigned)V >> up) & 0x7FFFFFu));
}

// Summary of x[0..m) (m <= kBlock) for a running sum near `guess`, in frame `tag` (see header).  F is left zero.
B3D_HD BlockSummary block_summary(const float* x, int m, float guess, unsigned tag) {
    BlockSummary r;
    r.Vg = 0; r.tag = kFail; r.margin = 0; r.pad = 0;
    for (int k = 0; k < 4; ++k) { r.o[k] = 0; r.F[k] = 0; }
    int Vg = 0;
    if (m <= 0 || (tag & kFail) || !to_units(f2u(guess), tag, Vg)) return r;
    // sign * 2^(23 - E) as a double: positions in frame units
    const double scale = u2d(((unsigned long long)((tag >> 8) & 1u) << 63) | ((unsigned long long)(1173u - (tag & 0xFFu)) << 52));
    double dmin;
    {
        const double pos = (double)Vg;
        const double a = pos - (double)kV23, b = (double)kV25 - pos, c = fabs(pos - (double)kV24);
        const double ab = a < b ? a : b;
        dmin = ab < c ? ab : c;
    }
    if (!(dmin > (double)kMarginSlack)) return r;                // the four starts would not share g's side
    const bool upper = Vg >= kV24;                               // grid 2: odd starts are not floats (and cannot occur)
    float s0 = from_units(Vg, tag), s2 = from_units(Vg + 2, tag);
    float s1 = upper ? s0 : from_units(Vg + 1, tag), s3 = upper ? s0 : from_units(Vg + 3, tag);
    for (int i = 0; i < m; ++i) {
        const float xi = x[i];
        const double pos = ((double)s0 + (double)xi) * scale;    // where the real sum falls (fp64: off by < 2^-28 units)
        const double a = pos - (double)kV23, b = (double)kV25 - pos, c = fabs(pos - (double)kV24);
        const double ab = a < b ? a : b;
        const double d = ab < c ? ab : c;
        dmin = d < dmin ? d : dmin;                              // a NaN position is caught by the to_units checks below
        s0 = s0 + xi; s1 = s1 + xi; s2 = s2 + xi; s3 = s3 + xi;  // the reference's additions, natively rounded
    }
    if (!(dmin > (double)kMarginSlack) || !(dmin < 1e9)) return r;
    int e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    if (!to_units(f2u(s0), tag, e0) || !to_units(f2u(s1), tag, e1) || !to_units(f2u(s2), tag, e2) || !to_units(f2u(s3), tag, e3)) return r;
    const int margin = (int)dmin - kMarginSlack;
    if (margin < 1) return r;                                    // 8 < dmin < 9: no start offset is covered (and the walk's one-compare
                                                                 // range test |d| < margin, done on unsigned values, needs margin >= 1)
    r.Vg = Vg; r.tag = tag; r.margin = margin;
    r.o[0] = e0 - Vg; r.o[1] = upper ? 0 : e1 - Vg - 1; r.o[2] = e2 - Vg - 2; r.o[3] = upper ? 0 : e3 - Vg - 3;
    return r;
}

// The two frames that contain g: the one whose inner border is nearest to g, and g's binade paired the other way.
// Frames of neighbouring binade pairs overlap and consecutive blocks in different frames cost the walk a round each, so
// both summaries are computed and choose_frame() lets a block keep its predecessor's frame whenever that one works.
B3D_HD unsigned other_frame_of(float g) {
    const unsigned own = frame_of(g);
    if (own & kFail) return kFail;
    return (own & 0x100u) | ((own & 0xFFu) + ((f2u(g) & 0x400000u) ? 0xFFFFFFFFu : 1u));
}
B3D_HD bool choose_second(unsigned tag_a, unsigned tag_b, unsigned predecessor) {    // a: nearest-border frame, b: the other
    if (!(tag_a & kFail) && tag_a == predecessor) return false;
    if (!(tag_b & kFail) && tag_b == predecessor) return true;
    return (tag_a & kFail) && !(tag_b & kFail);
}

// The zero-on-zero case (no frame exists around 0): true iff the guess is 0 and all m terms are zero.
B3D_HD bool block_is_zero_on_zero(const float* x, int m, float guess) {
    if ((f2u(guess) & 0x7FFFFFFFu) != 0u || m <= 0) return false;
    for (int i = 0; i < m; ++i) if ((f2u(x[i]) & 0x7FFFFFFFu) != 0u) return false;
    return true;
}

// h of a block: its advance minus the step of the guess to the next block, so that offsets from the guesses chain up:
// d_next = d + h[d & 3].  If the next guess lies outside this block's frame the walk breaks there anyway (other tag).
B3D_HD void block_map(const BlockSummary& b, float next_guess, int (&h)[4]) {
    int Vn = 0;
    const bool ok = !(b.tag & kFail) && to_units(f2u(next_guess), b.tag, Vn);
    for (int k = 0; k < 4; ++k) h[k] = ok ? b.o[k] - (Vn - b.Vg) : 0;           // all three below 2^25 in magnitude
}
// (a then b) as maps d -> d + h[d & 3]
B3D_HD void compose(const int (&a)[4], const int (&b)[4], int (&c)[4]) {
    // unsigned adds: prefixes that ran through unusable blocks may wrap; only differences inside one run are ever used
    for (int k = 0; k < 4; ++k) c[k] = (int)((unsigned)a[k] + (unsigned)b[((unsigned)k + (unsigned)a[k]) & 3u]);
}

}  // namespace ess
}  // namespace b3d

#if defined(__CUDACC__)
namespace b3d {
namespace ess {

// ---------------------------------------------------------------------------------
// device passes.  Layouts: terms[v][stride] (float, SoA; stride a multiple of kChunkTerms so staged chunks never straddle
// sums), bsum[v][stride / 16] (double block sums), ssum[v][stride / 512] (double super-block sums),
// summ[v][stride / 16] (BlockSummary).  Records that contribute nothing (ICP queries without a match) are stored as
// +0.0 terms: s + 0 == s bit for bit (s is never -0: it starts at +0 and round-to-nearest cancellation gives +0),
// so the sequence need not be compacted and record k is simply the reference's loop index.
// ---------------------------------------------------------------------------------
constexpr int kChunkSuper = 4;                                 // super-blocks staged per shared-memory buffer
constexpr int kChunkBlocks = kChunkSuper * kSuper;             // 128 block summaries (6 KB)
constexpr int kChunkTerms = kChunkBlocks * kBlock;             // 4096 terms (16 KB)
constexpr int kSummaryQuads = sizeof(BlockSummary) / 16;       // int4 per summary
inline size_t padded_terms(size_t n) { return (n + kChunkTerms - 1) / kChunkTerms * kChunkTerms; }

// Pass 1: one CTA of 512 threads per super-block of 512 records.  fn(k, out[NV]) forms record k's NV terms exactly as the
// reference forms them (all +0 when the record is skipped) and returns whether the record counts; the CTA stores the
// terms (coalesced per sum), their fp64 block and super-block sums, and adds its number of counted records to *count.
constexpr int kTermsThreads = kSuperTerms;
template <int NV, class TermFn>
__global__ void __launch_bounds__(kTermsThreads) terms_kernel(TermFn fn, unsigned n, const int* __restrict__ done, float* __restrict__ terms, size_t stride,
                                                               double* __restrict__ bsum, double* __restrict__ ssum, unsigned* __restrict__ count) {
    if (done && *done) return;
    __shared__ double wsum[kTermsThreads / 32][NV];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned k = blockIdx.x * kTermsThreads + threadIdx.x;
    float t[NV];
    bool counted = false;
    if (k < n) counted = fn(k, t);
    else {
#pragma unroll
        for (int v = 0; v < NV; ++v) t[v] = 0.0f;
    }
    if (count) {
        const unsigned c = __popc(__ballot_sync(0xffffffffu, counted));
        if (lane == 0 && c) atomicAdd(count, c);
    }
    const size_t nb_stride = stride / kBlock;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        terms[(size_t)v * stride + k] = t[v];                                  // k < stride always (stride is padded)
        double a = (double)t[v];
#pragma unroll
        for (int s = kBlock / 2; s >= 1; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
        if ((lane & (kBlock - 1)) == 0) bsum[(size_t)v * nb_stride + (k / kBlock)] = a;
#pragma unroll
        for (int s = kBlock; s < 32; s <<= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
        if (lane == 0) wsum[warp][v] = a;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < kTermsThreads / 32; ++w) a += wsum[w][threadIdx.x];
        ssum[(size_t)threadIdx.x * (stride / kSuperTerms) + blockIdx.x] = a;
    }
}

// Pass 2: one warp per (super-block, sum), 8 per CTA.  The CTA adds the fp64 sums of all the super-blocks before its
// first one (the guess only has to land near the running sum, any summation order does), each lane adds the block sums
// before its own block, runs its block's summary for that guess, and a warp scan under compose() turns the per-block
// maps into exclusive prefixes.
constexpr int kSummaryWarps = 8;
static __global__ void __launch_bounds__(kSummaryWarps * 32) summary_kernel(const float* __restrict__ terms, size_t stride, const double* __restrict__ bsum,
                                                                           const double* __restrict__ ssum, unsigned n, const int* __restrict__ done,
                                                                           BlockSummary* __restrict__ summ) {
    if (done && *done) return;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, v = blockIdx.y;
    const size_t nb_stride = stride / kBlock, ns_stride = stride / kSuperTerms;
    const unsigned n_super = (n + kSuperTerms - 1) / kSuperTerms;
    const unsigned sb0 = blockIdx.x * kSummaryWarps;
    __shared__ double red[kSummaryWarps];
    __shared__ double base_s;
    {
        const double* in = ssum + (size_t)v * ns_stride;
        double a = 0.0;
        for (unsigned i = threadIdx.x; i < sb0; i += kSummaryWarps * 32) a += in[i];
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
        if (lane == 0) red[warp] = a;
        __syncthreads();
        if (threadIdx.x == 0) { double b = 0.0; for (int w = 0; w < kSummaryWarps; ++w) b += red[w]; base_s = b; }
        __syncthreads();
    }
    const unsigned sb = sb0 + warp;
    if (sb >= n_super) return;
    double start = base_s;
    for (unsigned w = 0; w < warp; ++w) start += ssum[(size_t)v * ns_stride + sb0 + w];
    const unsigned b = sb * kSuper + lane;                                      // this lane's block
    const bool live = (size_t)b * kBlock < n;
    const double own = live ? bsum[(size_t)v * nb_stride + b] : 0.0;
    double inc = own;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const double o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (unsigned)d) inc += o; }
    // the block's guess must be the SAME float as its predecessor's next_guess (block_map() ties the two blocks' offsets
    // through it), so both come from one fp64 expression: an exclusive prefix formed as inc - own can round the other way
    // at a float tie
    double excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = 0.0;
    const float guess = (float)(start + excl), next_guess = (float)(start + inc);
    BlockSummary r, r2;
    r.Vg = 0; r.tag = kFail; r.margin = 0; r.pad = 0; r2.Vg = 0; r2.tag = kFail; r2.margin = 0; r2.pad = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { r.o[k] = 0; r.F[k] = 0; r2.o[k] = 0; r2.F[k] = 0; }
    if (live) {
        const float4* src = reinterpret_cast<const float4*>(terms + (size_t)v * stride + (size_t)b * kBlock);
        float x[kBlock];
#pragma unroll
        for (int q = 0; q < kBlock / 4; ++q) { const float4 t = src[q]; x[4 * q] = t.x; x[4 * q + 1] = t.y; x[4 * q + 2] = t.z; x[4 * q + 3] = t.w; }
        const int m = (int)min((unsigned)kBlock, n - b * kBlock);
        if (block_is_zero_on_zero(x, m, guess)) { r.tag = kZeroOnZero; r2.tag = kZeroOnZero; }
        else {
            r = block_summary(x, m, guess, frame_of(guess));
            r2 = block_summary(x, m, guess, other_frame_of(guess));
        }
    }
    {   // in block order: keep the predecessor's frame where it works (32 cheap steps; the summaries above were the work)
        unsigned cur = kFail;
        bool second = false;
#pragma unroll 4
        for (int j = 0; j < kSuper; ++j) {
            const bool pick = choose_second(r.tag, r2.tag, cur);
            if (lane == (unsigned)j) second = pick;
            const unsigned chosen = __shfl_sync(0xffffffffu, pick ? r2.tag : r.tag, j);
            if (!(chosen & kFail)) cur = chosen;
        }
        if (second) r = r2;
    }
    // Exclusive prefixes of the block maps under compose() — a SEGMENTED scan: the prefix restarts at every block whose
    // predecessor is unusable or uses another frame (those are the blocks the walk resumes at with a freshly measured
    // offset, which must not be tied to residues the previous run could reach).
    int e[4];
    block_map(r, next_guess, e);
    const unsigned prev_tag = __shfl_up_sync(0xffffffffu, r.tag, 1);
    const bool head = lane == 0 || prev_tag != r.tag || (prev_tag & kFail);
    bool seg = head;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = __shfl_up_sync(0xffffffffu, e[k], d);
        const bool aseg = __shfl_up_sync(0xffffffffu, (int)seg, d) != 0;
        if (lane >= (unsigned)d && !seg) {
            int c[4];
            compose(a, e, c);
#pragma unroll
            for (int k = 0; k < 4; ++k) e[k] = c[k];
            seg = aseg;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int up = __shfl_up_sync(0xffffffffu, e[k], 1); r.F[k] = head ? 0 : up; }
    finish_summary(r);
    int4* out = reinterpret_cast<int4*>(summ + (size_t)v * nb_stride + b);
    out[0] = make_int4(r.Vg, (int)r.tag, r.margin, 0);
    out[1] = make_int4(r.o[0], r.o[1], r.o[2], r.o[3]);                         // W after finish_summary()
    out[2] = make_int4(r.F[0], r.F[1], r.F[2], r.F[3]);
}

__device__ __forceinline__ int pick4(const int4 v, unsigned k) { return k == 0u ? v.x : (k == 1u ? v.y : (k == 2u ? v.z : v.w)); }

// The running sum as the walk carries it: a float, or (while summaries apply) frame units — the round after a clean round
// in the same frame then starts from the integer without a float round trip.
struct Running { float s; int V; unsigned tag; bool units; };
__device__ __forceinline__ float running_float(const Running& r) { return r.units ? from_units(r.V, r.tag) : r.s; }
__device__ __forceinline__ int sel4(const int4 v, unsigned k) { const int lo = (k & 1u) ? v.y : v.x, hi = (k & 1u) ? v.w : v.z; return (k & 2u) ? hi : lo; }

// One super-block (<= 32 block summaries, lane j holds block j's) walked by a converged warp; every lane carries the same
// running sum.  Round, from block a on: the true sum's offset from block a's guess, d_a, selects the start residue r0 whose
// scanned image at a equals d_a mod 4 (the scan is relative to the super-block's first block, so r0 = d_a mod 4 when a = 0);
// with base = d_a - F_a[r0] every lane derives its own offset d_j = base + F_j[r0] and checks frame and margin; everything
// before the first block that fails is folded in at once (the running sum becomes base + W[r0] of the last good lane, in
// units); if the failing block merely uses another frame the next round starts there, else its terms are added one by
// one (broadcast out of shared memory) and the walk resumes behind it.
struct RoundHead { unsigned tag_a; int Vg_a; int4 Fa; };
__device__ __forceinline__ RoundHead round_head(unsigned a, unsigned tag, int Vg, const int4 qf) {     // independent of the running sum
    RoundHead h;
    a &= 31u;
    h.tag_a = __shfl_sync(0xffffffffu, tag, a);
    h.Vg_a = __shfl_sync(0xffffffffu, Vg, a);
    h.Fa.x = __shfl_sync(0xffffffffu, qf.x, a); h.Fa.y = __shfl_sync(0xffffffffu, qf.y, a);
    h.Fa.z = __shfl_sync(0xffffffffu, qf.z, a); h.Fa.w = __shfl_sync(0xffffffffu, qf.w, a);
    return h;
}
__device__ __forceinline__ void walk_super(Running& run, const int4 qa, const int4 qw, const int4 qf, unsigned cnt, const float* __restrict__ tsm,
                                           unsigned n_terms, unsigned& rounds, unsigned& seq_blocks) {
    const unsigned lane = threadIdx.x & 31;
    const int Vg = qa.x, margin = qa.z;
    const unsigned tag = (unsigned)qa.y;
    const unsigned span = 2u * (unsigned)margin - 1u;              // |d| < margin  <=>  (unsigned)(d + margin - 1) < span
    unsigned a = 0;
    auto add_block = [&](unsigned blk) {
        float s = running_float(run);
        const unsigned k0 = blk * kBlock;
        const unsigned m = min((unsigned)kBlock, n_terms - k0);
        ++seq_blocks;
        if (m == kBlock) {
#pragma unroll
            for (int i = 0; i < kBlock; ++i) s = s + tsm[k0 + i];
        } else {
            for (unsigned i = 0; i < m; ++i) s = s + tsm[k0 + i];
        }
        run.s = s; run.units = false;
    };
    RoundHead h;
    h.tag_a = __shfl_sync(0xffffffffu, tag, 0); h.Vg_a = __shfl_sync(0xffffffffu, Vg, 0); h.Fa = make_int4(0, 0, 0, 0);
    while (a < cnt) {
        ++rounds;
        if (h.tag_a == kZeroOnZero && !run.units && f2u(run.s) == 0u) {      // a run of all-zero blocks on a sum that is still +0: no-ops
            const unsigned other = __ballot_sync(0xffffffffu, lane >= a && lane < cnt && tag != kZeroOnZero);
            a = other ? (unsigned)__ffs(other) - 1u : cnt;
            if (a < cnt) h = round_head(a, tag, Vg, qf);
            continue;
        }
        int Vs = run.V;
        bool inside = run.units && run.tag == h.tag_a;
        if (!inside) inside = to_units(f2u(running_float(run)), h.tag_a, Vs);
        const int da = Vs - h.Vg_a;
        const unsigned ra = (unsigned)da & 3u;
        unsigned r0 = 4u;                                          // the start residue whose image at a is d_a mod 4
        if ((((unsigned)h.Fa.w + 3u) & 3u) == ra) r0 = 3u;
        if ((((unsigned)h.Fa.z + 2u) & 3u) == ra) r0 = 2u;
        if ((((unsigned)h.Fa.y + 1u) & 3u) == ra) r0 = 1u;
        if (((unsigned)h.Fa.x & 3u) == ra) r0 = 0u;
        if (!inside || r0 == 4u) {                                 // outside block a's frame, or no residue maps onto d_a: term by term
            const RoundHead nxt = round_head(a + 1u, tag, Vg, qf);
            add_block(a);
            a += 1u; h = nxt;
            continue;
        }
        const int base = (int)((unsigned)da - (unsigned)sel4(h.Fa, r0));
        const int dj = (int)((unsigned)base + (unsigned)sel4(qf, r0));
        const bool ok = tag == h.tag_a && (unsigned)(dj + margin - 1) < span;
        const unsigned bad = __ballot_sync(0xffffffffu, lane >= a && lane < cnt && !ok);
        const unsigned first = bad ? (unsigned)__ffs(bad) - 1u : cnt;
        if (first > a) {
            run.V = __shfl_sync(0xffffffffu, (int)((unsigned)base + (unsigned)sel4(qw, r0)), first - 1u);
            run.tag = h.tag_a; run.units = true;
        }
        if (first >= cnt) break;
        const unsigned tag_f = __shfl_sync(0xffffffffu, tag, first);
        if (first > a && tag_f != h.tag_a && !(tag_f & kFail)) {    // only the frame changes here: the next round starts at this block
            a = first; h = round_head(a, tag, Vg, qf);
            continue;
        }
        const RoundHead nxt = round_head(first + 1u, tag, Vg, qf);
        add_block(first);
        a = first + 1u; h = nxt;
    }
}

// Pass 3 (device function; ONE warp per sum — launch a CTA of kChainThreads = 32): the sum's terms and summaries stream
// through a ring of shared-memory stages filled by TMA bulk copies (cp.async.bulk + mbarrier complete_tx; both arrays are
// contiguous per chunk), issued by lane 0 kRing - 1 chunks ahead, while the warp walks the current chunk.  Returns the
// exact sequential sum in every lane; stats[0] += walk rounds, stats[1] += blocks added term by term, stats[2] += SM cycles / 16.
constexpr int kChainThreads = 32;
constexpr int kRing = 6;                                       // 132 KB of shared memory: the chain kernels opt in to it
constexpr unsigned kChunkBytes = kChunkTerms * 4u + kChunkBlocks * (unsigned)sizeof(BlockSummary);
struct alignas(128) ChainSmem { float t[kRing][kChunkTerms]; int4 q[kRing][kSummaryQuads * kChunkBlocks]; unsigned long long full[kRing]; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ float chain(const float* __restrict__ terms, const BlockSummary* __restrict__ summ, unsigned n, ChainSmem& sm,
                                       unsigned* __restrict__ stats) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned nb = (n + kBlock - 1) / kBlock;
    const unsigned n_chunks = (n + kChunkTerms - 1) / kChunkTerms;
    if (lane == 0) {
        for (int i = 0; i < kRing; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.full[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](unsigned c) {                                 // lane 0 only: both arrays are padded to whole chunks
        const unsigned st = c % kRing, bar = smem_u32(&sm.full[st]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // the stage's last generic reads precede the async writes
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kChunkBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sm.t[st])), "l"(terms + (size_t)c * kChunkTerms), "r"((unsigned)(kChunkTerms * 4)), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sm.q[st])), "l"(summ + (size_t)c * kChunkBlocks), "r"((unsigned)(kChunkBlocks * sizeof(BlockSummary))), "r"(bar) : "memory");
    };
    if (lane == 0) for (unsigned c = 0; c < (unsigned)(kRing - 1) && c < n_chunks; ++c) issue(c);
    Running run; run.s = 0.0f; run.V = 0; run.tag = kFail; run.units = false;
    unsigned rounds = 0, seq_blocks = 0;
    const long long t0 = clock64();
    for (unsigned c = 0; c < n_chunks; ++c) {
        __syncwarp();                                              // every lane is done with chunk c - 1, whose stage is refilled now
        if (lane == 0 && c + kRing - 1 < n_chunks) issue(c + kRing - 1);
        const unsigned st = c % kRing;
        ring_wait(smem_u32(&sm.full[st]), (c / kRing) & 1u);
        const unsigned nb_here = min((unsigned)kChunkBlocks, nb - c * kChunkBlocks);
        const unsigned nt_here = min((unsigned)kChunkTerms, n - c * kChunkTerms);
        const int4* q = sm.q[st] + kSummaryQuads * lane;
        int4 qa = q[0], qw = q[1], qf = q[2];
#pragma unroll
        for (unsigned sblk = 0; sblk < (unsigned)kChunkSuper; ++sblk) {
            if (sblk * kSuper >= nb_here) break;
            const int4 ca = qa, cw = qw, cf = qf;
            if (sblk + 1 < (unsigned)kChunkSuper) {                // the next super-block's summaries load under this one's walk
                const int4* qn = q + kSummaryQuads * kSuper * (sblk + 1);
                qa = qn[0]; qw = qn[1]; qf = qn[2];
            }
            walk_super(run, ca, cw, cf, min((unsigned)kSuper, nb_here - sblk * kSuper), sm.t[st] + sblk * kSuperTerms,
                       nt_here - sblk * kSuperTerms, rounds, seq_blocks);
        }
    }
    if (stats && lane == 0) { stats[0] += rounds; stats[1] += seq_blocks; stats[2] += (unsigned)((clock64() - t0) >> 4); }
    return running_float(run);
}

}  // namespace ess
}  // namespace b3d
#endif
