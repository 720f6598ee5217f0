// Host harness for 3dvision_b200/csrc/b3d_ess.cuh (test infrastructure, CPU only).
// Compiles the SAME block_summary() / block_map() / compose() / to_units() / from_units() the kernels use with g++ and
// drives them the way the device passes do (fp64 block sums -> exclusive prefix -> float guesses -> block summaries ->
// per-super-block scan -> in-order walk, the warp spelled out as loops over 32 "lanes"), so the exactness argument in the
// header is checked against native sequential float addition without a GPU.
#include "../../3dvision_b200/csrc/b3d_ess.cuh"
#include <vector>
using namespace b3d::ess;
static const int kSummaryWarpsHost = 8;            // kSummaryWarps of the device section of the header

extern "C" float ess_seq_sum(const float* x, long n) {          // the reference's loop: acc += x[i], fp32, in order
    volatile float acc = 0.0f;
    for (long i = 0; i < n; ++i) acc = acc + x[i];
    return acc;
}

// walk_super() of the header, lane by lane (same checks, same order; the running sum is kept as a float throughout, which
// is what the device's unit form converts to).  st[0] += blocks folded in through summaries, st[1] += blocks added term by
// term, st[2] += rounds.
static float walk_super_host(float s, const BlockSummary* q, unsigned cnt, const float* t, unsigned n_terms, long* st) {
    unsigned a = 0;
    auto add_block = [&](unsigned blk) {
        const unsigned k0 = blk * kBlock;
        const unsigned m = (n_terms - k0) < (unsigned)kBlock ? (n_terms - k0) : (unsigned)kBlock;
        for (unsigned i = 0; i < m; ++i) { volatile float v = s + t[k0 + i]; s = v; }
        st[1]++;
    };
    while (a < cnt) {
        st[2]++;
        if (q[a].tag == kZeroOnZero && f2u(s) == 0u) {           // a run of all-zero blocks on a sum that is still +0
            unsigned nxt = a;
            while (nxt < cnt && q[nxt].tag == kZeroOnZero) ++nxt;
            st[0] += nxt - a;
            a = nxt;
            continue;
        }
        const unsigned tag_a = q[a].tag;
        const int* Fa = q[a].F;
        int Vs = 0;
        const bool inside = to_units(f2u(s), tag_a, Vs);
        const int da = Vs - q[a].Vg;
        const unsigned ra = (unsigned)da & 3u;
        unsigned r0 = 4u;
        for (int r = 3; r >= 0; --r) if ((((unsigned)Fa[r] + (unsigned)r) & 3u) == ra) r0 = (unsigned)r;
        if (!inside || r0 == 4u) { add_block(a); a += 1u; continue; }
        const int base = (int)((unsigned)da - (unsigned)Fa[r0]);
        unsigned first = cnt;
        for (unsigned l = a; l < cnt; ++l) {
            const int dj = (int)((unsigned)base + (unsigned)q[l].F[r0]);
            const bool ok = q[l].tag == tag_a && (unsigned)(dj + q[l].margin - 1) < 2u * (unsigned)q[l].margin - 1u;
            if (!ok) { first = l; break; }
        }
        if (first > a) { s = from_units((int)((unsigned)base + (unsigned)q[first - 1].o[r0]), tag_a); st[0] += first - a; }   // o holds W
        if (first >= cnt) break;
        const unsigned tag_f = q[first].tag;
        if (first > a && tag_f != tag_a && !(tag_f & kFail)) { a = first; continue; }
        add_block(first);
        a = first + 1u;
    }
    return s;
}

extern "C" float ess_parallel_sum(const float* x, long n, long* stats) {
    const long nb = (n + kBlock - 1) / kBlock;
    const long nbp = (nb + kSuper - 1) / kSuper * kSuper + kSuper;
    std::vector<double> bsum(nbp, 0.0);
    for (long b = 0; b < nb; ++b) {
        // the device adds the doubles in a butterfly; any fp64 order serves as a guess
        double a = 0.0;
        for (long i = b * kBlock; i < n && i < (b + 1) * kBlock; ++i) a += (double)x[i];
        bsum[b] = a;
    }
    std::vector<BlockSummary> summ(nbp);
    std::vector<float> next_guess(nbp, 0.0f);
    double run = 0.0;
    unsigned cur = kFail;
    for (long b = 0; b < nbp; ++b) {
        if (b % kSuper == 0) cur = kFail;                        // the frame choice restarts with every super-block (one warp each)
        BlockSummary r;
        r.Vg = 0; r.tag = kFail; r.margin = 0; r.pad = 0;
        for (int k = 0; k < 4; ++k) { r.o[k] = 0; r.F[k] = 0; }
        BlockSummary r2 = r;
        if (b < nb) {
            const long m = (n - b * kBlock) < kBlock ? (n - b * kBlock) : kBlock;
            if (block_is_zero_on_zero(x + b * kBlock, (int)m, (float)run)) { r.tag = kZeroOnZero; r2.tag = kZeroOnZero; }
            else {
                r = block_summary(x + b * kBlock, (int)m, (float)run, frame_of((float)run));
                r2 = block_summary(x + b * kBlock, (int)m, (float)run, other_frame_of((float)run));
            }
            run += bsum[b];
        }
        if (choose_second(r.tag, r2.tag, cur)) r = r2;
        if (!(r.tag & kFail)) cur = r.tag;
        next_guess[b] = (float)run;
        summ[b] = r;
    }
    for (long b0 = 0; b0 < nbp; b0 += kSuper) {                  // the summary kernel's scan: exclusive prefixes under compose()
        int e[4] = {0, 0, 0, 0};
        for (long j = 0; j < kSuper; ++j) {
            BlockSummary& q = summ[b0 + j];
            const bool head = j == 0 || summ[b0 + j - 1].tag != q.tag || (summ[b0 + j - 1].tag & kFail);   // segmented: see summary_kernel
            if (head) for (int k = 0; k < 4; ++k) e[k] = 0;
            for (int k = 0; k < 4; ++k) q.F[k] = e[k];
            int h[4], c[4];
            block_map(q, next_guess[b0 + j], h);
            compose(e, h, c);
            for (int k = 0; k < 4; ++k) e[k] = c[k];
            finish_summary(q);
        }
    }
    float s = 0.0f;
    long st[3] = {0, 0, 0};
    for (long b0 = 0; b0 < nb; b0 += kSuper) {
        const long cnt = (nb - b0) < kSuper ? (nb - b0) : kSuper;
        s = walk_super_host(s, &summ[b0], (unsigned)cnt, x + b0 * kBlock, (unsigned)(n - b0 * kBlock), st);
    }
    if (stats) { stats[0] = st[0]; stats[1] = st[1]; stats[2] = st[2]; }
    return s;
}

// Smallest margin over the usable block summaries of x (both frames, guesses = the fp64 running sum), or a large number if none.
extern "C" long ess_min_usable_margin(const float* x, long n) {
    const long nb = (n + kBlock - 1) / kBlock;
    double run = 0.0;
    long least = 1L << 40;
    for (long b = 0; b < nb; ++b) {
        const long m = (n - b * kBlock) < kBlock ? (n - b * kBlock) : kBlock;
        const float g = (float)run;
        const unsigned tags[2] = {frame_of(g), other_frame_of(g)};
        for (unsigned tag : tags) {
            const BlockSummary r = block_summary(x + b * kBlock, (int)m, g, tag);
            if (!(r.tag & kFail) && r.margin < least) least = r.margin;
        }
        for (long i = 0; i < m; ++i) run += (double)x[b * kBlock + i];
    }
    return least;
}

// ---------------------------------------------------------------------------------------------------------------------
// The device passes spelled out lane by lane (terms_kernel's fp64 butterflies, summary_kernel's guesses / frame choice /
// Hillis-Steele segmented scan, walk_super's unit-carrying running sum), so that a term sequence on which the GPU and the
// simpler driver above disagree can be replayed on the CPU.  Same header functions, same order of every operation.
// ---------------------------------------------------------------------------------------------------------------------
namespace devlike {
struct Running { float s; int V; unsigned tag; bool units; };
static float running_float(const Running& r) { return r.units ? from_units(r.V, r.tag) : r.s; }
static int sel4(const int* v, unsigned k) { return v[k & 3u]; }

static void walk_super(Running& run, const BlockSummary* q, unsigned cnt, const float* t, unsigned n_terms, long* st) {
    unsigned a = 0;
    auto add_block = [&](unsigned blk) {
        float s = running_float(run);
        const unsigned k0 = blk * kBlock;
        const unsigned m = (n_terms - k0) < (unsigned)kBlock ? (n_terms - k0) : (unsigned)kBlock;
        for (unsigned i = 0; i < m; ++i) { volatile float v = s + t[k0 + i]; s = v; }
        run.s = s; run.units = false;
        st[1]++;
    };
    unsigned tag_a = q[0].tag; int Vg_a = q[0].Vg; int Fa[4] = {0, 0, 0, 0};
    auto head = [&](unsigned i) { i &= 31u; tag_a = q[i].tag; Vg_a = q[i].Vg; for (int k = 0; k < 4; ++k) Fa[k] = q[i].F[k]; };
    while (a < cnt) {
        st[2]++;
        if (tag_a == kZeroOnZero && !run.units && f2u(run.s) == 0u) {
            unsigned nxt = a;
            while (nxt < cnt && q[nxt].tag == kZeroOnZero) ++nxt;
            st[0] += nxt - a;
            a = nxt;
            if (a < cnt) head(a);
            continue;
        }
        int Vs = run.V;
        bool inside = run.units && run.tag == tag_a;
        if (!inside) inside = to_units(f2u(running_float(run)), tag_a, Vs);
        const int da = Vs - Vg_a;
        const unsigned ra = (unsigned)da & 3u;
        unsigned r0 = 4u;
        if ((((unsigned)Fa[3] + 3u) & 3u) == ra) r0 = 3u;
        if ((((unsigned)Fa[2] + 2u) & 3u) == ra) r0 = 2u;
        if ((((unsigned)Fa[1] + 1u) & 3u) == ra) r0 = 1u;
        if (((unsigned)Fa[0] & 3u) == ra) r0 = 0u;
        if (!inside || r0 == 4u) { add_block(a); a += 1u; head(a); continue; }
        const int base = (int)((unsigned)da - (unsigned)sel4(Fa, r0));
        unsigned first = cnt;
        for (unsigned l = a; l < cnt; ++l) {
            const int dj = (int)((unsigned)base + (unsigned)sel4(q[l].F, r0));
            const unsigned span = 2u * (unsigned)q[l].margin - 1u;
            const bool ok = q[l].tag == tag_a && (unsigned)(dj + q[l].margin - 1) < span;
            if (!ok) { first = l; break; }
        }
        if (first > a) { run.V = (int)((unsigned)base + (unsigned)sel4(q[first - 1].o, r0)); run.tag = tag_a; run.units = true; st[0] += first - a; }
        if (first >= cnt) break;
        const unsigned tag_f = q[first].tag;
        if (first > a && tag_f != tag_a && !(tag_f & kFail)) { a = first; head(a); continue; }
        add_block(first);
        a = first + 1u; head(a);
    }
}
}  // namespace devlike

extern "C" float ess_parallel_sum_devicelike(const float* x, long n, long* stats) {
    const long nb = (n + kBlock - 1) / kBlock;
    const long n_super = (n + kSuperTerms - 1) / kSuperTerms;
    const long nbp = n_super * kSuper;
    // terms_kernel: fp64 block sums by xor butterflies (s = 16, 8, 4, 2, 1), super-block sums in warp order
    std::vector<double> bsum(nbp + kSuper, 0.0), ssum(n_super + 1, 0.0);
    for (long sb = 0; sb < n_super; ++sb) {
        double acc = 0.0;
        for (int w = 0; w < kSuper; ++w) {
            double a[32];
            for (int l = 0; l < 32; ++l) { const long k = (sb * kSuper + w) * kBlock + l; a[l] = k < n ? (double)x[k] : 0.0; }
            for (int s = 16; s >= 1; s >>= 1) { double o[32]; for (int l = 0; l < 32; ++l) o[l] = a[l ^ s]; for (int l = 0; l < 32; ++l) a[l] += o[l]; }
            bsum[sb * kSuper + w] = a[0];
            acc += a[0];
        }
        ssum[sb] = acc;
    }
    std::vector<BlockSummary> summ(nbp + kSuper);
    for (long sb = 0; sb < n_super; ++sb) {
        const long sb0 = sb / kSummaryWarpsHost * kSummaryWarpsHost;
        // base_s: the CTA adds ssum[0..sb0) strided over 256 threads, warp butterflies, then the 8 warp sums in order
        double red[kSummaryWarpsHost];
        for (int w = 0; w < kSummaryWarpsHost; ++w) {
            double a[32];
            for (int l = 0; l < 32; ++l) { a[l] = 0.0; for (long i = w * 32 + l; i < sb0; i += kSummaryWarpsHost * 32) a[l] += ssum[i]; }
            for (int s = 16; s >= 1; s >>= 1) { double o[32]; for (int l = 0; l < 32; ++l) o[l] = a[l ^ s]; for (int l = 0; l < 32; ++l) a[l] += o[l]; }
            red[w] = a[0];
        }
        double base_s = 0.0;
        for (int w = 0; w < kSummaryWarpsHost; ++w) base_s += red[w];
        double start = base_s;
        for (long w = sb0; w < sb; ++w) start += ssum[w];
        double own[32], inc[32];
        for (int l = 0; l < 32; ++l) { const long b = sb * kSuper + l; own[l] = (b * kBlock < n) ? bsum[b] : 0.0; inc[l] = own[l]; }
        for (int d = 1; d < 32; d <<= 1) { double o[32]; for (int l = 0; l < 32; ++l) o[l] = l >= d ? inc[l - d] : 0.0; for (int l = 0; l < 32; ++l) if (l >= d) inc[l] += o[l]; }
        BlockSummary r[32], r2[32];
        float next_guess[32];
        for (int l = 0; l < 32; ++l) {
            const long b = sb * kSuper + l;
            const float guess = (float)(start + (l ? inc[l - 1] : 0.0));           // == next_guess[l - 1], bit for bit
            next_guess[l] = (float)(start + inc[l]);
            BlockSummary z; z.Vg = 0; z.tag = kFail; z.margin = 0; z.pad = 0; for (int k = 0; k < 4; ++k) { z.o[k] = 0; z.F[k] = 0; }
            r[l] = z; r2[l] = z;
            if (b * kBlock < n) {
                const int m = (int)((n - b * kBlock) < kBlock ? (n - b * kBlock) : kBlock);
                if (block_is_zero_on_zero(x + b * kBlock, m, guess)) { r[l].tag = kZeroOnZero; r2[l].tag = kZeroOnZero; }
                else { r[l] = block_summary(x + b * kBlock, m, guess, frame_of(guess)); r2[l] = block_summary(x + b * kBlock, m, guess, other_frame_of(guess)); }
            }
        }
        unsigned cur = kFail;
        for (int j = 0; j < 32; ++j) {
            const bool pick = choose_second(r[j].tag, r2[j].tag, cur);
            const unsigned chosen = pick ? r2[j].tag : r[j].tag;
            if (pick) r[j] = r2[j];
            if (!(chosen & kFail)) cur = chosen;
        }
        int e[32][4]; bool seg[32], headf[32];
        for (int l = 0; l < 32; ++l) {
            block_map(r[l], next_guess[l], e[l]);
            const unsigned prev_tag = l ? r[l - 1].tag : r[l].tag;
            headf[l] = l == 0 || prev_tag != r[l].tag || (prev_tag & kFail);
            seg[l] = headf[l];
        }
        for (int d = 1; d < 32; d <<= 1) {
            int ne[32][4]; bool nseg[32];
            for (int l = 0; l < 32; ++l) {
                for (int k = 0; k < 4; ++k) ne[l][k] = e[l][k];
                nseg[l] = seg[l];
                if (l >= d && !seg[l]) { int c[4]; compose(e[l - d], e[l], c); for (int k = 0; k < 4; ++k) ne[l][k] = c[k]; nseg[l] = seg[l - d]; }
            }
            for (int l = 0; l < 32; ++l) { for (int k = 0; k < 4; ++k) e[l][k] = ne[l][k]; seg[l] = nseg[l]; }
        }
        for (int l = 0; l < 32; ++l) {
            for (int k = 0; k < 4; ++k) r[l].F[k] = headf[l] ? 0 : e[l - 1][k];
            finish_summary(r[l]);
            summ[sb * kSuper + l] = r[l];
        }
    }
    devlike::Running run; run.s = 0.0f; run.V = 0; run.tag = kFail; run.units = false;
    long st[3] = {0, 0, 0};
    for (long b0 = 0; b0 < nb; b0 += kSuper) {
        const long cnt = (nb - b0) < kSuper ? (nb - b0) : kSuper;
        devlike::walk_super(run, &summ[b0], (unsigned)cnt, x + b0 * kBlock, (unsigned)(n - b0 * kBlock), st);
    }
    if (stats) { stats[0] = st[0]; stats[1] = st[1]; stats[2] = st[2]; }
    return devlike::running_float(run);
}
