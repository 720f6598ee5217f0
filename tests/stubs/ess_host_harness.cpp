// Host harness for 3dvision_b200/csrc/b3d_ess.cuh (test infrastructure, CPU only).
// Compiles the SAME block_summary() / block_map() / compose() / to_units() / from_units() the kernels use with g++ and
// drives them the way the device passes do (fp64 block sums -> exclusive prefix -> float guesses -> block summaries ->
// per-super-block scan -> in-order walk, the warp spelled out as loops over 32 "lanes"), so the exactness argument in the
// header is checked against native sequential float addition without a GPU.
#include "../../3dvision_b200/csrc/b3d_ess.cuh"
#include <vector>
using namespace b3d::ess;

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
