// Host check of mcac_b200/csrc/seq_cumsum.cuh: the closed-form segments of a run of equal weights against the reference's own loop
// (aggregat_list.cpp:131-140: cum[i] = cum[i-1] + w[i], one addition after the other), bit for bit, on random and adversarial runs
// (full and short mantissas, round-to-even ties, starts below / inside / far above the weight, binade crossings, stagnating sums).
// usage: seq_cumsum_host <cases> <seed>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <algorithm>
#include "../../mcac_b200/csrc/seq_cumsum.cuh"

static uint64_t bits(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }

static std::mt19937_64 rng;
static double random_weight() {
    // mantissa with a random number of trailing zero bits, exponent in a modest range
    const int keep = 1 + (int)(rng() % 53);  // significant bits kept
    uint64_t m = (rng() & ((1ULL << 52) - 1)) | (1ULL << 52);
    m &= ~((1ULL << (53 - keep)) - 1);
    if ((rng() & 3) == 0) m |= 1ULL << (53 - keep);  // lowest kept bit set: ties one binade up
    const int e = (int)(rng() % 40) - 20;
    return std::ldexp((double)m, e - 52);
}

static long long check_run(double x0, int i0, int cnt, double v, long long &segs_max) {
    std::vector<seqsum::Seg> segs(seqsum::kMaxSegs);
    int ns = 0;
    const double x_end = seqsum::run_segments(x0, i0, cnt, v, segs.data(), ns, seqsum::kMaxSegs);
    if (ns > seqsum::kMaxSegs) { std::printf("FAIL: %d segments (x0=%a v=%a cnt=%d)\n", ns, x0, v, cnt); return -1; }
    if (ns > segs_max) segs_max = ns;
    // segments are consecutive and cover the run
    int at = i0;
    for (int s = 0; s < ns; s++) {
        if (segs[s].i0 != at || segs[s].cnt <= 0) { std::printf("FAIL: segment %d starts at %d, expected %d\n", s, segs[s].i0, at); return -1; }
        at += segs[s].cnt;
    }
    if (at != i0 + cnt) { std::printf("FAIL: segments end at %d, run at %d\n", at, i0 + cnt); return -1; }
    double acc = x0;
    int s = 0;
    for (int i = i0; i < i0 + cnt; i++) {
        acc = acc + v;
        while (i >= segs[s].i0 + segs[s].cnt) s++;
        const double got = seqsum::segment_value(segs[s], i);
        if (bits(got) != bits(acc)) {
            std::printf("FAIL: entry %d: %a, sequential %a (x0=%a v=%a cnt=%d seg %d/%d i0=%d c=%a)\n", i, got, acc, x0, v, cnt, s, ns, segs[s].i0,
                        segs[s].c);
            return -1;
        }
        if ((i & 1023) == 0 && seqsum::find_segment(segs.data(), ns, i) != s) { std::printf("FAIL: find_segment(%d)\n", i); return -1; }
    }
    if (bits(acc) != bits(x_end)) { std::printf("FAIL: running sum behind the run %a, sequential %a\n", x_end, acc); return -1; }
    return cnt;
}

// The head of the table (seqsum::head_*): the chunked integer-prefix form, run the way a CTA of `T` threads runs it (the scans across
// chunks are serial here), against the plain loop.  `jitter`: relative perturbation of the approximate sums before the chunks (another
// summation order); whatever they are, a head that `head_stitch` accepts must be the sequential sum bit for bit.
template <class WA, class KA, class CA>
static int check_head_on(const std::vector<double> &w0, WA w, KA K, CA c, int T, double jitter, long long &fails, long long &irr_max);
static int check_head(const std::vector<double> &w, int T, double jitter, long long &fails, long long &irr_max) {
    const int xs = (int)w.size();
    if (T == 512) {  // the device's layout: padded views
        const int pn = seqsum::padded_size(xs + 1);
        std::vector<double> wp(pn);
        std::vector<long long> Kp(pn);
        std::vector<unsigned short> cp(pn);
        seqsum::Padded<double> wv{wp.data()};
        for (int i = 0; i < xs; i++) wv[i] = w[i];
        return check_head_on(w, wv, seqsum::Padded<long long>{Kp.data()}, seqsum::Padded<unsigned short>{cp.data()}, T, jitter, fails, irr_max);
    }
    std::vector<double> wc(w);
    wc.push_back(0.);
    std::vector<long long> K(xs + 1);
    std::vector<unsigned short> c(xs + 1);
    return check_head_on(w, wc.data(), K.data(), c.data(), T, jitter, fails, irr_max);
}
template <class WA, class KA, class CA>
static int check_head_on(const std::vector<double> &w0, WA w, KA K, CA c, int T, double jitter, long long &fails, long long &irr_max) {
    const int xs = (int)w0.size();
    int P = 1;
    while (P < xs) P <<= 1;
    const int E = std::max(1, P / T);
    std::vector<int> irr_idx(seqsum::kMaxIrr), irr_e(seqsum::kMaxIrr);
    std::vector<double> base(seqsum::kMaxIrr), pex(T + 1), endP(T + 1);
    std::vector<seqsum::ChunkAgg> agg(T);
    double run = 0.;
    for (int t = 0; t < T; t++) {
        const int lo = std::min(xs, t * E), hi = std::min(xs, lo + E);
        pex[t] = run * (1.0 + jitter * ((double)(rng() % 2001) - 1000.0) / 1000.0);
        run = run + seqsum::head_chunk_sum(w, lo, hi, 0.);
        endP[t] = seqsum::head_chunk_sum(w, lo, hi, pex[t]);
    }
    for (int t = 0; t < T; t++) {
        const int lo = std::min(xs, t * E), hi = std::min(xs, lo + E);
        agg[t] = seqsum::head_chunk_classify(w, lo, hi, pex[t], t > 0 ? endP[t - 1] : 0., K, c);
    }
    seqsum::ChunkAgg acc;
    acc.has_irr = 0; acc.tail = 0; acc.n_irr = 0;
    for (int t = 0; t < T; t++) {
        const int lo = std::min(xs, t * E), hi = std::min(xs, lo + E);
        seqsum::head_chunk_finish(w, lo, hi, pex[t], acc.tail, acc.n_irr, K, c, irr_idx.data(), irr_e.data());
        acc = seqsum::agg_combine(acc, agg[t]);
    }
    const int M = acc.n_irr;
    if (M > irr_max) irr_max = M;
    if (!seqsum::head_stitch(w, K, irr_idx.data(), irr_e.data(), M, xs, base.data())) { fails++; return 0; }
    double s = 0.;
    for (int i = 0; i < xs; i++) {
        s = s + w0[i];
        const double got = seqsum::head_value(i, K, c, irr_idx.data(), irr_e.data(), base.data());
        if (bits(got) != bits(s)) {
            std::printf("FAIL: head entry %d of %d: %a, sequential %a (T=%d jitter=%g M=%d)\n", i, xs, got, s, T, jitter, M);
            return -1;
        }
    }
    return xs;
}

static int head_cases(int cases, long long &entries) {
    long long fails = 0, fails_jit = 0, irr_max = 0, n_jit = 0;
    for (int cidx = 0; cidx < cases; cidx++) {
        const int xs = (cidx % 11 == 0) ? (int)(rng() % 40) : (int)(rng() % 8193);
        std::vector<double> w(xs);
        const double W = random_weight();
        const int kind = (int)(rng() % 5);
        for (int i = 0; i < xs; i++) {
            switch (kind) {
                case 0: w[i] = W * ((double)(rng() % 1000000 + 1) / 1000001.0); break;              // full mantissas below W
                case 1: w[i] = W / (double)(2 + rng() % 40); break;                                 // a few classes (dimers, trimers, ...)
                case 2: w[i] = std::ldexp((double)(1 + rng() % 255), -8 - (int)(rng() % 3)) * 3.0; break;  // short mantissas: ties everywhere
                case 3: w[i] = W * std::ldexp((double)(rng() % 1000 + 1) / 1001.0, -(int)(rng() % 40)); break;  // many magnitudes
                default: w[i] = random_weight(); break;
            }
        }
        std::sort(w.begin(), w.end());
        const int T = (cidx & 1) ? 512 : 64;
        int r = check_head(w, T, 0., fails, irr_max);
        if (r < 0) return 1;
        entries += r;
        const double jit = (cidx % 3 == 0) ? 1e-13 : 1e-9;  // (1e-9: far rougher than any summation order: more stretches are refused, none is wrong)
        r = check_head(w, T, jit, fails_jit, irr_max);
        if (r < 0) return 1;
        n_jit++;
    }
    std::printf("head: %d cases, refused %lld (plain) / %lld of %lld (jittered), at most %lld irregular steps\n", cases, fails, fails_jit, n_jit, irr_max);
    return 0;
}

// seqsum::bitonic_pass (the building CTA's sort): the passes in the order k_event issues them, every thread of a 512-thread CTA one
// after the other between two barriers, on the padded view; must leave the array sorted with its +inf padding behind.
static int sort_cases(int cases) {
    for (int cidx = 0; cidx < cases; cidx++) {
        int P = 64 << (cidx % 8);  // 64 .. 8192
        const int xs = (int)(rng() % (P + 1)), nthr = 512;
        std::vector<double> plain(P, INFINITY), pad(seqsum::padded_size(P));
        for (int i = 0; i < xs; i++) plain[i] = (cidx % 3 == 0) ? (double)(rng() % 50) : random_weight();
        seqsum::Padded<double> v{pad.data()};
        for (int i = 0; i < P; i++) v[i] = plain[i];
        for (int k = 2, lg = 1; k <= P; k <<= 1, lg++)
            for (int top = lg - 1; top >= 0;) {
                const int g = std::min(3, top + 1), bb = top - g + 1;
                for (int tid = 0; tid < nthr; tid++) {
                    if (g == 3) seqsum::bitonic_pass<3>(v, P, k, bb, tid, nthr);
                    else if (g == 2) seqsum::bitonic_pass<2>(v, P, k, bb, tid, nthr);
                    else seqsum::bitonic_pass<1>(v, P, k, bb, tid, nthr);
                }
                top -= g;
            }
        std::sort(plain.begin(), plain.end());
        for (int i = 0; i < P; i++)
            if (bits(v[i]) != bits(plain[i])) { std::printf("FAIL: bitonic passes, P = %d, entry %d\n", P, i); return 1; }
    }
    std::printf("sort: %d tables of 64 .. 8192 entries\n", cases);
    return 0;
}

int main(int argc, char **argv) {
    const int cases = argc > 1 ? std::atoi(argv[1]) : 200;
    rng.seed(argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 1);
    long long entries = 0, segs_max = 0;
    for (int c = 0; c < cases; c++) {
        const double v = random_weight();
        double x0 = 0.;
        switch (rng() % 5) {
            case 0: x0 = 0.; break;
            case 1: x0 = random_weight(); break;
            case 2: x0 = v * (double)(rng() % 100000) * 0.37; break;                       // inside the run's own range
            case 3: x0 = std::ldexp(random_weight(), (int)(rng() % 50)); break;          // far above: the sum barely moves or stagnates
            default: x0 = std::ldexp(random_weight(), -(int)(rng() % 30)); break;        // far below
        }
        const int cnt = (c % 7 == 0) ? 1 + (int)(rng() % 3000000) : 1 + (int)(rng() % 200000);
        const long long r = check_run(x0, (int)(rng() % 10000), cnt, v, segs_max);
        if (r < 0) return 1;
        entries += r;
    }
    // every short-mantissa weight against every short-mantissa start: all the round-to-even situations of the first binades, and
    // (scaled starts) of binades far above the weight, where most steps are ties or stagnate
    for (int m = 1; m <= 96; m++)
        for (int j = 0; j <= 96; j++)
            for (int sh = 0; sh <= 54; sh += 6) {
                const long long r = check_run(std::ldexp((double)j, sh - 5), 7, 700, std::ldexp((double)m, -5), segs_max);
                if (r < 0) return 1;
                entries += r;
            }
    // the shape the pick table has: a head of lighter weights summed one by one, then ~10^6 copies of the largest weight
    for (int c = 0; c < 8; c++) {
        const double W = random_weight();
        double acc = 0.;
        const int head = (int)(rng() % 8192);
        for (int i = 0; i < head; i++) acc = acc + W * ((double)(rng() % 1000000 + 1) / 1000001.0);
        const long long r = check_run(acc, head, 1000000 + (int)(rng() % 100000), W, segs_max);
        if (r < 0) return 1;
        entries += r;
    }
    std::printf("ok %d cases, %lld entries, at most %lld segments per run\n", cases, entries, segs_max);
    if (head_cases(cases, entries)) return 1;
    if (sort_cases(std::min(cases, 160))) return 1;
    return 0;
}
