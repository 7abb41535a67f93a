// mcac_b200 — the reference's cumulative_time_steps (aggregat_list.cpp:131-140: cum[i] = cum[i-1] + w[sorted[i]], one addition after
// the other) in closed form over a RUN of equal weights: in a monodisperse run every monomer has the same weight W, and W is the largest
// one, so the sorted table is a short head of lighter aggregates followed by ~10^6 copies of W.
//
// Adding the same v over and over: while x and x + v stay inside one binade [2^e, 2^(e+1)) the spacing u of the doubles is constant, x is
// a multiple of u, and fl(x + v) = x + k*u with a k that does not depend on x (round-to-nearest of v on the u grid) — except when v/u
// ends in exactly .5: the tie goes to the even neighbour, so the step depends on the parity of x/u; after one such step x/u is even and
// every later step is the same again.  So inside a binade the sequential sums are x1, x1 + C, x1 + 2C, ... as soon as two consecutive
// steps agree (C = x2 - x1 = x3 - x2, exact differences), every one of them representable, i.e. cum[i0 + j] = fma(j, C, x1) exactly.
// `run_segments` walks the binades of a run with real additions at every irregular step (first steps, binade crossings, parity steps)
// and emits one segment per regular stretch; `segment_value` evaluates an entry.  tests/native/seq_cumsum_host.cpp checks both against
// the plain sequential loop, bit for bit.
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define SC_HD __host__ __device__ __forceinline__
#else
#define SC_HD inline
#endif

namespace seqsum {

struct Seg {
    int i0, cnt;   // entries [i0, i0 + cnt)
    double x0, c;  // cum[i0 + j] = fma(j, c, x0)
};
constexpr int kMaxSegs = 256;  // a run that starts below its own weight crosses ~log2(cnt) binades, a few segments each

SC_HD double segment_value(const Seg &s, int i) { return fma((double)(i - s.i0), s.c, s.x0); }

// the segment that holds entry i (segments are consecutive and ascending; i inside [segs[0].i0, end of the last one))
SC_HD int find_segment(const Seg *segs, int ns, int i) {
    int lo = 0, hi = ns;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (segs[mid].i0 <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// smallest power of two strictly above x (x > 0, finite, normal)
SC_HD double binade_top(double x) {
    int e;
    frexp(x, &e);  // x = m * 2^e, m in [0.5, 1)
    return ldexp(1.0, e);
}

// `cnt` entries of weight v (> 0) at indexes [i0, i0 + cnt), added one after the other to the running sum x.  Appends their segments
// to segs[ns ...] (ns keeps counting past max_segs, nothing is stored there) and returns the running sum behind the run.
SC_HD double run_segments(double x, int i0, int cnt, double v, Seg *segs, int &ns, int max_segs) {
    auto emit = [&](int i, int c_, double x0, double c) {
        if (ns < max_segs) { segs[ns].i0 = i; segs[ns].cnt = c_; segs[ns].x0 = x0; segs[ns].c = c; }
        ns++;
    };
    int i = i0;
    const int end = i0 + cnt;
    while (i < end) {
        const double x1 = x + v;  // one real addition: the first step of a run / of a binade / after a parity step
        const int left = end - (i + 1);  // entries of the run behind this one
        bool regular = left >= 2 && x1 >= v && x1 > 0. && x1 < INFINITY;  // (x1 >= v: the spacing of x1 is at least that of v)
        double top = 0., c = 0.;
        if (regular) {
            top = binade_top(x1);
            const double x2 = x1 + v, x3 = x2 + v;
            c = x2 - x1;  // exact when both lie in one binade
            regular = x3 < top && x3 - x2 == c;
        }
        if (!regular) {
            emit(i, 1, x1, 0.);
            x = x1;
            i++;
            continue;
        }
        // regular steps: x1 + j*c for j = 0 .. m, all below `top`; the estimate is corrected with exact evaluations
        double mf = floor((top - x1) / c);
        if (!(mf <= (double)left)) mf = (double)left;  // (c == 0: v is below half a spacing, the sum no longer moves)
        int m = (int)mf;
        while (m > 0 && !(fma((double)m, c, x1) < top)) m--;
        while (m < left && fma((double)(m + 1), c, x1) < top) m++;
        // entries i .. i + m have values x1 + j*c (j = 0 .. m): the step INTO entry i + m + 1 starts from an x inside the binade
        // but may land outside it, so it is taken by a real addition again
        emit(i, m + 1, x1, c);
        x = fma((double)m, c, x1);
        i += m + 1;
    }
    return x;
}

// ---- the HEAD of the table (a few thousand different weights, ascending) without the chain of dependent additions ------------------
// While the running sum stays in one binade (spacing u), fl(s + w) = s + u * rint(w / u) unless w / u ends in exactly .5 — so the
// sequential sums are INTEGER prefix sums of k_i = rint(w_i / u), which any parallel scan gives exactly.  The steps that are not of this
// kind ("irregular": the first element, a step whose result lands in another binade, a round-to-even tie) are taken by real additions,
// one thread walking from one irregular step to the next: s_r = fl(s + w_r), then s + u * K behind the regular stretch that follows.
// Which binade a step lands in is taken from APPROXIMATE prefix sums P_i (any summation order) and verified on the exact values: every
// stretch must start and end in the binade its elements were classified for, otherwise `head_stitch` reports failure (the caller then
// adds the head up one by one).  A thread owns a chunk [lo, hi) of consecutive elements; `pex` = approximate sum before the chunk.
// (the arrays are any indexable views: plain pointers, or `Padded` ones whose stride between the chunks of neighbouring threads is odd,
// so that the threads of a warp walking their own chunks do not meet in one shared-memory bank)
template <class T>
struct Padded {
    T *p;
    SC_HD T &operator[](int i) const { return p[i + (i >> 4)]; }
};
SC_HD int padded_size(int n) { return n + (n >> 4) + 1; }
constexpr int kMaxIrr = 512;                 // irregular steps a head may have (about one per binade crossed + the rare ties)
constexpr unsigned short kIrrMark = 0xffff;

struct ChunkAgg {
    int has_irr;     // the chunk holds an irregular step
    long long tail;  // sum of k behind its last irregular step (of the whole chunk when it has none)
    int n_irr;
};
// combine(a, b): the aggregate of a's elements followed by b's
SC_HD ChunkAgg agg_combine(const ChunkAgg &a, const ChunkAgg &b) {
    ChunkAgg r;
    r.has_irr = a.has_irr | b.has_irr;
    r.tail = b.has_irr ? b.tail : a.tail + b.tail;
    r.n_irr = a.n_irr + b.n_irr;
    return r;
}
template <class WA>
SC_HD double head_chunk_sum(WA w, int lo, int hi, double pex) {
    double p = pex;
    for (int i = lo; i < hi; i++) p = p + w[i];
    return p;
}
// classification of a chunk; p_before = P of element lo - 1 as ITS chunk computed it (head_chunk_sum of the previous chunk).
// K[i] = sum of k since the last irregular step inside the chunk, c[i] = kIrrMark on irregular steps.
template <class WA, class KA, class CA>
SC_HD ChunkAgg head_chunk_classify(WA w, int lo, int hi, double pex, double p_before, KA K, CA c) {
    ChunkAgg g;
    g.has_irr = 0; g.tail = 0; g.n_irr = 0;
    double p = pex;
    int e_prev = lo > 0 ? ilogb(p_before) : 0;
    for (int i = lo; i < hi; i++) {
        const double pc = p + w[i];
        const int e = ilogb(pc);
        bool irr = i == 0 || e != e_prev;
        long long k = 0;
        if (!irr) {
            const double t = ldexp(w[i], 52 - e);  // w / u, exact
            if (!(t < 9007199254740992.0)) irr = true;
            else {
                const double fl = floor(t);
                if (t - fl == 0.5) irr = true;  // round-to-even tie: depends on the parity of s / u
                else k = (long long)rint(t);
            }
        }
        if (irr) { g.has_irr = 1; g.tail = 0; g.n_irr++; K[i] = 0; c[i] = kIrrMark; }
        else { g.tail += k; K[i] = g.tail; c[i] = 0; }
        p = pc;
        e_prev = e;
    }
    return g;
}
// carry_in / irr_before: tail and number of irregular steps of everything before the chunk.  Leaves K[i] = sum of k since the last
// irregular step, c[i] = index of that step in the list (irr_idx, irr_e = binade its stretch was classified for).
template <class WA, class KA, class CA>
SC_HD void head_chunk_finish(WA w, int lo, int hi, double pex, long long carry_in, int irr_before, KA K, CA c, int *irr_idx, int *irr_e) {
    int m = irr_before - 1;
    bool seen = false;
    double p = pex;
    for (int i = lo; i < hi; i++) {
        p = p + w[i];
        if (c[i] == kIrrMark) {
            m++;
            seen = true;
            if (m < kMaxIrr) { irr_idx[m] = i; irr_e[m] = ilogb(p); }
        } else if (!seen) K[i] += carry_in;
        c[i] = (unsigned short)m;
    }
}
// one thread: the exact sums at the M irregular steps; false = a stretch is not where the approximate sums put it
template <class WA, class KA>
SC_HD bool head_stitch(WA w, KA K, const int *irr_idx, const int *irr_e, int M, int xs, double *base) {
    if (M > kMaxIrr) return false;
    double s = 0.;
    for (int m = 0; m < M; m++) {
        const int r = irr_idx[m];
        const double sr = s + w[r];  // the real addition
        base[m] = sr;
        if (!(sr > 0.) || !(sr < INFINITY)) return false;
        const int e = ilogb(sr);
        if (e != irr_e[m]) return false;
        const int q = (m + 1 < M ? irr_idx[m + 1] : xs) - 1;
        s = q > r ? fma((double)K[q], ldexp(1.0, e - 52), sr) : sr;
        if (ilogb(s) != e) return false;
    }
    return true;
}
template <class KA, class CA>
SC_HD double head_value(int i, KA K, CA c, const int *irr_idx, const int *irr_e, const double *base) {
    const int m = c[i];
    if (irr_idx[m] == i) return base[m];
    return fma((double)K[i], ldexp(1.0, irr_e[m] - 52), base[m]);
}

// ---- sorting the head's weights ------------------------------------------------------------------------------------------
// One pass of a bitonic sorting network over v[0, P) (P a power of two), stage k: the G steps whose partners are 2^(b+G-1) .. 2^b apart.
// The 2^G entries base | (m << b) are closed under those steps: a thread takes them through all G steps in registers — one load and
// one store per entry and pass instead of one per step (the network is bound by shared-memory bandwidth).  Barrier between passes.
template <int G, class VA>
SC_HD void bitonic_pass(const VA &v, int P, int k, int b, int tid, int nthr) {
    constexpr int NG = 1 << G;
    for (int q = tid; q < (P >> G); q += nthr) {
        const int base = ((q >> b) << (b + G)) | (q & ((1 << b) - 1));
        const bool asc = (base & k) == 0;  // (k lies above every bit the group varies)
        double r[NG];
#pragma unroll
        for (int m = 0; m < NG; m++) r[m] = v[base | (m << b)];
#pragma unroll
        for (int st = G - 1; st >= 0; st--) {
#pragma unroll
            for (int m = 0; m < NG; m++) {
                if ((m & (1 << st)) == 0) {
                    const double x = r[m], y = r[m | (1 << st)];
                    const bool sw = (x > y) == asc;
                    r[m] = sw ? y : x;
                    r[m | (1 << st)] = sw ? x : y;
                }
            }
        }
#pragma unroll
        for (int m = 0; m < NG; m++) v[base | (m << b)] = r[m];
    }
}

}  // namespace seqsum
