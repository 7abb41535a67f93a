// mcac_b200 — K9 fast path: AggregatList::sort_time_steps (aggregat_list.cpp:109-141) for TIE-DOMINATED weight tables.
//
// In monodisperse runs (C1, C3, C5) almost every aggregate is a monomer, all monomers share one 1/dt weight W, and W is the
// LARGEST weight of the table (a bigger aggregate has a longer time step).  libstdc++'s introsort on such a table is a chain:
// at every level the median-of-3 pivot is a W element, the Hoare partition sends the few lighter ("sparse") elements to the
// left part and leaves a right part made of W elements only, whose further fate does not depend on the data (closed form,
// `all_equal_final`).  So the top levels of the replayed sort need no pass over the table at all:
//   * `plan_build` (ONE CTA, shared memory): simulates the top levels on the sparse elements only — their sorted positions
//     (kept sorted in closed form, nothing is re-sorted), the number of Hoare swaps K and the partition point of each level —
//     and leaves per-level rank tables in global memory;
//   * `dense_route` (every thread of the grid, one W element each): follows an element through those levels with O(1) rank
//     queries per level until it falls into an all-W right part (final position in closed form) or reaches the segment that
//     is handed to the general level-synchronous sort (`k_event`), which finishes the small mixed remainder.
// The result is the same permutation std::sort produces (tests/native/tie_sort_host.cpp checks this header against
// libstdc++'s own __introsort_loop on the host; the GPU parity tests check it against the oracle).
// The code is host/device neutral: a `Team` supplies tid / nthr / sync / team_min (serial on the host, one CTA on the device).
#pragma once
#ifdef __CUDACC__
#define TS_HD __host__ __device__ __forceinline__
// (rank queries out of line — __noinline__ — were tried to shrink the level loop of the sparse simulation: 104 us instead of 92 us
// per sort at ~3900 sparse elements, profiles/r2_tuning.md; they stay inlined)
#define TS_HD_CALL __host__ __device__ __forceinline__
#else
#define TS_HD inline
#define TS_HD_CALL inline
#endif

namespace tiesort {

constexpr int kLeaf = 16;            // libstdc++ _S_threshold
constexpr int kMaxLevels = 14;       // top levels handled on the sparse elements
constexpr int kBuckets = 8192;       // buckets of a rank table
constexpr int kTblStride = kBuckets + 4;   // (rows stay 16-byte aligned)
constexpr int kMaxSparse = 8192;     // most sparse elements the one-CTA simulation takes
constexpr int kIntMax = 0x7fffffff;
constexpr int kMiscInts = 64;        // plan_build's s_misc: 16 entries of level state + 48 of scan scratch

struct Level {
    int f, l;    // segment [f, l) partitioned at this level
    int pick;    // position __move_median_to_first swaps with f (it holds a W element)
    int K;       // Hoare swaps of __unguarded_partition
    int cut;     // its return value: children [f, cut) (mixed: next level) and [cut, l) (W only)
    int depth;   // depth_limit after this level's decrement
    int shift;   // bucket shift of this level's rank table (position half)
    int shiftD;  // bucket shift of its dense-coordinate half
};
struct Plan {
    int n_levels;                     // levels simulated sparsely
    int hand_f, hand_l, hand_depth;   // segment handed to the general sort, before its pivot move; depth_limit it starts with
    int fail;                         // introsort's depth limit hit (heap-sort branch): caller falls back
    int x;                            // sparse elements
    int overlap;                      // set by the caller: the simulating CTA fills and sorts the handed-over segment itself
    int nb;                           // buckets of a rank table (power of two >= x, 256 .. kBuckets)
    int ready;                        // levels published so far (tables + lv[] visible device-wide): the routing pass follows it level by level
    int done;                         // set by the caller after the last level and the fields above are final
    int cum_gathered;                 // caller's: the CTA that builds the exact cumulative table has read the staged sparse weights
    double W;
    Level lv[kMaxLevels];
};

// entries of the sorted list R with position < q; tbl[b] = entries whose bucket ((pos - f) >> shift) is below b
// (a bucket holds about one entry: the first two are fetched together instead of one after the other — the lists are readable two
// entries past their end)
template <class PR, class PT>
TS_HD_CALL int rank_lt(PR R, PT tbl, int f, int shift, int q) {
    const int b = (q - f) >> shift;
    int j = tbl[b];
    const int e = tbl[b + 1];
    const int r0 = R[j], r1 = R[j + 1];
    if (j < e && r0 < q) {
        j++;
        if (j < e && r1 < q) {
            j++;
            while (j < e && R[j] < q) j++;
        }
    }
    return j;
}
// The simulating team keeps ONE packed table: low 16 bits = the position table above, high 16 bits = the same kind of table over
// d_i = R[i] - (f + 1) - i (the W positions of [f+1, l) before entry i; non-decreasing), bucket d_i >> shiftD.
template <class PR, class PT>
TS_HD_CALL int rank_lt_pk(PR R, PT pk, int f, int shift, int q) {
    const int b = (q - f) >> shift;
    int j = pk[b] & 0xffff;
    const int e = pk[b + 1] & 0xffff;
    const int r0 = R[j], r1 = R[j + 1];
    if (j < e && r0 < q) {
        j++;
        if (j < e && r1 < q) {
            j++;
            while (j < e && R[j] < q) j++;
        }
    }
    return j;
}
// entries with d_i < v
template <class PR, class PT>
TS_HD_CALL int rank_dense_lt_pk(PR R, PT pk, int f, int shiftD, int v) {
    const int b = v >> shiftD;
    int j = pk[b] >> 16;
    const int e = pk[b + 1] >> 16;
    const int r0 = R[j], r1 = R[j + 1];
    const int t = v + f + 1;  // d_i < v  <=>  R[i] - i < v + f + 1
    if (j < e && r0 - j < t) {
        j++;
        if (j < e && r1 - j < t) {
            j++;
            while (j < e && R[j] - j < t) j++;
        }
    }
    return j;
}
// k-th (0-based) position of [f+1, l) that holds no sparse element, in one look-up: the sparse entries before it are those with d_i <= k
template <class PR, class PT>
TS_HD_CALL int select_dense_direct(PR R, PT pk, int f, int shiftD, int k) {
    return f + 1 + k + rank_dense_lt_pk(R, pk, f, shiftD, k + 1);
}
// k-th (0-based) position of [f+1, l) that holds no sparse element
template <class PR, class PT>
TS_HD_CALL int select_dense(PR R, PT tbl, int f, int shift, int k) {
    int q = f + 1 + k;
    for (int it = 0; it < kMaxSparse + 2; it++) {  // monotone fixed point; q grows by at least one sparse element per round
        const int q2 = f + 1 + k + rank_lt(R, tbl, f, shift, q + 1);
        if (q2 == q) break;
        q = q2;
    }
    return q;
}

// Final position of the element at `pos` of a segment [cf, cl) whose keys are all equal; the segment was created by a level whose
// depth_limit (after its decrement) was dep_parent.  Equal keys: median-of-3 picks `mid`, the Hoare partition mirrors
// [cf+1, cl-1], the cut falls at cf+1+(m-1)/2, leaves do not move.
TS_HD int all_equal_final(int cf, int cl, int pos, int dep_parent, bool &bad) {
    if (cl - cf <= kLeaf) return pos;
    int bits = 0;  // m < 2^bits: more levels than the halving below can take
    while (((unsigned)(cl - cf) >> bits) != 0) bits++;
    if (dep_parent >= bits) {  // the depth limit cannot be reached: segment-relative form without the bookkeeping
        unsigned m = (unsigned)(cl - cf), r = (unsigned)(pos - cf);
        int base = cf;
        while (m > (unsigned)kLeaf) {
            const unsigned mid = m >> 1;
            if (r == 0) r = mid;
            else if (r == mid) r = 0;
            if (r) r = m - r;
            const unsigned c = 1 + ((m - 1) >> 1);
            if (r < c) m = c;
            else { r -= c; base += (int)c; m -= c; }
        }
        return base + (int)r;
    }
    int dep = dep_parent;
    if (dep == 0) { bad = true; return pos; }
    dep--;
    while (cl - cf > kLeaf) {
        const int m = cl - cf, mid = cf + m / 2;
        if (pos == cf) pos = mid;
        else if (pos == mid) pos = cf;
        if (pos > cf) pos = cf + cl - pos;
        const int cutp = cf + 1 + (m - 1) / 2;
        if (pos < cutp) cl = cutp;
        else cf = cutp;
        if (cl - cf > kLeaf) {
            if (dep == 0) { bad = true; return pos; }
            dep--;
        }
    }
    return pos;
}

// A W element that starts at `pos`: its final position (handed == false) or its position inside the handed-over segment.
template <class PR, class PT>
TS_HD int dense_route(const Plan &P, PR R, PT tbl, int r_stride, int t_stride, int pos, bool &handed, bool &bad) {
    handed = false;
    for (int t = 0; t < P.n_levels; t++) {
        const Level L = P.lv[t];
        PR Rt = R + (size_t)t * r_stride;
        PT Tt = tbl + (size_t)t * t_stride;
        if (pos == L.f) pos = L.pick;
        else if (pos == L.pick) pos = L.f;
        if (pos > L.f) {
            const int ka = pos - (L.f + 1) - rank_lt_pk(Rt, Tt, L.f, L.shift, pos);  // rank among the left stoppers (W elements)
            if (ka < L.K) pos = L.l - 1 - ka;                                          // every position is a right stopper
            else {
                const int kb = L.l - 1 - pos;
                if (kb < L.K) pos = select_dense_direct(Rt, Tt, L.f, L.shiftD, kb);
            }
        }
        if (pos >= L.cut) return all_equal_final(L.cut, L.l, pos, L.depth, bad);
    }
    handed = true;
    return pos;
}

// The W element found at position `pos` of the handed-over segment after the simulated levels: where it started (its label).
// Every level is an involution on positions (pivot move, Hoare swaps), and a W element of the left part was a W position before
// the partition as well (sparse elements never leave the left part), so the same look-ups run backwards.
template <class PR, class PT>
TS_HD int dense_origin(const Plan &P, PR R, PT tbl, int r_stride, int t_stride, int pos) {
    for (int t = P.n_levels - 1; t >= 0; t--) {
        const Level L = P.lv[t];
        if (pos == L.f) { pos = L.pick; continue; }  // the pivot came from `pick`
        const int ka = pos - (L.f + 1) - rank_lt(R + (size_t)t * r_stride, tbl + (size_t)t * t_stride, L.f, L.shift, pos);
        if (ka < L.K) pos = L.l - 1 - ka;            // swapped in from the right stopper B[ka]
        if (pos == L.pick) pos = L.f;                // the element the pivot move had put there
    }
    return pos;
}

// Bucket tables of the ascending list S (x entries, positions in [f, l]), packed (see rank_lt_pk): a histogram of the entries'
// buckets (shared-memory adds, one entry apart: about one entry per bucket) followed by one scan over the <= nb + 1 buckets,
// each thread a contiguous chunk.  The scan also writes the packed table to the global copy Tg the routing pass reads and its
// position half to the caller's archive Ta (may be null); the entry pass copies the list to Rg / Ra and finds K, the number of Hoare swaps: the first k
// with not (k < n_a and A[k] < B[k]), where A[k] is the k-th W position and B[k] = l-1-k  <=>  2k + #{sparse before A[k]} >= M-1;
// entry j owns the k with exactly j sparse elements before A[k].  Three team barriers, the last one at the end.
TS_HD int bits_of(unsigned v) {  // number of significant bits (0 for 0)
#ifdef __CUDA_ARCH__
    return 32 - __clz((int)v);
#else
    return v ? 32 - __builtin_clz(v) : 0;
#endif
}
template <class Team>
TS_HD void build_table_and_k(Team &tm, int x, int nb, const int *__restrict__ S, int f, int l, int *__restrict__ s_pk, int *__restrict__ Rg,
                             int *__restrict__ Tg, int *__restrict__ Ra, int *__restrict__ Ta, int *k_min, int *scratch, int &shift_out, int &shiftD_out) {
    // smallest shifts with ((l - f) >> shift) <= nb - 1 and (n_a >> shiftD) <= nb - 1 (nb is a power of two)
    const int M = l - f - 1, n_a = M - x;
    const int nb_bits = bits_of((unsigned)nb) - 1;
    const int shift = bits_of((unsigned)(l - f)) > nb_bits ? bits_of((unsigned)(l - f)) - nb_bits : 0;
    const int shiftD = bits_of((unsigned)n_a) > nb_bits ? bits_of((unsigned)n_a) - nb_bits : 0;
    const int nbk = ((l - f) >> shift) + 1, nbkD = (n_a >> shiftD) + 1;
    const int ntab = (nbk > nbkD ? nbk : nbkD) + 1;  // entries 0 .. ntab-1 of the packed table are used (written in groups of four)
    const int ngrp = (ntab + 3) >> 2;
    shift_out = shift;
    shiftD_out = shiftD;
#ifdef __CUDA_ARCH__
    for (int g = tm.tid; g < ngrp; g += tm.nthr) reinterpret_cast<int4 *>(s_pk)[g] = make_int4(0, 0, 0, 0);
#else
    for (int b = tm.tid; b < 4 * ngrp; b += tm.nthr) s_pk[b] = 0;
#endif
    tm.sync();
    int kbest = kIntMax;
    const int need0 = M - 1;
#pragma unroll 4
    for (int j = tm.tid; j < x; j += tm.nthr) {
        const int sp = j ? S[j - 1] : f, sc = S[j];  // (sp = f makes lo = 0 for the first entry)
        const int dj = sc - (f + 1) - j;
        tm.add(&s_pk[((sc - f) >> shift) + 1], 1);
        tm.add(&s_pk[(dj >> shiftD) + 1], 1 << 16);
        Rg[j] = sc;
        if (Ra) Ra[j] = sc;
        const int lo = sp - f - j;  // = sp - (f + 1) - (j - 1)
        const int kmin = (need0 - j + 1) >> 1;
        int k = lo > kmin ? lo : kmin;
        k = k > 0 ? k : 0;
        if (k < dj && k < kbest) kbest = k;
    }
    if (tm.tid == 0) {  // the k behind the last sparse entry
        const int lo = x ? S[x - 1] - f - x : 0;
        const int kmin = (need0 - x + 1) >> 1;
        int k = lo > kmin ? lo : kmin;
        k = k > 0 ? k : 0;
        if (k < n_a && k < kbest) kbest = k;
    }
    tm.team_min(k_min, kbest);  // (one shared-memory atomic per warp, not one per entry: about half of the entries are candidates)
    tm.sync();
    // scan: each thread a contiguous chunk of groups; an odd number of groups per chunk keeps the 16-byte accesses of neighbouring
    // threads in different banks
    const int gchunk = ((ngrp + tm.nthr - 1) / tm.nthr) | 1;
    const int g0 = tm.tid * gchunk < ngrp ? tm.tid * gchunk : ngrp, g1 = g0 + gchunk < ngrp ? g0 + gchunk : ngrp;
    int sum = 0;
#ifdef __CUDA_ARCH__
    for (int g = g0; g < g1; g++) {
        const int4 v = reinterpret_cast<const int4 *>(s_pk)[g];
        sum += (v.x + v.y) + (v.z + v.w);
    }
#else
    for (int b = 4 * g0; b < 4 * g1; b++) sum += s_pk[b];
#endif
    int run = tm.excl_scan(sum, scratch);
    for (int g = g0; g < g1; g++) {
#ifdef __CUDA_ARCH__
        int4 v = reinterpret_cast<const int4 *>(s_pk)[g];
        v.x += run; v.y += v.x; v.z += v.y; v.w += v.z;
        run = v.w;
        reinterpret_cast<int4 *>(s_pk)[g] = v;
        reinterpret_cast<int4 *>(Tg)[g] = v;
        if (Ta && 4 * g <= nbk) { Ta[4 * g] = v.x & 0xffff; Ta[4 * g + 1] = v.y & 0xffff; Ta[4 * g + 2] = v.z & 0xffff; Ta[4 * g + 3] = v.w & 0xffff; }
#else
        for (int b = 4 * g; b < 4 * g + 4; b++) {
            run += s_pk[b];
            s_pk[b] = run;
            Tg[b] = run;
            if (Ta && 4 * g <= nbk) Ta[b] = run & 0xffff;
        }
#endif
    }
    tm.sync();
}

// The sparse simulation.  st_pos (ascending) / st_w: label and weight of the x elements whose weight is below W; n: table size;
// depth0 = 2*floor(log2(n)).  Scratch of the team: four lists of x ints (a_s, a_i, b_s, b_i), s_tbl (nb + 3), s_misc (16).
// On return plan, R[t*xcap ..], tbl[t*kTblStride ..] describe the levels t < n_levels, and (a_s[j], a_i[j]) are the position and
// the index into st_pos / st_w of the sparse elements inside the handed-over segment (ascending positions).
// Per level: pivot samples -> pivot move (one list entry changes place; only when the element at f is sparse) -> bucket table
// for rank queries + K (parallel minimum), one pass -> every sparse element computes its new position AND its new rank in closed
// form from rank queries on the current table (the right-hand elements move to the K first W positions in reverse order, the
// left-hand ones stay), so the list stays sorted without sorting.  Two team barriers per level (three with a pivot move).
// (s_tbl: the packed bucket table, nb + 4 entries, 16-byte aligned; s_misc: kMiscInts entries; the four lists must be readable two entries past x)
template <class Team>
TS_HD void plan_build(Team &tm, int n, int x, const int *st_pos, const double *st_w, double W, int depth0, int hand_min, Plan *plan,
                      int *R, int *tbl, int xcap, int *a_s, int *a_i, int *b_s, int *b_i, int *s_tbl, int *s_misc,
                      int *arch_R = nullptr, int *arch_T = nullptr, int arch_levels = 0) {
    int nb = 256;
    while (nb < x && nb < kBuckets) nb <<= 1;
    int f = 0, l = n, depth = depth0, t = 0;
    int *cs = a_s, *ci = a_i, *os = b_s, *oi = b_i;  // current list (positions, ids) and the other buffer
    // s_misc: two sets of {sparse element at the pivot samples pa, pb, pc and at f; running minimum for K}, used alternately by levels
    for (int k = tm.tid; k < 4; k += tm.nthr) s_misc[k] = -1;
    if (tm.tid == 0) s_misc[4] = kIntMax;
    tm.sync();
    {
        const int pa = f + 1, pb = f + (l - f) / 2, pc = l - 1;
        for (int j = tm.tid; j < x; j += tm.nthr) {
            const int q = st_pos[j];
            cs[j] = q;
            ci[j] = j;
            if (q == pa) s_misc[0] = j;
            if (q == pb) s_misc[1] = j;
            if (q == pc) s_misc[2] = j;
            if (q == f) s_misc[3] = j;
        }
    }
    tm.sync();
    for (;;) {
        const int len = l - f;
        if (len <= hand_min || len <= kLeaf || t == kMaxLevels || depth == 0) break;
        int *mc = s_misc + 8 * (t & 1), *mn = s_misc + 8 * ((t + 1) & 1);  // this level's set, the next level's set
        const int pa = f + 1, pb = f + len / 2, pc = l - 1;
        // __move_median_to_first(first, first+1, mid, last-1) with the plain `<` of sort_indexes
        const double ka = mc[0] < 0 ? W : st_w[ci[mc[0]]], kb = mc[1] < 0 ? W : st_w[ci[mc[1]]], kc = mc[2] < 0 ? W : st_w[ci[mc[2]]];
        int pick;
        double kp;
        if (ka < kb) {
            if (kb < kc) { pick = pb; kp = kb; }
            else if (ka < kc) { pick = pc; kp = kc; }
            else { pick = pa; kp = ka; }
        } else if (ka < kc) { pick = pa; kp = ka; }
        else if (kb < kc) { pick = pc; kp = kc; }
        else { pick = pb; kp = kb; }
        const bool at_f = mc[3] >= 0;  // a sparse element at f is cs[0]
        if (kp != W) break;            // a sparse pivot: the rest goes to the general sort
        tm.lap(0);
        depth--;
        for (int k = tm.tid; k < 4; k += tm.nthr) mn[k] = -1;  // (last read before the barrier that ended the previous level)
        if (tm.tid == 0) mn[4] = kIntMax;
        if (at_f) {  // pivot move: the element at f goes to `pick`  (current list -> other buffer, which becomes the current one)
            for (int j = tm.tid; j < x; j += tm.nthr) {
                const int q = cs[j];
                if (j > 0) {
                    const int r = (j - 1) + (q > pick ? 1 : 0);
                    os[r] = q;
                    oi[r] = ci[j];
                    if (q < pick && (j == x - 1 || cs[j + 1] > pick)) { os[j] = pick; oi[j] = ci[0]; }
                } else if (x == 1 || cs[1] > pick) { os[0] = pick; oi[0] = ci[0]; }
            }
            int *ts_ = cs; cs = os; os = ts_;
            int *ti_ = ci; ci = oi; oi = ti_;
            tm.sync();
        }
        tm.lap(1);
        const int M = len - 1, n_a = M - x;
        // (arch_*: a second copy of the tables of the first arch_levels levels, strides x and nb + 3, kept by the caller)
        const bool ar = arch_R && t < arch_levels;
        int shift, shiftD;
        build_table_and_k(tm, x, nb, cs, f, l, s_tbl, R + (size_t)t * xcap, tbl + (size_t)t * kTblStride, ar ? arch_R + (size_t)t * x : nullptr,
                          ar ? arch_T + (size_t)t * (nb + 3) : nullptr, &mc[4], s_misc + 16, shift, shiftD);
        tm.lap(2);
        tm.lap(3);
        const int K = mc[4] < n_a ? mc[4] : n_a;
        const int aK = K < n_a ? select_dense_direct(cs, s_tbl, f, shiftD, K) : kIntMax;
        const int bK = K > 0 ? l - K : l;
        const int cut = aK < bK ? aK : bK;
        // moves (current list -> other buffer, still ascending) + the pivot samples of the next level [f, cut)
        const int npa = f + 1, npb = f + (cut - f) / 2, npc = cut - 1;
        tm.lap(4);
        for (int j = tm.tid; j < x; j += tm.nthr) {
            const int q = cs[j], kb2 = l - 1 - q;
            int np, nr;
            if (kb2 < K) {  // right stopper of a swap: goes to the kb2-th W position; the movers end up in reverse order
                const int before = rank_dense_lt_pk(cs, s_tbl, f, shiftD, kb2 + 1);  // sparse entries before that position
                np = f + 1 + kb2 + before;
                nr = before + (x - 1 - j);
            } else {        // stays; the movers that land before it: those with kb < c = W positions before q
                const int c = q - (f + 1) - j;
                const int first_mover = (l - c > l - K) ? l - c : l - K;
                np = q;
                nr = j + (x - rank_lt_pk(cs, s_tbl, f, shift, first_mover));
            }
            os[nr] = np;
            oi[nr] = ci[j];
            if (np == npa) mn[0] = nr;
            if (np == npb) mn[1] = nr;
            if (np == npc) mn[2] = nr;
            if (np == f) mn[3] = nr;
        }
        tm.lap(5);
        if (tm.tid == 0) {
            Level L;
            L.f = f; L.l = l; L.pick = pick; L.K = K; L.cut = cut; L.depth = depth; L.shift = shift; L.shiftD = shiftD;
            plan->lv[t] = L;
        }
        { int *ts_ = cs; cs = os; os = ts_; }
        { int *ti_ = ci; ci = oi; oi = ti_; }
        l = cut;
        t++;
        tm.sync();
        tm.publish(&plan->ready, t);  // level t - 1 (its tables, its Level record) can be used by other teams
        tm.lap(6);
    }
    if (cs != a_s) {  // the caller reads the final list from (a_s, a_i)
        tm.sync();
        for (int j = tm.tid; j < x; j += tm.nthr) { a_s[j] = cs[j]; a_i[j] = ci[j]; }
    }
    if (tm.tid == 0) {
        plan->n_levels = t;
        plan->hand_f = f;
        plan->hand_l = l;
        plan->hand_depth = depth;
        plan->fail = (depth == 0 && l - f > kLeaf) ? 1 : 0;
        plan->x = x;
        plan->overlap = 0;
        plan->nb = nb;
        plan->W = W;
    }
    tm.sync();
}

struct SerialTeam {
    int tid = 0, nthr = 1;
    void sync() {}
    void team_min(int *p, int v) { if (v < *p) *p = v; }
    void lap(int) {}
    void publish(int *p, int v) { *p = v; }
    void add(int *p, int v) { *p += v; }
    int excl_scan(int, int *) { return 0; }
};

}  // namespace tiesort
