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
constexpr int kTblStride = kBuckets + 3;
constexpr int kMaxSparse = 8192;     // most sparse elements the one-CTA simulation takes
constexpr int kIntMax = 0x7fffffff;

struct Level {
    int f, l;    // segment [f, l) partitioned at this level
    int pick;    // position __move_median_to_first swaps with f (it holds a W element)
    int K;       // Hoare swaps of __unguarded_partition
    int cut;     // its return value: children [f, cut) (mixed: next level) and [cut, l) (W only)
    int depth;   // depth_limit after this level's decrement
    int shift;   // bucket shift of this level's rank table
    int pad;
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
    double W;
    Level lv[kMaxLevels];
};

// entries of the sorted list R with position < q; tbl[b] = entries whose bucket ((pos - f) >> shift) is below b
template <class PR, class PT>
TS_HD_CALL int rank_lt(PR R, PT tbl, int f, int shift, int q) {
    const int b = (q - f) >> shift;
    int j = tbl[b];
    const int e = tbl[b + 1];
    while (j < e && R[j] < q) j++;
    return j;
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
            const int ka = pos - (L.f + 1) - rank_lt(Rt, Tt, L.f, L.shift, pos);  // rank among the left stoppers (W elements)
            if (ka < L.K) pos = L.l - 1 - ka;                                       // every position is a right stopper
            else {
                const int kb = L.l - 1 - pos;
                if (kb < L.K) pos = select_dense(Rt, Tt, L.f, L.shift, kb);
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

// Bucket table of the ascending list S (x entries, positions in [f, l]): tbl[b] = entries whose bucket ((pos - f) >> shift) is
// below b, for b = 0 .. nbk.  Written by "boundary marking" (entry j fills the buckets between its predecessor's and its own),
// together with the global copies (Rg, Tg) the routing pass reads.
// Bucket table of the ascending list S (x entries, positions in [f, l]): tbl[b] = entries whose bucket ((pos - f) >> shift) is
// below b, for b = 0 .. nbk.  Written by "boundary marking" (entry j fills the buckets between its predecessor's and its own),
// together with the global copies (Rg, Tg) the routing pass reads and the caller's archive (Ra, Ta; may be null).  The same pass
// finds K, the number of Hoare swaps: the first k with not (k < n_a and A[k] < B[k]), where A[k] is the k-th W position and
// B[k] = l-1-k  <=>  2k + #{sparse before A[k]} >= M-1; entry j owns the k with exactly j sparse elements before A[k].
template <class Team>
TS_HD int build_table_and_k(Team &tm, int x, int nb, const int *S, int f, int l, int *s_tbl, int *Rg, int *Tg, int *Ra, int *Ta, int *k_min) {
    int shift = 0;
    while (((l - f) >> shift) > nb - 1) shift++;
    const int nbk = ((l - f) >> shift) + 1;
    const int M = l - f - 1, n_a = M - x;
    int kbest = kIntMax;
    for (int j = tm.tid; j <= x; j += tm.nthr) {
        const int sp = j == 0 ? 0 : S[j - 1], sc = j == x ? 0 : S[j];
        const int bprev = j == 0 ? -1 : (sp - f) >> shift;
        const int bcur = j == x ? nbk : (sc - f) >> shift;
        for (int b = bprev + 1; b <= bcur; b++) {
            s_tbl[b] = j;
            Tg[b] = j;
            if (Ta) Ta[b] = j;
        }
        if (j < x) {
            Rg[j] = sc;
            if (Ra) Ra[j] = sc;
        }
        const int lo = j == 0 ? 0 : sp - (f + 1) - (j - 1);
        const int hi = j == x ? n_a : sc - (f + 1) - j;
        const int need = M - 1 - j;
        const int kmin = need <= 0 ? 0 : (need + 1) / 2;
        const int k = lo > kmin ? lo : kmin;
        if (k < hi && k < kbest) kbest = k;
    }
    tm.team_min(k_min, kbest);  // (one shared-memory atomic per warp, not one per entry: about half of the entries are candidates)
    return shift;
}

// The sparse simulation.  st_pos (ascending) / st_w: label and weight of the x elements whose weight is below W; n: table size;
// depth0 = 2*floor(log2(n)).  Scratch of the team: four lists of x ints (a_s, a_i, b_s, b_i), s_tbl (nb + 3), s_misc (16).
// On return plan, R[t*xcap ..], tbl[t*kTblStride ..] describe the levels t < n_levels, and (a_s[j], a_i[j]) are the position and
// the index into st_pos / st_w of the sparse elements inside the handed-over segment (ascending positions).
// Per level: pivot samples -> pivot move (one list entry changes place; only when the element at f is sparse) -> bucket table
// for rank queries + K (parallel minimum), one pass -> every sparse element computes its new position AND its new rank in closed
// form from rank queries on the current table (the right-hand elements move to the K first W positions in reverse order, the
// left-hand ones stay), so the list stays sorted without sorting.  Two team barriers per level (three with a pivot move).
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
        const int shift = build_table_and_k(tm, x, nb, cs, f, l, s_tbl, R + (size_t)t * xcap, tbl + (size_t)t * kTblStride,
                                            ar ? arch_R + (size_t)t * x : nullptr, ar ? arch_T + (size_t)t * (nb + 3) : nullptr, &mc[4]);
        tm.lap(2);
        tm.sync();
        tm.lap(3);
        const int K = mc[4] < n_a ? mc[4] : n_a;
        const int aK = K < n_a ? select_dense(cs, s_tbl, f, shift, K) : kIntMax;
        const int bK = K > 0 ? l - K : l;
        const int cut = aK < bK ? aK : bK;
        // moves (current list -> other buffer, still ascending) + the pivot samples of the next level [f, cut)
        const int npa = f + 1, npb = f + (cut - f) / 2, npc = cut - 1;
        tm.lap(4);
        for (int j = tm.tid; j < x; j += tm.nthr) {
            const int q = cs[j], kb2 = l - 1 - q;
            int np, nr;
            if (kb2 < K) {  // right stopper of a swap: goes to the kb2-th W position; the movers end up in reverse order
                np = select_dense(cs, s_tbl, f, shift, kb2);
                nr = (np - (f + 1) - kb2) + (x - 1 - j);
            } else {        // stays; the movers that land before it: those with kb < c = W positions before q
                const int c = q - (f + 1) - j;
                const int first_mover = (l - c > l - K) ? l - c : l - K;
                np = q;
                nr = j + (x - rank_lt(cs, s_tbl, f, shift, first_mover));
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
            L.f = f; L.l = l; L.pick = pick; L.K = K; L.cut = cut; L.depth = depth; L.shift = shift; L.pad = 0;
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
};

}  // namespace tiesort
