// mcac_b200 — the general MC step loop of mcac::calcul (src/calcul.cpp:66-281) as ONE persistent CTA per realization.
//
// Small and medium realizations (ensembles of examples/classic.ini: 10^1..10^2 aggregates of up to 10^3 spheres; the early boxes of
// every run) are latency-bound when every step is ~25 launches and 3 host synchronisations.  Here a whole CTA walks the steps of one
// realization by itself: pick table (labels -> 1/dt weights -> replayed introsort -> cumulative sums), pick + direction, Verlet cell
// rebuild, contact search, orientation redraws (interaction potentials), move, surface growth, merge + morphology update,
// nucleation, refresh / PhysicalModel::update — the same device functions the stand-alone kernels of mcac_kernels.cuh wrap, in
// calcul()'s order, separated by block barriers.  Nothing goes back to the host until the realization needs something only the
// host can do (domain duplication, growing its tables, more random draws) or has done the steps it was asked for.
// `k_ensemble_loop` runs many realizations in one launch: CTAs take realizations from a queue (one CTA per realization at a time),
// so a GPU holds as many realizations in flight as it has CTA slots and no host thread sits behind each of them.
#pragma once
#include "mcac_kernels.cuh"

namespace mcacb {

constexpr int kLoopThreads = 512;
constexpr int kLoopDynSmem = 5 * kLoopThreads * (int)sizeof(double);
enum LoopExit { LOOP_STEPS_DONE = 0, LOOP_FINISHED = 1, LOOP_NEED_DUP = 2, LOOP_NEED_REGROW = 3, LOOP_NEED_RNG = 4, LOOP_TOO_BIG = 5,
                LOOP_ERROR = 6, LOOP_EVENT_STOP = 7 };

struct LoopState {  // device-resident, one per handle: what the host reads back after the launch
    int exit_reason;
    int flipped;        // the sphere pool was compacted an odd number of times: d.s_* and alt.s_* have changed places
    int pick_valid, labels_valid;
    long long steps;    // MC steps done by this launch
    long long events, nucleated, sorts, compactions, dups;
    // SM cycles of the CTA per part of the step: 0 pick table (labels + sort), 1 cells + contact search (+ redraws), 2 update block of
    // calcul.cpp:184-206, 3 nucleation + bookkeeping + refresh, 4 loop top (checks, compaction), 5 move + clocks, 6 growth, 7 merge
    long long phase_cycles[8];
};
struct LoopArgs {
    int *q_slot;
    double *q_dir, *q_dist;
    SearchResult *q_res;
    SortBufs sb;
    int *sorted_label, *scan_tmp;    // scan_tmp: >= max(agg_cap, n_cells) + 2 ints
    double4 *alt_posr, *alt_relv;    // alternate sphere buffers for the in-kernel pool compaction
    double *alt_surf, *alt_veff, *alt_seff, *alt_dcen;
    int *alt_id, *alt_charge;
    mcac_step_record *rec;
    long long rec_cap, rec_base;
    long long max_steps;
    long long dup_threshold, full_freq;
    int with_nucleation, with_potentials, growth, individual, pick_last, with_collisions, with_domain_duplication;
    int cum_sequential_max, stable, depth_override;
    int pick_valid, labels_valid, stop_at_event;
    int max_slots;                   // the loop hands back (LOOP_TOO_BIG) when the aggregate table outgrows this
    int prune;                       // 0: every sphere pair of an examined suspect is tested (MCAC_B200_NO_PRUNE)
    int dups_allowed;                // domain duplications this launch may do by itself (<= 4; 0: always hand them to the host)
    double dup_box_volume[4];        // PhysicalModel::box_volume after the 1st .. 4th duplication from now: std::pow(box_length, 3) is
                                     // evaluated by the HOST (the reference's libm), like every other box quantity
    LoopState *out;
};

// ---- block-wide helpers (every thread of the CTA calls them) --------------------------------------------------------------
// exclusive scan of in[0..n) -> out[0..n], out[n] = total (in and out may alias)
__device__ __forceinline__ void cta_scan_int(const int *in, int n, int *out) {
    __shared__ int ws[32];
    __shared__ int carry;
    const int tid = threadIdx.x, nth = blockDim.x;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n; b0 += nth) {
        const int i = b0 + tid;
        const int v = i < n ? in[i] : 0;
        int total;
        const int pre = block_exclusive_scan(v, &total, ws);
        const int base = carry;
        __syncthreads();
        if (i < n) out[i] = base + pre;
        if (tid == 0) carry = base + total;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
    __syncthreads();
}
__device__ __forceinline__ void cta_scan_ll(const long long *in, int n, long long *out) {
    __shared__ long long ws[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, nth = blockDim.x;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n; b0 += nth) {
        const int i = b0 + tid;
        const long long v = i < n ? in[i] : 0;
        long long total;
        const long long pre = block_exclusive_scan_ll(v, &total, ws);
        const long long base = carry;
        __syncthreads();
        if (i < n) out[i] = base + pre;
        if (tid == 0) carry = base + total;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
    __syncthreads();
}

// labels = rank of the live slots (k_alive_to_labels)
__device__ __forceinline__ void cta_labels(const DevState &d, int *scan) {
    const int n = d.sc->n_agg_slots;
    cta_scan_int(d.a_alive, n, scan);
    for (int s = threadIdx.x; s < n; s += blockDim.x) {
        if (d.a_alive[s]) { d.label_of_slot[s] = scan[s]; d.slot_of_label[scan[s]] = s; }
        else d.label_of_slot[s] = -1;
    }
    __syncthreads();
}

// K2 in one CTA: counting sort of the live aggregates into their stored Verlet cells (+ bounding spheres in cell order)
__device__ __forceinline__ void cta_build_cells(const DevState &d) {
    const int tid = threadIdx.x, nth = blockDim.x;
    const int n = d.sc->n_agg_slots, nc = d.n_cells;
    for (int c = tid; c <= nc; c += nth) d.cell_fill[c] = 0;
    __syncthreads();
    for (int s = tid; s < n; s += nth)
        if (d.a_alive[s]) atomicAdd(&d.cell_fill[(d.a_cx[s] * d.n_div + d.a_cy[s]) * d.n_div + d.a_cz[s]], 1);
    __syncthreads();
    cta_scan_int(d.cell_fill, nc, d.cell_start);
    for (int c = tid; c <= nc; c += nth) d.cell_fill[c] = 0;
    __syncthreads();
    for (int s = tid; s < n; s += nth) {
        if (!d.a_alive[s]) continue;
        const int c = (d.a_cx[s] * d.n_div + d.a_cy[s]) * d.n_div + d.a_cz[s];
        const int pos = d.cell_start[c] + atomicAdd(&d.cell_fill[c], 1);
        d.cell_items[pos] = s;
        d.cell_posr[pos] = d.a_posr[s];
    }
    __syncthreads();
}

// AggregatList::sort_time_steps(max_time_step) in one CTA: the multi-launch form of engine.cu (k_make_keys, k_sort_init, the level
// loop k_sort_pivot / flags / scan / scatter / swap / split, k_sort_heap behind the depth limit, k_sort_leaves, the sequential or
// tree cumulative sum, k_sort_finish) with block barriers instead of launches — same arithmetic, same permutation.
__device__ __forceinline__ void cta_sort_time_steps(const DevState &d, const LoopArgs &a) {
    __shared__ int s_active, s_fail;
    const int tid = threadIdx.x, nth = blockDim.x;
    Scalars &sc = *d.sc;
    SortBufs b = a.sb;
    const int n = sc.n_agg;
    b.n = n;
    b.stable = a.stable;
    const double factor = sc.max_time_step;
    for (int l = tid; l < n; l += nth) {
        const double k = factor / d.a_ts[d.slot_of_label[l]];
        d.keys[l] = k;
        b.perm[l] = l;
        b.wk[l] = k;
        b.segf[l] = 0;
        b.segl[l] = n;
    }
    int lg = 0;
    while ((1LL << (lg + 1)) <= n) lg++;
    int depth = a.depth_override >= 0 ? a.depth_override : 2 * lg;
    bool active = n > kSortLeaf, fail = false;
    if (active && depth == 0) { fail = true; active = false; }
    __syncthreads();
    while (active && !fail) {
        depth--;
        // __move_median_to_first by the segment leaders
        for (int i = tid; i < n; i += nth) {
            if (b.segf[i] != i) continue;
            const int f = i, l = b.segl[i];
            if (l - f <= kSortLeaf) continue;
            const int pa = f + 1, pb = f + (l - f) / 2, pc = l - 1;
            const double ka = b.wk[pa], kb = b.wk[pb], kc = b.wk[pc];
            const int la = b.perm[pa], lb = b.perm[pb], lc = b.perm[pc];
            int pick;
            if (w_less(ka, la, kb, lb, b.stable)) {
                if (w_less(kb, lb, kc, lc, b.stable)) pick = pb;
                else if (w_less(ka, la, kc, lc, b.stable)) pick = pc;
                else pick = pa;
            } else if (w_less(ka, la, kc, lc, b.stable)) pick = pa;
            else if (w_less(kb, lb, kc, lc, b.stable)) pick = pc;
            else pick = pb;
            const double kf = b.wk[f];
            const int lf = b.perm[f];
            b.wk[f] = b.wk[pick]; b.perm[f] = b.perm[pick];
            b.wk[pick] = kf; b.perm[pick] = lf;
        }
        __syncthreads();
        // "not < pivot" / "not > pivot" flags
        for (int i = tid; i < n; i += nth) {
            const int f = b.segf[i], l = b.segl[i];
            long long fl = 0;
            if (l - f > kSortLeaf && i > f) {
                const double kp = b.wk[f], kx = b.wk[i];
                const int lp = b.perm[f], lx = b.perm[i];
                if (!w_less(kx, lx, kp, lp, b.stable)) fl |= 1LL;
                if (!w_less(kp, lp, kx, lx, b.stable)) fl |= (1LL << 32);
            }
            b.flags[i] = fl;
        }
        if (tid == 0) { b.flags[n] = 0; s_active = 0; s_fail = 0; }
        __syncthreads();
        cta_scan_ll(b.flags, n + 1, b.pre);
        // the k-th left stopper / k-th right stopper of every segment
        for (int i = tid; i < n; i += nth) {
            const int f = b.segf[i], l = b.segl[i];
            if (l - f <= kSortLeaf || i <= f) continue;
            const int base = f + 1;
            const long long p0 = b.pre[base], pi_ = b.pre[i], pl = b.pre[l];
            const long long fl = b.flags[i];
            if (fl & 1LL) b.tmp_a[base + (int)((pi_ & 0xffffffffLL) - (p0 & 0xffffffffLL))] = i;
            if (fl >> 32) {
                const int n_b = (int)((pl >> 32) - (p0 >> 32));
                b.tmp_b[base + n_b - 1 - (int)((pi_ >> 32) - (p0 >> 32))] = i;
            }
        }
        __syncthreads();
        // the Hoare swaps while the two scans have not crossed, and where they stop
        for (int j = tid; j < n; j += nth) {
            const int f = b.segf[j], l = b.segl[j];
            if (l - f <= kSortLeaf || j <= f) continue;
            const int base = f + 1, k = j - base;
            const long long p0 = b.pre[base], pl = b.pre[l];
            const int n_a = (int)((pl & 0xffffffffLL) - (p0 & 0xffffffffLL)), n_b = (int)((pl >> 32) - (p0 >> 32));
            const int m = n_a < n_b ? n_a : n_b;
            const bool sw = k < m && b.tmp_a[base + k] < b.tmp_b[base + k];
            const bool next_sw = (k + 1 < m) && b.tmp_a[base + k + 1] < b.tmp_b[base + k + 1];
            if (sw) {
                const int pa = b.tmp_a[base + k], pb = b.tmp_b[base + k];
                const double ka = b.wk[pa];
                const int la = b.perm[pa];
                b.wk[pa] = b.wk[pb]; b.perm[pa] = b.perm[pb];
                b.wk[pb] = ka; b.perm[pb] = la;
            }
            int s = -1;
            if (sw && !next_sw) s = k + 1;
            else if (k == 0 && !sw) s = 0;
            if (s >= 0) {
                const int a_s = s < n_a ? b.tmp_a[base + s] : 0x7fffffff;
                const int b_prev = s > 0 ? b.tmp_b[base + s - 1] : l;
                b.cut[f] = a_s < b_prev ? a_s : b_prev;
            }
        }
        __syncthreads();
        // split
        for (int i = tid; i < n; i += nth) {
            const int f = b.segf[i], l = b.segl[i];
            if (l - f <= kSortLeaf) continue;
            const int c = b.cut[f];
            int nf = f, nl = l;
            if (i < c) nl = c; else nf = c;
            b.segf[i] = nf;
            b.segl[i] = nl;
            if (i == nf && nl - nf > kSortLeaf) {
                if (depth > 0) s_active = 1; else s_fail = 1;  // would enter the heap-sort branch of introsort
            }
        }
        __syncthreads();
        active = s_active != 0;
        fail = s_fail != 0;
        __syncthreads();
    }
    if (fail) {  // introsort's depth limit: heap-sort branch of every segment still longer than 16
        for (int i = tid; i < n; i += nth) {
            if (b.segf[i] != i) continue;
            const int l = b.segl[i];
            if (l - i > kSortLeaf) heapsort::heap_sort_segment(heapsort::HeapView{b.wk, b.perm, b.stable}, i, l);
        }
        __syncthreads();
    }
    // __final_insertion_sort per leaf
    for (int i = tid; i < n; i += nth) {
        if (b.segf[i] != i) continue;
        const int f = i, l = b.segl[i];
        if (l - f > kSortLeaf) continue;
        for (int x = f + 1; x < l; x++) {
            const double kv = b.wk[x];
            const int lv = b.perm[x];
            int y = x - 1;
            while (y >= f && w_less(kv, lv, b.wk[y], b.perm[y], b.stable)) {
                b.wk[y + 1] = b.wk[y]; b.perm[y + 1] = b.perm[y];
                y--;
            }
            b.wk[y + 1] = kv; b.perm[y + 1] = lv;
        }
    }
    __syncthreads();
    // cumulative_time_steps, sequential: the reference's rounding (aggregat_list.cpp:133-140).  The loop leaves tables of more than
    // cum_sequential_max aggregates to the multi-launch path (LOOP_TOO_BIG), like the tree-summed form of k_cum_*.
    if (tid == 0) {
        double acc = b.wk[0];
        d.cum[0] = acc;
        for (int i = 1; i < n; i++) { acc = acc + b.wk[i]; d.cum[i] = acc; }
    }
    __syncthreads();
    for (int i = tid; i < n; i += nth) {
        a.sorted_label[i] = b.perm[i];
        d.sorted_slot[i] = d.slot_of_label[b.perm[i]];
    }
    if (tid == 0) { sc.n_pick = n; sc.cum_total = d.cum[n - 1]; sc.pick_dense_from = n; sc.last_sort_tie = 0; }
    __syncthreads();
}

// pool compaction (k_compact_counts / move / finish): live aggregates re-packed in slot order into the alternate buffers; the
// caller's DevState copy swaps the two sets of pointers
__device__ __forceinline__ void cta_compact_pool(DevState &d, LoopArgs &a, int *scan) {
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    const int n = d.sc->n_agg_slots;
    for (int s = tid; s < n; s += nth) scan[s] = d.a_alive[s] ? d.a_n[s] : 0;
    __syncthreads();
    cta_scan_int(scan, n, scan);
    for (int slot = warp; slot < n; slot += nwarps) {
        if (!d.a_alive[slot]) continue;
        const int off = d.a_off[slot], cnt = d.a_n[slot], to = scan[slot];
        for (int i = lane; i < cnt; i += 32) {
            a.alt_posr[to + i] = d.s_posr[off + i];
            a.alt_relv[to + i] = d.s_relv[off + i];
            a.alt_surf[to + i] = d.s_surf[off + i];
            a.alt_veff[to + i] = d.s_veff[off + i];
            a.alt_seff[to + i] = d.s_seff[off + i];
            a.alt_dcen[to + i] = d.s_dcen[off + i];
            const int id = d.s_id[off + i];
            a.alt_id[to + i] = id;
            a.alt_charge[to + i] = d.s_charge[off + i];
            d.slot_of_id[id] = to + i;
        }
        __syncwarp();
        if (lane == 0) d.a_off[slot] = to;
    }
    __syncthreads();
    if (tid == 0) d.sc->pool_top = scan[n];
    // every thread swaps its own copy of the pointers (DevState and LoopArgs live in shared memory: one thread does it)
    __syncthreads();
    if (tid == 0) {
        double4 *p4;
        double *pd;
        int *pi;
        p4 = d.s_posr; d.s_posr = a.alt_posr; a.alt_posr = p4;
        p4 = d.s_relv; d.s_relv = a.alt_relv; a.alt_relv = p4;
        pd = d.s_surf; d.s_surf = a.alt_surf; a.alt_surf = pd;
        pd = d.s_veff; d.s_veff = a.alt_veff; a.alt_veff = pd;
        pd = d.s_seff; d.s_seff = a.alt_seff; a.alt_seff = pd;
        pd = d.s_dcen; d.s_dcen = a.alt_dcen; a.alt_dcen = pd;
        pi = d.s_id; d.s_id = a.alt_id; a.alt_id = pi;
        pi = d.s_charge; d.s_charge = a.alt_charge; a.alt_charge = pi;
        a.out->flipped ^= 1;
        a.out->compactions += 1;
    }
    __syncthreads();
}

// refresh() + get_total_volume/surface + PhysicalModel::update at the end of a general step (k_refresh_partials + k_step_totals /
// k_refresh_if_event): max is exact; the two sums are combined in a fixed order (per-thread strided partials, warp butterfly,
// warps in order) — they only feed the reported concentrations / volume fraction
__device__ __forceinline__ void cta_refresh(const DevState &d, bool growth, bool totals_only = false) {
    __shared__ double sm[3][32];
    Scalars &sc = *d.sc;
    if (!totals_only && !growth && !sc.event) return;  // (uniform: every thread reads the same flag after the barrier that followed its write)
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    const int n = sc.n_agg_slots;
    double mx = 0., sv = 0., ss = 0.;
    for (int s = tid; s < n; s += nth) {
        if (!d.a_alive[s]) continue;
        const double ts = d.a_ts[s];
        mx = (mx < ts) ? ts : mx;
        sv += d.a_vol[s];
        ss += d.a_surf[s];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double omx = __shfl_xor_sync(kFull, mx, o);
        mx = (mx < omx) ? omx : mx;
        sv += __shfl_xor_sync(kFull, sv, o);
        ss += __shfl_xor_sync(kFull, ss, o);
    }
    if (lane == 0) { sm[0][warp] = mx; sm[1][warp] = sv; sm[2][warp] = ss; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < nwarps; w++) {
            mx = (mx < sm[0][w]) ? sm[0][w] : mx;
            sv += sm[1][w];
            ss += sm[2][w];
        }
        if (sc.event && !totals_only) {
            sc.max_time_step = mx;
            sc.avg_npp = static_cast<double>(sc.n_sph) / static_cast<double>(sc.n_agg);
        }
        sc.total_volume = sv;
        sc.total_surface = ss;
        sc.total_volume_concent = sv / sc.box_volume;
        sc.total_surface_concent = ss / sc.box_volume;
        sc.aggregate_concentration = static_cast<double>(sc.n_agg) / sc.box_volume;
        sc.monomer_concentration = static_cast<double>(sc.n_sph) / sc.box_volume;
        sc.volume_fraction = sv / sc.box_volume;
    }
    __syncthreads();
}

// AggregatList::duplication (aggregat_list.cpp:142-190) inside the loop, when the tables have room for it: the box doubles, every
// aggregate gets 7 translated copies appended in (aggregate, shift) order — labels n0 + 7a + c - 1 and fresh sphere indices in member
// order, exactly as the copy constructor chain of the reference numbers them (aggregat_storage.cpp:117-159) — then the copies are moved
// by (i,j,k) * old box with the NEW box's periodicity, every Verlet cell is recomputed and PhysicalModel::update refreshes the
// concentrations (:186-189).  Slots are stable: the live slots of the originals come first, the copies are appended behind them, so
// the rank of a slot among the live slots is the reference's label.  Returns false (nothing touched) when the tables are too small:
// the host then does the same through the upload boundary, with bigger tables.
__device__ __forceinline__ bool cta_duplicate(DevState &d, LoopArgs &a, bool &labels_valid, int dup_index) {
    __shared__ int s_ok;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    Scalars &sc = *d.sc;
    const int n0 = sc.n_agg, m0 = sc.n_sph;
    if (d.sph_cap - sc.pool_top < 7 * m0 && sc.pool_top > m0) cta_compact_pool(d, a, a.scan_tmp);
    if (tid == 0)
        s_ok = (sc.n_agg_slots + 7LL * n0 + 64 <= d.agg_cap && sc.pool_top + 7LL * m0 + (a.with_nucleation ? m0 + 64 : 0) <= d.sph_cap &&
                8LL * m0 + 64 <= d.sph_cap) ? 1 : 0;
    __syncthreads();
    if (!s_ok) return false;
    if (!labels_valid) { cta_labels(d, a.scan_tmp); labels_valid = true; }
    int *scan = a.scan_tmp;
    for (int l = tid; l < n0; l += nth) scan[l] = d.a_n[d.slot_of_label[l]];
    __syncthreads();
    cta_scan_int(scan, n0, scan);  // spheres of the aggregates with a smaller label
    const int base_slot = sc.n_agg_slots, base_pool = sc.pool_top;
    const double old_l = sc.box_length;
    __syncthreads();
    for (int t = warp; t < 7 * n0; t += nwarps) {
        const int al = t / 7, c = t % 7 + 1;
        const int src = d.slot_of_label[al], dst = base_slot + t;
        const int n = d.a_n[src], off_src = d.a_off[src];
        const int off_dst = base_pool + 7 * scan[al] + (c - 1) * n;
        const int id0 = m0 + 7 * scan[al] + (c - 1) * n;
        if (lane == 0) {
            d.a_posr[dst] = d.a_posr[src];
            d.a_rg[dst] = d.a_rg[src]; d.a_fagg[dst] = d.a_fagg[src]; d.a_lpm[dst] = d.a_lpm[src]; d.a_ts[dst] = d.a_ts[src];
            d.a_vol[dst] = d.a_vol[src]; d.a_surf[dst] = d.a_surf[src];
            d.a_rx[dst] = d.a_rx[src]; d.a_ry[dst] = d.a_ry[src]; d.a_rz[dst] = d.a_rz[src]; d.a_ptime[dst] = d.a_ptime[src];
            d.a_dp[dst] = d.a_dp[src]; d.a_dgdp[dst] = d.a_dgdp[src]; d.a_ovl[dst] = d.a_ovl[src]; d.a_cn[dst] = d.a_cn[src];
            d.a_dm[dst] = d.a_dm[src]; d.a_ch[dst] = d.a_ch[src]; d.a_bulk[dst] = d.a_bulk[src]; d.a_alpha[dst] = d.a_alpha[src];
            d.a_n[dst] = n; d.a_off[dst] = off_dst;
            d.a_cx[dst] = d.a_cx[src]; d.a_cy[dst] = d.a_cy[src]; d.a_cz[dst] = d.a_cz[src];
            d.a_charge[dst] = 0;  // the copy constructor resets the charge (aggregat_storage.cpp:138)
            d.a_alive[dst] = 1;
            d.a_dirty[dst] = d.a_dirty[src];
            d.label_of_slot[dst] = n0 + t;
            d.slot_of_label[n0 + t] = dst;
        }
        for (int k = lane; k < n; k += 32) {
            d.s_posr[off_dst + k] = d.s_posr[off_src + k];
            d.s_relv[off_dst + k] = d.s_relv[off_src + k];
            d.s_surf[off_dst + k] = d.s_surf[off_src + k];
            d.s_veff[off_dst + k] = d.s_veff[off_src + k];
            d.s_seff[off_dst + k] = d.s_seff[off_src + k];
            d.s_dcen[off_dst + k] = d.s_dcen[off_src + k];
            d.s_id[off_dst + k] = id0 + k;
            d.s_charge[off_dst + k] = 0;
            d.slot_of_id[id0 + k] = off_dst + k;
        }
    }
    __syncthreads();
    if (tid == 0) {
        sc.box_length = old_l * 2;
        sc.n_monomeres *= 8;
        sc.box_volume = a.dup_box_volume[dup_index];
        sc.n_agg = 8 * n0;
        sc.n_agg_slots = base_slot + 7 * n0;
        sc.n_sph = 8 * m0;
        sc.pool_top = base_pool + 7 * m0;
        a.out->dups += 1;
    }
    __syncthreads();
    const double box = sc.box_length;
    const int n_slots = sc.n_agg_slots;
    for (int s = warp; s < n_slots; s += nwarps) {  // k_dup_finish
        if (!d.a_alive[s]) continue;
        if (s >= base_slot) {
            const int c = (s - base_slot) % 7 + 1;
            agg_translate<false>(d, s, ((c >> 2) & 1) * old_l, ((c >> 1) & 1) * old_l, (c & 1) * old_l, box, lane, 32);
        } else if (lane == 0) {
            const double4 p = d.a_posr[s];
            d.a_cx[s] = cell_of(p.x, d.n_div, box);
            d.a_cy[s] = cell_of(p.y, d.n_div, box);
            d.a_cz[s] = cell_of(p.z, d.n_div, box);
        }
    }
    __syncthreads();
    cta_refresh(d, true, true);  // PhysicalModel::update: totals and concentrations for the new box; max_time_step / avg_npp are kept
    return true;
}

// ---- the loop ---------------------------------------------------------------------------------------------------------------
__device__ void step_loop(DevState &d, LoopArgs &a) {
    __shared__ double upd_scratch[kLoopThreads / 32][kUpdateScratch / 4];
    __shared__ double picked_scratch[kUpdateScratch];
    // terms of the ordered sums (agg_update's `stage`): 5 x 512 doubles for the CTA-wide update, 5 x 32 per warp otherwise — dynamic
    // shared memory (kLoopDynSmem bytes), the static part of this kernel is already at 31 KB
    double *upd_stage = reinterpret_cast<double *>(dyn_smem);
    __shared__ int s_draws, s_ntry, s_draws_at_search, s_again;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    Scalars &sc = *d.sc;
    LoopState &out = *a.out;
    bool pick_valid = a.pick_valid != 0, labels_valid = a.labels_valid != 0;
    long long steps = 0;
    int reason = LOOP_STEPS_DONE, dups_done = 0;
    if (tid == 0) {
        out.steps = 0; out.events = 0; out.nucleated = 0; out.sorts = 0; out.compactions = 0; out.dups = 0; out.flipped = 0; out.exit_reason = LOOP_STEPS_DONE;
        for (int k = 0; k < 8; k++) out.phase_cycles[k] = 0;
    }
    long long t_prev = clock64();
    auto lap = [&](int k) { if (tid == 0) { const long long t = clock64(); out.phase_cycles[k] += t - t_prev; t_prev = t; } };
    __syncthreads();
    while (steps < a.max_steps) {
        // ---- loop top of calcul(): PhysicalModel::finished, duplication test, room in the tables, draws staged
        if (sc.error != 0) { reason = LOOP_ERROR; break; }
        if (sc.n_agg < 1 || sc.n_agg <= d.n_agg_limit || (d.n_iter_limit > 0 && sc.n_iter_without_event >= d.n_iter_limit) ||
            (d.time_limit > 0 && sc.time >= d.time_limit) || (d.npp_limit > 0 && sc.avg_npp >= static_cast<double>(d.npp_limit))) {
            reason = LOOP_FINISHED;
            break;
        }
        if (sc.event && a.with_domain_duplication && sc.n_agg <= a.dup_threshold && !(d.u_sg < 0.0)) {
            if (dups_done >= a.dups_allowed || !cta_duplicate(d, a, labels_valid, dups_done)) { reason = LOOP_NEED_DUP; break; }
            dups_done++;
            pick_valid = false;
        }
        if (sc.n_agg_slots > a.max_slots || sc.n_agg > a.cum_sequential_max) { reason = LOOP_TOO_BIG; break; }
        // (with nucleation the regrow test below wants 64 more free slots than the plain compaction test: compact for those as well,
        // a regrow goes through the host)
        if (d.sph_cap - sc.pool_top < sc.n_sph + (a.with_nucleation ? 64 : 0) && sc.pool_top > sc.n_sph) cta_compact_pool(d, a, a.scan_tmp);
        if (a.with_nucleation && (d.agg_cap - sc.n_agg_slots < 64 || d.sph_cap - sc.pool_top < sc.n_sph + 64)) { reason = LOOP_NEED_REGROW; break; }
        if (sc.rand_pos < d.rng_buf_base || sc.rand_pos - d.rng_buf_base + 8192 + 64 > d.rng_buf_n) { reason = LOOP_NEED_RNG; break; }
        lap(4);
        // ---- pick table
        if (!labels_valid) { cta_labels(d, a.scan_tmp); labels_valid = true; }
        if (!a.pick_last && (sc.event || a.growth || !pick_valid)) {
            cta_sort_time_steps(d, a);
            pick_valid = true;
            if (tid == 0) out.sorts += 1;
        }
        lap(0);
        // ---- pick + direction
        if (a.pick_last) {
            dev_pick_last(d, a.q_slot);
            __syncthreads();
            dev_prepare_direction(d, a.q_slot, a.q_dir, a.q_dist, 0);
        } else if (tid == 0) {
            dev_prepare_query(d, 0, a.q_slot, a.q_dir, a.q_dist);
        }
        if (tid == 0) { s_draws = a.pick_last ? 2 : 3; s_ntry = 1; s_draws_at_search = a.pick_last ? 2 : 3; }
        __syncthreads();
        // ---- contact search + orientation loop (calcul.cpp:119-141)
        if (a.with_collisions) {
            cta_build_cells(d);
            search_wide_one<0, kLoopThreads, true>(d, 0, a.q_slot, a.q_dir, a.q_dist, a.q_res, nullptr, a.prune ? a.alt_id : nullptr, a.alt_charge);
            __syncthreads();
            while (a.with_potentials) {
                dev_check_regime(d, a.q_res, a.q_dist, s_draws);
                __syncthreads();
                if (tid == 0) {
                    s_again = 0;
                    if (sc.error == 0) {
                        s_draws += sc.p_regime_draws;
                        if (sc.p_regime != 0) {
                            s_ntry += 1;
                            if (s_draws + 8 > 8192) { sc.error = 1; sc.error_detail = DETAIL_RNG_NOT_STAGED; }
                            else s_again = 1;
                        }
                    }
                }
                __syncthreads();
                if (!s_again) break;
                dev_prepare_direction(d, a.q_slot, a.q_dir, a.q_dist, s_draws);
                if (tid == 0) { s_draws += 2; s_draws_at_search = s_draws; }
                __syncthreads();
                search_wide_one<0, kLoopThreads, true>(d, 0, a.q_slot, a.q_dir, a.q_dist, a.q_res, nullptr, a.prune ? a.alt_id : nullptr, a.alt_charge);
                __syncthreads();
            }
            if (sc.error != 0) { reason = LOOP_ERROR; break; }
        }
        lap(1);
        // ---- move + clocks, growth, deferred merge
        const long long iter_before = sc.n_iter_without_event;
        StepArgs sa;
        sa.q_slot = a.q_slot; sa.q_dir = a.q_dir; sa.q_dist = a.q_dist; sa.res = a.q_res;
        sa.rec = a.rec; sa.rec_cap = a.rec_cap; sa.rec_index = a.rec_base + steps;
        sa.pick_last = a.pick_last; sa.with_collisions = a.with_collisions; sa.n_try = s_ntry; sa.draws = s_draws; sa.draws_at_search = s_draws_at_search;
        __syncthreads();
        dev_step_move(d, sa);
        __syncthreads();
        lap(5);
        if (sc.error != 0) { reason = LOOP_ERROR; break; }
        if (a.growth) {  // k_grow_pending
            int lo = 0, hi = sc.pool_top;
            double dt = sc.p_dt;
            if (a.individual) { lo = d.a_off[sc.p_slot]; hi = lo + d.a_n[sc.p_slot]; dt = sc.p_dt_indiv; }
            if (a.individual) { if (tid == 0) d.a_dirty[sc.p_slot] = kDirtyAll; }
            else for (int s2 = tid; s2 < sc.n_agg_slots; s2 += nth) d.a_dirty[s2] = kDirtyAll;
            for (int t = lo + tid; t < hi; t += nth) {
                double4 p = d.s_posr[t];
                const double new_r = p.w + d.u_sg * dt;
                const double r2 = new_r * new_r;
                const double r3 = r2 * new_r;
                p.w = new_r;
                d.s_posr[t] = p;
                double4 rel = d.s_relv[t];
                rel.w = volume_factor() * r3;
                d.s_relv[t] = rel;
                d.s_surf[t] = surface_factor() * r2;
                if (new_r <= d.rp_min_oxid) { sc.error = 1; sc.error_detail = DETAIL_SPHERE_REMOVAL; }
            }
            __syncthreads();
        }
        lap(6);
        dev_step_merge(d, sa.rec, sa.rec_cap, sa.rec_index, upd_stage);
        __syncthreads();
        lap(7);
        // ---- update block of calcul.cpp:184-206
        if (a.growth) {
            const bool full = (iter_before % a.full_freq) == 0;
            const bool everyone = !(a.individual && !sc.b_merged);
            if (!everyone) {
                const int slot = sc.p_slot;
                if (slot >= 0 && slot < sc.n_agg_slots && d.a_alive[slot]) agg_update<true>(d, slot, full, tid, nth, picked_scratch, sc.box_length, upd_stage);
            } else {
                // small aggregates one per thread, medium ones one per warp, big ones by the whole CTA one after the other (a full
                // update is O(n^2): a 10^3-sphere aggregate left to one warp would keep the other 15 waiting for milliseconds)
                const int n_slots = sc.n_agg_slots;
                for (int s = tid; s < n_slots; s += nth)
                    if (d.a_alive[s] && d.a_n[s] <= kSingleMax) agg_update_single(d, s, full, sc.box_length);
                __syncthreads();
                // (only an O(n^2) pass needs the whole CTA: a big aggregate that is clean, or a partial update, is a chain of ordered adds —
                // sixteen of those run side by side on sixteen warps)
                for (int s = warp; s < n_slots; s += nwarps)
                    if (d.a_alive[s] && d.a_n[s] > kSingleMax && !(full && d.a_n[s] > kUpdateWarpMax && (d.a_dirty[s] & kDirtyFull)))
                        agg_update<false>(d, s, full, lane, 32, upd_scratch[warp], sc.box_length, upd_stage + warp * (5 * 32));
                __syncthreads();
                if (full)
                    for (int s = 0; s < n_slots; s++)
                        if (d.a_alive[s] && d.a_n[s] > kUpdateWarpMax && (d.a_dirty[s] & kDirtyFull))
                            agg_update<true>(d, s, full, tid, nth, picked_scratch, sc.box_length, upd_stage);
            }
            __syncthreads();
        }
        lap(2);
        // ---- nucleation, event bookkeeping, refresh / PhysicalModel::update
        if (a.with_nucleation) dev_nucleate(d, 0., 1);
        else if (tid == 0) sc.n_nucleated = 0;
        __syncthreads();
        dev_step_event(d);
        __syncthreads();
        cta_refresh(d, a.growth != 0);
        lap(3);
        if (sc.error != 0) { reason = LOOP_ERROR; break; }
        steps += 1;
        const bool merged = sc.b_merged != 0, nucl = sc.n_nucleated > 0;
        if (tid == 0) { out.events += merged ? 1 : 0; out.nucleated += sc.n_nucleated; }
        if (merged) { pick_valid = false; labels_valid = false; }
        if (nucl) pick_valid = false;
        if (a.stop_at_event && (merged || nucl)) { reason = LOOP_EVENT_STOP; break; }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
        out.exit_reason = reason;
        out.steps = steps;
        out.pick_valid = pick_valid ? 1 : 0;
        out.labels_valid = labels_valid ? 1 : 0;
    }
}

// AggregatList::duplication alone (the speculative-batch path and the per-call C ABI): one CTA, in place when the tables have room
// (LoopState::dups == 1 afterwards; 0: nothing was touched, the host re-allocates and does it through the upload boundary)
__global__ void __launch_bounds__(kLoopThreads) k_duplicate(DevState d_in, LoopArgs a_in) {
    __shared__ DevState d;
    __shared__ LoopArgs a;
    if (threadIdx.x == 0) {
        d = d_in; a = a_in;
        LoopState &out = *a_in.out;
        out.steps = 0; out.events = 0; out.nucleated = 0; out.sorts = 0; out.compactions = 0; out.dups = 0; out.flipped = 0;
        for (int k = 0; k < 8; k++) out.phase_cycles[k] = 0;
    }
    __syncthreads();
    bool labels_valid = a.labels_valid != 0;
    const bool ok = cta_duplicate(d, a, labels_valid, 0);
    __syncthreads();
    if (threadIdx.x == 0) {
        a.out->exit_reason = ok ? LOOP_STEPS_DONE : LOOP_NEED_DUP;
        a.out->labels_valid = labels_valid ? 1 : 0;
        a.out->pick_valid = 0;
    }
}
// one realization, one CTA
__global__ void __launch_bounds__(kLoopThreads) k_step_loop(DevState d_in, LoopArgs a_in) {
    __shared__ DevState d;
    __shared__ LoopArgs a;
    if (threadIdx.x == 0) { d = d_in; a = a_in; }
    __syncthreads();
    step_loop(d, a);
}
// Scalars of many realizations into one contiguous array (one copy to the host instead of one per handle)
__global__ void k_gather_scalars(Scalars *const *ptrs, int n, Scalars *out) {
    const int per = (int)(sizeof(Scalars) / sizeof(int));
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < (long long)n * per; k += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(k / per), w = (int)(k % per);
        reinterpret_cast<int *>(out + r)[w] = reinterpret_cast<const int *>(ptrs[r])[w];
    }
}
// many realizations, CTAs take them from a queue; `ds` is updated in place (compaction swaps the sphere buffers)
// sc_all / ls_all (optional): every realization's Scalars and LoopState are also left in these contiguous arrays when it leaves the
// loop, so that the host reads the whole round back with two copies instead of two per realization
__global__ void __launch_bounds__(kLoopThreads) k_ensemble_loop(DevState *ds, LoopArgs *as, int n, int *next, Scalars *sc_all, LoopState *ls_all) {
    __shared__ DevState d;
    __shared__ LoopArgs a;
    __shared__ int s_r;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_r = atomicAdd(next, 1);
        __syncthreads();
        const int r = s_r;
        if (r >= n) return;
        if (threadIdx.x == 0) { d = ds[r]; a = as[r]; }
        __syncthreads();
        step_loop(d, a);
        __syncthreads();
        if (threadIdx.x == 0) { ds[r] = d; as[r] = a; }
        if (sc_all) {
            const int *src = reinterpret_cast<const int *>(d.sc);
            int *dst = reinterpret_cast<int *>(sc_all + r);
            for (int k = threadIdx.x; k < (int)(sizeof(Scalars) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
            const int *ls = reinterpret_cast<const int *>(a.out);
            int *ld = reinterpret_cast<int *>(ls_all + r);
            for (int k = threadIdx.x; k < (int)(sizeof(LoopState) / sizeof(int)); k += blockDim.x) ld[k] = ls[k];
        }
    }
}

}  // namespace mcacb
