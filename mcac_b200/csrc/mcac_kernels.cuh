// mcac_b200 — hand-written sm_100a kernels of the Monte-Carlo aggregation step.
// K1 contact search, K2 cell list (counting sort), K3 translate/commit, K4 merge, K5-K7 aggregate update,
// K8 surface growth, K9 pick table helpers, K10 RNG stream.  No tensor cores: nothing here is a dense
// contraction (FP64 scalar pipes + HBM/L2 streams).  Compiled with --fmad=false (see mcac_math.cuh).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "../../include/mcac_b200.h"
#include "mcac_device.cuh"
#include "tie_sort.cuh"
#include "heap_sort.cuh"
#include "seq_cumsum.cuh"

namespace mcacb {

constexpr int kSearchThreads = 128;
constexpr int kCandCap = 512;   // eligible aggregates per search kept in shared memory
constexpr int kCommitThreads = 512;
constexpr int kMaxBatch = 512;
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// small cooperative helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_inclusive_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
// exclusive scan of one int per thread over the whole block (blockDim multiple of 32, <= 1024); returns prefix, total via *total
__device__ __forceinline__ int block_exclusive_scan(int v, int *total, int *warp_sums /* >= 32 ints of smem */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int inc = warp_inclusive_scan(v, lane);
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < nw ? warp_sums[lane] : 0;
        s = warp_inclusive_scan(s, lane);
        warp_sums[lane] = s;
    }
    __syncthreads();
    const int base = w ? warp_sums[w - 1] : 0;
    *total = warp_sums[nw - 1];
    __syncthreads();
    return base + inc - v;
}
// lexicographic (distance, index) min across a warp: strict `<` on the distance keeps the FIRST minimum in visiting
// order, which is what the reference's nested loops do (sphere_contact.cpp:131-137, aggregat_distance.cpp:30-41).
__device__ __forceinline__ void warp_argmin(double &d, long long &idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(kFull, d, o);
        const long long oi = __shfl_xor_sync(kFull, idx, o);
        if (od < d || (od == d && oi < idx)) {
            d = od;
            idx = oi;
        }
    }
}
__device__ __forceinline__ void atomic_max_positive_double(double *addr, double v) {
    atomicMax(reinterpret_cast<long long *>(addr), __double_as_longlong(v));  // valid for non-negative doubles
}

// ------------------------------------------------------------------------------------------------
// K10 — glibc rand() stream: one thread advances the 31-word ring kept in registers/local memory.
// ------------------------------------------------------------------------------------------------
__global__ void k_rng_fill(GlibcRandState *st, int *out, int n) {  // n multiple of 31, ring position 0 on entry and exit
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t r[31];
#pragma unroll
    for (int k = 0; k < 31; k++) r[k] = st->ring[(st->pos + k) % 31];
    for (int base = 0; base + 31 <= n; base += 31) {
#pragma unroll
        for (int k = 0; k < 31; k++) {  // r[i] = r[i-31] + r[i-3]; compile-time ring indices keep the state in registers
            r[k] = r[k] + r[(k + 28) % 31];
            out[base + k] = (int)(r[k] >> 1);
        }
    }
#pragma unroll
    for (int k = 0; k < 31; k++) st->ring[k] = r[k];
    st->pos = 0;
}

// ------------------------------------------------------------------------------------------------
// K9 — pick: one thread per speculative step j reads its three draws (pick, theta, phi), does
// pick_random (lower_bound on the cumulative table, aggregat_list.cpp:59-66) and random_direction.
// ------------------------------------------------------------------------------------------------
// random_direction() for the draws at buffer positions (p, p + 1): from the host-evaluated table in strict replay mode
__device__ __forceinline__ Vec3 dev_direction(const DevState &d, long long p) {
    if (d.dir_tab) return Vec3{d.dir_tab[3 * p], d.dir_tab[3 * p + 1], d.dir_tab[3 * p + 2]};
    return direction_from_draws(uniform_from_rand(d.rng_buf[p]), uniform_from_rand(d.rng_buf[p + 1]));
}
__device__ __forceinline__ void dev_prepare_query(const DevState &d, int j, int *q_slot, double *q_dir, double *q_dist) {
    Scalars &sc = *d.sc;
    const long long p = sc.rand_pos + 3LL * j - d.rng_buf_base;
    if (p < 0 || p + 2 >= d.rng_buf_n) {  // the host did not stage these draws: refuse instead of reading outside the buffer
        sc.error = 1; sc.error_detail = DETAIL_RNG_NOT_STAGED;
        q_slot[j] = -1;
        q_dir[3 * j] = q_dir[3 * j + 1] = q_dir[3 * j + 2] = 0.;
        q_dist[j] = 0.;
        return;
    }
    const double u_pick = uniform_from_rand(d.rng_buf[p]);
    const int n = sc.n_pick;
    const double val = u_pick * d.cum[n - 1];
    int lo = 0, hi = n;  // std::lower_bound: first index with cum[i] >= val (every i < lo has cum[i] < val, every i >= hi cum[i] >= val)
    const int df = sc.pick_dense_from;
    if (df > 0 && df < n) {
        // Tie-dominated table: behind the sparse head the table grows by the same weight per entry, so the answer is within a few
        // entries of an interpolated guess.  Two probes around the guess only narrow [lo, hi) when they confirm it; the search below
        // is the same lower_bound on whatever range is left (exact for any guess).
        const double c0 = d.cum[df - 1];
        if (val > c0) {
            lo = df;
            const double g = (val - c0) / sc.pick_w;
            const long long p = (long long)df + (g < 2.0e9 ? (long long)g : 2000000000LL);
            const int q1 = (int)max((long long)df - 1, min((long long)n - 1, p - 3)), q2 = (int)max((long long)df - 1, min((long long)n - 1, p + 3));
            const double v1 = d.cum[q1], v2 = d.cum[q2];
            if (v1 < val) lo = max(lo, q1 + 1);
            if (v2 >= val) hi = min(hi, q2);
        } else hi = df - 1;
    }
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (d.cum[mid] < val) lo = mid + 1; else hi = mid;
    }
    int slot = d.sorted_slot[lo];
    if (n < 1 || lo >= n || slot < 0 || slot >= sc.n_agg_slots) {  // corrupt pick table
        sc.error = 1; sc.error_detail = DETAIL_PICK_TABLE;
        slot = -1;
    }
    const Vec3 dir = dev_direction(d, p + 1);
    q_slot[j] = slot;
    q_dir[3 * j] = dir.x;
    q_dir[3 * j + 1] = dir.y;
    q_dir[3 * j + 2] = dir.z;
    q_dist[j] = slot >= 0 ? d.a_lpm[slot] : 0.;
}
__global__ void k_prepare_queries(DevState d, int nq, int *q_slot, double *q_dir, double *q_dist) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nq) return;
    dev_prepare_query(d, j, q_slot, q_dir, q_dist);
}
// explicit queries given by label (the per-call C ABI)
__global__ void k_labels_to_slots(DevState d, int nq, const long long *labels, int *q_slot) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nq) return;
    const long long l = labels[j];
    q_slot[j] = (l >= 0 && l < d.sc->n_agg) ? d.slot_of_label[l] : -1;
}

// ------------------------------------------------------------------------------------------------
// K1 — swept-sphere contact search, one CTA per query.
// Replaces AggregatList::distance_to_next_contact + get_neighborhood + filter_neighborhood +
// distance_to_contact x3 (aggregat_list.cpp:447-532, aggregat_distance.cpp:24-44, sphere_contact.cpp:47-139).
//  phase 1: every aggregate of the Verlet cells along the displacement (verlet.cpp:52-98) gets the
//           bounding-sphere sweep test; the eligible ones (d_b < distance, strict) are kept with the
//           reference's visiting key = (cell scan rank, label).
//  phase 2: one warp per eligible aggregate sweeps moving spheres x other spheres with the exact FP64 pair
//           test and a warp-shuffle argmin whose tie-break is the (i, j) visiting order.
//  phase 3: the suspects are ranked by (d_b, key) — the multimap order — and scanned with the reference's
//           two early breaks, so zero-distance / stale-rmax corner cases resolve exactly as on the CPU.
// ------------------------------------------------------------------------------------------------
// Scratch of the three-kernel form of ONE search whose aggregates hold many spheres (late DLCA stages: 10^2..10^4 spheres per
// aggregate): phase 1 (k_search_big_p1) leaves the eligible suspects here, the sphere-sphere sweep of all of them is cut into
// tiles of (moving sphere, other sphere) pairs spread over the whole grid (k_search_big_p2), phase 3 (k_search_big_p3) reduces the
// tiles in pair order and runs the reference's ordered scan.
constexpr long long kPruneMinPairs = 4096;  // sphere pairs of two aggregates from which the ordered sweep prunes with enclosing balls
constexpr int kBigTiles = 8192;
struct BigSearch {
    int m, m_all, nb, pad;
    long long tile_pairs, n_tiles;
    int c_slot[kCandCap];
    double c_db[kCandCap];
    unsigned long long c_key[kCandCap];
    long long tile_begin[kCandCap + 1];
    double t_best[kBigTiles];
    long long t_pair[kBigTiles];
};
// kPhase 0: the whole search in this CTA; 1: phase 1 only -> `big`; 3: phases 2' (tile reduction) + 3 from `big`.  NT = threads of the
// CTA.  kOrdered (with kPhase 0): the eligible suspects are ranked in the multimap order FIRST and swept in that order by the whole CTA
// with the reference's two breaks (aggregat_list.cpp:459-482), so suspects the reference never examines are never swept — the form the
// per-realization step loop uses, where a search between 10^2..10^3-sphere aggregates is most of a step.
template <int kPhase, int NT = kSearchThreads, bool kOrdered = false>
__device__ __forceinline__ void search_wide_one(const DevState &d, int q, const int *__restrict__ q_slot,
                                                const double *__restrict__ q_dir, const double *__restrict__ q_dist,
                                                SearchResult *__restrict__ out, BigSearch *__restrict__ big = nullptr,
                                                int *list_i = nullptr, int *list_j = nullptr /* kOrdered: >= spheres of an aggregate each */) {
    __shared__ int seg_beg[NT];
    __shared__ int seg_pre[NT + 1];
    __shared__ int warp_sums[32];
    __shared__ int c_slot[kCandCap];
    __shared__ double c_db[kCandCap];
    __shared__ unsigned long long c_key[kCandCap];
    __shared__ double c_d[kCandCap];
    __shared__ long long c_pair[kCandCap];
    __shared__ int c_order[kCandCap];
    __shared__ int s_count, s_nb;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int slot = q_slot[q];
    SearchResult res;
    res.distance = INFINITY;
    res.moving_slot = res.other_slot = res.other_agg = -1;
    res.n_bounding = 0;
    res.n_sphere_pairs = 0;
    res.status = 0;
    res.pad = 0;
    if (slot < 0) {
        if (tid == 0) { res.status = 3; out[q] = res; }  // VerletError
        return;
    }
    const Scalars &sc = *d.sc;
    const double box = sc.box_length;
    const int n_div = d.n_div;
    const double4 me = d.a_posr[slot];
    const double dist = q_dist[q];
    const double dx = q_dir[3 * q], dy = q_dir[3 * q + 1], dz = q_dir[3 * q + 2];
    const CellRange rg = verlet_range(me.x, me.y, me.z, dist * dx, dist * dy, dist * dz, me.w + sc.maxradius, n_div, box);
    const int ni = rg.hi[0] - rg.lo[0] + 1, nj = rg.hi[1] - rg.lo[1] + 1, nk = rg.hi[2] - rg.lo[2] + 1;
    if (tid == 0) { s_count = 0; s_nb = 0; }
    __syncthreads();

    if (kPhase != 3) {
    // ---- phase 1: rows (i, j) of the cell range; along k the CSR entries are contiguous (<= 2 segments with wrap)
    const int ks = wrap_cell(rg.lo[2], n_div);
    const bool wraps = ks + nk > n_div;
    const int nseg = ni * nj * (wraps ? 2 : 1);
    for (int seg0 = 0; seg0 < nseg; seg0 += NT) {
        int len = 0, beg = 0;
        const int s = seg0 + tid;
        if (s < nseg) {
            const int row = wraps ? (s >> 1) : s;
            const int part = wraps ? (s & 1) : 0;
            const int ii = wrap_cell(rg.lo[0] + row / nj, n_div), jj = wrap_cell(rg.lo[1] + row % nj, n_div);
            const int base = (ii * n_div + jj) * n_div;
            int ka, kb;  // inclusive wrapped k span of this segment
            if (!wraps) { ka = ks; kb = ks + nk - 1; }
            else if (part == 0) { ka = ks; kb = n_div - 1; }
            else { ka = 0; kb = nk - (n_div - ks) - 1; }
            beg = d.cell_start[base + ka];
            len = d.cell_start[base + kb + 1] - beg;
        }
        int total;
        const int pre = block_exclusive_scan(len, &total, warp_sums);
        seg_beg[tid] = beg;
        seg_pre[tid] = pre;
        if (tid == 0) seg_pre[NT] = total;
        __syncthreads();
        for (int f = tid; f < total; f += NT) {
            int lo = 0, hi = NT;  // last segment whose prefix <= f
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (seg_pre[mid] <= f) lo = mid; else hi = mid;
            }
            const int o = d.cell_items[seg_beg[lo] + (f - seg_pre[lo])];
            if (o == slot) continue;
            const double4 oa = d.a_posr[o];
            const double db = pair_contact_distance(me.x, me.y, me.z, me.w, oa.x, oa.y, oa.z, oa.w, dx, dy, dz, dist, box);
            atomicAdd(&s_nb, 1);
            if (db < dist) {
                const int pos = atomicAdd(&s_count, 1);
                if (pos < d.cand_cap) {
                    const int r0 = range_rank(d.a_cx[o], rg.lo[0], rg.hi[0], n_div);
                    const int r1 = range_rank(d.a_cy[o], rg.lo[1], rg.hi[1], n_div);
                    const int r2 = range_rank(d.a_cz[o], rg.lo[2], rg.hi[2], n_div);
                    const unsigned long long cell_rank = (unsigned long long)((r0 * nj + r1) * (long long)nk + r2);
                    c_slot[pos] = o;
                    c_db[pos] = db;
                    c_key[pos] = (cell_rank << 32) | (unsigned)o;
                }
            }
        }
        __syncthreads();
    }
    } else {  // the eligible suspects were left in `big` by phase 1
        for (int t = tid; t < big->m; t += NT) { c_slot[t] = big->c_slot[t]; c_db[t] = big->c_db[t]; c_key[t] = big->c_key[t]; }
        if (tid == 0) { s_count = big->m_all; s_nb = big->nb; }
        __syncthreads();
    }
    const int m_all = s_count;
    const int m = m_all < d.cand_cap ? m_all : d.cand_cap;
    if (m_all > d.cand_cap) res.status = 1;  // more eligible suspects than the shared-memory list holds: the first contact is unknown

    const int n_src = d.a_n[slot], off_src = d.a_off[slot];
    if (kPhase == 0 && kOrdered) {
        __shared__ double o_best[32];
        __shared__ long long o_pair[32];
        for (int t = tid; t < m; t += NT) {
            const double db = c_db[t];
            const unsigned long long key = c_key[t];
            int rank = 0;
            for (int u = 0; u < m; u++) {
                const double du = c_db[u];
                rank += (du < db || (du == db && c_key[u] < key)) ? 1 : 0;
            }
            c_order[rank] = t;
        }
        __syncthreads();
        double closest = INFINITY;  // the scan state is kept identically by every thread
        int who = -1;
        long long who_pair = 0, examined = 0, executed = 0;
        for (int r = 0; r < m; r++) {
            const int t = c_order[r];
            if (closest <= 0.) break;
            if (closest < c_db[t]) break;
            const int o = c_slot[t];
            const int n_o = d.a_n[o], off_o = d.a_off[o];
            const long long npairs = (long long)n_src * n_o;
            double best = INFINITY;
            long long best_p = npairs;
            if (list_i && npairs >= kPruneMinPairs) {
                // Pruned sweep (same result): only moving spheres that can reach the other aggregate's enclosing ball and other spheres
                // the moving aggregate's enclosing ball can reach are paired (sweep_may_touch, mcac_math.cuh: a conservative
                // necessary condition for a finite pair distance).  The balls are taken about the aggregate centres with radii
                // recomputed from the relative positions (not the stored rmax, which growth can leave stale).
                __shared__ double o_rad[2][32];
                __shared__ int s_ni, s_nj;
                const double mrx = d.a_rx[slot], mry = d.a_ry[slot], mrz = d.a_rz[slot];
                const double orx = d.a_rx[o], ory = d.a_ry[o], orz = d.a_rz[o];
                double rm = 0., ro = 0.;
                for (int i2 = tid; i2 < n_src; i2 += NT) {
                    const double4 rel = d.s_relv[off_src + i2];
                    const double ex = rel.x - mrx, ey = rel.y - mry, ez = rel.z - mrz;
                    const double t = sqrt(ex * ex + ey * ey + ez * ez) + d.s_posr[off_src + i2].w;
                    rm = t > rm ? t : rm;
                }
                for (int j2 = tid; j2 < n_o; j2 += NT) {
                    const double4 rel = d.s_relv[off_o + j2];
                    const double ex = rel.x - orx, ey = rel.y - ory, ez = rel.z - orz;
                    const double t = sqrt(ex * ex + ey * ey + ez * ez) + d.s_posr[off_o + j2].w;
                    ro = t > ro ? t : ro;
                }
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) {
                    const double t0 = __shfl_xor_sync(kFull, rm, sft), t1 = __shfl_xor_sync(kFull, ro, sft);
                    rm = t0 > rm ? t0 : rm;
                    ro = t1 > ro ? t1 : ro;
                }
                if (lane == 0) { o_rad[0][warp] = rm; o_rad[1][warp] = ro; }
                if (tid == 0) { s_ni = 0; s_nj = 0; }
                __syncthreads();
                rm = o_rad[0][0]; ro = o_rad[1][0];
                for (int w = 1; w < nwarps; w++) { rm = o_rad[0][w] > rm ? o_rad[0][w] : rm; ro = o_rad[1][w] > ro ? o_rad[1][w] : ro; }
                const double4 oc = d.a_posr[o];
                for (int i2 = tid; i2 < n_src; i2 += NT) {
                    const double4 p = d.s_posr[off_src + i2];
                    if (sweep_may_touch(p.x, p.y, p.z, p.w, oc.x, oc.y, oc.z, ro, dx, dy, dz, dist, box)) list_i[atomicAdd(&s_ni, 1)] = i2;
                }
                for (int j2 = tid; j2 < n_o; j2 += NT) {
                    const double4 p = d.s_posr[off_o + j2];
                    if (sweep_may_touch(me.x, me.y, me.z, rm, p.x, p.y, p.z, p.w, dx, dy, dz, dist, box)) list_j[atomicAdd(&s_nj, 1)] = j2;
                }
                __syncthreads();
                const int ni = s_ni, nj = s_nj;
                const long long nsel = (long long)ni * nj;
                if (nj > 0) {
                    const int qs = NT / nj, rs = NT % nj;
                    int ii = tid / nj, jj = tid % nj;
                    for (long long p = tid; p < nsel; p += NT) {
                        const int i2 = list_i[ii], j2 = list_j[jj];
                        const double4 a = d.s_posr[off_src + i2];
                        const double4 b = d.s_posr[off_o + j2];
                        const double c = pair_contact_distance(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, dx, dy, dz, dist, box);
                        const long long pp = (long long)i2 * n_o + j2;
                        if (c < best || (c == best && pp < best_p)) { best = c; best_p = pp; }  // the lists are unordered
                        jj += rs; ii += qs;
                        if (jj >= nj) { jj -= nj; ii++; }
                    }
                }
                executed += nsel + n_src + n_o;
                __syncthreads();  // the lists are reused by the next suspect
            } else {
            // pair p = i * n_o + j (moving sphere i outer, other sphere j inner: the reference's visiting order), NT pairs at a time
            const int qs = NT / n_o, rs = NT % n_o;
            int i = tid / n_o, j = tid % n_o;
            for (long long p = tid; p < npairs; p += NT) {
                const double4 a = d.s_posr[off_src + i];
                const double4 b = d.s_posr[off_o + j];
                const double c = pair_contact_distance(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, dx, dy, dz, dist, box);
                if (c < best) { best = c; best_p = p; }
                j += rs; i += qs;
                if (j >= n_o) { j -= n_o; i++; }
            }
            executed += npairs;
            }
            warp_argmin(best, best_p);
            if (lane == 0) { o_best[warp] = best; o_pair[warp] = best_p; }
            __syncthreads();
            best = o_best[0]; best_p = o_pair[0];
            for (int w = 1; w < nwarps; w++)
                if (o_best[w] < best || (o_best[w] == best && o_pair[w] < best_p)) { best = o_best[w]; best_p = o_pair[w]; }
            __syncthreads();
            examined += npairs;
            if (best < closest) { closest = best; who = t; who_pair = best_p; }
        }
        if (tid == 0) {
            res.n_bounding = s_nb;
            res.n_sphere_pairs = examined;
            if (who >= 0) {
                const int o = c_slot[who];
                const int n_o = d.a_n[o];
                res.distance = closest;
                res.moving_slot = off_src + (int)(who_pair / n_o);
                res.other_slot = d.a_off[o] + (int)(who_pair % n_o);
                res.other_agg = o;
            }
            out[q] = res;
            d.sc->pair_exec += executed;
        }
        return;
    }
    if (kPhase == 1) {  // hand the suspects and the tile partition of their sphere pairs to the grid-wide sweep
        __syncthreads();
        for (int t = tid; t < m; t += NT) { big->c_slot[t] = c_slot[t]; big->c_db[t] = c_db[t]; big->c_key[t] = c_key[t]; }
        if (tid == 0) {
            long long total = 0;
            for (int t = 0; t < m; t++) total += (long long)n_src * d.a_n[c_slot[t]];
            // sum over suspects of ceil(pairs / tile) <= total / tile + m: sized so that it never exceeds kBigTiles
            long long tile = (total + (kBigTiles - kCandCap) - 1) / (kBigTiles - kCandCap);
            if (tile < 2048) tile = 2048;
            long long acc = 0;
            for (int t = 0; t < m; t++) {
                big->tile_begin[t] = acc;
                acc += ((long long)n_src * d.a_n[c_slot[t]] + tile - 1) / tile;
            }
            big->tile_begin[m] = acc;
            big->n_tiles = acc;
            big->tile_pairs = tile;
            big->m = m; big->m_all = m_all; big->nb = s_nb;
        }
        return;
    }
    // ---- phase 2: exact sphere-sphere sweep, one warp per eligible aggregate
    if (kPhase == 3) {  // tiles of one suspect are in ascending pair order: strict `<` keeps the first minimum
        for (int k = tid; k < m; k += NT) {
            double best = INFINITY;
            long long best_p = (long long)n_src * d.a_n[c_slot[k]];
            for (long long t = big->tile_begin[k]; t < big->tile_begin[k + 1]; t++)
                if (big->t_best[t] < best) { best = big->t_best[t]; best_p = big->t_pair[t]; }
            c_d[k] = best;
            c_pair[k] = best_p;
        }
    } else
    for (int k = warp; k < m; k += nwarps) {
        const int o = c_slot[k];
        const int n_o = d.a_n[o], off_o = d.a_off[o];
        const long long npairs = (long long)n_src * n_o;
        double best = INFINITY;
        long long best_p = npairs;
        for (long long p = lane; p < npairs; p += 32) {
            const int i = (int)(p / n_o), j = (int)(p - (long long)i * n_o);
            const double4 a = d.s_posr[off_src + i];
            const double4 b = d.s_posr[off_o + j];
            const double c = pair_contact_distance(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, dx, dy, dz, dist, box);
            if (c < best) { best = c; best_p = p; }
        }
        warp_argmin(best, best_p);
        if (lane == 0) { c_d[k] = best; c_pair[k] = best_p; }
    }
    __syncthreads();

    // ---- phase 3: multimap order + the reference's scan with its two breaks (aggregat_list.cpp:459-482)
    for (int t = tid; t < m; t += NT) {
        const double db = c_db[t];
        const unsigned long long key = c_key[t];
        int rank = 0;
        for (int u = 0; u < m; u++) {
            const double du = c_db[u];
            rank += (du < db || (du == db && c_key[u] < key)) ? 1 : 0;
        }
        c_order[rank] = t;
    }
    __syncthreads();
    if (tid == 0) {
        double closest = INFINITY;
        int who = -1;
        long long examined = 0;
        for (int r = 0; r < m; r++) {
            const int t = c_order[r];
            if (closest <= 0.) break;
            if (closest < c_db[t]) break;
            examined += (long long)n_src * d.a_n[c_slot[t]];
            if (c_d[t] < closest) { closest = c_d[t]; who = t; }
        }
        res.n_bounding = s_nb;
        res.n_sphere_pairs = examined;
        if (who >= 0) {
            const int o = c_slot[who];
            const int n_o = d.a_n[o];
            const long long p = c_pair[who];
            res.distance = closest;
            res.moving_slot = off_src + (int)(p / n_o);
            res.other_slot = d.a_off[o] + (int)(p % n_o);
            res.other_agg = o;
        }
        out[q] = res;
    }
}
// wide form: one CTA per listed query (the searches the group kernel hands over: > kGroupCap eligible suspects or
// aggregate pairs with many sphere pairs); list == nullptr runs every query (debug / comparison)
__global__ void __launch_bounds__(kSearchThreads) k_search_wide(DevState d, int nq, const int *__restrict__ list, const int *__restrict__ count,
                                                                const int *__restrict__ q_slot, const double *__restrict__ q_dir,
                                                                const double *__restrict__ q_dist, SearchResult *__restrict__ out) {
    const int n = list ? *count : nq;
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        search_wide_one<0>(d, list ? list[e] : e, q_slot, q_dir, q_dist, out);
        __syncthreads();
    }
}

// the three-kernel form (one query, q = 0)
__global__ void __launch_bounds__(kSearchThreads) k_search_big_p1(DevState d, const int *__restrict__ q_slot, const double *__restrict__ q_dir,
                                                                  const double *__restrict__ q_dist, SearchResult *__restrict__ out, BigSearch *big) {
    if (threadIdx.x == 0) { big->m = 0; big->n_tiles = 0; }
    __syncthreads();
    search_wide_one<1>(d, 0, q_slot, q_dir, q_dist, out, big);
}
__global__ void __launch_bounds__(256) k_search_big_p2(DevState d, const int *__restrict__ q_slot, const double *__restrict__ q_dir,
                                                       const double *__restrict__ q_dist, BigSearch *big) {
    __shared__ double sb[8];
    __shared__ long long sp[8];
    const int slot = q_slot[0];
    if (slot < 0) return;
    const int m = big->m;
    const long long n_tiles = big->n_tiles, tile = big->tile_pairs;
    const double box = d.sc->box_length, dist = q_dist[0], dx = q_dir[0], dy = q_dir[1], dz = q_dir[2];
    const int n_src = d.a_n[slot], off_src = d.a_off[slot];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int lo = 0, hi = m;  // suspect of this tile: last k with tile_begin[k] <= t
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (big->tile_begin[mid] <= t) lo = mid; else hi = mid;
        }
        const int o = big->c_slot[lo];
        const int n_o = d.a_n[o], off_o = d.a_off[o];
        const long long npairs = (long long)n_src * n_o;
        const long long p0 = (t - big->tile_begin[lo]) * tile, p1 = (p0 + tile < npairs) ? p0 + tile : npairs;
        double best = INFINITY;
        long long best_p = npairs;
        for (long long p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
            const int i = (int)(p / n_o), j = (int)(p - (long long)i * n_o);
            const double4 a = d.s_posr[off_src + i];
            const double4 b = d.s_posr[off_o + j];
            const double c = pair_contact_distance(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, dx, dy, dz, dist, box);
            if (c < best) { best = c; best_p = p; }
        }
        warp_argmin(best, best_p);
        if (lane == 0) { sb[warp] = best; sp[warp] = best_p; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); w++)
                if (sb[w] < best || (sb[w] == best && sp[w] < best_p)) { best = sb[w]; best_p = sp[w]; }
            big->t_best[t] = best;
            big->t_pair[t] = best_p;
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kSearchThreads) k_search_big_p3(DevState d, const int *__restrict__ q_slot, const double *__restrict__ q_dir,
                                                                  const double *__restrict__ q_dist, SearchResult *__restrict__ out, BigSearch *big) {
    search_wide_one<3>(d, 0, q_slot, q_dir, q_dist, out, big);
}

// ------------------------------------------------------------------------------------------------
// K1 (main form) — one GROUP of G lanes (a quarter, half or whole warp) per query; no block-level barriers, no
// shared-memory atomics.  Same three phases and the same visiting order as above:
//  phase 1: the group's lanes take the (i, j) rows of the Verlet range, scan their lengths with shuffles and sweep
//           the concatenated candidate list G at a time.  The bounding spheres are read from `cell_posr`, a copy of
//           (x, y, z, rmax) laid out in CELL ORDER by K2, so the candidates of a row are one contiguous 32-byte-vector
//           stream instead of a gather over the aggregate table.
//  phase 3 before phase 2: the eligible suspects (<= kGroupCap) are ranked in the multimap order first, then the
//           sphere-sphere sweeps run in that order with the reference's two early breaks — suspects the reference
//           never examines are never swept.
// Queries whose suspect list overflows or whose aggregate pairs are large are appended to `wide_list`.
// ------------------------------------------------------------------------------------------------
constexpr int kGroupCap = 32;
constexpr long long kGroupMaxPairs = 4096;
template <int G>
__device__ __forceinline__ void group_argmin(double &dmin, long long &idx, unsigned gmask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(gmask, dmin, o);
        const long long oi = __shfl_xor_sync(gmask, idx, o);
        if (od < dmin || (od == dmin && oi < idx)) { dmin = od; idx = oi; }
    }
}
struct GroupHit {
    double closest;
    int who;
    long long who_pair, examined;
};
// phases 3 + 2 of one query for its group of G lanes; returns true when the query must go to the wide kernel
template <int G>
__device__ __noinline__ bool group_sweep_suspects(const int *__restrict__ a_n, const int *__restrict__ a_off,
                                                  const double4 *__restrict__ s_posr, int slot, int m, const int *c_slot, const double *c_db,
                                                  const unsigned long long *c_key, unsigned char *c_order, double dx, double dy, double dz,
                                                  double dist, double box, unsigned gmask, int gl, GroupHit &hit) {
    __syncwarp(gmask);
    // ---- phase 3 first: rank in multimap order (bounding distance, then cell scan rank, then label)
    for (int t = gl; t < m; t += G) {
        const double db = c_db[t];
        const unsigned long long key = c_key[t];
        int rank = 0;
        for (int u = 0; u < m; u++) {
            const double du = c_db[u];
            rank += (du < db || (du == db && c_key[u] < key)) ? 1 : 0;
        }
        c_order[rank] = (unsigned char)t;
    }
    __syncwarp(gmask);
    // ---- phase 2 in that order, with the reference's two breaks (aggregat_list.cpp:459-482)
    const int n_src = a_n[slot], off_src = a_off[slot];
    double closest = INFINITY;
    int who = -1;
    long long who_pair = 0, examined = 0;
    for (int r = 0; r < m; r++) {
        const int t = c_order[r];
        if (closest <= 0.) break;
        if (closest < c_db[t]) break;
        const int o = c_slot[t];
        const int n_o = a_n[o], off_o = a_off[o];
        const long long npairs = (long long)n_src * n_o;
        if (npairs > kGroupMaxPairs) return true;
        double best = INFINITY;
        long long best_p = npairs;
        for (long long p = gl; p < npairs; p += G) {
            const int i = (int)(p / n_o), j = (int)(p - (long long)i * n_o);
            const double4 a = s_posr[off_src + i];
            const double4 b = s_posr[off_o + j];
            const double c = pair_contact_distance(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, dx, dy, dz, dist, box);
            if (c < best) { best = c; best_p = p; }
        }
        group_argmin<G>(best, best_p, gmask);
        examined += npairs;
        if (best < closest) { closest = best; who = t; who_pair = best_p; }
    }
    hit.closest = closest;
    hit.who = who;
    hit.who_pair = who_pair;
    hit.examined = examined;
    return false;
}
template <int G, int kMinBlocks>
__global__ void __launch_bounds__(kSearchThreads, kMinBlocks) k_search_group(DevState d, int nq, const int *__restrict__ q_slot,
                                                                 const double *__restrict__ q_dir, const double *__restrict__ q_dist,
                                                                 SearchResult *__restrict__ out, int *__restrict__ wide_list,
                                                                 int *__restrict__ wide_count, int *__restrict__ next_count) {
    constexpr int GPB = kSearchThreads / G;
    if (blockIdx.x == 0 && threadIdx.x == 0) *next_count = 0;
    __shared__ int c_slot[GPB][kGroupCap];
    __shared__ double c_db[GPB][kGroupCap];
    __shared__ unsigned long long c_key[GPB][kGroupCap];
    __shared__ unsigned char c_order[GPB][kGroupCap];
    const int g = threadIdx.x / G, gl = threadIdx.x % G;
    const int q = blockIdx.x * GPB + g;
    if (q >= nq) return;  // whole groups leave together
    const unsigned gmask = (G == 32) ? kFull : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
    const unsigned lane_lt = (1u << (threadIdx.x & 31)) - 1u;
    const int slot = q_slot[q];
    SearchResult res;
    res.distance = INFINITY;
    res.moving_slot = res.other_slot = res.other_agg = -1;
    res.n_bounding = 0;
    res.n_sphere_pairs = 0;
    res.status = 0;
    res.pad = 0;
    if (slot < 0) {
        if (gl == 0) { res.status = 3; out[q] = res; }  // VerletError
        return;
    }
    const Scalars &sc = *d.sc;
    const double box = sc.box_length;
    const int n_div = d.n_div;
    const double4 me = d.a_posr[slot];
    const double dist = q_dist[q];
    const double dx = q_dir[3 * q], dy = q_dir[3 * q + 1], dz = q_dir[3 * q + 2];
    const CellRange rg = verlet_range(me.x, me.y, me.z, dist * dx, dist * dy, dist * dz, me.w + sc.maxradius, n_div, box);
    const int ni = rg.hi[0] - rg.lo[0] + 1, nj = rg.hi[1] - rg.lo[1] + 1, nk = rg.hi[2] - rg.lo[2] + 1;

    // ---- phase 1
    const int ks = wrap_cell(rg.lo[2], n_div);
    const bool wraps = ks + nk > n_div;
    const int nseg = ni * nj * (wraps ? 2 : 1);
    int m_all = 0, nb = 0;
    for (int seg0 = 0; seg0 < nseg; seg0 += G) {
        int len = 0, beg = 0;
        const int s = seg0 + gl;
        if (s < nseg) {
            const int row = wraps ? (s >> 1) : s;
            const int part = wraps ? (s & 1) : 0;
            const int ii = wrap_cell(rg.lo[0] + row / nj, n_div), jj = wrap_cell(rg.lo[1] + row % nj, n_div);
            const int base = (ii * n_div + jj) * n_div;
            int ka, kb;  // inclusive wrapped k span of this segment
            if (!wraps) { ka = ks; kb = ks + nk - 1; }
            else if (part == 0) { ka = ks; kb = n_div - 1; }
            else { ka = 0; kb = nk - (n_div - ks) - 1; }
            beg = d.cell_start[base + ka];
            len = d.cell_start[base + kb + 1] - beg;
        }
        int incl = len;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            const int t = __shfl_up_sync(gmask, incl, o, G);
            if (gl >= o) incl += t;
        }
        const int excl = incl - len;
        const int total = __shfl_sync(gmask, incl, G - 1, G);
        for (int f0 = 0; f0 < total; f0 += G) {
            const int f = f0 + gl;
            int L = 0;  // last segment whose exclusive prefix <= f
#pragma unroll
            for (int step = G / 2; step > 0; step >>= 1) {
                const int t = __shfl_sync(gmask, excl, L + step, G);
                if (t <= f) L += step;
            }
            const int sb = __shfl_sync(gmask, beg, L, G), se = __shfl_sync(gmask, excl, L, G);
            bool elig = false;
            double db = 0.;
            int o = -1;
            if (f < total) {
                const int idx = sb + (f - se);
                o = d.cell_items[idx];
                const double4 oa = d.cell_posr[idx];  // issued with the id load, not behind it
                if (o != slot) {
                    db = pair_contact_distance(me.x, me.y, me.z, me.w, oa.x, oa.y, oa.z, oa.w, dx, dy, dz, dist, box);
                    nb++;
                    elig = db < dist;
                }
            }
            const unsigned bal = __ballot_sync(gmask, elig) & gmask;
            if (elig) {
                const int pos = m_all + __popc(bal & lane_lt);
                if (pos < kGroupCap) {
                    const int r0 = range_rank(d.a_cx[o], rg.lo[0], rg.hi[0], n_div);
                    const int r1 = range_rank(d.a_cy[o], rg.lo[1], rg.hi[1], n_div);
                    const int r2 = range_rank(d.a_cz[o], rg.lo[2], rg.hi[2], n_div);
                    const unsigned long long cell_rank = (unsigned long long)((r0 * nj + r1) * (long long)nk + r2);
                    c_slot[g][pos] = o;
                    c_db[g][pos] = db;
                    c_key[g][pos] = (cell_rank << 32) | (unsigned)o;
                }
            }
            m_all += __popc(bal);
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) nb += __shfl_xor_sync(gmask, nb, o);
    bool wide = m_all > kGroupCap;
    GroupHit hit;
    hit.closest = INFINITY;
    hit.who = -1;
    hit.who_pair = 0;
    hit.examined = 0;
    if (m_all > 0 && !wide)  // rare: kept out of line so that the sweep above keeps a small register footprint
        wide = group_sweep_suspects<G>(d.a_n, d.a_off, d.s_posr, slot, m_all, c_slot[g], c_db[g], c_key[g], c_order[g], dx, dy, dz, dist, box, gmask, gl, hit);
    if (gl != 0) return;
    if (wide) {
        wide_list[atomicAdd(wide_count, 1)] = q;
        return;
    }
    res.n_bounding = nb;
    res.n_sphere_pairs = hit.examined;
    if (hit.who >= 0) {
        const int o = c_slot[g][hit.who];
        const int n_o = d.a_n[o];
        res.distance = hit.closest;
        res.moving_slot = d.a_off[slot] + (int)(hit.who_pair / n_o);
        res.other_slot = d.a_off[o] + (int)(hit.who_pair % n_o);
        res.other_agg = o;
    }
    out[q] = res;
}

// ------------------------------------------------------------------------------------------------
// K2 — Verlet cell list by counting sort (replaces Verlet ctor/add/remove, verlet.cpp:26-51, and
// Aggregate::update_verlet bookkeeping): histogram of the stored cell of each live aggregate ->
// exclusive scan -> scatter.  The in-cell order is irrelevant: K1 re-derives the reference's visiting key.
// ------------------------------------------------------------------------------------------------
__global__ void k_cell_count(DevState d) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.sc->n_agg_slots || !d.a_alive[s]) return;
    const int c = (d.a_cx[s] * d.n_div + d.a_cy[s]) * d.n_div + d.a_cz[s];
    atomicAdd(&d.cell_fill[c], 1);
}
__global__ void k_cell_scatter(DevState d) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.sc->n_agg_slots || !d.a_alive[s]) return;
    const int c = (d.a_cx[s] * d.n_div + d.a_cy[s]) * d.n_div + d.a_cz[s];
    const int pos = d.cell_start[c] + atomicAdd(&d.cell_fill[c], 1);
    d.cell_items[pos] = s;
    d.cell_posr[pos] = d.a_posr[s];  // bounding sphere in cell order: K1 streams it
}
// generic 3-phase exclusive scan of ints (deterministic): in[0..n) -> out[0..n], out[n] = total
constexpr int kScanBlock = 1024;
constexpr int kScanItems = 4;
__global__ void k_scan_partials(const int *in, int n, int *block_sums) {
    __shared__ int ws[32];
    const int base = blockIdx.x * kScanBlock * kScanItems + threadIdx.x * kScanItems;
    int v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) v += (base + k < n) ? in[base + k] : 0;
    int total;
    block_exclusive_scan(v, &total, ws);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void k_scan_block_sums(int *block_sums, int nb) {  // single block; nb <= few thousand
    __shared__ int ws[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += blockDim.x) {
        const int i = b0 + threadIdx.x;
        const int v = i < nb ? block_sums[i] : 0;
        int total;
        const int pre = block_exclusive_scan(v, &total, ws);
        if (i < nb) block_sums[i] = carry + pre;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nb] = carry;
}
__global__ void k_scan_apply(const int *in, int n, const int *block_sums, int nb, int *out) {
    __shared__ int ws[32];
    const int base = blockIdx.x * kScanBlock * kScanItems + threadIdx.x * kScanItems;
    int x[kScanItems];
    int v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { x[k] = (base + k < n) ? in[base + k] : 0; v += x[k]; }
    int total;
    int pre = block_exclusive_scan(v, &total, ws) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = pre;
        pre += x[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_sums[nb];
}

// label <-> slot maps (label = rank of the slot among live slots): flags -> scan -> scatter
__global__ void k_alive_to_labels(DevState d, const int *scan /* exclusive scan of a_alive */) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.sc->n_agg_slots) return;
    if (d.a_alive[s]) {
        d.label_of_slot[s] = scan[s];
        d.slot_of_label[scan[s]] = s;
    } else {
        d.label_of_slot[s] = -1;
    }
}

// ------------------------------------------------------------------------------------------------
// Aggregate-level device functions used by commit / merge / growth kernels
// ------------------------------------------------------------------------------------------------
// Aggregate::set_position (aggregat.cpp:109-118): periodic wrap of the centre + stored Verlet cell
__device__ __forceinline__ void agg_store_position(const DevState &d, int slot, double x, double y, double z, double box) {
    const double nx = periodic_position(x, box), ny = periodic_position(y, box), nz = periodic_position(z, box);
    double4 a = d.a_posr[slot];
    a.x = nx; a.y = ny; a.z = nz;
    d.a_posr[slot] = a;
    d.a_cx[slot] = cell_of(nx, d.n_div, box);
    d.a_cy[slot] = cell_of(ny, d.n_div, box);
    d.a_cz[slot] = cell_of(nz, d.n_div, box);
}
constexpr int kDirtyFull = 1, kDirtyPartial = 2, kDirtyAll = 3;  // DevState::a_dirty (see agg_update)
template <bool kBlock>
__device__ __forceinline__ void group_sync() {
    if (kBlock) __syncthreads(); else __syncwarp();
}
// K3 — Aggregate::translate (aggregat.cpp:148-161), cooperative over the `nth` threads of a warp or of the CTA.
// Every thread recomputes the new centre (cheap); the centre itself is rewritten after a group barrier.
template <bool kBlock>
__device__ __forceinline__ void agg_translate(const DevState &d, int slot, double vx, double vy, double vz, double box, int tid, int nth) {
    const double4 a = d.a_posr[slot];
    const double nx = periodic_position(a.x + vx, box), ny = periodic_position(a.y + vy, box), nz = periodic_position(a.z + vz, box);
    const double refx = nx - d.a_rx[slot], refy = ny - d.a_ry[slot], refz = nz - d.a_rz[slot];
    const int off = d.a_off[slot], n = d.a_n[slot];
    for (int i = tid; i < n; i += nth) {
        const double4 rel = d.s_relv[off + i];
        double4 p = d.s_posr[off + i];
        p.x = refx + rel.x; p.y = refy + rel.y; p.z = refz + rel.z;  // root sphere: rel == 0 -> refpos
        d.s_posr[off + i] = p;
    }
    group_sync<kBlock>();
    if (tid == 0) {
        agg_store_position(d, slot, a.x + vx, a.y + vy, a.z + vz, box);
        atomicOr(&d.a_dirty[slot], kDirtyPartial);  // (result unused: no round trip)
    }
    group_sync<kBlock>();
}

// K5-K7 — Aggregate::update() / update_partial() (aggregat.cpp:247-288, 321-483, 719-764) for one aggregate,
// cooperative over the threads [0,nth) of a warp (kBlock = false) or of the whole CTA (kBlock = true).
// All 1-D sums (volume, surface, centre of mass, gyration sums, mean diameter) are accumulated by ONE lane in
// `myspheres` order, exactly like the reference's loops, so they come out bit-identical; the O(n^2) contact
// pass gives one sphere per lane (ascending partner index) and its overlap statistics are combined in a fixed
// tree (deterministic, <= 1e-15 relative from the reference's hash-map order).
// scratch: kUpdateScratch doubles of shared memory private to the group.
constexpr int kUpdateScratch = 192;  // [0,8) results, [8,40) per-warp maxima, [40,40+7*16) per-warp partial sums
// stage (optional): 5 * nth doubles of shared memory private to the group.  The ordered sums are sequential by definition (one add
// after the other in `myspheres` order); with `stage` the group computes the TERMS of a chunk of nth spheres in parallel, and the
// accumulating lanes only run the add chains over shared memory — same operations in the same order, without one global-memory
// round trip per sphere on the critical path (a 10^2..10^3-sphere aggregate of the classic.ini ensemble: ~100 us -> a few us).
template <bool kBlock>
__device__ void agg_update(const DevState &d, int slot, bool full, int tid, int nth, double *scratch, double box, double *stage = nullptr) {
    const int off = d.a_off[slot], n = d.a_n[slot];
    const int method = d.volsurf_method;
    const int w = tid >> 5, nw = (nth + 31) >> 5;
    // Aggregate::update() of an aggregate whose spheres did not change since its last full update recomputes, from the same relative
    // positions and radii, exactly the contact graph, overlap statistics, effective volumes / surfaces and V, S that are stored: the
    // O(n^2) pass is skipped (individual surface reactions grow one aggregate per step; the reference's loop updates all of them)
    // The same holds for update_partial(): an aggregate that has neither changed nor MOVED since its last update gets, from the same
    // inputs, exactly the centre, radii, friction and time step it already has (the first update after a move is kept: it re-derives
    // the aggregate position from the root sphere, which can differ from the translated one in the last bit).  With individual
    // surface reactions one aggregate per step changes, and the loop over all of them costs one flag read each.
    // a_dirty: kDirtyFull = spheres changed (contact pass due), kDirtyPartial = update_partial due.
    const int flags = d.a_dirty[slot];
    group_sync<kBlock>();  // everybody has read the flags before they are rewritten
    if (full && !(flags & kDirtyFull)) full = false;
    if (!full && !(flags & kDirtyPartial)) return;
    if (full) {
        // ---- contact pass (update_distances_and_overlapping + the contact-graph loops of compute_volume_surface)
        double vals[7] = {0., 0., 0., 0., 0., 0., 0.};  // intersections, sum c_ij, c_s10, c_v20, c_v30, vp_sum, sp_sum
        for (int i = tid; i < n; i += nth) {
            const double4 ri = d.s_relv[off + i];
            const double r_i = d.s_posr[off + i].w;
            const double s_i = d.s_surf[off + i];
            double veff = ri.w, seff = s_i;
            for (int j = 0; j < n; j++) {
                if (j == i) continue;
                const double4 rj = d.s_relv[off + j];
                const double r_j = d.s_posr[off + j].w;
                const double ex = ri.x - rj.x, ey = ri.y - rj.y, ez = ri.z - rj.z;
                const double d2 = ex * ex + ey * ey + ez * ez;
                const double rs = r_i + r_j;
                const double lim = (1. + kCoordinationEpsilon) * rs;
                // far apart (with a margin of 1e-12 that dwarfs the rounding of sqrt and of the products): no square root needed —
                // almost every pair of a 10^2..10^3-sphere aggregate; the reference's own test decides the others
                if (d2 > (lim * lim) * (1. + 1e-12)) continue;
                const double dist = sqrt(d2);
                if (dist <= lim) {
                    const double c_ij = (rs - dist) / rs;
                    vals[0] += 1.;
                    vals[1] += c_ij;
                    if (method == MCAC_VS_ALPHAS) {
                        const double vp = pow(r_i, 3.) + pow(r_j, 3.);
                        const double sp = r_i * r_i + r_j * r_j;
                        vals[5] += vp;
                        vals[6] += sp;
                        vals[2] += c_ij * sp;
                        vals[3] += (c_ij * c_ij) * vp;
                        vals[4] += pow(c_ij, 3.) * vp;
                    } else if (method == MCAC_VS_CAPS) {
                        double caps[4];
                        if (i < j) {
                            lens_caps(r_i, ri.w, s_i, r_j, rj.w, d.s_surf[off + j], dist, caps);
                            veff = veff - caps[0];
                            seff = seff - caps[2];
                        } else {
                            lens_caps(r_j, rj.w, d.s_surf[off + j], r_i, ri.w, s_i, dist, caps);
                            veff = veff - caps[1];
                            seff = seff - caps[3];
                        }
                    }
                }
            }
            if (method == MCAC_VS_CAPS) {
                veff = (veff < 0.0) ? 0.0 : veff;
                seff = (seff < 0.0) ? 0.0 : seff;
            }
            d.s_veff[off + i] = veff;
            d.s_seff[off + i] = seff;
        }
#pragma unroll
        for (int k = 0; k < 7; k++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) vals[k] += __shfl_xor_sync(kFull, vals[k], o);
        }
        if (kBlock) {
            if ((tid & 31) == 0)
                for (int k = 0; k < 7; k++) scratch[40 + w * 7 + k] = vals[k];
            __syncthreads();
            for (int k = 0; k < 7; k++) {
                double acc = 0.;
                for (int ww = 0; ww < nw; ww++) acc += scratch[40 + ww * 7 + k];
                vals[k] = acc;
            }
        }
        group_sync<kBlock>();  // s_veff / s_seff visible
        if (stage) {  // V and S: two add chains (lanes 0, 1) over staged chunks
            double acc = 0.;
            for (int base = 0; base < n; base += nth) {
                const int i = base + tid;
                if (i < n) { stage[tid] = d.s_veff[off + i]; stage[nth + tid] = d.s_seff[off + i]; }
                group_sync<kBlock>();
                if (tid < 2) {
                    const double *src = stage + tid * nth;
                    const int m = (n - base < nth) ? n - base : nth;
                    for (int k = 0; k < m; k++) acc = acc + src[k];
                }
                group_sync<kBlock>();
            }
            if (tid < 2) scratch[tid] = acc;
            group_sync<kBlock>();
        }
        if (tid == 0) {
            double V = 0., S = 0.;
            if (stage) { V = scratch[0]; S = scratch[1]; }
            else
            for (int i = 0; i < n; i++) {
                V = V + d.s_veff[off + i];
                S = S + d.s_seff[off + i];
            }
            double ovl = 0., cn = 0.;
            const double intersections = vals[0];
            if (intersections > 0.) {
                ovl = vals[1] / intersections;
                cn = intersections / static_cast<double>(n);
                if (method == MCAC_VS_ALPHAS) {
                    const double c_s10 = vals[2] / vals[6], c_v20 = vals[3] / vals[5], c_v30 = vals[4] / vals[5];
                    const double min_cn = 2 * (1.0 - 1.0 / static_cast<double>(n));
                    const double extreme = d.a_alpha[slot];
                    V *= volume_alpha_correction(cn, c_v20, c_v30, min_cn, extreme);
                    S *= surface_alpha_correction(cn, c_s10, min_cn, extreme);
                }
            }
            d.a_ovl[slot] = ovl;
            d.a_cn[slot] = cn;
            d.a_vol[slot] = V;
            d.a_surf[slot] = S;
            if (V <= 0 || S <= 0) d.sc->error = 8;  // VolSurfError, aggregat.cpp:427-429
        }
        group_sync<kBlock>();
    }
    // ---- update_partial: centre of mass (lanes 0..2), mean diameter / mean sphere volume (lanes 3, 4)
    const double V = d.a_vol[slot];
    if (stage) {
        double acc = 0.;
        for (int base = 0; base < n; base += nth) {
            const int i = base + tid;
            if (i < n) {
                const double4 rel = d.s_relv[off + i];
                const double ve = d.s_veff[off + i];
                stage[tid] = rel.x * ve;
                stage[nth + tid] = rel.y * ve;
                stage[2 * nth + tid] = rel.z * ve;
                stage[3 * nth + tid] = d.s_posr[off + i].w;
                stage[4 * nth + tid] = rel.w;
            }
            group_sync<kBlock>();
            if (tid < 5) {
                const double *src = stage + tid * nth;
                const int m = (n - base < nth) ? n - base : nth;
                for (int k = 0; k < m; k++) acc += src[k];
            }
            group_sync<kBlock>();
        }
        if (tid < 3) scratch[tid] = acc / V;
        else if (tid == 3) scratch[3] = 2 * acc / static_cast<double>(n);  // dp
        else if (tid == 4) scratch[4] = acc / static_cast<double>(n);      // vol_pp
    } else if (tid < 3) {
        double acc = 0.;
        for (int i = 0; i < n; i++) {
            const double4 rel = d.s_relv[off + i];
            const double c = (tid == 0) ? rel.x : (tid == 1) ? rel.y : rel.z;
            acc += c * d.s_veff[off + i];
        }
        scratch[tid] = acc / V;
    } else if (tid == 3) {
        double acc = 0.;
        for (int i = 0; i < n; i++) acc += d.s_posr[off + i].w;
        scratch[3] = 2 * acc / static_cast<double>(n);  // dp
    } else if (tid == 4) {
        double acc = 0.;
        for (int i = 0; i < n; i++) acc += d.s_relv[off + i].w;
        scratch[4] = acc / static_cast<double>(n);  // vol_pp
    }
    group_sync<kBlock>();
    const double cx = scratch[0], cy = scratch[1], cz = scratch[2];
    double rmax = 0.;
    double gacc = 0.;  // gyration sums (lanes 0, 1), aggregat.cpp:465-483
    if (stage) {
        for (int base = 0; base < n; base += nth) {
            const int i = base + tid;
            if (i < n) {
                const double4 rel = d.s_relv[off + i];
                const double ex = rel.x - cx, ey = rel.y - cy, ez = rel.z - cz;
                const double dc = sqrt(ex * ex + ey * ey + ez * ez);
                d.s_dcen[off + i] = dc;
                const double r_i = d.s_posr[off + i].w;
                const double e = r_i + dc;
                rmax = (rmax < e) ? e : rmax;
                const double wgt = d.s_veff[off + i];
                stage[tid] = wgt * (dc * dc);
                stage[nth + tid] = wgt * (r_i * r_i);
            }
            group_sync<kBlock>();
            if (tid < 2) {
                const double *src = stage + tid * nth;
                const int m = (n - base < nth) ? n - base : nth;
                for (int k = 0; k < m; k++) gacc = gacc + src[k];
            }
            group_sync<kBlock>();
        }
    } else
    for (int i = tid; i < n; i += nth) {
        const double4 rel = d.s_relv[off + i];
        const double ex = rel.x - cx, ey = rel.y - cy, ez = rel.z - cz;
        const double dc = sqrt(ex * ex + ey * ey + ez * ez);
        d.s_dcen[off + i] = dc;
        const double e = d.s_posr[off + i].w + dc;
        rmax = (rmax < e) ? e : rmax;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(kFull, rmax, o);
        rmax = (rmax < other) ? other : rmax;
    }
    if (kBlock) {
        if ((tid & 31) == 0) scratch[8 + w] = rmax;
        __syncthreads();
        for (int ww = 0; ww < nw; ww++) rmax = (rmax < scratch[8 + ww]) ? scratch[8 + ww] : rmax;
    }
    group_sync<kBlock>();  // s_dcen visible
    if (stage) {
        if (tid < 2) scratch[5 + tid] = gacc;
    } else if (tid < 2) {  // gyration sums, aggregat.cpp:465-483
        double acc = 0.;
        for (int i = 0; i < n; i++) {
            const double wgt = d.s_veff[off + i];
            const double q = (tid == 0) ? d.s_dcen[off + i] : d.s_posr[off + i].w;
            acc = acc + wgt * (q * q);
        }
        scratch[5 + tid] = acc;
    }
    group_sync<kBlock>();
    if (tid == 0) {
        const double4 root = d.s_posr[off];
        agg_store_position(d, slot, root.x + cx, root.y + cy, root.z + cz, box);
        double4 a = d.a_posr[slot];
        a.w = rmax;
        d.a_posr[slot] = a;
        d.a_rx[slot] = cx; d.a_ry[slot] = cy; d.a_rz[slot] = cz;
        const double rg = sqrt(fabs((scratch[5] + 3. / 5. * scratch[6]) / V));
        d.a_rg[slot] = rg;
        const double dp = scratch[3], vol_pp = scratch[4];
        d.a_dp[slot] = dp;
        double ch = d.a_ch[slot];
        const Mobility mob = mobility_epilogue(d.gas, V, vol_pp, dp, &ch);
        d.a_ch[slot] = ch;
        d.a_bulk[slot] = mob.bulk_density;
        d.a_fagg[slot] = mob.f_agg;
        d.a_dm[slot] = mob.d_m;
        d.a_ts[slot] = mob.time_step;
        d.a_lpm[slot] = mob.lpm;
        d.a_dgdp[slot] = 2 * rg / dp;
        if (d.sc->maxradius < rmax) atomic_max_positive_double(&d.sc->maxradius, rmax);  // only ever grows: the plain read filters almost all
        d.a_dirty[slot] = full ? 0 : (flags & kDirtyFull);  // (a partial update leaves a due contact pass due)
    }
    group_sync<kBlock>();
}

// The same update for a SMALL aggregate (n <= kSingleMax) by ONE thread: a population of monomers / dimers (growth mode updates
// every aggregate every step, calcul.cpp:184-206) then runs one aggregate per lane instead of one per warp.  Every sum is taken
// in the same order as agg_update (sequential `myspheres` order; the overlap statistics in the butterfly order of its warp
// reduction), so both forms give bit-identical results.
constexpr int kSingleMax = 8;
constexpr int kUpdateWarpMax = 96;  // spheres of an aggregate up to which ONE WARP updates it when every aggregate is updated
__device__ void agg_update_single(const DevState &d, int slot, bool full, double box) {
    const int off = d.a_off[slot], n = d.a_n[slot];
    const int method = d.volsurf_method;
    const int flags = d.a_dirty[slot];  // (see agg_update)
    if (full && !(flags & kDirtyFull)) full = false;
    if (!full && !(flags & kDirtyPartial)) return;
    if (full) {
        double v[7][kSingleMax];
#pragma unroll
        for (int k = 0; k < 7; k++)
#pragma unroll
            for (int i = 0; i < kSingleMax; i++) v[k][i] = 0.;
        for (int i = 0; i < n; i++) {
            const double4 ri = d.s_relv[off + i];
            const double r_i = d.s_posr[off + i].w;
            const double s_i = d.s_surf[off + i];
            double veff = ri.w, seff = s_i;
            for (int j = 0; j < n; j++) {
                if (j == i) continue;
                const double4 rj = d.s_relv[off + j];
                const double r_j = d.s_posr[off + j].w;
                const double ex = ri.x - rj.x, ey = ri.y - rj.y, ez = ri.z - rj.z;
                const double dist = sqrt(ex * ex + ey * ey + ez * ez);
                const double rs = r_i + r_j;
                if (dist <= (1. + kCoordinationEpsilon) * rs) {
                    const double c_ij = (rs - dist) / rs;
                    v[0][i] += 1.;
                    v[1][i] += c_ij;
                    if (method == MCAC_VS_ALPHAS) {
                        const double vp = pow(r_i, 3.) + pow(r_j, 3.);
                        const double sp = r_i * r_i + r_j * r_j;
                        v[5][i] += vp;
                        v[6][i] += sp;
                        v[2][i] += c_ij * sp;
                        v[3][i] += (c_ij * c_ij) * vp;
                        v[4][i] += pow(c_ij, 3.) * vp;
                    } else if (method == MCAC_VS_CAPS) {
                        double caps[4];
                        if (i < j) {
                            lens_caps(r_i, ri.w, s_i, r_j, rj.w, d.s_surf[off + j], dist, caps);
                            veff = veff - caps[0];
                            seff = seff - caps[2];
                        } else {
                            lens_caps(r_j, rj.w, d.s_surf[off + j], r_i, ri.w, s_i, dist, caps);
                            veff = veff - caps[1];
                            seff = seff - caps[3];
                        }
                    }
                }
            }
            if (method == MCAC_VS_CAPS) {
                veff = (veff < 0.0) ? 0.0 : veff;
                seff = (seff < 0.0) ? 0.0 : seff;
            }
            d.s_veff[off + i] = veff;
            d.s_seff[off + i] = seff;
        }
        double vals[7];
#pragma unroll
        for (int k = 0; k < 7; k++)  // lane 0 of the xor-butterfly over lanes 0..7
            vals[k] = ((v[k][0] + v[k][4]) + (v[k][2] + v[k][6])) + ((v[k][1] + v[k][5]) + (v[k][3] + v[k][7]));
        double V = 0., S = 0.;
        for (int i = 0; i < n; i++) {
            V = V + d.s_veff[off + i];
            S = S + d.s_seff[off + i];
        }
        double ovl = 0., cn = 0.;
        const double intersections = vals[0];
        if (intersections > 0.) {
            ovl = vals[1] / intersections;
            cn = intersections / static_cast<double>(n);
            if (method == MCAC_VS_ALPHAS) {
                const double c_s10 = vals[2] / vals[6], c_v20 = vals[3] / vals[5], c_v30 = vals[4] / vals[5];
                const double min_cn = 2 * (1.0 - 1.0 / static_cast<double>(n));
                const double extreme = d.a_alpha[slot];
                V *= volume_alpha_correction(cn, c_v20, c_v30, min_cn, extreme);
                S *= surface_alpha_correction(cn, c_s10, min_cn, extreme);
            }
        }
        d.a_ovl[slot] = ovl;
        d.a_cn[slot] = cn;
        d.a_vol[slot] = V;
        d.a_surf[slot] = S;
        if (V <= 0 || S <= 0) d.sc->error = 8;  // VolSurfError, aggregat.cpp:427-429
    }
    // ---- update_partial
    const double V = d.a_vol[slot];
    double ax = 0., ay = 0., az = 0., ar = 0., av = 0.;
    for (int i = 0; i < n; i++) {
        const double4 rel = d.s_relv[off + i];
        const double ve = d.s_veff[off + i];
        ax += rel.x * ve;
        ay += rel.y * ve;
        az += rel.z * ve;
        ar += d.s_posr[off + i].w;
        av += rel.w;
    }
    const double cx = ax / V, cy = ay / V, cz = az / V;
    const double dp = 2 * ar / static_cast<double>(n), vol_pp = av / static_cast<double>(n);
    double rmax = 0., g0 = 0., g1 = 0.;
    for (int i = 0; i < n; i++) {
        const double4 rel = d.s_relv[off + i];
        const double ex = rel.x - cx, ey = rel.y - cy, ez = rel.z - cz;
        const double dc = sqrt(ex * ex + ey * ey + ez * ez);
        d.s_dcen[off + i] = dc;
        const double r = d.s_posr[off + i].w;
        const double e = r + dc;
        rmax = (rmax < e) ? e : rmax;
        const double wgt = d.s_veff[off + i];
        g0 = g0 + wgt * (dc * dc);
        g1 = g1 + wgt * (r * r);
    }
    const double4 root = d.s_posr[off];
    agg_store_position(d, slot, root.x + cx, root.y + cy, root.z + cz, box);
    double4 a = d.a_posr[slot];
    a.w = rmax;
    d.a_posr[slot] = a;
    d.a_rx[slot] = cx; d.a_ry[slot] = cy; d.a_rz[slot] = cz;
    const double rg = sqrt(fabs((g0 + 3. / 5. * g1) / V));
    d.a_rg[slot] = rg;
    d.a_dp[slot] = dp;
    double ch = d.a_ch[slot];
    const Mobility mob = mobility_epilogue(d.gas, V, vol_pp, dp, &ch);
    d.a_ch[slot] = ch;
    d.a_bulk[slot] = mob.bulk_density;
    d.a_fagg[slot] = mob.f_agg;
    d.a_dm[slot] = mob.d_m;
    d.a_ts[slot] = mob.time_step;
    d.a_lpm[slot] = mob.lpm;
    d.a_dgdp[slot] = 2 * rg / dp;
    if (d.sc->maxradius < rmax) atomic_max_positive_double(&d.sc->maxradius, rmax);
    d.a_dirty[slot] = full ? 0 : (flags & kDirtyFull);
}

// K4 — AggregatList::merge + Aggregate::merge + ListStorage::merge/remove (aggregat_list.cpp:367-410,
// aggregat.cpp:486-544): CTA-cooperative.  Returns (to every thread, uniformly) 1 when the aggregates were united.
__device__ int agg_merge(const DevState &d, int ms, int os, int moving_agg, int other_agg, double *scratch, double box, double *stage = nullptr) {
    const int tid = threadIdx.x, nth = blockDim.x;
    const double4 pm = d.s_posr[ms], po = d.s_posr[os];
    if (!spheres_in_contact(pm.x, pm.y, pm.z, pm.w, po.x, po.y, po.z, po.w, box)) return 0;  // aggregat_list.cpp:375
    const int kept = moving_agg < other_agg ? moving_agg : other_agg;  // min(label) == min(slot)
    const int removed = moving_agg < other_agg ? other_agg : moving_agg;
    const int my = (kept == moving_agg) ? ms : os, oth = (kept == moving_agg) ? os : ms;
    const int n_k = d.a_n[kept], n_r = d.a_n[removed], off_k = d.a_off[kept], off_r = d.a_off[removed];
    const double newtime = d.a_ptime[kept] + d.a_ptime[removed] - d.sc->time;  // :387
    const int total_charge = d.a_charge[kept] + d.a_charge[removed];
    const double4 root = d.s_posr[off_k];
    const double4 pmy = d.s_posr[my], poth = d.s_posr[oth];
    const double4 rmy = d.s_relv[my], roth = d.s_relv[oth];
    const double dcx = periodic_distance(poth.x - pmy.x, box), dcy = periodic_distance(poth.y - pmy.y, box),
                 dcz = periodic_distance(poth.z - pmy.z, box);
    const double fx = rmy.x + dcx - roth.x, fy = rmy.y + dcy - roth.y, fz = rmy.z + dcz - roth.z;  // aggregat.cpp:516-521
    const int dst = d.sc->pool_top;
    __syncthreads();  // everybody has read pool_top / the contact spheres before anything is rewritten
    if (dst + n_k + n_r > d.sph_cap) {
        if (tid == 0) { d.sc->error = 1; d.sc->error_detail = DETAIL_POOL_FULL; }
        return 0;
    }
    for (int i = tid; i < n_k + n_r; i += nth) {
        const bool from_removed = i >= n_k;
        const int src = from_removed ? off_r + (i - n_k) : off_k + i;
        double4 p = d.s_posr[src];
        double4 rel = d.s_relv[src];
        if (from_removed) {
            rel.x += fx; rel.y += fy; rel.z += fz;                              // Sphere::relative_translate(diffpos)
            p.x = rel.x + root.x; p.y = rel.y + root.y; p.z = rel.z + root.z;  // newpos = rel; newpos += refpos
        }
        const int t = dst + i;
        d.s_posr[t] = p;
        d.s_relv[t] = rel;
        d.s_surf[t] = d.s_surf[src];
        d.s_veff[t] = d.s_veff[src];
        d.s_seff[t] = d.s_seff[src];
        d.s_dcen[t] = d.s_dcen[src];
        const int id = d.s_id[src];
        d.s_id[t] = id;
        d.s_charge[t] = d.s_charge[src];
        d.slot_of_id[id] = t;
    }
    __syncthreads();
    if (tid == 0) {
        d.sc->pool_top = dst + n_k + n_r;
        d.a_off[kept] = dst;
        d.a_n[kept] = n_k + n_r;
        d.a_alpha[kept] = 1.0 / static_cast<double>(n_k + n_r);
        d.a_alive[removed] = 0;
        d.a_n[removed] = 0;
        d.a_dirty[kept] = kDirtyAll;
        d.sc->n_agg -= 1;
    }
    __syncthreads();
    // Aggregate::update() of the merged aggregate: a small one (the usual case: two monomers, a monomer and a dimer) by ONE thread —
    // the CTA-cooperative form is a chain of ~20 block barriers for a handful of spheres
    if (n_k + n_r <= kSingleMax) {
        if (tid == 0) agg_update_single(d, kept, true, box);
        __syncthreads();
    } else {
        agg_update<true>(d, kept, true, tid, nth, scratch, box, stage);
    }
    if (tid == 0) {
        d.a_ptime[kept] = newtime;
        d.a_charge[kept] = total_charge;
    }
    __syncthreads();
    return 1;
}

// ------------------------------------------------------------------------------------------------
// K3 + orchestration — in-order commit of a speculative batch (single CTA).
// Steps 0..B-1 were searched in parallel against the state at the start of the batch.  A step is valid iff no
// EARLIER step of the batch moved the same aggregate or an aggregate that is, before or after its move, an
// eligible suspect of this step (bounding-sphere sweep test of aggregat_list.cpp:510-532): then the sequential
// algorithm would have computed exactly the same result.  The first contact ends the batch (its merge
// re-sorts the pick table); the first conflict makes the next batch restart at that step.
// ------------------------------------------------------------------------------------------------
struct BatchArgs {
    int nq;
    const int *q_slot;
    const double *q_dir;
    const double *q_dist;
    const SearchResult *res;
    mcac_step_record *rec;  // device buffer or nullptr
    long long rec_cap, rec_base;
    long long max_steps;    // steps still allowed in this run call
    long long steps_limit_abs;  // > 0: Scalars::steps_done may not pass this (batches submitted before the previous one was read back)
    long long *prof;            // diagnostics (may be null): [k] += SM cycles of thread 0 in phase k of k_commit, [15] += launches
};
__global__ void __launch_bounds__(kCommitThreads) k_commit(DevState d, BatchArgs b) {
    __shared__ int sh_slot[kMaxBatch];
    __shared__ unsigned char sh_contact[kMaxBatch];
    __shared__ double sh_time[kMaxBatch + 1];
    __shared__ double scratch[kUpdateScratch];
    __shared__ double cap[8];  // contact step: dt, proper time, position right after the move
    __shared__ int s_conf, s_contact, s_limit, s_finished, s_bad;
    __shared__ double4 sh_posr[kMaxBatch];  // movers of the batch staged once: the O(B^2) conflict test then runs from shared memory
    __shared__ double sh_dist[kMaxBatch];
    __shared__ double sh_dir[3 * kMaxBatch];
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
    Scalars &sc = *d.sc;
    const double box = sc.box_length;
    const int nq = b.nq;
    const int n_agg_before = sc.n_agg;
    const long long steps_before = sc.steps_done, rand_before = sc.rand_pos, iter_before = sc.n_iter_without_event;
    const double dt_base = sc.max_time_step / sc.cum_total;  // AggregatList::get_time_step(max), aggregat_list.cpp:54-58
    long long t_prev = clock64();
    auto plap = [&](int k) { if (b.prof && tid == 0) { const long long t = clock64(); atomicAdd(reinterpret_cast<unsigned long long *>(b.prof + k), (unsigned long long)(t - t_prev)); t_prev = t; } };
    if (sc.b_need == 99 || sc.error != 0) {  // the pick table of this batch comes from a sort that gave up (host redoes it), or an
        // earlier kernel of the batch reported an error (its queries are not usable: q_slot may be -1): commit nothing
        if (tid == 0) { sc.b_committed = 0; sc.b_stop_reason = STOP_NONE; sc.b_contact = 0; sc.b_merged = 0; }
        return;
    }
    // PhysicalModel::finished (physical_model.cpp:288-337) at the top of the first step of the batch: the host checks the same
    // conditions before it submits a batch, except for batches it submits ahead of reading the previous one back
    if (sc.n_agg < 1 || sc.n_agg <= d.n_agg_limit ||
        (d.npp_limit > 0 && static_cast<double>(sc.n_sph) / static_cast<double>(sc.n_agg) >= static_cast<double>(d.npp_limit))) {
        if (tid == 0) { sc.b_committed = 0; sc.b_stop_reason = STOP_FINISHED; sc.b_contact = 0; sc.b_merged = 0; }
        return;
    }
    const long long max_steps = b.steps_limit_abs > 0 ? min(b.max_steps, b.steps_limit_abs - steps_before) : b.max_steps;
    if (tid == 0) { s_conf = nq; s_contact = nq; s_limit = nq; s_finished = 0; s_bad = nq; }
    __syncthreads();
    for (int j = tid; j < nq; j += nth) {
        const int qs = b.q_slot[j];
        sh_slot[j] = qs;
        // a search that did not complete (status 1: more eligible suspects than the list holds, 3: VerletError) or a query without
        // an aggregate has no usable distance: the first such step bounds what this batch may commit
        if (qs < 0 || b.res[j].status != 0) { atomicMin(&s_bad, j); sh_contact[j] = 0; sh_posr[j] = make_double4(0., 0., 0., 0.); sh_dist[j] = 0.;
            sh_dir[3 * j] = sh_dir[3 * j + 1] = sh_dir[3 * j + 2] = 0.; continue; }
        sh_contact[j] = (b.res[j].distance <= b.q_dist[j]) ? 1 : 0;  // `next_contact <= full_distance`, calcul.cpp:128
        sh_posr[j] = d.a_posr[qs];
        sh_dist[j] = b.q_dist[j];
        sh_dir[3 * j] = b.q_dir[3 * j]; sh_dir[3 * j + 1] = b.q_dir[3 * j + 1]; sh_dir[3 * j + 2] = b.q_dir[3 * j + 2];
    }
    __syncthreads();
    plap(0);
    // ---- stop conditions evaluated at the top of every step (PhysicalModel::finished, physical_model.cpp:288-337)
    for (int j = tid; j < nq; j += nth)
        if (sh_contact[j]) atomicMin(&s_contact, j);
    __syncthreads();
    if (tid == 0) {  // (a chain of dependent additions: only as far as the first contact, behind which nothing is committed)
        double t = sc.time;
        int lim = nq;
        if ((long long)lim > max_steps) lim = (int)(max_steps > 0 ? max_steps : 0);
        const int upto = min(nq, s_contact + 1);
        const bool limits = d.time_limit > 0 || d.n_iter_limit > 0;
        for (int j = 0; j < upto; j++) {
            sh_time[j] = t;
            if (limits) {
                const bool fin = (d.time_limit > 0 && t >= d.time_limit) || (d.n_iter_limit > 0 && iter_before + j >= d.n_iter_limit);
                if (fin && j < lim) { lim = j; s_finished = 1; }
            }
            t = t + dt_base;  // free flight: dt * (lpm/lpm + 0) == dt
        }
        sh_time[upto] = t;
        s_limit = lim;
    }
    __syncthreads();
    plap(1);
    // ---- conflicts with earlier movers of the batch: all (j, i<j) pairs up to the first contact (nothing behind it can be
    // committed by this batch), flattened over the block
    const int nj = min(min(min(nq, s_limit), s_contact + 1), s_bad);
    for (int p = tid; p < nj * nj; p += nth) {
        const int j = p / nj, i = p - j * nj;
        if (i >= j || j >= s_conf) continue;  // benign race on s_conf: only ever shrinks the work
        const int sj = sh_slot[j], si = sh_slot[i];
        bool conflict = (si == sj);
        if (!conflict) {
            const double4 aj = sh_posr[j], ai = sh_posr[i];
            const double dj = sh_dist[j], di = sh_dist[i];
            const double guard = (dj + di + aj.w + ai.w) * (1. + 1e-9) + 1e-9 * box;
            // centres are wrapped into [0, box): minimum-image separation per axis, branch-free
            double ex = fabs(aj.x - ai.x), ey = fabs(aj.y - ai.y), ez = fabs(aj.z - ai.z);
            ex = fmin(ex, box - ex); ey = fmin(ey, box - ey); ez = fmin(ez, box - ez);
            if (ex <= guard && ey <= guard && ez <= guard) {
                // exact: was / will `i` be an eligible suspect of j (bounding-sphere sweep, strict <)?
                const double djx = sh_dir[3 * j], djy = sh_dir[3 * j + 1], djz = sh_dir[3 * j + 2];
                const double before = pair_contact_distance(aj.x, aj.y, aj.z, aj.w, ai.x, ai.y, ai.z, ai.w, djx, djy, djz, dj, box);
                const double nx = periodic_position(ai.x + sh_dir[3 * i] * di, box), ny = periodic_position(ai.y + sh_dir[3 * i + 1] * di, box),
                             nz = periodic_position(ai.z + sh_dir[3 * i + 2] * di, box);
                const double after = pair_contact_distance(aj.x, aj.y, aj.z, aj.w, nx, ny, nz, ai.w, djx, djy, djz, dj, box);
                conflict = before < dj || after < dj;
            }
        }
        if (conflict) atomicMin(&s_conf, j);
    }
    __syncthreads();
    plap(2);
    int stop = s_limit;
    int reason = s_finished ? STOP_FINISHED : STOP_BATCH_END;
    if (s_bad < stop) {
        // the reference would have thrown inside this step's distance_to_next_contact: report it instead of committing a
        // distance that is not the first contact.  Steps before it are committed (they are what the reference did).
        if (s_bad == 0) {
            if (tid == 0) {
                const bool overflow = sh_slot[0] >= 0 && b.res[0].status == 1;
                sc.error = overflow ? 1 : 3;  // UNKNOWN_ERROR (suspect list overflow) / VerletError
                sc.error_detail = overflow ? DETAIL_SUSPECT_OVERFLOW : DETAIL_NOT_ON_VERLET;
                sc.b_committed = 0; sc.b_stop_reason = STOP_NONE; sc.b_contact = 0; sc.b_merged = 0;
            }
            return;
        }
        stop = s_bad; reason = STOP_BATCH_END;
    }
    if (s_conf < stop) { stop = s_conf; reason = STOP_CONFLICT; }
    bool do_contact = s_contact < stop;  // the contact step is valid (no conflict before it) and allowed
    if (do_contact) { stop = s_contact; reason = STOP_CONTACT; }
    if (do_contact) {  // the merged block is written at the pool top: without room for it the batch ends before the contact step
        const int other = b.res[stop].other_agg;
        if (other >= 0 && sc.pool_top + d.a_n[sh_slot[stop]] + d.a_n[other] > d.sph_cap) { do_contact = false; reason = STOP_POOL; }
    }
    // ---- commit the free-flight steps [0, stop): distinct aggregates; single-sphere aggregates one THREAD per step (all of them
    // in flight at once: the chain of dependent loads of a move is paid once, not once per round of warps), the others one warp per step
    for (int j = tid; j < stop; j += nth) {
        const int sj = sh_slot[j];
        if (d.a_n[sj] != 1) continue;
        const double dj = b.q_dist[j];
        const double vx = b.q_dir[3 * j] * dj, vy = b.q_dir[3 * j + 1] * dj, vz = b.q_dir[3 * j + 2] * dj;
        const double4 a = d.a_posr[sj];  // same arithmetic as agg_translate
        const double nx = periodic_position(a.x + vx, box), ny = periodic_position(a.y + vy, box), nz = periodic_position(a.z + vz, box);
        const double refx = nx - d.a_rx[sj], refy = ny - d.a_ry[sj], refz = nz - d.a_rz[sj];
        const int off = d.a_off[sj];
        const double4 rel = d.s_relv[off];
        double4 p = d.s_posr[off];
        p.x = refx + rel.x; p.y = refy + rel.y; p.z = refz + rel.z;
        d.s_posr[off] = p;
        agg_store_position(d, sj, a.x + vx, a.y + vy, a.z + vz, box);
        atomicOr(&d.a_dirty[sj], kDirtyPartial);
        d.a_ptime[sj] += dt_base * (dj / dj + 0.0);
    }
    for (int j = warp; j < stop; j += nwarps) {
        const int sj = sh_slot[j];
        if (d.a_n[sj] == 1) continue;
        const double dj = b.q_dist[j];
        agg_translate<false>(d, sj, b.q_dir[3 * j] * dj, b.q_dir[3 * j + 1] * dj, b.q_dir[3 * j + 2] * dj, box, lane, 32);
        if (lane == 0) d.a_ptime[sj] += dt_base * (dj / dj + 0.0);  // calcul.cpp:147-149 with move == full, n_try == 1
    }
    __syncthreads();
    plap(3);
    int merged = 0;
    if (do_contact) {
        const int j = stop;
        const int sj = sh_slot[j];
        const SearchResult r = b.res[j];
        const double full = b.q_dist[j], move = r.distance;
        agg_translate<true>(d, sj, b.q_dir[3 * j] * move, b.q_dir[3 * j + 1] * move, b.q_dir[3 * j + 2] * move, box, tid, nth);
        if (tid == 0) {
            const double dt = dt_base * (move / full + 0.0);
            d.a_ptime[sj] += dt;
            sc.time = sh_time[j] + dt;
            const double4 a = d.a_posr[sj];
            cap[0] = dt; cap[1] = d.a_ptime[sj]; cap[2] = a.x; cap[3] = a.y; cap[4] = a.z;
        }
        __syncthreads();
        plap(4);
        merged = agg_merge(d, r.moving_slot, r.other_slot, sj, r.other_agg, scratch, box);
        plap(5);
    } else if (tid == 0) {
        sc.time = sh_time[stop];
    }
    __syncthreads();
    // ---- trace records (tests / replay), one per committed step.  Old sphere slots stay readable after a merge.
    const int done = stop + (do_contact ? 1 : 0);
    if (b.rec) {
        for (int j = tid; j < done; j += nth) {
            const long long at = b.rec_base + j;
            if (at >= b.rec_cap) continue;
            const SearchResult r = b.res[j];
            const bool is_contact = do_contact && j == stop;
            const bool has = r.other_agg >= 0;
            mcac_step_record o;
            o.step = steps_before + j;
            o.rand_calls = rand_before + 3LL * (j + 1);
            o.source = d.label_of_slot[sh_slot[j]];
            o.dir[0] = b.q_dir[3 * j]; o.dir[1] = b.q_dir[3 * j + 1]; o.dir[2] = b.q_dir[3 * j + 2];
            o.full_distance = b.q_dist[j];
            o.distance = r.distance;
            o.moving_sphere = has ? (long long)d.s_id[r.moving_slot] : -1;
            o.other_sphere = has ? (long long)d.s_id[r.other_slot] : -1;
            o.moving_label = has ? (long long)d.label_of_slot[sh_slot[j]] : -1;
            o.other_label = has ? (long long)d.label_of_slot[r.other_agg] : -1;
            o.n_agg_before = n_agg_before;
            o.time_before = sh_time[j];
            if (is_contact) {
                o.dt = cap[0]; o.proper_time_after = cap[1];
                o.pos_after[0] = cap[2]; o.pos_after[1] = cap[3]; o.pos_after[2] = cap[4];
            } else {
                const double4 a = d.a_posr[sh_slot[j]];
                o.dt = dt_base * (b.q_dist[j] / b.q_dist[j] + 0.0);
                o.proper_time_after = d.a_ptime[sh_slot[j]];
                o.pos_after[0] = a.x; o.pos_after[1] = a.y; o.pos_after[2] = a.z;
            }
            o.merged = is_contact ? merged : 0;
            o.n_try = 1;
            b.rec[at] = o;
        }
    }
    __syncthreads();
    {   // pair-test counters of the committed steps (block reduction)
        long long ps = 0, pb = 0;
        for (int j = tid; j < done; j += nth) { ps += b.res[j].n_sphere_pairs; pb += b.res[j].n_bounding; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { ps += __shfl_xor_sync(kFull, ps, o); pb += __shfl_xor_sync(kFull, pb, o); }
        if (lane == 0) { atomicAdd((unsigned long long *)&sc.pair_sphere, (unsigned long long)ps); atomicAdd((unsigned long long *)&sc.pair_bounding, (unsigned long long)pb); }
    }
    __syncthreads();
    if (tid == 0) {
        sc.steps_done = steps_before + done;
        sc.rand_pos = rand_before + 3LL * done;
        sc.searches += done;
        if (reason == STOP_CONFLICT) sc.conflicts += 1;
        if (merged) {
            sc.n_iter_without_event = 0;
            sc.total_events += 1;
            sc.event = 1;
        } else if (done > 0) {
            sc.n_iter_without_event = iter_before + done;
            sc.event = 0;
        }
        sc.b_committed = done;
        sc.b_stop_reason = reason;
        sc.b_contact = do_contact ? 1 : 0;
        sc.b_merged = merged;
    }
    plap(6);
    if (b.prof && tid == 0) atomicAdd(reinterpret_cast<unsigned long long *>(b.prof + 15), 1ULL);
}

// ------------------------------------------------------------------------------------------------
// General step (growth / pick_last / no-collision configurations): one MC step per launch sequence, in the exact
// order of calcul() (src/calcul.cpp:93-234): move + clocks here, then K8 growth, then the deferred merge, then updates.
// ------------------------------------------------------------------------------------------------
// AggregatList::pick_last (aggregat_list.cpp:67-81): first minimum of proper_time in label (= slot) order
__device__ __forceinline__ void dev_pick_last(const DevState &d, int *q_slot) {
    __shared__ double sv[32];
    __shared__ int si[32];
    const int n = d.sc->n_agg_slots;
    double best = INFINITY;
    int who = 0x7fffffff;
    for (int s = threadIdx.x; s < n; s += blockDim.x) {
        if (!d.a_alive[s]) continue;
        const double t = d.a_ptime[s];
        if (t < best || (t == best && s < who)) { best = t; who = s; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(kFull, best, o);
        const int ow = __shfl_xor_sync(kFull, who, o);
        if (ob < best || (ob == best && ow < who)) { best = ob; who = ow; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = who; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++)
            if (sv[w] < best || (sv[w] == best && si[w] < who)) { best = sv[w]; who = si[w]; }
        q_slot[0] = who;
    }
}
__global__ void __launch_bounds__(1024) k_pick_last(DevState d, int *q_slot) { dev_pick_last(d, q_slot); }
// direction (2 draws) + lpm for an already picked aggregate (PICK_LAST, or a re-drawn orientation); thread 0 of the CTA
__device__ __forceinline__ void dev_prepare_direction(const DevState &d, int *q_slot, double *q_dir, double *q_dist, long long draw_offset) {
    if (threadIdx.x != 0) return;
    const long long p = d.sc->rand_pos + draw_offset - d.rng_buf_base;
    const Vec3 dir = dev_direction(d, p);
    q_dir[0] = dir.x; q_dir[1] = dir.y; q_dir[2] = dir.z;
    q_dist[0] = d.a_lpm[q_slot[0]];
}
__global__ void k_prepare_direction(DevState d, int *q_slot, double *q_dir, double *q_dist, long long draw_offset) {
    if (blockIdx.x != 0) return;
    dev_prepare_direction(d, q_slot, q_dir, q_dist, draw_offset);
}
struct StepArgs {
    const int *q_slot;
    const double *q_dir;
    const double *q_dist;
    const SearchResult *res;
    mcac_step_record *rec;
    long long rec_cap, rec_index;
    int pick_last, with_collisions, n_try, draws;
    int draws_at_search;  // draws consumed when the last search returned (the tap convention of the step records)
};
__device__ __forceinline__ void dev_step_move(const DevState &d, const StepArgs &a) {
    const int tid = threadIdx.x, nth = blockDim.x;
    Scalars &sc = *d.sc;
    const double box = sc.box_length;
    const int slot = a.q_slot[0];
    const double full = a.q_dist[0];
    SearchResult r;
    r.distance = INFINITY; r.moving_slot = r.other_slot = r.other_agg = -1; r.n_bounding = 0; r.n_sphere_pairs = 0; r.status = 0;
    if (sc.error != 0) return;  // an earlier kernel of this step failed (the pick may be -1): nothing is moved
    if (slot < 0) { if (tid == 0) { sc.error = 3; sc.error_detail = DETAIL_NOT_ON_VERLET; } return; }
    if (a.with_collisions) r = a.res[0];
    if (r.status != 0) {  // incomplete search: there is no first contact to move to
        if (tid == 0) { sc.error = r.status == 1 ? 1 : 3; sc.error_detail = r.status == 1 ? DETAIL_SUSPECT_OVERFLOW : DETAIL_NOT_ON_VERLET; }
        return;
    }
    const bool contact = a.with_collisions && r.distance <= full;
    const double move = contact ? r.distance : full;
    const double time_before = sc.time;
    const int n_agg_before = sc.n_agg;
    double deltatemps = a.pick_last ? d.a_ts[slot] : sc.max_time_step / sc.cum_total;  // calcul.cpp:103,106
    double deltatemps_indiv = d.a_ts[slot];                                             // :108
    agg_translate<true>(d, slot, a.q_dir[0] * move, a.q_dir[1] * move, a.q_dir[2] * move, box, tid, nth);
    if (tid == 0) {
        const double factor = move / full + static_cast<double>(a.n_try - 1);           // :147-148
        deltatemps = deltatemps * factor;
        deltatemps_indiv = deltatemps_indiv * factor;
        d.a_ptime[slot] += deltatemps;                                                  // :149
        const double dt_rec = deltatemps;
        if (a.pick_last) deltatemps = deltatemps / double(n_agg_before);                // :150-152
        sc.time = time_before + deltatemps;
        sc.p_contact = contact ? 1 : 0;
        sc.p_ms = r.moving_slot; sc.p_os = r.other_slot; sc.p_magg = slot; sc.p_oagg = r.other_agg;
        sc.p_dt = deltatemps; sc.p_dt_indiv = deltatemps_indiv; sc.p_slot = slot;
        sc.searches += a.with_collisions ? 1 : 0;
        sc.pair_sphere += r.n_sphere_pairs;
        sc.pair_bounding += r.n_bounding;
        if (a.rec && a.rec_index < a.rec_cap) {
            mcac_step_record o;
            const bool has = r.other_agg >= 0;
            o.step = sc.steps_done;
            o.rand_calls = sc.rand_pos + a.draws_at_search;
            o.source = d.label_of_slot[slot];
            o.dir[0] = a.q_dir[0]; o.dir[1] = a.q_dir[1]; o.dir[2] = a.q_dir[2];
            o.full_distance = full;
            o.distance = r.distance;
            o.moving_sphere = has ? (long long)d.s_id[r.moving_slot] : -1;
            o.other_sphere = has ? (long long)d.s_id[r.other_slot] : -1;
            o.moving_label = has ? (long long)d.label_of_slot[slot] : -1;
            o.other_label = has ? (long long)d.label_of_slot[r.other_agg] : -1;
            o.n_agg_before = n_agg_before;
            o.time_before = time_before;
            o.dt = dt_rec;
            o.proper_time_after = d.a_ptime[slot];
            const double4 p = d.a_posr[slot];
            o.pos_after[0] = p.x; o.pos_after[1] = p.y; o.pos_after[2] = p.z;
            o.merged = 0;
            o.n_try = a.n_try;
            a.rec[a.rec_index] = o;
        }
        sc.steps_done += 1;
        sc.rand_pos += a.draws;
    }
}
__global__ void __launch_bounds__(kCommitThreads) k_step_move(DevState d, StepArgs a) { dev_step_move(d, a); }
// AggregatList::check_InterPotentialRegime (aggregat_list.cpp:313-366) for the contact found by the last search:
// sticking / repulsion / bouncing decided with <= 2 draws taken at stream offset `draw_offset` of this step.
__device__ __forceinline__ void dev_check_regime(const DevState &d, const SearchResult *res, const double *q_dist, long long draw_offset) {
    if (d.sc->error != 0) return;  // an earlier kernel of this step failed: leave the state as it is
    if (threadIdx.x != 0) return;
    Scalars &sc = *d.sc;
    sc.p_regime = 0;
    sc.p_regime_draws = 0;
    const SearchResult r = res[0];
    if (!(r.distance <= q_dist[0])) return;  // no contact: the move is effective
    const double d_moving = 2.0 * d.s_posr[r.moving_slot].w, d_other = 2.0 * d.s_posr[r.other_slot].w;
    double p_stick = 1.0, p_coll = 1.0;
    if (d.with_external_potentials) {
        // Interpotential::get_Ebar_Ewell (physical_model_interpotential.cpp:144-197)
        const int q1 = 0, q2 = 0;  // aggregate charges: electric charges are not built (with_electric_charges = false)
        int i1 = 0, j1 = 0, k = -1, l = -1;
        while (i1 < d.ip_n1 && !(d_moving < d.ip_dp1[i1])) i1++;  // std::upper_bound
        while (j1 < d.ip_n2 && !(d_other < d.ip_dp2[j1])) j1++;
        for (int c = 0; c < d.ip_nq; c++) { if (d.ip_charge[c] == q1 && k < 0) k = c; if (d.ip_charge[c] == q2 && l < 0) l = c; }
        if (i1 == 0 || i1 == d.ip_n1 || j1 == 0 || j1 == d.ip_n2 || k < 0 || l < 0) { sc.error = 11; return; }  // InterPotentialError
        const int i0 = i1 - 1, j0 = j1 - 1;
        const double a = (d_moving - d.ip_dp1[i0]) / (d.ip_dp1[i1] - d.ip_dp1[i0]);
        const double b = (d_other - d.ip_dp2[j0]) / (d.ip_dp2[j1] - d.ip_dp2[j0]);
        auto at = [&](int ii, int jj) { return (((size_t)k * d.ip_nq + l) * d.ip_n1 + ii) * d.ip_n2 + jj; };
        const double e_bar = interpolate_2d(d.ip_ebar[at(i0, j0)], d.ip_ebar[at(i0, j1)], d.ip_ebar[at(i1, j0)], d.ip_ebar[at(i1, j1)], a, b);
        const double e_well = interpolate_2d(d.ip_ewell[at(i0, j0)], d.ip_ewell[at(i0, j1)], d.ip_ewell[at(i1, j0)], d.ip_ewell[at(i1, j1)], a, b);
        const double e_stick = fabs(e_well) + fabs(e_bar);
        p_stick = erf(sqrt(e_stick)) - sqrt(e_stick) * exp(-e_stick);
        p_coll = 1.0 - erf(sqrt(e_bar)) + sqrt(e_bar) * exp(-e_bar);
    } else {  // Hou et al., J. Aerosol Sci. (2020) 105478
        const double kbt = kBoltzmann * d.gas.temperature;
        double dd = d_moving * d_other / (d_moving + d_other);
        dd = dd * (1e+09);
        const double e_well = (-6.6891e-23) * pow(dd, 3.) + (1.1244e-21) * (dd * dd) + (1.1394e-20) * dd - 5.5373e-21;
        p_stick = 1.0 - (1.0 + fabs(e_well) / kbt) * exp(-fabs(e_well) / kbt);
    }
    const long long p = sc.rand_pos + draw_offset - d.rng_buf_base;
    sc.p_regime_draws = 1;
    if (uniform_from_rand(d.rng_buf[p]) > p_coll) { sc.p_regime = 1; return; }  // REPULSION
    sc.p_regime_draws = 2;
    if (uniform_from_rand(d.rng_buf[p + 1]) > p_stick) { sc.p_regime = 2; return; }  // BOUNCING
}
__global__ void k_check_regime(DevState d, const SearchResult *res, const double *q_dist, long long draw_offset) {
    if (blockIdx.x != 0) return;
    dev_check_regime(d, res, q_dist, draw_offset);
}

// AggregatList::add(n) for nucleation (aggregat_list.cpp:82-99 -> Aggregate::init(nucleation = true), aggregat.cpp:162-229):
// each new monomer draws its diameter (1 draw) then positions (3 draws per try) until it overlaps no existing aggregate
// (bounding sphere first, then member spheres: aggregat_distance.cpp:45-58), becomes the last aggregate / last sphere and gets
// Aggregate::update().  Single CTA; the free-space test of a try is spread over the threads.
__device__ __forceinline__ void dev_nucleate(const DevState &d, double deltatemps_unused, int use_pending_dt) {
    if (d.sc->error != 0) return;  // an earlier kernel of this step failed: leave the state as it is
    __shared__ double scratch[kUpdateScratch];
    __shared__ double cand[4];
    __shared__ int s_hit, s_stop, s_count;
    const int tid = threadIdx.x, nth = blockDim.x;
    Scalars &sc = *d.sc;
    const double box = sc.box_length;
    if (tid == 0) {
        // PhysicalModel::nucleation(dt) (physical_model.cpp:499-502) + the accumulator test of calcul.cpp:211-216
        const double dt = use_pending_dt ? sc.p_dt : deltatemps_unused;
        sc.nucleation_accum += d.flux_nucleation * sc.box_volume * dt;
        int n_new = 0;
        if (sc.nucleation_accum > 1.0) {
            n_new = static_cast<int>(floor(sc.nucleation_accum));
            sc.nucleation_accum -= static_cast<double>(n_new);
        }
        s_count = n_new;
        sc.n_nucleated = n_new;
        s_stop = 0;
    }
    __syncthreads();
    const int n_new = s_count;
    for (int m = 0; m < n_new; m++) {
        if (tid == 0) {
            const long long p = sc.rand_pos - d.rng_buf_base;
            if (p + 1 >= d.rng_buf_n) { sc.error = 1; sc.error_detail = DETAIL_RNG_NOT_STAGED; s_stop = 1; }
            else {
                const double diameter = diameter_from_draw(uniform_from_rand(d.rng_buf[p]), d.nucl_mean_diameter, d.nucl_dispersion_diameter, d.init_mode_normal);
                cand[3] = diameter * 0.5;
                sc.rand_pos += 1;
            }
        }
        __syncthreads();
        if (s_stop) return;
        const int max_tries = sc.n_sph - m + n_new;  // `n_try < external_storage->spheres.size()`: the list already holds the n new spheres
        bool placed = false;
        for (int attempt = 0; attempt < max_tries && !placed; attempt++) {
            if (tid == 0) {
                const long long p = sc.rand_pos - d.rng_buf_base;
                if (p + 3 >= d.rng_buf_n) { sc.error = 1; sc.error_detail = DETAIL_RNG_NOT_STAGED; s_stop = 1; }
                else {
                    cand[0] = uniform_from_rand(d.rng_buf[p]) * box;
                    cand[1] = uniform_from_rand(d.rng_buf[p + 1]) * box;
                    cand[2] = uniform_from_rand(d.rng_buf[p + 2]) * box;
                    sc.rand_pos += 3;
                }
                s_hit = 0;
            }
            __syncthreads();
            if (s_stop) return;
            const double px = cand[0], py = cand[1], pz = cand[2], pr = cand[3];
            for (int s = tid; s < sc.n_agg_slots; s += nth) {
                if (!d.a_alive[s]) continue;
                const double4 a = d.a_posr[s];
                if (!spheres_in_contact(px, py, pz, pr, a.x, a.y, a.z, a.w, box)) continue;
                const int off = d.a_off[s], n = d.a_n[s];
                for (int k = 0; k < n; k++) {
                    const double4 q = d.s_posr[off + k];
                    if (spheres_in_contact(px, py, pz, pr, q.x, q.y, q.z, q.w, box)) { s_hit = 1; break; }
                }
            }
            __syncthreads();
            placed = (s_hit == 0);
            __syncthreads();
        }
        if (!placed) { if (tid == 0) sc.error = 6; return; }  // TooDenseError
        const int slot = sc.n_agg_slots, sp = sc.pool_top, id = sc.n_sph;
        if (slot >= d.agg_cap || sp >= d.sph_cap || id >= d.sph_cap) { if (tid == 0) { sc.error = 1; sc.error_detail = DETAIL_POOL_FULL; } return; }
        if (tid == 0) {
            const double r = cand[3];
            d.s_posr[sp] = make_double4(cand[0], cand[1], cand[2], r);
            d.s_relv[sp] = make_double4(0., 0., 0., volume_factor() * pow(r, 3.));
            d.s_surf[sp] = surface_factor() * (r * r);
            d.s_id[sp] = id;
            d.s_charge[sp] = 0;
            d.slot_of_id[id] = sp;
            d.a_n[slot] = 1;
            d.a_off[slot] = sp;
            d.a_alpha[slot] = 1.0;
            d.a_alive[slot] = 1;
            d.a_dirty[slot] = kDirtyAll;
            d.a_charge[slot] = 0;
            d.a_ptime[slot] = sc.time;
            d.a_ch[slot] = 0.;
            d.a_posr[slot] = make_double4(cand[0], cand[1], cand[2], 0.);
            d.label_of_slot[slot] = sc.n_agg;
            d.slot_of_label[sc.n_agg] = slot;
            sc.n_agg_slots = slot + 1;
            sc.pool_top = sp + 1;
            sc.n_sph = id + 1;
            sc.n_agg += 1;
        }
        __syncthreads();
        agg_update<true>(d, slot, true, tid, nth, scratch, box);
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kCommitThreads) k_nucleate(DevState d, double deltatemps_unused, int use_pending_dt) {
    dev_nucleate(d, deltatemps_unused, use_pending_dt);
}

// the deferred AggregatList::merge of the step (calcul.cpp:174-181) + event bookkeeping (:222-229)
__device__ __forceinline__ void dev_step_merge(const DevState &d, mcac_step_record *rec, long long rec_cap, long long rec_index, double *stage = nullptr) {
    if (d.sc->error != 0) return;  // an earlier kernel of this step failed: leave the state as it is
    __shared__ double scratch[kUpdateScratch];
    Scalars &sc = *d.sc;
    int merged = 0;
    if (sc.p_contact) merged = agg_merge(d, sc.p_ms, sc.p_os, sc.p_magg, sc.p_oagg, scratch, sc.box_length, stage);
    __syncthreads();
    if (threadIdx.x == 0) {
        sc.b_merged = merged;
        sc.b_committed = 1;
        if (rec && rec_index < rec_cap) rec[rec_index].merged = merged;
    }
}
__global__ void __launch_bounds__(kCommitThreads) k_step_merge(DevState d, mcac_step_record *rec, long long rec_cap, long long rec_index) {
    dev_step_merge(d, rec, rec_cap, rec_index);
}
// event bookkeeping at the end of a general step (calcul.cpp:222-229): event = merge || nucleation; thread 0 of the CTA
__device__ __forceinline__ void dev_step_event(const DevState &d) {
    if (threadIdx.x != 0) return;
    Scalars &sc = *d.sc;
    if (sc.error != 0) return;
    const bool ev = sc.b_merged || sc.n_nucleated > 0;
    if (ev) { sc.n_iter_without_event = 0; sc.total_events += 1; sc.event = 1; }
    else { sc.n_iter_without_event += 1; sc.event = 0; }
}
__global__ void k_step_event(DevState d) {
    if (blockIdx.x != 0) return;
    dev_step_event(d);
}

// ------------------------------------------------------------------------------------------------
// Event pipeline: AggregatList::refresh + get_total_volume/surface (aggregat_list.cpp:100-108, 28-45) as a
// deterministic two-phase reduction, and the 1/dt weights of sort_time_steps (:124-131) in label order.
// ------------------------------------------------------------------------------------------------
constexpr int kReduceThreads = 256;
__global__ void __launch_bounds__(kReduceThreads) k_refresh_partials(DevState d, double *partials /* 3 x gridDim */) {
    __shared__ double sm[3][kReduceThreads / 32];
    const int n = d.sc->n_agg_slots;
    double mx = 0., sv = 0., ss = 0.;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
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
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sm[0][w] = mx; sm[1][w] = sv; sm[2][w] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int ww = 1; ww < kReduceThreads / 32; ww++) {
            mx = (mx < sm[0][ww]) ? sm[0][ww] : mx;
            sv += sm[1][ww];
            ss += sm[2][ww];
        }
        partials[blockIdx.x] = mx;
        partials[gridDim.x + blockIdx.x] = sv;
        partials[2 * gridDim.x + blockIdx.x] = ss;
    }
}
__global__ void k_refresh_final(DevState d, const double *partials, int nb) {  // one warp, fixed combination order
    const int lane = threadIdx.x;
    if (blockIdx.x != 0 || lane >= 32) return;
    double mx = 0., sv = 0., ss = 0.;
    for (int b = lane; b < nb; b += 32) {
        mx = (mx < partials[b]) ? partials[b] : mx;
        sv += partials[nb + b];
        ss += partials[2 * nb + b];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double omx = __shfl_xor_sync(kFull, mx, o);
        mx = (mx < omx) ? omx : mx;
        sv += __shfl_xor_sync(kFull, sv, o);
        ss += __shfl_xor_sync(kFull, ss, o);
    }
    if (lane != 0) return;
    Scalars &sc = *d.sc;
    sc.max_time_step = mx;
    sc.avg_npp = static_cast<double>(sc.n_sph) / static_cast<double>(sc.n_agg);
    sc.total_volume = sv;
    sc.total_surface = ss;
    // PhysicalModel::update, physical_model.cpp:489-498
    sc.total_volume_concent = sv / sc.box_volume;
    sc.total_surface_concent = ss / sc.box_volume;
    sc.aggregate_concentration = static_cast<double>(sc.n_agg) / sc.box_volume;
    sc.monomer_concentration = static_cast<double>(sc.n_sph) / sc.box_volume;
    sc.volume_fraction = sv / sc.box_volume;
}
// end of a general step: refresh() only after an event (calcul.cpp:232-234), PhysicalModel::update always (:272-277)
__global__ void k_step_totals(DevState d, const double *partials, int nb) {
    const int lane = threadIdx.x;
    if (blockIdx.x != 0 || lane >= 32) return;
    double mx = 0., sv = 0., ss = 0.;
    for (int b = lane; b < nb; b += 32) {
        mx = (mx < partials[b]) ? partials[b] : mx;
        sv += partials[nb + b];
        ss += partials[2 * nb + b];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double omx = __shfl_xor_sync(kFull, mx, o);
        mx = (mx < omx) ? omx : mx;
        sv += __shfl_xor_sync(kFull, sv, o);
        ss += __shfl_xor_sync(kFull, ss, o);
    }
    if (lane != 0) return;
    Scalars &sc = *d.sc;
    if (sc.event) {
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
__global__ void k_refresh_if_event(DevState d, const double *partials, int nb) {  // no growth: everything only after an event
    if (!d.sc->event) return;
    const int lane = threadIdx.x;
    if (blockIdx.x != 0 || lane >= 32) return;
    double mx = 0., sv = 0., ss = 0.;
    for (int b = lane; b < nb; b += 32) {
        mx = (mx < partials[b]) ? partials[b] : mx;
        sv += partials[nb + b];
        ss += partials[2 * nb + b];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double omx = __shfl_xor_sync(kFull, mx, o);
        mx = (mx < omx) ? omx : mx;
        sv += __shfl_xor_sync(kFull, sv, o);
        ss += __shfl_xor_sync(kFull, ss, o);
    }
    if (lane != 0) return;
    Scalars &sc = *d.sc;
    sc.max_time_step = mx;
    sc.avg_npp = static_cast<double>(sc.n_sph) / static_cast<double>(sc.n_agg);
    sc.total_volume = sv;
    sc.total_surface = ss;
    sc.total_volume_concent = sv / sc.box_volume;
    sc.total_surface_concent = ss / sc.box_volume;
    sc.aggregate_concentration = static_cast<double>(sc.n_agg) / sc.box_volume;
    sc.monomer_concentration = static_cast<double>(sc.n_sph) / sc.box_volume;
    sc.volume_fraction = sv / sc.box_volume;
}
// PhysicalModel::update only (growth mode updates the concentrations every step but max_time_step / avg_npp only on events)
__global__ void k_totals_final(DevState d, const double *partials, int nb) {
    const int lane = threadIdx.x;
    if (blockIdx.x != 0 || lane >= 32) return;
    double sv = 0., ss = 0.;
    for (int b = lane; b < nb; b += 32) { sv += partials[nb + b]; ss += partials[2 * nb + b]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sv += __shfl_xor_sync(kFull, sv, o); ss += __shfl_xor_sync(kFull, ss, o); }
    if (lane != 0) return;
    Scalars &sc = *d.sc;
    sc.total_volume = sv;
    sc.total_surface = ss;
    sc.total_volume_concent = sv / sc.box_volume;
    sc.total_surface_concent = ss / sc.box_volume;
    sc.aggregate_concentration = static_cast<double>(sc.n_agg) / sc.box_volume;
    sc.monomer_concentration = static_cast<double>(sc.n_sph) / sc.box_volume;
    sc.volume_fraction = sv / sc.box_volume;
}
__global__ void k_make_keys(DevState d, double factor) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= d.sc->n_agg) return;
    d.keys[l] = factor / d.a_ts[d.slot_of_label[l]];
}
// sorted labels -> slots, cumulative table already uploaded / scanned
__global__ void k_sorted_labels_to_slots(DevState d, const int *sorted_label, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d.sorted_slot[i] = d.slot_of_label[sorted_label[i]];
    if (i == 0) { d.sc->n_pick = n; d.sc->cum_total = d.cum[n - 1]; d.sc->pick_dense_from = n; }
}

// ------------------------------------------------------------------------------------------------
// K8 — surface growth of every sphere (Sphere::croissance_surface, sphere.cpp:113-120 through the chain
// aggregat_list.cpp:549-579): elementwise over the sphere pool, 8 B in / 24 B out per sphere.
// ------------------------------------------------------------------------------------------------
__global__ void k_grow(DevState d, double dt, int only_slot) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    int lo = 0, hi = d.sc->pool_top;
    if (only_slot >= 0) { lo = d.a_off[only_slot]; hi = lo + d.a_n[only_slot]; }
    const int t = lo + s;
    // radii change: the aggregates concerned need their next full update in full (there are never more aggregate slots than spheres)
    if (only_slot >= 0) { if (s == 0) d.a_dirty[only_slot] = kDirtyAll; }
    else if (s < d.sc->n_agg_slots) d.a_dirty[s] = kDirtyAll;
    if (t >= hi) return;
    double4 p = d.s_posr[t];
    const double new_r = p.w + d.u_sg * dt;  // PhysicalModel::grow, physical_model.cpp:587-590
    const double r2 = new_r * new_r;
    const double r3 = r2 * new_r;
    p.w = new_r;
    d.s_posr[t] = p;
    double4 rel = d.s_relv[t];
    rel.w = volume_factor() * r3;
    d.s_relv[t] = rel;
    d.s_surf[t] = surface_factor() * r2;
    if (new_r <= d.rp_min_oxid) { d.sc->error = 1; d.sc->error_detail = DETAIL_SPHERE_REMOVAL; }  // sphere removal by oxidation (u_sg < 0) is outside the built path
}
// growth of the general step: dt and the picked aggregate are device scalars written by k_step_move
__global__ void k_grow_pending(DevState d, int individual) {
    if (d.sc->error != 0) return;  // an earlier kernel of this step failed: leave the state as it is
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const Scalars &sc = *d.sc;
    int lo = 0, hi = sc.pool_top;
    double dt = sc.p_dt;
    if (individual) { lo = d.a_off[sc.p_slot]; hi = lo + d.a_n[sc.p_slot]; dt = sc.p_dt_indiv; }
    const int t = lo + s;
    if (individual) { if (s == 0) d.a_dirty[sc.p_slot] = kDirtyAll; }
    else if (s < sc.n_agg_slots) d.a_dirty[s] = kDirtyAll;
    if (t >= hi) return;
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
    if (new_r <= d.rp_min_oxid) { d.sc->error = 1; d.sc->error_detail = DETAIL_SPHERE_REMOVAL; }
}
// update block of calcul.cpp:184-206 for the general step: mode 0 = every aggregate, mode 1 = individual reactions
// (only the picked aggregate unless a merge happened, in which case every aggregate, as the reference does)
__global__ void __launch_bounds__(256) k_update_step(DevState d, int full, int individual) {
    if (d.sc->error != 0) return;  // an earlier kernel of this step failed: leave the state as it is
    __shared__ double scratch[8][kUpdateScratch / 4];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 8 + w;
    const Scalars &sc = *d.sc;
    if (slot >= sc.n_agg_slots || !d.a_alive[slot]) return;
    if (individual && !sc.b_merged) return;  // done by k_update_picked
    if (d.a_n[slot] <= kSingleMax) return;  // done by k_update_small
    if (full && d.a_n[slot] > kUpdateWarpMax && (d.a_dirty[slot] & kDirtyFull)) return;  // an O(n^2) pass of a big aggregate: done by k_update_big
    agg_update<false>(d, slot, full != 0, lane, 32, scratch[w], sc.box_length);
}
// ... and the big ones (n > kUpdateWarpMax) by a whole CTA each: a full update is O(n^2), a 10^3-sphere aggregate left to one warp would
// take milliseconds.  (The overlap statistics are combined in the tree of the group that computes them: every path — this kernel and
// the per-realization step loop — gives an aggregate of a given size to the same kind of group, so all paths agree bit for bit.)
__global__ void __launch_bounds__(kCommitThreads) k_update_big(DevState d, int full, int individual) {
    __shared__ double scratch[kUpdateScratch];
    const Scalars &sc = *d.sc;
    if (d.sc->error != 0) return;
    if (individual && !sc.b_merged) return;
    if (!full) return;  // partial updates are chains of ordered adds: one warp each (k_update_step / k_update_all)
    for (int slot = blockIdx.x; slot < sc.n_agg_slots; slot += gridDim.x) {
        if (!d.a_alive[slot] || d.a_n[slot] <= kUpdateWarpMax || !(d.a_dirty[slot] & kDirtyFull)) continue;
        agg_update<true>(d, slot, true, threadIdx.x, blockDim.x, scratch, sc.box_length);
    }
}
// individual surface reactions without a merge: only the picked aggregate is updated (calcul.cpp:196-203) — by a whole CTA, so
// that a 10^3-sphere aggregate's O(n^2) contact pass is not left to one warp
__global__ void __launch_bounds__(kCommitThreads) k_update_picked(DevState d, int full) {
    if (d.sc->error != 0) return;  // an earlier kernel of this step failed: leave the state as it is
    __shared__ double scratch[kUpdateScratch];
    const Scalars &sc = *d.sc;
    if (sc.b_merged) return;  // every aggregate is updated by k_update_small / k_update_step
    const int slot = sc.p_slot;
    if (slot < 0 || slot >= sc.n_agg_slots || !d.a_alive[slot]) return;
    agg_update<true>(d, slot, full != 0, threadIdx.x, blockDim.x, scratch, sc.box_length);
}
// the small aggregates (n <= kSingleMax) of the same update: one aggregate per THREAD
__global__ void __launch_bounds__(128) k_update_small(DevState d, int full, int individual, int all) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const Scalars &sc = *d.sc;
    if (slot >= sc.n_agg_slots || !d.a_alive[slot]) return;
    if (!all && individual && !sc.b_merged) return;  // done by k_update_picked
    if (d.a_n[slot] > kSingleMax) return;
    agg_update_single(d, slot, full != 0, sc.box_length);
}
// K5-K7 over ALL aggregates (growth mode, calcul.cpp:184-206): one warp per aggregate of more than kSingleMax spheres
__global__ void __launch_bounds__(256) k_update_all(DevState d, int full, int only_slot) {
    __shared__ double scratch[8][kUpdateScratch / 4];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int slot = blockIdx.x * 8 + w;
    if (only_slot >= 0) { if (slot != 0) return; slot = only_slot; }
    if (slot >= d.sc->n_agg_slots || !d.a_alive[slot]) return;
    if (only_slot < 0 && d.a_n[slot] <= kSingleMax) return;  // done by k_update_small
    if (only_slot < 0 && full && d.a_n[slot] > kUpdateWarpMax && (d.a_dirty[slot] & kDirtyFull)) return;  // done by k_update_big
    agg_update<false>(d, slot, full != 0, lane, 32, scratch[w], d.sc->box_length);
}
// Aggregate::update() / update_partial() of ONE aggregate (per-call C ABI), by the group every other path gives an aggregate of its size
__global__ void __launch_bounds__(kCommitThreads) k_update_one(DevState d, int full, int slot) {
    __shared__ double scratch[kUpdateScratch];
    if (slot < 0 || slot >= d.sc->n_agg_slots || !d.a_alive[slot]) return;
    const int n = d.a_n[slot];
    if (n <= kSingleMax) { if (threadIdx.x == 0) agg_update_single(d, slot, full != 0, d.sc->box_length); }
    else if (n <= kUpdateWarpMax || !full || !(d.a_dirty[slot] & kDirtyFull)) { if (threadIdx.x < 32) agg_update<false>(d, slot, full != 0, threadIdx.x, 32, scratch, d.sc->box_length); }
    else agg_update<true>(d, slot, full != 0, threadIdx.x, blockDim.x, scratch, d.sc->box_length);
}
// the 21 AggregatesFields of one aggregate slot (per-call C ABI)
__global__ void k_aggregate_fields(DevState d, int slot, double *out /* 22 */) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double4 p = d.a_posr[slot];
    const double v[22] = {d.a_rg[slot], d.a_fagg[slot], d.a_lpm[slot], d.a_ts[slot], p.w, d.a_vol[slot], d.a_surf[slot], p.x, p.y, p.z, d.a_rx[slot],
                          d.a_ry[slot], d.a_rz[slot], d.a_ptime[slot], d.a_dp[slot], d.a_dgdp[slot], d.a_ovl[slot], d.a_cn[slot],
                          static_cast<double>(d.a_charge[slot]), d.a_dm[slot], d.a_ch[slot], static_cast<double>(d.a_n[slot])};
    for (int k = 0; k < 22; k++) out[k] = v[k];
}
// single-aggregate entry points of the per-call C ABI
__global__ void __launch_bounds__(kCommitThreads) k_translate_one(DevState d, int slot, double vx, double vy, double vz) {
    agg_translate<true>(d, slot, vx, vy, vz, d.sc->box_length, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(kCommitThreads) k_merge_one(DevState d, int ms_id, int os_id, int *merged_out) {
    __shared__ double scratch[kUpdateScratch];
    const int ms = d.slot_of_id[ms_id], os = d.slot_of_id[os_id];
    // owning aggregates from the pool: search the aggregate whose block holds the slot (labels are not stored per sphere)
    __shared__ int owner[2];
    if (threadIdx.x == 0) { owner[0] = -1; owner[1] = -1; }
    __syncthreads();
    for (int s = threadIdx.x; s < d.sc->n_agg_slots; s += blockDim.x) {
        if (!d.a_alive[s]) continue;
        const int off = d.a_off[s], n = d.a_n[s];
        if (ms >= off && ms < off + n) owner[0] = s;
        if (os >= off && os < off + n) owner[1] = s;
    }
    __syncthreads();
    const int merged = (owner[0] >= 0 && owner[1] >= 0 && owner[0] != owner[1])
                           ? agg_merge(d, ms, os, owner[0], owner[1], scratch, d.sc->box_length) : 0;
    if (threadIdx.x == 0) {
        *merged_out = merged;
        if (merged) { d.sc->total_events += 1; d.sc->event = 1; d.sc->n_iter_without_event = 0; }
    }
}

// ------------------------------------------------------------------------------------------------
// Pool compaction: live aggregates re-packed in slot (= label) order into the alternate sphere buffers.
// ------------------------------------------------------------------------------------------------
__global__ void k_compact_counts(DevState d, int *counts) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.sc->n_agg_slots) return;
    counts[s] = d.a_alive[s] ? d.a_n[s] : 0;
}
__global__ void __launch_bounds__(256) k_compact_move(DevState d, DevState dst, const int *new_off) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 8 + w;
    if (slot >= d.sc->n_agg_slots || !d.a_alive[slot]) return;
    const int off = d.a_off[slot], n = d.a_n[slot], to = new_off[slot];
    for (int i = lane; i < n; i += 32) {
        dst.s_posr[to + i] = d.s_posr[off + i];
        dst.s_relv[to + i] = d.s_relv[off + i];
        dst.s_surf[to + i] = d.s_surf[off + i];
        dst.s_veff[to + i] = d.s_veff[off + i];
        dst.s_seff[to + i] = d.s_seff[off + i];
        dst.s_dcen[to + i] = d.s_dcen[off + i];
        const int id = d.s_id[off + i];
        dst.s_id[to + i] = id;
        dst.s_charge[to + i] = d.s_charge[off + i];
        d.slot_of_id[id] = to + i;
    }
    __syncwarp();
    if (lane == 0) d.a_off[slot] = to;
}
__global__ void k_compact_finish(DevState d, const int *new_off) {
    if (threadIdx.x == 0 && blockIdx.x == 0) d.sc->pool_top = new_off[d.sc->n_agg_slots];
}


// Domain duplication (AggregatList::duplication, aggregat_list.cpp:142-190), device part: the 7 copies of every
// aggregate (labels n0 + 7a + c-1, c = 4i+2j+k) are translated by (i,j,k)*old_box with the NEW box length, and every
// aggregate's Verlet cell is recomputed for the doubled box.  One warp per aggregate.
__global__ void __launch_bounds__(256) k_dup_finish(DevState d, int n0, double old_l) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 8 + w;
    if (slot >= d.sc->n_agg_slots) return;
    const double box = d.sc->box_length;
    if (slot >= n0) {
        const int c = (slot - n0) % 7 + 1;
        agg_translate<false>(d, slot, ((c >> 2) & 1) * old_l, ((c >> 1) & 1) * old_l, (c & 1) * old_l, box, lane, 32);
    } else if (lane == 0) {
        const double4 a = d.a_posr[slot];
        d.a_cx[slot] = cell_of(a.x, d.n_div, box);
        d.a_cy[slot] = cell_of(a.y, d.n_div, box);
        d.a_cz[slot] = cell_of(a.z, d.n_div, box);
    }
}

// ------------------------------------------------------------------------------------------------
// K9 — AggregatList::sort_time_steps on the device (aggregat_list.cpp:109-141).
// The reference sorts label indices with libstdc++'s std::sort (introsort) and the tie order among EQUAL 1/dt
// weights decides which aggregate a draw picks (in monodisperse runs whole classes tie).  std::sort's
// result is a deterministic function of the comparisons only, so it is replayed here in a level-synchronous form:
//   * __introsort_loop: every segment longer than 16 does median-of-3 -> __unguarded_partition.  The Hoare
//     partition is "k-th element from the left that is not < pivot swaps with the k-th from the right that is
//     not > pivot, while they have not crossed", i.e. two prefix counts + one pairing pass — parallel per level;
//   * __final_insertion_sort is a stable sort, and segments are already ordered relative to each other, so it
//     equals a stable sort inside each leaf (<= 16 elements).
// depth_limit exhaustion is reported through `fail`: the caller then runs k_sort_heap (libstdc++'s heap-sort branch).
// `stable` != 0 orders ties by label instead (MCAC_ORDER_STABLE).
// ------------------------------------------------------------------------------------------------
struct SortBufs {
    int *perm;          // labels being sorted
    double *wk;         // their weights, moved together with perm
    int *segf, *segl;   // per element: its current segment [segf, segl)
    long long *flags, *pre;  // low 32 bits: "not < pivot" ; high 32 bits: "not > pivot" ; and their exclusive scan
    int *tmp_a, *tmp_b, *cut;
    int *fin_perm;      // final order (cooperative kernel: all-equal segments are finished analytically, out of place)
    double *fin_wk;
    int *active;        // [0]: some segment still longer than 16 ; [1]: fail
    int n, stable;
};
constexpr int kSortLeaf = 16;
__device__ __forceinline__ bool w_less(double ka, int la, double kb, int lb, int stable) {
    return ka < kb || (stable && ka == kb && la < lb);
}
__global__ void k_sort_init(SortBufs b, const double *keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    b.perm[i] = i;
    b.wk[i] = keys[i];
    b.segf[i] = 0;
    b.segl[i] = b.n;
    if (i == 0) { b.active[0] = b.n > kSortLeaf ? 1 : 0; b.active[1] = 0; }
}
// __move_median_to_first(first, first+1, mid, last-1) by the segment leader
__global__ void k_sort_pivot(SortBufs b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n || b.segf[i] != i) return;
    const int f = i, l = b.segl[i];
    if (l - f <= kSortLeaf) return;
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
__global__ void k_sort_flags(SortBufs b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
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
__global__ void k_sort_scatter(SortBufs b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    const int f = b.segf[i], l = b.segl[i];
    if (l - f <= kSortLeaf || i <= f) return;
    const int base = f + 1;
    const long long p0 = b.pre[base], pi_ = b.pre[i], pl = b.pre[l];
    const long long fl = b.flags[i];
    if (fl & 1LL) b.tmp_a[base + (int)((pi_ & 0xffffffffLL) - (p0 & 0xffffffffLL))] = i;
    if (fl >> 32) {
        const int n_b = (int)((pl >> 32) - (p0 >> 32));
        b.tmp_b[base + n_b - 1 - (int)((pi_ >> 32) - (p0 >> 32))] = i;
    }
}
__global__ void k_sort_swap(SortBufs b) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= b.n) return;
    const int f = b.segf[j], l = b.segl[j];
    if (l - f <= kSortLeaf || j <= f) return;
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
    if (s >= 0) {  // where the two scans of __unguarded_partition stop after `s` swaps
        const int a_s = s < n_a ? b.tmp_a[base + s] : 0x7fffffff;
        const int b_prev = s > 0 ? b.tmp_b[base + s - 1] : l;
        b.cut[f] = a_s < b_prev ? a_s : b_prev;
    }
}
__global__ void k_sort_split(SortBufs b, int depth_left) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    const int f = b.segf[i], l = b.segl[i];
    if (l - f <= kSortLeaf) return;
    const int c = b.cut[f];
    int nf = f, nl = l;
    if (i < c) nl = c; else nf = c;
    b.segf[i] = nf;
    b.segl[i] = nl;
    if (i == nf && nl - nf > kSortLeaf) {
        if (depth_left > 0) b.active[0] = 1; else b.active[1] = 1;  // would enter the heap-sort branch of introsort
    }
}
// __final_insertion_sort restricted to a leaf: stable insertion sort of <= 16 elements by the leaf leader
__global__ void k_sort_leaves(SortBufs b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n || b.segf[i] != i) return;
    const int f = i, l = b.segl[i];
    if (l - f > kSortLeaf) return;
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
// introsort's depth limit (std::__introsort_loop, bits/stl_algo.h: `if (__depth_limit == 0) { std::__partial_sort(first, last, last); return; }`):
// every segment still longer than 16 when the limit is reached is heap-sorted — make_heap followed by sort_heap, restated here with
// libstdc++'s __adjust_heap / __push_heap so that equal weights leave in the same order.  One thread per segment (the branch is a
// safety net of introsort: it needs 2*log2(n) consecutive bad pivots); tests/native/heap_sort_host.cpp checks the same code against
// std::partial_sort on the host.
__global__ void k_sort_heap(SortBufs b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n || b.segf[i] != i) return;
    const int l = b.segl[i];
    if (l - i <= kSortLeaf) return;
    heapsort::heap_sort_segment(heapsort::HeapView{b.wk, b.perm, b.stable}, i, l);
}
// 64-bit packed exclusive scan (same 3-phase structure as the int scan)
__device__ __forceinline__ long long warp_inclusive_scan_ll(long long v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long n = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
__device__ __forceinline__ long long block_exclusive_scan_ll(long long v, long long *total, long long *warp_sums) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long inc = warp_inclusive_scan_ll(v, lane);
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        long long s = lane < nw ? warp_sums[lane] : 0;
        s = warp_inclusive_scan_ll(s, lane);
        warp_sums[lane] = s;
    }
    __syncthreads();
    const long long base = w ? warp_sums[w - 1] : 0;
    *total = warp_sums[nw - 1];
    __syncthreads();
    return base + inc - v;
}
__global__ void k_scan64_partials(const long long *in, int n, long long *block_sums) {
    __shared__ long long ws[32];
    const int base = blockIdx.x * kScanBlock * kScanItems + threadIdx.x * kScanItems;
    long long v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) v += (base + k < n) ? in[base + k] : 0;
    long long total;
    block_exclusive_scan_ll(v, &total, ws);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void k_scan64_block_sums(long long *block_sums, int nb) {
    __shared__ long long ws[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += blockDim.x) {
        const int i = b0 + threadIdx.x;
        const long long v = i < nb ? block_sums[i] : 0;
        long long total;
        const long long pre = block_exclusive_scan_ll(v, &total, ws);
        if (i < nb) block_sums[i] = carry + pre;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nb] = carry;
}
__global__ void k_scan64_apply(const long long *in, int n, const long long *block_sums, int nb, long long *out) {
    __shared__ long long ws[32];
    const int base = blockIdx.x * kScanBlock * kScanItems + threadIdx.x * kScanItems;
    long long x[kScanItems];
    long long v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { x[k] = (base + k < n) ? in[base + k] : 0; v += x[k]; }
    long long total;
    long long pre = block_exclusive_scan_ll(v, &total, ws) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = pre;
        pre += x[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_sums[nb];
}
// cumulative_time_steps (aggregat_list.cpp:133-140).  Sequential form = the reference's rounding, one thread.
__global__ void k_cum_sequential(const double *wk, double *cum, int n) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double acc = wk[0];
    cum[0] = acc;
    for (int i = 1; i < n; i++) { acc = acc + wk[i]; cum[i] = acc; }
}
// Parallel form for large N: fixed three-phase tree (deterministic; rounds differently from the sequential sum by a few ulp)
__global__ void k_cum_partials(const double *wk, int n, double *block_sums) {
    __shared__ double sm[32];
    const int base = blockIdx.x * kScanBlock * kScanItems + threadIdx.x * kScanItems;
    double v = 0.;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) v += (base + k < n) ? wk[base + k] : 0.;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int w = 0; w < kScanBlock / 32; w++) t += sm[w];
        block_sums[blockIdx.x] = t;
    }
}
__global__ void k_cum_block_sums(double *block_sums, int nb) {  // exclusive, sequential over <= few hundred blocks
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double acc = 0.;
    for (int b = 0; b < nb; b++) { const double t = block_sums[b]; block_sums[b] = acc; acc += t; }
}
__global__ void k_cum_apply(const double *wk, int n, const double *block_sums, double *cum) {
    __shared__ double warp_tot[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int base = blockIdx.x * kScanBlock * kScanItems + threadIdx.x * kScanItems;
    double x[kScanItems];
    double v = 0.;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { x[k] = (base + k < n) ? wk[base + k] : 0.; v += x[k]; }
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    double wbase = 0.;
    for (int ww = 0; ww < w; ww++) wbase += warp_tot[ww];
    double run = block_sums[blockIdx.x] + wbase + (inc - v);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        run += x[k];
        if (base + k < n) cum[base + k] = run;
    }
}
__global__ void k_sort_finish(DevState d, SortBufs b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    d.sorted_slot[i] = d.slot_of_label[b.perm[i]];
    if (i == 0) { d.sc->n_pick = b.n; d.sc->cum_total = d.cum[b.n - 1]; d.sc->pick_dense_from = b.n; }
}

// ------------------------------------------------------------------------------------------------
// The whole per-event pipeline in ONE cooperative launch (grid barriers instead of ~180 launches + host polls):
// labels (rank of live slots) -> refresh()/PhysicalModel::update -> 1/dt weights -> replayed introsort ->
// cumulative table -> pick table in slots.  Same arithmetic and same visiting orders as the stand-alone kernels
// above; every scan is a chunk-contiguous two-phase scan (block b owns elements [b*chunk, (b+1)*chunk)).
// ------------------------------------------------------------------------------------------------
constexpr int kEventThreads = 512;
// layout of the scratch arrays: part_ll (4096 entries) = [0, gridDim) per-block counts / chunk sums of the level scans, [kPartSparseCounts, +gridDim)
// sparse elements staged per block; part_d (16384) = five per-block partials of gridDim entries each, [kPartCumChunks, +8192) chunk sums of the cumulative table
constexpr int kPartSparseCounts = 2048, kPartCumChunks = 8192;
// exact cumulative table of a tie-dominated pick table (seq_cumsum.cuh): part_ll[kPartCumSegN] = segments of the W run (-1: none, the
// summation tree is used), part_d[kPartCumSegs, + 3 * kMaxSegs) = the segments
constexpr int kPartCumSegN = 4000, kPartCumSegs = 6144;
constexpr int kMaxWin = 512, kWinMin = 256, kWinBase = 16;  // SortBufs::active holds kWinBase + 4 * kMaxWin entries
struct EventArgs {
    SortBufs sb;
    long long *part_ll;   // >= gridDim + 1
    double *part_d;       // >= 4 * (gridDim + 1)
    int *scan_tmp;        // >= n_agg_slots + 1 (exclusive scan of a_alive)
    int *sorted_label;    // the reference's index_sorted_time_steps (labels), for mcac_gpu_get_pick_table
    int do_labels, do_refresh /* 1: max_time_step + avg_npp */, do_totals /* PhysicalModel::update */, do_sort;
    int cum_sequential_max, stable;
    int use_factor;       // sort_time_steps(factor) called with an explicit factor (per-call C ABI)
    double factor;
    int local_span;       // span (elements) that fits the shared-memory staging of the block-local levels
    int switch_span;      // span (elements) below which block 0 finishes the sort alone (>= local_span)
    int smem_cap;         // entries of dynamic shared memory per array available to the block-local levels (0 = none)
    int force_fail;       // test hook: report introsort's depth-limit failure although the sort succeeded
    int depth_override;   // test hook (MCAC_B200_SORT_DEPTH): introsort depth limit instead of 2*log2(n); < 0 = off
    int no_windows;       // tuning hook (MCAC_B200_NO_SORT_WINDOWS): the local levels by block 0 alone, as before
    int win_cap;          // a window is taken by its CTA once every window spans at most this many elements (<= smem_cap)
    long long *work;      // [0] += sum over levels of the active span (elements touched by the level passes), [1] += levels
    // tie-dominated tables (tie_sort.cuh): top levels simulated on the sparse elements only
    tiesort::Plan *ts_plan;
    int *ts_R, *ts_tbl;   // (kMaxLevels + 1) x ts_xcap sorted positions / x kTblStride bucket tables
    int ts_xcap;          // 0 = fast path off
    int ts_min_n;         // smallest table the fast path is tried on
    int smem_bytes;       // dynamic shared memory of the launch
    int skip_if_no_event; // return at once when Scalars::event == 0 (nothing changed since the last pick table)
    int ts_no_overlap;    // test / tuning hook: all CTAs route, then the general sort starts (no overlap with the block-local levels)
    int exact_cum_max_sparse;  // most sparse elements the exact cumulative table is built for (above: the summation tree)
    int no_exact_cum;     // tuning hook (MCAC_B200_NO_EXACT_CUM): cumulative table of a big tie-dominated table by the summation tree, as before
};
struct BlockTeam {  // tiesort's Team for one CTA
    int tid, nthr;
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void team_min(int *p, int v) {  // called by every thread of the CTA
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(kFull, v, o); v = t < v ? t : v; }
        if ((tid & 31) == 0 && v != 0x7fffffff) atomicMin(p, v);
    }
    __device__ __forceinline__ void lap(int) {}
    // called after a CTA barrier: what the team wrote before it becomes visible device-wide, then the progress word
    __device__ __forceinline__ void publish(int *p, int v) {
        if (tid == 0) { __threadfence(); *reinterpret_cast<volatile int *>(p) = v; }
    }
    __device__ __forceinline__ void add(int *p, int v) { atomicAdd(p, v); }  // (shared memory; result unused: a reduction)
    // exclusive scan over the CTA's threads; scratch: >= 32 ints of shared memory, free again after the next CTA barrier
    __device__ __forceinline__ int excl_scan(int v, int *scratch) {
        const int lane = tid & 31, w = tid >> 5;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) scratch[w] = inc;
        __syncthreads();
        int ws = lane < (nthr >> 5) ? scratch[lane] : 0;  // every warp scans the (<= 32) warp totals itself
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, ws, o); if (lane >= o) ws += t; }
        const int base = __shfl_sync(kFull, ws, w > 0 ? w - 1 : 0);
        return (w > 0 ? base : 0) + inc - v;
    }
};
struct ProbeTeam {  // BlockTeam + phase clocks of thread 0 (k_plan_probe)
    int tid, nthr;
    long long *acc, prev;
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void team_min(int *p, int v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(kFull, v, o); v = t < v ? t : v; }
        if ((tid & 31) == 0 && v != 0x7fffffff) atomicMin(p, v);
    }
    __device__ __forceinline__ void lap(int k) {
        if (tid == 0) { const long long t = clock64(); atomicAdd(reinterpret_cast<unsigned long long *>(acc + k), (unsigned long long)(t - prev)); prev = t; }
    }
    __device__ __forceinline__ void publish(int *p, int v) {
        if (tid == 0) { __threadfence(); *reinterpret_cast<volatile int *>(p) = v; }
    }
    __device__ __forceinline__ void add(int *p, int v) { atomicAdd(p, v); }  // (shared memory; result unused: a reduction)
    // exclusive scan over the CTA's threads; scratch: >= 32 ints of shared memory, free again after the next CTA barrier
    __device__ __forceinline__ int excl_scan(int v, int *scratch) {
        const int lane = tid & 31, w = tid >> 5;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) scratch[w] = inc;
        __syncthreads();
        int ws = lane < (nthr >> 5) ? scratch[lane] : 0;  // every warp scans the (<= 32) warp totals itself
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, ws, o); if (lane >= o) ws += t; }
        const int base = __shfl_sync(kFull, ws, w > 0 ? w - 1 : 0);
        return (w > 0 ? base : 0) + inc - v;
    }
};
namespace cgx = cooperative_groups;
extern __shared__ __align__(16) unsigned char dyn_smem[];
constexpr int kSortStageBytesPerEntry = 48;  // wk, flags, pre (8 B each) + perm, segf, segl, tmp_a, tmp_b, cut (4 B each)

__device__ __forceinline__ double block_sum_fixed(double v, double *sm /* >= 32 */) {  // fixed tree: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sm[w];
    __syncthreads();
    return t;
}
__device__ __forceinline__ double block_max_fixed(double v, double *sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(kFull, v, o); v = (v < t) ? t : v; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t = (t < sm[w]) ? sm[w] : t;
    __syncthreads();
    return t;
}
__device__ __forceinline__ double block_min_fixed(double v, double *sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(kFull, v, o); v = (t < v) ? t : v; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = sm[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) t = (sm[w] < t) ? sm[w] : t;
    __syncthreads();
    return t;
}
__device__ __forceinline__ long long block_sum_ll(long long v, long long *sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sm[w];
    __syncthreads();
    return t;
}

template <int kMinBlocks>
__global__ void __launch_bounds__(kEventThreads, kMinBlocks) k_event(DevState d, EventArgs a) {
    cgx::grid_group grid = cgx::this_grid();
    __shared__ long long sm_ll[32];
    __shared__ double sm_d[32];
    __shared__ long long sh_carry;
    const int tid = threadIdx.x, nthr = blockDim.x, nblk = gridDim.x, blk = blockIdx.x;
    const long long gtid = (long long)blk * nthr + tid, gsize = (long long)nblk * nthr;
    Scalars &sc = *d.sc;
    if (a.skip_if_no_event && sc.event == 0) return;  // submitted ahead of the read-back of a batch that turned out not to merge
    if (gtid == 0 && a.ts_plan) { a.ts_plan->ready = 0; a.ts_plan->done = 0; a.ts_plan->cum_gathered = 0; }  // (two grid barriers before anybody looks at them)
    const int n_slots = sc.n_agg_slots;
    // phase clocks of block 0 (SM cycles) accumulated into a.work[2 + k]: diagnostics for the K9 breakdown in profiles/
    long long t_prev = clock64();
    // (atomicAdd without a use of its result is a fire-and-forget reduction: no round trip to L2 on the critical path)
    auto work_add = [&](int k, long long v) { atomicAdd(reinterpret_cast<unsigned long long *>(a.work + k), (unsigned long long)v); };
    auto lap = [&](int k) { if (gtid == 0 && a.work) { const long long t = clock64(); work_add(2 + k, t - t_prev); t_prev = t; } };

    // ---------------- phase A: live-slot count per chunk + refresh partials
    const int chunk_s = ((n_slots + nblk - 1) / nblk + nthr - 1) / nthr * nthr;
    {
        long long cnt = 0;
        double mx = 0., sv = 0., ss = 0., mn = __longlong_as_double(0x7ff0000000000000LL);
        const int lo = blk * chunk_s, hi = min(n_slots, lo + chunk_s);
        for (int s0 = lo + tid; s0 < hi; s0 += 4 * nthr) {  // four rounds of loads in flight; same summation order as one by one
            int al[4];
            double ts[4], vv[4], sf[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int s = s0 + j * nthr;
                const bool ok = s < hi;
                al[j] = ok ? d.a_alive[s] : 0;
                ts[j] = ok ? d.a_ts[s] : 0.;
                vv[j] = ok ? d.a_vol[s] : 0.;
                sf[j] = ok ? d.a_surf[s] : 0.;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (!al[j]) continue;
                cnt += 1;
                mx = (mx < ts[j]) ? ts[j] : mx;
                mn = (ts[j] < mn) ? ts[j] : mn;
                sv += vv[j];
                ss += sf[j];
            }
        }
        const long long c = block_sum_ll(cnt, sm_ll);
        const double bmx = block_max_fixed(mx, sm_d), bsv = block_sum_fixed(sv, sm_d), bss = block_sum_fixed(ss, sm_d);
        const double bmn = block_min_fixed(mn, sm_d);
        if (tid == 0) {
            a.part_ll[blk] = c;
            a.part_d[blk] = bmx;
            a.part_d[nblk + blk] = bsv;
            a.part_d[2 * nblk + blk] = bss;
            a.part_d[4 * nblk + blk] = bmn;
        }
    }
    grid.sync();
    lap(0);
    // ---------------- phase B: labels; every block derives n_agg / max / totals in the same fixed order
    int n_agg = 0;
    double factor = sc.max_time_step;
    bool ts_try = false;
    double ts_W = 0.;
    __shared__ int sh_exc;
    constexpr int kMaxCoopBlocks = 320;
    __shared__ long long sh_pl[kMaxCoopBlocks], sh_res_ll[2];
    __shared__ double sh_pd[4][kMaxCoopBlocks], sh_res_d[4];
    {
        // the per-block partials are fetched once, in parallel, into shared memory; warp 0 then combines them in block order
        // (sequential, the same in every block: deterministic) and publishes the results
        long long base = 0, total = 0;
        double mx = 0., sv = 0., ss = 0., mn = __longlong_as_double(0x7ff0000000000000LL);
        if (nblk <= kMaxCoopBlocks) {
            for (int b = tid; b < nblk; b += nthr) {
                sh_pl[b] = a.part_ll[b];
                sh_pd[0][b] = a.part_d[b];
                sh_pd[1][b] = a.part_d[nblk + b];
                sh_pd[2][b] = a.part_d[2 * nblk + b];
                sh_pd[3][b] = a.part_d[4 * nblk + b];
            }
            __syncthreads();
            if (tid < 32) {
                for (int b = 0; b < nblk; b++) {
                    const long long c = sh_pl[b];
                    if (b < blk) base += c;
                    total += c;
                    const double t = sh_pd[0][b];
                    mx = (mx < t) ? t : mx;
                    const double tm = sh_pd[3][b];
                    mn = (tm < mn) ? tm : mn;
                    sv += sh_pd[1][b];
                    ss += sh_pd[2][b];
                }
                if (tid == 0) { sh_res_ll[0] = base; sh_res_ll[1] = total; sh_res_d[0] = mx; sh_res_d[1] = mn; sh_res_d[2] = sv; sh_res_d[3] = ss; }
            }
            __syncthreads();
            base = sh_res_ll[0]; total = sh_res_ll[1]; mx = sh_res_d[0]; mn = sh_res_d[1]; sv = sh_res_d[2]; ss = sh_res_d[3];
        } else {
            for (int b = 0; b < nblk; b++) {
                const long long c = a.part_ll[b];
                if (b < blk) base += c;
                total += c;
                const double t = a.part_d[b];
                mx = (mx < t) ? t : mx;
                const double tm = a.part_d[4 * nblk + b];
                mn = (tm < mn) ? tm : mn;
                sv += a.part_d[nblk + b];
                ss += a.part_d[2 * nblk + b];
            }
        }
        n_agg = (int)total;
        if (a.do_refresh) factor = mx;
        if (a.use_factor) factor = a.factor;
        // tie-dominated fast path (tie_sort.cuh): W = the largest weight = factor / (smallest time step); the elements below W are
        // staged per chunk in label order (labels in tmp_a, weights in pre) while the labels are made
        ts_try = a.do_sort && a.ts_xcap > 0 && !a.stable && n_agg >= a.ts_min_n && n_agg > a.local_span;
        ts_W = factor / mn;
        if (a.do_labels || ts_try) {
            // four consecutive slots per thread and round: one block scan per 4 * blockDim slots
            const int lo = blk * chunk_s, hi = min(n_slots, lo + chunk_s);
            double *st_w = reinterpret_cast<double *>(a.sb.pre);
            unsigned char *is_sparse = reinterpret_cast<unsigned char *>(a.sb.cut);  // per label, read by the routing pass
            if (tid == 0) { sh_carry = base; sh_exc = 0; }
            __syncthreads();
            for (int t0 = lo; t0 < hi; t0 += 4 * nthr) {
                const int s0 = t0 + 4 * tid;
                int al[4], ex[4];
                double w[4];
                if (s0 + 3 < hi) {
                    const int4 v = *reinterpret_cast<const int4 *>(d.a_alive + s0);  // lo and t0 are multiples of 512
                    al[0] = v.x; al[1] = v.y; al[2] = v.z; al[3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) al[j] = (s0 + j < hi) ? d.a_alive[s0 + j] : 0;
                }
                int na = 0, ne = 0;
                double tsv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) tsv[j] = (ts_try && s0 + j < hi) ? d.a_ts[s0 + j] : 1.;  // not behind the liveness load
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    w[j] = 0.;
                    ex[j] = 0;
                    if (ts_try && al[j]) { w[j] = factor / tsv[j]; ex[j] = (w[j] != ts_W) ? 1 : 0; }
                    na += al[j];
                    ne += ex[j];
                }
                int tot;
                __shared__ int ws[32];
                const int pre2 = block_exclusive_scan(na | (ne << 16), &tot, ws);  // <= 2048 of either per round
                int lab = (int)sh_carry + (pre2 & 0xffff), k = sh_exc + (pre2 >> 16);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int s = s0 + j;
                    if (s >= hi) break;
                    if (a.do_labels) {
                        if (al[j]) { d.label_of_slot[s] = lab; d.slot_of_label[lab] = s; }
                        else d.label_of_slot[s] = -1;
                    }
                    if (ts_try && al[j]) is_sparse[lab] = (unsigned char)ex[j];
                    if (ex[j]) {
                        if (k < tiesort::kMaxSparse) { a.sb.tmp_a[lo + k] = lab; st_w[lo + k] = w[j]; }
                        k++;
                    }
                    lab += al[j];
                }
                __syncthreads();
                if (tid == 0) { sh_carry += tot & 0xffff; sh_exc += tot >> 16; }
                __syncthreads();
            }
            if (tid == 0) a.part_ll[kPartSparseCounts + blk] = sh_exc;
        }
        if (gtid == 0) {
            if (a.do_refresh) {
                sc.max_time_step = mx;
                sc.avg_npp = static_cast<double>(sc.n_sph) / static_cast<double>(n_agg);
            }
            if (a.do_totals) {
                sc.total_volume = sv;
                sc.total_surface = ss;
                sc.total_volume_concent = sv / sc.box_volume;
                sc.total_surface_concent = ss / sc.box_volume;
                sc.aggregate_concentration = static_cast<double>(n_agg) / sc.box_volume;
                sc.monomer_concentration = static_cast<double>(sc.n_sph) / sc.box_volume;
                sc.volume_fraction = sv / sc.box_volume;
            }
        }
    }
    if (!a.do_sort) { lap(1); return; }
    grid.sync();
    lap(1);
    // ---------------- phase C: weights in label order + sort state
    SortBufs b = a.sb;
    const int n = n_agg;
    b.n = n;
    b.stable = a.stable;
    int lg = 0;
    while ((1LL << (lg + 1)) <= n) lg++;
    int depth = a.depth_override >= 0 ? a.depth_override : 2 * lg;
    int n_sort = n, delta = 0;  // the general sort below works on [0, n_sort) and writes its result at +delta
    // ---- tie-dominated table: the top levels on the sparse elements only (tie_sort.cuh)
    bool ts_on = false, ts_ovl = false;
    // exact_cum: the cumulative table of a big tie-dominated table is the reference's sequential sum, bit for bit (seq_cumsum.cuh).  Only
    // the sorted VALUES matter for it, and those are known without the introsort replay: the sparse (lighter) weights in ascending
    // order, then W over and over.  The LAST CTA sorts the sparse weights on its own (bitonic, shared memory), adds them up one after
    // the other, and walks the binades of the W run — beside the sparse simulation and the sort levels, off the critical path; behind
    // the sort every thread of the grid evaluates its entries of the W run in closed form.
    bool exact_cum = false;
    int cum_xs = 0, cum_P = 1;
    if (ts_try) {
        long long x = 0;
        for (int bb = tid; bb < nblk; bb += nthr) x += a.part_ll[kPartSparseCounts + bb];
        x = block_sum_ll(x, sm_ll);
        // shared memory of the simulating CTA: four lists of x entries, one bucket table, and behind them an archive of the
        // per-level tables (as many levels as fit) for the walk back from the handed-over segment
        int ts_nb = 256;
        while (ts_nb < x && ts_nb < tiesort::kBuckets) ts_nb <<= 1;
        const int xs_pad = ((int)x + 2 + 3) & ~3;  // (the lists are read up to two entries past their end)
        const int used_ints = 4 * xs_pad + (ts_nb + 4) + tiesort::kMiscInts + 2 * (nblk + 1) + 4;
        ts_on = x <= a.ts_xcap && x <= tiesort::kMaxSparse && used_ints * (int)sizeof(int) <= a.smem_bytes &&
                a.smem_bytes >= 4 * nthr * (int)sizeof(int);  // (the routing pass parks >= 4 positions per thread in shared memory)
        if (ts_on) {
            const int xs = (int)x;
            cum_xs = xs;
            cum_P = 64;
            while (cum_P < xs) cum_P <<= 1;
            {   // shared memory of the building CTA (laid out in cum_exact_build)
                const int pp = seqsum::padded_size(cum_P);
                const int need = pp * (int)(sizeof(double) + sizeof(long long) + sizeof(unsigned short)) + seqsum::kMaxSegs * (int)sizeof(seqsum::Seg) +
                                 (nblk + 4) * (int)sizeof(int) + seqsum::kMaxIrr * (int)(sizeof(double) + 2 * sizeof(int)) + nthr * (int)sizeof(double) + 64;
                exact_cum = n > a.cum_sequential_max && !a.no_exact_cum && nblk >= 4 && xs <= a.exact_cum_max_sparse && need <= a.smem_bytes;
            }
            int *st_pos = b.tmp_b;                                 // compact staged labels / weights of the sparse elements
            double *st_w = reinterpret_cast<double *>(b.flags);
            // (the sparse elements' weights first, when there is room: the pivot samples of every level read them)
            const bool w_smem = (used_ints + 2 * xs_pad) * (int)sizeof(int) <= a.smem_bytes;
            double *s_w = reinterpret_cast<double *>(dyn_smem);
            int *sm = reinterpret_cast<int *>(dyn_smem) + (w_smem ? 2 * xs_pad : 0);
            int *a_s = sm, *a_i = a_s + xs_pad, *b_s = a_i + xs_pad, *b_i = b_s + xs_pad, *s_tbl = b_i + xs_pad, *s_misc = s_tbl + ts_nb + 4,
                *s_base = s_misc + tiesort::kMiscInts, *arch_R = s_base + 2 * (nblk + 1) + 4;
            const int arch_levels = min(tiesort::kMaxLevels, (a.smem_bytes / (int)sizeof(int) - used_ints - (w_smem ? 2 * xs_pad : 0)) / (xs + ts_nb + 3 + 1));
            int *arch_T = arch_R + (size_t)arch_levels * xs;
            if (blk == 0) {
                __shared__ int ts_ws[32];
                for (int b0 = 0; b0 < nblk; b0 += nthr) {  // exclusive scan of the per-chunk counts (one round: nblk <= blockDim)
                    const int bb = b0 + tid;
                    const int c = bb < nblk ? (int)a.part_ll[kPartSparseCounts + bb] : 0;
                    int tot;
                    const int pre = block_exclusive_scan(c, &tot, ts_ws);
                    const int carry = b0 == 0 ? 0 : s_base[b0];
                    if (bb < nblk) s_base[bb] = carry + pre;
                    if (tid == 0) s_base[min(b0 + nthr, nblk)] = carry + tot;
                    __syncthreads();
                }
                const double *chunk_w = reinterpret_cast<const double *>(b.pre);
                for (int id0 = tid; id0 < xs; id0 += 4 * nthr) {  // gather the per-chunk stages into one ascending list (four loads in flight)
                    int pv[4];
                    double wv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int id = id0 + u * nthr;
                        int lo = 0, hi = nblk;
                        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_base[mid] <= id) lo = mid; else hi = mid; }
                        const int src = id < xs ? lo * chunk_s + (id - s_base[lo]) : 0;
                        pv[u] = b.tmp_a[src];
                        wv[u] = chunk_w[src];
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int id = id0 + u * nthr;
                        if (id >= xs) break;
                        st_pos[id] = pv[u];
                        st_w[id] = wv[u];
                        if (w_smem) s_w[id] = wv[u];
                    }
                }
                if (tid == 0) { b.active[0] = 0; b.active[1] = 0; b.active[2] = 0; b.active[3] = 0; }  // (before the first level is published)
                __syncthreads();
                lap(14);
                BlockTeam tm{tid, nthr};
                tiesort::plan_build(tm, n, xs, st_pos, w_smem ? s_w : st_w, ts_W, depth, a.local_span, a.ts_plan, a.ts_R, a.ts_tbl, a.ts_xcap, a_s, a_i,
                                    b_s, b_i, s_tbl, s_misc, arch_R, arch_T, arch_levels);
                lap(15);
                // the sparse elements of the handed-over segment
                const int hf = a.ts_plan->hand_f, hlen = a.ts_plan->hand_l - hf;
                // overlap: this CTA fills the W elements of a small handed-over segment itself (walk back through the archived
                // tables) and goes straight on to the block-local sort levels, while the other CTAs route the rest of the table
                if (tid == 0)
                    a.ts_plan->overlap = (a.ts_plan->n_levels <= arch_levels && hlen <= a.local_span && nblk > 1 && !a.ts_no_overlap) ? 1 : 0;
                __syncthreads();
                tm.publish(&a.ts_plan->done, 1);  // the plan is final: the routing CTAs leave their level loop
                for (int j0 = tid; j0 < xs; j0 += 4 * nthr) {  // (four loads in flight)
                    int pp[4], lab[4];
                    double wv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int j = j0 + u * nthr < xs ? j0 + u * nthr : tid;
                        const int id = a_i[j];
                        pp[u] = a_s[j] - hf;
                        lab[u] = st_pos[id];
                        wv[u] = w_smem ? s_w[id] : st_w[id];
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (j0 + u * nthr >= xs) break;
                        b.perm[pp[u]] = lab[u];
                        b.wk[pp[u]] = wv[u];
                        b.segf[pp[u]] = 0;
                        b.segl[pp[u]] = hlen;
                    }
                }
            }
            // Routing pass of the other CTAs: every W element follows the simulated levels to its final position or to its place in
            // the handed-over segment.  It runs BESIDE the simulation: a CTA waits for level t to be published (Plan::ready), moves
            // all its elements through it (positions parked in shared memory between levels, kRoute elements advanced together so
            // that their independent table look-ups are in flight at the same time), and is one level behind when the simulation ends.
            __shared__ tiesort::Plan sh_plan;
            __shared__ int sh_ready, sh_done;
            const int *__restrict__ R = a.ts_R;
            const int *__restrict__ T = a.ts_tbl;
            const unsigned char *__restrict__ is_sparse = reinterpret_cast<const unsigned char *>(b.cut);
            bool any_bad = false;
            constexpr int kRoute = 4;
            // (the CTA that builds the exact cumulative table does not route: exact_cum asks for a grid of four CTAs at least)
            const bool cum_builder = exact_cum && blk == nblk - 1, cum_dedicated = exact_cum;
            const bool router = nblk > 1 ? (blk != 0 && !(cum_builder && cum_dedicated)) : true;  // (a one-CTA launch routes after its own simulation)
            const long long rtid = nblk > 1 ? gtid - nthr : gtid, rsize = nblk > 1 ? gsize - (cum_dedicated ? 2 : 1) * nthr : gsize;
            auto cum_exact_build = [&]() {
                const long long t_cb = clock64();
                // dynamic shared memory: sparse weights in ascending order, then their running sums (padded view: the threads of a warp walk
                // chunks of consecutive entries) | segments of the W run | first sparse element of every chunk | K of the parallel head
                // (seq_cumsum.cuh; first the plain array the weights are sorted in) | exact sums at the irregular steps | approximate sums
                // behind the threads' chunks | the irregular steps' indexes / binades | the entries' stretch
                const int pp = seqsum::padded_size(cum_P);
                double *s_v = reinterpret_cast<double *>(dyn_smem);
                seqsum::Seg *s_seg = reinterpret_cast<seqsum::Seg *>(s_v + pp);
                int *s_cb = reinterpret_cast<int *>(s_seg + seqsum::kMaxSegs);
                long long *s_K = reinterpret_cast<long long *>(s_cb + ((nblk + 2 + 1) & ~1));
                double *s_base = reinterpret_cast<double *>(s_K + pp);
                double *s_endp = s_base + seqsum::kMaxIrr;
                int *s_iidx = reinterpret_cast<int *>(s_endp + nthr);
                int *s_ie = s_iidx + seqsum::kMaxIrr;
                unsigned short *s_c = reinterpret_cast<unsigned short *>(s_ie + seqsum::kMaxIrr);
                double *s_sort = reinterpret_cast<double *>(s_K);  // (K is not needed before the sort is over)
                const seqsum::Padded<double> v_p{s_v};
                const seqsum::Padded<long long> K_p{s_K};
                const seqsum::Padded<unsigned short> c_p{s_c};
                __shared__ int cb_ws[32];
                __shared__ int cb_ns;
                __syncthreads();
                for (int b0 = 0; b0 < nblk; b0 += nthr) {  // exclusive scan of the per-chunk counts
                    const int bb = b0 + tid;
                    const int c = bb < nblk ? (int)a.part_ll[kPartSparseCounts + bb] : 0;
                    int tot;
                    const int pre = block_exclusive_scan(c, &tot, cb_ws);
                    const int carry = b0 == 0 ? 0 : s_cb[b0];
                    if (bb < nblk) s_cb[bb] = carry + pre;
                    if (tid == 0) s_cb[min(b0 + nthr, nblk)] = carry + tot;
                    __syncthreads();
                }
                const double *chunk_w = reinterpret_cast<const double *>(b.pre);
                for (int id0 = tid; id0 < cum_P; id0 += 4 * nthr) {  // the per-chunk stages into one list (four loads in flight), +inf behind it
                    double wv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int id = id0 + u * nthr;
                        int lo = 0, hi = nblk;
                        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_cb[mid] <= id) lo = mid; else hi = mid; }
                        wv[u] = id < xs ? chunk_w[lo * chunk_s + (id - s_cb[lo])] : __longlong_as_double(0x7ff0000000000000LL);
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int id = id0 + u * nthr;
                        if (id < cum_P) s_sort[id] = wv[u];
                    }
                }
                __syncthreads();
                const int lane_ = tid & 31, warp_ = tid >> 5, nwarp_ = nthr >> 5;
                long long t_cl = t_cb;
                auto cb_lap = [&](int k) { if (a.work && tid == 0) { const long long t = clock64(); work_add(k, t - t_cl); t_cl = t; } };
                cb_lap(32);
                // (the stages lie in scratch of the sort levels: the simulating CTA waits for this word before it starts its own levels)
                if (tid == 0) *reinterpret_cast<volatile int *>(&a.ts_plan->cum_gathered) = 1;
                // Ascending sort (any correct sort gives the reference's sequence of values): a bitonic network in the padded array, three
                // steps per pass in registers (bitonic_pass).  Its result is checked — sorted, and the same multiset (two sums over the bit
                // patterns) — and a result that fails is redone from the plain copy by the one-step-at-a-time network below.
                auto pattern_sums = [&](auto &&at, long long &hi_sum, long long &lo_sum, long long &unsorted) {
                    long long h_ = 0, l_ = 0, u_ = 0;
                    for (int i = tid; i < cum_P; i += nthr) {
                        const double x = at(i);
                        const long long bits = __double_as_longlong(x);
                        h_ += bits >> 20;
                        l_ += bits & 0xfffff;
                        if (i + 1 < cum_P && x > at(i + 1)) u_++;
                    }
                    hi_sum = block_sum_ll(h_, sm_ll);
                    lo_sum = block_sum_ll(l_, sm_ll);
                    unsorted = block_sum_ll(u_, sm_ll);
                };
                long long in_hi, in_lo, in_uns;
                pattern_sums([&](int i) { return s_sort[i]; }, in_hi, in_lo, in_uns);
                for (int i = tid; i < cum_P; i += nthr) v_p[i] = s_sort[i];
                __syncthreads();
                for (int k = 2, lg = 1; k <= cum_P; k <<= 1, lg++)
                    for (int top = lg - 1; top >= 0;) {
                        const int g = min(3, top + 1), bb = top - g + 1;
                        if (g == 3) seqsum::bitonic_pass<3>(v_p, cum_P, k, bb, tid, nthr);
                        else if (g == 2) seqsum::bitonic_pass<2>(v_p, cum_P, k, bb, tid, nthr);
                        else seqsum::bitonic_pass<1>(v_p, cum_P, k, bb, tid, nthr);
                        __syncthreads();
                        top -= g;
                    }
                long long out_hi, out_lo, out_uns;
                pattern_sums([&](int i) { return v_p[i]; }, out_hi, out_lo, out_uns);
                if (out_hi != in_hi || out_lo != in_lo || out_uns != 0) {  // (block-uniform)
                    // the steps whose partner is less than a tile away stay inside one warp's tile (warp barriers only); the pairs of
                    // a step are disjoint: a thread loads up to four of them before it stores any
                    auto cmp_swap4 = [&](int t0, int stride, int cnt, int j, int k) {  // pairs t0, t0 + stride, .. (cnt <= 4 of them)
                        int ii[4];
                        double va[4], vb[4];
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (u < cnt) {
                                const int t = t0 + u * stride;
                                ii[u] = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                                va[u] = s_sort[ii[u]];
                                vb[u] = s_sort[ii[u] | j];
                            }
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (u < cnt && (va[u] > vb[u]) == ((ii[u] & k) == 0)) { s_sort[ii[u]] = vb[u]; s_sort[ii[u] | j] = va[u]; }
                    };
                    const int tile = max(64, cum_P / nwarp_), half = tile >> 1;  // (cum_P >= 64; at most one tile per warp)
                    auto tile_steps = [&](int k, int j_from) {  // steps j_from, j_from / 2, .., 1 of stage k, every tile by one warp
                        for (int tl = warp_; tl < cum_P / tile; tl += nwarp_)
                            for (int j = j_from; j > 0; j >>= 1) {
                                for (int t0 = tl * half + lane_; t0 < (tl + 1) * half; t0 += 4 * 32)
                                    cmp_swap4(t0, 32, min(4, ((tl + 1) * half - t0 + 31) / 32), j, k);
                                __syncwarp();
                            }
                    };
                    for (int k = 2; k <= tile; k <<= 1) tile_steps(k, k >> 1);
                    __syncthreads();
                    for (int k = 2 * tile; k <= cum_P; k <<= 1) {
                        for (int j = k >> 1; j >= tile; j >>= 1) {
                            for (int t0 = tid; t0 < (cum_P >> 1); t0 += 4 * nthr) cmp_swap4(t0, nthr, min(4, ((cum_P >> 1) - t0 + nthr - 1) / nthr), j, k);
                            __syncthreads();
                        }
                        tile_steps(k, half);
                        __syncthreads();
                    }
                    for (int i = tid; i < cum_P; i += nthr) v_p[i] = s_sort[i];
                    if (a.work && tid == 0) work_add(36, 1);
                    __syncthreads();
                }
                cb_lap(33);
                __shared__ int cb_head_ok;
                {
                    // the head's sequential sums as integer prefix sums between the irregular steps (seq_cumsum.cuh): a thread owns E
                    // consecutive elements
                    const int E = max(1, cum_P / nthr), lo = min(xs, tid * E), hi = min(xs, lo + E);
                    __shared__ double cb_wd[32];
                    __shared__ long long cb_wt[32];
                    __shared__ int cb_wf[32], cb_wn[32];
                    const double ls = seqsum::head_chunk_sum(v_p, lo, hi, 0.);
                    double inc = ls;  // approximate sum before the chunk: warp scan + the warps before this one
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(kFull, inc, o); if (lane_ >= o) inc += t; }
                    if (lane_ == 31) cb_wd[warp_] = inc;
                    __syncthreads();
                    double pex = inc - ls;
                    for (int w = 0; w < warp_; w++) pex += cb_wd[w];
                    s_endp[tid] = seqsum::head_chunk_sum(v_p, lo, hi, pex);
                    __syncthreads();
                    const seqsum::ChunkAgg g = seqsum::head_chunk_classify(v_p, lo, hi, pex, tid > 0 ? s_endp[tid - 1] : 0., K_p, c_p);
                    // exclusive scan of the chunk aggregates (agg_combine) over the threads
                    int f = g.has_irr, ni = g.n_irr;
                    long long tl = g.tail;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int f2 = __shfl_up_sync(kFull, f, o), n2 = __shfl_up_sync(kFull, ni, o);
                        const long long t2 = __shfl_up_sync(kFull, tl, o);
                        if (lane_ >= o) { if (!f) tl += t2; f |= f2; ni += n2; }
                    }
                    if (lane_ == 31) { cb_wf[warp_] = f; cb_wt[warp_] = tl; cb_wn[warp_] = ni; }
                    __syncthreads();
                    int pf = 0, pn = 0;  // everything before this warp
                    long long pt = 0;
                    for (int w = 0; w < warp_; w++) { pt = cb_wf[w] ? cb_wt[w] : pt + cb_wt[w]; pf |= cb_wf[w]; pn += cb_wn[w]; }
                    // ... followed by the lanes before this one (inclusive value of lane - 1)
                    const int f1 = __shfl_up_sync(kFull, f, 1), n1 = __shfl_up_sync(kFull, ni, 1);
                    const long long t1 = __shfl_up_sync(kFull, tl, 1);
                    long long carry_in = pt;
                    int irr_before = pn;
                    if (lane_ > 0) { carry_in = f1 ? t1 : pt + t1; irr_before = pn + n1; }
                    int M = 0;
                    for (int w = 0; w < nwarp_; w++) M += cb_wn[w];
                    seqsum::head_chunk_finish(v_p, lo, hi, pex, carry_in, irr_before, K_p, c_p, s_iidx, s_ie);
                    __syncthreads();
                    if (tid == 0) cb_head_ok = seqsum::head_stitch(v_p, K_p, s_iidx, s_ie, M, xs, s_base) ? 1 : 0;
                    __syncthreads();
                    if (cb_head_ok)  // (nothing reads the weights any more: the sums go in their place)
                        for (int i = lo; i < hi; i++) v_p[i] = seqsum::head_value(i, K_p, c_p, s_iidx, s_ie, s_base);
                    __syncthreads();
                }
                if (tid == 0) {
                    double acc = 0.;  // (0 + w == w: the reference starts with cum[0] = w[0])
                    if (cb_head_ok) {
                        cb_lap(34);
                        acc = xs > 0 ? v_p[xs - 1] : 0.;
                    } else {  // a stretch was refused: one addition after the other
                        for (int i = 0; i < xs; i++) { acc = acc + v_p[i]; v_p[i] = acc; }
                    }
                    int ns = 0;
                    seqsum::run_segments(acc, xs, n - xs, ts_W, s_seg, ns, seqsum::kMaxSegs);
                    cb_ns = ns <= seqsum::kMaxSegs ? ns : -1;
                    if (a.work) work_add(31, cb_head_ok ? 1 : 0);
                }
                __syncthreads();
                const int ns = cb_ns;
                for (int i = tid; i < xs; i += nthr) d.cum[i] = v_p[i];
                long long *seg_g = reinterpret_cast<long long *>(a.part_d + kPartCumSegs);
                for (int k = tid; k < 3 * max(ns, 0); k += nthr) seg_g[k] = reinterpret_cast<const long long *>(s_seg)[k];
                if (tid == 0) {
                    a.part_ll[kPartCumSegN] = ns;
                    if (a.work) { cb_lap(35); work_add(28, clock64() - t_cb); work_add(29, 1); work_add(30, max(ns, 0)); }
                }
                __syncthreads();
            };
            if (cum_builder && cum_dedicated) {
                cum_exact_build();
                if (tid == 0) {  // the plan this CTA goes on with must be the final one
                    const volatile int *vd = &a.ts_plan->done;
                    while (*vd == 0) __nanosleep(200);
                    __threadfence();
                }
                __syncthreads();
            }
            int *s_pos = reinterpret_cast<int *>(dyn_smem);  // [e][tid]; free in the routing CTAs (and in a lone CTA after plan_build)
            const int e_cap = min(64, max(kRoute, (a.smem_bytes / (int)sizeof(int) / nthr) & ~(kRoute - 1)));  // (64: the `live` mask)
            const int e_all = (int)((n + rsize - 1) / rsize);
            // all-equal right part [cut, l) entered at `p`: final position by the closed form, straight into the pick table
            auto finish_dense = [&](const tiesort::Level &L, int p, int lab) {
                bool bad = false;
                const int fp = tiesort::all_equal_final(L.cut, L.l, p, L.depth, bad);
                if (bad) { any_bad = true; return; }
                d.sorted_slot[fp] = d.slot_of_label[lab];
                a.sorted_label[fp] = lab;
            };
            int known_ready = 0, known_done = 0;
            for (int e0 = 0; router && e0 < e_all; e0 += e_cap) {
                const int E = min(e_cap, e_all - e0);
                unsigned long long live = 0;
                for (int e = 0; e < E; e++) {
                    const long long i = rtid + (long long)(e0 + e) * rsize;
                    s_pos[e * nthr + tid] = (int)i;
                    if (i < n && is_sparse[i] == 0) live |= 1ULL << e;
                }
                for (int t = 0;; t++) {
                    if (t >= known_ready && !known_done) {  // wait for level t (or for the end of the simulation)
                        __syncthreads();
                        if (tid == 0) {
                            const volatile int *vr = &a.ts_plan->ready, *vd = &a.ts_plan->done;
                            int r = *vr, dn = 0;
                            while (r <= t) {
                                dn = *vd;
                                if (dn) { r = *vr; break; }
                                __nanosleep(100);
                                r = *vr;
                            }
                            __threadfence();
                            sh_ready = r;
                            sh_done = dn;
                        }
                        __syncthreads();
                        known_ready = sh_ready;
                        known_done = sh_done;
                    }
                    if (t >= known_ready) break;  // (done, and every published level taken)
                    const tiesort::Level L = a.ts_plan->lv[t];
                    const int *__restrict__ Rt = R + (size_t)t * a.ts_xcap;
                    const int *__restrict__ Tt = T + (size_t)t * tiesort::kTblStride;
                    for (int g = 0; g < E; g += kRoute) {
                        const unsigned lv4 = (unsigned)(live >> g) & ((1u << kRoute) - 1);
                        if (!lv4) continue;
                        int pos[kRoute], lo_[kRoute], hi_[kRoute];
#pragma unroll
                        for (int j = 0; j < kRoute; j++) {  // pivot move + bucket bounds (dead lanes look at a valid dummy position)
                            int p = (lv4 >> j & 1) ? s_pos[(g + j) * nthr + tid] : L.f + 1;
                            if (p == L.f) p = L.pick;
                            else if (p == L.pick) p = L.f;
                            pos[j] = p;
                            const int bkt = (p - L.f) >> L.shift;
                            lo_[j] = Tt[bkt] & 0xffff;  // (packed table: position half)
                            hi_[j] = Tt[bkt + 1] & 0xffff;
                        }
#pragma unroll
                        for (int j = 0; j < kRoute; j++) {
                            if (!(lv4 >> j & 1)) continue;
                            int p = pos[j];
                            if (p > L.f) {
                                int r = lo_[j];
                                while (r < hi_[j] && Rt[r] < p) r++;
                                const int ka = p - (L.f + 1) - r;
                                if (ka < L.K) p = L.l - 1 - ka;
                                else {
                                    const int kb = L.l - 1 - p;
                                    if (kb < L.K) p = tiesort::select_dense_direct(Rt, Tt, L.f, L.shiftD, kb);
                                }
                            }
                            if (p >= L.cut) {  // into the all-W right part [cut, l)
                                finish_dense(L, p, (int)(rtid + (long long)(e0 + g + j) * rsize));
                                live &= ~(1ULL << (g + j));
                            } else s_pos[(g + j) * nthr + tid] = p;
                        }
                    }
                }
                if (e0 == 0) {  // the plan is final here
                    for (int k = tid; k < (int)(sizeof(tiesort::Plan) / sizeof(int)); k += nthr)
                        reinterpret_cast<int *>(&sh_plan)[k] = reinterpret_cast<const volatile int *>(a.ts_plan)[k];
                    __syncthreads();
                }
                if (!sh_plan.overlap && !sh_plan.fail) {  // (overlap: the simulating CTA has filled the handed-over segment)
                    const int hf = sh_plan.hand_f, hlen = sh_plan.hand_l - hf;
                    for (int e = 0; e < E; e++) {
                        if (!(live >> e & 1)) continue;
                        const int p = s_pos[e * nthr + tid] - hf;
                        b.perm[p] = (int)(rtid + (long long)(e0 + e) * rsize);
                        b.wk[p] = ts_W;
                        b.segf[p] = 0;
                        b.segl[p] = hlen;
                    }
                }
                __syncthreads();  // s_pos is reused by the next round
            }
            if (!router || e_all == 0) {  // the simulating CTA (and routers without elements): the final plan
                __syncthreads();
                for (int k = tid; k < (int)(sizeof(tiesort::Plan) / sizeof(int)); k += nthr)
                    reinterpret_cast<int *>(&sh_plan)[k] = reinterpret_cast<const volatile int *>(a.ts_plan)[k];
                __syncthreads();
            }
            lap(16);  // (laps 14-16 -> work[16..18]: the sparse simulation in three parts; the host adds them up for phase 8)
            if (sh_plan.fail) {  // introsort's heap-sort branch: the host redoes this sort on the multi-launch device path (k_sort_heap)
                if (gtid == 0) sc.b_need = 99;
                return;
            }
            const int hf = sh_plan.hand_f, hlen = sh_plan.hand_l - hf;
            ts_ovl = sh_plan.overlap != 0;
            if (ts_ovl && blk == 0) {
                unsigned char *s_mark = reinterpret_cast<unsigned char *>(b_s);  // list b is free after plan_build: xs_pad ints >= hlen bytes?
                const bool marks = hlen <= 4 * xs_pad;
                if (marks) {
                    for (int k = tid; k < hlen; k += nthr) s_mark[k] = 0;
                    __syncthreads();
                    for (int j = tid; j < xs; j += nthr) s_mark[a_s[j] - hf] = 1;
                    __syncthreads();
                }
                for (int p = hf + tid; p < hf + hlen; p += nthr) {
                    if (marks) {
                        if (s_mark[p - hf]) continue;
                    } else {
                        int lo = 0, hi = xs;  // a_s: ascending positions of the sparse elements at hand-over
                        while (lo < hi) { const int mid = (lo + hi) >> 1; if (a_s[mid] < p) lo = mid + 1; else hi = mid; }
                        if (lo < xs && a_s[lo] == p) continue;
                    }
                    b.perm[p - hf] = tiesort::dense_origin(sh_plan, arch_R, arch_T, xs, ts_nb + 3, p);
                    b.wk[p - hf] = ts_W;
                    b.segf[p - hf] = 0;
                    b.segl[p - hf] = hlen;
                }
                __syncthreads();
            }
            if (any_bad) b.active[2] = 1;
            n_sort = hlen;
            delta = hf;
            depth = sh_plan.hand_depth;
            if (gtid == 0 && a.work) { work_add(12, 1); work_add(13, sh_plan.n_levels); work_add(14, xs); work_add(15, hlen); }
            if (!ts_ovl) grid.sync();
            lap(9);
        }
    }
    if (!ts_on) {
        for (long long i = gtid; i < n; i += gsize) {
            b.perm[i] = (int)i;
            b.wk[i] = factor / d.a_ts[d.slot_of_label[i]];
            b.segf[i] = 0;
            b.segl[i] = n;
        }
        if (gtid == 0) { b.active[0] = 0; b.active[1] = 0; b.active[2] = 0; b.active[3] = 0; }
        grid.sync();
    }
    bool active = n_sort > kSortLeaf;
    auto pivot_of = [&](int f, int l) {  // __move_median_to_first(first, first+1, mid, last-1)
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
    };
    // b.active: [0],[1] ping-pong "some segment still > 16", [2] fail, [4..7] ping-pong (min, max) of the span of active segments,
    // [kWinBase + parity * 2 * kMaxWin + 2 * k, + 1]: (min, max) of the active segments that START in window k of the span (below)
    if (gtid == 0) {
        if (active) pivot_of(0, n_sort);
        b.active[4] = 0; b.active[5] = n_sort; b.active[6] = 0x7fffffff; b.active[7] = 0;
    }
    int level = 0;
    bool fail = false, local = false;
    // Work is restricted to the span [amin, amax) of the still-active segments (all-equal segments leave the loop
    // analytically, so in tie-dominated tables the span halves every level).  Once the span fits kSortLocal elements,
    // block 0 finishes the remaining levels alone with __syncthreads() instead of grid barriers.
    // Two thresholds: below `switch_span` elements block 0 works alone (block barriers instead of grid barriers; the state still in
    // HBM / L2: a span of ~10^4 elements gains nothing from more CTAs and each grid barrier costs ~3 us), and once the span also fits
    // the shared-memory staging area (smem_cap, sized by local_span) the remaining levels run out of shared memory.
    // Windows: the segments of a level are independent of each other, so as soon as every group of active segments that start in
    // the same window of the span (width >= kWinMin, at most one window per CTA) fits the shared-memory staging area, each CTA
    // takes its window's segments through ALL the remaining levels alone — block barriers and shared memory instead of four grid
    // barriers per level, and the windows run in parallel.  The new segment leaders record the windows of the next level.
    const int kSortSwitch = a.switch_span > a.local_span ? a.switch_span : a.local_span;
    const int n_win = min(nblk, kMaxWin);
    const int win_cap = min(a.smem_cap, a.win_cap);
    const bool win_allowed = win_cap > kSortLeaf + 2 && !ts_ovl && nblk > 1 && !a.no_windows;
    bool win_rec = false, windowed = false;  // win_rec: the previous level recorded the windows of this one
    long long t_win = 0;                     // diagnostics: work[19..21] += cycles / levels / 1 per window, work[22] += block 0's wait for the others
    int level_win = 0;
    long long t_wl = 0;                      // work[23..27] += window cycles in: level top + staging, flags + scan, (rest of scan), swaps, split
    auto wlap = [&](int k) { if (a.work && windowed && tid == 0) { const long long t = clock64(); work_add(23 + k, t - t_wl); t_wl = t; } };
    int eblk = blk, enblk = nblk;
    long long etid = gtid, esize = gsize;
    auto barrier = [&]() { if (local) __syncthreads(); else grid.sync(); };
    __shared__ int sh_act[8];
    int *act = b.active;
    bool staged = false;
    int st_min = 0, st_max = 0;
    SortBufs gb = b;
    if (!ts_ovl) grid.sync();  // first pivot + span in place
    else if (blk == 0) {
        if (exact_cum && tid == 0) {  // the staged sparse weights share the level passes' scratch: not before their last reader is done
            const volatile int *vg = &a.ts_plan->cum_gathered;
            while (*vg == 0) __nanosleep(100);
        }
        __syncthreads();
    } else active = false;       // overlap mode: the other CTAs are done with their part (routing) and wait behind the loop
    lap(2);
    while (active) {
        depth--;
        int amin = act[4 + 2 * (level & 1)], amax = act[5 + 2 * (level & 1)];
        if (!local && win_rec) {
            const int *win = b.active + kWinBase + (level & 1) * 2 * kMaxWin;
            int fits = 1;
            for (int k = tid; k < n_win; k += nthr) {
                const int lo = win[2 * k], hi = win[2 * k + 1];
                if (hi > lo && hi - lo + 2 > win_cap) fits = 0;
            }
            if (__syncthreads_and(fits)) {
                local = true;
                windowed = true;
                t_win = clock64();
                t_wl = t_win;
                level_win = level;
                const int lo = blk < n_win ? win[2 * blk] : 0x7fffffff, hi = blk < n_win ? win[2 * blk + 1] : 0;
                if (hi <= lo) break;  // no segment starts in this CTA's window: wait at the barrier behind the loop
                eblk = 0; enblk = 1; etid = tid; esize = nthr;
                if (tid < 8) sh_act[tid] = b.active[tid];
                __syncthreads();
                if (tid == 0) { sh_act[4 + 2 * (level & 1)] = lo; sh_act[5 + 2 * (level & 1)] = hi; }
                act = sh_act;
                __syncthreads();
                amin = lo;
                amax = hi;
            }
        }
        if (!local && amax - amin <= kSortSwitch) {
            local = true;
            if (blk != 0) break;  // the other blocks wait at the barrier behind the loop
            eblk = 0; enblk = 1; etid = tid; esize = nthr;
            if (tid < 8) sh_act[tid] = b.active[tid];
            act = sh_act;
            __syncthreads();
        }
        // (uniform over the grid) this level's new leaders record the windows of the next level
        const int win_w = max(kWinMin, (amax - amin + n_win - 1) / n_win);
        const bool win_now = !local && win_allowed && (long long)(amax - amin) <= (long long)n_win * win_cap;
        int *win_next = b.active + kWinBase + ((level + 1) & 1) * 2 * kMaxWin;
        if (local && !staged) {  // (block 0 only from here on)
            if (amax - amin + 2 <= a.smem_cap) {
                // The remaining levels run out of shared memory: the span's sort state is staged once (element i lives at
                // index i - amin; entry `amax` is a sentinel that reads as a finished segment), so each of the ~log2(span / 16)
                // block-local levels costs shared-memory latency instead of five dependent rounds through L2.
                const int cap = a.smem_cap;
                unsigned char *p = dyn_smem;
                double *s_wk = reinterpret_cast<double *>(p); p += sizeof(double) * cap;
                long long *s_flags = reinterpret_cast<long long *>(p); p += sizeof(long long) * cap;
                long long *s_pre = reinterpret_cast<long long *>(p); p += sizeof(long long) * cap;
                int *s_perm = reinterpret_cast<int *>(p); p += sizeof(int) * cap;
                int *s_segf = reinterpret_cast<int *>(p); p += sizeof(int) * cap;
                int *s_segl = reinterpret_cast<int *>(p); p += sizeof(int) * cap;
                int *s_ta = reinterpret_cast<int *>(p); p += sizeof(int) * cap;
                int *s_tb = reinterpret_cast<int *>(p); p += sizeof(int) * cap;
                int *s_cut = reinterpret_cast<int *>(p);
                for (int i = amin + tid; i < amax; i += nthr) {
                    s_perm[i - amin] = b.perm[i];
                    s_wk[i - amin] = b.wk[i];
                    s_segf[i - amin] = b.segf[i];
                    s_segl[i - amin] = b.segl[i];
                }
                if (tid == 0) { s_segf[amax - amin] = 0x7fffffff; s_segl[amax - amin] = 0; }
                staged = true;
                st_min = amin; st_max = amax;
                gb = b;
                b.perm = s_perm - amin; b.wk = s_wk - amin; b.segf = s_segf - amin; b.segl = s_segl - amin;
                b.flags = s_flags - amin; b.pre = s_pre - amin; b.tmp_a = s_ta - amin; b.tmp_b = s_tb - amin; b.cut = s_cut - amin;
            }
            __syncthreads();
        }
        wlap(0);
        const int span = amax - amin + 1;  // + 1 so that pre[amax] exists
        if (gtid == 0 && a.work) { work_add(0, span); work_add(1, 1); }
        const int chunk = ((span + enblk - 1) / enblk + nthr - 1) / nthr * nthr;
        // Staged window of at most kLean * blockDim entries: flags and their scan stay in registers — both counters packed in one
        // int (a window holds < 65536 entries), the warp scans of the kLean rounds independent of each other (in flight together),
        // one pass over the (round, warp) totals by warp 0.  Two CTA barriers instead of eight, no 64-bit shuffles.
        constexpr int kLean = 8;
        const bool lean = local && staged && span <= kLean * nthr;
        if (lean) {
            // (rolled loops on purpose: this kernel's phases are paced by instruction fetch — the time of a block of code follows the
            // size of what it executes, profiles/r2_tuning.md — so the per-round values are parked in the unused `flags` staging
            // array instead of an unrolled register array)
            __shared__ int sh_lean[kLean * (kEventThreads / 32)];
            const int lane = tid & 31, wrp = tid >> 5, nw = nthr >> 5;
            const int nr = (span + nthr - 1) / nthr;
#pragma unroll 1
            for (int r = 0; r < nr; r++) {
                const int i = amin + r * nthr + tid;
                int fl = 0;
                if (i < amax && i < n_sort) {
                    const int f = b.segf[i], l = b.segl[i];
                    if (l - f > kSortLeaf && i > f) {
                        const double kp = b.wk[f], kx = b.wk[i];
                        const int lp = b.perm[f], lx = b.perm[i];
                        if (!w_less(kx, lx, kp, lp, b.stable)) fl |= 1;
                        if (!w_less(kp, lp, kx, lx, b.stable)) fl |= 1 << 16;
                    }
                }
                const int inc = warp_inclusive_scan(fl, lane);
                if (i <= amax) b.flags[i] = (long long)(unsigned)(inc - fl) | ((long long)fl << 32);  // exclusive prefix inside the warp, own flags
                if (lane == 31) sh_lean[r * nw + wrp] = inc;
            }
            __syncthreads();
            if (tid < 32) {  // exclusive scan of the nr * nw <= 128 totals, four per lane
                const int m = nr * nw;
                int t4[4], sum = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) { t4[k] = 4 * lane + k < m ? sh_lean[4 * lane + k] : 0; sum += t4[k]; }
                int ex = warp_inclusive_scan(sum, lane) - sum;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (4 * lane + k < m) sh_lean[4 * lane + k] = ex;
                    ex += t4[k];
                }
            }
            __syncthreads();
#pragma unroll 1
            for (int r = 0; r < nr; r++) {
                const int i = amin + r * nthr + tid;
                if (i <= amax) {
                    const long long pk = b.flags[i];
                    const int fl = (int)(pk >> 32);
                    const int pre_i = sh_lean[r * nw + wrp] + (int)(pk & 0xffffffffLL);
                    const int pa_ = pre_i & 0xffff, pb_ = pre_i >> 16;
                    b.pre[i] = (long long)pa_ | ((long long)pb_ << 32);
                    if (fl & 1) b.tmp_a[amin + pa_] = i;
                    if (fl >> 16) b.tmp_b[amin + pb_] = i;
                }
            }
        } else {
        // ---- flags + chunk sums
        {
            const int lo = amin + eblk * chunk, hi = min(amax + 1, lo + chunk);
            long long acc = 0;
            for (int i = lo + tid; i < hi; i += nthr) {
                long long fl = 0;
                if (i < n_sort) {
                    const int f = b.segf[i], l = b.segl[i];
                    if (l - f > kSortLeaf && i > f) {
                        const double kp = b.wk[f], kx = b.wk[i];
                        const int lp = b.perm[f], lx = b.perm[i];
                        if (!w_less(kx, lx, kp, lp, b.stable)) fl |= 1LL;
                        if (!w_less(kp, lp, kx, lx, b.stable)) fl |= (1LL << 32);
                    }
                }
                b.flags[i] = fl;
                acc += fl;
            }
            if (enblk > 1) {  // chunk sums feed the other blocks' scan bases
                const long long t = block_sum_ll(acc, sm_ll);
                if (tid == 0) a.part_ll[eblk] = t;
            }
        }
        barrier();
        wlap(1);
        // ---- exclusive scan of the packed flags over the span
        {
            long long carry = 0;  // the same value in every thread
            if (eblk > 0) {
                for (int bb = tid; bb < eblk; bb += nthr) carry += a.part_ll[bb];
                carry = block_sum_ll(carry, sm_ll);
            }
            const int lo = amin + eblk * chunk, hi = min(amax + 1, lo + chunk);
            // 16 rounds of blockDim elements at a time: warp scans of every round first, ONE block scan over the (round, warp)
            // totals, then the warp scans again with their bases — 5 CTA barriers per 8192 elements instead of 5 per 512
            constexpr int kRounds = 16;
            __shared__ long long sh_wt[kRounds * (kEventThreads / 32) + 1];
            const int lane = tid & 31, wrp = tid >> 5, nw = nthr >> 5;
            for (int t0 = lo; t0 < hi; t0 += kRounds * nthr) {
                const int nr = min(kRounds, (hi - t0 + nthr - 1) / nthr);
                for (int r = 0; r < nr; r++) {
                    const int i = t0 + r * nthr + tid;
                    const long long inc = warp_inclusive_scan_ll((i < hi) ? b.flags[i] : 0, lane);
                    if (lane == 31) sh_wt[r * nw + wrp] = inc;
                }
                __syncthreads();
                {
                    const int m = nr * nw;
                    long long tot;
                    const long long pre = block_exclusive_scan_ll(tid < m ? sh_wt[tid] : 0, &tot, sm_ll);
                    if (tid < m) sh_wt[tid] = pre;
                    if (tid == 0) sh_wt[kRounds * (kEventThreads / 32)] = tot;
                    __syncthreads();
                }
                for (int r = 0; r < nr; r++) {
                    const int i = t0 + r * nthr + tid;
                    const long long v = (i < hi) ? b.flags[i] : 0;
                    const long long inc = warp_inclusive_scan_ll(v, lane);
                    if (i < hi) {
                        // the prefix over the whole span is also the slot of a stopper: the left stoppers of a segment [f, l) end up in
                        // tmp_a[amin + pre[f+1].a .. amin + pre[l].a) in position order, its right stoppers likewise in tmp_b — no
                        // per-segment reads, no separate scatter pass
                        const long long pre_i = carry + sh_wt[r * nw + wrp] + inc - v;
                        b.pre[i] = pre_i;
                        if (v & 1LL) b.tmp_a[amin + (int)(pre_i & 0xffffffffLL)] = i;
                        if (v >> 32) b.tmp_b[amin + (int)(pre_i >> 32)] = i;
                    }
                }
                carry += sh_wt[kRounds * (kEventThreads / 32)];
                __syncthreads();
            }
        }
        }  // (not lean)
        barrier();
        wlap(2);
        // ---- pairwise swaps + where the two scans stop.
        // A segment whose elements ALL equal the pivot (whole tie classes: every monomer of a monodisperse run) needs no
        // more memory passes: introsort's moves on equal keys do not depend on the data (median-of-3 picks `mid`, the
        // Hoare partition mirrors [f+1, l-1], the cut falls at f+1+(m-1)/2, leaves do not move), so every element
        // computes its final position in registers and leaves the level loop.
        for (long long j = amin + etid; j < amax; j += esize) {
            const int f = b.segf[j], l = b.segl[j];
            if (l - f <= kSortLeaf) continue;
            const int base = f + 1, k = (int)j - base;
            const long long p0 = b.pre[base], pl = b.pre[l];
            const int p0a = (int)(p0 & 0xffffffffLL), plb = (int)(pl >> 32);
            const int n_a = (int)(pl & 0xffffffffLL) - p0a, n_b = plb - (int)(p0 >> 32);
            if (n_a == l - f - 1 && n_b == l - f - 1) {
                int cf = f, cl = l, pos = (int)j, dep = depth;
                bool first = true, bad = false;
                while (cl - cf > kSortLeaf) {
                    const int m = cl - cf;
                    if (!first) {  // the pivot move of the current level has already been applied to the arrays
                        const int mid = cf + m / 2;
                        if (pos == cf) pos = mid; else if (pos == mid) pos = cf;
                    }
                    first = false;
                    if (pos > cf) pos = cf + cl - pos;
                    const int cutp = cf + 1 + (m - 1) / 2;
                    if (pos < cutp) cl = cutp; else cf = cutp;
                    if (cl - cf > kSortLeaf) {
                        if (dep == 0) { bad = true; break; }
                        dep--;
                    }
                }
                if (bad) { act[2] = 1; continue; }
                b.fin_perm[pos + delta] = b.perm[j];
                b.fin_wk[pos + delta] = b.wk[j];
                b.segf[j] = 0x7fffffff;  // done: inactive in every later phase
                b.segl[j] = 0;
                continue;
            }
            if (j <= f) continue;
            // k-th left stopper: tmp_a[A0 + k]; k-th right stopper counted from the right: tmp_b[B1 - k]
            const int *A = b.tmp_a + amin + p0a, *B = b.tmp_b + amin + plb - 1;
            const int m = n_a < n_b ? n_a : n_b;
            const bool sw = k < m && A[k] < B[-k];
            const bool next_sw = (k + 1 < m) && A[k + 1] < B[-(k + 1)];
            if (sw) {
                const int pa = A[k], pb = B[-k];
                const double ka = b.wk[pa];
                const int la = b.perm[pa];
                b.wk[pa] = b.wk[pb]; b.perm[pa] = b.perm[pb];
                b.wk[pb] = ka; b.perm[pb] = la;
            }
            int s_ = -1;
            if (sw && !next_sw) s_ = k + 1;
            else if (k == 0 && !sw) s_ = 0;
            if (s_ >= 0) {
                const int a_s = s_ < n_a ? A[s_] : 0x7fffffff;
                const int b_prev = s_ > 0 ? B[-(s_ - 1)] : l;
                b.cut[f] = a_s < b_prev ? a_s : b_prev;
            }
        }
        if (etid == 0) {
            act[(level + 1) & 1] = 0;
            act[4 + 2 * ((level + 1) & 1)] = 0x7fffffff;
            act[5 + 2 * ((level + 1) & 1)] = 0;
        }
        if (win_now && etid < 2 * n_win) win_next[etid] = (etid & 1) ? 0 : 0x7fffffff;
        barrier();
        wlap(3);
        // ---- split + (fused) median-of-3 pivots of the next level by the new leaders
        int lead_min = 0x7fffffff, lead_max = 0;  // span of the new segments this thread leads
        // (the swap pass may already have hit the depth limit inside an all-equal segment: some of its elements are finished, the others
        // are not, and the segment has no cut — nothing below may touch it; the loop ends on `fail` right after this pass)
        const bool failed_in_swap = act[2] != 0;
        for (long long i = amin + etid; i < amax && !failed_in_swap; i += esize) {
            const int f = b.segf[i], l = b.segl[i];
            if (l - f <= kSortLeaf) continue;
            const int c = b.cut[f];
            int nf = f, nl = l;
            if (i < c) nl = c; else nf = c;
            b.segf[i] = nf;
            b.segl[i] = nl;
            if (i == nf && nl - nf > kSortLeaf) {
                if (depth > 0) {
                    lead_min = nf < lead_min ? nf : lead_min;
                    lead_max = nl > lead_max ? nl : lead_max;
                    pivot_of(nf, nl);
                    if (win_now) {
                        const int k = (nf - amin) / win_w;
                        atomicMin(&win_next[2 * k], nf);
                        atomicMax(&win_next[2 * k + 1], nl);
                    }
                } else act[2] = 1;  // would enter introsort's heap-sort branch
            }
        }
        // one atomic pair per warp, not per new segment (hundreds of same-address atomics per level in the deep levels)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int t0 = __shfl_xor_sync(kFull, lead_min, o), t1 = __shfl_xor_sync(kFull, lead_max, o);
            lead_min = t0 < lead_min ? t0 : lead_min;
            lead_max = t1 > lead_max ? t1 : lead_max;
        }
        if ((tid & 31) == 0 && lead_max > 0) {
            act[level & 1] = 1;
            atomicMin(&act[4 + 2 * ((level + 1) & 1)], lead_min);
            atomicMax(&act[5 + 2 * ((level + 1) & 1)], lead_max);
        }
        barrier();
        wlap(4);
        active = act[level & 1] != 0;
        fail = act[2] != 0;
        if (fail) break;
        lap(local ? 4 : 3);
        win_rec = win_now;
        level++;
    }
    // cumulative_time_steps, first half: sums of fixed chunks of the sorted weights (the general sort's output inside
    // [delta, delta + n_sort), W everywhere else); see the second half below
    auto fw = [&](int i) { return (i >= delta && i < delta + n_sort) ? b.fin_wk[i] : ts_W; };
    constexpr int kCumRounds = 16, kCumWarps = kEventThreads / 32;
    const int chunk_c = kCumRounds * nthr, n_chunks = (n + chunk_c - 1) / chunk_c;
    double *chunk_sum = a.part_d + kPartCumChunks;  // behind the per-block partials: room for 8192 chunks (6.7e7 entries)
    auto cum_chunk_sums = [&]() {
        if (n <= a.cum_sequential_max) return;
        for (int c = blk; c < n_chunks; c += nblk) {
            const int lo = c * chunk_c, hi = min(n, lo + chunk_c);
            double acc = 0.;
            for (int i = lo + tid; i < hi; i += nthr) acc += fw(i);
            const double t = block_sum_fixed(acc, sm_d);
            if (tid == 0) chunk_sum[c] = t;
        }
    };
    // __final_insertion_sort of one leaf [i, segl[i]) by its first element's thread
    auto sort_leaf = [&](int i) {
        const int f = i, l = b.segl[i];
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
        for (int x = f; x < l; x++) { b.fin_perm[x + delta] = b.perm[x]; b.fin_wk[x + delta] = b.wk[x]; }
    };
    // overlap mode: the handed-over segment belongs to block 0 alone, which also sorts its leaves (out of shared memory when the
    // levels were staged): the grid-wide leaf pass and its barrier are skipped
    const bool leaves_by_block0 = ts_ovl && n_sort > kSortLeaf;
    if (local && (blk == 0 || windowed) && act == sh_act) {  // (a CTA without a window of its own has nothing staged)
        if (leaves_by_block0) {
            __syncthreads();
            if (!fail)
                for (int i = tid; i < n_sort; i += nthr)
                    if (b.segf[i] == i) sort_leaf(i);
            __syncthreads();
            if (staged) b = gb;
        } else if (staged) {  // back to HBM for the leaf pass
            __syncthreads();
            for (int i = st_min + tid; i < st_max; i += nthr) {
                gb.perm[i] = b.perm[i];
                gb.wk[i] = b.wk[i];
                gb.segf[i] = b.segf[i];
                gb.segl[i] = b.segl[i];
            }
            b = gb;
        }
        if (tid == 0 && sh_act[2]) b.active[2] = 1;
    }
    // overlap mode: the chunk sums do not wait for the barrier — every chunk but the first is W only, and the first one (it holds the
    // handed-over segment: hand_l <= 4096 < chunk) belongs to block 0, which has just finished it
    const bool cum_early = !exact_cum && leaves_by_block0 && delta + n_sort <= chunk_c;
    if (cum_early) cum_chunk_sums();
    if (a.work && tid == 0 && windowed && act == sh_act) { work_add(19, clock64() - t_win); work_add(20, level - level_win); work_add(21, 1); }
    const long long t_wait = clock64();
    grid.sync();  // blocks that left the loop early wait here for block 0's local levels
    if (a.work && gtid == 0) work_add(22, clock64() - t_wait);
    lap(4);
    fail = b.active[2] != 0 || a.force_fail != 0;
    if (fail) {  // the host redoes this sort on the multi-launch device path, which replays the heap-sort branch (k_sort_heap)
        if (gtid == 0) sc.b_need = 99;
        return;
    }
    // ---- __final_insertion_sort per leaf
    if (!leaves_by_block0) {
        for (long long i = gtid; i < n_sort; i += gsize)
            if (b.segf[i] == i) sort_leaf((int)i);
        grid.sync();
    }
    lap(5);
    // ---- cumulative_time_steps, second half
    if (n <= a.cum_sequential_max) {
        if (gtid == 0) {
            double acc = fw(0);
            d.cum[0] = acc;
            for (int i = 1; i < n; i++) { acc = acc + fw(i); d.cum[i] = acc; }
        }
    } else if (exact_cum && (int)__ldcg(a.part_ll + kPartCumSegN) >= 0) {
        // the sparse head is in place (written by the last CTA, which also left the segments of the W run): every entry of the run in
        // closed form, the reference's sequential sum bit for bit (seq_cumsum.cuh)
        const int ns = (int)__ldcg(a.part_ll + kPartCumSegN);
        long long *s_seg = reinterpret_cast<long long *>(dyn_smem);
        const long long *seg_g = reinterpret_cast<const long long *>(a.part_d + kPartCumSegs);
        __syncthreads();
        for (int k = tid; k < 3 * ns; k += nthr) s_seg[k] = __ldcg(seg_g + k);
        __syncthreads();
        const seqsum::Seg *segs = reinterpret_cast<const seqsum::Seg *>(dyn_smem);
        for (long long i = cum_xs + gtid; i < n; i += gsize) {
            const int sg = seqsum::find_segment(segs, ns, (int)i);
            d.cum[i] = seqsum::segment_value(segs[sg], (int)i);
        }
    } else {
        // Fixed tree, independent of the launch shape: chunks of kCumRounds * blockDim entries (CTA b takes chunks b, b + gridDim, ...),
        // chunk sums combined in chunk order, inside a chunk one round of blockDim entries after the other —
        // cum[i] = (carry + (wtot[0] + .. + wtot[w-1])) + inc, carry' = (carry + wbase of the last warp) + its total.
        if (!cum_early) {
            cum_chunk_sums();
            grid.sync();
        }
        __shared__ double carry_d;
        __shared__ double s_wtot[kCumRounds][kCumWarps], s_wbase[kCumRounds][kCumWarps], s_carry[kCumRounds];
        const int lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
        auto warp_inc = [&](double v) {
            double inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double tt = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += tt;
            }
            return inc;
        };
        // tie-dominated table: behind the handed-over segment every entry is W, and the warp scan of 32 equal values is the same
        // vector in every such warp — computed once, not 2 x 16 times per chunk (the scans are shuffle-throughput-bound)
        const double inc_w = warp_inc(ts_W);
        const int dense_from = ts_on ? delta + n_sort : n;
        auto all_w = [&](int i) { const int w0 = i - lane; return w0 >= dense_from && w0 + 31 < n; };  // (warp-uniform)
        double base_run = 0.;  // (thread 0) sum of the chunks before the next one of this CTA, in chunk order
        int base_upto = 0;
        for (int c = blk; c < n_chunks; c += nblk) {
            const int lo = c * chunk_c, hi = min(n, lo + chunk_c);
            // the chunk sums [base_upto, c) through shared memory (at most gridDim of them per step), added by one thread in order
            const int cnt = c - base_upto;
            for (int k0 = 0; k0 < cnt; k0 += kMaxCoopBlocks) {
                const int m = min(kMaxCoopBlocks, cnt - k0);
                __syncthreads();
                for (int k = tid; k < m; k += nthr) sh_pd[0][k] = chunk_sum[base_upto + k0 + k];
                __syncthreads();
                if (tid == 0)
                    for (int k = 0; k < m; k++) base_run += sh_pd[0][k];
            }
            base_upto = c;
            if (tid == 0) carry_d = base_run;
            __syncthreads();
            const int nr = (hi - lo + nthr - 1) / nthr;
            for (int r = 0; r < nr; r++) {
                const int i = lo + r * nthr + tid;
                const double inc = all_w(i) ? inc_w : warp_inc((i < hi) ? fw(i) : 0.);
                if (lane == 31) s_wtot[r][w] = inc;
            }
            __syncthreads();
            if (tid < nr * nw) {
                const int r = tid / nw, ww_ = tid % nw;
                double wbase = 0.;
                for (int ww = 0; ww < ww_; ww++) wbase += s_wtot[r][ww];
                s_wbase[r][ww_] = wbase;
            }
            __syncthreads();
            if (tid == 0) {
                double cc = carry_d;
                for (int r = 0; r < nr; r++) {
                    s_carry[r] = cc;
                    cc = cc + s_wbase[r][nw - 1] + s_wtot[r][nw - 1];
                }
            }
            __syncthreads();
            for (int r = 0; r < nr; r++) {
                const int i = lo + r * nthr + tid;
                const double inc = all_w(i) ? inc_w : warp_inc((i < hi) ? fw(i) : 0.);
                if (i < hi) d.cum[i] = s_carry[r] + s_wbase[r][w] + inc;
            }
            __syncthreads();
        }
    }
    grid.sync();
    lap(6);
    for (long long i = delta + gtid; i < delta + n_sort; i += gsize) {  // (the routing pass wrote the rest of a tie-dominated table)
        const int lab = b.fin_perm[i];
        d.sorted_slot[i] = d.slot_of_label[lab];
        a.sorted_label[i] = lab;
    }
    if (gtid == 0) {
        sc.n_pick = n;
        sc.cum_total = d.cum[n - 1];
        sc.pick_dense_from = ts_on ? delta + n_sort : n;  // behind the handed-over segment every entry is a W element
        sc.pick_w = ts_W;
        sc.last_sort_tie = ts_on ? 1 : 0;
    }
    lap(7);
}

// Tuning probe (mcac_gpu_kernel_bench 11): the sparse simulation of tie_sort.cuh alone, one CTA with k_event's shared-memory layout,
// on a synthetic tie-dominated table (x sparse elements among n).  cycles[0] += SM cycles of plan_build, cycles[1] += its levels.
__global__ void __launch_bounds__(kEventThreads) k_plan_probe(int n, int xs, const int *st_pos, const double *st_w, double W, int depth0, int hand_min,
                                                              tiesort::Plan *plan, int *R, int *tbl, int xcap, int smem_bytes, long long *cycles) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    int ts_nb = 256;
    while (ts_nb < xs && ts_nb < tiesort::kBuckets) ts_nb <<= 1;
    const int xs_pad = (xs + 2 + 3) & ~3;
    const int used_ints = 4 * xs_pad + (ts_nb + 4) + tiesort::kMiscInts;
    int *sm = reinterpret_cast<int *>(dyn_smem);
    int *a_s = sm, *a_i = a_s + xs_pad, *b_s = a_i + xs_pad, *b_i = b_s + xs_pad, *s_tbl = b_i + xs_pad, *s_misc = s_tbl + ts_nb + 4,
        *arch_R = s_misc + tiesort::kMiscInts;
    const int arch_levels = min(tiesort::kMaxLevels, (smem_bytes / (int)sizeof(int) - used_ints) / (xs + ts_nb + 3 + 1));
    int *arch_T = arch_R + (size_t)arch_levels * xs;
    ProbeTeam tm{tid, nthr, cycles + 4, 0};
    __syncthreads();
    const long long t0 = clock64();
    tm.prev = t0;
    tiesort::plan_build(tm, n, xs, st_pos, st_w, W, depth0, hand_min, plan, R, tbl, xcap, a_s, a_i, b_s, b_i, s_tbl, s_misc, arch_R, arch_T, arch_levels);
    if (tid == 0) {
        atomicAdd(reinterpret_cast<unsigned long long *>(cycles), (unsigned long long)(clock64() - t0));
        atomicAdd(reinterpret_cast<unsigned long long *>(cycles + 1), (unsigned long long)plan->n_levels);
        atomicAdd(reinterpret_cast<unsigned long long *>(cycles + 2), (unsigned long long)(plan->hand_l - plan->hand_f));
    }
}

// ------------------------------------------------------------------------------------------------
// K11 — per-realization morphology statistics staged for the ensemble all-gather: histogram of log2(Np),
// histogram of Rg on [0, rg_max), and the log-log regression sums of AggregatList::get_instantaneous_fractal_law
// (aggregat_list_fractal_law.cpp:23-33 -> linreg, tools.cpp:126-157: x = dg_over_dp, y = Np).
// out: [0,nb) Np histogram, [nb,2nb) Rg histogram, then n, sum_np, sumx, sumx2, sumxy, sumy, sumy2, sum_rg
// ------------------------------------------------------------------------------------------------
// Deterministic: every CTA reduces its (fixed) grid-stride share in a fixed tree and leaves 8 partial sums; the CTA that
// finishes last (ticket counter) combines the partials in CTA order.  Histograms are integer counts (atomics on integers are
// exact and order-independent).  No floating-point atomics: two launches on the same state give bit-identical rows.
__global__ void __launch_bounds__(256) k_morphology_stats(DevState d, int nb, double rg_max, double *out, double *part /* gridDim x 8 */,
                                                          unsigned int *ghist /* 2 * nb counts + 1 ticket, zeroed by the caller */) {
    __shared__ double red[8][8];
    __shared__ unsigned int hist[2 * 64];  // per-CTA histograms (nb <= 64 bins each): one global atomic per bin and CTA, not per aggregate
    __shared__ int s_last;
    const int n = d.sc->n_agg_slots;
    const bool local_hist = nb <= 64;
    for (int b = threadIdx.x; b < 2 * 64; b += blockDim.x) hist[b] = 0u;
    __syncthreads();
    double acc[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        if (!d.a_alive[s]) continue;
        const double np_ = static_cast<double>(d.a_n[s]);
        const double rg = d.a_rg[s];
        int b1 = 31 - __clz(d.a_n[s]);
        b1 = b1 < nb ? b1 : nb - 1;
        int b2 = static_cast<int>(rg / rg_max * nb);
        b2 = b2 < 0 ? 0 : (b2 < nb ? b2 : nb - 1);
        if (local_hist) {
            atomicAdd(&hist[b1], 1u);
            atomicAdd(&hist[64 + b2], 1u);
        } else {
            atomicAdd(&ghist[b1], 1u);
            atomicAdd(&ghist[nb + b2], 1u);
        }
        const double lx = log(d.a_dgdp[s]), ly = log(np_);
        acc[0] += 1.; acc[1] += np_; acc[2] += lx; acc[3] += lx * lx; acc[4] += lx * ly; acc[5] += ly; acc[6] += ly * ly; acc[7] += rg;
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(kFull, acc[k], o);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 8; k++) red[w][k] = acc[k];
    __syncthreads();
    if (threadIdx.x < 8) {
        double t = 0.;
        for (int ww = 0; ww < 8; ww++) t += red[ww][threadIdx.x];
        part[blockIdx.x * 8 + threadIdx.x] = t;
    }
    if (local_hist)
        for (int b = threadIdx.x; b < 2 * nb; b += blockDim.x) {
            const unsigned int c = hist[b < nb ? b : 64 + (b - nb)];
            if (c) atomicAdd(&ghist[b], c);
        }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ghist[2 * nb], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < 8) {  // partial sums combined in CTA order, whichever CTA happens to be the last one
        double t = 0.;
        for (unsigned int bb = 0; bb < gridDim.x; bb++) t += __ldcg(&part[bb * 8 + threadIdx.x]);
        out[2 * nb + threadIdx.x] = t;
    }
    for (int b = threadIdx.x; b < 2 * nb; b += blockDim.x) out[b] = static_cast<double>(__ldcg(&ghist[b]));
}
// FP64 pipe microbenchmarks (the denominators of the FP64 roofline, measured, not quoted): every thread runs 8
// independent dependency chains of either DFMA (mode 0: what the pipe can do) or DMUL + DADD (mode 1: what it can do for this
// library, which is built --fmad=false to match the reference's rounding).  flops = 2 per chain step in both modes.
template <int kMode>
__global__ void __launch_bounds__(256) k_fp64_peak(double *sink, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = seed + 1e-3 * (threadIdx.x + 8 * k);
    const double m = 1.0 + 1e-9, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (kMode == 0) a[k] = __fma_rn(a[k], m, c);
            else a[k] = __dadd_rn(__dmul_rn(a[k], m), c);
        }
    }
    double t = 0.;
#pragma unroll
    for (int k = 0; k < 8; k++) t += a[k];
    if (t == 123.456) sink[0] = t;  // keeps the chains alive
}
// summary of a sweep of independent searches (no commit): contacts, checksum of the finite distances, pair counters
__global__ void __launch_bounds__(256) k_sweep_summary(const SearchResult *res, const double *q_dist, int nq, double *out /* 4 */) {
    double cnt = 0., sum = 0., ps = 0., pb = 0.;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        const SearchResult r = res[q];
        if (r.distance <= q_dist[q]) { cnt += 1.; sum += r.distance; }
        ps += (double)r.n_sphere_pairs;
        pb += (double)r.n_bounding;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(kFull, cnt, o);
        sum += __shfl_xor_sync(kFull, sum, o);
        ps += __shfl_xor_sync(kFull, ps, o);
        pb += __shfl_xor_sync(kFull, pb, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], cnt);
        atomicAdd(&out[1], sum);
        atomicAdd(&out[2], ps);
        atomicAdd(&out[3], pb);
    }
}

// ------------------------------------------------------------------------------------------------
// State boundary (mcac_gpu_upload_state / mcac_gpu_download_state): the host keeps the reference's layout
// (field-major SoA in creation / label order, include/constants.hpp:34-69; ordered `myspheres` as CSR).  The raw
// arrays cross PCIe once, as they are; the re-layout into the aggregate-major pool (and back) is done here, in HBM.
// ------------------------------------------------------------------------------------------------
struct HostLayout {  // device staging copy of the host arrays (all int64 / double, host layout)
    long long n_sph, n_agg;
    double *sph;                 // 9 x n_sph
    long long *sph_label;        // n_sph (download only)
    long long *sph_charge;       // n_sph
    double *agg;                 // 21 x n_agg
    long long *agg_n;            // n_agg (download only)
    long long *agg_charge;       // n_agg
    long long *agg_cells;        // 3 x n_agg
    long long *offsets;          // n_agg + 1
    long long *members;          // n_sph
    double *per_member;          // 3 x n_sph
};
__global__ void __launch_bounds__(256) k_upload_spheres(DevState d, HostLayout s) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= s.n_sph) return;
    const long long id = s.members[k], n = s.n_sph;
    d.s_posr[k] = make_double4(s.sph[id], s.sph[n + id], s.sph[2 * n + id], s.sph[3 * n + id]);
    d.s_relv[k] = make_double4(s.sph[6 * n + id], s.sph[7 * n + id], s.sph[8 * n + id], s.sph[4 * n + id]);
    d.s_surf[k] = s.sph[5 * n + id];
    d.s_veff[k] = s.per_member[k];
    d.s_seff[k] = s.per_member[n + k];
    d.s_dcen[k] = s.per_member[2 * n + k];
    d.s_id[k] = (int)id;
    d.s_charge[k] = (int)s.sph_charge[id];
    d.slot_of_id[id] = (int)k;
}
__global__ void __launch_bounds__(256) k_upload_aggregates(DevState d, HostLayout s, double density) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= s.n_agg) return;
    const long long m = s.n_agg;
    const double *A = s.agg;
    d.a_posr[a] = make_double4(A[7 * m + a], A[8 * m + a], A[9 * m + a], A[4 * m + a]);
    d.a_rg[a] = A[a]; d.a_fagg[a] = A[m + a]; d.a_lpm[a] = A[2 * m + a]; d.a_ts[a] = A[3 * m + a];
    d.a_vol[a] = A[5 * m + a]; d.a_surf[a] = A[6 * m + a];
    d.a_rx[a] = A[10 * m + a]; d.a_ry[a] = A[11 * m + a]; d.a_rz[a] = A[12 * m + a]; d.a_ptime[a] = A[13 * m + a];
    d.a_dp[a] = A[14 * m + a]; d.a_dgdp[a] = A[15 * m + a]; d.a_ovl[a] = A[16 * m + a]; d.a_cn[a] = A[17 * m + a];
    d.a_dm[a] = A[19 * m + a]; d.a_ch[a] = A[20 * m + a];
    const long long o = s.offsets[a], n = s.offsets[a + 1] - o;
    d.a_bulk[a] = density;
    d.a_alpha[a] = 1.0 / static_cast<double>(n);
    d.a_n[a] = (int)n;
    d.a_off[a] = (int)o;
    d.a_cx[a] = (int)s.agg_cells[a]; d.a_cy[a] = (int)s.agg_cells[m + a]; d.a_cz[a] = (int)s.agg_cells[2 * m + a];
    d.a_charge[a] = (int)s.agg_charge[a];
    d.a_alive[a] = 1;
    d.a_dirty[a] = kDirtyAll;
    d.label_of_slot[a] = (int)a;
    d.slot_of_label[a] = (int)a;
}
__global__ void __launch_bounds__(256) k_download_counts(DevState d, int n_agg, int *cnt) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < n_agg) cnt[l] = d.a_n[d.slot_of_label[l]];
}
// one warp per label: aggregate row by lane 0, its spheres (a contiguous block of the pool) by all lanes
__global__ void __launch_bounds__(256) k_download_state(DevState d, HostLayout s, const int *off_of_label) {
    const long long l = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (l >= s.n_agg) return;
    const long long m = s.n_agg, ns = s.n_sph;
    const int slot = d.slot_of_label[l];
    const int n = d.a_n[slot], o = d.a_off[slot];
    const long long off = off_of_label[l];
    if (lane == 0) {
        double *A = s.agg;
        const double4 p = d.a_posr[slot];
        A[l] = d.a_rg[slot]; A[m + l] = d.a_fagg[slot]; A[2 * m + l] = d.a_lpm[slot]; A[3 * m + l] = d.a_ts[slot];
        A[4 * m + l] = p.w; A[5 * m + l] = d.a_vol[slot]; A[6 * m + l] = d.a_surf[slot];
        A[7 * m + l] = p.x; A[8 * m + l] = p.y; A[9 * m + l] = p.z;
        A[10 * m + l] = d.a_rx[slot]; A[11 * m + l] = d.a_ry[slot]; A[12 * m + l] = d.a_rz[slot]; A[13 * m + l] = d.a_ptime[slot];
        A[14 * m + l] = d.a_dp[slot]; A[15 * m + l] = d.a_dgdp[slot]; A[16 * m + l] = d.a_ovl[slot]; A[17 * m + l] = d.a_cn[slot];
        A[18 * m + l] = 0.; A[19 * m + l] = d.a_dm[slot]; A[20 * m + l] = d.a_ch[slot];
        s.agg_n[l] = n;
        s.agg_charge[l] = d.a_charge[slot];
        s.agg_cells[l] = d.a_cx[slot]; s.agg_cells[m + l] = d.a_cy[slot]; s.agg_cells[2 * m + l] = d.a_cz[slot];
        s.offsets[l] = off;
        if (l == m - 1) s.offsets[m] = off + n;
    }
    if (off + n > ns) return;  // inconsistent membership: reported by the host from the scan total
    for (int k = lane; k < n; k += 32) {
        const int t = o + k;
        const long long id = d.s_id[t];
        const double4 p = d.s_posr[t], r = d.s_relv[t];
        s.sph[id] = p.x; s.sph[ns + id] = p.y; s.sph[2 * ns + id] = p.z; s.sph[3 * ns + id] = p.w;
        s.sph[4 * ns + id] = r.w; s.sph[5 * ns + id] = d.s_surf[t];
        s.sph[6 * ns + id] = r.x; s.sph[7 * ns + id] = r.y; s.sph[8 * ns + id] = r.z;
        s.sph_label[id] = l;
        s.sph_charge[id] = d.s_charge[t];
        s.members[off + k] = id;
        s.per_member[off + k] = d.s_veff[t];
        s.per_member[ns + off + k] = d.s_seff[t];
        s.per_member[2 * ns + off + k] = d.s_dcen[t];
    }
}

// cost of the grid-wide barrier of the cooperative event kernel at its launch shape (diagnostic for K9's roofline)
__global__ void __launch_bounds__(kEventThreads) k_barrier_probe(int n, int *sink) {
    cgx::grid_group grid = cgx::this_grid();
    for (int i = 0; i < n; i++) grid.sync();
    if (sink && blockIdx.x == 0 && threadIdx.x == 0) *sink = n;
}

}  // namespace mcacb
