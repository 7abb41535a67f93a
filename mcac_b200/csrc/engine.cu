// mcac_b200 — engine: device memory of one realization, launch orchestration, and the C ABI (include/mcac_b200.h).
//
// Host code here only sequences kernels and moves state across the PCIe boundary; all per-step arithmetic of the
// hot path runs in the kernels of mcac_kernels.cuh.  There is no CPU fallback: every entry point fails with
// UNKNOWN_ERROR when no CUDA device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <type_traits>
#include <mutex>
#include <vector>

#include "../../include/mcac_b200.h"
#include "mcac_kernels.cuh"
#include "mcac_steploop.cuh"

using namespace mcacb;

namespace {
constexpr int kRngBuf = 31 * 33826;  // ~1.05 M draws, multiple of 31
constexpr int kStatsMaxBlocks = 1024, kStatsHistCap = 4096;
enum { E_OK = 0, E_UNKNOWN = 1, E_IO = 2, E_VERLET = 3, E_INPUT = 4, E_MERGE = 9 };

struct Buf {  // raw device allocation
    void *p = nullptr;
    size_t bytes = 0;
};

// Host-side image of a realization in the reference's own layout (label / creation order); used at the
// upload / download boundary and by the (rare, host-orchestrated) domain duplication.
struct HostState {
    long long n_sph = 0, n_agg = 0;
    std::vector<double> sph;        // 9 x n_sph field-major
    std::vector<long long> sph_label, sph_charge;
    std::vector<double> agg;        // 21 x n_agg field-major
    std::vector<long long> agg_n, agg_charge, agg_cells, offsets, members;
    std::vector<double> per_member;  // 3 x n_sph
    double scalars[20] = {0};
};
}  // namespace

struct mcac_gpu {
    mcac_params prm{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // side stream: the Verlet cell rebuild runs beside the event kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int *block_sums2 = nullptr;
    bool cells_on_side = false;
    int force_sort_fail = 0;    // MCAC_B200_FORCE_SORT_FAIL=k: every k-th device sort reports failure (exercises the fallback)
    long long sort_calls = 0;
    bool stop_at_event = false;   // mcac_gpu_run returns after the step that made an event (merge / nucleation)
    bool debug_sync = false;
    long long nucl_headroom = 0;  // MCAC_B200_NUCL_HEADROOM: fixed (small) slot headroom, to exercise the regrow path in tests
    bool overlap = true;        // MCAC_B200_NO_OVERLAP=1 serialises the rebuild and synchronises after every event kernel  // a rebuild is in flight on stream2 (joined before the next search)
    std::string err;
    DevState d{};
    DevState alt{};  // alternate sphere buffers for pool compaction (only s_* pointers used)
    std::vector<void *> owned;       // size-dependent allocations (re-made by upload / duplication)
    std::vector<void *> persistent;  // RNG stream, scalars, batch scratch: live as long as the handle
    Scalars *h_sc = nullptr;  // pinned mirror
    Scalars *h_sc_ring[2] = {nullptr, nullptr};  // pinned read-back slots of the batches in flight (pipelined submission)
    cudaEvent_t ev_cycle[2] = {nullptr, nullptr};
    bool pipeline = true;       // MCAC_B200_NO_PIPELINE=1: read every batch back before the next one is submitted
    Scalars sc_host{};
    // scratch
    int *q_slot = nullptr;
    double *q_dir = nullptr, *q_dist = nullptr;
    long long *q_label = nullptr;
    SearchResult *q_res = nullptr;
    int *scan_tmp = nullptr, *scan_out = nullptr, *block_sums = nullptr, *sorted_label = nullptr, *merged_flag = nullptr;
    double *partials = nullptr, *stats_dev = nullptr, *stats_part = nullptr;  // K11: output row, per-CTA partial sums
    unsigned int *stats_hist = nullptr;                                       // K11: integer histogram counts + ticket
    mcac_step_record *rec_dev = nullptr;
    long long rec_cap = 0;
    long long rng_generated = 0;  // stream position after the last generated draw
    long long dup_threshold = 0;
    bool labels_valid = false, pick_valid = false, cells_valid = false, uploaded = false;
    long long launches = 0;
    int n_sm = 148;
    int profile = 0;
    std::vector<cudaEvent_t> ev_pool;  // start/stop pairs recorded around K1 / commit launches when profile != 0
    std::vector<int> ev_kind;
    SortBufs sortb{};
    long long *scan64_sums = nullptr;
    double *cum_sums = nullptr;
    int *h_flags = nullptr;  // pinned: sort `active` flags
    long long sort_levels = 0, sort_fallbacks = 0, sort_heap_levels = 0;
    long long sort_fallbacks_seen = 0, sort_heap_seen = 0;
    int sort_depth_override = -1;  // MCAC_B200_SORT_DEPTH (test hook): introsort depth limit, to reach the heap-sort branch
    int coop_blocks = 0;      // grid of the cooperative event kernel (0 = not available / disabled)
    int coop_bps = 1, sort_local_span = 4096;
    int sort_switch_span = 0;      // MCAC_B200_SORT_SWITCH: span below which block 0 of the event kernel sorts alone, still out of L2
                                   // (0 = local_span: measured slower at 16384, profiles/r2_tuning.md)
    int event_spare_sms = 24; // MCAC_B200_EVENT_SPARE_SMS (sweep in profiles/r1_tuning.md)
    int event_smem_cap = 0;   // shared-memory staging of the block-local sort levels (entries; 0 = levels stay in HBM/L2)
    size_t event_dyn_bytes = 0;  // dynamic shared memory of the event kernel's launches
    long long *event_work = nullptr;
    long long *commit_prof = nullptr;  // MCAC_B200_K9_DEBUG: phase clocks of k_commit
    double loop_cost_per_step = 0.;    // SM cycles per MC step of this realization's last step-loop launch (ensemble queue order)
    long long event_work_seen[40] = {0};
    // tie-dominated pick tables (tie_sort.cuh): plan + per-level rank tables; allocated with the state when the table can be large
    tiesort::Plan *ts_plan = nullptr;
    int *ts_R = nullptr, *ts_tbl = nullptr;
    int ts_xcap = 0;
    int ts_min_n = 32768;     // MCAC_B200_TIE_MIN_N (0 disables the fast path)
    int ts_max_sparse = tiesort::kMaxSparse;  // MCAC_B200_TIE_MAX_SPARSE
    bool ts_no_overlap = false;               // MCAC_B200_TIE_NO_OVERLAP
    int exact_cum_max_sparse = 8192;          // MCAC_B200_EXACT_CUM_MAX_SPARSE: most lighter aggregates the exact cumulative table is built for
    bool no_exact_cum = false;                // MCAC_B200_NO_EXACT_CUM: big tie-dominated tables summed by the fixed tree (tuning: the cost of the exact table)
    long long *part_ll = nullptr;
    double *part_d = nullptr;
    int cum_sequential_max = 65536;  // below this size cumulative_time_steps is summed sequentially (the reference's rounding)
    int *sweep_slot = nullptr;
    double *sweep_dir = nullptr, *sweep_dist = nullptr;
    SearchResult *sweep_res = nullptr;
    long long sweep_cap = 0;
    BigSearch *big = nullptr;   // scratch of the three-kernel search between many-sphere aggregates
    double big_search_npp = 24.;  // mean spheres per aggregate from which single searches take that form (MCAC_B200_BIG_NPP)
    int *wide_list = nullptr, *wide_count = nullptr;  // queries handed from the group search kernel to the wide one
    long long wide_cap = 0;
    int wide_parity = 0;
    int search_group = -1;    // lanes per query of K1 (4, 8, 16, 32; 0 = wide kernel only; -1 = by launch size)
    int search_min_blocks = 8;  // occupancy target (__launch_bounds__ min blocks) of the narrow-group kernels
    // per-realization step loop (mcac_steploop.cuh): the general step of calcul() as one persistent CTA
    bool fused = true;            // MCAC_B200_NO_LOOP=1: every general step goes through the multi-launch sequence
    int fused_max_slots = 16384;  // MCAC_B200_LOOP_MAX_SLOTS: larger aggregate tables leave the loop to the multi-launch path
    LoopState *loop_dev = nullptr, *loop_host = nullptr;
    bool strict_dir = false;      // MCAC_B200_STRICT_DIRECTION / mcac_gpu_set_strict_direction: directions evaluated by the host's glibc
    bool dir_tab_valid = false;
    double *dir_tab = nullptr;
    bool loop_dups = true;        // MCAC_B200_NO_LOOP_DUP=1: the loop hands every domain duplication to the host
    long long reserve_sph = 0, reserve_agg = 0;  // mcac_gpu_reserve / MCAC_B200_RESERVE_SPHERES, _AGGREGATES: table sizes to allocate at least
    bool loop_prune = true;       // MCAC_B200_NO_PRUNE=1: the ordered sweep tests every sphere pair of an examined suspect
    long long loop_launches = 0, loop_steps = 0;
    long long loop_cycles[8] = {0}, loop_cycles_seen[8] = {0};
    void *stage = nullptr;  // device staging of the host-layout arrays at the upload / download boundary
    size_t stage_bytes = 0;
};

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
            return E_UNKNOWN;                                                                     \
        }                                                                                         \
    } while (0)
#define TRY(call)              \
    do {                       \
        int r_ = (call);       \
        if (r_ != E_OK) return r_; \
    } while (0)

namespace {
inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
// MCAC_B200_DEBUG_SYNC=1: synchronise after every launch group of the step loops and name the group that faulted
int dbg_sync(mcac_gpu *h, const char *tag);
#define DBG(tag)                                                     \
    if (h->debug_sync && (rc = dbg_sync(h, tag)) != E_OK) break


template <class T>
int dev_alloc(mcac_gpu *h, T **p, size_t n) {
    void *q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    h->owned.push_back(q);
    *p = (T *)q;
    return E_OK;
}
template <class T>
int dev_alloc_persistent(mcac_gpu *h, T **p, size_t n) {
    void *q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    h->persistent.push_back(q);
    *p = (T *)q;
    return E_OK;
}
void free_all(mcac_gpu *h) {
    for (void *p : h->owned) cudaFree(p);
    h->owned.clear();
}
int alloc_persistent(mcac_gpu *h) {
    DevState &d = h->d;
    TRY(dev_alloc_persistent(h, &d.rng, 1));
    TRY(dev_alloc_persistent(h, &d.rng_buf, kRngBuf + 64));
    TRY(dev_alloc_persistent(h, &d.sc, 1));
    TRY(dev_alloc_persistent(h, &h->q_slot, kMaxBatch));
    TRY(dev_alloc_persistent(h, &h->q_dir, 3 * kMaxBatch));
    TRY(dev_alloc_persistent(h, &h->q_dist, kMaxBatch));
    TRY(dev_alloc_persistent(h, &h->q_label, kMaxBatch));
    TRY(dev_alloc_persistent(h, &h->q_res, kMaxBatch));
    TRY(dev_alloc_persistent(h, &h->partials, 3 * 1024));
    TRY(dev_alloc_persistent(h, &h->merged_flag, 4));
    TRY(dev_alloc_persistent(h, &h->stats_dev, 4096));
    TRY(dev_alloc_persistent(h, &h->loop_dev, 1));
    TRY(dev_alloc_persistent(h, &h->stats_part, 8 * (size_t)kStatsMaxBlocks));
    TRY(dev_alloc_persistent(h, &h->stats_hist, (size_t)kStatsHistCap));
    h->alt.sc = d.sc;
    return E_OK;
}

int alloc_state(mcac_gpu *h, long long agg_cap, long long sph_cap) {
    DevState &d = h->d;
    d.agg_cap = (int)agg_cap;
    d.sph_cap = (int)sph_cap;
    for (DevState *t : {&h->d, &h->alt}) {
        TRY(dev_alloc(h, &t->s_posr, sph_cap));
        TRY(dev_alloc(h, &t->s_relv, sph_cap));
        TRY(dev_alloc(h, &t->s_surf, sph_cap));
        TRY(dev_alloc(h, &t->s_veff, sph_cap));
        TRY(dev_alloc(h, &t->s_seff, sph_cap));
        TRY(dev_alloc(h, &t->s_dcen, sph_cap));
        TRY(dev_alloc(h, &t->s_id, sph_cap));
        TRY(dev_alloc(h, &t->s_charge, sph_cap));
    }
    TRY(dev_alloc(h, &d.slot_of_id, sph_cap));
    TRY(dev_alloc(h, &d.a_posr, agg_cap));
    for (double **p : {&d.a_rg, &d.a_fagg, &d.a_lpm, &d.a_ts, &d.a_vol, &d.a_surf, &d.a_rx, &d.a_ry, &d.a_rz, &d.a_ptime, &d.a_dp,
                       &d.a_dgdp, &d.a_ovl, &d.a_cn, &d.a_dm, &d.a_ch, &d.a_bulk, &d.a_alpha, &d.cum, &d.keys})
        TRY(dev_alloc(h, p, agg_cap));
    for (int **p : {&d.a_n, &d.a_off, &d.a_cx, &d.a_cy, &d.a_cz, &d.a_charge, &d.a_alive, &d.a_dirty, &d.label_of_slot, &d.slot_of_label,
                    &d.sorted_slot, &d.cell_items, &h->sorted_label})
        TRY(dev_alloc(h, p, agg_cap));
    TRY(dev_alloc(h, &d.cell_posr, agg_cap));
    TRY(dev_alloc(h, &d.cell_start, (size_t)d.n_cells + 1));
    TRY(dev_alloc(h, &d.cell_fill, (size_t)d.n_cells + 1));
    SortBufs &sb = h->sortb;
    TRY(dev_alloc(h, &sb.perm, agg_cap));
    TRY(dev_alloc(h, &sb.wk, agg_cap));
    TRY(dev_alloc(h, &sb.segf, agg_cap));
    TRY(dev_alloc(h, &sb.segl, agg_cap));
    TRY(dev_alloc(h, &sb.flags, agg_cap + 1));
    TRY(dev_alloc(h, &sb.pre, agg_cap + 2));
    TRY(dev_alloc(h, &sb.tmp_a, agg_cap + 1));
    TRY(dev_alloc(h, &sb.tmp_b, agg_cap + 1));
    TRY(dev_alloc(h, &sb.cut, agg_cap + 1));
    TRY(dev_alloc(h, &sb.fin_perm, agg_cap + 1));
    TRY(dev_alloc(h, &sb.fin_wk, agg_cap + 1));
    TRY(dev_alloc(h, &sb.active, kWinBase + 4 * kMaxWin));
    h->ts_plan = nullptr; h->ts_R = nullptr; h->ts_tbl = nullptr; h->ts_xcap = 0;
    if (h->ts_min_n > 0 && agg_cap >= h->ts_min_n) {
        h->ts_xcap = std::max(1, std::min(h->ts_max_sparse, tiesort::kMaxSparse));
        TRY(dev_alloc(h, &h->ts_plan, 1));
        TRY(dev_alloc(h, &h->ts_R, (size_t)tiesort::kMaxLevels * h->ts_xcap + 8));  // (+8: rank look-ups read two entries past a row)
        TRY(dev_alloc(h, &h->ts_tbl, (size_t)tiesort::kMaxLevels * tiesort::kTblStride));
    }
    TRY(dev_alloc(h, &h->scan64_sums, agg_cap / (kScanBlock * kScanItems) + 8));
    TRY(dev_alloc(h, &h->cum_sums, agg_cap / (kScanBlock * kScanItems) + 8));
    const size_t scan_n = (size_t)std::max<long long>(agg_cap, d.n_cells) + 2;
    TRY(dev_alloc(h, &h->scan_tmp, scan_n));
    TRY(dev_alloc(h, &h->scan_out, scan_n));
    TRY(dev_alloc(h, &h->block_sums, scan_n / (kScanBlock * kScanItems) + 8));
    TRY(dev_alloc(h, &h->block_sums2, scan_n / (kScanBlock * kScanItems) + 8));
    return E_OK;
}

// deterministic exclusive scan of n ints (in -> out[0..n], out[n] = total)
int scan_ints(mcac_gpu *h, const int *in, int n, int *out, cudaStream_t st = nullptr, int *sums = nullptr) {
    if (!st) st = h->stream;
    if (!sums) sums = h->block_sums;
    const int per = kScanBlock * kScanItems;
    const int nb = std::max(1, div_up(n, per));
    k_scan_partials<<<nb, kScanBlock, 0, st>>>(in, n, sums);
    k_scan_block_sums<<<1, 1024, 0, st>>>(sums, nb);
    k_scan_apply<<<nb, kScanBlock, 0, st>>>(in, n, sums, nb, out);
    h->launches += 3;
    CK(cudaGetLastError());
    return E_OK;
}

int dbg_sync(mcac_gpu *h, const char *tag) {
    const cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) return E_OK;
    h->err = std::string("after ") + tag + ": " + cudaGetErrorString(e) + " (realization step " + std::to_string(h->sc_host.steps_done) +
             ", n_agg " + std::to_string(h->sc_host.n_agg) + ", n_sph " + std::to_string(h->sc_host.n_sph) + ")";
    return E_UNKNOWN;
}
// device-side error -> message + ErrorCodes value (Scalars::error is already one of the reference's ErrorCodes)
int device_error(mcac_gpu *h, const Scalars &s, const char *where) {
    const char *what = "";
    switch (s.error_detail) {
    case DETAIL_SUSPECT_OVERFLOW: what = ": a contact search found more eligible suspects than its list holds (the first contact is not known)"; break;
    case DETAIL_NOT_ON_VERLET: what = ": Aggregate not on the verlet list ???"; break;
    case DETAIL_RNG_NOT_STAGED: what = ": random draws of the step were not staged"; break;
    case DETAIL_PICK_TABLE: what = ": corrupt pick table"; break;
    case DETAIL_SPHERE_REMOVAL: what = ": a sphere shrank below rp_min_oxid (sphere removal / split are not built)"; break;
    case DETAIL_POOL_FULL: what = ": sphere pool / aggregate table full"; break;
    default: break;
    }
    static const char *names[] = {"NO_ERROR", "UNKNOWN_ERROR", "IO_ERROR", "VERLET_ERROR", "INPUT_ERROR", "ABANDON_ERROR", "TOO_DENSE_ERROR",
                                  "SBL_ERROR", "VOL_SURF_ERROR", "MERGE_ERROR", "ARVO_ERROR", "INTERPOTENTIAL_ERROR"};
    const int code = (s.error >= 1 && s.error <= 11) ? s.error : E_UNKNOWN;
    h->err = std::string(where) + ": device-side " + names[code] + what + " (step " + std::to_string(s.steps_done) + ")";
    return code;
}
int pull_scalars(mcac_gpu *h) {
    CK(cudaMemcpyAsync(h->h_sc, h->d.sc, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->sc_host = *h->h_sc;
    return E_OK;
}
int push_scalars(mcac_gpu *h) {
    *h->h_sc = h->sc_host;
    CK(cudaMemcpyAsync(h->d.sc, h->h_sc, sizeof(Scalars), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return E_OK;
}

int refresh_labels(mcac_gpu *h) {
    if (h->labels_valid) return E_OK;
    const int n = h->sc_host.n_agg_slots;
    TRY(scan_ints(h, h->d.a_alive, n, h->scan_out));
    k_alive_to_labels<<<div_up(n, 256), 256, 0, h->stream>>>(h->d, h->scan_out);
    h->launches++;
    CK(cudaGetLastError());
    h->labels_valid = true;
    return E_OK;
}

// K2: counting sort of live aggregates into their stored Verlet cells
int build_cells(mcac_gpu *h, bool side = false) {
    if (h->cells_valid) return E_OK;
    DevState &d = h->d;
    const int n = h->sc_host.n_agg_slots;
    cudaStream_t st = side ? h->stream2 : h->stream;
    // side == true: the caller guarantees that the state the rebuild reads (positions, cells, liveness) is final, i.e. that the
    // main stream was synchronised after the last kernel that wrote it
    CK(cudaMemsetAsync(d.cell_fill, 0, sizeof(int) * ((size_t)d.n_cells + 1), st));
    k_cell_count<<<div_up(n, 256), 256, 0, st>>>(d);
    h->launches++;
    TRY(scan_ints(h, d.cell_fill, d.n_cells, d.cell_start, st, side ? h->block_sums2 : h->block_sums));
    CK(cudaMemsetAsync(d.cell_fill, 0, sizeof(int) * ((size_t)d.n_cells + 1), st));
    k_cell_scatter<<<div_up(n, 256), 256, 0, st>>>(d);
    h->launches++;
    CK(cudaGetLastError());
    if (side) {
        CK(cudaEventRecord(h->ev_join, h->stream2));
        h->cells_on_side = true;
    }
    h->cells_valid = true;
    return E_OK;
}
// the main stream waits for a rebuild that was forked to the side stream
int join_cells(mcac_gpu *h) {
    if (!h->cells_on_side) return E_OK;
    CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    h->cells_on_side = false;
    return E_OK;
}

int refresh_reduce(mcac_gpu *h) {
    const int nb = std::min(1024, std::max(1, div_up(h->sc_host.n_agg_slots, kReduceThreads)));
    k_refresh_partials<<<nb, kReduceThreads, 0, h->stream>>>(h->d, h->partials);
    k_refresh_final<<<1, 32, 0, h->stream>>>(h->d, h->partials, nb);
    h->launches += 2;
    CK(cudaGetLastError());
    return E_OK;
}

// AggregatList::sort_time_steps (aggregat_list.cpp:124-141), multi-launch form.  The 1/dt weights are computed on the device in
// label order; MCAC_ORDER_LIBSTDCXX reproduces the reference's std::sort (introsort) order among EQUAL weights, which decides the
// pick in monodisperse runs — partition levels, the heap-sort branch behind introsort's depth limit and the final
// insertion sort are all replayed on the device (there is no host sort anywhere in the product).
// Device form: weights (K9 keys) -> replayed introsort (see mcac_kernels.cuh) -> cumulative table -> slots.
int sort_time_steps(mcac_gpu *h, double factor) {
    DevState &d = h->d;
    TRY(refresh_labels(h));
    const int n = h->sc_host.n_agg;
    const int nb = div_up(n, 256);
    k_make_keys<<<nb, 256, 0, h->stream>>>(d, factor);
    h->launches++;
    SortBufs sb = h->sortb;
    sb.n = n;
    sb.stable = h->prm.sort_order == MCAC_ORDER_STABLE ? 1 : 0;
    k_sort_init<<<nb, 256, 0, h->stream>>>(sb, d.keys);
    h->launches++;
    int lg = 0;
    while ((1LL << (lg + 1)) <= n) lg++;
    int depth = h->sort_depth_override >= 0 ? h->sort_depth_override : 2 * lg;  // std::__lg(n) * 2
    bool active = n > kSortLeaf, fail = false;
    if (active && depth == 0) { fail = true; active = false; }  // (test hook) the very first segment already takes the heap-sort branch
    const int per = kScanBlock * kScanItems, snb = std::max(1, div_up(n + 1, per));
    while (active && !fail) {
        depth--;
        k_sort_pivot<<<nb, 256, 0, h->stream>>>(sb);
        k_sort_flags<<<nb, 256, 0, h->stream>>>(sb);
        CK(cudaMemsetAsync(sb.flags + n, 0, sizeof(long long), h->stream));
        k_scan64_partials<<<snb, kScanBlock, 0, h->stream>>>(sb.flags, n + 1, h->scan64_sums);
        k_scan64_block_sums<<<1, 1024, 0, h->stream>>>(h->scan64_sums, snb);
        k_scan64_apply<<<snb, kScanBlock, 0, h->stream>>>(sb.flags, n + 1, h->scan64_sums, snb, sb.pre);
        k_sort_scatter<<<nb, 256, 0, h->stream>>>(sb);
        k_sort_swap<<<nb, 256, 0, h->stream>>>(sb);
        CK(cudaMemsetAsync(sb.active, 0, 2 * sizeof(int), h->stream));
        k_sort_split<<<nb, 256, 0, h->stream>>>(sb, depth);
        h->launches += 8;
        CK(cudaMemcpyAsync(h->h_flags, sb.active, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        active = h->h_flags[0] != 0;
        fail = h->h_flags[1] != 0;
        h->sort_levels++;
    }
    CK(cudaGetLastError());
    if (fail) {  // introsort's depth limit was hit: the segments still longer than 16 take the heap-sort branch
        k_sort_heap<<<nb, 256, 0, h->stream>>>(sb);
        h->launches++;
        h->sort_heap_levels++;
    }
    k_sort_leaves<<<nb, 256, 0, h->stream>>>(sb);
    h->launches++;
    if (n <= h->cum_sequential_max) {
        k_cum_sequential<<<1, 32, 0, h->stream>>>(sb.wk, d.cum, n);
        h->launches++;
    } else {
        const int cnb = std::max(1, div_up(n, per));
        k_cum_partials<<<cnb, kScanBlock, 0, h->stream>>>(sb.wk, n, h->cum_sums);
        k_cum_block_sums<<<1, 32, 0, h->stream>>>(h->cum_sums, cnb);
        k_cum_apply<<<cnb, kScanBlock, 0, h->stream>>>(sb.wk, n, h->cum_sums, d.cum);
        h->launches += 3;
    }
    CK(cudaMemcpyAsync(h->sorted_label, sb.perm, sizeof(int) * n, cudaMemcpyDeviceToDevice, h->stream));
    k_sort_finish<<<nb, 256, 0, h->stream>>>(d, sb);
    h->launches++;
    CK(cudaGetLastError());
    TRY(pull_scalars(h));
    h->pick_valid = true;
    return E_OK;
}

// The per-event pipeline as ONE cooperative launch (k_event): labels, refresh / PhysicalModel::update, weights, replayed
// introsort, cumulative table.  The multi-launch device form takes over when cooperative launch is unavailable or when
// introsort's depth limit is hit (it replays the heap-sort branch).
int event_pipeline(mcac_gpu *h, bool do_refresh, bool do_totals, bool do_sort, const double *factor = nullptr, bool defer_sync = false,
                   bool skip_if_no_event = false) {
    if (h->coop_blocks <= 0) {
        if (do_refresh || do_totals) { h->labels_valid = false; TRY(refresh_labels(h)); TRY(refresh_reduce(h)); TRY(pull_scalars(h)); }
        if (do_sort) TRY(sort_time_steps(h, factor ? *factor : h->sc_host.max_time_step));
        return E_OK;
    }
    EventArgs a{};
    a.use_factor = factor ? 1 : 0;
    a.factor = factor ? *factor : 0.;
    a.sb = h->sortb;
    a.part_ll = h->part_ll;
    a.part_d = h->part_d;
    a.scan_tmp = h->scan_tmp;
    a.sorted_label = h->sorted_label;
    a.do_labels = h->labels_valid ? 0 : 1;
    a.do_refresh = do_refresh ? 1 : 0;
    a.do_totals = do_totals ? 1 : 0;
    a.do_sort = do_sort ? 1 : 0;
    a.cum_sequential_max = h->cum_sequential_max;
    a.stable = h->prm.sort_order == MCAC_ORDER_STABLE ? 1 : 0;
    a.local_span = h->sort_local_span;
    a.switch_span = std::max(h->sort_switch_span, h->sort_local_span);
    a.work = h->event_work;
    a.smem_cap = h->event_smem_cap;
    a.no_windows = getenv("MCAC_B200_NO_SORT_WINDOWS") ? 1 : 0;
    a.win_cap = 2048;  // (span + 1 <= 4 * blockDim: the register form of the flag scan) (profiles/r2_tuning.md: 4096 -> 94 us, 2048 -> 84 us, 1024 -> 88 us of sort levels per event at N = 1e6)
    if (const char *e = getenv("MCAC_B200_SORT_WINDOW")) a.win_cap = std::max(64, atoi(e)) + 2;
    a.smem_bytes = (int)h->event_dyn_bytes;
    a.ts_plan = h->ts_plan;
    a.ts_R = h->ts_R;
    a.ts_tbl = h->ts_tbl;
    a.ts_xcap = h->ts_xcap;
    a.ts_min_n = h->ts_min_n;
    a.skip_if_no_event = skip_if_no_event ? 1 : 0;
    a.ts_no_overlap = h->ts_no_overlap ? 1 : 0;
    a.no_exact_cum = h->no_exact_cum ? 1 : 0;
    a.exact_cum_max_sparse = h->exact_cum_max_sparse;
    a.depth_override = h->sort_depth_override;
    a.force_fail = (do_sort && h->force_sort_fail > 0 && (++h->sort_calls % h->force_sort_fail) == 0) ? 1 : 0;
    DevState dcopy = h->d;
    void *args[] = {&dcopy, &a};
    const void *fn = h->coop_bps == 2 ? (const void *)k_event<2> : (const void *)k_event<1>;
    // grid: one CTA per 2048 aggregate slots, at most one CTA per SM — a small realization (ensembles, the early boxes of C1) gets a
    // single CTA whose barriers are __syncthreads-cheap and which leaves the other SMs to the other realizations' streams
    const int want_blocks = std::max(1, div_up(h->sc_host.n_agg_slots + (h->prm.with_nucleation ? 4096 : 0), 2048));
    // a few SMs are left to the side stream (the overlapped Verlet cell rebuild cannot share an SM with a 512-thread, 120-register CTA):
    // 8 when the event kernel takes > 400 us (general sort replay: the rebuild hides behind it anyway), event_spare_sms when the last pick
    // table came from the sparse path (~200 us: the rebuild on 8 SMs would outlast it; sweep in profiles/r1_tuning.md)
    const int spare = !h->overlap ? 0 : (h->sc_host.last_sort_tie ? h->event_spare_sms : std::min(8, h->event_spare_sms));
    const int grid_blocks = std::min(std::max(1, h->coop_blocks - spare), want_blocks);
    CK(cudaLaunchCooperativeKernel(fn, dim3(grid_blocks), dim3(kEventThreads), args, h->event_dyn_bytes, h->stream));
    h->launches++;
    h->labels_valid = true;
    if (defer_sync) {  // the caller reads the scalars after its next kernels and handles a failed sort (b_need == 99) there
        if (do_sort) h->pick_valid = true;
        return E_OK;
    }
    TRY(pull_scalars(h));
    if (do_sort) {
        if (h->sc_host.b_need == 99) {  // introsort's depth limit was hit: multi-launch path, which replays the heap-sort branch
            h->sc_host.b_need = 0;
            TRY(push_scalars(h));
            h->sort_fallbacks++;
            TRY(sort_time_steps(h, factor ? *factor : h->sc_host.max_time_step));
        }
        h->pick_valid = true;
    }
    return E_OK;
}

// strict replay mode: random_direction() (src/tools/tools.cpp:82-89) of every pair of consecutive staged draws, evaluated HERE with the
// host's glibc sin / cos / acos — the very functions the reference calls — and handed to the device as a table
int build_direction_table(mcac_gpu *h) {
    DevState &d = h->d;
    const int n = d.rng_buf_n;
    if (!h->dir_tab) TRY(dev_alloc_persistent(h, &h->dir_tab, 3 * (size_t)(kRngBuf + 64)));
    std::vector<int> draws((size_t)std::max(n, 1));
    std::vector<double> tab(3 * (size_t)std::max(n, 1), 0.);
    if (n > 0) {
        CK(cudaMemcpyAsync(draws.data(), d.rng_buf, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int p = 0; p + 1 < n; p++) {
            const Vec3 v = direction_from_draws(uniform_from_rand(draws[(size_t)p]), uniform_from_rand(draws[(size_t)p + 1]));
            tab[3 * (size_t)p] = v.x; tab[3 * (size_t)p + 1] = v.y; tab[3 * (size_t)p + 2] = v.z;
        }
        CK(cudaMemcpyAsync(h->dir_tab, tab.data(), sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    d.dir_tab = h->dir_tab;
    h->dir_tab_valid = true;
    return E_OK;
}
int ensure_rng(mcac_gpu *h, long long need_until) {  // draws [rand_pos, need_until) must be in rng_buf
    DevState &d = h->d;
    const long long pos = h->sc_host.rand_pos;
    if (need_until <= h->rng_generated && pos >= d.rng_buf_base) {
        if (h->strict_dir && !h->dir_tab_valid) TRY(build_direction_table(h));
        return E_OK;
    }
    const long long keep = std::max<long long>(0, h->rng_generated - pos);  // generated but unconsumed
    if (keep > 0 && pos > d.rng_buf_base) {
        std::vector<int> tmp((size_t)keep);
        CK(cudaMemcpyAsync(tmp.data(), d.rng_buf + (pos - d.rng_buf_base), sizeof(int) * keep, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpyAsync(d.rng_buf, tmp.data(), sizeof(int) * keep, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    d.rng_buf_base = pos;
    long long room = kRngBuf - keep;
    room -= room % 31;
    k_rng_fill<<<1, 32, 0, h->stream>>>(d.rng, d.rng_buf + keep, (int)room);
    h->launches++;
    CK(cudaGetLastError());
    h->rng_generated = pos + keep + room;
    d.rng_buf_n = (int)(keep + room);
    h->dir_tab_valid = false;
    if (h->strict_dir) TRY(build_direction_table(h));
    return E_OK;
}

int compact_pool(mcac_gpu *h) {
    DevState &d = h->d;
    const int n = h->sc_host.n_agg_slots;
    k_compact_counts<<<div_up(n, 256), 256, 0, h->stream>>>(d, h->scan_tmp);
    TRY(scan_ints(h, h->scan_tmp, n, h->scan_out));
    k_compact_move<<<div_up(n, 8), 256, 0, h->stream>>>(d, h->alt, h->scan_out);
    k_compact_finish<<<1, 32, 0, h->stream>>>(d, h->scan_out);
    h->launches += 3;
    CK(cudaGetLastError());
    std::swap(d.s_posr, h->alt.s_posr);
    std::swap(d.s_relv, h->alt.s_relv);
    std::swap(d.s_surf, h->alt.s_surf);
    std::swap(d.s_veff, h->alt.s_veff);
    std::swap(d.s_seff, h->alt.s_seff);
    std::swap(d.s_dcen, h->alt.s_dcen);
    std::swap(d.s_id, h->alt.s_id);
    std::swap(d.s_charge, h->alt.s_charge);
    TRY(pull_scalars(h));
    return E_OK;
}

// raw host arrays in the reference's layout (what the C ABI receives / fills)
struct HostView {
    long long n_sph, n_agg;
    const double *sph;
    const long long *sph_charge;
    const double *agg;
    const long long *agg_charge, *agg_cells, *offsets, *members;
    const double *per_member;
};
struct HostOut {
    double *sph;
    long long *sph_label, *sph_charge;
    double *agg;
    long long *agg_n, *agg_charge, *agg_cells, *offsets, *members;
    double *per_member, *scalars;
};
int upload(mcac_gpu *h, const HostView &s, double maxradius, double max_time_step);
bool is_speculative(const mcac_gpu *h);
int upload(mcac_gpu *h, const HostState &s, double maxradius, double max_time_step, bool keep_scalars);
int download(mcac_gpu *h, HostState &s);
int download_to(mcac_gpu *h, const HostOut &o);

// AggregatList::duplication (aggregat_list.cpp:142-190): box x2, 7 translated copies of every aggregate appended in
// (i,j,k) order, Verlet rebuilt.  Rare (once per 8x drop of N_agg) and purely a re-layout, so it is orchestrated
// through the download/upload boundary; the copies' positions are produced by the same K3 translate kernel.
int duplicate_on_device(mcac_gpu *h, bool *done);
int duplicate(mcac_gpu *h) {
    {   // in place on the device when the tables have room for 8x the state (k_duplicate, mcac_steploop.cuh)
        bool done = false;
        TRY(duplicate_on_device(h, &done));
        if (done) return E_OK;
    }
    HostState s;
    TRY(download(h, s));
    const long long n0 = s.n_agg, m0 = s.n_sph;
    const double old_l = s.scalars[1];
    HostState t;
    t.n_agg = 8 * n0;
    t.n_sph = 8 * m0;
    t.sph.assign((size_t)(9 * t.n_sph), 0.);
    t.sph_label.assign((size_t)t.n_sph, 0);
    t.sph_charge.assign((size_t)t.n_sph, 0);
    t.agg.assign((size_t)(21 * t.n_agg), 0.);
    t.agg_n.assign((size_t)t.n_agg, 0);
    t.agg_charge.assign((size_t)t.n_agg, 0);
    t.agg_cells.assign((size_t)(3 * t.n_agg), 0);
    t.offsets.assign((size_t)t.n_agg + 1, 0);
    t.members.assign((size_t)t.n_sph, 0);
    t.per_member.assign((size_t)(3 * t.n_sph), 0.);
    // originals keep label / sphere index; copy c of aggregate a gets label n0 + 7a + c and fresh sphere ids in member order
    std::vector<long long> src_of((size_t)t.n_agg);
    std::vector<int> shift((size_t)t.n_agg, 0);
    for (long long a = 0; a < n0; a++) src_of[(size_t)a] = a;
    {
        long long l = n0;
        for (long long a = 0; a < n0; a++)
            for (int c = 1; c < 8; c++) { src_of[(size_t)l] = a; shift[(size_t)l] = c; l++; }
    }
    for (long long i = 0; i < m0; i++) {
        for (int f = 0; f < 9; f++) t.sph[(size_t)(f * t.n_sph + i)] = s.sph[(size_t)(f * m0 + i)];
        t.sph_charge[(size_t)i] = s.sph_charge[(size_t)i];
    }
    long long next_id = m0, off = 0;
    for (long long l = 0; l < t.n_agg; l++) {
        const long long a = src_of[(size_t)l];
        for (int f = 0; f < 21; f++) t.agg[(size_t)(f * t.n_agg + l)] = s.agg[(size_t)(f * n0 + a)];
        t.agg_n[(size_t)l] = s.agg_n[(size_t)a];
        t.agg_charge[(size_t)l] = (l < n0) ? s.agg_charge[(size_t)a] : 0;  // copy ctor resets the charge (aggregat_storage.cpp:138)
        t.offsets[(size_t)l] = off;
        for (long long k = 0; k < s.agg_n[(size_t)a]; k++) {
            const long long sid = s.members[(size_t)(s.offsets[(size_t)a] + k)];
            long long nid = sid;
            if (l >= n0) {
                nid = next_id++;
                for (int f = 0; f < 9; f++) t.sph[(size_t)(f * t.n_sph + nid)] = s.sph[(size_t)(f * m0 + sid)];
                t.sph_charge[(size_t)nid] = 0;
            }
            t.members[(size_t)(off + k)] = nid;
            for (int f = 0; f < 3; f++) t.per_member[(size_t)(f * t.n_sph + off + k)] = s.per_member[(size_t)(f * m0 + s.offsets[(size_t)a] + k)];
        }
        off += s.agg_n[(size_t)a];
    }
    t.offsets[(size_t)t.n_agg] = off;
    std::memcpy(t.scalars, s.scalars, sizeof(t.scalars));
    // new box (aggregat_list.cpp:145-148)
    h->prm.box_length = old_l * 2;
    h->prm.n_monomeres *= 8;
    h->prm.box_volume = std::pow(h->prm.box_length, 3);
    t.scalars[1] = h->prm.box_length;
    t.scalars[15] = h->prm.box_volume;
    t.scalars[17] = (double)h->prm.n_monomeres;
    Scalars keep = h->sc_host;
    h->uploaded = false;
    TRY(upload(h, t, t.scalars[2], t.scalars[3], false));
    // restore run-time scalars that upload() resets
    h->sc_host.time = keep.time;
    h->sc_host.n_iter_without_event = keep.n_iter_without_event;
    h->sc_host.total_events = keep.total_events;
    h->sc_host.steps_done = keep.steps_done;
    h->sc_host.rand_pos = keep.rand_pos;
    h->sc_host.pair_sphere = keep.pair_sphere;
    h->sc_host.pair_bounding = keep.pair_bounding;
    h->sc_host.pair_exec = keep.pair_exec;
    h->sc_host.searches = keep.searches;
    h->sc_host.conflicts = keep.conflicts;
    h->sc_host.event = keep.event;
    h->sc_host.avg_npp = keep.avg_npp;
    h->sc_host.nucleation_accum = keep.nucleation_accum;
    TRY(push_scalars(h));
    // translate the copies with the new box length and recompute every Verlet cell (aggregat_list.cpp:169-183)
    k_dup_finish<<<div_up(t.n_agg, 8), 256, 0, h->stream>>>(h->d, (int)n0, old_l);
    h->launches++;
    CK(cudaGetLastError());
    h->cells_valid = false;
    // PhysicalModel::update at the end of duplication (:186-189): totals are unchanged x8, concentrations recomputed
    TRY(refresh_reduce(h));
    TRY(pull_scalars(h));
    h->sc_host.max_time_step = keep.max_time_step;  // duplication does not call refresh()
    h->sc_host.avg_npp = keep.avg_npp;
    TRY(push_scalars(h));
    return E_OK;
}
}  // namespace

namespace {
void fill_devstate_params(mcac_gpu *h) {
    DevState &d = h->d;
    const mcac_params &p = h->prm;
    d.n_div = p.n_verlet_divisions;
    d.n_cells = p.n_verlet_divisions * p.n_verlet_divisions * p.n_verlet_divisions;
    d.gas.mean_free_path = p.gaz_mean_free_path;
    d.gas.viscosity = p.viscosity;
    d.gas.temperature = p.temperature;
    d.gas.fractal_dimension = p.fractal_dimension;
    d.gas.density = p.density;
    d.gas.with_maturity = p.with_maturity;
    d.u_sg = p.u_sg;
    d.rp_min_oxid = p.rp_min_oxid;
    d.volsurf_method = p.volsurf_method;
    d.pick_method = p.pick_method;
    d.with_collisions = p.with_collisions;
    d.n_iter_limit = p.n_iter_without_event_limit;
    d.n_agg_limit = p.number_of_aggregates_limit;
    d.npp_limit = p.mean_monomere_per_aggregate_limit;
    d.time_limit = p.physical_time_limit;
    d.with_external_potentials = p.with_external_potentials;
    d.nucl_mean_diameter = p.mean_diameter_nucleation;
    d.nucl_dispersion_diameter = p.dispersion_diameter_nucleation;
    d.flux_nucleation = p.flux_nucleation;
    d.init_mode_normal = p.normal_initialisation;
    d.cand_cap = kCandCap;
    if (const char *e = getenv("MCAC_B200_CAND_CAP")) d.cand_cap = std::max(1, std::min(kCandCap, atoi(e)));
}

// carve the staging buffer into the host-layout arrays (256-byte aligned); returns the bytes needed
size_t carve_stage(HostLayout &L, char *base, long long n_sph, long long n_agg) {
    size_t off = 0;
    auto take = [&](auto **p, size_t count) {
        using T = std::remove_pointer_t<std::remove_pointer_t<decltype(p)>>;
        *p = reinterpret_cast<T *>(base + off);
        off += (count * sizeof(T) + 255) / 256 * 256;
    };
    L.n_sph = n_sph;
    L.n_agg = n_agg;
    take(&L.sph, (size_t)(9 * n_sph));
    take(&L.sph_label, (size_t)n_sph);
    take(&L.sph_charge, (size_t)n_sph);
    take(&L.agg, (size_t)(21 * n_agg));
    take(&L.agg_n, (size_t)n_agg);
    take(&L.agg_charge, (size_t)n_agg);
    take(&L.agg_cells, (size_t)(3 * n_agg));
    take(&L.offsets, (size_t)n_agg + 1);
    take(&L.members, (size_t)n_sph);
    take(&L.per_member, (size_t)(3 * n_sph));
    return off;
}
int ensure_stage(mcac_gpu *h, size_t bytes) {
    if (h->stage_bytes >= bytes) return E_OK;
    if (h->stage) cudaFree(h->stage);
    h->stage = nullptr;
    h->stage_bytes = 0;
    const size_t want = bytes + bytes / 8;
    CK(cudaMalloc(&h->stage, want));
    h->stage_bytes = want;
    return E_OK;
}

int upload(mcac_gpu *h, const HostState &s, double maxradius, double max_time_step, bool keep_scalars) {
    (void)keep_scalars;
    HostView v{s.n_sph, s.n_agg, s.sph.data(), s.sph_charge.data(), s.agg.data(), s.agg_charge.data(), s.agg_cells.data(),
               s.offsets.data(), s.members.data(), s.per_member.data()};
    return upload(h, v, maxradius, max_time_step);
}

// Host SoA -> HBM: the arrays cross PCIe as they are (one async copy each, full bandwidth from pinned memory); the
// gather into the aggregate-major pool is K-side work (k_upload_spheres / k_upload_aggregates).
int upload(mcac_gpu *h, const HostView &s, double maxradius, double max_time_step) {
    DevState &d = h->d;
    fill_devstate_params(h);
    const long long n_agg = s.n_agg, n_sph = s.n_sph;
    if (n_agg <= 0 || n_sph <= 0) { h->err = "upload_state: empty state"; return E_INPUT; }
    if (s.offsets[0] != 0 || s.offsets[n_agg] != n_sph) { h->err = "upload_state: membership offsets do not cover the spheres"; return E_INPUT; }
    const long long headroom = h->prm.with_nucleation ? std::max<long long>(h->nucl_headroom > 0 ? 0 : n_agg, h->nucl_headroom > 0 ? h->nucl_headroom : 4096) : 0;  // slots for nucleated monomers
    // tables: what the state needs now, or more when the caller reserved room (mcac_gpu_reserve) or — small realizations that duplicate
    // their domain — for the next duplication (x8), so that the step loop can do it without the host
    long long want_sph = std::max(n_sph, h->reserve_sph), want_agg = std::max(n_agg, h->reserve_agg);
    if (h->prm.with_domain_duplication && h->reserve_sph == 0 && n_sph <= 100000) { want_sph = std::max(want_sph, 8 * n_sph); want_agg = std::max(want_agg, 8 * n_agg); }
    const long long need_agg = want_agg + headroom, need_sph = 3 * (want_sph + headroom) + 1024;
    if (h->owned.empty() || d.agg_cap < need_agg || d.sph_cap < need_sph) {  // otherwise the resident allocation is reused
        free_all(h);
        TRY(alloc_state(h, need_agg, need_sph));
    }
    HostLayout L;
    TRY(ensure_stage(h, carve_stage(L, nullptr, n_sph, n_agg)));
    carve_stage(L, (char *)h->stage, n_sph, n_agg);
    auto up = [&](void *dst, const void *src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream); };
    CK(up(L.sph, s.sph, sizeof(double) * 9 * n_sph));
    if (s.sph_charge) CK(up(L.sph_charge, s.sph_charge, sizeof(long long) * n_sph));
    else CK(cudaMemsetAsync(L.sph_charge, 0, sizeof(long long) * n_sph, h->stream));
    CK(up(L.agg, s.agg, sizeof(double) * 21 * n_agg));
    if (s.agg_charge) CK(up(L.agg_charge, s.agg_charge, sizeof(long long) * n_agg));
    else CK(cudaMemsetAsync(L.agg_charge, 0, sizeof(long long) * n_agg, h->stream));
    CK(up(L.agg_cells, s.agg_cells, sizeof(long long) * 3 * n_agg));
    CK(up(L.offsets, s.offsets, sizeof(long long) * (n_agg + 1)));
    CK(up(L.members, s.members, sizeof(long long) * n_sph));
    CK(up(L.per_member, s.per_member, sizeof(double) * 3 * n_sph));
    k_upload_spheres<<<div_up(n_sph, 256), 256, 0, h->stream>>>(d, L);
    k_upload_aggregates<<<div_up(n_agg, 256), 256, 0, h->stream>>>(d, L, h->prm.density);
    h->launches += 2;
    CK(cudaGetLastError());
    Scalars &sc = h->sc_host;
    const Scalars old = sc;
    std::memset(&sc, 0, sizeof(sc));
    sc.time = h->prm.time;
    sc.box_length = h->prm.box_length;
    sc.box_volume = h->prm.box_volume;
    sc.maxradius = maxradius;
    sc.max_time_step = max_time_step;
    sc.avg_npp = static_cast<double>(n_sph) / static_cast<double>(n_agg);
    sc.nucleation_accum = h->prm.nucleation_accum;
    sc.n_monomeres = h->prm.n_monomeres;
    sc.n_agg = sc.n_agg_slots = (int)n_agg;
    sc.n_sph = sc.pool_top = (int)n_sph;
    sc.event = 1;
    sc.rand_pos = old.rand_pos;
    sc.pair_exec = old.pair_exec;
    sc.aggregate_concentration = static_cast<double>(n_agg) / sc.box_volume;
    sc.monomer_concentration = static_cast<double>(n_sph) / sc.box_volume;
    TRY(push_scalars(h));  // synchronizes the stream: the host arrays may be reused on return
    h->labels_valid = true;
    h->pick_valid = false;
    h->cells_valid = false;
    h->uploaded = true;
    return E_OK;
}

// HBM -> host SoA in the reference's layout (label order, creation-order sphere indices); any pointer may be null
int download_to(mcac_gpu *h, const HostOut &o) {
    DevState &d = h->d;
    TRY(pull_scalars(h));
    TRY(refresh_labels(h));
    const Scalars &sc = h->sc_host;
    const long long n_agg = sc.n_agg, n_sph = sc.n_sph;
    HostLayout L;
    TRY(ensure_stage(h, carve_stage(L, nullptr, n_sph, n_agg)));
    carve_stage(L, (char *)h->stage, n_sph, n_agg);
    k_download_counts<<<div_up(n_agg, 256), 256, 0, h->stream>>>(d, (int)n_agg, h->scan_tmp);
    TRY(scan_ints(h, h->scan_tmp, (int)n_agg, h->scan_out));
    int total = 0;
    CK(cudaMemcpyAsync(&total, h->scan_out + n_agg, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (total != n_sph) { h->err = "download_state: membership does not match the sphere count"; return E_UNKNOWN; }
    k_download_state<<<div_up(n_agg * 32, 256), 256, 0, h->stream>>>(d, L, h->scan_out);
    h->launches += 2;
    CK(cudaGetLastError());
    auto dn = [&](void *dst, const void *src, size_t bytes) {
        return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream) : cudaSuccess;
    };
    CK(dn(o.sph, L.sph, sizeof(double) * 9 * n_sph));
    CK(dn(o.sph_label, L.sph_label, sizeof(long long) * n_sph));
    CK(dn(o.sph_charge, L.sph_charge, sizeof(long long) * n_sph));
    CK(dn(o.agg, L.agg, sizeof(double) * 21 * n_agg));
    CK(dn(o.agg_n, L.agg_n, sizeof(long long) * n_agg));
    CK(dn(o.agg_charge, L.agg_charge, sizeof(long long) * n_agg));
    CK(dn(o.agg_cells, L.agg_cells, sizeof(long long) * 3 * n_agg));
    CK(dn(o.offsets, L.offsets, sizeof(long long) * (n_agg + 1)));
    CK(dn(o.members, L.members, sizeof(long long) * n_sph));
    CK(dn(o.per_member, L.per_member, sizeof(double) * 3 * n_sph));
    CK(cudaStreamSynchronize(h->stream));
    if (o.scalars) {
        const double sv[20] = {sc.time, sc.box_length, sc.maxradius, sc.max_time_step, sc.avg_npp, sc.volume_fraction,
                               sc.aggregate_concentration, sc.monomer_concentration, sc.total_volume_concent, sc.total_surface_concent,
                               h->prm.u_sg, h->prm.gaz_mean_free_path, 0., 0., h->prm.viscosity, sc.box_volume,
                               (double)sc.n_iter_without_event, (double)sc.n_monomeres, h->prm.temperature, sc.nucleation_accum};
        std::memcpy(o.scalars, sv, sizeof(sv));
    }
    return E_OK;
}

int download(mcac_gpu *h, HostState &s) {
    TRY(pull_scalars(h));
    const long long n_agg = h->sc_host.n_agg, n_sph = h->sc_host.n_sph;
    s.n_agg = n_agg;
    s.n_sph = n_sph;
    s.sph.resize((size_t)(9 * n_sph));
    s.sph_label.resize((size_t)n_sph);
    s.sph_charge.resize((size_t)n_sph);
    s.agg.resize((size_t)(21 * n_agg));
    s.agg_n.resize((size_t)n_agg);
    s.agg_charge.resize((size_t)n_agg);
    s.agg_cells.resize((size_t)(3 * n_agg));
    s.offsets.resize((size_t)n_agg + 1);
    s.members.resize((size_t)n_sph);
    s.per_member.resize((size_t)(3 * n_sph));
    return download_to(h, HostOut{s.sph.data(), s.sph_label.data(), s.sph_charge.data(), s.agg.data(), s.agg_n.data(), s.agg_charge.data(),
                                  s.agg_cells.data(), s.offsets.data(), s.members.data(), s.per_member.data(), s.scalars});
}

// PhysicalModel::finished (physical_model.cpp:288-337) on the mirrored scalars; wall-clock limits are not part of the path
bool finished(const mcac_gpu *h) {
    const Scalars &sc = h->sc_host;
    const mcac_params &p = h->prm;
    if (sc.n_agg < 1) return true;
    if (sc.n_agg <= p.number_of_aggregates_limit) return true;
    if (p.n_iter_without_event_limit > 0 && sc.n_iter_without_event >= p.n_iter_without_event_limit) return true;
    if (p.physical_time_limit > 0 && sc.time >= p.physical_time_limit) return true;
    if (p.mean_monomere_per_aggregate_limit > 0 && sc.avg_npp >= (double)p.mean_monomere_per_aggregate_limit) return true;
    return false;
}

void prof_begin(mcac_gpu *h, int kind) {
    if (!h->profile) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    h->ev_pool.push_back(a);
    h->ev_pool.push_back(b);
    h->ev_kind.push_back(kind);
    cudaEventRecord(a, h->stream);
}
void prof_end(mcac_gpu *h) {
    if (!h->profile) return;
    cudaEventRecord(h->ev_pool.back(), h->stream);
}
void prof_collect(mcac_gpu *h, mcac_run_report *rep) {
    double ms[4] = {0., 0., 0., 0.};
    long long cnt[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < h->ev_kind.size(); i++) {
        float t = 0.f;
        cudaEventSynchronize(h->ev_pool[2 * i + 1]);
        cudaEventElapsedTime(&t, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]);
        ms[h->ev_kind[i]] += t;
        cnt[h->ev_kind[i]]++;
        cudaEventDestroy(h->ev_pool[2 * i]);
        cudaEventDestroy(h->ev_pool[2 * i + 1]);
    }
    h->ev_pool.clear();
    h->ev_kind.clear();
    if (rep) {
        rep->search_ms = ms[0]; rep->commit_ms = ms[1]; rep->search_launches = cnt[0]; rep->commit_launches = cnt[1];
        rep->event_ms = ms[2]; rep->cells_ms = ms[3]; rep->event_launches = cnt[2]; rep->cells_launches = cnt[3];
    }
}

// K1: group-per-query kernel, then the wide kernel over the (usually empty) list of queries it handed over
int search_kernels(mcac_gpu *h, int nq, const int *q_slot, const double *q_dir, const double *q_dist, SearchResult *res) {
    if (h->wide_cap < nq) {
        if (h->wide_list) cudaFree(h->wide_list);
        h->wide_list = nullptr;
        h->wide_cap = 0;
        CK(cudaMalloc((void **)&h->wide_list, sizeof(int) * ((size_t)nq + 2)));
        CK(cudaMemsetAsync(h->wide_list + nq, 0, 2 * sizeof(int), h->stream));
        h->wide_cap = nq;
    }
    // two counters used alternately: each launch of the group kernel clears the one the NEXT search will use
    h->wide_parity ^= 1;
    int *count = h->wide_list + h->wide_cap + h->wide_parity, *next_count = h->wide_list + h->wide_cap + (h->wide_parity ^ 1);
    // group width: wide groups for small launches (latency-bound: more lanes per query), narrow ones for big launches
    // (throughput-bound: more queries in flight); MCAC_B200_SEARCH_GROUP / _MB override (0 = wide kernel only)
    const int G = h->search_group >= 0 ? h->search_group : (nq <= 4096 ? 32 : 8);
#define MCAC_LAUNCH_GROUP(GG, MB) \
    k_search_group<GG, MB><<<div_up(nq, kSearchThreads / GG), kSearchThreads, 0, h->stream>>>(h->d, nq, q_slot, q_dir, q_dist, res, h->wide_list, count, next_count)
    const int MB = h->search_min_blocks;
    if (G == 0) CK(cudaMemsetAsync(next_count, 0, sizeof(int), h->stream));
    else if (G == 4) { if (MB >= 8) MCAC_LAUNCH_GROUP(4, 8); else if (MB >= 6) MCAC_LAUNCH_GROUP(4, 6); else MCAC_LAUNCH_GROUP(4, 4); }
    else if (G == 8) { if (MB >= 8) MCAC_LAUNCH_GROUP(8, 8); else if (MB >= 6) MCAC_LAUNCH_GROUP(8, 6); else MCAC_LAUNCH_GROUP(8, 4); }
    else if (G == 16) { if (MB >= 8) MCAC_LAUNCH_GROUP(16, 8); else MCAC_LAUNCH_GROUP(16, 4); }
    else MCAC_LAUNCH_GROUP(32, 4);
#undef MCAC_LAUNCH_GROUP
    const bool wide_only = G == 0;
    const int wide_grid = wide_only ? nq : std::min(nq, 8 * h->n_sm);
    k_search_wide<<<wide_grid, kSearchThreads, 0, h->stream>>>(h->d, nq, wide_only ? nullptr : h->wide_list, count, q_slot, q_dir, q_dist, res);
    h->launches += wide_only ? 1 : 2;
    CK(cudaGetLastError());
    return E_OK;
}
// one search between aggregates of many spheres (general step, late stages): phase 1 / tiled sphere sweep over the grid / phase 3
int search_big(mcac_gpu *h) {
    TRY(build_cells(h));
    TRY(join_cells(h));
    if (!h->big) TRY(dev_alloc_persistent(h, &h->big, 1));
    k_search_big_p1<<<1, kSearchThreads, 0, h->stream>>>(h->d, h->q_slot, h->q_dir, h->q_dist, h->q_res, h->big);
    k_search_big_p2<<<2 * h->n_sm, 256, 0, h->stream>>>(h->d, h->q_slot, h->q_dir, h->q_dist, h->big);
    k_search_big_p3<<<1, kSearchThreads, 0, h->stream>>>(h->d, h->q_slot, h->q_dir, h->q_dist, h->q_res, h->big);
    h->launches += 3;
    CK(cudaGetLastError());
    return E_OK;
}
int search_launch(mcac_gpu *h, int nq) {
    if (nq == 1 && h->sc_host.avg_npp >= h->big_search_npp) return search_big(h);
    TRY(build_cells(h));
    TRY(join_cells(h));
    return search_kernels(h, nq, h->q_slot, h->q_dir, h->q_dist, h->q_res);
}

// ---- per-realization step loop (mcac_steploop.cuh) --------------------------------------------------------------------------
bool is_speculative(const mcac_gpu *h) {
    return h->prm.pick_method == MCAC_PICK_RANDOM && h->prm.with_collisions && !h->prm.with_surface_reactions && !h->prm.with_potentials &&
           !h->prm.with_nucleation;
}
bool loop_usable(const mcac_gpu *h) {
    return h->fused && !h->debug_sync && !h->profile && h->sc_host.n_agg_slots <= h->fused_max_slots &&
           h->sc_host.n_agg <= h->cum_sequential_max;
}
void loop_fill_args(mcac_gpu *h, LoopArgs &a, long long max_steps, mcac_step_record *rec, long long rec_cap, long long rec_base) {
    const mcac_params &p = h->prm;
    a.q_slot = h->q_slot; a.q_dir = h->q_dir; a.q_dist = h->q_dist; a.q_res = h->q_res;
    a.sb = h->sortb;
    a.sorted_label = h->sorted_label;
    a.scan_tmp = h->scan_tmp;
    a.alt_posr = h->alt.s_posr; a.alt_relv = h->alt.s_relv; a.alt_surf = h->alt.s_surf; a.alt_veff = h->alt.s_veff;
    a.alt_seff = h->alt.s_seff; a.alt_dcen = h->alt.s_dcen; a.alt_id = h->alt.s_id; a.alt_charge = h->alt.s_charge;
    a.rec = rec; a.rec_cap = rec_cap; a.rec_base = rec_base;
    a.max_steps = max_steps;
    a.dup_threshold = h->dup_threshold;
    a.full_freq = std::max<long long>(1, p.full_aggregate_update_frequency);
    a.with_nucleation = p.with_nucleation; a.with_potentials = p.with_potentials; a.growth = p.with_surface_reactions;
    a.individual = p.individual_surf_reactions; a.pick_last = p.pick_method == MCAC_PICK_LAST ? 1 : 0;
    a.with_collisions = p.with_collisions; a.with_domain_duplication = p.with_domain_duplication;
    a.cum_sequential_max = h->cum_sequential_max;
    a.stable = p.sort_order == MCAC_ORDER_STABLE ? 1 : 0;
    a.depth_override = h->sort_depth_override;
    a.pick_valid = h->pick_valid ? 1 : 0; a.labels_valid = h->labels_valid ? 1 : 0;
    a.stop_at_event = h->stop_at_event ? 1 : 0;
    a.max_slots = h->fused_max_slots;
    a.prune = h->loop_prune ? 1 : 0;
    a.dups_allowed = h->loop_dups ? 4 : 0;
    for (int k = 0; k < 4; k++) a.dup_box_volume[k] = std::pow(h->prm.box_length * (double)(2 << k), 3);  // duplicate(): box x2, std::pow(box, 3)
    a.out = h->loop_dev;
}
// host-side bookkeeping after a loop launch: the pool may have been compacted (buffers swapped), validity flags
void loop_apply(mcac_gpu *h, const LoopState &ls) {
    if (ls.flipped) {
        DevState &d = h->d;
        std::swap(d.s_posr, h->alt.s_posr); std::swap(d.s_relv, h->alt.s_relv); std::swap(d.s_surf, h->alt.s_surf);
        std::swap(d.s_veff, h->alt.s_veff); std::swap(d.s_seff, h->alt.s_seff); std::swap(d.s_dcen, h->alt.s_dcen);
        std::swap(d.s_id, h->alt.s_id); std::swap(d.s_charge, h->alt.s_charge);
    }
    for (long long k = 0; k < ls.dups; k++) {  // the host's mirror of the box (duplicate(): aggregat_list.cpp:145-148)
        h->prm.box_length = h->prm.box_length * 2;
        h->prm.n_monomeres *= 8;
        h->prm.box_volume = std::pow(h->prm.box_length, 3);
    }
    h->pick_valid = ls.pick_valid != 0;
    h->labels_valid = ls.labels_valid != 0;
    h->cells_valid = false;
    h->loop_launches++;
    h->loop_steps += ls.steps;
    for (int k = 0; k < 8; k++) h->loop_cycles[k] += ls.phase_cycles[k];
}
int duplicate_on_device(mcac_gpu *h, bool *done) {
    *done = false;
    if (!h->loop_dups || h->debug_sync) return E_OK;
    const Scalars &sc = h->sc_host;
    if (sc.n_agg_slots + 7LL * sc.n_agg + 64 > h->d.agg_cap || 8LL * sc.n_sph + 64 > h->d.sph_cap || 9LL * sc.n_sph + 64 > h->d.sph_cap) return E_OK;
    LoopArgs la;
    loop_fill_args(h, la, 0, nullptr, 0, 0);
    k_duplicate<<<1, kLoopThreads, kLoopDynSmem, h->stream>>>(h->d, la);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->loop_host, h->loop_dev, sizeof(LoopState), cudaMemcpyDeviceToHost, h->stream));
    TRY(pull_scalars(h));
    const LoopState ls = *h->loop_host;
    if (ls.dups != 1) return E_OK;  // the kernel found the tables too small after all: nothing was touched
    loop_apply(h, ls);
    h->loop_launches--;  // (not a step-loop launch)
    h->pick_valid = false;
    *done = true;
    return E_OK;
}
// slots for nucleated monomers: regrow through the upload boundary when the headroom is nearly used up.  This renumbers the
// aggregate slots (slot = label again), so it must come BEFORE the pick table of the step is built.
int regrow_tables(mcac_gpu *h) {
    HostState hs;
    TRY(download(h, hs));
    const Scalars keep = h->sc_host;
    h->uploaded = false;
    TRY(upload(h, hs, keep.maxradius, keep.max_time_step, false));
    Scalars &sc = h->sc_host;
    const int n_slots = sc.n_agg_slots, pool = sc.pool_top;
    sc = keep;
    sc.n_agg_slots = n_slots; sc.pool_top = pool;
    TRY(push_scalars(h));
    return E_OK;
}
}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int mcac_gpu_create(const mcac_params *params, int device, mcac_gpu **out) {
    if (!params || !out) return E_INPUT;
    mcac_gpu *h = new mcac_gpu();
    *out = h;
    h->prm = *params;
    h->device = device;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= device) {
        h->err = "mcac_b200 needs a CUDA device (sm_100a); there is no CPU fallback";
        return E_UNKNOWN;
    }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->n_sm = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CK(cudaMallocHost((void **)&h->h_sc, sizeof(Scalars)));
    for (int k = 0; k < 2; k++) {
        CK(cudaMallocHost((void **)&h->h_sc_ring[k], sizeof(Scalars)));
        CK(cudaEventCreateWithFlags(&h->ev_cycle[k], cudaEventDisableTiming));
    }
    CK(cudaMallocHost((void **)&h->h_flags, 4 * sizeof(int)));
    CK(cudaMallocHost((void **)&h->loop_host, sizeof(LoopState)));
    fill_devstate_params(h);
    TRY(alloc_persistent(h));
    {
        int occ = 0, coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        if (getenv("MCAC_B200_NO_OVERLAP")) h->overlap = false;
        if (getenv("MCAC_B200_NO_PIPELINE")) h->pipeline = false;
        if (getenv("MCAC_B200_DEBUG_SYNC")) h->debug_sync = true;
        if (const char *e = getenv("MCAC_B200_EVENT_SPARE_SMS")) h->event_spare_sms = std::max(0, atoi(e));
        if (const char *e = getenv("MCAC_B200_NUCL_HEADROOM")) h->nucl_headroom = std::max(80, atoi(e));
        if (const char *e = getenv("MCAC_B200_BIG_NPP")) h->big_search_npp = atof(e);
        if (const char *e = getenv("MCAC_B200_FORCE_SORT_FAIL")) h->force_sort_fail = atoi(e);
        if (const char *e = getenv("MCAC_B200_SORT_DEPTH")) h->sort_depth_override = std::max(0, atoi(e));
        if (getenv("MCAC_B200_NO_LOOP")) h->fused = false;
        if (getenv("MCAC_B200_NO_PRUNE")) h->loop_prune = false;
        if (getenv("MCAC_B200_NO_LOOP_DUP")) h->loop_dups = false;
        if (const char *e = getenv("MCAC_B200_RESERVE_SPHERES")) h->reserve_sph = std::max(0LL, atoll(e));
        if (const char *e = getenv("MCAC_B200_RESERVE_AGGREGATES")) h->reserve_agg = std::max(0LL, atoll(e));
        if (getenv("MCAC_B200_STRICT_DIRECTION")) h->strict_dir = true;
        if (const char *e = getenv("MCAC_B200_LOOP_MAX_SLOTS")) h->fused_max_slots = std::max(1, atoi(e));
        if (const char *e = getenv("MCAC_B200_SEARCH_GROUP")) h->search_group = atoi(e);
        if (const char *e = getenv("MCAC_B200_SEARCH_MB")) h->search_min_blocks = atoi(e);
        if (const char *e = getenv("MCAC_B200_COOP_BPS")) h->coop_bps = atoi(e) >= 2 ? 2 : 1;
        if (const char *e = getenv("MCAC_B200_SORT_LOCAL")) h->sort_local_span = std::max(64, atoi(e));
        if (const char *e = getenv("MCAC_B200_SORT_SWITCH")) h->sort_switch_span = std::max(64, atoi(e));
        if (const char *e = getenv("MCAC_B200_TIE_MIN_N")) h->ts_min_n = std::max(0, atoi(e));
        if (const char *e = getenv("MCAC_B200_TIE_MAX_SPARSE")) h->ts_max_sparse = std::max(1, atoi(e));
        if (getenv("MCAC_B200_TIE_NO_OVERLAP")) h->ts_no_overlap = true;
        if (getenv("MCAC_B200_NO_EXACT_CUM")) h->no_exact_cum = true;
        if (const char *e = getenv("MCAC_B200_EXACT_CUM_MAX_SPARSE")) h->exact_cum_max_sparse = std::max(0, atoi(e));
        // block-local sort levels staged in shared memory: (local_span + 2) entries of 48 B, if the SM has room for them
        h->event_smem_cap = 0;
        if (!getenv("MCAC_B200_NO_SORT_SMEM")) {
            int max_optin = 0;
            cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
            const long long want = (long long)(h->sort_local_span + 2) * kSortStageBytesPerEntry;
            if (want + 4096 <= max_optin / h->coop_bps) h->event_smem_cap = h->sort_local_span + 2;
        }
        size_t dyn = (size_t)h->event_smem_cap * kSortStageBytesPerEntry;
        if (h->ts_min_n > 0 && !getenv("MCAC_B200_NO_SORT_SMEM")) {  // scratch of the one-CTA sparse simulation (tie_sort.cuh)
            int max_optin = 0;
            cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
            const size_t need = sizeof(int) * (4 * ((size_t)std::min(h->ts_max_sparse, tiesort::kMaxSparse) + 4) + tiesort::kTblStride + tiesort::kMiscInts + 2 * 1024);
            if ((long long)need + 4096 <= max_optin / h->coop_bps) dyn = std::max(dyn, need);
        }
        const void *efn = h->coop_bps == 2 ? (const void *)k_event<2> : (const void *)k_event<1>;
        // (the attribute belongs to the function, not to the handle: it is only ever raised, or a handle created later with a smaller
        // need — the sparse path switched off, say — would take the launches of the earlier ones below their dynamic size)
        static size_t attr_set[64][2] = {};
        static std::mutex attr_mu;
        {
            std::lock_guard<std::mutex> lk(attr_mu);
            size_t &cur = attr_set[device & 63][h->coop_bps == 2 ? 1 : 0];
            if (dyn > cur) {
                if (cudaFuncSetAttribute(efn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) {
                    cudaGetLastError();
                    h->event_smem_cap = 0;
                    dyn = 0;
                } else cur = dyn;
            }
        }
        h->event_dyn_bytes = dyn;
        cudaError_t oe = h->coop_bps == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_event<2>, kEventThreads, dyn)
                                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_event<1>, kEventThreads, dyn);
        if (coop && oe == cudaSuccess && occ > 0) h->coop_blocks = h->n_sm * std::min(occ, h->coop_bps);
        if (getenv("MCAC_B200_NO_COOP")) h->coop_blocks = 0;
        // the step-loop kernels use 31 KB of static + 20 KB of dynamic shared memory: opt in above 48 KB
        cudaFuncSetAttribute(k_step_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, kLoopDynSmem);
        cudaFuncSetAttribute(k_ensemble_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, kLoopDynSmem);
        cudaGetLastError();
        if (getenv("MCAC_B200_K9_DEBUG")) {
            TRY(dev_alloc_persistent(h, &h->commit_prof, 16));
            CK(cudaMemset(h->commit_prof, 0, 16 * sizeof(long long)));
        }
        TRY(dev_alloc_persistent(h, &h->event_work, 40));
        CK(cudaMemset(h->event_work, 0, 40 * sizeof(long long)));
        TRY(dev_alloc_persistent(h, &h->part_ll, 4096));
        TRY(dev_alloc_persistent(h, &h->part_d, 4 * 4096));
    }
    GlibcRandState st;
    glibc_srand(st, params->random_seed);
    CK(cudaMemcpy(h->d.rng, &st, sizeof(st), cudaMemcpyHostToDevice));
    h->d.rng_buf_base = 0;
    h->d.rng_buf_n = 0;
    h->rng_generated = 0;
    return E_OK;
}

int mcac_gpu_destroy(mcac_gpu *h) {
    if (!h) return E_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_all(h);
    for (void *p : h->persistent) cudaFree(p);
    if (h->rec_dev) cudaFree(h->rec_dev);
    if (h->stage) cudaFree(h->stage);
    if (h->wide_list) cudaFree(h->wide_list);
    for (void *p : {(void *)h->sweep_slot, (void *)h->sweep_dir, (void *)h->sweep_dist, (void *)h->sweep_res}) if (p) cudaFree(p);
    if (h->h_sc) cudaFreeHost(h->h_sc);
    for (int k = 0; k < 2; k++) {
        if (h->h_sc_ring[k]) cudaFreeHost(h->h_sc_ring[k]);
        if (h->ev_cycle[k]) cudaEventDestroy(h->ev_cycle[k]);
    }
    if (h->h_flags) cudaFreeHost(h->h_flags);
    if (h->loop_host) cudaFreeHost(h->loop_host);
    if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return E_OK;
}

const char *mcac_gpu_last_error(const mcac_gpu *h) { return h ? h->err.c_str() : "null handle"; }

int mcac_host_alloc_pinned(int64_t bytes, void **out) {
    if (!out || bytes < 0) return E_INPUT;
    *out = nullptr;
    return cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault) == cudaSuccess ? E_OK : E_UNKNOWN;
}
int mcac_host_free_pinned(void *p) { return (!p || cudaFreeHost(p) == cudaSuccess) ? E_OK : E_UNKNOWN; }
void *mcac_gpu_stream(mcac_gpu *h) { return h ? (void *)h->stream : nullptr; }

// srand(seed) followed by `consumed` draws already taken by the host-side initial placement (a23)
int mcac_gpu_set_rng(mcac_gpu *h, uint32_t seed, int64_t consumed) {
    CK(cudaSetDevice(h->device));
    GlibcRandState st;
    glibc_srand(st, seed);
    for (int64_t i = 0; i < consumed; i++) glibc_rand_next(st);
    GlibcRandState rot;  // rotate so that the device generator always starts a block at ring position 0
    for (int i = 0; i < 31; i++) rot.ring[i] = st.ring[(st.pos + i) % 31];
    rot.pos = 0;
    CK(cudaMemcpy(h->d.rng, &rot, sizeof(rot), cudaMemcpyHostToDevice));
    h->sc_host.rand_pos = consumed;
    h->d.rng_buf_base = consumed;
    h->d.rng_buf_n = 0;
    h->rng_generated = consumed;
    h->dir_tab_valid = false;
    if (h->uploaded) TRY(push_scalars(h));
    return E_OK;
}

int mcac_gpu_set_interpotential(mcac_gpu *h, int32_t n1, int32_t n2, int32_t nq, const int32_t *val_charge, const double *val_dp1,
                                const double *val_dp2, const double *e_barr, const double *e_well) {
    CK(cudaSetDevice(h->device));
    DevState &d = h->d;
    const size_t nt = (size_t)nq * nq * n1 * n2;
    int *c; double *a, *b, *eb, *ew;
    TRY(dev_alloc_persistent(h, &c, (size_t)nq));
    TRY(dev_alloc_persistent(h, &a, (size_t)n1));
    TRY(dev_alloc_persistent(h, &b, (size_t)n2));
    TRY(dev_alloc_persistent(h, &eb, nt));
    TRY(dev_alloc_persistent(h, &ew, nt));
    CK(cudaMemcpy(c, val_charge, sizeof(int) * nq, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(a, val_dp1, sizeof(double) * n1, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b, val_dp2, sizeof(double) * n2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(eb, e_barr, sizeof(double) * nt, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ew, e_well, sizeof(double) * nt, cudaMemcpyHostToDevice));
    d.ip_n1 = n1; d.ip_n2 = n2; d.ip_nq = nq;
    d.ip_charge = c; d.ip_dp1 = a; d.ip_dp2 = b; d.ip_ebar = eb; d.ip_ewell = ew;
    return E_OK;
}

int mcac_gpu_upload_state(mcac_gpu *h, int64_t n_sph, int64_t n_agg, const double *sphere_fields, const int64_t *sphere_charge,
                          const double *agg_fields, const int64_t *agg_charge, const int64_t *agg_cells, const int64_t *offsets,
                          const int64_t *members, const double *per_member, double maxradius, double max_time_step) {
    CK(cudaSetDevice(h->device));
    if (!sphere_fields || !agg_fields || !agg_cells || !offsets || !members || !per_member) { h->err = "upload_state: null array"; return E_INPUT; }
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
    const HostView s{n_sph, n_agg, sphere_fields, (const long long *)sphere_charge, agg_fields, (const long long *)agg_charge,
                     (const long long *)agg_cells, (const long long *)offsets, (const long long *)members, per_member};
    const bool had_state = h->uploaded;
    const Scalars keep = h->sc_host;
    if (had_state) TRY(pull_scalars(h));
    const Scalars live = h->sc_host;
    h->uploaded = false;
    TRY(upload(h, s, maxradius, max_time_step));
    if (had_state) {  // the realization keeps running: clocks, counters and the RNG position are the handle's, not the host's
        Scalars &sc = h->sc_host;
        sc.time = live.time; sc.n_iter_without_event = live.n_iter_without_event; sc.total_events = live.total_events;
        sc.steps_done = live.steps_done; sc.rand_pos = live.rand_pos; sc.pair_sphere = live.pair_sphere;
        sc.pair_bounding = live.pair_bounding; sc.searches = live.searches; sc.conflicts = live.conflicts; sc.pair_exec = live.pair_exec;
        sc.nucleation_accum = live.nucleation_accum; sc.event = 1;
        sc.total_volume = live.total_volume; sc.total_surface = live.total_surface; sc.volume_fraction = live.volume_fraction;
        TRY(push_scalars(h));
    } else {
        h->dup_threshold = n_agg / 8;  // calcul.cpp:58
    }
    (void)keep;
    return E_OK;
}

int mcac_gpu_sizes(mcac_gpu *h, int64_t *n_sph, int64_t *n_agg) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    if (n_sph) *n_sph = h->sc_host.n_sph;
    if (n_agg) *n_agg = h->sc_host.n_agg;
    return E_OK;
}

int mcac_gpu_download_state(mcac_gpu *h, double *sphere_fields, int64_t *sphere_label, int64_t *sphere_charge, double *agg_fields,
                            int64_t *agg_n_spheres, int64_t *agg_charge, int64_t *agg_cells, int64_t *offsets, int64_t *members,
                            double *per_member, double *scalars) {
    CK(cudaSetDevice(h->device));
    return download_to(h, HostOut{sphere_fields, (long long *)sphere_label, (long long *)sphere_charge, agg_fields, (long long *)agg_n_spheres,
                                  (long long *)agg_charge, (long long *)agg_cells, (long long *)offsets, (long long *)members, per_member,
                                  scalars});
}

int mcac_gpu_contact_search_batch(mcac_gpu *h, int64_t n, const int64_t *source_labels, const double *directions, const double *distances,
                                  mcac_contact *out, int64_t *pair_tests) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    TRY(refresh_labels(h));
    std::vector<SearchResult> res(kMaxBatch);
    std::vector<int> sid;
    int64_t ps = 0, pb = 0;
    for (int64_t base = 0; base < n; base += kMaxBatch) {
        const int nq = (int)std::min<int64_t>(kMaxBatch, n - base);
        CK(cudaMemcpyAsync(h->q_label, source_labels + base, sizeof(int64_t) * nq, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->q_dir, directions + 3 * base, sizeof(double) * 3 * nq, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->q_dist, distances + base, sizeof(double) * nq, cudaMemcpyHostToDevice, h->stream));
        k_labels_to_slots<<<div_up(nq, 128), 128, 0, h->stream>>>(h->d, nq, (const long long *)h->q_label, h->q_slot);
        h->launches++;
        TRY(search_launch(h, nq));
        CK(cudaMemcpyAsync(res.data(), h->q_res, sizeof(SearchResult) * nq, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        // ids: sphere slots -> creation indices, aggregate slots -> labels
        std::vector<int> want;
        for (int j = 0; j < nq; j++) {
            out[base + j].distance = res[(size_t)j].distance;
            out[base + j].moving_sphere = out[base + j].other_sphere = out[base + j].moving_label = out[base + j].other_label = -1;
            ps += res[(size_t)j].n_sphere_pairs;
            pb += res[(size_t)j].n_bounding;
            if (res[(size_t)j].status == 3) { h->err = "Aggregate not on the verlet list ???"; return E_VERLET; }
            if (res[(size_t)j].status == 1) { h->err = "contact search: suspect list overflow"; return E_UNKNOWN; }
        }
        for (int j = 0; j < nq; j++) {
            const SearchResult &r = res[(size_t)j];
            if (r.other_agg < 0) continue;
            int ids[2], lab;
            CK(cudaMemcpyAsync(&ids[0], h->d.s_id + r.moving_slot, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(&ids[1], h->d.s_id + r.other_slot, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(&lab, h->d.label_of_slot + r.other_agg, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            out[base + j].moving_sphere = ids[0];
            out[base + j].other_sphere = ids[1];
            out[base + j].moving_label = source_labels[base + j];
            out[base + j].other_label = lab;
        }
    }
    if (pair_tests) { pair_tests[0] = ps; pair_tests[1] = pb; }
    return E_OK;
}

int mcac_gpu_contact_search(mcac_gpu *h, int64_t source_label, const double direction[3], double distance, mcac_contact *out) {
    return mcac_gpu_contact_search_batch(h, 1, &source_label, direction, &distance, out, nullptr);
}

int mcac_gpu_translate(mcac_gpu *h, int64_t label, const double vector[3]) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    TRY(refresh_labels(h));
    if (label < 0 || label >= h->sc_host.n_agg) { h->err = "translate: bad label"; return E_INPUT; }
    int slot;
    CK(cudaMemcpyAsync(&slot, h->d.slot_of_label + label, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    k_translate_one<<<1, 128, 0, h->stream>>>(h->d, slot, vector[0], vector[1], vector[2]);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->cells_valid = false;
    return E_OK;
}

int mcac_gpu_merge(mcac_gpu *h, const mcac_contact *c, int *merged) {
    CK(cudaSetDevice(h->device));
    if (merged) *merged = 0;
    if (!c || c->moving_sphere < 0 || c->other_sphere < 0) return E_OK;  // expired weak_ptrs: merge() returns false
    TRY(pull_scalars(h));
    if (h->d.sph_cap - h->sc_host.pool_top < h->sc_host.n_sph) TRY(compact_pool(h));
    k_merge_one<<<1, kCommitThreads, 0, h->stream>>>(h->d, (int)c->moving_sphere, (int)c->other_sphere, h->merged_flag);
    h->launches++;
    CK(cudaGetLastError());
    int m = 0;
    CK(cudaMemcpyAsync(&m, h->merged_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (merged) *merged = m;
    if (m) { h->labels_valid = false; h->cells_valid = false; h->pick_valid = false; }
    TRY(pull_scalars(h));
    if (h->sc_host.error) return device_error(h, h->sc_host, "mcac_gpu_merge");
    return E_OK;
}

int mcac_gpu_grow(mcac_gpu *h, double dt, int64_t label) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    TRY(refresh_labels(h));
    int slot = -1, n = h->sc_host.pool_top;
    if (label >= h->sc_host.n_agg) { h->err = "grow: bad label"; return E_INPUT; }
    if (label >= 0) {
        CK(cudaMemcpyAsync(&slot, h->d.slot_of_label + label, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        n = h->sc_host.n_sph;
    }
    k_grow<<<div_up(n, 256), 256, 0, h->stream>>>(h->d, dt, slot);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return E_OK;
}

int mcac_gpu_update(mcac_gpu *h, int64_t label, int full) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    TRY(refresh_labels(h));
    int slot = -1;
    if (label >= h->sc_host.n_agg) { h->err = "update: bad label"; return E_INPUT; }
    if (label >= 0) {
        CK(cudaMemcpyAsync(&slot, h->d.slot_of_label + label, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    const int nblk = label >= 0 ? 1 : div_up(h->sc_host.n_agg_slots, 8);
    if (label < 0) { k_update_small<<<div_up(h->sc_host.n_agg_slots, 128), 128, 0, h->stream>>>(h->d, full, 0, 1); h->launches++; }
    if (label >= 0) {
        k_update_one<<<1, kCommitThreads, 0, h->stream>>>(h->d, full, slot);
        h->launches++;
    } else {
        k_update_all<<<nblk, 256, 0, h->stream>>>(h->d, full, slot);
        k_update_big<<<std::min(h->sc_host.n_agg_slots, 4 * h->n_sm), kCommitThreads, 0, h->stream>>>(h->d, full, 0);
        h->launches += 2;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->cells_valid = false;
    TRY(pull_scalars(h));
    if (h->sc_host.error) return device_error(h, h->sc_host, "mcac_gpu_update");
    return E_OK;
}

int mcac_gpu_aggregate_fields(mcac_gpu *h, int64_t label, double fields[21], int64_t *n_spheres) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    TRY(refresh_labels(h));
    if (label < 0 || label >= h->sc_host.n_agg) { h->err = "aggregate_fields: bad label"; return E_INPUT; }
    int slot;
    CK(cudaMemcpyAsync(&slot, h->d.slot_of_label + label, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    k_aggregate_fields<<<1, 32, 0, h->stream>>>(h->d, slot, h->stats_dev);
    h->launches++;
    double v[22];
    CK(cudaMemcpyAsync(v, h->stats_dev, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (fields) std::memcpy(fields, v, 21 * sizeof(double));
    if (n_spheres) *n_spheres = (int64_t)v[21];
    return E_OK;
}

int mcac_gpu_refresh(mcac_gpu *h, double *max_time_step, double *avg_npp, double *total_volume, double *total_surface) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    TRY(refresh_reduce(h));
    TRY(pull_scalars(h));
    if (max_time_step) *max_time_step = h->sc_host.max_time_step;
    if (avg_npp) *avg_npp = h->sc_host.avg_npp;
    if (total_volume) *total_volume = h->sc_host.total_volume;
    if (total_surface) *total_surface = h->sc_host.total_surface;
    return E_OK;
}

int mcac_gpu_sort_time_steps(mcac_gpu *h, double factor) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    TRY(event_pipeline(h, false, false, true, &factor));
    return E_OK;
}

int mcac_gpu_get_pick_table(mcac_gpu *h, int64_t *index_sorted, double *cumulative, int64_t *n_out) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    const int n = h->sc_host.n_pick;
    if (n_out) *n_out = n;
    std::vector<int> lab((size_t)n);
    CK(cudaMemcpyAsync(lab.data(), h->sorted_label, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    if (cumulative) CK(cudaMemcpyAsync(cumulative, h->d.cum, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (index_sorted) for (int i = 0; i < n; i++) index_sorted[i] = lab[(size_t)i];
    return E_OK;
}

int mcac_gpu_pick_random(mcac_gpu *h, double u, int64_t *label, double *deltatemps) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    if (!h->pick_valid) { h->err = "pick_random before sort_time_steps"; return E_INPUT; }
    const int n = h->sc_host.n_pick;
    if (n < 1) { h->err = "pick_random: empty pick table"; return E_INPUT; }
    std::vector<double> cum((size_t)n);
    std::vector<int> lab((size_t)n);
    CK(cudaMemcpyAsync(cum.data(), h->d.cum, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(lab.data(), h->sorted_label, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const double val = u * cum[(size_t)n - 1];
    const long k = std::lower_bound(cum.begin(), cum.end(), val) - cum.begin();
    if (label) *label = lab[(size_t)k];
    if (deltatemps) *deltatemps = h->sc_host.max_time_step / cum[(size_t)n - 1];
    return E_OK;
}

int mcac_gpu_pick_last(mcac_gpu *h, int64_t *label) {
    CK(cudaSetDevice(h->device));
    HostState s;
    TRY(download(h, s));
    int64_t latest = 0;
    double t = s.agg[(size_t)(13 * s.n_agg)];
    for (int64_t l = 0; l < s.n_agg; l++)
        if (s.agg[(size_t)(13 * s.n_agg + l)] < t) { t = s.agg[(size_t)(13 * s.n_agg + l)]; latest = l; }
    if (label) *label = latest;
    return E_OK;
}

int mcac_gpu_duplicate(mcac_gpu *h) {
    CK(cudaSetDevice(h->device));
    TRY(pull_scalars(h));
    TRY(duplicate(h));
    return E_OK;
}

int mcac_gpu_rand(mcac_gpu *h, int64_t n, int32_t *out) {
    CK(cudaSetDevice(h->device));
    if (h->uploaded) TRY(pull_scalars(h));
    for (int64_t done = 0; done < n;) {
        const int64_t chunk = std::min<int64_t>(n - done, kRngBuf / 2);
        TRY(ensure_rng(h, h->sc_host.rand_pos + chunk));
        CK(cudaMemcpyAsync(out + done, h->d.rng_buf + (h->sc_host.rand_pos - h->d.rng_buf_base), sizeof(int) * chunk,
                           cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->sc_host.rand_pos += chunk;
        done += chunk;
    }
    if (h->uploaded) TRY(push_scalars(h));
    return E_OK;
}

int mcac_gpu_run(mcac_gpu *h, int64_t max_steps, int32_t batch, mcac_step_record *records, int64_t n_records, mcac_run_report *report) {
    CK(cudaSetDevice(h->device));
    if (!h->uploaded) { h->err = "run before upload_state"; return E_INPUT; }
    if (h->prm.with_external_potentials && h->prm.with_potentials && !h->d.ip_ebar) {
        h->err = "mcac_gpu_run: with_external_potentials needs mcac_gpu_set_interpotential first";
        return E_INPUT;
    }
    // speculative batches need a pick sequence that is fixed between events: random pick, no per-step growth / redraws / nucleation
    const bool speculative = is_speculative(h);
    const int B = !speculative ? 1 : (batch > 0 ? std::min<int>(batch, kMaxBatch) : 256);
    TRY(pull_scalars(h));
    const Scalars at_start = h->sc_host;
    const long long launches0 = h->launches;
    if (records && n_records > 0) {
        if (h->rec_cap < n_records) {
            if (h->rec_dev) cudaFree(h->rec_dev);
            CK(cudaMalloc((void **)&h->rec_dev, sizeof(mcac_step_record) * (size_t)n_records));
            h->rec_cap = n_records;
        }
    }
    cudaEvent_t ev0, ev1;
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaEventRecord(ev0, h->stream));
    int64_t steps = 0, batches = 0, sorts = 0, dups = 0, nucleated_total = 0;
    int rc = E_OK;
    bool fin = false, need_refresh = false, fallback_sorted = false, loop_too_big = false;
    while (!speculative && steps < max_steps) {  // ---- general step: one MC step per iteration, calcul() order
        if (finished(h)) { fin = true; break; }
        const mcac_params &p = h->prm;
        const bool growth = p.with_surface_reactions != 0, pick_last = p.pick_method == MCAC_PICK_LAST;
        if (h->sc_host.event && p.with_domain_duplication && h->sc_host.n_agg <= h->dup_threshold && !(p.u_sg < 0.0)) {
            if ((rc = duplicate(h)) != E_OK) break;
            dups++;
            DBG("duplicate");
        }
        if (!loop_too_big && loop_usable(h)) {
            // the per-realization step loop: one persistent CTA walks the steps until the call is served or the realization needs
            // the host (duplication, table regrow, more staged draws) — mcac_steploop.cuh
            if ((rc = ensure_rng(h, h->sc_host.rand_pos + 8192 + 64 + 31)) != E_OK) break;
            LoopArgs la;
            loop_fill_args(h, la, max_steps - steps, (records && n_records > 0) ? h->rec_dev : nullptr, n_records, steps);
            k_step_loop<<<1, kLoopThreads, kLoopDynSmem, h->stream>>>(h->d, la);
            h->launches++;
            if (cudaGetLastError() != cudaSuccess) { h->err = "step loop launch failed"; rc = E_UNKNOWN; break; }
            if (cudaMemcpyAsync(h->loop_host, h->loop_dev, sizeof(LoopState), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) { rc = E_UNKNOWN; break; }
            if ((rc = pull_scalars(h)) != E_OK) break;
            const LoopState ls = *h->loop_host;
            loop_apply(h, ls);
            steps += ls.steps;
            batches += ls.steps;
            dups += ls.dups;
            sorts += ls.sorts;
            nucleated_total += ls.nucleated;
            if (ls.exit_reason == LOOP_ERROR || h->sc_host.error) { rc = device_error(h, h->sc_host, "mcac_gpu_run"); break; }
            if (ls.exit_reason == LOOP_FINISHED) { fin = true; break; }
            if (ls.exit_reason == LOOP_EVENT_STOP) break;
            if (ls.exit_reason == LOOP_NEED_REGROW) { if ((rc = regrow_tables(h)) != E_OK) break; }
            if (ls.exit_reason == LOOP_TOO_BIG) loop_too_big = true;
            // LOOP_NEED_DUP / LOOP_NEED_RNG / LOOP_STEPS_DONE: served at the top of the next iteration
            continue;
        }
        if (h->d.sph_cap - h->sc_host.pool_top < h->sc_host.n_sph)
            if ((rc = compact_pool(h)) != E_OK) break;
        DBG("compact_pool");
        if (p.with_nucleation && (h->d.agg_cap - h->sc_host.n_agg_slots < 64 || h->d.sph_cap - h->sc_host.pool_top < h->sc_host.n_sph + 64))
            if ((rc = regrow_tables(h)) != E_OK) break;
        if (!pick_last && (h->sc_host.event || growth || !h->pick_valid)) {
            if ((rc = event_pipeline(h, false, false, true)) != E_OK) break;
            sorts++;
            DBG("event_pipeline");
        }
        if ((rc = refresh_labels(h)) != E_OK) break;
        DBG("refresh_labels");
        DBG("nucleation regrow");
        if ((rc = ensure_rng(h, h->sc_host.rand_pos + 8192)) != E_OK) break;
        DBG("ensure_rng");
        int draws = pick_last ? 2 : 3, n_try = 1, draws_at_search = pick_last ? 2 : 3;
        if (pick_last) {
            k_pick_last<<<1, 1024, 0, h->stream>>>(h->d, h->q_slot);
            k_prepare_direction<<<1, 32, 0, h->stream>>>(h->d, h->q_slot, h->q_dir, h->q_dist, 0);
            h->launches += 2;
        } else {
            k_prepare_queries<<<1, 128, 0, h->stream>>>(h->d, 1, h->q_slot, h->q_dir, h->q_dist);
            h->launches++;
        }
        DBG("pick / direction");
        if (p.with_collisions) {
            prof_begin(h, 0);
            if ((rc = search_launch(h, 1)) != E_OK) break;
            prof_end(h);
            DBG("search");
            // orientation loop of calcul.cpp:119-141: a non-sticking contact redraws the direction (n_try++)
            while (p.with_potentials) {
                k_check_regime<<<1, 32, 0, h->stream>>>(h->d, h->q_res, h->q_dist, draws);
                h->launches++;
                if ((rc = pull_scalars(h)) != E_OK) break;
                if (h->sc_host.error) break;
                draws += h->sc_host.p_regime_draws;
                if (h->sc_host.p_regime == 0) break;
                n_try++;
                if (draws + 8 > 8192) { h->err = "orientation loop: too many redraws"; rc = E_UNKNOWN; break; }
                k_prepare_direction<<<1, 32, 0, h->stream>>>(h->d, h->q_slot, h->q_dir, h->q_dist, draws);
                draws += 2;
                draws_at_search = draws;
                h->launches++;
                if ((rc = search_launch(h, 1)) != E_OK) break;
            }
            if (rc != E_OK) break;
            if (h->sc_host.error) { rc = device_error(h, h->sc_host, "mcac_gpu_run"); break; }
        }
        StepArgs sa;
        sa.q_slot = h->q_slot; sa.q_dir = h->q_dir; sa.q_dist = h->q_dist; sa.res = h->q_res;
        sa.rec = (records && n_records > 0) ? h->rec_dev : nullptr;
        sa.rec_cap = n_records; sa.rec_index = steps;
        sa.pick_last = pick_last; sa.with_collisions = p.with_collisions; sa.n_try = n_try; sa.draws = draws; sa.draws_at_search = draws_at_search;
        prof_begin(h, 1);
        k_step_move<<<1, kCommitThreads, 0, h->stream>>>(h->d, sa);
        DBG("k_step_move");
        if (growth) k_grow_pending<<<div_up(p.individual_surf_reactions ? h->sc_host.n_sph : h->sc_host.pool_top, 256), 256, 0, h->stream>>>(h->d, p.individual_surf_reactions);
        DBG("k_grow_pending");
        k_step_merge<<<1, kCommitThreads, 0, h->stream>>>(h->d, sa.rec, sa.rec_cap, sa.rec_index);
        DBG("k_step_merge");
        prof_end(h);
        h->launches += growth ? 3 : 2;
        if (growth) {  // calcul.cpp:184-206 — the frequency test uses the counter BEFORE this step's bookkeeping
            const int full = (h->sc_host.n_iter_without_event % p.full_aggregate_update_frequency == 0) ? 1 : 0;
            if (p.individual_surf_reactions) { k_update_picked<<<1, kCommitThreads, 0, h->stream>>>(h->d, full); h->launches++; }
            k_update_small<<<div_up(h->sc_host.n_agg_slots, 128), 128, 0, h->stream>>>(h->d, full, p.individual_surf_reactions, 0);
            k_update_step<<<div_up(h->sc_host.n_agg_slots, 8), 256, 0, h->stream>>>(h->d, full, p.individual_surf_reactions);
            k_update_big<<<std::min(h->sc_host.n_agg_slots, 4 * h->n_sm), kCommitThreads, 0, h->stream>>>(h->d, full, p.individual_surf_reactions);
            h->launches += 3;
        }
        DBG("update kernels");
        if (p.with_nucleation) {  // calcul.cpp:208-220
            k_nucleate<<<1, kCommitThreads, 0, h->stream>>>(h->d, 0., 1);
            h->launches++;
        } else {
            CK(cudaMemsetAsync(&h->d.sc->n_nucleated, 0, sizeof(int), h->stream));
        }
        DBG("k_nucleate");
        k_step_event<<<1, 32, 0, h->stream>>>(h->d);
        h->launches++;
        {   // refresh() after an event, PhysicalModel::update after an event or in growth mode (calcul.cpp:232-234, 272-277)
            const int nb = std::min(1024, std::max(1, div_up(h->sc_host.n_agg_slots, kReduceThreads)));
            k_refresh_partials<<<nb, kReduceThreads, 0, h->stream>>>(h->d, h->partials);
            if (growth) k_step_totals<<<1, 32, 0, h->stream>>>(h->d, h->partials, nb);
            else k_refresh_if_event<<<1, 32, 0, h->stream>>>(h->d, h->partials, nb);
            h->launches += 2;
        }
        if (cudaGetLastError() != cudaSuccess) { h->err = "general step launch failed"; rc = E_UNKNOWN; break; }
        if ((rc = pull_scalars(h)) != E_OK) break;
        batches++;
        h->cells_valid = false;
        if (h->sc_host.error) { rc = device_error(h, h->sc_host, "mcac_gpu_run"); break; }
        steps += 1;
        if (h->sc_host.b_merged) { h->pick_valid = false; h->labels_valid = false; }
        if (h->sc_host.n_nucleated > 0) h->pick_valid = false;
        nucleated_total += h->sc_host.n_nucleated;
        if (h->stop_at_event && (h->sc_host.b_merged || h->sc_host.n_nucleated > 0)) break;
    }
    // Pipelined submission (speculative mode): batch i+1 — event kernel (it returns at once if batch i did not merge), cell rebuild,
    // queries, search, commit — is submitted BEFORE the scalars of batch i are read back, so the device never waits for the host
    // between batches.  The kernels themselves enforce what the host would have checked first (step limit, finished(), room in
    // the pool); anything unusual drains the pipeline and is handled by the one-batch-at-a-time iteration below.
    const long long steps_limit_abs = at_start.steps_done + max_steps;
    const bool pipe_mode = speculative && h->pipeline && h->overlap && h->coop_blocks > 0 &&
                           !(records && n_records > 0) && !h->stop_at_event && !h->prm.with_domain_duplication && !h->debug_sync;
    while (speculative && steps < max_steps) {
        if (finished(h)) { fin = true; break; }
        if (pipe_mode && !fallback_sorted && max_steps - steps > 2LL * B) {
            auto resources_ok = [&](int cycles) {  // from the last scalars read back: `cycles` more batches cannot run out of anything
                const Scalars &s = h->sc_host;
                return h->d.sph_cap - s.pool_top >= s.n_sph && s.rand_pos >= h->d.rng_buf_base &&
                       s.rand_pos + 3LL * B * cycles + 64 <= h->rng_generated && max_steps - steps > (long long)B * cycles;
            };
            if ((rc = ensure_rng(h, h->sc_host.rand_pos + 3LL * B * 3 + 64)) != E_OK) break;
            int inflight = 0, slot = 0, oldest = 0;
            bool spec_sorted[2] = {false, false};  // the cycle was submitted ahead (its event kernel sorts iff the previous cycle merged)
            bool anomaly = false, need_fallback_sort = false;
            Scalars last{};
            bool have_last = false;
            auto enqueue = [&]() -> int {
                const bool spec = inflight > 0;
                if (spec) {
                    CK(cudaEventRecord(h->ev_fork, h->stream));  // the rebuild on the side stream reads what the previous commit wrote
                    CK(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
                    h->cells_valid = false;
                    h->labels_valid = false;  // the batch in flight may merge: labels, refresh and totals are redone with the sort
                    prof_begin(h, 2);
                    TRY(event_pipeline(h, true, true, true, nullptr, true, true));
                    prof_end(h);
                    TRY(build_cells(h, true));
                } else if (h->sc_host.event || !h->pick_valid) {
                    prof_begin(h, 2);
                    TRY(event_pipeline(h, need_refresh, need_refresh, true, nullptr, true));
                    prof_end(h);
                    TRY(build_cells(h, true));
                    need_refresh = false;
                    sorts++;
                }
                spec_sorted[slot] = spec;
                k_prepare_queries<<<div_up(B, 128), 128, 0, h->stream>>>(h->d, B, h->q_slot, h->q_dir, h->q_dist);
                h->launches++;
                if (!h->cells_valid) {
                    prof_begin(h, 3);
                    TRY(build_cells(h));
                    prof_end(h);
                }
                prof_begin(h, 0);
                TRY(search_launch(h, B));
                prof_end(h);
                BatchArgs ba;
                ba.nq = B;
                ba.q_slot = h->q_slot; ba.q_dir = h->q_dir; ba.q_dist = h->q_dist; ba.res = h->q_res;
                ba.rec = nullptr; ba.rec_cap = 0; ba.rec_base = 0;
                ba.max_steps = max_steps;
                ba.steps_limit_abs = steps_limit_abs;
                ba.prof = h->commit_prof;
                prof_begin(h, 1);
                k_commit<<<1, kCommitThreads, 0, h->stream>>>(h->d, ba);
                prof_end(h);
                h->launches++;
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(h->h_sc_ring[slot], h->d.sc, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaEventRecord(h->ev_cycle[slot], h->stream));
                h->cells_valid = false;  // the commit moves aggregates
                slot ^= 1;
                inflight++;
                return E_OK;
            };
            auto retire = [&]() -> int {
                CK(cudaEventSynchronize(h->ev_cycle[oldest]));
                const Scalars s = *h->h_sc_ring[oldest];
                const bool was_spec = spec_sorted[oldest];
                oldest ^= 1;
                inflight--;
                if (was_spec && have_last && last.b_merged && s.b_need != 99) sorts++;  // its event kernel re-sorted after that merge
                last = s;
                have_last = true;
                h->sc_host = s;
                if (s.b_need == 99) { anomaly = true; need_fallback_sort = true; return E_OK; }
                batches++;
                if (s.error) return device_error(h, s, "mcac_gpu_run");
                steps += s.b_committed;
                if (s.b_merged) h->sc_host.avg_npp = static_cast<double>(s.n_sph) / static_cast<double>(s.n_agg);
                if (s.b_stop_reason == STOP_FINISHED) { fin = true; anomaly = true; }
                if (s.b_stop_reason == STOP_POOL || s.b_committed == 0) anomaly = true;
                if (inflight == 0) {  // host-side validity flags as of the last batch submitted
                    h->cells_valid = false;
                    if (s.b_merged) { h->pick_valid = false; h->labels_valid = false; need_refresh = true; }
                    else need_refresh = false;  // an earlier merge was refreshed and re-sorted by the event kernel that followed it
                }
                return E_OK;
            };
            if (resources_ok(1)) rc = enqueue();
            while (rc == E_OK && inflight > 0) {
                if (!anomaly && inflight < 2 && resources_ok(2) && !finished(h)) {
                    if ((rc = enqueue()) != E_OK) break;
                }
                if ((rc = retire()) != E_OK) break;
                if (anomaly) {  // drain: what is still in flight did nothing harmful (the kernels check the same conditions)
                    while (inflight > 0 && (rc = retire()) == E_OK) {}
                    break;
                }
                if (inflight == 0 && !(resources_ok(1) && max_steps - steps > 2LL * B)) break;
                if (inflight == 0 && (rc = enqueue()) != E_OK) break;
            }
            if (rc != E_OK) break;
            if (have_last) {  // host-side validity flags as of the last batch read back
                h->cells_valid = false;
                if (need_fallback_sort) {
                    h->sc_host.b_need = 0;
                    if ((rc = push_scalars(h)) != E_OK) break;
                    h->sort_fallbacks++;
                    if ((rc = sort_time_steps(h, h->sc_host.max_time_step)) != E_OK) break;
                    h->pick_valid = true;
                    fallback_sorted = true;
                }
                if (fin) break;
                if (steps >= max_steps) break;
                if (finished(h)) { fin = true; break; }
            }
        }
        if ((h->sc_host.event || !h->pick_valid) && !fallback_sorted) {
            // top of the loop after an event (calcul.cpp:72-101): duplication test, then sort_time_steps(max)
            if (h->sc_host.event && h->prm.with_domain_duplication && h->sc_host.n_agg <= h->dup_threshold && !(h->prm.u_sg < 0.0)) {
                if ((rc = duplicate(h)) != E_OK) break;
                dups++;
            }
            // the Verlet cell rebuild only reads positions: it runs on the side stream beside the event kernel
            // (the last commit was synchronised by pull_scalars, so the side stream needs no fork event; the event kernel is
            // submitted first and the rebuild's seven small launches fill in beside it)
            prof_begin(h, 2);
            if ((rc = event_pipeline(h, need_refresh, need_refresh, true, nullptr, h->overlap)) != E_OK) break;
            prof_end(h);
            if (h->overlap && (rc = build_cells(h, true)) != E_OK) break;
            need_refresh = false;
            sorts++;
        }
        if (h->d.sph_cap - h->sc_host.pool_top < h->sc_host.n_sph)
            if ((rc = compact_pool(h)) != E_OK) break;
        const int nq = (int)std::min<int64_t>(B, max_steps - steps);
        if ((rc = ensure_rng(h, h->sc_host.rand_pos + 3LL * nq)) != E_OK) break;
        k_prepare_queries<<<div_up(nq, 128), 128, 0, h->stream>>>(h->d, nq, h->q_slot, h->q_dir, h->q_dist);
        h->launches++;
        if (!h->cells_valid) {
            prof_begin(h, 3);
            if ((rc = build_cells(h)) != E_OK) break;
            prof_end(h);
        }
        prof_begin(h, 0);
        if ((rc = search_launch(h, nq)) != E_OK) break;
        prof_end(h);
        BatchArgs ba;
        ba.nq = nq;
        ba.q_slot = h->q_slot;
        ba.q_dir = h->q_dir;
        ba.q_dist = h->q_dist;
        ba.res = h->q_res;
        ba.rec = (records && n_records > 0) ? h->rec_dev : nullptr;
        ba.rec_cap = n_records;
        ba.rec_base = steps;
        ba.max_steps = max_steps - steps;
        ba.steps_limit_abs = 0;
        ba.prof = h->commit_prof;
        prof_begin(h, 1);
        k_commit<<<1, kCommitThreads, 0, h->stream>>>(h->d, ba);
        prof_end(h);
        h->launches++;
        if (cudaGetLastError() != cudaSuccess) { h->err = "k_commit launch failed"; rc = E_UNKNOWN; break; }
        if ((rc = pull_scalars(h)) != E_OK) break;
        if (h->sc_host.b_need == 99) {  // the (unsynchronised) event kernel hit introsort's depth limit: k_commit did nothing;
            h->sc_host.b_need = 0;      // redo the sort on the multi-launch path (it replays the heap-sort branch), then the batch
            if ((rc = push_scalars(h)) != E_OK) break;
            h->sort_fallbacks++;
            if ((rc = sort_time_steps(h, h->sc_host.max_time_step)) != E_OK) break;
            h->pick_valid = true;
            fallback_sorted = true;  // the event block of this step is done: go straight to the batch
            continue;
        }
        fallback_sorted = false;
        batches++;
        h->cells_valid = false;
        if (h->sc_host.error) { rc = device_error(h, h->sc_host, "mcac_gpu_run"); break; }
        steps += h->sc_host.b_committed;
        if (h->sc_host.b_merged) {  // refresh() + PhysicalModel::update + re-sort happen in the event pipeline at the loop top
            h->pick_valid = false;
            h->labels_valid = false;
            need_refresh = true;
            h->sc_host.avg_npp = static_cast<double>(h->sc_host.n_sph) / static_cast<double>(h->sc_host.n_agg);  // for finished()
        }
        if (h->sc_host.b_stop_reason == STOP_FINISHED) { fin = true; break; }
        if (h->sc_host.b_committed == 0) { h->err = "batch made no progress"; rc = E_UNKNOWN; break; }
        if (h->stop_at_event && h->sc_host.b_merged) break;
    }
    if (rc != E_OK) {  // keep the message of the call that failed
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        return rc;
    }
    if (join_cells(h) != E_OK) rc = E_UNKNOWN;
    if (rc == E_OK && need_refresh) {  // the call ended on a merge: refresh() / PhysicalModel::update belong to that step
        rc = event_pipeline(h, true, true, false);
        need_refresh = false;
    }
    CK(cudaEventRecord(ev1, h->stream));
    CK(cudaEventSynchronize(ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (rc != E_OK) return rc;
    if (records && n_records > 0 && steps > 0)
        CK(cudaMemcpy(records, h->rec_dev, sizeof(mcac_step_record) * (size_t)std::min<int64_t>(steps, n_records), cudaMemcpyDeviceToHost));
    if (report) {
        const Scalars &sc = h->sc_host;
        std::memset(report, 0, sizeof(*report));
        report->steps = steps;
        report->events = sc.total_events - at_start.total_events;
        report->searches = sc.searches - at_start.searches;
        report->pair_tests_sphere = sc.pair_sphere - at_start.pair_sphere;
        report->pair_tests_bounding = sc.pair_bounding - at_start.pair_bounding;
        report->pair_tests_executed = sc.pair_exec - at_start.pair_exec;
        for (int k = 0; k < 8; k++) { report->loop_phase_cycles[k] = h->loop_cycles[k] - h->loop_cycles_seen[k]; h->loop_cycles_seen[k] = h->loop_cycles[k]; }
        report->batches = batches;
        report->conflicts = sc.conflicts - at_start.conflicts;
        report->duplications = dups;
        report->sorts = sorts;
        report->kernel_launches = h->launches - launches0;
        report->n_aggregates = sc.n_agg;
        report->n_spheres = sc.n_sph;
        report->finished = (fin || finished(h)) ? 1 : 0;
        report->time = sc.time;
        report->box_length = sc.box_length;
        report->avg_npp = sc.avg_npp;
        report->max_time_step = sc.max_time_step;
        report->volume_fraction = sc.volume_fraction;
        report->device_ms = ms;
        report->n_iter_without_event = sc.n_iter_without_event;
        report->nucleated = nucleated_total;
        report->total_volume = sc.total_volume;
        report->total_surface = sc.total_surface;
        long long w[40] = {0};
        cudaMemcpy(w, h->event_work, sizeof(w), cudaMemcpyDeviceToHost);
        report->sort_span_elements = w[0] - h->event_work_seen[0];
        report->sort_levels = w[1] - h->event_work_seen[1];
        for (int k = 0; k < 8; k++) report->event_phase_cycles[k] = w[2 + k] - h->event_work_seen[2 + k];
        for (int k = 0; k < 2; k++) report->tie_phase_cycles[k] = w[10 + k] - h->event_work_seen[10 + k];
        report->tie_sorts = w[12] - h->event_work_seen[12];
        report->tie_levels = w[13] - h->event_work_seen[13];
        report->tie_sparse = w[14] - h->event_work_seen[14];
        report->tie_handed = w[15] - h->event_work_seen[15];
        for (int k = 0; k < 3; k++) {
            report->tie_sim_cycles[k] = w[16 + k] - h->event_work_seen[16 + k];
            report->tie_phase_cycles[0] += report->tie_sim_cycles[k];
        }
        if (h->commit_prof) {
            long long c[16];
            cudaMemcpy(c, h->commit_prof, sizeof(c), cudaMemcpyDeviceToHost);
            cudaMemset(h->commit_prof, 0, sizeof(c));
            const double nl = std::max(1.0, (double)c[15]);
            fprintf(stderr, "k_commit cycles per launch (%lld launches): stage %.0f, stop conditions %.0f, conflicts %.0f, free-flight moves %.0f, contact move %.0f, merge + update %.0f, tail %.0f\n",
                    c[15], c[0] / nl, c[1] / nl, c[2] / nl, c[3] / nl, c[4] / nl, c[5] / nl, c[6] / nl);
        }
        if (getenv("MCAC_B200_K9_DEBUG") && w[29] > h->event_work_seen[29]) {
            const double nc = (double)(w[29] - h->event_work_seen[29]);
            fprintf(stderr, "k9 exact cumulative tables: %.0f, %.0f cycles of the building CTA and %.1f segments of the W run per table; head by integer prefix sums in %.0f\n", nc,
                    (w[28] - h->event_work_seen[28]) / nc, (w[30] - h->event_work_seen[30]) / nc, (double)(w[31] - h->event_work_seen[31]));
            fprintf(stderr, "k9 building CTA cycles per table: gather %.0f, sort %.0f (redone by the one-step network: %.0f tables), head %.0f, W run + write-out %.0f\n",
                    (w[32] - h->event_work_seen[32]) / nc, (w[33] - h->event_work_seen[33]) / nc, (double)(w[36] - h->event_work_seen[36]),
                    (w[34] - h->event_work_seen[34]) / nc, (w[35] - h->event_work_seen[35]) / nc);
        }
        if (getenv("MCAC_B200_K9_DEBUG") && w[21] > h->event_work_seen[21]) {
            const double nw = (double)(w[21] - h->event_work_seen[21]), ns = std::max(1.0, (double)(w[12] - h->event_work_seen[12]));
            fprintf(stderr, "k9 windows: %.1f per sort, mean %.0f cycles and %.2f levels per window; block 0 waits %.0f cycles per sort behind its own levels\n",
                    nw / ns, (w[19] - h->event_work_seen[19]) / nw, (w[20] - h->event_work_seen[20]) / nw, (w[22] - h->event_work_seen[22]) / ns);
            fprintf(stderr, "k9 window cycles per window: top+staging %.0f, flags+scan %.0f, scan tail %.0f, swaps %.0f, split %.0f\n",
                    (w[23] - h->event_work_seen[23]) / nw, (w[24] - h->event_work_seen[24]) / nw, (w[25] - h->event_work_seen[25]) / nw,
                    (w[26] - h->event_work_seen[26]) / nw, (w[27] - h->event_work_seen[27]) / nw);
        }
        for (int k = 0; k < 40; k++) h->event_work_seen[k] = w[k];
        report->sort_fallbacks = h->sort_fallbacks - h->sort_fallbacks_seen;
        report->sort_heap_branches = h->sort_heap_levels - h->sort_heap_seen;
        h->sort_fallbacks_seen = h->sort_fallbacks;
        h->sort_heap_seen = h->sort_heap_levels;
    }
    prof_collect(h, report);
    return E_OK;
}

int mcac_gpu_set_profile(mcac_gpu *h, int32_t profile) { h->profile = profile; return E_OK; }
int mcac_gpu_reserve(mcac_gpu *h, int64_t n_spheres, int64_t n_aggregates) {
    if (n_spheres < 0 || n_aggregates < 0) { h->err = "reserve: negative size"; return E_INPUT; }
    h->reserve_sph = n_spheres;
    h->reserve_agg = n_aggregates;
    return E_OK;
}
int mcac_gpu_set_strict_direction(mcac_gpu *h, int32_t on) {
    h->strict_dir = on != 0;
    h->dir_tab_valid = false;
    if (!h->strict_dir) h->d.dir_tab = nullptr;
    return E_OK;
}
int mcac_gpu_set_stop_at_event(mcac_gpu *h, int32_t on) { h->stop_at_event = on != 0; return E_OK; }

int mcac_gpu_morphology_stats_device(mcac_gpu *h, int32_t n_bins, double rg_max, void *device_out) {
    CK(cudaSetDevice(h->device));
    if (n_bins < 1 || !device_out) { h->err = "morphology_stats: bad arguments"; return E_INPUT; }
    if (2 * (size_t)n_bins + 1 > kStatsHistCap) { h->err = "morphology_stats: too many bins"; return E_INPUT; }
    TRY(pull_scalars(h));
    CK(cudaMemsetAsync(h->stats_hist, 0, sizeof(unsigned int) * (2 * (size_t)n_bins + 1), h->stream));
    const int nb = std::min(kStatsMaxBlocks, std::max(1, div_up(h->sc_host.n_agg_slots, 256)));
    k_morphology_stats<<<nb, 256, 0, h->stream>>>(h->d, n_bins, rg_max, (double *)device_out, h->stats_part, h->stats_hist);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return E_OK;
}
int mcac_gpu_morphology_stats(mcac_gpu *h, int32_t n_bins, double rg_max, double *out) {
    if (2 * (size_t)n_bins + 8 > 4096) { h->err = "morphology_stats: too many bins"; return E_INPUT; }
    TRY(mcac_gpu_morphology_stats_device(h, n_bins, rg_max, h->stats_dev));
    CK(cudaMemcpy(out, h->stats_dev, sizeof(double) * (2 * (size_t)n_bins + 8), cudaMemcpyDeviceToHost));
    return E_OK;
}

int mcac_gpu_search_sweep(mcac_gpu *h, int64_t n, int32_t repeats, mcac_sweep_report *report) {
    CK(cudaSetDevice(h->device));
    if (!h->uploaded || n < 1) { h->err = "search_sweep: no state"; return E_INPUT; }
    if (3 * n > kRngBuf - 64) n = (kRngBuf - 64) / 3;
    TRY(pull_scalars(h));
    if (!h->pick_valid) TRY(sort_time_steps(h, h->sc_host.max_time_step));
    if (h->sweep_cap < n) {
        for (void *p : {(void *)h->sweep_slot, (void *)h->sweep_dir, (void *)h->sweep_dist, (void *)h->sweep_res}) if (p) cudaFree(p);
        CK(cudaMalloc((void **)&h->sweep_slot, sizeof(int) * n));
        CK(cudaMalloc((void **)&h->sweep_dir, sizeof(double) * 3 * n));
        CK(cudaMalloc((void **)&h->sweep_dist, sizeof(double) * n));
        CK(cudaMalloc((void **)&h->sweep_res, sizeof(SearchResult) * n));
        h->sweep_cap = n;
    }
    TRY(ensure_rng(h, h->sc_host.rand_pos + 3 * n));
    TRY(build_cells(h));
    TRY(join_cells(h));
    k_prepare_queries<<<div_up(n, 128), 128, 0, h->stream>>>(h->d, (int)n, h->sweep_slot, h->sweep_dir, h->sweep_dist);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int reps = repeats > 0 ? repeats : 1;
    TRY(search_kernels(h, (int)n, h->sweep_slot, h->sweep_dir, h->sweep_dist, h->sweep_res));  // warm-up
    CK(cudaEventRecord(e0, h->stream));
    for (int r = 0; r < reps; r++) TRY(search_kernels(h, (int)n, h->sweep_slot, h->sweep_dir, h->sweep_dist, h->sweep_res));
    CK(cudaEventRecord(e1, h->stream));
    h->launches += 1;
    CK(cudaMemsetAsync(h->stats_dev, 0, sizeof(double) * 4, h->stream));
    k_sweep_summary<<<std::min(1024, div_up(n, 256)), 256, 0, h->stream>>>(h->sweep_res, h->sweep_dist, (int)n, h->stats_dev);
    double out[4];
    CK(cudaMemcpyAsync(out, h->stats_dev, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (report) {
        report->n_queries = n;
        report->contacts = (int64_t)out[0];
        report->distance_checksum = out[1];
        report->pair_tests_sphere = (int64_t)out[2];
        report->pair_tests_bounding = (int64_t)out[3];
        report->kernel_ms = ms / reps;
    }
    return E_OK;
}

// Per-kernel timings on the resident state (bench.py's per-kernel roofline table; profiles/).  `which`:
//  0 K2 cell list rebuild, 1 K8 growth of every sphere (dt = 0), 2 K6/K7 update_partial of every aggregate, 3 K5-K7 full update,
//  4 K9 event pipeline with sort, 5 event pipeline without sort (labels + refresh + totals), 6 100 grid barriers at K9's launch shape,
//  7 K10 RNG fill (kRngBuf draws), 8 K11 morphology statistics, 9 / 10 FP64 pipe peak (DFMA / DMUL+DADD; units = flops),
//  11 the sparse simulation of the tie-dominated sort alone (tuning probe).
// Growth / update rewrite derived fields from the resident radii (a replayed trajectory should not continue from this state).
int mcac_gpu_kernel_bench(mcac_gpu *h, int32_t which, int32_t reps, double *ms_out, int64_t *units_out) {
    CK(cudaSetDevice(h->device));
    if (!h->uploaded || reps < 1 || !ms_out) { h->err = "kernel_bench: bad arguments / no state"; return E_INPUT; }
    TRY(pull_scalars(h));
    TRY(refresh_labels(h));
    const Scalars sc = h->sc_host;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int64_t units = 0;
    int rc = E_OK;
    for (int r = -1; r < reps && rc == E_OK; r++) {  // r == -1: warm-up
        if (r == 0) CK(cudaEventRecord(e0, h->stream));
        switch (which) {
        case 0: h->cells_valid = false; rc = build_cells(h); units = sc.n_agg; break;
        case 1: k_grow<<<div_up(sc.pool_top, 256), 256, 0, h->stream>>>(h->d, 0., -1); units = sc.n_sph; break;
        case 2:
        case 3:
            k_update_small<<<div_up(sc.n_agg_slots, 128), 128, 0, h->stream>>>(h->d, which == 3, 0, 1);
            k_update_all<<<div_up(sc.n_agg_slots, 8), 256, 0, h->stream>>>(h->d, which == 3, -1);
            k_update_big<<<std::min(sc.n_agg_slots, 4 * h->n_sm), kCommitThreads, 0, h->stream>>>(h->d, which == 3, 0);
            units = sc.n_agg;
            break;
        case 4: h->labels_valid = false; rc = event_pipeline(h, true, true, true); units = sc.n_agg; break;
        case 5: h->labels_valid = false; rc = event_pipeline(h, true, true, false); units = sc.n_agg; break;
        case 6: {
            if (h->coop_blocks <= 0) { h->err = "kernel_bench: no cooperative launch"; rc = E_INPUT; break; }
            int n = 100;
            int *sink = nullptr;
            void *args[] = {&n, &sink};
            if (cudaLaunchCooperativeKernel((const void *)k_barrier_probe, dim3(h->coop_blocks), dim3(kEventThreads), args, 0, h->stream) != cudaSuccess) rc = E_UNKNOWN;
            units = n;
            break;
        }
        case 7: {  // draws go to the staging buffer and the stream state is put back afterwards: the realization's stream is untouched
            if ((rc = ensure_stage(h, sizeof(int) * (size_t)kRngBuf)) != E_OK) break;
            cudaMemcpyAsync(h->stats_dev, h->d.rng, sizeof(GlibcRandState), cudaMemcpyDeviceToDevice, h->stream);
            k_rng_fill<<<1, 32, 0, h->stream>>>(h->d.rng, (int *)h->stage, kRngBuf);
            cudaMemcpyAsync(h->d.rng, h->stats_dev, sizeof(GlibcRandState), cudaMemcpyDeviceToDevice, h->stream);
            units = kRngBuf;
            break;
        }
        case 8:
            cudaMemsetAsync(h->stats_hist, 0, sizeof(unsigned int) * (2 * 24 + 1), h->stream);
            k_morphology_stats<<<std::min(kStatsMaxBlocks, std::max(1, div_up(sc.n_agg_slots, 256))), 256, 0, h->stream>>>(h->d, 24, 2e-6, h->stats_dev, h->stats_part, h->stats_hist);
            units = sc.n_agg;
            break;
        case 9:
        case 10: {  // FP64 pipe peaks: 9 = DFMA, 10 = DMUL + DADD (this library is built --fmad=false); units = flops per launch
            const int iters = 4096, blocks = h->n_sm * 8;
            if (which == 9) k_fp64_peak<0><<<blocks, 256, 0, h->stream>>>(h->stats_dev, iters, 1.0);
            else k_fp64_peak<1><<<blocks, 256, 0, h->stream>>>(h->stats_dev, iters, 1.0);
            units = (int64_t)blocks * 256 * 8 * 2 * iters;
            break;
        }
        case 11: {  // tuning probe: the sparse simulation alone (MCAC_B200_PROBE_X sparse elements among n_agg, synthetic positions)
            if (!h->ts_plan || h->event_dyn_bytes == 0) { h->err = "kernel_bench: tie-sort path off"; rc = E_INPUT; break; }
            int x = 3915;
            if (const char *e = getenv("MCAC_B200_PROBE_X")) x = std::max(1, std::min(atoi(e), h->ts_xcap));
            const int n = (int)sc.n_agg;
            if (r == -1) {
                std::vector<int> pos(x);
                std::vector<double> w(x);
                unsigned long long lcg = 88172645463325252ULL;
                auto rnd = [&]() { lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(lcg >> 11) / 9007199254740992.0; };
                const double stride = (double)n / x;
                for (int j = 0; j < x; j++) {
                    pos[j] = std::min(n - 1, (int)(j * stride + rnd() * std::max(1.0, stride - 1.0)));
                    if (j > 0 && pos[j] <= pos[j - 1]) pos[j] = pos[j - 1] + 1;
                    w[j] = 0.2 + 0.7 * rnd();
                }
                if ((rc = ensure_stage(h, (sizeof(int) + sizeof(double)) * (size_t)x + 64)) != E_OK) break;
                CK(cudaMemcpyAsync(h->stage, w.data(), sizeof(double) * x, cudaMemcpyHostToDevice, h->stream));
                CK(cudaMemcpyAsync((char *)h->stage + sizeof(double) * x, pos.data(), sizeof(int) * x, cudaMemcpyHostToDevice, h->stream));
                CK(cudaMemsetAsync(h->event_work, 0, 32 * sizeof(long long), h->stream));
                CK(cudaStreamSynchronize(h->stream));
                cudaFuncSetAttribute(k_plan_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->event_dyn_bytes);
            }
            int lg = 0;
            while ((1LL << (lg + 1)) <= n) lg++;
            k_plan_probe<<<1, kEventThreads, h->event_dyn_bytes, h->stream>>>(n, x, (const int *)((char *)h->stage + sizeof(double) * x), (const double *)h->stage, 1.0,
                                                                             2 * lg, h->sort_local_span, h->ts_plan, h->ts_R, h->ts_tbl, h->ts_xcap,
                                                                             (int)h->event_dyn_bytes, h->event_work);
            units = x;
            if (r == reps - 1) {
                long long cyc[12];
                CK(cudaMemcpyAsync(cyc, h->event_work, sizeof(cyc), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
                fprintf(stderr, "plan_probe: n %d x %d cycles/launch %.0f levels %.2f handed %.0f | per launch: pivot %.0f pivot-move %.0f table %.0f barrier %.0f aK %.0f moves %.0f end %.0f\n",
                        n, x, (double)cyc[0] / (reps + 1), (double)cyc[1] / (reps + 1), (double)cyc[2] / (reps + 1), (double)cyc[4] / (reps + 1),
                        (double)cyc[5] / (reps + 1), (double)cyc[6] / (reps + 1), (double)cyc[7] / (reps + 1), (double)cyc[8] / (reps + 1),
                        (double)cyc[9] / (reps + 1), (double)cyc[10] / (reps + 1));
                CK(cudaMemsetAsync(h->event_work, 0, 32 * sizeof(long long), h->stream));
            }
            break;
        }
        default: h->err = "kernel_bench: unknown kernel"; rc = E_INPUT;
        }
    }
    if (rc == E_OK) {
        CK(cudaEventRecord(e1, h->stream));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        *ms_out = ms / reps;
        if (units_out) *units_out = units;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    h->cells_valid = false;
    h->pick_valid = false;
    return rc;
}

// Ensemble of independent realizations (the path does not shard, replicas do).
//
// Realizations whose steps run in the per-realization step loop (general step: growth / potentials / nucleation / pick_last; the
// ensemble of examples/classic.ini) are advanced by ONE launch of k_ensemble_loop per round: the CTAs of the grid take realizations
// from a queue and walk their MC steps without the host, so the device holds as many realizations in flight as it has CTA slots.  A
// round ends when every realization has either done its steps or needs the host (domain duplication, table regrow, more staged
// draws); `threads` host threads serve those in parallel (each handle has its own stream), then the next round starts.
// Other realizations (speculative batches of the collision-only configurations, or MCAC_B200_NO_LOOP) are driven by the host threads,
// handle k on thread k mod threads.  No data is shared between handles; the first non-zero error code of any realization is returned.
static void fill_report_basic(mcac_gpu *h, const Scalars &at_start, long long launches0, mcac_run_report *report, int64_t steps, int64_t dups,
                              int64_t sorts, int64_t nucleated, bool fin) {
    const Scalars &sc = h->sc_host;
    std::memset(report, 0, sizeof(*report));
    report->steps = steps;
    report->events = sc.total_events - at_start.total_events;
    report->searches = sc.searches - at_start.searches;
    report->pair_tests_sphere = sc.pair_sphere - at_start.pair_sphere;
    report->pair_tests_bounding = sc.pair_bounding - at_start.pair_bounding;
    report->pair_tests_executed = sc.pair_exec - at_start.pair_exec;
    for (int k = 0; k < 8; k++) { report->loop_phase_cycles[k] = h->loop_cycles[k] - h->loop_cycles_seen[k]; h->loop_cycles_seen[k] = h->loop_cycles[k]; }
    report->batches = steps;
    report->duplications = dups;
    report->sorts = sorts;
    report->kernel_launches = h->launches - launches0;
    report->n_aggregates = sc.n_agg;
    report->n_spheres = sc.n_sph;
    report->finished = (fin || finished(h)) ? 1 : 0;
    report->time = sc.time;
    report->box_length = sc.box_length;
    report->avg_npp = sc.avg_npp;
    report->max_time_step = sc.max_time_step;
    report->volume_fraction = sc.volume_fraction;
    report->n_iter_without_event = sc.n_iter_without_event;
    report->nucleated = nucleated;
    report->total_volume = sc.total_volume;
    report->total_surface = sc.total_surface;
}

int mcac_ensemble_run(mcac_gpu **handles, int32_t n, int64_t max_steps, int32_t batch, int32_t threads, mcac_run_report *reports) {
    if (!handles || n < 1) return E_INPUT;
    const auto t_call = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    double service_ms = 0., launch_ms = 0., apply_ms = 0.;
    std::atomic<long long> n_dup_services{0}, n_regrow_services{0}, n_rng_services{0};
    const int T = std::max(1, std::min<int>(threads, n));
    std::vector<int> rcs((size_t)n, E_OK);
    auto parallel_for = [&](const std::vector<int> &items, auto &&fn) {
        const int W = std::max(1, std::min<int>(T, (int)items.size()));
        auto work = [&](int t) { for (size_t i = (size_t)t; i < items.size(); i += (size_t)W) fn(items[i]); };
        std::vector<std::thread> pool;
        for (int t = 1; t < W; t++) pool.emplace_back(work, t);
        work(0);
        for (auto &th : pool) th.join();
    };
    // ---- realizations that take the step loop, all on one device
    std::vector<int> loop_set, host_set;
    for (int k = 0; k < n; k++) {
        mcac_gpu *h = handles[k];
        const bool ok = h && h->uploaded && !is_speculative(h) && h->fused && !h->debug_sync && !h->profile && !h->stop_at_event &&
                        h->device == handles[0]->device &&
                        !(h->prm.with_external_potentials && h->prm.with_potentials && !h->d.ip_ebar);
        (ok ? loop_set : host_set).push_back(k);
    }
    if (!host_set.empty())
        parallel_for(host_set, [&](int k) { rcs[(size_t)k] = mcac_gpu_run(handles[k], max_steps, batch, nullptr, 0, reports ? reports + k : nullptr); });
    if (!loop_set.empty()) {
        const int m = (int)loop_set.size();
        cudaSetDevice(handles[loop_set[0]]->device);
        struct Track { Scalars at_start; long long launches0; int64_t steps = 0, dups = 0, sorts = 0, nucleated = 0; bool fin = false, host_only = false; };
        std::vector<Track> tr((size_t)m);
        // (the Scalars of all realizations are fetched below with one gather kernel + one copy, once the scratch arrays exist)
        cudaStream_t es = nullptr;
        DevState *ds_dev = nullptr;
        LoopArgs *as_dev = nullptr;
        int *next_dev = nullptr;
        Scalars *sc_all_dev = nullptr;
        LoopState *ls_all_dev = nullptr;
        std::vector<DevState> ds_host((size_t)m);
        std::vector<LoopArgs> as_host((size_t)m);
        std::vector<Scalars> sc_all_host((size_t)m);
        std::vector<LoopState> ls_all_host((size_t)m);
        int rc_all = E_OK;
        // scratch of the call (argument arrays, queue counter, read-back arrays, stream, events): kept per device between calls and only
        // ever grown — cudaMalloc / cudaFree / stream creation per call are device-wide synchronising driver calls whose cost was seen
        // to reach 200 ms per call on some boxes with 1024 realizations resident
        struct EnsScratch { int cap = 0; cudaStream_t es = nullptr; cudaEvent_t ev_a = nullptr, ev_b = nullptr; DevState *ds = nullptr; LoopArgs *as = nullptr;
                            int *next = nullptr; Scalars *sc = nullptr; LoopState *ls = nullptr; };
        static EnsScratch scratch_of[64];
        static std::mutex scratch_mu;
        std::unique_lock<std::mutex> scratch_lock(scratch_mu);  // (one ensemble call at a time per process: they share the scratch)
        EnsScratch &S = scratch_of[handles[loop_set[0]]->device & 63];
        if (!S.es && (cudaStreamCreateWithFlags(&S.es, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&S.ev_a) != cudaSuccess ||
                      cudaEventCreate(&S.ev_b) != cudaSuccess || cudaMalloc((void **)&S.next, sizeof(int)) != cudaSuccess))
            rc_all = E_UNKNOWN;
        if (rc_all == E_OK && S.cap < m) {
            if (S.ds) { cudaFree(S.ds); cudaFree(S.as); cudaFree(S.sc); cudaFree(S.ls); S.ds = nullptr; S.as = nullptr; S.sc = nullptr; S.ls = nullptr; S.cap = 0; }
            if (cudaMalloc((void **)&S.ds, sizeof(DevState) * (size_t)m) != cudaSuccess || cudaMalloc((void **)&S.as, sizeof(LoopArgs) * (size_t)m) != cudaSuccess ||
                cudaMalloc((void **)&S.sc, sizeof(Scalars) * (size_t)m) != cudaSuccess || cudaMalloc((void **)&S.ls, sizeof(LoopState) * (size_t)m) != cudaSuccess)
                rc_all = E_UNKNOWN;
            else S.cap = m;
        }
        es = S.es; ds_dev = S.ds; as_dev = S.as; next_dev = S.next; sc_all_dev = S.sc; ls_all_dev = S.ls;
        if (rc_all == E_OK) {  // current Scalars of every realization: one gather + one copy (every handle's stream is idle between calls)
            std::vector<Scalars *> ptrs((size_t)m);
            for (int i = 0; i < m; i++) ptrs[(size_t)i] = handles[loop_set[(size_t)i]]->d.sc;
            Scalars **ptrs_dev = reinterpret_cast<Scalars **>(as_dev);  // (scratch: LoopArgs is larger than a pointer)
            bool ok = cudaDeviceSynchronize() == cudaSuccess &&
                      cudaMemcpyAsync(ptrs_dev, ptrs.data(), sizeof(Scalars *) * (size_t)m, cudaMemcpyHostToDevice, es) == cudaSuccess;
            if (ok) {
                k_gather_scalars<<<std::max(1, std::min(m, 296)), 256, 0, es>>>(ptrs_dev, m, sc_all_dev);
                ok = cudaGetLastError() == cudaSuccess &&
                     cudaMemcpyAsync(sc_all_host.data(), sc_all_dev, sizeof(Scalars) * (size_t)m, cudaMemcpyDeviceToHost, es) == cudaSuccess &&
                     cudaStreamSynchronize(es) == cudaSuccess;
            }
            if (!ok) rc_all = E_UNKNOWN;
            for (int i = 0; i < m && ok; i++) {
                mcac_gpu *h = handles[loop_set[(size_t)i]];
                h->sc_host = sc_all_host[(size_t)i];
                h->loop_host->exit_reason = LOOP_STEPS_DONE;
                tr[(size_t)i].at_start = h->sc_host;
                tr[(size_t)i].launches0 = h->launches;
            }
        }
        const double prologue_ms = since(t_call);
        int occ = 1, n_sm = handles[loop_set[0]]->n_sm;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ensemble_loop, kLoopThreads, kLoopDynSmem);
        cudaEvent_t ev_a = S.ev_a, ev_b = S.ev_b;
        double kernel_ms = 0.;
        long long rounds = 0;
        while (rc_all == E_OK) {
            // ---- host services of the round: what the loop asked for, or what calcul() does at the loop top.  Most rounds need none
            // (the loop duplicates, compacts and re-sorts by itself): the handles are scanned here, and only those that need the host
            // are served, in parallel
            const auto t_service = std::chrono::steady_clock::now();
            std::vector<int> active, need;  // indices into loop_set
            for (int i = 0; i < m; i++) {
                const int k = loop_set[(size_t)i];
                if (rcs[(size_t)k] != E_OK || tr[(size_t)i].fin || tr[(size_t)i].steps >= max_steps) continue;
                mcac_gpu *h = handles[k];
                if (finished(h)) { tr[(size_t)i].fin = true; continue; }
                active.push_back(i);
                const mcac_params &p = h->prm;
                const bool dup = h->sc_host.event && p.with_domain_duplication && h->sc_host.n_agg <= h->dup_threshold && !(p.u_sg < 0.0) &&
                                 h->loop_host->exit_reason == LOOP_NEED_DUP;  // (otherwise the loop does it, if its tables have room)
                const bool regrow = p.with_nucleation &&
                                    (h->d.agg_cap - h->sc_host.n_agg_slots < 64 || h->d.sph_cap - h->sc_host.pool_top < h->sc_host.n_sph + 64);
                const bool rng = !(h->sc_host.rand_pos + 8192 + 64 + 31 <= h->rng_generated && h->sc_host.rand_pos >= h->d.rng_buf_base) ||
                                 (h->strict_dir && !h->dir_tab_valid);
                tr[(size_t)i].host_only = !loop_usable(h);
                if (dup || regrow || rng) need.push_back(i);
            }
            if (active.empty()) break;
            if (!need.empty())
                parallel_for(need, [&](int i) {
                    const int k = loop_set[(size_t)i];
                    mcac_gpu *h = handles[k];
                    Track &t = tr[(size_t)i];
                    cudaSetDevice(h->device);
                    int rc = E_OK;
                    const mcac_params &p = h->prm;
                    if (h->sc_host.event && p.with_domain_duplication && h->sc_host.n_agg <= h->dup_threshold && !(p.u_sg < 0.0) &&
                        h->loop_host->exit_reason == LOOP_NEED_DUP) {
                        rc = duplicate(h);
                        t.dups++;
                        n_dup_services++;
                        h->loop_host->exit_reason = LOOP_STEPS_DONE;
                    }
                    if (rc == E_OK && p.with_nucleation &&
                        (h->d.agg_cap - h->sc_host.n_agg_slots < 64 || h->d.sph_cap - h->sc_host.pool_top < h->sc_host.n_sph + 64)) {
                        rc = regrow_tables(h);
                        n_regrow_services++;
                    }
                    if (rc == E_OK) {
                        const long long before = h->rng_generated;
                        rc = ensure_rng(h, h->sc_host.rand_pos + 8192 + 64 + 31);
                        if (h->rng_generated != before) n_rng_services++;
                    }
                    if (rc == E_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = E_UNKNOWN;
                    t.host_only = !loop_usable(h);
                    rcs[(size_t)k] = rc;
                });
            // realizations that outgrew the loop finish on the host-driven path
            std::vector<int> big, run;
            for (int i : active) {
                const int k = loop_set[(size_t)i];
                if (rcs[(size_t)k] != E_OK || tr[(size_t)i].fin) continue;
                (tr[(size_t)i].host_only ? big : run).push_back(i);
            }
            if (!big.empty())
                parallel_for(big, [&](int i) {
                    const int k = loop_set[(size_t)i];
                    mcac_run_report rep{};
                    rcs[(size_t)k] = mcac_gpu_run(handles[k], max_steps - tr[(size_t)i].steps, batch, nullptr, 0, &rep);
                    tr[(size_t)i].steps += rep.steps; tr[(size_t)i].dups += rep.duplications; tr[(size_t)i].sorts += rep.sorts;
                    tr[(size_t)i].nucleated += rep.nucleated;
                    if (rep.finished || rep.steps == 0) tr[(size_t)i].fin = true;
                });
            service_ms += since(t_service);
            if (run.empty()) continue;
            const auto t_launch = std::chrono::steady_clock::now();
            // ---- one launch for the whole round
            const int nr = (int)run.size();
            // queue order: longest expected first (cycles per MC step of the realization's last launch x the steps it still has to
            // do), so that the CTAs that finish early are not left waiting for a long realization taken last
            {
                auto cost = [&](int i) {
                    const mcac_gpu *h = handles[loop_set[(size_t)i]];
                    const double per_step = h->loop_cost_per_step > 0. ? h->loop_cost_per_step : (double)h->sc_host.n_sph;
                    return per_step * (double)(max_steps - tr[(size_t)i].steps);
                };
                std::stable_sort(run.begin(), run.end(), [&](int x, int y) { return cost(x) > cost(y); });
            }
            for (int j = 0; j < nr; j++) {
                mcac_gpu *h = handles[loop_set[(size_t)run[(size_t)j]]];
                ds_host[(size_t)j] = h->d;
                loop_fill_args(h, as_host[(size_t)j], max_steps - tr[(size_t)run[(size_t)j]].steps, nullptr, 0, 0);
            }
            const int grid = std::max(1, std::min(nr, n_sm * std::max(1, occ)));
            if (cudaMemcpyAsync(ds_dev, ds_host.data(), sizeof(DevState) * (size_t)nr, cudaMemcpyHostToDevice, es) != cudaSuccess ||
                cudaMemcpyAsync(as_dev, as_host.data(), sizeof(LoopArgs) * (size_t)nr, cudaMemcpyHostToDevice, es) != cudaSuccess ||
                cudaMemsetAsync(next_dev, 0, sizeof(int), es) != cudaSuccess) { rc_all = E_UNKNOWN; break; }
            cudaEventRecord(ev_a, es);
            k_ensemble_loop<<<grid, kLoopThreads, kLoopDynSmem, es>>>(ds_dev, as_dev, nr, next_dev, sc_all_dev, ls_all_dev);
            cudaEventRecord(ev_b, es);
            rounds++;
            // the whole round comes back with two copies (every realization left its Scalars / LoopState in the contiguous arrays)
            bool ok = cudaGetLastError() == cudaSuccess &&
                      cudaMemcpyAsync(sc_all_host.data(), sc_all_dev, sizeof(Scalars) * (size_t)nr, cudaMemcpyDeviceToHost, es) == cudaSuccess &&
                      cudaMemcpyAsync(ls_all_host.data(), ls_all_dev, sizeof(LoopState) * (size_t)nr, cudaMemcpyDeviceToHost, es) == cudaSuccess &&
                      cudaStreamSynchronize(es) == cudaSuccess;
            if (!ok) {
                for (int j = 0; j < nr; j++) {
                    mcac_gpu *h = handles[loop_set[(size_t)run[(size_t)j]]];
                    h->err = std::string("ensemble step loop: ") + cudaGetErrorString(cudaGetLastError());
                }
                rc_all = E_UNKNOWN;
                break;
            }
            {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, ev_a, ev_b) == cudaSuccess) kernel_ms += ms;
            }
            launch_ms += since(t_launch);
            const auto t_apply = std::chrono::steady_clock::now();
            for (int j = 0; j < nr; j++) {
                const int i = run[(size_t)j], k = loop_set[(size_t)i];
                mcac_gpu *h = handles[k];
                Track &t = tr[(size_t)i];
                h->launches++;
                h->sc_host = sc_all_host[(size_t)j];
                const LoopState ls = ls_all_host[(size_t)j];
                *h->loop_host = ls;
                if (ls.steps > 0) {
                    long long cyc = 0;
                    for (int q = 0; q < 8; q++) cyc += ls.phase_cycles[q];
                    h->loop_cost_per_step = (double)cyc / (double)ls.steps;
                }
                loop_apply(h, ls);
                t.steps += ls.steps; t.sorts += ls.sorts; t.nucleated += ls.nucleated; t.dups += ls.dups;
                if (ls.exit_reason == LOOP_ERROR || h->sc_host.error) { rcs[(size_t)k] = device_error(h, h->sc_host, "mcac_ensemble_run"); continue; }
                if (ls.exit_reason == LOOP_FINISHED) t.fin = true;
                if (ls.exit_reason == LOOP_STEPS_DONE && ls.steps == 0 && t.steps < max_steps) { h->err = "step loop made no progress"; rcs[(size_t)k] = E_UNKNOWN; }
            }
            apply_ms += since(t_apply);
            if (getenv("MCAC_B200_K9_DEBUG")) {  // the round's three most expensive realizations (the queue takes them first)
                std::vector<std::pair<long long, int>> cost;
                for (int j = 0; j < nr; j++) {
                    long long cyc = 0;
                    for (int q = 0; q < 8; q++) cyc += ls_all_host[(size_t)j].phase_cycles[q];
                    cost.push_back({cyc, j});
                }
                std::sort(cost.begin(), cost.end());
                long long total = 0;
                for (auto &c : cost) total += c.first;
                fprintf(stderr, "ensemble round: %d realizations, mean %.0f cycles, median %lld, max %lld\n", nr, (double)total / nr, cost[(size_t)nr / 2].first,
                        cost.back().first);
                for (int t = 0; t < std::min(3, nr); t++) {
                    const int j = cost[(size_t)(nr - 1 - t)].second;
                    const LoopState &ls = ls_all_host[(size_t)j];
                    const Scalars &sc = sc_all_host[(size_t)j];
                    fprintf(stderr, "  #%d: handle %d steps %lld n_sph %d n_agg %d events %lld | cycles: pick %lld search %lld update %lld nucl+refresh %lld top %lld move %lld growth %lld merge %lld\n",
                            t, loop_set[(size_t)run[(size_t)j]], ls.steps, (int)sc.n_sph, (int)sc.n_agg, ls.events, ls.phase_cycles[0], ls.phase_cycles[1], ls.phase_cycles[2],
                            ls.phase_cycles[3], ls.phase_cycles[4], ls.phase_cycles[5], ls.phase_cycles[6], ls.phase_cycles[7]);
                }
            }
        }
        if (getenv("MCAC_B200_K9_DEBUG"))
            fprintf(stderr, "ensemble call: prologue %.2f ms, services %.2f, launch+wait %.2f, bookkeeping %.2f, so far %.2f ms, %lld rounds\n", prologue_ms, service_ms,
                    launch_ms, apply_ms, since(t_call), rounds);
        scratch_lock.unlock();
        for (int i = 0; i < m; i++) {
            const int k = loop_set[(size_t)i];
            if (rc_all != E_OK && rcs[(size_t)k] == E_OK) rcs[(size_t)k] = rc_all;
            if (reports && rcs[(size_t)k] == E_OK)
            {
                fill_report_basic(handles[k], tr[(size_t)i].at_start, tr[(size_t)i].launches0, reports + k, tr[(size_t)i].steps, tr[(size_t)i].dups,
                                  tr[(size_t)i].sorts, tr[(size_t)i].nucleated, tr[(size_t)i].fin);
                reports[k].device_ms = kernel_ms;  // CUDA-event time of the k_ensemble_loop launches of this call (shared by its realizations)
                reports[k].conflicts = rounds;     // rounds (= launches) of this call
                // host side of the call (ms, shared by its realizations): services, launch + wait + read-back, bookkeeping, whole call so far
                reports[k].search_ms = service_ms; reports[k].commit_ms = launch_ms; reports[k].event_ms = apply_ms; reports[k].cells_ms = since(t_call);
                // host services of the call: duplications / table regrows through the upload boundary, RNG refills
                reports[k].search_launches = n_dup_services; reports[k].commit_launches = n_regrow_services; reports[k].event_launches = n_rng_services;
            }
        }
    }
    int rc = E_OK;
    for (int k = 0; k < n; k++) if (rcs[(size_t)k] != E_OK && rc == E_OK) rc = rcs[(size_t)k];
    return rc;
}

}  // extern "C"
