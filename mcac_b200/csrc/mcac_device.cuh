// mcac_b200 — device-resident state of one realization (data layout in HBM) and launch-side helpers.
//
// Layout:
//  * Sphere pool, AGGREGATE-MAJOR: the spheres of one aggregate occupy a contiguous block of slots
//    [a_off, a_off + a_n) in the reference's `myspheres` order (include/aggregats/aggregat.hpp:81), so
//    that translate / contact search / update stream them with coalesced 32-byte vector loads
//    (x,y,z,r packed in one double4).  A merge writes the joined block at the top of the pool (bump
//    allocation, compacted when the pool fills up).  `s_id` is the reference's global sphere index
//    (creation order, include/constants.hpp:34-45 arrays), `slot_of_id` its inverse.
//  * Aggregates live in STABLE slots; the reference's compact label (index == label,
//    src/aggregats/aggregat_list_storage.cpp:37-44) is the rank of the slot among the live slots
//    (`label_of_slot`, refreshed on events).  Slot order == label order, so every label tie-break of the
//    reference can be evaluated on slots.
//  * Verlet cells (src/verlet/verlet.cpp): CSR over n_div^3 cells of aggregate slots, built by counting sort.
#pragma once
#include <cuda_runtime.h>

#include "mcac_math.cuh"

namespace mcacb {

struct Scalars {
    // PhysicalModel / AggregatList scalars that change during the run
    double time, box_length, box_volume, maxradius, max_time_step, avg_npp, cum_total;
    double total_volume, total_surface, nucleation_accum;
    double volume_fraction, aggregate_concentration, monomer_concentration, total_volume_concent, total_surface_concent;
    long long n_iter_without_event, total_events, steps_done, rand_pos, n_monomeres;
    long long pair_sphere, pair_bounding, searches, conflicts;
    int n_agg, n_agg_slots, n_sph, pool_top, n_pick;
    int event;  // calcul()'s `event` flag for the NEXT step (src/calcul.cpp:57,222)
    int error;  // sticky device-side error (ErrorCodes)
    // outcome of the last batch
    int b_committed, b_stop_reason, b_contact, b_merged, b_need;
    // general step: contact found by the search, merged after growth (calcul.cpp:174-181)
    int p_contact, p_ms, p_os, p_magg, p_oagg, p_slot;
    double p_dt, p_dt_indiv;
    int p_regime, p_regime_draws;  // check_InterPotentialRegime outcome of the last try (0 sticking) and draws it consumed
    int n_nucleated;
    int error_detail;  // what `error` means when the ErrorCodes value alone does not say (ErrorDetail)
    // pick table of a tie-dominated run: cumulative_time_steps is affine (slope pick_w) from index pick_dense_from on, up to
    // rounding — a hint for the lower_bound of pick_random (pick_dense_from == n_pick: no such stretch)
    double pick_w;
    int pick_dense_from;
    int last_sort_tie;  // 1: the last pick table was made by the sparse path of the event kernel (it is short: see event_spare_sms)
    // sphere-pair tests actually EXECUTED by the step loop's pruned sweeps (pair_sphere counts the tests the reference runs)
    long long pair_exec;
};
// finer cause of a device-side error (Scalars::error_detail); the host turns it into the message of mcac_gpu_last_error
enum ErrorDetail { DETAIL_NONE = 0, DETAIL_SUSPECT_OVERFLOW = 12, DETAIL_NOT_ON_VERLET = 13, DETAIL_RNG_NOT_STAGED = 21, DETAIL_PICK_TABLE = 22,
                   DETAIL_SPHERE_REMOVAL = 31, DETAIL_POOL_FULL = 32 };
enum StopReason { STOP_NONE = 0, STOP_CONTACT = 1, STOP_CONFLICT = 2, STOP_FINISHED = 3, STOP_BATCH_END = 4, STOP_POOL = 5 /* no room for the merged block: host compacts */ };

struct SearchResult {
    double distance;
    int moving_slot, other_slot;  // sphere slots (pool positions)
    int other_agg;                // aggregate slot
    int n_bounding;               // bounding-sphere prefilter tests of this search (aggregat_list.cpp:510-532)
    long long n_sphere_pairs;     // sphere-sphere tests the reference would run (examined prefix, :459-482)
    int status;
    int pad;
};

struct DevState {
    // ---- sphere pool (slot-indexed)
    double4 *s_posr;  // x, y, z, r
    double4 *s_relv;  // rx, ry, rz, volume         (position relative to the aggregate's root sphere; 4/3 pi r^3)
    double *s_surf;   // 4 pi r^2
    double *s_veff, *s_seff, *s_dcen;  // Aggregate::volumes / surfaces / distances_center (member order)
    int *s_id, *s_charge;
    int *slot_of_id;
    // ---- aggregates (slot-indexed)
    double4 *a_posr;  // x, y, z, rmax
    double *a_rg, *a_fagg, *a_lpm, *a_ts, *a_vol, *a_surf, *a_rx, *a_ry, *a_rz, *a_ptime, *a_dp, *a_dgdp, *a_ovl, *a_cn, *a_dm, *a_ch,
        *a_bulk, *a_alpha;
    int *a_n, *a_off, *a_cx, *a_cy, *a_cz, *a_charge, *a_alive;
    // 1: radii or membership changed since the aggregate's last FULL update (its contact graph, overlap statistics and effective
    // volumes / surfaces must be recomputed); 0: a full update would recompute exactly what is stored, so it runs as a partial one
    int *a_dirty;
    int *label_of_slot, *slot_of_label;
    // ---- Verlet cells
    int *cell_start, *cell_fill, *cell_items;
    double4 *cell_posr;  // (x, y, z, rmax) of cell_items[i], i.e. the aggregate bounding spheres in cell order (rebuilt with the cells)
    // ---- pick table (sorted 1/dt weights)
    int *sorted_slot;
    double *cum, *keys;
    // ---- RNG
    GlibcRandState *rng;
    int *rng_buf;
    long long rng_buf_base;  // stream position of rng_buf[0]
    int rng_buf_n;
    // strict replay mode: random_direction() of the draws (rng_buf[p], rng_buf[p+1]) evaluated on the HOST with glibc's sin / cos / acos
    // (3 doubles per buffer position), so that directions are the reference's bit for bit; nullptr = CUDA's sincos / acos (<= 2 ulp)
    const double *dir_tab;
    // ---- misc
    Scalars *sc;
    int agg_cap, sph_cap, n_div, n_cells;
    Gas gas;
    double u_sg, rp_min_oxid;
    int volsurf_method, pick_method, with_collisions;
    long long n_iter_limit, n_agg_limit, npp_limit;
    double time_limit;
    // interaction potentials (src/physical_model/physical_model_interpotential.cpp): table [q1][q2][dp1][dp2]
    int with_external_potentials, ip_n1, ip_n2, ip_nq;
    const int *ip_charge;
    const double *ip_dp1, *ip_dp2, *ip_ebar, *ip_ewell;
    // nucleation (physical_model.cpp:499-502, 557-578): diameter law of new monomers
    double nucl_mean_diameter, nucl_dispersion_diameter, flux_nucleation;
    int init_mode_normal;
    int cand_cap;  // eligible suspects a wide search keeps (<= kCandCap; MCAC_B200_CAND_CAP shrinks it to test the overflow error)
};

}  // namespace mcacb
