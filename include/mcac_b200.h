/*
 * mcac_b200 — C ABI of the B200-native (sm_100a) Monte-Carlo aggregation hot path.
 *
 * The reference (giraldeau/MCAC) has no FFI: its "boundary" is the C++ class surface that
 * mcac::calcul consumes.  Every entry point below replaces the body of one of those
 * methods; the citation after each prototype is the reference interface it stands in for.
 * Conventions: extern "C", plain pointers and sizes, POD structs only, no exceptions across the ABI.
 * Every function returns an `int` equal to the reference's ErrorCodes (include/constants.hpp:70-83):
 *   0 NO_ERROR, 1 UNKNOWN_ERROR (incl. CUDA errors), 2 IO, 3 VERLET, 4 INPUT, 5 ABANDON, 6 TOO_DENSE,
 *   7 SBL, 8 VOL_SURF, 9 MERGE, 10 ARVO, 11 InterPotential.
 * A handle owns all device memory of ONE realization, one CUDA stream, and its own RNG stream
 * (no process-global state); it is not thread-safe.  Aggregates are addressed by LABEL (the
 * reference's compact 0..N_agg-1 index) and spheres by their global creation index, exactly as in
 * the reference's HDF5 output (`Label`, sphere order).
 */
#ifndef MCAC_B200_H
#define MCAC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcac_gpu mcac_gpu; /* opaque */

enum { MCAC_PICK_RANDOM = 0, MCAC_PICK_LAST = 1 };                          /* constants.hpp:91-96  */
enum { MCAC_VS_CAPS = 0, MCAC_VS_SBL = 1, MCAC_VS_ARVO = 2, MCAC_VS_ALPHAS = 3, MCAC_VS_NONE = 4 }; /* :103-110 */
enum { MCAC_ORDER_LIBSTDCXX = 0, MCAC_ORDER_STABLE = 1 };

/* Physics + numerics of one realization: the PhysicalModel fields the hot path reads
 * (include/physical_model/physical_model.hpp:31-77), already derived by the host-side reader. */
typedef struct mcac_params {
    double box_length;              /* PhysicalModel::box_length                                   */
    double time;                    /* PhysicalModel::time (initial_time)                          */
    double temperature, pressure;   /* used through viscosity / gaz_mean_free_path                  */
    double viscosity, gaz_mean_free_path;
    double density;                 /* bulk density when with_maturity == 0                        */
    double fractal_dimension;
    double u_sg;                    /* surface growth velocity dr/dt                                */
    double rp_min_oxid;
    double flux_nucleation, nucleation_accum, box_volume;
    double mean_diameter_nucleation, dispersion_diameter_nucleation; /* law of the nucleated monomers [nm] */
    double physical_time_limit;     /* limits (PhysicalModel::finished, physical_model.cpp:288-337) */
    int64_t number_of_aggregates_limit;
    int64_t n_iter_without_event_limit;
    int64_t mean_monomere_per_aggregate_limit;
    int64_t n_monomeres;            /* N0 of the current box (×8 per duplication)                   */
    int64_t full_aggregate_update_frequency;
    int32_t n_verlet_divisions;
    int32_t pick_method, volsurf_method;
    int32_t with_collisions, with_surface_reactions, individual_surf_reactions, with_domain_duplication;
    int32_t with_maturity, with_potentials, with_external_potentials, with_nucleation, with_dynamic_random_charges;
    int32_t normal_initialisation;  /* monomeres_initialisation_type == NORMAL_INITIALISATION */
    int32_t sort_order;             /* MCAC_ORDER_LIBSTDCXX replays std::sort's tie order */
    uint32_t random_seed;           /* srand() argument (src/tools/tools.cpp:41-50)                 */
} mcac_params;

/* AggregateContactInfo (include/tools/contact_info.hpp:22-85) with ids instead of weak_ptrs.
 * "none" = distance +inf and every id -1, like a default-constructed SphereContactInfo. */
typedef struct mcac_contact {
    double distance;
    int64_t moving_sphere, other_sphere; /* global sphere indices */
    int64_t moving_label, other_label;   /* aggregate labels       */
} mcac_contact;

/* One MC step as the reference's taps see it (tests compare these records with the oracle). */
typedef struct mcac_step_record {
    int64_t step, rand_calls, source;
    double dir[3], full_distance, distance;
    int64_t moving_sphere, other_sphere, moving_label, other_label, n_agg_before;
    double time_before, dt, proper_time_after, pos_after[3];
    int64_t merged, n_try;
} mcac_step_record;

typedef struct mcac_run_report {
    int64_t steps, events, searches;       /* done by this call                                      */
    int64_t pair_tests_sphere, pair_tests_bounding; /* contact-pair tests evaluated on the device     */
    int64_t batches, conflicts, duplications, sorts, kernel_launches;
    int64_t n_aggregates, n_spheres, finished;
    double time, box_length, avg_npp, max_time_step, volume_fraction;
    double device_ms;                      /* CUDA-event time of the whole call on the handle's stream */
    double search_ms, commit_ms;           /* CUDA-event time spent in the K1 / commit+merge kernels (profile != 0) */
    int64_t search_launches, commit_launches;
    double event_ms, cells_ms;             /* per-event pipeline (k_event) / Verlet cell rebuild (K2) */
    int64_t event_launches, cells_launches;
    int64_t sort_span_elements, sort_levels; /* sum over sort levels of the active span (elements) / number of levels */
    /* SM cycles of block 0 of the event kernel per phase: 0 reduce, 1 labels, 2 sort init, 3 grid-wide levels, 4 block-local levels,
     * 5 leaf insertion sorts, 6 cumulative table, 7 pick table */
    int64_t event_phase_cycles[8];
    int64_t n_iter_without_event, nucleated;   /* PhysicalModel::n_iter_without_event after the call; monomers nucleated by it */
    double total_volume, total_surface;       /* AggregatList::get_total_volume / surface as of the last PhysicalModel::update */
    /* tie-dominated pick tables (csrc/tie_sort.cuh): SM cycles of the sparse simulation (one CTA) / of the routing pass (grid),
     * sorts that took the fast path, levels it simulated, sparse elements and elements handed to the general sort (sums) */
    int64_t tie_phase_cycles[2];
    int64_t tie_sorts, tie_levels, tie_sparse, tie_handed;
    /* SM cycles inside the sparse simulation: gathering the staged sparse elements, the simulated levels, hand-over + grid barrier */
    int64_t tie_sim_cycles[3];
    /* pick-table sorts of this call that the one-launch event kernel handed back to the multi-launch device path (introsort's depth
     * limit, or the MCAC_B200_FORCE_SORT_FAIL test hook), and sorts that replayed libstdc++'s heap-sort branch.  Every sort runs on
     * the device either way: there is no host sort in the product. */
    int64_t sort_fallbacks, sort_heap_branches;
    /* sphere-pair tests the per-realization step loop actually executed: its ordered sweep prunes pairs that cannot touch with
     * enclosing balls (same result); pair_tests_sphere stays the number of tests the reference runs.  0 on the other paths
     * (they execute what they count). */
    int64_t pair_tests_executed;
    /* SM cycles of the step loop's CTA per part of the step: 0 pick table (labels + sort), 1 cell rebuild + contact search (+ redraws),
     * 2 update block (calcul.cpp:184-206), 3 nucleation + bookkeeping + refresh, 4 loop top (checks, pool compaction), 5 move + clocks,
     * 6 surface growth, 7 merge (incl. Aggregate::update of the merged aggregate) */
    int64_t loop_phase_cycles[8];
} mcac_run_report;

/* One launch of K1 over `n` independent speculative searches drawn from the handle's RNG stream (pick + direction
 * exactly as in calcul(), nothing committed, stream position restored): the batched form of the contact search. */
typedef struct mcac_sweep_report {
    int64_t n_queries, contacts, pair_tests_sphere, pair_tests_bounding;
    double distance_checksum, kernel_ms;
} mcac_sweep_report;

/* --- lifetime ------------------------------------------------------------------------------- */
/* AggregatList::AggregatList(PhysicalModel*) minus placement (src/aggregats/aggregat_list_storage.cpp:52-88) */
int mcac_gpu_create(const mcac_params *params, int device, mcac_gpu **out);
int mcac_gpu_destroy(mcac_gpu *h);                       /* AggregatList::~AggregatList                */
const char *mcac_gpu_last_error(const mcac_gpu *h);      /* BaseException::what(), include/exceptions.hpp:30-60 */

/* init_random(seed) (src/tools/tools.cpp:41-50) followed by `consumed` draws already taken by the host-side
 * initial placement, so the device stream continues exactly where the reference's would */
int mcac_gpu_set_rng(mcac_gpu *h, uint32_t seed, int64_t consumed);

/* Interpotential(file) (src/physical_model/physical_model_interpotential.cpp:46-121): energy-barrier / well tables indexed
 * [charge1][charge2][dp1][dp2], already parsed by the host layer */
int mcac_gpu_set_interpotential(mcac_gpu *h, int32_t n_dp1, int32_t n_dp2, int32_t n_charge, const int32_t *val_charge,
                                const double *val_dp1, const double *val_dp2, const double *e_barr, const double *e_well);

/* --- state (host SoA <-> HBM) ---------------------------------------------------------------- */
/* Sphere fields in SpheresFields order, field-major: X,Y,Z,R,VOLUME,SURFACE,RX,RY,RZ (constants.hpp:34-45);
 * aggregate fields in AggregatesFields order, field-major, 21 x n_agg (constants.hpp:46-69);
 * membership = ordered `myspheres` lists as CSR (offsets n_agg+1, members n_sph);
 * per_member = Aggregate::volumes, surfaces, distances_center (3 x n_sph, member order). */
int mcac_gpu_upload_state(mcac_gpu *h, int64_t n_sph, int64_t n_agg, const double *sphere_fields, const int64_t *sphere_charge,
                          const double *agg_fields, const int64_t *agg_charge, const int64_t *agg_cells /*3 x n_agg*/,
                          const int64_t *offsets, const int64_t *members, const double *per_member, double maxradius,
                          double max_time_step);
int mcac_gpu_sizes(mcac_gpu *h, int64_t *n_sph, int64_t *n_agg);
/* Page-locked host memory for the arrays above (the storage behind the reference's ListStorage vectors,
 * include/list_storage/list_storage.hpp:32-34, when the caller wants full-bandwidth async transfers); pageable memory works too. */
int mcac_host_alloc_pinned(int64_t bytes, void **out);
int mcac_host_free_pinned(void *p);
int mcac_gpu_download_state(mcac_gpu *h, double *sphere_fields, int64_t *sphere_label, int64_t *sphere_charge, double *agg_fields,
                            int64_t *agg_n_spheres, int64_t *agg_charge, int64_t *agg_cells, int64_t *offsets, int64_t *members,
                            double *per_member, double *scalars /* 20, same order as the oracle's orc_get_scalars */);

/* --- the hot calls of mcac::calcul (src/calcul.cpp:66-281) ----------------------------------- */
/* AggregatList::distance_to_next_contact(source, direction, distance)   aggregat_list.cpp:447-484 */
int mcac_gpu_contact_search(mcac_gpu *h, int64_t source_label, const double direction[3], double distance, mcac_contact *out);
/* many independent searches against the same state in ONE launch (speculative batches, ensembles) */
int mcac_gpu_contact_search_batch(mcac_gpu *h, int64_t n, const int64_t *source_labels, const double *directions /*n x 3*/,
                                  const double *distances, mcac_contact *out, int64_t *pair_tests /*2: sphere, bounding; may be NULL*/);
/* Aggregate::translate(vector)                                          aggregat.cpp:148-161     */
int mcac_gpu_translate(mcac_gpu *h, int64_t label, const double vector[3]);
/* AggregatList::merge(contact_info) incl. Aggregate::merge / update / remove  aggregat_list.cpp:367-410 */
int mcac_gpu_merge(mcac_gpu *h, const mcac_contact *contact, int *merged);
/* AggregatList::croissance_surface(dt[, index]); label < 0 = all aggregates   aggregat_list.cpp:549-579 */
int mcac_gpu_grow(mcac_gpu *h, double dt, int64_t label);
/* Aggregate::update() (full != 0) / update_partial() for one label or all (label < 0)  aggregat.cpp:247-288 */
int mcac_gpu_update(mcac_gpu *h, int64_t label, int full);
/* Aggregate::get_lpm / get_time_step / ... (include/aggregats/aggregat.hpp:83-133): the 21 AggregatesFields of one aggregate (constants.hpp:
 * 46-69 order: RG, F_AGG, LPM, TIME_STEP, RMAX, VOLUME, SURFACE, X, Y, Z, RX, RY, RZ, TIME, DP, DG_OVER_DP, OVERLAPPING, COORDINATION_NUMBER,
 * ELECTRIC_CHARGE, D_M, CH_RATIO) and its number of spheres */
int mcac_gpu_aggregate_fields(mcac_gpu *h, int64_t label, double fields[21], int64_t *n_spheres);
/* AggregatList::refresh() + get_total_volume/surface() + PhysicalModel::update  aggregat_list.cpp:100-108,28-45 */
int mcac_gpu_refresh(mcac_gpu *h, double *max_time_step, double *avg_npp, double *total_volume, double *total_surface);
/* AggregatList::sort_time_steps(factor)                                 aggregat_list.cpp:124-141 */
int mcac_gpu_sort_time_steps(mcac_gpu *h, double factor);
int mcac_gpu_get_pick_table(mcac_gpu *h, int64_t *index_sorted /*labels*/, double *cumulative, int64_t *n);
/* AggregatList::pick_random() with the draw supplied / pick_last()      aggregat_list.cpp:59-81  */
int mcac_gpu_pick_random(mcac_gpu *h, double u, int64_t *label, double *deltatemps);
int mcac_gpu_pick_last(mcac_gpu *h, int64_t *label);
/* AggregatList::duplication()                                           aggregat_list.cpp:142-190 */
int mcac_gpu_duplicate(mcac_gpu *h);
/* device copy of the glibc rand() stream (tools.cpp:51-55): next n raw draws, advancing the handle's stream */
int mcac_gpu_rand(mcac_gpu *h, int64_t n, int32_t *out);

/* --- the whole loop: mcac::calcul(physicalmodel, aggregates) --------------------------------- */
/* Runs MC steps on the device until PhysicalModel::finished() or max_steps.  `records` (capacity n_records,
 * may be NULL) receives one mcac_step_record per step; `batch` = speculative batch width (0 = default). */
int mcac_gpu_run(mcac_gpu *h, int64_t max_steps, int32_t batch, mcac_step_record *records, int64_t n_records, mcac_run_report *report);

/* --- ensemble of independent realizations (the statistically required use: many seeds of one .ini) -------------------- */
/* Runs mcac_gpu_run(max_steps, batch) on each of the n handles, `threads` host threads driving them concurrently (handle k on
 * thread k mod threads; every handle has its own stream and RNG stream, nothing is shared).  reports: n entries or NULL.
 * Returns the first non-zero error code of any realization. */
int mcac_ensemble_run(mcac_gpu **handles, int32_t n, int64_t max_steps, int32_t batch, int32_t threads, mcac_run_report *reports);

int mcac_gpu_search_sweep(mcac_gpu *h, int64_t n, int32_t repeats, mcac_sweep_report *report);
/* Times one kernel of the path on the resident state (CUDA events on the handle's stream, `reps` launches after one warm-up):
 * which = 0 K2 cell rebuild, 1 K8 growth (all spheres), 2 update_partial (all aggregates), 3 full update, 4 K9 event pipeline with
 * sort, 5 without sort, 6 100 grid barriers at K9's launch shape, 7 K10 RNG fill, 8 K11 statistics, 9 / 10 FP64 pipe peak with DFMA /
 * with DMUL + DADD (the library is built --fmad=false): units = flops per launch; 11 tuning probe: the one-CTA sparse simulation of K9's
 * tie-dominated sort alone, on a synthetic table of the resident size (MCAC_B200_PROBE_X sparse elements).  units = items per launch
 * otherwise. */
int mcac_gpu_kernel_bench(mcac_gpu *h, int32_t which, int32_t reps, double *ms_per_launch, int64_t *units);
/* on != 0: mcac_gpu_run returns right after the step that made an event (merge or nucleation: `event` of calcul.cpp:222), so that a
 * host loop can do what calcul() does between events (advancement.dat rows, the progress table, output files) */
int mcac_gpu_set_stop_at_event(mcac_gpu *h, int32_t on);
/* Table sizes to allocate at least at the next mcac_gpu_upload_state / regrow (std::vector::reserve of the reference's ListStorage
 * vectors, include/list_storage/list_storage.hpp:32-34): with room for the next domain duplications the device loop does them itself. */
int mcac_gpu_reserve(mcac_gpu *h, int64_t n_spheres, int64_t n_aggregates);
/* on != 0: strict replay mode.  random_direction() (src/tools/tools.cpp:82-89) is evaluated on the host with glibc's sin / cos / acos for
 * every staged pair of draws and read from a table by the kernels, so directions — and with them contact distances, positions and
 * clocks — are the reference's bit for bit instead of within 2 ulp of CUDA's sincos / acos (costs one host pass per ~10^6 draws). */
int mcac_gpu_set_strict_direction(mcac_gpu *h, int32_t on);
/* profile != 0: mcac_gpu_run brackets its K1 / commit launches with CUDA events (reported in mcac_run_report) */
int mcac_gpu_set_profile(mcac_gpu *h, int32_t profile);

/* --- ensemble statistics (K11): per-realization morphology histogram staged for the NCCL all-gather ---- */
/* out[0..n_bins) = histogram of log2(Np), out[n_bins..2n_bins) = histogram of Rg over [0,rg_max),
 * then {n_agg, sum Np, sum lx, sum lx^2, sum lx*ly, sum ly, sum ly^2, sum Rg} with lx = log(dg/dp), ly = log(Np): the sums of
 * AggregatList::get_instantaneous_fractal_law -> linreg (aggregat_list_fractal_law.cpp:23-33, tools.cpp:126-157). */
int mcac_gpu_morphology_stats(mcac_gpu *h, int32_t n_bins, double rg_max, double *out /* 2*n_bins + 8 */);
/* same, written to a device buffer (for torch.distributed all_gather without a host round trip) */
int mcac_gpu_morphology_stats_device(mcac_gpu *h, int32_t n_bins, double rg_max, void *device_out);
/* raw CUDA stream of the handle (cudaStream_t) so callers can order their own work / events on it */
void *mcac_gpu_stream(mcac_gpu *h);

/* --- host-side mirror: PhysicalModel(ini) + AggregatList placement (src/main.cpp:26-56) -------------------------- */
/* These run on the host (the rejection-sampling placement is sequential in the RNG stream) and need no GPU. */
typedef struct mcac_host_model mcac_host_model; /* opaque: PhysicalModel + the placed monomers */
const char *mcac_host_last_error(void);
/* PhysicalModel::PhysicalModel(ini) (physical_model.cpp:32-287); place != 0 also runs the AggregatList ctor placement */
int mcac_host_model_create(const char *ini_text, int place, mcac_host_model **out);
void mcac_host_model_destroy(mcac_host_model *m);
int mcac_host_model_params(const mcac_host_model *m, mcac_params *out);
int mcac_host_model_sizes(const mcac_host_model *m, int64_t *n_sph, int64_t *n_agg);
int mcac_host_model_metadata(const mcac_host_model *m, char *buf, int64_t cap); /* PhysicalModel::xmf_write, io/physical_model.cpp:30-49 */
int mcac_host_model_derived(const mcac_host_model *m, double out[12]);
/* the echo of the parsed .ini the reference writes to <output_dir>/params.ini (inipp::Ini::generate, physical_model.cpp:271-272) */
int mcac_host_model_ini_echo(const mcac_host_model *m, char *buf, int64_t cap);
int mcac_host_model_state(const mcac_host_model *m, double *sphere_fields, double *agg_fields, int64_t *agg_cells, int64_t *offsets,
                          int64_t *members, double *per_member, double *scalars /*maxradius,max_time_step,avg_npp*/,
                          int64_t *rand_consumed);
/* --- output files of the reference (src/io/*): <prefix>_<k>.h5 (HDF5 heavy data) + <prefix>_<k>.xmf (XDMF light data) ----------- */
/* One series = the reference's ThreadedIO for "Spheres" or "Aggregats" (threaded_io.cpp, writer.cpp:72-142): n_time_per_file grids per
 * file pair, k zero-padded to ceil(log10(n_for_width)) + 4 digits (format.cpp:60-65).  Written by a libhdf5-free HDF5 writer
 * (superblock v0, contiguous datasets Data0, Data1, ...: what XdmfHDF5Writer emits with deflate off).  physics = "name=value\n" lines
 * of PhysicalModel::xmf_write (io/physical_model.cpp:30-49).  type: 0 = float64, 1 = int32, 2 = int64. */
typedef struct mcac_io_writer mcac_io_writer;
int mcac_io_writer_create(const char *prefix, const char *grid_name, int64_t n_time_per_file, int64_t n_for_width, const char *physics,
                          mcac_io_writer **out);
int mcac_io_begin_step(mcac_io_writer *w, double time);                                    /* XdmfUnstructuredGrid + XdmfTime      */
int mcac_io_positions(mcac_io_writer *w, const double *xyz_interleaved, int64_t n_points); /* the_positions(), writer.cpp:37-43   */
int mcac_io_attribute(mcac_io_writer *w, const char *name, int32_t type, const void *data, int64_t count, int32_t scalar_on_nodes);
int mcac_io_end_step(mcac_io_writer *w);                                                   /* ThreadedIO::write                     */
int mcac_io_writer_destroy(mcac_io_writer *w);                                             /* ~ThreadedIO: flushes the open file    */
/* SphereList::save() + AggregatList::save() (io/sphere_list.cpp:36-57, io/aggregat_list.cpp:36-67) of the state resident in HBM:
 * one grid appended to each of the two series */
int mcac_gpu_save(mcac_gpu *h, mcac_io_writer *spheres, mcac_io_writer *aggregates);

/* main(): PhysicalModel(ini) -> AggregatList(&physicalmodel) with the state resident in HBM of `device` */
int mcac_sim_create(const char *ini_text, int device, mcac_gpu **out);

#ifdef __cplusplus
}
#endif
#endif /* MCAC_B200_H */
