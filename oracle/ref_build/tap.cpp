// TEST INFRASTRUCTURE ONLY — link-time taps on the UNMODIFIED reference binary.
//
// Built into oracle/_ref/MCAC_tap with `-Wl,--wrap=<symbol>`: every cross-TU call of a wrapped
// reference function lands in __wrap_<symbol> here, is logged, and is forwarded to __real_<symbol>.
// No reference logic is restated in this file; it only *observes* the reference:
//   rand()                                     src/tools/tools.cpp:51-55      (draw counter)
//   AggregatList::distance_to_next_contact     src/aggregats/aggregat_list.cpp:447-484
//   AggregatList::merge                        src/aggregats/aggregat_list.cpp:367-410
//   AggregatList::sort_time_steps              src/aggregats/aggregat_list.cpp:124-141
//   Aggregate::time_forward                    src/aggregats/aggregat.cpp:106-108 (1 call per MC step, calcul.cpp:149)
//   distance_to_contact(Aggregate,Aggregate)   src/aggregats/aggregat_distance.cpp:24-44 (pair-test counter)
//   Verlet::get_neighborhood (3-arg)           src/verlet/verlet.cpp:52-98     (prefilter counter)
//   mcac::calcul                               src/calcul.cpp:55-290           (initial/final state, summary)
//
// Environment:
//   MCAC_TAP_DIR          directory for the outputs (default ".")
//   MCAC_TAP_MAX_STEPS    stop logging per-step records after this many steps (default 0 = none)
//   MCAC_TAP_STATE_STEPS  comma list of MC step numbers; full state dumped when that step's move is done
//                         (inside time_forward, i.e. after translate, before growth/merge)
//   MCAC_TAP_SORT_CALLS   comma list of sort_time_steps call numbers whose result is dumped
//   MCAC_TAP_EXIT_STEP    call _exit(0) once this many steps were done (bounded samples)
//   MCAC_TAP_CHUNK        record the wall clock (s since calcul started) every CHUNK steps -> summary.txt `chunk_times`
// Outputs (raw little-endian, layouts in tests/ref_trace.py):
//   steps.bin  searches.bin  merges.bin  sort_<k>.bin  state_<step>.bin  state_init.bin  state_final.bin  summary.txt
#include "aggregats/aggregat_list.hpp"
#include "spheres/sphere.hpp"
#include "calcul.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <unistd.h>

using mcac::AggregatList;
using mcac::Aggregate;
using mcac::AggregateContactInfo;
using mcac::PhysicalModel;
using mcac::Verlet;

#define SYM_SEARCH _ZNK4mcac12AggregatList24distance_to_next_contactEmRKSt5arrayIdLm3EEd
#define SYM_MERGE _ZN4mcac12AggregatList5mergeENS_20AggregateContactInfoE
#define SYM_TFWD _ZN4mcac9Aggregate12time_forwardEd
#define SYM_SORT _ZN4mcac12AggregatList15sort_time_stepsEd
#define SYM_CALCUL _ZN4mcac6calculERNS_13PhysicalModelERNS_12AggregatListE
#define SYM_AGGDIST _ZN4mcac19distance_to_contactERKSt10shared_ptrINS_9AggregateEES4_RKSt5arrayIdLm3EEd
#define SYM_VERLET _ZNK4mcac6Verlet16get_neighborhoodERKSt5arrayIdLm3EES4_d
#define STR2(x) #x
#define STR(x) STR2(x)
#define REAL(x) asm("__real_" STR(x))
#define WRAP(x) asm("__wrap_" STR(x))

extern "C" int __real_rand(void);
AggregateContactInfo real_search(const AggregatList *, size_t, const std::array<double, 3> &, double) REAL(SYM_SEARCH);
bool real_merge(AggregatList *, AggregateContactInfo) REAL(SYM_MERGE);
void real_time_forward(Aggregate *, double) REAL(SYM_TFWD);
void real_sort(AggregatList *, double) REAL(SYM_SORT);
void real_calcul(PhysicalModel &, AggregatList &) REAL(SYM_CALCUL);
AggregateContactInfo real_aggdist(const std::shared_ptr<Aggregate> &, const std::shared_ptr<Aggregate> &,
                                  const std::array<double, 3> &, double) REAL(SYM_AGGDIST);
std::vector<size_t> real_verlet(const Verlet *, const std::array<double, 3> &, const std::array<double, 3> &,
                                double) REAL(SYM_VERLET);

namespace {
struct Tap {
    std::string dir = ".";
    long long max_steps = 0, exit_step = -1, chunk = 0;
    std::vector<double> chunk_times;
    std::set<long long> state_steps, sort_calls;
    FILE *f_steps = nullptr, *f_search = nullptr, *f_merge = nullptr;
    long long n_rand = 0, n_steps = 0, n_search = 0, n_merge_calls = 0, n_merge_ok = 0, n_sort = 0;
    long long n_pair_sphere = 0, n_pair_bound = 0, n_aggdist = 0;
    const AggregatList *list = nullptr;
    PhysicalModel *pm = nullptr;
    std::chrono::steady_clock::time_point t0;
    bool in_calcul = false;
    Tap() {
        if (const char *e = getenv("MCAC_TAP_DIR")) dir = e;
        if (const char *e = getenv("MCAC_TAP_MAX_STEPS")) max_steps = atoll(e);
        if (const char *e = getenv("MCAC_TAP_EXIT_STEP")) exit_step = atoll(e);
        if (const char *e = getenv("MCAC_TAP_CHUNK")) chunk = atoll(e);
        parse(getenv("MCAC_TAP_STATE_STEPS"), state_steps);
        parse(getenv("MCAC_TAP_SORT_CALLS"), sort_calls);
    }
    static void parse(const char *e, std::set<long long> &out) {
        if (!e) return;
        std::string s(e);
        size_t p = 0;
        while (p < s.size()) {
            size_t q = s.find(',', p);
            if (q == std::string::npos) q = s.size();
            if (q > p) out.insert(atoll(s.substr(p, q - p).c_str()));
            p = q + 1;
        }
    }
    FILE *open(const std::string &name) {
        FILE *f = fopen((dir + "/" + name).c_str(), "wb");
        if (!f) { perror(("tap: cannot open " + dir + "/" + name).c_str()); exit(99); }
        return f;
    }
    bool logging() const { return n_steps < max_steps; }
    void summary(const char *why);
};
Tap tap;

template <class T> void put(FILE *f, const T &v) { fwrite(&v, sizeof(T), 1, f); }
void put_i64(FILE *f, long long v) { put(f, v); }
void put_f64(FILE *f, double v) { put(f, v); }

// Full SoA snapshot: header, 9 sphere fields, labels, charges, 21 aggregate fields, per-aggregate ints,
// CSR membership (myspheres order), per-member volumes/surfaces/distances_center.
void dump_state(const AggregatList &al, const std::string &name) {
    FILE *f = tap.open(name);
    const PhysicalModel &pm = *al.physicalmodel;
    long long n_sph = (long long)al.spheres.size(), n_agg = (long long)al.size();
    put_i64(f, 0x4d434143534e4150LL);  // "MCACSNAP"
    put_i64(f, tap.n_steps);
    put_i64(f, tap.n_rand);
    put_i64(f, n_sph);
    put_i64(f, n_agg);
    put_i64(f, (long long)pm.n_monomeres);
    put_i64(f, (long long)pm.n_iter_without_event);
    put_f64(f, pm.time);
    put_f64(f, pm.box_length);
    put_f64(f, al.maxradius);
    put_f64(f, al.max_time_step);
    put_f64(f, al.avg_npp);
    put_f64(f, pm.volume_fraction);
    put_f64(f, pm.aggregate_concentration);
    put_f64(f, pm.monomer_concentration);
    put_f64(f, pm.total_volume_concent);
    put_f64(f, pm.total_surface_concent);
    for (int fld = 0; fld < mcac::SpheresFields::SPHERE_NFIELDS; fld++)
        fwrite((*al.spheres.storage)[fld].data(), sizeof(double), (size_t)n_sph, f);
    for (long long i = 0; i < n_sph; i++) put_i64(f, (long long)al.spheres[i]->agg_label);
    for (long long i = 0; i < n_sph; i++) put_i64(f, (long long)al.spheres[i]->electric_charge);
    for (int fld = 0; fld < mcac::AggregatesFields::AGGREGAT_NFIELDS; fld++)
        fwrite((*al.storage)[fld].data(), sizeof(double), (size_t)n_agg, f);
    for (long long i = 0; i < n_agg; i++) put_i64(f, (long long)al[i]->n_spheres);
    for (long long i = 0; i < n_agg; i++) put_i64(f, (long long)al[i]->label);
    for (long long i = 0; i < n_agg; i++) put_i64(f, (long long)al[i]->electric_charge);
    for (int d = 0; d < 3; d++)
        for (long long i = 0; i < n_agg; i++) put_i64(f, (long long)al[i]->index_verlet[d]);
    for (long long i = 0; i < n_agg; i++) put_f64(f, al[i]->bulk_density);
    for (long long i = 0; i < n_agg; i++) put_f64(f, al[i]->alpha_vs_extreme);
    long long off = 0;
    for (long long i = 0; i < n_agg; i++) { put_i64(f, off); off += (long long)al[i]->myspheres.size(); }
    put_i64(f, off);
    for (long long i = 0; i < n_agg; i++)
        for (size_t k = 0; k < al[i]->myspheres.size(); k++) put_i64(f, (long long)al[i]->myspheres[k]->get_index());
    for (int which = 0; which < 3; which++)
        for (long long i = 0; i < n_agg; i++) {
            const std::vector<double> &v = which == 0 ? al[i]->volumes : which == 1 ? al[i]->surfaces : al[i]->distances_center;
            for (size_t k = 0; k < al[i]->myspheres.size(); k++) put_f64(f, k < v.size() ? v[k] : 0.0);
        }
    fclose(f);
    // the reference's own morphology regression on this state (AggregatList::get_instantaneous_fractal_law ->
    // linreg, aggregat_list_fractal_law.cpp:23-33; const, not called by calcul): pins K11 / ensemble.fractal_law
    const auto law = al.get_instantaneous_fractal_law();
    FILE *g = tap.open(name + ".fractal");
    put_f64(g, std::get<0>(law) ? 1.0 : 0.0);
    put_f64(g, std::get<1>(law));
    put_f64(g, std::get<2>(law));
    put_f64(g, std::get<3>(law));
    fclose(g);
}

void Tap::summary(const char *why) {
    double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    FILE *f = open("summary.txt");
    fprintf(f, "reason %s\n", why);
    fprintf(f, "steps %lld\nrand_calls %lld\nsearch_calls %lld\nmerge_calls %lld\nmerges %lld\nsort_calls %lld\n", n_steps, n_rand,
            n_search, n_merge_calls, n_merge_ok, n_sort);
    fprintf(f, "pair_tests_sphere %lld\npair_tests_bounding %lld\naggregate_pair_calls %lld\n", n_pair_sphere, n_pair_bound, n_aggdist);
    fprintf(f, "calcul_wall_s %.6f\n", wall);
    if (!chunk_times.empty()) {
        fprintf(f, "chunk_times");
        for (double t : chunk_times) fprintf(f, " %.6f", t);
        fprintf(f, "\n");
    }
    if (list) fprintf(f, "n_agg %zu\nn_sph %zu\n", list->size(), list->spheres.size());
    if (pm) fprintf(f, "time %.17g\nbox_length %.17g\n", pm->time, pm->box_length);
    fclose(f);
    if (f_steps) fflush(f_steps);
    if (f_search) fflush(f_search);
    if (f_merge) fflush(f_merge);
}
}  // namespace

extern "C" int __wrap_rand(void) {
    tap.n_rand++;
    return __real_rand();
}

AggregateContactInfo wrap_search(const AggregatList *self, size_t source, const std::array<double, 3> &dir, double dist) WRAP(SYM_SEARCH);
AggregateContactInfo wrap_search(const AggregatList *self, size_t source, const std::array<double, 3> &dir, double dist) {
    tap.list = self;
    AggregateContactInfo res = real_search(self, source, dir, dist);
    if (tap.in_calcul) {
        tap.n_search++;
        if (tap.logging()) {
            if (!tap.f_search) tap.f_search = tap.open("searches.bin");
            FILE *f = tap.f_search;
            auto ms = res.moving_sphere.lock();
            auto os = res.other_sphere.lock();
            auto ma = res.moving_aggregate.lock();
            auto oa = res.other_aggregate.lock();
            put_i64(f, tap.n_steps);
            put_i64(f, tap.n_rand);
            put_i64(f, (long long)source);
            put_f64(f, dir[0]); put_f64(f, dir[1]); put_f64(f, dir[2]);
            put_f64(f, dist);
            put_f64(f, res.distance);
            put_i64(f, ms ? (long long)ms->get_index() : -1);
            put_i64(f, os ? (long long)os->get_index() : -1);
            put_i64(f, ma ? (long long)ma->get_label() : -1);
            put_i64(f, oa ? (long long)oa->get_label() : -1);
            put_i64(f, (long long)self->size());
            put_f64(f, self->physicalmodel->time);
        }
    }
    return res;
}

bool wrap_merge(AggregatList *self, AggregateContactInfo info) WRAP(SYM_MERGE);
bool wrap_merge(AggregatList *self, AggregateContactInfo info) {
    bool ok = real_merge(self, info);
    tap.n_merge_calls++;
    if (ok) tap.n_merge_ok++;
    if (tap.n_steps <= tap.max_steps) {
        if (!tap.f_merge) tap.f_merge = tap.open("merges.bin");
        put_i64(tap.f_merge, tap.n_steps - 1);  // step whose move produced the contact
        put_i64(tap.f_merge, ok ? 1 : 0);
        put_i64(tap.f_merge, (long long)self->size());
        put_i64(tap.f_merge, (long long)self->spheres.size());
    }
    return ok;
}

// One call per MC step (calcul.cpp:149), right after translate: marks the end of the move.
void wrap_time_forward(Aggregate *self, double dt) WRAP(SYM_TFWD);
void wrap_time_forward(Aggregate *self, double dt) {
    real_time_forward(self, dt);
    if (!tap.in_calcul) return;
    if (tap.logging()) {
        if (!tap.f_steps) tap.f_steps = tap.open("steps.bin");
        FILE *f = tap.f_steps;
        std::array<double, 3> p = self->get_position();
        put_i64(f, tap.n_steps);
        put_i64(f, tap.n_rand);
        put_i64(f, (long long)self->get_label());
        put_f64(f, dt);
        put_f64(f, self->get_proper_time());
        put_f64(f, p[0]); put_f64(f, p[1]); put_f64(f, p[2]);
        put_f64(f, self->get_lpm());
    }
    if (tap.list && tap.state_steps.count(tap.n_steps)) dump_state(*tap.list, "state_" + std::to_string(tap.n_steps) + ".bin");
    tap.n_steps++;
    if (tap.chunk > 0 && tap.n_steps % tap.chunk == 0)
        tap.chunk_times.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - tap.t0).count());
    if (tap.exit_step >= 0 && tap.n_steps >= tap.exit_step) {
        tap.summary("exit_step");
        if (tap.list) dump_state(*tap.list, "state_final.bin");
        fflush(nullptr);
        _exit(0);
    }
}

void wrap_sort(AggregatList *self, double factor) WRAP(SYM_SORT);
void wrap_sort(AggregatList *self, double factor) {
    tap.list = self;
    real_sort(self, factor);
    if (tap.sort_calls.count(tap.n_sort)) {
        FILE *f = tap.open("sort_" + std::to_string(tap.n_sort) + ".bin");
        long long n = (long long)self->size();
        put_i64(f, tap.n_steps);
        put_i64(f, n);
        put_f64(f, factor);
        for (long long i = 0; i < n; i++) put_i64(f, (long long)self->index_sorted_time_steps[(size_t)i]);
        fwrite(self->cumulative_time_steps.data(), sizeof(double), (size_t)n, f);
        for (long long i = 0; i < n; i++) put_f64(f, (*self)[(size_t)i]->get_time_step());
        fclose(f);
    }
    tap.n_sort++;
}

AggregateContactInfo wrap_aggdist(const std::shared_ptr<Aggregate> &a, const std::shared_ptr<Aggregate> &b,
                                  const std::array<double, 3> &dir, double dist) WRAP(SYM_AGGDIST);
AggregateContactInfo wrap_aggdist(const std::shared_ptr<Aggregate> &a, const std::shared_ptr<Aggregate> &b,
                                  const std::array<double, 3> &dir, double dist) {
    if (tap.in_calcul) {
        tap.n_aggdist++;
        tap.n_pair_sphere += (long long)(a->size() * b->size());
    }
    return real_aggdist(a, b, dir, dist);
}

std::vector<size_t> wrap_verlet(const Verlet *self, const std::array<double, 3> &pos, const std::array<double, 3> &vec, double d) WRAP(SYM_VERLET);
std::vector<size_t> wrap_verlet(const Verlet *self, const std::array<double, 3> &pos, const std::array<double, 3> &vec, double d) {
    std::vector<size_t> r = real_verlet(self, pos, vec, d);
    if (tap.in_calcul && !r.empty()) tap.n_pair_bound += (long long)r.size() - 1;  // self is removed before the prefilter
    return r;
}

void wrap_calcul(PhysicalModel &pm, AggregatList &al) WRAP(SYM_CALCUL);
void wrap_calcul(PhysicalModel &pm, AggregatList &al) {
    tap.list = &al;
    tap.pm = &pm;
    tap.in_calcul = true;
    tap.t0 = std::chrono::steady_clock::now();
    dump_state(al, "state_init.bin");
    real_calcul(pm, al);
    tap.summary("finished");
    dump_state(al, "state_final.bin");
    tap.in_calcul = false;
}
