// aggregat_list_gpu.cpp — the reference-side binding of INTEGRATION.md §3, compilable against the reference's REAL headers.
//
// A maintainer of giraldeau/MCAC compiles this translation unit instead of the CPU bodies of src/aggregats/aggregat_list.cpp
// (-DWITH_MCAC_B200): the class surface mcac::calcul consumes (include/aggregats/aggregat_list.hpp:43-116) keeps its exact
// signatures and every body forwards to the C ABI of include/mcac_b200.h.  tests/test_host_layer.py compiles it with
// -I/root/reference/include (when the reference is present) so that a drift between the C ABI and the reference's interface
// is a build error, not prose.  The device handle is kept in a side table keyed by the AggregatList (the maintainer would add a
// `mcac_gpu *gpu` member under the same #ifdef instead).
#include <unordered_map>

#include "aggregats/aggregat_list.hpp"
#include "exceptions.hpp"
#include "spheres/sphere.hpp"
#include "mcac_b200.h"
#include "tools/tools.hpp"

namespace mcac {
namespace {
std::unordered_map<const AggregatList *, mcac_gpu *> &handles() {
    static std::unordered_map<const AggregatList *, mcac_gpu *> table;
    return table;
}
mcac_gpu *gpu_of(const AggregatList *a) { return handles().at(a); }
// ErrorCodes -> the matching exception class of include/exceptions.hpp (main.cpp:43-54 maps them back to exit codes)
void check(mcac_gpu *h, int rc) {
    if (rc == static_cast<int>(ErrorCodes::NO_ERROR)) return;
    const std::string msg = mcac_gpu_last_error(h);
    switch (static_cast<ErrorCodes>(rc)) {
    case ErrorCodes::VERLET_ERROR: throw VerletError(msg);
    case ErrorCodes::MERGE_ERROR: throw MergeError(msg);
    case ErrorCodes::VOL_SURF_ERROR: throw VolSurfError();
    case ErrorCodes::INPUT_ERROR: throw InputError(msg);
    case ErrorCodes::TOO_DENSE_ERROR: throw TooDenseError();
    case ErrorCodes::IO_ERROR: throw IOError(msg);
    default: throw BaseException("mcac_b200: " + msg);
    }
}
}  // namespace

void mcac_b200_attach(const AggregatList *a, mcac_gpu *h) { handles()[a] = h; }

// aggregat_list.hpp:69-71 / aggregat_list.cpp:447-484
AggregateContactInfo AggregatList::distance_to_next_contact(const size_t source, const std::array<double, 3> &direction,
                                                            const double distance) const {
    mcac_gpu *gpu = gpu_of(this);
    mcac_contact c;
    check(gpu, mcac_gpu_contact_search(gpu, static_cast<int64_t>(source), direction.data(), distance, &c));
    AggregateContactInfo info;  // default: distance = +inf, expired weak_ptrs (contact_info.hpp:25-27)
    info.distance = c.distance;
    if (c.other_label >= 0) {  // ids -> the reference's weak_ptrs
        info.moving_sphere = spheres[static_cast<size_t>(c.moving_sphere)];
        info.other_sphere = spheres[static_cast<size_t>(c.other_sphere)];
        info.moving_aggregate = list[static_cast<size_t>(c.moving_label)];
        info.other_aggregate = list[static_cast<size_t>(c.other_label)];
    }
    return info;
}
// aggregat_list.hpp:63 / aggregat_list.cpp:367-410
bool AggregatList::merge(AggregateContactInfo contact_info) {
    const std::shared_ptr<Sphere> ms = contact_info.moving_sphere.lock(), os = contact_info.other_sphere.lock();
    if (!ms || !os) return false;
    mcac_gpu *gpu = gpu_of(this);
    mcac_contact c{contact_info.distance, static_cast<int64_t>(ms->get_index()), static_cast<int64_t>(os->get_index()), -1, -1};
    int merged = 0;
    check(gpu, mcac_gpu_merge(gpu, &c, &merged));
    return merged != 0;
}
// aggregat_list.cpp:124-141, 59-66, 67-81, 54-58, 100-108
void AggregatList::sort_time_steps(double factor) { check(gpu_of(this), mcac_gpu_sort_time_steps(gpu_of(this), factor)); }
size_t AggregatList::pick_random() const {
    int64_t label = 0;
    double dt = 0.;
    check(gpu_of(this), mcac_gpu_pick_random(gpu_of(this), random(), &label, &dt));  // random(): the reference's own draw (tools.cpp:51-55)
    return static_cast<size_t>(label);
}
size_t AggregatList::pick_last() const {
    int64_t label = 0;
    check(gpu_of(this), mcac_gpu_pick_last(gpu_of(this), &label));
    return static_cast<size_t>(label);
}
double AggregatList::get_time_step(double max) const {
    int64_t label = 0;
    double dt = 0.;
    check(gpu_of(this), mcac_gpu_pick_random(gpu_of(this), 0., &label, &dt));  // dt = max_time_step / cumulative.back()
    return dt * (max / max_time_step);
}
void AggregatList::refresh() {
    double tv = 0., ts = 0.;
    check(gpu_of(this), mcac_gpu_refresh(gpu_of(this), &max_time_step, &avg_npp, &tv, &ts));
}
double AggregatList::get_total_volume() const {
    double mx, npp, tv = 0., ts = 0.;
    check(gpu_of(this), mcac_gpu_refresh(gpu_of(this), &mx, &npp, &tv, &ts));
    return tv;
}
double AggregatList::get_total_surface() const {
    double mx, npp, tv = 0., ts = 0.;
    check(gpu_of(this), mcac_gpu_refresh(gpu_of(this), &mx, &npp, &tv, &ts));
    return ts;
}
// aggregat_list.cpp:549-579
bool AggregatList::croissance_surface(double dt) {
    check(gpu_of(this), mcac_gpu_grow(gpu_of(this), dt, -1));
    return false;
}
bool AggregatList::croissance_surface(const double dt, const size_t index) {
    check(gpu_of(this), mcac_gpu_grow(gpu_of(this), dt, static_cast<int64_t>(index)));
    return false;
}
// aggregat_list.cpp:142-190
void AggregatList::duplication() { check(gpu_of(this), mcac_gpu_duplicate(gpu_of(this))); }

// Aggregate::translate / update / update_partial (aggregat.hpp:83-133, aggregat.cpp:148-161, 247-288) by label
void mcac_b200_translate(const AggregatList *a, size_t label, std::array<double, 3> vector) {
    check(gpu_of(a), mcac_gpu_translate(gpu_of(a), static_cast<int64_t>(label), vector.data()));
}
void mcac_b200_update(const AggregatList *a, long label, bool full) { check(gpu_of(a), mcac_gpu_update(gpu_of(a), label, full ? 1 : 0)); }
}  // namespace mcac
