// mcac_b200 host layer — AggregatList facade + calcul() (see aggregat_list.hpp).
#include "aggregat_list.hpp"

#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>

#include "../csrc/mcac_math.cuh"
#include "placement.hpp"

namespace mcac {

void AggregatList::check(int rc) const {
    if (rc == NO_ERROR) return;
    const std::string msg = gpu ? mcac_gpu_last_error(gpu) : "no device handle";
    switch (rc) {
        case VERLET_ERROR: throw VerletError(msg);
        case MERGE_ERROR: throw MergeError(msg);
        case VOL_SURF_ERROR: throw VolSurfError();
        case INPUT_ERROR: throw InputError(msg);
        case TOO_DENSE_ERROR: throw TooDenseError();
        default: throw DeviceError(msg);
    }
}

AggregatList::AggregatList(PhysicalModel *pm, int device) : physicalmodel(pm) {
    InitialState st = place_monomers(*pm);
    mcac_params prm = pm->to_params();
    check(mcac_gpu_create(&prm, device, &gpu));
    check(mcac_gpu_set_rng(gpu, prm.random_seed, st.rand_consumed));
    check(mcac_gpu_upload_state(gpu, st.n_sph, st.n_agg, st.sphere_fields.data(), st.sphere_charge.data(), st.agg_fields.data(),
                                st.agg_charge.data(), st.agg_cells.data(), st.offsets.data(), st.members.data(), st.per_member.data(),
                                st.maxradius, st.max_time_step));
    max_time_step = st.max_time_step;
    avg_npp = st.avg_npp;
}
AggregatList::~AggregatList() noexcept {
    if (gpu) mcac_gpu_destroy(gpu);
}
size_t AggregatList::size() const {
    int64_t ns = 0, na = 0;
    check(mcac_gpu_sizes(gpu, &ns, &na));
    return static_cast<size_t>(na);
}
size_t AggregatList::n_spheres() const {
    int64_t ns = 0, na = 0;
    check(mcac_gpu_sizes(gpu, &ns, &na));
    return static_cast<size_t>(ns);
}
double AggregatList::random() {
    int32_t v = 0;
    check(mcac_gpu_rand(gpu, 1, &v));
    return mcacb::uniform_from_rand(v);
}
std::array<double, 3> AggregatList::random_direction() {  // theta first, then phi (tools.cpp:82-89)
    const double u_theta = random();
    const double u_phi = random();
    const mcacb::Vec3 v = mcacb::direction_from_draws(u_theta, u_phi);
    return {v.x, v.y, v.z};
}
void AggregatList::sort_time_steps(double factor) { check(mcac_gpu_sort_time_steps(gpu, factor)); }
double AggregatList::get_time_step(double max) const { return max / last_cum_total; }
size_t AggregatList::pick_random() {
    int64_t label = 0;
    double dt = 0.;
    const double u = random();
    check(mcac_gpu_pick_random(gpu, u, &label, &dt));
    last_cum_total = max_time_step / dt;
    return static_cast<size_t>(label);
}
size_t AggregatList::pick_last() {
    int64_t label = 0;
    check(mcac_gpu_pick_last(gpu, &label));
    return static_cast<size_t>(label);
}
void AggregatList::refresh() { check(mcac_gpu_refresh(gpu, &max_time_step, &avg_npp, &total_volume, &total_surface)); }
void AggregatList::duplication() {
    check(mcac_gpu_duplicate(gpu));
    physicalmodel->box_length *= 2;
    physicalmodel->n_monomeres *= 8;
    physicalmodel->box_volume = std::pow(physicalmodel->box_length, 3);
}
AggregateContactInfo AggregatList::distance_to_next_contact(size_t source, const std::array<double, 3> &direction, double distance) const {
    mcac_contact c;
    check(mcac_gpu_contact_search(gpu, static_cast<int64_t>(source), direction.data(), distance, &c));
    AggregateContactInfo info;
    info.distance = c.distance;
    info.moving_sphere = c.moving_sphere; info.other_sphere = c.other_sphere;
    info.moving_aggregate = c.moving_label; info.other_aggregate = c.other_label;
    return info;
}
bool AggregatList::merge(AggregateContactInfo ci) {
    mcac_contact c{ci.distance, ci.moving_sphere, ci.other_sphere, ci.moving_aggregate, ci.other_aggregate};
    int merged = 0;
    check(mcac_gpu_merge(gpu, &c, &merged));
    return merged != 0;
}
bool AggregatList::croissance_surface(double dt) { check(mcac_gpu_grow(gpu, dt, -1)); return false; }
bool AggregatList::croissance_surface(double dt, size_t index) { check(mcac_gpu_grow(gpu, dt, static_cast<int64_t>(index))); return false; }
void AggregatList::translate(size_t label, const std::array<double, 3> &v) { check(mcac_gpu_translate(gpu, static_cast<int64_t>(label), v.data())); }
void AggregatList::update(long label) { check(mcac_gpu_update(gpu, label, 1)); }
void AggregatList::update_partial(long label) { check(mcac_gpu_update(gpu, label, 0)); }
mcac_run_report AggregatList::run(long max_steps, int batch) {
    mcac_run_report rep{};
    check(mcac_gpu_run(gpu, max_steps, batch, nullptr, 0, &rep));
    avg_npp = rep.avg_npp;
    max_time_step = rep.max_time_step;
    physicalmodel->time = rep.time;
    physicalmodel->box_length = rep.box_length;
    physicalmodel->box_volume = std::pow(rep.box_length, 3);
    physicalmodel->volume_fraction = rep.volume_fraction;
    physicalmodel->aggregate_concentration = static_cast<double>(rep.n_aggregates) / physicalmodel->box_volume;
    physicalmodel->monomer_concentration = static_cast<double>(rep.n_spheres) / physicalmodel->box_volume;
    return rep;
}

// advancement.dat, 9 space-separated columns (src/calcul.cpp:29-43)
static void save_advancement(const PhysicalModel &pm, const AggregatList &aggregates, const std::string &dir) {
    std::ofstream out(dir + "/advancement.dat", std::ios_base::app);
    out << pm.time << " " << pm.aggregate_concentration << " " << pm.volume_fraction << " " << aggregates.get_avg_npp() << " " << pm.temperature
        << " " << pm.box_volume << " " << pm.monomer_concentration << " " << pm.u_sg << " " << pm.flux_nucleation << std::endl;
}

// mcac::calcul (src/calcul.cpp:55-290): the loop itself runs on the device (mcac_gpu_run); the host reports progress between
// chunks in the layout of the reference's stdout table (:237-270) and appends advancement.dat rows.
void calcul(PhysicalModel &pm, AggregatList &aggregates) {
    pm.print();
    const std::string dir = pm.output_dir.empty() ? "." : pm.output_dir;
    std::cout << std::setw(8) << "#" << " | " << std::setw(9) << "Npp_avg" << " | " << std::setw(8) << "NAgg" << " | " << std::setw(10) << "Time"
              << " | " << std::setw(10) << "steps/s" << " | " << std::setw(10) << "MC steps" << std::endl;
    long long total_events = 0, total_steps = 0;
    save_advancement(pm, aggregates, dir);
    while (true) {
        const mcac_run_report rep = aggregates.run(200000);
        total_events += rep.events;
        total_steps += rep.steps;
        save_advancement(pm, aggregates, dir);
        std::cout.precision(3);
        std::cout << std::scientific << std::setw(8) << total_events << " | " << std::setw(8) << rep.avg_npp << " | " << std::setw(8)
                  << rep.n_aggregates << " | " << std::setw(8) << rep.time << "s | " << std::setw(10)
                  << (rep.device_ms > 0 ? 1e3 * rep.steps / rep.device_ms : 0.) << " | " << std::setw(10) << total_steps << std::endl;
        if (rep.finished || rep.steps == 0) break;
    }
    std::cout << " Final residence time=" << std::setw(4) << pm.time << "s" << std::endl;
    std::cout << "Final number of aggregates : " << aggregates.size() << std::endl;
    std::cout << "\nThe End\n" << std::endl;
}
}  // namespace mcac
