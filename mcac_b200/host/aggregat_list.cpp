// mcac_b200 host layer — AggregatList facade + calcul() (see aggregat_list.hpp).
#include "aggregat_list.hpp"

#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>

#include <cstdlib>
#include <memory>

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
std::array<double, 21> AggregatList::fields(size_t label, size_t *n_spheres) const {
    std::array<double, 21> f{};
    int64_t n = 0;
    check(mcac_gpu_aggregate_fields(gpu, static_cast<int64_t>(label), f.data(), &n));
    if (n_spheres) *n_spheres = static_cast<size_t>(n);
    return f;
}
void Aggregate::translate(const std::array<double, 3> &vector) { list->translate(label, vector); }
void Aggregate::update() { list->update(static_cast<long>(label)); }
void Aggregate::update_partial() { list->update_partial(static_cast<long>(label)); }
double Aggregate::get_lpm() const { return list->fields(label)[2]; }
double Aggregate::get_time_step() const { return list->fields(label)[3]; }
double Aggregate::get_rg() const { return list->fields(label)[0]; }
size_t Aggregate::size() const { size_t n = 0; list->fields(label, &n); return n; }
size_t SphereList::size() const { return list->n_spheres(); }
mcac_run_report AggregatList::run(long max_steps, int batch) {
    mcac_run_report rep{};
    check(mcac_gpu_run(gpu, max_steps, batch, nullptr, 0, &rep));
    avg_npp = rep.avg_npp;
    max_time_step = rep.max_time_step;
    physicalmodel->time = rep.time;
    physicalmodel->box_length = rep.box_length;
    physicalmodel->box_volume = std::pow(rep.box_length, 3);
    // PhysicalModel::update runs after an event, after a duplication, and every step with surface reactions (calcul.cpp:272-277,
    // aggregat_list.cpp:186-189): only then do the concentrations and the volume fraction change
    if (rep.events > 0 || rep.nucleated > 0 || rep.duplications > 0 || physicalmodel->with_surface_reactions) {
        physicalmodel->volume_fraction = rep.volume_fraction;
        physicalmodel->aggregate_concentration = static_cast<double>(rep.n_aggregates) / physicalmodel->box_volume;
        physicalmodel->monomer_concentration = static_cast<double>(rep.n_spheres) / physicalmodel->box_volume;
        total_volume = rep.total_volume;
        total_surface = rep.total_surface;
    }
    return rep;
}

// advancement.dat, 9 space-separated columns (src/calcul.cpp:29-43)
static void save_advancement(const PhysicalModel &pm, const AggregatList &aggregates, const std::string &dir) {
    std::ofstream out(dir + "/advancement.dat", std::ios_base::app);
    out << pm.time << " " << pm.aggregate_concentration << " " << pm.volume_fraction << " " << aggregates.get_avg_npp() << " " << pm.temperature
        << " " << pm.box_volume << " " << pm.monomer_concentration << " " << pm.u_sg << " " << pm.flux_nucleation << std::endl;
}
static void print_bool(bool b, int width) {  // src/calcul.cpp:44-51
    if (b) std::cout << std::setw(width / 2 + width % 2) << "X" << std::setw(width / 2 + 3) << " | ";
    else std::cout << std::setw(width + 3) << " | ";
}
// PhysicalModel::time_to_write (physical_model.cpp:338-356)
static bool time_to_write(const PhysicalModel &pm, size_t total_events, size_t n_iter_without_event, size_t &last_timestep_written) {
    if (pm.write_Delta_t > 0) {
        const size_t timestep = static_cast<size_t>(std::floor(pm.time / pm.write_Delta_t));
        if (timestep > last_timestep_written) {
            last_timestep_written = timestep;
            return true;
        }
    }
    if (n_iter_without_event == 0 && pm.write_events_frequency > 0 && total_events % pm.write_events_frequency == 0) return true;
    if (n_iter_without_event > 0 && pm.write_between_event_frequency > 0 && n_iter_without_event % pm.write_between_event_frequency == 0)
        return true;
    return false;
}

// mcac::calcul (src/calcul.cpp:55-290).  The MC steps run on the device (mcac_gpu_run); the host does what calcul() does around
// them, at the same points of the loop: PhysicalModel::finished, time_to_write -> advancement.dat row (same 9 columns, same stream
// formatting), the "Duplication" lines and the per-event progress table.  mcac_gpu_run is called in slices that end where the
// reference's loop top could write: right after an event, or when n_iter_without_event reaches a multiple of
// write_between_event_frequency (every step if write_Delta_t is set).
void calcul(PhysicalModel &pm, AggregatList &aggregates) {
    pm.print();
    const std::string dir = pm.output_dir.empty() ? "." : pm.output_dir;
    mcac_gpu *gpu = aggregates.handle();
    mcac_gpu_set_stop_at_event(gpu, 1);
    size_t total_events = 0, n_iter = 0, last_timestep_written = 0;
    long long total_steps = 0;
    long long n_sph = static_cast<long long>(aggregates.n_spheres()), n_agg = static_cast<long long>(aggregates.size());
    pm.cpu_last_event = pm.cpu_start = clock();  // physical_model.cpp:286
    // the two output series of the reference (SphereList / AggregatList writers: <output_dir>/Spheres_<k>.{h5,xmf}, Aggregats_<k>...);
    // MCAC_B200_NO_HEAVY_OUTPUT=1 keeps advancement.dat only (what the reference does when built without WITH_HDF5)
    struct IoCloser { void operator()(mcac_io_writer *w) const { mcac_io_writer_destroy(w); } };
    std::unique_ptr<mcac_io_writer, IoCloser> io_spheres, io_aggregats;
    if (!std::getenv("MCAC_B200_NO_HEAVY_OUTPUT")) {
        std::string physics;
        for (const auto &kv : pm.xmf_write()) physics += kv.first + "=" + kv.second + "\n";
        mcac_io_writer *ws = nullptr, *wa = nullptr;
        const int64_t n0 = static_cast<int64_t>(pm.n_monomeres);
        if (mcac_io_writer_create((dir + "/Spheres").c_str(), "Spheres", static_cast<int64_t>(pm.n_time_per_file), n0, physics.c_str(), &ws) ||
            mcac_io_writer_create((dir + "/Aggregats").c_str(), "Aggregats", static_cast<int64_t>(pm.n_time_per_file), n0, physics.c_str(), &wa)) {
            if (ws) mcac_io_writer_destroy(ws);
            throw IOError(mcac_host_last_error());
        }
        io_spheres.reset(ws);
        io_aggregats.reset(wa);
    }
    auto save_state = [&]() {  // aggregates.spheres.save(); aggregates.save();  (calcul.cpp:68-69, 283-284)
        if (!io_spheres) return;
        const int rc = mcac_gpu_save(gpu, io_spheres.get(), io_aggregats.get());
        if (rc) throw BaseException(static_cast<ErrorCodes>(rc), mcac_host_last_error());
    };
    const clock_t cpu_start = pm.cpu_start;
    bool first = true;
    while (true) {
        // loop top of calcul(): PhysicalModel::finished.  The state-dependent rules are also enforced inside mcac_gpu_run at every
        // step (a slice stops where the reference's loop would); STOPCODE and the cpu / cpu_event clocks are host-only and are
        // seen here, i.e. after every event and at least every write_between_event_frequency steps.
        pm.n_iter_without_event = n_iter;
        if (pm.finished(static_cast<size_t>(n_agg), aggregates.get_avg_npp())) break;
        if (!first && pm.finished_flag) break;
        if (time_to_write(pm, total_events, n_iter, last_timestep_written)) {
            save_state();
            save_advancement(pm, aggregates, dir);
        }
        first = false;
        long slice = 1;
        if (!(pm.write_Delta_t > 0)) {
            const size_t f = pm.write_between_event_frequency;
            slice = f > 0 ? static_cast<long>(f - n_iter % f) : 1000000L;
        }
        const mcac_run_report rep = aggregates.run(slice);
        if (rep.duplications > 0)
            std::cout << "Duplication : " << n_sph << " spheres in " << n_agg << " aggregates duplicated into " << 8 * n_sph << " spheres in "
                      << 8 * n_agg << " aggregates" << std::endl;
        const bool event = rep.events > 0 || rep.nucleated > 0;
        total_steps += rep.steps;
        pm.finished_flag = rep.finished != 0;
        if (rep.steps == 0) {
            if (!pm.finished_flag) throw DeviceError("calcul: the device loop made no progress");
            break;
        }
        if (event) {
            total_events++;
            pm.cpu_last_event = clock();  // calcul.cpp:238
            if (total_events % 20 == 1)
                std::cout << std::setw(8) << "#" << " | " << std::setw(9) << "Npp_avg" << " | " << std::setw(8) << "NAgg" << " | " << std::setw(10)
                          << "Time" << " | " << std::setw(10) << "CPU" << " | " << std::setw(7) << "contact" << " | " << std::setw(5) << "merge"
                          << " | " << std::setw(5) << "split" << " | " << std::setw(9) << "disappear" << " | " << std::setw(10) << "nucleation"
                          << " | " << std::setw(8) << "after" << std::endl;
            const double elapse = double(clock() - cpu_start) / CLOCKS_PER_SEC;
            std::cout.precision(3);
            std::cout << std::scientific;
            std::cout << std::setw(8) << total_events << " | " << std::setw(8) << rep.avg_npp << " | " << std::setw(8) << rep.n_aggregates << " | "
                      << std::setw(8) << rep.time << "s" << " | " << std::setw(8) << elapse << "s" << " | ";
            print_bool(rep.events > 0, 7);  // contact (a contact that is kept is a merge on this path)
            print_bool(rep.events > 0, 5);  // merge
            print_bool(false, 5);           // split     (u_sg < 0 only: out of scope)
            print_bool(false, 9);           // disappear (oxidation only: out of scope)
            std::cout << std::setw(10) << rep.nucleated << " | " << std::setw(8) << (n_iter + static_cast<size_t>(rep.steps) - 1) << std::endl;
            std::cout.unsetf(std::ios_base::floatfield);
            std::cout.precision(6);
        }
        n_iter = static_cast<size_t>(rep.n_iter_without_event);
        n_sph = rep.n_spheres;
        n_agg = rep.n_aggregates;
    }
    save_advancement(pm, aggregates, dir);
    save_state();
    io_spheres.reset();   // ~ThreadedIO: the file in progress is written
    io_aggregats.reset();
    std::cout << " Final residence time=" << std::setw(4) << pm.time << "s" << std::endl;
    std::cout << "Final number of aggregates : " << aggregates.size() << std::endl;
    std::cout << "Output files saved on: \"" << dir << "\"" << std::endl;
    std::cout << std::endl;
    std::cout << "\nThe End\n" << std::endl;
    (void)total_steps;
}
}  // namespace mcac
