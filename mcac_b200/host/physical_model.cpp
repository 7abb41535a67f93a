// mcac_b200 host layer — PhysicalModel: .ini reader, derived constants, stop rules.
// Mirrors src/physical_model/physical_model.cpp of the reference (cited per function); physics closures come
// from csrc/mcac_math.cuh so the host placement and the device kernels evaluate the same expressions.
#include "physical_model.hpp"

#include <unistd.h>

#include <cmath>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "../csrc/mcac_math.cuh"

namespace mcac {
namespace {
std::string strip(const std::string &s) {
    const char *ws = " \t\r\n";
    const size_t b = s.find_first_not_of(ws);
    if (b == std::string::npos) return {};
    return s.substr(b, s.find_last_not_of(ws) - b + 1);
}
// [section] / key = value / ';' or '#' comments; first occurrence of a key wins (inipp semantics)
using IniMap = std::map<std::string, std::map<std::string, std::string>>;
IniMap read_ini(std::istream &in) {
    IniMap out;
    std::string raw, section;
    while (std::getline(in, raw)) {
        const std::string line = strip(raw);
        if (line.empty() || line[0] == ';' || line[0] == '#') continue;
        if (line[0] == '[') {
            const size_t close = line.find(']');
            if (close != std::string::npos) section = strip(line.substr(1, close - 1));
            continue;
        }
        const size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        out[section].emplace(strip(line.substr(0, eq)), strip(line.substr(eq + 1)));
    }
    return out;
}
// inipp::extract: whole-token conversion (boolalpha) or leave the default untouched
template <class T>
void take(IniMap &ini, const char *section, const char *key, T &dst) {
    std::istringstream is(ini[section][key]);
    T v;
    char extra;
    if ((is >> std::boolalpha >> v) && !(is >> extra)) dst = v;
}
std::string take_string(IniMap &ini, const char *section, const char *key) { return ini[section][key]; }
// inipp::Ini::generate: "[section]" / "key=value" lines in map order, a blank line after every section.  Keys that were looked up
// but are absent from the file appear with an empty value, exactly as in the reference (extract() goes through operator[]).
std::string generate_ini(const IniMap &ini) {
    std::ostringstream os;
    for (const auto &sec : ini) {
        os << '[' << sec.first << ']' << std::endl;
        for (const auto &kv : sec.second) os << kv.first << '=' << kv.second << std::endl;
        os << std::endl;
    }
    return os.str();
}
// src/tools/tools.cpp:28-40 (http://www.concentric.net/~Ttwang/tech/inthash.htm): the clock / pid hash behind random_seed < 0
unsigned long mix(unsigned long a, unsigned long b, unsigned long c) {
    a = a - b; a = a - c; a = a ^ (c >> 13);
    b = b - c; b = b - a; b = b ^ (a << 8);
    c = c - a; c = c - b; c = c ^ (b >> 13);
    a = a - b; a = a - c; a = a ^ (c >> 12);
    b = b - c; b = b - a; b = b ^ (a << 16);
    c = c - a; c = c - b; c = c ^ (b >> 5);
    a = a - b; a = a - c; a = a ^ (c >> 3);
    b = b - c; b = b - a; b = b ^ (a << 10);
    c = c - a; c = c - b; c = c ^ (b >> 15);
    return c;
}
mcacb::Gas gas_of(const PhysicalModel &p) {
    return {p.gaz_mean_free_path, p.viscosity, p.temperature, p.fractal_dimension, p.density, p.with_maturity ? 1 : 0};
}
}  // namespace

PhysicalModel::PhysicalModel(const std::string &ini_file) {
    std::ifstream f(ini_file);
    if (!f) throw InputError("File does not exist");  // extract_path(), physical_model.cpp:618-626
    parse(f);
}

// physical_model.cpp:101-187 (keys), :228-269 (derived constants)
void PhysicalModel::parse(std::istream &in) {
    IniMap ini = read_ini(in);
    take(ini, "monomers", "number", n_monomeres);
    take(ini, "monomers", "density", density);
    take(ini, "monomers", "dispersion_diameter", dispersion_diameter);
    take(ini, "monomers", "mean_diameter", mean_diameter);
    std::string word = take_string(ini, "monomers", "initialisation_mode");
    if (!word.empty()) {
        monomeres_initialisation_type = word == "lognormal" ? LOG_NORMAL_INITIALISATION : word == "normal" ? NORMAL_INITIALISATION : INVALID_INITIALISATION;
        if (monomeres_initialisation_type == INVALID_INITIALISATION) throw InputError("Monomere initialisation mode unknown: " + word);
    }
    take(ini, "environment", "initial_time", time);
    take(ini, "environment", "volume_fraction", volume_fraction);
    take(ini, "environment", "temperature", temperature);
    take(ini, "environment", "pressure", pressure);
    take(ini, "environment", "fractal_prefactor", fractal_prefactor);
    take(ini, "environment", "fractal_dimension", fractal_dimension);
    take(ini, "surface_growth", "with_surface_reactions", with_surface_reactions);
    take(ini, "surface_growth", "flux_surfgrowth", flux_surfgrowth);
    word = take_string(ini, "surface_growth", "volsurf_method");
    if (!word.empty()) {
        static const char *names[] = {"caps", "sbl", "arvo", "alphas", "none"};
        volsurf_method = INVALID_VOLSURF_METHOD;
        for (int i = 0; i < 5; i++) if (word == names[i]) volsurf_method = static_cast<VolSurfMethods>(i);
        if (volsurf_method == INVALID_VOLSURF_METHOD) throw InputError("Invalid method to calculate Vols/Surf: " + word);
    }
    take(ini, "surface_growth", "full_aggregate_update_frequency", full_aggregate_update_frequency);
    take(ini, "oxidation", "rp_min", rp_min_oxid);
    mean_diameter_nucleation = mean_diameter;
    dispersion_diameter_nucleation = dispersion_diameter;
    mass_nuclei = (mcacb::pi() / 6.) * std::pow(mean_diameter_nucleation * (1e-09), 3) * density *
                  std::exp(std::pow(4.5 * std::log(dispersion_diameter_nucleation), 2));
    take(ini, "nucleation", "with_nucleation", with_nucleation);
    take(ini, "nucleation", "flux", flux_nucleation);
    take(ini, "nucleation", "mean_diameter", mean_diameter_nucleation);
    take(ini, "nucleation", "dispersion_diameter", dispersion_diameter_nucleation);
    take(ini, "nucleation", "mass_nuclei", mass_nuclei);
    take(ini, "limits", "number_of_aggregates", number_of_aggregates_limit);
    take(ini, "limits", "n_iter_without_event", n_iter_without_event_limit);
    take(ini, "limits", "cpu", cpu_limit);
    take(ini, "limits", "cpu_event", cpu_event_limit);
    take(ini, "limits", "physical_time", physical_time_limit);
    take(ini, "limits", "mean_monomere_per_aggregate", mean_monomere_per_aggregate_limit);
    take(ini, "numerics", "with_domain_duplication", with_domain_duplication);
    take(ini, "numerics", "with_domain_reduction", with_domain_reduction);
    take(ini, "numerics", "individual_surf_reactions", individual_surf_reactions);
    take(ini, "numerics", "with_collisions", with_collisions);
    take(ini, "numerics", "enforce_volume_fraction", enforce_volume_fraction);
    take(ini, "numerics", "n_verlet_divisions", n_verlet_divisions);
    take(ini, "numerics", "random_seed", random_seed);
    // init_random (tools.cpp:41-50): a negative seed is replaced by a hash of the CPU clock, the time and the pid; srand() then takes
    // the int converted to unsigned.  The seed actually used is kept (print() reports it) so that the run can be replayed.
    if (random_seed < 0) random_seed = static_cast<int>(mix(static_cast<unsigned long>(clock()), static_cast<unsigned long>(::time(nullptr)),
                                                            static_cast<unsigned long>(getpid())));
    random_seed_used = static_cast<uint32_t>(random_seed);
    word = take_string(ini, "numerics", "pick_method");
    if (!word.empty()) {
        pick_method = word == "random" ? PICK_RANDOM : word == "last" ? PICK_LAST : INVALID_PICK_METHOD;
        if (pick_method == INVALID_PICK_METHOD) throw InputError("Invalid pick method: " + word);
    }
    {   // mcac_b200 extension; looked up without operator[] so that it does not show up in the params.ini echo
        const auto sec = ini.find("numerics");
        if (sec != ini.end()) {
            const auto kv = sec->second.find("sort_order");
            if (kv != sec->second.end() && !kv->second.empty()) {
                if (kv->second == "libstdcxx") sort_order = MCAC_ORDER_LIBSTDCXX;
                else if (kv->second == "stable") sort_order = MCAC_ORDER_STABLE;
                else throw InputError("Invalid sort_order: " + kv->second);
            }
        }
    }
    take(ini, "inter_potential", "with_potentials", with_potentials);
    take(ini, "inter_potential", "with_electric_charges", with_electric_charges);
    take(ini, "inter_potential", "with_external_potentials", with_external_potentials);
    take(ini, "inter_potential", "with_dynamic_random_charges", with_dynamic_random_charges);
    interpotential_file = take_string(ini, "inter_potential", "interpotential_file");
    take(ini, "inter_potential", "with_maturity", with_maturity);
    take(ini, "flame_coupling", "with_flame_coupling", with_flame_coupling);
    { const std::string ff = take_string(ini, "flame_coupling", "flame_file"); if (!ff.empty()) flame_file = ff; }
    { const std::string od = take_string(ini, "output", "output_dir"); if (!od.empty()) output_dir = od; }
    take(ini, "output", "n_time_per_file", n_time_per_file);
    take(ini, "output", "write_between_event_frequency", write_between_event_frequency);
    take(ini, "output", "write_events_frequency", write_events_frequency);
    take(ini, "output", "write_Delta_t", write_Delta_t);
    ini_echo = generate_ini(ini);  // what the reference writes to <output_dir>/params.ini (physical_model.cpp:271-272)
    // Options of the reference that this path does not implement are refused here: a run that silently ignored them would finish
    // with a trajectory that differs from the reference's and no diagnostic.
    if (with_flame_coupling) throw InputError("flame coupling is not part of the mcac_b200 hot path");
    if (with_electric_charges) throw InputError("with_electric_charges is not built in mcac_b200 (initial / merged aggregate charges)");
    if (with_dynamic_random_charges)
        throw InputError("with_dynamic_random_charges is not built in mcac_b200 (the reference draws a charge inside every merge)");
    if (with_domain_reduction) throw InputError("with_domain_reduction is not built in mcac_b200 (AggregatList::reduction)");
    if (with_surface_reactions && flux_surfgrowth < 0.)
        throw InputError("flux_surfgrowth < 0 (oxidation: sphere removal and AggregatList::split) is not built in mcac_b200");
    if (volsurf_method == EXACT_SBL || volsurf_method == EXACT_ARVO)
        throw InputError("volsurf_method sbl / arvo are not built in mcac_b200 (use caps, alphas or none)");

    const double n = static_cast<double>(n_monomeres);
    double tot_volume_pp = 0., tot_surface_pp = 0.;
    if (monomeres_initialisation_type == NORMAL_INITIALISATION) {  // :229-245
        const double rel = dispersion_diameter / mean_diameter;
        box_length = mean_diameter * 1E-9 * std::pow(n * mcacb::pi() / 6. / volume_fraction * (1. + 3. * std::pow(rel, 2)), 1. / 3.);
        mean_massic_radius = 0.5 * 1E-9 *
                             (std::pow(mean_diameter, 4) + 6 * std::pow(mean_diameter, 2) * std::pow(dispersion_diameter, 2) +
                              3 * std::pow(dispersion_diameter, 4)) /
                             (std::pow(mean_diameter, 3) + 3 * mean_diameter * std::pow(dispersion_diameter, 2));
        const double mr = 0.5 * mean_diameter * 1E-9, dr = 0.5 * dispersion_diameter * 1E-9;
        tot_volume_pp = n * (4.0 * mcacb::pi() / 3.0) * (mr) * (std::pow(mr, 2) + 3.0 * std::pow(dr, 2));
        tot_surface_pp = n * (4.0 * mcacb::pi()) * (std::pow(mr, 2) + std::pow(dr, 2));
    } else {  // :246-258
        const double ln_s = std::log(dispersion_diameter);
        box_length = mean_diameter * 1E-9 * std::pow(n * mcacb::pi() / 6. / volume_fraction * std::exp(9. / 2. * std::pow(ln_s, 2)), 1. / 3.);
        mean_massic_radius = 0.5 * mean_diameter * 1E-9 * std::exp(1.5 * std::pow(ln_s, 2));
        const double mr = 0.5 * mean_diameter * 1E-9;
        tot_volume_pp = n * (4.0 * mcacb::pi() / 3.0) * std::pow(mr, 3) * std::exp(4.5 * std::pow(ln_s, 2));
        tot_surface_pp = n * (4.0 * mcacb::pi()) * std::pow(mr, 2) * std::exp(2 * std::pow(ln_s, 2));
    }
    box_volume = std::pow(box_length, 3);
    update_temperature(temperature);
    u_sg = flux_surfgrowth / density;
    aggregate_concentration = n / box_volume;
    monomer_concentration = aggregate_concentration;
    total_surface_concent = tot_surface_pp / box_volume;
    total_volume_concent = tot_volume_pp / box_volume;
}

// physical_model.cpp:536-546
void PhysicalModel::update_temperature(double t) noexcept {
    temperature = t;
    viscosity = 18.203E-6 * (110 + 293.15) / (110 + temperature) * std::pow(temperature / 293.15, 1.5);
    gaz_mean_free_path = 66.5E-9 * (101300 / pressure) * (temperature / 293.15) * (1. + 110 / 293.15) / (1. + 110 / temperature);
    friction_exponnant = 0.689 * (1. + std::erf(((gaz_mean_free_path / mean_massic_radius) + 4.454) / 10.628));
}
double PhysicalModel::cunningham(double r) const { return mcacb::cunningham(gas_of(*this), r); }
double PhysicalModel::friction_exponent(double r) const { return mcacb::friction_exponent(gas_of(*this), r); }
double PhysicalModel::friction_coeff(double V, double v, double r) const { return mcacb::friction_coeff(gas_of(*this), V, v, r); }
double PhysicalModel::diffusivity(double f_agg) const { return mcacb::kBoltzmann * temperature / f_agg; }
double PhysicalModel::relax_time(double masse, double f_agg) { return masse / f_agg; }
double PhysicalModel::mobility_diameter(double V, double v, double r) const { return mcacb::mobility_diameter(gas_of(*this), V, v, r); }
double PhysicalModel::grow(double r, double dt) const { return r + u_sg * dt; }

// physical_model.cpp:489-498
void PhysicalModel::update(size_t n_aggregates, size_t n_monomers, double total_volume, double total_surface) noexcept {
    total_volume_concent = total_volume / box_volume;
    total_surface_concent = total_surface / box_volume;
    aggregate_concentration = static_cast<double>(n_aggregates) / box_volume;
    monomer_concentration = static_cast<double>(n_monomers) / box_volume;
    volume_fraction = total_volume / box_volume;
}
// physical_model.cpp:288-337, same order of tests and same messages.  The state-dependent rules are also evaluated by the device
// loop at every step; the STOPCODE file and the CPU clocks can only be seen by the host, which calls this at the top of every
// slice of calcul() (after every event and at least every write_between_event_frequency steps).
bool PhysicalModel::finished(size_t n_agg, double avg_npp) const {
    if (!output_dir.empty() && std::filesystem::exists(std::filesystem::path(output_dir) / "STOPCODE")) {
        std::cout << "STOPCODE" << std::endl << std::endl;
        return true;
    }
    if (n_agg < 1) { std::cout << "All the aggregates disappeared" << std::endl << std::endl; return true; }
    if (n_agg <= number_of_aggregates_limit) { std::cout << "We reach the AggMin condition" << std::endl << std::endl; return true; }
    if (n_iter_without_event_limit > 0 && n_iter_without_event >= static_cast<size_t>(n_iter_without_event_limit)) {
        std::cout << "We reach the WaitLimit condition" << std::endl << std::endl;
        return true;
    }
    if (cpu_limit > 0) {
        const double elapse = double(clock() - cpu_start) / CLOCKS_PER_SEC;
        if (elapse >= cpu_limit) { std::cout << "We reach the CPULimit condition" << std::endl << std::endl; return true; }
    }
    if (cpu_event_limit > 0) {
        const double elapse = double(clock() - cpu_last_event) / CLOCKS_PER_SEC;
        if (elapse >= cpu_event_limit) { std::cout << "We reach the CPULimitEvent condition" << std::endl << std::endl; return true; }
    }
    if (physical_time_limit > 0 && time >= physical_time_limit) {
        std::cout << "We reach the Maximum physical time condition " << time << "/" << physical_time_limit << std::endl;
        return true;
    }
    if (mean_monomere_per_aggregate_limit > 0 && avg_npp >= mean_monomere_per_aggregate_limit) {
        std::cout << "We reach the NPP_avg_limit condition " << avg_npp << "/" << mean_monomere_per_aggregate_limit << std::endl;
        return true;
    }
    return false;
}
void PhysicalModel::print() const {
    std::cout << "PARTICLES PROPERTIES:\n density  : " << density << " (kg/m^3)\n Dpm      : " << mean_diameter << " (nm)\n sigmaDpm : "
              << dispersion_diameter << "\n dfe      : " << fractal_dimension << "\n kfe      : " << fractal_prefactor << "\n rpeqmass : "
              << mean_massic_radius << " (m)\n gamma_   : " << friction_exponnant << "\nFLUID PROPERTIES:\n Pressure    : " << pressure
              << " (Pa)\n Temperature : " << temperature << " (K)\n viscosity   : " << viscosity << " (kg/m*s)\n lambda      : "
              << gaz_mean_free_path << " (m)\nSIMULATION OPTIONS:\n Initial Nagg : " << n_monomeres << "\n Box size     : " << box_length
              << " (m)\n FV           : " << volume_fraction << "\n Seed         : " << random_seed << std::endl;
}

mcac_params PhysicalModel::to_params() const {
    mcac_params p{};
    p.box_length = box_length;
    p.time = time;
    p.temperature = temperature;
    p.pressure = pressure;
    p.viscosity = viscosity;
    p.gaz_mean_free_path = gaz_mean_free_path;
    p.density = density;
    p.fractal_dimension = fractal_dimension;
    p.u_sg = u_sg;
    p.rp_min_oxid = rp_min_oxid;
    p.flux_nucleation = flux_nucleation;
    p.nucleation_accum = nucleation_accum;
    p.box_volume = box_volume;
    p.mean_diameter_nucleation = mean_diameter_nucleation;
    p.dispersion_diameter_nucleation = dispersion_diameter_nucleation;
    p.normal_initialisation = monomeres_initialisation_type == NORMAL_INITIALISATION ? 1 : 0;
    p.physical_time_limit = physical_time_limit;
    p.number_of_aggregates_limit = static_cast<int64_t>(number_of_aggregates_limit);
    p.n_iter_without_event_limit = n_iter_without_event_limit;
    p.mean_monomere_per_aggregate_limit = mean_monomere_per_aggregate_limit;
    p.n_monomeres = static_cast<int64_t>(n_monomeres);
    p.full_aggregate_update_frequency = static_cast<int64_t>(full_aggregate_update_frequency);
    p.n_verlet_divisions = static_cast<int32_t>(n_verlet_divisions);
    p.pick_method = pick_method == PICK_LAST ? MCAC_PICK_LAST : MCAC_PICK_RANDOM;
    p.volsurf_method = static_cast<int32_t>(volsurf_method);
    p.with_collisions = with_collisions;
    p.with_surface_reactions = with_surface_reactions;
    p.individual_surf_reactions = individual_surf_reactions;
    p.with_domain_duplication = with_domain_duplication;
    p.with_maturity = with_maturity;
    p.with_potentials = with_potentials;
    p.with_external_potentials = with_external_potentials;
    p.with_nucleation = with_nucleation;
    p.with_dynamic_random_charges = with_dynamic_random_charges;
    p.sort_order = sort_order;
    p.random_seed = random_seed_used;
    return p;
}

// src/io/physical_model.cpp:30-49 + include/io/format.hpp: values printed through operator<< (6 significant digits);
// pinned by pymcac/tests/test_read.py:31-48
std::vector<std::pair<std::string, std::string>> PhysicalModel::xmf_write() const {
    auto fmt = [](double v) { std::ostringstream o; o << v; return o.str(); };
    return {{"flux_surfgrowth", fmt(flux_surfgrowth)}, {"u_sg", fmt(u_sg)}, {"dfe", fmt(fractal_dimension)}, {"kfe", fmt(fractal_prefactor)},
            {"lambda", fmt(gaz_mean_free_path)}, {"rpeqmass", fmt(mean_massic_radius)}, {"gamma_", fmt(friction_exponnant)},
            {"P [Pa]", fmt(pressure)}, {"T [K]", fmt(temperature)}, {"Mu", fmt(viscosity)}, {"Rho [kg/m3]", fmt(density)},
            {"Dpm [nm]", fmt(mean_diameter)}, {"sigmaDpm [nm]", fmt(dispersion_diameter)}, {"FV [ppt]", fmt(volume_fraction)},
            {"L", fmt(box_length)}, {"N []", std::to_string(static_cast<int>(n_monomeres))}};
}
std::map<std::string, std::string> PhysicalModel::golden_metadata() const {
    std::map<std::string, std::string> out;
    for (const auto &kv : xmf_write()) out[kv.first] = kv.second;
    return out;
}
}  // namespace mcac
