// ============================================================================================
// TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the per-step Monte-Carlo aggregation
// hot path of giraldeau/MCAC.  Nothing under mcac_b200/ may include, link or call this file; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement bit-for-bit
// (every step's pick, direction, contact distance, contact pair, merge result, final SoA state)
// against traces of the UNMODIFIED reference binary (oracle/_ref/MCAC_tap, built by
// oracle/ref_build/Makefile from /root/reference) committed under tests/golden/, and against the
// golden physics constants of pymcac/tests/test_read.py:31-48.
//
// Each function cites the reference file:line it restates.  Data layout is deliberately flat
// (index == label) but the *operation order* of every floating-point expression, every container
// whose iteration order is observable (std::set cells, std::multimap suspects,
// std::unordered_map contact graph, std::sort) and every RNG draw follows the reference exactly.
// Build: g++ -std=c++17 -O2 -ffp-contract=off (no FMA contraction: the reference's effective
// build has none).
// ============================================================================================
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace orc {
using vec3 = std::array<double, 3>;
static inline vec3 operator+(const vec3 &a, const vec3 &b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
static inline vec3 operator-(const vec3 &a, const vec3 &b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
static inline vec3 operator*(const vec3 &a, double f) { return {a[0] * f, a[1] * f, a[2] * f}; }
static inline vec3 operator*(double f, const vec3 &a) { return {f * a[0], f * a[1], f * a[2]}; }

// include/constants.hpp:112-127
static const double PI = std::atan(1.0) * 4;
static const double VOLUME_FACTOR = 4 * PI / 3;
static const double SURFACE_FACTOR = 4 * PI;
static const double BOLTZMANN = 1.38066E-23;
static const double FLUID_MFP_REF = 66.5E-9;
static const double TEMPERATURE_REF = 293.15;
static const double SUTHERLAND = 110;
static const double PRESSURE_REF = 101300;
static const double VISCOSITY_REF = 18.203E-6;
static const double KE_E2 = 9.0e+09 * std::pow(1.60217662e-19, 2);
static const double CH_MATURE = 10, CH_YOUNG = 1.1, RHO_MATURE = 1800, RHO_YOUNG = 1200;
static const double CONTACT_EPSILON = 1e-28;
static const double COORDINATION_EPSILON = 1e-10;

enum ErrorCodes { NO_ERROR, UNKNOWN_ERROR, IO_ERROR, VERLET_ERROR, INPUT_ERROR, ABANDON_ERROR, TOO_DENSE_ERROR, SBL_ERROR,
                  VOL_SURF_ERROR, MERGE_ERROR, ARVO_ERROR, INTERPOTENTIAL_ERROR };  // constants.hpp:70-83
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
enum Pick { PICK_RANDOM, PICK_LAST };
enum InitMode { LOG_NORMAL, NORMAL };
enum VolSurf { CAPS, SBL, ARVO, ALPHAS, NONE };  // constants.hpp:103-110 (order matters for parsing only)
enum Regime { STICKING, REPULSION, BOUNCING };

// ---------------------------------------------------------------------------------------------
// glibc rand()/srand() TYPE_3 additive feedback generator — SURVEY.md Appendix B; called through
// src/tools/tools.cpp:41-55 (`random()` = rand()/RAND_MAX, inclusive of 1.0).
// ---------------------------------------------------------------------------------------------
struct GlibcRand {
    uint32_t ring[31];
    int pos = 0;
    long long calls = 0;
    void seed(unsigned int s) {
        std::vector<uint32_t> r(344);
        if (s == 0) s = 1;
        r[0] = s;
        for (int i = 1; i < 31; i++) {
            long hi = (long)((int32_t)r[i - 1]) / 127773, lo = (long)((int32_t)r[i - 1]) % 127773;
            long w = 16807 * lo - 2836 * hi;
            if (w < 0) w += 2147483647;
            r[i] = (uint32_t)w;
        }
        for (int i = 31; i < 34; i++) r[i] = r[i - 31];
        for (int i = 34; i < 344; i++) r[i] = r[i - 31] + r[i - 3];
        for (int i = 0; i < 31; i++) ring[i] = r[344 - 31 + i];
        pos = 0;
        calls = 0;
    }
    int next() {  // ring[pos] holds r[i-31], ring[(pos+28)%31] holds r[i-3]
        uint32_t v = ring[pos] + ring[(pos + 28) % 31];
        ring[pos] = v;
        pos = (pos + 1) % 31;
        calls++;
        return (int)(v >> 1);
    }
    double uniform() {  // tools.cpp:51-55
        double v = next();
        v = v / 2147483647;
        return v;
    }
};

// tools.cpp:56-81
static double inverfc(double p) {
    double x, t, pp;
    if (p >= 2.) return -100.;
    if (p <= 0.0) return 100.;
    pp = (p < 1.0) ? p : 2. - p;
    t = std::sqrt(-2. * std::log(pp / 2.));
    x = -0.70711 * ((2.30753 + t * 0.27061) / (1. + t * (0.99229 + t * 0.04481)) - t);
    for (int j = 0; j < 2; j++) {
        double err = std::erfc(x) - pp;
        x += err / (1.12837916709551257 * std::exp(-(x * x)) - x * err);
    }
    return (p < 1.0 ? x : -x);
}
static double inverf(double p) { return inverfc(1. - p); }

// include/physical_model/physical_model.hpp:99-126
static inline double periodic_distance(double dist, double dim) {
    double d(dist), half(0.5 * dim);
    while (d < -half) d += dim;
    while (d >= half) d -= dim;
    return d;
}
static inline double periodic_position(double p, double dim) {
    double q(p);
    while (q < 0) q += dim;
    while (q >= dim) q -= dim;
    return q;
}
// src/spheres/sphere_distances.cpp:68-83
static inline double distance_2(const vec3 &p1, const vec3 &p2, double box) {
    vec3 diff = p1 - p2;
    double dx(periodic_distance(diff[0], box)), dy(periodic_distance(diff[1], box)), dz(periodic_distance(diff[2], box));
    return dx * dx + dy * dy + dz * dz;
}
static inline double relative_distance_2(const vec3 &p1, const vec3 &p2) {
    vec3 diff = p1 - p2;
    return diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2];
}
// src/spheres/sphere_distances.cpp:84-90 on raw (pos, r)
static inline bool contact_spheres(const vec3 &p1, double r1, const vec3 &p2, double r2, double box) {
    double distance = distance_2(p1, p2, box);
    double dist_contact = (r1 + r2) * (r1 + r2);
    return (distance - dist_contact <= CONTACT_EPSILON);
}

// ---------------------------------------------------------------------------------------------
// THE pair kernel — src/spheres/sphere_contact.cpp:47-125
// ---------------------------------------------------------------------------------------------
static double pair_distance_to_contact(const vec3 &pos1, double r1, vec3 pos2, double r2, const vec3 &dir, double dist,
                                       double box_length) {
    double dist_contact = r1 + r2;
    double dist_contact_2 = dist_contact * dist_contact;
    if (distance_2(pos1, pos2, box_length) <= dist_contact_2) return 0.;
    vec3 total_displacement = dir * dist;
    std::array<int, 3> nper{0, 0, 0};
    vec3 zone{0., 0., 0.};
    for (size_t l = 0; l < 3; ++l) {
        double base = std::min(pos1[l], pos1[l] + total_displacement[l]) - dist_contact;
        double end = std::max(pos1[l], pos1[l] + total_displacement[l]) + dist_contact;
        zone[l] = end - base;
        pos2[l] = std::fmod((pos2[l] - base), box_length);
        if (pos2[l] < 0) pos2[l] += box_length;
        pos2[l] += base;
        nper[l] = static_cast<int>(std::floor(zone[l] / box_length));
    }
    double res = std::numeric_limits<double>::infinity();
    for (int i = 0; i <= nper[0]; i++)
        for (int j = 0; j <= nper[1]; j++)
            for (int k = 0; k <= nper[2]; k++) {
                vec3 pos3{pos2[0] + i * box_length, pos2[1] + j * box_length, pos2[2] + k * box_length};
                vec3 diff = pos3 - pos1;
                if (std::abs(diff[0]) > zone[0]) continue;
                if (std::abs(diff[1]) > zone[1]) continue;
                if (std::abs(diff[2]) > zone[2]) continue;
                double proj = diff[0] * dir[0] + diff[1] * dir[1] + diff[2] * dir[2];
                if (proj < 0) continue;
                bool end_contact = relative_distance_2(pos1 + total_displacement, pos3) <= dist_contact_2;
                if ((!end_contact) && dist < proj) continue;
                vec3 cross{diff[1] * dir[2] - diff[2] * dir[1], diff[2] * dir[0] - diff[0] * dir[2],
                           diff[0] * dir[1] - diff[1] * dir[0]};
                double dist_to_axis = cross[0] * cross[0] + cross[1] * cross[1] + cross[2] * cross[2];
                if (dist_to_axis > dist_contact_2) continue;
                res = std::min(res, proj - std::sqrt(dist_contact_2 - dist_to_axis));
            }
    return res;
}

// ---------------------------------------------------------------------------------------------
// INI reader (stand-in for inipp, see oracle/ref_build/shim/inipp.h) + PhysicalModel
// src/physical_model/physical_model.cpp:32-287
// ---------------------------------------------------------------------------------------------
struct Ini {
    std::map<std::string, std::map<std::string, std::string>> sections;
    static std::string trim(const std::string &s) {
        size_t b = s.find_first_not_of(" \t\r\n");
        if (b == std::string::npos) return "";
        size_t e = s.find_last_not_of(" \t\r\n");
        return s.substr(b, e - b + 1);
    }
    void parse(std::istream &is) {
        std::string line, section;
        while (std::getline(is, line)) {
            line = trim(line);
            if (line.empty() || line[0] == ';' || line[0] == '#') continue;
            if (line[0] == '[') {
                size_t end = line.find(']');
                if (end != std::string::npos) section = trim(line.substr(1, end - 1));
                continue;
            }
            size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
            if (!sections[section].count(key)) sections[section][key] = val;
        }
    }
    template <class T> void get(const char *sec, const char *key, T &dst) {
        std::istringstream is(sections[sec][key]);
        T result;
        char c;
        if ((is >> std::boolalpha >> result) && !(is >> c)) dst = result;
    }
    void get(const char *sec, const char *key, std::string &dst) { dst = sections[sec][key]; }
};

// src/physical_model/physical_model_interpotential.cpp:46-197
struct Interpotential {
    int max_charge = 6, min_charge = -6;
    double fixed_T = 0;
    size_t n1 = 0, n2 = 0, nq = 0;
    std::vector<int> val_charge;
    std::vector<double> val_dp1, val_dp2, E_barr, E_well;  // [q1][q2][i][j]
    size_t at(size_t k, size_t l, size_t i, size_t j) const { return ((k * nq + l) * n1 + i) * n2 + j; }
    void load(const std::string &file) {
        std::ifstream f(file);
        if (!f) throw Error(IO_ERROR, " Interpotential file does not exist: " + file);
        double d;
        f >> fixed_T;
        f >> d; n1 = (size_t)d; val_dp1.resize(n1);
        f >> d; n2 = (size_t)d; val_dp2.resize(n2);
        f >> d; nq = (size_t)d; val_charge.resize(nq);
        E_barr.assign(nq * nq * n1 * n2, 0.);
        E_well.assign(nq * nq * n1 * n2, 0.);
        for (size_t i = 0; i < nq; i++) { f >> d; val_charge[i] = (int)d; }
        for (size_t i = 0; i < n1; i++) f >> val_dp1[i];
        for (size_t i = 0; i < n2; i++) f >> val_dp2[i];
        for (size_t i = 0; i < n1; i++)
            for (size_t j = 0; j < n2; j++)
                for (size_t k = 0; k < nq; k++)
                    for (size_t l = 0; l < nq; l++) f >> E_barr[at(k, l, i, j)] >> E_well[at(k, l, i, j)];
        min_charge = *std::min_element(val_charge.begin(), val_charge.end());
        max_charge = *std::max_element(val_charge.begin(), val_charge.end());
    }
    // tools.cpp:162-175
    static double interpolate_2d(double f11, double f12, double f21, double f22, double dx, double dy) {
        double Df_x = f21 - f11, Df_y = f12 - f11, Df_xy = (f11 + f22) - (f21 + f12);
        return Df_x * dx + Df_y * dy + Df_xy * dx * dy + f11;
    }
    std::pair<double, double> get(double dp1, double dp2, int q1, int q2) const {
        auto n1it = std::upper_bound(val_dp1.begin(), val_dp1.end(), dp1);
        if (n1it == val_dp1.begin() || n1it == val_dp1.end()) throw Error(INTERPOTENTIAL_ERROR, "PP diameter out of range");
        auto n2it = std::upper_bound(val_dp2.begin(), val_dp2.end(), dp2);
        if (n2it == val_dp2.begin() || n2it == val_dp2.end()) throw Error(INTERPOTENTIAL_ERROR, "PP diameter out of range");
        auto c1 = std::find(val_charge.begin(), val_charge.end(), q1);
        auto c2 = std::find(val_charge.begin(), val_charge.end(), q2);
        if (c1 == val_charge.end() || c2 == val_charge.end()) throw Error(INTERPOTENTIAL_ERROR, "charges out of range");
        size_t i1 = (size_t)(n1it - val_dp1.begin()), i0 = i1 - 1, j1 = (size_t)(n2it - val_dp2.begin()), j0 = j1 - 1;
        size_t k = (size_t)(c1 - val_charge.begin()), l = (size_t)(c2 - val_charge.begin());
        double ddp1 = val_dp1[i1] - val_dp1[i0], ddp2 = val_dp2[j1] - val_dp2[j0];
        double a = (dp1 - val_dp1[i0]) / ddp1, b = (dp2 - val_dp2[j0]) / ddp2;
        double eb = interpolate_2d(E_barr[at(k, l, i0, j0)], E_barr[at(k, l, i0, j1)], E_barr[at(k, l, i1, j0)],
                                   E_barr[at(k, l, i1, j1)], a, b);
        double ew = interpolate_2d(E_well[at(k, l, i0, j0)], E_well[at(k, l, i0, j1)], E_well[at(k, l, i1, j0)],
                                   E_well[at(k, l, i1, j1)], a, b);
        return {eb, ew};
    }
};

struct PhysicalModel {
    // physical_model.cpp:33-98 (constructor defaults)
    double fractal_dimension = 1.4, fractal_prefactor = 1.8, flux_surfgrowth = 0., u_sg = 0., flux_nucleation = 0.,
           nucleation_accum = 0.0, pressure = PRESSURE_REF, temperature = TEMPERATURE_REF, gaz_mean_free_path = FLUID_MFP_REF,
           viscosity = VISCOSITY_REF, density = 1800., mean_diameter = 30., dispersion_diameter = 1.0, mean_massic_radius = 0.,
           mass_nuclei = 0.0, mean_diameter_nucleation = 5.0, dispersion_diameter_nucleation = 1.0, friction_exponnant = 0.,
           time = 0., volume_fraction = 1e-3, box_length = 0., box_volume = 0., aggregate_concentration = 0.0,
           monomer_concentration = 0.0, total_surface_concent = 0.0, total_volume_concent = 0.0, rp_min_oxid = 0.166e-09;
    size_t n_verlet_divisions = 10;
    Pick pick_method = PICK_RANDOM;
    VolSurf volsurf_method = NONE;
    size_t n_monomeres = 2500, n_time_per_file = 10;
    InitMode init_mode = LOG_NORMAL;
    size_t n_iter_without_event = 0;
    double cpu_limit = -1, cpu_event_limit = -1, physical_time_limit = -1, write_Delta_t = -1;
    int mean_monomere_per_aggregate_limit = -1;
    size_t number_of_aggregates_limit = 1;
    int n_iter_without_event_limit = -1, random_seed = -1;
    size_t write_events_frequency = 1, write_between_event_frequency = 100, full_aggregate_update_frequency = 1;
    std::string interpotential_file = "interpotential_file";
    bool with_domain_duplication = true, with_domain_reduction = false, with_nucleation = false, with_collisions = true,
         with_surface_reactions = false, with_flame_coupling = false, enforce_volume_fraction = true,
         individual_surf_reactions = false, with_potentials = false, with_external_potentials = false,
         with_dynamic_random_charges = false, with_electric_charges = false, with_maturity = false;
    Interpotential intpotential;
    GlibcRand rng;

    double random() { return rng.uniform(); }
    double random_normal(double mean, double sigma) { return mean + std::sqrt(2.) * sigma * inverf(2. * random() - 1.0); }  // tools.cpp:78-81

    void load(std::istream &is, const std::string &base_dir) {
        Ini ini;
        ini.parse(is);
        std::string s;
        ini.get("monomers", "number", n_monomeres);
        ini.get("monomers", "density", density);
        ini.get("monomers", "dispersion_diameter", dispersion_diameter);
        ini.get("monomers", "mean_diameter", mean_diameter);
        ini.get("monomers", "initialisation_mode", s);
        if (s != "") {
            if (s == "lognormal") init_mode = LOG_NORMAL;
            else if (s == "normal") init_mode = NORMAL;
            else throw Error(INPUT_ERROR, "Monomere initialisation mode unknown: " + s);
        }
        ini.get("environment", "initial_time", time);
        ini.get("environment", "volume_fraction", volume_fraction);
        ini.get("environment", "temperature", temperature);
        ini.get("environment", "pressure", pressure);
        ini.get("environment", "fractal_prefactor", fractal_prefactor);
        ini.get("environment", "fractal_dimension", fractal_dimension);
        ini.get("surface_growth", "with_surface_reactions", with_surface_reactions);
        ini.get("surface_growth", "flux_surfgrowth", flux_surfgrowth);
        ini.get("surface_growth", "volsurf_method", s);
        if (s != "") {
            static const char *names[] = {"caps", "sbl", "arvo", "alphas", "none"};
            int found = -1;
            for (int i = 0; i < 5; i++) if (s == names[i]) found = i;
            if (found < 0) throw Error(INPUT_ERROR, "Invalid method to calculate Vols/Surf: " + s);
            volsurf_method = (VolSurf)found;
        }
        ini.get("surface_growth", "full_aggregate_update_frequency", full_aggregate_update_frequency);
        ini.get("oxidation", "rp_min", rp_min_oxid);
        mean_diameter_nucleation = mean_diameter;
        dispersion_diameter_nucleation = dispersion_diameter;
        mass_nuclei = (PI / 6.) * std::pow(mean_diameter_nucleation * (1e-09), 3) * density *
                      std::exp(std::pow(4.5 * std::log(dispersion_diameter_nucleation), 2));
        ini.get("nucleation", "with_nucleation", with_nucleation);
        ini.get("nucleation", "flux", flux_nucleation);
        ini.get("nucleation", "mean_diameter", mean_diameter_nucleation);
        ini.get("nucleation", "dispersion_diameter", dispersion_diameter_nucleation);
        ini.get("nucleation", "mass_nuclei", mass_nuclei);
        ini.get("limits", "number_of_aggregates", number_of_aggregates_limit);
        ini.get("limits", "n_iter_without_event", n_iter_without_event_limit);
        ini.get("limits", "cpu", cpu_limit);
        ini.get("limits", "cpu_event", cpu_event_limit);
        ini.get("limits", "physical_time", physical_time_limit);
        ini.get("limits", "mean_monomere_per_aggregate", mean_monomere_per_aggregate_limit);
        ini.get("numerics", "with_domain_duplication", with_domain_duplication);
        ini.get("numerics", "with_domain_reduction", with_domain_reduction);
        ini.get("numerics", "individual_surf_reactions", individual_surf_reactions);
        ini.get("numerics", "with_collisions", with_collisions);
        ini.get("numerics", "enforce_volume_fraction", enforce_volume_fraction);
        ini.get("numerics", "n_verlet_divisions", n_verlet_divisions);
        ini.get("numerics", "pick_method", s);
        ini.get("numerics", "random_seed", random_seed);
        if (random_seed < 0) throw Error(INPUT_ERROR, "oracle needs numerics.random_seed >= 0 (clock/pid seeds are not replayable)");
        rng.seed((unsigned)random_seed);  // tools.cpp:41-50
        if (s != "") {
            if (s == "random") pick_method = PICK_RANDOM;
            else if (s == "last") pick_method = PICK_LAST;
            else throw Error(INPUT_ERROR, "Invalid pick method: " + s);
        }
        ini.get("inter_potential", "with_potentials", with_potentials);
        ini.get("inter_potential", "with_electric_charges", with_electric_charges);
        ini.get("inter_potential", "with_external_potentials", with_external_potentials);
        ini.get("inter_potential", "with_dynamic_random_charges", with_dynamic_random_charges);
        ini.get("inter_potential", "interpotential_file", interpotential_file);
        ini.get("inter_potential", "with_maturity", with_maturity);
        ini.get("flame_coupling", "with_flame_coupling", with_flame_coupling);
        if (with_flame_coupling) throw Error(INPUT_ERROR, "flame coupling is out of scope (SURVEY.md §2 row 9)");
        ini.get("output", "write_between_event_frequency", write_between_event_frequency);
        ini.get("output", "write_events_frequency", write_events_frequency);
        ini.get("output", "write_Delta_t", write_Delta_t);
        // physical_model.cpp:228-269
        if (init_mode == NORMAL) {
            box_length = mean_diameter * 1E-9 *
                         std::pow(static_cast<double>(n_monomeres) * PI / 6. / volume_fraction *
                                      (1. + 3. * std::pow(dispersion_diameter / mean_diameter, 2)),
                                  1. / 3.);
            mean_massic_radius = 0.5 * 1E-9 *
                                 (std::pow(mean_diameter, 4) + 6 * std::pow(mean_diameter, 2) * std::pow(dispersion_diameter, 2) +
                                  3 * std::pow(dispersion_diameter, 4)) /
                                 (std::pow(mean_diameter, 3) + 3 * mean_diameter * std::pow(dispersion_diameter, 2));
            double mean_radius = 0.5 * mean_diameter * 1E-9, dispersion_radius = 0.5 * dispersion_diameter * 1E-9;
            double tot_volume_pp = static_cast<double>(n_monomeres) * (4.0 * PI / 3.0) * (mean_radius) *
                                   (std::pow(mean_radius, 2) + 3.0 * std::pow(dispersion_radius, 2));
            double tot_surface_pp =
                static_cast<double>(n_monomeres) * (4.0 * PI) * (std::pow(mean_radius, 2) + std::pow(dispersion_radius, 2));
            box_volume = std::pow(box_length, 3);
            total_surface_concent = tot_surface_pp / box_volume;
            total_volume_concent = tot_volume_pp / box_volume;
        } else {
            box_length = mean_diameter * 1E-9 *
                         std::pow(static_cast<double>(n_monomeres) * PI / 6. / volume_fraction *
                                      std::exp(9. / 2. * std::pow(std::log(dispersion_diameter), 2)),
                                  1. / 3.);
            mean_massic_radius = 0.5 * mean_diameter * 1E-9 * std::exp(1.5 * std::pow(std::log(dispersion_diameter), 2));
            double mean_radius = 0.5 * mean_diameter * 1E-9;
            double tot_volume_pp = static_cast<double>(n_monomeres) * (4.0 * PI / 3.0) * std::pow(mean_radius, 3) *
                                   std::exp(4.5 * std::pow(std::log(dispersion_diameter), 2));
            double tot_surface_pp = static_cast<double>(n_monomeres) * (4.0 * PI) * std::pow(mean_radius, 2) *
                                    std::exp(2 * std::pow(std::log(dispersion_diameter), 2));
            box_volume = std::pow(box_length, 3);
            total_surface_concent = tot_surface_pp / box_volume;
            total_volume_concent = tot_volume_pp / box_volume;
        }
        update_temperature(temperature);
        u_sg = flux_surfgrowth / density;
        aggregate_concentration = static_cast<double>(n_monomeres) / box_volume;
        monomer_concentration = aggregate_concentration;
        if (with_external_potentials) {
            // the reference resolves this path against the CWD (physical_model_interpotential.cpp:49); base_dir = caller's choice
            std::string p = interpotential_file;
            if (!p.empty() && p[0] != '/' && !base_dir.empty()) p = base_dir + "/" + p;
            intpotential.load(p);
        }
    }
    // physical_model.cpp:536-546
    void update_temperature(double t) {
        temperature = t;
        viscosity = VISCOSITY_REF * (SUTHERLAND + TEMPERATURE_REF) / (SUTHERLAND + temperature) * std::pow(temperature / TEMPERATURE_REF, 1.5);
        gaz_mean_free_path = FLUID_MFP_REF * (PRESSURE_REF / pressure) * (temperature / TEMPERATURE_REF) *
                             (1. + SUTHERLAND / TEMPERATURE_REF) / (1. + SUTHERLAND / temperature);
        friction_exponnant = 0.689 * (1. + std::erf(((gaz_mean_free_path / mean_massic_radius) + 4.454) / 10.628));
    }
    // physical_model.cpp:550-555
    double cunningham(double r) const {
        double a = 1.142, b = 0.558, c = 0.999;
        return 1.0 + a * gaz_mean_free_path / r + b * gaz_mean_free_path / r * std::exp(-c * r / gaz_mean_free_path);
    }
    // physical_model.cpp:557-578
    double random_diameter(double mean, double disp) {
        double diameter = 0;
        if (init_mode == NORMAL) diameter = random_normal(mean, disp);
        else {
            if (disp < 1.0) throw Error(INPUT_ERROR, "dispersion_diameter cannot be lower than 1");
            diameter = mean * std::pow(disp, std::sqrt(2.) * inverf(2. * random() - 1.0));
        }
        if (diameter <= 0) diameter = mean;
        return diameter * 1E-9;
    }
    double grow(double r, double dt) const { return r + u_sg * dt; }  // :587-590
    double friction_exponent(double r) const { return 0.689 * (1. + std::erf(((gaz_mean_free_path / r) + 4.454) / 10.628)); }  // :591-593
    double friction_coeff(double V, double v, double r) const {  // :594-600
        double fe = friction_exponent(r), cc = cunningham(r);
        return (6. * PI * viscosity * r / cc) * std::pow(V / v, fe / fractal_dimension);
    }
    double diffusivity(double f) const { return BOLTZMANN * temperature / f; }  // :601-603
    double mobility_diameter(double V, double v, double r) const {  // :607-617
        double fe = friction_exponent(r);
        double ra = r * std::pow(V / v, fe / fractal_dimension / 2.0);
        double cc_pp = cunningham(r), cc_a = cunningham(ra);
        return (cc_a / cc_pp) * 2.0 * r * std::pow(V / v, fe / fractal_dimension);
    }
    int get_random_charge(double d_m) {  // :627-637
        double kbT = BOLTZMANN * temperature;
        double sigma_q = std::sqrt(d_m * kbT / (2.0 * KE_E2));
        int q = static_cast<int>(std::round(random_normal(0.0, sigma_q)));
        return std::min(std::max(q, intpotential.min_charge), intpotential.max_charge);
    }
    void update(size_t n_agg, size_t n_sph, double V, double S) {  // :489-498
        total_volume_concent = V / box_volume;
        total_surface_concent = S / box_volume;
        aggregate_concentration = static_cast<double>(n_agg) / box_volume;
        monomer_concentration = static_cast<double>(n_sph) / box_volume;
        volume_fraction = V / box_volume;
    }
    bool finished(size_t n_agg, double avg_npp) const {  // :288-337 (STOPCODE / cpu limits are wall-clock: not restated)
        if (n_agg < 1) return true;
        if (n_agg <= number_of_aggregates_limit) return true;
        if (n_iter_without_event_limit > 0 && n_iter_without_event >= static_cast<size_t>(n_iter_without_event_limit)) return true;
        if (physical_time_limit > 0 && time >= physical_time_limit) return true;
        if (mean_monomere_per_aggregate_limit > 0 && avg_npp >= mean_monomere_per_aggregate_limit) return true;
        return false;
    }
};

// ---------------------------------------------------------------------------------------------
// State.  Spheres: include/constants.hpp:34-45 + include/spheres/sphere.hpp:50-54.
// Aggregates: constants.hpp:46-69 + include/aggregats/aggregat.hpp:39-82.
// ---------------------------------------------------------------------------------------------
struct Spheres {
    std::vector<double> x, y, z, r, volume, surface, rx, ry, rz;
    std::vector<long> label;
    std::vector<int> charge;
    size_t size() const { return x.size(); }
    void add(size_t n) {
        for (auto *v : {&x, &y, &z, &r, &volume, &surface, &rx, &ry, &rz}) v->insert(v->end(), n, 0.);
        label.insert(label.end(), n, 0);
        charge.insert(charge.end(), n, 0);
    }
    vec3 pos(size_t i) const { return {x[i], y[i], z[i]}; }
    vec3 rel(size_t i) const { return {rx[i], ry[i], rz[i]}; }
    void set_pos(size_t i, const vec3 &p) { x[i] = p[0]; y[i] = p[1]; z[i] = p[2]; }
    void update_vol_and_surf(size_t i) {  // sphere.cpp:149-154
        volume[i] = VOLUME_FACTOR * std::pow(r[i], 3);
        surface[i] = SURFACE_FACTOR * std::pow(r[i], 2);
    }
};

struct Aggregate {
    double rg = 0, f_agg = 0, lpm = 0, time_step = 0, rmax = 0, volume = 0, surface = 0, x = 0, y = 0, z = 0, rx = 0, ry = 0, rz = 0,
           proper_time = 0, dp = 0, dg_over_dp = 0, overlapping = 0, coordination_number = 0, electric_charge_field = 0, d_m = 0,
           CH_ratio = 0;
    int electric_charge = 0;
    size_t n_spheres = 0;
    double bulk_density = 0, alpha_vs_extreme = 0;
    std::array<size_t, 3> index_verlet{{0, 0, 0}};
    bool in_verlet = false;
    std::vector<size_t> myspheres;  // ordered global sphere indices
    std::vector<std::unordered_map<size_t, double>> distances;
    std::vector<double> distances_center, volumes, surfaces;
    vec3 pos() const { return {x, y, z}; }
    vec3 rel() const { return {rx, ry, rz}; }
};

struct ContactInfo {  // include/tools/contact_info.hpp:22-85 with ids instead of weak_ptrs
    double distance = std::numeric_limits<double>::infinity();
    long moving_sphere = -1, other_sphere = -1, moving_aggregate = -1, other_aggregate = -1;
};

struct StepRecord {  // what tests compare against the tap (tests/ref_trace.py)
    long long step, rand_calls, source;
    double dir[3], full_distance, distance;
    long long moving_sphere, other_sphere, moving_label, other_label, n_agg_before;
    double time_before, dt, proper_time_after, pos_after[3];
    long long merged, n_try;
};

struct Counters {
    long long steps = 0, events = 0, searches = 0, pair_sphere = 0, pair_bounding = 0, sorts = 0, duplications = 0;
};

struct System {
    PhysicalModel pm;
    Spheres spheres;
    std::vector<std::unique_ptr<Aggregate>> list;
    // Verlet: src/verlet/verlet.cpp:26-51 — n_div^3 std::set<size_t> of aggregate labels
    size_t n_div = 0;
    double verlet_width = 0;
    std::vector<std::set<size_t>> grid;
    double maxradius = 0., avg_npp = 1, max_time_step = 0.;
    std::vector<size_t> index_sorted_time_steps;
    std::vector<double> cumulative_time_steps;
    // calcul() locals (src/calcul.cpp:57-61)
    bool event = true;
    size_t duplication_threshold = 0, total_events = 0;
    Counters counters;
    bool stop_after_move = false;   // test knob: return from step() right after time_forward (where the tap's exit_step dump is taken)
    bool stable_sort_ties = false;  // test knob: replace std::sort by a (key,label) stable order (NOT reference behaviour)

    size_t size() const { return list.size(); }
    std::set<size_t> &cell(const std::array<size_t, 3> &c) { return grid[(c[0] * n_div + c[1]) * n_div + c[2]]; }
    const std::set<size_t> &cell(size_t i, size_t j, size_t k) const { return grid[(i * n_div + j) * n_div + k]; }
    void verlet_reset(size_t n, double width) {
        n_div = n;
        verlet_width = width;
        grid.assign(n * n * n, {});
    }

    // ------------------------------------------------------------------ Aggregate methods
    // aggregat.cpp:699-705
    std::array<size_t, 3> compute_index_verlet(const Aggregate &a) const {
        double step = double(pm.n_verlet_divisions) / pm.box_length;
        return {size_t(std::floor(a.x * step)), size_t(std::floor(a.y * step)), size_t(std::floor(a.z * step))};
    }
    // aggregat.cpp:92-96, 706-717, 109-118
    void set_verlet(size_t label) {
        Aggregate &a = *list[label];
        a.in_verlet = true;
        a.index_verlet = compute_index_verlet(a);
        cell(a.index_verlet).insert(label);
    }
    void unset_verlet(size_t label) {
        Aggregate &a = *list[label];
        if (a.in_verlet) cell(a.index_verlet).erase(label);
        a.in_verlet = false;
    }
    void set_position(size_t label, const vec3 &p) {
        Aggregate &a = *list[label];
        a.x = periodic_position(p[0], pm.box_length);
        a.y = periodic_position(p[1], pm.box_length);
        a.z = periodic_position(p[2], pm.box_length);
        if (a.in_verlet) {
            std::array<size_t, 3> n = compute_index_verlet(a);
            if (n != a.index_verlet) {
                cell(a.index_verlet).erase(label);
                a.index_verlet = n;
                cell(n).insert(label);
            }
        }
    }
    // aggregat.cpp:148-161
    void translate(size_t label, const vec3 &v) {
        Aggregate &a = *list[label];
        set_position(label, a.pos() + v);
        vec3 refpos = a.pos() - a.rel();
        spheres.set_pos(a.myspheres[0], refpos);
        for (size_t s : a.myspheres) spheres.set_pos(s, refpos + spheres.rel(s));
    }
    // aggregat.cpp:719-764
    void update_distances_and_overlapping(Aggregate &a) {
        a.overlapping = a.coordination_number = 0.0;
        double c_ij(0);
        int intersections(0);
        const size_t n = a.n_spheres;
        a.distances.resize(n);
        a.distances_center.resize(n);
        for (size_t i = 0; i < n; i++) {
            size_t old_size = a.distances[i].size();
            a.distances[i].clear();
            a.distances[i].reserve(old_size);
        }
        for (size_t i = 0; i < n; i++) {
            size_t si = a.myspheres[i];
            for (size_t j = i + 1; j < n; j++) {
                size_t sj = a.myspheres[j];
                double dist = std::sqrt(relative_distance_2(spheres.rel(si), spheres.rel(sj)));
                if (dist <= (1. + COORDINATION_EPSILON) * (spheres.r[si] + spheres.r[sj])) {
                    a.distances[i][j] = dist;
                    a.distances[j][i] = dist;
                    c_ij = (spheres.r[si] + spheres.r[sj] - dist) / (spheres.r[si] + spheres.r[sj]);
                    a.overlapping += 2.0 * c_ij;
                    intersections += 2;
                }
            }
        }
        if (intersections > 0) {
            a.overlapping /= static_cast<double>(intersections);
            a.coordination_number = static_cast<double>(intersections) / static_cast<double>(n);
        }
    }
    // aggregat.cpp:289-319
    static double volume_alpha_correction(double cn, double c_20, double c_30, double min_cn, double extreme) {
        double diff = std::abs(cn - min_cn);
        double correction = 0.25 * (3.0 * c_20 - c_30) * cn - c_30 * diff * 0.62741833 - pow(diff, 1.5) * 0.00332425;
        if (correction < 0.0) correction = 1.0;
        correction = std::min(correction, 1.0);
        double alpha_v = 1.0 - correction;
        return std::max(alpha_v, extreme);
    }
    static double surface_alpha_correction(double cn, double c_10, double min_cn, double extreme) {
        double diff = std::abs(cn - min_cn);
        double correction = 0.5 * c_10 * cn - (c_10 * c_10) * diff * 0.70132500 - (diff * diff) * 0.00450000;
        if (correction < 0.0) correction = 1.0;
        correction = std::min(correction, 1.0);
        double alpha_s = 1.0 - correction;
        return std::max(alpha_s, extreme);
    }
    // src/spheres/sphere_intersection.cpp:29-66 → (v1, v2, s1, s2)
    void intersection(size_t s1, size_t s2, double dist, double out[4]) const {
        out[0] = out[1] = out[2] = out[3] = 0.;
        if (dist <= 0) return;
        double r1 = spheres.r[s1], r2 = spheres.r[s2];
        if (dist < r1 + r2) {
            if (dist >= std::fdim(r1, r2)) {
                double h_1 = (r2 * r2 - (r1 - dist) * (r1 - dist)) / (2. * dist);
                double h_2 = (r1 * r1 - (r2 - dist) * (r2 - dist)) / (2. * dist);
                out[0] = PI * (h_1 * h_1) * (3 * r1 - h_1) / 3.;
                out[1] = PI * (h_2 * h_2) * (3 * r2 - h_2) / 3.;
                out[2] = 2 * PI * r1 * h_1;
                out[3] = 2 * PI * r2 * h_2;
            } else if (r1 < r2) {
                out[0] = spheres.volume[s1];
                out[2] = spheres.surface[s1];
            } else {
                out[1] = spheres.volume[s2];
                out[3] = spheres.surface[s2];
            }
        }
    }
    // aggregat.cpp:321-430
    void compute_volume_surface(Aggregate &a) {
        a.volume = a.surface = 0.0;
        const size_t n = a.n_spheres;
        a.volumes.resize(n);
        a.surfaces.resize(n);
        if (pm.volsurf_method == SBL) throw Error(SBL_ERROR, "SBL not available");
        if (pm.volsurf_method == ARVO) throw Error(ARVO_ERROR, "ARVO not available");
        for (size_t i = 0; i < n; i++) {
            a.volumes[i] = spheres.volume[a.myspheres[i]];
            a.surfaces[i] = spheres.surface[a.myspheres[i]];
        }
        if (pm.volsurf_method == CAPS) {
            for (size_t i = 0; i < n; i++) {
                for (const auto &[j, dist] : a.distances[i]) {
                    if (j <= i) continue;
                    double in[4];
                    intersection(a.myspheres[i], a.myspheres[j], dist, in);
                    a.volumes[i] = a.volumes[i] - in[0];
                    a.surfaces[i] = a.surfaces[i] - in[2];
                    a.volumes[j] = a.volumes[j] - in[1];
                    a.surfaces[j] = a.surfaces[j] - in[3];
                }
                a.volumes[i] = std::max(a.volumes[i], 0.0);
                a.surfaces[i] = std::max(a.surfaces[i], 0.0);
            }
        }
        for (size_t i = 0; i < n; i++) {
            a.volume = a.volume + a.volumes[i];
            a.surface = a.surface + a.surfaces[i];
        }
        if (pm.volsurf_method == ALPHAS) {
            a.overlapping = 0.0;
            double c_v30(0.0), c_v20(0.0), c_s10(0.0), vp_sum(0.0), sp_sum(0.0);
            size_t intersections(0);
            for (size_t i = 0; i < n; i++) {
                for (const auto &it : a.distances[i]) {
                    double radius_1 = spheres.r[a.myspheres[i]];
                    double radius_2 = spheres.r[a.myspheres[it.first]];
                    double c_ij = (radius_1 + radius_2 - it.second) / (radius_1 + radius_2);
                    double vp1 = std::pow(radius_1, 3), vp2 = std::pow(radius_2, 3);
                    double sp1 = radius_1 * radius_1, sp2 = radius_2 * radius_2;
                    vp_sum += (vp1 + vp2);
                    sp_sum += (sp1 + sp2);
                    a.overlapping += c_ij;
                    c_s10 += c_ij * (sp1 + sp2);
                    c_v20 += (c_ij * c_ij) * (vp1 + vp2);
                    c_v30 += std::pow(c_ij, 3) * (vp1 + vp2);
                }
                intersections += a.distances[i].size();
            }
            if (intersections > 0) {
                c_s10 /= sp_sum;
                c_v20 /= vp_sum;
                c_v30 /= vp_sum;
                double min_cn = 2 * (1.0 - 1.0 / static_cast<double>(n));
                a.overlapping /= static_cast<double>(intersections);
                a.coordination_number = static_cast<double>(intersections) / static_cast<double>(n);
                a.volume *= volume_alpha_correction(a.coordination_number, c_v20, c_v30, min_cn, a.alpha_vs_extreme);
                a.surface *= surface_alpha_correction(a.coordination_number, c_s10, min_cn, a.alpha_vs_extreme);
            }
        }
        if (a.volume <= 0 || a.surface <= 0) throw Error(VOL_SURF_ERROR, "VolSurfError");
    }
    // aggregat.cpp:431-483, 247-282
    void update_partial(size_t label) {
        Aggregate &a = *list[label];
        const size_t n = a.n_spheres;
        vec3 r{0., 0., 0.};
        for (size_t i = 0; i < n; i++) {
            vec3 c = spheres.rel(a.myspheres[i]) * a.volumes[i];
            r[0] += c[0]; r[1] += c[1]; r[2] += c[2];
        }
        r[0] /= a.volume; r[1] /= a.volume; r[2] /= a.volume;
        for (size_t i = 0; i < n; i++) {
            vec3 diff = spheres.rel(a.myspheres[i]) - r;
            a.distances_center[i] = std::sqrt(diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2]);
        }
        set_position(label, spheres.pos(a.myspheres[0]) + r);
        a.rx = r[0]; a.ry = r[1]; a.rz = r[2];
        a.rmax = 0.0;
        for (size_t i = 0; i < n; i++) a.rmax = std::max(a.rmax, spheres.r[a.myspheres[i]] + a.distances_center[i]);
        double arg(0.), brg(0.);
        for (size_t i = 0; i < n; i++) {
            arg = arg + a.volumes[i] * (a.distances_center[i] * a.distances_center[i]);
            brg = brg + a.volumes[i] * (spheres.r[a.myspheres[i]] * spheres.r[a.myspheres[i]]);
        }
        a.rg = std::sqrt(std::abs((arg + 3. / 5. * brg) / (a.volume)));
        a.dp = 0.;
        double vol_pp(0.0);
        for (size_t i = 0; i < n; i++) {
            vol_pp += spheres.volume[a.myspheres[i]];
            a.dp += spheres.r[a.myspheres[i]];
        }
        a.dp = 2 * (a.dp) / static_cast<double>(n);
        vol_pp = vol_pp / static_cast<double>(n);
        // set_bulk_density aggregat.cpp:119-147
        if (pm.with_maturity) {
            double dpp_nm = (a.dp) * 1e+09;
            a.CH_ratio = 0.5 * (std::erf((dpp_nm - 4.0) / 1.0) + 1.0) * (CH_MATURE - CH_YOUNG) + CH_YOUNG;
            a.bulk_density = RHO_YOUNG + (RHO_MATURE - RHO_YOUNG) / (CH_MATURE - CH_YOUNG) * ((a.CH_ratio) - CH_YOUNG);
            if (a.bulk_density < RHO_YOUNG || a.bulk_density > RHO_MATURE) throw Error(INPUT_ERROR, "Problem with bulk density");
        } else {
            a.bulk_density = pm.density;
        }
        a.f_agg = pm.friction_coeff(a.volume, vol_pp, 0.5 * (a.dp));
        a.d_m = pm.mobility_diameter(a.volume, vol_pp, 0.5 * (a.dp));
        double masse = a.bulk_density * (a.volume);
        double relax_time = masse / a.f_agg;
        a.time_step = 3. * relax_time;
        double diffusivity = pm.diffusivity(a.f_agg);
        a.lpm = sqrt(6. * diffusivity * (a.time_step));
        a.dg_over_dp = 2 * (a.rg) / (a.dp);
        if (a.rmax > maxradius) maxradius = a.rmax;
    }
    void update(size_t label) {  // aggregat.cpp:284-288
        update_distances_and_overlapping(*list[label]);
        compute_volume_surface(*list[label]);
        update_partial(label);
    }

    // ------------------------------------------------------------------ init (a23)
    // verlet.cpp:52-98 → labels in (i,j,k) scan order, ascending inside a cell
    std::vector<size_t> verlet_neighborhood(const vec3 &src, const vec3 &direction, double distance) const {
        double xp{src[0] + distance + std::max(direction[0], 0.)}, xm{src[0] - distance + std::min(direction[0], 0.)};
        double yp{src[1] + distance + std::max(direction[1], 0.)}, ym{src[1] - distance + std::min(direction[1], 0.)};
        double zp{src[2] + distance + std::max(direction[2], 0.)}, zm{src[2] - distance + std::min(direction[2], 0.)};
        const double nd = static_cast<double>(n_div), width = verlet_width;
        auto bi1{static_cast<int>(std::floor(nd * xm / width))}, bi2{static_cast<int>(std::floor(nd * xp / width) + 1)};
        auto bj1{static_cast<int>(std::floor(nd * ym / width))}, bj2{static_cast<int>(std::floor(nd * yp / width) + 1)};
        auto bk1{static_cast<int>(std::floor(nd * zm / width))}, bk2{static_cast<int>(std::floor(nd * zp / width) + 1)};
        if (bi2 - bi1 >= static_cast<int>(n_div)) { bi1 = 0; bi2 = static_cast<int>(n_div) - 1; }
        if (bj2 - bj1 >= static_cast<int>(n_div)) { bj1 = 0; bj2 = static_cast<int>(n_div) - 1; }
        if (bk2 - bk1 >= static_cast<int>(n_div)) { bk1 = 0; bk2 = static_cast<int>(n_div) - 1; }
        std::vector<size_t> out;
        for (int i = bi1; i <= bi2; i++)
            for (int j = bj1; j <= bj2; j++)
                for (int k = bk1; k <= bk2; k++) {
                    auto ii = static_cast<size_t>(periodic_position(i, static_cast<int>(n_div)));
                    auto jj = static_cast<size_t>(periodic_position(j, static_cast<int>(n_div)));
                    auto kk = static_cast<size_t>(periodic_position(k, static_cast<int>(n_div)));
                    const auto &c = cell(ii, jj, kk);
                    out.insert(out.end(), c.begin(), c.end());
                }
        return out;
    }
    // aggregat_distance.cpp:45-58 + aggregat_list.cpp:534-548
    bool test_free_space(const vec3 &pos, double radius) const {
        std::vector<size_t> nb = verlet_neighborhood(pos, {0, 0, 0}, radius + maxradius);
        for (size_t suspect : nb) {
            const Aggregate &a = *list[suspect];
            if (!contact_spheres(pos, radius, a.pos(), a.rmax, pm.box_length)) continue;
            for (size_t s : a.myspheres)
                if (contact_spheres(pos, radius, spheres.pos(s), spheres.r[s], pm.box_length)) return false;
        }
        return true;
    }
    // aggregat.cpp:162-229
    void init_aggregate(size_t new_label, size_t sphere_index, bool nucleation) {
        Aggregate &a = *list[new_label];
        a.proper_time = pm.time;
        double diameter = nucleation ? pm.random_diameter(pm.mean_diameter_nucleation, pm.dispersion_diameter_nucleation)
                                     : pm.random_diameter(pm.mean_diameter, pm.dispersion_diameter);
        for (size_t n_try = 0; n_try < spheres.size(); n_try++) {
            double px = pm.random() * pm.box_length;
            double py = pm.random() * pm.box_length;
            double pz = pm.random() * pm.box_length;
            vec3 newpos{px, py, pz};
            if (test_free_space(newpos, diameter * 0.5)) {
                set_position(new_label, newpos);
                a.proper_time = pm.time;
                set_verlet(new_label);
                spheres.label[sphere_index] = int(new_label);
                spheres.set_pos(sphere_index, newpos);
                spheres.r[sphere_index] = diameter * 0.5;
                spheres.rx[sphere_index] = spheres.ry[sphere_index] = spheres.rz[sphere_index] = 0.;
                spheres.update_vol_and_surf(sphere_index);
                a.myspheres = {sphere_index};
                a.n_spheres = 1;
                a.alpha_vs_extreme = 1.0 / static_cast<double>(a.n_spheres);
                a.d_m = diameter;
                a.electric_charge = pm.with_electric_charges ? pm.get_random_charge(a.d_m) : 0;
                spheres.charge[sphere_index] = a.electric_charge;
                update(new_label);
                return;
            }
        }
        throw Error(TOO_DENSE_ERROR, "TooDenseError");
    }
    void refresh() {  // aggregat_list.cpp:100-108
        max_time_step = list[0]->time_step;
        for (const auto &a : list) max_time_step = std::max(a->time_step, max_time_step);
        avg_npp = static_cast<double>(spheres.size()) / static_cast<double>(size());
    }
    double get_total_volume() const {  // aggregat_list.cpp:28-45
        double t(0.0);
        for (const auto &a : list) t += a->volume;
        return t;
    }
    double get_total_surface() const {
        double t(0.0);
        for (const auto &a : list) t += a->surface;
        return t;
    }
    // aggregat_list_storage.cpp:52-88
    void construct() {
        const size_t n = pm.n_monomeres;
        spheres.add(n);
        verlet_reset(pm.n_verlet_divisions, pm.box_length);
        for (size_t i = 0; i < n; i++) list.push_back(std::make_unique<Aggregate>());
        for (size_t i = 0; i < n; i++) init_aggregate(i, i, false);
        refresh();
        if (pm.enforce_volume_fraction) {
            double current_total_volume = get_total_volume();
            double prescribed_total_volume = pm.volume_fraction * std::pow(pm.box_length, 3);
            double correction = std::pow(prescribed_total_volume / current_total_volume, 1. / 3.);
            for (size_t s = 0; s < spheres.size(); s++) {
                spheres.r[s] = spheres.r[s] * correction;
                spheres.update_vol_and_surf(s);
            }
            for (const auto &a : list) compute_volume_surface(*a);
        }
        duplication_threshold = size() / 8;  // calcul.cpp:58
    }
    // aggregat_list.cpp:82-99
    void add(size_t n) {
        size_t n_agg0 = size(), n_sph0 = spheres.size();
        spheres.add(n);
        for (size_t i = 0; i < n; i++) list.push_back(std::make_unique<Aggregate>());
        for (size_t i = 0; i < n; i++) init_aggregate(n_agg0 + i, n_sph0 + i, true);
        refresh();
    }

    // ------------------------------------------------------------------ pick (a6, a7)
    void sort_time_steps(double factor) {  // aggregat_list.cpp:109-141
        counters.sorts++;
        const size_t n = size();
        std::vector<double> tp_t(n);
        for (size_t i = 0; i < n; i++) tp_t[i] = factor / (list[i]->time_step);
        std::vector<size_t> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        if (stable_sort_ties) std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return tp_t[a] < tp_t[b]; });
        else std::sort(idx.begin(), idx.end(), [&tp_t](size_t i_1, size_t i_2) { return tp_t[i_1] < tp_t[i_2]; });
        index_sorted_time_steps = idx;
        cumulative_time_steps.resize(n);
        cumulative_time_steps[0] = tp_t[idx[0]];
        for (size_t i = 1; i < n; i++) cumulative_time_steps[i] = cumulative_time_steps[i - 1] + tp_t[idx[i]];
    }
    size_t pick_random() {  // :59-66
        double val_alea = pm.random() * cumulative_time_steps[size() - 1];
        long n = std::lower_bound(cumulative_time_steps.begin(), cumulative_time_steps.end(), val_alea) - cumulative_time_steps.begin();
        return index_sorted_time_steps[static_cast<size_t>(n)];
    }
    size_t pick_last() const {  // :67-81
        double time = list[0]->proper_time;
        size_t latest = 0;
        for (size_t i = 0; i < size(); i++)
            if (list[i]->proper_time < time) { time = list[i]->proper_time; latest = i; }
        return latest;
    }
    vec3 random_direction() {  // tools.cpp:82-89
        double thetarandom = pm.random() * 2 * PI;
        double phirandom = std::acos(1 - 2 * pm.random());
        return {std::sin(phirandom) * std::cos(thetarandom), std::sin(phirandom) * std::sin(thetarandom), std::cos(phirandom)};
    }

    // ------------------------------------------------------------------ contact search (a8–a11)
    std::vector<size_t> get_neighborhood(size_t source, const vec3 &direction, double distance) const {  // aggregat_list.cpp:485-508
        double mindist(list[source]->rmax + maxradius);
        std::vector<size_t> nb = verlet_neighborhood(list[source]->pos(), distance * direction, mindist);
        for (size_t i = 0; i < nb.size(); i++)
            if (nb[i] == source) { nb.erase(nb.begin() + long(i)); return nb; }
        throw Error(VERLET_ERROR, "Aggregate not on the verlet list ???");
    }
    std::multimap<double, size_t> filter_neighborhood(size_t moving, const vec3 &direction, const std::vector<size_t> &nb,
                                                      double distance) {  // :510-532
        std::multimap<double, size_t> sorted;
        const Aggregate &me = *list[moving];
        for (size_t other : nb) {
            const Aggregate &o = *list[other];
            double d = pair_distance_to_contact(me.pos(), me.rmax, o.pos(), o.rmax, direction, distance, pm.box_length);
            counters.pair_bounding++;
            if (d < distance) sorted.insert({d, other});
        }
        return sorted;
    }
    ContactInfo aggregate_distance_to_contact(size_t a1, size_t a2, const vec3 &direction, double distance) {  // aggregat_distance.cpp:24-44 + sphere_contact.cpp:126-139
        ContactInfo closest;
        closest.moving_aggregate = -1;
        const Aggregate &A = *list[a1], &B = *list[a2];
        for (size_t i : A.myspheres) {
            double half = std::numeric_limits<double>::infinity();
            long half_other = -1;
            for (size_t j : B.myspheres) {
                double d = pair_distance_to_contact(spheres.pos(i), spheres.r[i], spheres.pos(j), spheres.r[j], direction, distance,
                                                    pm.box_length);
                if (d < half) { half = d; half_other = (long)j; }
            }
            counters.pair_sphere += (long long)B.myspheres.size();
            if (half < closest.distance) {
                closest.distance = half;
                closest.other_sphere = half_other;
                closest.moving_sphere = (long)i;
                closest.moving_aggregate = (long)a1;
                closest.other_aggregate = (long)a2;
            }
        }
        return closest;
    }
    ContactInfo distance_to_next_contact(size_t source, const vec3 &direction, double distance) {  // aggregat_list.cpp:447-484
        counters.searches++;
        std::vector<size_t> nb(get_neighborhood(source, direction, distance));
        std::multimap<double, size_t> filtered(filter_neighborhood(source, direction, nb, distance));
        ContactInfo closest;
        for (auto suspect : filtered) {
            auto [suspect_distance, id] = suspect;
            if (closest.distance <= 0.) break;
            if (closest.distance < suspect_distance) break;
            ContactInfo potential = aggregate_distance_to_contact(source, id, direction, distance);
            if (potential.distance < closest.distance) closest = potential;
        }
        return closest;
    }

    // ------------------------------------------------------------------ potentials (a25)
    Regime check_InterPotentialRegime(const ContactInfo &ci) {  // aggregat_list.cpp:313-366
        double D_moving = 2.0 * spheres.r[(size_t)ci.moving_sphere], D_other = 2.0 * spheres.r[(size_t)ci.other_sphere];
        double P_stick(1.0), P_coll(1.0);
        if (pm.with_external_potentials) {
            int q_moving = list[(size_t)ci.moving_aggregate]->electric_charge, q_other = list[(size_t)ci.other_aggregate]->electric_charge;
            auto [E_bar, E_well] = pm.intpotential.get(D_moving, D_other, q_moving, q_other);
            double E_stick = std::abs(E_well) + std::abs(E_bar);
            P_stick = std::erf(std::sqrt(E_stick)) - std::sqrt(E_stick) * std::exp(-E_stick);
            P_coll = 1.0 - std::erf(std::sqrt(E_bar)) + std::sqrt(E_bar) * std::exp(-E_bar);
        } else {
            double kbT = BOLTZMANN * (pm.temperature);
            double D = D_moving * D_other / (D_moving + D_other);
            D = D * (1e+09);
            double E_well = (-6.6891e-23) * std::pow(D, 3) + (1.1244e-21) * std::pow(D, 2) + (1.1394e-20) * D - 5.5373e-21;
            P_stick = 1.0 - (1.0 + std::abs(E_well) / kbT) * std::exp(-std::abs(E_well) / kbT);
        }
        if (pm.random() > P_coll) return REPULSION;
        if (pm.random() > P_stick) return BOUNCING;
        return STICKING;
    }

    // ------------------------------------------------------------------ merge (a15, a16)
    void remove_aggregate(size_t id) {  // aggregat_list_storage.cpp:37-44 + list_storage_methods.hpp:79-87 + aggregat.cpp:80-91
        unset_verlet(id);  // ~Aggregate(), aggregat_storage.cpp:113-115
        list.erase(list.begin() + long(id));
        for (size_t i = id; i < list.size(); i++) {
            Aggregate &a = *list[i];
            size_t old_label = i + 1;
            if (a.in_verlet) cell(a.index_verlet).erase(old_label);
            for (size_t s : a.myspheres) spheres.label[s]--;
            if (a.in_verlet) cell(a.index_verlet).insert(i);
        }
    }
    bool merge(const ContactInfo &ci) {  // aggregat_list.cpp:367-410 + aggregat.cpp:486-544
        if (ci.moving_sphere < 0 || ci.other_sphere < 0) return false;
        size_t ms = (size_t)ci.moving_sphere, os = (size_t)ci.other_sphere;
        if (!contact_spheres(spheres.pos(ms), spheres.r[ms], spheres.pos(os), spheres.r[os], pm.box_length)) return false;
        auto keeped = static_cast<size_t>(std::min(spheres.label[ms], spheres.label[os]));
        auto removed = static_cast<size_t>(std::max(spheres.label[ms], spheres.label[os]));
        double newtime = (list[keeped]->proper_time) + (list[removed]->proper_time) - pm.time;
        int total_charge = list[keeped]->electric_charge + list[removed]->electric_charge;
        {
            Aggregate &me = *list[keeped], &other = *list[removed];
            size_t mysphere, othersphere;
            if ((size_t)ci.moving_aggregate == keeped) { mysphere = ms; othersphere = os; }
            else if ((size_t)ci.other_aggregate == keeped) { mysphere = os; othersphere = ms; }
            else throw Error(MERGE_ERROR, "ListAggregate want to merge but the aggregate refuses");
            vec3 refpos = spheres.pos(me.myspheres[0]);
            vec3 ref_root_to_contact = spheres.rel(mysphere);
            vec3 d = spheres.pos(othersphere) - spheres.pos(mysphere);
            vec3 diffcontact{periodic_distance(d[0], pm.box_length), periodic_distance(d[1], pm.box_length),
                             periodic_distance(d[2], pm.box_length)};
            vec3 other_root_to_contact = spheres.rel(othersphere);
            vec3 diffpos = ref_root_to_contact + diffcontact - other_root_to_contact;
            for (size_t s : other.myspheres) {
                spheres.label[s] = long(keeped);
                spheres.rx[s] += diffpos[0]; spheres.ry[s] += diffpos[1]; spheres.rz[s] += diffpos[2];
                vec3 newpos = spheres.rel(s);
                newpos[0] += refpos[0]; newpos[1] += refpos[1]; newpos[2] += refpos[2];
                spheres.set_pos(s, newpos);
            }
            me.myspheres.insert(me.myspheres.end(), other.myspheres.begin(), other.myspheres.end());
            me.n_spheres = me.myspheres.size();
            me.alpha_vs_extreme = 1.0 / static_cast<double>(me.n_spheres);
            update(keeped);
        }
        remove_aggregate(removed);
        list[keeped]->proper_time = newtime;
        if (pm.with_dynamic_random_charges) list[keeped]->electric_charge = pm.get_random_charge(list[keeped]->d_m);
        else list[keeped]->electric_charge = total_charge;
        return true;
    }

    // ------------------------------------------------------------------ growth (a21)
    bool croissance_surface_one(double dt, size_t index) {  // aggregat_list.cpp:567-579 + aggregat.cpp:230-243 + sphere.cpp:113-120
        Aggregate &a = *list[index];
        for (size_t s : a.myspheres) {
            double new_r = pm.grow(spheres.r[s], dt);
            double new_r_2 = new_r * new_r;
            double new_r_3 = new_r_2 * new_r;
            spheres.r[s] = new_r;
            spheres.volume[s] = VOLUME_FACTOR * new_r_3;
            spheres.surface[s] = SURFACE_FACTOR * new_r_2;
        }
        for (size_t s : a.myspheres)
            if (spheres.r[s] <= pm.rp_min_oxid)
                throw Error(UNKNOWN_ERROR, "sphere removal by oxidation (u_sg<0) is outside the restated path (SURVEY.md §8a a26)");
        return false;
    }
    bool croissance_surface_all(double dt) {  // aggregat_list.cpp:549-566
        for (size_t i = 0; i < size(); i++) croissance_surface_one(dt, i);
        return false;
    }

    // ------------------------------------------------------------------ duplication (a24)
    void duplication() {  // aggregat_list.cpp:142-190 + aggregat_storage.cpp:117-159
        counters.duplications++;
        size_t old_n_agg = size();
        double old_l = pm.box_length;
        pm.box_length *= 2;
        pm.n_monomeres *= 8;
        pm.box_volume = std::pow(pm.box_length, 3);
        for (size_t i = 0; i < size(); i++) unset_verlet(i);
        for (size_t iagg = 0; iagg < old_n_agg; iagg++)
            for (int i = 0; i <= 1; i++)
                for (int j = 0; j <= 1; j++)
                    for (int k = 0; k <= 1; k++)
                        if (i != 0 or j != 0 or k != 0) {
                            const Aggregate &src = *list[iagg];
                            auto na = std::make_unique<Aggregate>();
                            // storage row copy (21 fields) + copied members; electric_charge member reset to 0 by the copy ctor
                            *na = Aggregate(src);
                            na->electric_charge = 0;
                            na->in_verlet = false;
                            na->index_verlet = {{0, 0, 0}};
                            size_t new_label = size();
                            na->myspheres.clear();
                            for (size_t s : src.myspheres) {
                                size_t ns = spheres.size();
                                spheres.add(1);
                                spheres.x[ns] = spheres.x[s]; spheres.y[ns] = spheres.y[s]; spheres.z[ns] = spheres.z[s];
                                spheres.r[ns] = spheres.r[s]; spheres.volume[ns] = spheres.volume[s]; spheres.surface[ns] = spheres.surface[s];
                                spheres.rx[ns] = spheres.rx[s]; spheres.ry[ns] = spheres.ry[s]; spheres.rz[ns] = spheres.rz[s];
                                spheres.charge[ns] = 0;  // sphere_storage.cpp:120-135 copy ctor resets electric_charge
                                spheres.label[ns] = long(new_label);
                                na->myspheres.push_back(ns);
                            }
                            list.push_back(std::move(na));
                            vec3 vec_move = {i * old_l, j * old_l, k * old_l};
                            translate(new_label, vec_move);
                        }
        verlet_reset(pm.n_verlet_divisions, pm.box_length);
        for (size_t i = 0; i < size(); i++) set_verlet(i);
        pm.update(size(), spheres.size(), get_total_volume(), get_total_surface());
    }

    // ------------------------------------------------------------------ one MC step (a1) — src/calcul.cpp:66-281
    // returns false when finished() was true (no step done)
    bool step(StepRecord *rec) {
        if (pm.finished(size(), avg_npp)) return false;
        if (event) {
            if (pm.with_domain_duplication && size() <= duplication_threshold && !(pm.u_sg < 0.0)) duplication();
            if (pm.with_domain_reduction) throw Error(INPUT_ERROR, "domain reduction is outside the restated path");
        }
        double deltatemps(0);
        size_t num_agg(0);
        if (pm.pick_method == PICK_RANDOM) {
            double max = max_time_step;
            if (event || pm.with_surface_reactions || pm.with_flame_coupling) sort_time_steps(max);
            num_agg = pick_random();
            deltatemps = max / cumulative_time_steps[size() - 1];
        } else {
            num_agg = pick_last();
            deltatemps = list[num_agg]->time_step;
        }
        double deltatemps_indiv = list[num_agg]->time_step;
        double full_distance = list[num_agg]->lpm;
        vec3 vectdir = {{0, 0, 0}};
        bool effective_move = false;
        double move_distance = full_distance;
        int n_try(0);
        bool contact = false;
        ContactInfo next_contact;
        long long rand_after_search = -1;  // tap convention: draw counter when the last search returned
        while (!effective_move) {
            vectdir = random_direction();
            effective_move = true;
            move_distance = full_distance;
            n_try++;
            if (pm.with_collisions) {
                next_contact = distance_to_next_contact(num_agg, vectdir, full_distance);
                rand_after_search = pm.rng.calls;
                contact = next_contact.distance <= full_distance;
                if (contact) {
                    move_distance = next_contact.distance;
                    if (pm.with_potentials) {
                        Regime regime = check_InterPotentialRegime(next_contact);
                        if (regime != STICKING) effective_move = false;
                    }
                }
            }
        }
        if (rec) {
            rec->step = counters.steps;
            rec->rand_calls = rand_after_search >= 0 ? rand_after_search : pm.rng.calls;
            rec->source = (long long)num_agg;
            rec->dir[0] = vectdir[0]; rec->dir[1] = vectdir[1]; rec->dir[2] = vectdir[2];
            rec->full_distance = full_distance;
            rec->distance = next_contact.distance;
            rec->moving_sphere = next_contact.moving_sphere;
            rec->other_sphere = next_contact.other_sphere;
            rec->moving_label = next_contact.moving_aggregate;
            rec->other_label = next_contact.other_aggregate;
            rec->n_agg_before = (long long)size();
            rec->time_before = pm.time;
            rec->n_try = n_try;
        }
        translate(num_agg, vectdir * move_distance);
        deltatemps = deltatemps * (move_distance / full_distance + static_cast<double>(n_try - 1));
        deltatemps_indiv = deltatemps_indiv * (move_distance / full_distance + static_cast<double>(n_try - 1));
        list[num_agg]->proper_time += deltatemps;
        if (rec) {
            rec->dt = deltatemps;
            rec->proper_time_after = list[num_agg]->proper_time;
            rec->pos_after[0] = list[num_agg]->x; rec->pos_after[1] = list[num_agg]->y; rec->pos_after[2] = list[num_agg]->z;
        }
        if (stop_after_move) { counters.steps++; return true; }
        if (pm.pick_method == PICK_LAST) deltatemps = deltatemps / double(size());
        pm.time = pm.time + deltatemps;
        bool split = false, disappear = false;
        if (pm.with_surface_reactions) {
            if (pm.individual_surf_reactions) disappear = croissance_surface_one(deltatemps_indiv, num_agg);
            else disappear = croissance_surface_all(deltatemps);
            if (pm.u_sg < 0.0) throw Error(INPUT_ERROR, "splitting (u_sg<0) is outside the restated path");
        }
        bool merged = false;
        if (contact)
            if (!pm.individual_surf_reactions || !(split || disappear)) merged = merge(next_contact);
        if (rec) rec->merged = merged ? 1 : 0;
        if (pm.with_surface_reactions || pm.with_flame_coupling) {
            bool full = pm.n_iter_without_event % pm.full_aggregate_update_frequency == 0;
            if (pm.individual_surf_reactions && !merged && !split && !disappear) {
                if (full) update(num_agg); else update_partial(num_agg);
            } else {
                for (size_t i = 0; i < size(); i++) { if (full) update(i); else update_partial(i); }
            }
        }
        bool nucleation(false);
        if (pm.with_nucleation) {
            pm.nucleation_accum += pm.flux_nucleation * pm.box_volume * deltatemps;  // physical_model.cpp:499-502
            if (pm.nucleation_accum > 1.0) {
                nucleation = true;
                int monomers_to_add = static_cast<int>(std::floor(pm.nucleation_accum));
                pm.nucleation_accum -= static_cast<double>(monomers_to_add);
                add((size_t)monomers_to_add);
                sort_time_steps(max_time_step);
            }
        }
        event = split || merged || disappear || nucleation;
        if (event) { pm.n_iter_without_event = 0; total_events++; counters.events++; }
        else pm.n_iter_without_event++;
        if (event) refresh();
        if (event || pm.with_surface_reactions) pm.update(size(), spheres.size(), get_total_volume(), get_total_surface());
        counters.steps++;
        return true;
    }
};
}  // namespace orc

// =============================================================================================
// C API (ctypes) — test harness surface only
// =============================================================================================
using orc::System;
static thread_local std::string g_err;
static thread_local int g_code = 0;
#define ORC_TRY try {
#define ORC_CATCH(ret)                                                                   \
    }                                                                                    \
    catch (const orc::Error &e) { g_err = e.what(); g_code = e.code; return ret; }       \
    catch (const std::exception &e) { g_err = e.what(); g_code = orc::UNKNOWN_ERROR; return ret; }

extern "C" {
const char *orc_last_error() { return g_err.c_str(); }
int orc_last_code() { return g_code; }

void *orc_create(const char *ini_text, const char *base_dir, int construct) {
    ORC_TRY
    auto *s = new System();
    std::istringstream is(ini_text);
    s->pm.load(is, base_dir ? base_dir : "");
    if (construct) s->construct();
    return s;
    ORC_CATCH(nullptr)
}
void orc_destroy(void *h) { delete (System *)h; }
void orc_set_stop_after_move(void *h, int on) { ((System *)h)->stop_after_move = on != 0; }
void orc_set_stable_sort(void *h, int on) { ((System *)h)->stable_sort_ties = on != 0; }

// runs up to max_steps MC steps; fills recs[0..cap) if given; returns steps done, or -1 on error
long long orc_run(void *h, long long max_steps, orc::StepRecord *recs, long long cap) {
    ORC_TRY
    System &s = *(System *)h;
    long long done = 0;
    while (done < max_steps) {
        orc::StepRecord tmp;
        orc::StepRecord *r = recs ? (done < cap ? &recs[done] : &tmp) : nullptr;
        if (!s.step(r)) break;
        done++;
    }
    return done;
    ORC_CATCH(-1)
}
int orc_finished(void *h) { System &s = *(System *)h; return s.pm.finished(s.size(), s.avg_npp) ? 1 : 0; }
long long orc_n_spheres(void *h) { return (long long)((System *)h)->spheres.size(); }
long long orc_n_aggregates(void *h) { return (long long)((System *)h)->size(); }
void orc_get_counters(void *h, long long out[8]) {
    System &s = *(System *)h;
    out[0] = s.counters.steps; out[1] = s.counters.events; out[2] = s.counters.searches; out[3] = s.counters.pair_sphere;
    out[4] = s.counters.pair_bounding; out[5] = s.counters.sorts; out[6] = s.counters.duplications; out[7] = s.pm.rng.calls;
}
// scalars: time, box_length, maxradius, max_time_step, avg_npp, volume_fraction, aggregate_concentration,
// monomer_concentration, total_volume_concent, total_surface_concent, u_sg, gaz_mean_free_path, mean_massic_radius,
// friction_exponnant, viscosity, box_volume, n_iter_without_event, n_monomeres, temperature, nucleation_accum
void orc_get_scalars(void *h, double out[20]) {
    System &s = *(System *)h;
    const orc::PhysicalModel &p = s.pm;
    double v[20] = {p.time, p.box_length, s.maxradius, s.max_time_step, s.avg_npp, p.volume_fraction, p.aggregate_concentration,
                    p.monomer_concentration, p.total_volume_concent, p.total_surface_concent, p.u_sg, p.gaz_mean_free_path,
                    p.mean_massic_radius, p.friction_exponnant, p.viscosity, p.box_volume, (double)p.n_iter_without_event,
                    (double)p.n_monomeres, p.temperature, p.nucleation_accum};
    std::memcpy(out, v, sizeof(v));
}
// 9 sphere fields (field-major, n each) + labels + charges
void orc_get_spheres(void *h, double *fields, long long *label, long long *charge) {
    System &s = *(System *)h;
    size_t n = s.spheres.size();
    const std::vector<double> *f[9] = {&s.spheres.x, &s.spheres.y, &s.spheres.z, &s.spheres.r, &s.spheres.volume,
                                       &s.spheres.surface, &s.spheres.rx, &s.spheres.ry, &s.spheres.rz};
    for (int k = 0; k < 9; k++) std::memcpy(fields + k * n, f[k]->data(), n * sizeof(double));
    for (size_t i = 0; i < n; i++) { label[i] = s.spheres.label[i]; charge[i] = s.spheres.charge[i]; }
}
// 21 aggregate fields in AggregatesFields order (field-major), n_spheres, cells[3][n], charge, CSR offsets (n+1), members,
// per-member volumes / surfaces / distances_center
void orc_get_aggregates(void *h, double *fields, long long *n_spheres, long long *cells, long long *charge, long long *offsets,
                        long long *members, double *per_member) {
    System &s = *(System *)h;
    size_t n = s.size(), nm = s.spheres.size();
    long long off = 0;
    for (size_t i = 0; i < n; i++) {
        const orc::Aggregate &a = *s.list[i];
        double v[21] = {a.rg, a.f_agg, a.lpm, a.time_step, a.rmax, a.volume, a.surface, a.x, a.y, a.z, a.rx, a.ry, a.rz, a.proper_time,
                        a.dp, a.dg_over_dp, a.overlapping, a.coordination_number, a.electric_charge_field, a.d_m, a.CH_ratio};
        for (int k = 0; k < 21; k++) fields[k * n + i] = v[k];
        n_spheres[i] = (long long)a.n_spheres;
        for (int d = 0; d < 3; d++) cells[d * n + i] = (long long)a.index_verlet[d];
        charge[i] = a.electric_charge;
        offsets[i] = off;
        for (size_t k = 0; k < a.myspheres.size(); k++) {
            members[off + (long long)k] = (long long)a.myspheres[k];
            if (per_member) {
                per_member[off + (long long)k] = k < a.volumes.size() ? a.volumes[k] : 0.;
                per_member[nm + off + (long long)k] = k < a.surfaces.size() ? a.surfaces[k] : 0.;
                per_member[2 * nm + off + (long long)k] = k < a.distances_center.size() ? a.distances_center[k] : 0.;
            }
        }
        off += (long long)a.myspheres.size();
    }
    offsets[n] = off;
}
void orc_get_pick_table(void *h, long long *idx, double *cum) {
    System &s = *(System *)h;
    for (size_t i = 0; i < s.index_sorted_time_steps.size(); i++) { idx[i] = (long long)s.index_sorted_time_steps[i]; cum[i] = s.cumulative_time_steps[i]; }
}
long long orc_pick_table_size(void *h) { return (long long)((System *)h)->index_sorted_time_steps.size(); }

// Re-synchronise the restatement with a state downloaded from the device (same layouts as orc_get_spheres / orc_get_aggregates):
// everything a step reads is replaced — sphere and aggregate fields, ordered membership, Verlet cells (recomputed from the stored
// cell of every aggregate), clocks and counters of calcul() — and the RNG stream is positioned `rand_consumed` draws after
// srand(seed).  `event` is forced so that the next step re-sorts the pick table (valid whenever no aggregate changed since the
// last event, i.e. at any point of a run without surface growth).  The box must be the one of the .ini (no duplication yet).
// scalars = {time, maxradius, max_time_step, avg_npp, n_iter_without_event}.
int orc_set_state(void *h, long long n_sph, long long n_agg, const double *sph_fields, const long long *sph_label, const double *agg_fields,
                  const long long *cells, const long long *offsets, const long long *members, const double *per_member,
                  const double *scalars, long long rand_consumed) {
    ORC_TRY
    System &s = *(System *)h;
    orc::Spheres &sp = s.spheres;
    std::vector<double> *f[9] = {&sp.x, &sp.y, &sp.z, &sp.r, &sp.volume, &sp.surface, &sp.rx, &sp.ry, &sp.rz};
    for (int k = 0; k < 9; k++) f[k]->assign(sph_fields + k * n_sph, sph_fields + (k + 1) * n_sph);
    sp.label.assign((size_t)n_sph, 0);
    sp.charge.assign((size_t)n_sph, 0);
    for (long long i = 0; i < n_sph; i++) sp.label[(size_t)i] = (long)sph_label[i];
    s.list.clear();
    s.verlet_reset(s.pm.n_verlet_divisions, s.pm.box_length);
    for (long long i = 0; i < n_agg; i++) {
        auto a = std::make_unique<orc::Aggregate>();
        double *v[21] = {&a->rg, &a->f_agg, &a->lpm, &a->time_step, &a->rmax, &a->volume, &a->surface, &a->x, &a->y, &a->z, &a->rx, &a->ry,
                         &a->rz, &a->proper_time, &a->dp, &a->dg_over_dp, &a->overlapping, &a->coordination_number,
                         &a->electric_charge_field, &a->d_m, &a->CH_ratio};
        for (int k = 0; k < 21; k++) *v[k] = agg_fields[k * n_agg + i];
        const long long lo = offsets[i], hi = offsets[i + 1];
        a->n_spheres = (size_t)(hi - lo);
        a->alpha_vs_extreme = 1.0 / static_cast<double>(a->n_spheres);
        for (long long k = lo; k < hi; k++) {
            a->myspheres.push_back((size_t)members[k]);
            a->volumes.push_back(per_member[k]);
            a->surfaces.push_back(per_member[n_sph + k]);
            a->distances_center.push_back(per_member[2 * n_sph + k]);
        }
        a->index_verlet = {(size_t)cells[i], (size_t)cells[n_agg + i], (size_t)cells[2 * n_agg + i]};
        a->in_verlet = true;
        {   // contact graph (Aggregate::distances) rebuilt from the relative positions and the CURRENT radii; the stored overlap
            // statistics are kept.  With surface growth this equals the reference's graph only right after a full-update step
            // (calcul.cpp:184-206): re-synchronise there.
            const double ovl = a->overlapping, cn = a->coordination_number;
            const std::vector<double> dc = a->distances_center;
            s.update_distances_and_overlapping(*a);
            a->overlapping = ovl;
            a->coordination_number = cn;
            a->distances_center = dc;
        }
        s.list.push_back(std::move(a));
        s.cell(s.list.back()->index_verlet).insert((size_t)i);
    }
    s.pm.time = scalars[0];
    s.maxradius = scalars[1];
    s.max_time_step = scalars[2];
    s.avg_npp = scalars[3];
    s.pm.n_iter_without_event = (size_t)scalars[4];
    s.event = true;
    s.pm.rng.seed((unsigned)s.pm.random_seed);
    for (long long i = 0; i < rand_consumed; i++) s.pm.rng.next();
    return 0;
    ORC_CATCH(-1)
}
// Aggregate::update() (aggregat.cpp:284-288) of every aggregate from its spheres' radii and relative positions: the per-aggregate
// morphology (V, S, Rg, rmax, f_agg, d_m, lpm, time_step ...) a downloaded device state is checked against
int orc_update_all(void *h) {
    ORC_TRY
    System &s = *(System *)h;
    for (size_t i = 0; i < s.size(); i++) s.update(i);
    return 0;
    ORC_CATCH(-1)
}

// ---- unit-level entry points
double orc_pair_distance(const double p1[3], double r1, const double p2[3], double r2, const double dir[3], double dist, double L) {
    return orc::pair_distance_to_contact({p1[0], p1[1], p1[2]}, r1, {p2[0], p2[1], p2[2]}, r2, {dir[0], dir[1], dir[2]}, dist, L);
}
void orc_pair_distance_batch(long long n, const double *p1, const double *r1, const double *p2, const double *r2, const double *dir,
                             const double *dist, double L, double *out) {
    for (long long i = 0; i < n; i++)
        out[i] = orc_pair_distance(p1 + 3 * i, r1[i], p2 + 3 * i, r2[i], dir + 3 * i, dist[i], L);
}
int orc_search(void *h, long long source, const double dir[3], double dist, double *out_distance, long long out_ids[4]) {
    ORC_TRY
    System &s = *(System *)h;
    orc::ContactInfo c = s.distance_to_next_contact((size_t)source, {dir[0], dir[1], dir[2]}, dist);
    *out_distance = c.distance;
    out_ids[0] = c.moving_sphere; out_ids[1] = c.other_sphere; out_ids[2] = c.moving_aggregate; out_ids[3] = c.other_aggregate;
    return 0;
    ORC_CATCH(g_code)
}
void orc_rand_stream(unsigned seed, long long n, int *out) {
    orc::GlibcRand g;
    g.seed(seed);
    for (long long i = 0; i < n; i++) out[i] = g.next();
}
void orc_direction(double u1, double u2, double out[3]) {  // tools.cpp:82-89 on given draws
    double thetarandom = u1 * 2 * orc::PI;
    double phirandom = std::acos(1 - 2 * u2);
    out[0] = std::sin(phirandom) * std::cos(thetarandom);
    out[1] = std::sin(phirandom) * std::sin(thetarandom);
    out[2] = std::cos(phirandom);
}
// physics closures on the handle's PhysicalModel: out = {f_agg, d_m, time_step, lpm, cunningham(r), friction_exponent(r)}
void orc_physics(void *h, double V, double v, double r, double out[6]) {
    System &s = *(System *)h;
    double f = s.pm.friction_coeff(V, v, r);
    out[0] = f;
    out[1] = s.pm.mobility_diameter(V, v, r);
    double ts = 3. * (s.pm.density * V / f);
    out[2] = ts;
    out[3] = sqrt(6. * s.pm.diffusivity(f) * ts);
    out[4] = s.pm.cunningham(r);
    out[5] = s.pm.friction_exponent(r);
}
void orc_sort_introsort(long long n, const double *keys, long long *idx) {  // the libstdc++ order the reference relies on
    std::vector<size_t> v((size_t)n);
    std::iota(v.begin(), v.end(), 0);
    std::sort(v.begin(), v.end(), [keys](size_t a, size_t b) { return keys[a] < keys[b]; });
    for (long long i = 0; i < n; i++) idx[i] = (long long)v[(size_t)i];
}
// mcac::linreg (src/tools/tools.cpp:126-157) as AggregatList::get_instantaneous_fractal_law calls it
// (src/aggregats/aggregat_list_fractal_law.cpp:23-33: x = dg_over_dp, y = number of spheres, in label order).
// out = {ok, a, b, r}; `r` keeps the reference's pow(..., 2) where a square root is meant (:153-155).
void orc_linreg(long long n_, const double *x, const double *y, double *out) {
    double sumx = 0.0, sumx_2 = 0.0, sumxy = 0.0, sumy = 0.0, sumy_2 = 0.0;
    const double n = static_cast<double>(n_);
    for (long long i = 0; i < n_; i++) {
        const double log_x = std::log(x[i]), log_y = std::log(y[i]);
        sumx += log_x;
        sumx_2 += std::pow(log_x, 2);
        sumxy += log_x * log_y;
        sumy += log_y;
        sumy_2 += std::pow(log_y, 2);
    }
    const double denom = (n * sumx_2 - std::pow(sumx, 2));
    if (n_ == 0 || std::abs(denom) < 1e-9) { out[0] = out[1] = out[2] = out[3] = 0.; return; }
    out[0] = 1.;
    out[1] = (n * sumxy - sumx * sumy) / denom;
    out[2] = (sumy * sumx_2 - sumx * sumxy) / denom;
    out[3] = (sumxy - sumx * sumy / n) / std::pow((sumx_2 - std::pow(sumx, 2) / n) * (sumy_2 - std::pow(sumy, 2) / n), 2);
}
// the same sort with introsort's depth limit forced to `depth` (std::sort uses 2 * floor(log2(n))): libstdc++'s own
// __introsort_loop + __final_insertion_sort, so that the heap-sort branch behind the limit can be checked on the device
void orc_sort_introsort_depth(long long n, const double *keys, long long depth, long long *idx) {
    std::vector<size_t> v((size_t)n);
    std::iota(v.begin(), v.end(), 0);
    auto cmp = __gnu_cxx::__ops::__iter_comp_iter([keys](size_t a, size_t b) { return keys[a] < keys[b]; });
    if (n > 1) {
        std::__introsort_loop(v.begin(), v.end(), (long)depth, cmp);
        std::__final_insertion_sort(v.begin(), v.end(), cmp);
    }
    for (long long i = 0; i < n; i++) idx[i] = (long long)v[(size_t)i];
}
}  // extern "C"
