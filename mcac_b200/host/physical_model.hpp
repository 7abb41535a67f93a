// mcac_b200 host layer — mirror of the reference's PhysicalModel for the hot path
// (include/physical_model/physical_model.hpp:29-96, src/physical_model/physical_model.cpp:32-287,489-637).
// Same public field names and meaning, so reference call sites (`physicalmodel.time`, `.box_length`,
// `.finished(...)`, `.update(...)`) keep working; the .ini reader accepts the same sections/keys with the
// same defaults.  Options whose code is not built (flame coupling, electric charges, domain reduction,
// oxidation, sbl / arvo) are refused with InputError instead of being ignored.
#pragma once
#include <cstddef>
#include <cstdint>
#include <istream>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mcac_b200.h"

namespace mcac {

enum ErrorCodes { NO_ERROR, UNKNOWN_ERROR, IO_ERROR, VERLET_ERROR, INPUT_ERROR, ABANDON_ERROR, TOO_DENSE_ERROR, SBL_ERROR,
                  VOL_SURF_ERROR, MERGE_ERROR, ARVO_ERROR, InterPotential_ERROR };  // include/constants.hpp:70-83
enum PickMethods { PICK_RANDOM, PICK_LAST, INVALID_PICK_METHOD };
enum MonomeresInitialisationMode { LOG_NORMAL_INITIALISATION, NORMAL_INITIALISATION, INVALID_INITIALISATION };
enum VolSurfMethods { SPHERICAL_CAPS, EXACT_SBL, EXACT_ARVO, ALPHAS, NONE, INVALID_VOLSURF_METHOD };

// include/exceptions.hpp:30-162: every error carries the process exit code
struct BaseException : std::runtime_error {
    ErrorCodes code;
    BaseException(ErrorCodes c, const std::string &what) : std::runtime_error(what), code(c) {}
};
struct InputError : BaseException { explicit InputError(const std::string &w) : BaseException(INPUT_ERROR, w) {} };
struct IOError : BaseException { explicit IOError(const std::string &w) : BaseException(IO_ERROR, w) {} };
struct TooDenseError : BaseException { TooDenseError() : BaseException(TOO_DENSE_ERROR, "Too dense") {} };
struct VerletError : BaseException { explicit VerletError(const std::string &w) : BaseException(VERLET_ERROR, w) {} };
struct MergeError : BaseException { explicit MergeError(const std::string &w) : BaseException(MERGE_ERROR, w) {} };
struct VolSurfError : BaseException { VolSurfError() : BaseException(VOL_SURF_ERROR, "Volume or surface <= 0") {} };
struct DeviceError : BaseException { explicit DeviceError(const std::string &w) : BaseException(UNKNOWN_ERROR, w) {} };

class PhysicalModel {
  public:
    double fractal_dimension = 1.4, fractal_prefactor = 1.8;
    double flux_surfgrowth = 0., u_sg = 0.;
    double flux_nucleation = 0., nucleation_accum = 0.;
    double pressure = 101300, temperature = 293.15, gaz_mean_free_path = 66.5E-9, viscosity = 18.203E-6, density = 1800.;
    double mean_diameter = 30., dispersion_diameter = 1.0;
    double mass_nuclei = 0., mean_diameter_nucleation = 5.0, dispersion_diameter_nucleation = 1.0;
    double mean_massic_radius = 0., friction_exponnant = 0.;
    double time = 0.;
    double volume_fraction = 1e-3, box_length = 0., box_volume = 0., aggregate_concentration = 0., monomer_concentration = 0.;
    double total_surface_concent = 0., total_volume_concent = 0.;
    double rp_min_oxid = 0.166e-09;
    size_t n_verlet_divisions = 10;
    PickMethods pick_method = PICK_RANDOM;
    VolSurfMethods volsurf_method = NONE;
    size_t n_monomeres = 2500;
    size_t n_time_per_file = 10;
    MonomeresInitialisationMode monomeres_initialisation_type = LOG_NORMAL_INITIALISATION;
    size_t n_iter_without_event = 0;
    double cpu_limit = -1, cpu_event_limit = -1, physical_time_limit = -1;
    double write_Delta_t = -1;
    bool finished_flag = false;  // PhysicalModel::finished() as evaluated by the device loop at the last loop top
    int mean_monomere_per_aggregate_limit = -1;
    size_t number_of_aggregates_limit = 1;
    int n_iter_without_event_limit = -1;
    int random_seed = -1;              // after parse(): the seed in use (a negative one was replaced, tools.cpp:41-50)
    uint32_t random_seed_used = 0;     // srand() argument
    long cpu_start = 0, cpu_last_event = 0;  // clock() at construction / at the last event (physical_model.cpp:286, calcul.cpp:231)
    std::string ini_echo;              // inipp::Ini::generate of the parsed file: the content of <output_dir>/params.ini
    size_t write_events_frequency = 1, write_between_event_frequency = 100, full_aggregate_update_frequency = 1;
    std::string output_dir = "MCAC_output", flame_file = "flame_input", interpotential_file = "interpotential_file";
    bool with_domain_duplication = true, with_domain_reduction = false, with_nucleation = false, with_collisions = true;
    bool with_surface_reactions = false, with_flame_coupling = false, enforce_volume_fraction = true;
    bool individual_surf_reactions = false;
    bool with_potentials = false, with_external_potentials = false, with_dynamic_random_charges = false;
    bool with_electric_charges = false, with_maturity = false;
    int sort_order = MCAC_ORDER_LIBSTDCXX;  // mcac_b200 extension ([numerics] sort_order = libstdcxx|stable)

    PhysicalModel() = default;
    explicit PhysicalModel(const std::string &ini_file);  // reads the file; throws InputError like the reference
    void parse(std::istream &ini);                        // same, from a stream

    double cunningham(double r) const;
    double friction_exponent(double sphere_radius) const;
    double friction_coeff(double aggregate_volume, double sphere_volume, double sphere_radius) const;
    double diffusivity(double f_agg) const;
    static double relax_time(double masse, double f_agg);
    double mobility_diameter(double aggregate_volume, double sphere_volume, double sphere_radius) const;
    double grow(double r, double dt) const;
    void update(size_t n_aggregates, size_t n_monomers, double total_volume, double total_surface) noexcept;
    void update_temperature(double new_temperature) noexcept;
    bool finished(size_t number_of_aggregates, double mean_monomere_per_aggregate) const;
    void print() const;

    mcac_params to_params() const;  // what the device needs
    std::map<std::string, std::string> golden_metadata() const;  // the 6-significant-digit strings of io/physical_model.cpp:30-49
    std::vector<std::pair<std::string, std::string>> xmf_write() const;  // the same, in the order PhysicalModel::xmf_write emits them
};

}  // namespace mcac
