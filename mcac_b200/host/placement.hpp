// mcac_b200 host layer — initial monomer placement (reference a23: AggregatList ctor, Aggregate::init x2,
// test_free_space, random_diameter, enforce_volume_fraction block).  Runs once per realization on
// the host because the rejection sampling is strictly sequential in the RNG stream; the result is uploaded to HBM.
#pragma once
#include <cstdint>
#include <vector>

#include "../csrc/mcac_math.cuh"
#include "physical_model.hpp"

namespace mcac {

// The realization in the reference's own order (label == index, spheres in creation order); same layout as the
// C ABI's upload/download (include/mcac_b200.h).
struct InitialState {
    int64_t n_sph = 0, n_agg = 0;
    std::vector<double> sphere_fields;  // 9 x n_sph : X,Y,Z,R,VOLUME,SURFACE,RX,RY,RZ
    std::vector<int64_t> sphere_charge;
    std::vector<double> agg_fields;     // 21 x n_agg : AggregatesFields order
    std::vector<int64_t> agg_charge, agg_cells, offsets, members;
    std::vector<double> per_member;     // 3 x n_sph : volumes, surfaces, distances_center
    double maxradius = 0., max_time_step = 0., avg_npp = 1.;
    int64_t rand_consumed = 0;          // rand() calls made so far (the device stream continues from here)
};

// mcac::random() family on a private glibc-compatible stream (src/tools/tools.cpp:41-81)
class HostRandom {
  public:
    explicit HostRandom(uint32_t seed) { mcacb::glibc_srand(state_, seed); }
    double uniform() { calls_++; return mcacb::uniform_from_rand(mcacb::glibc_rand_next(state_)); }
    double normal(double mean, double sigma);
    int64_t calls() const { return calls_; }
    static double inverf(double p);
  private:
    mcacb::GlibcRandState state_{};
    int64_t calls_ = 0;
};

InitialState place_monomers(PhysicalModel &pm);  // throws TooDenseError / InputError like the reference

}  // namespace mcac
