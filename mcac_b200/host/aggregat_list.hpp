// mcac_b200 host layer — the class surface mcac::calcul consumes (include/aggregats/aggregat_list.hpp:43-116,
// include/aggregats/aggregat.hpp:83-133, include/tools/contact_info.hpp:22-85 of the reference), with every body
// forwarded to the C ABI of include/mcac_b200.h.  State lives in HBM; this object owns the handle.
#pragma once
#include <array>
#include <cstddef>
#include <limits>
#include <vector>

#include "../../include/mcac_b200.h"
#include "physical_model.hpp"

namespace mcac {

// AggregateContactInfo with ids instead of weak_ptrs (default = "no contact": +inf, ids -1)
struct AggregateContactInfo {
    double distance = std::numeric_limits<double>::infinity();
    long moving_sphere = -1, other_sphere = -1, moving_aggregate = -1, other_aggregate = -1;
    bool operator<=(double d) const { return distance <= d; }
    bool operator<(const AggregateContactInfo &o) const { return distance < o.distance; }
};

class AggregatList;

// One aggregate of the list, resident in HBM and addressed by its label: the methods mcac::calcul calls on `aggregates[i]`
// (include/aggregats/aggregat.hpp:83-133).  `aggregates[num_agg]->translate(v)` reads as in src/calcul.cpp:144.
class Aggregate {
  public:
    Aggregate(AggregatList *owner, size_t label) : list(owner), label(label) {}
    Aggregate *operator->() { return this; }
    void translate(const std::array<double, 3> &vector);
    void update();
    void update_partial();
    double get_lpm() const;         // *lpm
    double get_time_step() const;   // *time_step
    double get_rg() const;
    size_t size() const;            // n_spheres
    size_t get_label() const { return label; }

  private:
    AggregatList *list;
    size_t label;
};

// The sphere list of the AggregatList (include/spheres/sphere_list.hpp): size and output
class SphereList {
  public:
    explicit SphereList(AggregatList *owner) : list(owner) {}
    size_t size() const;

  private:
    AggregatList *list;
};

class AggregatList {
  public:
    explicit AggregatList(PhysicalModel *physicalmodel, int device = 0);  // placement (host) + upload + RNG continuation
    ~AggregatList() noexcept;
    AggregatList(const AggregatList &) = delete;
    AggregatList &operator=(const AggregatList &) = delete;

    size_t size() const;
    size_t n_spheres() const;
    double get_avg_npp() const { return avg_npp; }
    double get_max_time_step() const { return max_time_step; }
    double get_time_step(double max) const;  // max / cumulative.back()
    double get_total_volume() const { return total_volume; }
    double get_total_surface() const { return total_surface; }
    double random();                          // mcac::random() on the handle's glibc-compatible stream
    std::array<double, 3> random_direction();
    size_t pick_random();
    size_t pick_last();
    void sort_time_steps(double factor);
    void refresh();
    void duplication();
    AggregateContactInfo distance_to_next_contact(size_t source, const std::array<double, 3> &direction, double distance) const;
    bool merge(AggregateContactInfo contact_info);
    bool croissance_surface(double dt);
    bool croissance_surface(double dt, size_t index);
    void translate(size_t label, const std::array<double, 3> &vector);  // aggregates[label]->translate(vector)
    void update(long label = -1);                                       // Aggregate::update()   (label < 0: all)
    void update_partial(long label = -1);                               // Aggregate::update_partial()
    mcac_run_report run(long max_steps, int batch = 0);                 // the whole calcul() loop on the device
    mcac_gpu *handle() const { return gpu; }
    Aggregate operator[](size_t label) { return Aggregate(this, label); }  // aggregates[i]->...
    SphereList spheres{this};
    std::array<double, 21> fields(size_t label, size_t *n_spheres = nullptr) const;  // the AggregatesFields of one aggregate

  private:
    void check(int rc) const;
    PhysicalModel *physicalmodel;
    mcac_gpu *gpu = nullptr;
    double avg_npp = 1., max_time_step = 0., total_volume = 0., total_surface = 0., last_cum_total = 0.;
};

void calcul(PhysicalModel &physicalmodel, AggregatList &aggregates);  // src/calcul.cpp:55-290

}  // namespace mcac
