// mcac_b200 host layer — initial placement (see placement.hpp).  Reference behaviour restated:
//  * per monomer: 1 draw -> diameter (physical_model.cpp:557-578), then up to N tries of 3 draws -> position,
//    accepted when no already-placed monomer is in contact (aggregat.cpp:162-193, aggregat_list.cpp:534-548,
//    aggregat_distance.cpp:45-58, sphere_distances.cpp:84-90);
//  * the accepted monomer becomes a one-sphere aggregate whose morphology / mobility fields are those of
//    Aggregate::update() (aggregat.cpp:194-229, 247-288);
//  * optional global radius rescale to the prescribed volume fraction (aggregat_list_storage.cpp:75-87).
// The neighbour search is our own uniform hash grid: the reference's Verlet scan is a superset filter, so any
// exact neighbour search takes the same accept/reject decisions and therefore consumes the same RNG draws.
#include "placement.hpp"

#include <cmath>

#include "../csrc/mcac_math.cuh"

namespace mcac {
namespace {
using mcacb::inverse_erfc;  // inverfc / inverf of src/tools/tools.cpp:56-77: the one restatement, shared with the device (csrc/mcac_math.cuh)

struct Grid {  // periodic hash grid of placed monomers
    int n = 1;
    double box = 1., width = 1.;
    std::vector<std::vector<int>> cells;
    void init(double box_length, double typical_radius, int64_t count) {
        box = box_length;
        int want = static_cast<int>(box / (6. * typical_radius));
        const int by_count = static_cast<int>(std::cbrt(static_cast<double>(count))) + 1;
        if (want > by_count) want = by_count;
        n = want < 1 ? 1 : (want > 256 ? 256 : want);
        width = box / n;
        cells.assign(static_cast<size_t>(n) * n * n, {});
    }
    int coord(double x) const {
        int c = static_cast<int>(std::floor(mcacb::periodic_position(x, box) / width));
        return c >= n ? n - 1 : (c < 0 ? 0 : c);
    }
    std::vector<int> &at(int i, int j, int k) { return cells[(static_cast<size_t>(i) * n + j) * n + k]; }
};
}  // namespace

double HostRandom::inverf(double p) { return inverse_erfc(1. - p); }
double HostRandom::normal(double mean, double sigma) { return mean + std::sqrt(2.) * sigma * inverf(2. * uniform() - 1.0); }

InitialState place_monomers(PhysicalModel &pm) {
    if (pm.with_electric_charges) throw InputError("with_electric_charges is not built in mcac_b200 (initial / merged aggregate charges)");
    const int64_t n = static_cast<int64_t>(pm.n_monomeres);
    const double box = pm.box_length;
    HostRandom rng(pm.random_seed_used);  // srand(seed): a negative [numerics] random_seed was replaced by parse() (tools.cpp:41-50)
    const mcacb::Gas gas{pm.gaz_mean_free_path, pm.viscosity, pm.temperature, pm.fractal_dimension, pm.density, pm.with_maturity ? 1 : 0};
    InitialState st;
    st.n_sph = st.n_agg = n;
    st.sphere_fields.assign(static_cast<size_t>(9 * n), 0.);
    st.sphere_charge.assign(static_cast<size_t>(n), 0);
    st.agg_fields.assign(static_cast<size_t>(21 * n), 0.);
    st.agg_charge.assign(static_cast<size_t>(n), 0);
    st.agg_cells.assign(static_cast<size_t>(3 * n), 0);
    st.offsets.resize(static_cast<size_t>(n) + 1);
    st.members.resize(static_cast<size_t>(n));
    st.per_member.assign(static_cast<size_t>(3 * n), 0.);
    auto S = [&](int f, int64_t i) -> double & { return st.sphere_fields[static_cast<size_t>(f * n + i)]; };
    auto A = [&](int f, int64_t i) -> double & { return st.agg_fields[static_cast<size_t>(f * n + i)]; };
    Grid grid;
    grid.init(box, 0.5e-9 * pm.mean_diameter, n);
    double largest_radius = 0.;  // of the placed monomers: bounds the search reach
    const int n_div = static_cast<int>(pm.n_verlet_divisions);

    for (int64_t i = 0; i < n; i++) {
        // --- diameter: one draw (lognormal: Dpm * sigma^(sqrt2 * inverf(2u-1)); normal: mean + sqrt2*sigma*inverf(2u-1))
        double diameter = 0.;
        if (pm.monomeres_initialisation_type == NORMAL_INITIALISATION) {
            diameter = rng.normal(pm.mean_diameter, pm.dispersion_diameter);
        } else {
            if (pm.dispersion_diameter < 1.0) throw InputError("dispersion_diameter cannot be lower than 1");
            diameter = pm.mean_diameter * std::pow(pm.dispersion_diameter, std::sqrt(2.) * HostRandom::inverf(2. * rng.uniform() - 1.0));
        }
        if (diameter <= 0) diameter = pm.mean_diameter;
        diameter = diameter * 1E-9;
        const double radius = diameter * 0.5;
        // --- position: rejection sampling, 3 draws per try, at most N tries
        bool placed = false;
        double px = 0., py = 0., pz = 0.;
        for (int64_t attempt = 0; attempt < n && !placed; attempt++) {
            px = rng.uniform() * box;
            py = rng.uniform() * box;
            pz = rng.uniform() * box;
            const double reach = radius + largest_radius;
            const int span = static_cast<int>(std::floor(reach / grid.width)) + 1;
            const int ci = grid.coord(px), cj = grid.coord(py), ck = grid.coord(pz);
            bool free_space = true;
            const int lo = (2 * span + 1 >= grid.n) ? 0 : -span, hi = (2 * span + 1 >= grid.n) ? grid.n - 1 : span;
            for (int a = lo; a <= hi && free_space; a++)
                for (int b = lo; b <= hi && free_space; b++)
                    for (int c = lo; c <= hi && free_space; c++) {
                        const int ii = (2 * span + 1 >= grid.n) ? a : mcacb::wrap_cell(ci + a, grid.n);
                        const int jj = (2 * span + 1 >= grid.n) ? b : mcacb::wrap_cell(cj + b, grid.n);
                        const int kk = (2 * span + 1 >= grid.n) ? c : mcacb::wrap_cell(ck + c, grid.n);
                        for (int other : grid.at(ii, jj, kk)) {
                            // bounding sphere of the one-monomer aggregate first, then the monomer itself
                            if (!mcacb::spheres_in_contact(px, py, pz, radius, A(7, other), A(8, other), A(9, other), A(4, other), box)) continue;
                            if (mcacb::spheres_in_contact(px, py, pz, radius, S(0, other), S(1, other), S(2, other), S(3, other), box)) {
                                free_space = false;
                                break;
                            }
                        }
                    }
            placed = free_space;
        }
        if (!placed) throw TooDenseError();
        // --- Sphere::init_val + Aggregate::update() of a single sphere
        const double vol = mcacb::volume_factor() * std::pow(radius, 3);
        const double surf = mcacb::surface_factor() * (radius * radius);
        S(0, i) = px; S(1, i) = py; S(2, i) = pz; S(3, i) = radius; S(4, i) = vol; S(5, i) = surf;
        const double V = 0.0 + vol, Sf = 0.0 + surf;
        const double cx = (0. + 0. * vol) / V, cy = (0. + 0. * vol) / V, cz = (0. + 0. * vol) / V;
        const double dcen = std::sqrt((0. - cx) * (0. - cx) + (0. - cy) * (0. - cy) + (0. - cz) * (0. - cz));
        const double ax = mcacb::periodic_position(px + cx, box), ay = mcacb::periodic_position(py + cy, box),
                     az = mcacb::periodic_position(pz + cz, box);
        const double rmax = (0.0 < radius + dcen) ? radius + dcen : 0.0;
        const double arg = 0. + vol * (dcen * dcen), brg = 0. + vol * (radius * radius);
        const double rg = std::sqrt(std::fabs((arg + 3. / 5. * brg) / V));
        const double dp = 2 * (0. + radius) / 1.0, vol_pp = (0.0 + vol) / 1.0;
        double ch_ratio = 0.;
        const mcacb::Mobility mob = mcacb::mobility_epilogue(gas, V, vol_pp, dp, &ch_ratio);
        A(0, i) = rg; A(1, i) = mob.f_agg; A(2, i) = mob.lpm; A(3, i) = mob.time_step; A(4, i) = rmax; A(5, i) = V; A(6, i) = Sf;
        A(7, i) = ax; A(8, i) = ay; A(9, i) = az; A(10, i) = cx; A(11, i) = cy; A(12, i) = cz; A(13, i) = pm.time; A(14, i) = dp;
        A(15, i) = 2 * rg / dp; A(16, i) = 0.; A(17, i) = 0.; A(18, i) = 0.; A(19, i) = mob.d_m; A(20, i) = ch_ratio;
        st.agg_cells[static_cast<size_t>(i)] = mcacb::cell_of(ax, n_div, box);
        st.agg_cells[static_cast<size_t>(n + i)] = mcacb::cell_of(ay, n_div, box);
        st.agg_cells[static_cast<size_t>(2 * n + i)] = mcacb::cell_of(az, n_div, box);
        st.offsets[static_cast<size_t>(i)] = i;
        st.members[static_cast<size_t>(i)] = i;
        st.per_member[static_cast<size_t>(i)] = vol;
        st.per_member[static_cast<size_t>(n + i)] = surf;
        st.per_member[static_cast<size_t>(2 * n + i)] = dcen;
        if (rmax > st.maxradius) st.maxradius = rmax;
        if (radius > largest_radius) largest_radius = radius;
        grid.at(grid.coord(px), grid.coord(py), grid.coord(pz)).push_back(static_cast<int>(i));
    }
    st.offsets[static_cast<size_t>(n)] = n;
    // AggregatList::refresh (aggregat_list.cpp:100-108)
    st.max_time_step = A(3, 0);
    for (int64_t i = 0; i < n; i++) st.max_time_step = (st.max_time_step < A(3, i)) ? A(3, i) : st.max_time_step;
    st.avg_npp = 1.0;
    // enforce_volume_fraction (aggregat_list_storage.cpp:75-87): V/S follow, Rg / f_agg / lpm / dt stay until the next update()
    if (pm.enforce_volume_fraction) {
        double current_total_volume = 0.0;
        for (int64_t i = 0; i < n; i++) current_total_volume += A(5, i);
        const double prescribed_total_volume = pm.volume_fraction * std::pow(box, 3);
        const double correction = std::pow(prescribed_total_volume / current_total_volume, 1. / 3.);
        for (int64_t i = 0; i < n; i++) {
            const double r = S(3, i) * correction;
            S(3, i) = r;
            S(4, i) = mcacb::volume_factor() * std::pow(r, 3);
            S(5, i) = mcacb::surface_factor() * (r * r);
            st.per_member[static_cast<size_t>(i)] = S(4, i);
            st.per_member[static_cast<size_t>(n + i)] = S(5, i);
            A(5, i) = 0.0 + S(4, i);
            A(6, i) = 0.0 + S(5, i);
        }
    }
    st.rand_consumed = rng.calls();
    return st;
}
}  // namespace mcac
