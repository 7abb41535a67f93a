// mcac_b200 host layer — C shims over the host mirror (PhysicalModel + initial placement) so that tests, bench.py
// and foreign-language hosts can build a realization from an .ini text and hand it to the device engine.
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/mcac_b200.h"
#include "physical_model.hpp"
#include "placement.hpp"

struct mcac_host_model {
    mcac::PhysicalModel pm;
    mcac::InitialState st;
    std::string err;
};

static thread_local std::string g_host_err;

extern "C" {
const char *mcac_host_last_error() { return g_host_err.c_str(); }

int mcac_host_model_create(const char *ini_text, int place, mcac_host_model **out) {
    *out = nullptr;
    auto *m = new mcac_host_model();
    try {
        std::istringstream is(ini_text ? ini_text : "");
        m->pm.parse(is);
        if (place) m->st = mcac::place_monomers(m->pm);
    } catch (const mcac::BaseException &e) {
        g_host_err = e.what();
        const int code = e.code;
        delete m;
        return code;
    } catch (const std::exception &e) {
        g_host_err = e.what();
        delete m;
        return mcac::UNKNOWN_ERROR;
    }
    *out = m;
    return mcac::NO_ERROR;
}
void mcac_host_model_destroy(mcac_host_model *m) { delete m; }
int mcac_host_model_params(const mcac_host_model *m, mcac_params *out) { *out = m->pm.to_params(); return 0; }
int mcac_host_model_sizes(const mcac_host_model *m, int64_t *n_sph, int64_t *n_agg) { *n_sph = m->st.n_sph; *n_agg = m->st.n_agg; return 0; }
// writes "key=value\n" lines of the golden metadata (io/physical_model.cpp:30-49)
int mcac_host_model_metadata(const mcac_host_model *m, char *buf, int64_t cap) {
    std::string s;
    for (const auto &kv : m->pm.golden_metadata()) s += kv.first + "=" + kv.second + "\n";
    if ((int64_t)s.size() + 1 > cap) return mcac::INPUT_ERROR;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}
int mcac_host_model_ini_echo(const mcac_host_model *m, char *buf, int64_t cap) {
    const std::string &s = m->pm.ini_echo;
    if ((int64_t)s.size() + 1 > cap) return mcac::INPUT_ERROR;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}
// derived PhysicalModel scalars for parity checks: box_length, box_volume, viscosity, gaz_mean_free_path, mean_massic_radius,
// friction_exponnant, u_sg, aggregate_concentration, total_volume_concent, total_surface_concent, mass_nuclei, volume_fraction
int mcac_host_model_derived(const mcac_host_model *m, double out[12]) {
    const mcac::PhysicalModel &p = m->pm;
    const double v[12] = {p.box_length, p.box_volume, p.viscosity, p.gaz_mean_free_path, p.mean_massic_radius, p.friction_exponnant, p.u_sg,
                          p.aggregate_concentration, p.total_volume_concent, p.total_surface_concent, p.mass_nuclei, p.volume_fraction};
    std::memcpy(out, v, sizeof(v));
    return 0;
}
int mcac_host_model_state(const mcac_host_model *m, double *sphere_fields, double *agg_fields, int64_t *agg_cells, int64_t *offsets,
                          int64_t *members, double *per_member, double *scalars /* maxradius, max_time_step, avg_npp */, int64_t *rand_consumed) {
    const mcac::InitialState &s = m->st;
    auto cp = [](auto *dst, const auto &v) { if (dst) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(sphere_fields, s.sphere_fields);
    cp(agg_fields, s.agg_fields);
    cp(agg_cells, s.agg_cells);
    cp(offsets, s.offsets);
    cp(members, s.members);
    cp(per_member, s.per_member);
    if (scalars) { scalars[0] = s.maxradius; scalars[1] = s.max_time_step; scalars[2] = s.avg_npp; }
    if (rand_consumed) *rand_consumed = s.rand_consumed;
    return 0;
}
// Interpotential(file) of the reference (physical_model_interpotential.cpp:46-121): 4 header values, the charge / dp1 / dp2 axes,
// then (E_barr, E_well) pairs ordered dp1-major, dp2, charge1, charge2; handed to the device as [q1][q2][dp1][dp2] tables.
static int load_interpotential(mcac_gpu *h, const std::string &file) {
    std::ifstream f(file);
    if (!f) { g_host_err = " Interpotential file does not exist: " + file; return mcac::IO_ERROR; }
    double t, d;
    f >> t;
    f >> d; const int n1 = static_cast<int>(d);
    f >> d; const int n2 = static_cast<int>(d);
    f >> d; const int nq = static_cast<int>(d);
    std::vector<int32_t> q((size_t)nq);
    std::vector<double> a((size_t)n1), b((size_t)n2), eb((size_t)nq * nq * n1 * n2), ew(eb.size());
    for (int i = 0; i < nq; i++) { f >> d; q[(size_t)i] = static_cast<int>(d); }
    for (int i = 0; i < n1; i++) f >> a[(size_t)i];
    for (int i = 0; i < n2; i++) f >> b[(size_t)i];
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++)
            for (int k = 0; k < nq; k++)
                for (int l = 0; l < nq; l++) {
                    const size_t at = (((size_t)k * nq + l) * n1 + i) * n2 + j;
                    f >> eb[at] >> ew[at];
                }
    if (!f) { g_host_err = "Interpotential file is truncated: " + file; return mcac::IO_ERROR; }
    return mcac_gpu_set_interpotential(h, n1, n2, nq, q.data(), a.data(), b.data(), eb.data(), ew.data());
}

// PhysicalModel(ini) + AggregatList(&physicalmodel) of the reference's main() (src/main.cpp:26-56): placement on the host,
// state uploaded to HBM, RNG stream continued on the device.
int mcac_sim_create(const char *ini_text, int device, mcac_gpu **out) {
    mcac_host_model *m = nullptr;
    *out = nullptr;
    const int rc = mcac_host_model_create(ini_text, 1, &m);
    if (rc) return rc;
    mcac_params prm = m->pm.to_params();
    mcac_gpu *h = nullptr;
    int r = mcac_gpu_create(&prm, device, &h);
    if (r == 0) r = mcac_gpu_set_rng(h, prm.random_seed, m->st.rand_consumed);
    if (r == 0)
        r = mcac_gpu_upload_state(h, m->st.n_sph, m->st.n_agg, m->st.sphere_fields.data(), m->st.sphere_charge.data(), m->st.agg_fields.data(),
                                  m->st.agg_charge.data(), m->st.agg_cells.data(), m->st.offsets.data(), m->st.members.data(),
                                  m->st.per_member.data(), m->st.maxradius, m->st.max_time_step);
    if (r == 0 && m->pm.with_potentials && m->pm.with_external_potentials) {
        r = load_interpotential(h, m->pm.interpotential_file);  // path relative to the CWD, like the reference (:49)
        if (r) { mcac_host_model_destroy(m); if (h) mcac_gpu_destroy(h); return r; }
    }
    if (r) g_host_err = h ? mcac_gpu_last_error(h) : "mcac_gpu_create failed";
    mcac_host_model_destroy(m);
    if (r) { if (h) mcac_gpu_destroy(h); return r; }
    *out = h;
    return 0;
}
}
