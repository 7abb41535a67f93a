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
#include "xdmf_writer.hpp"

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

// ---- output files (xdmf_writer.hpp) ---------------------------------------------------------------------------------------------
struct mcac_io_writer {
    mcac::XdmfSeriesWriter w;
};
#define MCAC_IO_TRY(stmt)                                                                       \
    try { stmt; } catch (const mcac::BaseException &e) { g_host_err = e.what(); return e.code; } \
    catch (const std::exception &e) { g_host_err = e.what(); return mcac::UNKNOWN_ERROR; }       \
    return mcac::NO_ERROR
int mcac_io_writer_create(const char *prefix, const char *grid_name, int64_t n_time_per_file, int64_t n_for_width, const char *physics,
                          mcac_io_writer **out) {
    if (!prefix || !grid_name || !out || n_time_per_file < 1 || n_for_width < 1) { g_host_err = "mcac_io_writer_create: bad arguments"; return mcac::INPUT_ERROR; }
    std::vector<std::pair<std::string, std::string>> kv;
    std::istringstream is(physics ? physics : "");
    std::string line;
    while (std::getline(is, line)) {
        const size_t eq = line.find('=');
        if (eq != std::string::npos) kv.emplace_back(line.substr(0, eq), line.substr(eq + 1));
    }
    MCAC_IO_TRY(*out = new mcac_io_writer{mcac::XdmfSeriesWriter(prefix, grid_name, (size_t)n_time_per_file, (size_t)n_for_width, kv)});
}
int mcac_io_begin_step(mcac_io_writer *w, double time) { MCAC_IO_TRY(w->w.begin_step(time)); }
int mcac_io_positions(mcac_io_writer *w, const double *xyz, int64_t n_points) { MCAC_IO_TRY(w->w.positions(xyz, (uint64_t)n_points)); }
int mcac_io_attribute(mcac_io_writer *w, const char *name, int32_t type, const void *data, int64_t count, int32_t scalar_on_nodes) {
    if (type < 0 || type > 2) { g_host_err = "mcac_io_attribute: type must be 0 (f64), 1 (i32) or 2 (i64)"; return mcac::INPUT_ERROR; }
    MCAC_IO_TRY(w->w.attribute(name, static_cast<mcac::H5File::Type>(type), data, (uint64_t)count, scalar_on_nodes != 0));
}
int mcac_io_end_step(mcac_io_writer *w) { MCAC_IO_TRY(w->w.end_step()); }
int mcac_io_writer_destroy(mcac_io_writer *w) {
    if (!w) return mcac::NO_ERROR;
    int rc = mcac::NO_ERROR;
    try { w->w.flush(); } catch (const mcac::BaseException &e) { g_host_err = e.what(); rc = e.code; }
    delete w;
    return rc;
}
// SphereList::get_data / AggregatList::get_data (io/sphere_list.cpp:36-57, io/aggregat_list.cpp:36-67) from the downloaded SoA
int mcac_gpu_save(mcac_gpu *h, mcac_io_writer *spheres, mcac_io_writer *aggregates) {
    int64_t ns = 0, na = 0;
    int rc = mcac_gpu_sizes(h, &ns, &na);
    if (rc) { g_host_err = mcac_gpu_last_error(h); return rc; }
    std::vector<double> sf((size_t)(9 * ns)), af((size_t)(21 * na)), scal(20);
    std::vector<int64_t> slab((size_t)ns), sch((size_t)ns), anp((size_t)na), ach((size_t)na);
    rc = mcac_gpu_download_state(h, sf.data(), slab.data(), sch.data(), af.data(), anp.data(), ach.data(), nullptr, nullptr, nullptr, nullptr, scal.data());
    if (rc) { g_host_err = mcac_gpu_last_error(h); return rc; }
    const double time = scal[0];
    auto interleave = [](const double *x, const double *y, const double *z, int64_t n) {
        std::vector<double> p((size_t)(3 * n));
        for (int64_t i = 0; i < n; i++) { p[(size_t)(3 * i)] = x[i]; p[(size_t)(3 * i + 1)] = y[i]; p[(size_t)(3 * i + 2)] = z[i]; }
        return p;
    };
    auto narrow = [](const std::vector<int64_t> &v) { return std::vector<int32_t>(v.begin(), v.end()); };
    try {
        if (spheres) {  // SpheresFields: X,Y,Z,R,... (constants.hpp:34-45)
            mcac::XdmfSeriesWriter &w = spheres->w;
            w.begin_step(time);
            w.attribute("Time", mcac::H5File::F64, &time, 1, false);
            const std::vector<double> pos = interleave(&sf[0], &sf[(size_t)ns], &sf[(size_t)(2 * ns)], ns);
            w.positions(pos.data(), (uint64_t)ns);
            const std::vector<int32_t> q = narrow(sch);
            w.attribute("electric_charge", mcac::H5File::I32, q.data(), (uint64_t)ns);
            w.attribute("Radius", mcac::H5File::F64, &sf[(size_t)(3 * ns)], (uint64_t)ns);
            w.attribute("Label", mcac::H5File::I64, slab.data(), (uint64_t)ns);
            w.end_step();
        }
        if (aggregates) {  // AggregatesFields order (constants.hpp:46-69): RG 0, F_AGG 1, LPM 2, TIME_STEP 3, RMAX 4, VOLUME 5, SURFACE 6, X 7, Y 8,
            // Z 9, RX..RZ 10-12, TIME 13, DP 14, DG_OVER_DP 15, OVERLAPPING 16, COORDINATION_NUMBER 17, ELECTRIC_CHARGE 18, D_M 19, CH_RATIO 20
            mcac::XdmfSeriesWriter &w = aggregates->w;
            auto fld = [&](int k) { return &af[(size_t)(k * na)]; };
            w.begin_step(time);
            w.attribute("Time", mcac::H5File::F64, &time, 1, false);
            const std::vector<double> pos = interleave(fld(7), fld(8), fld(9), na);
            w.positions(pos.data(), (uint64_t)na);
            w.attribute("Rg", mcac::H5File::F64, fld(0), (uint64_t)na);
            w.attribute("Np", mcac::H5File::I64, anp.data(), (uint64_t)na);
            w.attribute("f_agg", mcac::H5File::F64, fld(1), (uint64_t)na);
            w.attribute("lpm", mcac::H5File::F64, fld(2), (uint64_t)na);
            w.attribute("Deltat", mcac::H5File::F64, fld(3), (uint64_t)na);
            w.attribute("Rmax", mcac::H5File::F64, fld(4), (uint64_t)na);
            w.attribute("Volume", mcac::H5File::F64, fld(5), (uint64_t)na);
            w.attribute("Surface", mcac::H5File::F64, fld(6), (uint64_t)na);
            w.attribute("proper_time", mcac::H5File::F64, fld(13), (uint64_t)na);
            w.attribute("coordination_number", mcac::H5File::F64, fld(17), (uint64_t)na);
            w.attribute("overlapping", mcac::H5File::F64, fld(16), (uint64_t)na);
            const std::vector<int32_t> q = narrow(ach);
            w.attribute("electric_charge", mcac::H5File::I32, q.data(), (uint64_t)na);
            w.attribute("d_m", mcac::H5File::F64, fld(19), (uint64_t)na);
            std::vector<int64_t> label((size_t)na);
            for (int64_t i = 0; i < na; i++) label[(size_t)i] = i;
            w.attribute("Label", mcac::H5File::I64, label.data(), (uint64_t)na);
            w.end_step();
        }
    } catch (const mcac::BaseException &e) { g_host_err = e.what(); return e.code; }
    return mcac::NO_ERROR;
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
