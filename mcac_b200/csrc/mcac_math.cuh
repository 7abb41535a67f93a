// mcac_b200 — scalar FP64 building blocks shared by every kernel (and by the host-side initial placement).
//
// All functions are __host__ __device__ and are compiled WITHOUT fused multiply-add (nvcc --fmad=false;
// the reference's effective build has no FMA) so that +,-,*,/,sqrt,fmod,floor give the same
// IEEE-754 results on sm_100a as on the reference's x86-64 build.  Only the transcendental calls
// (sincos/acos in the direction, exp/erf/pow in the physics closures) differ from glibc by <= 2 ulp.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define MCAC_HD __host__ __device__ __forceinline__
#else
#define MCAC_HD inline
#endif

namespace mcacb {

constexpr double kContactEpsilon = 1e-28;       // include/constants.hpp:126
constexpr double kCoordinationEpsilon = 1e-10;  // include/constants.hpp:127
constexpr double kBoltzmann = 1.38066E-23;      // include/constants.hpp:115

// constants.hpp:112-114 — `atan(1.0)*4`, 4*pi/3, 4*pi evaluated in double
MCAC_HD double pi() { return 3.141592653589793; }
MCAC_HD double volume_factor() { return 4 * 3.141592653589793 / 3; }
MCAC_HD double surface_factor() { return 4 * 3.141592653589793; }

struct Vec3 {
    double x, y, z;
};

// include/physical_model/physical_model.hpp:99-109
MCAC_HD double periodic_distance(double dist, double dim) {
    double d = dist;
    const double half = 0.5 * dim;
    while (d < -half) d += dim;
    while (d >= half) d -= dim;
    return d;
}
// include/physical_model/physical_model.hpp:113-122
MCAC_HD double periodic_position(double p, double dim) {
    double q = p;
    while (q < 0) q += dim;
    while (q >= dim) q -= dim;
    return q;
}
// src/spheres/sphere_distances.cpp:68-76
MCAC_HD double distance2_periodic(double ax, double ay, double az, double bx, double by, double bz, double box) {
    const double dx = periodic_distance(ax - bx, box);
    const double dy = periodic_distance(ay - by, box);
    const double dz = periodic_distance(az - bz, box);
    return dx * dx + dy * dy + dz * dz;
}
// src/spheres/sphere_distances.cpp:84-90 — `contact()`: d^2 - (r1+r2)^2 <= 1e-28
MCAC_HD bool spheres_in_contact(double ax, double ay, double az, double ra, double bx, double by, double bz, double rb, double box) {
    const double d2 = distance2_periodic(ax, ay, az, bx, by, bz, box);
    const double rc = (ra + rb) * (ra + rb);
    return (d2 - rc <= kContactEpsilon);
}

// fmod(x, box) for box > 0, bit for bit: fmod is exact, so is x -/+ box on box <= |x| < 2 box (Sterbenz), and the sign of a
// zero result follows x as in fmod.  Sphere coordinates stay within a box length of the origin, so the libm loop is rare.
MCAC_HD double fmod_box(double x, double box) {
    const double ax = fabs(x);
    if (ax < box) return x;
    if (ax < 2 * box) {
        const double r = (x < 0) ? x + box : x - box;
        return (r == 0.) ? ((x < 0) ? -0. : 0.) : r;
    }
    return fmod(x, box);
}

// --------------------------------------------------------------------------------------------------
// THE pair test (K1 inner op): distance the sphere (p1, r1) can travel along `dir` (|dir| = 1, at most
// `dist`) before touching sphere (p2, r2) or one of its periodic images; +inf if it never does.
// Replaces distance_to_contact(Sphere, Sphere, dir, dist), src/spheres/sphere_contact.cpp:47-125.
// ~100 FP64 ops + 3 fmod + 3 div + <= 1 sqrt on the full path; 32 B of candidate data.
// --------------------------------------------------------------------------------------------------
MCAC_HD double pair_contact_distance(double p1x, double p1y, double p1z, double r1, double p2x, double p2y, double p2z, double r2,
                                     double dx, double dy, double dz, double dist, double box) {
    const double rsum = r1 + r2;
    const double rsum2 = rsum * rsum;
    if (distance2_periodic(p1x, p1y, p1z, p2x, p2y, p2z, box) <= rsum2) return 0.;
    const double p1[3] = {p1x, p1y, p1z};
    const double dir[3] = {dx, dy, dz};
    const double disp[3] = {dx * dist, dy * dist, dz * dist};
    double p2[3] = {p2x, p2y, p2z};
    double zone[3];
    int nper[3];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const double moved = p1[l] + disp[l];
        const double lo = (moved < p1[l]) ? moved : p1[l];  // std::min(p1, p1 + disp)
        const double hi = (p1[l] < moved) ? moved : p1[l];  // std::max(p1, p1 + disp)
        const double base = lo - rsum;
        const double end = hi + rsum;
        zone[l] = end - base;
        double w = fmod_box(p2[l] - base, box);
        if (w < 0) w += box;
        p2[l] = w + base;
        // floor(zone / box): a quotient of 0 < zone < box rounds to at most 1 - 2^-53, so the division is only needed beyond
        nper[l] = (zone[l] >= 0. && zone[l] < box) ? 0 : static_cast<int>(floor(zone[l] / box));
    }
    double res = INFINITY;
    const double endx = p1[0] + disp[0], endy = p1[1] + disp[1], endz = p1[2] + disp[2];
    for (int i = 0; i <= nper[0]; i++)
        for (int j = 0; j <= nper[1]; j++)
            for (int k = 0; k <= nper[2]; k++) {
                const double qx = p2[0] + i * box, qy = p2[1] + j * box, qz = p2[2] + k * box;
                const double fx = qx - p1[0], fy = qy - p1[1], fz = qz - p1[2];
                if (fabs(fx) > zone[0]) continue;
                if (fabs(fy) > zone[1]) continue;
                if (fabs(fz) > zone[2]) continue;
                const double proj = fx * dir[0] + fy * dir[1] + fz * dir[2];
                if (proj < 0) continue;  // contact in the past
                const double ex = endx - qx, ey = endy - qy, ez = endz - qz;
                const bool end_contact = (ex * ex + ey * ey + ez * ez) <= rsum2;
                if ((!end_contact) && dist < proj) continue;  // too far to reach
                const double cx = fy * dir[2] - fz * dir[1];
                const double cy = fz * dir[0] - fx * dir[2];
                const double cz = fx * dir[1] - fy * dir[0];
                const double axis2 = cx * cx + cy * cy + cz * cz;
                if (axis2 > rsum2) continue;  // passes beside it
                const double c = proj - sqrt(rsum2 - axis2);
                res = (c < res) ? c : res;  // std::min(res, c)
            }
    return res;
}

// Conservative companion of the pair test, for pruning sphere-pair sweeps between big aggregates without changing their result:
// can ANY sphere lying inside the ball (c2, r2) — or the ball itself — give a finite pair_contact_distance against a moving sphere
// that lies inside the ball (p1, r1)?  A finite pair result needs a periodic image q of the static sphere with 0 <= proj <= dist
// (or the end-of-move overlap: proj <= dist + r1' + r2') and an axis distance <= r1' + r2', or an overlap at the start: in every
// case q is within r1' + r2' of the segment [p1', p1' + (dist + r1' + r2') dir].  Moving both spheres to the centres of their
// enclosing balls changes that distance by at most (r1 - r1') + (r2 - r2') and lengthens the segment, hence the test below:
// some image of c2 within r1 + r2 of the segment [p1, p1 + (dist + r1 + r2) dir].  False => every enclosed pair returns +inf.
MCAC_HD bool sweep_may_touch(double p1x, double p1y, double p1z, double r1, double c2x, double c2y, double c2z, double r2,
                             double dx, double dy, double dz, double dist, double box) {
    const double rsum = (r1 + r2) * (1. + 1e-9) + 1e-12 * box;  // slack for the rounding of positions / radii of the enclosed spheres
    const double rsum2 = rsum * rsum;
    const double len = dist + rsum;
    const double p1[3] = {p1x, p1y, p1z}, dir[3] = {dx, dy, dz};
    double c[3] = {c2x, c2y, c2z};
    int nper[3];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const double moved = p1[l] + dir[l] * len;
        const double lo = ((moved < p1[l]) ? moved : p1[l]) - rsum;
        const double hi = ((p1[l] < moved) ? moved : p1[l]) + rsum;
        double w = fmod(c[l] - lo, box);
        if (w < 0) w += box;
        c[l] = w + lo;  // the image in [lo, lo + box); further images at + k box while they are <= hi
        nper[l] = static_cast<int>(floor((hi - lo) / box));
    }
    for (int i = 0; i <= nper[0]; i++)
        for (int j = 0; j <= nper[1]; j++)
            for (int k = 0; k <= nper[2]; k++) {
                const double fx = c[0] + i * box - p1[0], fy = c[1] + j * box - p1[1], fz = c[2] + k * box - p1[2];
                double t = fx * dir[0] + fy * dir[1] + fz * dir[2];
                t = t < 0. ? 0. : (t > len ? len : t);
                const double ex = fx - t * dir[0], ey = fy - t * dir[1], ez = fz - t * dir[2];
                if (ex * ex + ey * ey + ez * ez <= rsum2) return true;
            }
    return false;
}

// --------------------------------------------------------------------------------------------------
// Verlet cell range swept by a move (src/verlet/verlet.cpp:52-79): inclusive, un-wrapped cell indices.
// `reach` = rmax(source) + maxradius; (vx,vy,vz) = distance * direction.
// --------------------------------------------------------------------------------------------------
struct CellRange {
    int lo[3], hi[3];
};
// floor(x / w) evaluated as the reference does (correctly rounded quotient, then floor), without the division in the common
// case: x * inv_w differs from the rounded quotient by a few ulp, so away from an integer both floors agree.
MCAC_HD double floor_div(double x, double w, double inv_w) {
    const double q = x * inv_w;
    const double fq = floor(q);
    const double frac = q - fq;
    if (fabs(q) < 1048576. && frac > 1e-6 && frac < 0.999999) return fq;
    return floor(x / w);
}
MCAC_HD CellRange verlet_range(double sx, double sy, double sz, double vx, double vy, double vz, double reach, int n_div, double width) {
    const double src[3] = {sx, sy, sz};
    const double v[3] = {vx, vy, vz};
    CellRange r;
    const double nd = static_cast<double>(n_div);
    const double inv_w = 1. / width;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const double vp = (0. < v[a]) ? v[a] : 0.;  // std::max(direction, 0.)
        const double vm = (v[a] < 0.) ? v[a] : 0.;  // std::min(direction, 0.)
        const double p = src[a] + reach + vp;
        const double m = src[a] - reach + vm;
        int b1 = static_cast<int>(floor_div(nd * m, width, inv_w));
        int b2 = static_cast<int>(floor_div(nd * p, width, inv_w) + 1);
        if (b2 - b1 >= n_div) {
            b1 = 0;
            b2 = n_div - 1;
        }
        r.lo[a] = b1;
        r.hi[a] = b2;
    }
    return r;
}
MCAC_HD int wrap_cell(int i, int n) {  // periodic_position(i, n_div) on integers
    int m = i % n;
    return m < 0 ? m + n : m;
}
// Aggregate::compute_index_verlet, src/aggregats/aggregat.cpp:699-705
MCAC_HD int cell_of(double x, int n_div, double box) {
    const double step = double(n_div) / box;
    return static_cast<int>(static_cast<unsigned long long>(floor(x * step)));
}
// rank of cell coordinate c inside [lo, hi] scanned with periodic wrap; -1 when outside the range
MCAC_HD int range_rank(int c, int lo, int hi, int n) {
    const int off = wrap_cell(c - lo, n);
    return (off <= hi - lo) ? off : -1;
}

// --------------------------------------------------------------------------------------------------
// Physics closures, src/physical_model/physical_model.cpp:550-617
// --------------------------------------------------------------------------------------------------
struct Gas {
    double mean_free_path, viscosity, temperature, fractal_dimension, density;
    int with_maturity;
};
MCAC_HD double cunningham(const Gas &g, double r) {
    const double a = 1.142, b = 0.558, c = 0.999;
    return 1.0 + a * g.mean_free_path / r + b * g.mean_free_path / r * exp(-c * r / g.mean_free_path);
}
MCAC_HD double friction_exponent(const Gas &g, double r) { return 0.689 * (1. + erf(((g.mean_free_path / r) + 4.454) / 10.628)); }
MCAC_HD double friction_coeff(const Gas &g, double agg_volume, double sphere_volume, double r) {
    const double fe = friction_exponent(g, r);
    const double cc = cunningham(g, r);
    return (6. * pi() * g.viscosity * r / cc) * pow(agg_volume / sphere_volume, fe / g.fractal_dimension);
}
MCAC_HD double mobility_diameter(const Gas &g, double agg_volume, double sphere_volume, double r) {
    const double fe = friction_exponent(g, r);
    const double ra = r * pow(agg_volume / sphere_volume, fe / g.fractal_dimension / 2.0);
    const double cc_pp = cunningham(g, r);
    const double cc_a = cunningham(g, ra);
    return (cc_a / cc_pp) * 2.0 * r * pow(agg_volume / sphere_volume, fe / g.fractal_dimension);
}
// Aggregate::set_bulk_density / set_CH_ratio, src/aggregats/aggregat.cpp:119-147; returns density, writes CH ratio
MCAC_HD double bulk_density(const Gas &g, double dp, double *ch_ratio) {
    if (!g.with_maturity) return g.density;
    const double dpp_nm = dp * 1e+09;
    const double ch = 0.5 * (erf((dpp_nm - 4.0) / 1.0) + 1.0) * (10. - 1.1) + 1.1;
    *ch_ratio = ch;
    return 1200. + (1800. - 1200.) / (10. - 1.1) * (ch - 1.1);
}
// tail of Aggregate::update_partial, aggregat.cpp:264-276: (V, mean sphere V, dp, rg) -> f_agg, d_m, time_step, lpm
struct Mobility {
    double f_agg, d_m, time_step, lpm, bulk_density;
};
MCAC_HD Mobility mobility_epilogue(const Gas &g, double agg_volume, double vol_pp, double dp, double *ch_ratio) {
    Mobility m;
    m.bulk_density = bulk_density(g, dp, ch_ratio);
    m.f_agg = friction_coeff(g, agg_volume, vol_pp, 0.5 * dp);
    m.d_m = mobility_diameter(g, agg_volume, vol_pp, 0.5 * dp);
    const double masse = m.bulk_density * agg_volume;
    const double relax_time = masse / m.f_agg;
    m.time_step = 3. * relax_time;
    const double diffusivity = kBoltzmann * g.temperature / m.f_agg;
    m.lpm = sqrt(6. * diffusivity * m.time_step);
    return m;
}

// random_direction(), src/tools/tools.cpp:82-89, on the two uniform draws (theta first, then phi)
MCAC_HD Vec3 direction_from_draws(double u_theta, double u_phi) {
    const double theta = u_theta * 2 * pi();
    const double phi = acos(1 - 2 * u_phi);
    double st, ct, sp, cp;
#if defined(__CUDA_ARCH__)
    sincos(theta, &st, &ct);
    sincos(phi, &sp, &cp);
#else
    st = sin(theta); ct = cos(theta); sp = sin(phi); cp = cos(phi);
#endif
    return {sp * ct, sp * st, cp};
}
// mcac::random(), src/tools/tools.cpp:51-55: rand()/RAND_MAX in [0,1] inclusive
MCAC_HD double uniform_from_rand(int32_t v) { return static_cast<double>(v) / 2147483647.; }

// Lens caps removed from each sphere of an overlapping pair, Intersection::Intersection,
// src/spheres/sphere_intersection.cpp:29-66.  out = {volume_1, volume_2, surface_1, surface_2}
MCAC_HD void lens_caps(double r1, double v1, double s1, double r2, double v2, double s2, double dist, double out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0.;
    if (dist <= 0) return;
    if (dist < r1 + r2) {
        const double fdim12 = (r1 > r2) ? r1 - r2 : 0.;  // std::fdim(r1, r2) — NOT |r1-r2|
        if (dist >= fdim12) {
            const double h1 = (r2 * r2 - (r1 - dist) * (r1 - dist)) / (2. * dist);
            const double h2 = (r1 * r1 - (r2 - dist) * (r2 - dist)) / (2. * dist);
            out[0] = pi() * (h1 * h1) * (3 * r1 - h1) / 3.;
            out[1] = pi() * (h2 * h2) * (3 * r2 - h2) / 3.;
            out[2] = 2 * pi() * r1 * h1;
            out[3] = 2 * pi() * r2 * h2;
        } else if (r1 < r2) {
            out[0] = v1;
            out[2] = s1;
        } else {
            out[1] = v2;
            out[3] = s2;
        }
    }
}
// aggregat.cpp:289-319
MCAC_HD double volume_alpha_correction(double cn, double c20, double c30, double min_cn, double extreme) {
    const double diff = fabs(cn - min_cn);
    double correction = 0.25 * (3.0 * c20 - c30) * cn - c30 * diff * 0.62741833 - pow(diff, 1.5) * 0.00332425;
    if (correction < 0.0) correction = 1.0;
    correction = (1.0 < correction) ? 1.0 : correction;
    const double alpha = 1.0 - correction;
    return (alpha < extreme) ? extreme : alpha;
}
MCAC_HD double surface_alpha_correction(double cn, double c10, double min_cn, double extreme) {
    const double diff = fabs(cn - min_cn);
    double correction = 0.5 * c10 * cn - (c10 * c10) * diff * 0.70132500 - (diff * diff) * 0.00450000;
    if (correction < 0.0) correction = 1.0;
    correction = (1.0 < correction) ? 1.0 : correction;
    const double alpha = 1.0 - correction;
    return (alpha < extreme) ? extreme : alpha;
}

// inverfc / inverf of src/tools/tools.cpp:56-77 (rational start + two Halley steps), used by random_diameter
MCAC_HD double inverse_erfc(double p) {
    if (p >= 2.) return -100.;
    if (p <= 0.0) return 100.;
    const double pp = (p < 1.0) ? p : 2. - p;
    const double t = sqrt(-2. * log(pp / 2.));
    double x = -0.70711 * ((2.30753 + t * 0.27061) / (1. + t * (0.99229 + t * 0.04481)) - t);
    for (int it = 0; it < 2; it++) {
        const double err = erfc(x) - pp;
        x += err / (1.12837916709551257 * exp(-(x * x)) - x * err);
    }
    return (p < 1.0 ? x : -x);
}
// PhysicalModel::random_diameter (physical_model.cpp:557-578) on the uniform draw; metres
MCAC_HD double diameter_from_draw(double u, double mean, double dispersion, int normal_law) {
    double diameter;
    if (normal_law) diameter = mean + sqrt(2.) * dispersion * inverse_erfc(1. - (2. * u - 1.0));
    else diameter = mean * pow(dispersion, sqrt(2.) * inverse_erfc(1. - (2. * u - 1.0)));
    if (diameter <= 0) diameter = mean;
    return diameter * 1E-9;
}
// interpolate_2d, src/tools/tools.cpp:162-175
MCAC_HD double interpolate_2d(double f11, double f12, double f21, double f22, double dx, double dy) {
    const double df_x = f21 - f11, df_y = f12 - f11, df_xy = (f11 + f22) - (f21 + f12);
    return df_x * dx + df_y * dy + df_xy * dx * dy + f11;
}

// --------------------------------------------------------------------------------------------------
// glibc rand() TYPE_3 additive-feedback stream: r[i] = r[i-31] + r[i-3] (mod 2^32),
// output r[i] >> 1.  State = the last 31 words in a ring.  Reproduced so that a trajectory can be replayed
// draw for draw against the reference (which calls libc srand/rand, src/tools/tools.cpp:41-55).
// --------------------------------------------------------------------------------------------------
struct GlibcRandState {
    uint32_t ring[31];
    int32_t pos;
};
inline void glibc_srand(GlibcRandState &s, uint32_t seed) {
    uint32_t r[344];
    if (seed == 0) seed = 1;
    r[0] = seed;
    for (int i = 1; i < 31; i++) {
        const long hi = (long)((int32_t)r[i - 1]) / 127773, lo = (long)((int32_t)r[i - 1]) % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (uint32_t)w;
    }
    for (int i = 31; i < 34; i++) r[i] = r[i - 31];
    for (int i = 34; i < 344; i++) r[i] = r[i - 31] + r[i - 3];
    for (int i = 0; i < 31; i++) s.ring[i] = r[344 - 31 + i];
    s.pos = 0;
}
MCAC_HD int32_t glibc_rand_next(GlibcRandState &s) {
    int p3 = s.pos + 28;
    if (p3 >= 31) p3 -= 31;
    const uint32_t v = s.ring[s.pos] + s.ring[p3];
    s.ring[s.pos] = v;
    s.pos = (s.pos + 1 == 31) ? 0 : s.pos + 1;
    return (int32_t)(v >> 1);
}

}  // namespace mcacb
