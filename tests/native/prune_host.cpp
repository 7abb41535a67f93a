// Host check of sweep_may_touch (mcac_b200/csrc/mcac_math.cuh), the enclosing-ball pruning of the step loop's sphere-pair sweeps: it
// must never reject a pair whose exact test (pair_contact_distance, the restatement of src/spheres/sphere_contact.cpp:47-125) is
// finite — for spheres anywhere inside their aggregates' enclosing balls, un-wrapped positions, boxes smaller than the sweep
// (several periodic images), grazing and overlapping starts.  usage: prune_host <cases> <seed>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../mcac_b200/csrc/mcac_math.cuh"

using namespace mcacb;

int main(int argc, char **argv) {
    const int cases = argc > 1 ? atoi(argv[1]) : 2000;
    std::mt19937_64 rng(argc > 2 ? atoll(argv[2]) : 1);
    std::uniform_real_distribution<double> U(0., 1.);
    long long finite = 0, pruned = 0, total = 0;
    for (int c = 0; c < cases; c++) {
        const double box = 1e-6 * (0.2 + 3. * U(rng));
        const double Rm = box * (0.01 + 0.4 * U(rng)), Ro = box * (0.01 + 0.4 * U(rng));
        const double dist = box * (c % 5 == 0 ? 2.5 * U(rng) : 0.3 * U(rng));
        double dir[3];
        {
            const double th = 2 * pi() * U(rng), ph = std::acos(1 - 2 * U(rng));
            dir[0] = std::sin(ph) * std::cos(th); dir[1] = std::sin(ph) * std::sin(th); dir[2] = std::cos(ph);
        }
        // centres: the other one placed near the path of the mover so that contacts are common; random box shifts (un-wrapped spheres)
        double cm[3], co[3];
        const double along = dist * (U(rng) * 1.4 - 0.2), off = (Rm + Ro) * 1.3 * U(rng);
        double perp[3] = {U(rng) - .5, U(rng) - .5, U(rng) - .5};
        const double pd = perp[0] * dir[0] + perp[1] * dir[1] + perp[2] * dir[2];
        double pn = 0;
        for (int l = 0; l < 3; l++) { perp[l] -= pd * dir[l]; pn += perp[l] * perp[l]; }
        pn = std::sqrt(pn) + 1e-300;
        for (int l = 0; l < 3; l++) {
            cm[l] = box * U(rng);
            co[l] = cm[l] + along * dir[l] + off * perp[l] / pn + box * (double)((int)(rng() % 5) - 2);
        }
        auto inside = [&](const double *ctr, double R, double out[4]) {  // a sphere enclosed by the ball (ctr, R), sometimes touching it
            const double r = R * (0.02 + 0.3 * U(rng));
            double v[3] = {U(rng) - .5, U(rng) - .5, U(rng) - .5};
            const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + 1e-300;
            const double rho = (U(rng) < 0.3 ? 1.0 : U(rng)) * (R - r);
            for (int l = 0; l < 3; l++) out[l] = ctr[l] + rho * v[l] / n + box * (double)((int)(rng() % 3) - 1);
            out[3] = r;
        };
        std::vector<double> A(4 * 12), B(4 * 12);
        for (int i = 0; i < 12; i++) { inside(cm, Rm, &A[4 * i]); inside(co, Ro, &B[4 * i]); }
        for (int i = 0; i < 12; i++) {
            const double *a = &A[4 * i];
            const bool keep_i = sweep_may_touch(a[0], a[1], a[2], a[3], co[0], co[1], co[2], Ro, dir[0], dir[1], dir[2], dist, box);
            for (int j = 0; j < 12; j++) {
                const double *b = &B[4 * j];
                const bool keep_j = sweep_may_touch(cm[0], cm[1], cm[2], Rm, b[0], b[1], b[2], b[3], dir[0], dir[1], dir[2], dist, box);
                const double d = pair_contact_distance(a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3], dir[0], dir[1], dir[2], dist, box);
                total++;
                if (!(keep_i && keep_j)) pruned++;
                if (std::isfinite(d)) {
                    finite++;
                    if (!keep_i || !keep_j) {
                        printf("case %d: pair (%d,%d) has distance %.17g but was pruned (keep_i %d keep_j %d)\n", c, i, j, d, (int)keep_i, (int)keep_j);
                        return 1;
                    }
                }
            }
        }
    }
    printf("ok %d cases: %lld pairs, %lld finite, %lld pruned\n", cases, total, finite, pruned);
    return finite > 1000 && pruned > 1000 ? 0 : 2;
}
