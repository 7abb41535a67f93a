// Host check of mcac_b200/csrc/tie_sort.cuh (the K9 fast path for tie-dominated pick tables) against libstdc++'s own
// std::sort: plan_build + dense_route must place every W element where std::sort puts it, and the handed-over segment,
// continued with libstdc++'s __introsort_loop at the recorded depth, must complete the same permutation.
// usage: tie_sort_host <cases> <seed>   -> prints "ok <cases> ..." or the first mismatch (exit 1)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include "../../mcac_b200/csrc/tie_sort.cuh"

using namespace tiesort;

static int run_case(int n, const std::vector<double> &w, int hand_min, long long *levels_out, long long *handed_out) {
    // truth: sort_indexes of the reference (aggregat_list.cpp:109-123)
    std::vector<size_t> truth(n);
    std::iota(truth.begin(), truth.end(), 0);
    auto cmp = [&w](size_t a, size_t b) { return w[a] < w[b]; };
    std::sort(truth.begin(), truth.end(), cmp);

    double W = w[0];
    for (double v : w) W = std::max(W, v);
    std::vector<int> st_pos;
    std::vector<double> st_w;
    for (int i = 0; i < n; i++)
        if (w[i] != W) { st_pos.push_back(i); st_w.push_back(w[i]); }
    const int x = (int)st_pos.size();
    if (x > kMaxSparse) return 0;
    int lg = 0;
    while ((1LL << (lg + 1)) <= n) lg++;
    const int xcap = std::max(1, x);
    std::vector<int> R((size_t)kMaxLevels * xcap + 2), tbl((size_t)kMaxLevels * kTblStride);
    std::vector<int> a_s(xcap + 2), a_i(xcap + 2), b_s(xcap + 2), b_i(xcap + 2), s_tbl(kTblStride), s_misc(kMiscInts);
    Plan plan{};
    SerialTeam tm;
    int nb = 256;
    while (nb < x && nb < kBuckets) nb <<= 1;
    std::vector<int> arch_R((size_t)kMaxLevels * xcap + 2), arch_T((size_t)kMaxLevels * (nb + 3));
    plan_build(tm, n, x, st_pos.data(), st_w.data(), W, 2 * lg, hand_min, &plan, R.data(), tbl.data(), xcap, a_s.data(), a_i.data(),
               b_s.data(), b_i.data(), s_tbl.data(), s_misc.data(), arch_R.data(), arch_T.data(), kMaxLevels);
    if (plan.fail) { std::printf("plan.fail on n=%d x=%d\n", n, x); return 0; }
    *levels_out += plan.n_levels;
    std::vector<long long> out(n, -1);
    const int hf = plan.hand_f, hl = plan.hand_l;
    *handed_out += hl - hf;
    std::vector<char> is_sparse(n, 0);
    for (int id = 0; id < x; id++) is_sparse[st_pos[id]] = 1;
    for (int j = 0; j < x; j++) {
        const int p = a_s[j];
        if (j > 0 && a_s[j - 1] >= p) { std::printf("sparse list not ascending at %d\n", j); return 1; }
        if (p < hf || p >= hl || out[p] != -1) { std::printf("sparse element %d at %d outside the handed segment [%d,%d)\n", j, p, hf, hl); return 1; }
        out[p] = st_pos[a_i[j]];
    }
    for (int i = 0; i < n; i++) {
        if (is_sparse[i]) continue;
        bool handed = false, bad = false;
        const int p = dense_route(plan, R.data(), tbl.data(), xcap, kTblStride, i, handed, bad);
        if (bad) { std::printf("dense_route: depth limit on n=%d\n", n); return 0; }
        if (p < 0 || p >= n || out[p] != -1) { std::printf("n=%d x=%d: element %d -> %d collides / out of range\n", n, x, i, p); return 1; }
        if (handed && (p < hf || p >= hl)) { std::printf("handed element outside the segment\n"); return 1; }
        if (!handed && p >= hf && p < hl) { std::printf("final element inside the handed segment\n"); return 1; }
        out[p] = i;
    }
    for (int i = 0; i < n; i++)
        if (out[i] < 0) { std::printf("n=%d: position %d left empty\n", n, i); return 1; }
    // the inverse walk (block 0 of the event kernel fills the handed-over segment with it), on the archived tables
    for (int p = hf; p < hl; p++) {
        if (is_sparse[out[p]]) continue;
        const int o = dense_origin(plan, arch_R.data(), arch_T.data(), x, nb + 3, p);
        if (o != (int)out[p]) { std::printf("dense_origin: n=%d x=%d position %d -> %d, expected %lld\n", n, x, p, o, out[p]); return 1; }
    }
    std::vector<size_t> arr(out.begin(), out.end());
    if (hl - hf > 1) {
        if (hl - hf > kLeaf) std::__introsort_loop(arr.begin() + hf, arr.begin() + hl, (long)plan.hand_depth, __gnu_cxx::__ops::__iter_comp_iter(cmp));
    }
    std::__final_insertion_sort(arr.begin(), arr.end(), __gnu_cxx::__ops::__iter_comp_iter(cmp));
    for (int i = 0; i < n; i++)
        if (arr[i] != truth[i]) {
            std::printf("MISMATCH n=%d x=%d levels=%d hand=[%d,%d) at %d: got %zu want %zu\n", n, x, plan.n_levels, hf, hl, i, arr[i], truth[i]);
            return 1;
        }
    return 0;
}

int main(int argc, char **argv) {
    const int cases = argc > 1 ? std::atoi(argv[1]) : 200;
    const unsigned seed = argc > 2 ? (unsigned)std::atoi(argv[2]) : 1u;
    std::mt19937_64 rng(seed);
    long long levels = 0, handed = 0, total_n = 0;
    for (int c = 0; c < cases; c++) {
        int n;
        switch (c % 5) {
            case 0: n = 17 + (int)(rng() % 200); break;
            case 1: n = 200 + (int)(rng() % 5000); break;
            case 2: n = 5000 + (int)(rng() % 60000); break;
            case 3: n = 60000 + (int)(rng() % 300000); break;
            default: n = 1 << (10 + (int)(rng() % 9)); n += (int)(rng() % 3) - 1; break;
        }
        const double frac = (c % 7 == 0) ? 0. : std::pow(10., -4. + 3.5 * (double)(rng() % 1000) / 1000.);
        const int distinct = 1 + (int)(rng() % 6);  // sparse weights drawn from a few classes (dimers, trimers...) or continuous
        std::vector<double> w(n, 7.25);
        std::uniform_real_distribution<double> U(0., 1.);
        for (int i = 0; i < n; i++)
            if (U(rng) < frac) w[i] = (distinct == 6) ? 1. + 6. * U(rng) : 1. + (double)(rng() % distinct);
        const int hand_min = (c % 3 == 0) ? 16 : (c % 3 == 1 ? 256 : 4096);
        if (run_case(n, w, hand_min, &levels, &handed)) return 1;
        total_n += n;
    }
    // all_equal_final: the segment-relative fast form against the depth-tracking form
    for (int c = 0; c < 20000; c++) {
        const int cf = (int)(rng() % 1000), m = 1 + (int)(rng() % (c % 2 ? 3000000 : 300)), cl = cf + m, pos = cf + (int)(rng() % m);
        int bits = 0;
        while (((unsigned)m >> bits) != 0) bits++;
        bool bad1 = false, bad2 = false;
        const int p1 = all_equal_final(cf, cl, pos, 1000, bad1), p2 = all_equal_final(cf, cl, pos, bits - 1, bad2);
        if (bad1 || (!bad2 && p1 != p2) || p1 < cf || p1 >= cl) { std::printf("all_equal_final mismatch m=%d pos=%d: %d vs %d\n", m, pos - cf, p1, p2); return 1; }
    }
    std::printf("ok %d cases, %lld elements, %lld sparse levels, %lld elements handed over\n", cases, total_n, levels, handed);
    return 0;
}
