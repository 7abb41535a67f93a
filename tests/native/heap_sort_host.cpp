// Host check of mcac_b200/csrc/heap_sort.cuh against libstdc++: (1) heap_sort_segment == std::partial_sort(first, last, last) on
// index arrays with many equal keys (the order among ties is what matters), (2) a whole introsort with a forced small depth limit
// — libstdc++'s __introsort_loop + __final_insertion_sort — equals partition levels replayed by libstdc++ itself down to the limit
// followed by heap_sort_segment on every segment still longer than 16 (the structure k_sort_heap relies on).
// usage: heap_sort_host <cases> <seed>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include "../../mcac_b200/csrc/heap_sort.cuh"

int main(int argc, char **argv) {
    const int cases = argc > 1 ? atoi(argv[1]) : 200;
    std::mt19937_64 rng(argc > 2 ? atoll(argv[2]) : 1);
    for (int c = 0; c < cases; c++) {
        const int n = 2 + (int)(rng() % 3000);
        const int distinct = 1 + (int)(rng() % (c % 3 == 0 ? 3 : n));
        std::vector<double> key(n);
        for (auto &k : key) k = (double)(rng() % distinct) * 0.25;
        // (1) the heap sort alone
        std::vector<size_t> truth(n);
        std::iota(truth.begin(), truth.end(), 0);
        auto cmp = [&key](size_t a, size_t b) { return key[a] < key[b]; };
        std::partial_sort(truth.begin(), truth.end(), truth.end(), cmp);
        std::vector<double> wk(key);
        std::vector<int> perm(n);
        std::iota(perm.begin(), perm.end(), 0);
        heapsort::heap_sort_segment(heapsort::HeapView{wk.data(), perm.data(), 0}, 0, n);
        for (int i = 0; i < n; i++)
            if ((size_t)perm[i] != truth[i] || wk[i] != key[truth[i]]) { printf("heap mismatch case %d n %d at %d\n", c, n, i); return 1; }
        // (2) sub-range: only [f, l) is touched
        if (n > 40) {
            const int f = (int)(rng() % (n / 2)), l = f + 17 + (int)(rng() % (n - f - 17));
            std::vector<size_t> t2(n);
            std::iota(t2.begin(), t2.end(), 0);
            std::partial_sort(t2.begin() + f, t2.begin() + l, t2.begin() + l, cmp);
            std::vector<double> w2(key);
            std::vector<int> p2(n);
            std::iota(p2.begin(), p2.end(), 0);
            heapsort::heap_sort_segment(heapsort::HeapView{w2.data(), p2.data(), 0}, f, l);
            for (int i = 0; i < n; i++)
                if ((size_t)p2[i] != t2[i]) { printf("sub-range mismatch case %d\n", c); return 1; }
        }
        // (3) introsort with a small depth limit == std::sort's own loop at that limit
        for (int depth = 0; depth <= 3; depth++) {
            std::vector<size_t> a(n), b(n);
            std::iota(a.begin(), a.end(), 0);
            b = a;
            auto icmp = __gnu_cxx::__ops::__iter_comp_iter(cmp);
            std::__introsort_loop(a.begin(), a.end(), (long)depth, icmp);
            std::__final_insertion_sort(a.begin(), a.end(), icmp);
            // replay: `depth` partition levels by libstdc++'s own partition, then heap_sort_segment on what is still > 16
            struct Seg { int f, l; };
            std::vector<Seg> segs{{0, n}};
            for (int lev = 0; lev < depth; lev++) {
                std::vector<Seg> next;
                for (const Seg s : segs) {
                    if (s.l - s.f <= 16) { next.push_back(s); continue; }
                    auto cut = std::__unguarded_partition_pivot(b.begin() + s.f, b.begin() + s.l, icmp);
                    const int cpos = (int)(cut - b.begin());
                    next.push_back({s.f, cpos});
                    next.push_back({cpos, s.l});
                }
                segs.swap(next);
            }
            std::vector<double> w3(n);
            std::vector<int> p3(n);
            for (int i = 0; i < n; i++) { p3[i] = (int)b[i]; w3[i] = key[b[i]]; }
            for (const Seg s : segs)
                if (s.l - s.f > 16) heapsort::heap_sort_segment(heapsort::HeapView{w3.data(), p3.data(), 0}, s.f, s.l);
            std::vector<size_t> fin(n);
            for (int i = 0; i < n; i++) fin[i] = (size_t)p3[i];
            std::__final_insertion_sort(fin.begin(), fin.end(), icmp);
            if (fin != a) { printf("introsort-with-limit mismatch case %d depth %d n %d\n", c, depth, n); return 1; }
        }
    }
    printf("ok %d cases\n", cases);
    return 0;
}
