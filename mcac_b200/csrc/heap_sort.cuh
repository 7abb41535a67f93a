// mcac_b200 — libstdc++'s heap-sort branch of std::sort (std::__introsort_loop, bits/stl_algo.h:
//   `if (__depth_limit == 0) { std::__partial_sort(__first, __last, __last, __comp); return; }`),
// restated so that EQUAL weights leave the heap in the same order as in the reference's sort_indexes
// (src/aggregats/aggregat_list.cpp:109-123): __make_heap + __sort_heap built on __adjust_heap / __push_heap
// (bits/stl_heap.h).  Host/device neutral: tests/native/heap_sort_host.cpp runs it against std::partial_sort on the CPU.
#pragma once
#ifdef __CUDACC__
#define HS_HD __host__ __device__
#else
#define HS_HD
#endif

namespace heapsort {
struct HeapView {
    double *wk;
    int *perm;
    int stable;
    HS_HD bool less_at(int i, int j) const {
        return wk[i] < wk[j] || (stable && wk[i] == wk[j] && perm[i] < perm[j]);
    }
    HS_HD bool less_val(int i, double kv, int lv) const {  // element i < value
        return wk[i] < kv || (stable && wk[i] == kv && perm[i] < lv);
    }
};
HS_HD inline void heap_adjust(const HeapView &v, int first, int hole, int len, double kv, int lv) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (v.less_at(first + child, first + child - 1)) child--;
        v.wk[first + hole] = v.wk[first + child]; v.perm[first + hole] = v.perm[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        v.wk[first + hole] = v.wk[first + child - 1]; v.perm[first + hole] = v.perm[first + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;  // __push_heap
    while (hole > top && v.less_val(first + parent, kv, lv)) {
        v.wk[first + hole] = v.wk[first + parent]; v.perm[first + hole] = v.perm[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    v.wk[first + hole] = kv; v.perm[first + hole] = lv;
}
HS_HD inline void heap_sort_segment(const HeapView &v, int first, int last) {  // std::__partial_sort(first, last, last)
    const int len = last - first;
    if (len < 2) return;
    for (int parent = (len - 2) / 2;; parent--) {  // __make_heap
        heap_adjust(v, first, parent, len, v.wk[first + parent], v.perm[first + parent]);
        if (parent == 0) break;
    }
    for (int end = last; end - first > 1;) {  // __sort_heap: __pop_heap(first, end - 1, end - 1)
        --end;
        const double kv = v.wk[end];
        const int lv = v.perm[end];
        v.wk[end] = v.wk[first]; v.perm[end] = v.perm[first];
        heap_adjust(v, first, 0, end - first, kv, lv);
    }
}
}  // namespace heapsort
