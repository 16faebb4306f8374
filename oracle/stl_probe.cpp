// TEST INFRASTRUCTURE ONLY.  Live probe of the toolchain's std::unordered_map<size_t,int>:
// used by tests/test_oracle.py to pin oracle_geom.c's orc_stl_order() (and the bucket schedule
// table embedded in the CUDA product) against the real libstdc++ of this image.
#include <cstddef>
#include <cstdint>
#include <unordered_map>

extern "C" {

// Iteration order (as insertion ranks) of an unordered_map filled with keys[0..n) (distinct).
void probe_unordered_order(const uint64_t* keys, int n, int* order) {
    std::unordered_map<size_t, int> m;
    for (int i = 0; i < n; i++)
        if (m.count((size_t)keys[i]) < 1) m.emplace((size_t)keys[i], i);
    int k = 0;
    for (auto& v : m) order[k++] = v.second;
}

// Same for unordered_map<int,int> (label histograms, grid_subsampling.h:19).
void probe_unordered_order_int(const int* keys, int n, int* order) {
    std::unordered_map<int, int> m;
    for (int i = 0; i < n; i++)
        if (m.count(keys[i]) < 1) m.emplace(keys[i], i);
    int k = 0;
    for (auto& v : m) order[k++] = v.second;
}

// Bucket-count schedule: bucket_count() after each growth while inserting up to max_elems keys.
int probe_bucket_schedule(uint64_t max_elems, uint64_t* sched, int cap) {
    std::unordered_map<size_t, char> m;
    int k = 0;
    size_t last = m.bucket_count();
    for (uint64_t i = 0; i < max_elems && k < cap; i++) {
        m.emplace((size_t)i, 0);
        if (m.bucket_count() != last) {
            last = m.bucket_count();
            sched[k++] = last;
        }
    }
    return k;
}

// Pure policy walk (no allocation) to extend the schedule to huge sizes.
int probe_policy_schedule(uint64_t* sched, int cap) {
    std::__detail::_Prime_rehash_policy pol;
    size_t n_bkt = 1, n_elt = 0;
    int k = 0;
    while (k < cap) {
        auto r = pol._M_need_rehash(n_bkt, n_elt, 1);
        if (r.first) {
            n_bkt = r.second;
            sched[k++] = n_bkt;
            if (n_bkt > (1ull << 31)) break;
        }
        n_elt = n_bkt;  // jump to the next growth point: element number n_bkt+1 triggers it
    }
    return k;
}
}
