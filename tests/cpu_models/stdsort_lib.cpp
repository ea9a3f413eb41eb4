// Test helper: libstdc++ std::sort on (key,val) pairs compared by key only — the permutation the device
// introsort must reproduce.  Built as a tiny shared library for ctypes.
#include <algorithm>
#include <cstdint>
#include <vector>
struct KV { uint64_t key; uint32_t val; };
extern "C" void std_sort_segments(uint64_t* keys, uint32_t* vals, const uint64_t* segOff, uint32_t nSegs) {
    for (uint32_t s = 0; s < nSegs; ++s) {
        std::vector<KV> a(segOff[s + 1] - segOff[s]);
        for (size_t i = 0; i < a.size(); ++i) a[i] = {keys[segOff[s] + i], vals[segOff[s] + i]};
        std::sort(a.begin(), a.end(), [](const KV& x, const KV& y) { return x.key < y.key; });
        for (size_t i = 0; i < a.size(); ++i) { keys[segOff[s] + i] = a[i].key; vals[segOff[s] + i] = a[i].val; }
    }
}
