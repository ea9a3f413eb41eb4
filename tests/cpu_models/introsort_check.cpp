// Host build of flye_b200/csrc/introsort_warp.cuh (32 lanes run as loops) checked against std::sort.
// Usage: introsort_check [n_trials] [seed]   -> prints "OK <arrays> <elements>" or the first mismatch.
#define FG_WARP_HOST 1
#include "../../flye_b200/csrc/introsort_warp.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>

struct VecSink {
    struct T { long f, l; int d; };
    std::vector<T> tasks;
    void operator()(long f, long l, int d) { if (d >= 0) tasks.push_back({f, l, d}); }   // d < 0: finished in place (heap sort)
};
static long g_small = 0;   // > 0: two-level mode (top pass hands ranges <= g_small to a second pass)
static int g_table = 0;     // second pass / one-pass sort uses the rank-table partition of the shared-memory kernel

static bool checkOne(std::vector<fg::Elem> a, const char* what) {
    std::vector<fg::Elem> ref = a;
    std::sort(ref.begin(), ref.end(), [](const fg::Elem& x, const fg::Elem& y) { return x.key < y.key; });
    unsigned char tab[128];
    if (g_small > 0 && a.size() > 1) {
        VecSink sink;
        fg::warpIntrosortRange(a.data(), 0, (int)a.size(), fg::introsortDepth((long)a.size()), (int)g_small, sink, tab);
        // second pass on a staged copy of each task, in reverse order to show that the order is irrelevant
        for (size_t t = sink.tasks.size(); t-- > 0;) {
            auto& T = sink.tasks[t];
            std::vector<fg::Elem> tmp(a.begin() + T.f, a.begin() + T.l);
            fg::NoSink none;
            if (g_table) {
                std::vector<unsigned short> t16(2 * tmp.size() + 2);
                std::vector<uint32_t> bits(tmp.size() / 32 + 3);
                if (g_table == 1) fg::warpIntrosortRange<true>(tmp.data(), 0, (int)tmp.size(), T.d, 0, none, (unsigned char*)t16.data(), false, (int)tmp.size());
                else if (g_table == 2) fg::warpIntrosortSmem<16>(tmp.data(), (int)tmp.size(), T.d, t16.data(), t16.data() + tmp.size(), bits.data());
                else fg::warpIntrosortSmem<32>(tmp.data(), (int)tmp.size(), T.d, t16.data(), t16.data() + tmp.size(), bits.data());
            } else
                fg::warpIntrosortRange(tmp.data(), 0, (int)tmp.size(), T.d, 0, none, tab);
            std::copy(tmp.begin(), tmp.end(), a.begin() + T.f);
        }
    } else if (g_small < 0)
        fg::seqIntrosort(a.data(), (long)a.size());
    else if (g_table && a.size() < 65536) {
        fg::NoSink none;
        std::vector<unsigned short> t16(2 * a.size() + 2);
        std::vector<uint32_t> bits(a.size() / 32 + 3);
        const int d0 = fg::introsortDepth((long)a.size());
        if (g_table == 1) fg::warpIntrosortRange<true>(a.data(), 0, (int)a.size(), d0, 0, none, (unsigned char*)t16.data(), false, (int)a.size());
        else if (g_table == 2) fg::warpIntrosortSmem<16>(a.data(), (int)a.size(), d0, t16.data(), t16.data() + a.size(), bits.data());
        else fg::warpIntrosortSmem<32>(a.data(), (int)a.size(), d0, t16.data(), t16.data() + a.size(), bits.data());
    } else
        fg::warpIntrosort(a.data(), (int)a.size(), tab);
    for (size_t i = 0; i < a.size(); ++i)
        if (a[i].key != ref[i].key || a[i].val != ref[i].val) {
            printf("MISMATCH %s n=%zu at %zu: got (%llu,%u) want (%llu,%u)\n", what, a.size(), i, a[i].key, a[i].val,
                   ref[i].key, ref[i].val);
            return false;
        }
    return true;
}

// Model of the tie-following hit sort (overlap.cu: rangeIsTieFree): the recursion is only followed into ranges whose
// sorted ranks contain a duplicated key (inside or across an end); every other range is taken from the STABLE-sorted array.
// Must equal std::sort of the whole array.
static long g_followed = 0, g_skipped = 0;
static void tieFollow(std::vector<fg::Elem>& e, const std::vector<fg::Elem>& stab, const std::vector<unsigned>& tieP, long f, long l, int d) {
    if (l - f < 2) { if (l - f == 1 && tieP[l] == tieP[f ? f - 1 : 0]) e[f] = stab[f]; return; }
    if (tieP[l] == tieP[f ? f - 1 : 0]) { ++g_skipped; std::copy(stab.begin() + f, stab.begin() + l, e.begin() + f); return; }
    ++g_followed;
    if (l - f <= 16) { fg::seqInsertionSort(e.data(), f, l); return; }
    if (d == 0) { fg::seqHeapSort(e.data() + f, l - f); return; }
    --d;
    fg::seqMedianToFirst(e.data(), f, f + 1, f + (l - f) / 2, l - 1);
    const long cut = fg::seqUnguardedPartition(e.data(), f + 1, l, f);
    tieFollow(e, stab, tieP, cut, l, d);
    tieFollow(e, stab, tieP, f, cut, d);
}
static bool checkTieFollow(std::vector<fg::Elem> a, const char* what) {
    std::vector<fg::Elem> ref = a, stab = a;
    std::sort(ref.begin(), ref.end(), [](const fg::Elem& x, const fg::Elem& y) { return x.key < y.key; });
    std::stable_sort(stab.begin(), stab.end(), [](const fg::Elem& x, const fg::Elem& y) { return x.key < y.key; });
    const long n = (long)a.size();
    std::vector<unsigned> tieP(n + 1, 0);
    for (long i = 0; i < n; ++i) tieP[i + 1] = tieP[i] + ((i + 1 < n && stab[i].key == stab[i + 1].key) ? 1u : 0u);
    tieFollow(a, stab, tieP, 0, n, fg::introsortDepth(n));
    for (long i = 0; i < n; ++i)
        if (a[i].key != ref[i].key || a[i].val != ref[i].val) {
            printf("MISMATCH tie-follow %s n=%ld at %ld: got (%llu,%u) want (%llu,%u)\n", what, n, i, a[i].key, a[i].val, ref[i].key, ref[i].val);
            return false;
        }
    return true;
}

// Model of sortHugeKernel's partition step (overlap.cu): stop lists per slice, s = #{m : L_m < R_m} by a monotone search,
// s parallel swaps, cut = min(L_s, R_{s-1}) — against the literal __move_median_to_first + __unguarded_partition.
static bool checkHugePartition(std::vector<fg::Elem> a, const char* what) {
    const long n = (long)a.size();
    if (n <= 16) return true;
    std::vector<fg::Elem> ref = a;
    fg::seqMedianToFirst(ref.data(), 0, 1, n / 2, n - 1);
    const long cutRef = fg::seqUnguardedPartition(ref.data(), 1, n, 0);
    fg::seqMedianToFirst(a.data(), 0, 1, n / 2, n - 1);
    const unsigned long long p = a[0].key;
    const int W = 16;
    const long sliceLen = (((n - 1 + W - 1) / W) + 31) & ~31L;
    std::vector<long> L, R;   // concatenation of the slices' lists = all stops, ascending positions
    for (int w = 0; w < W; ++w)
        for (long i = std::min(n, 1 + w * sliceLen); i < std::min(n, 1 + (w + 1) * sliceLen); ++i) {
            if (a[i].key >= p) L.push_back(i);
            if (a[i].key <= p) R.push_back(i);
        }
    auto Lm = [&](long m) { return L[m]; };
    auto Rm = [&](long m) { return R[(long)R.size() - 1 - m]; };
    long lo = 0, hi = (long)std::min(L.size(), R.size());   // all m < lo hold, no m >= hi holds
    while (lo < hi) {   // the kernel probes 512 values of m per round; any search over a monotone predicate gives the same s
        const long mid = (lo + hi) / 2;
        if (Lm(mid) < Rm(mid)) lo = mid + 1; else hi = mid;
    }
    const long s = lo;
    for (long m = 0; m < s; ++m) std::swap(a[Lm(m)], a[Rm(m)]);
    const long Ls = s < (long)L.size() ? Lm(s) : n;
    const long cut = s == 0 ? Ls : std::min(Ls, Rm(s - 1));
    if (cut != cutRef) { printf("MISMATCH huge partition cut %s n=%ld: %ld vs %ld\n", what, n, cut, cutRef); return false; }
    for (long i = 0; i < n; ++i)
        if (a[i].key != ref[i].key || a[i].val != ref[i].val) { printf("MISMATCH huge partition %s n=%ld at %ld\n", what, n, i); return false; }
    return true;
}

// median-of-3 killer (Musser) to drive introsort into its heap-sort fallback
static std::vector<unsigned long long> killer(size_t n) {
    std::vector<unsigned long long> v(n);
    size_t k = n / 2;
    for (size_t i = 0; i < k; ++i) { v[i] = (i % 2 == 0) ? i + 1 : k + i + (k % 2 == 0 ? 0 : 1); }
    for (size_t i = 0; i < k; ++i) { if (i % 2 == 0) v[i] = i + 1; else v[i] = k + i + 1; v[k + i] = 2 * (i + 1); }
    return v;
}

int main(int argc, char** argv) {
    int trials = argc > 1 ? atoi(argv[1]) : 3000;
    unsigned seed = argc > 2 ? atoi(argv[2]) : 12345;
    g_small = argc > 3 ? atol(argv[3]) : 0;
    g_table = argc > 4 ? atoi(argv[4]) : 0;   // 1: table partition, 2/3: warpIntrosortSmem<16/32>
    std::mt19937_64 rng(seed);
    size_t arrays = 0, elements = 0;
    auto mk = [&](size_t n, int mode, unsigned long long distinct) {
        std::vector<fg::Elem> a(n);
        for (size_t i = 0; i < n; ++i) {
            unsigned long long k;
            switch (mode) {
                case 0: k = rng() % distinct; break;                       // random with ties
                case 1: k = i / (1 + rng() % 3); break;                    // presorted with runs
                case 2: k = (n - i) / 2; break;                            // reversed with pairs
                case 3: k = 7; break;                                      // all equal
                case 4: k = i < n / 2 ? i : n - i; break;                  // organ pipe
                case 5: k = (i * 2654435761ULL) % distinct; break;         // structured
                default: k = rng(); break;                                 // distinct
            }
            a[i] = {k, (unsigned)i, 0u};
        }
        return a;
    };
    for (int t = 0; t < trials; ++t) {
        size_t n;
        int r = t % 10;
        if (r < 4) n = 1 + rng() % 200; else if (r < 8) n = 1 + rng() % 5000; else n = 1 + rng() % 70000;
        int mode = rng() % 7;
        unsigned long long distinct = 1 + rng() % (r % 2 ? 50 : (n + 1));
        auto a = mk(n, mode, distinct);
        ++arrays; elements += n;
        if (!checkOne(a, "random")) return 1;
        if (g_small == 0 && g_table == 0) {
            if (!checkTieFollow(a, "random")) return 1;
            if (!checkHugePartition(a, "random")) return 1;
            // mostly distinct keys with a few duplicated ones: the case the fast path is made for
            auto b = mk(n, 6, 1);
            for (int dup = 0; dup < 3 && n > 4; ++dup) { const size_t x = rng() % n, y = rng() % n; b[x].key = b[y].key; if (rng() % 2) b[rng() % n].key = b[y].key; }
            if (!checkTieFollow(b, "few-ties")) return 1;
        }
    }
    for (size_t n : {17u, 18u, 31u, 32u, 33u, 34u, 48u, 49u, 50u, 63u, 64u, 65u, 66u, 96u, 97u, 128u, 129u, 1000u, 4096u, 100000u}) {
        for (int mode = 0; mode < 7; ++mode) { auto a = mk(n, mode, 5); ++arrays; elements += n; if (!checkOne(a, "edge")) return 1; }
        auto kv = killer(n);
        std::vector<fg::Elem> a(n);
        for (size_t i = 0; i < n; ++i) a[i] = {kv[i], (unsigned)i, 0u};
        ++arrays; elements += n;
        if (!checkOne(a, "killer")) return 1;
        if (g_small == 0 && g_table == 0) { if (!checkHugePartition(a, "killer")) return 1; a[n / 3].key = a[n / 2].key; if (!checkTieFollow(a, "killer+tie")) return 1; }
    }
    // exhaustive small alphabets around the chunk sizes: all arrays over {0,1,2} of length 17..20 is 3^20 -> sample
    for (int t = 0; t < trials * 5; ++t) {
        size_t n = 17 + rng() % 60;
        std::vector<fg::Elem> a(n);
        for (size_t i = 0; i < n; ++i) a[i] = {(unsigned long long)(rng() % 3), (unsigned)i, 0u};
        ++arrays; elements += n;
        if (!checkOne(a, "tiny-alphabet")) return 1;
    }
    printf("OK %zu %zu heapsorts=%ld tie-follow ranges followed=%ld skipped=%ld\n", arrays, elements, fg::g_heapSortCalls, g_followed, g_skipped);
    return 0;
}
