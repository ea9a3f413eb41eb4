// CPU model check of flye_b200/csrc/glibc_logf.cuh: the host build of the device's logf against the container's logf (glibc —
// what the reference's std::log(float) calls, overlap.cpp:423), bit for bit.
//   logf_check [stride]   compares every stride-th positive normal float (stride 1 = all 2,130,706,432 of them, ~40 s) plus every
//                         float of [1, 4) and, on random overlap records, kmerDivergence against the reference's expression.
// Prints "OK n=<floats compared> records=<...>" or the first mismatches.
#include "../../flye_b200/csrc/glibc_logf.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

static bool same(float a, float b) { return std::memcmp(&a, &b, 4) == 0; }

int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)std::strtoul(argv[1], nullptr, 10) : 97;
    unsigned long long n = 0, bad = 0;
    auto check = [&](uint32_t ix) {
        float x; std::memcpy(&x, &ix, 4);
        volatile float xv = x;
        const float ref = std::log((float)xv), mine = fg::glibcLogf(x);
        ++n;
        if (!fg::glibcLogfInDomain(x) || !same(ref, mine)) { if (bad++ < 5) std::printf("logf(%a): glibc %a, device model %a\n", x, ref, mine); }
    };
    for (uint64_t ix = 0x00800000u; ix < 0x7f800000u; ix += stride) check((uint32_t)ix);
    for (uint32_t ix = 0x3f800000u; ix < 0x40800000u; ++ix) check(ix);   // [1, 4): where 1 / matchRate of a good overlap lives
    const float outside[] = {0.0f, -1.0f, INFINITY, NAN, 1e-40f};
    for (float x : outside) if (fg::glibcLogfInDomain(x)) { ++bad; std::printf("%a should be outside the domain\n", x); }

    // whole expression on random records (overlap.cpp:417-423)
    std::mt19937 rng(12345);
    unsigned long long records = 0, degenerate = 0;
    for (int it = 0; it < 4000000; ++it) {
        const int32_t curRange = 1 + (int32_t)(rng() % 60000), extRange = 1 + (int32_t)(rng() % 60000);
        const int32_t filtered = (it % 7 == 0) ? (int32_t)(rng() % (uint32_t)(std::max(curRange, extRange) + 2)) : (int32_t)(rng() % 50);
        const int32_t chain = (it % 11 == 0) ? 0 : (int32_t)(rng() % 9000);
        const float sampleRates[] = {1.0f, 5.5f, 6.0132f, 0.999f};
        const float sampleRate = sampleRates[rng() % 4];
        const int k = (rng() & 1) ? 15 : 17;
        // the reference's statement sequence
        volatile float normLen = std::max(curRange, extRange) - filtered;
        volatile float matchRate = (float)chain * sampleRate / normLen;
        matchRate = std::min((float)matchRate, 1.0f);
        volatile float inv = 1 / matchRate;
        volatile float ref = std::log((float)inv) / k;
        bool ok = false;
        const float mine = fg::kmerDivergence(curRange, extRange, filtered, chain, sampleRate, k, ok);
        ++records;
        if (!ok) { ++degenerate; continue; }   // zero chain, zero / negative normLen, ...: the device reports these, the host recomputes
        if (!same((float)ref, mine)) { if (bad++ < 5) std::printf("record %d: reference %a, device model %a\n", it, (float)ref, mine); }
    }
    if (bad) { std::printf("FAILED: %llu mismatches\n", bad); return 1; }
    std::printf("OK n=%llu records=%llu degenerate=%llu\n", n, records, degenerate);
    return 0;
}
