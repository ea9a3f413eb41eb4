// flye_b200 — glibc's single-precision logf, bit for bit, for device code.
//
// OverlapDetector computes seqDivergence = std::log(1 / matchRate) / kmerSize in float (overlap.cpp:417-423): the bits of the
// result are part of the parity contract, and std::log(float) is glibc's logf.  glibc >= 2.27 evaluates logf as
//     x = 2^k * z,  z in [OFF, 2 OFF),  OFF = 0x3f330000;   16 subintervals with a table of (1/c, log c);
//     r = z/c - 1;  log x = k ln2 + log c + r + r^2 (A0 r^2 + A1 r + A2)          all in double, one final rounding to float
// (sysdeps/ieee754/flt-32/e_logf.c; the table is __logf_data).  Every operation below is an IEEE-754 double operation with
// round-to-nearest, which the device performs exactly as the host does, so the function returns glibc's bits — not merely a
// result within its error bound.  tests/cpu_models/logf_check.cpp compares the host build of this header with the container's
// logf over the positive normal floats (every one of the 2.13e9 when asked to; the same result with and without contraction of
// the multiply-adds, so glibc's FMA and non-FMA variants are both covered), and fg_ctx runs the DEVICE build against the host's
// logf on a sample before the device epilogue is switched on (overlap.cu: deviceEpilogueUsable).
#pragma once
#include <cstdint>
#include <cstring>
#ifndef __CUDACC__
#include <cmath>
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace fg {

// true for the arguments glibcLogf handles: positive, finite, normal (everything seqDivergence ever sees except degenerate
// records, which take the host path)
__host__ __device__ inline bool glibcLogfInDomain(float x) {
    uint32_t ix;
#ifdef __CUDA_ARCH__
    ix = __float_as_uint(x);
#else
    std::memcpy(&ix, &x, 4);
#endif
    return ix - 0x00800000u < 0x7f800000u - 0x00800000u;
}

__host__ __device__ inline float glibcLogf(float x) {
    // {1/c, log c} for the 16 subintervals of [OFF, 2 OFF)
    const double T[16][2] = {
        {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
        {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
        {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
        {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
        {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
        {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
        {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
        {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
    const double Ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    uint32_t ix;
#ifdef __CUDA_ARCH__
    ix = __float_as_uint(x);
#else
    std::memcpy(&ix, &x, 4);
#endif
    if (ix == 0x3f800000u) return 0.0f;   // log(1) is exactly +0
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;   // arithmetic shift
    const uint32_t iz = ix - (tmp & (0x1ffu << 23));
    float zf;
#ifdef __CUDA_ARCH__
    zf = __uint_as_float(iz);
#else
    std::memcpy(&zf, &iz, 4);
#endif
    const double z = (double)zf;
#ifdef __CUDA_ARCH__
    const double r = __fma_rn(z, T[i][0], -1.0);
    const double y0 = __fma_rn((double)k, Ln2, T[i][1]);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(A1, r, A2);
    y = __fma_rn(A0, r2, y);
    y = __fma_rn(y, r2, __dadd_rn(y0, r));
    return __double2float_rn(y);
#else
    const double r = std::fma(z, T[i][0], -1.0);
    const double y0 = std::fma((double)k, Ln2, T[i][1]);
    const double r2 = r * r;
    double y = std::fma(A1, r, A2);
    y = std::fma(A0, r2, y);
    y = std::fma(y, r2, y0 + r);
    return (float)y;
#endif
}

// seqDivergence of one overlap from the k-mer chain (overlap.cpp:409-423), the reference's float expression operation by
// operation.  `ok` = false when an intermediate leaves glibcLogf's domain (the caller then takes the host path for the batch).
__host__ __device__ inline float kmerDivergence(int32_t curRange, int32_t extRange, int32_t filtered, int32_t chainLength, float sampleRate,
                                                int k, bool& ok) {
#ifdef __CUDA_ARCH__
    const float normLen = __int2float_rn(max(curRange, extRange) - filtered);
    float matchRate = __fdiv_rn(__fmul_rn(__int2float_rn(chainLength), sampleRate), normLen);
    matchRate = (1.0f < matchRate) ? 1.0f : matchRate;   // std::min(matchRate, 1.0f): (b < a) ? b : a
    const float inv = __fdiv_rn(1.0f, matchRate);
    ok = glibcLogfInDomain(inv);
    return ok ? __fdiv_rn(glibcLogf(inv), __int2float_rn(k)) : 0.0f;
#else
    volatile float normLen = (float)((curRange > extRange ? curRange : extRange) - filtered);
    volatile float mr = (float)chainLength * sampleRate;
    volatile float matchRate = mr / normLen;
    matchRate = (1.0f < matchRate) ? 1.0f : (float)matchRate;
    volatile float inv = 1.0f / matchRate;
    ok = glibcLogfInDomain(inv);
    if (!ok) return 0.0f;
    volatile float lg = glibcLogf(inv);
    return lg / (float)k;
#endif
}

}  // namespace fg
