// flye_b200 — pure k-mer arithmetic shared by the kernels and by host code (compiles with a plain C++ compiler too:
// tests/cpu_models/dense_check.cpp checks the dense-index maps on the CPU).
#pragma once
#include <cstdint>
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace fg {

// ---------------------------------------------------------------------------------------------
// k-mer arithmetic.  Reads are packed like DnaSequence (reference sequence.h:54-69): base j at bits
// 2*(j%32) of word j/32.  `v` below is the 2k-bit little-endian window (base p in the lowest bits).
// The reference's Kmer keeps the FIRST base most significant (kmer.h:32-36), so
//     forward k-mer  = 2-bit-group reversal of v          (fwdFromWindow)
//     reverse compl. = ~v & mask                            (kmer.h:39-52 collapses to this)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t kmerMask(int k) { return k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1); }

__host__ __device__ inline uint64_t rev2(uint64_t x) {   // reverse the order of the 32 2-bit groups
#ifdef __CUDA_ARCH__
    x = __brevll(x);
#else
    x = ((x >> 32) | (x << 32));
    x = ((x & 0xFFFF0000FFFF0000ULL) >> 16) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    x = ((x & 0xFF00FF00FF00FF00ULL) >> 8) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x & 0xF0F0F0F0F0F0F0F0ULL) >> 4) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x & 0xCCCCCCCCCCCCCCCCULL) >> 2) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((x & 0x5555555555555555ULL) << 1);
#endif
    return ((x & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((x & 0x5555555555555555ULL) << 1);
}

__host__ __device__ inline uint64_t fwdFromWindow(uint64_t v, int k) { return rev2(v) >> (64 - 2 * k); }

// 2k-bit window starting at base `pos` of a packed read (words must be readable up to pos+k-1 >> 5, +1)
__host__ __device__ inline uint64_t windowAt(const uint64_t* words, uint32_t pos, int k) {
    uint32_t w = pos >> 5, sh = (pos & 31) * 2;
    uint64_t lo = words[w] >> sh;
    if (sh && sh + 2 * k > 64) lo |= words[w + 1] << (64 - sh);
    return lo & kmerMask(k);
}

// canonical k-mer of the forward-strand position; isRc = reverse complement strictly smaller (kmer.h:54-63)
__host__ __device__ inline uint64_t canonFromWindow(uint64_t v, int k, bool& isRc) {
    uint64_t f = fwdFromWindow(v, k), r = (~v) & kmerMask(k);
    isRc = r < f;
    return isRc ? r : f;
}

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {   // Kmer::hash, kmer.h:91-98
    uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__host__ __device__ inline uint64_t mix64(uint64_t h) {   // table hash (murmur3 finaliser)
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    return h;
}

// ---------------------------------------------------------------------------------------------
// dense index of a canonical k-mer class {f, rc(f)} (count_index.cu):
//   odd k   exactly one of the two has A or C as its MIDDLE base; that one is the representative and the high bit of its middle
//           base (always 0) is squeezed out: indices in [0, 4^k / 2)
//   even k  min(f, rc(f)): indices in [0, 4^k)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t denseIndexOfPair(uint64_t f, uint64_t r, int k) {   // f, r: a k-mer and its reverse complement
    if (k & 1) {   // base j of f sits at bits 2(k-1-j), 2(k-1-j)+1; the middle one (j = (k-1)/2) at bits k-1, k
        const uint64_t rep = ((f >> k) & 1ULL) ? r : f;
        return (rep & ((1ULL << k) - 1ULL)) | ((rep >> (k + 1)) << k);
    }
    return f < r ? f : r;
}
__host__ __device__ inline uint64_t revCompKmer(uint64_t f, int k) { return (~rev2(f << (64 - 2 * k))) & kmerMask(k); }   // kmer.h:39-52
__host__ __device__ inline uint64_t denseSpaceOf(int k) { return (k & 1) ? (1ULL << (2 * k - 1)) : (1ULL << (2 * k)); }
__host__ __device__ inline uint64_t denseIndexFromWindow(uint64_t v, int k) {
    return denseIndexOfPair(fwdFromWindow(v, k), (~v) & kmerMask(k), k);
}
// canonical k-mer (Kmer::standardForm, kmer.h:54-63) of a dense index
__host__ __device__ inline uint64_t canonFromDenseIndex(uint64_t idx, int k) {
    if (!(k & 1)) return idx;
    const uint64_t rep = (idx & ((1ULL << k) - 1ULL)) | ((idx >> k) << (k + 1));
    const uint64_t rc = revCompKmer(rep, k);
    return rep < rc ? rep : rc;
}

}  // namespace fg
