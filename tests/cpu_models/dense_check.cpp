// CPU check of the dense k-mer class index used by the counting kernels (flye_b200/csrc/kmer_math.cuh):
//   * a k-mer and its reverse complement get the same index, below the size of the counter array;
//   * the index maps back to the canonical k-mer (Kmer::standardForm: min of the two);
//   * different classes get different indices (checked exhaustively for small k, by sampling for k = 15..17);
//   * owner / slot split of the multi-GPU layout (owner = index % N, slot = index / N) is a bijection and slots fit 32 bits.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <map>
#include <vector>
#include "../../flye_b200/csrc/kmer_math.cuh"
using namespace fg;

static uint64_t rng_state = 88172645463325252ULL;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

static int fail(const char* what, int k, uint64_t v) { printf("FAIL %s k=%d v=%llx\n", what, k, (unsigned long long)v); return 1; }

int main() {
    for (int k = 1; k <= 17; ++k) {
        const uint64_t mask = kmerMask(k), space = denseSpaceOf(k);
        const bool exhaustive = k <= 9;
        const uint64_t n = exhaustive ? (1ULL << (2 * k)) : 150000;
        std::vector<uint8_t> seen(exhaustive ? space : 0, 0);
        std::map<uint64_t, uint64_t> sample;
        uint64_t classes = 0;
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t v = exhaustive ? i : (rnd() & mask);            // window: base p in the lowest bits
            const uint64_t f = fwdFromWindow(v, k), r = (~v) & mask;       // forward k-mer and its reverse complement (Kmer repr.)
            if (revCompKmer(f, k) != r) return fail("revCompKmer", k, v);
            const uint64_t canon = f < r ? f : r;
            const uint64_t idx = denseIndexFromWindow(v, k);
            if (idx >= space) return fail("index out of range", k, v);
            // the window of the reverse complement strand: rc in window form = group reversal of r
            const uint64_t vrc = rev2(r) >> (64 - 2 * k);
            if (denseIndexFromWindow(vrc, k) != idx) return fail("strand symmetry", k, v);
            if (denseIndexOfPair(r, f, k) != idx) return fail("pair symmetry", k, v);
            if (canonFromDenseIndex(idx, k) != canon) return fail("inverse", k, v);
            if (exhaustive) { if (!seen[idx]) { seen[idx] = 1; ++classes; } }
            else { auto it = sample.find(idx); if (it == sample.end()) sample[idx] = canon; else if (it->second != canon) return fail("collision", k, v); }
            for (uint32_t N : {2u, 3u, 4u, 8u}) {
                const uint64_t owner = idx % N, slot = idx / N;
                if (slot * N + owner != idx || slot >= (1ULL << 32)) return fail("owner/slot", k, v);
            }
        }
        if (exhaustive) {
            // number of classes: (4^k + palindromes) / 2, palindromes (k-mer == its reverse complement) exist for even k only: 4^(k/2)
            const uint64_t all = 1ULL << (2 * k), pal = (k & 1) ? 0 : (1ULL << k);
            if (classes != (all + pal) / 2) { printf("FAIL class count k=%d: %llu\n", k, (unsigned long long)classes); return 1; }
        }
    }
    printf("ok\n");
    return 0;
}
