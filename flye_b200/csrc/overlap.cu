// flye_b200 — OverlapDetector::getSeqOverlaps for a batch of query sequences on the device.
//
// Replaces src/sequence/overlap.cpp:99-508 (and overlapTest :29-69).  Pipeline per call (DESIGN.md §2 "Exact shortcuts" has the
// arguments why the shortcuts below do not change a bit of the reference's output):
//   Q1 queryLookupKernel  every query k-mer (either strand) -> index table probe -> hit count, list start,
//                         repetitive bit (curFilteredPos, :178-182), self-hit bit (:189-190)
//   Q2 scans              hit offsets per slot / per query; prefix popcounts of the repetitive bitmap
//   Q3 expandKernel<2>    load-balanced expansion of the position lists into packed KmerMatch records in the reference's
//                         emission order (query position ascending, list order) (:176-196); flags the queries that contain an
//                         (extId, curPos) tie
//   Q4 the hit sort (:201-204)
//        segTileSortKernel    every query: stable 8-bit LSD radix sort by extId, one CTA per query, tile by tile through shared
//                             memory with coalesced copy-out — std::sort's result for the queries without ties; the last pass also
//                             writes the target-group start flags (segRadixSortKernel / ...ClusterKernel: earlier versions, kept
//                             as tested variants)
//        queries with ties    re-expanded into a scratch copy and sorted by the std::sort-exact introsort emulation
//                             (sortHugeKernel = CTA-wide partition of long ranges, sortLevel/TailKernel = warp per range,
//                             sortSmallKernel = shared-memory tasks), which only follows the ranges that contain ties
//   Q5 group kernels      target groups, uniqueMatches / bounding-box / overhang prefilters (:216-262)
//   Q6 per (query,target) pair: pairPrepKernel (DP axis, re-keying, run detection; pairs already sorted by extPos skip the
//                         re-sort of :269-275, the others take the exact sort); chainRunDpKernel = chaining DP over runs of
//                         matches, half-warp per pair (:277-323); chainFillKernel = scores / back pointers of all matches;
//                         exact sort of the score order (:331-334) only for pairs whose scores are not strictly increasing;
//                         chainWalkKernel<true> = chain walk run by run, overlapTest, filtered-position count (:338-427) and
//                         primary selection (:431-458), one warp per pair
//   Q7 gather, edit distance (editdist.cu, bounded by the divergence threshold) + host epilogue: seqDivergence with the
//                         reference's float expression and glibc logf (:417-423), divergence filter (:470-473), maxOverlaps
//                         cut (:218-219), parallel compaction; fg_overlaps_refilter = setDivergenceThreshold afterwards.
// FG_HIT_RADIX / FG_SEG_SORT / FG_DP_MODE / FG_SORT_HUGE / FG_DEVICE_EPILOGUE select the older, slower paths (kept as cross-checks, tests/test_gpu_parity.py).
#include "ctx.cuh"
#include "introsort_warp.cuh"
#include "glibc_logf.cuh"

#include <cub/cub.cuh>
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>

namespace fg {

static constexpr int QTILE = 2048;

struct OvParams {
    int k, maxJump, minOverlap, maxOverhang;
    bool checkOverhang, forceLocal, onlyMaxExt;
    bool sameSet;       // queries come from the indexed read set (ids comparable, self hits possible)
    float minUniqueF;   // 0.01f * minOverlap (overlap.cpp:110,235)
};

// per-slot info word: [39:0] first entry, bit 61 forward position stored reverse-complemented,
// bit 62 the list contains the query's own position, bit 63 query k-mer was reverse-complemented
static constexpr uint64_t INFO_FIRST_MASK = (1ULL << 40) - 1;
static constexpr uint64_t INFO_FWDRC = 1ULL << 61, INFO_SELF = 1ULL << 62, INFO_QRC = 1ULL << 63;

// ------------------------------------------------------------------------------------------------
// Q1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) queryLookupKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                         const uint32_t* __restrict__ len, const uint64_t* __restrict__ slotOff,
                                                         const uint32_t* __restrict__ selBits, const uint32_t* __restrict__ qIds,
                                                         const uint64_t* __restrict__ qSlotOff, const uint2* __restrict__ qTiles,
                                                         int k, bool sameSet, Table index, const uint32_t* __restrict__ idxBits,
                                                         uint32_t* __restrict__ hitCnt, uint64_t* __restrict__ slotInfo, uint32_t* __restrict__ filtBits) {
    const uint2 t = qTiles[blockIdx.x];
    const uint32_t id = qIds[t.x], r = id >> 1;
    const bool strand = id & 1;
    const uint32_t L = len[r], n = L - k;
    const uint32_t cnt = min((uint32_t)QTILE, n - t.y);
    const uint32_t cntPad = (cnt + 31u) & ~31u;
    const uint64_t* words = seq + wordOff[r];
    const uint64_t qbase = qSlotOff[t.x] + t.y;
    const uint64_t fbase = sameSet ? slotOff[r] : 0;
    const uint64_t mask = kmerMask(k);
    const int lane = threadIdx.x & 31;
    for (uint32_t i0 = (threadIdx.x >> 5) * 32; i0 < cntPad; i0 += (blockDim.x >> 5) * 32) {
        const uint32_t i = i0 + lane, p = t.y + i;
        uint32_t c = 0; uint64_t info = 0; bool rep = false;
        if (i < cnt) {
            const uint32_t q = strand ? (L - k - p) : p;   // forward-strand position covering the same bases
            const uint64_t v = windowAt(words, q, k);
            const uint64_t f = fwdFromWindow(v, k), rcv = (~v) & mask;
            const uint64_t qk = strand ? rcv : f, qrc = strand ? f : rcv;
            const bool flagQ = qrc < qk;            // Kmer::standardForm on the query k-mer (kmer.h:54-63)
            const uint64_t canon = flagQ ? qrc : qk;
            const bool fwdRc = rcv < f;
            uint64_t payload;
            // (absent k-mers — most of a noisy read's — are answered by the L2-resident presence bitmap, without a probe in HBM)
            if ((!idxBits || testBit(idxBits, denseIndexOfPair(f, rcv, k))) && tableFind(index, canon, payload)) {
                const uint64_t size = payload & IDX_SIZE_MASK;
                if (size == IDX_REPETITIVE) rep = true;
                else {
                    bool self = false;
                    if (sameSet && q < n) self = (selBits[(fbase + q) >> 5] >> (q & 31)) & 1u;
                    c = (uint32_t)size - (self ? 1u : 0u);
                    info = (payload >> IDX_SIZE_BITS) | (fwdRc ? INFO_FWDRC : 0) | (self ? INFO_SELF : 0) | (flagQ ? INFO_QRC : 0);
                }
            }
        }
        hitCnt[qbase + i] = c;
        slotInfo[qbase + i] = info;
        const uint32_t m = __ballot_sync(0xffffffffu, rep);
        if (lane == 0) filtBits[(qbase + i0) >> 5] = m;
    }
}

struct CastU64 { __host__ __device__ uint64_t operator()(uint32_t x) const { return x; } };
struct CastU8 { __host__ __device__ uint32_t operator()(uint8_t x) const { return x; } };
struct PopcU32 { __host__ __device__ uint32_t operator()(uint32_t x) const {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
} };

__global__ void gatherQueryHitOffKernel(const uint64_t* __restrict__ hitOff, const uint64_t* __restrict__ qSlotOff, uint32_t nQ,
                                        uint64_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nQ) out[i] = hitOff[qSlotOff[i]];
}

// ------------------------------------------------------------------------------------------------
// Q3: expansion.  One CTA per query tile; the tile's slot offsets are staged in shared memory and every
// output record finds its slot with a shared-memory binary search, so global writes are coalesced.
// ------------------------------------------------------------------------------------------------
// MODE 0: Elem records (key = extId << 32 | curPos, val = extPos) into `hits`; with `onlyTied` set, only the tiles of
//         queries whose flag is set are written (the re-expansion of the queries that need the exact sort).
// MODE 1: input of the library's radix sort (fallback of the tie-free fast path): keys[h] = (q - qFirst) << idBits | extId,
//         vals[h] = h, payload[h] = curPos << 32 | extPos.
// MODE 2: input of segRadixSortKernel: packed[h] = extId << 2*posBits | curPos << posBits | extPos, and the tie test: two hits
//         of one query position on the same target (neighbours in the position list) flag the query in qTie.
template <int MODE>
__global__ void __launch_bounds__(256) expandKernel(const uint32_t* __restrict__ len, const uint32_t* __restrict__ qlen, const uint2* __restrict__ entries,
                                                    const uint32_t* __restrict__ qIds, const uint64_t* __restrict__ qSlotOff,
                                                    const uint2* __restrict__ qTiles, int k, const uint64_t* __restrict__ hitOff,
                                                    const uint64_t* __restrict__ slotInfo, uint64_t hitBase, Elem* __restrict__ hits,
                                                    uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, unsigned long long* __restrict__ payload,
                                                    uint32_t qFirst, int idBits, const uint8_t* __restrict__ onlyTied, int posBits,
                                                    uint8_t* __restrict__ qTie) {
    __shared__ uint32_t rel[QTILE + 1];
    const uint2 t = qTiles[blockIdx.x];
    if (MODE == 0 && onlyTied && !onlyTied[t.x - qFirst]) return;
    const uint32_t id = qIds[t.x], r = id >> 1;
    const bool strand = id & 1;
    const uint32_t L = qlen[r], n = L - k;
    const uint32_t cnt = min((uint32_t)QTILE, n - t.y);
    const uint64_t qbase = qSlotOff[t.x] + t.y;
    const uint64_t h0 = hitOff[qbase];
    for (uint32_t i = threadIdx.x; i <= cnt; i += blockDim.x) rel[i] = (uint32_t)(hitOff[qbase + i] - h0);
    __syncthreads();
    const uint32_t T = rel[cnt];
    Elem* out = hits + (h0 - hitBase);
    for (uint32_t h = threadIdx.x; h < T; h += blockDim.x) {
        uint32_t lo = 0, hi = cnt;   // largest s with rel[s] <= h
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (rel[mid] <= h) lo = mid; else hi = mid; }
        const uint32_t s = lo, j = h - rel[s], p = t.y + s;
        const uint64_t info = slotInfo[qbase + s];
        const uint64_t first = info & INFO_FIRST_MASK;
        uint2 e = entries[first + j];
        if (info & INFO_SELF) {
            // the query's own position as it is stored in the index (vertex_index.cpp:78-85)
            const uint32_t q = strand ? (L - k - p) : p;
            const uint64_t self = (info & INFO_FWDRC) ? (((uint64_t)(2 * r + 1) << 32) | (L - q - k)) : (((uint64_t)(2 * r) << 32) | q);
            if ((((uint64_t)e.x << 32) | e.y) >= self) e = entries[first + j + 1];   // lists are sorted and duplicate free
        }
        uint32_t extId = e.x; int32_t extPos = (int32_t)e.y;
        if (info & INFO_QRC) { extPos = (int32_t)len[extId >> 1] - extPos - k; extId ^= 1u; }   // vertex_index.h:166-173
        if (MODE == 1) {
            const uint32_t g = (uint32_t)(h0 - hitBase) + h;
            keys[g] = ((t.x - qFirst) << idBits) | extId;
            vals[g] = g;
            payload[g] = ((unsigned long long)p << 32) | (uint32_t)extPos;
        } else if (MODE == 2) {
            payload[(h0 - hitBase) + h] = ((unsigned long long)extId << (2 * posBits)) | ((unsigned long long)p << posBits) | (uint32_t)extPos;
            if (h + 1 < rel[s + 1]) {   // the next hit comes from the same query position: same target = an (extId, curPos) tie
                uint2 e2 = entries[first + j + 1];
                if (info & INFO_SELF) {
                    const uint32_t q = strand ? (L - k - p) : p;
                    const uint64_t self = (info & INFO_FWDRC) ? (((uint64_t)(2 * r + 1) << 32) | (L - q - k)) : (((uint64_t)(2 * r) << 32) | q);
                    if ((((uint64_t)e2.x << 32) | e2.y) >= self) e2 = entries[first + j + 2];
                }
                if (e2.x == e.x) qTie[t.x - qFirst] = 1;
            }
        } else {
            Elem o; o.key = ((unsigned long long)extId << 32) | p; o.val = (unsigned int)extPos; o.aux = 0;
            out[h] = o;
        }
    }
}

// Tie-free fast path of the hit sort (overlap.cpp:201-204).  std::sort's result is the unique sorted sequence whenever
// no two elements compare equal, and the expansion emits a query's hits by ascending curPos — so for a query without an
// (extId, curPos) tie a STABLE sort by extId alone reproduces libstdc++'s output.  This kernel gathers the radix-sorted
// hits into Elem records and flags the ties (two adjacent elements with equal query, extId and curPos): the queries that
// contain one are expanded again and go through the exact introsort emulation.
__global__ void __launch_bounds__(256) gatherSortedHitsKernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                              const unsigned long long* __restrict__ payload, uint64_t M, int idBits,
                                                              Elem* __restrict__ hits, uint8_t* __restrict__ qTie, uint8_t* __restrict__ tieFlag) {
    const uint32_t idMask = (1u << idBits) - 1u;
    const int lane = threadIdx.x & 31;
    // whole warps walk the array together (the tie test looks at the neighbouring lane)
    const uint64_t warpStride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) - lane; i0 < M; i0 += warpStride) {
        const uint64_t i = i0 + lane;
        uint32_t kk = 0; unsigned long long pl = 0;
        if (i < M) { kk = keys[i]; pl = payload[vals[i]]; }
        uint32_t nk = __shfl_down_sync(0xffffffffu, kk, 1);
        uint32_t np = __shfl_down_sync(0xffffffffu, (uint32_t)(pl >> 32), 1);
        if (lane == 31 && i + 1 < M) { nk = keys[i + 1]; np = (uint32_t)(payload[vals[i + 1]] >> 32); }
        if (i < M) {
            Elem o; o.key = ((unsigned long long)(kk & idMask) << 32) | (pl >> 32); o.val = (unsigned int)pl; o.aux = 0;
            hits[i] = o;
            const bool tie = i + 1 < M && nk == kk && np == (uint32_t)(pl >> 32);
            tieFlag[i] = tie;   // input of the prefix count that lets the exact sort skip tie-free ranges
            if (tie) qTie[kk >> idBits] = 1;
        }
    }
}

// Segmented stable LSD radix sort of the packed hits by extId, one CTA per query (its hits are one contiguous segment,
// already in curPos order).  8-bit digits; every warp owns a contiguous chunk of the segment, so (warp, digit) counts turn
// into private running offsets and the scatter needs no synchronisation between warps: a warp moves 32 consecutive
// elements per step, ranks equal digits with match.any, and keeps the order of equal digits (stable).  The counts of the
// next pass are accumulated while scattering (the destination tells the owning warp).  The segment ping-pongs between two
// buffers (L2 resident for all but the longest reads); the last pass writes the Elem records.
static constexpr int SEG_WARPS = 16;
// DEEP = false: two-step register prefetch in the scatter passes, four loads in flight in the count pass, one pair of loads per
// iteration of the flag loop (the kernel of rounds 1-2).  DEEP = true (default, FG_SRS_DEEP=0 switches back): the same algorithm with
// more independent loads in flight per thread — eight in the count pass, a four-step prefetch ring in the scatter passes, four
// positions per iteration of the flag loop.  The kernel is bound by load latency x bytes in flight (Little's law: 48 warps x 2 x
// 256 B per SM sustain ~1.9 TB/s at the ~2 us a load takes here — exactly the DRAM read rate ncu reports), not by bytes or issue.
template <int MIN_CTAS, bool DEEP>
__global__ void __launch_bounds__(SEG_WARPS * 32, MIN_CTAS) segRadixSortKernel(unsigned long long* bufA, unsigned long long* bufB,
                                                                         const uint64_t* __restrict__ qHitOff, uint32_t qFirst, uint64_t hitBase,
                                                                         int posBits, int nPass, Elem* __restrict__ hits, uint8_t* __restrict__ groupFlags) {
    __shared__ uint32_t hist[2][SEG_WARPS][256];
    __shared__ uint32_t base[256];
    const uint64_t start = qHitOff[qFirst + blockIdx.x] - hitBase;
    const uint32_t n = (uint32_t)(qHitOff[qFirst + blockIdx.x + 1] - qHitOff[qFirst + blockIdx.x]);
    if (n == 0) return;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t C = (((n + SEG_WARPS - 1) / SEG_WARPS) + 31u) & ~31u;   // elements per warp, whole steps
    const uint32_t wBeg = min(n, w * C), wEnd = min(n, wBeg + C);
    const unsigned long long* src = bufA + start;
    unsigned long long* dst = bufB + start;
    const int idShift = 2 * posBits;
    const unsigned long long posMask = (1ULL << posBits) - 1ULL;
    uint32_t (*cur)[256] = hist[0];
    uint32_t (*nxt)[256] = hist[1];
    for (int i = threadIdx.x; i < SEG_WARPS * 256; i += blockDim.x) (&cur[0][0])[i] = 0u;
    __syncthreads();
    // counts of the first pass: order does not matter here, one shared-memory atomic per element (four independent loads in flight)
    constexpr int CU = DEEP ? 8 : 4;
    for (uint32_t i0 = wBeg; i0 < wEnd; i0 += 32 * CU) {
        unsigned long long x[CU];
#pragma unroll
        for (int u = 0; u < CU; ++u) { const uint32_t i = i0 + 32 * u + lane; x[u] = i < wEnd ? src[i] : ~0ULL; }
#pragma unroll
        for (int u = 0; u < CU; ++u)
            if (i0 + 32 * u + lane < wEnd) atomicAdd(&cur[w][(uint32_t)(x[u] >> idShift) & 255u], 1u);
    }
    __syncthreads();
    unsigned long long xn, xn2;   // register prefetch: the loads of steps t+1 and t+2 overlap step t
    for (int pass = 0; pass < nPass; ++pass) {
        const bool last = pass == nPass - 1;
        const int sh = idShift + 8 * pass;
        // (warp, digit) counts -> offsets: exclusive over the warps of a digit, then over the digits
        uint32_t tot = 0;
        if (threadIdx.x < 256) {
            for (int v = 0; v < SEG_WARPS; ++v) { const uint32_t c = cur[v][threadIdx.x]; cur[v][threadIdx.x] = tot; tot += c; }
        }
        if (threadIdx.x < 256) {   // warps 0..7: block-wide exclusive scan of the 256 digit totals
            uint32_t inc = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
            base[threadIdx.x] = inc - tot;               // exclusive inside the warp
            if (lane == 31) nxt[0][w] = inc;             // warp totals, parked in the other histogram for a moment
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            uint32_t add = 0;
            for (int v = 0; v < w; ++v) add += nxt[0][v];
            base[threadIdx.x] += add;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SEG_WARPS * 256; i += blockDim.x) (&nxt[0][0])[i] = 0u;
        __syncthreads();
        // one step: 32 consecutive elements of the warp's chunk, `x` = this lane's element
        auto step = [&](const uint32_t i0, const unsigned long long x) {
            const bool act = i0 + lane < wEnd;
            const uint32_t am = __ballot_sync(0xffffffffu, act);
            if (act) {
                const uint32_t d = (uint32_t)(x >> sh) & 255u;
                const uint32_t peers = __match_any_sync(am, d);
                const uint32_t o = base[d] + cur[w][d];
                __syncwarp(am);
                if ((peers & ((1u << lane) - 1u)) == 0u) cur[w][d] += __popc(peers);
                const uint32_t pos = o + __popc(peers & ((1u << lane) - 1u));
                if (!last) {
                    dst[pos] = x;
                    atomicAdd(&nxt[pos / C][(uint32_t)(x >> (sh + 8)) & 255u], 1u);
                } else {
                    Elem e; e.key = ((x >> idShift) << 32) | ((x >> posBits) & posMask); e.val = (unsigned int)(x & posMask); e.aux = 0;
                    hits[start + pos] = e;
                }
            }
            __syncwarp();
        };
        if (DEEP) {
            unsigned long long ring[4];   // the elements of the next four steps (slot = step number mod 4, compile-time after unrolling)
#pragma unroll
            for (int u = 0; u < 4; ++u) ring[u] = wBeg + 32 * u + lane < wEnd ? src[wBeg + 32 * u + lane] : 0ULL;
            for (uint32_t i0 = wBeg; i0 < wEnd; i0 += 128) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t s0 = i0 + 32 * u;
                    if (s0 < wEnd) {   // warp uniform
                        const unsigned long long x = ring[u];
                        if (s0 + 128 + lane < wEnd) ring[u] = src[s0 + 128 + lane];
                        step(s0, x);
                    }
                }
            }
        } else {
            xn = wBeg + lane < wEnd ? src[wBeg + lane] : 0ULL;
            xn2 = wBeg + 32 + lane < wEnd ? src[wBeg + 32 + lane] : 0ULL;
            for (uint32_t i0 = wBeg; i0 < wEnd; i0 += 32) {
                const unsigned long long x = xn;
                xn = xn2;
                if (i0 + lane + 64 < wEnd) xn2 = src[i0 + lane + 64];
                step(i0, x);
            }
        }
        __syncthreads();
        const unsigned long long* t = dst; dst = const_cast<unsigned long long*>(src); src = t;
        uint32_t (*th)[256] = cur; cur = nxt; nxt = th;
    }
    // start flags of the target groups (overlap.cpp:216-221), while the sorted segment is still in L2
    if (DEEP) {
        // every thread owns four consecutive positions per iteration: five independent key loads, one 4-byte flag store
        const Elem* h = hits + start;
        for (uint32_t i = 4 * threadIdx.x; i < n; i += 4 * blockDim.x) {
            uint32_t id[5];
            id[0] = i ? (uint32_t)(h[i - 1].key >> 32) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) id[u + 1] = i + u < n ? (uint32_t)(h[i + u].key >> 32) : 0u;
            uint32_t f = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) f |= (uint32_t)((i + u == 0) || id[u + 1] != id[u]) << (8 * u);
            if (((start + i) & 3u) == 0 && i + 4 <= n) *reinterpret_cast<uint32_t*>(groupFlags + start + i) = f;
            else for (int u = 0; u < 4 && i + u < n; ++u) groupFlags[start + i + u] = (uint8_t)(f >> (8 * u));
        }
        return;
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        groupFlags[start + i] = i == 0 || (uint32_t)(hits[start + i].key >> 32) != (uint32_t)(hits[start + i - 1].key >> 32);
}

// Tile version of the segmented sort (the default; FG_SEG_SORT=0 selects the kernel above).  The kernel above scatters straight from registers: the 32 elements of a
// warp step go to up to 32 different sectors (ncu: 15 sectors per store request), partly written sectors are evicted from the L2
// before their neighbours arrive, and DRAM sees 59 B per hit where 24 are compulsory — the run time did not move when the loads
// were pipelined deeper, it is the scattered traffic that bounds it.  Here a CTA takes its segment tile by tile (4096 elements):
// the tile is ranked stably by digit (per-warp counts with match.any, exclusive scan over the warps and the digits), reordered in
// shared memory, and copied out in runs of equal digits — consecutive threads write consecutive addresses, whole sectors.  The
// digit's position in the segment advances tile by tile, so equal digits keep their order (stable LSD passes, as above).
static constexpr int TS_WARPS = 16, TS_E = 8, TS_TILE = TS_WARPS * 32 * TS_E;
struct TsSmem {
    unsigned long long stage[TS_TILE];
    uint32_t warpCnt[TS_WARPS][256];   // in-tile counts per (warp, digit); after the scan: first rank of the warp inside the digit's run
    uint32_t digitBase[256];           // where the digit's next element goes in the segment
    uint32_t tileStart[256];           // where the digit's run starts in the staged tile
    uint32_t nxtCnt[256];              // digit counts of the next pass, gathered while copying out
    uint32_t digitStart[256];          // where the digit's elements start in the segment (digitBase before the first tile)
    uint32_t warpTot[8];
};
template <int MIN_CTAS>
__global__ void __launch_bounds__(TS_WARPS * 32, MIN_CTAS) segTileSortKernel(unsigned long long* bufA, unsigned long long* bufB,
                                                                            const uint64_t* __restrict__ qHitOff, uint32_t qFirst, uint64_t hitBase,
                                                                            int posBits, int nPass, Elem* __restrict__ hits, uint8_t* __restrict__ groupFlags) {
    extern __shared__ __align__(16) unsigned char tsRaw[];
    TsSmem& S = *reinterpret_cast<TsSmem*>(tsRaw);
    const uint64_t start = qHitOff[qFirst + blockIdx.x] - hitBase;
    const uint32_t n = (uint32_t)(qHitOff[qFirst + blockIdx.x + 1] - qHitOff[qFirst + blockIdx.x]);
    if (n == 0) return;
    const uint32_t tid = threadIdx.x;
    const int w = tid >> 5, lane = tid & 31;
    const unsigned long long* src = bufA + start;
    unsigned long long* dst = bufB + start;
    const int idShift = 2 * posBits;
    const unsigned long long posMask = (1ULL << posBits) - 1ULL;
    // exclusive prefix over the 256 values the threads 0..255 hold (the others pass 0); one barrier inside, and the caller has a
    // barrier between two calls (warpTot is reused)
    auto exclusive256 = [&](const uint32_t v) -> uint32_t {
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (tid < 256 && lane == 31) S.warpTot[w] = inc;
        __syncthreads();
        uint32_t add = 0;
        if (tid < 256) for (int v2 = 0; v2 < w; ++v2) add += S.warpTot[v2];
        return inc - v + add;
    };
    // digit counts of the first pass
    if (tid < 256) S.nxtCnt[tid] = 0u;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += 4 * TS_WARPS * 32) {
        unsigned long long x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const uint32_t i = i0 + u * TS_WARPS * 32 + tid; x[u] = i < n ? src[i] : 0ULL; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * TS_WARPS * 32 + tid < n) atomicAdd(&S.nxtCnt[(uint32_t)(x[u] >> idShift) & 255u], 1u);
    }
    __syncthreads();
    for (int pass = 0; pass < nPass; ++pass) {
        const bool last = pass == nPass - 1;
        const int sh = idShift + 8 * pass;
        {
            const uint32_t c = tid < 256 ? S.nxtCnt[tid] : 0u;
            const uint32_t ex = exclusive256(c);
            if (tid < 256) { S.digitBase[tid] = ex; S.digitStart[tid] = ex; S.nxtCnt[tid] = 0u; }
        }
        uint32_t prevRun = 0;   // thread d < 256: elements of digit d in the previous tile (advances digitBase[d])
        for (uint32_t t0 = 0; t0 < n; t0 += TS_TILE) {
            const uint32_t tileN = min((uint32_t)TS_TILE, n - t0);
            // (1) load, rank inside the warp: warp w owns the tile positions [256 w, 256 w + 256), step e its e-th 32 elements
            for (int j = lane; j < 256; j += 32) S.warpCnt[w][j] = 0u;
            __syncwarp();
            const uint32_t wb = t0 + (uint32_t)w * 32u * TS_E;
            unsigned long long x[TS_E];
            uint32_t r[TS_E];
#pragma unroll
            for (int e = 0; e < TS_E; ++e) { const uint32_t i = wb + e * 32 + lane; x[e] = i < n ? src[i] : 0ULL; }
#pragma unroll
            for (int e = 0; e < TS_E; ++e) {
                const bool valid = wb + e * 32 + lane < n;
                const uint32_t am = __ballot_sync(0xffffffffu, valid);
                r[e] = 0;
                if (valid) {
                    const uint32_t d = (uint32_t)(x[e] >> sh) & 255u;
                    const uint32_t peers = __match_any_sync(am, d);
                    const uint32_t lower = peers & ((1u << lane) - 1u);
                    r[e] = S.warpCnt[w][d] + __popc(lower);
                    __syncwarp(am);
                    if (lower == 0u) S.warpCnt[w][d] += __popc(peers);
                }
                __syncwarp();
            }
            __syncthreads();   // A: every warp's counts are in; the previous tile is copied out
            // (2) per digit: exclusive over the warps; the runs' starts in the tile; the digit's place in the segment moves on
            uint32_t run = 0;
            if (tid < 256) {
                for (int v = 0; v < TS_WARPS; ++v) { const uint32_t c = S.warpCnt[v][tid]; S.warpCnt[v][tid] = run; run += c; }
                S.digitBase[tid] += prevRun;
                prevRun = run;
            }
            const uint32_t ts = exclusive256(run);   // B inside
            if (tid < 256) S.tileStart[tid] = ts;
            __syncthreads();   // C
            // (3) reorder in shared memory
#pragma unroll
            for (int e = 0; e < TS_E; ++e) {
                if (wb + e * 32 + lane < n) {
                    const uint32_t d = (uint32_t)(x[e] >> sh) & 255u;
                    S.stage[S.tileStart[d] + S.warpCnt[w][d] + r[e]] = x[e];
                }
            }
            __syncthreads();   // D
            // (4) copy out: consecutive threads, consecutive addresses inside a run
            for (uint32_t s0 = tid; s0 < tileN; s0 += TS_WARPS * 32) {
                const unsigned long long y = S.stage[s0];
                const uint32_t d = (uint32_t)(y >> sh) & 255u;
                const uint32_t g = S.digitBase[d] + (s0 - S.tileStart[d]);
                if (!last) {
                    dst[g] = y;
                    atomicAdd(&S.nxtCnt[(uint32_t)(y >> (sh + 8)) & 255u], 1u);
                } else {
                    Elem o; o.key = ((y >> idShift) << 32) | ((y >> posBits) & posMask); o.val = (unsigned int)(y & posMask); o.aux = 0;
                    hits[start + g] = o;
                    // start flag of the target group (overlap.cpp:216-221): the predecessor in the final order is the staged
                    // neighbour inside a run; the last element of the digit's earlier tiles at a run's head (written before this
                    // tile's barriers); an element with another leading digit — another target — at the digit's very first element
                    const uint32_t id = (uint32_t)(y >> idShift);
                    bool flag = true;
                    if (s0 > S.tileStart[d]) flag = id != (uint32_t)(S.stage[s0 - 1] >> idShift);
                    else if (g != S.digitStart[d]) flag = id != (uint32_t)(hits[start + g - 1].key >> 32);
                    groupFlags[start + g] = flag;
                }
            }
            // (no barrier: the next tile's step (1) touches neither the stage nor the digit tables, and its barrier A comes before
            // anything of this tile is overwritten)
        }
        __syncthreads();   // the pass is written before it is read back / before the counts of the next pass are scanned
        const unsigned long long* t = dst; dst = const_cast<unsigned long long*>(src); src = t;
    }
}

// Cluster version of the segmented sort (FG_SEG_SORT=1; NOT the default — measured on B200, 1 lane, whole configs[0] / configs[1]
// pass: 9.3 / 16.2 ms with clusters of 1 / 2 CTAs against 8.3 / 12.9 ms for the kernel above, and 20-29 ms with clusters of 4-8:
// the sort is bound by the latency of its warp steps (load -> match.any -> shared-memory offsets -> scattered store), which
// 444 independent CTAs hide better than 148 CTAs that meet at cluster barriers, not by the bytes it moves.  Kept as a tested
// variant and as the record of that experiment.)  The kernel above keeps 444 segments in flight, each with its source, its
// ping-pong copy and its 16-byte output: 100-500 MB, so every pass goes through HBM (59 B of DRAM traffic per hit measured where
// 24 are compulsory).  Here the grid is PERSISTENT — one 1024-thread CTA per SM, grouped into thread-block clusters of 1, 2, 4 or
// 8 CTAs that take one query at a time from a counter — and the intermediate copy lives in a scratch slot that belongs to the
// cluster and is reused for every query it sorts: sources + scratch of all resident clusters stay inside the 126 MB L2 (the
// host picks the cluster size from the segment sizes), so HBM only sees the packed hits once on the way in and the Elem records
// on the way out.  A cluster splits a segment into one contiguous chunk per warp; the per-(warp, digit) counts of the CTAs are
// combined through distributed shared memory (every CTA reads its peers' digit totals), the counts of the next pass are
// accumulated while scattering with shared-memory atomics on the CTA that owns the destination chunk (remote for c > 1).
static constexpr int SRS_WARPS = 32;
struct SrsSmem {
    uint32_t hist[2][SRS_WARPS][256];
    uint32_t base[256];
    uint32_t ctaTot[256];
    uint32_t warpTot[8];
    uint32_t query;
};
__global__ void __launch_bounds__(SRS_WARPS * 32, 1) segRadixSortClusterKernel(const unsigned long long* __restrict__ packed, unsigned long long* __restrict__ scratch,
                                                                                uint64_t scratchStride, int nBufs, const uint64_t* __restrict__ qHitOff,
                                                                                uint32_t qFirst, uint32_t nq, uint64_t hitBase, int posBits, int nPass, Elem* __restrict__ hits,
                                                                                uint8_t* __restrict__ groupFlags, uint32_t* __restrict__ nextQuery) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char srsRaw[];
    SrsSmem& sm = *reinterpret_cast<SrsSmem*>(srsRaw);
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t cSize = cluster.num_blocks(), cRank = cluster.block_rank();
    const uint32_t clusterId = blockIdx.x / cSize;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t gw = cRank * SRS_WARPS + w, nW = cSize * SRS_WARPS;
    const int idShift = 2 * posBits;
    const unsigned long long posMask = (1ULL << posBits) - 1ULL;
    unsigned long long* const slot = scratch + (uint64_t)clusterId * (uint64_t)nBufs * scratchStride;   // nBufs = 2 when three passes ping-pong
    const uint32_t* remoteQuery = cluster.map_shared_rank(&sm.query, 0);
    for (;;) {
        if (cRank == 0 && threadIdx.x == 0) sm.query = atomicAdd(nextQuery, 1u);
        cluster.sync();
        const uint32_t q = *remoteQuery;
        cluster.sync();                                   // everybody has read it before rank 0 fetches the next one
        if (q >= nq) break;
        const uint64_t start = qHitOff[qFirst + q] - hitBase;
        const uint32_t n = (uint32_t)(qHitOff[qFirst + q + 1] - qHitOff[qFirst + q]);
        if (n == 0) continue;
        const uint32_t C = (((n + nW - 1) / nW) + 31u) & ~31u;   // elements per warp, whole steps
        const uint32_t wBeg = (uint32_t)min((uint64_t)n, (uint64_t)gw * C), wEnd = (uint32_t)min((uint64_t)n, (uint64_t)wBeg + C);
        const unsigned long long* src = packed + start;
        uint32_t (*cur)[256] = sm.hist[0];
        uint32_t (*nxt)[256] = sm.hist[1];
        for (int i = threadIdx.x; i < SRS_WARPS * 256; i += blockDim.x) (&cur[0][0])[i] = 0u;
        __syncthreads();
        {   // counts of the first pass
            unsigned long long xn = wBeg + lane < wEnd ? __ldcg(src + wBeg + lane) : 0ULL;
            for (uint32_t i0 = wBeg; i0 < wEnd; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool act = i < wEnd;
                const uint32_t am = __ballot_sync(0xffffffffu, act);
                const unsigned long long xc = xn;
                if (i + 32 < wEnd) xn = __ldcg(src + i + 32);
                if (act) {
                    const uint32_t d = (uint32_t)(xc >> idShift) & 255u;
                    const uint32_t peers = __match_any_sync(am, d);
                    if ((peers & ((1u << lane) - 1u)) == 0u) cur[w][d] += __popc(peers);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int pass = 0; pass < nPass; ++pass) {
            const bool last = pass == nPass - 1;
            const int sh = idShift + 8 * pass;
            unsigned long long* dst = slot + (uint64_t)(pass & (nBufs - 1)) * scratchStride;
            // (warp, digit) counts -> exclusive offsets over the warps of this CTA; the CTA's digit totals for the cluster
            if (threadIdx.x < 256) {
                uint32_t tot = 0;
                for (int v = 0; v < SRS_WARPS; ++v) { const uint32_t c = cur[v][threadIdx.x]; cur[v][threadIdx.x] = tot; tot += c; }
                sm.ctaTot[threadIdx.x] = tot;
            }
            for (int i = threadIdx.x; i < SRS_WARPS * 256; i += blockDim.x) (&nxt[0][0])[i] = 0u;
            cluster.sync();                               // totals visible, next-pass counters cleared in every CTA
            uint32_t below = 0, all = 0;
            if (threadIdx.x < 256) {
                for (uint32_t r = 0; r < cSize; ++r) {
                    const uint32_t t = r == cRank ? sm.ctaTot[threadIdx.x] : cluster.map_shared_rank(sm.ctaTot, r)[threadIdx.x];
                    all += t; if (r < cRank) below += t;
                }
                uint32_t inc = all;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
                sm.base[threadIdx.x] = inc - all + below;   // exclusive inside the warp of digits, plus the peers in front of this CTA
                if (lane == 31) sm.warpTot[w] = inc;
            }
            __syncthreads();
            if (threadIdx.x < 256) {
                uint32_t add = 0;
                for (int v = 0; v < w; ++v) add += sm.warpTot[v];
                sm.base[threadIdx.x] += add;
            }
            __syncthreads();
            unsigned long long xn = wBeg + lane < wEnd ? __ldcg(src + wBeg + lane) : 0ULL;
            for (uint32_t i0 = wBeg; i0 < wEnd; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool act = i < wEnd;
                const uint32_t am = __ballot_sync(0xffffffffu, act);
                const unsigned long long x = xn;
                if (i + 32 < wEnd) xn = __ldcg(src + i + 32);
                if (act) {
                    const uint32_t d = (uint32_t)(x >> sh) & 255u;
                    const uint32_t peers = __match_any_sync(am, d);
                    const uint32_t o = sm.base[d] + cur[w][d];
                    __syncwarp(am);
                    if ((peers & ((1u << lane) - 1u)) == 0u) cur[w][d] += __popc(peers);
                    const uint32_t pos = o + __popc(peers & ((1u << lane) - 1u));
                    if (!last) {
                        dst[pos] = x;
                        const uint32_t g2 = pos / C, r2 = g2 / SRS_WARPS, d2 = (uint32_t)(x >> (sh + 8)) & 255u;
                        uint32_t* cnt = &nxt[g2 % SRS_WARPS][d2];
                        if (r2 == cRank) atomicAdd(cnt, 1u);                       // shared-memory atomic
                        else atomicAdd(cluster.map_shared_rank(cnt, r2), 1u);      // the owner's shared memory, through the cluster
                    } else {
                        Elem e; e.key = ((x >> idShift) << 32) | ((x >> posBits) & posMask); e.val = (unsigned int)(x & posMask); e.aux = 0;
                        hits[start + pos] = e;
                    }
                }
                __syncwarp();
            }
            __threadfence();
            cluster.sync();                               // the scattered copy and the remote counts are complete
            src = dst;
            uint32_t (*th)[256] = cur; cur = nxt; nxt = th;
        }
        // start flags of the target groups (overlap.cpp:216-221), while the sorted segment is still in the L2
        const uint32_t per = (n + cSize - 1) / cSize, fBeg = min(n, cRank * per), fEnd = min(n, fBeg + per);
        const uint32_t* ids = reinterpret_cast<const uint32_t*>(hits + start);   // high half of the key = word 1 of every 4-word record
        for (uint32_t i = fBeg + threadIdx.x; i < fEnd; i += blockDim.x)
            groupFlags[start + i] = i == 0 || __ldcg(ids + 4ULL * i + 1) != __ldcg(ids + 4ULL * (i - 1) + 1);
    }
}

// tie flags (input of the prefix count of rangeIsTieFree) for the queries that contain ties; all other flags stay 0
__global__ void __launch_bounds__(256) tieFlagKernel(const Elem* __restrict__ hits, const uint64_t* __restrict__ qHitOff, uint32_t qFirst, uint64_t hitBase,
                                                     const uint8_t* __restrict__ qTie, uint8_t* __restrict__ tieFlag) {
    if (!qTie[blockIdx.x]) return;
    const uint64_t start = qHitOff[qFirst + blockIdx.x] - hitBase;
    const uint32_t n = (uint32_t)(qHitOff[qFirst + blockIdx.x + 1] - qHitOff[qFirst + blockIdx.x]);
    for (uint32_t i = threadIdx.x; i + 1 < n; i += blockDim.x) tieFlag[start + i] = hits[start + i].key == hits[start + i + 1].key;
}

// ------------------------------------------------------------------------------------------------
// Q4: segmented std::sort-exact sorting, two levels.
//   sortSeedKernel / sortLevelKernel   level-synchronous introsort recursion in global memory: per level one warp
//                    partitions one range (software-prefetched streaming Hoare partition) and emits the two halves as
//                    tasks; ranges of <= SORT_SMALL elements go to the task list of the second kernel.  All ranges of
//                    all segments of a level run concurrently, so a long read does not serialise behind one warp.
//   sortSmallKernel  persistent warps pull tasks, stage the range in shared memory (16 KB per warp), finish the
//                    introsort there and write it back — most recursion levels run at shared-memory latency and
//                    the tasks balance the load across the chip whatever the segment-length distribution is.
// ------------------------------------------------------------------------------------------------
static constexpr int SORT_SMALL_MAX = 2048;
// How the shared-memory kernel finishes the tasks of one segmented sort: the largest task (elements) and the variant
// of sortSmallKernel.  Measured on B200: the k-mer hits of a read arrive in no useful order -> rank-table partition +
// stable leaf pass by ranking (variant 3), 384-element tasks (28 resident warps per SM); the two per-pair sorts get
// nearly sorted input, where one lane per leaf is cheaper (variant 1), 512-element tasks.
struct SortCfg { int smallN, variant; };
static int envInt(const char* name, int dflt, int lo, int hi) { const char* e = getenv(name); const int x = e ? atoi(e) : dflt; return x < lo ? lo : (x > hi ? hi : x); }
static SortCfg sortCfgHits() {
    static const SortCfg c{envInt("FG_SORT_SMALL_HITS", 384, 64, SORT_SMALL_MAX), envInt("FG_SORT_VAR_HITS", 3, 0, 3)};
    return c;
}
static SortCfg sortCfgPairs() {
    static const SortCfg c{envInt("FG_SORT_SMALL", 512, 64, SORT_SMALL_MAX), envInt("FG_SORT_VAR", 1, 0, 3)};
    return c;
}
static int sortSmallN() { return std::min(sortCfgHits().smallN, sortCfgPairs().smallN); }   // sizes the task lists
struct Seg { uint32_t start, n; };
struct SortTask { uint32_t start, n; int depth; };

// counters of one segmented sort (8 words): [0] nSegs (input)  [1] nSmall  [2] next small task  [3] nBig ping  [4] nBig pong
//                                           [5] elements in the ping list  [6] elements in the pong list
// tieP (optional): tieP[i] = number of positions j < i of the stably sorted array with key[j] == key[j+1].  A range of the
// recursion holds the keys of the sorted ranks it covers; when none of them is duplicated (inside the range or across its
// ends) the elements are determined by their keys and std::sort's result for these positions is the stable-sorted array
// itself, which the output already holds — the emulation only has to follow the ranges that contain ties.
__device__ __forceinline__ bool rangeIsTieFree(const uint32_t* __restrict__ tieP, uint32_t start, uint32_t n) {
    return tieP && tieP[start + n] == tieP[start ? start - 1 : 0];
}
__device__ __forceinline__ void emitRange(uint32_t start, uint32_t n, int depth, SortTask* big, uint32_t* nBig, uint32_t capBig,
                                          SortTask* small, uint32_t* nSmall, uint32_t capSmall, uint32_t smallN, const uint32_t* __restrict__ tieP) {
    // (a single element that is part of a tie still has to be delivered: which of the equal elements ended up here is
    // exactly what the emulation decides)
    if (n < (tieP ? 1u : 2u) || rangeIsTieFree(tieP, start, n)) return;
    SortTask k; k.start = start; k.n = n; k.depth = depth;
    if (n > smallN) { const uint32_t t = atomicAdd(nBig, 1u); atomicAdd(nBig + 2, n); if (t < capBig) big[t] = k; }
    else { const uint32_t t = atomicAdd(nSmall, 1u); if (t < capSmall) small[t] = k; }
}

// every segment becomes one task: a "big" one (partitioned level by level in global memory) or a "small" one
__global__ void __launch_bounds__(256) sortSeedKernel(const Seg* __restrict__ segs, const uint32_t* __restrict__ nSegsPtr, SortTask* __restrict__ big,
                                                      uint32_t* __restrict__ nBig, uint32_t capBig, SortTask* __restrict__ small,
                                                      uint32_t* __restrict__ nSmall, uint32_t capSmall, uint32_t smallN, const uint32_t* __restrict__ tieP,
                                                      SortTask* __restrict__ huge, uint32_t* __restrict__ nHuge, uint32_t capHuge, uint32_t hugeMin) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *nSegsPtr) return;
    const Seg sg = segs[i];
    if (huge && sg.n > hugeMin && !rangeIsTieFree(tieP, sg.start, sg.n)) {
        const uint32_t q = atomicAdd(nHuge, 1u);
        if (q < capHuge) { SortTask k; k.start = sg.start; k.n = sg.n; k.depth = introsortDepth((long)sg.n); huge[q] = k; }
        return;
    }
    emitRange(sg.start, sg.n, introsortDepth((long)sg.n), big, nBig, capBig, small, nSmall, capSmall, smallN, tieP);
}

// one introsort level: every warp partitions one big range (or heap-sorts it when its depth budget is spent,
// stl_algo.h:1925-1929) and emits the two halves as tasks of the next level / of the shared-memory kernel
// `outArr`: where finished (fully sorted) ranges are delivered; equal to arr except in the tie-following mode, where arr is a
// scratch copy in the original order and outArr already holds the stable-sorted elements.
__device__ __forceinline__ void copyOutRange(const Elem* arr, Elem* outArr, uint32_t start, uint32_t n) {
    if (outArr == arr) return;
    __syncwarp();
    for (uint32_t i = laneId(); i < n; i += 32) outArr[start + i] = arr[start + i];
}

__global__ void __launch_bounds__(128) sortLevelKernel(Elem* __restrict__ arr, const SortTask* __restrict__ in, const uint32_t* __restrict__ nInPtr,
                                                       uint32_t capBig, SortTask* __restrict__ out, uint32_t* __restrict__ nOut,
                                                       SortTask* __restrict__ small, uint32_t* __restrict__ nSmall, uint32_t capSmall, uint32_t smallN,
                                                       Elem* outArr, const uint32_t* __restrict__ tieP) {
    __shared__ unsigned char tabs[4][WARP_TAB_BYTES];
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= min(*nInPtr, capBig)) return;
    const SortTask t = in[w];
    Elem* a = arr + t.start;
    if (t.depth == 0) {
        if (laneId() == 0) seqHeapSort(a, (long)t.n);
        copyOutRange(arr, outArr, t.start, t.n);
        return;
    }
    const idx_t cut = warpPartitionAny(a, 0, (idx_t)t.n, tabs[threadIdx.x >> 5], /*inGlobal*/ true);
    if (laneId() == 0) {
        emitRange(t.start, (uint32_t)cut, t.depth - 1, out, nOut, capBig, small, nSmall, capSmall, smallN, tieP);
        emitRange(t.start + (uint32_t)cut, t.n - (uint32_t)cut, t.depth - 1, out, nOut, capBig, small, nSmall, capSmall, smallN, tieP);
    }
}

// One partition step of a HUGE range by a whole CTA (16 warps).  A warp that streams a long range alone needs ~27 ns per
// element (every 32-element chunk waits for its own loads and for the pairing of its stops), and the level-synchronous
// recursion can only move on when the longest range of a level is done — with a few long reads to sort that is the whole
// cost.  Here the same partition (see introsort_warp.cuh: L-stops ascending = positions with key >= pivot, R-stops
// descending = positions with key <= pivot, swap the pairs (L_m, R_m) with L_m < R_m, cut = min(L_s, R_{s-1})) is done in
// three data-parallel phases: (1) every warp lists the L- and R-stops of its slice (positions, into scratch), (2) the
// number s of swapped pairs is found by a 512-way search over the monotone predicate L_m < R_m, (3) the s swaps are done
// by all threads.  Result and permutation are identical to the sequential loop.
static constexpr int HUGE_WARPS = 16;
struct StopLists {
    const uint32_t* Lpos; const uint32_t* Rpos; const uint32_t* geBefore; const uint32_t* leAfter; const uint32_t* leCnt;
    uint32_t sliceLen;
    // position (relative to the range) of the m-th L-stop / of the m-th R-stop counted from the right end
    __device__ __forceinline__ uint32_t L(uint32_t m) const {
        int w = 0;
#pragma unroll
        for (int v = 1; v < HUGE_WARPS; ++v) w += geBefore[v] <= m;
        return Lpos[1u + w * sliceLen + (m - geBefore[w])];
    }
    __device__ __forceinline__ uint32_t R(uint32_t m) const {
        int w = HUGE_WARPS - 1;
#pragma unroll
        for (int v = HUGE_WARPS - 2; v >= 0; --v) w -= leAfter[v] <= m;
        return Rpos[1u + w * sliceLen + (leCnt[w] - 1u - (m - leAfter[w]))];
    }
};

__global__ void __launch_bounds__(HUGE_WARPS * 32) sortHugeKernel(Elem* __restrict__ arr, const SortTask* __restrict__ in, const uint32_t* __restrict__ nInPtr,
                                                                  uint32_t capHuge, SortTask* __restrict__ outHuge, uint32_t* __restrict__ nOutHuge, uint32_t hugeMin,
                                                                  SortTask* __restrict__ big, uint32_t* __restrict__ nBig, uint32_t capBig,
                                                                  SortTask* __restrict__ small, uint32_t* __restrict__ nSmall, uint32_t capSmall, uint32_t smallN,
                                                                  Elem* outArr, const uint32_t* __restrict__ tieP, uint32_t* __restrict__ scratchL,
                                                                  uint32_t* __restrict__ scratchR) {
    __shared__ uint32_t geCnt[HUGE_WARPS], leCnt[HUGE_WARPS], geBefore[HUGE_WARPS + 1], leAfter[HUGE_WARPS + 1];
    __shared__ uint32_t sLo, sHi, sS;
    if (blockIdx.x >= min(*nInPtr, capHuge)) return;
    const SortTask t = in[blockIdx.x];
    Elem* a = arr + t.start;
    const uint32_t n = t.n;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (t.depth == 0) {   // stl_algo.h:1925-1929 (not expected for ranges of this size)
        if (threadIdx.x == 0) seqHeapSort(a, (long)n);
        __syncthreads();
        if (outArr != arr) for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) outArr[t.start + i] = a[i];
        return;
    }
    if (threadIdx.x == 0) seqMedianToFirst(a, 0, 1, (long)(n / 2), (long)n - 1);
    __syncthreads();
    const unsigned long long p = a[0].key;
    uint32_t* Lp = scratchL + t.start; uint32_t* Rp = scratchR + t.start;
    const uint32_t sliceLen = (((n - 1 + HUGE_WARPS - 1) / HUGE_WARPS) + 31u) & ~31u;
    const uint32_t b = min(n, 1u + w * sliceLen), e = min(n, b + sliceLen);
    // (1) stops of this warp's slice [b, e), ascending positions
    uint32_t cg = 0, cl = 0;
    for (uint32_t i0 = b; i0 < e; i0 += 32) {
        const uint32_t i = i0 + lane;
        const unsigned long long key = i < e ? a[i].key : 0ULL;
        const bool ge = i < e && key >= p, le = i < e && key <= p;
        const uint32_t gm = __ballot_sync(0xffffffffu, ge), lm = __ballot_sync(0xffffffffu, le);
        const uint32_t lt = (1u << lane) - 1u;
        if (ge) Lp[b + cg + __popc(gm & lt)] = i;
        if (le) Rp[b + cl + __popc(lm & lt)] = i;
        cg += __popc(gm); cl += __popc(lm);
    }
    if (lane == 0) { geCnt[w] = cg; leCnt[w] = cl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t g = 0;
        for (int v = 0; v < HUGE_WARPS; ++v) { geBefore[v] = g; g += geCnt[v]; }
        geBefore[HUGE_WARPS] = g;
        uint32_t l = 0;
        for (int v = HUGE_WARPS - 1; v >= 0; --v) { leAfter[v] = l; l += leCnt[v]; }
        leAfter[HUGE_WARPS] = l;
        sLo = 0; sHi = min(g, l);
    }
    __syncthreads();
    const uint32_t GE = geBefore[HUGE_WARPS];
    StopLists st{Lp, Rp, geBefore, leAfter, leCnt, sliceLen};
    // (2) s = number of m with L_m < R_m (a prefix of the m): all m < lo hold, no m >= hi holds
    for (;;) {
        const uint32_t lo = sLo, hi = sHi;
        if (lo >= hi) { if (threadIdx.x == 0) sS = lo; break; }
        const uint32_t step = (hi - lo + blockDim.x - 1) / blockDim.x;
        const uint32_t m = lo + threadIdx.x * step;
        const bool pred = m < hi && st.L(m) < st.R(m);
        const uint32_t c = (uint32_t)__syncthreads_count(pred);
        if (threadIdx.x == 0) {
            if (step == 1) { sLo = lo + c; sHi = lo + c; }
            else if (c == 0) { sHi = lo; }
            else { sLo = lo + (c - 1) * step + 1; sHi = min(hi, lo + c * step); }
        }
        __syncthreads();
    }
    __syncthreads();
    const uint32_t s = sS;
    // (3) the swaps (the two position sets are disjoint: every L_m with m < s lies below every R_m' with m' < s)
    for (uint32_t m = threadIdx.x; m < s; m += blockDim.x) {
        const uint32_t x = st.L(m), y = st.R(m);
        const Elem ex = a[x], ey = a[y];
        a[x] = ey; a[y] = ex;
    }
    if (threadIdx.x == 0) {
        const uint32_t Ls = s < GE ? st.L(s) : n;
        const uint32_t cut = s == 0 ? Ls : min(Ls, st.R(s - 1));
        const uint32_t cs[2] = {t.start, t.start + cut}, cn[2] = {cut, n - cut};
        for (int c = 0; c < 2; ++c) {
            if (cn[c] > hugeMin && !rangeIsTieFree(tieP, cs[c], cn[c])) {
                const uint32_t q = atomicAdd(nOutHuge, 1u);
                if (q < capHuge) { SortTask k; k.start = cs[c]; k.n = cn[c]; k.depth = t.depth - 1; outHuge[q] = k; }
            } else emitRange(cs[c], cn[c], t.depth - 1, big, nBig, capBig, small, nSmall, capSmall, smallN, tieP);
        }
    }
}

// the long tail of the recursion (few, unevenly split ranges left): one warp finishes one remaining range in global
// memory, handing every sub-range of <= smallN elements to the shared-memory kernel
struct TaskSinkDev {
    SortTask* tasks; uint32_t* counter; uint32_t cap; uint32_t base;
    const Elem* arr; Elem* outArr; const uint32_t* tieP;
    __device__ __forceinline__ void operator()(long f, long l, int d) const {
        if (d < 0) { copyOutRange(arr, outArr, base + (uint32_t)f, (uint32_t)(l - f)); return; }   // heap-sorted in place: finished
        if (l - f < (tieP ? 1 : 2)) return;
        if (laneId() == 0 && !rangeIsTieFree(tieP, base + (uint32_t)f, (uint32_t)(l - f))) {
            const uint32_t t = atomicAdd(counter, 1u);
            if (t < cap) { SortTask k; k.start = base + (uint32_t)f; k.n = (uint32_t)(l - f); k.depth = d; tasks[t] = k; }
        }
    }
};

__global__ void __launch_bounds__(128) sortTailKernel(Elem* __restrict__ arr, const SortTask* __restrict__ in, const uint32_t* __restrict__ nInPtr,
                                                      uint32_t capBig, SortTask* __restrict__ small, uint32_t* __restrict__ nSmall,
                                                      uint32_t capSmall, uint32_t smallN, Elem* outArr, const uint32_t* __restrict__ tieP) {
    __shared__ unsigned char tabs[4][WARP_TAB_BYTES];
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= min(*nInPtr, capBig)) return;
    const SortTask t = in[w];
    TaskSinkDev sink{small, nSmall, capSmall, t.start, arr, outArr, tieP};
    warpIntrosortRange(arr + t.start, 0, (idx_t)t.n, t.depth, (idx_t)smallN, sink, tabs[threadIdx.x >> 5], /*inGlobal*/ true);
}

// VARIANT 0: streaming two-ended partition + one lane per mini range (the code path of the global-memory kernels);
//         1: rank-table partition (warpPartitionTable), leaves as in 0;
//         2: warpIntrosortSmem<16>: rank-table partition down to the 16-element leaves, stable leaf pass by ranking;
//         3: warpIntrosortSmem<32>: as 2, ranges of 17..32 elements partitioned one lane per range.
// Dynamic shared memory: 4 warps x smallN elements, then (VARIANT > 0) 4 x 2 rank tables of smallN 16-bit entries, then
// (VARIANT > 1) 4 cut bitmaps of smallN / 32 + 2 words.
template <int VARIANT>
__global__ void __launch_bounds__(128) sortSmallKernel(Elem* __restrict__ arr, const SortTask* __restrict__ tasks,
                                                       const uint32_t* __restrict__ taskCounter, uint32_t taskCap, uint32_t* __restrict__ next,
                                                       uint32_t smallN, Elem* outArr) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ unsigned char tabs[4][WARP_TAB_BYTES];
    const int wid = threadIdx.x >> 5;
    Elem* sm = reinterpret_cast<Elem*>(smemRaw) + wid * smallN;
    unsigned char* tab = VARIANT ? smemRaw + 4 * smallN * sizeof(Elem) + wid * smallN * 4 : tabs[wid];
    uint32_t* bits = reinterpret_cast<uint32_t*>(smemRaw + 4 * smallN * (sizeof(Elem) + 4)) + wid * (smallN / 32 + 2);
    const int lane = threadIdx.x & 31;
    const uint32_t nTasks = min(*taskCounter, taskCap);
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(next, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= nTasks) return;
        const SortTask k = tasks[t];
        Elem* g = arr + k.start;
        for (uint32_t i = lane; i < k.n; i += 32) sm[i] = g[i];
        __syncwarp();
        NoSink none;
        if (VARIANT == 2) warpIntrosortSmem<16>(sm, (idx_t)k.n, k.depth, reinterpret_cast<unsigned short*>(tab), reinterpret_cast<unsigned short*>(tab) + smallN, bits);
        else if (VARIANT == 3) warpIntrosortSmem<32>(sm, (idx_t)k.n, k.depth, reinterpret_cast<unsigned short*>(tab), reinterpret_cast<unsigned short*>(tab) + smallN, bits);
        else warpIntrosortRange<VARIANT == 1>(sm, 0, (idx_t)k.n, k.depth, 0, none, tab, false, (int)smallN);
        __syncwarp();
        Elem* go = outArr + k.start;
        for (uint32_t i = lane; i < k.n; i += 32) go[i] = sm[i];
        __syncwarp();
    }
}

__global__ void querySegsKernel(const uint64_t* __restrict__ qHitOff, uint32_t qFirst, uint32_t nQ, uint64_t hitBase, Seg* __restrict__ segs,
                                const uint8_t* __restrict__ onlyTied, uint32_t* __restrict__ nTied) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nQ) return;
    Seg s; s.start = (uint32_t)(qHitOff[qFirst + i] - hitBase); s.n = (uint32_t)(qHitOff[qFirst + i + 1] - qHitOff[qFirst + i]);
    if (onlyTied) {
        if (!onlyTied[i]) s.n = 0;   // already in std::sort's order (segments of < 2 elements are never seeded)
        else atomicAdd(nTied, 1u);
    }
    segs[i] = s;
}

// ------------------------------------------------------------------------------------------------
// Q5: groups
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) groupFlagKernel(const Elem* __restrict__ hits, uint64_t M, uint8_t* __restrict__ flags) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < M; i += (uint64_t)gridDim.x * blockDim.x)
        flags[i] = (i == 0) || ((hits[i].key >> 32) != (hits[i - 1].key >> 32));
}
__global__ void queryStartFlagKernel(const uint64_t* __restrict__ qHitOff, uint32_t qFirst, uint32_t nQ, uint64_t hitBase,
                                     uint64_t M, uint8_t* __restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nQ) return;
    uint64_t a = qHitOff[qFirst + i] - hitBase;
    if (a < M && qHitOff[qFirst + i + 1] > qHitOff[qFirst + i]) flags[a] = 1;
}

// candidate = group with enough hits to possibly pass the uniqueMatches test
__global__ void __launch_bounds__(256) groupCandidateKernel(const uint32_t* __restrict__ gStart, uint32_t G, uint64_t M, float minUniqueF,
                                                            uint8_t* __restrict__ cand) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    uint64_t a = gStart[g], b = (g + 1 < G) ? gStart[g + 1] : M;
    cand[g] = !((float)(b - a) < minUniqueF);
}

struct PairInfo { uint32_t start, n, qi, extId; };

// one warp per candidate group: uniqueMatches, bounding box, overhang (overlap.cpp:222-262)
__global__ void __launch_bounds__(256) pairFilterKernel(const Elem* __restrict__ hits, const uint32_t* __restrict__ gStart, uint32_t G, uint64_t M,
                                                        const uint32_t* __restrict__ candIds, uint32_t C,
                                                        const uint64_t* __restrict__ qHitOff, uint32_t qFirst, uint32_t nQ, uint64_t hitBase,
                                                        const uint32_t* __restrict__ qIds, const uint32_t* __restrict__ len,
                                                        const uint32_t* __restrict__ qlen, OvParams P,
                                                        uint8_t* __restrict__ pass, PairInfo* __restrict__ info) {
    const uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    const uint32_t g = candIds[c];
    const uint32_t a = gStart[g], b = (g + 1 < G) ? gStart[g + 1] : (uint32_t)M;
    uint32_t uniq = 0; int32_t minExt = INT32_MAX, maxExt = INT32_MIN;
    for (uint32_t i = a + lane; i < b; i += 32) {
        const Elem e = hits[i];
        const uint32_t cur = (uint32_t)e.key;
        const uint32_t prev = (i == a) ? 0u : (uint32_t)hits[i - 1].key;   // prevPos starts at 0 (:223)
        uniq += (cur != prev);
        minExt = min(minExt, (int32_t)e.val); maxExt = max(maxExt, (int32_t)e.val);
    }
    uniq = __reduce_add_sync(0xffffffffu, uniq);
    minExt = __reduce_min_sync(0xffffffffu, minExt);
    maxExt = __reduce_max_sync(0xffffffffu, maxExt);
    if (lane) return;
    // query of this group: last q with qHitOff[q] - hitBase <= a
    uint32_t lo = 0, hi = nQ;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (qHitOff[qFirst + mid] - hitBase <= a) lo = mid; else hi = mid; }
    const uint32_t qi = qFirst + lo;
    const uint32_t extId = (uint32_t)(hits[a].key >> 32);
    const int32_t curLen = (int32_t)qlen[qIds[qi] >> 1], extLen = (int32_t)len[extId >> 1];
    const int32_t minCur = (int32_t)(uint32_t)hits[a].key, maxCur = (int32_t)(uint32_t)hits[b - 1].key;
    bool ok = !((float)uniq < P.minUniqueF);
    if (maxCur - minCur < P.minOverlap || maxExt - minExt < P.minOverlap) ok = false;
    if (P.checkOverhang && !P.forceLocal) {
        if (min(minCur, minExt) > P.maxOverhang) ok = false;
        if (min(curLen - maxCur, extLen - maxExt) > P.maxOverhang) ok = false;
    }
    pass[c] = ok;
    PairInfo pi; pi.start = a; pi.n = b - a; pi.qi = qi; pi.extId = extId;
    info[c] = pi;
}

// ------------------------------------------------------------------------------------------------
// Q6: chaining
// ------------------------------------------------------------------------------------------------
struct Cand { int32_t curBegin, curEnd, extBegin, extEnd, score, chainLength, filtered, pad; };
// run of the run-compressed chaining DP (see chainRunDpKernel): [a, b] = its matches
struct __align__(16) Run { int32_t a, curA, extA, backA; int32_t b, curB, extB, scoreB; };
static_assert(sizeof(Run) == sizeof(Cand), "run records live in the candidate array until the chain walk");
static constexpr uint32_t RUNS_PENDING = 0xffffffffu;   // nRuns of a pair whose matches still have to be re-sorted by extPos

__device__ __forceinline__ bool overlapTestDev(const OvParams& P, uint32_t curId, uint32_t extId, int32_t curBegin, int32_t curEnd,
                                               int32_t extBegin, int32_t extEnd, int32_t curLen, int32_t extLen) {
    // overlap.cpp:29-69
    const int32_t cr = curEnd - curBegin, er = extEnd - extBegin;
    if (cr < P.minOverlap || er < P.minOverlap) return false;
    const float lengthDiff = (float)abs(cr - er);
    if (lengthDiff > __fmul_rn(0.5f, (float)min(cr, er))) return false;
    if (P.sameSet && curId == extId) {
        const int32_t intersect = min(curEnd, extEnd) - max(curBegin, extBegin);
        if (intersect > cr / 2) return false;
    }
    if (P.sameSet && curId == (extId ^ 1u)) {
        const int32_t intersect = min(curEnd, extLen - extBegin) - max(curBegin, extLen - extEnd);
        if (intersect > cr / 2) return false;
    }
    const int32_t lrOverhang = max(min(curBegin, extBegin), min(curLen - curEnd, extLen - extEnd));
    if (!P.forceLocal && P.checkOverhang && lrOverhang > P.maxOverhang) return false;
    return true;
}

__device__ __forceinline__ uint32_t filtRank(const uint32_t* __restrict__ filtBits, const uint32_t* __restrict__ filtPrefix, uint64_t qbase, uint32_t x) {
    const uint64_t w = (qbase + x) >> 5;
    return filtPrefix[w] + __popc(filtBits[w] & ((1u << (x & 31)) - 1u));
}

// pair flags
static constexpr uint32_t PAIR_EXTSORTED = 1u;
static constexpr uint32_t PAIR_PRESORTED = 2u;   // scores increase strictly with the match index: the score order is n-1, n-2, ..., 0

// (a) per pair: decide the DP axis (overlap.cpp:269), re-key the elements of ext-sorted pairs by extPos and
// append them to the list of segments that need the std::sort-exact re-sort (:272-274)
// `runs` != nullptr (run-compressed DP): the same pass also finds the runs of the pair (see chainRunsKernel, which then only
// has to visit the pairs that were re-sorted in between).
__global__ void __launch_bounds__(256) pairPrepKernel(Elem* __restrict__ hits, const PairInfo* __restrict__ pairs, const uint32_t* __restrict__ pairIds,
                                                      uint32_t nPairs, const uint32_t* __restrict__ qIds, const uint32_t* __restrict__ len,
                                                      const uint32_t* __restrict__ qlen, uint32_t* __restrict__ pairFlags, Seg* __restrict__ extSegs, uint32_t* __restrict__ nExtSegs,
                                                      Seg* __restrict__ allSegs, int k, Run* __restrict__ runs, int32_t* __restrict__ runOf,
                                                      uint32_t* __restrict__ nRuns) {
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nPairs) return;
    const int lane = threadIdx.x & 31;
    const PairInfo pi = pairs[pairIds[w]];
    const int32_t curLen = (int32_t)qlen[qIds[pi.qi] >> 1], extLen = (int32_t)len[pi.extId >> 1];
    const bool extSorted = extLen > curLen;
    Seg sg; sg.start = pi.start; sg.n = pi.n;
    if (lane == 0) {
        pairFlags[w] = extSorted ? PAIR_EXTSORTED : 0u;
        allSegs[w] = sg;
    }
    if (!extSorted && !runs) return;
    // A pair whose extPos values already increase strictly along the curPos order is in the unique sorted order of
    // distinct keys, i.e. in std::sort's output order: no re-sort (the usual case for a clean overlap).
    Elem* h = hits + pi.start;
    Run* rn = runs ? runs + pi.start : nullptr;
    int32_t* ro = runs ? runOf + pi.start : nullptr;
    const int32_t n = (int32_t)pi.n;
    bool sorted = true;
    int32_t base = 0, carryC = 0, carryE = 0;
    for (int32_t i0 = 0; i0 < n; i0 += 32) {
        const int32_t i = i0 + lane;
        int32_t c = 0, x = 0;
        if (i < n) {
            const Elem e = h[i];
            c = (int32_t)(uint32_t)e.key; x = (int32_t)e.val;
            if (extSorted) { Elem t; t.key = (unsigned long long)e.val; t.val = (unsigned int)e.key; t.aux = 0; h[i] = t; }
        }
        int32_t pc = __shfl_up_sync(0xffffffffu, c, 1), px = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane == 0) { pc = carryC; px = carryE; }
        if (extSorted && i < n && i > 0 && !((uint32_t)px < (uint32_t)x)) sorted = false;
        if (runs) {
            const int32_t dc = c - pc, de = x - px;
            // match 0 is its own run (score 0, :323); match 1 always starts one (its predecessor's score is not >= k)
            const bool head = i < n && (i <= 1 || !(dc == de && 0 < dc && dc < k));
            const uint32_t m = __ballot_sync(0xffffffffu, head);
            const int32_t r = base + __popc(m & (0xffffffffu >> (31 - lane))) - 1;
            if (i < n) ro[i] = r;
            if (head) {
                rn[r].a = i; rn[r].curA = c; rn[r].extA = x;
                if (i > 0) { rn[r - 1].b = i - 1; rn[r - 1].curB = pc; rn[r - 1].extB = px; }
            }
            if (i == n - 1) { rn[r].b = i; rn[r].curB = c; rn[r].extB = x; }
            base += __popc(m);
        }
        carryC = __shfl_sync(0xffffffffu, c, 31); carryE = __shfl_sync(0xffffffffu, x, 31);
    }
    sorted = __all_sync(0xffffffffu, sorted);
    if (lane == 0) {
        if (extSorted && !sorted) extSegs[atomicAdd(nExtSegs, 1u)] = sg;
        if (runs) nRuns[w] = (extSorted && !sorted) ? RUNS_PENDING : (uint32_t)base;   // pending: chainRunsKernel after the re-sort
    }
}

__device__ __forceinline__ int32_t elemCur(const Elem& e, bool extSorted) { return extSorted ? (int32_t)e.val : (int32_t)(uint32_t)e.key; }
__device__ __forceinline__ int32_t elemExt(const Elem& e, bool extSorted) { return extSorted ? (int32_t)(uint32_t)e.key : (int32_t)e.val; }

// (b) chaining DP (overlap.cpp:277-323): one HALF-WARP (16 lanes) per pair, two pairs per warp, both halves in lock
// step so that every collective uses the full-warp mask (partial masks compile to a slow WARPSYNC/COLLECTIVE path).  The
// lanes of a half evaluate 16 predecessors j = i-1, i-2, ... per step in the reference's traversal order; the first
// lane that breaks ends the scan.  The last 16 matches (cur, ext, score) live in lane registers (match j in sub-lane
// j%16) and reach the lanes by shuffle, so the common case (a look-back of ~10 matches) never waits for memory; longer
// look-backs continue from global memory.  Break rule 1 (jumpDiv == 0 && dc < k on an IMPROVING predecessor) is tested
// with warp reductions per candidate instead of a prefix-max scan.  Pairs are visited in order of decreasing size
// (pairOrder) so that the two halves of a warp have the same number of matches to go through.
__device__ __forceinline__ int32_t halfMax(int32_t v, bool upper) {   // max over the 16 lanes of the caller's half
    const int32_t a = __reduce_max_sync(0xffffffffu, upper ? INT32_MIN : v);
    const int32_t b = __reduce_max_sync(0xffffffffu, upper ? v : INT32_MIN);
    return upper ? b : a;
}

__global__ void __launch_bounds__(128) chainDpKernel(const Elem* __restrict__ hits, const PairInfo* __restrict__ pairs, const uint32_t* __restrict__ pairIds,
                                                     const uint32_t* __restrict__ pairOrder, uint32_t nPairs, const uint32_t* __restrict__ pairFlags,
                                                     OvParams P, int32_t* __restrict__ score, int32_t* __restrict__ back, Elem* __restrict__ ord,
                                                     unsigned long long* __restrict__ cellCount) {
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;   // half-warp index
    const int sl = threadIdx.x & 15;
    const bool upper = threadIdx.x & 16;
    const int hs = threadIdx.x & 16;                                       // bit offset of this half in a ballot
    const bool valid = slot < nPairs;
    const uint32_t w = valid ? pairOrder[slot] : 0u;
    PairInfo pi; pi.start = 0; pi.n = 0; pi.qi = 0; pi.extId = 0;
    if (valid) pi = pairs[pairIds[w]];
    const bool extSorted = valid && (pairFlags[w] & PAIR_EXTSORTED);
    const int32_t n = (int32_t)pi.n;
    const int32_t nMax = max(n, __shfl_xor_sync(0xffffffffu, n, 16));
    const int k = P.k;
    const Elem* h = hits + pi.start;
    int32_t* sc = score + pi.start; int32_t* bk = back + pi.start; Elem* od = ord + pi.start;

    // block b = matches [16b, 16b+16): nx* holds the current block, pf* the next one (loaded a block ahead)
    int32_t nxC = 0, nxE = 0, pfC = 0, pfE = 0;
    if (sl < n) { const Elem e = h[sl]; nxC = elemCur(e, extSorted); nxE = elemExt(e, extSorted); }
    if (16 + sl < n) { const Elem e = h[16 + sl]; pfC = elemCur(e, extSorted); pfE = elemExt(e, extSorted); }
    int32_t wC = nxC, wE = nxE, wS = 0;   // window registers: sub-lane 0 holds match 0 (score 0); others not yet valid
    if (sl == 0 && n > 0) { sc[0] = 0; bk[0] = -1; Elem t; t.key = 0x7fffffffULL; t.val = 0; t.aux = 0; od[0] = t; }
    unsigned long long cells = 0;
    for (int32_t i = 1; i < nMax; ++i) {
        const bool act = i < n;
        const int l0 = i & 15;
        if (l0 == 0) {
            nxC = pfC; nxE = pfE;
            const int32_t nb = i + 16 + sl;
            if (nb < n) { const Elem e = h[nb]; pfC = elemCur(e, extSorted); pfE = elemExt(e, extSorted); }
        }
        const int32_t curN = __shfl_sync(0xffffffffu, nxC, l0, 16), extN = __shfl_sync(0xffffffffu, nxE, l0, 16);
        int32_t best = 0, bestId = 0;
        bool stop = !act;
        for (int32_t jb = i - 1;; jb -= 16) {
            const bool live = !stop && jb >= 0;                            // uniform inside a half
            if (!__any_sync(0xffffffffu, live)) break;
            const int32_t j = jb - sl;
            const bool in = live && j >= 0;
            int32_t cj, ej, sj;
            if (jb == i - 1) {   // the 16 most recent matches: registers
                const int src = j & 15;
                cj = __shfl_sync(0xffffffffu, wC, src, 16); ej = __shfl_sync(0xffffffffu, wE, src, 16); sj = __shfl_sync(0xffffffffu, wS, src, 16);
            } else {
                if (jb == i - 17) __syncwarp();   // order sub-lane 0's score stores before these loads
                cj = 0; ej = 0; sj = 0;
                if (in) { const Elem e = h[j]; cj = elemCur(e, extSorted); ej = elemExt(e, extSorted); sj = sc[j]; }
            }
            const int32_t dc = curN - cj, de = extN - ej;
            const bool ok = in && 0 < dc && dc < P.maxJump && 0 < de && de < P.maxJump;
            const int32_t jd = abs(dc - de);
            const int32_t gap = jd > 100 ? 2 * jd : (jd >> 1);            // int32(float(LG_GAP|SM_GAP) * jd), :299
            const int32_t s = ok ? sj + min(min(dc, de), k) - gap : INT32_MIN;
            if (jb == i - 1) {
                // fast path (the usual case on low-error reads): the nearest predecessor already improves and satisfies
                // the first break rule, so the reference's scan ends after one step.  Taken when it holds for both halves.
                const uint32_t fb = __ballot_sync(0xffffffffu, sl == 0 && (!act || (ok && jd == 0 && dc < k && s > 0)));
                if (fb == 0x00010001u) {
                    const int32_t s0 = __shfl_sync(0xffffffffu, s, 0, 16);
                    if (act) { best = s0; bestId = i - 1; ++cells; }
                    break;
                }
            }
            // second break rule (sorted-axis distance) first: nothing beyond its first lane is ever evaluated
            const uint32_t far = (__ballot_sync(0xffffffffu, in && (extSorted ? de : dc) > P.maxJump) >> hs) & 0xffffu;
            int stopLane = far ? (__ffs(far) - 1) : 15;
            bool brk = far != 0;
            uint32_t pot = (__ballot_sync(0xffffffffu, ok && jd == 0 && dc < k) >> hs) & ((2u << stopLane) - 1u);
            while (__any_sync(0xffffffffu, pot != 0)) {
                const int t = pot ? (__ffs(pot) - 1) : 0;
                const int32_t prior = max(best, halfMax((pot && sl < t) ? s : INT32_MIN, upper));
                const int32_t st = __shfl_sync(0xffffffffu, s, t, 16);
                if (pot) {
                    if (st > prior) { stopLane = t; brk = true; pot = 0; }
                    else pot &= pot - 1;
                }
            }
            const bool part = ok && sl <= stopLane;
            const int32_t mx = halfMax(part ? s : INT32_MIN, upper);
            const uint32_t wm = (__ballot_sync(0xffffffffu, part && s == mx) >> hs) & 0xffffu;
            if (live) {
                if (mx > best) { best = mx; bestId = jb - (__ffs(wm) - 1); }
                cells += min(jb + 1, stopLane + 1);
                stop = brk;
            }
        }
        if (act) {
            const int32_t sci = max(best, k);
            if (sl == l0) { wC = curN; wE = extN; wS = sci; }   // match i replaces match i-16 in the window
            if (sl == 0) {
                sc[i] = sci; bk[i] = best > k ? bestId : -1;
                Elem t; t.key = (unsigned long long)(0x7fffffff - sci); t.val = (unsigned int)i; t.aux = 0; od[i] = t;   // (c) input of the score sort
            }
        }
    }
    if (sl == 0 && cells) atomicAdd(cellCount, cells);
}

// The same DP with exact pruning of the look-back.  Most cells of a low-error data set are spent by matches that lie off
// the chain's diagonal: no predecessor can improve them, yet the reference walks back through every match within
// maxJump.  Here every completed block of 16 matches leaves a summary (maximum score, range of diagonals cur - ext,
// sorted-axis coordinate of its first match) in the registers of its half-warp.  For a block of predecessors
//     s(j) = score[j] + min(dc, de, k) - gap(|diag_i - diag_j|)  <=  maxScore + k - gap(dist(diag_i, [minDiag, maxDiag]))
// (gap is monotone), so when that bound is <= the best score found so far no member of the block can improve — and a
// predecessor that does not improve can neither become the back pointer (strict >, overlap.cpp:303) nor trigger the
// first break rule (:307).  A block that also lies entirely within maxJump on the sorted axis (second break rule,
// :314) is therefore skipped without looking at its matches; 16 blocks are tested per step, one per lane.  Blocks that
// fail the test, that contain the maxJump boundary, or that are older than the 16 summaries kept in registers are
// evaluated exactly as in chainDpKernel.  Scores, back pointers and the cell count (skipped cells are counted: the
// reference evaluates them) are identical.
__global__ void __launch_bounds__(128) chainDpPrunedKernel(const Elem* __restrict__ hits, const PairInfo* __restrict__ pairs,
                                                           const uint32_t* __restrict__ pairIds, const uint32_t* __restrict__ pairOrder, uint32_t nPairs,
                                                           const uint32_t* __restrict__ pairFlags, OvParams P, int32_t* __restrict__ score,
                                                           int32_t* __restrict__ back, Elem* __restrict__ ord, unsigned long long* __restrict__ cellCount) {
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;   // half-warp index
    const int sl = threadIdx.x & 15;
    const bool upper = threadIdx.x & 16;
    const int hs = threadIdx.x & 16;                                       // bit offset of this half in a ballot
    const bool valid = slot < nPairs;
    const uint32_t w = valid ? pairOrder[slot] : 0u;
    PairInfo pi; pi.start = 0; pi.n = 0; pi.qi = 0; pi.extId = 0;
    if (valid) pi = pairs[pairIds[w]];
    const bool extSorted = valid && (pairFlags[w] & PAIR_EXTSORTED);
    const int32_t n = (int32_t)pi.n;
    const int32_t nMax = max(n, __shfl_xor_sync(0xffffffffu, n, 16));
    const int k = P.k;
    const Elem* h = hits + pi.start;
    int32_t* sc = score + pi.start; int32_t* bk = back + pi.start; Elem* od = ord + pi.start;

    int32_t nxC = 0, nxE = 0, pfC = 0, pfE = 0;
    if (sl < n) { const Elem e = h[sl]; nxC = elemCur(e, extSorted); nxE = elemExt(e, extSorted); }
    if (16 + sl < n) { const Elem e = h[16 + sl]; pfC = elemCur(e, extSorted); pfE = elemExt(e, extSorted); }
    int32_t wC = nxC, wE = nxE, wS = 0;   // window registers: match j of the last 16 in sub-lane j % 16
    int32_t sumS = 0, sumLo = 0, sumHi = 0, sumFirst = 0;   // summary of the completed block b with b % 16 == sl
    if (sl == 0 && n > 0) { sc[0] = 0; bk[0] = -1; Elem t; t.key = 0x7fffffffULL; t.val = 0; t.aux = 0; od[0] = t; }
    unsigned long long cells = 0;
    for (int32_t i = 1; i < nMax; ++i) {
        const bool act = i < n;
        const int l0 = i & 15;
        if (l0 == 0) {
            nxC = pfC; nxE = pfE;
            const int32_t nb = i + 16 + sl;
            if (nb < n) { const Elem e = h[nb]; pfC = elemCur(e, extSorted); pfE = elemExt(e, extSorted); }
        }
        const int32_t curN = __shfl_sync(0xffffffffu, nxC, l0, 16), extN = __shfl_sync(0xffffffffu, nxE, l0, 16);
        int32_t best = 0, bestId = 0;
        bool stop = !act;

        // exact evaluation of the 16 predecessors jb, jb-1, ... (those with j <= jLim) for the halves with `want`
        auto evalStep = [&](const int32_t jb, const int32_t jLim, const bool fromRegs, const bool want) {
            const int32_t j = jb - sl;
            const bool in = want && j >= 0 && j <= jLim;
            int32_t cj = 0, ej = 0, sj = 0;
            if (fromRegs) {
                const int src = j & 15;
                cj = __shfl_sync(0xffffffffu, wC, src, 16); ej = __shfl_sync(0xffffffffu, wE, src, 16); sj = __shfl_sync(0xffffffffu, wS, src, 16);
            } else if (in) { const Elem e = h[j]; cj = elemCur(e, extSorted); ej = elemExt(e, extSorted); sj = sc[j]; }
            const int32_t dc = curN - cj, de = extN - ej;
            const bool ok = in && 0 < dc && dc < P.maxJump && 0 < de && de < P.maxJump;
            const int32_t jd = abs(dc - de);
            const int32_t gap = jd > 100 ? 2 * jd : (jd >> 1);            // int32(float(LG_GAP|SM_GAP) * jd), :299
            const int32_t s = ok ? sj + min(min(dc, de), k) - gap : INT32_MIN;
            if (fromRegs) {
                // fast path (the usual case on low-error reads): the nearest predecessor already improves and satisfies
                // the first break rule, so the reference's scan ends after one step.  Taken when it holds for both halves.
                const uint32_t fb = __ballot_sync(0xffffffffu, sl == 0 && (!act || (ok && jd == 0 && dc < k && s > 0)));
                if (fb == 0x00010001u) {
                    const int32_t s0 = __shfl_sync(0xffffffffu, s, 0, 16);
                    if (act) { best = s0; bestId = i - 1; ++cells; stop = true; }
                    return;
                }
            }
            // second break rule (sorted-axis distance) first: nothing beyond its first lane is ever evaluated
            const uint32_t far = (__ballot_sync(0xffffffffu, in && (extSorted ? de : dc) > P.maxJump) >> hs) & 0xffffu;
            int stopLane = far ? (__ffs(far) - 1) : 15;
            bool brk = far != 0;
            uint32_t pot = (__ballot_sync(0xffffffffu, ok && jd == 0 && dc < k) >> hs) & ((2u << stopLane) - 1u);
            while (__any_sync(0xffffffffu, pot != 0)) {
                const int t = pot ? (__ffs(pot) - 1) : 0;
                const int32_t prior = max(best, halfMax((pot && sl < t) ? s : INT32_MIN, upper));
                const int32_t st = __shfl_sync(0xffffffffu, s, t, 16);
                if (pot) {
                    if (st > prior) { stopLane = t; brk = true; pot = 0; }
                    else pot &= pot - 1;
                }
            }
            const bool part = ok && sl <= stopLane;
            const int32_t mx = halfMax(part ? s : INT32_MIN, upper);
            const uint32_t wm = (__ballot_sync(0xffffffffu, part && s == mx) >> hs) & 0xffffu;
            if (want) {
                if (mx > best) { best = mx; bestId = jb - (__ffs(wm) - 1); }
                cells += max(0, min(jb + 1, stopLane + 1) - max(0, jb - jLim));
                if (brk) stop = true;
            }
        };

        evalStep(i - 1, i - 1, true, !stop);
        // look-back beyond the 16 most recent matches: block by block, skipping what provably cannot matter
        int32_t jLim = i - 17;                 // highest predecessor not looked at yet
        int32_t nb = jLim >= 0 ? jLim >> 4 : -1;   // its block
        if (nb < 0) stop = true;
        const int32_t cb = i >> 4;             // blocks cb-16 .. cb-1 have their summaries in registers
        const int32_t diagN = curN - extN, sortedN = extSorted ? extN : curN;
        while (__any_sync(0xffffffffu, !stop)) {
            __syncwarp();   // orders sub-lane 0's score stores before the loads below
            const bool live = !stop;
            const int32_t B = nb - sl;
            const int src = B & 15;
            const int32_t bS = __shfl_sync(0xffffffffu, sumS, src, 16), bLo = __shfl_sync(0xffffffffu, sumLo, src, 16);
            const int32_t bHi = __shfl_sync(0xffffffffu, sumHi, src, 16), bFirst = __shfl_sync(0xffffffffu, sumFirst, src, 16);
            const int32_t jdMin = diagN < bLo ? bLo - diagN : (diagN > bHi ? diagN - bHi : 0);
            const int32_t gapMin = jdMin > 100 ? 2 * jdMin : (jdMin >> 1);
            const bool skippable = live && B >= 0 && B >= cb - 16 && bS + k - gapMin <= best && sortedN - bFirst <= P.maxJump;
            const uint32_t keep = (__ballot_sync(0xffffffffu, !skippable) >> hs) & 0xffffu;
            const int nskip = keep ? __ffs(keep) - 1 : 16;
            if (live && nskip > 0) {
                cells += (unsigned long long)(jLim - 16 * nb + 1) + 16ULL * (nskip - 1);
                nb -= nskip;
                jLim = 16 * nb + 15;
                if (nb < 0) stop = true;
            }
            const bool evalWant = live && !stop && nskip < 16;
            if (__any_sync(0xffffffffu, evalWant)) evalStep(16 * nb + 15, jLim, false, evalWant);
            if (evalWant) {
                --nb; jLim = 16 * nb + 15;
                if (nb < 0) stop = true;
            }
        }
        if (act) {
            const int32_t sci = max(best, k);
            if (sl == l0) { wC = curN; wE = extN; wS = sci; }   // match i replaces match i-16 in the window
            if (sl == 0) {
                sc[i] = sci; bk[i] = best > k ? bestId : -1;
                Elem t; t.key = (unsigned long long)(0x7fffffff - sci); t.val = (unsigned int)i; t.aux = 0; od[i] = t;   // (c) input of the score sort
            }
        }
        if (l0 == 15) {   // block i >> 4 is complete: its summary goes to sub-lane (i >> 4) % 16
            const int32_t dg = wC - wE;
            const int32_t mS = halfMax(wS, upper), mHi = halfMax(dg, upper), mLo = -halfMax(-dg, upper);
            const int32_t fst = __shfl_sync(0xffffffffu, extSorted ? wE : wC, 0, 16);
            if (sl == (cb & 15)) { sumS = mS; sumLo = mLo; sumHi = mHi; sumFirst = fst; }
        }
    }
    if (sl == 0 && cells) atomicAdd(cellCount, cells);
}

// ------------------------------------------------------------------------------------------------
// Run-compressed chaining DP (default).  A RUN is a maximal sequence of consecutive matches i-1 -> i of a pair with
// 0 < dc == de < k (same diagonal, closer than k).  For such an i the reference's scan (overlap.cpp:284-316) looks at
// j = i-1 only: it is valid, scores score[i-1] + dc > 0, and satisfies the first break rule (:307) — so
//     score[i] = score[i-1] + dc,  back[i] = i-1                      for every non-first match of a run (i >= 2),
// and scores inside a run are score[a] + cur[i] - cur[a].  Only the first match of every run (its HEAD) needs a scan, and
// the scan can go run by run instead of match by match: walking a run from its last match towards its first, dc and de
// grow by the same amount the score shrinks, so score[j] + min(dc, de, k) never increases; the first VALID match of the
// run is the only one that can pass the strict `>` of :303, and with it the only one that can trigger the break of :307.
// The second break rule (:314-315) depends on the sorted-axis coordinate, which is monotone over the whole list, so it
// cuts the scan at a run boundary or makes the rest of a run invalid.  Scores and back pointers are identical to the
// match-by-match scan; the work drops from O(matches within maxJump) to O(runs within maxJump) per head and to O(1)
// per other match (on low-error reads a run holds ~30 matches).
//   chainRunsKernel   finds the runs of a pair (one warp per pair), writes their records and the run of every match; the pairs
//                     that need no re-sort by extPos get this done by pairPrepKernel, which reads their matches anyway
//   chainRunDpKernel  the scan of the heads: half-warp per pair, 16 runs per step, the 16 most recent runs in registers
//   chainFillKernel   scores / back pointers / score-sort input of all matches from the run records
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) chainRunsKernel(const Elem* __restrict__ hits, const PairInfo* __restrict__ pairs, const uint32_t* __restrict__ pairIds,
                                                       uint32_t nPairs, const uint32_t* __restrict__ pairFlags, int k, Run* __restrict__ runs,
                                                       int32_t* __restrict__ runOf, uint32_t* __restrict__ nRuns) {
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nPairs) return;
    const int lane = threadIdx.x & 31;
    if (nRuns[w] != RUNS_PENDING) return;   // pairPrepKernel already found the runs (no re-sort in between)
    const PairInfo pi = pairs[pairIds[w]];
    const bool extSorted = pairFlags[w] & PAIR_EXTSORTED;
    const int32_t n = (int32_t)pi.n;
    const Elem* h = hits + pi.start;
    Run* rn = runs + pi.start;
    int32_t* ro = runOf + pi.start;
    int32_t base = 0, carryC = 0, carryE = 0;
    for (int32_t i0 = 0; i0 < n; i0 += 32) {
        const int32_t i = i0 + lane;
        int32_t c = 0, x = 0;
        if (i < n) { const Elem e = h[i]; c = elemCur(e, extSorted); x = elemExt(e, extSorted); }
        int32_t pc = __shfl_up_sync(0xffffffffu, c, 1), px = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane == 0) { pc = carryC; px = carryE; }
        const int32_t dc = c - pc, de = x - px;
        // match 0 is its own run (score 0, :323); match 1 always starts one (its predecessor's score is not >= k)
        const bool head = i < n && (i <= 1 || !(dc == de && 0 < dc && dc < k));
        const uint32_t m = __ballot_sync(0xffffffffu, head);
        const int32_t r = base + __popc(m & (0xffffffffu >> (31 - lane))) - 1;
        if (i < n) ro[i] = r;
        if (head) {
            rn[r].a = i; rn[r].curA = c; rn[r].extA = x;
            if (i > 0) { rn[r - 1].b = i - 1; rn[r - 1].curB = pc; rn[r - 1].extB = px; }
        }
        if (i == n - 1) { rn[r].b = i; rn[r].curB = c; rn[r].extB = x; }
        base += __popc(m);
        carryC = __shfl_sync(0xffffffffu, c, 31); carryE = __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) nRuns[w] = (uint32_t)base;
}

__global__ void __launch_bounds__(128) chainRunDpKernel(const Elem* __restrict__ hits, const PairInfo* __restrict__ pairs, const uint32_t* __restrict__ pairIds,
                                                        const uint32_t* __restrict__ pairOrder, uint32_t nPairs, const uint32_t* __restrict__ pairFlags,
                                                        const uint32_t* __restrict__ nRuns, OvParams P, Run* runs, unsigned long long* __restrict__ cellCount) {
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;   // half-warp index
    const int sl = threadIdx.x & 15;
    const bool upper = threadIdx.x & 16;
    const int hs = threadIdx.x & 16;                                       // bit offset of this half in a ballot
    const bool valid = slot < nPairs;
    const uint32_t w = valid ? pairOrder[slot] : 0u;
    PairInfo pi; pi.start = 0; pi.n = 0; pi.qi = 0; pi.extId = 0;
    if (valid) pi = pairs[pairIds[w]];
    const bool extSorted = valid && (pairFlags[w] & PAIR_EXTSORTED);
    const int32_t R = valid ? (int32_t)nRuns[w] : 0;
    const int32_t Rmax = max(R, __shfl_xor_sync(0xffffffffu, R, 16));
    const int k = P.k;
    const Elem* h = hits + pi.start;
    Run* rn = runs + pi.start;
    const int4* rn4 = reinterpret_cast<const int4*>(rn);   // [2r] = {a, curA, extA | run of backA, backA}, [2r+1] = {b, curB, extB, scoreB}

    // records of the runs [16b, 16b+16): nx* = current block, pf* = next block (loaded a block ahead; scoreB not used from these)
    int4 nxH = make_int4(0, 0, 0, 0), nxT = nxH, pfH = nxH, pfT = nxH;
    if (sl < R) { nxH = rn4[2 * sl]; nxT = rn4[2 * sl + 1]; }
    if (16 + sl < R) { pfH = rn4[2 * (16 + sl)]; pfT = rn4[2 * (16 + sl) + 1]; }
    // window registers: last match of run q (index, cur, ext, score) in sub-lane q % 16; run 0 = match 0 with score 0
    int32_t wB = nxT.x, wC = nxT.y, wE = nxT.z, wS = 0;
    if (sl == 0 && R > 0) { rn[0].scoreB = 0; rn[0].backA = -1; }
    unsigned long long cells = 0;
    for (int32_t r = 1; r < Rmax; ++r) {
        const bool act = r < R;
        const int l0 = r & 15;
        if (l0 == 0) {
            nxH = pfH; nxT = pfT;
            const int32_t nb = r + 16 + sl;
            if (nb < R) { pfH = rn4[2 * nb]; pfT = rn4[2 * nb + 1]; }
        }
        const int32_t curN = __shfl_sync(0xffffffffu, nxH.y, l0, 16), extN = __shfl_sync(0xffffffffu, nxH.z, l0, 16);
        const int32_t sortedN = extSorted ? extN : curN;
        int32_t best = 0, bestId = 0, bestRun = 0;
        bool stop = !act;
        for (int32_t rb = r - 1;; rb -= 16) {
            const bool live = !stop && rb >= 0;                            // uniform inside a half
            if (!__any_sync(0xffffffffu, live)) break;
            const int32_t q = rb - sl;
            const bool in = live && q >= 0;
            int32_t jB, cj, ej, sj;
            if (rb == r - 1) {   // the 16 most recent runs: registers
                const int src = q & 15;
                jB = __shfl_sync(0xffffffffu, wB, src, 16); cj = __shfl_sync(0xffffffffu, wC, src, 16);
                ej = __shfl_sync(0xffffffffu, wE, src, 16); sj = __shfl_sync(0xffffffffu, wS, src, 16);
            } else {
                if (rb == r - 17) __syncwarp();   // order sub-lane 0's score stores before these loads
                jB = 0; cj = 0; ej = 0; sj = 0;
                if (in) { const int4 t = rn4[2 * q + 1]; jB = t.x; cj = t.y; ej = t.z; sj = t.w; }
            }
            const int32_t sortedB = extSorted ? ej : cj;   // the first match of this run the reference's scan visits
            bool has = in;
            if (in && !(cj < curN && ej < extN)) {
                // the run's last match is not a predecessor of the head: its first valid match is the last one with
                // cur < curN and ext < extN (ext = cur - diagonal inside a run)
                const int32_t a = rn[q].a, cA = rn[q].curA;
                const int32_t dg = cj - ej;
                const int32_t X = min(curN, extN + dg);
                if (a == jB || !(cA < X)) has = false;
                else {
                    int32_t lo = a, hi = jB, cLo = cA;   // cur[lo] < X <= cur[hi]
                    while (hi - lo > 1) {
                        const int32_t mid = (lo + hi) >> 1;
                        const int32_t cm = elemCur(h[mid], extSorted);
                        if (cm < X) { lo = mid; cLo = cm; } else hi = mid;
                    }
                    sj -= cj - cLo; cj = cLo; ej = cLo - dg; jB = lo;
                }
            }
            const int32_t dc = curN - cj, de = extN - ej;
            const bool ok = has && dc < P.maxJump && de < P.maxJump;   // dc > 0 and de > 0 by construction
            const int32_t jd = abs(dc - de);
            const int32_t gap = jd > 100 ? 2 * jd : (jd >> 1);            // int32(float(LG_GAP|SM_GAP) * jd), :299
            const int32_t s = ok ? sj + min(min(dc, de), k) - gap : INT32_MIN;
            // second break rule (sorted-axis distance): the scan ends at the first run whose last match is beyond maxJump
            const uint32_t far = (__ballot_sync(0xffffffffu, in && sortedN - sortedB > P.maxJump) >> hs) & 0xffffu;
            int stopLane = far ? (__ffs(far) - 1) : 15;
            bool brk = far != 0;
            uint32_t pot = (__ballot_sync(0xffffffffu, ok && jd == 0 && dc < k) >> hs) & ((2u << stopLane) - 1u);
            while (__any_sync(0xffffffffu, pot != 0)) {
                const int t = pot ? (__ffs(pot) - 1) : 0;
                const int32_t prior = max(best, halfMax((pot && sl < t) ? s : INT32_MIN, upper));
                const int32_t st = __shfl_sync(0xffffffffu, s, t, 16);
                if (pot) {
                    if (st > prior) { stopLane = t; brk = true; pot = 0; }
                    else pot &= pot - 1;
                }
            }
            const bool part = ok && sl <= stopLane;
            const int32_t mx = halfMax(part ? s : INT32_MIN, upper);
            const uint32_t wm = (__ballot_sync(0xffffffffu, part && s == mx) >> hs) & 0xffffu;
            const int32_t jW = __shfl_sync(0xffffffffu, jB, wm ? (__ffs(wm) - 1) : 0, 16);
            if (live) {
                if (mx > best) { best = mx; bestId = jW; bestRun = rb - (__ffs(wm) - 1); }
                cells += min(rb + 1, stopLane + 1);
                stop = brk;
            }
        }
        const int32_t bIdx = __shfl_sync(0xffffffffu, nxT.x, l0, 16), curB = __shfl_sync(0xffffffffu, nxT.y, l0, 16);
        const int32_t extB = __shfl_sync(0xffffffffu, nxT.z, l0, 16);
        if (act) {
            const int32_t sB = max(best, k) + (curB - curN);   // score of the run's last match
            if (sl == l0) { wB = bIdx; wC = curB; wE = extB; wS = sB; }   // run r replaces run r-16 in the window
            // extA of a finished head is not read again: the slot keeps the run of the back pointer for the chain walk
            if (sl == 0) { rn[r].scoreB = sB; rn[r].backA = best > k ? bestId : -1; rn[r].extA = bestRun; }
        }
    }
    if (sl == 0 && cells) atomicAdd(cellCount, cells);
}

// scores, back pointers and the input of the score sort (:331-334) of every match.  When the scores of a pair increase
// strictly with the match index, their descending order is the unique sorted sequence of distinct keys — std::sort's
// output is n-1, n-2, ..., 0 — so the pair is flagged PAIR_PRESORTED and taken off the list of segments to sort.
// PAIRED (default; FG_FILL_PAIRED=0 switches back): two 32-match blocks per iteration — the loads of the second block (run index ->
// run record -> match) are issued next to the first block's, which halves the number of dependent round trips of a warp.
template <bool PAIRED>
__global__ void __launch_bounds__(256) chainFillKernel(const Elem* __restrict__ hits, const PairInfo* __restrict__ pairs, const uint32_t* __restrict__ pairIds,
                                                       uint32_t nPairs, uint32_t* __restrict__ pairFlags, const Run* __restrict__ runs,
                                                       int32_t* __restrict__ score, int32_t* __restrict__ back, Elem* __restrict__ ord,
                                                       Seg* __restrict__ allSegs, unsigned long long* __restrict__ nPresorted) {
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nPairs) return;
    const int lane = threadIdx.x & 31;
    const PairInfo pi = pairs[pairIds[w]];
    const bool extSorted = pairFlags[w] & PAIR_EXTSORTED;
    const int32_t n = (int32_t)pi.n;
    const Elem* h = hits + pi.start;
    const int4* rn4 = reinterpret_cast<const int4*>(runs + pi.start);
    int32_t* sc = score + pi.start; int32_t* bk = back + pi.start; Elem* od = ord + pi.start;
    bool incr = true;
    int32_t carry = INT32_MIN;
    if (PAIRED) {
        for (int32_t i0 = 0; i0 < n; i0 += 64) {
            const int32_t iA = i0 + lane, iB = iA + 32;
            const bool vA = iA < n, vB = iB < n;
            int32_t rA = 0, rB = 0;
            if (vA) rA = bk[iA];   // run of this match (chainRunsKernel)
            if (vB) rB = bk[iB];
            int4 HA = make_int4(0, 0, 0, 0), TA = HA, HB = HA, TB = HA;
            Elem eA, eB; eA.key = 0; eA.val = 0; eA.aux = 0; eB = eA;
            if (vA) { HA = rn4[2 * rA]; TA = rn4[2 * rA + 1]; eA = h[iA]; }
            if (vB) { HB = rn4[2 * rB]; TB = rn4[2 * rB + 1]; eB = h[iB]; }
            int32_t sA = INT32_MAX, sB = INT32_MAX;
            if (vA) { sA = TA.w - (TA.y - elemCur(eA, extSorted)); sc[iA] = sA; bk[iA] = (iA == HA.x) ? HA.w : iA - 1; }
            if (vB) { sB = TB.w - (TB.y - elemCur(eB, extSorted)); sc[iB] = sB; bk[iB] = (iB == HB.x) ? HB.w : iB - 1; }
            int32_t ps = __shfl_up_sync(0xffffffffu, sA, 1);
            if (lane == 0) ps = carry;
            if (vA && !(ps < sA)) incr = false;
            carry = __shfl_sync(0xffffffffu, sA, 31);
            ps = __shfl_up_sync(0xffffffffu, sB, 1);
            if (lane == 0) ps = carry;
            if (vB && !(ps < sB)) incr = false;
            carry = __shfl_sync(0xffffffffu, sB, 31);
        }
    } else {
    for (int32_t i0 = 0; i0 < n; i0 += 32) {
        const int32_t i = i0 + lane;
        int32_t s = INT32_MAX;
        if (i < n) {
            const int32_t r = bk[i];   // run of this match (chainRunsKernel)
            const int4 H = rn4[2 * r], T = rn4[2 * r + 1];
            s = T.w - (T.y - elemCur(h[i], extSorted));
            sc[i] = s;
            bk[i] = (i == H.x) ? H.w : i - 1;
        }
        int32_t ps = __shfl_up_sync(0xffffffffu, s, 1);
        if (lane == 0) ps = carry;
        if (i < n && !(ps < s)) incr = false;
        carry = __shfl_sync(0xffffffffu, s, 31);
    }
    }
    incr = __all_sync(0xffffffffu, incr);
    if (incr) {   // the chain walk takes the order n-1 .. 0 from the flag: nothing to write, nothing to sort
        if (lane == 0) { pairFlags[w] |= PAIR_PRESORTED; allSegs[w].n = 0; atomicAdd(nPresorted, 1ULL); }
        return;
    }
    for (int32_t i0 = 0; i0 < n; i0 += 32) {
        const int32_t i = i0 + lane;
        if (i < n) { Elem t; t.key = (unsigned long long)(0x7fffffff - sc[i]); t.val = (unsigned int)i; t.aux = 0; od[i] = t; }
    }
}

__global__ void __launch_bounds__(256) runCountKeyKernel(const uint32_t* __restrict__ nRuns, uint32_t nPairs, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ idx) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nPairs) return;
    keys[w] = ~nRuns[w];
    idx[w] = w;
}

// order in which chainDpKernel visits the pairs: decreasing number of matches
__global__ void __launch_bounds__(256) pairSizeKeyKernel(const PairInfo* __restrict__ pairs, const uint32_t* __restrict__ pairIds, uint32_t nPairs,
                                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nPairs) return;
    keys[w] = ~pairs[pairIds[w]].n;
    idx[w] = w;
}

// (d)+(e) chain walk in std::sort order of the scores (overlap.cpp:331-427), overlapTest, filtered positions, then
// primary selection (:431-458).  One warp per pair: the back-pointer table is staged in shared memory (the walk is
// pure pointer chasing, so its latency is what matters), the lanes test 32 chain starts at a time and lane 0
// walks the chains in order, re-testing the remaining starts after every walk because a walk consumes pointers.
static constexpr int WALK_CAP = 2048;   // matches per warp whose back pointers are staged in shared memory as 16-bit deltas (4 KB)

// RUNS = true (after the run-compressed DP): the walk goes run by run.  Inside a run every match points to its
// predecessor, and every walk that enters a run continues down to the run's head or to an already consumed match, so the
// consumed matches of a run are always a prefix [a, ce) of it: one 16-bit watermark per run replaces the per-match
// pointers, a walk costs one step per run instead of one per match, and each run can start at most one chain.  Run table
// in shared memory (4 x 16 bit per run: first match, watermark, back pointer of the head and its run); pairs with more
// than RUN_CAP runs or more than 65535 matches take the match-by-match path on the global back pointers.
static constexpr int RUN_CAP = WALK_CAP / 4;

template <bool RUNS>
__global__ void __launch_bounds__(128) chainWalkKernel(const Elem* __restrict__ hits, const PairInfo* __restrict__ pairs, const uint32_t* __restrict__ pairIds,
                                                       uint32_t nPairs, const uint32_t* __restrict__ pairFlags, const uint32_t* __restrict__ qIds,
                                                       const uint64_t* __restrict__ qSlotOff, const uint32_t* __restrict__ len,
                                                       const uint32_t* __restrict__ qlen, const uint32_t* __restrict__ filtBits, const uint32_t* __restrict__ filtPrefix, OvParams P,
                                                       const int32_t* __restrict__ score, int32_t* __restrict__ back, Elem* __restrict__ ord,
                                                       Cand* __restrict__ cands, uint32_t* __restrict__ nCandOut, uint32_t* __restrict__ nKeptOut,
                                                       const uint32_t* __restrict__ nRuns) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nPairs) return;
    const int lane = threadIdx.x & 31;
    const PairInfo pi = pairs[pairIds[w]];
    const bool extSorted = pairFlags[w] & PAIR_EXTSORTED;
    const int32_t n = (int32_t)pi.n;
    const int k = P.k;
    const uint32_t curId = qIds[pi.qi], extId = pi.extId;
    const int32_t curLen = (int32_t)qlen[curId >> 1], extLen = (int32_t)len[extId >> 1];
    const Elem* h = hits + pi.start;
    const int32_t* sc = score + pi.start;
    Elem* od = ord + pi.start; Cand* cd = cands + pi.start;
    const uint64_t qbase = qSlotOff[pi.qi];
    const uint32_t qn = (uint32_t)(curLen - k);

    // back pointers: delta = pos - back[pos] (>= 1), 0 = none.  n <= WALK_CAP < 65536, so a delta fits 16 bits.
    int32_t* bkG = back + pi.start;
    unsigned short* sb = reinterpret_cast<unsigned short*>(smemRaw) + (threadIdx.x >> 5) * WALK_CAP;
    const bool presorted = RUNS && (pairFlags[w] & PAIR_PRESORTED);
    const int32_t R = RUNS ? (int32_t)nRuns[w] : 0;
    const bool byRuns = RUNS && R <= RUN_CAP && n <= 65535;
    const bool staged = !byRuns && n <= WALK_CAP;
    // run table (byRuns); the candidates written below reuse the global run records, so everything needed is staged
    unsigned short* sA = sb; unsigned short* sCe = sb + RUN_CAP; unsigned short* sBk = sb + 2 * RUN_CAP; unsigned short* sBr = sb + 3 * RUN_CAP;
    if (byRuns) {
        const int4* rn4 = reinterpret_cast<const int4*>(cands + pi.start);   // {a, curA, run of backA, backA}
        for (int32_t r = lane; r < R; r += 32) {
            const int4 H = rn4[2 * r];
            sA[r] = (unsigned short)H.x; sCe[r] = (unsigned short)H.x;
            sBk[r] = H.w < 0 ? (unsigned short)0xffff : (unsigned short)H.w; sBr[r] = (unsigned short)H.z;
        }
    } else if (staged)
        for (int32_t i = lane; i < n; i += 32) { const int32_t b = bkG[i]; sb[i] = b < 0 ? 0 : (unsigned short)(i - b); }
    __syncwarp();
    auto runOf = [&](int32_t pos) {   // last run with first match <= pos
        int32_t lo = 0, hi = R;
        while (hi - lo > 1) { const int32_t mid = (lo + hi) >> 1; if ((int32_t)sA[mid] <= pos) lo = mid; else hi = mid; }
        return lo;
    };
    auto hasBack = [&](int32_t pos, int32_t r) {
        if (byRuns) return pos >= (int32_t)sCe[r] && (pos > (int32_t)sA[r] || sBk[r] != 0xffff);
        return staged ? sb[pos] != 0 : bkG[pos] >= 0;
    };

    uint32_t nCand = 0;   // meaningful in lane 0
    for (int32_t t0 = 0; t0 < n; t0 += 32) {
        const int32_t t = t0 + lane;
        const int32_t cs = t < n ? (presorted ? n - 1 - t : (int32_t)od[t].val) : 0;
        const int32_t cr = (byRuns && t < n) ? runOf(cs) : 0;
        uint32_t m = __ballot_sync(0xffffffffu, t < n && hasBack(cs, cr));
        while (m) {
            const int src = __ffs(m) - 1;
            const int32_t chainStart = __shfl_sync(0xffffffffu, cs, src);
            const int32_t startRun = __shfl_sync(0xffffffffu, cr, src);
            if (lane == 0) {
                int32_t firstMatch = 0, chainLength = 0, pos = chainStart;
                if (byRuns) {
                    int32_t r = startRun;
                    for (;;) {
                        const int32_t a = sA[r], ce = sCe[r];
                        if (pos < ce) { firstMatch = pos; ++chainLength; break; }   // a consumed match ends the chain (it is still its first match)
                        sCe[r] = (unsigned short)(pos + 1);
                        if (ce > a) { firstMatch = ce - 1; chainLength += pos - ce + 2; break; }
                        firstMatch = a; chainLength += pos - a + 1;
                        const unsigned short nb = sBk[r];
                        if (nb == 0xffff) break;
                        pos = nb; r = sBr[r];
                    }
                } else if (staged) {
                    for (;;) { firstMatch = pos; ++chainLength; const unsigned short d = sb[pos]; sb[pos] = 0; if (!d) break; pos -= d; }
                } else {
                    // consumed nodes keep their pointer recoverable (value -2-np) for the keepAlignment re-walk
                    while (pos != -1) {
                        firstMatch = pos; ++chainLength;
                        const int32_t v = bkG[pos];
                        const int32_t np = v >= 0 ? v : -1;
                        if (v >= 0) bkG[pos] = -2 - v;
                        pos = np;
                    }
                }
                const Elem eF = h[firstMatch], eL = h[chainStart];
                const int32_t curBegin = elemCur(eF, extSorted), extBegin = elemExt(eF, extSorted);
                const int32_t curEnd = elemCur(eL, extSorted) + k - 1, extEnd = elemExt(eL, extSorted) + k - 1;
                if (overlapTestDev(P, curId, extId, curBegin, curEnd, extBegin, extEnd, curLen, extLen)) {
                    const uint32_t hiPos = min((uint32_t)curEnd + 1u, qn), loPos = min((uint32_t)curBegin, qn);
                    Cand c;
                    c.curBegin = curBegin; c.curEnd = curEnd; c.extBegin = extBegin; c.extEnd = extEnd;
                    c.score = sc[chainStart] - sc[firstMatch] + k - 1;
                    c.chainLength = chainLength;
                    c.filtered = (int32_t)(filtRank(filtBits, filtPrefix, qbase, hiPos) - filtRank(filtBits, filtPrefix, qbase, loPos));
                    c.pad = chainStart;
                    cd[nCand++] = c;
                }
            }
            __syncwarp();
            m = __ballot_sync(0xffffffffu, lane > src && t < n && hasBack(cs, cr));
        }
    }
    __syncwarp();   // all lanes are done reading ord[] before lane 0 reuses its slots
    if (lane != 0) return;
    // primary selection: std::sort by score descending (the ord slots of this pair are free now), then best /
    // containment filter
    for (uint32_t i = 0; i < nCand; ++i) { Elem t; t.key = (unsigned long long)(0x7fffffff - cd[i].score); t.val = i; t.aux = 0; od[i] = t; }
    seqIntrosort(od, (long)nCand);
    uint32_t kept = 0;
    if (P.onlyMaxExt) { if (nCand) { od[0].aux = 1; kept = 1; } }
    else {
        for (uint32_t i = 0; i < nCand; ++i) {
            const Cand c = cd[od[i].val];
            bool contained = false;
            for (uint32_t j = 0; j < i && !contained; ++j) {
                if (!od[j].aux) continue;
                const Cand p = cd[od[j].val];
                contained = p.curBegin <= c.curBegin && c.curEnd <= p.curEnd && p.extBegin <= c.extBegin && c.extEnd <= p.extEnd &&
                            p.score > c.score;
            }
            od[i].aux = contained ? 0u : 1u;
            kept += !contained;
        }
    }
    nCandOut[w] = nCand;
    nKeptOut[w] = kept;
}

// Q7: compact the kept candidates of every pair into the output, in the reference's order
__global__ void __launch_bounds__(256) gatherOverlapsKernel(const PairInfo* __restrict__ pairs, const uint32_t* __restrict__ pairIds, uint32_t nPairs,
                                                            const uint32_t* __restrict__ qIds, const uint32_t* __restrict__ len,
                                                            const uint32_t* __restrict__ qlen, const Elem* __restrict__ ord, const Cand* __restrict__ cands,
                                                            const uint32_t* __restrict__ nCand, const uint64_t* __restrict__ outOff,
                                                            uint32_t qiBase, const uint32_t* __restrict__ pairFlags, fg_overlap* __restrict__ out) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nPairs) return;
    const PairInfo pi = pairs[pairIds[w]];
    const unsigned long long extSortedBit = (pairFlags[w] & PAIR_EXTSORTED) ? (1ULL << 63) : 0ULL;
    const uint32_t curId = qIds[pi.qi];
    uint64_t o = outOff[w];
    const uint32_t nc = nCand[w];
    for (uint32_t i = 0; i < nc; ++i) {
        const Elem e = ord[pi.start + i];
        if (!e.aux) continue;
        const Cand c = cands[pi.start + e.val];
        fg_overlap r;
        r.cur_id = curId; r.cur_begin = c.curBegin; r.cur_end = c.curEnd; r.cur_len = (int32_t)qlen[curId >> 1];
        r.ext_id = pi.extId; r.ext_begin = c.extBegin; r.ext_end = c.extEnd; r.ext_len = (int32_t)len[pi.extId >> 1];
        r.score = c.score; r.seq_divergence = 0.f; r.chain_length = c.chainLength; r.filtered_positions = c.filtered;
        r.edit_distance = -1; r.aln_len = 0; r.reserved = pi.qi - qiBase;
        r.aln_first = (unsigned long long)pi.start | extSortedBit;   // scratch for alignKernel: where the pair's matches live ...
        r.aln_count = (uint32_t)c.pad;                               // ... and the match its chain starts at
        out[o++] = r;
    }
}

// keepAlignment (overlap.cpp:368-377, 398-405): the (cur,ext) anchors of a kept chain, re-walked from its start through
// the back pointers (a consumed pointer v < -1 stands for -2-v) down to its first match.  write == false: count only.
__global__ void __launch_bounds__(256) alignKernel(fg_overlap* __restrict__ ov, uint32_t nOv, const Elem* __restrict__ hits,
                                                   const int32_t* __restrict__ back, int k, bool write,
                                                   const uint64_t* __restrict__ alnOff, uint32_t* __restrict__ alnCount,
                                                   int32_t* __restrict__ alnPairs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nOv) return;
    fg_overlap o = ov[i];
    const bool extSorted = o.aln_first >> 63;
    const uint32_t start = (uint32_t)(o.aln_first & 0xffffffffULL);
    const Elem* h = hits + start;
    const int32_t* bk = back + start;
    int32_t pos = (int32_t)o.aln_count, cnt = 0, lastCur = 0;
    // pass 1 (count) or pass 2 (write in reversed order): list = begin, recorded anchors ascending, end
    int32_t* out = write ? alnPairs + 2 * alnOff[i] : nullptr;
    const int32_t total = write ? (int32_t)alnCount[i] : 0;
    for (;;) {
        const Elem e = h[pos];
        const int32_t cur = elemCur(e, extSorted), ext = elemExt(e, extSorted);
        if (cnt == 0 || lastCur - cur > k) {
            if (write) { const int32_t at = total - 2 - cnt; out[2 * at] = cur; out[2 * at + 1] = ext; }
            ++cnt; lastCur = cur;
        }
        if (cur == o.cur_begin && ext == o.ext_begin) break;
        const int32_t v = bk[pos];
        pos = v >= 0 ? v : -2 - v;
        if (pos < 0) break;   // cannot happen: the chain ends at its first match
    }
    if (!write) { alnCount[i] = (uint32_t)cnt + 2u; return; }
    out[0] = o.cur_begin; out[1] = o.ext_begin;
    out[2 * (total - 1)] = o.cur_end; out[2 * (total - 1) + 1] = o.ext_end;
}

__global__ void __launch_bounds__(256) alignFinishKernel(fg_overlap* __restrict__ ov, uint32_t nOv, const uint64_t* __restrict__ alnOff,
                                                         const uint32_t* __restrict__ alnCount, uint64_t base, bool keep) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nOv) return;
    ov[i].aln_first = keep ? base + alnOff[i] : 0ULL;
    ov[i].aln_count = keep ? alnCount[i] : 0u;
}

// Q7 on the device (the default epilogue; DESIGN.md §2 "Exact shortcuts" 7).  seqDivergence of every gathered record with the
// reference's float expression — glibc's logf evaluated bit for bit (glibc_logf.cuh) — or, with nuclAlignment, from the edit
// distance (alignment.cpp:240-245); the divergence test (overlap.cpp:470-473) against the call's or the query's threshold; the
// number of records every query keeps.  Records whose arithmetic leaves the domain the device handles (zero chain, empty
// alignment: never seen on real input) are counted in kept[nQueriesOfSub]; the sub-batch then takes the host epilogue.
static_assert(sizeof(fg_overlap) == 72 && offsetof(fg_overlap, reserved) == 68 && offsetof(fg_overlap, aln_count) == 64, "gatherKeptKernel copies 9 words per record");
__global__ void __launch_bounds__(256) divergenceKernel(fg_overlap* __restrict__ ov, uint32_t nOv, float sampleRate, int k, bool nucl, float maxDiv,
                                                        const float* __restrict__ queryMaxDiv, uint32_t qFirst, uint32_t nq,
                                                        uint8_t* __restrict__ keep, uint32_t* __restrict__ kept) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nOv) return;
    fg_overlap& o = ov[i];
    bool ok = true;
    float div;
    if (nucl) {
        const int32_t ed = o.edit_distance, al = o.aln_len;
        if (ed < 0) div = 1.0f;
        else if (al <= 0) { ok = false; div = 0.0f; }
        else div = __fdiv_rn(__int2float_rn(ed), __ull2float_rn((unsigned long long)al));   // (float)editDistance / size_t
    } else {
        div = kmerDivergence(o.cur_end - o.cur_begin, o.ext_end - o.ext_begin, o.filtered_positions, o.chain_length, sampleRate, k, ok);
    }
    o.seq_divergence = div;
    const uint32_t q = o.reserved;
    const bool pass = div < (queryMaxDiv ? queryMaxDiv[q] : maxDiv);
    keep[i] = pass;
    if (pass) atomicAdd(&kept[q - qFirst], 1u);
    if (!ok) atomicAdd(&kept[nq], 1u);
}

// the kept records, in order, into the buffer that goes to the host; `reserved` becomes 0 (= passed the divergence test)
__global__ void __launch_bounds__(256) gatherKeptKernel(const unsigned long long* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t nSel,
                                                        unsigned long long* __restrict__ dst) {
    const uint64_t total = 9ULL * nSel, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint32_t j = (uint32_t)(t / 9), w = (uint32_t)(t - 9ULL * j);
        unsigned long long v = src[9ULL * idx[j] + w];
        if (w == 8) v &= 0xffffffffULL;
        dst[t] = v;
    }
}

__global__ void logfSelfTestKernel(const float* __restrict__ x, uint32_t n, float* __restrict__ y) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = glibcLogfInDomain(x[i]) ? glibcLogf(x[i]) : -1.0f;
}

// segmented sort driver.  dCounters: 5 device words (layout above) with [0] = nSegs already set and the rest zero.
static constexpr uint32_t SORT_HUGE_MIN = 8192;   // ranges above this size are partitioned by a whole CTA (sortHugeKernel)
struct SortWorkspace {
    DevBuf<SortTask> small, bigA, bigB, hugeA, hugeB;
    uint32_t capSmall = 0, capBig = 0, capHuge = 0;
    void ensure(uint64_t totalElems, uint32_t maxSegs) {
        capSmall = (uint32_t)(totalElems / 8 + maxSegs + 4096);
        capBig = (uint32_t)(totalElems / sortSmallN() + maxSegs + 64);
        capSmall = (uint32_t)(totalElems * (128.0 / sortSmallN()) / 8 + maxSegs + 4096);
        small.ensure(capSmall); bigA.ensure(capBig); bigB.ensure(capBig);
        capHuge = (uint32_t)(totalElems / SORT_HUGE_MIN + maxSegs + 64);
        hugeA.ensure(capHuge); hugeB.ensure(capHuge);
    }
};

// outArr / tieP: tie-following mode (see rangeIsTieFree): arr = scratch copy of the segments in their original order,
// outArr = the stable-sorted elements; only ranges that contain ties are emulated and delivered to outArr.
// hugeScratch (optional): two arrays of 32-bit words indexed like arr (stop lists of sortHugeKernel) and hugeCounters, two
// zeroed device words; without them every range takes the warp-per-range path.
static void sortSegments(fg_ctx* ctx, Elem* arr, const Seg* dSegs, uint32_t* dCounters, uint32_t maxSegs, SortWorkspace& ws,
                         const char* topName, const char* smallName, SortCfg cfg, Elem* outArr = nullptr, const uint32_t* tieP = nullptr,
                         uint32_t* hugeScratchL = nullptr, uint32_t* hugeScratchR = nullptr, uint32_t* hugeCounters = nullptr) {
    if (!outArr) outArr = arr;
    const bool useHuge = hugeScratchL && hugeScratchR && hugeCounters && envInt("FG_SORT_HUGE", 1, 0, 1);
    const int variant = cfg.variant;
    const uint32_t smallN = (uint32_t)cfg.smallN;
    const int smemBytes = 4 * ((int)smallN * ((int)sizeof(Elem) + (variant ? 4 : 0)) + (variant > 1 ? ((int)smallN / 32 + 2) * 4 : 0));
    auto smallKernel = variant == 0 ? sortSmallKernel<0> : variant == 1 ? sortSmallKernel<1> : variant == 2 ? sortSmallKernel<2> : sortSmallKernel<3>;
    static std::once_flag attrSet[64];   // function attributes are per device
    std::call_once(attrSet[ctx->device & 63], [&] {
        const int maxSmem = 4 * (SORT_SMALL_MAX * ((int)sizeof(Elem) + 4) + (SORT_SMALL_MAX / 32 + 2) * 4);
        FG_CUDA(cudaFuncSetAttribute(sortSmallKernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        FG_CUDA(cudaFuncSetAttribute(sortSmallKernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        FG_CUDA(cudaFuncSetAttribute(sortSmallKernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        FG_CUDA(cudaFuncSetAttribute(sortSmallKernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
    });
    const int blocksPerSm = std::max(1, std::min(8, (int)((220 * 1024) / (smemBytes + 1024))));
    if (!maxSegs) return;
    {
        PhaseTimer pt(ctx, topName);
        sortSeedKernel<<<(maxSegs + 255) / 256, 256, 0, streamOf(ctx)>>>(dSegs, dCounters, ws.bigA.p, dCounters + 3, ws.capBig, ws.small.p,
                                                                      dCounters + 1, ws.capSmall, smallN, tieP,
                                                                      useHuge ? ws.hugeA.p : nullptr, hugeCounters, ws.capHuge, SORT_HUGE_MIN);
        checkLaunch(ctx, "sortSeedKernel");
        if (useHuge) {   // CTA-wide partition steps until no range is longer than SORT_HUGE_MIN
            SortTask* hin = ws.hugeA.p; SortTask* hout = ws.hugeB.p;
            uint32_t* nHin = hugeCounters; uint32_t* nHout = hugeCounters + 1;
            for (int level = 0; level < 200; ++level) {
                uint32_t hCnt = 0;
                FG_CUDA(cudaMemcpyAsync(&hCnt, nHin, 4, cudaMemcpyDeviceToHost, streamOf(ctx)));
                FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
                if (hCnt > ws.capHuge) throw Error(FG_ERR_INTERNAL, "sort task list overflow");
                if (!hCnt) break;
                FG_CUDA(cudaMemsetAsync(nHout, 0, 4, streamOf(ctx)));
                sortHugeKernel<<<hCnt, HUGE_WARPS * 32, 0, streamOf(ctx)>>>(arr, hin, nHin, ws.capHuge, hout, nHout, SORT_HUGE_MIN, ws.bigA.p, dCounters + 3,
                                                                         ws.capBig, ws.small.p, dCounters + 1, ws.capSmall, smallN, outArr, tieP,
                                                                         hugeScratchL, hugeScratchR);
                checkLaunch(ctx, "sortHugeKernel");
                std::swap(hin, hout); std::swap(nHin, nHout);
            }
        }
        SortTask* in = ws.bigA.p; SortTask* out = ws.bigB.p;
        uint32_t* nIn = dCounters + 3; uint32_t* nOut = dCounters + 4;
        uint64_t firstElems = 0;
        for (int level = 0; level < 200; ++level) {
            uint32_t hIn[3] = {0, 0, 0};   // count, (other list), elements
            FG_CUDA(cudaMemcpyAsync(hIn, nIn, 12, cudaMemcpyDeviceToHost, streamOf(ctx)));
            FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
            const uint32_t cnt = hIn[0], elems = hIn[2];
            if (cnt > ws.capBig) throw Error(FG_ERR_INTERNAL, "sort task list overflow");
            if (!cnt) break;
            if (level == 0) firstElems = elems;
            // tail: little work left in unevenly split ranges -> finish each range with one warp instead of paying a
            // launch + host round trip per remaining level
            // ... and a handful of short ranges from the start (the few pairs whose matches are not presorted): the same
            const bool fewShort = level == 0 && cnt <= 8192 && (uint64_t)elems <= (uint64_t)cnt * 4096;
            if (fewShort || (level >= 4 && ((uint64_t)elems * 16 < firstElems || level >= 48))) {
                sortTailKernel<<<(cnt + 3) / 4, 128, 0, streamOf(ctx)>>>(arr, in, nIn, ws.capBig, ws.small.p, dCounters + 1, ws.capSmall, smallN, outArr, tieP);
                checkLaunch(ctx, "sortTailKernel");
                break;
            }
            FG_CUDA(cudaMemsetAsync(nOut, 0, 4, streamOf(ctx)));
            FG_CUDA(cudaMemsetAsync(nOut + 2, 0, 4, streamOf(ctx)));
            sortLevelKernel<<<(cnt + 3) / 4, 128, 0, streamOf(ctx)>>>(arr, in, nIn, ws.capBig, out, nOut, ws.small.p, dCounters + 1, ws.capSmall, smallN, outArr, tieP);
            checkLaunch(ctx, "sortLevelKernel");
            std::swap(in, out); std::swap(nIn, nOut);
        }
    }
    {
        PhaseTimer pt(ctx, smallName);
        smallKernel<<<148 * blocksPerSm, 128, smemBytes, streamOf(ctx)>>>(arr, ws.small.p, dCounters + 1, ws.capSmall, dCounters + 2, smallN, outArr);
        checkLaunch(ctx, "sortSmallKernel");
    }
}

// test hook: sort caller-provided segments with the device introsort
void debugWarpSort(fg_ctx* ctx, uint64_t* keys, uint32_t* vals, const uint64_t* segOffsets, uint32_t nSegs) {
    if (!nSegs) return;
    const uint64_t n = segOffsets[nSegs];
    if (n >= (1ULL << 31)) throw Error(FG_ERR_ARG, "too many elements");
    std::vector<Elem> h(n);
    for (uint64_t i = 0; i < n; ++i) { h[i].key = keys[i]; h[i].val = vals[i]; h[i].aux = 0; }
    std::vector<Seg> hs(nSegs);
    for (uint32_t i = 0; i < nSegs; ++i) { hs[i].start = (uint32_t)segOffsets[i]; hs[i].n = (uint32_t)(segOffsets[i + 1] - segOffsets[i]); }
    DevBuf<Elem> d(std::max<uint64_t>(n, 1));
    DevBuf<Seg> dSegs(nSegs);
    DevBuf<uint32_t> counters(16), hugeScratch(2 * std::max<uint64_t>(n, 1));
    SortWorkspace ws; ws.ensure(n, nSegs);
    const uint32_t cap = ws.capSmall;
    uint32_t hc[16] = {nSegs, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    FG_CUDA(cudaMemcpyAsync(d.p, h.data(), n * sizeof(Elem), cudaMemcpyHostToDevice, streamOf(ctx)));
    FG_CUDA(cudaMemcpyAsync(dSegs.p, hs.data(), nSegs * sizeof(Seg), cudaMemcpyHostToDevice, streamOf(ctx)));
    FG_CUDA(cudaMemcpyAsync(counters.p, hc, sizeof hc, cudaMemcpyHostToDevice, streamOf(ctx)));
    sortSegments(ctx, d.p, dSegs.p, counters.p, nSegs, ws, "dbg_sort_top", "dbg_sort_small", envInt("FG_DEBUG_SORT_PAIRS", 0, 0, 1) ? sortCfgPairs() : sortCfgHits(),
                 nullptr, nullptr, hugeScratch.p, hugeScratch.p + n, counters.p + 8);
    FG_CUDA(cudaMemcpyAsync(h.data(), d.p, n * sizeof(Elem), cudaMemcpyDeviceToHost, streamOf(ctx)));
    FG_CUDA(cudaMemcpyAsync(hc, counters.p, sizeof hc, cudaMemcpyDeviceToHost, streamOf(ctx)));
    FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
    if (hc[1] > cap) throw Error(FG_ERR_INTERNAL, "sort task list overflow");
    for (uint64_t i = 0; i < n; ++i) { keys[i] = h[i].key; vals[i] = h[i].val; }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
template <class InIt, class OutT>
static void exclusiveScanToPlus1(fg_ctx* ctx, InIt in, OutT* outPlus1Base, uint64_t n) {
    // outPlus1Base[0] = 0, outPlus1Base[i+1] = sum(in[0..i])
    FG_CUDA(cudaMemsetAsync(outPlus1Base, 0, sizeof(OutT), streamOf(ctx)));
    if (!n) return;
    size_t tmpBytes = 0;
    FG_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmpBytes, in, outPlus1Base + 1, (int)n, streamOf(ctx)));
    DevBuf<char> tmp(tmpBytes);
    FG_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmpBytes, in, outPlus1Base + 1, (int)n, streamOf(ctx)));
    ++ctx->launches;
    FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
}

template <class FlagT>
static uint32_t selectFlagged(fg_ctx* ctx, const FlagT* flags, uint32_t n, DevBuf<uint32_t>& out) {
    out.ensure(std::max<uint32_t>(n, 1));
    if (!n) return 0;
    DevBuf<uint32_t> dNum(1);
    cub::CountingInputIterator<uint32_t> it(0);
    size_t tmpBytes = 0;
    FG_CUDA(cub::DeviceSelect::Flagged(nullptr, tmpBytes, it, flags, out.p, dNum.p, (int)n, streamOf(ctx)));
    DevBuf<char> tmp(tmpBytes);
    FG_CUDA(cub::DeviceSelect::Flagged(tmp.p, tmpBytes, it, flags, out.p, dNum.p, (int)n, streamOf(ctx)));
    ++ctx->launches;
    uint32_t h = 0;
    FG_CUDA(cudaMemcpyAsync(&h, dNum.p, 4, cudaMemcpyDeviceToHost, streamOf(ctx)));
    FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
    return h;
}

// The device epilogue (divergenceKernel) is only used on a context whose DEVICE logf has reproduced the HOST's logf — the function
// the reference calls — on a sample of 2^16 arguments: every binade, dense around the values 1 / matchRate takes.  The arithmetic is
// proven on the CPU (tests/cpu_models/logf_check.cpp); this guards against a host libm that is not the glibc the proof was run on.
// FG_DEVICE_EPILOGUE=0 keeps the host epilogue (recordDivergences + sliceEpilogue).  Main thread, before the lanes start.
static bool deviceEpilogueUsable(fg_ctx* ctx) {
    if (const char* e = getenv("FG_DEVICE_EPILOGUE")) if (atoi(e) == 0) return false;   // read per call: the tests switch it
    if (ctx->devEpilogueState) return ctx->devEpilogueState > 0;
    const uint32_t n = 1u << 16;
    std::vector<float> x(n), y(n);
    uint64_t rs = 0x9E3779B97F4A7C15ULL;
    for (uint32_t i = 0; i < n; ++i) {
        rs = splitmix64(rs);
        uint32_t bits;
        if (i & 1) bits = 0x3f800000u + (uint32_t)(rs % 0x01800000u);                       // [1, 8)
        else bits = 0x00800000u + (uint32_t)(rs % (0x7f800000u - 0x00800000u));             // any positive normal float
        std::memcpy(&x[i], &bits, 4);
    }
    DevBuf<float> dX(n), dY(n);
    FG_CUDA(cudaMemcpyAsync(dX.p, x.data(), n * 4, cudaMemcpyHostToDevice, streamOf(ctx)));
    logfSelfTestKernel<<<n / 256, 256, 0, streamOf(ctx)>>>(dX.p, n, dY.p);
    checkLaunch(ctx, "logfSelfTestKernel");
    FG_CUDA(cudaMemcpyAsync(y.data(), dY.p, n * 4, cudaMemcpyDeviceToHost, streamOf(ctx)));
    FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
    bool same = true;
    for (uint32_t i = 0; i < n && same; ++i) {
        volatile float xv = x[i];
        const float ref = std::log((float)xv);
        same = std::memcmp(&ref, &y[i], 4) == 0;
    }
    ctx->devEpilogueState = same ? 1 : -1;
    return same;
}

// one chunk of queries (< 2^30 k-mer slots): lookup, expansion and the per-sub-batch pipeline; the raw overlap records are
// appended to the context's pinned buffer with `reserved` = position of the query in the whole call
// Work counters of a call, added up by the lanes.
struct BatchTotals { std::atomic<uint64_t> pairs{0}, dpPairs{0}, cells{0}, tied{0}, presorted{0}, gathered{0}; };
// Sub-batches hand their raw records to the shared pinned buffer IN ORDER (sub-batch i directly behind sub-batch i-1), so the
// buffer holds the records in query order whatever lane produced them: a sub-batch learns its offset when its predecessor has.
struct OrderedCommit {
    std::mutex m; std::condition_variable cv;
    size_t next = 0;       // index of the sub-batch whose turn it is
    size_t nRaw = 0;       // records reserved so far
    bool failed = false;
    // per query of the whole call, filled by the lanes (sliceEpilogue): first record in the pinned buffer, records kept
    std::vector<size_t> qStart; std::vector<uint32_t> kept;
    std::atomic<bool> badOrder{false};
    size_t reserve(size_t index, size_t n) {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return next == index || failed; });
        if (failed) throw Error(FG_ERR_INTERNAL, "another lane failed");
        const size_t off = nRaw;
        nRaw += n; next = index + 1;
        lk.unlock(); cv.notify_all();
        return off;
    }
    void fail() { { std::lock_guard<std::mutex> lk(m); failed = true; } cv.notify_all(); }
};

static void recordDivergences(fg_ctx* ctx, fg_overlap* recs, size_t n, const fg_overlap_params& prm);
static void sliceEpilogue(fg_ctx* ctx, fg_overlap* recs, size_t n, size_t sliceOff, const fg_overlap_params& prm, OrderedCommit& commit,
                          size_t qFirst, size_t qCount);

static void overlapsChunk(fg_ctx* ctx, const uint32_t* queryIds, uint32_t nQ, uint32_t qOffset, const fg_overlap_params& prm, const OvParams& P,
                          OrderedCommit& commit, size_t& subBase, uint64_t& totHits, BatchTotals& tot, const float* dQueryMaxDiv, const bool devEpilogue) {
    const int k = ctx->k;
    // where the query sequences live: the indexed reads themselves, or the second set of fg_queries_upload
    const std::vector<uint32_t>& qHLen = P.sameSet ? ctx->hLen : ctx->hQsLen;
    const uint64_t* qSeq = P.sameSet ? ctx->dSeq.p : ctx->dQsSeq.p;
    const uint64_t* qWordOff = P.sameSet ? ctx->dWordOff.p : ctx->dQsWordOff.p;
    const uint32_t* qLen = P.sameSet ? ctx->dLen.p : ctx->dQsLen.p;
    uint32_t maxSeqLen = 1;   // bounds curPos and extPos (packed hits of the segmented sort)
    for (uint32_t L : ctx->hLen) maxSeqLen = std::max(maxSeqLen, L);
    for (uint32_t L : qHLen) maxSeqLen = std::max(maxSeqLen, L);

    // query slot space and tiles
    std::unique_ptr<HostTimer> hostPrep(new HostTimer(ctx, "host_chunk_prep"));   // wall clock up to the first sub-batch (incl. "gather" lookups)
    std::vector<uint64_t> hQSlotOff(nQ + 1, 0);
    std::vector<uint2> hQTiles;
    std::vector<size_t> qTileFirst(nQ + 1, 0);
    for (uint32_t i = 0; i < nQ; ++i) {
        const uint32_t L = qHLen[queryIds[i] >> 1];
        const uint32_t n = L > (uint32_t)k ? L - k : 0;
        qTileFirst[i] = hQTiles.size();
        for (uint32_t t = 0; t < n; t += QTILE) hQTiles.push_back(make_uint2(i, t));
        hQSlotOff[i + 1] = hQSlotOff[i] + ((n + 31u) & ~31u);
    }
    qTileFirst[nQ] = hQTiles.size();
    const uint64_t nQSlots = hQSlotOff[nQ];
    std::vector<uint64_t> hQHitOff(nQ + 1, 0);

    std::unique_ptr<HostTimer> prepPart(new HostTimer(ctx, "prep_alloc_copy"));
    DevBuf<uint32_t> dQIds(std::max<uint32_t>(nQ, 1));
    DevBuf<uint64_t> dQSlotOff(nQ + 1);
    DevBuf<uint2> dQTiles(std::max<size_t>(hQTiles.size(), 1));
    DevBuf<uint32_t> hitCnt(nQSlots + 1), filtBits(nQSlots / 32 + 2), filtPrefix(nQSlots / 32 + 2);
    DevBuf<uint64_t> slotInfo(nQSlots + 1), hitOff(nQSlots + 2), dQHitOff(nQ + 1);
    if (nQ) FG_CUDA(cudaMemcpyAsync(dQIds.p, queryIds, nQ * 4ULL, cudaMemcpyHostToDevice, streamOf(ctx)));
    FG_CUDA(cudaMemcpyAsync(dQSlotOff.p, hQSlotOff.data(), (nQ + 1) * 8ULL, cudaMemcpyHostToDevice, streamOf(ctx)));
    if (!hQTiles.empty()) FG_CUDA(cudaMemcpyAsync(dQTiles.p, hQTiles.data(), hQTiles.size() * sizeof(uint2), cudaMemcpyHostToDevice, streamOf(ctx)));
    FG_CUDA(cudaMemsetAsync(filtBits.p, 0, filtBits.bytes(), streamOf(ctx)));
    FG_CUDA(cudaMemsetAsync(hitCnt.p, 0, hitCnt.bytes(), streamOf(ctx)));

    prepPart.reset(new HostTimer(ctx, "prep_lookup_wall"));
    if (!hQTiles.empty()) {
        if (nQSlots >= (1ULL << 31)) throw Error(FG_ERR_ARG, "query batch too large (>= 2^31 k-mer slots); split the call");
        {
            L2Pin pinBits(ctx, ctx->dIdxBits.p, ctx->dIdxBits.bytes(), 4);
            PhaseTimer pt(ctx, "lookup");
            queryLookupKernel<<<(unsigned)hQTiles.size(), 256, 0, streamOf(ctx)>>>(qSeq, qWordOff, qLen, ctx->dSlotOff.p,
                                                                                ctx->dSelBits.p, dQIds.p, dQSlotOff.p, dQTiles.p, k,
                                                                                P.sameSet, ctx->indexTable, ctx->dIdxBits.p, hitCnt.p, slotInfo.p, filtBits.p);
            checkLaunch(ctx, "queryLookupKernel");
            cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> it64(hitCnt.p, CastU64());
            exclusiveScanToPlus1(ctx, it64, hitOff.p, nQSlots);
            cub::TransformInputIterator<uint32_t, PopcU32, const uint32_t*> itPop(filtBits.p, PopcU32());
            exclusiveScanToPlus1(ctx, itPop, filtPrefix.p, nQSlots / 32);
            gatherQueryHitOffKernel<<<(nQ + 256) / 256, 256, 0, streamOf(ctx)>>>(hitOff.p, dQSlotOff.p, nQ, dQHitOff.p);
            checkLaunch(ctx, "gatherQueryHitOffKernel");
            FG_CUDA(cudaMemcpyAsync(hQHitOff.data(), dQHitOff.p, (nQ + 1) * 8ULL, cudaMemcpyDeviceToHost, streamOf(ctx)));
            FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
        }
    }
    totHits += hQHitOff[nQ];
    prepPart.reset(new HostTimer(ctx, "prep_budget"));

    // sub-batches of consecutive queries with a bounded number of hits, processed by a few lanes (host thread + stream +
    // arena each).  Hits per sub-batch: as many as comfortably fit (about 80 B of workspace per hit and lane), but at least
    // two sub-batches per lane when the call is big enough, so that the lanes can overlap each other's host work (counts
    // coming back from the device, result copies, the per-record epilogue).
    const int nLanes = envInt("FG_LANES", 2, 1, 8);
    if (!ctx->hitBudget) {
        uint64_t budget = 768ULL << 20;
        size_t freeB = 0, totalB = 0;
        if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
            uint64_t cached = ctx->arena.cachedFreeBytes();
            for (auto& l : ctx->lanes) cached += l->arena.cachedFreeBytes();
            const uint64_t usable = (uint64_t)((freeB + cached) * 0.55);
            budget = std::max<uint64_t>(64ULL << 20, std::min<uint64_t>(budget, usable / 80));
        }
        ctx->hitBudget = budget;
    }
    uint64_t budget = std::max<uint64_t>(32ULL << 20, ctx->hitBudget / nLanes);
    budget = std::min<uint64_t>(budget, std::max<uint64_t>(48ULL << 20, hQHitOff[nQ] / ((uint64_t)envInt("FG_SUBS_PER_LANE", 1, 1, 16) * nLanes) + 1));
    if (const char* e = getenv("FG_HIT_BUDGET")) budget = std::max<uint64_t>(1024, strtoull(e, nullptr, 10));
    struct Sub { uint32_t qa, qb; };
    std::vector<Sub> subs;
    for (uint32_t qa = 0; qa < nQ;) {
        uint32_t qb = qa + 1;
        while (qb < nQ && hQHitOff[qb + 1] - hQHitOff[qa] <= budget) ++qb;
        subs.push_back({qa, qb});
        qa = qb;
    }
    prepPart.reset();
    hostPrep.reset();
    HostTimer hostSub(ctx, "host_subbatches");   // wall clock of all sub-batches
    PinnedBuf<fg_overlap>& pinned = ctx->pinnedOut;

    auto runSub = [&](const uint32_t qa, const uint32_t qb, const size_t subIndex) {
        bool committed = false;
        auto commitNone = [&] { if (!committed) { commit.reserve(subIndex, 0); committed = true; } };
        const uint64_t hitBase = hQHitOff[qa], M = hQHitOff[qb] - hitBase;
        const uint32_t nq = qb - qa;
        if (M == 0) { commitNone(); return; }
        DevBuf<Elem> hits, ord; DevBuf<int32_t> score, back; DevBuf<Cand> cands; DevBuf<uint8_t> flags;
        DevBuf<uint32_t> gStart, candIds, pairIds;
        SortWorkspace ws; DevBuf<Seg> segsQ;
        DevBuf<uint32_t> counters(32);
        DevBuf<uint8_t> qTie;
        if (M >= (1ULL << 31)) throw Error(FG_ERR_ARG, "a single query produced >= 2^31 k-mer hits");
        hits.ensure(M); ord.ensure(M); score.ensure(M + 1); back.ensure(M); cands.ensure(M); flags.ensure(M + 1);
        ws.ensure(M, nq);
        const uint32_t taskCap = ws.capSmall;
        segsQ.ensure(nq);
        uint32_t hCounters[32] = {0};
        hCounters[0] = nq;
        FG_CUDA(cudaMemcpyAsync(counters.p, hCounters, sizeof hCounters, cudaMemcpyHostToDevice, streamOf(ctx)));
        const size_t tA = qTileFirst[qa], tB = qTileFirst[qb];
        // tie-free fast path: stable radix sort by (query, extId).  FG_HIT_RADIX: 1 (default) = segmented sort of packed 64-bit
        // hits, one CTA per query; 2 = the library's radix sort on 32-bit (query, extId) keys; 0 = exact emulation for every query
        int idBits = 1, qBits = 1, posBits = 1;
        while ((2ULL * ctx->nReads) >> idBits) ++idBits;
        while (((uint64_t)nq) >> qBits) ++qBits;
        while (((uint64_t)maxSeqLen) >> posBits) ++posBits;
        int radixMode = envInt("FG_HIT_RADIX", 1, 0, 2);
        if (radixMode == 1 && (idBits > 24 || idBits + 2 * posBits > 64)) radixMode = 2;
        if (radixMode == 2 && idBits + qBits > 32) radixMode = 0;
        const bool radixPath = radixMode != 0;
        uint8_t* tieFlags = reinterpret_cast<uint8_t*>(back.p);   // M + 1 bytes of the back-pointer array (free until the chaining)
        if (radixPath) {
            qTie.ensure(nq);
            FG_CUDA(cudaMemsetAsync(qTie.p, 0, nq, streamOf(ctx)));
        }
        if (radixMode == 1) {
            // scratch: ord (16 B per hit, free until the DP) holds the two 8-byte ping-pong buffers
            unsigned long long* bufA = reinterpret_cast<unsigned long long*>(ord.p);
            unsigned long long* bufB = bufA + M;
            {
                PhaseTimer pt(ctx, "expand");
                expandKernel<2><<<(unsigned)(tB - tA), 256, 0, streamOf(ctx)>>>(ctx->dLen.p, qLen, ctx->dEntries.p, dQIds.p, dQSlotOff.p, dQTiles.p + tA, k,
                                                                             hitOff.p, slotInfo.p, hitBase, nullptr, nullptr, nullptr, bufA, qa, idBits, nullptr,
                                                                             posBits, qTie.p);
                checkLaunch(ctx, "expandKernel");
            }
            int nPass = (idBits + 7) / 8;
            {   // (tests: more passes than the ids need exercise the three-pass ping-pong on small inputs; the extra digits are 0)
                const int forced = envInt("FG_SRS_MIN_PASSES", 1, 1, 3);
                if (forced > nPass && 2 * posBits + 8 * forced <= 64) nPass = forced;
            }
            const int segSort = envInt("FG_SEG_SORT", 2, 0, 2);   // 2 (default) = segTileSortKernel, 0 = segRadixSortKernel, 1 = segRadixSortClusterKernel
            if (segSort == 2) {
                static std::once_flag tsAttr[64];   // function attributes are per device
                std::call_once(tsAttr[ctx->device & 63], [&] {
                    FG_CUDA(cudaFuncSetAttribute(segTileSortKernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TsSmem)));
                    FG_CUDA(cudaFuncSetAttribute(segTileSortKernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TsSmem)));
                });
                PhaseTimer pt(ctx, "hit_sort_radix");
                auto sortK = envInt("FG_SEG_OCC", 3, 3, 4) == 4 ? segTileSortKernel<4> : segTileSortKernel<3>;
                sortK<<<nq, TS_WARPS * 32, sizeof(TsSmem), streamOf(ctx)>>>(bufA, bufB, dQHitOff.p, qa, hitBase, posBits, nPass, hits.p, flags.p);
                checkLaunch(ctx, "segTileSortKernel");
            } else if (segSort == 1) {   // 1 = segRadixSortClusterKernel (L2-resident; measured slower, see its comment)
                // cluster size: sources, scratch copies and fresh output of all resident clusters should fit the L2 together.
                // The typical hit lives in a segment of the hits-weighted mean size.
                double sum2 = 0.0; uint64_t maxSeg = 1;
                for (uint32_t q = qa; q < qb; ++q) { const uint64_t n = hQHitOff[q + 1] - hQHitOff[q]; sum2 += (double)n * (double)n; maxSeg = std::max(maxSeg, n); }
                const double segW = sum2 / (double)M;
                const double l2Budget = (double)envInt("FG_SRS_L2_MB", 64, 1, 4096) * 1048576.0;
                int cSize = 1;
                while (cSize < 8 && (148.0 / cSize) * segW * 32.0 > l2Budget) cSize *= 2;
                if (const int forced = envInt("FG_SRS_CLUSTER", 0, 0, 8)) cSize = forced == 1 || forced == 2 || forced == 4 || forced == 8 ? forced : cSize;
                static std::once_flag srsAttr[64];   // function attributes are per device
                std::call_once(srsAttr[ctx->device & 63], [&] {
                    FG_CUDA(cudaFuncSetAttribute(segRadixSortClusterKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SrsSmem)));
                });
                cudaLaunchConfig_t cfg{};
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = (unsigned)cSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.blockDim = dim3(SRS_WARPS * 32); cfg.dynamicSmemBytes = sizeof(SrsSmem); cfg.stream = streamOf(ctx);
                cfg.attrs = attr; cfg.numAttrs = 1;
                cfg.gridDim = dim3((unsigned)cSize);
                int maxClusters = 0;
                FG_CUDA(cudaOccupancyMaxActiveClusters(&maxClusters, segRadixSortClusterKernel, &cfg));
                if (maxClusters < 1) throw Error(FG_ERR_CUDA, "segRadixSortClusterKernel: no cluster of the requested size fits the device");
                const uint32_t nClusters = std::min<uint32_t>((uint32_t)maxClusters, nq);
                cfg.gridDim = dim3(nClusters * (unsigned)cSize);
                const uint64_t stride = (maxSeg + 31) & ~31ULL;
                const int nBufs = nPass >= 3 ? 2 : 1;
                DevBuf<unsigned long long> srsScratch((uint64_t)nClusters * nBufs * stride);
                DevBuf<uint32_t> nextQuery(1);
                FG_CUDA(cudaMemsetAsync(nextQuery.p, 0, 4, streamOf(ctx)));
                PhaseTimer pt(ctx, "hit_sort_radix");
                FG_CUDA(cudaLaunchKernelEx(&cfg, segRadixSortClusterKernel, (const unsigned long long*)bufA, srsScratch.p, stride, nBufs, (const uint64_t*)dQHitOff.p, qa, nq,
                                           hitBase, posBits, nPass, hits.p, flags.p, nextQuery.p));
                checkLaunch(ctx, "segRadixSortClusterKernel");
            } else {
                PhaseTimer pt(ctx, "hit_sort_radix");
                const bool deep = envInt("FG_SRS_DEEP", 1, 0, 1) != 0;   // more loads in flight per thread (see the kernel's comment)
                auto sortK = envInt("FG_SEG_OCC", 3, 3, 4) == 4 ? (deep ? segRadixSortKernel<4, true> : segRadixSortKernel<4, false>)
                                                                : (deep ? segRadixSortKernel<3, true> : segRadixSortKernel<3, false>);   // resident CTAs per SM (32 / 40 registers)
                sortK<<<nq, SEG_WARPS * 32, 0, streamOf(ctx)>>>(bufA, bufB, dQHitOff.p, qa, hitBase, posBits, nPass, hits.p, flags.p);
                checkLaunch(ctx, "segRadixSortKernel");
            }
        } else if (radixMode == 2) {
            // scratch: the arrays of the later chaining stages are free until the DP (ord = 16 B per hit: the two key and the two
            // value buffers; the second half of the 32-B candidate slots: the payload; its first half is the scratch copy below)
            uint32_t* keyA = reinterpret_cast<uint32_t*>(ord.p);
            uint32_t* keyB = keyA + M;
            uint32_t* valA = keyB + M;
            uint32_t* valB = valA + M;
            unsigned long long* payload = reinterpret_cast<unsigned long long*>(reinterpret_cast<Elem*>(cands.p) + M);
            {
                PhaseTimer pt(ctx, "expand");
                expandKernel<1><<<(unsigned)(tB - tA), 256, 0, streamOf(ctx)>>>(ctx->dLen.p, qLen, ctx->dEntries.p, dQIds.p, dQSlotOff.p, dQTiles.p + tA, k,
                                                                             hitOff.p, slotInfo.p, hitBase, nullptr, keyA, valA, payload, qa, idBits, nullptr, 0, nullptr);
                checkLaunch(ctx, "expandKernel");
            }
            cub::DoubleBuffer<uint32_t> dk(keyA, keyB), dv(valA, valB);
            {
                PhaseTimer pt(ctx, "hit_sort_radix_lib");
                size_t tb = 0;
                FG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)M, 0, idBits + qBits, streamOf(ctx)));
                DevBuf<char> tmpS(tb);
                FG_CUDA(cub::DeviceRadixSort::SortPairs(tmpS.p, tb, dk, dv, (int)M, 0, idBits + qBits, streamOf(ctx)));
                ctx->launches += 2 + (idBits + qBits + 7) / 8;
            }
            {
                PhaseTimer pt(ctx, "hit_sort_gather");
                gatherSortedHitsKernel<<<gridFor(M, 256, 16), 256, 0, streamOf(ctx)>>>(dk.Current(), dv.Current(), payload, M, idBits, hits.p, qTie.p, tieFlags);
                checkLaunch(ctx, "gatherSortedHitsKernel");
            }
        }
        if (radixPath) {
            querySegsKernel<<<(nq + 255) / 256, 256, 0, streamOf(ctx)>>>(dQHitOff.p, qa, nq, hitBase, segsQ.p, qTie.p, counters.p + 24);
            checkLaunch(ctx, "querySegsKernel");
            uint32_t hTied = 0;
            FG_CUDA(cudaMemcpyAsync(&hTied, counters.p + 24, 4, cudaMemcpyDeviceToHost, streamOf(ctx)));
            FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
            tot.tied += hTied;
            if (hTied && radixMode == 1) {
                PhaseTimer pt(ctx, "hit_sort_top");
                FG_CUDA(cudaMemsetAsync(tieFlags, 0, M + 1, streamOf(ctx)));
                tieFlagKernel<<<nq, 256, 0, streamOf(ctx)>>>(hits.p, dQHitOff.p, qa, hitBase, qTie.p, tieFlags);
                checkLaunch(ctx, "tieFlagKernel");
            }
            if (hTied) {
                // queries with ties: expanded again in the original order into a scratch array (the candidate array, free
                // until the chain walk) and sorted by the exact introsort emulation, which only follows the ranges that
                // contain ties; everything else of these queries keeps the stable-sorted order already in `hits`
                Elem* scratch = reinterpret_cast<Elem*>(cands.p);
                uint32_t* tieP = reinterpret_cast<uint32_t*>(score.p);
                {
                    PhaseTimer pt(ctx, "hit_sort_top");
                    cub::TransformInputIterator<uint32_t, CastU8, const uint8_t*> itF(tieFlags, CastU8());
                    size_t tb = 0;
                    FG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, itF, tieP, (int)(M + 1), streamOf(ctx)));
                    DevBuf<char> tmpS(tb);
                    FG_CUDA(cub::DeviceScan::ExclusiveSum(tmpS.p, tb, itF, tieP, (int)(M + 1), streamOf(ctx)));
                    ++ctx->launches;
                }
                {
                    PhaseTimer pt(ctx, "hit_sort_top");   // re-expansion of the queries with ties
                    expandKernel<0><<<(unsigned)(tB - tA), 256, 0, streamOf(ctx)>>>(ctx->dLen.p, qLen, ctx->dEntries.p, dQIds.p, dQSlotOff.p, dQTiles.p + tA, k,
                                                                                     hitOff.p, slotInfo.p, hitBase, scratch, nullptr, nullptr, nullptr, qa, 0, qTie.p, 0, nullptr);
                    checkLaunch(ctx, "expandKernel");
                }
                sortSegments(ctx, scratch, segsQ.p, counters.p, nq, ws, "hit_sort_top", "hit_sort_small", sortCfgHits(), hits.p, tieP,
                             reinterpret_cast<uint32_t*>(ord.p), reinterpret_cast<uint32_t*>(ord.p) + M, counters.p + 25);
            }
        } else {
            {
                PhaseTimer pt(ctx, "expand");
                expandKernel<0><<<(unsigned)(tB - tA), 256, 0, streamOf(ctx)>>>(ctx->dLen.p, qLen, ctx->dEntries.p, dQIds.p, dQSlotOff.p, dQTiles.p + tA, k,
                                                                                 hitOff.p, slotInfo.p, hitBase, hits.p, nullptr, nullptr, nullptr, qa, 0, nullptr, 0, nullptr);
                checkLaunch(ctx, "expandKernel");
            }
            querySegsKernel<<<(nq + 255) / 256, 256, 0, streamOf(ctx)>>>(dQHitOff.p, qa, nq, hitBase, segsQ.p, nullptr, nullptr);
            checkLaunch(ctx, "querySegsKernel");
            sortSegments(ctx, hits.p, segsQ.p, counters.p, nq, ws, "hit_sort_top", "hit_sort_small", sortCfgHits(), nullptr, nullptr,
                         reinterpret_cast<uint32_t*>(ord.p), reinterpret_cast<uint32_t*>(ord.p) + M, counters.p + 25);
        }
        uint32_t G = 0, C = 0, Pn = 0;
        DevBuf<PairInfo> pairInfo;
        DevBuf<uint8_t> candFlag, passFlag;
        {
            PhaseTimer pt(ctx, "group");
            if (radixMode != 1) {   // (the segmented sort writes the flags itself)
                groupFlagKernel<<<gridFor(M), 256, 0, streamOf(ctx)>>>(hits.p, M, flags.p);
                checkLaunch(ctx, "groupFlagKernel");
            }
            queryStartFlagKernel<<<(nq + 255) / 256, 256, 0, streamOf(ctx)>>>(dQHitOff.p, qa, nq, hitBase, M, flags.p);
            checkLaunch(ctx, "queryStartFlagKernel");
            G = selectFlagged(ctx, flags.p, (uint32_t)M, gStart);
            candFlag.alloc(std::max<uint32_t>(G, 1));
            groupCandidateKernel<<<(G + 255) / 256, 256, 0, streamOf(ctx)>>>(gStart.p, G, M, P.minUniqueF, candFlag.p);
            checkLaunch(ctx, "groupCandidateKernel");
            C = selectFlagged(ctx, candFlag.p, G, candIds);
            pairInfo.alloc(std::max<uint32_t>(C, 1)); passFlag.alloc(std::max<uint32_t>(C, 1));
            if (C) {
                pairFilterKernel<<<(C + 7) / 8, 256, 0, streamOf(ctx)>>>(hits.p, gStart.p, G, M, candIds.p, C, dQHitOff.p, qa, nq, hitBase,
                                                                      dQIds.p, ctx->dLen.p, qLen, P, passFlag.p, pairInfo.p);
                checkLaunch(ctx, "pairFilterKernel");
                Pn = selectFlagged(ctx, passFlag.p, C, pairIds);
            }
        }
        tot.pairs += G; tot.dpPairs += Pn;
        if (Pn) {
            DevBuf<uint32_t> nCand(Pn), nKept(Pn);
            DevBuf<uint64_t> outOff(Pn + 1);
            DevBuf<unsigned long long> dCells(2);   // [0] predecessor evaluations of the DP, [1] pairs whose score order needed no sort
            FG_CUDA(cudaMemsetAsync(dCells.p, 0, 16, streamOf(ctx)));
            DevBuf<uint32_t> pairFlags(Pn), nRuns;
            DevBuf<Seg> extSegs(Pn), allSegs(Pn);
            // FG_DP_MODE: 2 = run-compressed DP and chain walk (default), 1 = match-by-match with pruned look-back, 0 = match-by-match
            // (the run-compressed DP relies on every in-run step passing the jump test, i.e. on k < max_jump — overlap.cpp:289 tests
            // dc < _maxJump; any other setting takes the literal match-by-match DP)
            const int dpMode = P.maxJump > k ? envInt("FG_DP_MODE", 2, 0, 2) : 0;   // read per call: the tests switch it
            {
                PhaseTimer pt(ctx, "chain_prep");
                if (dpMode == 2) nRuns.alloc(Pn);
                pairPrepKernel<<<(Pn + 7) / 8, 256, 0, streamOf(ctx)>>>(hits.p, pairInfo.p, pairIds.p, Pn, dQIds.p, ctx->dLen.p, qLen, pairFlags.p,
                                                                     extSegs.p, counters.p + 8, allSegs.p, k, dpMode == 2 ? reinterpret_cast<Run*>(cands.p) : nullptr,
                                                                     back.p, nRuns.p);
                checkLaunch(ctx, "pairPrepKernel");
            }
            sortSegments(ctx, hits.p, extSegs.p, counters.p + 8, Pn, ws, "chain_extsort_top", "chain_extsort_small", sortCfgPairs());
            {
                // visit the pairs by decreasing size: two pairs share a warp, 8 a block
                DevBuf<uint32_t> szKeyA(Pn), szKeyB(Pn), ordA(Pn), ordB(Pn);
                Run* runs = reinterpret_cast<Run*>(cands.p);
                if (dpMode == 2) {
                    PhaseTimer pt(ctx, "chain_runs");
                    chainRunsKernel<<<(Pn + 7) / 8, 256, 0, streamOf(ctx)>>>(hits.p, pairInfo.p, pairIds.p, Pn, pairFlags.p, k, runs, back.p, nRuns.p);
                    checkLaunch(ctx, "chainRunsKernel");
                }
                cub::DoubleBuffer<uint32_t> dk(szKeyA.p, szKeyB.p), dv(ordA.p, ordB.p);
                {
                    PhaseTimer pt(ctx, "chain_order");
                    if (dpMode == 2) {
                        runCountKeyKernel<<<(Pn + 255) / 256, 256, 0, streamOf(ctx)>>>(nRuns.p, Pn, szKeyA.p, ordA.p);
                        checkLaunch(ctx, "runCountKeyKernel");
                    } else {
                        pairSizeKeyKernel<<<(Pn + 255) / 256, 256, 0, streamOf(ctx)>>>(pairInfo.p, pairIds.p, Pn, szKeyA.p, ordA.p);
                        checkLaunch(ctx, "pairSizeKeyKernel");
                    }
                    size_t tb = 0;
                    FG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)Pn, 0, 32, streamOf(ctx)));
                    DevBuf<char> tmpS(tb);
                    FG_CUDA(cub::DeviceRadixSort::SortPairs(tmpS.p, tb, dk, dv, (int)Pn, 0, 32, streamOf(ctx)));
                    ctx->launches += 5;
                }
                if (dpMode == 2) {
                    {
                        PhaseTimer pt(ctx, "chain_dp");
                        chainRunDpKernel<<<(Pn + 7) / 8, 128, 0, streamOf(ctx)>>>(hits.p, pairInfo.p, pairIds.p, dv.Current(), Pn, pairFlags.p, nRuns.p, P, runs,
                                                                               dCells.p);   // whole warps: no early exit inside
                        checkLaunch(ctx, "chainRunDpKernel");
                    }
                    PhaseTimer pt(ctx, "chain_fill");
                    auto fillK = envInt("FG_FILL_PAIRED", 1, 0, 1) ? chainFillKernel<true> : chainFillKernel<false>;
                    fillK<<<(Pn + 7) / 8, 256, 0, streamOf(ctx)>>>(hits.p, pairInfo.p, pairIds.p, Pn, pairFlags.p, runs, score.p, back.p, ord.p,
                                                                          allSegs.p, dCells.p + 1);
                    checkLaunch(ctx, "chainFillKernel");
                } else {
                    PhaseTimer pt(ctx, "chain_dp");
                    auto dp = dpMode == 1 ? chainDpPrunedKernel : chainDpKernel;
                    dp<<<(Pn + 7) / 8, 128, 0, streamOf(ctx)>>>(hits.p, pairInfo.p, pairIds.p, dv.Current(), Pn, pairFlags.p, P, score.p, back.p, ord.p,
                                                             dCells.p);   // whole warps: no early exit inside
                    checkLaunch(ctx, "chainDpKernel");
                }
            }
            FG_CUDA(cudaMemcpyAsync(counters.p + 16, &Pn, 4, cudaMemcpyHostToDevice, streamOf(ctx)));
            sortSegments(ctx, ord.p, allSegs.p, counters.p + 16, Pn, ws, "chain_ordsort_top", "chain_ordsort_small", sortCfgPairs());
            {
                PhaseTimer pt(ctx, "chain_walk");
                const int walkSmem = 4 * WALK_CAP * (int)sizeof(unsigned short);
                static std::once_flag walkAttr[64];   // function attributes are per device
                std::call_once(walkAttr[ctx->device & 63], [&] {
                    FG_CUDA(cudaFuncSetAttribute(chainWalkKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, walkSmem));
                    FG_CUDA(cudaFuncSetAttribute(chainWalkKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, walkSmem));
                });
                auto walk = dpMode == 2 ? chainWalkKernel<true> : chainWalkKernel<false>;
                walk<<<(Pn + 3) / 4, 128, walkSmem, streamOf(ctx)>>>(hits.p, pairInfo.p, pairIds.p, Pn, pairFlags.p, dQIds.p, dQSlotOff.p,
                                                               ctx->dLen.p, qLen, filtBits.p, filtPrefix.p, P, score.p, back.p, ord.p,
                                                               cands.p, nCand.p, nKept.p, nRuns.p);
                checkLaunch(ctx, "chainWalkKernel");
            }
            {
                uint32_t hc[32];
                FG_CUDA(cudaMemcpyAsync(hc, counters.p, sizeof hc, cudaMemcpyDeviceToHost, streamOf(ctx)));
                FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
                if (hc[1] > taskCap || hc[9] > taskCap || hc[17] > taskCap) throw Error(FG_ERR_INTERNAL, "sort task list overflow");
            }
            HostTimer pt(ctx, "host_results");   // gather of the kept overlaps, device -> pinned host copies (includes "edit")
            cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> it64(nKept.p, CastU64());
            exclusiveScanToPlus1(ctx, it64, outOff.p, Pn);
            uint64_t nOut = 0; unsigned long long cells[2] = {0, 0};
            FG_CUDA(cudaMemcpyAsync(&nOut, outOff.p + Pn, 8, cudaMemcpyDeviceToHost, streamOf(ctx)));
            FG_CUDA(cudaMemcpyAsync(cells, dCells.p, 16, cudaMemcpyDeviceToHost, streamOf(ctx)));
            FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
            tot.cells += cells[0]; tot.presorted += cells[1]; tot.gathered += nOut;
            if (nOut) {
                DevBuf<fg_overlap> dOut(nOut);
                gatherOverlapsKernel<<<(Pn + 255) / 256, 256, 0, streamOf(ctx)>>>(pairInfo.p, pairIds.p, Pn, dQIds.p, ctx->dLen.p, qLen, ord.p, cands.p,
                                                                               nCand.p, outOff.p, 0u - qOffset, pairFlags.p, dOut.p);
                checkLaunch(ctx, "gatherOverlapsKernel");
                {   // kmerMatches of the kept overlaps, or just clearing the scratch the gather left in aln_first / aln_count
                    if (nOut >= (1ULL << 31)) throw Error(FG_ERR_ARG, "too many overlaps in one sub-batch");
                    const uint32_t nO = (uint32_t)nOut;
                    DevBuf<uint32_t> alnCount(nO); DevBuf<uint64_t> alnOff(nO + 1);
                    if (prm.keep_alignment) {
                        alignKernel<<<(nO + 255) / 256, 256, 0, streamOf(ctx)>>>(dOut.p, nO, hits.p, back.p, k, false, nullptr, alnCount.p, nullptr);
                        checkLaunch(ctx, "alignKernel");
                        cub::TransformInputIterator<uint64_t, CastU64, const uint32_t*> itc(alnCount.p, CastU64());
                        exclusiveScanToPlus1(ctx, itc, alnOff.p, nO);
                        uint64_t nPts = 0;
                        FG_CUDA(cudaMemcpyAsync(&nPts, alnOff.p + nO, 8, cudaMemcpyDeviceToHost, streamOf(ctx)));
                        FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
                        DevBuf<int32_t> alnPairs(std::max<uint64_t>(2 * nPts, 2));
                        alignKernel<<<(nO + 255) / 256, 256, 0, streamOf(ctx)>>>(dOut.p, nO, hits.p, back.p, k, true, alnOff.p, alnCount.p, alnPairs.p);
                        checkLaunch(ctx, "alignKernel");
                        std::lock_guard<std::mutex> alnLock(ctx->alnMutex);   // the pair array is shared by the lanes: append under a lock
                        const uint64_t base = ctx->resAln.size() / 2;
                        alignFinishKernel<<<(nO + 255) / 256, 256, 0, streamOf(ctx)>>>(dOut.p, nO, alnOff.p, alnCount.p, base, true);
                        checkLaunch(ctx, "alignFinishKernel");
                        ctx->resAln.resize(ctx->resAln.size() + 2 * nPts);
                        FG_CUDA(cudaMemcpyAsync(ctx->resAln.data() + 2 * base, alnPairs.p, 2 * nPts * 4, cudaMemcpyDeviceToHost, streamOf(ctx)));
                        FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
                    } else {
                        alignFinishKernel<<<(nO + 255) / 256, 256, 0, streamOf(ctx)>>>(dOut.p, nO, nullptr, nullptr, 0, false);
                        checkLaunch(ctx, "alignFinishKernel");
                    }
                }
                // grows the shared pinned buffer to `need` records: exclusively, when no lane is copying into or working on it
                auto growPinned = [&](std::shared_lock<std::shared_mutex>& rd, size_t need) {
                    if (pinned.n >= need) return;
                    rd.unlock();
                    {
                        std::unique_lock<std::shared_mutex> wr(ctx->pinnedMutex);
                        size_t reserved;   // slices of later sub-batches may already hold records: keep everything reserved so far
                        { std::lock_guard<std::mutex> lk(commit.m); reserved = commit.nRaw; }
                        if (pinned.n < need) pinned.ensureKeep(std::max(need, reserved), std::min(reserved, pinned.n));
                    }
                    rd.lock();
                };
                if (devEpilogue) {
                    // device epilogue: edit distances, seqDivergence (glibc's logf, bit for bit), the divergence test and the per-query
                    // counts are computed where the records are; only the kept records travel, already compact and in order
                    const uint32_t nO = (uint32_t)nOut;
                    if (prm.nucl_alignment) {   // overlap.cpp:463-468
                        PhaseTimer pe(ctx, "edit");
                        editDistances(ctx, dOut.p, nullptr, nO, prm.use_hpc != 0, !P.sameSet, prm.max_divergence, dQueryMaxDiv);
                    }
                    DevBuf<uint8_t> keepFlag(nO);
                    DevBuf<uint32_t> dKept(nq + 1), selIdx;   // [nq] = records the device does not handle
                    FG_CUDA(cudaMemsetAsync(dKept.p, 0, (nq + 1) * 4ULL, streamOf(ctx)));
                    uint32_t nSel = 0;
                    {
                        PhaseTimer pt(ctx, "epilogue_dev");
                        divergenceKernel<<<(nO + 255) / 256, 256, 0, streamOf(ctx)>>>(dOut.p, nO, ctx->stats.sample_rate, k, prm.nucl_alignment != 0, prm.max_divergence,
                                                                                  dQueryMaxDiv, qOffset + qa, nq, keepFlag.p, dKept.p);
                        checkLaunch(ctx, "divergenceKernel");
                        nSel = selectFlagged(ctx, keepFlag.p, nO, selIdx);
                    }
                    std::vector<uint32_t> hKept(nq + 1);
                    FG_CUDA(cudaMemcpyAsync(hKept.data(), dKept.p, (nq + 1) * 4ULL, cudaMemcpyDeviceToHost, streamOf(ctx)));
                    FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
                    if (hKept[nq] == 0) {
                        DevBuf<fg_overlap> dKeep(std::max<uint32_t>(nSel, 1));
                        if (nSel) {
                            PhaseTimer pt(ctx, "epilogue_dev");
                            gatherKeptKernel<<<gridFor(9ULL * nSel, 256, 16), 256, 0, streamOf(ctx)>>>(reinterpret_cast<const unsigned long long*>(dOut.p), selIdx.p, nSel,
                                                                                                  reinterpret_cast<unsigned long long*>(dKeep.p));
                            checkLaunch(ctx, "gatherKeptKernel");
                        }
                        const size_t myOff = commit.reserve(subIndex, nSel);
                        committed = true;
                        std::shared_lock<std::shared_mutex> rd(ctx->pinnedMutex);
                        growPinned(rd, myOff + nSel);
                        if (nSel) {
                            FG_CUDA(cudaMemcpyAsync(pinned.p + myOff, dKeep.p, nSel * sizeof(fg_overlap), cudaMemcpyDeviceToHost, streamOf(ctx)));
                            FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
                        }
                        size_t run = myOff;   // where every query's records start, how many it keeps (what sliceEpilogue derives on the host path)
                        for (uint32_t j = 0; j < nq; ++j) {
                            const size_t q = (size_t)qOffset + qa + j;
                            if (hKept[j]) commit.qStart[q] = run;
                            commit.kept[q] = hKept[j];
                            run += hKept[j];
                        }
                        if (run != myOff + nSel) throw Error(FG_ERR_INTERNAL, "device epilogue: per-query counts do not add up");
                        return;
                    }
                    // (a record outside the device's domain: the whole sub-batch takes the host epilogue below)
                }
                // this sub-batch's slice of the shared pinned buffer: directly behind its predecessor's
                const size_t myOff = commit.reserve(subIndex, nOut);
                committed = true;
                {
                    std::shared_lock<std::shared_mutex> rd(ctx->pinnedMutex);
                    growPinned(rd, myOff + nOut);
                    fg_overlap* dst = pinned.p + myOff;
                    FG_CUDA(cudaMemcpyAsync(dst, dOut.p, nOut * sizeof(fg_overlap), cudaMemcpyDeviceToHost, streamOf(ctx)));
                    FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
                    if (prm.nucl_alignment) {   // overlap.cpp:463-468
                        PhaseTimer pe(ctx, "edit");
                        editDistances(ctx, dOut.p, dst, (uint32_t)nOut, prm.use_hpc != 0, !P.sameSet, prm.max_divergence, dQueryMaxDiv);
                        FG_CUDA(cudaMemcpyAsync(dst, dOut.p, nOut * sizeof(fg_overlap), cudaMemcpyDeviceToHost, streamOf(ctx)));
                        FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
                    }
                    // seqDivergence of this slice's records (host arithmetic with glibc logf), while the other lanes keep the device busy
                    HostTimer he(ctx, "host_divergence");
                    recordDivergences(ctx, dst, nOut, prm);
                    sliceEpilogue(ctx, dst, nOut, myOff, prm, commit, (size_t)qOffset + qa, nq);   // threshold / maxOverlaps replay of this slice's queries
                }
            }
        } else {   // no pair survived the prefilters: the hit sort's task counter is still checked (a truncated task list means unsorted ranges)
            uint32_t hSmall = 0;
            FG_CUDA(cudaMemcpyAsync(&hSmall, counters.p + 1, 4, cudaMemcpyDeviceToHost, streamOf(ctx)));
            FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
            if (hSmall > taskCap) throw Error(FG_ERR_INTERNAL, "sort task list overflow");
        }
        commitNone();
    };

    // lanes
    while ((int)ctx->lanes.size() < nLanes) {
        std::unique_ptr<Lane> l(new Lane());
        FG_CUDA(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking));
        // a lane that runs out of memory asks the context's own arena for its cached blocks (its owner thread is inside
        // fg_overlaps_batch, waiting for or working as a lane: nothing allocates from it meanwhile)
        l->arena.onPressure = [ctx] { std::lock_guard<std::mutex> lk(ctx->pressureMutex); ctx->arena.trim(); };
        ctx->lanes.push_back(std::move(l));
    }
    std::atomic<size_t> nextSub{0};
    std::atomic<bool> aborted{false};
    std::exception_ptr firstError; std::mutex errM;
    const int nWorkers = (int)std::min<size_t>(nLanes, std::max<size_t>(subs.size(), 1));
    auto worker = [&](int li) {
        Lane* lane = ctx->lanes[li].get();
        cudaSetDevice(ctx->device);
        currentArena() = &lane->arena; currentLane() = lane;
        try {
            for (;;) {
                const size_t i = nextSub.fetch_add(1);
                if (i >= subs.size() || aborted.load()) break;
                runSub(subs[i].qa, subs[i].qb, subBase + i);
            }
        } catch (...) {
            { std::lock_guard<std::mutex> lk(errM); if (!firstError) firstError = std::current_exception(); }
            aborted = true;
            commit.fail();
            cudaGetLastError();
        }
        currentLane() = nullptr; currentArena() = nullptr;
    };
    for (auto& l : ctx->lanes) { l->timings.clear(); l->timingCalls.clear(); }
    std::vector<std::thread> threads;
    for (int li = 1; li < nWorkers; ++li) threads.emplace_back(worker, li);
    {   // the calling thread is lane 0
        Arena* mainArena = currentArena();
        worker(0);
        currentArena() = mainArena;
    }
    for (auto& th : threads) th.join();
    subBase += subs.size();
    for (auto& l : ctx->lanes) {   // the lanes' phase times join the call's list
        for (size_t i = 0; i < l->timings.size(); ++i) addTiming(ctx, l->timings[i].first.c_str(), l->timings[i].second, l->timingCalls[i]);
        l->timings.clear(); l->timingCalls.clear();
    }
    if (firstError) std::rethrow_exception(firstError);
}

// seqDivergence of raw records (overlap.cpp:409-423; alignment.cpp:240-245 with nucl_alignment): the reference's float
// expression, evaluated on the host with glibc logf so that the bits agree.  Runs on the context's host pool.
static void recordDivergences(fg_ctx* ctx, fg_overlap* recs, size_t n, const fg_overlap_params& prm) {
    const float sampleRate = ctx->stats.sample_rate;
    const int k = ctx->k;
    auto body = [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
            fg_overlap& o = recs[i];
            const int32_t curRange = o.cur_end - o.cur_begin, extRange = o.ext_end - o.ext_begin;
            volatile float normLen = std::max(curRange, extRange) - o.filtered_positions;
            volatile float mr = (float)o.chain_length * sampleRate;
            volatile float matchRate = mr / normLen;
            matchRate = std::min((float)matchRate, 1.0f);
            volatile float inv = 1 / matchRate;
            volatile float lg = std::log((float)inv);
            o.seq_divergence = lg / k;
            if (prm.nucl_alignment) {   // alignment.cpp:240-245: (float)editDistance / max(len, len)
                if (o.edit_distance < 0) o.seq_divergence = 1.0f;
                else { volatile float dv = (float)o.edit_distance / (size_t)o.aln_len; o.seq_divergence = dv; }
            }
        }
    };
    if (n < 4096) { body(0, n); return; }
    std::lock_guard<std::mutex> lk(ctx->hostPoolMutex);
    ctx->hostPool.parallelFor(n, body, n);
}

// Per-query part of the host epilogue for one sub-batch's slice of the pinned buffer (whole queries, in query order): where every
// query's records start, then the replay of the divergence threshold (overlap.cpp:470) and of the maxOverlaps cut (:218-219),
// which only depends on the query's own records.  Dropped records are marked in `reserved`; the call's tail only has to turn the
// kept counts into offsets and compact if anything was dropped.  Runs on the host pool while the other lane keeps the device busy.
static void sliceEpilogue(fg_ctx* ctx, fg_overlap* recs, size_t n, size_t sliceOff, const fg_overlap_params& prm, OrderedCommit& commit,
                          size_t qFirst, size_t qCount) {
    if (!n) return;
    const size_t nQ = commit.kept.size();
    std::vector<size_t>& qStart = commit.qStart;
    std::lock_guard<std::mutex> lk(ctx->hostPoolMutex);
    // (1) first record of every query of the slice (queries qFirst ... qFirst + qCount - 1 of the call, in order)
    ctx->hostPool.parallelFor(n, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
            const fg_overlap& o = recs[i];
            if (i == 0 || recs[i - 1].reserved != o.reserved) {
                if (o.reserved >= nQ || o.reserved < qFirst || o.reserved >= qFirst + qCount || (i && recs[i - 1].reserved > o.reserved)) { commit.badOrder = true; continue; }
                qStart[o.reserved] = sliceOff + i;
            }
        }
    }, n);
    if (commit.badOrder) return;
    // end of every query's records = start of the next query that has any
    std::vector<size_t> qEndOf(qCount);
    size_t nextStart = sliceOff + n;
    for (size_t j = qCount; j-- > 0;) { qEndOf[j] = nextStart; if (qStart[qFirst + j] != SIZE_MAX) nextStart = qStart[qFirst + j]; }
    // (2) replay per query
    ctx->hostPool.parallelFor(qCount, [&](size_t ja, size_t jb) {
        for (size_t j = ja; j < jb; ++j) {
            const size_t q = qFirst + j;
            if (qStart[q] == SIZE_MAX) continue;
            size_t pos = qStart[q] - sliceOff, detected = 0;
            const size_t qEnd = qEndOf[j] - sliceOff;
            const float maxDiv = prm.query_max_divergence ? prm.query_max_divergence[q] : prm.max_divergence;
            while (pos < qEnd) {
                size_t end = pos;   // one target group = run of equal ext_id
                while (end < qEnd && recs[end].ext_id == recs[pos].ext_id) ++end;
                const bool stop = prm.max_overlaps != 0 && detected >= (size_t)prm.max_overlaps;
                for (size_t i = pos; i < end; ++i) {
                    const bool pass = recs[i].seq_divergence < maxDiv;
                    const bool keep = !stop && (pass || prm.keep_rejected);
                    recs[i].reserved = keep ? (pass ? 0u : 1u) : 0xffffffffu;
                    detected += keep;
                }
                pos = end;
            }
            commit.kept[q] = (uint32_t)detected;
        }
    }, n);
}

void overlapsBatch(fg_ctx* ctx, const uint32_t* queryIds, uint32_t nQ, const fg_overlap_params& prm, fg_overlap_result* result) {
    if (!ctx->indexed) throw Error(FG_ERR_ARG, "no index: call fg_build_index_* first");
    if (prm.keep_rejected && prm.max_overlaps != 0) throw Error(FG_ERR_ARG, "keep_rejected needs max_overlaps = 0");
    ctx->timings.clear(); ctx->timingCalls.clear();
    const int k = ctx->k;
    const bool sameSet = prm.query_set == 0;
    if (!sameSet && !ctx->dQsSeq.p) throw Error(FG_ERR_ARG, "query_set = 1 but fg_queries_upload has not been called");
    // the self-hit removal of the lookup / expansion identifies a query position with ONE index entry; a palindromic k-mer (its own
    // reverse complement: only possible for even k) would have two.  Every preset of the reference uses an odd k.
    if (sameSet && (k & 1) == 0) throw Error(FG_ERR_ARG, "the device overlap path needs an odd k-mer size (palindromic k-mers)");
    const uint32_t nQuerySeqs = sameSet ? ctx->nReads : ctx->nQsReads;
    for (uint32_t i = 0; i < nQ; ++i)
        if (queryIds[i] >= 2 * nQuerySeqs) throw Error(FG_ERR_ARG, "query id out of range");

    OvParams P;
    P.sameSet = sameSet;
    P.k = k; P.maxJump = prm.max_jump; P.minOverlap = prm.min_overlap; P.maxOverhang = prm.max_overhang;
    P.checkOverhang = prm.max_overhang > 0; P.forceLocal = prm.force_local != 0; P.onlyMaxExt = prm.only_max_ext != 0;
    { volatile float a = 0.01f; volatile float m = a * prm.min_overlap; P.minUniqueF = m; }   // minKmerSruvivalRate * _minOverlap

    ctx->resOffsets.assign(nQ + 1, 0);
    ctx->resAln.clear();
    uint64_t totHits = 0;
    BatchTotals tot;
    OrderedCommit commit;
    commit.qStart.assign((size_t)nQ + 1, SIZE_MAX); commit.kept.assign(nQ, 0);
    size_t subBase = 0;
    PinnedBuf<fg_overlap>& pinned = ctx->pinnedOut;   // all chunks / sub-batches land here in query order; the epilogue compacts into a second buffer
    HostTimer hostAll(ctx, "host_total");
    const uint64_t mallocs0 = ctx->arena.mallocCalls;

    if (prm.nucl_alignment) prepareEditDistances(ctx, prm.use_hpc != 0, !sameSet);
    DevBuf<float> dQueryMaxDiv;   // per-query divergence thresholds (optional), indexed like the records' `reserved`
    if (prm.query_max_divergence && nQ) {
        dQueryMaxDiv.alloc(nQ);
        FG_CUDA(cudaMemcpyAsync(dQueryMaxDiv.p, prm.query_max_divergence, nQ * sizeof(float), cudaMemcpyHostToDevice, streamOf(ctx)));
    }
    // divergence, threshold and per-query counts on the device unless the call needs the sequential per-query replay (maxOverlaps),
    // the rejected records (partitionBadMappings) or kmerMatches — those keep the host epilogue
    const bool devEpilogue = prm.max_overlaps == 0 && !prm.keep_rejected && !prm.keep_alignment && deviceEpilogueUsable(ctx);
    // chunks of consecutive queries with < 2^30 k-mer slots each (device arrays are indexed with 32-bit counts)
    uint64_t chunkSlots = 1ULL << 30;
    if (const char* e = getenv("FG_CHUNK_SLOTS")) chunkSlots = std::max<uint64_t>(4096, std::min<uint64_t>(chunkSlots, strtoull(e, nullptr, 10)));
    for (uint32_t q0 = 0; q0 < nQ;) {
        uint64_t slots = 0;
        uint32_t q1 = q0;
        while (q1 < nQ) {
            const uint32_t L = (sameSet ? ctx->hLen : ctx->hQsLen)[queryIds[q1] >> 1];
            const uint64_t add = L > (uint32_t)k ? ((uint64_t)(L - k) + 31u) & ~31ULL : 0;
            if (q1 > q0 && slots + add >= chunkSlots) break;
            slots += add; ++q1;
        }
        overlapsChunk(ctx, queryIds + q0, q1 - q0, q0, prm, P, commit, subBase, totHits, tot, dQueryMaxDiv.p, devEpilogue);
        q0 = q1;
    }
    const size_t nRaw = commit.nRaw;
    const uint64_t totPairs = tot.pairs, totDpPairs = tot.dpPairs, totCells = tot.cells, totTied = tot.tied, totPresorted = tot.presorted;

    // host epilogue: divergence (overlap.cpp:417-423), threshold (:470), maxOverlaps (:218-219)
    HostTimer hostEpi(ctx, "host_epilogue");
    fg_overlap* hOut = pinned.p;
    auto parallelFor = [&](size_t n, const std::function<void(size_t, size_t)>& fn) { ctx->hostPool.parallelFor(n, fn, nRaw); };
    // the lanes have computed the divergences (recordDivergences), found the first record of every query and replayed the threshold
    // and maxOverlaps per query (sliceEpilogue); dropped records carry reserved = ~0
    if (commit.badOrder) throw Error(FG_ERR_INTERNAL, "overlap records out of query order");
    std::vector<size_t>& qStart = commit.qStart;
    std::vector<uint32_t>& kept = commit.kept;
    qStart[nQ] = nRaw;
    for (uint32_t q = nQ; q-- > 0;) if (qStart[q] == SIZE_MAX) qStart[q] = qStart[q + 1];   // queries without records
    size_t wpos = 0;
    for (uint32_t q = 0; q < nQ; ++q) { ctx->resOffsets[q] = wpos; wpos += kept[q]; }
    ctx->resOffsets[nQ] = wpos;
    if (wpos != nRaw) {   // compact: the kept records of every query go to their final place in a second buffer, in parallel
        fg_overlap* dst = ctx->compactBuffer(0, wpos);
        parallelFor(nQ, [&](size_t qa, size_t qb) {
            for (size_t q = qa; q < qb; ++q) {
                size_t w2 = ctx->resOffsets[q];
                for (size_t i = qStart[q]; i < qStart[q + 1]; ++i)
                    if (hOut[i].reserved != 0xffffffffu) dst[w2++] = hOut[i];
            }
        });
        hOut = dst;
    }

    ctx->timings.emplace_back("arena_mallocs", (float)(ctx->arena.mallocCalls - mallocs0)); ctx->timingCalls.push_back(1);
    ctx->timings.emplace_back("arena_gib", (float)(ctx->arena.totalBytes / 1073741824.0)); ctx->timingCalls.push_back(1);
    ctx->timings.emplace_back("raw_overlaps", (float)nRaw); ctx->timingCalls.push_back(1);   // records copied device -> host (device epilogue: the kept ones)
    ctx->timings.emplace_back("gathered_overlaps", (float)tot.gathered.load()); ctx->timingCalls.push_back(1);   // primary overlaps before the divergence test (what the edit-distance kernel aligns)
    ctx->timings.emplace_back("tied_queries", (float)totTied); ctx->timingCalls.push_back(1);   // queries that needed the exact hit sort
    ctx->timings.emplace_back("presorted_pairs", (float)totPresorted); ctx->timingCalls.push_back(1);   // pairs whose score order needed no sort
    result->n_queries = nQ;
    result->offsets = ctx->resOffsets.data();
    result->overlaps = hOut;
    result->aln_pairs = ctx->resAln.empty() ? nullptr : ctx->resAln.data();
    result->n_aln_pairs = ctx->resAln.size() / 2;
    result->n_hits = totHits; result->n_pairs = totPairs; result->n_dp_pairs = totDpPairs; result->n_dp_cells = totCells;
    ctx->lastResult = *result;
    ctx->lastMaxOverlaps = prm.max_overlaps != 0 || prm.keep_rejected ? -1 : 0;
    ctx->lastMaxDivergence = prm.max_divergence;
    if (prm.query_max_divergence) ctx->lastQueryMaxDivergence.assign(prm.query_max_divergence, prm.query_max_divergence + nQ);
    else ctx->lastQueryMaxDivergence.clear();
}

void overlapsRefilter(fg_ctx* ctx, uint32_t firstQuery, float maxDivergence, fg_overlap_result* result) {
    fg_overlap_result& last = ctx->lastResult;
    const uint32_t nQ = last.n_queries;
    if (!last.offsets || firstQuery > nQ) throw Error(FG_ERR_ARG, "fg_overlaps_refilter: no previous result / first_query out of range");
    // The records at hand are those that survived the batch's own threshold and maxOverlaps replay (overlap.cpp:218-219, 470): a
    // tighter threshold can be applied to them afterwards, a looser one cannot (the dropped records are gone, and with
    // nucl_alignment their edit distance was never finished), and after a maxOverlaps cut the reference would have gone on
    // into later target groups.
    if (ctx->lastMaxOverlaps != 0) throw Error(FG_ERR_ARG, "fg_overlaps_refilter: the batch was cut with max_overlaps (or kept rejected records); run it again with the new threshold");
    for (uint32_t q = firstQuery; q < nQ; ++q) {
        const float used = ctx->lastQueryMaxDivergence.empty() ? ctx->lastMaxDivergence : ctx->lastQueryMaxDivergence[q];
        if (maxDivergence > used) throw Error(FG_ERR_ARG, "fg_overlaps_refilter: the new threshold is looser than the one the batch was computed with");
    }
    HostTimer ht(ctx, "host_refilter");
    const fg_overlap* src = last.overlaps;
    const size_t nRaw = ctx->resOffsets[nQ];
    std::vector<uint64_t> oldOff(ctx->resOffsets.begin(), ctx->resOffsets.begin() + nQ + 1);
    std::vector<uint32_t> kept(nQ, 0);
    ctx->hostPool.parallelFor(nQ, [&](size_t qa, size_t qb) {
        for (size_t q = qa; q < qb; ++q) {
            uint32_t c = 0;
            if (q < firstQuery) c = (uint32_t)(oldOff[q + 1] - oldOff[q]);
            else for (uint64_t i = oldOff[q]; i < oldOff[q + 1]; ++i) c += src[i].seq_divergence < maxDivergence;
            kept[q] = c;
        }
    }, nRaw);
    size_t wpos = 0;
    for (uint32_t q = 0; q < nQ; ++q) { ctx->resOffsets[q] = wpos; wpos += kept[q]; }
    ctx->resOffsets[nQ] = wpos;
    if (wpos != nRaw) {
        const int which = (src == ctx->resCompact[0].get()) ? 1 : 0;
        fg_overlap* dst = ctx->compactBuffer(which, wpos);
        ctx->hostPool.parallelFor(nQ, [&](size_t qa, size_t qb) {
            for (size_t q = qa; q < qb; ++q) {
                size_t w2 = ctx->resOffsets[q];
                for (uint64_t i = oldOff[q]; i < oldOff[q + 1]; ++i)
                    if (q < firstQuery || src[i].seq_divergence < maxDivergence) dst[w2++] = src[i];
            }
        }, nRaw);
        last.overlaps = dst;
    }
    last.offsets = ctx->resOffsets.data();
    *result = last;
    // from now on the queries from firstQuery on stand under the new threshold
    if (ctx->lastQueryMaxDivergence.empty()) ctx->lastQueryMaxDivergence.assign(nQ, ctx->lastMaxDivergence);
    for (uint32_t q = firstQuery; q < nQ; ++q) ctx->lastQueryMaxDivergence[q] = maxDivergence;
}

}  // namespace fg
