// Warp-parallel, permutation-exact emulation of libstdc++ 13 std::sort (introsort).
//
// Why: the reference sorts with std::sort in four places on this path (overlap.cpp:201-204, :272-274,
// :333-334, :432-434) and the ORDER OF TIES that libstdc++'s introsort happens to produce is
// result-relevant.  Bit-exact overlaps therefore need the same permutation, not just
// a sorted array.  This file produces that permutation with one warp per array:
//
//   * the recursion structure, depth limit (2*floor(lg n)), median-of-3 pivot selection, the
//     16-element threshold and the heap-sort fallback follow bits/stl_algo.h:1848-1951 literally
//     (small sequential pieces run on one lane);
//   * the Hoare partition (__unguarded_partition, :1871-1889) — where all the time goes — is restated
//     data-parallel.  With L_0<L_1<... the positions whose value is >= pivot (ascending) and
//     R_0>R_1>... those whose value is <= pivot (descending), the sequential two-pointer loop swaps
//     (L_m,R_m) for exactly the m with L_m < R_m (a prefix, s pairs) and returns
//         cut = s==0 ? L_0 : min(L_s, R_{s-1}).
//     The warp streams 32-element chunks from both ends, finds stops with ballots and lets lane m move
//     pair m; when the chunks meet, the remaining pairs live inside the last chunk and are resolved from
//     the ballot masks of its ORIGINAL values (kept in registers).
//   * the final insertion sort (__final_insertion_sort, :1854-1865) is a stable sort that never moves an
//     element across a partition cut, i.e. an independent stable insertion sort of every <=16-element
//     leaf: one lane per leaf.
//
// The same source compiles for the host (FG_WARP_HOST) with the 32 lanes run as loops; that build is
// checked against std::sort itself in tests/test_introsort_model.py, so the logic below is validated on
// the CPU and the device build is checked against the CPU build on the GPU.
#pragma once
#include <cstdint>

#ifndef FG_WARP_HOST
#include "common.cuh"
#define FG_DEV __device__ __forceinline__
#define FG_FOR_LANES { const int lane = fg::laneId();
#define FG_END_LANES }
#define FG_LANEVAR(T, name) T name
#define FG_L(name) name
#define FG_BALLOT(out, pred) { const int lane = fg::laneId(); (void)lane; out = __ballot_sync(0xffffffffu, (pred)); }
#define FG_SYNCWARP() __syncwarp()
#define FG_POPC(x) __popc(x)
#define FG_CTZ(x) (__ffs(x) - 1)
#define FG_POPCLL(x) __popcll(x)
#define FG_CTZLL(x) (__ffsll((long long)(x)) - 1)
#define FG_PREFETCH_L2(ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr))
#define FG_REDUCE_MAX(out, expr) { const int lane = fg::laneId(); (void)lane; out = __reduce_max_sync(0xffffffffu, (int)(expr)); }
#define FG_CLZ(x) __clz((int)(x))
#define FG_ATOMIC_OR(ptr, v) atomicOr((ptr), (v))
#else
#include <algorithm>
namespace fg {
inline int nthLowBit(uint32_t mask, int n) {
    for (int b = 0; b < 32; ++b) if (mask >> b & 1) { if (n == 0) return b; --n; }
    return 32;
}
inline int nthHighBit(uint32_t mask, int n) {
    for (int b = 31; b >= 0; --b) if (mask >> b & 1) { if (n == 0) return b; --n; }
    return -1;
}
}
#define FG_DEV inline
#define FG_FOR_LANES for (int lane = 0; lane < 32; ++lane) {
#define FG_END_LANES }
#define FG_LANEVAR(T, name) T name[32]
#define FG_L(name) name[lane]
#define FG_BALLOT(out, pred) { out = 0; for (int lane = 0; lane < 32; ++lane) if (pred) out |= (1u << lane); }
#define FG_SYNCWARP() ((void)0)
#define FG_POPC(x) __builtin_popcount(x)
#define FG_CTZ(x) __builtin_ctz(x)
#define FG_POPCLL(x) __builtin_popcountll(x)
#define FG_CTZLL(x) __builtin_ctzll(x)
#define FG_PREFETCH_L2(ptr) ((void)0)
#define FG_REDUCE_MAX(out, expr) { out = INT32_MIN; for (int lane = 0; lane < 32; ++lane) { const int v_ = (int)(expr); if (v_ > out) out = v_; } }
#define FG_CLZ(x) ((x) ? __builtin_clz((unsigned)(x)) : 32)
#define FG_ATOMIC_OR(ptr, v) (*(ptr) |= (v))
#endif

namespace fg {

#ifdef FG_WARP_HOST
static long g_heapSortCalls = 0;   // test visibility: how often the depth-limit fallback ran
#endif

// 16-byte sort element: only `key` is compared (ascending); val/aux ride along.
struct
#ifndef FG_WARP_HOST
    __align__(16)
#endif
    Elem {
    unsigned long long key;
    unsigned int val;
    unsigned int aux;
};

FG_DEV bool elemLess(const Elem& a, const Elem& b) { return a.key < b.key; }
FG_DEV void elemSwap(Elem* a, long i, long j) { Elem t = a[i]; a[i] = a[j]; a[j] = t; }

// ---- sequential pieces (one lane) -----------------------------------------------------------------
// __move_median_to_first(result, a, b, c), stl_algo.h:85-111
FG_DEV void seqMedianToFirst(Elem* arr, long result, long a, long b, long c) {
    unsigned long long ka = arr[a].key, kb = arr[b].key, kc = arr[c].key;
    long pick;
    if (ka < kb) { if (kb < kc) pick = b; else if (ka < kc) pick = c; else pick = a; }
    else if (ka < kc) pick = a;
    else if (kb < kc) pick = c;
    else pick = b;
    elemSwap(arr, result, pick);
}

// stable insertion sort of a leaf (the part of __final_insertion_sort that touches this range)
FG_DEV void seqInsertionSort(Elem* arr, long f, long l) {
    for (long i = f + 1; i < l; ++i) {
        Elem v = arr[i];
        long j = i;
        while (j > f && v.key < arr[j - 1].key) { arr[j] = arr[j - 1]; --j; }
        arr[j] = v;
    }
}

// __partial_sort(first,last,last) = __heap_select + __sort_heap  (stl_algo.h:1905-1912, stl_heap.h)
FG_DEV void seqAdjustHeap(Elem* first, long hole, long len, Elem value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (first[child].key < first[child - 1].key) --child;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    long parent = (hole - 1) / 2;   // __push_heap
    while (hole > top && first[parent].key < value.key) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
FG_DEV void seqHeapSort(Elem* first, long len) {
    if (len >= 2)
        for (long parent = (len - 2) / 2;; --parent) {
            Elem v = first[parent];
            seqAdjustHeap(first, parent, len, v);
            if (parent == 0) break;
        }
    for (long last = len; last > 1;) {
        --last;
        Elem v = first[last];
        first[last] = first[0];
        seqAdjustHeap(first, 0, last, v);
    }
}

// literal sequential introsort (one thread) for the rare small arrays sorted inside per-thread code
FG_DEV long seqUnguardedPartition(Elem* a, long first, long last, long pivot) {   // stl_algo.h:1871-1889
    const unsigned long long p = a[pivot].key;
    for (;;) {
        while (a[first].key < p) ++first;
        --last;
        while (p < a[last].key) --last;
        if (!(first < last)) return first;
        elemSwap(a, first, last);
        ++first;
    }
}
// literal __introsort_loop + leaf insertion sorts on a[f,l) with an INHERITED depth budget (one thread).  STACK bounds the
// number of pending right halves (> 16 elements each, pairwise disjoint): (l-f)/17 at most.
template <int STACK>
FG_DEV void seqIntrosortRange(Elem* a, long f0, long l0, int d0) {
    if (l0 - f0 < 2) return;
    long stF[STACK], stL[STACK]; int stD[STACK];
    int sp = 0;
    long f = f0, l = l0; int d = d0;
    for (;;) {
        while (l - f > 16) {
            if (d == 0) { seqHeapSort(a + f, l - f); break; }
            --d;
            seqMedianToFirst(a, f, f + 1, f + (l - f) / 2, l - 1);
            const long cut = seqUnguardedPartition(a, f + 1, l, f);
            if (l - cut > 16) { stF[sp] = cut; stL[sp] = l; stD[sp] = d; ++sp; }
            l = cut;
        }
        if (sp == 0) break;
        --sp; f = stF[sp]; l = stL[sp]; d = stD[sp];
    }
    seqInsertionSort(a, f0, l0);   // stable; never moves an element across a cut
}

FG_DEV void seqIntrosort(Elem* a, long n) {
    if (n < 2) return;
    if (n > 16) {
        long stF[64], stL[64]; int stD[64];
        int sp = 0, lg = 0;
        while ((n >> (lg + 1)) != 0) ++lg;
        long f = 0, l = n; int d = 2 * lg;
        for (;;) {
            while (l - f > 16) {
                if (d == 0) { seqHeapSort(a + f, l - f); break; }
                --d;
                seqMedianToFirst(a, f, f + 1, f + (l - f) / 2, l - 1);
                const long cut = seqUnguardedPartition(a, f + 1, l, f);
                if (l - cut > 16) { stF[sp] = cut; stL[sp] = l; stD[sp] = d; ++sp; }
                l = cut;
            }
            if (sp == 0) break;
            --sp; f = stF[sp]; l = stL[sp]; d = stD[sp];
        }
    }
    seqInsertionSort(a, 0, n);   // __final_insertion_sort == stable insertion sort of the whole array
}

// ---- the data-parallel Hoare partition ---------------------------------------------------------------
// Partitions arr[f,l) (l-f > 16) exactly like
//     __move_median_to_first(f, f+1, f+(l-f)/2, l-1); return __unguarded_partition(f+1, l, f);
// and returns the cut.  All control flow is warp-uniform.  `tab` is 64 bytes of per-warp scratch (shared memory on
// the device): rank -> lane tables that let lane m find the partner of its stop without searching bit masks.
typedef int idx_t;   // positions inside one sorted array (< 2^31 elements)

// `inGlobal`: arr lives in global memory — lines a few chunks ahead of both cursors are pulled into L2 early.
static constexpr idx_t PF_AHEAD = 4 * 32;
static constexpr idx_t PF_AHEAD64 = 4 * 64;   // elements (4 chunks = 2 KB) between the register prefetch and the L2 prefetch

FG_DEV idx_t warpPartition(Elem* arr, idx_t f, idx_t l, unsigned char* tab, bool inGlobal = false) {
    unsigned char* tabL = tab;        // tabL[m] = lane of the m-th lowest pending ">= pivot" stop
    unsigned char* tabR = tab + 32;   // tabR[m] = lane of the m-th highest pending "<= pivot" stop
#ifndef FG_WARP_HOST
    if (fg::laneId() == 0)
#endif
        seqMedianToFirst(arr, f, f + 1, f + (l - f) / 2, l - 1);
    FG_SYNCWARP();
    const unsigned long long p = arr[f].key;

    idx_t lo = f + 1, hi = l;        // untouched middle [lo,hi)
    idx_t s = 0;                     // pairs swapped so far
    idx_t lastR = -1;                // position of R_{s-1}
    idx_t firstGeAbove = l;          // lowest position with original value >= p in retired right chunks
    idx_t Lb = 0, Rb = 0;            // base position of the current left / right chunk
    uint32_t geL = 0, leL = 0, pendL = 0, geR = 0, leR = 0, pendR = 0;
    bool haveR = false;
    FG_LANEVAR(Elem, eL);
    FG_LANEVAR(Elem, eR);
    // Software prefetch: the next full chunk of each side is loaded one refill ahead.  Positions of the
    // untouched middle are never written, so a prefetched chunk stays valid as long as the other side has
    // not claimed any of its 32 positions (hi - lo >= 32 at the time it is used).
    FG_LANEVAR(Elem, pfL);
    FG_LANEVAR(Elem, pfR);
    bool pfLok = false, pfRok = false;
    if (hi - lo >= 64) {
        FG_FOR_LANES FG_L(pfL) = arr[lo + lane]; FG_L(pfR) = arr[hi - 32 + lane]; FG_END_LANES
        pfLok = pfRok = true;
    }
    if (inGlobal) {
        FG_FOR_LANES
            for (idx_t a = 32; a <= PF_AHEAD; a += 32) {
                if (lo + a + lane < hi) FG_PREFETCH_L2(arr + lo + a + lane);
                if (hi - 32 - a + lane >= lo) FG_PREFETCH_L2(arr + hi - 32 - a + lane);
            }
        FG_END_LANES
    }

    for (;;) {
        if (pendL == 0 && lo < hi) {
            Lb = lo;
            const idx_t nL = (hi - lo < 32) ? (hi - lo) : 32;
            if (pfLok && nL == 32) { FG_FOR_LANES FG_L(eL) = FG_L(pfL); FG_END_LANES }
            else { FG_FOR_LANES if (lane < nL) FG_L(eL) = arr[Lb + lane]; FG_END_LANES }
            lo += nL;
            pfLok = hi - lo >= 32;
            if (pfLok) { FG_FOR_LANES FG_L(pfL) = arr[lo + lane]; FG_END_LANES }
            if (inGlobal) { FG_FOR_LANES if (lo + PF_AHEAD + lane < hi) FG_PREFETCH_L2(arr + lo + PF_AHEAD + lane); FG_END_LANES }
            FG_BALLOT(geL, lane < nL && FG_L(eL).key >= p);
            FG_BALLOT(leL, lane < nL && FG_L(eL).key <= p);
            pendL = geL;
        }
        if (pendR == 0 && lo < hi) {
            if (haveR && geR) firstGeAbove = Rb + FG_CTZ(geR);
            const idx_t nR = (hi - lo < 32) ? (hi - lo) : 32;
            Rb = hi - nR;
            if (pfRok && nR == 32) { FG_FOR_LANES FG_L(eR) = FG_L(pfR); FG_END_LANES }
            else { FG_FOR_LANES if (lane < nR) FG_L(eR) = arr[Rb + lane]; FG_END_LANES }
            hi = Rb;
            pfRok = hi - lo >= 32;
            if (pfRok) { FG_FOR_LANES FG_L(pfR) = arr[hi - 32 + lane]; FG_END_LANES }
            if (inGlobal) { FG_FOR_LANES if (hi - 32 - PF_AHEAD + lane >= lo) FG_PREFETCH_L2(arr + hi - 32 - PF_AHEAD + lane); FG_END_LANES }
            FG_BALLOT(geR, lane < nR && FG_L(eR).key >= p);
            FG_BALLOT(leR, lane < nR && FG_L(eR).key <= p);
            pendR = leR;
            haveR = true;
        }
        const int cL = FG_POPC(pendL), cR = FG_POPC(pendR);
        const int c = cL < cR ? cL : cR;
        if (c > 0) {
            // pair m: m-th lowest pending L  <->  m-th highest pending R   (every pending L < every pending R)
            FG_FOR_LANES
                if (pendL >> lane & 1) tabL[FG_POPC(pendL & ((1u << lane) - 1u))] = (unsigned char)lane;
                if (pendR >> lane & 1) tabR[FG_POPC(pendR & ~((2u << lane) - 1u))] = (unsigned char)lane;
            FG_END_LANES
            FG_SYNCWARP();
            FG_FOR_LANES
                if (pendL >> lane & 1) { const int m = FG_POPC(pendL & ((1u << lane) - 1u)); if (m < c) arr[Rb + tabR[m]] = FG_L(eL); }
                if (pendR >> lane & 1) { const int m = FG_POPC(pendR & ~((2u << lane) - 1u)); if (m < c) arr[Lb + tabL[m]] = FG_L(eR); }
            FG_END_LANES
            lastR = Rb + tabR[c - 1];
            // drop the c lowest stops of pendL and the c highest of pendR
            uint32_t keepL, keepR;
            FG_BALLOT(keepL, (pendL >> lane & 1) && FG_POPC(pendL & ((1u << lane) - 1u)) >= c);
            FG_BALLOT(keepR, (pendR >> lane & 1) && FG_POPC(pendR & ~((2u << lane) - 1u)) >= c);
            pendL = keepL; pendR = keepR;
            s += c;
            FG_SYNCWARP();   // the tables are rewritten by the next round
        }
        if (lo < hi && (pendL == 0 || pendR == 0)) continue;
        break;
    }
    FG_SYNCWARP();

    // The chunks have met (lo == hi).  Remaining pairs, if any, lie inside ONE chunk.
    const idx_t INF = l + 1;
    idx_t Ls;   // position of L_s, the first unpaired ">= pivot" stop in the original ordering
    if (pendL != 0 || pendR != 0) {
        // case A (pendL): unpaired L stops remain in the current left chunk; the next R stops are that chunk's
        //   "<= pivot" positions, descending.
        // case B (pendR): unpaired R stops remain in the current right chunk; the next L stops are that chunk's
        //   ">= pivot" positions, ascending, then those of the chunks above it.
        const bool caseA = pendL != 0;
        const uint32_t maskL = caseA ? pendL : geR, maskR = caseA ? leL : pendR;
        const idx_t base = caseA ? Lb : Rb;
        const int nl = FG_POPC(maskL), nr = FG_POPC(maskR);
        FG_FOR_LANES
            if (maskL >> lane & 1) tabL[FG_POPC(maskL & ((1u << lane) - 1u))] = (unsigned char)lane;
            if (maskR >> lane & 1) tabR[FG_POPC(maskR & ~((2u << lane) - 1u))] = (unsigned char)lane;
        FG_END_LANES
        FG_SYNCWARP();
        uint32_t ok;
        FG_BALLOT(ok, lane < nl && lane < nr && tabL[lane] < tabR[lane]);
        const int e = FG_POPC(ok);
        FG_FOR_LANES
            const Elem mine = caseA ? FG_L(eL) : FG_L(eR);
            if (maskL >> lane & 1) { const int m = FG_POPC(maskL & ((1u << lane) - 1u)); if (m < e) arr[base + tabR[m]] = mine; }
            if (maskR >> lane & 1) { const int m = FG_POPC(maskR & ~((2u << lane) - 1u)); if (m < e) arr[base + tabL[m]] = mine; }
        FG_END_LANES
        s += e;
        if (e > 0) lastR = base + tabR[e - 1];
        Ls = (e < nl) ? base + tabL[e] : (caseA ? INF : firstGeAbove);
    } else {
        Ls = (haveR && geR) ? Rb + FG_CTZ(geR) : firstGeAbove;
    }
    FG_SYNCWARP();
    if (s == 0) return Ls;
    return Ls < lastR ? Ls : lastR;
}

// ---- the same partition with 64-position chunks (two elements per lane) --------------------------------------
// Identical logic and result; every refill moves 1 KB per side, so the fixed cost of a round (ballots, rank tables,
// mask updates) is spread over twice as many elements and twice as many bytes are in flight per warp.  Slot s of a
// chunk (position base + s) lives in lane s%32, register s/32.  `tab` needs 128 bytes here.
FG_DEV int rankBelow64(unsigned long long m, int s) { return FG_POPCLL(m & ((1ULL << s) - 1ULL)); }
FG_DEV int rankAbove64(unsigned long long m, int s) { return s == 63 ? 0 : FG_POPCLL(m >> (s + 1)); }

FG_DEV idx_t warpPartition64(Elem* arr, idx_t f, idx_t l, unsigned char* tab, bool inGlobal = false) {
    unsigned char* tabL = tab;        // tabL[m] = slot of the m-th lowest pending ">= pivot" stop
    unsigned char* tabR = tab + 64;   // tabR[m] = slot of the m-th highest pending "<= pivot" stop
#ifndef FG_WARP_HOST
    if (fg::laneId() == 0)
#endif
        seqMedianToFirst(arr, f, f + 1, f + (l - f) / 2, l - 1);
    FG_SYNCWARP();
    const unsigned long long p = arr[f].key;

    idx_t lo = f + 1, hi = l;        // untouched middle [lo,hi)
    idx_t s = 0, lastR = -1, firstGeAbove = l, Lb = 0, Rb = 0;
    unsigned long long geL = 0, leL = 0, pendL = 0, geR = 0, leR = 0, pendR = 0;
    bool haveR = false;
    FG_LANEVAR(Elem, eL0); FG_LANEVAR(Elem, eL1); FG_LANEVAR(Elem, eR0); FG_LANEVAR(Elem, eR1);
    FG_LANEVAR(Elem, pL0); FG_LANEVAR(Elem, pL1); FG_LANEVAR(Elem, pR0); FG_LANEVAR(Elem, pR1);   // register prefetch
    bool pfLok = false, pfRok = false;
    if (hi - lo >= 128) {
        FG_FOR_LANES
            FG_L(pL0) = arr[lo + lane]; FG_L(pL1) = arr[lo + 32 + lane];
            FG_L(pR0) = arr[hi - 64 + lane]; FG_L(pR1) = arr[hi - 32 + lane];
        FG_END_LANES
        pfLok = pfRok = true;
    }
    if (inGlobal) {
        FG_FOR_LANES
            for (idx_t a = 64; a <= PF_AHEAD64; a += 32) {
                if (lo + a + lane < hi) FG_PREFETCH_L2(arr + lo + a + lane);
                if (hi - 32 - a + lane >= lo) FG_PREFETCH_L2(arr + hi - 32 - a + lane);
            }
        FG_END_LANES
    }

    for (;;) {
        if (pendL == 0 && lo < hi) {
            Lb = lo;
            const idx_t nL = (hi - lo < 64) ? (hi - lo) : 64;
            if (pfLok && nL == 64) { FG_FOR_LANES FG_L(eL0) = FG_L(pL0); FG_L(eL1) = FG_L(pL1); FG_END_LANES }
            else { FG_FOR_LANES if (lane < nL) FG_L(eL0) = arr[Lb + lane]; if (32 + lane < nL) FG_L(eL1) = arr[Lb + 32 + lane]; FG_END_LANES }
            lo += nL;
            pfLok = hi - lo >= 64;
            if (pfLok) { FG_FOR_LANES FG_L(pL0) = arr[lo + lane]; FG_L(pL1) = arr[lo + 32 + lane]; FG_END_LANES }
            if (inGlobal) {
                FG_FOR_LANES
                    if (lo + PF_AHEAD64 + lane < hi) FG_PREFETCH_L2(arr + lo + PF_AHEAD64 + lane);
                    if (lo + PF_AHEAD64 + 32 + lane < hi) FG_PREFETCH_L2(arr + lo + PF_AHEAD64 + 32 + lane);
                FG_END_LANES
            }
            uint32_t a0, a1, b0, b1;
            FG_BALLOT(a0, lane < nL && FG_L(eL0).key >= p); FG_BALLOT(a1, 32 + lane < nL && FG_L(eL1).key >= p);
            FG_BALLOT(b0, lane < nL && FG_L(eL0).key <= p); FG_BALLOT(b1, 32 + lane < nL && FG_L(eL1).key <= p);
            geL = a0 | ((unsigned long long)a1 << 32); leL = b0 | ((unsigned long long)b1 << 32);
            pendL = geL;
        }
        if (pendR == 0 && lo < hi) {
            if (haveR && geR) firstGeAbove = Rb + FG_CTZLL(geR);
            const idx_t nR = (hi - lo < 64) ? (hi - lo) : 64;
            Rb = hi - nR;
            if (pfRok && nR == 64) { FG_FOR_LANES FG_L(eR0) = FG_L(pR0); FG_L(eR1) = FG_L(pR1); FG_END_LANES }
            else { FG_FOR_LANES if (lane < nR) FG_L(eR0) = arr[Rb + lane]; if (32 + lane < nR) FG_L(eR1) = arr[Rb + 32 + lane]; FG_END_LANES }
            hi = Rb;
            pfRok = hi - lo >= 64;
            if (pfRok) { FG_FOR_LANES FG_L(pR0) = arr[hi - 64 + lane]; FG_L(pR1) = arr[hi - 32 + lane]; FG_END_LANES }
            if (inGlobal) {
                FG_FOR_LANES
                    if (hi - 32 - PF_AHEAD64 + lane >= lo) FG_PREFETCH_L2(arr + hi - 32 - PF_AHEAD64 + lane);
                    if (hi - 64 - PF_AHEAD64 + lane >= lo) FG_PREFETCH_L2(arr + hi - 64 - PF_AHEAD64 + lane);
                FG_END_LANES
            }
            uint32_t a0, a1, b0, b1;
            FG_BALLOT(a0, lane < nR && FG_L(eR0).key >= p); FG_BALLOT(a1, 32 + lane < nR && FG_L(eR1).key >= p);
            FG_BALLOT(b0, lane < nR && FG_L(eR0).key <= p); FG_BALLOT(b1, 32 + lane < nR && FG_L(eR1).key <= p);
            geR = a0 | ((unsigned long long)a1 << 32); leR = b0 | ((unsigned long long)b1 << 32);
            pendR = leR;
            haveR = true;
        }
        const int cL = FG_POPCLL(pendL), cR = FG_POPCLL(pendR);
        const int c = cL < cR ? cL : cR;
        if (c > 0) {
            FG_FOR_LANES
                for (int r = 0; r < 2; ++r) {
                    const int sl = lane + 32 * r;
                    if (pendL >> sl & 1) tabL[rankBelow64(pendL, sl)] = (unsigned char)sl;
                    if (pendR >> sl & 1) tabR[rankAbove64(pendR, sl)] = (unsigned char)sl;
                }
            FG_END_LANES
            FG_SYNCWARP();
            FG_FOR_LANES
                for (int r = 0; r < 2; ++r) {
                    const int sl = lane + 32 * r;
                    if (pendL >> sl & 1) { const int m = rankBelow64(pendL, sl); if (m < c) arr[Rb + tabR[m]] = r ? FG_L(eL1) : FG_L(eL0); }
                    if (pendR >> sl & 1) { const int m = rankAbove64(pendR, sl); if (m < c) arr[Lb + tabL[m]] = r ? FG_L(eR1) : FG_L(eR0); }
                }
            FG_END_LANES
            lastR = Rb + tabR[c - 1];
            uint32_t k0, k1, q0, q1;
            FG_BALLOT(k0, (pendL >> lane & 1) && rankBelow64(pendL, lane) >= c);
            FG_BALLOT(k1, (pendL >> (lane + 32) & 1) && rankBelow64(pendL, lane + 32) >= c);
            FG_BALLOT(q0, (pendR >> lane & 1) && rankAbove64(pendR, lane) >= c);
            FG_BALLOT(q1, (pendR >> (lane + 32) & 1) && rankAbove64(pendR, lane + 32) >= c);
            pendL = k0 | ((unsigned long long)k1 << 32); pendR = q0 | ((unsigned long long)q1 << 32);
            s += c;
            FG_SYNCWARP();
        }
        if (lo < hi && (pendL == 0 || pendR == 0)) continue;
        break;
    }
    FG_SYNCWARP();

    const idx_t INF = l + 1;
    idx_t Ls;
    if (pendL != 0 || pendR != 0) {
        const bool caseA = pendL != 0;
        const unsigned long long maskL = caseA ? pendL : geR, maskR = caseA ? leL : pendR;
        const idx_t base = caseA ? Lb : Rb;
        const int nl = FG_POPCLL(maskL), nr = FG_POPCLL(maskR);
        FG_FOR_LANES
            for (int r = 0; r < 2; ++r) {
                const int sl = lane + 32 * r;
                if (maskL >> sl & 1) tabL[rankBelow64(maskL, sl)] = (unsigned char)sl;
                if (maskR >> sl & 1) tabR[rankAbove64(maskR, sl)] = (unsigned char)sl;
            }
        FG_END_LANES
        FG_SYNCWARP();
        uint32_t ok0, ok1;
        FG_BALLOT(ok0, lane < nl && lane < nr && tabL[lane] < tabR[lane]);
        FG_BALLOT(ok1, lane + 32 < nl && lane + 32 < nr && tabL[lane + 32] < tabR[lane + 32]);
        const int e = FG_POPC(ok0) + FG_POPC(ok1);
        FG_FOR_LANES
            for (int r = 0; r < 2; ++r) {
                const int sl = lane + 32 * r;
                const Elem mine = caseA ? (r ? FG_L(eL1) : FG_L(eL0)) : (r ? FG_L(eR1) : FG_L(eR0));
                if (maskL >> sl & 1) { const int m = rankBelow64(maskL, sl); if (m < e) arr[base + tabR[m]] = mine; }
                if (maskR >> sl & 1) { const int m = rankAbove64(maskR, sl); if (m < e) arr[base + tabL[m]] = mine; }
            }
        FG_END_LANES
        s += e;
        if (e > 0) lastR = base + tabR[e - 1];
        Ls = (e < nl) ? base + tabL[e] : (caseA ? INF : firstGeAbove);
    } else {
        Ls = (haveR && geR) ? Rb + FG_CTZLL(geR) : firstGeAbove;
    }
    FG_SYNCWARP();
    if (s == 0) return Ls;
    return Ls < lastR ? Ls : lastR;
}

// ---- the same partition for a range staged in SHARED memory: rank tables instead of streaming ---------------------
// In shared memory there is nothing to stream: one pass over the range records the positions of all L stops (key >=
// pivot) and all R stops (key <= pivot) in two rank tables (ascending position; R_m is entry nR-1-m), then lane m swaps
// pair m for the prefix of pairs with L_m < R_m.  Same result as warpPartition with ~5x fewer instructions per element
// (one load, two ballots and two rank stores per 32 elements; no pending masks, no per-round bookkeeping).
// tL / tR: l - f - 1 entries each; positions are offsets into arr and must fit 16 bits.
FG_DEV idx_t warpPartitionTable(Elem* arr, idx_t f, idx_t l, unsigned short* tL, unsigned short* tR) {
#ifndef FG_WARP_HOST
    if (fg::laneId() == 0)
#endif
        seqMedianToFirst(arr, f, f + 1, f + (l - f) / 2, l - 1);
    FG_SYNCWARP();
    const unsigned long long p = arr[f].key;
    int nL = 0, nR = 0;
    for (idx_t b = f + 1; b < l; b += 32) {
        uint32_t g, e;
        FG_BALLOT(g, b + lane < l && arr[b + lane].key >= p);
        FG_BALLOT(e, b + lane < l && arr[b + lane].key <= p);
        FG_FOR_LANES
            const uint32_t below = (1u << lane) - 1u;
            if (g >> lane & 1) tL[nL + FG_POPC(g & below)] = (unsigned short)(b + lane);
            if (e >> lane & 1) tR[nR + FG_POPC(e & below)] = (unsigned short)(b + lane);
        FG_END_LANES
        nL += FG_POPC(g); nR += FG_POPC(e);
    }
    FG_SYNCWARP();
    const int lim = nL < nR ? nL : nR;
    int s = 0;
    for (int m0 = 0; m0 < lim; m0 += 32) {
        uint32_t ok;
        FG_BALLOT(ok, m0 + lane < lim && tL[m0 + lane] < tR[nR - 1 - m0 - lane]);
        FG_FOR_LANES if (ok >> lane & 1) elemSwap(arr, tL[m0 + lane], tR[nR - 1 - m0 - lane]); FG_END_LANES
        s += FG_POPC(ok);
        if (ok != 0xffffffffu) break;
    }
    FG_SYNCWARP();
    const idx_t Ls = s < nL ? (idx_t)tL[s] : l + 1;
    if (s == 0) return Ls;
    const idx_t lastR = tR[nR - s];
    return Ls < lastR ? Ls : lastR;
}

#ifndef FG_CHUNK64
#define FG_CHUNK64 0   // measured on B200: the 64-wide variant needs 107-124 registers and loses more to occupancy than it gains
#endif
FG_DEV idx_t warpPartitionAny(Elem* arr, idx_t f, idx_t l, unsigned char* tab, bool inGlobal) {
#if FG_CHUNK64
    return warpPartition64(arr, f, l, tab, inGlobal);
#else
    return warpPartition(arr, f, l, tab, inGlobal);
#endif
}
static constexpr int WARP_TAB_BYTES = 128;   // per-warp scratch of the partition

// ---- the whole sort -------------------------------------------------------------------------------------
// The introsort loop (__introsort_loop, stl_algo.h:1918-1940) on arr[f0,l0) with depth budget d0, followed by the
// leaf insertion sorts.  One warp; no shared memory: the explicit recursion stack (<= 2*lg n entries) and the
// pending-leaf list live in lane registers.
//
// Two-level use: with small > 0, every range of at most `small` elements is NOT processed here but handed to
// `sink(f, l, depthBudget)`; a second kernel finishes those ranges in shared memory (warpIntrosortRange with
// small = 0 on the staged copy).  Ranges are disjoint, so the order in which they are finished is irrelevant
// to the resulting permutation.
struct NoSink { FG_DEV void operator()(long, long, int) const {} };
#ifndef FG_WARP_MINI
#define FG_WARP_MINI 32
#endif
static constexpr int WARP_MINI = FG_WARP_MINI;

// TABLE: arr is in shared memory and `tab` holds two rank tables of `tabCap` 16-bit entries each (warpPartitionTable).
template <bool TABLE = false, class Sink>
FG_DEV void warpIntrosortRange(Elem* arr, idx_t f0, idx_t l0, int d0, idx_t small, Sink& sink, unsigned char* tab, bool inGlobal = false,
                               int tabCap = 0) {
    if (l0 - f0 < 2) return;
    FG_LANEVAR(idx_t, stF0); FG_LANEVAR(idx_t, stL0); FG_LANEVAR(int, stD0);   // stack entries 0..31
    FG_LANEVAR(idx_t, stF1); FG_LANEVAR(idx_t, stL1); FG_LANEVAR(int, stD1);   // stack entries 32..63
    FG_LANEVAR(idx_t, lfF);  FG_LANEVAR(idx_t, lfL); FG_LANEVAR(int, lfD);    // pending mini ranges
    int sp = 0, nLeaf = 0;
    FG_FOR_LANES FG_L(lfF) = 0; FG_L(lfL) = 0; FG_L(lfD) = 0; FG_END_LANES

    // Ranges of at most WARP_MINI elements are not worth a warp-wide partition (most lanes would idle): they are
    // queued and finished 32 at a time, one LANE per range, by the literal sequential algorithm with the range's
    // inherited depth budget.
    auto flushLeaves = [&]() {
        FG_SYNCWARP();
        FG_FOR_LANES if (lane < nLeaf) seqIntrosortRange<WARP_MINI / 17 + 1>(arr, FG_L(lfF), FG_L(lfL), FG_L(lfD)); FG_END_LANES
        FG_SYNCWARP();
        nLeaf = 0;
    };
    // a finished-partitioning range: leaf (<= 16) or, in two-level mode, a task for the second kernel
    auto retire = [&](idx_t f, idx_t l, int d) {
        if (l - f < 1) return;
        if (small > 0) { sink(f, l, d); return; }   // single elements too: the tie-following sink has to deliver them
        if (l - f < 2) return;
        FG_FOR_LANES if (lane == nLeaf) { FG_L(lfF) = f; FG_L(lfL) = l; FG_L(lfD) = d; } FG_END_LANES
        if (++nLeaf == 32) flushLeaves();
    };
    const idx_t stopAt = small > WARP_MINI ? small : WARP_MINI;   // ranges of at most this many elements are retired

    idx_t f = f0, l = l0;
    int d = d0;
    bool have = true;
    while (have) {
        bool heapSorted = false;
        while (l - f > stopAt) {
            if (d == 0) {
#ifndef FG_WARP_HOST
                if (fg::laneId() == 0)
#endif
                    seqHeapSort(arr + f, l - f);
#ifdef FG_WARP_HOST
                ++g_heapSortCalls;
#endif
                FG_SYNCWARP();
                heapSorted = true;
                if (small > 0) sink(f, l, -1);   // two-level mode: tell the sink that this range is finished in place
                break;
            }
            --d;
            const idx_t cut = TABLE ? warpPartitionTable(arr, f, l, reinterpret_cast<unsigned short*>(tab),
                                                         reinterpret_cast<unsigned short*>(tab) + tabCap)
                                    : warpPartitionAny(arr, f, l, tab, inGlobal);
            if (l - cut > stopAt) {   // "recurse" on the right part: push
                FG_FOR_LANES
                    if (lane == (sp & 31)) {
                        if (sp < 32) { FG_L(stF0) = cut; FG_L(stL0) = l; FG_L(stD0) = d; }
                        else { FG_L(stF1) = cut; FG_L(stL1) = l; FG_L(stD1) = d; }
                    }
                FG_END_LANES
                ++sp;
            } else retire(cut, l, d);
            l = cut;
        }
        if (!heapSorted) retire(f, l, d);
        if (sp == 0) have = false;
        else {
            --sp;
#ifndef FG_WARP_HOST
            idx_t tf = sp < 32 ? stF0 : stF1, tl = sp < 32 ? stL0 : stL1; int td = sp < 32 ? stD0 : stD1;
            f = __shfl_sync(0xffffffffu, tf, sp & 31);
            l = __shfl_sync(0xffffffffu, tl, sp & 31);
            d = __shfl_sync(0xffffffffu, td, sp & 31);
#else
            f = sp < 32 ? stF0[sp & 31] : stF1[sp & 31];
            l = sp < 32 ? stL0[sp & 31] : stL1[sp & 31];
            d = sp < 32 ? stD0[sp & 31] : stD1[sp & 31];
#endif
        }
    }
    if (nLeaf) flushLeaves();
}

// ---- the whole sort of a range staged in shared memory -------------------------------------------------------
// Same permutation as warpIntrosortRange, organised for shared memory:
//   * every partition of more than MINI elements is a warpPartitionTable; with MINI = 32 the ranges of 17..32 elements
//     are queued and partitioned 32 at a time, one lane per range, by the literal sequential loop;
//   * no leaf is insertion-sorted on its own.  Every cut is recorded in a bitmap (`bits`, (n >> 5) + 2 words); when the
//     recursion is over the array is a concatenation of leaves of at most 16 elements, and __final_insertion_sort —
//     a STABLE sort that never moves an element across a cut — is done for all leaves by ranking: windows of up to 32
//     positions made of whole leaves, one lane per element, rank = #(smaller keys in my leaf) + #(equal keys before me).
//     (A range that ran out of depth budget is heap-sorted as in the reference and all its positions are marked as
//     cuts; it is sorted already, so the stable pass would not move anything.)
// tL/tR: n entries each.
template <int MINI>
FG_DEV void warpIntrosortSmem(Elem* arr, idx_t n, int d0, unsigned short* tL, unsigned short* tR, uint32_t* bits) {
    if (n < 2) return;
    FG_FOR_LANES for (idx_t w = lane; w < (n >> 5) + 2; w += 32) bits[w] = 0u; FG_END_LANES
    FG_SYNCWARP();
    FG_LANEVAR(idx_t, stF0); FG_LANEVAR(idx_t, stL0); FG_LANEVAR(int, stD0);   // stack entries 0..31
    FG_LANEVAR(idx_t, stF1); FG_LANEVAR(idx_t, stL1); FG_LANEVAR(int, stD1);   // stack entries 32..63
    FG_LANEVAR(idx_t, lfF);  FG_LANEVAR(idx_t, lfL); FG_LANEVAR(int, lfD);    // pending mini ranges (MINI > 16)
    int sp = 0, nLeaf = 0;
    FG_FOR_LANES FG_L(lfF) = 0; FG_L(lfL) = 0; FG_L(lfD) = 0; FG_END_LANES

    auto markAll = [&](idx_t f, idx_t l) { for (idx_t i = f; i < l; ++i) FG_ATOMIC_OR(&bits[i >> 5], 1u << (i & 31)); };
    auto flushMinis = [&]() {
        FG_SYNCWARP();
        FG_FOR_LANES
            if (lane < nLeaf) {
                idx_t sf[2], sl[2]; int sd[2]; int q = 0;
                idx_t f = FG_L(lfF), l = FG_L(lfL); int d = FG_L(lfD);
                for (;;) {
                    while (l - f > 16) {
                        if (d == 0) { seqHeapSort(arr + f, l - f); markAll(f, l); break; }
                        --d;
                        seqMedianToFirst(arr, f, f + 1, f + (l - f) / 2, l - 1);
                        const idx_t cut = (idx_t)seqUnguardedPartition(arr, f + 1, l, f);
                        FG_ATOMIC_OR(&bits[cut >> 5], 1u << (cut & 31));
                        if (l - cut > 16) { sf[q] = cut; sl[q] = l; sd[q] = d; ++q; }
                        l = cut;
                    }
                    if (q == 0) break;
                    --q; f = sf[q]; l = sl[q]; d = sd[q];
                }
            }
        FG_END_LANES
        FG_SYNCWARP();
        nLeaf = 0;
    };

    idx_t f = 0, l = n;
    int d = d0;
    bool have = true;
    while (have) {
        while (l - f > MINI) {
            if (d == 0) {
#ifndef FG_WARP_HOST
                if (fg::laneId() == 0)
#endif
                { seqHeapSort(arr + f, l - f); markAll(f, l); }
#ifdef FG_WARP_HOST
                ++g_heapSortCalls;
#endif
                FG_SYNCWARP();
                break;
            }
            --d;
            const idx_t cut = warpPartitionTable(arr, f, l, tL, tR);
#ifndef FG_WARP_HOST
            if (fg::laneId() == 0)
#endif
                bits[cut >> 5] |= 1u << (cut & 31);
            if (l - cut > MINI) {   // "recurse" on the right part: push
                FG_FOR_LANES
                    if (lane == (sp & 31)) {
                        if (sp < 32) { FG_L(stF0) = cut; FG_L(stL0) = l; FG_L(stD0) = d; }
                        else { FG_L(stF1) = cut; FG_L(stL1) = l; FG_L(stD1) = d; }
                    }
                FG_END_LANES
                ++sp;
            } else if (MINI > 16 && l - cut > 16) {
                FG_FOR_LANES if (lane == nLeaf) { FG_L(lfF) = cut; FG_L(lfL) = l; FG_L(lfD) = d; } FG_END_LANES
                if (++nLeaf == 32) flushMinis();
            }
            l = cut;
        }
        if (MINI > 16 && l - f > 16 && l - f <= MINI) {
            FG_FOR_LANES if (lane == nLeaf) { FG_L(lfF) = f; FG_L(lfL) = l; FG_L(lfD) = d; } FG_END_LANES
            if (++nLeaf == 32) flushMinis();
        }
        if (sp == 0) have = false;
        else {
            --sp;
#ifndef FG_WARP_HOST
            idx_t tf = sp < 32 ? stF0 : stF1, tl = sp < 32 ? stL0 : stL1; int td = sp < 32 ? stD0 : stD1;
            f = __shfl_sync(0xffffffffu, tf, sp & 31);
            l = __shfl_sync(0xffffffffu, tl, sp & 31);
            d = __shfl_sync(0xffffffffu, td, sp & 31);
#else
            f = sp < 32 ? stF0[sp & 31] : stF1[sp & 31];
            l = sp < 32 ? stL0[sp & 31] : stL1[sp & 31];
            d = sp < 32 ? stD0[sp & 31] : stD1[sp & 31];
#endif
        }
    }
    if (MINI > 16 && nLeaf) flushMinis();
    FG_SYNCWARP();

    // the stable pass over all leaves
    for (idx_t pos = 0; pos < n;) {
        // M: bit t <-> position pos + t is a cut (bit 0 always: pos is one); end = last cut in (pos, pos + 32]
        const idx_t q = pos + 1;
        const uint32_t w0 = bits[q >> 5], w1 = bits[(q >> 5) + 1];
        const int sh = q & 31;
        const uint32_t m = sh ? (w0 >> sh) | (w1 << (32 - sh)) : w0;   // bit t <-> cut at pos + 1 + t
        const idx_t end = pos + 32 >= n ? n : pos + 1 + (31 - FG_CLZ(m));
        const uint32_t M = (m << 1) | 1u;
        FG_LANEVAR(idx_t, la); FG_LANEVAR(idx_t, lb); FG_LANEVAR(unsigned long long, lk); FG_LANEVAR(int, cnt);
        FG_FOR_LANES
            const idx_t i = pos + lane;
            const uint32_t upTo = M & (lane == 31 ? 0xffffffffu : ((2u << lane) - 1u));
            FG_L(la) = pos + (31 - FG_CLZ(upTo));
            const uint32_t above = lane == 31 ? 0u : (M & ~((2u << lane) - 1u));
            idx_t b = above ? pos + FG_CTZ(above) : end;
            if (b > end) b = end;
            FG_L(lb) = i < end ? b : FG_L(la);   // inactive lanes: empty leaf
            FG_L(lk) = i < end ? arr[i].key : 0ULL;
            FG_L(cnt) = 0;
        FG_END_LANES
        // leaves that are in order already (common when the input was nearly sorted) need nothing
        uint32_t unsorted;
        FG_BALLOT(unsorted, pos + lane < end && pos + lane > FG_L(la) && arr[pos + lane - 1].key > FG_L(lk));
        if (!unsorted) { pos = end; continue; }
        int maxLen;
        FG_REDUCE_MAX(maxLen, FG_L(lb) - FG_L(la));
        for (int dd = 0; dd < maxLen; ++dd) {
            FG_FOR_LANES
                const idx_t j = FG_L(la) + dd;
                if (j < FG_L(lb)) {
                    const unsigned long long kj = arr[j].key;
                    FG_L(cnt) += (kj < FG_L(lk)) || (kj == FG_L(lk) && j < pos + lane);
                }
            FG_END_LANES
        }
        FG_LANEVAR(Elem, ev);
        FG_FOR_LANES if (pos + lane < end) FG_L(ev) = arr[pos + lane]; FG_END_LANES
        FG_SYNCWARP();
        FG_FOR_LANES if (pos + lane < end) arr[FG_L(la) + FG_L(cnt)] = FG_L(ev); FG_END_LANES
        FG_SYNCWARP();
        pos = end;
    }
}

FG_DEV int introsortDepth(long n) {   // std::__lg(n) * 2
    int lg = 0;
    while ((n >> (lg + 1)) != 0) ++lg;
    return 2 * lg;
}

// std::sort(arr, arr + n) in one go
FG_DEV void warpIntrosort(Elem* arr, idx_t n, unsigned char* tab) {
    if (n < 2) return;
    NoSink none;
    warpIntrosortRange(arr, 0, n, introsortDepth(n), 0, none, tab);
}

}  // namespace fg
