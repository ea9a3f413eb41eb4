// flye_b200 — base-level divergence of primary overlaps (HiFi / corrected-read presets).
//
// Replaces getAlignmentErrEdlib (src/sequence/alignment.cpp:218-247): homopolymer compression of the two
// overlapping substrings (alignment.cpp:52-70) and the exact global unit-cost edit distance that
// edlibAlign(k=-1, EDLIB_MODE_NW, EDLIB_TASK_DISTANCE) returns (edlib.cpp:141-212).  The distance is a
// unique integer, so any exact algorithm is bit-compatible; the float division stays on the host.
//
// K9a hpcReadsKernel  once per sequence set (cached until the next upload): every read is homopolymer-compressed as
//                 a whole into a 2-bit packed sequence, with a run-start mask and a running count per 32-base word.
//                 The compression of a SUBSTRING is the runs that intersect it, i.e. a contiguous slice
//                 [rank(begin), rank(end-1)] of the read's compressed sequence, and the compression of a substring
//                 of the reverse-complement strand is the reverse complement of a slice — so an overlap needs no
//                 compaction pass of its own, only two rank look-ups.
// K9b wfaKernel   one warp per overlap: furthest-reaching wavefront (Ukkonen / Myers O(ND)) over the diagonals,
//                 32 diagonals per step.  Sequences are read through "views" (slice + strand) straight from the
//                 packed words; a lane extends its diagonal 32 bases per compare (xor + count trailing zeros), the
//                 few diagonals that keep matching (the true alignment path) are extended by the whole warp, 1024
//                 bases per step.  Wavefronts live in shared memory up to WF_DMAX edits and continue in global
//                 scratch beyond.  Work is O(d^2 + n) for distance d — HiFi overlaps have d ~ 10^2 on 10^4 bases.
#include "ctx.cuh"

namespace fg {

// ---- K9a ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t evenBits(uint64_t x) {   // bit 2t of x -> bit t
    x &= 0x5555555555555555ULL;
    x = (x | (x >> 1)) & 0x3333333333333333ULL;
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0fULL;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffULL;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffULL;
    x = (x | (x >> 16)) & 0x00000000ffffffffULL;
    return (uint32_t)x;
}

// one warp per read.  hpcSeq must be zero on entry (bases are OR-ed in).  For raw word w of the read:
// startMask[w] bit t = base 32w+t starts a run (differs from its predecessor; base 0 always does),
// startPrefix[w] = number of run starts in the words before w.
__global__ void __launch_bounds__(128) hpcReadsKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                      const uint32_t* __restrict__ len, uint32_t nReads,
                                                      unsigned long long* __restrict__ hpcSeq, uint32_t* __restrict__ startMask,
                                                      uint32_t* __restrict__ startPrefix) {
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nReads) return;
    const int lane = threadIdx.x & 31;
    const uint64_t off = wordOff[r];
    const uint32_t L = len[r], nWords = (L + 31) >> 5;
    uint32_t tot = 0;
    for (uint32_t w0 = 0; w0 < nWords; w0 += 32) {
        const uint32_t wi = w0 + lane;
        uint64_t word = 0; uint32_t mask = 0;
        if (wi < nWords) {
            word = seq[off + wi];
            const uint64_t prevBase = wi ? seq[off + wi - 1] >> 62 : 0ULL;
            const uint64_t x = word ^ ((word << 2) | prevBase);
            mask = evenBits(x | (x >> 1));
            if (wi == 0) mask |= 1u;
            const uint32_t nb = min(32u, L - 32u * wi);
            if (nb < 32) mask &= (1u << nb) - 1u;
        }
        const uint32_t cnt = __popc(mask);
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        const uint32_t o = tot + inc - cnt;
        if (wi < nWords) {
            startMask[off + wi] = mask; startPrefix[off + wi] = o;
            unsigned long long v = 0; int c = 0;
            for (uint32_t mm = mask; mm; mm &= mm - 1) { const int t = __ffs(mm) - 1; v |= ((word >> (2 * t)) & 3ULL) << (2 * c); ++c; }
            if (cnt) {
                const uint32_t sh = (o & 31) * 2;
                atomicOr(hpcSeq + off + (o >> 5), v << sh);
                if (sh && (o & 31) + cnt > 32) atomicOr(hpcSeq + off + (o >> 5) + 1, v >> (64 - sh));
            }
        }
        tot += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// ---- K9b ----------------------------------------------------------------------------------------------------
// A sequence as the wavefront kernel sees it: n bases, base p = words[r0 + p] (forward strand) or the complement of
// words[r1 - p] (reverse-complement strand), positions counted in 2-bit bases from `words`.
struct SeqView { const uint64_t* words; int32_t r0, r1, n; bool rc; };

__device__ __forceinline__ uint64_t fetchFwd(const uint64_t* __restrict__ words, int32_t h) {   // 32 bases from base h >= 0
    const uint64_t* w = words + (h >> 5);
    const int sh = (h & 31) * 2;
    return (w[0] >> sh) | ((w[1] << 1) << (63 - sh));
}
// 32 bases of the view starting at p (base t in bits 2t, 2t+1); bases at or beyond n are unspecified
__device__ __forceinline__ uint64_t fetchView(const SeqView& v, int32_t p) {
    if (!v.rc) return fetchFwd(v.words, v.r0 + p);
    const int32_t h = v.r1 - p - 31;
    const uint64_t x = h >= 0 ? fetchFwd(v.words, h) : fetchFwd(v.words, 0) << (2 * min(-h, 31));
    return rev2(~x);
}
// number of matching bases from (pa, pb), at most min(lim, 32)
__device__ __forceinline__ int matchLen32(const SeqView& a, int32_t pa, const SeqView& b, int32_t pb, int lim) {
    const uint64_t x = fetchView(a, pa) ^ fetchView(b, pb);
    const int e = x ? (__ffsll((long long)x) - 1) >> 1 : 32;
    return min(e, lim);
}
// the same by the whole warp, 1024 bases per step (every lane returns the result)
__device__ __forceinline__ int matchLenWarp(const SeqView& a, int32_t pa, const SeqView& b, int32_t pb, int lim, int lane) {
    for (int base = 0;; base += 1024) {
        const int o = base + 32 * lane;
        const int l = max(0, min(32, lim - o));
        const int e = l > 0 ? matchLen32(a, pa + o, b, pb + o, l) : 0;
        const uint32_t stop = __ballot_sync(0xffffffffu, e < 32);
        if (stop) { const int f = __ffs(stop) - 1; return base + 32 * f + __shfl_sync(0xffffffffu, e, f); }
    }
}

static constexpr int WF_DMAX = 254;                 // edits handled with the wavefronts in shared memory
static constexpr int WF_SMEM_INTS = 2 * WF_DMAX + 5;   // diagonals -DMAX-2 .. DMAX+2
static constexpr int WF_NEG = -(1 << 29);

// Steps s = sFrom .. sTo of the wavefront recurrence.  fr[k + base] = furthest row i reached on diagonal k = j - i with
// s edits (WF_NEG / -1 = unreachable); `prev` holds step sFrom-1 on entry.  Returns the distance when the end is
// reached, else -1 with `prev` holding step sTo.
// `band` (>= 1): only distances < band matter to the caller.  A path that ends on the target diagonal with fewer than `band`
// edits is, after s edits, at most band-1-s diagonals away from it — every other diagonal of step s is left out (its value
// reads as unreachable).  The optimal path of any distance < band lies inside this diamond, so such a distance is still found
// exactly, at the same step; the diamond has about half the area of the full triangle of diagonals.
// (kept out of line and free of min / max chains on purpose: the one-line forms `max(max(-s, -n), …)`, inlined into the second
// wfaSteps call site (global-memory phase), came out of nvcc 12.9 returning +n for s > WF_DMAX — caught by the host-mirror
// parity test; tests/test_gpu_parity.py::test_device_edit_distance_long_and_bounded pins the phase)
__device__ __noinline__ int bandLo(int s, int n, int target, int band) {
    int lo = -s;
    if (lo < -n) lo = -n;
    const int b = target - (band - 1 - s);
    if (lo < b) lo = b;
    return lo;
}
__device__ __noinline__ int bandHi(int s, int m, int target, int band) {
    int hi = s;
    if (hi > m) hi = m;
    const int b = target + (band - 1 - s);
    if (hi > b) hi = b;
    return hi;
}

__device__ int wfaSteps(const SeqView& A, const SeqView& B, int*& prev, int*& cur, int base, int sFrom, int sTo, int lane, int band) {
    const int n = A.n, m = B.n, target = m - n;
    for (int s = sFrom; s <= sTo; ++s) {
        const int lo = bandLo(s, n, target, band), hi = bandHi(s, m, target, band);
        const int plo = bandLo(s - 1, n, target, band), phi = bandHi(s - 1, m, target, band);
        if (lane < 2) prev[plo - 1 - lane + base] = WF_NEG; else if (lane < 4) prev[phi - 1 + lane + base] = WF_NEG;   // guards
        __syncwarp();
        bool done = false;
        for (int k0 = lo; k0 <= hi; k0 += 32) {
            const int k = k0 + lane;
            int best = -1;
            bool longRun = false;
            if (k <= hi) {
                const int a = prev[k - 1 + base], b = prev[k + base], c = prev[k + 1 + base];
                if (a >= 0 && a + k <= m) best = a;
                if (b >= 0 && b + 1 <= n && b + k + 1 <= m) best = max(best, b + 1);
                if (c >= 0 && c + 1 <= n) best = max(best, c + 1);
                if (best >= 0) {
                    const int lim = min(n - best, m - best - k);
                    const int e = lim > 0 ? matchLen32(A, best, B, best + k, lim) : 0;
                    best += e;
                    longRun = e == 32 && lim > 32;
                }
            }
            uint32_t lm = __ballot_sync(0xffffffffu, longRun);
            while (lm) {
                const int src = __ffs(lm) - 1;
                lm &= lm - 1;
                const int bb = __shfl_sync(0xffffffffu, best, src);
                const int kk = k0 + src;
                const int e = matchLenWarp(A, bb, B, bb + kk, min(n - bb, m - bb - kk), lane);
                if (lane == src) best = bb + e;
            }
            if (k <= hi) {
                cur[k + base] = best;
                if (k == target && best >= n) done = true;
            }
        }
        if (__any_sync(0xffffffffu, done)) return s;
        __syncwarp();
        int* t = prev; prev = cur; cur = t;
    }
    return -1;
}

// exact global edit distance of the two views, or `limit` as soon as the distance is known to be >= limit; one warp.
// smem: 2 * WF_SMEM_INTS ints; gA/gB: n + m + 8 ints each.
__device__ int wfaEditDistance(const SeqView& A, const SeqView& B, int* smem, int* gA, int* gB, int limit = 0x7fffffff, bool bandOn = true) {
    const int lane = threadIdx.x & 31;
    const int n = A.n, m = B.n;
    if (limit <= 0) return limit;
    if (abs(n - m) >= limit) return limit;   // the length difference alone costs that many edits
    if (n == 0) return m;
    if (m == 0) return n;
    int* prev = smem; int* cur = smem + WF_SMEM_INTS;
    const int base = WF_DMAX + 2;
    const int i0 = matchLenWarp(A, 0, B, 0, min(n, m), lane);   // s = 0: common prefix
    if (m == n && i0 >= n) return 0;
    if (lane == 0) prev[base] = i0;
    __syncwarp();
    const int sMax = min(n + m, limit - 1);   // steps worth taking: beyond them the distance is >= limit
    const int band = bandOn ? min(limit, n + m + 1) : n + m + 1;   // distances of interest are < band (the distance never exceeds n + m)
    const int dCap = min(WF_DMAX, sMax);
    int d = wfaSteps(A, B, prev, cur, base, 1, dCap, lane, band);
    if (d >= 0) return d;
    if (dCap == sMax) return sMax == n + m ? d : limit;
    // continue in global memory
    const int gBase = n + 2;
    const int plo = bandLo(dCap, n, m - n, band), phi = bandHi(dCap, m, m - n, band);
    for (int k = plo + lane; k <= phi; k += 32) gA[k + gBase] = prev[k + base];
    __syncwarp();
    prev = gA; cur = gB;
    d = wfaSteps(A, B, prev, cur, gBase, dCap + 1, sMax, lane, band);
    return (d < 0 && sMax < n + m) ? limit : d;
}

struct HpcSet { const uint64_t* seq; const uint64_t* hpc; const uint32_t* mask; const uint32_t* prefix; const uint64_t* wordOff; const uint32_t* len; };

// run index of raw base i of a read
__device__ __forceinline__ int32_t hpcRank(const HpcSet& S, uint64_t off, uint32_t i) {
    const uint32_t w = i >> 5;
    return (int32_t)(S.prefix[off + w] + __popc(S.mask[off + w] & (0xffffffffu >> (31 - (i & 31))))) - 1;
}
// the substring [begin, end) of sequence `id` (strand = id & 1), homopolymer-compressed or raw
__device__ SeqView makeView(const HpcSet& S, uint32_t id, int32_t begin, int32_t end, bool compress) {
    SeqView v;
    const uint32_t r = id >> 1;
    const uint64_t off = S.wordOff[r];
    const int32_t L = (int32_t)S.len[r];
    v.rc = id & 1;
    v.words = (compress ? S.hpc : S.seq) + off;
    if (end <= begin) { v.r0 = 0; v.r1 = -1; v.n = 0; return v; }
    const int32_t fb = v.rc ? L - end : begin, fe = v.rc ? L - begin : end;   // forward-strand coordinates, fe exclusive
    if (compress) { v.r0 = hpcRank(S, off, (uint32_t)fb); v.r1 = hpcRank(S, off, (uint32_t)(fe - 1)); }
    else { v.r0 = fb; v.r1 = fe - 1; }
    v.n = v.r1 - v.r0 + 1;
    return v;
}

// The divergence filter (overlap.cpp:470-473) keeps an overlap iff (float)editDistance / alnLen < maxDivergence, so an
// overlap whose distance reaches T = the smallest integer with (float)T / alnLen >= maxDivergence is dropped whatever the
// exact value is: the wavefronts stop there and T is reported.  (Most of the cost used to be spent on exactly those
// overlaps — two thirds of the HiFi candidates fail the 1 % threshold, with distances far beyond it.)
__device__ __forceinline__ int dropLimit(float maxDiv, int alnLen) {
    if (alnLen <= 0) return 0x7fffffff;
    if (!(maxDiv > 0.f)) return 0;   // nothing passes a threshold <= 0 (or NaN)
    const float L = (float)alnLen;
    long long T = (long long)ceilf(__fmul_rn(maxDiv, L));
    if (T > 0x7ffffff0LL) return 0x7fffffff;
    while (T > 0 && __fdiv_rn((float)(T - 1), L) >= maxDiv) --T;
    while (T < 0x7ffffff0LL && __fdiv_rn((float)T, L) < maxDiv) ++T;
    return (int)T;
}

__global__ void __launch_bounds__(128) wfaKernel(fg_overlap* __restrict__ ov, uint32_t nOv, HpcSet cur, HpcSet ext, bool compress,
                                                 int* __restrict__ wfScratch, uint64_t wfStride, uint32_t* __restrict__ nextJob,
                                                 float maxDiv, const float* __restrict__ queryMaxDiv, bool bandOn) {
    __shared__ int wfSm[4][2 * WF_SMEM_INTS];
    const uint32_t warpGlobal = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    int* gA = wfScratch + (uint64_t)warpGlobal * 2 * wfStride;
    int* gB = gA + wfStride;
    for (;;) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(nextJob, 1u);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= nOv) return;
        const fg_overlap o = ov[j];
        const SeqView A = makeView(cur, o.cur_id, o.cur_begin, o.cur_end, compress);
        const SeqView B = makeView(ext, o.ext_id, o.ext_begin, o.ext_end, compress);
        const int limit = dropLimit(queryMaxDiv ? queryMaxDiv[o.reserved] : maxDiv, max(A.n, B.n));
        const int d = wfaEditDistance(A, B, wfSm[threadIdx.x >> 5], gA, gB, limit, bandOn);
        if (lane == 0) { ov[j].edit_distance = d; ov[j].aln_len = max(A.n, B.n); }
        __syncwarp();
    }
}

// longest cur / ext range of a batch of overlaps (sizes the global wavefront scratch)
__global__ void __launch_bounds__(256) maxRangeKernel(const fg_overlap* __restrict__ ov, uint32_t nOv, uint32_t* __restrict__ out) {
    uint32_t m = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nOv; i += gridDim.x * blockDim.x)
        m = max(m, (uint32_t)max(max(0, ov[i].cur_end - ov[i].cur_begin), max(0, ov[i].ext_end - ov[i].ext_begin)));
    for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// homopolymer-compressed copy of a sequence set (see K9a)
static void buildHpc(fg_ctx* ctx, fg_ctx::HpcCache& c, const uint64_t* seq, const uint64_t* wordOff, const uint32_t* len, uint32_t nReads,
                     uint64_t nWords) {
    c.seq.alloc(nWords + 4); c.mask.alloc(nWords + 4); c.prefix.alloc(nWords + 4);
    FG_CUDA(cudaMemsetAsync(c.seq.p, 0, (nWords + 4) * 8, streamOf(ctx)));
    if (nReads) {
        hpcReadsKernel<<<(nReads + 3) / 4, 128, 0, streamOf(ctx)>>>(seq, wordOff, len, nReads, reinterpret_cast<unsigned long long*>(c.seq.p), c.mask.p,
                                                                  c.prefix.p);
        checkLaunch(ctx, "hpcReadsKernel");
    }
    c.valid = true;
}

// the homopolymer-compressed copies are built once per uploaded sequence set, on the context's own stream, before the lanes of
// fg_overlaps_batch start (they only read them)
void prepareEditDistances(fg_ctx* ctx, bool useHpc, bool querySet) {
    if (!useHpc) return;
    if (!ctx->hpcReads.valid) buildHpc(ctx, ctx->hpcReads, ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p, ctx->nReads, ctx->hWordOff.back());
    if (querySet && !ctx->hpcQueries.valid)
        buildHpc(ctx, ctx->hpcQueries, ctx->dQsSeq.p, ctx->dQsWordOff.p, ctx->dQsLen.p, ctx->nQsReads, ctx->nQsWords);
    FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
}

// device overlaps (already gathered) -> edit_distance / aln_len filled in.  querySet: the "cur" side refers to the second
// sequence set (fg_queries_upload).
void editDistances(fg_ctx* ctx, fg_overlap* dOv, const fg_overlap* hOv, uint32_t nOv, bool useHpc, bool querySet, float maxDivergence,
                   const float* dQueryMaxDivergence) {
    if (!nOv) return;
    uint32_t maxLen = 0;
    if (hOv) {
        for (uint32_t i = 0; i < nOv; ++i)
            maxLen = std::max(maxLen, (uint32_t)std::max(std::max(0, hOv[i].cur_end - hOv[i].cur_begin), std::max(0, hOv[i].ext_end - hOv[i].ext_begin)));
    } else {   // no host copy of the records (device epilogue): the longest range comes from a reduction on the device
        DevBuf<uint32_t> dMax(1);
        FG_CUDA(cudaMemsetAsync(dMax.p, 0, 4, streamOf(ctx)));
        maxRangeKernel<<<gridFor(nOv), 256, 0, streamOf(ctx)>>>(dOv, nOv, dMax.p);
        checkLaunch(ctx, "maxRangeKernel");
        FG_CUDA(cudaMemcpyAsync(&maxLen, dMax.p, 4, cudaMemcpyDeviceToHost, streamOf(ctx)));
        FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
    }
    HpcSet ext{ctx->dSeq.p, nullptr, nullptr, nullptr, ctx->dWordOff.p, ctx->dLen.p};
    HpcSet cur = ext;
    if (querySet) cur = HpcSet{ctx->dQsSeq.p, nullptr, nullptr, nullptr, ctx->dQsWordOff.p, ctx->dQsLen.p};
    if (useHpc) {
        if (!ctx->hpcReads.valid || (querySet && !ctx->hpcQueries.valid)) throw Error(FG_ERR_INTERNAL, "prepareEditDistances has not run");
        ext.hpc = ctx->hpcReads.seq.p; ext.mask = ctx->hpcReads.mask.p; ext.prefix = ctx->hpcReads.prefix.p;
        if (querySet) {
            cur.hpc = ctx->hpcQueries.seq.p; cur.mask = ctx->hpcQueries.mask.p; cur.prefix = ctx->hpcQueries.prefix.p;
        } else { cur.hpc = ext.hpc; cur.mask = ext.mask; cur.prefix = ext.prefix; }
    }
    DevBuf<uint32_t> nextJob(1);
    FG_CUDA(cudaMemsetAsync(nextJob.p, 0, 4, streamOf(ctx)));
    const int blocks = 148 * 10, warps = blocks * 4;   // 48 registers / thread: 40 resident warps per SM
    const uint64_t stride = 2ULL * maxLen + 8;
    DevBuf<int> wf((uint64_t)warps * 2 * stride);   // only touched by overlaps with more than WF_DMAX edits
    wfaKernel<<<blocks, 128, 0, streamOf(ctx)>>>(dOv, nOv, cur, ext, useHpc, wf.p, stride, nextJob.p, maxDivergence, dQueryMaxDivergence,
                                               getenv("FG_WFA_BAND") == nullptr || atoi(getenv("FG_WFA_BAND")) != 0);
    checkLaunch(ctx, "wfaKernel");
    FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
}

// test hook: exact edit distance of two base strings (values 0..3), optionally seen as reverse complements
__global__ void debugEdKernel(const uint64_t* a, int n, bool rcA, const uint64_t* b, int m, bool rcB, int* wf, uint64_t stride, int* out, int limit) {
    __shared__ int sm[2 * WF_SMEM_INTS];
    SeqView A{a, 0, n - 1, n, rcA}, B{b, 0, m - 1, m, rcB};
    const int d = wfaEditDistance(A, B, sm, wf, wf + stride, limit);
    if ((threadIdx.x & 31) == 0) *out = d;
}

int debugEditDistance(fg_ctx* ctx, const uint8_t* a, int n, const uint8_t* b, int m, int rcA, int rcB) {
    auto pack = [](const uint8_t* s, int len) {
        std::vector<uint64_t> w((size_t)len / 32 + 3, 0);
        for (int i = 0; i < len; ++i) {
            if (s[i] > 3) throw Error(FG_ERR_ARG, "bases must be 0..3");
            w[i >> 5] |= (uint64_t)s[i] << ((i & 31) * 2);
        }
        return w;
    };
    const std::vector<uint64_t> pa = pack(a, n), pb = pack(b, m);
    DevBuf<uint64_t> dA(pa.size()), dB(pb.size());
    const uint64_t stride = (uint64_t)n + m + 8;
    DevBuf<int> wf(2 * stride), dOut(1);
    FG_CUDA(cudaMemcpyAsync(dA.p, pa.data(), pa.size() * 8, cudaMemcpyHostToDevice, streamOf(ctx)));
    FG_CUDA(cudaMemcpyAsync(dB.p, pb.data(), pb.size() * 8, cudaMemcpyHostToDevice, streamOf(ctx)));
    // test hook of the test hook: FG_DEBUG_ED_LIMIT bounds the distance like the divergence threshold does in wfaKernel
    const char* lim = getenv("FG_DEBUG_ED_LIMIT");
    debugEdKernel<<<1, 32, 0, streamOf(ctx)>>>(dA.p, n, rcA != 0, dB.p, m, rcB != 0, wf.p, stride, dOut.p, lim ? atoi(lim) : 0x7fffffff);
    checkLaunch(ctx, "debugEdKernel");
    int d = -1;
    FG_CUDA(cudaMemcpyAsync(&d, dOut.p, 4, cudaMemcpyDeviceToHost, streamOf(ctx)));
    FG_CUDA(cudaStreamSynchronize(streamOf(ctx)));
    return d;
}

}  // namespace fg
