// flye_b200 — base-level divergence of primary overlaps (HiFi / corrected-read presets).
//
// Replaces getAlignmentErrEdlib (src/sequence/alignment.cpp:218-247): homopolymer compression of the two
// overlapping substrings (alignment.cpp:52-70) and the exact global unit-cost edit distance that
// edlibAlign(k=-1, EDLIB_MODE_NW, EDLIB_TASK_DISTANCE) returns (edlib.cpp:141-212).  The distance is a
// unique integer, so any exact algorithm is bit-compatible; the float division stays on the host.
//
// K9a hpcKernel   one warp per overlap and side: ballot-compaction of "base != previous base" into a
//                 byte-per-base scratch sequence (either strand of the read is read in place).
// K9b wfaKernel   one warp per overlap: furthest-reaching wavefront (Ukkonen / Myers O(ND)) over the
//                 diagonals, 32 diagonals per step, match extension by direct comparison.  Work is O(d^2 + n)
//                 for distance d — HiFi overlaps have d of a few hundred on 10^4 bases, so this is ~50x less
//                 work than the bit-vector band edlib uses, and it needs no band-doubling restarts.
#include "ctx.cuh"

namespace fg {

__device__ __forceinline__ uint32_t baseOf(const uint64_t* __restrict__ words, uint32_t L, bool strand, uint32_t i) {
    // DnaSequence::atRaw, sequence.h:120-129
    const uint32_t idx = strand ? (L - 1 - i) : i;
    const uint32_t c = (uint32_t)(words[idx >> 5] >> ((idx & 31) * 2)) & 3u;
    return strand ? 3u - c : c;
}

struct EdJob { uint64_t offA, offB; };   // byte offsets of the two compressed sequences in the scratch

__global__ void __launch_bounds__(256) hpcKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                 const uint32_t* __restrict__ len, const uint64_t* __restrict__ qseq,
                                                 const uint64_t* __restrict__ qWordOff, const uint32_t* __restrict__ qlen,
                                                 const fg_overlap* __restrict__ ov, uint32_t nOv,
                                                 const EdJob* __restrict__ jobs, bool compress, uint8_t* __restrict__ scratch,
                                                 uint32_t* __restrict__ outLen) {
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= 2 * nOv) return;
    const int lane = threadIdx.x & 31;
    const fg_overlap o = ov[w >> 1];
    const bool side = w & 1;   // 0: cur, 1: ext
    const uint32_t id = side ? o.ext_id : o.cur_id;
    const uint32_t start = (uint32_t)(side ? o.ext_begin : o.cur_begin);
    const uint32_t n = (uint32_t)(side ? (o.ext_end - o.ext_begin) : (o.cur_end - o.cur_begin));
    const uint32_t r = id >> 1, L = side ? len[r] : qlen[r];   // side 0 = the query ("cur"), possibly from the second set
    const bool strand = id & 1;
    const uint64_t* words = side ? seq + wordOff[r] : qseq + qWordOff[r];
    uint8_t* dst = scratch + (side ? jobs[w >> 1].offB : jobs[w >> 1].offA);
    uint32_t out = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t c = 255, prev = 255;
        if (i < n) {
            c = baseOf(words, L, strand, start + i);
            if (i > 0) prev = baseOf(words, L, strand, start + i - 1);
        }
        const bool keep = i < n && (!compress || i == 0 || c != prev);
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) dst[out + __popc(m & ((1u << lane) - 1u))] = (uint8_t)c;
        out += __popc(m);
    }
    if (lane == 0) outLen[w] = out;
}

// Exact global edit distance of a[0,n) and b[0,m).  fr[k] = furthest row i reached on diagonal k = j - i with
// the current number of edits; -1 = unreachable.  One warp; the two wavefront arrays live in global scratch
// (index k + n).
static constexpr int WFA_SHORT = 8;

__device__ int wfaEditDistance(const uint8_t* __restrict__ a, int n, const uint8_t* __restrict__ b, int m, int* __restrict__ wfA,
                               int* __restrict__ wfB) {
    const int lane = threadIdx.x & 31;
    if (n == 0) return m;
    if (m == 0) return n;
    int* prev = wfA; int* cur = wfB;
    {   // s = 0: common prefix, 32 bases per step
        const int lim = min(n, m);
        int i = 0;
        for (;;) {
            const int j = i + lane;
            const uint32_t mm = __ballot_sync(0xffffffffu, j >= lim || a[j] != b[j]);
            if (mm) { i += __ffs(mm) - 1; break; }
            i += 32;
        }
        if (lane == 0) prev[n] = i;
        if (m == n && i >= n) return 0;
    }
    __syncwarp();
    const int target = m - n;
    for (int s = 1; s <= n + m; ++s) {
        const int lo = max(-s, -n), hi = min(s, m);
        const int plo = max(-(s - 1), -n), phi = min(s - 1, m);
        bool done = false;
        for (int k0 = lo; k0 <= hi; k0 += 32) {
            const int k = k0 + lane;
            int best = -1;
            bool longRun = false;
            if (k <= hi) {
                if (k - 1 >= plo && k - 1 <= phi) { const int i = prev[k - 1 + n]; if (i >= 0 && i + k <= m) best = i; }
                if (k >= plo && k <= phi) { const int i = prev[k + n]; if (i >= 0 && i + 1 <= n && i + k + 1 <= m) best = max(best, i + 1); }
                if (k + 1 >= plo && k + 1 <= phi) { const int i = prev[k + 1 + n]; if (i >= 0 && i + 1 <= n) best = max(best, i + 1); }
                if (best >= 0) {   // short extension by the lane itself (off-diagonals stop after ~1 base)
                    const uint8_t* pa = a + best; const uint8_t* pb = b + best + k;
                    const int lim = min(min(n - best, m - best - k), WFA_SHORT);
                    int e = 0;
                    while (e < lim && pa[e] == pb[e]) ++e;
                    best += e;
                    longRun = e == WFA_SHORT;
                }
            }
            // the few diagonals that keep matching (the true alignment path) are extended by the whole warp, 32 bases
            // per step, instead of one lane crawling alone
            uint32_t lm = __ballot_sync(0xffffffffu, longRun);
            while (lm) {
                const int src = __ffs(lm) - 1;
                lm &= lm - 1;
                int bb = __shfl_sync(0xffffffffu, best, src);
                const int kk = k0 + src;
                const int lim = min(n - bb, m - bb - kk);
                int e = 0;
                for (;;) {
                    const int i = e + lane;
                    const bool mismatch = i >= lim || a[bb + i] != b[bb + kk + i];
                    const uint32_t mm = __ballot_sync(0xffffffffu, mismatch);
                    if (mm) { e += __ffs(mm) - 1; break; }
                    e += 32;
                }
                if (lane == src) best = bb + e;
            }
            if (k <= hi) {
                cur[k + n] = best;
                if (k == target && best >= n) done = true;
            }
        }
        if (__any_sync(0xffffffffu, done)) return s;
        __syncwarp();
        int* t = prev; prev = cur; cur = t;
    }
    return -1;   // unreachable: the distance never exceeds max(n,m)
}

__global__ void __launch_bounds__(128) wfaKernel(fg_overlap* __restrict__ ov, uint32_t nOv, const EdJob* __restrict__ jobs,
                                                 const uint32_t* __restrict__ hpcLen, const uint8_t* __restrict__ scratch,
                                                 int* __restrict__ wfScratch, uint64_t wfStride, uint32_t* __restrict__ nextJob) {
    const uint32_t warpGlobal = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    int* wfA = wfScratch + (uint64_t)warpGlobal * 2 * wfStride;
    int* wfB = wfA + wfStride;
    for (;;) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(nextJob, 1u);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= nOv) return;
        const int n = (int)hpcLen[2 * j], m = (int)hpcLen[2 * j + 1];
        const int d = wfaEditDistance(scratch + jobs[j].offA, n, scratch + jobs[j].offB, m, wfA, wfB);
        if (lane == 0) { ov[j].edit_distance = d; ov[j].aln_len = max(n, m); }
        __syncwarp();
    }
}

// device overlaps (already gathered) -> edit_distance / aln_len filled in
void editDistances(fg_ctx* ctx, fg_overlap* dOv, const fg_overlap* hOv, uint32_t nOv, bool useHpc, const uint64_t* qSeq,
                   const uint64_t* qWordOff, const uint32_t* qLen) {
    if (!nOv) return;
    std::vector<EdJob> jobs(nOv);
    uint64_t off = 0; uint32_t maxLen = 0;
    for (uint32_t i = 0; i < nOv; ++i) {
        const uint32_t la = (uint32_t)std::max(0, hOv[i].cur_end - hOv[i].cur_begin), lb = (uint32_t)std::max(0, hOv[i].ext_end - hOv[i].ext_begin);
        jobs[i].offA = off; off += (la + 15u) & ~15u;
        jobs[i].offB = off; off += (lb + 15u) & ~15u;
        maxLen = std::max(maxLen, std::max(la, lb));
    }
    DevBuf<EdJob> dJobs(nOv);
    DevBuf<uint8_t> scratch(off + 16);
    DevBuf<uint32_t> hpcLen(2 * (size_t)nOv), nextJob(1);
    FG_CUDA(cudaMemcpyAsync(dJobs.p, jobs.data(), nOv * sizeof(EdJob), cudaMemcpyHostToDevice, ctx->stream));
    FG_CUDA(cudaMemsetAsync(nextJob.p, 0, 4, ctx->stream));
    hpcKernel<<<(2 * nOv + 7) / 8, 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p, qSeq, qWordOff, qLen, dOv, nOv, dJobs.p, useHpc,
                                                          scratch.p, hpcLen.p);
    checkLaunch(ctx, "hpcKernel");
    const int blocks = 148 * 8, warps = blocks * 4;
    const uint64_t stride = 2ULL * maxLen + 8;
    DevBuf<int> wf((uint64_t)warps * 2 * stride);
    wfaKernel<<<blocks, 128, 0, ctx->stream>>>(dOv, nOv, dJobs.p, hpcLen.p, scratch.p, wf.p, stride, nextJob.p);
    checkLaunch(ctx, "wfaKernel");
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
}

// test hook: exact edit distance of byte strings (values are compared as they are)
__global__ void debugEdKernel(const uint8_t* a, int n, const uint8_t* b, int m, int* wf, uint64_t stride, int* out) {
    int d = wfaEditDistance(a, n, b, m, wf, wf + stride);
    if ((threadIdx.x & 31) == 0) *out = d;
}

int debugEditDistance(fg_ctx* ctx, const uint8_t* a, int n, const uint8_t* b, int m) {
    DevBuf<uint8_t> dA(n + 1), dB(m + 1);
    const uint64_t stride = (uint64_t)n + m + 8;
    DevBuf<int> wf(2 * stride), dOut(1);
    if (n) FG_CUDA(cudaMemcpyAsync(dA.p, a, n, cudaMemcpyHostToDevice, ctx->stream));
    if (m) FG_CUDA(cudaMemcpyAsync(dB.p, b, m, cudaMemcpyHostToDevice, ctx->stream));
    debugEdKernel<<<1, 32, 0, ctx->stream>>>(dA.p, n, dB.p, m, wf.p, stride, dOut.p);
    checkLaunch(ctx, "debugEdKernel");
    int d = -1;
    FG_CUDA(cudaMemcpyAsync(&d, dOut.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    return d;
}

}  // namespace fg
