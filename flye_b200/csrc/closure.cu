// flye_b200 — the rest of OverlapContainer::findAllOverlaps on the device:
//   ensureTransitivity(false)   overlap.cpp:576-627   every overlap is also recorded, reversed, with its ext sequence
//   filterOverlaps()            overlap.cpp:681-741   per sequence: overlaps with the same ext sequence whose ends differ by less
//                                                      than k on both sequences are clustered (union-find), the best-scoring member
//                                                      of every cluster stays, the list is sorted by curBegin
// Input: the getSeqOverlaps vectors of the FORWARD sequences (what lazySeqOverlaps caches, overlap.cpp:528-574).  The cache also
// holds their complements for the reverse strands, and the closure adds the reverse of every cached overlap to its ext sequence
// (a reverse of a reverse only duplicates an overlap, and duplicates share a cluster), so the list of a sequence s is, as a set,
//     { r, complement(r), reverse(r), reverse(complement(r)) : r an input record }  restricted to the records whose cur id is s.
// Kernels: closureExpandKernel writes the four variants with the sort key (owner, ext id); one stable radix sort makes the lists and
// their per-ext groups contiguous; closureClusterKernel (one warp per sequence) labels the clusters by min-label propagation inside
// the groups and elects the best member; a second sort by (owner, curBegin) of the elected records gives the output order.
// Every output record names the input record and the variant it is (`reserved` = 4 * index + variant), so the host side can
// rebuild kmerMatches with OverlapRange::reverse / complement.
#include "ctx.cuh"

#include <cub/cub.cuh>

namespace fg {

__device__ __forceinline__ fg_overlap closureVariant(const fg_overlap& r, int v) {
    fg_overlap o = r;
    if (v & 1) {   // complement (overlap.h:118-147): the same overlap seen from the opposite strands
        o.cur_begin = r.cur_len - r.cur_end - 1; o.cur_end = r.cur_len - r.cur_begin - 1;
        o.ext_begin = r.ext_len - r.ext_end - 1; o.ext_end = r.ext_len - r.ext_begin - 1;
        o.cur_id = r.cur_id ^ 1u; o.ext_id = r.ext_id ^ 1u;
    }
    if (v & 2) {   // reverse (overlap.h:95-116): the roles of the two sequences swapped
        fg_overlap t = o;
        o.cur_id = t.ext_id; o.cur_begin = t.ext_begin; o.cur_end = t.ext_end; o.cur_len = t.ext_len;
        o.ext_id = t.cur_id; o.ext_begin = t.cur_begin; o.ext_end = t.cur_end; o.ext_len = t.cur_len;
    }
    return o;
}

// variant-major layout (j = v * n + i): a stable sort then keeps, inside a group, the own records before the reversed ones and
// lower input indices first — the order the best-member election breaks score ties with
__global__ void __launch_bounds__(256) closureExpandKernel(const fg_overlap* __restrict__ in, uint64_t n, fg_overlap* __restrict__ out,
                                                           unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < 4 * n; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = j % n; const int v = (int)(j / n);
        fg_overlap o = closureVariant(in[i], v);
        o.reserved = (uint32_t)(4 * i + v);
        out[j] = o;
        keys[j] = ((unsigned long long)o.cur_id << 32) | o.ext_id;
        idx[j] = (uint32_t)j;
    }
}

// first sorted position of every sequence id (lists are contiguous after the sort by (owner, ext))
__global__ void closureOffsetsKernel(const unsigned long long* __restrict__ keys, uint64_t m, uint32_t nSeqs, uint64_t* __restrict__ off) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nSeqs) return;
    uint64_t lo = 0, hi = m;   // first position with owner >= s
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if ((keys[mid] >> 32) < s) lo = mid + 1; else hi = mid; }
    off[s] = lo;
}

__device__ __forceinline__ bool closureRelated(const fg_overlap& a, const fg_overlap& b, int maxEndsDiff) {
    // overlap.cpp:707-715 for the ordered pair (a, b)
    const int curDiff = (a.cur_end - a.cur_begin) - (min(a.cur_end, b.cur_end) - max(a.cur_begin, b.cur_begin));
    const int extDiff = (a.ext_end - a.ext_begin) - (min(a.ext_end, b.ext_end) - max(a.ext_begin, b.ext_begin));
    return curDiff < maxEndsDiff && extDiff < maxEndsDiff;
}

// one warp per sequence: clusters inside every ext group, best member of every cluster -> sort key of the survivors
__global__ void __launch_bounds__(128) closureClusterKernel(const fg_overlap* __restrict__ recs, const uint32_t* __restrict__ order,
                                                            const unsigned long long* __restrict__ keys, const uint64_t* __restrict__ off,
                                                            uint32_t nSeqs, int maxEndsDiff, uint32_t* __restrict__ label,
                                                            unsigned long long* __restrict__ best, unsigned long long* __restrict__ outKeys) {
    const uint32_t s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= nSeqs) return;
    const int lane = threadIdx.x & 31;
    const uint64_t a = off[s], b = off[s + 1];
    for (uint64_t g0 = a; g0 < b;) {
        uint64_t g1 = g0 + 1;
        while (g1 < b && keys[g1] == keys[g0]) ++g1;   // one ext sequence
        const uint64_t g = g1 - g0;
        for (uint64_t p = g0 + lane; p < g1; p += 32) { label[p] = (uint32_t)(p - g0); best[p] = 0ULL; }
        __syncwarp();
        if (g > 1) {
            // min-label propagation over the "related" graph (either direction of overlap.cpp:707-715) until nothing changes
            for (;;) {
                bool changed = false;
                for (uint64_t p = g0 + lane; p < g1; p += 32) {
                    const fg_overlap rp = recs[order[p]];
                    uint32_t lp = label[p];
                    for (uint64_t q = g0; q < g1; ++q) {
                        if (q == p) continue;
                        const uint32_t lq = label[q];
                        if (lq >= lp) continue;
                        const fg_overlap rq = recs[order[q]];
                        if (closureRelated(rp, rq, maxEndsDiff) || closureRelated(rq, rp, maxEndsDiff)) lp = lq;
                    }
                    if (lp < label[p]) { label[p] = lp; changed = true; }
                }
                __syncwarp();
                if (!__any_sync(0xffffffffu, changed)) break;
            }
        }
        // best member of every cluster: highest score, first in list order on ties (strict > in overlap.cpp:727)
        for (uint64_t p = g0 + lane; p < g1; p += 32) {
            const fg_overlap rp = recs[order[p]];
            atomicMax(&best[g0 + label[p]], ((unsigned long long)(uint32_t)rp.score << 32) | (0xffffffffu - (uint32_t)(p - g0)));
        }
        __syncwarp();
        for (uint64_t p = g0 + lane; p < g1; p += 32) {
            const bool keep = (0xffffffffu - (uint32_t)best[g0 + label[p]]) == (uint32_t)(p - g0);
            const fg_overlap rp = recs[order[p]];
            outKeys[p] = keep ? (((unsigned long long)s << 32) | (uint32_t)rp.cur_begin) : ~0ULL;
        }
        __syncwarp();
        g0 = g1;
    }
}

__global__ void __launch_bounds__(256) closureGatherKernel(const fg_overlap* __restrict__ recs, const uint32_t* __restrict__ order, uint64_t n,
                                                           fg_overlap* __restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = recs[order[i]];
}

__global__ void closureCountKeptKernel(const unsigned long long* __restrict__ keys, uint64_t m, unsigned long long* __restrict__ nKept) {
    // keys are sorted: the survivors come first, the dropped records carry ~0
    uint64_t lo = 0, hi = m;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (keys[mid] != ~0ULL) lo = mid + 1; else hi = mid; }
    *nKept = lo;
}

void overlapsClosure(fg_ctx* ctx, const fg_overlap* records, uint64_t n, uint32_t nSeqs, int maxEndsDiff, fg_overlap_result* result) {
    if (n >= (1ULL << 30)) throw Error(FG_ERR_ARG, "fg_overlaps_closure: too many records");
    ctx->timings.clear(); ctx->timingCalls.clear();
    ctx->resOffsets.assign((size_t)nSeqs + 1, 0);
    result->n_queries = nSeqs; result->offsets = ctx->resOffsets.data(); result->overlaps = nullptr;
    result->aln_pairs = nullptr; result->n_aln_pairs = 0; result->n_hits = result->n_pairs = result->n_dp_pairs = result->n_dp_cells = 0;
    if (n == 0) { ctx->lastResult = *result; ctx->lastMaxOverlaps = -1; return; }
    for (uint64_t i = 0; i < n; ++i)
        if (records[i].cur_id >= nSeqs || records[i].ext_id >= nSeqs) throw Error(FG_ERR_ARG, "fg_overlaps_closure: sequence id out of range");
    const uint64_t m = 4 * n;
    cudaStream_t st = streamOf(ctx);
    DevBuf<fg_overlap> dIn(n), dAll(m);
    DevBuf<unsigned long long> keysA(m), keysB(m), best(m);
    DevBuf<uint32_t> idxA(m), idxB(m), label(m);
    DevBuf<uint64_t> dOff((size_t)nSeqs + 2);
    uint64_t nKept = 0;
    {
        PhaseTimer pt(ctx, "closure_expand");
        FG_CUDA(cudaMemcpyAsync(dIn.p, records, n * sizeof(fg_overlap), cudaMemcpyHostToDevice, st));
        closureExpandKernel<<<gridFor(m), 256, 0, st>>>(dIn.p, n, dAll.p, keysA.p, idxA.p);
        checkLaunch(ctx, "closureExpandKernel");
    }
    cub::DoubleBuffer<unsigned long long> dk(keysA.p, keysB.p);
    cub::DoubleBuffer<uint32_t> dv(idxA.p, idxB.p);
    {
        PhaseTimer pt(ctx, "closure_sort");
        size_t tb = 0;
        FG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (uint64_t)m, 0, 64, st));
        DevBuf<char> tmp(tb);
        FG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, (uint64_t)m, 0, 64, st));
        ctx->launches += 9;
    }
    {
        PhaseTimer pt(ctx, "closure_cluster");
        closureOffsetsKernel<<<(nSeqs + 256) / 256, 256, 0, st>>>(dk.Current(), m, nSeqs, dOff.p);
        checkLaunch(ctx, "closureOffsetsKernel");
        unsigned long long* outKeys = dk.Alternate();   // free after the sort
        closureClusterKernel<<<(nSeqs + 3) / 4, 128, 0, st>>>(dAll.p, dv.Current(), dk.Current(), dOff.p, nSeqs, maxEndsDiff, label.p, best.p, outKeys);
        checkLaunch(ctx, "closureClusterKernel");
        // survivors in (sequence, curBegin) order; the dropped records sort to the end.  The order of equal curBegin is
        // unspecified in the reference as well (std::sort, cuckoo iteration order): compare results as multisets
        cub::DoubleBuffer<unsigned long long> dk2(outKeys, dk.Current());
        cub::DoubleBuffer<uint32_t> dv2(dv.Current(), dv.Alternate());
        size_t tb = 0;
        FG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk2, dv2, (uint64_t)m, 0, 64, st));
        DevBuf<char> tmp(tb);
        FG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk2, dv2, (uint64_t)m, 0, 64, st));
        ctx->launches += 9;
        DevBuf<unsigned long long> dKept(1);
        closureCountKeptKernel<<<1, 1, 0, st>>>(dk2.Current(), m, dKept.p);
        checkLaunch(ctx, "closureCountKeptKernel");
        FG_CUDA(cudaMemcpyAsync(&nKept, dKept.p, 8, cudaMemcpyDeviceToHost, st));
        closureOffsetsKernel<<<(nSeqs + 256) / 256, 256, 0, st>>>(dk2.Current(), m, nSeqs, dOff.p);
        checkLaunch(ctx, "closureOffsetsKernel");
        FG_CUDA(cudaStreamSynchronize(st));
        DevBuf<fg_overlap> dOut(std::max<uint64_t>(nKept, 1));
        if (nKept) {
            closureGatherKernel<<<gridFor(nKept), 256, 0, st>>>(dAll.p, dv2.Current(), nKept, dOut.p);
            checkLaunch(ctx, "closureGatherKernel");
        }
        fg_overlap* hOut = ctx->compactBuffer(0, std::max<uint64_t>(nKept, 1));
        if (nKept) FG_CUDA(cudaMemcpyAsync(hOut, dOut.p, nKept * sizeof(fg_overlap), cudaMemcpyDeviceToHost, st));
        std::vector<uint64_t> hOff((size_t)nSeqs + 1);
        FG_CUDA(cudaMemcpyAsync(hOff.data(), dOff.p, ((size_t)nSeqs + 1) * 8, cudaMemcpyDeviceToHost, st));
        FG_CUDA(cudaStreamSynchronize(st));
        for (uint32_t s = 0; s <= nSeqs; ++s) ctx->resOffsets[s] = std::min<uint64_t>(hOff[s], nKept);   // (the ~0 keys of dropped records lie beyond every id)
        result->overlaps = hOut;
    }
    ctx->lastResult = *result;
    ctx->lastMaxOverlaps = -1;   // not a result fg_overlaps_refilter can work on
}

}  // namespace fg
