// flye_b200 — k-mer counting and VertexIndex construction on the device.
//
// Replaces (reference file:line):
//   KmerCounter::count / getFreq                     src/sequence/vertex_index.cpp:499-616
//   VertexIndex::yieldFrequentKmers                  vertex_index.cpp:316-358
//   VertexIndex::buildIndexUnevenCoverage            vertex_index.cpp:25-125
//   VertexIndex::filterFrequentKmers                 vertex_index.cpp:173-212
//   VertexIndex::buildIndexMinimizers                vertex_index.cpp:389-483  (yieldMinimizers kmer.h:206-262)
//
// Data flow (all HBM-bound integer work, no tensor cores):
//   K2  extractKeys      packed reads -> canonical k-mer keys (one 4/8-byte key per position, coalesced)
//   K3  radix sort + run-length encode (CUB device primitives) -> (distinct k-mer, count); histogram;
//       open-addressing table of the k-mers with count >= 2 (absent == 1)
//   K4  selectKernel     one CTA per read: frequency of every position (table probe), in-CTA radix select of
//       the reference's per-read threshold, tandem-repeat filter -> one "selected" bit per position
//   K5  emit (two passes) writes (key,(seqId,pos)) for selected positions ALREADY ordered by global position
//       (forward-strand tiles ascending, then reverse-strand tiles descending), one stable radix sort by key,
//       run-length encode -> CSR (first,size) per key, classification (repetitive / valid), hash table.
#include "ctx.cuh"

#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>

namespace fg {

static constexpr int TILE_SLOTS = 2048;   // 64 bitmap words
static constexpr uint32_t MEM_CHUNK = 32u * 1024u * 1024u / 5u;   // vertex_index.h:286 (5-byte IndexChunk)

// ------------------------------------------------------------------------------------------------
// slot space
// ------------------------------------------------------------------------------------------------
void setKmerSize(fg_ctx* ctx, int k) {
    if (k < 1 || k > 31) throw Error(FG_ERR_ARG, "k-mer size must be in [1,31]");
    if (ctx->nReads == 0) throw Error(FG_ERR_ARG, "no reads uploaded");
    if (ctx->k == k && ctx->dSlotOff.p) return;
    ctx->k = k;
    ctx->hSlotOff.assign(ctx->nReads + 1, 0);
    ctx->hTiles.clear();
    uint64_t slots = 0, kmers = 0;
    for (uint32_t i = 0; i < ctx->nReads; ++i) {
        ctx->hSlotOff[i] = slots;
        uint32_t L = ctx->hLen[i];
        uint32_t n = L > (uint32_t)k ? L - k : 0;   // kmer.h:185-198: positions 0 .. L-k-1
        for (uint32_t t = 0; t < n; t += TILE_SLOTS) ctx->hTiles.push_back(make_uint2(i, t));
        slots += (n + 31u) & ~31u;
        kmers += n;
    }
    ctx->hSlotOff[ctx->nReads] = slots;
    ctx->nSlots = slots;
    ctx->nKmers = kmers;
    ctx->dSlotOff.alloc(ctx->nReads + 1);
    FG_CUDA(cudaMemcpyAsync(ctx->dSlotOff.p, ctx->hSlotOff.data(), (ctx->nReads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->dTiles.alloc(std::max<size_t>(ctx->hTiles.size(), 1));
    if (!ctx->hTiles.empty())
        FG_CUDA(cudaMemcpyAsync(ctx->dTiles.p, ctx->hTiles.data(), ctx->hTiles.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->counted = false;
    ctx->indexed = false;
}

// tile range [lo,hi) of the reads [first, first+count)
static void tileRange(const fg_ctx* ctx, uint32_t first, uint32_t count, size_t& lo, size_t& hi) {
    auto cmp = [](const uint2& t, uint32_t r) { return t.x < r; };
    lo = std::lower_bound(ctx->hTiles.begin(), ctx->hTiles.end(), first, cmp) - ctx->hTiles.begin();
    hi = std::lower_bound(ctx->hTiles.begin(), ctx->hTiles.end(), first + count, cmp) - ctx->hTiles.begin();
}

// ------------------------------------------------------------------------------------------------
// K2: canonical k-mer extraction.  One CTA per tile of <=2048 positions of one read; the tile's packed
// words are staged in shared memory, every thread then assembles its k-mers from two 64-bit words.
// Writes are fully coalesced (thread t -> output t, t+256, ...).
// ------------------------------------------------------------------------------------------------
template <class KeyT>
__global__ void __launch_bounds__(256) extractKeysKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                         const uint32_t* __restrict__ len, const uint2* __restrict__ tiles,
                                                         const uint64_t* __restrict__ tileOut, int k, KeyT* __restrict__ keys) {
    __shared__ uint64_t sw[TILE_SLOTS / 32 + 4];
    const uint2 t = tiles[blockIdx.x];
    const uint32_t L = len[t.x], n = L - k, p0 = t.y;
    const uint32_t cnt = min((uint32_t)TILE_SLOTS, n - p0);
    const uint64_t* words = seq + wordOff[t.x];
    const uint32_t w0 = p0 >> 5, nw = ((p0 + cnt + k - 1) >> 5) - w0 + 1;
    for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = words[w0 + i];
    if (threadIdx.x == 0) sw[nw] = 0;
    __syncthreads();
    KeyT* out = keys + tileOut[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        bool rc;
        uint64_t key = canonFromWindow(windowAt(sw, (p0 & 31) + i, k), k, rc);
        out[i] = (KeyT)key;
    }
}

// freq -> #distinct histogram: shared-memory privatised low bins, global atomics above, overflow list
static constexpr int HIST_SMEM_BINS = 1024;
static constexpr int HIST_GLOBAL_BINS = 1 << 16;
__global__ void __launch_bounds__(256) histKernel(const uint32_t* __restrict__ counts, uint64_t n, unsigned long long* __restrict__ hist,
                                                  uint32_t* __restrict__ overflow, uint32_t* __restrict__ nOverflow, uint32_t overflowCap) {
    __shared__ uint32_t sh[HIST_SMEM_BINS];
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t c = counts[i];
        if (c < HIST_SMEM_BINS) atomicAdd(&sh[c], 1u);
        else if (c < HIST_GLOBAL_BINS) atomicAdd(&hist[c], 1ULL);
        else { uint32_t s = atomicAdd(nOverflow, 1u); if (s < overflowCap) overflow[s] = c; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

template <class KeyT>
__global__ void __launch_bounds__(256) buildCountTableKernel(const KeyT* __restrict__ ukeys, const uint32_t* __restrict__ counts,
                                                             uint64_t n, Table table) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t c = counts[i];
        if (c >= 2) tableInsertUnique(table, (uint64_t)ukeys[i], c);
    }
}

__global__ void fillSlotsKernel(ulonglong2* slots, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        slots[i] = make_ulonglong2(EMPTY_KEY, 0ULL);
}

static Table makeTable(fg_ctx* ctx, DevBuf<ulonglong2>& buf, uint64_t nItems) {
    uint64_t cap = 1024;
    while (cap < nItems * 2) cap <<= 1;
    buf.alloc(cap);
    fillSlotsKernel<<<(unsigned)std::min<uint64_t>((cap + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(buf.p, cap);
    checkLaunch(ctx, "fillSlotsKernel");
    Table t; t.slots = buf.p; t.mask = cap - 1;
    return t;
}

template <class KeyT>
static void countKmersT(fg_ctx* ctx) {
    const int k = ctx->k;
    const uint32_t first = ctx->shardSet ? ctx->shardFirst : 0, count = ctx->shardSet ? ctx->shardCount : ctx->nReads;
    size_t tLo, tHi;
    tileRange(ctx, first, count, tLo, tHi);
    const size_t nTiles = tHi - tLo;
    // dense output offset of every tile
    std::vector<uint64_t> hTileOut(nTiles + 1, 0);
    for (size_t t = 0; t < nTiles; ++t) {
        const uint2 tl = ctx->hTiles[tLo + t];
        uint32_t n = ctx->hLen[tl.x] - k;
        hTileOut[t + 1] = hTileOut[t] + std::min<uint32_t>(TILE_SLOTS, n - tl.y);
    }
    const uint64_t N = hTileOut[nTiles];
    if (N >= (1ULL << 31)) throw Error(FG_ERR_ARG, "more than 2^31 k-mers in one counting shard; partition the reads over more GPUs");
    ctx->hist.clear();
    ctx->nDistinct = 0;
    ctx->dCountSlots.release();
    if (N == 0) {
        if (sharded(ctx)) throw Error(FG_ERR_ARG, "a rank's read shard has no k-mers");
        ctx->countTable = makeTable(ctx, ctx->dCountSlots, 0); ctx->counted = true; return;
    }

    DevBuf<uint64_t> dTileOut(nTiles + 1);
    FG_CUDA(cudaMemcpyAsync(dTileOut.p, hTileOut.data(), (nTiles + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    DevBuf<KeyT> keysA(N), keysB(N);
    {
        PhaseTimer pt(ctx, "extract");
        extractKeysKernel<KeyT><<<(unsigned)nTiles, 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p,
                                                                           ctx->dTiles.p + tLo, dTileOut.p, k, keysA.p);
        checkLaunch(ctx, "extractKeysKernel");
    }
    cub::DoubleBuffer<KeyT> db(keysA.p, keysB.p);
    {
        PhaseTimer pt(ctx, "count_sort");
        size_t tmpBytes = 0;
        FG_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, db, (int)N, 0, 2 * k, ctx->stream));
        DevBuf<char> tmp(tmpBytes);
        FG_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmpBytes, db, (int)N, 0, 2 * k, ctx->stream));
        ctx->launches += (2 * k + 7) / 8 + 1;
    }
    KeyT* sorted = db.Current();
    KeyT* uniq = db.Alternate();   // the other buffer is free now
    DevBuf<uint32_t> counts(N);
    DevBuf<uint64_t> dRuns(1);
    uint64_t nRuns = 0;
    {
        PhaseTimer pt(ctx, "count_reduce");
        size_t tmpBytes = 0;
        FG_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tmpBytes, sorted, uniq, counts.p, dRuns.p, (int)N, ctx->stream));
        DevBuf<char> tmp(tmpBytes);
        FG_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, tmpBytes, sorted, uniq, counts.p, dRuns.p, (int)N, ctx->stream));
        ctx->launches += 2;
        FG_CUDA(cudaMemcpyAsync(&nRuns, dRuns.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaStreamSynchronize(ctx->stream));

        // multi-GPU: every rank counted its own shard of the reads; all-gather the (k-mer,count) runs over NCCL and
        // reduce them by key, so that every rank ends up with the global counts (replicated count table)
        DevBuf<char> allKeys, allCounts; DevBuf<KeyT> mKeysB, mUniq; DevBuf<uint32_t> mCountsB, mCounts;
        if (sharded(ctx)) {
            std::vector<uint64_t> offK, offC;
            allGatherV(ctx, uniq, nRuns * sizeof(KeyT), allKeys, offK);
            allGatherV(ctx, counts.p, nRuns * 4ULL, allCounts, offC);
            const uint64_t T = offC.back() / 4;
            if (T >= (1ULL << 31)) throw Error(FG_ERR_ARG, "more than 2^31 (k-mer,count) runs to merge");
            mKeysB.alloc(T); mCountsB.alloc(T); mUniq.alloc(T); mCounts.alloc(T);
            cub::DoubleBuffer<KeyT> mk((KeyT*)allKeys.p, mKeysB.p);
            cub::DoubleBuffer<uint32_t> mv((uint32_t*)allCounts.p, mCountsB.p);
            size_t tb = 0;
            FG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, mk, mv, (int)T, 0, 2 * k, ctx->stream));
            DevBuf<char> t1(tb);
            FG_CUDA(cub::DeviceRadixSort::SortPairs(t1.p, tb, mk, mv, (int)T, 0, 2 * k, ctx->stream));
            tb = 0;
            FG_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, tb, mk.Current(), mUniq.p, mv.Current(), mCounts.p, dRuns.p, cub::Sum(), (int)T, ctx->stream));
            DevBuf<char> t2(tb);
            FG_CUDA(cub::DeviceReduce::ReduceByKey(t2.p, tb, mk.Current(), mUniq.p, mv.Current(), mCounts.p, dRuns.p, cub::Sum(), (int)T, ctx->stream));
            ctx->launches += (2 * k + 7) / 8 + 3;
            FG_CUDA(cudaMemcpyAsync(&nRuns, dRuns.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        const KeyT* gUniq = sharded(ctx) ? mUniq.p : uniq;
        const uint32_t* gCounts = sharded(ctx) ? mCounts.p : counts.p;

        // histogram (vertex_index.cpp:567-576)
        DevBuf<unsigned long long> dHist(HIST_GLOBAL_BINS);
        const uint32_t ovCap = 1u << 20;
        DevBuf<uint32_t> dOv(ovCap), dNOv(1);
        FG_CUDA(cudaMemsetAsync(dHist.p, 0, dHist.bytes(), ctx->stream));
        FG_CUDA(cudaMemsetAsync(dNOv.p, 0, 4, ctx->stream));
        if (sharded(ctx)) {
            // each rank histograms its slice of the merged runs; the freq -> #k-mers histogram is then reduced over NCCL
            const uint64_t lo = nRuns * ctx->rank / ctx->nRanks, hi = nRuns * (ctx->rank + 1) / ctx->nRanks;
            if (hi > lo) {
                histKernel<<<gridFor(hi - lo), 256, 0, ctx->stream>>>(gCounts + lo, hi - lo, dHist.p, dOv.p, dNOv.p, ovCap);
                checkLaunch(ctx, "histKernel");
            }
            allReduceSumU64(ctx, dHist.p, HIST_GLOBAL_BINS);
            DevBuf<char> ovAll; std::vector<uint64_t> ovOff;
            uint32_t myOv = 0;
            FG_CUDA(cudaMemcpyAsync(&myOv, dNOv.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
            if (myOv > ovCap) throw Error(FG_ERR_INTERNAL, "k-mer histogram overflow list exhausted");
            allGatherV(ctx, dOv.p, myOv * 4ULL, ovAll, ovOff);
            const uint32_t totOv = (uint32_t)(ovOff.back() / 4);
            if (totOv > ovCap) throw Error(FG_ERR_INTERNAL, "k-mer histogram overflow list exhausted");
            if (totOv) FG_CUDA(cudaMemcpyAsync(dOv.p, ovAll.p, totOv * 4ULL, cudaMemcpyDeviceToDevice, ctx->stream));
            FG_CUDA(cudaMemcpyAsync(dNOv.p, &totOv, 4, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            histKernel<<<gridFor(nRuns), 256, 0, ctx->stream>>>(gCounts, nRuns, dHist.p, dOv.p, dNOv.p, ovCap);
            checkLaunch(ctx, "histKernel");
        }
        std::vector<unsigned long long> hHist(HIST_GLOBAL_BINS);
        uint32_t nOv = 0;
        FG_CUDA(cudaMemcpyAsync(hHist.data(), dHist.p, dHist.bytes(), cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaMemcpyAsync(&nOv, dNOv.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
        if (nOv > ovCap) throw Error(FG_ERR_INTERNAL, "k-mer histogram overflow list exhausted");
        for (int f = 1; f < HIST_GLOBAL_BINS; ++f) if (hHist[f]) ctx->hist[f] = hHist[f];
        if (nOv) {
            std::vector<uint32_t> ov(nOv);
            FG_CUDA(cudaMemcpy(ov.data(), dOv.p, nOv * 4ULL, cudaMemcpyDeviceToHost));
            for (uint32_t c : ov) ctx->hist[c] += 1;
        }
        ctx->nDistinct = nRuns;
        uint64_t n2 = nRuns - (ctx->hist.count(1) ? ctx->hist[1] : 0);
        ctx->countTable = makeTable(ctx, ctx->dCountSlots, n2);
        buildCountTableKernel<KeyT><<<gridFor(nRuns), 256, 0, ctx->stream>>>(gUniq, gCounts, nRuns, ctx->countTable);
        checkLaunch(ctx, "buildCountTableKernel");
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ctx->counted = true;
}

void countKmers(fg_ctx* ctx, int k) {
    if (k > 17) throw Error(FG_ERR_KMER_SIZE, "Can't use flat counter for k-mer size > 17");   // vertex_index.cpp:504-507
    setKmerSize(ctx, k);
    ctx->timings.clear(); ctx->timingCalls.clear();
    if (2 * k <= 32) countKmersT<uint32_t>(ctx); else countKmersT<uint64_t>(ctx);
}

// KmerCounter::getFreq for a batch (tests)
__global__ void kmerFreqKernel(const uint64_t* kmers, uint32_t n, Table table, uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t payload;
    out[i] = tableFind(table, kmers[i], payload) ? (uint32_t)payload : 0xFFFFFFFFu;   // ~0: count is 0 or 1
}

void kmerFreqQuery(fg_ctx* ctx, const uint64_t* kmers, uint32_t n, uint32_t* out) {
    if (!ctx->counted) throw Error(FG_ERR_ARG, "fg_count_kmers has not run");
    if (!n) return;
    DevBuf<uint64_t> dK(n); DevBuf<uint32_t> dO(n);
    FG_CUDA(cudaMemcpyAsync(dK.p, kmers, n * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
    kmerFreqKernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(dK.p, n, ctx->countTable, dO.p);
    checkLaunch(ctx, "kmerFreqKernel");
    FG_CUDA(cudaMemcpyAsync(out, dO.p, n * 4ULL, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ------------------------------------------------------------------------------------------------
// K4: per-read selection (yieldFrequentKmers).  One CTA per read.
//   phase 1  freq[p] = global count of the canonical k-mer (table probe; absent = 1), rc bit per position
//   phase 2  minFreq = sortedDescending[maxKmers], maxKmers = size_t(selectRate * n) in float arithmetic
//            (vertex_index.cpp:339-340) by a 4 x 8-bit MSB radix select in shared memory
//   phase 3  count the positions that need the tandem test (freq > tandemFreq)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) selectKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                    const uint32_t* __restrict__ len, const uint64_t* __restrict__ slotOff,
                                                    int k, Table countTable, float selectRate, int tandemFreq,
                                                    uint32_t readFirst, uint32_t* __restrict__ freq, uint32_t* __restrict__ rcBits,
                                                    uint32_t* __restrict__ minFreqOut, unsigned long long* __restrict__ nTandemCand) {
    const uint32_t r = readFirst + blockIdx.x;
    const uint32_t L = len[r];
    if (L <= (uint32_t)k) { if (threadIdx.x == 0) minFreqOut[r] = 0; return; }
    const uint32_t n = L - k;
    const uint64_t* words = seq + wordOff[r];
    const uint64_t base = slotOff[r];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;

    __shared__ uint32_t hist[256];
    __shared__ uint32_t sPrefix, sRemaining, sMaxF;
    if (threadIdx.x == 0) sMaxF = 0;
    __syncthreads();
    uint32_t myMax = 0;
    for (uint32_t p0 = warp * 32; p0 < n; p0 += nWarps * 32) {
        const uint32_t p = p0 + lane;
        bool rc = false;
        if (p < n) {
            uint64_t key = canonFromWindow(windowAt(words, p, k), k, rc);
            uint64_t payload;
            const uint32_t f = tableFind(countTable, key, payload) ? (uint32_t)payload : 1u;
            freq[base + p] = f;
            myMax = max(myMax, f);
        }
        uint32_t m = __ballot_sync(0xffffffffu, p < n && rc);
        if (lane == 0) rcBits[(base + p0) >> 5] = m;
    }
    myMax = __reduce_max_sync(0xffffffffu, myMax);
    if (lane == 0) atomicMax(&sMaxF, myMax);
    __syncthreads();
    const uint32_t maxF = sMaxF;

    if (threadIdx.x == 0) {
        unsigned long long rank = __float2ull_rz(__fmul_rn(selectRate, (float)n));
        if (rank >= n) rank = n - 1;   // the reference would index out of bounds for selectRate >= 1
        sPrefix = 0; sRemaining = (uint32_t)rank;
    }
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (shift > 0 && (maxF >> shift) == 0) continue;   // every value has digit 0 here: nothing to decide
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = sPrefix;
        for (uint32_t p0 = warp * 32; p0 < n; p0 += nWarps * 32) {
            const uint32_t p = p0 + lane;
            bool take = false; uint32_t digit = 0;
            if (p < n) {
                const uint32_t f = freq[base + p];
                take = shift == 24 || (f >> (shift + 8)) == (prefix >> (shift + 8));
                digit = (f >> shift) & 255u;
            }
            // warp-aggregated histogram update: frequencies are tiny and repeat, so most lanes share a bin
            const uint32_t act = __ballot_sync(0xffffffffu, take);
            if (take) {
                const uint32_t peers = __match_any_sync(act, digit);
                if (lane == __ffs(peers) - 1) atomicAdd(&hist[digit], (uint32_t)__popc(peers));
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t remaining = sRemaining, cum = 0;
            int d = 255;
            for (; d > 0; --d) { if (cum + hist[d] > remaining) break; cum += hist[d]; }
            sRemaining = remaining - cum;
            sPrefix = prefix | ((uint32_t)d << shift);
        }
        __syncthreads();
    }
    const uint32_t minFreq = sPrefix;
    if (threadIdx.x == 0) minFreqOut[r] = minFreq;

    if (tandemFreq > 0) {
        uint32_t c = 0;
        for (uint32_t p = threadIdx.x; p < n; p += blockDim.x) {
            uint32_t f = freq[base + p];
            c += (f > (uint32_t)tandemFreq && f >= minFreq);
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(nTandemCand, (unsigned long long)c);
    }
}

// localFreq of yieldFrequentKmers (vertex_index.cpp:331,346-355) restricted to the positions it can matter
// for: key (read, canonical k-mer) -> occurrences of that k-mer in that read.
__global__ void __launch_bounds__(256) tandemCountKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                         const uint32_t* __restrict__ len, const uint64_t* __restrict__ slotOff,
                                                         const uint2* __restrict__ tiles, int k, int tandemFreq,
                                                         const uint32_t* __restrict__ freq, const uint32_t* __restrict__ minFreq,
                                                         Table tandem) {
    const uint2 t = tiles[blockIdx.x];
    const uint32_t r = t.x, n = len[r] - k;
    const uint32_t cnt = min((uint32_t)TILE_SLOTS, n - t.y);
    const uint64_t* words = seq + wordOff[r];
    const uint64_t base = slotOff[r];
    const uint32_t mf = minFreq[r];
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const uint32_t p = t.y + i;
        const uint32_t f = freq[base + p];
        if (f > (uint32_t)tandemFreq && f >= mf) {
            bool rc;
            uint64_t key = canonFromWindow(windowAt(words, p, k), k, rc);
            tableAddOne(tandem, ((uint64_t)r << (2 * k)) | key);
        }
    }
}

// selected bit per position: kept by the per-read threshold, passes the global minimum frequency gate of
// pass 1 (vertex_index.cpp:51), not a tandem k-mer (:346-355)
__global__ void __launch_bounds__(256) finalizeSelectionKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                               const uint32_t* __restrict__ len, const uint64_t* __restrict__ slotOff,
                                                               const uint2* __restrict__ tiles, int k, int minCov, int tandemFreq,
                                                               const uint32_t* __restrict__ freq, const uint32_t* __restrict__ minFreq,
                                                               Table tandem, bool haveTandem, uint32_t* __restrict__ selBits) {
    const uint2 t = tiles[blockIdx.x];
    const uint32_t r = t.x, n = len[r] - k;
    const uint32_t cnt = min((uint32_t)TILE_SLOTS, n - t.y);
    const uint64_t* words = seq + wordOff[r];
    const uint64_t base = slotOff[r];
    const uint32_t mf = minFreq[r];
    const int lane = threadIdx.x & 31;
    for (uint32_t i0 = (threadIdx.x >> 5) * 32; i0 < cnt; i0 += (blockDim.x >> 5) * 32) {
        const uint32_t i = i0 + lane, p = t.y + i;
        bool sel = false;
        if (i < cnt) {
            const uint32_t f = freq[base + p];
            sel = f >= mf && f >= (uint32_t)minCov;
            if (sel && haveTandem && tandemFreq > 0 && f > (uint32_t)tandemFreq) {
                bool rc;
                uint64_t key = canonFromWindow(windowAt(words, p, k), k, rc);
                uint64_t local = 0;
                tableFind(tandem, ((uint64_t)r << (2 * k)) | key, local);
                if (local > (uint64_t)tandemFreq) sel = false;
            }
        }
        uint32_t m = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) selBits[(base + t.y + i0) >> 5] = m;
    }
}

// ------------------------------------------------------------------------------------------------
// yieldMinimizers (kmer.h:206-262).  The monotone deque with its "skip equal hashes only after an expiry" quirk is
// order dependent, but it can be restarted EXACTLY at any position p whose hash is strictly below every hash still
// in the deque: the pop-back loop then empties the deque, the state becomes {p} whatever happened before, and p is
// emitted.  (If the walk started >= window positions before p, every element the true deque could still hold lies
// inside the simulated stretch, and any of them the simulation lost to a tie skip had the hash of an element it
// kept, so the true deque is emptied as well.)  So a read is cut into chunks of MIN_CHUNK positions; the thread of a
// chunk starts MIN_WARM positions early in silent mode, locks onto the first such position and reports the
// emissions of the steps of its own chunk.  A chunk that fails to lock before its first step retries with a 4x
// longer warm-up (ultimately from position 0, which is exact by definition).
// ------------------------------------------------------------------------------------------------
static constexpr int MIN_CHUNK = 512;   // TILE_SLOTS / 4
static constexpr int MIN_WARM = 128;

__global__ void __launch_bounds__(128) minimizerKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                       const uint32_t* __restrict__ len, const uint64_t* __restrict__ slotOff,
                                                       const uint2* __restrict__ tiles, uint32_t nChunks, int k, int window,
                                                       uint32_t* __restrict__ selBits, uint32_t* __restrict__ rcBits) {
    const uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= nChunks) return;
    const uint2 tile = tiles[ci >> 2];
    const uint32_t r = tile.x;
    const int32_t n = (int32_t)(len[r] - k);
    const int32_t c0 = (int32_t)tile.y + (int32_t)(ci & 3) * MIN_CHUNK;
    if (c0 >= n) return;
    const int32_t c1 = min(n, c0 + MIN_CHUNK);
    const uint64_t* words = seq + wordOff[r];
    const uint64_t base = slotOff[r] >> 5;
    const uint64_t mask = kmerMask(k);
    constexpr int QCAP = 64;                 // window < 64
    int32_t qPos[QCAP]; uint64_t qHash[QCAP];

    for (int32_t warm = MIN_WARM;; warm *= 4) {
        const int32_t start = max(0, c0 - warm);
        bool locked = start == 0 || window == 1;
        int head = 0, size = 0;
        int32_t lastEmitted = -1;
        uint32_t selWord = 0, rcWord = 0;
        uint64_t v = windowAt(words, (uint32_t)start, k);
        bool failed = false;
        for (int32_t p = start; p < c1; ++p) {
            if (p > start) {
                const uint32_t q = (uint32_t)p + k - 1;
                const uint64_t b = (words[q >> 5] >> ((q & 31) * 2)) & 3ULL;
                v = (v >> 2) | (b << (2 * k - 2));
            }
            if (p == c0 && !locked) { failed = true; break; }
            const uint64_t f = fwdFromWindow(v, k), rcv = (~v) & mask;
            const bool isRc = rcv < f;
            if (p >= c0 && isRc) rcWord |= 1u << (p & 31);
            int32_t emit;
            if (window == 1) emit = p;
            else {
                const uint64_t h = splitmix64(isRc ? rcv : f);
                while (size > 0 && qHash[(head + size - 1) & (QCAP - 1)] > h) --size;
                if (!locked && size == 0 && p - start >= window) { locked = true; lastEmitted = -1; }
                qPos[(head + size) & (QCAP - 1)] = p; qHash[(head + size) & (QCAP - 1)] = h; ++size;
                if (qPos[head] <= p - window) {
                    while (qPos[head] <= p - window) { head = (head + 1) & (QCAP - 1); --size; }
                    while (size >= 2 && qHash[head] == qHash[(head + 1) & (QCAP - 1)]) { head = (head + 1) & (QCAP - 1); --size; }
                }
                emit = qPos[head];
            }
            if (locked && emit != lastEmitted) {
                lastEmitted = emit;
                if (p >= c0) {   // the emissions of this chunk's steps are this thread's to report
                    if ((emit >> 5) == (p >> 5)) selWord |= 1u << (emit & 31);
                    else atomicOr(&selBits[base + (emit >> 5)], 1u << (emit & 31));
                }
            }
            if (p >= c0 && ((p & 31) == 31 || p == c1 - 1)) {
                if (selWord) atomicOr(&selBits[base + (p >> 5)], selWord);
                rcBits[base + (p >> 5)] = rcWord;
                selWord = 0; rcWord = 0;
            }
        }
        if (!failed) break;
    }
}

// The same walk with the deque in REGISTERS (window < CAP): the generic kernel keeps 64 (position, hash) slots per thread
// in local memory, 1.5 MB per SM at full occupancy — every deque access is an L2 round trip and the kernel is bound by
// that latency (ncu: 54 warps stalled on the long scoreboard per issue, 17 % of the DRAM peak).  The deque's hashes are
// non-decreasing from front to back and its positions increase, so "pop_back while back.hash > h" is a prefix count,
// and pops from the front are shifts of a small array; with compile-time indices everything stays in registers.
template <int CAP>
__global__ void __launch_bounds__(128) minimizerRegKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                          const uint32_t* __restrict__ len, const uint64_t* __restrict__ slotOff,
                                                          const uint2* __restrict__ tiles, uint32_t nChunks, int k, int window,
                                                          uint32_t* __restrict__ selBits, uint32_t* __restrict__ rcBits) {
    const uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= nChunks) return;
    const uint2 tile = tiles[ci >> 2];
    const uint32_t r = tile.x;
    const int32_t n = (int32_t)(len[r] - k);
    const int32_t c0 = (int32_t)tile.y + (int32_t)(ci & 3) * MIN_CHUNK;
    if (c0 >= n) return;
    const int32_t c1 = min(n, c0 + MIN_CHUNK);
    const uint64_t* words = seq + wordOff[r];
    const uint64_t base = slotOff[r] >> 5;
    const uint64_t mask = kmerMask(k);
    uint64_t qh[CAP]; int32_t qp[CAP];   // front = slot 0

    for (int32_t warm = MIN_WARM;; warm *= 4) {
        const int32_t start = max(0, c0 - warm);
        bool locked = start == 0;
        int size = 0;
#pragma unroll
        for (int j = 0; j < CAP; ++j) { qh[j] = 0; qp[j] = 0; }
        int32_t lastEmitted = -1;
        uint32_t selWord = 0, rcWord = 0;
        uint64_t v = windowAt(words, (uint32_t)start, k);
        uint64_t wcur = words[((uint32_t)start + k) >> 5];   // the word of the next base to shift in
        bool failed = false;
        for (int32_t p = start; p < c1; ++p) {
            if (p > start) {
                const uint32_t q = (uint32_t)p + k - 1;
                if ((q & 31) == 0) wcur = words[q >> 5];
                const uint64_t b = (wcur >> ((q & 31) * 2)) & 3ULL;
                v = (v >> 2) | (b << (2 * k - 2));
            }
            if (p == c0 && !locked) { failed = true; break; }
            const uint64_t f = fwdFromWindow(v, k), rcv = (~v) & mask;
            const bool isRc = rcv < f;
            if (p >= c0 && isRc) rcWord |= 1u << (p & 31);
            const uint64_t h = splitmix64(isRc ? rcv : f);
            int cnt = 0;   // pop_back while back.hash > h
#pragma unroll
            for (int j = 0; j < CAP; ++j) cnt += (j < size && qh[j] <= h) ? 1 : 0;
            if (!locked && cnt == 0 && p - start >= window) { locked = true; lastEmitted = -1; }
#pragma unroll
            for (int j = 0; j < CAP; ++j) if (j == cnt) { qh[j] = h; qp[j] = p; }
            size = cnt + 1;
            if (qp[0] <= p - window) {
                while (qp[0] <= p - window) {
#pragma unroll
                    for (int j = 0; j + 1 < CAP; ++j) { qh[j] = qh[j + 1]; qp[j] = qp[j + 1]; }
                    --size;
                }
                while (size >= 2 && qh[0] == qh[1]) {
#pragma unroll
                    for (int j = 0; j + 1 < CAP; ++j) { qh[j] = qh[j + 1]; qp[j] = qp[j + 1]; }
                    --size;
                }
            }
            const int32_t emit = qp[0];
            if (locked && emit != lastEmitted) {
                lastEmitted = emit;
                if (p >= c0) {   // the emissions of this chunk's steps are this thread's to report
                    if ((emit >> 5) == (p >> 5)) selWord |= 1u << (emit & 31);
                    else atomicOr(&selBits[base + (emit >> 5)], 1u << (emit & 31));
                }
            }
            if (p >= c0 && ((p & 31) == 31 || p == c1 - 1)) {
                if (selWord) atomicOr(&selBits[base + (p >> 5)], selWord);
                rcBits[base + (p >> 5)] = rcWord;
                selWord = 0; rcWord = 0;
            }
        }
        if (!failed) break;
    }
}

// ------------------------------------------------------------------------------------------------
// K5: emission in global-position order.  An "etile" is (tile, strand).  For every read the forward-strand
// etiles come first with positions ascending, then the reverse-strand etiles with positions descending, so the
// emitted (seqId,pos) values are ascending and ONE stable sort by key yields lists sorted by global position
// (vertex_index.cpp:109-114 sorts each list; global position order == (seqId,pos) order).
// ------------------------------------------------------------------------------------------------
__device__ inline uint32_t etileWord(const uint32_t* __restrict__ selBits, const uint32_t* __restrict__ rcBits, uint64_t w, bool strand) {
    uint32_t s = selBits[w], r = rcBits[w];
    return strand ? (s & r) : (s & ~r);
}

__global__ void __launch_bounds__(256) emitCountKernel(const uint64_t* __restrict__ slotOff, const uint32_t* __restrict__ len,
                                                       const uint2* __restrict__ tiles, const uint32_t* __restrict__ etiles,
                                                       uint32_t nEtiles, int k, const uint32_t* __restrict__ selBits,
                                                       const uint32_t* __restrict__ rcBits, uint32_t* __restrict__ etileCount) {
    const uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= nEtiles) return;
    const int lane = threadIdx.x & 31;
    const uint32_t et = etiles[e];
    const uint2 t = tiles[et >> 1];
    const bool strand = et & 1;
    const uint32_t n = len[t.x] - k;
    const uint32_t nWords = (min((uint32_t)TILE_SLOTS, n - t.y) + 31) >> 5;
    const uint64_t w0 = (slotOff[t.x] + t.y) >> 5;
    uint32_t c = 0;
    for (uint32_t w = lane; w < nWords; w += 32) c += __popc(etileWord(selBits, rcBits, w0 + w, strand));
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) etileCount[e] = c;
}

template <class KeyT>
__global__ void __launch_bounds__(256) emitWriteKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                       const uint64_t* __restrict__ slotOff, const uint32_t* __restrict__ len,
                                                       const uint2* __restrict__ tiles, const uint32_t* __restrict__ etiles,
                                                       uint32_t nEtiles, int k, const uint32_t* __restrict__ selBits,
                                                       const uint32_t* __restrict__ rcBits, const uint64_t* __restrict__ etileOff,
                                                       KeyT* __restrict__ outKeys, uint64_t* __restrict__ outVals) {
    const uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= nEtiles) return;
    const int lane = threadIdx.x & 31;
    const uint32_t et = etiles[e];
    const uint2 t = tiles[et >> 1];
    const bool strand = et & 1;
    const uint32_t r = t.x, L = len[r], n = L - k;
    const uint32_t nWords = (min((uint32_t)TILE_SLOTS, n - t.y) + 31) >> 5;   // <= 64
    const uint64_t w0 = (slotOff[r] + t.y) >> 5;
    const uint64_t* words = seq + wordOff[r];
    // lane owns two words; in emission order: ascending for the forward strand, descending for the reverse
    uint32_t wi[2], bits[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        uint32_t ord = lane * 2 + j;                       // rank of the word in emission order
        uint32_t w = strand ? (nWords - 1 - ord) : ord;    // only meaningful if ord < nWords
        wi[j] = w;
        bits[j] = ord < nWords ? etileWord(selBits, rcBits, w0 + w, strand) : 0u;
    }
    const uint32_t mine = __popc(bits[0]) + __popc(bits[1]);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    uint64_t o = etileOff[e] + (incl - mine);
    const uint32_t seqId = 2 * r + (strand ? 1u : 0u);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        uint32_t b = bits[j];
        while (b) {
            int bit = strand ? (31 - __clz(b)) : (__ffs(b) - 1);
            b &= ~(1u << bit);
            const uint32_t p = t.y + wi[j] * 32 + bit;
            bool rc;
            uint64_t key = canonFromWindow(windowAt(words, p, k), k, rc);
            const uint32_t pos = strand ? (L - p - k) : p;   // vertex_index.cpp:78-85
            outKeys[o] = (KeyT)key;
            outVals[o] = ((uint64_t)seqId << 32) | pos;
            ++o;
        }
    }
}

// totals for filterFrequentKmers (vertex_index.cpp:175-184)
__global__ void __launch_bounds__(256) capacityTotalsKernel(const uint32_t* __restrict__ cap, uint64_t n, uint32_t minCov,
                                                            unsigned long long* __restrict__ totals) {
    unsigned long long total = 0, uniq = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t c = cap[i];
        if (c >= minCov) { total += c; ++uniq; }
    }
    for (int d = 16; d; d >>= 1) { total += __shfl_down_sync(0xffffffffu, total, d); uniq += __shfl_down_sync(0xffffffffu, uniq, d); }
    if ((threadIdx.x & 31) == 0 && uniq) { atomicAdd(&totals[0], total); atomicAdd(&totals[1], uniq); }
}

// classification of every distinct emitted key (vertex_index.cpp:188-202 repetitive set; :73-74 frequency gate
// of pass 2) + table payload.  stats: [0] repetitive keys, [1] valid entries, [2] valid-or-empty keys, [3] too-frequent flag
template <class KeyT>
__global__ void __launch_bounds__(256) classifyKernel(const KeyT* __restrict__ ukeys, const uint32_t* __restrict__ cap,
                                                      const uint64_t* __restrict__ first, uint64_t n, uint64_t repFreq,
                                                      bool minimizerMode, Table countTable, uint64_t* __restrict__ keys64,
                                                      uint64_t* __restrict__ payload, unsigned long long* __restrict__ stats) {
    unsigned long long nRep = 0, nEnt = 0, nKeys = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = (uint64_t)ukeys[i];
        const uint32_t c = cap[i];
        uint64_t pl;
        if ((uint64_t)c > repFreq) { pl = IDX_REPETITIVE; ++nRep; }
        else {
            ++nKeys;
            if ((uint64_t)c + 1 > MEM_CHUNK) atomicExch(&stats[3], 1ULL);   // allocateIndexMemory, vertex_index.cpp:372-375
            bool valid = true;
            if (!minimizerMode) {
                uint64_t f = 1;
                tableFind(countTable, key, f);
                valid = f <= repFreq;
            }
            if (valid) { pl = (first[i] << IDX_SIZE_BITS) | c; nEnt += c; }
            else pl = ~0ULL;   // key stays in the reference's table with size 0: behaves as absent
        }
        keys64[i] = key;
        payload[i] = pl;
    }
    for (int d = 16; d; d >>= 1) {
        nRep += __shfl_down_sync(0xffffffffu, nRep, d); nEnt += __shfl_down_sync(0xffffffffu, nEnt, d);
        nKeys += __shfl_down_sync(0xffffffffu, nKeys, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nRep) atomicAdd(&stats[0], nRep);
        if (nEnt) atomicAdd(&stats[1], nEnt);
        if (nKeys) atomicAdd(&stats[2], nKeys);
    }
}

__global__ void __launch_bounds__(256) insertIndexKernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ payload,
                                                         uint64_t n, Table table) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (payload[i] != ~0ULL) tableInsertUnique(table, keys[i], payload[i]);
}

__global__ void __launch_bounds__(256) unpackEntriesKernel(const uint64_t* __restrict__ vals, uint64_t n, uint2* __restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = vals[i];
        out[i] = make_uint2((uint32_t)(v >> 32), (uint32_t)v);
    }
}

// shared tail of both index builders: selected/rc bitmaps -> entries -> CSR + table
template <class KeyT>
static void buildFromSelection(fg_ctx* ctx, DevBuf<uint32_t>& rcBits, int minCov, float repeatRate, float sampleRateIn) {
    const int k = ctx->k;
    const uint32_t firstRead = ctx->shardSet ? ctx->shardFirst : 0, nReadsShard = ctx->shardSet ? ctx->shardCount : ctx->nReads;
    size_t tLo, tHi;
    tileRange(ctx, firstRead, nReadsShard, tLo, tHi);
    // etile order
    std::vector<uint32_t> hEtiles;
    hEtiles.reserve(2 * (tHi - tLo));
    for (size_t a = tLo; a < tHi;) {
        size_t b = a;
        while (b < tHi && ctx->hTiles[b].x == ctx->hTiles[a].x) ++b;
        for (size_t t = a; t < b; ++t) hEtiles.push_back((uint32_t)(t << 1));
        for (size_t t = b; t-- > a;) hEtiles.push_back((uint32_t)(t << 1) | 1u);
        a = b;
    }
    const uint32_t nEtiles = (uint32_t)hEtiles.size();
    uint64_t E = 0;
    DevBuf<KeyT> keysA, keysB;
    DevBuf<uint64_t> valsA, valsB;
    {
        PhaseTimer pt(ctx, "emit");
        if (nEtiles) {
            DevBuf<uint32_t> dEtiles(nEtiles), dCount(nEtiles);
            DevBuf<uint64_t> dOff(nEtiles + 1);
            FG_CUDA(cudaMemcpyAsync(dEtiles.p, hEtiles.data(), nEtiles * 4ULL, cudaMemcpyHostToDevice, ctx->stream));
            const unsigned grid = (nEtiles + 7) / 8;
            emitCountKernel<<<grid, 256, 0, ctx->stream>>>(ctx->dSlotOff.p, ctx->dLen.p, ctx->dTiles.p, dEtiles.p, nEtiles, k,
                                                           ctx->dSelBits.p, rcBits.p, dCount.p);
            checkLaunch(ctx, "emitCountKernel");
            size_t tmpBytes = 0;
            FG_CUDA(cudaMemsetAsync(dOff.p, 0, 8, ctx->stream));
            FG_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmpBytes, dCount.p, dOff.p + 1, (int)nEtiles, ctx->stream));
            DevBuf<char> tmp(tmpBytes);
            FG_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmpBytes, dCount.p, dOff.p + 1, (int)nEtiles, ctx->stream));
            ++ctx->launches;
            FG_CUDA(cudaMemcpyAsync(&E, dOff.p + nEtiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
            if (E >= (1ULL << 31)) throw Error(FG_ERR_ARG, "more than 2^31 index entries in one shard; partition the reads over more GPUs");
            keysA.alloc(std::max<uint64_t>(E, 1)); keysB.alloc(std::max<uint64_t>(E, 1));
            valsA.alloc(std::max<uint64_t>(E, 1)); valsB.alloc(std::max<uint64_t>(E, 1));
            emitWriteKernel<KeyT><<<grid, 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dSlotOff.p, ctx->dLen.p,
                                                                 ctx->dTiles.p, dEtiles.p, nEtiles, k, ctx->dSelBits.p, rcBits.p,
                                                                 dOff.p, keysA.p, valsA.p);
            checkLaunch(ctx, "emitWriteKernel");
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (sharded(ctx)) {
        // shards are contiguous, ascending read ranges, so concatenating the ranks' entries in rank order keeps the
        // global-position order the single stable sort below relies on
        DevBuf<char> allK, allV; std::vector<uint64_t> offK, offV;
        allGatherV(ctx, keysA.p, E * sizeof(KeyT), allK, offK);
        allGatherV(ctx, valsA.p, E * 8ULL, allV, offV);
        E = offV.back() / 8;
        if (E >= (1ULL << 31)) throw Error(FG_ERR_ARG, "more than 2^31 index entries");
        keysA.alloc(std::max<uint64_t>(E, 1)); keysB.alloc(std::max<uint64_t>(E, 1));
        valsA.alloc(std::max<uint64_t>(E, 1)); valsB.alloc(std::max<uint64_t>(E, 1));
        if (E) {
            FG_CUDA(cudaMemcpyAsync(keysA.p, allK.p, E * sizeof(KeyT), cudaMemcpyDeviceToDevice, ctx->stream));
            FG_CUDA(cudaMemcpyAsync(valsA.p, allV.p, E * 8ULL, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        // the "selected" bitmap (needed for the self-hit test of every query) is replicated by one NCCL broadcast per
        // owner; each rank owns the whole bitmap words of its reads
        std::vector<uint32_t> firsts(ctx->nRanks + 1, 0);
        {
            DevBuf<char> all; std::vector<uint64_t> off;
            DevBuf<uint32_t> mine(1);
            FG_CUDA(cudaMemcpyAsync(mine.p, &firstRead, 4, cudaMemcpyHostToDevice, ctx->stream));
            allGatherV(ctx, mine.p, 4, all, off);
            FG_CUDA(cudaMemcpyAsync(firsts.data(), all.p, ctx->nRanks * 4ULL, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
            firsts[ctx->nRanks] = ctx->nReads;
        }
        groupStart();
        for (int r = 0; r < ctx->nRanks; ++r) {
            const uint64_t w0 = ctx->hSlotOff[firsts[r]] / 32, w1 = ctx->hSlotOff[firsts[r + 1]] / 32;
            broadcastBytes(ctx, ctx->dSelBits.p + w0, (w1 - w0) * 4, r);
        }
        groupEnd();
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ctx->dEntries.release(); ctx->dIndexSlots.release(); ctx->dUKeys.release(); ctx->dUPayload.release();
    ctx->stats = fg_index_stats{};
    ctx->stats.sample_rate = sampleRateIn;
    ctx->nEntriesStored = E; ctx->nUKeys = 0;
    if (E == 0) {
        ctx->indexTable = makeTable(ctx, ctx->dIndexSlots, 0);
        ctx->dEntries.alloc(1);
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->indexed = true;
        return;
    }
    cub::DoubleBuffer<KeyT> dbK(keysA.p, keysB.p);
    cub::DoubleBuffer<uint64_t> dbV(valsA.p, valsB.p);
    {
        PhaseTimer pt(ctx, "index_sort");
        size_t tmpBytes = 0;
        FG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, dbK, dbV, (int)E, 0, 2 * k, ctx->stream));
        DevBuf<char> tmp(tmpBytes);
        FG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, dbK, dbV, (int)E, 0, 2 * k, ctx->stream));
        ctx->launches += (2 * k + 7) / 8 + 1;
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    PhaseTimer pt(ctx, "index_table");
    // entries in their final (key, global position) order
    ctx->dEntries.alloc(E);
    unpackEntriesKernel<<<gridFor(E), 256, 0, ctx->stream>>>(dbV.Current(), E, ctx->dEntries.p);
    checkLaunch(ctx, "unpackEntriesKernel");
    KeyT* sortedKeys = dbK.Current();
    KeyT* ukeys = dbK.Alternate();
    DevBuf<uint32_t> cap(E);
    DevBuf<uint64_t> dRuns(1);
    uint64_t S = 0;
    {
        size_t tmpBytes = 0;
        FG_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tmpBytes, sortedKeys, ukeys, cap.p, dRuns.p, (int)E, ctx->stream));
        DevBuf<char> tmp(tmpBytes);
        FG_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, tmpBytes, sortedKeys, ukeys, cap.p, dRuns.p, (int)E, ctx->stream));
        ctx->launches += 2;
        FG_CUDA(cudaMemcpyAsync(&S, dRuns.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    DevBuf<uint64_t> first(S);
    {
        size_t tmpBytes = 0;
        FG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, cap.p, first.p, (int)S, ctx->stream));
        DevBuf<char> tmp(tmpBytes);
        FG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, cap.p, first.p, (int)S, ctx->stream));
        ++ctx->launches;
    }
    DevBuf<unsigned long long> dTotals(2), dStats(4);
    FG_CUDA(cudaMemsetAsync(dTotals.p, 0, 16, ctx->stream));
    FG_CUDA(cudaMemsetAsync(dStats.p, 0, 32, ctx->stream));
    capacityTotalsKernel<<<gridFor(S), 256, 0, ctx->stream>>>(cap.p, S, (uint32_t)minCov, dTotals.p);
    checkLaunch(ctx, "capacityTotalsKernel");
    unsigned long long hTotals[2];
    FG_CUDA(cudaMemcpyAsync(hTotals, dTotals.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    // vertex_index.cpp:185-186, same float arithmetic (host, IEEE single, no contraction)
    const size_t totalKmers = hTotals[0], uniqueKmers = hTotals[1];
    volatile float meanFrequency = (float)totalKmers / (uniqueKmers + 1);
    volatile float repF = repeatRate * meanFrequency;
    const size_t repetitiveFrequency = (size_t)repF;
    ctx->dUKeys.alloc(S); ctx->dUPayload.alloc(S);
    classifyKernel<KeyT><<<gridFor(S), 256, 0, ctx->stream>>>(ukeys, cap.p, first.p, S, repetitiveFrequency, ctx->minimizerMode,
                                                              ctx->countTable, ctx->dUKeys.p, ctx->dUPayload.p, dStats.p);
    checkLaunch(ctx, "classifyKernel");
    unsigned long long hStats[4];
    FG_CUDA(cudaMemcpyAsync(hStats, dStats.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hStats[3]) throw Error(FG_ERR_TOO_FREQ, "k-mer is too frequent");
    ctx->indexTable = makeTable(ctx, ctx->dIndexSlots, S);
    insertIndexKernel<<<gridFor(S), 256, 0, ctx->stream>>>(ctx->dUKeys.p, ctx->dUPayload.p, S, ctx->indexTable);
    checkLaunch(ctx, "insertIndexKernel");
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->nUKeys = S;
    ctx->stats.n_repetitive = hStats[0];
    ctx->stats.n_entries = hStats[1];
    ctx->stats.n_keys = hStats[2];
    ctx->stats.repetitive_frequency = repetitiveFrequency;
    ctx->stats.mean_frequency = meanFrequency;
    if (ctx->minimizerMode) {
        volatile float rate = (float)ctx->totalBases / (size_t)hStats[1];   // vertex_index.cpp:480-482
        ctx->stats.sample_rate = rate;
    }
    ctx->indexed = true;
}

void buildIndexSolid(fg_ctx* ctx, int minFreq, float selectRate, int tandemFreq, float repeatRate, float sampleRate) {
    if (!ctx->counted) throw Error(FG_ERR_ARG, "fg_count_kmers must run before fg_build_index_solid");
    ctx->timings.clear(); ctx->timingCalls.clear();
    ctx->minimizerMode = false;
    const int k = ctx->k;
    const uint32_t firstRead = ctx->shardSet ? ctx->shardFirst : 0, nReadsShard = ctx->shardSet ? ctx->shardCount : ctx->nReads;
    const uint64_t nWords = ctx->nSlots / 32 + 1;
    DevBuf<uint32_t> rcBits(nWords);
    ctx->dSelBits.alloc(nWords);
    FG_CUDA(cudaMemsetAsync(rcBits.p, 0, nWords * 4, ctx->stream));
    FG_CUDA(cudaMemsetAsync(ctx->dSelBits.p, 0, nWords * 4, ctx->stream));
    size_t tLo, tHi;
    tileRange(ctx, firstRead, nReadsShard, tLo, tHi);
    {
        PhaseTimer pt(ctx, "select");
        DevBuf<uint32_t> freq(std::max<uint64_t>(ctx->nSlots, 1));
        DevBuf<uint32_t> minFreqR(ctx->nReads);
        DevBuf<unsigned long long> dCand(1);
        FG_CUDA(cudaMemsetAsync(dCand.p, 0, 8, ctx->stream));
        FG_CUDA(cudaMemsetAsync(minFreqR.p, 0, ctx->nReads * 4ULL, ctx->stream));
        if (nReadsShard) {
            selectKernel<<<nReadsShard, 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p, ctx->dSlotOff.p, k,
                                                              ctx->countTable, selectRate, tandemFreq, firstRead, freq.p,
                                                              rcBits.p, minFreqR.p, dCand.p);
            checkLaunch(ctx, "selectKernel");
        }
        unsigned long long nCand = 0;
        FG_CUDA(cudaMemcpyAsync(&nCand, dCand.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
        DevBuf<ulonglong2> tandemSlots;
        Table tandem{};
        if (nCand && tHi > tLo) {
            if ((uint64_t)ctx->nReads >= (1ULL << (64 - 2 * k))) throw Error(FG_ERR_ARG, "too many reads for the tandem-filter key");
            tandem = makeTable(ctx, tandemSlots, nCand);
            tandemCountKernel<<<(unsigned)(tHi - tLo), 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p, ctx->dSlotOff.p,
                                                                             ctx->dTiles.p + tLo, k, tandemFreq, freq.p, minFreqR.p, tandem);
            checkLaunch(ctx, "tandemCountKernel");
        }
        if (tHi > tLo) {
            finalizeSelectionKernel<<<(unsigned)(tHi - tLo), 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p,
                                                                                   ctx->dSlotOff.p, ctx->dTiles.p + tLo, k, minFreq,
                                                                                   tandemFreq, freq.p, minFreqR.p, tandem, nCand != 0,
                                                                                   ctx->dSelBits.p);
            checkLaunch(ctx, "finalizeSelectionKernel");
        }
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (2 * k <= 32) buildFromSelection<uint32_t>(ctx, rcBits, minFreq, repeatRate, sampleRate);
    else buildFromSelection<uint64_t>(ctx, rcBits, minFreq, repeatRate, sampleRate);
}

void buildIndexMinimizers(fg_ctx* ctx, int k, int minCov, int window, float repeatRate) {
    if (window < 1) throw Error(FG_ERR_ARG, "wrong minimizer length");   // kmer.h:208
    if (window >= 64) throw Error(FG_ERR_ARG, "minimizer window must be < 64");
    setKmerSize(ctx, k);
    ctx->timings.clear(); ctx->timingCalls.clear();
    ctx->minimizerMode = true;
    const uint32_t firstRead = ctx->shardSet ? ctx->shardFirst : 0, nReadsShard = ctx->shardSet ? ctx->shardCount : ctx->nReads;
    const uint64_t nWords = ctx->nSlots / 32 + 1;
    DevBuf<uint32_t> rcBits(nWords);
    ctx->dSelBits.alloc(nWords);
    FG_CUDA(cudaMemsetAsync(rcBits.p, 0, nWords * 4, ctx->stream));
    FG_CUDA(cudaMemsetAsync(ctx->dSelBits.p, 0, nWords * 4, ctx->stream));
    {
        PhaseTimer pt(ctx, "select");
        if (nReadsShard) {
            size_t tLo, tHi;
            tileRange(ctx, firstRead, nReadsShard, tLo, tHi);
            const uint32_t nChunks = (uint32_t)(tHi - tLo) * (TILE_SLOTS / MIN_CHUNK);
            if (nChunks) {
                static const bool regDeque = getenv("FG_MINIMIZER_GENERIC") == nullptr;
                auto kern = (regDeque && window >= 2 && window < 15) ? minimizerRegKernel<16> : minimizerKernel;
                kern<<<(nChunks + 127) / 128, 128, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p, ctx->dSlotOff.p,
                                                                    ctx->dTiles.p + tLo, nChunks, k, window, ctx->dSelBits.p, rcBits.p);
                checkLaunch(ctx, "minimizerKernel");
            }
        }
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (2 * k <= 32) buildFromSelection<uint32_t>(ctx, rcBits, minCov, repeatRate, 1.0f);
    else buildFromSelection<uint64_t>(ctx, rcBits, minCov, repeatRate, 1.0f);
}

}  // namespace fg
