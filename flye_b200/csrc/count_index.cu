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
// K2 + K3: k-mer counting (KmerCounter::count, vertex_index.cpp:499-590) with a DENSE counter array in HBM.
//
// The reference counts in a flat array over all 4^k k-mers (4-bit saturating counters + an overflow map) and refuses
// k > 17 for that reason (vertex_index.cpp:504-507).  The same idea maps well onto 180 GB of HBM3e: one 32-bit counter
// per canonical k-mer class, incremented with a fire-and-forget RED.ADD per read position — no key array is ever
// materialised, nothing is sorted, and the footprint does not depend on the number of reads.
//
// Dense index of the class {f, rc(f)}:
//   odd k   exactly one of f / rc(f) has A or C as its MIDDLE base (the middle base of the reverse complement is the
//           complement of the middle base); that one is the representative and the high bit of its middle base (always
//           0) is squeezed out: 4^k / 2 counters (k = 15: 2.1 GB, k = 17: 34 GB)
//   even k  min(f, rc(f)): 4^k counters (k = 16: 17 GB)
// Multi-GPU: the class with dense index i is OWNED by rank i % N and counted there in slot i / N (< 2^32 for N >= 2), so
// the array, its clearing and its scan all shrink with the rank count.  Every rank partitions the positions of its read
// shard by owner (two passes over the packed reads: totals, then scatter of the 32-bit local slots), one grouped
// ncclSend/ncclRecv all-to-all moves them over NVLink, and the owner counts what it receives.  The scan then yields, per
// rank: its part of the freq -> #k-mers histogram (reduced with ncclAllReduce) and a 16-bit saturated copy of its counters;
// the copies are all-gathered, so that every rank can look up any class by direct indexing, as on one GPU (CountView::count16).
// ------------------------------------------------------------------------------------------------
static constexpr int MAX_RANKS = 64;
static bool envInt01(const char* name, int dflt) { const char* e = getenv(name); return (e ? atoi(e) : dflt) != 0; }

struct DenseMap {
    int k;
    uint32_t nOwners, owner;      // owner: this rank
    uint32_t ownerShift;          // log2(nOwners) if it is a power of two, else 0xffffffff
};
static inline uint64_t denseSpace(int k) { return denseSpaceOf(k); }

__device__ __forceinline__ void ownerOf(const DenseMap& m, uint64_t idx, uint32_t& owner, uint64_t& local) {
    if (m.ownerShift != 0xffffffffu) { owner = (uint32_t)(idx & (m.nOwners - 1u)); local = idx >> m.ownerShift; }
    else { local = idx / m.nOwners; owner = (uint32_t)(idx - local * m.nOwners); }
}

// One occurrence of class i.  With the "seen" bitmap (1 bit per class, L2-resident: 64 MB for k = 15) the FIRST occurrence only sets
// its bit — an atomic that never leaves the L2 — and the counter array in HBM holds count - 1: the singletons, more than half of the
// positions of a noisy read set, cost no DRAM traffic at all (a RED on a counter that misses the L2 moves ~100 B).
__device__ __forceinline__ void countOne(uint64_t i, uint32_t* __restrict__ dense, uint32_t* __restrict__ seenBits) {
    if (seenBits) {
        const uint32_t bit = 1u << (i & 31);
        if (!(atomicOr(&seenBits[i >> 5], bit) & bit)) return;
    }
    atomicAdd(&dense[i], 1u);
}

// single GPU: one CTA per tile of <= 2048 positions of one read; the tile's packed words are staged in shared memory,
// every position becomes one RED.ADD.U32 on its class counter
__global__ void __launch_bounds__(256) denseCountKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                        const uint32_t* __restrict__ len, const uint2* __restrict__ tiles, int k,
                                                        uint32_t* __restrict__ dense, uint32_t* __restrict__ seenBits) {
    __shared__ uint64_t sw[TILE_SLOTS / 32 + 4];
    const uint2 t = tiles[blockIdx.x];
    const uint32_t L = len[t.x], n = L - k, p0 = t.y;
    const uint32_t cnt = min((uint32_t)TILE_SLOTS, n - p0);
    const uint64_t* words = seq + wordOff[t.x];
    const uint32_t w0 = p0 >> 5, nw = ((p0 + cnt + k - 1) >> 5) - w0 + 1;
    for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = words[w0 + i];
    if (threadIdx.x == 0) sw[nw] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) countOne(denseIndexFromWindow(windowAt(sw, (p0 & 31) + i, k), k), dense, seenBits);
}

// multi-GPU pass A: how many positions of this rank's tiles belong to every owner
__global__ void __launch_bounds__(256) ownerTotalsKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                         const uint32_t* __restrict__ len, const uint2* __restrict__ tiles, DenseMap m,
                                                         unsigned long long* __restrict__ totals) {
    __shared__ uint64_t sw[TILE_SLOTS / 32 + 4];
    __shared__ uint32_t cnt[MAX_RANKS];
    const int k = m.k;
    const uint2 t = tiles[blockIdx.x];
    const uint32_t L = len[t.x], n = L - k, p0 = t.y;
    const uint32_t c = min((uint32_t)TILE_SLOTS, n - p0);
    const uint64_t* words = seq + wordOff[t.x];
    const uint32_t w0 = p0 >> 5, nw = ((p0 + c + k - 1) >> 5) - w0 + 1;
    for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = words[w0 + i];
    if (threadIdx.x == 0) sw[nw] = 0;
    if (threadIdx.x < MAX_RANKS) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
        uint32_t o; uint64_t local;
        ownerOf(m, denseIndexFromWindow(windowAt(sw, (p0 & 31) + i, k), k), o, local);
        atomicAdd(&cnt[o], 1u);
    }
    __syncthreads();
    if (threadIdx.x < m.nOwners && cnt[threadIdx.x]) atomicAdd(&totals[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
}

// multi-GPU pass B: the local slots go to the send segment of their owner (order inside a segment is irrelevant: the
// receiver only counts); a CTA reserves its share of every segment with one atomic per owner
__global__ void __launch_bounds__(256) ownerScatterKernel(const uint64_t* __restrict__ seq, const uint64_t* __restrict__ wordOff,
                                                          const uint32_t* __restrict__ len, const uint2* __restrict__ tiles, DenseMap m,
                                                          const unsigned long long* __restrict__ segStart, unsigned long long* __restrict__ cursor,
                                                          uint32_t* __restrict__ sendBuf) {
    __shared__ uint64_t sw[TILE_SLOTS / 32 + 4];
    __shared__ uint32_t cnt[MAX_RANKS];
    __shared__ unsigned long long base[MAX_RANKS];
    constexpr int PER = TILE_SLOTS / 256;
    const int k = m.k;
    const uint2 t = tiles[blockIdx.x];
    const uint32_t L = len[t.x], n = L - k, p0 = t.y;
    const uint32_t c = min((uint32_t)TILE_SLOTS, n - p0);
    const uint64_t* words = seq + wordOff[t.x];
    const uint32_t w0 = p0 >> 5, nw = ((p0 + c + k - 1) >> 5) - w0 + 1;
    for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = words[w0 + i];
    if (threadIdx.x == 0) sw[nw] = 0;
    if (threadIdx.x < MAX_RANKS) cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t own[PER], slot[PER], rk[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint32_t i = threadIdx.x + j * 256;
        own[j] = 0xffffffffu;
        if (i < c) {
            uint64_t local;
            ownerOf(m, denseIndexFromWindow(windowAt(sw, (p0 & 31) + i, k), k), own[j], local);
            slot[j] = (uint32_t)local;
            rk[j] = atomicAdd(&cnt[own[j]], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < m.nOwners && cnt[threadIdx.x])
        base[threadIdx.x] = segStart[threadIdx.x] + atomicAdd(&cursor[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j)
        if (own[j] != 0xffffffffu) sendBuf[base[own[j]] + rk[j]] = slot[j];
}

__global__ void __launch_bounds__(256) denseCountListKernel(const uint32_t* __restrict__ slots, uint64_t n, uint32_t* __restrict__ dense,
                                                            uint32_t* __restrict__ seenBits) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        countOne(slots[i], dense, seenBits);
}

// scan of the counters, pass 1: freq -> #k-mers histogram (vertex_index.cpp:567-576; shared-memory privatised low bins, global
// atomics above, overflow list beyond 2^16), number of distinct k-mers (hist[0]) and of k-mers with count >= 2 (*nSolid)
static constexpr int HIST_SMEM_BINS = 1024;
static constexpr int HIST_GLOBAL_BINS = 1 << 16;
__global__ void __launch_bounds__(256) denseHistKernel(const uint4* __restrict__ dense4, uint64_t n8, const uint8_t* __restrict__ seenBytes,
                                                       uint8_t* __restrict__ solidBytes, unsigned long long* __restrict__ hist,
                                                       uint32_t* __restrict__ overflow, uint32_t* __restrict__ nOverflow, uint32_t overflowCap,
                                                       unsigned long long* __restrict__ nSolid, uint4* __restrict__ count16,
                                                       uint32_t* __restrict__ overflowSlot) {
    __shared__ uint32_t sh[HIST_SMEM_BINS];
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint32_t distinct = 0, solid = 0;
    // one step = 8 consecutive counters = one byte of either bitmap
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n8; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t seen = seenBytes ? seenBytes[g] : 0u;
        uint32_t solidByte = 0;
        if (seenBytes && !seen) {   // (a counter is only ever touched after its bit)
            if (solidBytes) solidBytes[g] = 0;
            if (count16) count16[g] = make_uint4(0, 0, 0, 0);
            continue;
        }
        const uint4 q0 = dense4[2 * g], q1 = dense4[2 * g + 1];
        const uint32_t cs[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint32_t h16[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t c = cs[j] + ((seen >> j) & 1u);   // with the bitmap the array holds count - 1
            h16[j] = min(c, 65535u);
            if (!c) continue;
            ++distinct;
            if (c == 1u) continue;            // counted below as distinct - everything else
            ++solid; solidByte |= 1u << j;
            if (c < HIST_SMEM_BINS) atomicAdd(&sh[c], 1u);
            else if (c < HIST_GLOBAL_BINS) atomicAdd(&hist[c], 1ULL);
            else { const uint32_t s = atomicAdd(nOverflow, 1u); if (s < overflowCap) { overflow[s] = c; if (overflowSlot) overflowSlot[s] = (uint32_t)(8 * g + j); } }
        }
        if (solidBytes) solidBytes[g] = (uint8_t)solidByte;
        if (count16) count16[g] = make_uint4(h16[0] | (h16[1] << 16), h16[2] | (h16[3] << 16), h16[4] | (h16[5] << 16), h16[6] | (h16[7] << 16));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
    distinct = __reduce_add_sync(0xffffffffu, distinct);
    solid = __reduce_add_sync(0xffffffffu, solid);
    if ((threadIdx.x & 31) == 0) {
        if (distinct) { atomicAdd(&hist[0], (unsigned long long)distinct); atomicAdd(&hist[1], (unsigned long long)(distinct - solid)); }
        if (solid) atomicAdd(nSolid, (unsigned long long)solid);
    }
}

__global__ void __launch_bounds__(256) buildCountTableKernel(const uint64_t* __restrict__ ukeys, const uint32_t* __restrict__ counts,
                                                             uint64_t n, Table table) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        tableInsertUnique(table, ukeys[i], counts[i]);
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

void countKmers(fg_ctx* ctx, int k) {
    if (k > 17) throw Error(FG_ERR_KMER_SIZE, "Can't use flat counter for k-mer size > 17");   // vertex_index.cpp:504-507
    setKmerSize(ctx, k);
    ctx->timings.clear(); ctx->timingCalls.clear();
    const bool multi = sharded(ctx);
    if (multi && ctx->nRanks > MAX_RANKS) throw Error(FG_ERR_ARG, "too many ranks");
    const uint32_t first = ctx->shardSet ? ctx->shardFirst : 0, count = ctx->shardSet ? ctx->shardCount : ctx->nReads;
    size_t tLo, tHi;
    tileRange(ctx, first, count, tLo, tHi);
    const size_t nTiles = tHi - tLo;
    ctx->hist.clear();
    ctx->nDistinct = 0;
    ctx->dCountSlots.release(); ctx->dCount16.release(); ctx->dSolidBitsAll.release(); ctx->dDense.release(); ctx->dSolidBits.release();
    ctx->counts = CountView{};
    // counting invalidates the index (setKmerSize): give its memory back before the counters are allocated
    ctx->dEntries.release(); ctx->dIndexSlots.release(); ctx->dUKeys.release(); ctx->dUPayload.release();

    DenseMap m; m.k = k; m.nOwners = multi ? (uint32_t)ctx->nRanks : 1u; m.owner = multi ? (uint32_t)ctx->rank : 0u;
    m.ownerShift = 0xffffffffu;
    for (uint32_t s = 0; s < 31; ++s) if ((1u << s) == m.nOwners) m.ownerShift = s;
    // counters of this rank: slots of the classes i with i % nOwners == owner, padded to whole 16-byte groups
    const uint64_t nLocal = ((denseSpace(k) + m.nOwners - 1) / m.nOwners + 31) & ~31ULL;
    DevBuf<uint32_t>& dense = ctx->dDense;
    dense.alloc(nLocal);
    // bitmap fronts (see countOne): only while a bitmap fits the L2 comfortably (2^29 classes = 64 MB: k <= 15, or more ranks)
    const bool front = nLocal <= (1ULL << 29) && envInt01("FG_COUNT_FRONT", 1);
    DevBuf<uint32_t> seenBits;
    ctx->dSolidBits.release();
    if (front) { seenBits.alloc(nLocal / 32); ctx->dSolidBits.alloc(nLocal / 32); }
    {
        PhaseTimer pt(ctx, "count_clear");
        FG_CUDA(cudaMemsetAsync(dense.p, 0, nLocal * 4ULL, ctx->stream));
        if (front) FG_CUDA(cudaMemsetAsync(seenBits.p, 0, nLocal / 8, ctx->stream));
    }
    std::unique_ptr<L2Pin> pinSeen(front ? new L2Pin(ctx, seenBits.p, nLocal / 8) : nullptr);
    if (!multi) {
        PhaseTimer pt(ctx, "count");
        if (nTiles) {
            denseCountKernel<<<(unsigned)nTiles, 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p, ctx->dTiles.p + tLo, k, dense.p, seenBits.p);
            checkLaunch(ctx, "denseCountKernel");
        }
    } else {
        const int R = ctx->nRanks;
        if (nLocal > (1ULL << 32)) throw Error(FG_ERR_ARG, "dense k-mer slots do not fit 32 bits");   // (cannot happen for k <= 17, N >= 2: k = 17 on two ranks uses exactly 2^32 slots)
        DevBuf<unsigned long long> dTot(3 * MAX_RANKS);   // totals, segment starts, cursors
        std::vector<unsigned long long> hTot(R, 0), hStart(R, 0);
        DevBuf<uint32_t> sendBuf, recvBuf;
        std::vector<uint64_t> sendOffs(R + 1, 0), recvOffs;
        {
            PhaseTimer pt(ctx, "count_partition");
            FG_CUDA(cudaMemsetAsync(dTot.p, 0, dTot.bytes(), ctx->stream));
            if (nTiles) {
                ownerTotalsKernel<<<(unsigned)nTiles, 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p, ctx->dTiles.p + tLo, m, dTot.p);
                checkLaunch(ctx, "ownerTotalsKernel");
            }
            FG_CUDA(cudaMemcpyAsync(hTot.data(), dTot.p, R * 8ULL, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int r = 0; r < R; ++r) { hStart[r] = sendOffs[r] / 4; sendOffs[r + 1] = sendOffs[r] + hTot[r] * 4ULL; }
            sendBuf.alloc(std::max<uint64_t>(sendOffs[R] / 4, 1));
            FG_CUDA(cudaMemcpyAsync(dTot.p + MAX_RANKS, hStart.data(), R * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
            if (nTiles) {
                ownerScatterKernel<<<(unsigned)nTiles, 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p, ctx->dTiles.p + tLo, m,
                                                                              dTot.p + MAX_RANKS, dTot.p + 2 * MAX_RANKS, sendBuf.p);
                checkLaunch(ctx, "ownerScatterKernel");
            }
        }
        {
            PhaseTimer pt(ctx, "count_exchange");
            exchangeSizes(ctx, sendOffs, recvOffs);
            recvBuf.alloc(std::max<uint64_t>(recvOffs[R] / 4, 1));
            allToAllVInto(ctx, sendBuf.p, sendOffs, recvBuf.p, recvOffs);
        }
        sendBuf.release();
        PhaseTimer pt(ctx, "count");
        const uint64_t nRecv = recvOffs[R] / 4;
        if (nRecv) {
            denseCountListKernel<<<gridFor(nRecv, 256, 16), 256, 0, ctx->stream>>>(recvBuf.p, nRecv, dense.p, seenBits.p);
            checkLaunch(ctx, "denseCountListKernel");
        }
    }

    // histogram (vertex_index.cpp:567-576), distinct k-mers, k-mers with count >= 2
    DevBuf<unsigned long long> dHist(HIST_GLOBAL_BINS), dSolid(2);
    const uint32_t ovCap = 1u << 20;
    DevBuf<uint32_t> dOv(ovCap), dNOv(1), dOvSlot;
    DevBuf<uint16_t> local16;   // multi-GPU: this rank's counters, saturated to 16 bits, for the all-gather below
    if (multi) { local16.alloc(nLocal); dOvSlot.alloc(ovCap); }
    unsigned long long hSolid = 0;
    uint32_t nOv = 0;
    {
        PhaseTimer pt(ctx, "count_hist");
        FG_CUDA(cudaMemsetAsync(dHist.p, 0, dHist.bytes(), ctx->stream));
        FG_CUDA(cudaMemsetAsync(dNOv.p, 0, 4, ctx->stream));
        FG_CUDA(cudaMemsetAsync(dSolid.p, 0, 16, ctx->stream));
        pinSeen.reset();
        denseHistKernel<<<gridFor(nLocal / 8, 256, 16), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(dense.p), nLocal / 8,
                                                                              reinterpret_cast<const uint8_t*>(seenBits.p),
                                                                              reinterpret_cast<uint8_t*>(ctx->dSolidBits.p), dHist.p, dOv.p, dNOv.p,
                                                                              ovCap, dSolid.p, reinterpret_cast<uint4*>(local16.p), dOvSlot.p);
        checkLaunch(ctx, "denseHistKernel");
        FG_CUDA(cudaMemcpyAsync(&hSolid, dSolid.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaMemcpyAsync(&nOv, dNOv.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    agreeOrThrow(ctx, nOv > ovCap ? FG_ERR_INTERNAL : 0, "k-mer histogram overflow list exhausted");
    std::vector<uint32_t> ov(nOv);
    if (nOv) FG_CUDA(cudaMemcpyAsync(ov.data(), dOv.p, nOv * 4ULL, cudaMemcpyDeviceToHost, ctx->stream));
    if (!multi) {
        // one GPU: the counter array IS the count structure of the selection (no table, no second pass over the counters)
        ctx->counts.dense = dense.p; ctx->counts.solidBits = ctx->dSolidBits.p; ctx->counts.k = k;
    } else {
        // Every rank needs the count of every k-mer of its reads (selection) and of its index keys (classification).  The owners'
        // counters, saturated to 16 bits, are all-gathered: 2 bytes per class (k = 15: 1 GB, k = 17: 17 GB per rank) moved once over
        // NVLink, then every lookup is a direct index as on one GPU — no table of the solid k-mers is built, nothing is inserted
        // on every rank.  The (rare) counts >= 65535 go to a small replicated table; the "at least twice" bitmaps are gathered too
        // and stay in the L2 in front of the array.
        dense.release();
        const int R = ctx->nRanks;
        {
            PhaseTimer pt(ctx, "count_merge");
            allReduceSumU64(ctx, dHist.p, HIST_GLOBAL_BINS);
            std::vector<uint64_t> offs(R + 1);
            for (int r = 0; r <= R; ++r) offs[r] = (uint64_t)r * nLocal * 2ULL;
            ctx->dCount16.alloc((uint64_t)R * nLocal);
            allGatherVInto(ctx, local16.p, offs, ctx->dCount16.p);
            local16.release();
            if (front) {
                for (int r = 0; r <= R; ++r) offs[r] = (uint64_t)r * (nLocal / 8);
                ctx->dSolidBitsAll.alloc((uint64_t)R * (nLocal / 32));
                allGatherVInto(ctx, ctx->dSolidBits.p, offs, ctx->dSolidBitsAll.p);
            }
            ctx->dSolidBits.release();
            // counters beyond 16 bits (also beyond the histogram's dense bins): (class index, count) of every rank
            std::vector<uint32_t> slots(nOv);
            if (nOv) FG_CUDA(cudaMemcpyAsync(slots.data(), dOvSlot.p, nOv * 4ULL, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
            std::vector<uint64_t> mine(2 * (size_t)nOv);
            for (uint32_t i = 0; i < nOv; ++i) { mine[2 * i] = (uint64_t)slots[i] * m.nOwners + m.owner; mine[2 * i + 1] = ov[i]; }
            DevBuf<uint64_t> dMine(std::max<size_t>(mine.size(), 1));
            if (nOv) FG_CUDA(cudaMemcpyAsync(dMine.p, mine.data(), mine.size() * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
            DevBuf<char> ovAll; std::vector<uint64_t> ovOff;
            allGatherV(ctx, dMine.p, nOv * 16ULL, ovAll, ovOff);
            const size_t nBig = ovOff.back() / 16;
            std::vector<uint64_t> big(2 * nBig);
            if (nBig) FG_CUDA(cudaMemcpyAsync(big.data(), ovAll.p, nBig * 16ULL, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
            ov.resize(nBig);
            for (size_t i = 0; i < nBig; ++i) ov[i] = (uint32_t)big[2 * i + 1];
            ctx->counts.table = makeTable(ctx, ctx->dCountSlots, nBig);
            if (nBig) {
                DevBuf<uint64_t> bk(nBig); DevBuf<uint32_t> bc(nBig);
                std::vector<uint64_t> hk(nBig); std::vector<uint32_t> hc(nBig);
                for (size_t i = 0; i < nBig; ++i) { hk[i] = big[2 * i]; hc[i] = (uint32_t)big[2 * i + 1]; }
                FG_CUDA(cudaMemcpyAsync(bk.p, hk.data(), nBig * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
                FG_CUDA(cudaMemcpyAsync(bc.p, hc.data(), nBig * 4ULL, cudaMemcpyHostToDevice, ctx->stream));
                buildCountTableKernel<<<gridFor(nBig, 256, 16), 256, 0, ctx->stream>>>(bk.p, bc.p, nBig, ctx->counts.table);
                checkLaunch(ctx, "buildCountTableKernel");
                FG_CUDA(cudaStreamSynchronize(ctx->stream));
            }
        }
        ctx->counts.count16 = ctx->dCount16.p; ctx->counts.solidBits = ctx->dSolidBitsAll.p;
        ctx->counts.nLocal = nLocal; ctx->counts.nOwners = m.nOwners; ctx->counts.ownerShift = m.ownerShift; ctx->counts.k = k;
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    std::vector<unsigned long long> hHist(HIST_GLOBAL_BINS);
    FG_CUDA(cudaMemcpyAsync(hHist.data(), dHist.p, dHist.bytes(), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->nDistinct = hHist[0];
    for (int f = 1; f < HIST_GLOBAL_BINS; ++f) if (hHist[f]) ctx->hist[f] = hHist[f];
    for (uint32_t c : ov) ctx->hist[c] += 1;
    ctx->counted = true;
}

// KmerCounter::getFreq for a batch (tests)
__global__ void kmerFreqKernel(const uint64_t* kmers, uint32_t n, CountView counts, uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = countOfKmer(counts, kmers[i] & kmerMask(counts.k));
    out[i] = c >= 2u ? c : 0xFFFFFFFFu;   // ~0: count is 0 or 1
}

void kmerFreqQuery(fg_ctx* ctx, const uint64_t* kmers, uint32_t n, uint32_t* out) {
    if (!ctx->counted) throw Error(FG_ERR_ARG, "fg_count_kmers has not run");
    if (!ctx->counts.dense && !ctx->counts.count16) throw Error(FG_ERR_ARG, "the k-mer counters were released when the index was built");
    if (!n) return;
    DevBuf<uint64_t> dK(n); DevBuf<uint32_t> dO(n);
    FG_CUDA(cudaMemcpyAsync(dK.p, kmers, n * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
    kmerFreqKernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(dK.p, n, ctx->counts, dO.p);
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
                                                    int k, CountView counts, float selectRate, int tandemFreq,
                                                    uint32_t readFirst, uint64_t slotBase, uint32_t* __restrict__ freqShard, uint32_t* __restrict__ rcBits,
                                                    uint32_t* __restrict__ minFreqOut, unsigned long long* __restrict__ nTandemCand) {
    const uint32_t r = readFirst + blockIdx.x;
    const uint32_t L = len[r];
    if (L <= (uint32_t)k) { if (threadIdx.x == 0) minFreqOut[r] = 0; return; }
    const uint32_t n = L - k;
    const uint64_t* words = seq + wordOff[r];
    const uint64_t base = slotOff[r];
    uint32_t* freq = freqShard - slotBase;   // freq[] only covers the slots of this rank's read shard: indexed by global slot below
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
            const uint64_t v = windowAt(words, p, k);
            canonFromWindow(v, k, rc);
            const uint32_t f = countOfWindow(counts, v);
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
                                                         const uint2* __restrict__ tiles, int k, int tandemFreq, uint64_t slotBase,
                                                         const uint32_t* __restrict__ freqShard, const uint32_t* __restrict__ minFreq,
                                                         Table tandem) {
    const uint32_t* freq = freqShard - slotBase;
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
                                                               const uint2* __restrict__ tiles, int k, int minCov, int tandemFreq, uint64_t slotBase,
                                                               const uint32_t* __restrict__ freqShard, const uint32_t* __restrict__ minFreq,
                                                               Table tandem, bool haveTandem, uint32_t* __restrict__ selBits) {
    const uint32_t* freq = freqShard - slotBase;
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

// run-length encoding of the sorted keys in two passes over 2048-key tiles (any number of keys; the library's encoder takes an
// int): head counts per tile, one scan of the tile counts, then every head writes its key and its start position.  The start
// positions ARE the CSR offsets ("first" of every key); the capacity of key g is starts[g+1] - starts[g].
static constexpr int RLE_PER = 8, RLE_TILE = 256 * RLE_PER;
template <class KeyT>
__global__ void __launch_bounds__(256) rleCountKernel(const KeyT* __restrict__ keys, uint64_t n, uint32_t* __restrict__ tileHeads) {
    const uint64_t i0 = (uint64_t)blockIdx.x * RLE_TILE + (uint64_t)threadIdx.x * RLE_PER;
    uint32_t c = 0;
    if (i0 < n) {
        KeyT prev = i0 ? keys[i0 - 1] : (KeyT)0;
#pragma unroll
        for (int j = 0; j < RLE_PER; ++j)
            if (i0 + j < n) { const KeyT x = keys[i0 + j]; c += (i0 + j == 0) || x != prev; prev = x; }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ uint32_t ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += ws[w]; tileHeads[blockIdx.x] = t; }
}
template <class KeyT>
__global__ void __launch_bounds__(256) rleWriteKernel(const KeyT* __restrict__ keys, uint64_t n, const uint64_t* __restrict__ tileBase,
                                                      KeyT* __restrict__ ukeys, uint64_t* __restrict__ starts) {
    const uint64_t i0 = (uint64_t)blockIdx.x * RLE_TILE + (uint64_t)threadIdx.x * RLE_PER;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    KeyT x[RLE_PER]; uint32_t heads = 0;
    if (i0 < n) {
        KeyT prev = i0 ? keys[i0 - 1] : (KeyT)0;
#pragma unroll
        for (int j = 0; j < RLE_PER; ++j)
            if (i0 + j < n) { x[j] = keys[i0 + j]; if ((i0 + j == 0) || x[j] != prev) heads |= 1u << j; prev = x[j]; }
    }
    const uint32_t mine = __popc(heads);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    __shared__ uint32_t ws[8];
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (int v = 0; v < w; ++v) before += ws[v];
    uint64_t g = tileBase[blockIdx.x] + before + (incl - mine);
#pragma unroll
    for (int j = 0; j < RLE_PER; ++j)
        if (heads & (1u << j)) { ukeys[g] = x[j]; starts[g] = i0 + j; ++g; }
}

// totals for filterFrequentKmers (vertex_index.cpp:175-184)
__global__ void __launch_bounds__(256) capacityTotalsKernel(const uint64_t* __restrict__ starts, uint64_t n, uint32_t minCov,
                                                            unsigned long long* __restrict__ totals) {
    unsigned long long total = 0, uniq = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = starts[i + 1] - starts[i];
        if (c >= minCov) { total += c; ++uniq; }
    }
    for (int d = 16; d; d >>= 1) { total += __shfl_down_sync(0xffffffffu, total, d); uniq += __shfl_down_sync(0xffffffffu, uniq, d); }
    if ((threadIdx.x & 31) == 0 && uniq) { atomicAdd(&totals[0], total); atomicAdd(&totals[1], uniq); }
}

// classification of every distinct emitted key (vertex_index.cpp:188-202 repetitive set; :73-74 frequency gate
// of pass 2) + table payload.  stats: [0] repetitive keys, [1] valid entries, [2] valid-or-empty keys, [3] too-frequent flag
// entryBase: position of this rank's first entry in the replicated entry array (multi-GPU), 0 otherwise
template <class KeyT>
__global__ void __launch_bounds__(256) classifyKernel(const KeyT* __restrict__ ukeys, const uint64_t* __restrict__ starts, uint64_t n,
                                                      uint64_t repFreq, uint64_t entryBase, bool minimizerMode, CountView counts,
                                                      uint64_t* __restrict__ keys64, uint64_t* __restrict__ payload,
                                                      unsigned long long* __restrict__ stats) {
    unsigned long long nRep = 0, nEnt = 0, nKeys = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = (uint64_t)ukeys[i];
        const uint64_t c = starts[i + 1] - starts[i];
        uint64_t pl;
        if (c > repFreq) { pl = IDX_REPETITIVE; ++nRep; }
        else {
            ++nKeys;
            if (c + 1 > MEM_CHUNK) atomicExch(&stats[3], 1ULL);   // allocateIndexMemory, vertex_index.cpp:372-375
            bool valid = true;
            if (!minimizerMode) {
                valid = (uint64_t)countOfKmer(counts, key) <= repFreq;
            }
            if (valid) { pl = ((entryBase + starts[i]) << IDX_SIZE_BITS) | c; nEnt += c; }
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

// presence bitmap of the index over the k-mer classes (dense index, kmer_math.cuh): set for every key the table holds (valid or
// repetitive).  The queries look up ALL their k-mers, and in a noisy read set most of those are erroneous and absent: the
// L2-resident bitmap answers them without a table probe in HBM (queryLookupKernel).
__global__ void __launch_bounds__(256) indexBitsKernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ payload, uint64_t n, int k,
                                                       uint32_t* __restrict__ bits) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (payload[i] != ~0ULL) {
            const uint64_t key = keys[i], d = denseIndexOfPair(key, revCompKmer(key, k), k);
            atomicOr(&bits[d >> 5], 1u << (d & 31));
        }
}

__global__ void __launch_bounds__(256) unpackEntriesKernel(const uint64_t* __restrict__ vals, uint64_t n, uint2* __restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = vals[i];
        out[i] = make_uint2((uint32_t)(v >> 32), (uint32_t)v);
    }
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU: the index entries are sorted by the rank that OWNS their key (owner = multiply-shift of a hash of the key, so
// the shares are even whatever the key distribution is).  A STABLE partition keeps the emission order — global position
// ascending — inside every share; shards are ascending read ranges, so concatenating what the ranks send in rank order keeps
// it across ranks too, and the owner's single stable sort by key yields position-sorted lists as on one GPU.
// ------------------------------------------------------------------------------------------------
static constexpr int PART_PER = 8, PART_TILE = 256 * PART_PER;
// (explicit __umulhi: written as a 64-bit product shifted by 32, nvcc 12.9 folded the index arithmetic of partCountKernel's
// shared-memory atomic into `hi * (nOwners >> 30)` — every key counted for owner 0; cuobjdump showed LEA.HI + IMAD)
__device__ __forceinline__ uint32_t indexOwner(uint64_t key, uint32_t nOwners) {
    return __umulhi((uint32_t)(mix64(key) >> 32), nOwners);
}
template <class KeyT>
__global__ void __launch_bounds__(256) partCountKernel(const KeyT* __restrict__ keys, uint64_t n, uint32_t nOwners, uint32_t nTiles,
                                                       uint32_t* __restrict__ cntT) {
    __shared__ uint32_t cnt[MAX_RANKS];
    if (threadIdx.x < MAX_RANKS) cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * PART_TILE;
    for (uint32_t i = threadIdx.x; i < PART_TILE && base + i < n; i += 256) atomicAdd(&cnt[indexOwner((uint64_t)keys[base + i], nOwners)], 1u);
    __syncthreads();
    if (threadIdx.x < nOwners) cntT[(uint64_t)threadIdx.x * nTiles + blockIdx.x] = cnt[threadIdx.x];
}
template <class KeyT>
__global__ void __launch_bounds__(256) partScatterKernel(const KeyT* __restrict__ keys, const uint64_t* __restrict__ vals, uint64_t n,
                                                         uint32_t nOwners, uint32_t nTiles, const uint64_t* __restrict__ offT,
                                                         KeyT* __restrict__ outK, uint64_t* __restrict__ outV) {
    __shared__ uint32_t wc[8][MAX_RANKS];
    __shared__ uint64_t wbase[8][MAX_RANKS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * MAX_RANKS; i += 256) (&wc[0][0])[i] = 0;
    __syncthreads();
    // warp w owns the 256 consecutive items [base, base + 256) of the tile, 32 consecutive ones per round
    const uint64_t base = (uint64_t)blockIdx.x * PART_TILE + (uint64_t)w * (32 * PART_PER);
    KeyT key[PART_PER]; uint64_t val[PART_PER]; uint32_t own[PART_PER];
#pragma unroll
    for (int r = 0; r < PART_PER; ++r) {
        const uint64_t i = base + r * 32 + lane;
        const bool act = i < n;
        own[r] = 0xffffffffu;
        if (act) { key[r] = keys[i]; val[r] = vals[i]; own[r] = indexOwner((uint64_t)key[r], nOwners); }
        const uint32_t am = __ballot_sync(0xffffffffu, act);
        if (act) {
            const uint32_t peers = __match_any_sync(am, own[r]);
            if ((peers & ((1u << lane) - 1u)) == 0u) wc[w][own[r]] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < nOwners) {
        uint64_t run = offT[(uint64_t)threadIdx.x * nTiles + blockIdx.x];
        for (int v = 0; v < 8; ++v) { wbase[v][threadIdx.x] = run; run += wc[v][threadIdx.x]; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PART_PER; ++r) {
        const bool act = own[r] != 0xffffffffu;
        const uint32_t am = __ballot_sync(0xffffffffu, act);
        uint64_t pos = 0; uint32_t peers = 0;
        if (act) {
            peers = __match_any_sync(am, own[r]);
            pos = wbase[w][own[r]] + __popc(peers & ((1u << lane) - 1u));
        }
        __syncwarp();
        if (act) {
            if ((peers & ((1u << lane) - 1u)) == 0u) wbase[w][own[r]] += __popc(peers);
            outK[pos] = key[r]; outV[pos] = val[r];
        }
        __syncwarp();
    }
}

struct CastU32ToU64 { __host__ __device__ uint64_t operator()(uint32_t x) const { return x; } };
// out[0] = 0, out[i+1] = in[0] + ... + in[i] (64-bit sums of 32-bit counts)
static void prefixU32(fg_ctx* ctx, const uint32_t* in, uint64_t n, uint64_t* out) {
    FG_CUDA(cudaMemsetAsync(out, 0, 8, ctx->stream));
    if (!n) return;
    cub::TransformInputIterator<uint64_t, CastU32ToU64, const uint32_t*> it(in, CastU32ToU64());
    size_t tmpBytes = 0;
    FG_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmpBytes, it, out + 1, (uint64_t)n, ctx->stream));
    DevBuf<char> tmp(tmpBytes);
    FG_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmpBytes, it, out + 1, (uint64_t)n, ctx->stream));
    ++ctx->launches;
}

// shared tail of both index builders: selected/rc bitmaps -> entries -> CSR + table
template <class KeyT>
static void buildFromSelection(fg_ctx* ctx, DevBuf<uint32_t>& rcBits, int minCov, float repeatRate, float sampleRateIn) {
    const int k = ctx->k;
    const bool multi = sharded(ctx);
    const int R = multi ? ctx->nRanks : 1;
    if (R > MAX_RANKS) throw Error(FG_ERR_ARG, "too many ranks");
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
            prefixU32(ctx, dCount.p, nEtiles, dOff.p);
            FG_CUDA(cudaMemcpyAsync(&E, dOff.p + nEtiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
            keysA.alloc(std::max<uint64_t>(E, 1)); valsA.alloc(std::max<uint64_t>(E, 1));
            emitWriteKernel<KeyT><<<grid, 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dSlotOff.p, ctx->dLen.p,
                                                                 ctx->dTiles.p, dEtiles.p, nEtiles, k, ctx->dSelBits.p, rcBits.p,
                                                                 dOff.p, keysA.p, valsA.p);
            checkLaunch(ctx, "emitWriteKernel");
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
        } else { keysA.alloc(1); valsA.alloc(1); }
    }
    if (multi) {
        PhaseTimer pt(ctx, "index_exchange");
        // stable partition of this rank's entries by the owner of their key, then one all-to-all
        const uint32_t nTiles = (uint32_t)((E + PART_TILE - 1) / PART_TILE);
        DevBuf<KeyT> pk(std::max<uint64_t>(E, 1)); DevBuf<uint64_t> pv(std::max<uint64_t>(E, 1));
        std::vector<uint64_t> items(R + 1, 0);
        if (E) {
            DevBuf<uint32_t> cntT((uint64_t)R * nTiles);
            DevBuf<uint64_t> offT((uint64_t)R * nTiles + 1);
            partCountKernel<KeyT><<<nTiles, 256, 0, ctx->stream>>>(keysA.p, E, (uint32_t)R, nTiles, cntT.p);
            checkLaunch(ctx, "partCountKernel");
            prefixU32(ctx, cntT.p, (uint64_t)R * nTiles, offT.p);
            partScatterKernel<KeyT><<<nTiles, 256, 0, ctx->stream>>>(keysA.p, valsA.p, E, (uint32_t)R, nTiles, offT.p, pk.p, pv.p);
            checkLaunch(ctx, "partScatterKernel");
            for (int r = 0; r <= R; ++r)
                FG_CUDA(cudaMemcpyAsync(&items[r], offT.p + (uint64_t)r * nTiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        std::vector<uint64_t> recvItems, sOff(R + 1), rOff(R + 1);
        exchangeSizes(ctx, items, recvItems);
        E = recvItems[R];
        keysA.alloc(std::max<uint64_t>(E, 1)); valsA.alloc(std::max<uint64_t>(E, 1));
        for (int r = 0; r <= R; ++r) { sOff[r] = items[r] * sizeof(KeyT); rOff[r] = recvItems[r] * sizeof(KeyT); }
        allToAllVInto(ctx, pk.p, sOff, keysA.p, rOff);
        for (int r = 0; r <= R; ++r) { sOff[r] = items[r] * 8ULL; rOff[r] = recvItems[r] * 8ULL; }
        allToAllVInto(ctx, pv.p, sOff, valsA.p, rOff);
        // the "selected" bitmap (needed for the self-hit test of every query) is replicated by one NCCL broadcast per
        // owner; each rank owns the whole bitmap words of its reads
        std::vector<uint64_t> fOff;
        DevBuf<uint32_t> mine(1), all(R);
        FG_CUDA(cudaMemcpyAsync(mine.p, &firstRead, 4, cudaMemcpyHostToDevice, ctx->stream));
        allGatherSizes(ctx, 4, fOff);
        allGatherVInto(ctx, mine.p, fOff, all.p);
        std::vector<uint32_t> firsts(R + 1, 0);
        FG_CUDA(cudaMemcpyAsync(firsts.data(), all.p, R * 4ULL, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
        firsts[R] = ctx->nReads;
        groupStart();
        for (int r = 0; r < R; ++r) {
            const uint64_t w0 = ctx->hSlotOff[firsts[r]] / 32, w1 = ctx->hSlotOff[firsts[r + 1]] / 32;
            broadcastBytes(ctx, ctx->dSelBits.p + w0, (w1 - w0) * 4, r);
        }
        groupEnd();
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ctx->dEntries.release(); ctx->dIndexSlots.release(); ctx->dUKeys.release(); ctx->dUPayload.release(); ctx->dIdxBits.release();
    ctx->stats = fg_index_stats{};
    ctx->stats.sample_rate = sampleRateIn;
    ctx->nEntriesStored = 0; ctx->nUKeys = 0;

    // this rank's entries sorted by (key, global position): one stable radix sort by key
    keysB.alloc(std::max<uint64_t>(E, 1)); valsB.alloc(std::max<uint64_t>(E, 1));
    cub::DoubleBuffer<KeyT> dbK(keysA.p, keysB.p);
    cub::DoubleBuffer<uint64_t> dbV(valsA.p, valsB.p);
    if (E) {
        PhaseTimer pt(ctx, "index_sort");
        size_t tmpBytes = 0;
        FG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, dbK, dbV, (uint64_t)E, 0, 2 * k, ctx->stream));
        DevBuf<char> tmp(tmpBytes);
        FG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, dbK, dbV, (uint64_t)E, 0, 2 * k, ctx->stream));
        ctx->launches += (2 * k + 7) / 8 + 1;
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    PhaseTimer pt(ctx, "index_table");
    KeyT* sortedKeys = dbK.Current();
    KeyT* ukeys = dbK.Alternate();
    uint64_t S = 0;
    DevBuf<uint64_t> starts;
    {
        const uint32_t nTiles = (uint32_t)((E + RLE_TILE - 1) / RLE_TILE);
        DevBuf<uint32_t> tileHeads(std::max<uint32_t>(nTiles, 1));
        DevBuf<uint64_t> tileBase(nTiles + 1);
        if (nTiles) {
            rleCountKernel<KeyT><<<nTiles, 256, 0, ctx->stream>>>(sortedKeys, E, tileHeads.p);
            checkLaunch(ctx, "rleCountKernel");
        }
        prefixU32(ctx, tileHeads.p, nTiles, tileBase.p);
        FG_CUDA(cudaMemcpyAsync(&S, tileBase.p + nTiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
        starts.alloc(S + 1);
        if (nTiles) {
            rleWriteKernel<KeyT><<<nTiles, 256, 0, ctx->stream>>>(sortedKeys, E, tileBase.p, ukeys, starts.p);
            checkLaunch(ctx, "rleWriteKernel");
        }
        FG_CUDA(cudaMemcpyAsync(starts.p + S, &E, 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    // the sorted keys have done their job (the distinct ones are in `ukeys`): give the buffer back before the outputs are allocated
    (sortedKeys == keysA.p ? keysA : keysB).release();
    // totals of filterFrequentKmers (vertex_index.cpp:175-184) over ALL keys; entry / key counts of every rank
    DevBuf<unsigned long long> dTotals(4), dStats(4);
    FG_CUDA(cudaMemsetAsync(dTotals.p, 0, 32, ctx->stream));
    FG_CUDA(cudaMemsetAsync(dStats.p, 0, 32, ctx->stream));
    if (S) {
        capacityTotalsKernel<<<gridFor(S), 256, 0, ctx->stream>>>(starts.p, S, (uint32_t)minCov, dTotals.p);
        checkLaunch(ctx, "capacityTotalsKernel");
    }
    std::vector<uint64_t> eOff(2, 0), sOffs(2, 0);
    eOff[1] = E * sizeof(uint2); sOffs[1] = S * 8ULL;
    if (multi) {
        allReduceSumU64(ctx, dTotals.p, 2);
        allGatherSizes(ctx, E * sizeof(uint2), eOff);
        allGatherSizes(ctx, S * 8ULL, sOffs);
    }
    const int me = multi ? ctx->rank : 0;
    const uint64_t Etotal = eOff.back() / sizeof(uint2), Stotal = sOffs.back() / 8, entryBase = eOff[me] / sizeof(uint2);
    unsigned long long hTotals[2];
    FG_CUDA(cudaMemcpyAsync(hTotals, dTotals.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    // vertex_index.cpp:185-186, same float arithmetic (host, IEEE single, no contraction)
    const size_t totalKmers = hTotals[0], uniqueKmers = hTotals[1];
    volatile float meanFrequency = (float)totalKmers / (uniqueKmers + 1);
    volatile float repF = repeatRate * meanFrequency;
    const size_t repetitiveFrequency = (size_t)repF;
    // keys + payloads: written in place on one GPU, gathered from the owners otherwise
    ctx->dUKeys.alloc(std::max<uint64_t>(Stotal, 1)); ctx->dUPayload.alloc(std::max<uint64_t>(Stotal, 1));
    DevBuf<uint64_t> locKeys, locPayload;
    if (multi) { locKeys.alloc(std::max<uint64_t>(S, 1)); locPayload.alloc(std::max<uint64_t>(S, 1)); }
    if (S) {
        classifyKernel<KeyT><<<gridFor(S), 256, 0, ctx->stream>>>(ukeys, starts.p, S, repetitiveFrequency, entryBase, ctx->minimizerMode,
                                                                  ctx->counts, multi ? locKeys.p : ctx->dUKeys.p, multi ? locPayload.p : ctx->dUPayload.p, dStats.p);
        checkLaunch(ctx, "classifyKernel");
    }
    if (multi) allReduceSumU64(ctx, dStats.p, 4);
    unsigned long long hStats[4];
    FG_CUDA(cudaMemcpyAsync(hStats, dStats.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    // the k-mer counters were last needed by the classification: large ones (k >= 16: 17 / 34 GB) go back before the table is built
    if (ctx->dDense.bytes() > (4ULL << 30) || ctx->dCount16.bytes() > (4ULL << 30)) {
        ctx->dDense.release(); ctx->dSolidBits.release(); ctx->dCount16.release(); ctx->dSolidBitsAll.release(); ctx->dCountSlots.release();
        ctx->counts = CountView{};
    }
    (ukeys == keysA.p ? keysA : keysB).release();
    starts.release();
    if (hStats[3]) throw Error(FG_ERR_TOO_FREQ, "k-mer is too frequent");   // (all ranks see the reduced flag: they throw together)
    // entries in their final (key, global position) order; replicated on every rank
    {
        DevBuf<uint64_t>& curV = dbV.Current() == valsA.p ? valsA : valsB;
        DevBuf<uint64_t>& altV = dbV.Current() == valsA.p ? valsB : valsA;
        uint2* locEntries = reinterpret_cast<uint2*>(altV.p);   // the other value buffer is free: the unpacked entries go there
        if (E) { unpackEntriesKernel<<<gridFor(E), 256, 0, ctx->stream>>>(curV.p, E, locEntries); checkLaunch(ctx, "unpackEntriesKernel"); }
        if (!multi) {
            ctx->dEntries = altV.reinterpretAs<uint2>();   // ... and that buffer becomes the index's entry array
        } else {
            ctx->dEntries.alloc(std::max<uint64_t>(Etotal, 1));
            allGatherVInto(ctx, locEntries, eOff, ctx->dEntries.p);
            allGatherVInto(ctx, locKeys.p, sOffs, ctx->dUKeys.p);
            allGatherVInto(ctx, locPayload.p, sOffs, ctx->dUPayload.p);
            altV.release();
        }
        curV.release();
    }
    ctx->indexTable = makeTable(ctx, ctx->dIndexSlots, Stotal);
    ctx->dIdxBits.release();
    if (denseSpace(k) <= (1ULL << 29) && envInt01("FG_INDEX_BITS", 1)) {   // 64 MB at most: stays in the L2 during the lookups
        ctx->dIdxBits.alloc(denseSpace(k) / 32 + 1);
        FG_CUDA(cudaMemsetAsync(ctx->dIdxBits.p, 0, ctx->dIdxBits.bytes(), ctx->stream));
    }
    if (Stotal) {
        insertIndexKernel<<<gridFor(Stotal, 256, 16), 256, 0, ctx->stream>>>(ctx->dUKeys.p, ctx->dUPayload.p, Stotal, ctx->indexTable);
        checkLaunch(ctx, "insertIndexKernel");
        if (ctx->dIdxBits.p) {
            indexBitsKernel<<<gridFor(Stotal, 256, 16), 256, 0, ctx->stream>>>(ctx->dUKeys.p, ctx->dUPayload.p, Stotal, k, ctx->dIdxBits.p);
            checkLaunch(ctx, "indexBitsKernel");
        }
    }
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->nEntriesStored = Etotal;
    ctx->nUKeys = Stotal;
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
        L2Pin pinSolid(ctx, ctx->counts.solidBits, ctx->dSolidBits.p ? ctx->dSolidBits.bytes() : ctx->dSolidBitsAll.bytes(), 2);   // the "count >= 2" bitmap answers most lookups: keep it in the L2
        PhaseTimer pt(ctx, "select");
        const uint64_t slotBase = ctx->hSlotOff[firstRead], shardSlots = ctx->hSlotOff[firstRead + nReadsShard] - slotBase;
        DevBuf<uint32_t> freq(std::max<uint64_t>(shardSlots, 1));
        DevBuf<uint32_t> minFreqR(ctx->nReads);
        DevBuf<unsigned long long> dCand(1);
        FG_CUDA(cudaMemsetAsync(dCand.p, 0, 8, ctx->stream));
        FG_CUDA(cudaMemsetAsync(minFreqR.p, 0, ctx->nReads * 4ULL, ctx->stream));
        if (nReadsShard) {
            selectKernel<<<nReadsShard, 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p, ctx->dSlotOff.p, k,
                                                              ctx->counts, selectRate, tandemFreq, firstRead, slotBase, freq.p,
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
                                                                             ctx->dTiles.p + tLo, k, tandemFreq, slotBase, freq.p, minFreqR.p, tandem);
            checkLaunch(ctx, "tandemCountKernel");
        }
        if (tHi > tLo) {
            finalizeSelectionKernel<<<(unsigned)(tHi - tLo), 256, 0, ctx->stream>>>(ctx->dSeq.p, ctx->dWordOff.p, ctx->dLen.p,
                                                                                   ctx->dSlotOff.p, ctx->dTiles.p + tLo, k, minFreq,
                                                                                   tandemFreq, slotBase, freq.p, minFreqR.p, tandem, nCand != 0,
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
