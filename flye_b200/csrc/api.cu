// flye_b200 — the extern "C" boundary (include/flye_b200.h).  Every entry point locks the context,
// selects its device, converts fg::Error into a status code and keeps the message for fg_last_error.
#include "ctx.cuh"

#include <dlfcn.h>

using fg::DevBuf;
using fg::Error;

namespace {

template <class F>
int guarded(fg_ctx* ctx, F&& body) {
    if (!ctx) return FG_ERR_ARG;
    std::lock_guard<std::mutex> lock(ctx->mtx);
    try {
        FG_CUDA(cudaSetDevice(ctx->device));
        fg::currentArena() = &ctx->arena;
        body();
        return FG_OK;
    } catch (const Error& e) {
        ctx->lastError = e.what();
        cudaGetLastError();
        return e.code;
    } catch (const std::exception& e) {
        ctx->lastError = e.what();
        return FG_ERR_INTERNAL;
    }
}

// K1: ASCII -> 2-bit packing, one thread per output word, 32 coalesced byte loads each
__global__ void __launch_bounds__(256) packAsciiKernel(const char* __restrict__ bases, const uint64_t* __restrict__ baseOff,
                                                       const uint64_t* __restrict__ wordOff, const uint32_t* __restrict__ wordRead,
                                                       uint64_t nWords, uint64_t* __restrict__ packed, uint32_t* __restrict__ bad) {
    uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w >= nWords) return;
    const uint32_t r = wordRead[w];
    const uint64_t local = w - wordOff[r];
    const uint64_t b0 = baseOff[r] + local * 32, bEnd = baseOff[r + 1];
    uint64_t out = 0;
    for (int j = 0; j < 32 && b0 + j < bEnd; ++j) {
        const char c = bases[b0 + j] & ~0x20;
        uint64_t code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
        if (code == 4) { atomicAdd(bad, 1u); code = 0; }
        out |= code << (2 * j);
    }
    packed[w] = out;
}

void installReads(fg_ctx* ctx, uint32_t n) {
    ctx->hpcReads.valid = false;
    ctx->nReads = n;
    ctx->hBasePrefix.assign(n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) ctx->hBasePrefix[i + 1] = ctx->hBasePrefix[i] + ctx->hLen[i];
    ctx->totalBases = ctx->hBasePrefix[n];
    if (2 * ctx->totalBases >= (1ULL << 40)) throw Error(FG_ERR_OVERFLOW, "Input overflow");   // sequence_container.cpp:386-391
    ctx->dWordOff.alloc(n + 1);
    ctx->dLen.alloc(std::max<uint32_t>(n, 1));
    FG_CUDA(cudaMemcpyAsync(ctx->dWordOff.p, ctx->hWordOff.data(), (n + 1) * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
    if (n) FG_CUDA(cudaMemcpyAsync(ctx->dLen.p, ctx->hLen.data(), n * 4ULL, cudaMemcpyHostToDevice, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->k = 0; ctx->counted = false; ctx->indexed = false;
    // the hit budget of fg_overlaps_batch comes from the free device memory (cudaMemGetInfo: up to a few ms): keep it while the
    // read set keeps its size, recompute when a differently sized one arrives
    if (ctx->budgetBases == 0 || ctx->totalBases > ctx->budgetBases + ctx->budgetBases / 4 || ctx->totalBases + ctx->totalBases / 4 < ctx->budgetBases) {
        ctx->hitBudget = 0; ctx->budgetBases = ctx->totalBases;
    }
    ctx->dSlotOff.release();
    ctx->shardSet = false;
}

__global__ void lookupKernel(const uint64_t* __restrict__ kmers, uint32_t n, int k, fg::Table index, uint8_t* rep, uint32_t* size,
                             uint64_t* first, uint8_t* rc) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // kmers arrive in the reference's representation (first base most significant)
    uint64_t f = kmers[i] & fg::kmerMask(k);
    uint64_t r = (~fg::rev2(f << (64 - 2 * k))) & fg::kmerMask(k);
    bool flag = r < f;
    uint64_t canon = flag ? r : f, payload;
    rep[i] = 0; size[i] = 0; first[i] = 0; rc[i] = flag;
    if (fg::tableFind(index, canon, payload)) {
        uint64_t s = payload & fg::IDX_SIZE_MASK;
        if (s == fg::IDX_REPETITIVE) rep[i] = 1;
        else { size[i] = (uint32_t)s; first[i] = payload >> fg::IDX_SIZE_BITS; }
    }
}

}  // namespace

extern "C" {

int fg_ctx_create(int device, fg_ctx** out) {
    if (!out) return FG_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return FG_ERR_CUDA;   // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return FG_ERR_CUDA;
    // FG_L2_FETCH=32|64|128: L2 fetch granularity hint
    {
        const char* e = getenv("FG_L2_FETCH");
        const size_t g = e ? (size_t)atoi(e) : 0;   // (measured: no effect on these kernels on B200 — a DRAM access moves ~128 B either way; left as a switch)
        if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
        cudaGetLastError();
    }
    fg_ctx* ctx = new fg_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return FG_ERR_CUDA; }

    // an allocation of the context's own arena that fails makes the lanes give their cached workspaces back (they are idle
    // outside fg_overlaps_batch)
    ctx->arena.onPressure = [ctx] {
        std::lock_guard<std::mutex> lk(ctx->pressureMutex);
        for (auto& l : ctx->lanes) l->arena.trim();
    };
    *out = ctx;
    return FG_OK;
}

void fg_ctx_destroy(fg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStream_t s = ctx->stream;
    cudaDeviceSynchronize();
    try { fg::commDestroy(ctx); } catch (...) {}
    fg::currentArena() = &ctx->arena;
    ctx->lanes.clear();   // (their buffers are all released by now; frees the lanes' cached blocks and streams)
    delete ctx;
    fg::currentArena() = nullptr;
    cudaStreamDestroy(s);
}

const char* fg_last_error(const fg_ctx* ctx) { return ctx ? ctx->lastError.c_str() : "null context"; }
void* fg_stream(const fg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t fg_kernel_launches(const fg_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int fg_last_timings(const fg_ctx* ctx, const char** names, float* ms, int* calls, int cap) {
    if (!ctx) return 0;
    int n = 0;
    for (size_t i = 0; i < ctx->timings.size() && n < cap; ++i, ++n) {
        names[n] = ctx->timings[i].first.c_str(); ms[n] = ctx->timings[i].second;
        if (calls) calls[n] = ctx->timingCalls[i];
    }
    return n;
}

int fg_reads_upload(fg_ctx* ctx, const uint64_t* packed, const uint64_t* wordOffsets, const uint32_t* lengths, uint32_t n) {
    return guarded(ctx, [&] {
        if (n && (!packed || !wordOffsets || !lengths)) throw Error(FG_ERR_ARG, "null input");
        ctx->hLen.assign(lengths, lengths + n);
        ctx->hWordOff.assign(n + 1, 0);
        // re-lay the reads word aligned and contiguous on the device (the caller's layout may have gaps)
        for (uint32_t i = 0; i < n; ++i) ctx->hWordOff[i + 1] = ctx->hWordOff[i] + (lengths[i] + 31) / 32;
        const uint64_t nWords = ctx->hWordOff[n];
        ctx->dSeq.alloc(nWords + 4);
        FG_CUDA(cudaMemsetAsync(ctx->dSeq.p + nWords, 0, 4 * 8, ctx->stream));
        bool contiguous = true;
        for (uint32_t i = 0; i < n && contiguous; ++i) contiguous = wordOffsets[i] - wordOffsets[0] == ctx->hWordOff[i];
        if (contiguous && n) {
            FG_CUDA(cudaMemcpyAsync(ctx->dSeq.p, packed + wordOffsets[0], nWords * 8, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            for (uint32_t i = 0; i < n; ++i)
                FG_CUDA(cudaMemcpyAsync(ctx->dSeq.p + ctx->hWordOff[i], packed + wordOffsets[i],
                                        (ctx->hWordOff[i + 1] - ctx->hWordOff[i]) * 8, cudaMemcpyHostToDevice, ctx->stream));
        }
        installReads(ctx, n);
    });
}

int fg_queries_upload(fg_ctx* ctx, const uint64_t* packed, const uint64_t* wordOffsets, const uint32_t* lengths, uint32_t n) {
    return guarded(ctx, [&] {
        if (n && (!packed || !wordOffsets || !lengths)) throw Error(FG_ERR_ARG, "null input");
        ctx->hQsLen.assign(lengths, lengths + n);
        std::vector<uint64_t> off(n + 1, 0);
        for (uint32_t i = 0; i < n; ++i) off[i + 1] = off[i] + (lengths[i] + 31) / 32;
        ctx->dQsSeq.alloc(off[n] + 4);
        ctx->dQsWordOff.alloc(n + 1);
        ctx->dQsLen.alloc(std::max<uint32_t>(n, 1));
        FG_CUDA(cudaMemsetAsync(ctx->dQsSeq.p + off[n], 0, 4 * 8, ctx->stream));
        for (uint32_t i = 0; i < n; ++i)
            if (off[i + 1] > off[i])
                FG_CUDA(cudaMemcpyAsync(ctx->dQsSeq.p + off[i], packed + wordOffsets[i], (off[i + 1] - off[i]) * 8, cudaMemcpyHostToDevice,
                                        ctx->stream));
        FG_CUDA(cudaMemcpyAsync(ctx->dQsWordOff.p, off.data(), (n + 1) * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
        if (n) FG_CUDA(cudaMemcpyAsync(ctx->dQsLen.p, lengths, n * 4ULL, cudaMemcpyHostToDevice, ctx->stream));
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->nQsReads = n;
        ctx->nQsWords = off[n];
        ctx->hpcQueries.valid = false;
    });
}

int fg_reads_upload_ascii(fg_ctx* ctx, const char* bases, const uint64_t* baseOffsets, uint32_t n) {
    return guarded(ctx, [&] {
        if (n && (!bases || !baseOffsets)) throw Error(FG_ERR_ARG, "null input");
        ctx->hLen.resize(n);
        ctx->hWordOff.assign(n + 1, 0);
        for (uint32_t i = 0; i < n; ++i) {
            uint64_t L = baseOffsets[i + 1] - baseOffsets[i];
            if (L >= (1ULL << 31)) throw Error(FG_ERR_ARG, "read longer than 2^31");
            ctx->hLen[i] = (uint32_t)L;
            ctx->hWordOff[i + 1] = ctx->hWordOff[i] + (L + 31) / 32;
        }
        const uint64_t nWords = ctx->hWordOff[n], nBases = n ? baseOffsets[n] - baseOffsets[0] : 0;
        ctx->dSeq.alloc(nWords + 4);
        FG_CUDA(cudaMemsetAsync(ctx->dSeq.p + nWords, 0, 4 * 8, ctx->stream));
        if (nWords) {
            std::vector<uint32_t> wordRead(nWords);
            std::vector<uint64_t> relOff(n + 1);
            for (uint32_t i = 0; i <= n; ++i) relOff[i] = baseOffsets[i] - baseOffsets[0];
            for (uint32_t i = 0; i < n; ++i)
                for (uint64_t w = ctx->hWordOff[i]; w < ctx->hWordOff[i + 1]; ++w) wordRead[w] = i;
            DevBuf<char> dBases(nBases);
            DevBuf<uint64_t> dBaseOff(n + 1), dWordOff(n + 1);
            DevBuf<uint32_t> dWordRead(nWords), dBad(1);
            FG_CUDA(cudaMemcpyAsync(dBases.p, bases + baseOffsets[0], nBases, cudaMemcpyHostToDevice, ctx->stream));
            FG_CUDA(cudaMemcpyAsync(dBaseOff.p, relOff.data(), (n + 1) * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
            FG_CUDA(cudaMemcpyAsync(dWordOff.p, ctx->hWordOff.data(), (n + 1) * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
            FG_CUDA(cudaMemcpyAsync(dWordRead.p, wordRead.data(), nWords * 4ULL, cudaMemcpyHostToDevice, ctx->stream));
            FG_CUDA(cudaMemsetAsync(dBad.p, 0, 4, ctx->stream));
            packAsciiKernel<<<(unsigned)((nWords + 255) / 256), 256, 0, ctx->stream>>>(dBases.p, dBaseOff.p, dWordOff.p, dWordRead.p, nWords,
                                                                                       ctx->dSeq.p, dBad.p);
            fg::checkLaunch(ctx, "packAsciiKernel");
            uint32_t bad = 0;
            FG_CUDA(cudaMemcpyAsync(&bad, dBad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
            FG_CUDA(cudaStreamSynchronize(ctx->stream));
            if (bad) throw Error(FG_ERR_ARG, "reads contain letters other than ACGT (the reference substitutes rand() bases, "
                                             "sequence_container.cpp:318-328; substitute on the host before upload)");
        }
        installReads(ctx, n);
    });
}

int fg_count_kmers(fg_ctx* ctx, int k, uint64_t* nDistinct) {
    return guarded(ctx, [&] {
        fg::countKmers(ctx, k);
        if (nDistinct) *nDistinct = ctx->nDistinct;
    });
}

int fg_kmer_hist(fg_ctx* ctx, uint64_t* freqs, uint64_t* counts, uint64_t* nBins) {
    return guarded(ctx, [&] {
        if (!ctx->counted) throw Error(FG_ERR_ARG, "fg_count_kmers has not run");
        if (nBins) *nBins = ctx->hist.size();
        if (freqs && counts) { size_t i = 0; for (const auto& kv : ctx->hist) { freqs[i] = kv.first; counts[i] = kv.second; ++i; } }
    });
}

int fg_kmer_freq(fg_ctx* ctx, const uint64_t* kmers, uint32_t n, uint32_t* out) {
    return guarded(ctx, [&] { fg::kmerFreqQuery(ctx, kmers, n, out); });
}

int fg_build_index_solid(fg_ctx* ctx, int minFreq, float selectRate, int tandemFreq, float repeatRate, float sampleRate, fg_index_stats* stats) {
    return guarded(ctx, [&] {
        fg::buildIndexSolid(ctx, minFreq, selectRate, tandemFreq, repeatRate, sampleRate);
        if (stats) *stats = ctx->stats;
    });
}

int fg_build_index_minimizers(fg_ctx* ctx, int k, int minCov, int window, float repeatRate, fg_index_stats* stats) {
    return guarded(ctx, [&] {
        fg::buildIndexMinimizers(ctx, k, minCov, window, repeatRate);
        if (stats) *stats = ctx->stats;
    });
}

int fg_index_clear(fg_ctx* ctx) {
    return guarded(ctx, [&] {
        ctx->dEntries.release(); ctx->dIndexSlots.release(); ctx->dUKeys.release(); ctx->dUPayload.release();
        ctx->dCountSlots.release(); ctx->dCount16.release(); ctx->dSolidBitsAll.release(); ctx->dDense.release(); ctx->dSolidBits.release(); ctx->dIdxBits.release(); ctx->counts = fg::CountView{}; ctx->dSelBits.release();
        ctx->indexed = false; ctx->counted = false;
    });
}

int fg_index_lookup(fg_ctx* ctx, const uint64_t* kmers, uint32_t n, uint8_t* isRep, uint32_t* size, uint64_t* first, uint8_t* revComp) {
    return guarded(ctx, [&] {
        if (!ctx->indexed) throw Error(FG_ERR_ARG, "no index");
        if (!n) return;
        DevBuf<uint64_t> dK(n), dF(n); DevBuf<uint8_t> dR(n), dC(n); DevBuf<uint32_t> dS(n);
        FG_CUDA(cudaMemcpyAsync(dK.p, kmers, n * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
        lookupKernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(dK.p, n, ctx->k, ctx->indexTable, dR.p, dS.p, dF.p, dC.p);
        fg::checkLaunch(ctx, "lookupKernel");
        if (isRep) FG_CUDA(cudaMemcpyAsync(isRep, dR.p, n, cudaMemcpyDeviceToHost, ctx->stream));
        if (size) FG_CUDA(cudaMemcpyAsync(size, dS.p, n * 4ULL, cudaMemcpyDeviceToHost, ctx->stream));
        if (first) FG_CUDA(cudaMemcpyAsync(first, dF.p, n * 8ULL, cudaMemcpyDeviceToHost, ctx->stream));
        if (revComp) FG_CUDA(cudaMemcpyAsync(revComp, dC.p, n, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int fg_index_positions(fg_ctx* ctx, uint64_t first, uint32_t n, int revComp, uint32_t* seqIds, int32_t* positions) {
    return guarded(ctx, [&] {
        if (!ctx->indexed) throw Error(FG_ERR_ARG, "no index");
        if (first + n > ctx->nEntriesStored) throw Error(FG_ERR_ARG, "position range outside the list store");
        if (!n) return;
        std::vector<uint2> h(n);
        FG_CUDA(cudaMemcpyAsync(h.data(), ctx->dEntries.p + first, n * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t id = h[i].x; int32_t pos = (int32_t)h[i].y;
            if (revComp) { pos = (int32_t)ctx->hLen[id >> 1] - pos - ctx->k; id ^= 1u; }   // vertex_index.h:166-173
            seqIds[i] = id; positions[i] = pos;
        }
    });
}

int fg_index_export(fg_ctx* ctx, uint64_t* keys, uint8_t* isRep, uint64_t* first, uint32_t* size, uint64_t* nKeys,
                    uint32_t* entrySeqIds, int32_t* entryPos, uint64_t* nEntries) {
    return guarded(ctx, [&] {
        if (!ctx->indexed) throw Error(FG_ERR_ARG, "no index");
        const uint64_t S = ctx->nUKeys, E = ctx->nEntriesStored;
        if (nKeys) *nKeys = S;
        if (nEntries) *nEntries = E;
        FG_CUDA(cudaStreamSynchronize(ctx->stream));
        if (keys && S) {
            std::vector<uint64_t> pl(S), ks(S);
            FG_CUDA(cudaMemcpy(ks.data(), ctx->dUKeys.p, S * 8, cudaMemcpyDeviceToHost));
            FG_CUDA(cudaMemcpy(pl.data(), ctx->dUPayload.p, S * 8, cudaMemcpyDeviceToHost));
            // keys ascending: already so on one GPU; with the index sorted in per-owner shares (multi-GPU) order them here
            std::vector<uint64_t> order(S);
            for (uint64_t i = 0; i < S; ++i) order[i] = i;
            if (!std::is_sorted(ks.begin(), ks.end())) std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return ks[a] < ks[b]; });
            for (uint64_t j = 0; j < S; ++j) {
                const uint64_t i = order[j];
                keys[j] = ks[i];
                const bool absent = pl[i] == ~0ULL;
                const uint64_t s = pl[i] & fg::IDX_SIZE_MASK;
                if (isRep) isRep[j] = !absent && s == fg::IDX_REPETITIVE;
                if (size) size[j] = (absent || s == fg::IDX_REPETITIVE) ? 0 : (uint32_t)s;
                if (first) first[j] = (absent || s == fg::IDX_REPETITIVE) ? 0 : pl[i] >> fg::IDX_SIZE_BITS;
            }
        }
        if (entrySeqIds && entryPos && E) {
            std::vector<uint2> h(E);
            FG_CUDA(cudaMemcpy(h.data(), ctx->dEntries.p, E * sizeof(uint2), cudaMemcpyDeviceToHost));
            for (uint64_t i = 0; i < E; ++i) { entrySeqIds[i] = h[i].x; entryPos[i] = (int32_t)h[i].y; }
        }
    });
}

int fg_overlaps_batch(fg_ctx* ctx, const uint32_t* queryIds, uint32_t n, const fg_overlap_params* params, fg_overlap_result* result) {
    return guarded(ctx, [&] {
        if (!params || !result || (n && !queryIds)) throw Error(FG_ERR_ARG, "null argument");
        fg::overlapsBatch(ctx, queryIds, n, *params, result);
    });
}

int fg_overlaps_refilter(fg_ctx* ctx, uint32_t firstQuery, float maxDivergence, fg_overlap_result* result) {
    return guarded(ctx, [&] {
        if (!result) throw Error(FG_ERR_ARG, "null argument");
        fg::overlapsRefilter(ctx, firstQuery, maxDivergence, result);
    });
}

int fg_overlaps_closure(fg_ctx* ctx, const fg_overlap* records, uint64_t n, uint32_t nSeqs, int32_t maxEndsDiff, fg_overlap_result* result) {
    return guarded(ctx, [&] {
        if (!result || (n && !records)) throw Error(FG_ERR_ARG, "null argument");
        fg::overlapsClosure(ctx, records, n, nSeqs, maxEndsDiff, result);
    });
}

int fg_debug_edit_distance(fg_ctx* ctx, const uint8_t* a, int n, const uint8_t* b, int m, int* distance) {
    return guarded(ctx, [&] {
        if (n < 0 || m < 0 || !distance) throw Error(FG_ERR_ARG, "bad argument");
        *distance = fg::debugEditDistance(ctx, a, n, b, m, 0, 0);
    });
}

int fg_debug_edit_distance_rc(fg_ctx* ctx, const uint8_t* a, int n, int rc_a, const uint8_t* b, int m, int rc_b, int* distance) {
    return guarded(ctx, [&] {
        if (n < 0 || m < 0 || !distance) throw Error(FG_ERR_ARG, "bad argument");
        *distance = fg::debugEditDistance(ctx, a, n, b, m, rc_a, rc_b);
    });
}

int fg_align_cigar_batch(fg_ctx* ctx, const uint8_t* targets, const uint64_t* target_offsets, const uint8_t* queries, const uint64_t* query_offsets,
                       uint32_t n_pairs, uint32_t cigar_cap, uint32_t* cigars, uint32_t* n_cigar, int32_t* status) {
    return guarded(ctx, [&] {
        if (n_pairs && (!target_offsets || !query_offsets || !cigars || !n_cigar || !status)) throw Error(FG_ERR_ARG, "bad argument");
        fg::alignCigarBatch(ctx, targets, target_offsets, queries, query_offsets, n_pairs, cigar_cap, cigars, n_cigar, status);
    });
}

int fg_debug_int_peak(fg_ctx* ctx, double* gops) {
    return guarded(ctx, [&] {
        if (!gops) throw Error(FG_ERR_ARG, "null argument");
        *gops = fg::intPeak(ctx);
    });
}

int fg_debug_warp_sort(fg_ctx* ctx, uint64_t* keys, uint32_t* vals, const uint64_t* segOffsets, uint32_t nSegs) {
    return guarded(ctx, [&] { fg::debugWarpSort(ctx, keys, vals, segOffsets, nSegs); });
}

// ---- multi-GPU ------------------------------------------------------------------------------------------
int fg_comm_unique_id(uint8_t id[FG_NCCL_ID_BYTES]) {
    try { fg::commUniqueId(id); return FG_OK; } catch (const Error& e) { return e.code; }
}
int fg_comm_init(fg_ctx* ctx, int nRanks, int rank, const uint8_t id[FG_NCCL_ID_BYTES]) {
    return guarded(ctx, [&] { fg::commInit(ctx, nRanks, rank, id); });
}
int fg_comm_set_shard(fg_ctx* ctx, uint32_t firstRead, uint32_t nReads) {
    return guarded(ctx, [&] {
        if ((uint64_t)firstRead + nReads > ctx->nReads) throw Error(FG_ERR_ARG, "shard outside the read set");
        ctx->shardFirst = firstRead; ctx->shardCount = nReads; ctx->shardSet = true;
    });
}

}  // extern "C"
