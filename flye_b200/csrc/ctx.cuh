// flye_b200 — device context shared by the translation units of libflye_b200.so.
#pragma once
#include "common.cuh"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <shared_mutex>
#include <thread>
#include <utility>

namespace fg {
// Persistent host worker threads of a context (the per-record epilogue of fg_overlaps_batch: divergence with glibc logf,
// threshold / maxOverlaps replay).  Created on first use; creating 2 x 16 std::threads per call cost more than the work.
class HostPool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cvStart, cvDone;
    const std::function<void(size_t, size_t)>* job = nullptr;
    size_t jobN = 0;
    unsigned generation = 0, pending = 0;
    bool stopping = false;
    void run(unsigned t) {
        unsigned seen = 0;
        for (;;) {
            const std::function<void(size_t, size_t)>* fn; size_t n; unsigned T;
            {
                std::unique_lock<std::mutex> lk(m);
                cvStart.wait(lk, [&] { return stopping || generation != seen; });
                if (stopping) return;
                seen = generation; fn = job; n = jobN; T = (unsigned)workers.size();
            }
            (*fn)(n * t / T, n * (t + 1) / T);
            std::lock_guard<std::mutex> lk(m);
            if (--pending == 0) cvDone.notify_all();
        }
    }
public:
    unsigned maxThreads = 16;   // lowered when several ranks share the host (fg_comm_init: cores / ranks); fixed once the workers exist
    // fn(a, b) over a partition of [0, n) into contiguous ranges, one per worker; returns when all are done
    // (`work`: number of elementary items behind the n ranges, decides whether threads are worth it)
    void parallelFor(size_t n, const std::function<void(size_t, size_t)>& fn, size_t work = 0) {
        const unsigned T = std::max(1u, std::min(maxThreads, std::thread::hardware_concurrency()));
        if (std::max(n, work) < 20000 || n < T || T == 1) { fn(0, n); return; }
        if (workers.empty()) for (unsigned t = 0; t < T; ++t) workers.emplace_back(&HostPool::run, this, t);
        std::unique_lock<std::mutex> lk(m);
        job = &fn; jobN = n; pending = (unsigned)workers.size(); ++generation;
        cvStart.notify_all();
        cvDone.wait(lk, [&] { return pending == 0; });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(m); stopping = true; }
        cvStart.notify_all();
        for (auto& th : workers) th.join();
    }
};
}  // namespace fg

namespace fg {
// One execution lane of fg_overlaps_batch: the sub-batches of a call are processed by a few host worker threads, each with its
// own CUDA stream, its own caching arena (a block is only ever reused by work on the stream that released it) and its own
// phase-timing list.  While one lane waits for a count to come back from the device or copies results to the host, the
// kernels of the other lanes keep the SMs busy.  Outside fg_overlaps_batch there is one implicit lane: the context itself.
struct Lane {
    Arena arena;
    cudaStream_t stream = nullptr;
    std::vector<std::pair<std::string, float>> timings;
    std::vector<int> timingCalls;
    ~Lane() { if (stream) cudaStreamDestroy(stream); }
};
inline Lane*& currentLane() { static thread_local Lane* l = nullptr; return l; }
}  // namespace fg

struct fg_ctx {
    fg::Arena arena;                      // first member: destroyed last, after every DevBuf below
    int device = 0;
    cudaStream_t stream = nullptr;
    mutable std::mutex mtx;
    std::string lastError;
    std::atomic<uint64_t> launches{0};
    std::vector<std::unique_ptr<fg::Lane>> lanes;   // worker lanes of fg_overlaps_batch (streams and arenas are kept across calls)

    // ---- reads (forward strands, 2-bit packed, every read word aligned) ----
    uint32_t nReads = 0;
    uint64_t totalBases = 0;              // over forward strands
    std::vector<uint32_t> hLen;           // per read
    std::vector<uint64_t> hWordOff;       // nReads+1
    std::vector<uint64_t> hBasePrefix;    // nReads+1, prefix of lengths (global offset of id 2i = 2*prefix[i])
    fg::DevBuf<uint64_t> dSeq;            // packed words (+2 words of slack)
    fg::DevBuf<uint64_t> dWordOff;        // nReads+1
    fg::DevBuf<uint32_t> dLen;            // nReads

    // ---- optional second sequence set: queries that are not part of the indexed reads (read-to-graph alignment,
    //      reference: OverlapContainer(detector, queryContainer) with queryContainer != the indexed container) ----
    uint32_t nQsReads = 0;
    std::vector<uint32_t> hQsLen;
    fg::DevBuf<uint64_t> dQsSeq, dQsWordOff;
    fg::DevBuf<uint32_t> dQsLen;
    uint64_t nQsWords = 0;

    // ---- homopolymer-compressed copies of the two sequence sets (editdist.cu), built on first use, dropped on upload ----
    struct HpcCache { fg::DevBuf<uint64_t> seq; fg::DevBuf<uint32_t> mask, prefix; bool valid = false; };
    HpcCache hpcReads, hpcQueries;

    // ---- k-mer slot space: read i owns slots [slotOff[i], slotOff[i]+n_i), n_i = max(L_i-k,0) (kmer.h:185-198),
    //      slotOff is a prefix of roundup32(n_i) so that every read owns whole 32-bit bitmap words ----
    int k = 0;
    uint64_t nSlots = 0;                  // size of the slot space
    uint64_t nKmers = 0;                  // sum of n_i
    std::vector<uint64_t> hSlotOff;       // nReads+1
    fg::DevBuf<uint64_t> dSlotOff;
    // tiles of <=2048 slots inside one read (64 bitmap words): tile -> (read, first position)
    std::vector<uint2> hTiles;
    fg::DevBuf<uint2> dTiles;

    // this rank's share of the reads for count / emission (multi-GPU); default = all
    uint32_t shardFirst = 0, shardCount = 0;
    bool shardSet = false;

    // ---- k-mer counts ----
    bool counted = false;
    uint64_t nDistinct = 0;
    std::map<uint64_t, uint64_t> hist;    // freq -> #distinct canonical k-mers
    fg::DevBuf<ulonglong2> dCountSlots;   // multi-GPU: replicated table class index -> count for the (rare) counts >= 65535
    fg::DevBuf<uint16_t> dCount16;        // multi-GPU: every rank's saturated 16-bit counters, all-gathered (CountView::count16)
    fg::DevBuf<uint32_t> dSolidBitsAll;   // multi-GPU: every rank's "occurs at least twice" bitmap, all-gathered
    fg::DevBuf<uint32_t> dDense;          // one GPU: the dense counter array itself (count_index.cu), kept until the index is built
    fg::DevBuf<uint32_t> dSolidBits;      // 1 bit per class: occurs at least twice (only while the bitmap fits the L2, k <= 15)
    fg::CountView counts;

    // ---- index ----
    bool indexed = false;
    bool minimizerMode = false;
    fg_index_stats stats{};
    fg::DevBuf<uint32_t> dSelBits;        // slot bitmap: position contributed an index entry candidate
    fg::DevBuf<uint2> dEntries;           // (seqId,pos) sorted by (canonical k-mer, seqId, pos)
    uint64_t nEntriesStored = 0;
    fg::DevBuf<ulonglong2> dIndexSlots;
    fg::Table indexTable;
    fg::DevBuf<uint32_t> dIdxBits;        // 1 bit per k-mer class: has an index entry or is repetitive (only while the bitmap fits the L2)
    // sorted unique keys with their class, for export / tests
    fg::DevBuf<uint64_t> dUKeys;
    fg::DevBuf<uint64_t> dUPayload;       // same encoding as the table payload; ~0 = not in index
    uint64_t nUKeys = 0;

    // ---- timings of the last call ----
    std::vector<std::pair<std::string, float>> timings;
    std::vector<int> timingCalls;         // launches of the phase's kernel (parallel to timings)

    // ---- overlap results (host, library owned) ----
    std::vector<uint64_t> resOffsets;
    std::vector<int32_t> resAln;
    std::mutex alnMutex;
    fg::PinnedBuf<fg_overlap> pinnedOut;  // D2H staging, kept across calls
    std::shared_mutex pinnedMutex;        // lanes copy into / work on their slice under a shared lock; growing the buffer takes it exclusively
    fg::HostPool hostPool;
    std::mutex hostPoolMutex;             // one parallelFor at a time (the lanes share the pool)
    std::mutex pressureMutex;             // one arena at a time asks the others to give cached blocks back
    int devEpilogueState = 0;             // device epilogue of fg_overlaps_batch: 0 = not tested yet, 1 = the device's logf matched the host's, -1 = it did not
    uint64_t hitBudget = 0;               // k-mer hits per sub-batch and lane, derived once from the free device memory
    uint64_t budgetBases = 0;          // totalBases of the read set hitBudget was computed for
    // results after the divergence / maxOverlaps filter (when it removed something); two buffers: fg_overlaps_refilter
    // compacts from the one that holds the last result into the other
    std::unique_ptr<fg_overlap[]> resCompact[2];
    size_t resCompactCap[2] = {0, 0};
    fg_overlap_result lastResult{};             // what the last fg_overlaps_batch / fg_overlaps_refilter returned
    // what the last batch was computed under (fg_overlaps_refilter may only TIGHTEN that)
    int32_t lastMaxOverlaps = 0;
    float lastMaxDivergence = 0.f;
    std::vector<float> lastQueryMaxDivergence;
    fg_overlap* compactBuffer(int which, size_t n) {
        if (resCompactCap[which] < n) { resCompactCap[which] = n + n / 4 + 16; resCompact[which].reset(new fg_overlap[resCompactCap[which]]); }
        return resCompact[which].get();
    }

    // ---- NCCL ----
    void* ncclComm = nullptr;
    int nRanks = 1, rank = 0;
};

namespace fg {

// the stream of the calling thread: its lane's inside a worker of fg_overlaps_batch, the context's otherwise
inline cudaStream_t streamOf(const fg_ctx* ctx) { Lane* l = currentLane(); return l ? l->stream : ctx->stream; }

inline void addTiming(fg_ctx* ctx, const char* name, float ms, int calls = 1) {
    Lane* l = currentLane();
    auto& t = l ? l->timings : ctx->timings;
    auto& c = l ? l->timingCalls : ctx->timingCalls;
    for (size_t i = 0; i < t.size(); ++i)
        if (t[i].first == name) { t[i].second += ms; c[i] += calls; return; }
    t.emplace_back(name, ms); c.push_back(calls);
}

// RAII phase timer on the calling thread's stream (CUDA events; the phase list is what fg_last_timings reports; with several
// lanes at work the phases of different lanes overlap in time, so their sum can exceed the wall clock of the call)
struct PhaseTimer {
    fg_ctx* ctx;
    const char* name;
    cudaEvent_t a{}, b{};
    PhaseTimer(fg_ctx* c, const char* n) : ctx(c), name(n) {
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, streamOf(ctx));
    }
    ~PhaseTimer() {
        cudaEventRecord(b, streamOf(ctx));
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        addTiming(ctx, name, ms);
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
};

// host wall-clock of a host-side section, reported next to the device phases as "host_<name>"
struct HostTimer {
    fg_ctx* ctx; const char* name; std::chrono::steady_clock::time_point t0;
    HostTimer(fg_ctx* c, const char* n) : ctx(c), name(n), t0(std::chrono::steady_clock::now()) {}
    ~HostTimer() { addTiming(ctx, name, std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count()); }
};

// Keeps a small, randomly accessed structure (a presence bitmap) resident in the L2 while the kernels of the enclosing scope
// stream gigabytes past it: persisting-L2 access policy window on the calling thread's stream (sm_80+).  Best effort.
struct L2Pin {
    cudaStream_t st; bool on = false, carved = false;
    L2Pin(const fg_ctx* ctx, const void* p, size_t bytes, int which = 1) : st(streamOf(ctx)) {
        if (!p || !bytes) return;
        if (const char* e = getenv("FG_L2_PIN")) if (!(atoi(e) & which)) return;   // A/B switch: bit 0 count, bit 1 select, bit 2 lookup
        int dev = 0, maxWin = 0, maxPersist = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        if (maxWin <= 0 || maxPersist <= 0) { cudaGetLastError(); return; }
        const size_t carve = std::min<size_t>((size_t)maxPersist, bytes);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
        carved = true;
        cudaStreamAttrValue a{};
        a.accessPolicyWindow.base_ptr = const_cast<void*>(p);
        a.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)maxWin);
        a.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)a.accessPolicyWindow.num_bytes);
        a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        on = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a) == cudaSuccess;
        cudaGetLastError();
    }
    ~L2Pin() {
        if (!carved) return;
        // the kernels launched under the window have to be done before their persisting lines are dropped; the set-aside is
        // device wide and outlives the window, so it is given back as well (left in place it halves the L2 of every later kernel:
        // measured +19 ms on the overlap phase of configs[0])
        cudaStreamSynchronize(st);
        if (on) {
            cudaStreamAttrValue a{};
            a.accessPolicyWindow.num_bytes = 0;
            cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a);
        }
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        cudaGetLastError();
    }
};

inline void checkLaunch(fg_ctx* ctx, const char* what) {
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(FG_ERR_CUDA, std::string("launch of ") + what + ": " + cudaGetErrorString(e));
}

inline int gridFor(uint64_t n, int block = 256, int maxPerSm = 8) {
    uint64_t g = (n + block - 1) / block;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(g, (uint64_t)148 * maxPerSm));
}

// phases implemented in the other translation units
void setKmerSize(fg_ctx* ctx, int k);
void countKmers(fg_ctx* ctx, int k);
void kmerFreqQuery(fg_ctx* ctx, const uint64_t* kmers, uint32_t n, uint32_t* out);
void buildIndexSolid(fg_ctx* ctx, int minFreq, float selectRate, int tandemFreq, float repeatRate, float sampleRate);
void buildIndexMinimizers(fg_ctx* ctx, int k, int minCov, int window, float repeatRate);
void overlapsBatch(fg_ctx* ctx, const uint32_t* queryIds, uint32_t nQueries, const fg_overlap_params& p,
                   fg_overlap_result* result);
void overlapsRefilter(fg_ctx* ctx, uint32_t firstQuery, float maxDivergence, fg_overlap_result* result);
void overlapsClosure(fg_ctx* ctx, const fg_overlap* records, uint64_t n, uint32_t nSeqs, int maxEndsDiff, fg_overlap_result* result);

// NCCL plumbing (comm.cu)
void commUniqueId(uint8_t* id);
void commInit(fg_ctx* ctx, int nRanks, int rank, const uint8_t* id);
void commDestroy(fg_ctx* ctx);
void allGatherV(fg_ctx* ctx, const void* src, uint64_t bytes, DevBuf<char>& out, std::vector<uint64_t>& offs);
void allGatherSizes(fg_ctx* ctx, uint64_t bytes, std::vector<uint64_t>& offs);
void allGatherVInto(fg_ctx* ctx, const void* src, const std::vector<uint64_t>& offs, void* dst);
void exchangeSizes(fg_ctx* ctx, const std::vector<uint64_t>& sendOffs, std::vector<uint64_t>& recvOffs);
void allToAllVInto(fg_ctx* ctx, const void* sendBuf, const std::vector<uint64_t>& sendOffs, void* recvBuf, const std::vector<uint64_t>& recvOffs);
void agreeOrThrow(fg_ctx* ctx, int code, const std::string& msg);
void allReduceSumU64(fg_ctx* ctx, unsigned long long* buf, size_t count);
void broadcastBytes(fg_ctx* ctx, void* buf, uint64_t bytes, int root);
void groupStart();
void groupEnd();
inline bool sharded(const fg_ctx* ctx) { return ctx->nRanks > 1 && ctx->ncclComm && ctx->shardSet; }

void editDistances(fg_ctx* ctx, fg_overlap* dOv, const fg_overlap* hOv, uint32_t nOv, bool useHpc, bool querySet, float maxDivergence,
                   const float* dQueryMaxDivergence);
void prepareEditDistances(fg_ctx* ctx, bool useHpc, bool querySet);   // builds the HPC copies editDistances reads (main thread, before the lanes start)
double intPeak(fg_ctx* ctx);
int debugEditDistance(fg_ctx* ctx, const uint8_t* a, int n, const uint8_t* b, int m, int rcA, int rcB);
void debugWarpSort(fg_ctx* ctx, uint64_t* keys, uint32_t* vals, const uint64_t* segOffsets, uint32_t nSegs);
void alignCigarBatch(fg_ctx* ctx, const uint8_t* targets, const uint64_t* tOff, const uint8_t* queries, const uint64_t* qOff, uint32_t n, uint32_t cigarCap,
                   uint32_t* cigars, uint32_t* nCigar, int32_t* status);

}  // namespace fg
