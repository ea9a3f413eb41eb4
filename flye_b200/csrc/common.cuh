// flye_b200 — shared device/host helpers for the sm_100a read-overlap path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/flye_b200.h"
#include "kmer_math.cuh"

namespace fg {

// ---------------------------------------------------------------------------------------------
// errors: internal code throws fg::Error; the extern "C" layer converts to a status + message
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define FG_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            throw fg::Error(FG_ERR_CUDA, std::string(#call) + " failed at " + __FILE__ + ":" +      \
                                             std::to_string(__LINE__) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// device buffers: every context owns a caching arena.  A released block goes back to the arena's free list and
// is handed out again to the next request it fits (best fit, at most 2x oversize); cudaMalloc / cudaFree are
// only paid while the working set is still growing (first step) and at context destruction.  All work of a
// context is ordered on its single stream, so a block may be reused as soon as the host has released it.
// ---------------------------------------------------------------------------------------------
struct Arena {
    struct Block { void* p; size_t bytes; bool used; };
    std::vector<Block> blocks;
    size_t totalBytes = 0;
    uint64_t mallocCalls = 0;
    std::function<void()> onPressure;   // frees what OTHER arenas of the context have cached (set by the context)
    // size classes {1, 1.25, 1.5, 1.75} * 2^k (>= 256 KiB): a request only ever reuses a block of its own class, so
    // a phase that issues the same request sequence every step hits the cache on every request from step 2 on
    static size_t sizeClass(size_t bytes) {
        size_t c = 256u << 10;
        while (c < bytes) {
            const size_t q = c / 4;
            for (int i = 1; i <= 4; ++i) if (c + i * q >= bytes) return c + i * q;
            c *= 2;
        }
        return c;
    }
    void* alloc(size_t bytes) {
        if (!bytes) return nullptr;
        const size_t want = sizeClass(bytes);
        for (auto& b : blocks)
            if (!b.used && b.bytes == want) { b.used = true; return b.p; }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, want);
        ++mallocCalls;
        if (e != cudaSuccess) {   // give cached blocks back (this arena's, then the context's other arenas') and retry once
            cudaGetLastError();
            trim();
            e = cudaMalloc(&p, want);
            if (e != cudaSuccess && onPressure) { cudaGetLastError(); onPressure(); e = cudaMalloc(&p, want); }
            if (e != cudaSuccess)
                throw Error(FG_ERR_CUDA, std::string("cudaMalloc of ") + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e));
        }
        blocks.push_back({p, want, true});
        totalBytes += want;
        return p;
    }
    void release(void* p) {
        for (auto& b : blocks) if (b.p == p) { b.used = false; return; }
    }
    size_t cachedFreeBytes() const {
        size_t n = 0;
        for (const auto& b : blocks) if (!b.used) n += b.bytes;
        return n;
    }
    void trim() {   // free every unused block (synchronises the device)
        cudaDeviceSynchronize();
        std::vector<Block> keep;
        for (auto& b : blocks) { if (b.used) keep.push_back(b); else { cudaFree(b.p); totalBytes -= b.bytes; } }
        blocks.swap(keep);
    }
    ~Arena() { for (auto& b : blocks) cudaFree(b.p); }
};
inline Arena*& currentArena() { static thread_local Arena* a = nullptr; return a; }

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    Arena* owner = nullptr;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), owner(o.owner) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; owner = o.owner; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        owner = currentArena();
        if (!owner) throw Error(FG_ERR_INTERNAL, "device allocation outside a context call");
        if (count) p = (T*)owner->alloc(count * sizeof(T));
    }
    void ensure(size_t count) { if (count > n) alloc(count + count / 8); }
    // hands the block over under another element type (no copy, no allocation)
    template <class U> DevBuf<U> reinterpretAs() {
        DevBuf<U> o;
        o.p = reinterpret_cast<U*>(p); o.n = n * sizeof(T) / sizeof(U); o.owner = owner;
        p = nullptr; n = 0;
        return o;
    }
    void release() { if (p && owner) owner->release(p); p = nullptr; n = 0; }
    size_t bytes() const { return n * sizeof(T); }
};

template <class T>
struct PinnedBuf {
    T* p = nullptr;
    size_t n = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { release(); }
    void ensure(size_t count) {
        if (count <= n) return;
        release();
        n = count + count / 4 + 16;
        FG_CUDA(cudaMallocHost((void**)&p, n * sizeof(T)));
    }
    // grow, keeping the first `used` elements
    void ensureKeep(size_t count, size_t used) {
        if (count <= n) return;
        T* old = p; size_t oldN = n;
        p = nullptr; n = count + count / 2 + 16;
        FG_CUDA(cudaMallocHost((void**)&p, n * sizeof(T)));
        if (old) { if (used) memcpy(p, old, std::min(used, oldN) * sizeof(T)); cudaFreeHost(old); }
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};

// ---------------------------------------------------------------------------------------------
// open-addressing table in HBM: 16-byte slots {key, payload}; one probe = one 32-byte sector
// ---------------------------------------------------------------------------------------------
static constexpr uint64_t EMPTY_KEY = ~0ULL;

struct Table {
    ulonglong2* slots = nullptr;
    uint64_t mask = 0;   // capacity - 1 (capacity is a power of two)
};

__device__ inline bool tableFind(const Table& t, uint64_t key, uint64_t& payload) {
    uint64_t i = mix64(key) & t.mask;
    for (;;) {
        ulonglong2 s = __ldg(&t.slots[i]);
        if (s.x == key) { payload = s.y; return true; }
        if (s.x == EMPTY_KEY) return false;
        i = (i + 1) & t.mask;
    }
}

// keys are unique among inserters: plain claim + store
__device__ inline void tableInsertUnique(const Table& t, uint64_t key, uint64_t payload) {
    uint64_t i = mix64(key) & t.mask;
    for (;;) {
        unsigned long long prev = atomicCAS((unsigned long long*)&t.slots[i].x, (unsigned long long)EMPTY_KEY,
                                            (unsigned long long)key);
        if (prev == EMPTY_KEY) { t.slots[i].y = payload; return; }
        i = (i + 1) & t.mask;
    }
}

// counting insert: payload += 1
__device__ inline void tableAddOne(const Table& t, uint64_t key) {
    uint64_t i = mix64(key) & t.mask;
    for (;;) {
        unsigned long long prev = atomicCAS((unsigned long long*)&t.slots[i].x, (unsigned long long)EMPTY_KEY,
                                            (unsigned long long)key);
        if (prev == EMPTY_KEY || prev == key) { atomicAdd((unsigned long long*)&t.slots[i].y, 1ULL); return; }
        i = (i + 1) & t.mask;
    }
}

// k-mer counts as the selection / classification kernels see them: on one GPU the dense counter array of the counting
// phase itself (one 32-byte sector per lookup, no hashing, no probing); with the classes spread over several ranks the
// replicated table of the k-mers with count >= 2 (absent == 1).  Dense index: see count_index.cu.
struct CountView {
    const uint32_t* dense = nullptr;
    const uint32_t* solidBits = nullptr;   // optional front: bit set <=> the class occurs at least twice; `dense` then holds count - 1
    // multi-GPU: every rank's 16-bit saturated counters, all-gathered.  Class i lives at slot (i % nOwners) * nLocal + i / nOwners
    // (solidBits uses the same slot numbering); a value of 65535 continues in `table` (class index -> count).
    const uint16_t* count16 = nullptr;
    uint64_t nLocal = 0;
    uint32_t nOwners = 1, ownerShift = 0;   // ownerShift = log2(nOwners) if it is a power of two, else 0xffffffff
    Table table;
    int k = 0;
};
__device__ __forceinline__ uint64_t countSlotOf(const CountView& c, uint64_t i) {
    if (c.ownerShift != 0xffffffffu) return (i & (c.nOwners - 1u)) * c.nLocal + (i >> c.ownerShift);
    const uint64_t local = i / c.nOwners;
    return (i - local * c.nOwners) * c.nLocal + local;
}
__device__ __forceinline__ bool testBit(const uint32_t* __restrict__ bits, uint64_t i) { return (__ldg(&bits[i >> 5]) >> (i & 31)) & 1u; }
__device__ inline uint32_t countOfClass16(const CountView& c, uint64_t i) {
    const uint64_t slot = countSlotOf(c, i);
    if (c.solidBits && !testBit(c.solidBits, slot)) return 1u;   // callers only ask for k-mers that occur
    uint32_t v = c.count16[slot];
    if (v == 65535u) { uint64_t payload; if (tableFind(c.table, i, payload)) v = (uint32_t)payload; }
    return v;
}
// count of the class of the k-mer given as its 2k-bit window `v` (base p in the lowest bits) ...
__device__ inline uint32_t countOfWindow(const CountView& c, uint64_t v) {
    const uint64_t f = fwdFromWindow(v, c.k), r = (~v) & kmerMask(c.k);
    if (c.dense) {
        const uint64_t i = denseIndexOfPair(f, r, c.k);
        if (!c.solidBits) return c.dense[i];
        return testBit(c.solidBits, i) ? c.dense[i] + 1u : 1u;   // most positions of noisy reads are singletons: answered from the L2-resident bitmap
    }
    return countOfClass16(c, denseIndexOfPair(f, r, c.k));
}
// ... or as a k-mer in the reference's representation (either orientation); a k-mer that never occurs reports 0 (1 behind a
// bitmap front) — callers only ask for k-mers that occur
__device__ inline uint32_t countOfKmer(const CountView& c, uint64_t kmer) {
    const uint64_t r = revCompKmer(kmer, c.k);
    if (c.dense) {
        const uint64_t i = denseIndexOfPair(kmer, r, c.k);
        if (!c.solidBits) return c.dense[i];
        return testBit(c.solidBits, i) ? c.dense[i] + 1u : 1u;
    }
    return countOfClass16(c, denseIndexOfPair(kmer, r, c.k));
}

// index payload: [63:24] first entry, [23:0] size; all-ones size = repetitive k-mer
static constexpr uint64_t IDX_SIZE_BITS = 24;
static constexpr uint64_t IDX_SIZE_MASK = (1ULL << IDX_SIZE_BITS) - 1;
static constexpr uint64_t IDX_REPETITIVE = IDX_SIZE_MASK;

// ---------------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------------
__device__ inline int laneId() { return threadIdx.x & 31; }

// position of the n-th (0-based) lowest set bit; requires popc(mask) > n
__host__ __device__ inline int nthLowBit(uint32_t mask, int n) {
    int pos = 0;
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        uint32_t part = (mask >> pos) & ((1u << w) - 1u);
#ifdef __CUDA_ARCH__
        int c = __popc(part);
#else
        int c = __builtin_popcount(part);
#endif
        if (n >= c) { n -= c; pos += w; }
    }
    return pos;
}
__host__ __device__ inline uint32_t brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = (x >> 16) | (x << 16);
    x = ((x & 0xFF00FF00u) >> 8) | ((x & 0x00FF00FFu) << 8);
    x = ((x & 0xF0F0F0F0u) >> 4) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x & 0xCCCCCCCCu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xAAAAAAAAu) >> 1) | ((x & 0x55555555u) << 1);
    return x;
#endif
}
__host__ __device__ inline int nthHighBit(uint32_t mask, int n) { return 31 - nthLowBit(brev32(mask), n); }

}  // namespace fg
