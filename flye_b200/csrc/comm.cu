// flye_b200 — NCCL plumbing for the multi-GPU layout (one process per GPU, reads partitioned, index replicated;
// SURVEY.md §8e).  NCCL is resolved at run time with dlopen so that the library uses the SAME libnccl the host
// program already loaded (torch bundles its own) and still loads on machines without NCCL.
#include "ctx.cuh"

#include <dlfcn.h>

namespace fg {

namespace {
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclSuccess = 0 };
enum { ncclChar = 0, ncclUint64 = 5 };   // ncclDataType_t
enum { ncclSum = 0 };                    // ncclRedOp_t

struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;

    static Nccl& get() {
        static Nccl n;
        static bool ready = false;
        static std::mutex m;
        std::lock_guard<std::mutex> lock(m);
        if (ready) return n;
        if (!n.handle)
            for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
                n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
                if (n.handle) break;
            }
        if (!n.handle) throw Error(FG_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
        // (`ready` is only set once every symbol resolved: a failed attempt never leaves a half-filled table behind)
        auto sym = [&](const char* s) { void* p = dlsym(n.handle, s); if (!p) throw Error(FG_ERR_NCCL, std::string("missing NCCL symbol ") + s); return p; };
        n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
        n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
        n.AllGather = (decltype(n.AllGather))sym("ncclAllGather");
        n.AllReduce = (decltype(n.AllReduce))sym("ncclAllReduce");
        n.Broadcast = (decltype(n.Broadcast))sym("ncclBroadcast");
        n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
        n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
        n.Send = (decltype(n.Send))sym("ncclSend");
        n.Recv = (decltype(n.Recv))sym("ncclRecv");
        ready = true;
        return n;
    }
};

void ncclCheck(int rc, const char* what) {
    if (rc != ncclSuccess) throw Error(FG_ERR_NCCL, std::string(what) + ": " + Nccl::get().GetErrorString(rc));
}
}  // namespace

void commUniqueId(uint8_t* id) {
    ncclUniqueId u;
    ncclCheck(Nccl::get().GetUniqueId(&u), "ncclGetUniqueId");
    static_assert(sizeof(u) == FG_NCCL_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &u, sizeof u);
}

void commInit(fg_ctx* ctx, int nRanks, int rank, const uint8_t* id) {
    // the ranks of one box share its host cores: size the per-context host pool accordingly (FG_HOST_THREADS overrides)
    {
        unsigned t = std::max(2u, std::min(16u, std::thread::hardware_concurrency() / (unsigned)std::max(1, nRanks)));
        if (const char* e = getenv("FG_HOST_THREADS")) t = (unsigned)std::max(1, atoi(e));
        ctx->hostPool.maxThreads = t;
    }
    if (nRanks < 1 || rank < 0 || rank >= nRanks) throw Error(FG_ERR_ARG, "bad rank / world size");
    commDestroy(ctx);
    ctx->nRanks = nRanks; ctx->rank = rank;
    if (nRanks == 1) return;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm = nullptr;
    ncclCheck(Nccl::get().CommInitRank(&comm, nRanks, u, rank), "ncclCommInitRank");
    ctx->ncclComm = comm;
}

void commDestroy(fg_ctx* ctx) {
    if (ctx->ncclComm) { Nccl::get().CommDestroy((ncclComm_t)ctx->ncclComm); ctx->ncclComm = nullptr; }
    ctx->nRanks = 1; ctx->rank = 0;
}

// variable-size all-gather: rank r contributes `bytes` bytes at `src`; every rank receives the concatenation in rank
// order; offs[r] .. offs[r+1] is rank r's part.  Two steps so that the caller can allocate the destination with its own type:
// allGatherSizes (one tiny ncclAllGather + host sync), then allGatherVInto = one ncclBroadcast per rank inside a group
// (no padding, no staging copy).  Ranks that contribute 0 bytes take part like every other rank.
void allGatherSizes(fg_ctx* ctx, uint64_t bytes, std::vector<uint64_t>& offs) {
    Nccl& n = Nccl::get();
    const int R = ctx->nRanks;
    DevBuf<uint64_t> dSizes(R), dMine(1);
    FG_CUDA(cudaMemcpyAsync(dMine.p, &bytes, 8, cudaMemcpyHostToDevice, ctx->stream));
    ncclCheck(n.AllGather(dMine.p, dSizes.p, 1, ncclUint64, (ncclComm_t)ctx->ncclComm, ctx->stream), "ncclAllGather(sizes)");
    std::vector<uint64_t> sizes(R);
    FG_CUDA(cudaMemcpyAsync(sizes.data(), dSizes.p, R * 8ULL, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    offs.assign(R + 1, 0);
    for (int r = 0; r < R; ++r) offs[r + 1] = offs[r] + sizes[r];
}

void allGatherVInto(fg_ctx* ctx, const void* src, const std::vector<uint64_t>& offs, void* dst) {
    Nccl& n = Nccl::get();
    ncclComm_t comm = (ncclComm_t)ctx->ncclComm;
    ncclCheck(n.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < ctx->nRanks; ++r)
        if (offs[r + 1] > offs[r])
            ncclCheck(n.Broadcast(src, (char*)dst + offs[r], offs[r + 1] - offs[r], ncclChar, r, comm, ctx->stream), "ncclBroadcast");
    ncclCheck(n.GroupEnd(), "ncclGroupEnd");
}

void allGatherV(fg_ctx* ctx, const void* src, uint64_t bytes, DevBuf<char>& out, std::vector<uint64_t>& offs) {
    allGatherSizes(ctx, bytes, offs);
    out.alloc(std::max<uint64_t>(offs.back(), 1));
    allGatherVInto(ctx, src, offs, out.p);
}

// all-to-all with per-destination sizes: this rank sends the bytes [sendOffs[d], sendOffs[d+1]) of sendBuf to rank d; on return
// recvOffs[s] .. recvOffs[s+1] delimits what rank s sent here, in rank order.  exchangeSizes = one ncclAllGather of the size row;
// allToAllVInto = grouped ncclSend / ncclRecv over NVLink (the part for this rank itself is a device copy).
void exchangeSizes(fg_ctx* ctx, const std::vector<uint64_t>& sendOffs, std::vector<uint64_t>& recvOffs) {
    Nccl& n = Nccl::get();
    const int R = ctx->nRanks;
    std::vector<uint64_t> row(R);
    for (int d = 0; d < R; ++d) row[d] = sendOffs[d + 1] - sendOffs[d];
    DevBuf<uint64_t> dRow(R), dAll((size_t)R * R);
    FG_CUDA(cudaMemcpyAsync(dRow.p, row.data(), R * 8ULL, cudaMemcpyHostToDevice, ctx->stream));
    ncclCheck(n.AllGather(dRow.p, dAll.p, R, ncclUint64, (ncclComm_t)ctx->ncclComm, ctx->stream), "ncclAllGather(size matrix)");
    std::vector<uint64_t> all((size_t)R * R);
    FG_CUDA(cudaMemcpyAsync(all.data(), dAll.p, all.size() * 8ULL, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    recvOffs.assign(R + 1, 0);
    for (int s = 0; s < R; ++s) recvOffs[s + 1] = recvOffs[s] + all[(size_t)s * R + ctx->rank];
}

void allToAllVInto(fg_ctx* ctx, const void* sendBuf, const std::vector<uint64_t>& sendOffs, void* recvBuf, const std::vector<uint64_t>& recvOffs) {
    Nccl& n = Nccl::get();
    ncclComm_t comm = (ncclComm_t)ctx->ncclComm;
    const int R = ctx->nRanks, me = ctx->rank;
    if (sendOffs[me + 1] > sendOffs[me])
        FG_CUDA(cudaMemcpyAsync((char*)recvBuf + recvOffs[me], (const char*)sendBuf + sendOffs[me], sendOffs[me + 1] - sendOffs[me],
                                cudaMemcpyDeviceToDevice, ctx->stream));
    ncclCheck(n.GroupStart(), "ncclGroupStart");
    for (int p = 0; p < R; ++p) {
        if (p == me) continue;
        if (sendOffs[p + 1] > sendOffs[p])
            ncclCheck(n.Send((const char*)sendBuf + sendOffs[p], sendOffs[p + 1] - sendOffs[p], ncclChar, p, comm, ctx->stream), "ncclSend");
        if (recvOffs[p + 1] > recvOffs[p])
            ncclCheck(n.Recv((char*)recvBuf + recvOffs[p], recvOffs[p + 1] - recvOffs[p], ncclChar, p, comm, ctx->stream), "ncclRecv");
    }
    ncclCheck(n.GroupEnd(), "ncclGroupEnd");
}

// Rank-local failures inside a sharded phase must not leave the peers blocked in the next collective: every rank brings its
// status word here (0 = fine) and all of them throw together when any is non-zero.
void agreeOrThrow(fg_ctx* ctx, int code, const std::string& msg) {
    if (!sharded(ctx)) { if (code) throw Error(code, msg); return; }
    DevBuf<unsigned long long> d(1);
    const unsigned long long mine = code ? 1ULL : 0ULL;
    unsigned long long sum = 0;
    FG_CUDA(cudaMemcpyAsync(d.p, &mine, 8, cudaMemcpyHostToDevice, ctx->stream));
    allReduceSumU64(ctx, d.p, 1);
    FG_CUDA(cudaMemcpyAsync(&sum, d.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (code) throw Error(code, msg);
    if (sum) throw Error(FG_ERR_INTERNAL, "a peer rank failed in a sharded phase (see its fg_last_error)");
}

void allReduceSumU64(fg_ctx* ctx, unsigned long long* buf, size_t count) {
    ncclCheck(Nccl::get().AllReduce(buf, buf, count, ncclUint64, ncclSum, (ncclComm_t)ctx->ncclComm, ctx->stream), "ncclAllReduce");
}

// in-place broadcast of buf[0,bytes) from `root`
void broadcastBytes(fg_ctx* ctx, void* buf, uint64_t bytes, int root) {
    if (!bytes) return;
    ncclCheck(Nccl::get().Broadcast(buf, buf, bytes, ncclChar, root, (ncclComm_t)ctx->ncclComm, ctx->stream), "ncclBroadcast");
}

void groupStart() { ncclCheck(Nccl::get().GroupStart(), "ncclGroupStart"); }
void groupEnd() { ncclCheck(Nccl::get().GroupEnd(), "ncclGroupEnd"); }

}  // namespace fg
