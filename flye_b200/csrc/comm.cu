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
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;

    static Nccl& get() {
        static Nccl n;
        if (n.handle) return n;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
        }
        if (!n.handle) throw Error(FG_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
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
// order in `out`; offs[r] .. offs[r+1] is rank r's part.  Implemented as one ncclBroadcast per rank inside a group
// (no padding, no staging copy).
void allGatherV(fg_ctx* ctx, const void* src, uint64_t bytes, DevBuf<char>& out, std::vector<uint64_t>& offs) {
    Nccl& n = Nccl::get();
    ncclComm_t comm = (ncclComm_t)ctx->ncclComm;
    const int R = ctx->nRanks;
    DevBuf<uint64_t> dSizes(R), dMine(1);
    FG_CUDA(cudaMemcpyAsync(dMine.p, &bytes, 8, cudaMemcpyHostToDevice, ctx->stream));
    ncclCheck(n.AllGather(dMine.p, dSizes.p, 1, ncclUint64, comm, ctx->stream), "ncclAllGather(sizes)");
    std::vector<uint64_t> sizes(R);
    FG_CUDA(cudaMemcpyAsync(sizes.data(), dSizes.p, R * 8ULL, cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(cudaStreamSynchronize(ctx->stream));
    offs.assign(R + 1, 0);
    for (int r = 0; r < R; ++r) offs[r + 1] = offs[r] + sizes[r];
    out.alloc(std::max<uint64_t>(offs[R], 1));
    ncclCheck(n.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < R; ++r)
        if (sizes[r]) ncclCheck(n.Broadcast(src, out.p + offs[r], sizes[r], ncclChar, r, comm, ctx->stream), "ncclBroadcast");
    ncclCheck(n.GroupEnd(), "ncclGroupEnd");
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
