// flye_b200 — integer-issue ceiling of the chip, measured ("measure the chip's achievable IMAD/LOP3 rate with a
// microbenchmark; do not assume a spec number").  The DP and edit-distance kernels are bound by the INT32 ALU / shift pipes,
// not by HBM; bench.py reports them against this number.  Every thread runs 8 independent chains of IMAD, LOP3 and SHF
// (funnel shift) — the instruction mix of wfaKernel / chainRunDpKernel — with no memory traffic; all SMs, full occupancy.
#include "ctx.cuh"

namespace fg {

static constexpr int INT_CHAINS = 8, INT_ITERS = 2048;
__global__ void __launch_bounds__(256) intPeakKernel(uint32_t seed, uint32_t* __restrict__ out) {
    uint32_t x[INT_CHAINS];
#pragma unroll
    for (int c = 0; c < INT_CHAINS; ++c) x[c] = seed + threadIdx.x * 2654435761u + c * 40503u + blockIdx.x;
    const uint32_t m = seed | 1u, a = seed ^ 0x9e3779b9u;
#pragma unroll 1
    for (int it = 0; it < INT_ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < INT_CHAINS; ++c) {
            x[c] = x[c] * m + a;                               // IMAD
            x[c] = (x[c] & a) ^ (x[c] | m);                    // LOP3
            x[c] = __funnelshift_l(x[c], x[c], 7);             // SHF
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < INT_CHAINS; ++c) s ^= x[c];
    if (s == 0x12345678u) out[0] = s;   // keeps the chains alive
}

// Gop/s (one op = one IMAD, LOP3 or SHF thread instruction)
double intPeak(fg_ctx* ctx) {
    DevBuf<uint32_t> out(1);
    const int grid = 148 * 8;
    cudaEvent_t e0, e1;
    FG_CUDA(cudaEventCreate(&e0)); FG_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        FG_CUDA(cudaEventRecord(e0, ctx->stream));
        intPeakKernel<<<grid, 256, 0, ctx->stream>>>(12345u + rep, out.p);
        checkLaunch(ctx, "intPeakKernel");
        FG_CUDA(cudaEventRecord(e1, ctx->stream));
        FG_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        FG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)grid * 256 * INT_ITERS * INT_CHAINS * 3;
        if (rep) best = std::max(best, ops / (ms * 1e-3) / 1e9);   // first launch: warm-up
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

}  // namespace fg
