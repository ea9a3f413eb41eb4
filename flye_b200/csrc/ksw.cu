// flye_b200 — device build of the banded KSW2 alignment behind checkIdyAndTrim (ksw_core.cuh): fg_align_cigar_batch.  One thread per alignment: the routine is ksw2's own recurrence byte for byte (its results depend
// on ksw2's in-place block updates, see ksw_core.cuh), so the first device version parallelises over alignments only.  The band
// doubling of getAlignmentCigarKsw (alignment.cpp:150-165) is driven from the host: every round launches the pairs whose band
// was too narrow with twice the band.  The host mirror uses it for partitionBadMappings when FLYE_B200_DEVICE_KSW=1 (default: the
// reference's own checkIdyAndTrim on host threads).
#include "ctx.cuh"
#include "ksw_core.cuh"

#include <algorithm>
#include <vector>

namespace fg {

__global__ void __launch_bounds__(64) kswCigarKernel(const uint8_t* __restrict__ targets, const uint64_t* __restrict__ tOff, const uint8_t* __restrict__ queries,
                                                     const uint64_t* __restrict__ qOff, const uint32_t* __restrict__ pairIdx, uint32_t nActive, int band,
                                                     uint8_t* __restrict__ mem, const uint64_t* __restrict__ memOff, uint8_t* __restrict__ p,
                                                     const uint64_t* __restrict__ pOff, int* __restrict__ offs, const uint64_t* __restrict__ offOff,
                                                     uint32_t* __restrict__ cigars, uint32_t cigarCap, uint32_t* __restrict__ nCigar, int32_t* __restrict__ status) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nActive) return;
    const uint32_t i = pairIdx[a];
    const int tlen = (int)(tOff[i + 1] - tOff[i]), qlen = (int)(qOff[i + 1] - qOff[i]);
    const size_t rounds = (size_t)qlen + (size_t)tlen - 1;
    int n = 0;
    const int rc = kswExtz2Core(queries + qOff[i], qlen, targets + tOff[i], tlen, band, mem + memOff[a], p + pOff[a], offs + offOff[a],
                                offs + offOff[a] + rounds, cigars + (size_t)i * cigarCap, (int)cigarCap, &n);
    status[i] = rc;
    nCigar[i] = (uint32_t)n;
}

void alignCigarBatch(fg_ctx* ctx, const uint8_t* targets, const uint64_t* tOff, const uint8_t* queries, const uint64_t* qOff, uint32_t n, uint32_t cigarCap,
                   uint32_t* cigars, uint32_t* nCigar, int32_t* status) {
    if (!n) return;
    if (!cigarCap) throw Error(FG_ERR_ARG, "cigar_cap must be positive");
    for (uint32_t i = 0; i < n; ++i) {
        if (tOff[i + 1] < tOff[i] || qOff[i + 1] < qOff[i] || tOff[i + 1] - tOff[i] >= (1u << 30) || qOff[i + 1] - qOff[i] >= (1u << 30))
            throw Error(FG_ERR_ARG, "bad sequence offsets");
        status[i] = KSW_OK; nCigar[i] = 0;
    }
    DevBuf<uint8_t> dT(std::max<uint64_t>(tOff[n], 1)), dQ(std::max<uint64_t>(qOff[n], 1));
    DevBuf<uint64_t> dTOff(n + 1), dQOff(n + 1);
    DevBuf<uint32_t> dCigars((uint64_t)n * cigarCap), dNCigar(n);
    DevBuf<int32_t> dStatus(n);
    cudaStream_t st = streamOf(ctx);
    if (tOff[n]) FG_CUDA(cudaMemcpyAsync(dT.p, targets, tOff[n], cudaMemcpyHostToDevice, st));
    if (qOff[n]) FG_CUDA(cudaMemcpyAsync(dQ.p, queries, qOff[n], cudaMemcpyHostToDevice, st));
    FG_CUDA(cudaMemcpyAsync(dTOff.p, tOff, (n + 1) * 8ULL, cudaMemcpyHostToDevice, st));
    FG_CUDA(cudaMemcpyAsync(dQOff.p, qOff, (n + 1) * 8ULL, cudaMemcpyHostToDevice, st));
    FG_CUDA(cudaMemsetAsync(dNCigar.p, 0, n * 4ULL, st));
    FG_CUDA(cudaMemsetAsync(dStatus.p, 0, n * 4ULL, st));
    std::vector<uint32_t> active;
    for (uint32_t i = 0; i < n; ++i)
        if (tOff[i + 1] > tOff[i] && qOff[i + 1] > qOff[i]) active.push_back(i);   // an empty side gives an empty CIGAR (ksw2_extz2_sse.c:64)
    const uint64_t budget = 3ULL << 30;   // workspace bytes per launch
    for (int band = 64; !active.empty(); band *= 2) {
        std::vector<uint32_t> retry;
        size_t pos = 0;
        while (pos < active.size()) {
            std::vector<uint32_t> idx; std::vector<uint64_t> memOff, pOff, offOff;
            uint64_t mTot = 0, pTot = 0, oTot = 0;
            while (pos < active.size()) {
                const uint32_t i = active[pos];
                const KswSizes z = kswSizes((int)(qOff[i + 1] - qOff[i]), (int)(tOff[i + 1] - tOff[i]), band);
                if (z.memBytes + z.pBytes + 8 * z.rounds > budget) throw Error(FG_ERR_ARG, "alignment too large for the device workspace");
                if (!idx.empty() && mTot + pTot + 4 * oTot + z.memBytes + z.pBytes + 8 * z.rounds > budget) break;
                idx.push_back(i); memOff.push_back(mTot); pOff.push_back(pTot); offOff.push_back(oTot);
                mTot += (z.memBytes + 15) & ~15ULL; pTot += (z.pBytes + 15) & ~15ULL; oTot += 2 * z.rounds;
                ++pos;
            }
            const uint32_t m = (uint32_t)idx.size();
            DevBuf<uint8_t> dMem(mTot), dP(pTot);
            DevBuf<int> dOffs(std::max<uint64_t>(oTot, 1));
            DevBuf<uint32_t> dIdx(m);
            DevBuf<uint64_t> dMemOff(m), dPOff(m), dOffOff(m);
            FG_CUDA(cudaMemcpyAsync(dIdx.p, idx.data(), m * 4ULL, cudaMemcpyHostToDevice, st));
            FG_CUDA(cudaMemcpyAsync(dMemOff.p, memOff.data(), m * 8ULL, cudaMemcpyHostToDevice, st));
            FG_CUDA(cudaMemcpyAsync(dPOff.p, pOff.data(), m * 8ULL, cudaMemcpyHostToDevice, st));
            FG_CUDA(cudaMemcpyAsync(dOffOff.p, offOff.data(), m * 8ULL, cudaMemcpyHostToDevice, st));
            kswCigarKernel<<<(m + 63) / 64, 64, 0, st>>>(dT.p, dTOff.p, dQ.p, dQOff.p, dIdx.p, m, band, dMem.p, dMemOff.p, dP.p, dPOff.p, dOffs.p, dOffOff.p,
                                                       dCigars.p, cigarCap, dNCigar.p, dStatus.p);
            checkLaunch(ctx, "kswCigarKernel");
            FG_CUDA(cudaMemcpyAsync(status, dStatus.p, n * 4ULL, cudaMemcpyDeviceToHost, st));
            FG_CUDA(cudaStreamSynchronize(st));
            for (uint32_t i : idx) {
                const int longest = (int)std::max(qOff[i + 1] - qOff[i], tOff[i + 1] - tOff[i]);
                if (status[i] == KSW_BAND_TOO_NARROW && band <= longest) retry.push_back(i);   // alignment.cpp:163-164
            }
        }
        active.swap(retry);
    }
    FG_CUDA(cudaMemcpyAsync(cigars, dCigars.p, (uint64_t)n * cigarCap * 4ULL, cudaMemcpyDeviceToHost, st));
    FG_CUDA(cudaMemcpyAsync(nCigar, dNCigar.p, n * 4ULL, cudaMemcpyDeviceToHost, st));
    FG_CUDA(cudaMemcpyAsync(status, dStatus.p, n * 4ULL, cudaMemcpyDeviceToHost, st));
    FG_CUDA(cudaStreamSynchronize(st));
}

}  // namespace fg
