// CPU models of segRadixSortKernel and segTileSortKernel (flye_b200/csrc/overlap.cu): the same per-warp chunks, (warp, digit) counts, exclusive offsets,
// match.any ranking and next-pass counts accumulated while scattering, executed warp by warp — against std::stable_sort by extId.
// Also the claim the fast path rests on: for hits in curPos order without an (extId, curPos) tie, that stable sort equals
// std::sort by (extId, curPos).   Usage: segsort_check [trials] [seed]
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

static const int SEG_WARPS = 16;

static std::vector<uint64_t> modelSort(std::vector<uint64_t> bufA, int posBits, int nPass) {
    const uint32_t n = (uint32_t)bufA.size();
    std::vector<uint64_t> bufB(n), out(n);
    if (!n) return out;
    const uint32_t C = (((n + SEG_WARPS - 1) / SEG_WARPS) + 31u) & ~31u;
    const int idShift = 2 * posBits;
    std::vector<std::vector<uint32_t>> cur(SEG_WARPS, std::vector<uint32_t>(256, 0)), nxt = cur;
    const uint64_t* src = bufA.data(); uint64_t* dst = bufB.data();
    auto range = [&](int w, uint32_t& b, uint32_t& e) { b = std::min(n, (uint32_t)w * C); e = std::min(n, b + C); };
    for (int w = 0; w < SEG_WARPS; ++w) {   // counts of the first pass
        uint32_t b, e; range(w, b, e);
        for (uint32_t i = b; i < e; ++i) ++cur[w][(uint32_t)(src[i] >> idShift) & 255u];
    }
    for (int pass = 0; pass < nPass; ++pass) {
        const bool last = pass == nPass - 1;
        const int sh = idShift + 8 * pass;
        std::vector<uint32_t> base(256), tot(256);
        for (int d = 0; d < 256; ++d) { uint32_t t = 0; for (int v = 0; v < SEG_WARPS; ++v) { const uint32_t c = cur[v][d]; cur[v][d] = t; t += c; } tot[d] = t; }
        uint32_t acc = 0;
        for (int d = 0; d < 256; ++d) { base[d] = acc; acc += tot[d]; }
        for (auto& row : nxt) std::fill(row.begin(), row.end(), 0u);
        for (int w = 0; w < SEG_WARPS; ++w) {   // the warps never touch each other's offsets, so any order of the warps is the same
            uint32_t b, e; range(w, b, e);
            for (uint32_t i0 = b; i0 < e; i0 += 32) {
                uint32_t pos[32]; uint32_t dig[32]; int act = (int)std::min<uint32_t>(32, e - i0);
                for (int l = 0; l < act; ++l) dig[l] = (uint32_t)(src[i0 + l] >> sh) & 255u;
                for (int l = 0; l < act; ++l) {   // all lanes read the offset before the leaders update it
                    int rank = 0;
                    for (int p = 0; p < l; ++p) rank += dig[p] == dig[l];
                    pos[l] = base[dig[l]] + cur[w][dig[l]] + rank;
                }
                for (int l = 0; l < act; ++l) {   // leader = lowest lane of every digit
                    bool leader = true; int cnt = 0;
                    for (int p = 0; p < act; ++p) { if (dig[p] == dig[l]) { ++cnt; if (p < l) leader = false; } }
                    if (leader) cur[w][dig[l]] += cnt;
                }
                for (int l = 0; l < act; ++l) {
                    const uint64_t x = src[i0 + l];
                    if (!last) { dst[pos[l]] = x; ++nxt[pos[l] / C][(uint32_t)(x >> (sh + 8)) & 255u]; }
                    else out[pos[l]] = x;
                }
            }
        }
        std::swap(cur, nxt);
        const uint64_t* t = dst; dst = const_cast<uint64_t*>(src); src = t;
    }
    return out;
}

// segTileSortKernel: tiles of 4096 elements; warp w of the CTA owns the tile positions [256 w, 256 w + 256), step e its e-th 32
// elements; in-warp ranks with match.any, exclusive scan of the (warp, digit) counts over the warps, run starts of the staged tile,
// digitBase advancing tile by tile; copy-out position = digitBase[d] + (staged index - tileStart[d]).
static std::vector<uint64_t> modelTileSort(std::vector<uint64_t> bufA, int posBits, int nPass) {
    const int WARPS = 16, E = 8, TILE = WARPS * 32 * E;
    const uint32_t n = (uint32_t)bufA.size();
    std::vector<uint64_t> bufB(n), out(n), stage(TILE);
    if (!n) return out;
    const int idShift = 2 * posBits;
    const uint64_t* src = bufA.data(); uint64_t* dst = bufB.data();
    std::vector<uint32_t> nxtCnt(256, 0), digitBase(256), tileStart(256);
    for (uint32_t i = 0; i < n; ++i) ++nxtCnt[(uint32_t)(src[i] >> idShift) & 255u];
    for (int pass = 0; pass < nPass; ++pass) {
        const bool last = pass == nPass - 1;
        const int sh = idShift + 8 * pass;
        uint32_t acc = 0;
        for (int d = 0; d < 256; ++d) { digitBase[d] = acc; acc += nxtCnt[d]; nxtCnt[d] = 0; }
        std::vector<uint32_t> prevRun(256, 0);
        for (uint32_t t0 = 0; t0 < n; t0 += TILE) {
            const uint32_t tileN = std::min<uint32_t>(TILE, n - t0);
            std::vector<std::vector<uint32_t>> warpCnt(WARPS, std::vector<uint32_t>(256, 0));
            std::vector<uint32_t> rank(TILE, 0);   // r[e] of (warp, e, lane) at tile position 256 w + 32 e + lane
            for (int w = 0; w < WARPS; ++w)
                for (int e = 0; e < E; ++e) {
                    const uint32_t b = t0 + (uint32_t)w * 32u * E + (uint32_t)e * 32u;
                    uint32_t dig[32]; int act = 0;
                    for (int l = 0; l < 32; ++l) if (b + l < n) { dig[l] = (uint32_t)(src[b + l] >> sh) & 255u; act = l + 1; }
                    for (int l = 0; l < act; ++l) {   // every lane reads the count before the leaders add to it
                        int lower = 0;
                        for (int p = 0; p < l; ++p) lower += dig[p] == dig[l];
                        rank[b - t0 + l] = warpCnt[w][dig[l]] + lower;
                    }
                    for (int l = 0; l < act; ++l) {
                        bool leader = true; int cnt = 0;
                        for (int p = 0; p < act; ++p) if (dig[p] == dig[l]) { ++cnt; if (p < l) leader = false; }
                        if (leader) warpCnt[w][dig[l]] += cnt;
                    }
                }
            std::vector<uint32_t> run(256, 0);
            for (int d = 0; d < 256; ++d) {
                for (int v = 0; v < WARPS; ++v) { const uint32_t c = warpCnt[v][d]; warpCnt[v][d] = run[d]; run[d] += c; }
                digitBase[d] += prevRun[d];
                prevRun[d] = run[d];
            }
            uint32_t a2 = 0;
            for (int d = 0; d < 256; ++d) { tileStart[d] = a2; a2 += run[d]; }
            if (a2 != tileN) { printf("tile model: counts do not add up\n"); exit(1); }
            for (int w = 0; w < WARPS; ++w)
                for (int e = 0; e < E; ++e)
                    for (int l = 0; l < 32; ++l) {
                        const uint32_t i = t0 + (uint32_t)w * 32u * E + (uint32_t)e * 32u + l;
                        if (i >= n) continue;
                        const uint32_t d = (uint32_t)(src[i] >> sh) & 255u;
                        stage[tileStart[d] + warpCnt[w][d] + rank[i - t0]] = src[i];
                    }
            for (uint32_t s0 = 0; s0 < tileN; ++s0) {
                const uint64_t y = stage[s0];
                const uint32_t d = (uint32_t)(y >> sh) & 255u;
                const uint32_t g = digitBase[d] + (s0 - tileStart[d]);
                if (!last) { dst[g] = y; ++nxtCnt[(uint32_t)(y >> (sh + 8)) & 255u]; }
                else out[g] = y;
            }
        }
        const uint64_t* t = dst; dst = const_cast<uint64_t*>(src); src = t;
    }
    return out;
}

int main(int argc, char** argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 3000;
    std::mt19937_64 rng(argc > 2 ? atoi(argv[2]) : 5);
    long elems = 0, tieFree = 0;
    for (int t = 0; t < trials; ++t) {
        const int posBits = 10 + (int)(rng() % 11), idBits = 1 + (int)(rng() % 24), nPass = (idBits + 7) / 8;
        uint32_t n;
        switch (t % 6) { case 0: n = (uint32_t)(rng() % 40); break; case 1: n = 500 + (uint32_t)(rng() % 40); break; case 2: n = 512 * (1 + (uint32_t)(rng() % 4)); break;
                         case 3: n = (uint32_t)(rng() % 3000); break; default: n = (uint32_t)(rng() % 60000); }
        const uint32_t nIds = 1 + (uint32_t)(rng() % std::min<uint64_t>(1ULL << idBits, (t % 2) ? 40 : 100000));
        std::vector<uint64_t> hits(n);
        uint32_t cur = 0;
        for (uint32_t i = 0; i < n; ++i) {   // expansion order: curPos ascending
            if (rng() % 3) cur += 1 + (uint32_t)(rng() % 5);
            cur &= (1u << posBits) - 1u;
            const uint64_t ext = rng() & ((1ULL << posBits) - 1ULL), id = (rng() % nIds) & ((1ULL << idBits) - 1ULL);
            hits[i] = (id << (2 * posBits)) | ((uint64_t)cur << posBits) | ext;
        }
        std::stable_sort(hits.begin(), hits.end(), [&](uint64_t a, uint64_t b) { return ((a >> posBits) & ((1ULL << posBits) - 1)) < ((b >> posBits) & ((1ULL << posBits) - 1)); });
        std::vector<uint64_t> want = hits;
        std::stable_sort(want.begin(), want.end(), [&](uint64_t a, uint64_t b) { return (a >> (2 * posBits)) < (b >> (2 * posBits)); });
        const std::vector<uint64_t> got = modelSort(hits, posBits, nPass);
        if (got != want) { printf("MISMATCH trial %d n=%u idBits=%d posBits=%d\n", t, n, idBits, posBits); return 1; }
        if (modelTileSort(hits, posBits, nPass) != want) { printf("MISMATCH (tile model) trial %d n=%u idBits=%d posBits=%d\n", t, n, idBits, posBits); return 1; }
        bool tie = false;
        for (uint32_t i = 0; i + 1 < n; ++i) tie = tie || (want[i] >> posBits) == (want[i + 1] >> posBits);
        if (!tie) {   // distinct (extId, curPos) keys: std::sort has one possible result, and it is this one
            std::vector<uint64_t> s = hits;
            std::shuffle(s.begin(), s.end(), rng);
            std::sort(s.begin(), s.end(), [&](uint64_t a, uint64_t b) { return (a >> posBits) < (b >> posBits); });
            if (s != want) { printf("MISMATCH tie-free claim trial %d\n", t); return 1; }
            ++tieFree;
        }
        elems += n;
    }
    printf("OK %d %ld tieFree=%ld\n", trials, elems, tieFree);
    return 0;
}
