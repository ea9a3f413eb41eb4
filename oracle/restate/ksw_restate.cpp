// TEST INFRASTRUCTURE — CPU restatement of the alignment-trimming branch of getSeqOverlaps: what
// OverlapDetector does with a primary overlap that fails the divergence test when _partitionBadMappings is set
// (overlap.cpp:475-485): checkIdyAndTrim (alignment.cpp:306-495) = homopolymer compression (alignment.cpp:52-70), a banded
// global alignment with affine gaps (getAlignmentCigarKsw, alignment.cpp:102-216, which calls minimap2's ksw_extz2_sse,
// lib/minimap2/ksw2_extz2_sse.c — vendored in the reference tree, so pinned by it), and the search for the longest
// low-divergence stretches of that alignment.  No reference header is included and no code is shared; every function cites
// the lines it follows.  tests/test_oracle_trim.py pins it to the reference's own checkIdyAndTrim byte for byte.
//
// Why the DP is restated at the level of ksw2's bytes.  The reference takes the CIGAR, and the CIGAR of a banded
// Suzuki-Kasahara alignment is not a function of the two strings alone: ksw2 keeps the differences u, v, x, y of one
// anti-diagonal in 8-bit arrays indexed by the target position, updates them in place sixteen at a time, and never
// initialises the cells just outside the band — a cell entering the band reads whatever the 16-byte block update of an earlier
// anti-diagonal left at its position (values computed from stale scores), and the score block of a diagonal is written in
// unaligned 16-byte strides that run past the band.  Those bytes decide ties in the traceback.  The restatement therefore keeps
// the same byte arrays with the same layout in one zero-filled block (u | v | x | y | s | target | reversed query, then 16 spare
// bytes, as kcalloc(tlen_*6 + qlen_ + 1, 16) gives them) and performs the same per-byte operations in the same order,
// one byte at a time (lanes of a 16-byte operation do not interact except through the one-byte carries x1_ / v1_, which are
// kept).  The code path restated is the one the oracle's build of ksw2 takes: SSE2 without SSE4.1 (oracle/Makefile: -msse2),
// with_cigar, KSW_EZ_APPROX_MAX | KSW_EZ_APPROX_DROP, zdrop = -1 (never drops; the H0 tracking only feeds the unused score),
// left-aligned gaps, no KSW_EZ_GENERIC_SC, no KSW_EZ_EXTZ_ONLY.
#include "restate.h"

#include <algorithm>
#include <cstring>

namespace restate {

namespace {

struct KswOut {
    bool zdropped = false;
    std::vector<uint32_t> cigar;   // len << 4 | op, op 0 = M, 1 = I (query only), 2 = D (target only)
};

inline int8_t s8(uint8_t x) { return (int8_t)x; }

// ksw_extz2_sse (ksw2_extz2_sse.c:26-304) for m = 5, the flags named above and end_bonus = 0
KswOut kswExtz2(const std::vector<uint8_t>& query, const std::vector<uint8_t>& target, const int8_t* mat, int q, int e, int w) {
    KswOut out;
    const int m = 5, qlen = (int)query.size(), tlen = (int)target.size();
    if (qlen <= 0 || tlen <= 0) return out;                                   // :64
    const uint8_t qe2 = (uint8_t)((q + e) * 2), qv = (uint8_t)q;
    const uint8_t scMch = (uint8_t)mat[0], scMis = (uint8_t)mat[1];
    const uint8_t scN = mat[m * m - 1] == 0 ? (uint8_t)(-e) : (uint8_t)mat[m * m - 1];   // :76
    const uint8_t wild = (uint8_t)(m - 1), maxSc = (uint8_t)(mat[0] + (q + e) * 2);      // :77-78
    if (w < 0) w = std::max(tlen, qlen);
    const int wl = w, wr = w;
    const int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16;
    int nCol_ = std::min(qlen, tlen);
    nCol_ = (std::min(nCol_, w + 1) + 15) / 16 + 1;                          // :84-85
    int minSc = mat[1];
    for (int t = 1; t < m * m; ++t) minSc = std::min<int>(minSc, mat[t]);
    if (-minSc > 2 * (q + e)) return out;                                     // :91

    // one zero-filled block: u, v, x, y, s (tlen_*16 bytes each), sf = target, qr = reversed query, 16 spare bytes (:93-95)
    const size_t T16 = (size_t)tlen_ * 16;
    std::vector<uint8_t> mem(T16 * 6 + (size_t)qlen_ * 16 + 16, 0);
    uint8_t* u = mem.data(); uint8_t* v = u + T16; uint8_t* x = v + T16; uint8_t* y = x + T16; uint8_t* s = y + T16;
    uint8_t* sf = s + T16; uint8_t* qr = sf + T16;
    const size_t nRounds = (size_t)qlen + tlen - 1;
    const size_t nCol = (size_t)nCol_ * 16;
    std::vector<uint8_t> p(nRounds * nCol + 16, 0);                           // :101-102 (only written cells are ever read)
    std::vector<int> off(nRounds), offEnd(nRounds);
    for (int t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];               // :107
    std::memcpy(sf, target.data(), tlen);

    int lastSt = -1, lastEn = -1;
    for (int r = 0; r < qlen + tlen - 1; ++r) {                               // :110
        int st = 0, en = tlen - 1;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < ((r - wr + 1) >> 1)) st = (r - wr + 1) >> 1;                 // :119 (arithmetic shift, as in C)
        if (en > ((r + wl) >> 1)) en = (r + wl) >> 1;
        if (st > en) { out.zdropped = true; break; }                          // :121-124
        const int st0 = st, en0 = en;
        st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;                      // :126
        uint8_t x1, v1;
        if (st > 0) {                                                         // :128-132
            if (st - 1 >= lastSt && st - 1 <= lastEn) { x1 = x[st - 1]; v1 = v[st - 1]; }
            else x1 = v1 = 0;
        } else { x1 = 0; v1 = r ? qv : 0; }
        if (en >= r) { y[r] = 0; u[r] = r ? qv : 0; }                         // :133
        // scores of the diagonal, in unaligned 16-byte strides from st0 (:135-152); the stride that covers en0 runs up to 15
        // bytes past it (into the next array of the block when en0 is near the end of s)
        const uint8_t* qrr = qr + (qlen - 1 - r);
        for (int t = st0; t <= en0; t += 16)
            for (int i = 0; i < 16; ++i) {
                const uint8_t sq = sf[t + i], sq2 = qrr[t + i];
                uint8_t sc = sq == sq2 ? scMch : scMis;
                if (sq == wild || sq2 == wild) sc = scN;
                s[t + i] = sc;
            }
        // core loop, left-aligned gaps (:179-208 with the SSE2 emulation branches), byte by byte over the aligned range
        uint8_t* pr = p.data() + (size_t)r * nCol - st;                       // pr[t] = byte of target position t
        off[r] = st; offEnd[r] = en;
        uint8_t carryX = x1, carryV = v1;
        for (int t = st; t <= en; ++t) {
            uint8_t z = (uint8_t)(s[t] + qe2);                                // __dp_code_block1
            const uint8_t xt1 = carryX; carryX = x[t];
            const uint8_t vt1 = carryV; carryV = v[t];
            uint8_t a = (uint8_t)(xt1 + vt1);
            const uint8_t ut = u[t];
            uint8_t b = (uint8_t)(y[t] + ut);
            uint8_t d = s8(a) > s8(z) ? 1 : 0;                                // :185
            z = s8(z) > 0 ? z : 0;                                            // :191
            z = std::max(z, a);                                               // :192 (unsigned)
            if (s8(b) > s8(z)) d = 2;                                         // :193-194
            z = std::max(z, b);                                               // __dp_code_block2 (unsigned max, unsigned min)
            z = std::min(z, maxSc);
            u[t] = (uint8_t)(z - vt1);
            v[t] = (uint8_t)(z - ut);
            z = (uint8_t)(z - qv);
            a = (uint8_t)(a - z);
            b = (uint8_t)(b - z);
            if (s8(a) > 0) { x[t] = a; d |= 0x08; } else x[t] = 0;           // :197-199
            if (s8(b) > 0) { y[t] = b; d |= 0x10; } else y[t] = 0;           // :200-202
            pr[t] = d;
        }
        lastSt = st; lastEn = en;                                             // :296
    }
    if (out.zdropped) return out;
    // ksw_backtrack (ksw2.h:119-154) with is_rot = 1, is_rev = 0, min_intron_len = 0, from (tlen-1, qlen-1)
    std::vector<uint32_t> rev;
    auto push = [&](uint32_t op, int len) {
        if (rev.empty() || op != (rev.back() & 0xf)) rev.push_back((uint32_t)len << 4 | op);
        else rev.back() += (uint32_t)len << 4;
    };
    int i = tlen - 1, j = qlen - 1, state = 0;
    while (i >= 0 && j >= 0) {
        int force = -1;
        const int r = i + j;
        if (i < off[r]) force = 2;
        if (i > offEnd[r]) force = 1;
        const uint32_t tmp = force < 0 ? p[(size_t)r * nCol + i - off[r]] : 0;
        if (state == 0) state = tmp & 7;
        else if (!(tmp >> (state + 2) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (force >= 0) state = force;
        if (state == 0) { push(0, 1); --i; --j; }
        else if (state == 1 || state == 3) { push(2, 1); --i; }
        else { push(1, 1); --j; }
    }
    if (i >= 0) push(2, i + 1);
    if (j >= 0) push(1, j + 1);
    out.cigar.assign(rev.rbegin(), rev.rend());
    return out;
}

struct CigOp { char op; int len; };

// getAlignmentCigarKsw (alignment.cpp:102-216): band 64, doubled while the band cannot connect the corners
std::vector<CigOp> alignmentCigar(const std::vector<uint8_t>& trg, const std::vector<uint8_t>& qry) {
    const int8_t a = 2, b = -4;
    const int8_t mat[25] = {a, b, b, b, 0, b, a, b, b, 0, b, b, a, b, 0, b, b, b, a, 0, 0, 0, 0, 0, 0};
    KswOut ez;
    int band = 64;
    for (;;) {
        ez = kswExtz2(qry, trg, mat, 4, 2, band);
        if (!ez.zdropped) break;
        if (band > (int)std::max(qry.size(), trg.size())) break;
        band *= 2;
    }
    std::vector<CigOp> cigar;
    size_t posQry = 0, posTrg = 0;
    for (uint32_t c : ez.cigar) {
        const int size = (int)(c >> 4);
        const char op = "MID"[c & 0xf];
        if (op == 'M') {
            for (int k = 0; k < size; ++k) {
                const char match = trg[posTrg + k] == qry[posQry + k] ? '=' : 'X';
                if (k == 0 || match != cigar.back().op) cigar.push_back({match, 1});
                else ++cigar.back().len;
            }
            posQry += size; posTrg += size;
        } else if (op == 'I') { cigar.push_back({'I', size}); posQry += size; }
        else { cigar.push_back({'D', size}); posTrg += size; }
    }
    return cigar;
}

struct Compressed { std::vector<uint8_t> seq; std::vector<int32_t> offsetTable; };

Compressed compress(const Reads& r, uint32_t id, int32_t start, int32_t length, bool doCompression) {   // alignment.cpp:52-70
    Compressed c;
    for (int32_t i = 0; i < length; ++i) {
        const uint8_t base = r.at(id, (size_t)start + i);
        if (!doCompression || i == 0 || c.seq.back() != base) { c.seq.push_back(base); c.offsetTable.push_back(i); }
    }
    return c;
}

}  // namespace

std::vector<std::pair<char, int>> alignmentCigarKsw(const std::vector<uint8_t>& trg, const std::vector<uint8_t>& qry) {
    std::vector<std::pair<char, int>> out;
    for (const CigOp& c : alignmentCigar(trg, qry)) out.emplace_back(c.op, c.len);
    return out;
}

// checkIdyAndTrim (alignment.cpp:306-495)
std::vector<Overlap> checkIdyAndTrim(const Reads& r, const Overlap& ovlp, float maxDivergence, int32_t minOverlap, bool useHpc) {
    const Compressed cur = compress(r, ovlp.curId, ovlp.curBegin, ovlp.curEnd - ovlp.curBegin, useHpc);
    const Compressed ext = compress(r, ovlp.extId, ovlp.extBegin, ovlp.extEnd - ovlp.extBegin, useHpc);
    const std::vector<CigOp> cigar = alignmentCigar(cur.seq, ext.seq);      // target = cur, query = ext (:319-321)
    std::vector<int> sumErrors{0}, sumCurLen{0}, sumExtLen{0};
    for (const CigOp& op : cigar) {                                           // :337-354
        int curConsumed = op.len, extConsumed = op.len;
        if (op.op == 'I') curConsumed = 0;
        else if (op.op == 'D') extConsumed = 0;
        const int errLen = op.op != '=' ? op.len : 0;
        sumCurLen.push_back(sumCurLen.back() + curConsumed);
        sumExtLen.push_back(sumExtLen.back() + extConsumed);
        sumErrors.push_back(sumErrors.back() + errLen);
    }
    struct Interval { int start, end; float divergence; int realLen; };
    std::vector<Interval> good;
    const int n = (int)cigar.size();
    for (int intLen = n; intLen > 0; --intLen)                                // :367-388
        for (int intStart = 0; intStart < n - intLen + 1; ++intStart) {
            const int i = intStart, j = intStart + intLen - 1;
            if (cigar[i].op != '=' || cigar[j].op != '=') continue;
            const int rangeLen = std::max(sumCurLen[j + 1] - sumCurLen[i], sumExtLen[j + 1] - sumExtLen[i]);
            const int rangeErr = sumErrors[j + 1] - sumErrors[i];
            const float divergence = float(rangeErr) / rangeLen;
            if (divergence < maxDivergence) good.push_back({i, j, divergence, rangeLen});
        }
    // std::sort by realLen descending (:391-393): libstdc++'s permutation decides which of several equally long intervals the greedy
    // pass below sees first — the model of restate.cpp reproduces it (keys ascending with `<` = lengths descending with `>`)
    {
        std::vector<uint64_t> keys(good.size());
        for (size_t i = 0; i < good.size(); ++i) keys[i] = (uint64_t)(0x7fffffff - good[i].realLen);
        const std::vector<uint32_t> perm = introsort_model(keys);
        std::vector<Interval> sorted(good.size());
        for (size_t i = 0; i < good.size(); ++i) sorted[i] = good[perm[i]];
        good.swap(sorted);
    }
    std::vector<Interval> chosen;                                             // :396-411
    for (const Interval& iv : good) {
        bool intersects = false;
        for (const Interval& o : chosen)
            if (std::min(iv.end + 1, o.end + 1) - std::max(iv.start, o.start) > 0) { intersects = true; break; }
        if (!intersects) chosen.push_back(iv);
    }
    std::vector<Overlap> trimmed;                                             // :415-452
    for (const Interval& cand : chosen) {
        Overlap o = ovlp;
        o.seqDivergence = cand.divergence;
        size_t posQry = 0, posTrg = 0;
        for (int i = 0; i < n; ++i) {
            if (i == cand.start) { o.curBegin += cur.offsetTable[posTrg]; o.extBegin += ext.offsetTable[posQry]; }
            if (cigar[i].op == '=' || cigar[i].op == 'X') { posQry += cigar[i].len; posTrg += cigar[i].len; }
            else if (cigar[i].op == 'I') posQry += cigar[i].len;
            else posTrg += cigar[i].len;
            if (i == cand.end) { o.curEnd = ovlp.curBegin + cur.offsetTable[posTrg - 1]; o.extEnd = ovlp.extBegin + ext.offsetTable[posQry - 1]; }
        }
        if (o.curEnd - o.curBegin > minOverlap && o.extEnd - o.extBegin > minOverlap) trimmed.push_back(o);
    }
    return trimmed;
}

}  // namespace restate
