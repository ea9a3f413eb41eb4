// TEST INFRASTRUCTURE — not product code.  SURVEY §8f N3: getAlignmentCigarKsw and checkIdyAndTrim (alignment.cpp:102-216,
// 306-495; ksw2 inside) on generated sequence pairs — one driver, two builds whose outputs must be identical byte for byte
// (tests/test_oracle_trim.py):
//   (1) the UNMODIFIED reference (-DTRIM_REF: alignment.cpp + minimap2's ksw2_extz2_sse.c / kalloc.c as oracle/Makefile builds them),
//   (2) the CPU restatement oracle/restate/ksw_restate.cpp (no reference header).
// Cases: a target and a copy of it with block-wise varying divergence (clean blocks between noisy ones, so that the trimming
// finds several pieces), homopolymer-rich stretches, length differences beyond the first band (the band is doubled), either
// strand of either sequence, with and without homopolymer compression, three thresholds.
// Two more builds print the CIGAR lines only (the trimming itself stays host code):
//   (3) -DTRIM_CORE    the host build of flye_b200/csrc/ksw_core.cuh — the scalar routine the device kernel runs,
//   (4) -DTRIM_DEVICE  the device build of the same routine through the C ABI (fg_align_cigar_batch, one batch for all cases).
// And one more prints everything again:
//   (5) -DTRIM_MIRROR  the host mirror's own trimming tail (flye_b200/host/sequence/overlap.h: hpcRange + trimByCigar — what
//                      FLYE_B200_DEVICE_KSW=1 runs around the device alignment) fed with CIGARs of the routine's host build.
// Usage: trim_check N_CASES SEED      prints per case the CIGAR (digest + head) and the trimmed overlaps.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#if defined(TRIM_REF)
#include "sequence/sequence_container.h"
#include "sequence/alignment.h"
#elif defined(TRIM_CORE)
#include "../flye_b200/csrc/ksw_core.cuh"
#elif defined(TRIM_MIRROR)
#include "sequence/sequence_container.h"
#include "sequence/overlap.h"
#include "../flye_b200/csrc/ksw_core.cuh"
#elif defined(TRIM_DEVICE)
#include "flye_b200.h"
#else
#include "restate/restate.h"
#endif

#if defined(TRIM_CORE) || defined(TRIM_MIRROR)
// getAlignmentCigarKsw's band loop (alignment.cpp:150-165) around the shared scalar routine; the ksw-level CIGAR
static std::vector<uint32_t> coreCigar(const std::vector<uint8_t>& trg, const std::vector<uint8_t>& qry) {
    std::vector<uint32_t> cg(trg.size() + qry.size() + 2);
    int nCg = 0;
    for (int band = 64;; band *= 2) {
        const fg::KswSizes z = fg::kswSizes((int)qry.size(), (int)trg.size(), band);
        std::vector<uint8_t> mem(z.memBytes, 0xAB), p(z.pBytes, 0xCD);   // (the routine must not depend on what the scratch held)
        std::vector<int> off(2 * z.rounds);
        const int rc = fg::kswExtz2Core(qry.data(), (int)qry.size(), trg.data(), (int)trg.size(), band, mem.data(), p.data(), off.data(),
                                        off.data() + z.rounds, cg.data(), (int)cg.size(), &nCg);
        if (rc == fg::KSW_CIGAR_OVERFLOW) { fprintf(stderr, "cigar capacity\n"); exit(3); }
        if (rc != fg::KSW_BAND_TOO_NARROW) break;
        if (band > (int)std::max(qry.size(), trg.size())) { nCg = 0; break; }
    }
    cg.resize(nCg);
    return cg;
}
#endif

#if defined(TRIM_CORE) || defined(TRIM_DEVICE)
#define TRIM_CIGAR_ONLY 1
#endif
#if defined(TRIM_CORE) || defined(TRIM_DEVICE) || defined(TRIM_MIRROR)
static std::vector<uint8_t> baseCodes(const std::string& s) {
    std::vector<uint8_t> v(s.size());
    for (size_t i = 0; i < s.size(); ++i) v[i] = s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'G' ? 2 : 3;
    return v;
}
// ksw-level CIGAR (len << 4 | op) -> the =, X, I, D runs of getAlignmentCigarKsw (alignment.cpp:172-211)
static std::vector<std::pair<char, int>> decodeKsw(const uint32_t* cg, int n, const std::vector<uint8_t>& trg, const std::vector<uint8_t>& qry) {
    std::vector<std::pair<char, int>> out;
    size_t posQry = 0, posTrg = 0;
    for (int c = 0; c < n; ++c) {
        const int size = (int)(cg[c] >> 4);
        const char op = "MID"[cg[c] & 0xf];
        if (op == 'M') {
            for (int k = 0; k < size; ++k) {
                const char match = trg[posTrg + k] == qry[posQry + k] ? '=' : 'X';
                if (k == 0 || match != out.back().first) out.emplace_back(match, 1);
                else ++out.back().second;
            }
            posQry += size; posTrg += size;
        } else if (op == 'I') { out.emplace_back('I', size); posQry += size; }
        else { out.emplace_back('D', size); posTrg += size; }
    }
    return out;
}
#endif

static uint32_t rngState = 88172645u;
static uint32_t rnd() { rngState ^= rngState << 13; rngState ^= rngState >> 17; rngState ^= rngState << 5; return rngState; }
static uint64_t fnv(uint64_t h, uint64_t v) { for (int i = 0; i < 8; ++i) { h ^= (v >> (8 * i)) & 255u; h *= 1099511628211ULL; } return h; }

static std::string makeTarget(size_t n) {
    std::string s;
    while (s.size() < n) {
        const char c = "ACGT"[rnd() % 4];
        const int run = (rnd() % 5 == 0) ? 2 + rnd() % 6 : 1;   // homopolymer runs
        s.append((size_t)run, c);
    }
    s.resize(n);
    return s;
}

// copy with errors: blocks of 80-400 bases with their own error rate (mostly low, some very high)
static std::string mutate(const std::string& t, int bias) {
    std::string q;
    size_t i = 0;
    while (i < t.size()) {
        const size_t block = 80 + rnd() % 320;
        const uint32_t kind = rnd() % 10;
        const uint32_t rate = kind < 6 ? 20 + rnd() % 40 : kind < 8 ? 150 : 400;   // per mille
        for (size_t e = std::min(t.size(), i + block); i < e; ++i) {
            if (rnd() % 1000 < rate) {
                const uint32_t what = rnd() % (4 + bias);
                if (what == 0) { q += "ACGT"[rnd() % 4]; }                       // substitution
                else if (what == 1) { q += t[i]; q += "ACGT"[rnd() % 4]; }       // insertion
                else if (what == 2) { }                                            // deletion
                else if (what == 3) { q += t[i]; q += t[i]; }                      // homopolymer extension
                else { }                                                           // (bias: more deletions -> length difference)
            } else q += t[i];
        }
    }
    if (q.size() < 40) q += makeTarget(40);
    return q;
}

static std::string revComp(const std::string& s) {
    std::string r(s.rbegin(), s.rend());
    for (char& c : r) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
    return r;
}

struct Case {
    std::string t, q;
    bool curRc, extRc, useHpc; float maxDiv; int32_t minOverlap, curBegin, curEnd, extBegin, extEnd;
};
struct Piece { int32_t cb, ce, eb, ee; float div; };

int main(int argc, char** argv) {
    const int nCases = argc > 1 ? atoi(argv[1]) : 60;
    if (argc > 2) rngState = (uint32_t)strtoul(argv[2], nullptr, 10) | 1u;
    std::vector<Case> cases(nCases);
    for (int c = 0; c < nCases; ++c) {
        Case& k = cases[c];
        const size_t n = c % 7 == 0 ? 2500 + rnd() % 4000 : 200 + rnd() % 1500;
        k.t = makeTarget(n);
        k.q = mutate(k.t, c % 5 == 0 ? 6 : 0);
        k.curRc = c % 4 == 1; k.extRc = c % 4 == 2;
        k.useHpc = c % 2 == 0;
        k.maxDiv = c % 3 == 0 ? 0.05f : c % 3 == 1 ? 0.10f : 0.20f;
        k.minOverlap = c % 2 ? 50 : 150;
        // the overlap: most of both sequences (the stored sequence of a reverse strand is the reverse complement of what is aligned)
        k.curBegin = (int32_t)(rnd() % 10); k.curEnd = (int32_t)k.t.size() - (int32_t)(rnd() % 10);
        k.extBegin = (int32_t)(rnd() % 10); k.extEnd = (int32_t)k.q.size() - (int32_t)(rnd() % 10);
    }
    std::vector<std::vector<std::pair<char, int>>> cigars(nCases);
    std::vector<std::vector<Piece>> pieces(nCases);
#if defined(TRIM_DEVICE)
    {   // one batch through the C ABI
        std::vector<uint8_t> T, Q; std::vector<uint64_t> tOff{0}, qOff{0};
        std::vector<std::vector<uint8_t>> trgs(nCases), qrys(nCases);
        for (int c = 0; c < nCases; ++c) {
            const Case& k = cases[c];
            trgs[c] = baseCodes(k.t.substr(k.curBegin, k.curEnd - k.curBegin));
            qrys[c] = baseCodes(k.q.substr(k.extBegin, k.extEnd - k.extBegin));
            T.insert(T.end(), trgs[c].begin(), trgs[c].end()); tOff.push_back(T.size());
            Q.insert(Q.end(), qrys[c].begin(), qrys[c].end()); qOff.push_back(Q.size());
        }
        fg_ctx* ctx = nullptr;
        if (fg_ctx_create(0, &ctx) != FG_OK) { fprintf(stderr, "fg_ctx_create failed\n"); return 3; }
        const uint32_t cap = 16384;
        std::vector<uint32_t> cg((size_t)nCases * cap), nCg(nCases);
        std::vector<int32_t> status(nCases);
        if (fg_align_cigar_batch(ctx, T.data(), tOff.data(), Q.data(), qOff.data(), (uint32_t)nCases, cap, cg.data(), nCg.data(), status.data()) != FG_OK) {
            fprintf(stderr, "fg_align_cigar_batch: %s\n", fg_last_error(ctx));
            return 3;
        }
        for (int c = 0; c < nCases; ++c) {
            if (status[c] == 2) { fprintf(stderr, "case %d: cigar capacity\n", c); return 3; }
            cigars[c] = decodeKsw(cg.data() + (size_t)c * cap, (int)nCg[c], trgs[c], qrys[c]);
        }
        fg_ctx_destroy(ctx);
    }
#endif
    for (int c = 0; c < nCases; ++c) {
        const Case& k = cases[c];
        const std::string& t = k.t; const std::string& q = k.q;
        const bool curRc = k.curRc, extRc = k.extRc, useHpc = k.useHpc;
        const float maxDiv = k.maxDiv;
        const int32_t minOverlap = k.minOverlap, curBegin = k.curBegin, curEnd = k.curEnd, extBegin = k.extBegin, extEnd = k.extEnd;
        const std::string storedCur = curRc ? revComp(t) : t, storedExt = extRc ? revComp(q) : q;
        (void)storedCur; (void)storedExt; (void)useHpc; (void)maxDiv; (void)minOverlap;
        std::vector<std::pair<char, int>>& cigar = cigars[c];
        std::vector<Piece>& pcs = pieces[c];
        printf("case %d cur %zu%s ext %zu%s hpc %d maxDiv %.2f minOvlp %d\n", c, t.size(), curRc ? "-" : "+", q.size(), extRc ? "-" : "+", (int)useHpc,
               (double)maxDiv, minOverlap);
#if defined(TRIM_REF)
        SequenceContainer sc;
        const FastaRecord& curRec = sc.addSequence(DnaSequence(storedCur), "cur");
        const FastaRecord::Id curFwd = curRec.id;
        const FastaRecord& extRec = sc.addSequence(DnaSequence(storedExt), "ext");
        const FastaRecord::Id extFwd = extRec.id;
        const FastaRecord::Id curId = curRc ? curFwd.rc() : curFwd, extId = extRc ? extFwd.rc() : extFwd;
        const DnaSequence& curSeq = sc.getSeq(curId);
        const DnaSequence& extSeq = sc.getSeq(extId);
        std::vector<CigOp> cg;
        getAlignmentCigarKsw(curSeq, curBegin, curEnd - curBegin, extSeq, extBegin, extEnd - extBegin, maxDiv, cg);
        for (auto& o : cg) cigar.emplace_back(o.op, o.len);
        OverlapRange ov;
        ov.curId = curId; ov.curBegin = curBegin; ov.curEnd = curEnd; ov.curLen = (int32_t)t.size();
        ov.extId = extId; ov.extBegin = extBegin; ov.extEnd = extEnd; ov.extLen = (int32_t)q.size();
        ov.score = 1234; ov.seqDivergence = 0.5f;
        for (auto& p : checkIdyAndTrim(ov, curSeq, extSeq, maxDiv, minOverlap, useHpc)) pcs.push_back({p.curBegin, p.curEnd, p.extBegin, p.extEnd, p.seqDivergence});
#elif defined(TRIM_CORE)
        {
            const std::vector<uint8_t> trg = baseCodes(t.substr(curBegin, curEnd - curBegin)), qry = baseCodes(q.substr(extBegin, extEnd - extBegin));
            const std::vector<uint32_t> cg = coreCigar(trg, qry);
            cigar = decodeKsw(cg.data(), (int)cg.size(), trg, qry);
        }
#elif defined(TRIM_MIRROR)
        {
            SequenceContainer sc;
            const FastaRecord::Id curFwd = sc.addSequence(DnaSequence(storedCur), "cur").id;
            const FastaRecord::Id extFwd = sc.addSequence(DnaSequence(storedExt), "ext").id;
            const FastaRecord::Id curId = curRc ? curFwd.rc() : curFwd, extId = extRc ? extFwd.rc() : extFwd;
            const flye_b200::HpcRange rawCur = flye_b200::hpcRange(sc.getSeq(curId), curBegin, curEnd - curBegin, false);
            const flye_b200::HpcRange rawExt = flye_b200::hpcRange(sc.getSeq(extId), extBegin, extEnd - extBegin, false);
            const std::vector<uint32_t> cgRaw = coreCigar(rawCur.seq, rawExt.seq);
            cigar = decodeKsw(cgRaw.data(), (int)cgRaw.size(), rawCur.seq, rawExt.seq);
            OverlapRange ov;
            ov.curId = curId; ov.curBegin = curBegin; ov.curEnd = curEnd; ov.curLen = (int32_t)t.size();
            ov.extId = extId; ov.extBegin = extBegin; ov.extEnd = extEnd; ov.extLen = (int32_t)q.size();
            ov.score = 1234; ov.seqDivergence = 0.5f;
            const flye_b200::HpcRange hc = flye_b200::hpcRange(sc.getSeq(curId), ov.curBegin, ov.curRange(), useHpc);
            const flye_b200::HpcRange he = flye_b200::hpcRange(sc.getSeq(extId), ov.extBegin, ov.extRange(), useHpc);
            const std::vector<uint32_t> cg = coreCigar(hc.seq, he.seq);
            for (auto& p : flye_b200::trimByCigar(ov, hc, he, cg.data(), cg.size(), maxDiv, minOverlap))
                pcs.push_back({p.curBegin, p.curEnd, p.extBegin, p.extEnd, p.seqDivergence});
        }
#elif defined(TRIM_DEVICE)
        // (computed above)
#else
        restate::Reads reads;
        auto codes = [](const std::string& s) { std::vector<uint8_t> v(s.size()); for (size_t i = 0; i < s.size(); ++i) v[i] = s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'G' ? 2 : 3; return v; };
        reads.fwd.push_back(codes(storedCur)); reads.fwd.push_back(codes(storedExt));
        reads.buildOffsets();
        const uint32_t curId = curRc ? 1u : 0u, extId = extRc ? 3u : 2u;
        std::vector<uint8_t> trg, qry;
        for (int32_t i = curBegin; i < curEnd; ++i) trg.push_back(reads.at(curId, (size_t)i));
        for (int32_t i = extBegin; i < extEnd; ++i) qry.push_back(reads.at(extId, (size_t)i));
        cigar = restate::alignmentCigarKsw(trg, qry);
        restate::Overlap ov;
        ov.curId = curId; ov.curBegin = curBegin; ov.curEnd = curEnd; ov.curLen = (int32_t)t.size();
        ov.extId = extId; ov.extBegin = extBegin; ov.extEnd = extEnd; ov.extLen = (int32_t)q.size();
        ov.score = 1234; ov.seqDivergence = 0.5f;
        for (auto& p : restate::checkIdyAndTrim(reads, ov, maxDiv, minOverlap, useHpc)) pcs.push_back({p.curBegin, p.curEnd, p.extBegin, p.extEnd, p.seqDivergence});
#endif
        uint64_t h = 1469598103934665603ULL;
        for (auto& o : cigar) { h = fnv(h, (uint64_t)o.first); h = fnv(h, (uint64_t)o.second); }
        printf("  cigar ops %zu digest %016" PRIx64 " head", cigar.size(), h);
        for (size_t i = 0; i < cigar.size() && i < 12; ++i) printf(" %d%c", cigar[i].second, cigar[i].first);
#ifdef TRIM_CIGAR_ONLY
        printf("\n");
#else
        printf("\n  pieces %zu\n", pcs.size());
        for (auto& p : pcs) { uint32_t bits; memcpy(&bits, &p.div, 4); printf("    cur [%d,%d) ext [%d,%d) div %08x\n", p.cb, p.ce, p.eb, p.ee, bits); }
#endif
    }
    return 0;
}
