// TEST INFRASTRUCTURE — not product code.  SURVEY §8f N3: getAlignmentCigarKsw and checkIdyAndTrim (alignment.cpp:102-216,
// 306-495; ksw2 inside) on generated sequence pairs — one driver, two builds whose outputs must be identical byte for byte
// (tests/test_oracle_trim.py):
//   (1) the UNMODIFIED reference (-DTRIM_REF: alignment.cpp + minimap2's ksw2_extz2_sse.c / kalloc.c as oracle/Makefile builds them),
//   (2) the CPU restatement oracle/restate/ksw_restate.cpp (no reference header).
// Cases: a target and a copy of it with block-wise varying divergence (clean blocks between noisy ones, so that the trimming
// finds several pieces), homopolymer-rich stretches, length differences beyond the first band (the band is doubled), either
// strand of either sequence, with and without homopolymer compression, three thresholds.
// Usage: trim_check N_CASES SEED      prints per case the CIGAR (digest + head) and the trimmed overlaps.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef TRIM_REF
#include "sequence/sequence_container.h"
#include "sequence/alignment.h"
#else
#include "restate/restate.h"
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

int main(int argc, char** argv) {
    const int nCases = argc > 1 ? atoi(argv[1]) : 60;
    if (argc > 2) rngState = (uint32_t)strtoul(argv[2], nullptr, 10) | 1u;
    for (int c = 0; c < nCases; ++c) {
        const size_t n = c % 7 == 0 ? 2500 + rnd() % 4000 : 200 + rnd() % 1500;
        const std::string t = makeTarget(n);
        const std::string q = mutate(t, c % 5 == 0 ? 6 : 0);
        const bool curRc = c % 4 == 1, extRc = c % 4 == 2;
        const bool useHpc = c % 2 == 0;
        const float maxDiv = c % 3 == 0 ? 0.05f : c % 3 == 1 ? 0.10f : 0.20f;
        const int32_t minOverlap = c % 2 ? 50 : 150;
        // the overlap: most of both sequences (the stored sequence of a reverse strand is the reverse complement of what is aligned)
        const int32_t curBegin = (int32_t)(rnd() % 10), curEnd = (int32_t)t.size() - (int32_t)(rnd() % 10);
        const int32_t extBegin = (int32_t)(rnd() % 10), extEnd = (int32_t)q.size() - (int32_t)(rnd() % 10);
        const std::string storedCur = curRc ? revComp(t) : t, storedExt = extRc ? revComp(q) : q;
        printf("case %d cur %zu%s ext %zu%s hpc %d maxDiv %.2f minOvlp %d\n", c, t.size(), curRc ? "-" : "+", q.size(), extRc ? "-" : "+", (int)useHpc,
               (double)maxDiv, minOverlap);
        std::vector<std::pair<char, int>> cigar;
        struct Piece { int32_t cb, ce, eb, ee; float div; };
        std::vector<Piece> pieces;
#ifdef TRIM_REF
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
        for (auto& p : checkIdyAndTrim(ov, curSeq, extSeq, maxDiv, minOverlap, useHpc)) pieces.push_back({p.curBegin, p.curEnd, p.extBegin, p.extEnd, p.seqDivergence});
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
        for (auto& p : restate::checkIdyAndTrim(reads, ov, maxDiv, minOverlap, useHpc)) pieces.push_back({p.curBegin, p.curEnd, p.extBegin, p.extEnd, p.seqDivergence});
#endif
        uint64_t h = 1469598103934665603ULL;
        for (auto& o : cigar) { h = fnv(h, (uint64_t)o.first); h = fnv(h, (uint64_t)o.second); }
        printf("  cigar ops %zu digest %016" PRIx64 " head", cigar.size(), h);
        for (size_t i = 0; i < cigar.size() && i < 12; ++i) printf(" %d%c", cigar[i].second, cigar[i].first);
        printf("\n  pieces %zu\n", pieces.size());
        for (auto& p : pieces) { uint32_t bits; memcpy(&bits, &p.div, 4); printf("    cur [%d,%d) ext [%d,%d) div %08x\n", p.cb, p.ce, p.eb, p.ee, bits); }
    }
    return 0;
}
