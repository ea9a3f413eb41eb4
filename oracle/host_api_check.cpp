// TEST INFRASTRUCTURE — not product code.  The value types of the drop-in boundary, exercised
// through their public interface on random data — one driver, two builds (the unmodified reference; the host mirror in
// flye_b200/host), whose outputs must be identical byte for byte (tests/test_host_api.py):
//   DnaSequence   str / at / atRaw / substr / complement (views and views of views)
//   Kmer          construction, reverseComplement, standardForm, appendLeft / appendRight, hash, comparison
//   IterKmers     whole sequences and (start, length) windows, forward and complement views (positions 0 .. L-k-1)
//   yieldMinimizers   windows 1, 5, 10, 19
//   OverlapRange  reverse, complement, ranges, shifts, overhang, contains / containedBy, intersections, project (with and
//                 without kmerMatches), copies
//   SequenceContainer::writeFasta   both modes
// Usage: host_api_check OUT_PREFIX   (writes OUT_PREFIX.fasta / .pos.fasta, prints everything else)
#include <cinttypes>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "sequence/sequence_container.h"
#include "sequence/kmer.h"
#include "sequence/overlap.h"
#include "common/config.h"

static uint32_t rngState = 2463534242u;
static uint32_t rnd() { rngState ^= rngState << 13; rngState ^= rngState >> 17; rngState ^= rngState << 5; return rngState; }

static uint32_t idNum(FastaRecord::Id id) {
    int s = id.signedId();
    return s > 0 ? (uint32_t)(s - 1) * 2 : (uint32_t)(-s - 1) * 2 + 1;
}
static uint64_t fnv(uint64_t h, uint64_t v) { for (int i = 0; i < 8; ++i) { h ^= (v >> (8 * i)) & 255u; h *= 1099511628211ULL; } return h; }

static std::string randomSeq(size_t n, bool lowComplexity) {
    std::string s(n, 'A');
    for (size_t i = 0; i < n; ++i) {
        if (lowComplexity && i >= 3 && rnd() % 4) s[i] = s[i - 3];   // short tandem repeats: equal hashes inside minimizer windows
        else s[i] = "ACGT"[rnd() % 4];
    }
    return s;
}

static void kmerChecks(const DnaSequence& seq, const char* tag) {
    const size_t k = Parameters::get().kmerSize;
    uint64_t h = 1469598103934665603ULL; size_t n = 0; int printed = 0;
    for (auto kp : IterKmers(seq)) {
        Kmer km = kp.kmer;
        Kmer rc = km.reverseComplement();
        Kmer canon = km;
        const bool flipped = canon.standardForm();
        h = fnv(h, kp.position); h = fnv(h, km.numRepr()); h = fnv(h, rc.numRepr()); h = fnv(h, canon.numRepr()); h = fnv(h, flipped);
        h = fnv(h, canon.hash()); h = fnv(h, std::hash<Kmer>()(km)); h = fnv(h, (uint64_t)(km == rc) * 2 + (uint64_t)(km != canon));
        if (printed < 3) { printf("  %s kmer pos %d repr %zx rc %zx canon %zx flipped %d hash %zx\n", tag, kp.position, km.numRepr(), rc.numRepr(), canon.numRepr(), (int)flipped, canon.hash()); ++printed; }
        ++n;
    }
    printf("  %s IterKmers n %zu digest %016" PRIx64 "\n", tag, n, h);
    if (seq.length() > k + 40) {   // a window
        uint64_t h2 = 1469598103934665603ULL; size_t n2 = 0;
        for (auto kp : IterKmers(seq, 7, k + 25)) { h2 = fnv(h2, kp.position); h2 = fnv(h2, kp.kmer.numRepr()); ++n2; }
        printf("  %s IterKmers(7, k+25) n %zu digest %016" PRIx64 "\n", tag, n2, h2);
        Kmer a(seq, 5, k), b(seq, 6, k);
        Kmer a2 = a; a2.appendRight(seq.atRaw(5 + k));
        Kmer b2 = b; b2.appendLeft(seq.atRaw(5));
        printf("  %s append: right %d left %d less %d\n", tag, (int)(a2 == b), (int)(b2 == a), (int)(a < b));
    }
    for (int w : {1, 5, 10, 19}) {
        uint64_t hm = 1469598103934665603ULL;
        const auto mins = yieldMinimizers(seq, w);
        for (const auto& kp : mins) { Kmer km = kp.kmer; hm = fnv(hm, kp.position); hm = fnv(hm, km.numRepr()); }
        printf("  %s minimizers w=%d n %zu digest %016" PRIx64 " first %d last %d\n", tag, w, mins.size(), hm,
               mins.empty() ? -1 : mins.front().position, mins.empty() ? -1 : mins.back().position);
    }
}

static void printOvlp(const char* tag, const OverlapRange& o) {
    printf("  %s cur %u [%d,%d) %d ext %u [%d,%d) %d score %d div %.9g ranges %d %d %d shifts %d %d overhang %d", tag, idNum(o.curId), o.curBegin,
           o.curEnd, o.curLen, idNum(o.extId), o.extBegin, o.extEnd, o.extLen, o.score, (double)o.seqDivergence, o.curRange(), o.extRange(),
           o.minRange(), o.leftShift(), o.rightShift(), o.lrOverhang());
    if (o.kmerMatches) { printf(" matches"); for (auto& p : *o.kmerMatches) printf(" %d,%d", p.first, p.second); }
    printf("\n");
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "host_api_check OUT_PREFIX\n"); return 2; }
    const std::string out = argv[1];
    SequenceContainer c;
    std::vector<std::string> texts;
    for (int i = 0; i < 14; ++i) {
        const size_t n = i == 0 ? 15 : i == 1 ? 17 : i == 2 ? 18 : 30 + rnd() % 700;
        texts.push_back(randomSeq(n, i % 3 == 2));
        c.addSequence(DnaSequence(texts.back()), "seq_" + std::to_string(i));
    }
    c.buildPositionIndex();
    const uint32_t firstId = idNum(c.iterSeqs().front().id);
    for (size_t k : {(size_t)15, (size_t)17}) {
        Parameters::get().kmerSize = k;
        printf("k = %zu\n", k);
        for (const auto& rec : c.iterSeqs()) {
            const DnaSequence& s = rec.sequence;
            const std::string tag = rec.description;
            printf(" %s id %u len %zu str %s\n", tag.c_str(), idNum(rec.id) - firstId, s.length(), k == 15 ? s.str().c_str() : "-");
            if (k == 15) {
                const DnaSequence cc = s.complement();        // (a complement of a complement view stays a complement view)
                const size_t a = s.length() / 3, l = s.length() / 2;
                printf("  substr %s | compl.substr %s | compl.compl %s | at %c%c raw %d%d\n", s.substr(a, l).str().c_str(), cc.substr(a, l).str().c_str(),
                       cc.complement().str().c_str(), s.at(0), s.at(s.length() - 1), (int)s.atRaw(1), (int)cc.atRaw(1));
            }
            if (s.length() >= k) kmerChecks(s, tag.c_str());
        }
    }
    // OverlapRange
    const auto& seqs = c.iterSeqs();
    std::vector<OverlapRange> ovs;
    for (int i = 0; i < 10; ++i) {
        OverlapRange o;
        const FastaRecord& cur = seqs[6 + rnd() % (seqs.size() - 6)];
        const FastaRecord& ext = seqs[6 + rnd() % (seqs.size() - 6)];
        o.curId = cur.id; o.curLen = (int32_t)cur.sequence.length(); o.curBegin = (int32_t)(rnd() % 10); o.curEnd = o.curLen - 1 - (int32_t)(rnd() % 10);
        o.extId = ext.id; o.extLen = (int32_t)ext.sequence.length(); o.extBegin = (int32_t)(rnd() % 10); o.extEnd = o.extLen - 1 - (int32_t)(rnd() % 10);
        o.score = (int32_t)(rnd() % 5000); o.seqDivergence = (float)(rnd() % 1000) / 4096.0f;
        if (i % 2) {
            o.kmerMatches = new std::vector<std::pair<int32_t, int32_t>>();
            const int m = 6;
            for (int j = 0; j <= m; ++j)
                o.kmerMatches->push_back({o.curBegin + (o.curEnd - o.curBegin) * j / m, o.extBegin + (o.extEnd - o.extBegin) * j / m});
        }
        ovs.push_back(o);   // (deep copy of kmerMatches)
    }
    for (size_t i = 0; i < ovs.size(); ++i) {
        const OverlapRange& o = ovs[i];
        printf("overlap %zu\n", i);
        printOvlp("plain", o);
        printOvlp("reverse", o.reverse());
        printOvlp("complement", o.complement());
        printOvlp("rev.compl", o.reverse().complement());
        OverlapRange copy(o), assigned; assigned = o.complement();
        printOvlp("copy", copy); printOvlp("assigned", assigned);
        for (int32_t p : {o.curBegin - 3, o.curBegin, o.curBegin + 1, (o.curBegin + o.curEnd) / 2, o.curEnd - 1, o.curEnd, o.curEnd + 5}) {
            int32_t pr = -1; const char* err = "";
            try { pr = o.project(p); } catch (const std::exception& e) { err = e.what(); }
            printf("  project %d -> %d %s | contains %d\n", p, pr, err, (int)o.contains(p, pr));
        }
        const OverlapRange& q = ovs[(i + 1) % ovs.size()];
        OverlapRange inner(o); inner.curBegin += 2; inner.extEnd -= 2;
        printf("  containedBy next %d inner-in-o %d o-in-inner %d curIntersect %d extIntersect %d\n", (int)o.containedBy(q), (int)inner.containedBy(o),
               (int)o.containedBy(inner), o.curIntersect(q), o.extIntersect(q));
    }
    // writeFasta (the reference never closes the file: its bytes are complete when the process has exited — the test compares
    // the two files afterwards)
    SequenceContainer::writeFasta(c.iterSeqs(), out + ".fasta", false);
    SequenceContainer::writeFasta(c.iterSeqs(), out + ".pos.fasta", true);
    return 0;
}
