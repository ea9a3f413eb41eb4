// TEST INFRASTRUCTURE — not product code.  SURVEY §8(f) N4: read ingestion (FASTA / FASTQ, plain or gzip) and the OverlapRange
// text format (overlap.h:227-251), checked the way oracle/harness.cpp checks the hot path — one driver, two builds:
//   (1) the UNMODIFIED reference (sequence_container.cpp / sequence.cpp under /root/reference/src, build container only), and
//   (2) the host mirror in flye_b200/host (header-only above the C ABI),
// whose outputs must be identical byte for byte (tests/test_host_ingest.py).
// Usage: ingest_check MINLEN FILE...   prints, per file: every record (id, name, length, sequence, reverse-complement id), the
// global-position round trip of a few positions, N50; then a dump -> load -> dump round trip of OverlapRange records; a file the
// parser rejects prints the exception text.
#include <cinttypes>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "sequence/sequence_container.h"
#include "sequence/overlap.h"

static uint32_t idNum(FastaRecord::Id id) {
    int s = id.signedId();
    return s > 0 ? (uint32_t)(s - 1) * 2 : (uint32_t)(-s - 1) * 2 + 1;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "ingest_check MINLEN FILE...\n"); return 2; }
    const int minLen = atoi(argv[1]);
    for (int a = 2; a < argc; ++a) {
        const std::string path = argv[a];
        const std::string base = path.substr(path.find_last_of('/') + 1);
        SequenceContainer c;
        try {
            c.loadFromFile(path, minLen);
        } catch (const std::exception& e) {
            printf("FILE %s ERROR %s\n", base.c_str(), e.what());
            continue;
        }
        c.buildPositionIndex();
        printf("FILE %s records %zu n50 %d\n", base.c_str(), c.iterSeqs().size(), c.iterSeqs().empty() ? 0 : c.computeNxStat(0.5f));
        uint32_t firstId = 0; bool haveFirst = false;
        for (const auto& rec : c.iterSeqs()) {
            if (!haveFirst) { firstId = idNum(rec.id); haveFirst = true; }
            const size_t L = rec.sequence.length();
            printf("%u %s %zu %s rc=%u\n", idNum(rec.id) - firstId, rec.description.c_str(), L, rec.sequence.str().c_str(),
                   idNum(rec.id.rc()) - firstId);
            for (size_t p : {(size_t)0, L / 2, L - 1}) {
                if (!L) break;
                const size_t g = c.globalPosition(rec.id, (int32_t)p);
                FastaRecord::Id oid; int32_t opos = -1, olen = -1;
                c.seqPosition(g, oid, opos, olen);
                printf("  pos %zu -> id %u pos %d len %d\n", p, idNum(oid) - firstId, opos, olen);
            }
        }
        // OverlapRange text round trip over this container's names
        const auto& seqs = c.iterSeqs();
        if (seqs.size() >= 4) {
            std::stringstream ss;
            uint32_t x = 12345;
            auto rnd = [&] { x = x * 1664525u + 1013904223u; return x >> 8; };
            const int n = 8;
            for (int i = 0; i < n; ++i) {
                OverlapRange o;
                const FastaRecord& cur = seqs[rnd() % seqs.size()];
                const FastaRecord& ext = seqs[rnd() % seqs.size()];
                o.curId = cur.id; o.curLen = (int32_t)cur.sequence.length(); o.curBegin = (int32_t)(rnd() % 100); o.curEnd = o.curLen - (int32_t)(rnd() % 50);
                o.extId = ext.id; o.extLen = (int32_t)ext.sequence.length(); o.extBegin = (int32_t)(rnd() % 100); o.extEnd = o.extLen - (int32_t)(rnd() % 50);
                o.score = (int32_t)(rnd() % 100000);
                o.seqDivergence = (float)(rnd() % 10000) / 65536.0f;
                o.dump(ss, c, c);
                ss << "\n";
            }
            const std::string text = ss.str();
            printf("DUMP\n%s", text.c_str());
            std::stringstream in(text), again;
            for (int i = 0; i < n; ++i) {
                OverlapRange o;
                o.load(in, c, c);
                printf("LOADED %u %d %d %d %u %d %d %d %d %.9g\n", idNum(o.curId) - firstId, o.curBegin, o.curEnd, o.curLen, idNum(o.extId) - firstId,
                       o.extBegin, o.extEnd, o.extLen, o.score, (double)o.seqDivergence);
                o.dump(again, c, c);
                again << "\n";
            }
            printf("REDUMP %s\n", again.str() == text ? "identical" : "DIFFERENT");
        }
    }
    return 0;
}
