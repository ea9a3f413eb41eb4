// TEST INFRASTRUCTURE — not product code.
//
// One driver, two builds:
//   (1) against the UNMODIFIED reference headers/sources in /root/reference/src
//       (-> oracle/_ref/flye_ref_harness, recipe: oracle/Makefile), and
//   (2) against the B200 host mirror in flye_b200/host (-DFLYE_B200, linked to
//       libflye_b200.so -> build/flye_b200_harness).
// It uses only the public API of src/sequence/ exactly as assemble_main does
// (reference: src/assemble/main_assemble.cpp:158-242) and dumps everything the
// hot path computes, so the two builds can be compared byte for byte:
//   <out>.hist   k-mer frequency histogram            (vertex_index.cpp:567-576)
//   <out>.index  per distinct query k-mer: repetitive flag, list of (seqId,pos)
//                as seen through iterKmerPos           (vertex_index.h:220-246)
//   <out>.ovlp   ordered per-read getSeqOverlaps vectors, all fields, divergence
//                as raw float bits                     (overlap.cpp:99-508)
//   stdout       one JSON line with phase timings.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cinttypes>
#include <string>
#include <vector>
#include <set>
#include <mutex>
#include <functional>
#include <fstream>
#include <algorithm>

#include "sequence/sequence_container.h"
#include "sequence/vertex_index.h"
#include "sequence/overlap.h"
#include "common/config.h"
#include "common/parallel.h"

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Args {
    std::string reads, cfg, out, queryReads;
    int k = -1, threads = 1, minOverlap = 1000, maxOverlaps = 0, minReadLen = -1;
    bool forceLocal = false, dumpIndex = false, bothStrands = false, noEstimate = false;
    bool findAll = false, noOverlaps = false, keepAln = false, allExt = false, partitionBad = false;
    bool lazy = false, perRead = false;
    long maxQueries = -1;
    long chunk = 0;          // --chunk N: answer and dump N queries at a time (bounds the memory of a 10^8-overlap run)
};

static void usage() {
    fprintf(stderr,
        "harness --reads F --cfg C --out PREFIX [--k K] [--threads T] [--min-overlap M]\n"
        "        [--max-overlaps N] [--force-local] [--dump-index] [--both-strands]\n"
        "        [--no-estimate] [--find-all] [--no-overlaps] [--max-queries N] [--all-ext] [--partition-bad] [--lazy] [--per-read] [--chunk N]\n");
}

static uint32_t idNum(FastaRecord::Id id) {
    // Id has no public accessor for the raw number; signedId() is invertible.
    int s = id.signedId();
    return s > 0 ? (uint32_t)(s - 1) * 2 : (uint32_t)(-s - 1) * 2 + 1;
}

static void dumpOverlap(FILE* f, const OverlapRange& o) {
    uint32_t bits; float d = o.seqDivergence; memcpy(&bits, &d, 4);
    fprintf(f, "%u %d %d %d %u %d %d %d %d %08x\n", idNum(o.curId), o.curBegin, o.curEnd, o.curLen,
            idNum(o.extId), o.extBegin, o.extEnd, o.extLen, o.score, bits);
}

int main(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        auto nxt = [&]() { if (i + 1 >= argc) { usage(); exit(1); } return std::string(argv[++i]); };
        if (s == "--reads") a.reads = nxt(); else if (s == "--cfg") a.cfg = nxt();
        else if (s == "--out") a.out = nxt(); else if (s == "--k") a.k = atoi(nxt().c_str());
        else if (s == "--query-reads") a.queryReads = nxt();
        else if (s == "--threads") a.threads = atoi(nxt().c_str());
        else if (s == "--min-overlap") a.minOverlap = atoi(nxt().c_str());
        else if (s == "--min-read-len") a.minReadLen = atoi(nxt().c_str());
        else if (s == "--max-overlaps") a.maxOverlaps = atoi(nxt().c_str());
        else if (s == "--max-queries") a.maxQueries = atol(nxt().c_str());
        else if (s == "--chunk") a.chunk = atol(nxt().c_str());
        else if (s == "--force-local") a.forceLocal = true;
        else if (s == "--dump-index") a.dumpIndex = true;
        else if (s == "--both-strands") a.bothStrands = true;
        else if (s == "--no-estimate") a.noEstimate = true;
        else if (s == "--find-all") a.findAll = true;
        else if (s == "--no-overlaps") a.noOverlaps = true;
        else if (s == "--keep-aln") a.keepAln = true;
        else if (s == "--all-ext") a.allExt = true;
        else if (s == "--partition-bad") a.partitionBad = true;   // OverlapDetector(..., partitionBadMappings = true): the repeat stage's setting
        else if (s == "--lazy") a.lazy = true;          // lazySeqOverlaps(id) from --threads worker threads (how extender.cpp calls it)
        else if (s == "--per-read") a.perRead = true;   // one quickSeqOverlaps call per read from --threads threads, also on the mirror
        else { usage(); return 1; }
    }
    if (a.reads.empty() || a.cfg.empty() || a.out.empty()) { usage(); return 1; }

    Config::load(a.cfg);
    int kmerSize = a.k > 0 ? a.k : (int)Config::get("kmer_size");
    Parameters::get().numThreads = a.threads;
    Parameters::get().kmerSize = kmerSize;
    Parameters::get().minimumOverlap = a.minOverlap;
    Parameters::get().unevenCoverage = false;

    double t0 = now();
    SequenceContainer reads;
    reads.loadFromFile(a.reads, a.minReadLen >= 0 ? a.minReadLen : a.minOverlap);
    reads.buildPositionIndex();
    double tLoad = now() - t0;

    VertexIndex index(reads, (int)Config::get("assemble_kmer_sample"));
    index.outputProgress(false);
    const int MIN_FREQ = 2;
    const float selectRate = Config::get("meta_read_top_kmer_rate");
    const int tandemFreq = Config::get("meta_read_filter_kmer_freq");
    bool useMinimizers = Config::get("use_minimizers");
    double tCount = 0, tIndex = 0;
    if (useMinimizers) {
        t0 = now();
        index.buildIndexMinimizers(1, (int)Config::get("minimizer_window"));
        tIndex = now() - t0;
    } else {
        t0 = now();
        index.countKmers();
        tCount = now() - t0;
        {   // histogram must be read before buildIndex clears the counter? (it survives; dump now anyway)
            FILE* f = fopen((a.out + ".hist").c_str(), "w");
            for (const auto& kv : index.getKmerHist()) fprintf(f, "%zu %zu\n", (size_t)kv.first, (size_t)kv.second);
            fclose(f);
        }
        t0 = now();
        index.buildIndexUnevenCoverage(MIN_FREQ, selectRate, tandemFreq);
        tIndex = now() - t0;
    }

    if (a.dumpIndex) {
        // every distinct canonical k-mer that any read position (either strand's enumeration) can query
        std::set<uint64_t> keys;
        for (const auto& rec : reads.iterSeqs())
            for (auto kp : IterKmers(rec.sequence)) { kp.kmer.standardForm(); keys.insert(kp.kmer.numRepr()); }
        FILE* f = fopen((a.out + ".index").c_str(), "w");
        uint32_t sr; float srf = index.getSampleRate(); memcpy(&sr, &srf, 4);
        fprintf(f, "sampleRate %08x\n", sr);
        for (uint64_t key : keys) {
            Kmer km(key);
            bool rep = index.isRepetitive(km);
            size_t freq = index.kmerFreq(km);
            if (!rep && !freq) continue;
            fprintf(f, "%" PRIx64 " %d %zu", key, (int)rep, freq);
            // iterKmerPos on a k-mer that is not in the table throws in the reference (cuckoo find); getSeqOverlaps
            // itself only calls it after kmerFreq() != 0 (overlap.cpp:183-186)
            if (freq) for (const auto& rp : index.iterKmerPos(km)) fprintf(f, " %u:%d", idNum(rp.readId), rp.position);
            fputc('\n', f);
        }
        fclose(f);
    }

    double tEstimate = 0, tOverlaps = 0; size_t nQueries = 0, nOverlaps = 0;
    if (!a.noOverlaps) {
        OverlapDetector detector(reads, index, (int)Config::get("maximum_jump"),
                                 Parameters::get().minimumOverlap, (int)Config::get("maximum_overhang"),
                                 /*keepAlignment*/ a.keepAln, /*onlyMaxExt*/ !a.allExt, /*maxDivergence*/ 1.0f,
                                 (bool)Config::get("reads_base_alignment"), /*partitionBadMappings*/ a.partitionBad,
                                 (bool)Config::get("hpc_scoring_on"));
        // queries from a SECOND container against the index of the first, as ReadAligner::alignReads does
        // (read_aligner.cpp:178-217): ids keep running through the process-global counter
        SequenceContainer queryReads;
        if (!a.queryReads.empty()) queryReads.loadFromFile(a.queryReads, a.minReadLen >= 0 ? a.minReadLen : a.minOverlap);
        const SequenceContainer& qc = a.queryReads.empty() ? reads : queryReads;
        OverlapContainer container(detector, qc);
        if (!a.noEstimate) {
            t0 = now();
            container.estimateOverlaperParameters();
            container.setDivergenceThreshold((float)Config::get("assemble_ovlp_divergence"),
                                             (bool)Config::get("assemble_divergence_relative"));
            tEstimate = now() - t0;
        }
        std::vector<FastaRecord::Id> queries;
        for (const auto& rec : qc.iterSeqs())
            if (rec.id.strand() || a.bothStrands) queries.push_back(rec.id);
        if (a.maxQueries >= 0 && (size_t)a.maxQueries < queries.size()) queries.resize(a.maxQueries);
        nQueries = queries.size();

        FILE* f = fopen((a.out + ".ovlp").c_str(), "w");
        if (a.findAll) {
            t0 = now();
            container.findAllOverlaps();
            tOverlaps = now() - t0;
            for (auto id : queries) {
                auto ov = container.lazySeqOverlaps(id);   // copy: compare as multiset
                std::vector<std::string> lines;
                for (const auto& o : ov) {
                    char buf[256]; uint32_t bits; float d = o.seqDivergence; memcpy(&bits, &d, 4);
                    snprintf(buf, sizeof buf, "%u %d %d %d %u %d %d %d %d %08x", idNum(o.curId), o.curBegin, o.curEnd,
                             o.curLen, idNum(o.extId), o.extBegin, o.extEnd, o.extLen, o.score, bits);
                    lines.push_back(buf);
                }
                std::sort(lines.begin(), lines.end());
                fprintf(f, "# %u %zu\n", idNum(id), lines.size());
                for (auto& l : lines) fprintf(f, "%s\n", l.c_str());
                nOverlaps += lines.size();
            }
        } else {
            const size_t step = a.chunk > 0 ? (size_t)a.chunk : std::max<size_t>(queries.size(), 1);
            for (size_t base = 0; base < queries.size(); base += step) {
            const size_t cnt = std::min(step, queries.size() - base);
            std::vector<std::vector<OverlapRange>> results(cnt);
            t0 = now();
            std::vector<size_t> order(cnt);
            for (size_t i = 0; i < order.size(); ++i) order[i] = i;
            std::function<void(const size_t&)> work = [&](const size_t& i) {
                if (a.lazy) results[i] = container.lazySeqOverlaps(queries[base + i]);   // thread safe, cached (overlap.h:397-411)
                else results[i] = container.quickSeqOverlaps(queries[base + i], a.maxOverlaps, a.forceLocal);
            };
#ifdef FLYE_B200
            // the mirror's batch entry (same results as N quickSeqOverlaps calls, one device pass) unless the per-read
            // calling pattern itself is what is being tested
            if (!a.lazy && !a.perRead)
                results = container.quickSeqOverlapsBatch(std::vector<FastaRecord::Id>(queries.begin() + base, queries.begin() + base + cnt),
                                                          a.maxOverlaps, a.forceLocal);
            else processInParallel(order, work, Parameters::get().numThreads, false);
#else
            processInParallel(order, work, Parameters::get().numThreads, false);
#endif
            tOverlaps += now() - t0;
            for (size_t i = 0; i < cnt; ++i) {
                fprintf(f, "# %u %zu\n", idNum(queries[base + i]), results[i].size());
                for (const auto& o : results[i]) {
                    dumpOverlap(f, o);
                    if (a.keepAln && o.kmerMatches) {
                        fprintf(f, "  aln %zu", o.kmerMatches->size());
                        for (auto& p : *o.kmerMatches) fprintf(f, " %d,%d", p.first, p.second);
                        fputc('\n', f);
                    }
                }
                nOverlaps += results[i].size();
            }
            }
        }
        fclose(f);
    }

    size_t nFwd = 0, nBases = 0;
    for (const auto& rec : reads.iterSeqs()) if (rec.id.strand()) { ++nFwd; nBases += rec.sequence.length(); }
    printf("{\"reads\": %zu, \"bases\": %zu, \"k\": %d, \"threads\": %d, \"minimizers\": %d, "
           "\"t_load\": %.4f, \"t_count\": %.4f, \"t_index\": %.4f, \"t_estimate\": %.4f, \"t_overlaps\": %.4f, "
           "\"queries\": %zu, \"overlaps\": %zu}\n",
           nFwd, nBases, kmerSize, a.threads, (int)useMinimizers, tLoad, tCount, tIndex, tEstimate, tOverlaps,
           nQueries, nOverlaps);
    return 0;
}
