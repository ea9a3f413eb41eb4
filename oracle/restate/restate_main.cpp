// TEST INFRASTRUCTURE — command-line driver for the CPU restatement (see restate.h).
// Same arguments and byte-identical dump formats as oracle/harness.cpp, so that
//   diff <(flye_ref_harness ...) <(flye_restate ...)
// is the pinning test.
#include "restate.h"

#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <atomic>

using namespace restate;

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// common/config.h:36-72: "key = float" lines, '#' comments, "%include file" relative to the including file
static void loadCfg(const std::string& path, std::map<std::string, float>& kv) {
    std::ifstream in(path);
    if (!in) { fprintf(stderr, "Can't open config file: %s\n", path.c_str()); exit(1); }
    std::string dir; size_t slash = path.find_last_of("/\\");
    if (slash != std::string::npos) dir = path.substr(0, slash + 1);
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        if (line.compare(0, 8, "%include") == 0) {
            std::istringstream ss(line); std::string a, b; ss >> a >> b; loadCfg(dir + b, kv); continue;
        }
        size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        auto trim = [](std::string s) {
            size_t l = s.find_first_not_of(" \t\r"), r = s.find_last_not_of(" \t\r");
            return l == std::string::npos ? std::string() : s.substr(l, r - l + 1);
        };
        kv[trim(line.substr(0, eq))] = (float)atof(trim(line.substr(eq + 1)).c_str());
    }
}

static void dumpOverlap(FILE* f, const Overlap& o) {
    uint32_t bits; memcpy(&bits, &o.seqDivergence, 4);
    fprintf(f, "%u %d %d %d %u %d %d %d %d %08x\n", o.curId, o.curBegin, o.curEnd, o.curLen, o.extId, o.extBegin,
            o.extEnd, o.extLen, o.score, bits);
}

int main(int argc, char** argv) {
    std::string readsPath, cfgPath, out;
    int k = -1, threads = 1, minOverlap = 1000, maxOverlaps = 0, minReadLen = -1; long maxQueries = -1;
    bool forceLocal = false, dumpIndex = false, bothStrands = false, noEstimate = false, noOverlaps = false;
    bool keepAln = false, allExt = false, partitionBad = false;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        auto nxt = [&]() { if (i + 1 >= argc) exit(1); return std::string(argv[++i]); };
        if (s == "--reads") readsPath = nxt(); else if (s == "--cfg") cfgPath = nxt();
        else if (s == "--out") out = nxt(); else if (s == "--k") k = atoi(nxt().c_str());
        else if (s == "--threads") threads = atoi(nxt().c_str());
        else if (s == "--min-overlap") minOverlap = atoi(nxt().c_str());
        else if (s == "--min-read-len") minReadLen = atoi(nxt().c_str());
        else if (s == "--max-overlaps") maxOverlaps = atoi(nxt().c_str());
        else if (s == "--max-queries") maxQueries = atol(nxt().c_str());
        else if (s == "--force-local") forceLocal = true;
        else if (s == "--dump-index") dumpIndex = true;
        else if (s == "--both-strands") bothStrands = true;
        else if (s == "--no-estimate") noEstimate = true;
        else if (s == "--no-overlaps") noOverlaps = true;
        else if (s == "--keep-aln") keepAln = true;
        else if (s == "--all-ext") allExt = true;
        else if (s == "--partition-bad") partitionBad = true;
        else { fprintf(stderr, "unknown option %s\n", s.c_str()); return 1; }
    }
    std::map<std::string, float> cfg;
    loadCfg(cfgPath, cfg);
    Params p;
    p.k = k > 0 ? k : (int)cfg.at("kmer_size");
    p.minOverlap = minOverlap;
    p.maxJump = (int)cfg.at("maximum_jump");
    p.maxOverhang = (int)cfg.at("maximum_overhang");
    p.keepAlignment = keepAln; p.onlyMaxExt = !allExt; p.partitionBadMappings = partitionBad;
    p.nuclAlignment = (bool)cfg.at("reads_base_alignment");
    p.useHpc = (bool)cfg.at("hpc_scoring_on");
    p.maxDivergence = 1.0f;
    p.useMinimizers = (bool)cfg.at("use_minimizers");
    if (p.useMinimizers) p.minimizerWindow = (int)cfg.at("minimizer_window");
    p.selectRate = cfg.at("meta_read_top_kmer_rate");
    p.tandemFreq = (int)cfg.at("meta_read_filter_kmer_freq");
    p.repeatKmerRate = cfg.at("repeat_kmer_rate");
    p.sampleRate = (int)cfg.at("assemble_kmer_sample");
    p.ovlpDivergence = cfg.at("assemble_ovlp_divergence");
    p.divergenceRelative = (bool)cfg.at("assemble_divergence_relative");

    double t0 = now();
    Reads reads = loadFasta(readsPath, minReadLen >= 0 ? minReadLen : minOverlap);
    double tLoad = now() - t0;

    Index index;
    double tCount = 0, tIndex = 0;
    t0 = now();
    if (p.useMinimizers) buildIndexMinimizers(reads, p, index);
    else {
        buildIndexSolid(reads, p, index);
        FILE* f = fopen((out + ".hist").c_str(), "w");
        for (const auto& kv : index.hist) fprintf(f, "%zu %zu\n", (size_t)kv.first, (size_t)kv.second);
        fclose(f);
    }
    tIndex = now() - t0;

    if (dumpIndex) {
        std::set<uint64_t> keys;
        for (uint32_t id = 0; id < reads.numIds(); ++id)
            for (uint64_t km : iterKmers(reads, id, p.k)) { kmerCanonical(km, p.k); keys.insert(km); }
        FILE* f = fopen((out + ".index").c_str(), "w");
        uint32_t sr; memcpy(&sr, &index.sampleRate, 4);
        fprintf(f, "sampleRate %08x\n", sr);
        for (uint64_t key : keys) {
            bool rep = index.repetitive.count(key);
            auto it = index.lists.find(key);
            size_t freq = it == index.lists.end() ? 0 : it->second.size();
            if (!rep && !freq) continue;
            fprintf(f, "%" PRIx64 " %d %zu", key, (int)rep, freq);
            if (it != index.lists.end())
                for (uint64_t g : it->second) { SeqPos sp = seqPosition(reads, g); fprintf(f, " %u:%d", sp.id, sp.pos); }
            fputc('\n', f);
        }
        fclose(f);
    }

    double tEstimate = 0, tOverlaps = 0; size_t nQueries = 0, nOverlaps = 0;
    if (!noOverlaps) {
        float maxDiv = p.maxDivergence;
        if (!noEstimate) { t0 = now(); maxDiv = estimateMaxDivergence(reads, index, p); tEstimate = now() - t0; }
        std::vector<uint32_t> queries;
        for (uint32_t id = 0; id < reads.numIds(); ++id) if (!(id & 1) || bothStrands) queries.push_back(id);
        if (maxQueries >= 0 && (size_t)maxQueries < queries.size()) queries.resize(maxQueries);
        nQueries = queries.size();
        std::vector<std::vector<Overlap>> results(queries.size());
        t0 = now();
        std::atomic<size_t> next(0);
        auto worker = [&]() {
            for (;;) { size_t i = next++; if (i >= queries.size()) return;
                results[i] = getSeqOverlaps(reads, index, p, queries[i], forceLocal, maxOverlaps, maxDiv); }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < std::max(1, threads); ++t) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
        tOverlaps = now() - t0;
        FILE* f = fopen((out + ".ovlp").c_str(), "w");
        for (size_t i = 0; i < queries.size(); ++i) {
            fprintf(f, "# %u %zu\n", queries[i], results[i].size());
            for (const auto& o : results[i]) {
                dumpOverlap(f, o);
                if (keepAln && !o.kmerMatches.empty()) {
                    fprintf(f, "  aln %zu", o.kmerMatches.size());
                    for (auto& pr : o.kmerMatches) fprintf(f, " %d,%d", pr.first, pr.second);
                    fputc('\n', f);
                }
            }
            nOverlaps += results[i].size();
        }
        fclose(f);
    }
    size_t nBases = 0; for (const auto& s : reads.fwd) nBases += s.size();
    printf("{\"reads\": %zu, \"bases\": %zu, \"k\": %d, \"threads\": %d, \"minimizers\": %d, "
           "\"t_load\": %.4f, \"t_count\": %.4f, \"t_index\": %.4f, \"t_estimate\": %.4f, \"t_overlaps\": %.4f, "
           "\"queries\": %zu, \"overlaps\": %zu, \"dp_cells\": %llu}\n",
           reads.fwd.size(), nBases, p.k, threads, (int)p.useMinimizers, tLoad, tCount, tIndex, tEstimate, tOverlaps,
           nQueries, nOverlaps, (unsigned long long)restate::dpCellsVisited.load());
    return 0;
}
