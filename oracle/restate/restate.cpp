// TEST INFRASTRUCTURE — see restate.h.  Sequential CPU restatement of the reference hot path.
#include "restate.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <limits>
#include <numeric>
#include <stdexcept>

namespace restate {

std::atomic<unsigned long long> dpCellsVisited{0};


// ------------------------------------------------------------------------------------------------
// reads
// ------------------------------------------------------------------------------------------------
static int baseCode(char c) {
    switch (c) {   // sequence.h:166-173
        case 'A': case 'a': return 0; case 'C': case 'c': return 1;
        case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return -1;
    }
}

void Reads::buildOffsets() {   // sequence_container.cpp:359-369: offsets over ALL ids in id order
    offset.assign(numIds() + 1, 0);
    uint64_t off = 0;
    for (size_t id = 0; id < numIds(); ++id) { offset[id] = off; off += fwd[id >> 1].size(); }
    offset[numIds()] = off;
}

Reads loadFasta(const std::string& path, int minReadLength) {
    // sequence_container.cpp:81-107,145-227: keep reads strictly longer than minReadLength.
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    Reads r; std::string line; std::vector<uint8_t> cur; bool have = false;
    auto flush = [&]() {
        if (have && (long)cur.size() > (long)minReadLength) r.fwd.push_back(cur);
        cur.clear();
    };
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') { flush(); have = true; continue; }
        for (char c : line) {
            int b = baseCode(c);
            if (b < 0) throw std::runtime_error("non-ACGT letter: inputs must be ACGT only (SURVEY 9.2)");
            cur.push_back((uint8_t)b);
        }
    }
    flush();
    r.buildOffsets();
    return r;
}

SeqPos seqPosition(const Reads& r, uint64_t g) {   // sequence_container.h:220-235 (hint table = a search)
    size_t id = std::upper_bound(r.offset.begin(), r.offset.end(), g) - r.offset.begin() - 1;
    return {(uint32_t)id, (int32_t)(g - r.offset[id])};
}

// ------------------------------------------------------------------------------------------------
// k-mers
// ------------------------------------------------------------------------------------------------
uint64_t kmerRevComp(uint64_t kmer, int k) {   // kmer.h:39-52
    uint64_t out = 0;
    for (int i = 0; i < k; ++i) { out = (out << 2) | (~kmer & 3); kmer >>= 2; }
    return out;
}

bool kmerCanonical(uint64_t& kmer, int k) {   // kmer.h:54-63
    uint64_t rc = kmerRevComp(kmer, k);
    if (rc < kmer) { kmer = rc; return true; }
    return false;
}

uint64_t kmerHash(uint64_t x) {   // kmer.h:91-98
    uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

std::vector<uint64_t> iterKmers(const Reads& r, uint32_t id, int k) {
    // kmer.h:131-204: positions 0 .. L-k-1 — the last k-mer is never visited; empty if L <= k.
    std::vector<uint64_t> out;
    long L = r.len(id);
    if (L <= k) return out;
    const uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint64_t km = 0;
    for (int i = 0; i < k; ++i) km = (km << 2) | r.at(id, i);
    out.reserve(L - k);
    for (long p = 0; p < L - k; ++p) {
        out.push_back(km);
        km = ((km << 2) | r.at(id, p + k)) & mask;   // appendRight, kmer.h:65-73
    }
    return out;
}

std::vector<int32_t> yieldMinimizers(const Reads& r, uint32_t id, int k, int window) {   // kmer.h:206-262
    std::vector<uint64_t> kmers = iterKmers(r, id, k);
    std::vector<int32_t> out;
    if (window == 1) { for (size_t p = 0; p < kmers.size(); ++p) out.push_back((int32_t)p); return out; }
    struct QE { int32_t pos; uint64_t hash; };
    std::deque<QE> q;
    for (int32_t p = 0; p < (int32_t)kmers.size(); ++p) {
        uint64_t c = kmers[p]; kmerCanonical(c, k);
        uint64_t h = kmerHash(c);
        while (!q.empty() && q.back().hash > h) q.pop_back();
        q.push_back({p, h});
        if (q.front().pos <= p - window) {
            while (q.front().pos <= p - window) q.pop_front();
            while (q.size() >= 2 && q[0].hash == q[1].hash) q.pop_front();
        }
        if (out.empty() || out.back() != q.front().pos) out.push_back(q.front().pos);
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// index
// ------------------------------------------------------------------------------------------------
namespace {
struct Selected { uint64_t kmerFwd; int32_t pos; size_t freq; };

// vertex_index.cpp:316-358
std::vector<Selected> yieldFrequentKmers(const Reads& r, uint32_t id, int k, float selectRate, int tandemFreq,
                                         const std::unordered_map<uint64_t, uint32_t>& counts) {
    std::unordered_map<uint64_t, size_t> localFreq;
    std::vector<Selected> top;
    std::vector<uint64_t> kmers = iterKmers(r, id, k);
    for (size_t p = 0; p < kmers.size(); ++p) {
        uint64_t c = kmers[p]; kmerCanonical(c, k);
        auto it = counts.find(c);
        size_t freq = it == counts.end() ? 0 : it->second;
        ++localFreq[c];
        top.push_back({kmers[p], (int32_t)p, freq});
    }
    if (top.empty()) return {};
    // the reference sorts by freq descending (unstable); only the kept SET matters
    std::vector<size_t> f(top.size());
    for (size_t i = 0; i < top.size(); ++i) f[i] = top[i].freq;
    std::sort(f.begin(), f.end(), std::greater<size_t>());
    const size_t maxKmers = selectRate * top.size();   // float product, then truncation (:339)
    const size_t minFreq = f[maxKmers];
    std::vector<Selected> kept;
    for (const auto& s : top) {
        if (s.freq < minFreq) continue;
        if (tandemFreq > 0) {
            uint64_t c = s.kmerFwd; kmerCanonical(c, k);
            if (localFreq[c] > (size_t)tandemFreq) continue;
        }
        kept.push_back(s);
    }
    return kept;
}

// vertex_index.cpp:173-212 on a capacity table; returns keys that become repetitive
void filterFrequent(std::unordered_map<uint64_t, uint32_t>& capacity, int minCoverage, float rate, Index& out) {
    size_t total = 0, uniq = 0;
    for (const auto& kv : capacity)
        if (kv.second >= (size_t)minCoverage) { total += kv.second; uniq += 1; }
    float meanFrequency = (float)total / (uniq + 1);
    out.repetitiveFrequency = rate * meanFrequency;
    for (const auto& kv : capacity)
        if (kv.second > out.repetitiveFrequency) out.repetitive.insert(kv.first);
    for (uint64_t key : out.repetitive) capacity.erase(key);
}
}  // namespace

void buildIndexSolid(const Reads& r, const Params& p, Index& out) {
    const int k = p.k;
    if (k > 17) throw std::runtime_error("Can't use flat counter for k-mer size > 17");   // :504-507
    // KmerCounter::count (:499-590): exact count of canonical k-mers over forward reads
    std::unordered_map<uint64_t, uint32_t> counts;
    for (size_t i = 0; i < r.fwd.size(); ++i)
        for (uint64_t km : iterKmers(r, (uint32_t)(2 * i), k)) { kmerCanonical(km, k); ++counts[km]; }
    out.hist.clear();
    for (const auto& kv : counts) out.hist[kv.second] += 1;   // :567-576

    const int minFreq = 2;   // MIN_FREQ, main_assemble.cpp:207
    // pass 1 (:41-59): capacity of every selected k-mer with global freq >= minFreq
    std::unordered_map<uint64_t, uint32_t> capacity;
    for (size_t i = 0; i < r.fwd.size(); ++i)
        for (const auto& s : yieldFrequentKmers(r, (uint32_t)(2 * i), k, p.selectRate, p.tandemFreq, counts)) {
            if (s.freq < (size_t)minFreq) continue;
            uint64_t c = s.kmerFwd; kmerCanonical(c, k);
            ++capacity[c];
        }
    filterFrequent(capacity, minFreq, p.repeatKmerRate, out);   // :61
    out.lists.clear();
    for (const auto& kv : capacity) out.lists[kv.first];        // keys stay even if they end up empty
    // pass 2 (:65-104)
    for (size_t i = 0; i < r.fwd.size(); ++i) {
        uint32_t id = (uint32_t)(2 * i);
        for (const auto& s : yieldFrequentKmers(r, id, k, p.selectRate, p.tandemFreq, counts)) {
            if (s.freq < (size_t)minFreq || s.freq > out.repetitiveFrequency) continue;
            uint64_t c = s.kmerFwd; uint32_t target = id; int32_t pos = s.pos;
            if (kmerCanonical(c, k)) { pos = r.len(id) - pos - k; target = id ^ 1; }
            auto it = out.lists.find(c);
            if (it == out.lists.end()) continue;
            if (it->second.size() == capacity[c]) continue;   // "Index size mismatch" guard (:91-95)
            it->second.push_back(r.offset[target] + pos);
        }
    }
    for (auto& kv : out.lists) std::sort(kv.second.begin(), kv.second.end());   // :109-114
    out.sampleRate = (float)p.sampleRate;   // VertexIndex ctor argument (main_assemble.cpp:195-197)
}

void buildIndexMinimizers(const Reads& r, const Params& p, Index& out) {   // vertex_index.cpp:389-483
    const int k = p.k;
    size_t totalLen = 0;
    for (const auto& s : r.fwd) totalLen += s.size();
    std::unordered_map<uint64_t, uint32_t> capacity;
    std::vector<std::vector<int32_t>> mins(r.fwd.size());
    for (size_t i = 0; i < r.fwd.size(); ++i) {
        mins[i] = yieldMinimizers(r, (uint32_t)(2 * i), k, p.minimizerWindow);
        std::vector<uint64_t> kmers = iterKmers(r, (uint32_t)(2 * i), k);
        for (int32_t pos : mins[i]) { uint64_t c = kmers[pos]; kmerCanonical(c, k); ++capacity[c]; }
    }
    filterFrequent(capacity, /*minCoverage*/ 1, p.repeatKmerRate, out);
    out.lists.clear();
    for (const auto& kv : capacity) out.lists[kv.first];
    for (size_t i = 0; i < r.fwd.size(); ++i) {
        uint32_t id = (uint32_t)(2 * i);
        std::vector<uint64_t> kmers = iterKmers(r, id, k);
        for (int32_t mpos : mins[i]) {
            uint64_t c = kmers[mpos]; uint32_t target = id; int32_t pos = mpos;
            if (kmerCanonical(c, k)) { pos = r.len(id) - pos - k; target = id ^ 1; }
            if (out.repetitive.count(c)) continue;   // :442
            auto it = out.lists.find(c);
            if (it == out.lists.end()) continue;
            if (it->second.size() == capacity[c]) continue;
            it->second.push_back(r.offset[target] + pos);
        }
    }
    size_t totalEntries = 0;
    for (auto& kv : out.lists) { std::sort(kv.second.begin(), kv.second.end()); totalEntries += kv.second.size(); }
    out.sampleRate = (float)totalLen / totalEntries;   // :480-482
    out.hist.clear();
}

// ------------------------------------------------------------------------------------------------
// base-level divergence
// ------------------------------------------------------------------------------------------------
std::vector<uint8_t> hpcSubstr(const Reads& r, uint32_t id, int32_t start, int32_t length, bool compress) {
    // alignment.cpp:52-70
    std::vector<uint8_t> out;
    for (int32_t i = 0; i < length; ++i) {
        uint8_t c = r.at(id, (size_t)start + i);
        if (!compress || i == 0 || out.back() != c) out.push_back(c);
    }
    return out;
}

int editDistanceNW(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {
    // Exact global (NW) unit-cost edit distance — what edlibAlign(k=-1, EDLIB_MODE_NW, EDLIB_TASK_DISTANCE)
    // returns (alignment.cpp:231-236; edlib.cpp:141-212 iterates k = 64,128,... until the distance fits).
    // Restated as Ukkonen's banded DP with the same doubling idea; the value is unique, so any exact
    // algorithm matches.
    const int n = (int)a.size(), m = (int)b.size();
    if (n == 0) return m;
    if (m == 0) return n;
    const int INF = std::numeric_limits<int>::max() / 2;
    for (int band = 64;; band *= 2) {
        if (std::abs(n - m) > band) continue;
        // row i holds columns j in [i-band, i+band]
        std::vector<int> prev(2 * band + 3, INF), cur(2 * band + 3, INF);
        auto idx = [&](int i, int j) { return j - i + band + 1; };
        for (int j = 0; j <= std::min(m, band); ++j) prev[idx(0, j)] = j;
        for (int i = 1; i <= n; ++i) {
            std::fill(cur.begin(), cur.end(), INF);
            int lo = std::max(0, i - band), hi = std::min(m, i + band);
            for (int j = lo; j <= hi; ++j) {
                int best = INF;
                if (j == 0) best = i;
                else {
                    int d = prev[idx(i - 1, j - 1)];
                    if (d < INF) best = d + (a[i - 1] != b[j - 1]);
                    int l = cur[idx(i, j - 1)];
                    if (l < INF) best = std::min(best, l + 1);
                }
                if (j - (i - 1) <= band) { int u = prev[idx(i - 1, j)]; if (u < INF) best = std::min(best, u + 1); }
                cur[idx(i, j)] = best;
            }
            std::swap(prev, cur);
        }
        int d = prev[idx(n, m)];
        if (d <= band) return d;
    }
}

static float alignmentErrEdlib(const Reads& r, const Overlap& o, bool useHpc) {   // alignment.cpp:218-247
    auto trg = hpcSubstr(r, o.curId, o.curBegin, o.curEnd - o.curBegin, useHpc);
    auto qry = hpcSubstr(r, o.extId, o.extBegin, o.extEnd - o.extBegin, useHpc);
    int ed = editDistanceNW(qry, trg);
    return (float)ed / std::max(qry.size(), trg.size());
}

// ------------------------------------------------------------------------------------------------
// overlaps
// ------------------------------------------------------------------------------------------------
namespace {
struct KmerMatch { int32_t curPos; int32_t extPos; uint32_t extId; };   // overlap.cpp:74-82

int32_t curRange(const Overlap& o) { return o.curEnd - o.curBegin; }
int32_t extRange(const Overlap& o) { return o.extEnd - o.extBegin; }

bool overlapTest(const Overlap& o, const Params& p, bool forceLocal) {   // overlap.cpp:29-69
    if (curRange(o) < p.minOverlap || extRange(o) < p.minOverlap) return false;
    static const float OVLP_DIV = 0.5;
    float lengthDiff = abs(curRange(o) - extRange(o));
    if (lengthDiff > OVLP_DIV * std::min(curRange(o), extRange(o))) return false;
    if (o.curId == o.extId) {
        int32_t intersect = std::min(o.curEnd, o.extEnd) - std::max(o.curBegin, o.extBegin);
        if (intersect > curRange(o) / 2) return false;
    }
    if (o.curId == (o.extId ^ 1)) {
        int32_t intersect = std::min(o.curEnd, o.extLen - o.extBegin) - std::max(o.curBegin, o.extLen - o.extEnd);
        if (intersect > curRange(o) / 2) return false;
    }
    const bool checkOverhang = p.maxOverhang > 0;   // overlap.h:324
    int32_t lrOverhang = std::max(std::min(o.curBegin, o.extBegin),
                                  std::min(o.curLen - o.curEnd, o.extLen - o.extEnd));   // overlap.h:195-199
    if (!forceLocal && checkOverhang && lrOverhang > p.maxOverhang) return false;
    return true;
}

bool containedBy(const Overlap& a, const Overlap& b) {   // overlap.h:207-213
    if (a.curId != b.curId || a.extId != b.extId) return false;
    return b.curBegin <= a.curBegin && a.curEnd <= b.curEnd && b.extBegin <= a.extBegin && a.extEnd <= b.extEnd;
}
}  // namespace

std::vector<Overlap> getSeqOverlaps(const Reads& r, const Index& idx, const Params& p, uint32_t queryId,
                                    bool forceLocal, int maxOverlaps, float maxDivergence) {
    // overlap.cpp:99-508
    const int kmerSize = p.k;
    const float minKmerSruvivalRate = 0.01;
    const float LG_GAP = 2;
    const float SM_GAP = 0.5;
    const bool checkOverhang = p.maxOverhang > 0;
    int32_t curLen = r.len(queryId);
    std::vector<int32_t> curFilteredPos;
    std::vector<KmerMatch> vecMatches;

    // hit gathering (:176-196; vertex_index.h:158-174,220-246)
    std::vector<uint64_t> kmers = iterKmers(r, queryId, kmerSize);
    for (int32_t pos = 0; pos < (int32_t)kmers.size(); ++pos) {
        uint64_t c = kmers[pos];
        bool revComp = kmerCanonical(c, kmerSize);
        if (idx.repetitive.count(c)) { curFilteredPos.push_back(pos); continue; }
        auto it = idx.lists.find(c);
        if (it == idx.lists.end() || it->second.empty()) continue;
        for (uint64_t g : it->second) {
            SeqPos sp = seqPosition(r, g);
            if (revComp) { sp.pos = r.len(sp.id) - sp.pos - kmerSize; sp.id ^= 1; }
            if (sp.id == queryId && sp.pos == pos) continue;   // no trivial matches
            vecMatches.push_back({pos, sp.pos, sp.id});
        }
    }

    std::sort(vecMatches.begin(), vecMatches.end(), [](const KmerMatch& k1, const KmerMatch& k2) {
        return k1.extId != k2.extId ? k1.extId < k2.extId : k1.curPos < k2.curPos; });   // :201-204

    std::vector<Overlap> detected;
    std::vector<KmerMatch> matchesList;
    static thread_local std::vector<int32_t> scoreTable;
    std::vector<int32_t> backtrackTable;
    size_t extRangeBegin = 0, extRangeEnd = 0;
    while (extRangeEnd < vecMatches.size()) {
        if (maxOverlaps != 0 && detected.size() >= (size_t)maxOverlaps) break;   // :218-219
        extRangeBegin = extRangeEnd;
        size_t uniqueMatches = 0;
        int32_t prevPos = 0;
        while (extRangeEnd < vecMatches.size() && vecMatches[extRangeBegin].extId == vecMatches[extRangeEnd].extId) {
            if (vecMatches[extRangeEnd].curPos != prevPos) { ++uniqueMatches; prevPos = vecMatches[extRangeEnd].curPos; }
            ++extRangeEnd;
        }
        if (uniqueMatches < minKmerSruvivalRate * p.minOverlap) continue;   // :235

        matchesList.assign(vecMatches.begin() + extRangeBegin, vecMatches.begin() + extRangeEnd);
        uint32_t extId = matchesList.front().extId;
        int32_t extLen = r.len(extId);

        int32_t minCur = matchesList.front().curPos, maxCur = matchesList.back().curPos;   // :246-262
        int32_t minExt = std::numeric_limits<int32_t>::max(), maxExt = std::numeric_limits<int32_t>::min();
        for (const auto& m : matchesList) { minExt = std::min(minExt, m.extPos); maxExt = std::max(maxExt, m.extPos); }
        if (maxCur - minCur < p.minOverlap || maxExt - minExt < p.minOverlap) continue;
        if (checkOverhang && !forceLocal) {
            if (std::min(minCur, minExt) > p.maxOverhang) continue;
            if (std::min(curLen - maxCur, extLen - maxExt) > p.maxOverhang) continue;
        }

        scoreTable.assign(matchesList.size(), 0);        // :266-323 chaining DP
        backtrackTable.assign(matchesList.size(), -1);
        bool extSorted = extLen > curLen;
        if (extSorted)
            std::sort(matchesList.begin(), matchesList.end(),
                      [](const KmerMatch& k1, const KmerMatch& k2) { return k1.extPos < k2.extPos; });
        unsigned long long cells = 0;   // predecessors the scan visits (the integer work of the stage)
        for (int32_t i = 1; i < (int32_t)scoreTable.size(); ++i) {
            int32_t maxScore = 0, maxId = 0;
            int32_t curNext = matchesList[i].curPos, extNext = matchesList[i].extPos;
            for (int32_t j = i - 1; j >= 0; --j) {
                ++cells;
                int32_t curPrev = matchesList[j].curPos, extPrev = matchesList[j].extPos;
                if (0 < curNext - curPrev && curNext - curPrev < p.maxJump &&
                    0 < extNext - extPrev && extNext - extPrev < p.maxJump) {
                    int32_t matchScore = std::min(std::min(curNext - curPrev, extNext - extPrev), kmerSize);
                    int32_t jumpDiv = abs((curNext - curPrev) - (extNext - extPrev));
                    int32_t gapCost = (jumpDiv > 100 ? LG_GAP : SM_GAP) * jumpDiv;
                    int32_t nextScore = scoreTable[j] + matchScore - gapCost;
                    if (nextScore > maxScore) {
                        maxScore = nextScore; maxId = j;
                        if (jumpDiv == 0 && curNext - curPrev < kmerSize) break;
                    }
                }
                if (extSorted && extNext - extPrev > p.maxJump) break;
                if (!extSorted && curNext - curPrev > p.maxJump) break;
            }
            scoreTable[i] = std::max(maxScore, kmerSize);
            if (maxScore > kmerSize) backtrackTable[i] = maxId;
        }
        dpCellsVisited += cells;

        std::vector<Overlap> extOverlaps;                 // :326-427 chain extraction
        std::vector<std::pair<int32_t, int32_t>> kmerMatches;
        std::vector<size_t> orderedScores(backtrackTable.size());
        std::iota(orderedScores.begin(), orderedScores.end(), 0);
        std::sort(orderedScores.begin(), orderedScores.end(),
                  [](size_t a, size_t b) { return scoreTable[a] > scoreTable[b]; });
        for (int32_t chainStart : orderedScores) {
            if (backtrackTable[chainStart] == -1) continue;
            int32_t lastMatch = chainStart, firstMatch = 0, chainLength = 0;
            kmerMatches.clear();
            int32_t pos = chainStart;
            while (pos != -1) {
                firstMatch = pos;
                ++chainLength;
                if (p.keepAlignment) {
                    if (kmerMatches.empty() || kmerMatches.back().first - matchesList[pos].curPos > kmerSize)
                        kmerMatches.emplace_back(matchesList[pos].curPos, matchesList[pos].extPos);
                }
                int32_t newPos = backtrackTable[pos];
                backtrackTable[pos] = -1;
                pos = newPos;
            }
            Overlap o;
            o.curId = queryId; o.extId = matchesList.front().extId;
            o.curBegin = matchesList[firstMatch].curPos; o.extBegin = matchesList[firstMatch].extPos;
            o.curLen = curLen; o.extLen = extLen;
            o.curEnd = matchesList[lastMatch].curPos + kmerSize - 1;
            o.extEnd = matchesList[lastMatch].extPos + kmerSize - 1;
            o.score = scoreTable[lastMatch] - scoreTable[firstMatch] + kmerSize - 1;
            o.seqDivergence = 0.0f;
            if (overlapTest(o, p, forceLocal)) {
                if (p.keepAlignment) {
                    kmerMatches.emplace_back(o.curBegin, o.extBegin);
                    std::reverse(kmerMatches.begin(), kmerMatches.end());
                    kmerMatches.emplace_back(o.curEnd, o.extEnd);
                    o.kmerMatches.swap(kmerMatches);
                    kmerMatches.clear();
                }
                int32_t filteredPositions = 0;           // :409-425
                for (auto fp : curFilteredPos) {
                    if (fp < o.curBegin) continue;
                    if (fp > o.curEnd) break;
                    ++filteredPositions;
                }
                float normLen = std::max(curRange(o), extRange(o)) - filteredPositions;
                float matchRate = (float)chainLength * idx.sampleRate / normLen;
                matchRate = std::min(matchRate, 1.0f);
                o.seqDivergence = std::log(1 / matchRate) / kmerSize;
                extOverlaps.push_back(o);
            }
        }

        std::vector<Overlap> primary;                     // :431-458
        std::sort(extOverlaps.begin(), extOverlaps.end(),
                  [](const Overlap& r1, const Overlap& r2) { return r1.score > r2.score; });
        if (p.onlyMaxExt) {
            if (!extOverlaps.empty()) primary.push_back(extOverlaps.front());
        } else {
            for (const auto& o : extOverlaps) {
                bool isContained = false;
                for (const auto& prim : primary)
                    if (containedBy(o, prim) && prim.score > o.score) { isContained = true; break; }
                if (!isContained) primary.push_back(o);
            }
        }
        for (auto& o : primary) {                         // :461-473
            if (p.nuclAlignment) o.seqDivergence = alignmentErrEdlib(r, o, p.useHpc);
            if (o.seqDivergence < maxDivergence) detected.push_back(o);
            else if (p.partitionBadMappings)              // :475-485
                for (const auto& piece : checkIdyAndTrim(r, o, maxDivergence, p.minOverlap, p.useHpc)) detected.push_back(piece);
        }
    }
    return detected;
}

float estimateMaxDivergence(const Reads& r, const Index& idx, const Params& p) {   // overlap.cpp:744-827
    const int MAX_SEQS = 1000;
    std::vector<float> trueDivergence;
    for (int i = 0; i < MAX_SEQS; ++i) {
        uint32_t id = rand() % r.numIds();
        auto overlaps = getSeqOverlaps(r, idx, p, id, false, 0, p.maxDivergence);
        const Overlap* best = nullptr;
        for (const auto& o : overlaps)
            if (!best || curRange(o) > curRange(*best)) best = &o;
        if (best) trueDivergence.push_back(best->seqDivergence);
    }
    float mean = 0.5f;
    if (!trueDivergence.empty()) {   // median = quantile(vec, 50), utils.h:31-51
        std::sort(trueDivergence.begin(), trueDivergence.end());
        size_t target = std::min(trueDivergence.size() * (size_t)50 / 100, trueDivergence.size() - 1);
        mean = trueDivergence[target];
    }
    return (p.divergenceRelative ? mean : 0.0f) + p.ovlpDivergence;
}

// ------------------------------------------------------------------------------------------------
// literal model of libstdc++ 13 std::sort (bits/stl_algo.h:1848-1951, bits/stl_heap.h)
// ------------------------------------------------------------------------------------------------
namespace {
struct KI { uint64_t key; uint32_t idx; };
inline bool lt(const KI& a, const KI& b) { return a.key < b.key; }

void pushHeap(KI* first, long hole, long top, KI value) {          // __push_heap
    long parent = (hole - 1) / 2;
    while (hole > top && lt(first[parent], value)) { first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2; }
    first[hole] = value;
}
void adjustHeap(KI* first, long hole, long len, KI value) {       // __adjust_heap
    const long top = hole; long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (lt(first[child], first[child - 1])) --child;
        first[hole] = first[child]; hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); first[hole] = first[child - 1]; hole = child - 1; }
    pushHeap(first, hole, top, value);
}
void heapSortRange(KI* first, KI* last) {   // __partial_sort(first,last,last) = __heap_select + __sort_heap
    long len = last - first;
    if (len >= 2) for (long parent = (len - 2) / 2;; --parent) { KI v = first[parent]; adjustHeap(first, parent, len, v); if (parent == 0) break; }
    while (last - first > 1) { --last; KI v = *last; *last = *first; adjustHeap(first, 0, last - first, v); }
}
void moveMedianToFirst(KI* result, KI* a, KI* b, KI* c) {         // __move_median_to_first, :85-111
    if (lt(*a, *b)) { if (lt(*b, *c)) std::swap(*result, *b); else if (lt(*a, *c)) std::swap(*result, *c); else std::swap(*result, *a); }
    else if (lt(*a, *c)) std::swap(*result, *a);
    else if (lt(*b, *c)) std::swap(*result, *c);
    else std::swap(*result, *b);
}
KI* unguardedPartition(KI* first, KI* last, KI* pivot) {          // :1871-1889
    while (true) {
        while (lt(*first, *pivot)) ++first;
        --last;
        while (lt(*pivot, *last)) --last;
        if (!(first < last)) return first;
        std::swap(*first, *last);
        ++first;
    }
}
void introsortLoop(KI* first, KI* last, long depth) {             // :1918-1940
    while (last - first > 16) {
        if (depth == 0) { heapSortRange(first, last); return; }
        --depth;
        KI* mid = first + (last - first) / 2;
        moveMedianToFirst(first, first + 1, mid, last - 1);
        KI* cut = unguardedPartition(first + 1, last, first);
        introsortLoop(cut, last, depth);
        last = cut;
    }
}
void insertionSort(KI* first, KI* last) {   // guarded + unguarded variants give the same (stable) result
    for (KI* i = first + 1; i < last; ++i) {
        KI v = *i; KI* j = i;
        while (j > first && lt(v, *(j - 1))) { *j = *(j - 1); --j; }
        *j = v;
    }
}
}  // namespace

std::vector<uint32_t> introsort_model(const std::vector<uint64_t>& keys) {
    std::vector<KI> a(keys.size());
    for (size_t i = 0; i < keys.size(); ++i) a[i] = {keys[i], (uint32_t)i};
    if (a.size() > 1) {
        long lg = 63 - __builtin_clzll((unsigned long long)a.size());
        introsortLoop(a.data(), a.data() + a.size(), 2 * lg);
        insertionSort(a.data(), a.data() + a.size());   // __final_insertion_sort, :1854-1865
    }
    std::vector<uint32_t> perm(a.size());
    for (size_t i = 0; i < a.size(); ++i) perm[i] = a[i].idx;
    return perm;
}

}  // namespace restate
