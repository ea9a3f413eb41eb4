// flye_b200 host mirror — OverlapRange / OverlapDetector / OverlapContainer with the reference's interface
// (src/sequence/overlap.h:20-523, overlap.cpp:510-927).  getSeqOverlaps runs on the device through
// fg_overlaps_batch; the container keeps the reference's lazy cache semantics on top of batched device calls.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "vertex_index.h"
#include "sequence_container.h"
#include "../common/config.h"
#include "../common/logger.h"
#if __has_include("../common/progress_bar.h")
#include "../common/progress_bar.h"
#endif

#if __has_include("IntervalTree.h")
#include "IntervalTree.h"
#else
template <class T> struct Interval {   // what lib/interval_tree provides, for stand-alone builds
    Interval(int32_t s, int32_t e, const T& v) : start(s), stop(e), value(v) {}
    int32_t start, stop;
    T value;
};
template <class T> class IntervalTree {
public:
    IntervalTree() {}
    explicit IntervalTree(std::vector<Interval<T>>& ivs) : _ivs(ivs) {}
    std::vector<Interval<T>> findOverlapping(int32_t start, int32_t stop) const {
        std::vector<Interval<T>> out;
        for (const auto& iv : _ivs) if (iv.stop >= start && iv.start <= stop) out.push_back(iv);
        return out;
    }
private:
    std::vector<Interval<T>> _ivs;
};
#endif

struct OverlapRange {
    OverlapRange(FastaRecord::Id curId = FastaRecord::ID_NONE, FastaRecord::Id extId = FastaRecord::ID_NONE, int32_t curInit = 0,
                 int32_t extInit = 0, int32_t curLen = 0, int32_t extLen = 0)
        : curId(curId), curBegin(curInit), curEnd(curInit), curLen(curLen), extId(extId), extBegin(extInit), extEnd(extInit),
          extLen(extLen), score(0), seqDivergence(0.0f), kmerMatches(nullptr) {}
    ~OverlapRange() { delete kmerMatches; }
    OverlapRange(const OverlapRange& o) : kmerMatches(nullptr) { assign(o); }
    OverlapRange(OverlapRange&& o) : kmerMatches(nullptr) { assign(o, false); kmerMatches = o.kmerMatches; o.kmerMatches = nullptr; }
    OverlapRange& operator=(const OverlapRange& o) { if (this != &o) assign(o); return *this; }

    int32_t curRange() const { return curEnd - curBegin; }
    int32_t extRange() const { return extEnd - extBegin; }
    int32_t minRange() const { return std::min(curRange(), extRange()); }

    OverlapRange reverse() const {   // swap the roles of the two sequences
        OverlapRange r(*this);
        std::swap(r.curId, r.extId); std::swap(r.curBegin, r.extBegin);
        std::swap(r.curEnd, r.extEnd); std::swap(r.curLen, r.extLen);
        if (r.kmerMatches) {
            for (auto& p : *r.kmerMatches) std::swap(p.first, p.second);
            std::sort(r.kmerMatches->begin(), r.kmerMatches->end(),
                      [](const std::pair<int32_t, int32_t>& a, const std::pair<int32_t, int32_t>& b) { return a.first < b.first; });
        }
        return r;
    }
    OverlapRange complement() const {   // the same overlap seen from the opposite strands
        OverlapRange c(*this);
        c.curBegin = curLen - curEnd - 1; c.curEnd = curLen - curBegin - 1;
        c.extBegin = extLen - extEnd - 1; c.extEnd = extLen - extBegin - 1;
        c.curId = curId.rc(); c.extId = extId.rc();
        if (c.kmerMatches) {
            for (auto& p : *c.kmerMatches) { p.first = curLen - p.first - 1; p.second = extLen - p.second - 1; }
            std::reverse(c.kmerMatches->begin(), c.kmerMatches->end());
        }
        return c;
    }
    int32_t project(int32_t curPos) const {
        if (curPos <= curBegin) return extBegin;
        if (curPos >= curEnd) return extEnd;
        if (!kmerMatches) {
            float ratio = (float)extRange() / curRange();
            int32_t p = extBegin + float(curPos - curBegin) * ratio;
            return std::max(extBegin, std::min(p, extEnd));
        }
        size_t i = std::lower_bound(kmerMatches->begin(), kmerMatches->end(), curPos,
                                    [](const std::pair<int32_t, int32_t>& pr, int32_t v) { return pr.first < v; }) - kmerMatches->begin();
        if (i == 0 || i == kmerMatches->size()) throw std::runtime_error("Error in overlap projection");
        const auto &a = (*kmerMatches)[i - 1], &b = (*kmerMatches)[i];
        float ratio = (float)(b.second - a.second) / (b.first - a.first);
        int32_t p = a.second + float(curPos - a.first) * ratio;
        return std::max(a.second, std::min(p, b.second));
    }
    int32_t leftShift() const { return curBegin - extBegin; }
    int32_t rightShift() const { return (extLen - extEnd) - (curLen - curEnd); }
    int32_t lrOverhang() const { return std::max(std::min(curBegin, extBegin), std::min(curLen - curEnd, extLen - extEnd)); }
    bool contains(int32_t curPos, int32_t extPos) const {
        return curBegin <= curPos && curPos <= curEnd && extBegin <= extPos && extPos <= extEnd;
    }
    bool containedBy(const OverlapRange& o) const {
        if (curId != o.curId || extId != o.extId) return false;
        return o.curBegin <= curBegin && curEnd <= o.curEnd && o.extBegin <= extBegin && extEnd <= o.extEnd;
    }
    int32_t curIntersect(const OverlapRange& o) const { return std::min(curEnd, o.curEnd) - std::max(curBegin, o.curBegin); }
    int32_t extIntersect(const OverlapRange& o) const { return std::min(extEnd, o.extEnd) - std::max(extBegin, o.extBegin); }
    void dump(std::ostream& os, const SequenceContainer& curContainer, const SequenceContainer& extContainer) {
        os << curContainer.seqName(curId) << " " << curBegin << " " << curEnd << " " << curLen << " " << extContainer.seqName(extId)
           << " " << extBegin << " " << extEnd << " " << extLen << " " << -1 << " " << -1 << " " << score << " " << seqDivergence;
    }
    void load(std::istream& is, const SequenceContainer& curContainer, const SequenceContainer& extContainer) {
        int32_t pad1, pad2; std::string curName, extName;
        is >> curName >> curBegin >> curEnd >> curLen >> extName >> extBegin >> extEnd >> extLen >> pad1 >> pad2 >> score >> seqDivergence;
        curId = curContainer.recordByName(curName).id;
        extId = extContainer.recordByName(extName).id;
    }

    FastaRecord::Id curId; int32_t curBegin, curEnd, curLen;
    FastaRecord::Id extId; int32_t extBegin, extEnd, extLen;
    int32_t score;
    float seqDivergence;
    std::vector<std::pair<int32_t, int32_t>>* kmerMatches;

private:
    void assign(const OverlapRange& o, bool deep = true) {
        curId = o.curId; curBegin = o.curBegin; curEnd = o.curEnd; curLen = o.curLen;
        extId = o.extId; extBegin = o.extBegin; extEnd = o.extEnd; extLen = o.extLen;
        score = o.score; seqDivergence = o.seqDivergence;
        if (!deep) return;
        if (!o.kmerMatches) { delete kmerMatches; kmerMatches = nullptr; }
        else { if (!kmerMatches) kmerMatches = new std::vector<std::pair<int32_t, int32_t>>(); *kmerMatches = *o.kmerMatches; }
    }
};

// KSW2 trimming of a primary overlap that failed the divergence test (_partitionBadMappings, overlap.cpp:475-485).  The
// function is the reference's own (alignment.cpp:306-495, with lib/minimap2's ksw2): inside flye-modules it is linked in as
// before; a stand-alone build that never sets partitionBadMappings does not need it (weak reference).
std::vector<OverlapRange> checkIdyAndTrim(OverlapRange& ovlp, const DnaSequence& curSeq, const DnaSequence& extSeq, float maxDivergence,
                                          int32_t minOverlap, bool useHpc) __attribute__((weak));

// The same trimming with the banded alignment on the device (FLYE_B200_DEVICE_KSW=1; default off until the repeat stage has run with
// it on a GPU): homopolymer compression (alignment.cpp:52-70), then fg_align_cigar_batch for all rejected overlaps of a call in one
// batch, then the host tail below — the CIGAR's =/X split (alignment.cpp:172-211), the search for the longest stretches below the
// divergence threshold, their greedy non-overlapping choice (std::sort by length: the same libstdc++ permutation as in the
// reference) and the mapping back to uncompressed coordinates (alignment.cpp:330-452).  tests/test_oracle_trim.py runs the tail on
// CIGARs of the routine's host build and compares the pieces with the reference's checkIdyAndTrim.
namespace flye_b200 {
struct HpcRange { std::vector<uint8_t> seq; std::vector<int32_t> offsetTable; };
inline HpcRange hpcRange(const DnaSequence& s, int32_t start, int32_t length, bool compress) {
    HpcRange h;
    h.seq.reserve(length); h.offsetTable.reserve(length);
    for (int32_t i = 0; i < length; ++i) {
        const uint8_t b = (uint8_t)s.atRaw((size_t)start + i);
        if (!compress || i == 0 || h.seq.back() != b) { h.seq.push_back(b); h.offsetTable.push_back(i); }
    }
    return h;
}
// kswCigar: n entries len << 4 | op (0 = M, 1 = I: ext only, 2 = D: cur only) of the alignment of ext.seq (query) against cur.seq (target)
inline std::vector<OverlapRange> trimByCigar(const OverlapRange& ovlp, const HpcRange& cur, const HpcRange& ext, const uint32_t* kswCigar, size_t n,
                                             float maxDivergence, int32_t minOverlap) {
    struct Op { char op; int len; };
    std::vector<Op> cigar;
    size_t posQry = 0, posTrg = 0;
    for (size_t c = 0; c < n; ++c) {
        const int size = (int)(kswCigar[c] >> 4);
        const uint32_t code = kswCigar[c] & 0xf;
        if (code == 0) {
            for (int k = 0; k < size; ++k) {
                const char match = cur.seq[posTrg + k] == ext.seq[posQry + k] ? '=' : 'X';
                if (k == 0 || match != cigar.back().op) cigar.push_back({match, 1});
                else ++cigar.back().len;
            }
            posQry += size; posTrg += size;
        } else if (code == 1) { cigar.push_back({'I', size}); posQry += size; }
        else { cigar.push_back({'D', size}); posTrg += size; }
    }
    const int m = (int)cigar.size();
    std::vector<int> sumErrors(1, 0), sumCurLen(1, 0), sumExtLen(1, 0);
    for (const Op& op : cigar) {
        const int curConsumed = op.op == 'I' ? 0 : op.len, extConsumed = op.op == 'D' ? 0 : op.len;
        sumCurLen.push_back(sumCurLen.back() + curConsumed);
        sumExtLen.push_back(sumExtLen.back() + extConsumed);
        sumErrors.push_back(sumErrors.back() + (op.op != '=' ? op.len : 0));
    }
    struct IntervalDiv { int start, end; float divergence; int realLen; };
    std::vector<IntervalDiv> good;
    for (int intLen = m; intLen > 0; --intLen)
        for (int intStart = 0; intStart < m - intLen + 1; ++intStart) {
            const int i = intStart, j = intStart + intLen - 1;
            if (cigar[i].op != '=' || cigar[j].op != '=') continue;
            const int rangeLen = std::max(sumCurLen[j + 1] - sumCurLen[i], sumExtLen[j + 1] - sumExtLen[i]);
            const float divergence = float(sumErrors[j + 1] - sumErrors[i]) / rangeLen;
            if (divergence < maxDivergence) good.push_back({i, j, divergence, rangeLen});
        }
    std::sort(good.begin(), good.end(), [](const IntervalDiv& a, const IntervalDiv& b) { return a.realLen > b.realLen; });
    std::vector<IntervalDiv> chosen;
    for (const IntervalDiv& iv : good) {
        bool intersects = false;
        for (const IntervalDiv& o : chosen)
            if (std::min(iv.end + 1, o.end + 1) - std::max(iv.start, o.start) > 0) { intersects = true; break; }
        if (!intersects) chosen.push_back(iv);
    }
    std::vector<OverlapRange> trimmed;
    for (const IntervalDiv& cand : chosen) {
        OverlapRange o = ovlp;
        o.seqDivergence = cand.divergence;
        size_t pq = 0, pt = 0;
        for (int i = 0; i < m; ++i) {
            if (i == cand.start) { o.curBegin += cur.offsetTable[pt]; o.extBegin += ext.offsetTable[pq]; }
            if (cigar[i].op == '=' || cigar[i].op == 'X') { pq += cigar[i].len; pt += cigar[i].len; }
            else if (cigar[i].op == 'I') pq += cigar[i].len;
            else pt += cigar[i].len;
            if (i == cand.end) { o.curEnd = ovlp.curBegin + cur.offsetTable[pt - 1]; o.extEnd = ovlp.extBegin + ext.offsetTable[pq - 1]; }
        }
        if (o.curRange() > minOverlap && o.extRange() > minOverlap) trimmed.push_back(o);
    }
    return trimmed;
}
inline bool deviceKswEnabled() { const char* e = std::getenv("FLYE_B200_DEVICE_KSW"); return e && std::atoi(e) != 0; }
}  // namespace flye_b200

struct OvlpDivStats {
    static const size_t MAX_STATS = 1000000;
    OvlpDivStats() : divVec(MAX_STATS), vecSize(0) {}
    void add(float val) {
        size_t slot = vecSize.fetch_add(1);
        if (slot < MAX_STATS) divVec[slot] = val; else vecSize = MAX_STATS;
    }
    std::vector<float> divVec;
    std::atomic<size_t> vecSize;
};

class OverlapDetector {
public:
    OverlapDetector(const SequenceContainer& seqContainer, const VertexIndex& vertexIndex, int maxJump, int minOverlap, int maxOverhang,
                    bool keepAlignment, bool onlyMaxExt, float maxDivergence, bool nuclAlignment, bool partitionBadMappings, bool useHpc)
        : _maxJump(maxJump), _minOverlap(minOverlap), _maxOverhang(maxOverhang), _maxCurOverlaps(0), _checkOverhang(maxOverhang > 0),
          _keepAlignment(keepAlignment), _onlyMaxExt(onlyMaxExt), _nuclAlignment(nuclAlignment),
          _partitionBadMappings(partitionBadMappings), _useHpc(useHpc), _maxDivergence(maxDivergence), _vertexIndex(vertexIndex),
          _seqContainer(seqContainer) {
        if (partitionBadMappings && !checkIdyAndTrim && !flye_b200::deviceKswEnabled())
            throw std::runtime_error("flye_b200: partitionBadMappings needs the reference's alignment.cpp (checkIdyAndTrim) linked in (or FLYE_B200_DEVICE_KSW=1)");
    }
    friend class OverlapContainer;

private:
    // getSeqOverlaps for many query sequences in one device pass (reference: one call per read, overlap.cpp:99-508)
    // `queryContainer` may be a different container than the indexed one (read-to-graph alignment): its forward
    // sequences are then uploaded once as the device's second sequence set
    std::vector<std::vector<OverlapRange>> getSeqOverlapsBatch(const std::vector<FastaRecord::Id>& queries, bool forceLocal,
                                                               OvlpDivStats& divStats, int maxOverlaps,
                                                               const SequenceContainer* queryContainer = nullptr) const {
        const auto& dev = _vertexIndex.device();
        if (&_vertexIndex.container() != &_seqContainer) throw std::runtime_error("flye_b200: index and target containers differ");
        if (!queryContainer) queryContainer = &_seqContainer;
        const bool sameSet = queryContainer == &_seqContainer;
        const uint32_t base = (uint32_t)_seqContainer.idOffset();              // ids of the indexed (target) sequences
        const uint32_t qbase = (uint32_t)queryContainer->idOffset();           // ids of the query sequences
        std::vector<uint32_t> ids(queries.size());
        for (size_t i = 0; i < queries.size(); ++i) ids[i] = queries[i].rawId() - qbase;
        fg_overlap_params p{};
        p.query_set = sameSet ? 0 : 1;
        p.max_jump = _maxJump; p.min_overlap = _minOverlap; p.max_overhang = _maxOverhang; p.max_overlaps = maxOverlaps;
        p.force_local = forceLocal; p.keep_alignment = _keepAlignment; p.only_max_ext = _onlyMaxExt; p.nucl_alignment = _nuclAlignment;
        p.use_hpc = _useHpc; p.max_divergence = _maxDivergence;
        p.keep_rejected = _partitionBadMappings;   // the primary overlaps that fail the test come back flagged and are trimmed below
        // results live in library memory until the next call: copy out under the lock
        std::lock_guard<std::mutex> lock(batchMutex());
        if (!sameSet) dev->uploadQueries(*queryContainer);
        fg_overlap_result res;
        dev->check(fg_overlaps_batch(dev->ctx, ids.data(), (uint32_t)ids.size(), &p, &res));
        std::vector<std::vector<OverlapRange>> out(queries.size());
        struct Rejected { size_t q, at; OverlapRange ovlp; Rejected(size_t q, size_t at, OverlapRange&& o) : q(q), at(at), ovlp(std::move(o)) {} };
        std::vector<Rejected> rejected;
        for (size_t q = 0; q < queries.size(); ++q) {
            out[q].reserve(res.offsets[q + 1] - res.offsets[q]);
            for (uint64_t i = res.offsets[q]; i < res.offsets[q + 1]; ++i) {
                const fg_overlap& o = res.overlaps[i];
                OverlapRange r(FastaRecord::Id(qbase + o.cur_id), FastaRecord::Id(base + o.ext_id), o.cur_begin, o.ext_begin, o.cur_len, o.ext_len);
                r.curEnd = o.cur_end; r.extEnd = o.ext_end; r.score = o.score; r.seqDivergence = o.seq_divergence;
                if (_keepAlignment && o.aln_count) {
                    r.kmerMatches = new std::vector<std::pair<int32_t, int32_t>>();
                    r.kmerMatches->reserve(o.aln_count);
                    for (uint64_t a = o.aln_first; a < o.aln_first + o.aln_count; ++a)
                        r.kmerMatches->emplace_back(res.aln_pairs[2 * a], res.aln_pairs[2 * a + 1]);
                }
                if (o.reserved & 1u) {   // failed the divergence test: its parts may pass (overlap.cpp:475-485); host side, the reference's KSW2
                    rejected.emplace_back(q, out[q].size(), std::move(r));
                    continue;
                }
                out[q].push_back(std::move(r));
            }
            // per-10kb-window divergence statistics (overlap.cpp:488-506).  The reference records every PRIMARY overlap before the
            // divergence test; the device only returns the rejected ones when asked to (keep_rejected, i.e. partitionBadMappings),
            // so otherwise the logged quantiles (overlapDivergenceStats) are those of the kept overlaps — results are unaffected
            const int STAT_WND = 10000;
            if (!out[q].empty()) {
                std::vector<const OverlapRange*> wnd(out[q].front().curLen / STAT_WND + 1, nullptr);
                for (const auto& r : out[q]) { auto& w = wnd[r.curBegin / STAT_WND]; if (!w || r.curRange() > w->curRange()) w = &r; }
                for (auto* w : wnd) if (w && w->curRange() > 0) divStats.add(w->seqDivergence);
            }
        }
        if (!rejected.empty() && flye_b200::deviceKswEnabled()) {
            // the banded alignments of all rejected overlaps in one device batch; compression and the interval search on host threads
            std::vector<flye_b200::HpcRange> cur(rejected.size()), ext(rejected.size());
            std::vector<std::vector<OverlapRange>> pieces(rejected.size());
            auto parallel = [&](const std::function<void(size_t)>& fn) {
                std::atomic<size_t> next(0);
                auto work = [&]() { for (size_t i; (i = next.fetch_add(1)) < rejected.size();) fn(i); };
                std::vector<std::thread> pool;
                const size_t nThreads = std::min<size_t>(std::max<size_t>(1, Parameters::get().numThreads), rejected.size());
                for (size_t t = 1; t < nThreads; ++t) pool.emplace_back(work);
                work();
                for (auto& t : pool) t.join();
            };
            parallel([&](size_t i) {
                const OverlapRange& o = rejected[i].ovlp;
                cur[i] = flye_b200::hpcRange(queryContainer->getSeq(o.curId), o.curBegin, o.curRange(), _useHpc);
                ext[i] = flye_b200::hpcRange(_seqContainer.getSeq(o.extId), o.extBegin, o.extRange(), _useHpc);
            });
            std::vector<uint8_t> T, Q; std::vector<uint64_t> tOff(1, 0), qOff(1, 0);
            size_t longest = 1;
            for (size_t i = 0; i < rejected.size(); ++i) {
                T.insert(T.end(), cur[i].seq.begin(), cur[i].seq.end()); tOff.push_back(T.size());
                Q.insert(Q.end(), ext[i].seq.begin(), ext[i].seq.end()); qOff.push_back(Q.size());
                longest = std::max(longest, cur[i].seq.size() + ext[i].seq.size() + 2);
            }
            const uint32_t cap = (uint32_t)std::min<size_t>(longest, (size_t)1 << 20);   // a CIGAR has at most tlen + qlen entries
            std::vector<uint32_t> cigars(rejected.size() * (size_t)cap), nCigar(rejected.size());
            std::vector<int32_t> status(rejected.size());
            dev->check(fg_align_cigar_batch(dev->ctx, T.data(), tOff.data(), Q.data(), qOff.data(), (uint32_t)rejected.size(), cap, cigars.data(), nCigar.data(),
                                          status.data()));
            for (int32_t st : status) if (st == 2) throw std::runtime_error("flye_b200: CIGAR capacity exceeded in the device alignment");
            parallel([&](size_t i) {
                pieces[i] = flye_b200::trimByCigar(rejected[i].ovlp, cur[i], ext[i], cigars.data() + i * (size_t)cap, nCigar[i], _maxDivergence, _minOverlap);
            });
            for (size_t i = rejected.size(); i-- > 0;)   // back to front: the insertion points of earlier ones stay valid
                out[rejected[i].q].insert(out[rejected[i].q].begin() + rejected[i].at, pieces[i].begin(), pieces[i].end());
        } else if (!rejected.empty()) {
            // KSW2 trimming on the host, in parallel; the pieces take the place of the rejected overlap in its query's list
            std::vector<std::vector<OverlapRange>> pieces(rejected.size());
            std::atomic<size_t> next(0);
            auto work = [&]() {
                for (size_t i; (i = next.fetch_add(1)) < rejected.size();) {
                    Rejected& r = rejected[i];
                    pieces[i] = checkIdyAndTrim(r.ovlp, queryContainer->getSeq(r.ovlp.curId), _seqContainer.getSeq(r.ovlp.extId), _maxDivergence,
                                                _minOverlap, _useHpc);
                }
            };
            std::vector<std::thread> pool;
            const size_t nThreads = std::min<size_t>(std::max<size_t>(1, Parameters::get().numThreads), rejected.size());
            for (size_t t = 1; t < nThreads; ++t) pool.emplace_back(work);
            work();
            for (auto& t : pool) t.join();
            for (size_t i = rejected.size(); i-- > 0;)   // back to front: the insertion points of earlier ones stay valid
                out[rejected[i].q].insert(out[rejected[i].q].begin() + rejected[i].at, pieces[i].begin(), pieces[i].end());
        }
        return out;
    }
    std::vector<OverlapRange> getSeqOverlaps(const FastaRecord& rec, bool forceLocal, OvlpDivStats& stats, int maxOverlaps) const {
        return std::move(getSeqOverlapsBatch({rec.id}, forceLocal, stats, maxOverlaps)[0]);   // queries of the indexed container
    }
    static std::mutex& batchMutex() { static std::mutex m; return m; }

    const int _maxJump, _minOverlap, _maxOverhang, _maxCurOverlaps;
    const bool _checkOverhang, _keepAlignment, _onlyMaxExt, _nuclAlignment, _partitionBadMappings, _useHpc;
    mutable float _maxDivergence;
    const VertexIndex& _vertexIndex;
    const SequenceContainer& _seqContainer;
};

class OverlapContainer {
public:
    OverlapContainer(const OverlapDetector& ovlpDetect, const SequenceContainer& queryContainer)
        : _ovlpDetect(ovlpDetect), _queryContainer(queryContainer), _indexSize(0), _meanTrueOvlpDiv(0) {}
    struct IndexVecWrapper {
        IndexVecWrapper() : fwdOverlaps(new std::vector<OverlapRange>), revOverlaps(new std::vector<OverlapRange>), cached(false),
                            suggestChimeric(false) {}
        std::shared_ptr<std::vector<OverlapRange>> fwdOverlaps, revOverlaps;
        bool cached, suggestChimeric;
    };

    // thread safe: computes and caches on first use; both strands are stored (overlap.cpp:528-574).  The callers ask for one
    // read at a time from many threads (extender.cpp:32,44,98,244,339; chimera.cpp:77), the device wants batches: the FIRST miss
    // computes the overlaps of every forward read of the query container in one device pass (exactly what the reference's cache
    // holds once every read has been asked for: getSeqOverlaps(read, forceLocal = false, maxOverlaps = 0) under the detector's
    // current threshold), later calls are cache hits.  FLYE_B200_LAZY_PREFETCH=0 restores one device pass per miss.
    const std::vector<OverlapRange>& lazySeqOverlaps(FastaRecord::Id readId) {
        const bool flipped = !readId.strand();
        if (flipped) readId = readId.rc();
        {
            std::lock_guard<std::mutex> lock(_cacheMutex);
            auto it = _overlapIndex.find(readId);
            if (it != _overlapIndex.end() && it->second.cached) return flipped ? *it->second.revOverlaps : *it->second.fwdOverlaps;
        }
        static const bool prefetchAll = !(getenv("FLYE_B200_LAZY_PREFETCH") && atoi(getenv("FLYE_B200_LAZY_PREFETCH")) == 0);
        if (prefetchAll) {
            std::lock_guard<std::mutex> once(_prefetchMutex);   // the threads that missed meanwhile wait here and then find their read cached
            if (!_prefetchedAll) {
                std::vector<FastaRecord::Id> all;
                for (const auto& seq : _queryContainer.iterSeqs()) if (seq.id.strand()) all.push_back(seq.id);
                prefetch(all);
                _prefetchedAll = true;
            }
        }
        prefetch({readId});
        std::lock_guard<std::mutex> lock(_cacheMutex);
        const auto& w = _overlapIndex[readId];
        return flipped ? *w.revOverlaps : *w.fwdOverlaps;
    }
    bool hasSelfOverlaps(FastaRecord::Id readId) {
        this->lazySeqOverlaps(readId);
        std::lock_guard<std::mutex> lock(_cacheMutex);
        return _overlapIndex[readId.strand() ? readId : readId.rc()].suggestChimeric;
    }
    std::vector<OverlapRange> quickSeqOverlaps(FastaRecord::Id readId, int maxOverlaps = 0, bool forceLocal = false) {
        return std::move(_ovlpDetect.getSeqOverlapsBatch({readId}, forceLocal, _divergenceStats, maxOverlaps, &_queryContainer)[0]);
    }
    // mirror extension: the same results as N quickSeqOverlaps calls from one device pass
    std::vector<std::vector<OverlapRange>> quickSeqOverlapsBatch(const std::vector<FastaRecord::Id>& ids, int maxOverlaps = 0,
                                                                 bool forceLocal = false) {
        return _ovlpDetect.getSeqOverlapsBatch(ids, forceLocal, _divergenceStats, maxOverlaps, &_queryContainer);
    }
    // mirror extension: fill the lazy cache for many forward reads with one device pass
    void prefetch(const std::vector<FastaRecord::Id>& fwdIds) {
        std::vector<FastaRecord::Id> todo;
        {
            std::lock_guard<std::mutex> lock(_cacheMutex);
            for (auto id : fwdIds) { auto it = _overlapIndex.find(id); if (it == _overlapIndex.end() || !it->second.cached) todo.push_back(id); }
        }
        if (todo.empty()) return;
        auto res = _ovlpDetect.getSeqOverlapsBatch(todo, false, _divergenceStats, _ovlpDetect._maxCurOverlaps, &_queryContainer);
        std::lock_guard<std::mutex> lock(_cacheMutex);
        for (size_t i = 0; i < todo.size(); ++i) {
            auto& w = _overlapIndex[todo[i]];
            if (w.cached) continue;
            w.revOverlaps->reserve(res[i].size());
            for (const auto& o : res[i]) w.revOverlaps->push_back(o.complement());
            _indexSize += res[i].size();
            *w.fwdOverlaps = std::move(res[i]);
            w.cached = true;
        }
    }
    size_t indexSize() { return _indexSize; }

    void estimateOverlaperParameters() {   // overlap.cpp:744-817: 1000 ids drawn with rand() % size, either strand
        Logger::get().debug() << "Estimating k-mer identity bias";
        const int MAX_SEQS = 1000;
        std::vector<FastaRecord::Id> ids;
        for (int i = 0; i < MAX_SEQS; ++i) ids.push_back(_queryContainer.iterSeqs()[rand() % _queryContainer.iterSeqs().size()].id);
        auto res = _ovlpDetect.getSeqOverlapsBatch(ids, false, _divergenceStats, 0, &_queryContainer);
        std::vector<float> divs;
        for (const auto& ovs : res) {
            const OverlapRange* best = nullptr;
            for (const auto& o : ovs) if (!best || o.curRange() > best->curRange()) best = &o;
            if (best) divs.push_back(best->seqDivergence);
        }
        if (!divs.empty()) {
            std::sort(divs.begin(), divs.end());
            _meanTrueOvlpDiv = divs[std::min(divs.size() * (size_t)50 / 100, divs.size() - 1)];
            _divergenceStats.vecSize = 0;
        } else {
            Logger::get().warning() << "No overlaps found - unable to estimate parameters";
            _meanTrueOvlpDiv = 0.5f;
        }
        Logger::get().debug() << "Initial divergence estimate : " << _meanTrueOvlpDiv;
    }
    void setDivergenceThreshold(float threshold, bool isRelative) {
        _ovlpDetect._maxDivergence = (isRelative ? _meanTrueOvlpDiv : 0.0f) + threshold;
        Logger::get().debug() << "Max divergence threshold set to " << _ovlpDetect._maxDivergence;
    }

    // the functions below are NOT thread safe (as in the reference)
    void ensureTransitivity(bool onlyMaxExt) {   // overlap.cpp:576-627
        std::vector<FastaRecord::Id> all;
        for (const auto& kv : _overlapIndex) { all.push_back(kv.first); all.push_back(kv.first.rc()); }
        for (auto seq : all) {
            std::vector<OverlapRange> toAdd;
            for (const auto& cur : unsafeSeqOverlaps(seq)) {
                if (onlyMaxExt) {
                    bool found = false;
                    for (auto& ext : unsafeSeqOverlaps(cur.extId))
                        if (ext.extId == cur.curId) { if (cur.score > ext.score) ext = cur.reverse(); found = true; break; }
                    if (!found) toAdd.push_back(cur.reverse());
                } else toAdd.push_back(cur.reverse());
            }
            for (const auto& o : toAdd) unsafeSeqOverlaps(o.curId).push_back(o);
        }
    }
    void overlapDivergenceStats() { overlapDivergenceStats(_divergenceStats, _ovlpDetect._maxDivergence); }
    void overlapDivergenceStats(const OvlpDivStats& stats, float divCutoff) {
        std::vector<float> v(stats.divVec.begin(), stats.divVec.begin() + std::min<size_t>(stats.vecSize, OvlpDivStats::MAX_STATS));
        if (v.empty()) return;
        std::sort(v.begin(), v.end());
        auto q = [&](int pct) { return v[std::min(v.size() * (size_t)pct / 100, v.size() - 1)]; };
        Logger::get().info() << "Median overlap divergence: " << q(50);
        Logger::get().debug() << "Sequence divergence distribution: Q25 = " << q(25) << ", Q50 = " << q(50) << ", Q75 = " << q(75)
                              << " (cutoff " << divCutoff << ")";
    }
    void findAllOverlaps() {   // overlap.cpp:630-668: all forward reads in one device pass, then closure + de-duplication
        std::vector<FastaRecord::Id> all;
        for (const auto& seq : _queryContainer.iterSeqs()) if (seq.id.strand()) all.push_back(seq.id);
        prefetch(all);
        static const bool hostClosure = getenv("FLYE_B200_HOST_CLOSURE") && atoi(getenv("FLYE_B200_HOST_CLOSURE")) != 0;
        if (!hostClosure && &_queryContainer == &_ovlpDetect._seqContainer) { deviceClosure(all); return; }
        ensureTransitivity(false);
        filterOverlaps();
    }
    void buildIntervalTree() {
        for (auto& kv : _overlapIndex)
            for (auto id : {kv.first, kv.first.rc()}) {
                std::vector<Interval<const OverlapRange*>> ivs;
                for (const auto& o : unsafeSeqOverlaps(id)) ivs.emplace_back(o.curBegin, o.curEnd, &o);
                _ovlpTree[id] = IntervalTree<const OverlapRange*>(ivs);
            }
    }
    std::vector<Interval<const OverlapRange*>> getCoveringOverlaps(FastaRecord::Id seqId, int32_t start, int32_t end) const {
        return _ovlpTree.at(seqId).findOverlapping(start, end);
    }

private:
    // ensureTransitivity(false) + filterOverlaps() on the device (fg_overlaps_closure): the cached vectors of the forward
    // sequences go down as flat records, the lists of every sequence (both strands) come back as (input record, variant)
    // pairs; the variants are rebuilt with reverse() / complement(), which also carry kmerMatches along
    void deviceClosure(const std::vector<FastaRecord::Id>& fwdIds) {
        const auto& dev = _ovlpDetect._vertexIndex.device();
        const uint32_t base = (uint32_t)_queryContainer.idOffset();
        std::vector<fg_overlap> flat;
        std::vector<const OverlapRange*> src;
        for (auto id : fwdIds)
            for (const auto& o : *_overlapIndex[id].fwdOverlaps) {
                fg_overlap r{};
                r.cur_id = o.curId.rawId() - base; r.cur_begin = o.curBegin; r.cur_end = o.curEnd; r.cur_len = o.curLen;
                r.ext_id = o.extId.rawId() - base; r.ext_begin = o.extBegin; r.ext_end = o.extEnd; r.ext_len = o.extLen;
                r.score = o.score; r.seq_divergence = o.seqDivergence;
                flat.push_back(r); src.push_back(&o);
            }
        const uint32_t nSeqs = (uint32_t)_queryContainer.iterSeqs().size();
        std::vector<std::vector<OverlapRange>> lists(nSeqs);
        {
            std::lock_guard<std::mutex> lock(OverlapDetector::batchMutex());   // results live in library memory until the next call
            fg_overlap_result res;
            dev->check(fg_overlaps_closure(dev->ctx, flat.data(), flat.size(), nSeqs, (int32_t)Parameters::get().kmerSize, &res));
            for (uint32_t s = 0; s < nSeqs; ++s) {
                lists[s].reserve(res.offsets[s + 1] - res.offsets[s]);
                for (uint64_t i = res.offsets[s]; i < res.offsets[s + 1]; ++i) {
                    const uint32_t code = res.overlaps[i].reserved;
                    OverlapRange o = *src[code >> 2];
                    if (code & 1u) o = o.complement();
                    if (code & 2u) o = o.reverse();
                    lists[s].push_back(std::move(o));
                }
            }
        }
        size_t total = 0;
        for (uint32_t s = 0; s < nSeqs; ++s) {
            const FastaRecord::Id id(base + s);
            auto& w = _overlapIndex[id.strand() ? id : id.rc()];
            w.cached = true;
            total += lists[s].size();
            *(id.strand() ? w.fwdOverlaps : w.revOverlaps) = std::move(lists[s]);
        }
        _indexSize = total;
    }
    std::vector<OverlapRange>& unsafeSeqOverlaps(FastaRecord::Id id) {
        auto& w = _overlapIndex[id.strand() ? id : id.rc()];
        return id.strand() ? *w.fwdOverlaps : *w.revOverlaps;
    }
    void filterOverlaps() {   // overlap.cpp:681-741: cluster near-identical overlaps (ends within k), keep the best of each
        const int MAX_ENDS_DIFF = (int)Parameters::get().kmerSize;
        for (const auto& seq : _queryContainer.iterSeqs()) {
            auto& ovs = unsafeSeqOverlaps(seq.id);
            std::vector<size_t> parent(ovs.size());
            for (size_t i = 0; i < parent.size(); ++i) parent[i] = i;
            auto find = [&](size_t x) { while (parent[x] != x) x = parent[x] = parent[parent[x]]; return x; };
            for (size_t i = 0; i < ovs.size(); ++i)
                for (size_t j = 0; j < ovs.size(); ++j) {
                    if (ovs[i].extId != ovs[j].extId) continue;
                    if (ovs[i].curRange() - ovs[i].curIntersect(ovs[j]) < MAX_ENDS_DIFF &&
                        ovs[i].extRange() - ovs[i].extIntersect(ovs[j]) < MAX_ENDS_DIFF) parent[find(i)] = find(j);
                }
            std::unordered_map<size_t, size_t> best;   // cluster -> index of the first maximum score
            for (size_t i = 0; i < ovs.size(); ++i) {
                auto it = best.find(find(i));
                if (it == best.end()) best[find(i)] = i; else if (ovs[i].score > ovs[it->second].score) it->second = i;
            }
            std::vector<OverlapRange> kept;
            for (const auto& kv : best) kept.push_back(ovs[kv.second]);
            std::sort(kept.begin(), kept.end(), [](const OverlapRange& a, const OverlapRange& b) { return a.curBegin < b.curBegin; });
            ovs = std::move(kept);
        }
    }

    const OverlapDetector& _ovlpDetect;
    const SequenceContainer& _queryContainer;
    OvlpDivStats _divergenceStats;
    std::mutex _cacheMutex, _prefetchMutex;
    bool _prefetchedAll = false;
    std::unordered_map<FastaRecord::Id, IndexVecWrapper> _overlapIndex;
    std::atomic<size_t> _indexSize;
    std::unordered_map<FastaRecord::Id, IntervalTree<const OverlapRange*>> _ovlpTree;
    float _meanTrueOvlpDiv;
};

// iterate overlaps without large overhangs (overlap.h:457-523)
class OvlpIterator {
public:
    OvlpIterator(std::vector<OverlapRange>::const_iterator it, std::vector<OverlapRange>::const_iterator end, bool onlyNoOverhang)
        : it(it), end(end), onlyNoOverhang(onlyNoOverhang) { skip(); }
    bool operator==(const OvlpIterator& o) const { return it == o.it && onlyNoOverhang == o.onlyNoOverhang; }
    bool operator!=(const OvlpIterator& o) const { return !(*this == o); }
    const OverlapRange& operator*() const { return *it; }
    OvlpIterator& operator++() { ++it; skip(); return *this; }
private:
    void skip() {
        if (!onlyNoOverhang) return;
        static const int MAX_OVERHANG = Config::get("maximum_overhang");
        while (it != end && it->lrOverhang() > MAX_OVERHANG) ++it;
    }
    std::vector<OverlapRange>::const_iterator it, end;
    bool onlyNoOverhang;
};
class IterNoOverhang {
public:
    IterNoOverhang(const std::vector<OverlapRange>& ovlps) : ovlps(ovlps) {}
    OvlpIterator begin() { return OvlpIterator(ovlps.begin(), ovlps.end(), true); }
    OvlpIterator end() { return OvlpIterator(ovlps.end(), ovlps.end(), true); }
private:
    const std::vector<OverlapRange>& ovlps;
};
