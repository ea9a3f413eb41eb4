// flye_b200 host mirror — VertexIndex with the reference's interface (src/sequence/vertex_index.h:66-294).
// Every method forwards to the C ABI (include/flye_b200.h); the k-mer counter, the per-read selection, the position
// lists and the repetitive set all live in HBM.  Errors come back as std::runtime_error with the reference's texts.
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
// the reference's vertex_index.h / overlap.h pull libcuckoo in, and its callers (extender.h, chimera.h, main_*.cpp) rely on
// getting it — and <thread>, <cassert> — through these headers
#if __has_include(<cuckoohash_map.hh>)
#include <cuckoohash_map.hh>
#endif
#include <memory>
#include <stdexcept>
#include <vector>

#include "kmer.h"
#include "sequence_container.h"
#include "../common/config.h"
#include "../common/logger.h"
#include "flye_b200.h"   // the C ABI: put <flye_b200 repo>/include on the include path (-I)

typedef std::map<size_t, size_t> KmerDistribution;

namespace flye_b200 {
// one device context per VertexIndex (reads + counts + index of ONE container), shared with the OverlapDetector /
// OverlapContainer built on top of it; the CUDA device is taken from FLYE_B200_DEVICE (default 0)
struct DeviceContext {
    fg_ctx* ctx = nullptr;
    const void* uploadedFrom = nullptr;   // the SequenceContainer whose reads are resident
    explicit DeviceContext(int device) {
        if (fg_ctx_create(device, &ctx) != FG_OK)
            throw std::runtime_error("flye_b200: no usable CUDA device (this build has no CPU path)");
    }
    ~DeviceContext() { fg_ctx_destroy(ctx); }
    void check(int rc) const { if (rc != FG_OK) throw std::runtime_error(fg_last_error(ctx)); }
    // forward strands only; the reverse complements are views on the device as they are on the host
    void uploadReads(const SequenceContainer& sc) {
        if (uploadedFrom == &sc) return;
        std::vector<uint64_t> words, offsets{0};
        std::vector<uint32_t> lengths;
        for (const auto& rec : sc.iterSeqs()) {
            if (!rec.id.strand()) continue;
            words.insert(words.end(), rec.sequence.packedWords(), rec.sequence.packedWords() + rec.sequence.numWords());
            offsets.push_back(words.size());
            lengths.push_back((uint32_t)rec.sequence.length());
        }
        words.push_back(0);
        check(fg_reads_upload(ctx, words.data(), offsets.data(), lengths.data(), (uint32_t)lengths.size()));
        uploadedFrom = &sc;
    }
    const void* queriesFrom = nullptr;    // the SequenceContainer uploaded as the second (query-only) set
    void uploadQueries(const SequenceContainer& sc) {
        if (queriesFrom == &sc) return;
        std::vector<uint64_t> words, offsets{0};
        std::vector<uint32_t> lengths;
        for (const auto& rec : sc.iterSeqs()) {
            if (!rec.id.strand()) continue;
            words.insert(words.end(), rec.sequence.packedWords(), rec.sequence.packedWords() + rec.sequence.numWords());
            offsets.push_back(words.size());
            lengths.push_back((uint32_t)rec.sequence.length());
        }
        words.push_back(0);
        check(fg_queries_upload(ctx, words.data(), offsets.data(), lengths.data(), (uint32_t)lengths.size()));
        queriesFrom = &sc;
    }
    static std::shared_ptr<DeviceContext> create() {
        const char* e = std::getenv("FLYE_B200_DEVICE");
        return std::make_shared<DeviceContext>(e ? std::atoi(e) : 0);
    }
};
}  // namespace flye_b200

class VertexIndex {
public:
    VertexIndex(const SequenceContainer& seqContainer, float sampleRate)
        : _seqContainer(seqContainer), _outputProgress(false), _sampleRate(sampleRate), _dev(flye_b200::DeviceContext::create()) {}
    ~VertexIndex() { this->clear(); }
    VertexIndex(const VertexIndex&) = delete;
    void operator=(const VertexIndex&) = delete;

    struct ReadPosition {
        ReadPosition(FastaRecord::Id readId = FastaRecord::ID_NONE, int32_t position = 0) : readId(readId), position(position) {}
        FastaRecord::Id readId;
        int32_t position;
    };
    // iterKmerPos returns a materialised list (the reference iterates its hash-table bucket in place)
    typedef std::vector<ReadPosition> IterHelper;

    void countKmers() {
        _dev->uploadReads(_seqContainer);
        uint64_t distinct = 0;
        _dev->check(fg_count_kmers(_dev->ctx, (int)Parameters::get().kmerSize, &distinct));
        uint64_t bins = 0;
        _dev->check(fg_kmer_hist(_dev->ctx, nullptr, nullptr, &bins));
        std::vector<uint64_t> f(bins), c(bins);
        if (bins) _dev->check(fg_kmer_hist(_dev->ctx, f.data(), c.data(), &bins));
        _hist.clear();
        for (uint64_t i = 0; i < bins; ++i) _hist[f[i]] = c[i];
    }
    void buildIndexUnevenCoverage(int minCoverage, float selectRate, int tandemFreq) {
        if (_outputProgress) Logger::get().info() << "Filling index table";
        fg_index_stats st;
        _dev->check(fg_build_index_solid(_dev->ctx, minCoverage, selectRate, tandemFreq, (float)Config::get("repeat_kmer_rate"),
                                         _sampleRate, &st));
        report(st);
    }
    void buildIndexMinimizers(int minCoverage, int wndLen) {
        if (_outputProgress) Logger::get().info() << "Building minimizer index";
        _dev->uploadReads(_seqContainer);
        fg_index_stats st;
        _dev->check(fg_build_index_minimizers(_dev->ctx, (int)Parameters::get().kmerSize, minCoverage, wndLen,
                                              (float)Config::get("repeat_kmer_rate"), &st));
        _sampleRate = st.sample_rate;
        report(st);
    }
    void clear() { if (_dev && _dev->ctx) fg_index_clear(_dev->ctx); }

    IterHelper iterKmerPos(Kmer kmer) const {
        uint64_t repr = kmer.numRepr(), first = 0; uint8_t rep = 0, rc = 0; uint32_t size = 0;
        _dev->check(fg_index_lookup(_dev->ctx, &repr, 1, &rep, &size, &first, &rc));
        IterHelper out;
        if (!size) return out;
        std::vector<uint32_t> ids(size); std::vector<int32_t> pos(size);
        _dev->check(fg_index_positions(_dev->ctx, first, size, rc, ids.data(), pos.data()));
        const uint32_t base = (uint32_t)_seqContainer.idOffset();
        for (uint32_t i = 0; i < size; ++i) out.emplace_back(FastaRecord::Id(base + ids[i]), pos[i]);
        return out;
    }
    bool isRepetitive(Kmer kmer) const {
        uint64_t repr = kmer.numRepr(); uint8_t rep = 0;
        _dev->check(fg_index_lookup(_dev->ctx, &repr, 1, &rep, nullptr, nullptr, nullptr));
        return rep != 0;
    }
    size_t kmerFreq(Kmer kmer) const {
        uint64_t repr = kmer.numRepr(); uint32_t size = 0;
        _dev->check(fg_index_lookup(_dev->ctx, &repr, 1, nullptr, &size, nullptr, nullptr));
        return size;
    }
    void outputProgress(bool set) { _outputProgress = set; }
    const KmerDistribution& getKmerHist() const { return _hist; }
    float getSampleRate() const { return _sampleRate; }

    // mirror-only
    const std::shared_ptr<flye_b200::DeviceContext>& device() const { return _dev; }
    const SequenceContainer& container() const { return _seqContainer; }

private:
    void report(const fg_index_stats& st) {
        Logger::get().debug() << "Mean k-mer frequency: " << st.mean_frequency;
        Logger::get().debug() << "Repetitive k-mer frequency: " << st.repetitive_frequency;
        Logger::get().debug() << "Selected k-mers: " << st.n_keys;
        Logger::get().debug() << "Index size: " << st.n_entries;
    }
    const SequenceContainer& _seqContainer;
    bool _outputProgress;
    float _sampleRate;
    KmerDistribution _hist;
    std::shared_ptr<flye_b200::DeviceContext> _dev;
};
