// flye_b200 host mirror — Kmer / IterKmers / yieldMinimizers with the reference's interface (src/sequence/kmer.h).
// Host-side convenience for callers; the device computes the same values from the packed reads.
#pragma once
#include <deque>
#include <vector>

#include "sequence_container.h"
#include "../common/config.h"

static_assert(sizeof(size_t) == 8, "32-bit architectures are not supported");

class Kmer {
public:
    typedef size_t KmerRepr;
    explicit Kmer(KmerRepr repr = 0) : _repr(repr) {}
    Kmer(const DnaSequence& seq, size_t start, size_t length) : _repr(0) {
        if (length != Parameters::get().kmerSize) throw std::runtime_error("Kmer length inconsistency");
        for (size_t i = start; i < start + length; ++i) _repr = (_repr << 2) | seq.atRaw(i);   // first base most significant
    }
    Kmer reverseComplement() {
        KmerRepr in = _repr, out = 0;
        for (size_t i = 0; i < Parameters::get().kmerSize; ++i) { out = (out << 2) | (~in & 3); in >>= 2; }
        return Kmer(out);
    }
    bool standardForm() {   // true iff replaced by a strictly smaller reverse complement
        const Kmer rc = reverseComplement();
        if (rc._repr < _repr) { _repr = rc._repr; return true; }
        return false;
    }
    void appendRight(DnaSequence::NuclType nt) {
        const size_t k = Parameters::get().kmerSize;
        _repr = ((_repr << 2) | nt) & (k >= 32 ? ~(KmerRepr)0 : (((KmerRepr)1 << (2 * k)) - 1));
    }
    void appendLeft(DnaSequence::NuclType nt) { _repr = (_repr >> 2) | (nt << (2 * Parameters::get().kmerSize - 2)); }
    bool operator==(const Kmer& o) const { return _repr == o._repr; }
    bool operator!=(const Kmer& o) const { return _repr != o._repr; }
    bool operator<(const Kmer& o) { return _repr < o._repr; }
    size_t hash() const {
        size_t x = _repr, z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    size_t numRepr() { return _repr; }
private:
    KmerRepr _repr;
};
namespace std { template <> struct hash<Kmer> { size_t operator()(const Kmer& k) const { return k.hash(); } }; }

struct KmerPosition {
    KmerPosition(Kmer kmer, int32_t position) : kmer(kmer), position(position) {}
    Kmer kmer;
    int32_t position;
};

class KmerIterator {
public:
    typedef std::forward_iterator_tag iterator_category;
    KmerIterator(const DnaSequence* seq, size_t position) : _seq(seq), _position(position) {
        if (position != seq->length() - Parameters::get().kmerSize) _kmer = Kmer(*seq, 0, Parameters::get().kmerSize);
    }
    bool operator==(const KmerIterator& o) const { return _seq == o._seq && _position == o._position; }
    bool operator!=(const KmerIterator& o) const { return !(*this == o); }
    KmerPosition operator*() const { return KmerPosition(_kmer, (int32_t)_position); }
    KmerIterator& operator++() {
        _kmer.appendRight(_seq->atRaw(_position + Parameters::get().kmerSize));
        ++_position;
        return *this;
    }
private:
    const DnaSequence* _seq;
    size_t _position;
    Kmer _kmer;
};

// positions 0 .. L-k-1: the end sentinel is L-k, so the last k-mer of a sequence is never visited (kmer.h:185-198)
class IterKmers {
public:
    IterKmers(const DnaSequence& sequence, size_t start = 0, size_t length = std::string::npos)
        : _sequence(sequence), _start(start), _length(length) {}
    KmerIterator begin() {
        if (_sequence.length() < Parameters::get().kmerSize + _start) return end();
        return KmerIterator(&_sequence, _start);
    }
    KmerIterator end() {
        const size_t e = _length == std::string::npos ? _sequence.length() : _length + _start;
        return KmerIterator(&_sequence, e - Parameters::get().kmerSize);
    }
private:
    const DnaSequence& _sequence;
    const size_t _start, _length;
};

inline std::vector<KmerPosition> yieldMinimizers(const DnaSequence& sequence, int window) {   // kmer.h:206-262
    if (window < 1) throw std::runtime_error("wrong minimizer length");
    std::vector<KmerPosition> out;
    struct Entry { KmerPosition kp; size_t hash; };
    std::deque<Entry> q;
    for (auto kp : IterKmers(sequence)) {
        if (window == 1) { out.push_back(kp); continue; }
        Kmer canon = kp.kmer;
        canon.standardForm();
        const size_t h = canon.hash();
        while (!q.empty() && q.back().hash > h) q.pop_back();
        q.push_back({kp, h});
        if (q.front().kp.position <= kp.position - window) {
            while (q.front().kp.position <= kp.position - window) q.pop_front();
            while (q.size() >= 2 && q[0].hash == q[1].hash) q.pop_front();
        }
        if (out.empty() || out.back().position != q.front().kp.position) out.push_back(q.front().kp);
    }
    return out;
}
