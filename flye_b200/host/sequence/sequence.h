// flye_b200 host mirror — DnaSequence with the reference's interface and EXACTLY its packed layout
// (src/sequence/sequence.h:15-181): 2 bits per base, 32 bases per 64-bit word, base j at bits 2*(j%32) of word j/32,
// A0 C1 G2 T3; the reverse complement is a view (index flip + ~code&3) sharing the buffer.  The packed words are
// what fg_reads_upload takes, so uploading a read set is a straight copy.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

class DnaSequence {
public:
    typedef size_t NuclType;

    DnaSequence() : _store(std::make_shared<Store>()), _rc(false) {}
    explicit DnaSequence(const std::string& text) : _store(std::make_shared<Store>()), _rc(false) {
        _store->length = text.size();
        _store->words.assign((text.size() + 31) / 32, 0);
        for (size_t i = 0; i < text.size(); ++i) _store->words[i >> 5] |= (uint64_t)dnaToId(text[i]) << ((i & 31) * 2);
    }

    size_t length() const { return _store->length; }
    NuclType atRaw(size_t i) const {
        const size_t j = _rc ? _store->length - 1 - i : i;
        const NuclType code = (_store->words[j >> 5] >> ((j & 31) * 2)) & 3;
        return _rc ? (~code & 3) : code;
    }
    char at(size_t i) const { return "ACGT"[atRaw(i)]; }
    DnaSequence complement() const {
        DnaSequence c(*this);
        c._rc = true;   // same quirk as the reference: a complement of a complement stays a complement view
        return c;
    }
    DnaSequence substr(size_t start, size_t length) const {
        if (length == 0) throw std::runtime_error("Zero length subtring");
        if (start >= _store->length) throw std::runtime_error("Incorrect substring start");
        if (start + length > _store->length) length = _store->length - start;
        DnaSequence out;
        out._store->length = length;
        out._store->words.assign((length + 31) / 32, 0);
        for (size_t i = 0; i < length; ++i) out._store->words[i >> 5] |= (uint64_t)atRaw(start + i) << ((i & 31) * 2);
        return out;
    }
    std::string str() const {
        std::string s(length(), 'A');
        for (size_t i = 0; i < s.size(); ++i) s[i] = at(i);
        return s;
    }
    static size_t dnaToId(char c) {
        switch (c) {
            case 'A': case 'a': return 0; case 'C': case 'c': return 1;
            case 'G': case 'g': return 2; case 'T': case 't': return 3;
            default: return (size_t)-1;
        }
    }
    static char idToDna(size_t id) { return "ACGT"[id]; }

    // mirror-only accessors used for the device upload
    const uint64_t* packedWords() const { return _store->words.data(); }
    size_t numWords() const { return _store->words.size(); }
    bool isComplementView() const { return _rc; }

private:
    struct Store {
        size_t length = 0;
        std::vector<uint64_t> words;
    };
    std::shared_ptr<Store> _store;
    bool _rc;
};
