// flye_b200 host mirror — FastaRecord / SequenceContainer with the reference's interface
// (src/sequence/sequence_container.h:14-270, .cpp:16-392).  Ids: a process-global counter hands out even ids to
// forward strands, id^1 is the reverse complement; global positions are prefix offsets over ALL ids in id order.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include <zlib.h>

#include "sequence.h"
#include "../common/logger.h"

struct FastaRecord {
    class Id {
    public:
        Id() : _id(std::numeric_limits<uint32_t>::max()) {}
        explicit Id(uint32_t id) : _id(id) {}
        bool operator==(const Id& o) const { return _id == o._id; }
        bool operator!=(const Id& o) const { return _id != o._id; }
        bool operator<(const Id& o) const { return _id < o._id; }
        Id rc() const { return Id(_id ^ 1u); }
        bool strand() const { return !(_id & 1u); }
        size_t hash() const {
            size_t x = _id, z = (x += 0x9E3779B97F4A7C15ULL);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            return z ^ (z >> 31);
        }
        int signedId() const { return (_id & 1u) ? -((int)_id + 1) / 2 : (int)_id / 2 + 1; }
        uint32_t rawId() const { return _id; }   // mirror-only
        friend std::ostream& operator<<(std::ostream& s, const Id& id) { return s << std::to_string(id._id); }
        friend std::istream& operator>>(std::istream& s, Id& id) { std::string b; s >> b; id._id = std::stoi(b); return s; }
        friend class SequenceContainer;
    private:
        uint32_t _id;
    };
    static const Id ID_NONE;
    typedef std::tuple<Id, Id> IdPair;

    FastaRecord() {}
    FastaRecord(const DnaSequence& s, const std::string& d, Id i) : id(i), sequence(s), description(d) {}
    Id id;
    DnaSequence sequence;
    std::string description;
};
inline const FastaRecord::Id FastaRecord::ID_NONE = FastaRecord::Id();

namespace std {
template <> struct hash<FastaRecord::Id> { size_t operator()(const FastaRecord::Id& h) const noexcept { return h.hash(); } };
template <> struct hash<FastaRecord::IdPair> {
    size_t operator()(const FastaRecord::IdPair& k) const {
        size_t l = std::get<0>(k).hash(), r = std::get<1>(k).hash();
        return l ^ (r + 0x9ddfea08eb382d69ULL + (l << 6) + (l >> 2));
    }
};
}  // namespace std

class SequenceContainer {
public:
    class ParseException : public std::runtime_error {
    public:
        ParseException(const std::string& what) : std::runtime_error(what) {}
    };
    typedef std::vector<FastaRecord> SequenceIndex;

    SequenceContainer() : _seqIdOffest(0), _offsetInitialized(false) {}

    void loadFromFile(const std::string& filename, int minReadLength = 0) {
        std::vector<FastaRecord> records;
        parse(filename, isFasta(filename), records);
        for (const auto& r : records)
            if (r.sequence.length() > (size_t)minReadLength) addRecord(r);
    }
    const FastaRecord& addSequence(const DnaSequence& sequence, const std::string& description) {
        Id fwd = addRecord(FastaRecord(sequence, description, FastaRecord::ID_NONE));
        return _seqIndex[fwd._id - _seqIdOffest];
    }
    static void writeFasta(const std::vector<FastaRecord>& records, const std::string& fileName, bool onlyPositiveStrand = false) {
        FILE* f = fopen(fileName.c_str(), "w");
        if (!f) throw std::runtime_error("Can't open " + fileName);
        for (const auto& rec : records) {
            if (onlyPositiveStrand && !rec.id.strand()) continue;
            const std::string name = onlyPositiveStrand ? rec.description.substr(1) : rec.description;
            fprintf(f, ">%s\n", name.c_str());
            const std::string s = rec.sequence.str();
            for (size_t c = 0; c < s.size(); c += 80) fprintf(f, "%s\n", s.substr(c, 80).c_str());
        }
        fclose(f);
    }
    static size_t getMaxSeqId() { return g_nextSeqId; }
    const SequenceIndex& iterSeqs() const { return _seqIndex; }
    const FastaRecord& getRecord(FastaRecord::Id id) const { return _seqIndex[id._id - _seqIdOffest]; }
    const DnaSequence& getSeq(FastaRecord::Id id) const { return _seqIndex[id._id - _seqIdOffest].sequence; }
    int32_t seqLen(FastaRecord::Id id) const { return (int32_t)_seqIndex[id._id - _seqIdOffest].sequence.length(); }
    std::string seqName(FastaRecord::Id id) const { return _seqIndex[id._id - _seqIdOffest].description; }
    const FastaRecord& recordByName(const std::string& name) const { return getRecord(_nameIndex.at(name)); }

    int computeNxStat(float fraction) const {
        std::vector<int32_t> lens;
        int64_t total = 0;
        for (const auto& r : _seqIndex) { lens.push_back((int32_t)r.sequence.length()); total += r.sequence.length(); }
        std::sort(lens.begin(), lens.end(), std::greater<int32_t>());
        int64_t cum = 0;
        for (int32_t l : lens) { cum += l; if (cum > fraction * total) return l; }
        return 0;
    }

    void buildPositionIndex() {
        _offsets.clear();
        size_t off = 0;
        for (const auto& r : _seqIndex) { _offsets.push_back(off); off += r.sequence.length(); }
        _offsets.push_back(off);
        if (off >= MAX_SEQUENCE) {
            Logger::get().error() << "Maximum sequence limit reached (" << MAX_SEQUENCE / 2 << ")";
            throw std::runtime_error("Input overflow");
        }
    }
    size_t globalPosition(FastaRecord::Id id, int32_t position) const { return _offsets[id._id - _seqIdOffest] + position; }
    void seqPosition(size_t globPos, FastaRecord::Id& outSeqId, int32_t& outPosition, int32_t& outLen) const {
        const size_t i = std::upper_bound(_offsets.begin(), _offsets.end(), globPos) - _offsets.begin() - 1;
        outSeqId = FastaRecord::Id((uint32_t)(_seqIdOffest + i));
        outPosition = (int32_t)(globPos - _offsets[i]);
        outLen = (int32_t)_seqIndex[i].sequence.length();
    }
    size_t idOffset() const { return _seqIdOffest; }   // mirror-only: first id of this container

    static size_t g_nextSeqId;

private:
    typedef FastaRecord::Id Id;

    Id addRecord(const FastaRecord& rec) {   // sequence_container.cpp:48-79
        if (!_offsetInitialized) { _offsetInitialized = true; _seqIdOffest = g_nextSeqId; }
        if (_seqIndex.size() != g_nextSeqId - _seqIdOffest) throw std::runtime_error("something wrong with sequence ids!");
        const Id fwd((uint32_t)g_nextSeqId);
        g_nextSeqId += 2;
        _seqIndex.emplace_back(rec.sequence, "+" + rec.description, fwd);
        if (_nameIndex.count(_seqIndex.back().description))
            throw ParseException("The input contain reads with duplicated IDs. Make sure all reads have unique IDs and restart. "
                                 "The first problematic ID was: " + rec.description);
        _nameIndex[_seqIndex.back().description] = fwd;
        _seqIndex.emplace_back(rec.sequence.complement(), "-" + rec.description, fwd.rc());
        _nameIndex[_seqIndex.back().description] = fwd.rc();
        return fwd;
    }

    static bool isFasta(const std::string& fileName) {
        std::string base = fileName;
        if (base.size() > 3 && base.substr(base.size() - 3) == ".gz") base.resize(base.size() - 3);
        const size_t dot = base.rfind('.');
        if (dot == std::string::npos) throw ParseException("Can't identify input file type");
        const std::string suffix = base.substr(dot + 1);
        if (suffix == "fasta" || suffix == "fa") return true;
        if (suffix == "fastq" || suffix == "fq") return false;
        throw ParseException("Can't identify input file type");
    }
    static void cleanHeader(std::string& h) {   // text between the marker and the first blank
        size_t end = 1;
        while (end < h.size() && !std::isspace((unsigned char)h[end])) ++end;
        h = h.substr(1, end - 1);
        if (h.empty()) throw ParseException("empty header");
    }
    // validateSequence (sequence_container.cpp:318-328) means to replace non-ACGT letters by rand() bases, but compares the
    // size_t code with `-1U` (32 bits), which never matches where size_t has 64: the letter stays, rand() is NOT called (its sequence
    // feeds estimateOverlaperParameters later), and DnaSequence's constructor ORs the all-ones code into the chunk — the letter
    // and every later base of its 32-base chunk read as T.  Kept bit for bit (tests/test_host_ingest.py compares with the reference).
    static void cleanSequence(std::string& s) {
        for (char& c : s)
            if (DnaSequence::dnaToId(c) == -1U) c = "ACGT"[rand() % 4];
    }
    // one line without its '\n' (a trailing '\r' stays: the callers strip it after the emptiness test, as the reference does)
    static bool nextLine(gzFile fd, std::vector<char>& buf, std::string& line) {
        line.clear();
        bool got = false;
        while (gzgets(fd, buf.data(), (int)buf.size())) {
            got = true;
            line += buf.data();
            if (!line.empty() && line.back() == '\n') { line.pop_back(); break; }
        }
        return got;
    }
    // readFasta / readFastq (sequence_container.cpp:145-303).  The line number of an error message counts the non-empty lines
    // consumed so far plus one, as the reference's does (blank lines are not counted; an error at the end of the file names the
    // line after the last one).
    static void parse(const std::string& fileName, bool fasta, std::vector<FastaRecord>& out) {
        gzFile fd = gzopen(fileName.c_str(), "rb");
        if (!fd) throw ParseException("Can't open reads file");
        std::vector<char> buf(1 << 20);
        std::string line, header, seq;
        int lineNo = 1, state = 0;
        try {
            while (nextLine(fd, buf, line)) {
                if (line.empty()) { if (!fasta) state = (state + 1) % 4; continue; }
                if (line.back() == '\r') line.pop_back();
                const char first = line.empty() ? '\0' : line[0];
                if (fasta) {
                    if (first == '>') {
                        if (!header.empty()) {
                            if (seq.empty()) throw ParseException("empty sequence");
                            out.emplace_back(DnaSequence(seq), header, FastaRecord::ID_NONE);
                            seq.clear();
                        }
                        header = line;
                        cleanHeader(header);
                    } else { cleanSequence(line); seq += line; }
                } else {
                    if (state == 0) { if (first != '@') throw ParseException("Fastq format error"); header = line; cleanHeader(header); }
                    else if (state == 1) { cleanSequence(line); out.emplace_back(DnaSequence(line), header, FastaRecord::ID_NONE); }
                    else if (state == 2 && first != '+') throw ParseException("Fastq fromat error");
                    state = (state + 1) % 4;
                }
                ++lineNo;
            }
            if (fasta) {
                if (seq.empty()) throw ParseException("empty sequence");
                if (header.empty()) throw ParseException("Fasta fromat error");
                out.emplace_back(DnaSequence(seq), header, FastaRecord::ID_NONE);
            }
        } catch (ParseException& e) {
            gzclose(fd);
            throw ParseException("parse error in " + fileName + " on line " + std::to_string(lineNo) + ": " + e.what());
        }
        gzclose(fd);
    }

    SequenceIndex _seqIndex;
    size_t _seqIdOffest;
    bool _offsetInitialized;
    std::unordered_map<std::string, Id> _nameIndex;
    static const size_t MAX_SEQUENCE = 1ULL << 40;
    std::vector<size_t> _offsets;
};
inline size_t SequenceContainer::g_nextSeqId = 0;
