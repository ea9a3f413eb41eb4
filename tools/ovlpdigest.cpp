// TEST / BENCH INFRASTRUCTURE — not product code.
//
// Digest of an ordered overlap dump (the `.ovlp` text format of oracle/harness.cpp) that can be computed
//   (a) from the text file the reference harness wrote (CLI below), and
//   (b) from the binary records fg_overlaps_batch returned (shared-library entry, ctypes), on any partition of the
//       queries over ranks,
// and compared without ever holding both dumps on one machine:
//     D_q = sha256(block of query q: its "# id n" header line + its overlap lines, '\n' terminated)
//     C_j = sha256(D_{1000 j} || ... || D_{1000 j + 999})        (per-chunk digests localise a mismatch)
//     F   = sha256(C_0 || C_1 || ...)
// Equal F  <=>  the dumps are byte identical (up to sha256 collisions).
//
//   ovlpdigest FILE.ovlp            -> JSON {"queries":..,"overlaps":..,"final":"..","chunks":["..",..]}
//   libovlpdigest.so: fg_digest_queries(...) fills D_q for binary records; fg_digest_combine(...) folds D_q into C_j / F.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Sha256 {
    uint32_t h[8]; uint64_t len = 0; uint8_t buf[64]; size_t fill = 0;
    Sha256() { reset(); }
    void reset() {
        static const uint32_t init[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        memcpy(h, init, sizeof h); len = 0; fill = 0;
    }
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t* p) {
        static const uint32_t K[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
            0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
            0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
            0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
            0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
            0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
        uint32_t w[64];
        for (int i = 0; i < 16; ++i) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g), t1 = hh + S1 + ch + K[i] + w[i];
            const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c), t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const void* data, size_t n) {
        const uint8_t* p = (const uint8_t*)data; len += n;
        if (fill) { const size_t t = std::min(n, 64 - fill); memcpy(buf + fill, p, t); fill += t; p += t; n -= t; if (fill == 64) { block(buf); fill = 0; } }
        while (n >= 64) { block(p); p += 64; n -= 64; }
        if (n) { memcpy(buf, p, n); fill = n; }
    }
    void final(uint8_t out[32]) {
        const uint64_t bits = len * 8; uint8_t pad[72] = {0x80};
        const size_t padLen = (fill < 56 ? 56 - fill : 120 - fill);
        update(pad, padLen);
        uint8_t lb[8]; for (int i = 0; i < 8; ++i) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(lb, 8);
        for (int i = 0; i < 8; ++i) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
    }
};

std::string hex(const uint8_t d[32]) { char s[65]; for (int i = 0; i < 32; ++i) snprintf(s + 2 * i, 3, "%02x", d[i]); return std::string(s, 64); }

// the layout of fg_overlap (include/flye_b200.h), 72 bytes
struct Rec {
    uint32_t cur_id; int32_t cur_begin, cur_end, cur_len; uint32_t ext_id; int32_t ext_begin, ext_end, ext_len; int32_t score; float div;
    int32_t chain_length, filtered, edit, aln_len; uint64_t aln_first; uint32_t aln_count, reserved;
};
static_assert(sizeof(Rec) == 72, "fg_overlap layout");

const int CHUNK = 1000;

}  // namespace

extern "C" {

// D_q for queries [0, n): query q has id query_ids[q] and the records [offsets[q], offsets[q+1]) — formatted exactly as
// oracle/harness.cpp: dumpOverlap does.  out: n x 32 bytes.
void fg_digest_queries(const uint32_t* query_ids, const uint64_t* offsets, const void* records, uint64_t n, uint8_t* out) {
    const Rec* r = (const Rec*)records;
    char line[256];
    for (uint64_t q = 0; q < n; ++q) {
        Sha256 s;
        int m = snprintf(line, sizeof line, "# %u %zu\n", query_ids[q], (size_t)(offsets[q + 1] - offsets[q]));
        s.update(line, m);
        for (uint64_t i = offsets[q]; i < offsets[q + 1]; ++i) {
            uint32_t bits; memcpy(&bits, &r[i].div, 4);
            m = snprintf(line, sizeof line, "%u %d %d %d %u %d %d %d %d %08x\n", r[i].cur_id, r[i].cur_begin, r[i].cur_end, r[i].cur_len, r[i].ext_id,
                         r[i].ext_begin, r[i].ext_end, r[i].ext_len, r[i].score, bits);
            s.update(line, m);
        }
        s.final(out + 32 * q);
    }
}

// folds n per-query digests into chunk digests (ceil(n / 1000) x 32 bytes, may be NULL) and the final digest (32 bytes)
void fg_digest_combine(const uint8_t* dq, uint64_t n, uint8_t* chunks, uint8_t* fin) {
    Sha256 f;
    for (uint64_t a = 0; a < n; a += CHUNK) {
        Sha256 c; c.update(dq + 32 * a, 32 * std::min<uint64_t>(CHUNK, n - a));
        uint8_t d[32]; c.final(d);
        if (chunks) memcpy(chunks + 32 * (a / CHUNK), d, 32);
        f.update(d, 32);
    }
    f.final(fin);
}

}  // extern "C"

#ifndef OVLPDIGEST_LIB
int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: ovlpdigest FILE.ovlp\n"); return 1; }
    FILE* f = fopen(argv[1], "r");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    static char iobuf[1 << 22]; setvbuf(f, iobuf, _IOFBF, sizeof iobuf);
    std::vector<uint8_t> dq; Sha256 cur; bool open = false; uint64_t nOv = 0;
    char* line = nullptr; size_t cap = 0; ssize_t m;
    auto close = [&]() { if (open) { dq.resize(dq.size() + 32); cur.final(dq.data() + dq.size() - 32); cur.reset(); } };
    while ((m = getline(&line, &cap, f)) > 0) {
        if (line[0] == '#') { close(); open = true; } else if (line[0] != ' ') ++nOv;   // ("  aln ..." lines belong to the record before)
        cur.update(line, (size_t)m);
    }
    close();
    fclose(f);
    const uint64_t n = dq.size() / 32, nc = (n + CHUNK - 1) / CHUNK;
    std::vector<uint8_t> chunks(32 * std::max<uint64_t>(nc, 1)); uint8_t fin[32];
    fg_digest_combine(dq.data(), n, chunks.data(), fin);
    printf("{\"queries\": %llu, \"overlaps\": %llu, \"chunk_queries\": %d, \"final\": \"%s\", \"chunks\": [", (unsigned long long)n, (unsigned long long)nOv, CHUNK,
           hex(fin).c_str());
    for (uint64_t c = 0; c < nc; ++c) printf("%s\"%s\"", c ? ", " : "", hex(chunks.data() + 32 * c).c_str());
    printf("]}\n");
    return 0;
}
#endif
