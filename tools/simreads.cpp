// Deterministic long-read simulator for the benchmark / parity configurations
// (SURVEY.md §8(d)).  Not part of the product path and not part of the oracle:
// it only manufactures inputs.  Output: FASTA on disk (what the reference
// harness loads) — every letter is one of ACGT so that no loader ever has to
// substitute random bases (reference: sequence_container.cpp:318-328).
//
//   simreads --out reads.fasta --genome-len 4600000 --coverage 50 --mean-len 7500
//            --shape 2 --error 0.12 --seed 1 [--min-len 1500]
//            [--genome-fasta real.fa] [--meta N --meta-total 65000000]
//
// RNG: splitmix64-seeded xoshiro256**; same stream on every platform.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>

struct Rng {
    uint64_t s[4];
    static uint64_t sm64(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (auto& v : s) v = sm64(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0,1)
    uint64_t below(uint64_t n) { return (uint64_t)(uni() * (double)n); }
};

static const char NT[4] = {'A', 'C', 'G', 'T'};

static std::string randomGenome(Rng& rng, size_t len) {
    std::string g(len, 'A');
    for (size_t i = 0; i < len; i += 32) {
        uint64_t r = rng.next();
        for (size_t j = i; j < std::min(len, i + 32); ++j) { g[j] = NT[r & 3]; r >>= 2; }
    }
    return g;
}

static std::string loadFastaConcat(const char* path) {
    FILE* f = fopen(path, "r");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(1); }
    std::string g; char buf[1 << 16];
    while (fgets(buf, sizeof buf, f)) {
        if (buf[0] == '>') continue;
        for (char* p = buf; *p; ++p) {
            char c = *p & ~0x20;
            if (c == 'A' || c == 'C' || c == 'G' || c == 'T') g.push_back(c);
        }
    }
    fclose(f);
    return g;
}

static char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; default: return 'A'; } }

int main(int argc, char** argv) {
    std::string out = "reads.fasta", genomeFasta;
    size_t genomeLen = 1000000; double coverage = 30, meanLen = 7500, err = 0.12; int shape = 2;
    uint64_t seed = 1; long minLen = 1500; int meta = 0; size_t metaTotal = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto nxt = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(1); } return argv[++i]; };
        if (a == "--out") out = nxt();
        else if (a == "--genome-len") genomeLen = strtoull(nxt(), 0, 10);
        else if (a == "--coverage") coverage = atof(nxt());
        else if (a == "--mean-len") meanLen = atof(nxt());
        else if (a == "--shape") shape = atoi(nxt());
        else if (a == "--error") err = atof(nxt());
        else if (a == "--seed") seed = strtoull(nxt(), 0, 10);
        else if (a == "--min-len") minLen = atol(nxt());
        else if (a == "--genome-fasta") genomeFasta = nxt();
        else if (a == "--meta") meta = atoi(nxt());
        else if (a == "--meta-total") metaTotal = strtoull(nxt(), 0, 10);
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    Rng rng(seed);
    // genomes and their coverages
    std::vector<std::string> genomes; std::vector<double> covs;
    if (meta > 0) {
        // sizes log-uniform, rescaled to metaTotal; coverage_i = 4*2^(i/2.5), capped at 400
        std::vector<double> w(meta); double sum = 0;
        for (int i = 0; i < meta; ++i) { w[i] = std::exp(rng.uni() * std::log(10.0)); sum += w[i]; }
        for (int i = 0; i < meta; ++i) {
            size_t len = (size_t)(w[i] / sum * (double)metaTotal);
            genomes.push_back(randomGenome(rng, std::max<size_t>(len, 20000)));
            covs.push_back(std::min(400.0, 4.0 * std::pow(2.0, i / 2.5)));
        }
    } else {
        genomes.push_back(genomeFasta.empty() ? randomGenome(rng, genomeLen) : loadFastaConcat(genomeFasta.c_str()));
        covs.push_back(coverage);
    }
    FILE* fo = fopen(out.c_str(), "w");
    if (!fo) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
    static char iobuf[1 << 22]; setvbuf(fo, iobuf, _IOFBF, sizeof iobuf);
    const double pIns = err * 0.4, pDel = err * 0.4, pSub = err * 0.2;
    size_t nReads = 0, nBases = 0; std::string read;
    for (size_t gi = 0; gi < genomes.size(); ++gi) {
        const std::string& g = genomes[gi];
        double target = covs[gi] * (double)g.size(), done = 0;
        while (done < target) {
            double len = 0;
            for (int s = 0; s < shape; ++s) len += -std::log(1.0 - rng.uni());
            len *= meanLen / shape;
            size_t L = (size_t)std::max((double)minLen, len);
            L = std::min(L, g.size());
            size_t start = rng.below(g.size() - L + 1);
            bool rev = rng.next() & 1;
            read.clear();
            for (size_t i = 0; i < L; ++i) {
                char c = rev ? comp(g[start + L - 1 - i]) : g[start + i];
                double u = rng.uni();
                if (u < pDel) continue;
                if (u < pDel + pIns) { read.push_back(NT[rng.next() & 3]); read.push_back(c); continue; }
                if (u < pDel + pIns + pSub) { char d; do d = NT[rng.next() & 3]; while (d == c); read.push_back(d); continue; }
                read.push_back(c);
            }
            done += (double)L;
            if (read.empty()) continue;
            fprintf(fo, ">r%zu_g%zu_%zu_%c\n", nReads, gi, start, rev ? '-' : '+');
            fwrite(read.data(), 1, read.size(), fo); fputc('\n', fo);
            ++nReads; nBases += read.size();
        }
    }
    fclose(fo);
    fprintf(stderr, "simreads: %zu reads, %zu bases -> %s\n", nReads, nBases, out.c_str());
    return 0;
}
