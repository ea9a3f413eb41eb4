// CPU model of the threshold-bounded, band-pruned wavefront edit distance (flye_b200/csrc/editdist.cu: wfaSteps / wfaEditDistance /
// dropLimit) against the textbook O(nm) edit distance: for every limit, a distance below the limit must be reported exactly and
// a distance at or above it as ">= limit"; dropLimit(maxDiv, L) must be the smallest T with (float)T / L >= maxDiv.
// Usage: wfa_check [trials] [seed]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

static const int NEG = -(1 << 29);
static int bandLo(int s, int n, int target, int band) { return std::max(std::max(-s, -n), target - (band - 1 - s)); }
static int bandHi(int s, int m, int target, int band) { return std::min(std::min(s, m), target + (band - 1 - s)); }

// the kernel's recurrence, one diagonal at a time: fr[k] = furthest row i on diagonal k = j - i after s edits
static int wfaModel(const std::string& A, const std::string& B, int limit) {
    const int n = (int)A.size(), m = (int)B.size();
    if (limit <= 0) return limit;
    if (std::abs(n - m) >= limit) return limit;
    if (n == 0) return m;
    if (m == 0) return n;
    const int target = m - n, off = n + 3;
    std::vector<int> prev(n + m + 8, 12345), cur(n + m + 8, 54321);   // garbage: only the guards may make a diagonal unreachable
    int i0 = 0;
    while (i0 < n && i0 < m && A[i0] == B[i0]) ++i0;
    if (m == n && i0 >= n) return 0;
    prev[off] = i0;
    const int sMax = std::min(n + m, limit - 1), band = std::min(limit, n + m + 1);
    for (int s = 1; s <= sMax; ++s) {
        const int lo = bandLo(s, n, target, band), hi = bandHi(s, m, target, band);
        const int plo = bandLo(s - 1, n, target, band), phi = bandHi(s - 1, m, target, band);
        prev[plo - 1 + off] = prev[plo - 2 + off] = prev[phi + 1 + off] = prev[phi + 2 + off] = NEG;   // guards
        bool done = false;
        for (int k = lo; k <= hi; ++k) {
            const int a = prev[k - 1 + off], b = prev[k + off], c = prev[k + 1 + off];
            int best = -1;
            if (a >= 0 && a + k <= m) best = a;
            if (b >= 0 && b + 1 <= n && b + k + 1 <= m) best = std::max(best, b + 1);
            if (c >= 0 && c + 1 <= n) best = std::max(best, c + 1);
            if (best >= 0) while (best < n && best + k < m && A[best] == B[best + k]) ++best;
            cur[k + off] = best;
            if (k == target && best >= n) done = true;
        }
        if (done) return s;
        std::swap(prev, cur);
    }
    return sMax == n + m ? -1 : limit;
}
static int editDp(const std::string& A, const std::string& B) {
    const int n = (int)A.size(), m = (int)B.size();
    std::vector<int> row(m + 1), nr(m + 1);
    for (int j = 0; j <= m; ++j) row[j] = j;
    for (int i = 1; i <= n; ++i) {
        nr[0] = i;
        for (int j = 1; j <= m; ++j) nr[j] = std::min(std::min(row[j] + 1, nr[j - 1] + 1), row[j - 1] + (A[i - 1] != B[j - 1]));
        std::swap(row, nr);
    }
    return row[m];
}
static int dropLimit(float maxDiv, int alnLen) {
    if (alnLen <= 0) return 0x7fffffff;
    if (!(maxDiv > 0.f)) return 0;
    const float L = (float)alnLen;
    volatile float prod = maxDiv * L;
    long long T = (long long)std::ceil((float)prod);
    if (T > 0x7ffffff0LL) return 0x7fffffff;
    auto q = [&](long long t) { volatile float v = (float)t / L; return (float)v; };
    while (T > 0 && q(T - 1) >= maxDiv) --T;
    while (T < 0x7ffffff0LL && q(T) < maxDiv) ++T;
    return (int)T;
}

int main(int argc, char** argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 20000;
    std::mt19937_64 rng(argc > 2 ? atoi(argv[2]) : 11);
    long exact = 0, bounded = 0;
    for (int t = 0; t < trials; ++t) {
        const int n = (int)(rng() % 160), alpha = (t % 3 == 0) ? 2 : 4;
        std::string A(n, 'A'), B;
        for (auto& c : A) c = "ACGT"[rng() % alpha];
        const double err = (t % 4 == 0) ? 0.3 : (t % 4 == 1) ? 0.02 : 0.1;
        for (char c : A) {
            const double u = (rng() % 10000) / 10000.0;
            if (u < err / 3) continue;
            if (u < 2 * err / 3) { B.push_back("ACGT"[rng() % alpha]); B.push_back(c); continue; }
            B.push_back(u < err ? "ACGT"[rng() % alpha] : c);
        }
        if (t % 7 == 0) B = B.substr(0, B.size() / 2);
        const int d = editDp(A, B);
        for (int limit : {0x7fffffff, d + 1, d, d - 1, 1, 2, (int)(rng() % 40), d + 5, std::max(1, d / 2)}) {
            const int got = wfaModel(A, B, limit);
            if (d < limit) { if (got != d) { printf("MISMATCH exact: n=%zu m=%zu d=%d limit=%d got=%d\n", A.size(), B.size(), d, limit, got); return 1; } ++exact; }
            else { if (got < limit) { printf("MISMATCH bound: n=%zu m=%zu d=%d limit=%d got=%d\n", A.size(), B.size(), d, limit, got); return 1; } ++bounded; }
        }
    }
    for (int t = 0; t < trials * 5; ++t) {   // dropLimit: smallest T failing the reference's float test
        const int L = 1 + (int)(rng() % 40000);
        const float thr = (t % 5 == 0) ? 0.01f : (t % 5 == 1) ? 1.0f : (float)((rng() % 100000) / 100000.0 * 0.3);
        const int T = dropLimit(thr, L);
        auto fails = [&](long long x) { volatile float v = (float)x / (float)L; return !((float)v < thr); };
        if (!fails(T) || (T > 0 && fails(T - 1))) { printf("MISMATCH dropLimit thr=%g L=%d T=%d\n", thr, L, T); return 1; }
    }
    printf("OK exact=%ld bounded=%ld\n", exact, bounded);
    return 0;
}
