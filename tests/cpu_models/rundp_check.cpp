// CPU model of the run-compressed chaining DP and of the run-wise chain walk (flye_b200/csrc/overlap.cu: pairPrepKernel /
// chainRunsKernel, chainRunDpKernel, chainFillKernel, chainWalkKernel<true>) checked against the literal algorithm of the
// reference (src/sequence/overlap.cpp:277-323 and :338-383) on random match lists: diagonal runs, indels, noise, ties.
// Usage: rundp_check [trials] [seed]  -> "OK <pairs> <matches> heads=<..> " or the first mismatch.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

struct Match { int32_t cur, ext; };
struct Params { int k, maxJump; bool extSorted; };

// ---- literal reference -----------------------------------------------------------------------------------------------
static void literalDp(const std::vector<Match>& m, const Params& P, std::vector<int32_t>& score, std::vector<int32_t>& back) {
    const int n = (int)m.size();
    score.assign(n, 0); back.assign(n, -1);
    for (int i = 1; i < n; ++i) {
        int32_t maxScore = 0, maxId = 0;
        const int32_t curNext = m[i].cur, extNext = m[i].ext;
        for (int j = i - 1; j >= 0; --j) {
            const int32_t curPrev = m[j].cur, extPrev = m[j].ext;
            if (0 < curNext - curPrev && curNext - curPrev < P.maxJump && 0 < extNext - extPrev && extNext - extPrev < P.maxJump) {
                const int32_t matchScore = std::min(std::min(curNext - curPrev, extNext - extPrev), (int32_t)P.k);
                const int32_t jumpDiv = std::abs((curNext - curPrev) - (extNext - extPrev));
                const int32_t gapCost = (int32_t)((jumpDiv > 100 ? 2.0f : 0.5f) * jumpDiv);
                const int32_t nextScore = score[j] + matchScore - gapCost;
                if (nextScore > maxScore) {
                    maxScore = nextScore; maxId = j;
                    if (jumpDiv == 0 && curNext - curPrev < P.k) break;
                }
            }
            if (P.extSorted && extNext - extPrev > P.maxJump) break;
            if (!P.extSorted && curNext - curPrev > P.maxJump) break;
        }
        score[i] = std::max(maxScore, (int32_t)P.k);
        if (maxScore > P.k) back[i] = maxId;
    }
}
struct Chain { int start, first, length; bool operator==(const Chain& o) const { return start == o.start && first == o.first && length == o.length; } };
static std::vector<Chain> literalWalk(const std::vector<int32_t>& order, std::vector<int32_t> back) {
    std::vector<Chain> out;
    for (int32_t cs : order) {
        if (back[cs] == -1) continue;
        int first = 0, len = 0, pos = cs;
        while (pos != -1) { first = pos; ++len; const int np = back[pos]; back[pos] = -1; pos = np; }
        out.push_back({cs, first, len});
    }
    return out;
}

// ---- model of the kernels --------------------------------------------------------------------------------------------
struct Run { int32_t a, curA, extA, backA, backRun, b, curB, extB, scoreB; };
static std::vector<Run> findRuns(const std::vector<Match>& m, int k, std::vector<int32_t>& runOf) {
    std::vector<Run> runs;
    const int n = (int)m.size();
    runOf.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        const int32_t dc = i ? m[i].cur - m[i - 1].cur : 0, de = i ? m[i].ext - m[i - 1].ext : 0;
        const bool head = i <= 1 || !(dc == de && 0 < dc && dc < k);
        if (head) {
            if (i > 0) { runs.back().b = i - 1; runs.back().curB = m[i - 1].cur; runs.back().extB = m[i - 1].ext; }
            Run r{}; r.a = i; r.curA = m[i].cur; r.extA = m[i].ext; r.backA = -1; r.backRun = 0;
            runs.push_back(r);
        }
        runOf[i] = (int)runs.size() - 1;
    }
    if (n) { runs.back().b = n - 1; runs.back().curB = m[n - 1].cur; runs.back().extB = m[n - 1].ext; }
    return runs;
}
static long g_heads = 0, g_walkBack = 0, g_runEvals = 0;
static void runDp(const std::vector<Match>& m, const Params& P, std::vector<Run>& runs) {
    const int R = (int)runs.size();
    if (!R) return;
    runs[0].scoreB = 0; runs[0].backA = -1;
    for (int r = 1; r < R; ++r) {
        ++g_heads;
        const int32_t curN = runs[r].curA, extN = runs[r].extA, sortedN = P.extSorted ? extN : curN;
        int32_t best = 0, bestId = 0, bestRun = 0;
        for (int q = r - 1; q >= 0; --q) {   // lanes of the kernel look at 16 runs per step; the order of decisions is this one
            ++g_runEvals;
            int32_t jB = runs[q].b, cj = runs[q].curB, ej = runs[q].extB, sj = runs[q].scoreB;
            const int32_t sortedB = P.extSorted ? ej : cj;
            if (sortedN - sortedB > P.maxJump) break;                      // second break rule: this run and all older ones
            bool has = true;
            if (!(cj < curN && ej < extN)) {
                const int32_t a = runs[q].a, cA = runs[q].curA, dg = cj - ej, X = std::min(curN, extN + dg);
                if (a == jB || !(cA < X)) has = false;
                else {
                    ++g_walkBack;
                    int32_t lo = a, hi = jB, cLo = cA;
                    while (hi - lo > 1) { const int32_t mid = (lo + hi) >> 1; if (m[mid].cur < X) { lo = mid; cLo = m[mid].cur; } else hi = mid; }
                    sj -= cj - cLo; cj = cLo; ej = cLo - dg; jB = lo;
                }
            }
            const int32_t dc = curN - cj, de = extN - ej;
            const bool ok = has && dc < P.maxJump && de < P.maxJump;
            if (!ok) continue;
            const int32_t jd = std::abs(dc - de), gap = jd > 100 ? 2 * jd : (jd >> 1);
            const int32_t s = sj + std::min(std::min(dc, de), (int32_t)P.k) - gap;
            if (s > best) {
                best = s; bestId = jB; bestRun = q;
                if (jd == 0 && dc < P.k) break;                             // first break rule
            }
        }
        runs[r].scoreB = std::max(best, (int32_t)P.k) + (runs[r].curB - curN);
        runs[r].backA = best > P.k ? bestId : -1; runs[r].backRun = bestRun;
    }
}
static void fill(const std::vector<Match>& m, const std::vector<Run>& runs, const std::vector<int32_t>& runOf, std::vector<int32_t>& score,
                 std::vector<int32_t>& back) {
    const int n = (int)m.size();
    score.assign(n, 0); back.assign(n, -1);
    for (int i = 0; i < n; ++i) {
        const Run& r = runs[runOf[i]];
        score[i] = r.scoreB - (r.curB - m[i].cur);
        back[i] = i == r.a ? r.backA : i - 1;
    }
}
static std::vector<Chain> runWalk(const std::vector<int32_t>& order, const std::vector<Run>& runs, const std::vector<int32_t>& runOf) {
    std::vector<int32_t> ce(runs.size());
    for (size_t r = 0; r < runs.size(); ++r) ce[r] = runs[r].a;
    std::vector<Chain> out;
    for (int32_t cs : order) {
        int r = runOf[cs];
        if (!(cs >= ce[r] && (cs > runs[r].a || runs[r].backA >= 0))) continue;
        int first = 0, len = 0, pos = cs;
        for (;;) {
            const int32_t a = runs[r].a, c = ce[r];
            if (pos < c) { first = pos; ++len; break; }
            ce[r] = pos + 1;
            if (c > a) { first = c - 1; len += pos - c + 2; break; }
            first = a; len += pos - a + 1;
            if (runs[r].backA < 0) break;
            pos = runs[r].backA; r = runs[r].backRun;
        }
        out.push_back({cs, first, len});
    }
    return out;
}

int main(int argc, char** argv) {
    const int trials = argc > 1 ? atoi(argv[1]) : 20000;
    std::mt19937_64 rng(argc > 2 ? atoi(argv[2]) : 7);
    auto U = [&](int lo, int hi) { return lo + (int)(rng() % (uint64_t)(hi - lo + 1)); };
    long pairs = 0, matches = 0, presorted = 0;
    for (int t = 0; t < trials; ++t) {
        Params P; P.k = (t % 3 == 0) ? 15 : 17; P.maxJump = (t % 5 == 0) ? 300 : 1500; P.extSorted = t & 1;
        // a main chain of diagonal runs separated by indels / gaps, plus off-diagonal noise, repeats and exact duplicates
        std::vector<Match> m;
        int cur = U(0, 200), ext = U(0, 200);
        const int nRuns = U(1, 25), errStyle = t % 4;
        for (int r = 0; r < nRuns; ++r) {
            const int len = errStyle == 0 ? U(1, 3) : U(1, 40);
            for (int i = 0; i < len; ++i) { m.push_back({cur, ext}); const int step = U(1, errStyle == 1 ? 30 : P.k + 4); cur += step; ext += step; }
            const int kind = U(0, 5);
            if (kind == 0) cur += U(1, 5); else if (kind == 1) ext += U(1, 5); else if (kind == 2) { cur += U(1, 400); ext += U(1, 400); }
            else if (kind == 3) { cur += U(100, 2500); ext += U(100, 2500); } else if (kind == 4) ext -= std::min(ext, U(1, 60));
        }
        const int noise = U(0, 12);
        for (int i = 0; i < noise; ++i) m.push_back({U(0, cur + 50), U(0, ext + 50)});
        if (U(0, 3) == 0 && !m.empty()) for (int i = 0; i < 3; ++i) { Match x = m[rng() % m.size()]; if (U(0, 1)) x.ext += U(0, 40); m.push_back(x); }   // ties
        if (P.extSorted) std::stable_sort(m.begin(), m.end(), [](const Match& a, const Match& b) { return a.ext < b.ext; });
        else std::stable_sort(m.begin(), m.end(), [](const Match& a, const Match& b) { return a.cur < b.cur; });
        std::vector<int32_t> s0, b0, s1, b1, runOf;
        literalDp(m, P, s0, b0);
        std::vector<Run> runs = findRuns(m, P.k, runOf);
        runDp(m, P, runs);
        fill(m, runs, runOf, s1, b1);
        if (s0 != s1 || b0 != b1) {
            for (size_t i = 0; i < m.size(); ++i)
                if (s0[i] != s1[i] || b0[i] != b1[i]) {
                    printf("MISMATCH dp trial %d match %zu of %zu: literal (%d,%d) model (%d,%d) extSorted=%d\n", t, i, m.size(), s0[i], b0[i], s1[i], b1[i], (int)P.extSorted);
                    return 1;
                }
        }
        // the walk, in std::sort's score order
        std::vector<int32_t> order(m.size());
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return s0[a] > s0[b]; });
        if (!(literalWalk(order, b0) == runWalk(order, runs, runOf))) { printf("MISMATCH walk trial %d (%zu matches)\n", t, m.size()); return 1; }
        bool incr = true;
        for (size_t i = 1; i < m.size(); ++i) incr = incr && s0[i - 1] < s0[i];
        if (incr) {   // presorted pairs: the order n-1 .. 0 is std::sort's order
            ++presorted;
            for (size_t i = 0; i < m.size(); ++i) if (order[i] != (int32_t)(m.size() - 1 - i)) { printf("MISMATCH presorted order trial %d\n", t); return 1; }
        }
        ++pairs; matches += (long)m.size();
    }
    printf("OK %ld %ld heads=%ld runEvals=%ld walkBacks=%ld presorted=%ld\n", pairs, matches, g_heads, g_runEvals, g_walkBack, presorted);
    return 0;
}
