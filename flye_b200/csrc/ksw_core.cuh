// flye_b200 — the banded affine-gap global alignment behind checkIdyAndTrim (alignment.cpp:102-216 -> minimap2's
// ksw_extz2_sse, lib/minimap2/ksw2_extz2_sse.c) as one scalar routine for host and device.
//
// The reference takes the CIGAR, and the CIGAR of a banded Suzuki-Kasahara alignment is not a function of the two strings alone:
// ksw2 keeps the differences u, v, x, y of one anti-diagonal in 8-bit arrays indexed by the target position, updates them in
// place sixteen at a time and never initialises the cells next to the band — a cell that enters the band reads what the 16-byte
// block update of an earlier anti-diagonal left at its position, and the scores of a diagonal are written in unaligned 16-byte
// strides that run past the band.  Those bytes decide ties in the traceback.  This routine keeps ksw2's byte arrays with ksw2's
// layout in one zero-filled block (u | v | x | y | s | target | reversed query | 16 spare bytes, what kcalloc(tlen_*6 + qlen_ + 1,
// 16) provides) and performs the same per-byte operations in the same order; the lanes of a 16-byte operation interact only
// through the one-byte carries x1_ / v1_, which are kept.  Path: SSE2 without SSE4.1 (what oracle/Makefile builds), with_cigar,
// KSW_EZ_APPROX_MAX | KSW_EZ_APPROX_DROP, zdrop = -1, left-aligned gaps, match 2, mismatch -4, gap open 4, gap extend 2, m = 5.
//
// Checked on the CPU: oracle/trim_check.cpp built with -DTRIM_CORE runs this header's host build and must print the CIGAR lines
// of the unmodified reference (tests/test_oracle_trim.py).  The device build is reached through fg_align_cigar_batch (ksw.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace fg {

struct KswSizes { size_t memBytes, pBytes, rounds, nCol; };

// workspace of one alignment with band w
__host__ __device__ inline KswSizes kswSizes(int qlen, int tlen, int w) {
    KswSizes z;
    const size_t tlen_ = ((size_t)tlen + 15) / 16, qlen_ = ((size_t)qlen + 15) / 16;
    int n = qlen < tlen ? qlen : tlen;
    n = ((n < w + 1 ? n : w + 1) + 15) / 16 + 1;                              // ksw2_extz2_sse.c:84-85
    z.nCol = (size_t)n * 16;
    z.rounds = (size_t)qlen + (size_t)tlen - 1;
    z.memBytes = tlen_ * 16 * 6 + qlen_ * 16 + 16;
    z.pBytes = z.rounds * z.nCol + 16;
    return z;
}

enum { KSW_OK = 0, KSW_BAND_TOO_NARROW = 1, KSW_CIGAR_OVERFLOW = 2 };

// One ksw_extz2_sse call + ksw_backtrack.  mem (memBytes), p (pBytes), off / offEnd (rounds ints each) are scratch; the CIGAR
// (len << 4 | op; 0 = M, 1 = I: query only, 2 = D: target only) is written front to back into cigar[0 .. *nCigar).
__host__ __device__ inline int kswExtz2Core(const uint8_t* query, int qlen, const uint8_t* target, int tlen, int w, uint8_t* mem, uint8_t* p,
                                            int* off, int* offEnd, uint32_t* cigar, int cigarCap, int* nCigar) {
    *nCigar = 0;
    if (qlen <= 0 || tlen <= 0) return KSW_OK;                                // :64 (empty CIGAR)
    const int q = 4, e = 2;
    const uint8_t qe2 = (uint8_t)((q + e) * 2), qv = (uint8_t)q;
    const uint8_t scMch = (uint8_t)2, scMis = (uint8_t)-4, scN = (uint8_t)-e;  // mat[0], mat[1]; mat[24] == 0 -> -e (:76)
    const uint8_t wild = 4, maxSc = (uint8_t)(2 + (q + e) * 2);               // :77-78
    if (w < 0) w = tlen > qlen ? tlen : qlen;
    const KswSizes z = kswSizes(qlen, tlen, w);
    const size_t T16 = (((size_t)tlen + 15) / 16) * 16;
    for (size_t i = 0; i < z.memBytes; ++i) mem[i] = 0;                        // kcalloc
    uint8_t* u = mem; uint8_t* v = u + T16; uint8_t* x = v + T16; uint8_t* y = x + T16; uint8_t* s = y + T16;
    uint8_t* sf = s + T16; uint8_t* qr = sf + T16;
    for (int t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];               // :107
    for (int t = 0; t < tlen; ++t) sf[t] = target[t];
    const size_t nCol = z.nCol;
    int lastSt = -1, lastEn = -1;
    for (int r = 0; r < qlen + tlen - 1; ++r) {                               // :110
        int st = 0, en = tlen - 1;
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;                   // :119 (arithmetic shift)
        if (en > ((r + w) >> 1)) en = (r + w) >> 1;
        if (st > en) return KSW_BAND_TOO_NARROW;                              // :121-124 (ez->zdropped)
        const int st0 = st, en0 = en;
        st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;                      // :126
        uint8_t x1, v1;
        if (st > 0) {                                                         // :128-132
            if (st - 1 >= lastSt && st - 1 <= lastEn) { x1 = x[st - 1]; v1 = v[st - 1]; }
            else x1 = v1 = 0;
        } else { x1 = 0; v1 = r ? qv : 0; }
        if (en >= r) { y[r] = 0; u[r] = r ? qv : 0; }                         // :133
        const uint8_t* qrr = qr + (qlen - 1 - r);
        for (int t = st0; t <= en0; t += 16)                                   // :135-152, unaligned 16-byte strides
            for (int i = 0; i < 16; ++i) {
                const uint8_t sq = sf[t + i], sq2 = qrr[t + i];
                uint8_t sc = sq == sq2 ? scMch : scMis;
                if (sq == wild || sq2 == wild) sc = scN;
                s[t + i] = sc;
            }
        uint8_t* pr = p + (size_t)r * nCol - st;
        off[r] = st; offEnd[r] = en;
        uint8_t carryX = x1, carryV = v1;
        for (int t = st; t <= en; ++t) {                                      // :179-208 (SSE2 emulation branches)
            uint8_t zz = (uint8_t)(s[t] + qe2);
            const uint8_t xt1 = carryX; carryX = x[t];
            const uint8_t vt1 = carryV; carryV = v[t];
            uint8_t a = (uint8_t)(xt1 + vt1);
            const uint8_t ut = u[t];
            uint8_t b = (uint8_t)(y[t] + ut);
            uint8_t d = (int8_t)a > (int8_t)zz ? 1 : 0;
            zz = (int8_t)zz > 0 ? zz : 0;
            zz = zz > a ? zz : a;                                             // unsigned max
            if ((int8_t)b > (int8_t)zz) d = 2;
            zz = zz > b ? zz : b;
            zz = zz < maxSc ? zz : maxSc;
            u[t] = (uint8_t)(zz - vt1);
            v[t] = (uint8_t)(zz - ut);
            zz = (uint8_t)(zz - qv);
            a = (uint8_t)(a - zz);
            b = (uint8_t)(b - zz);
            if ((int8_t)a > 0) { x[t] = a; d |= 0x08; } else x[t] = 0;
            if ((int8_t)b > 0) { y[t] = b; d |= 0x10; } else y[t] = 0;
            pr[t] = d;
        }
        lastSt = st; lastEn = en;
    }
    // ksw_backtrack (ksw2.h:119-154: is_rot, !is_rev, no introns) from (tlen-1, qlen-1): written back to front into the tail of
    // cigar[], then moved to the front
    int i = tlen - 1, j = qlen - 1, state = 0, n = 0;
    bool overflow = false;
    auto push = [&](uint32_t op, int len) {
        if (n && op == (cigar[cigarCap - n] & 0xf)) { cigar[cigarCap - n] += (uint32_t)len << 4; return; }
        if (n == cigarCap) { overflow = true; return; }
        ++n;
        cigar[cigarCap - n] = (uint32_t)len << 4 | op;
    };
    while (i >= 0 && j >= 0 && !overflow) {
        int force = -1;
        const int r = i + j;
        if (i < off[r]) force = 2;
        if (i > offEnd[r]) force = 1;
        const uint32_t tmp = force < 0 ? p[(size_t)r * nCol + i - off[r]] : 0;
        if (state == 0) state = tmp & 7;
        else if (!(tmp >> (state + 2) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (force >= 0) state = force;
        if (state == 0) { push(0, 1); --i; --j; }
        else if (state == 1 || state == 3) { push(2, 1); --i; }
        else { push(1, 1); --j; }
    }
    if (!overflow && i >= 0) push(2, i + 1);
    if (!overflow && j >= 0) push(1, j + 1);
    if (overflow) return KSW_CIGAR_OVERFLOW;
    for (int k = 0; k < n; ++k) cigar[k] = cigar[cigarCap - n + k];          // the walk went from the end: this is already front to back
    *nCigar = n;
    return KSW_OK;
}

}  // namespace fg
