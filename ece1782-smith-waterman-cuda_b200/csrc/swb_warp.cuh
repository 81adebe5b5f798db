// The warp program of the Smith-Waterman scan: one warp scores one tile (32/G pairs of DB sequences,
// G lanes per pair) against the query rows staged in shared memory.
//
// Recurrence (reference: src/SWSolver.cu:246, src/cpu.cpp:45-72), linear gap g:
//     H(i,j) = max(0, H(i-1,j-1) + S(q_i, d_j), H(i,j-1) - g, H(i-1,j) - g),   score = max H
// restated so that the additions leave the ALU pipe (B200: VIADDMNMX / VIMNMX3 / PRMT share one 64-lane pipe, the
// packed add VIADD.16x2 issues on another one at the same rate -- swb_microbench kinds 0..3, 11..13):
//     d(i,j) = H(i-1,j-1) + S                                      vadd2      (off the chain)
//     H(i,j) = max(0, d(i,j), H(i-1,j) - g, H(i,j-1) - g)          vimax3.relu
//     H(i,j) - g                                                   vadd2      (kept per row; next cell's up-term)
//     score  = max(0, max d): a maximal H is never the end of a gap, so the running maximum takes d -- which also keeps
//     ptxas from folding the first vadd2 back into a VIADDMNMX.
// (SWB_V16_FORM 0 is the round-1 cell: c = viaddmax.relu(diag, S, left); H = viaddmax(H_up, -g, c); left' = vadd2:
//  one fused op on the chain, 3.5 ALU-pipe instructions per cell pair. Measured on one B200, 20 reference queries:
//  9,062 -> 9,395 GCUPS with the new cell, profiles/r2x_*.)
//
// Three arithmetic policies run the same program:
//   V16   two DB sequences in the halves of a 32-bit word, signed s16x2 DPX instructions
//         (prmt, vimax3.relu, 1/2 vimax3 = 2.5 ALU-pipe instructions per cell pair, 2 vadd2 beside them)
//   V32   two int32 lanes; exact for any score; re-scores flagged one-lane tiles (and everything when V16R cannot run).
//   V16R  packed s16x2 RELATIVE to a base that moves with the columns ("rebased"): exact for any score at 3.5
//         ALU-pipe instructions per cell pair; lane-group tiles only. Re-scores flagged lane-group tiles and scores
//         long-against-long tiles directly (their true scores pass 32767 anyway).
//   V16A / V32A  the affine-gap versions of V16 / V32 (two values per element: H and F along the chain, H and E per row)
// (Two more policies were measured slower on B200 and removed, see DESIGN.md: one that moved the additions to the FMA
//  pipe as IMADs in a biased domain -- register-file operand bandwidth -- and one that packed two QUERIES into the
//  halves of a word -- 4 B of shared-memory profile per packed cell.)
//
// Work split: a lane keeps K consecutive query rows in registers ("strip") and walks along the DB
// columns. G lanes of a group hold G consecutive strips and run a wavefront: lane g works on column
// t-g at step t and hands its bottom H and the residue pair to lane g+1 by __shfl_up_sync. Rows beyond
// K*G are covered by further passes ("super-strips"); the row between two passes goes through the
// boundary scratch in global memory. G = 1 is the pure inter-task case (no shuffles), G = 32 the
// intra-task warp wavefront for long sequences. K is chosen per group size and query on the host.
//
// The same source is compiled for the device (DevBackend, swb_kernels.cu) and, for CPU validation of
// the indexing / wavefront logic, for the host with a fiber-per-lane backend (tests/emu). The host
// build is test infrastructure only and is never linked into libswb.so.
#pragma once
#include <cuda_runtime.h>
#include "swb_types.h"

#ifndef SWB_SPLIT_HYST
// columns a pipelined pass lets its predecessor get ahead once it has had to wait. Measured on configs[3]: 0 / 16 / 32 /
// 64 / 256 / 1024 -> 5,046 / 4,945 / 5,044 / 4,972 / 4,607 / 3,910 GCUPS: any margin only lengthens the fill of a tile's
// pipeline, so a pass resumes as soon as the columns it needs are there.
#define SWB_SPLIT_HYST 0u
#endif
#ifndef SWB_V16_FORM
// the cell of the V16 policy: 0 = two viaddmax + one vadd2 (one fused op on the chain down a column),
// 1 = one vimax3.relu + two vadd2 (the additions issue beside the ALU pipe; two ops on the chain)
#define SWB_V16_FORM 1
#endif
#ifndef SWB_PF_CHUNKS
#define SWB_PF_CHUNKS 6u  // chunks (of 4 columns) that the L2 prefetch of the one-lane tiles runs ahead
#endif
#ifndef SWB_BLOCK_CHUNKS
#define SWB_BLOCK_CHUNKS 16u  // one-lane tiles: chunks per column block that the passes of a group share (see swb_run_tile)
#endif

// ---------------------------------------------------------------------------------------------
// byte permute with sign replication (PTX prmt.b32 generic mode: selector nibble bit 3 = replicate
// the sign of the selected byte)
SWB_HD uint32_t swb_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
#else
    uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t n = (sel >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(src >> (8 * (n & 7))) & 0xff;
        if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        d |= byte << (8 * i);
    }
    return d;
#endif
}

// ---------------------------------------------------------------------------------------------
// shared pieces of the packed-s16 policy
struct V16Base {
    typedef uint32_t T;
    static const bool is16 = true;
    static SWB_HD T splat(int v) { uint32_t u = (uint32_t)v & 0xffffu; return u | (u << 16); }
    static SWB_HD T max2(T a, T b)
    {
#ifdef __CUDA_ARCH__
        return __vmaxs2(a, b);
#else
        return __vimax3_s16x2(a, b, b);
#endif
    }
    static SWB_HD T add(T a, T b)
    {
#ifdef __CUDA_ARCH__
        return __vadd2(a, b);
#else
        return ((a + b) & 0xffffu) | ((((a >> 16) + (b >> 16)) & 0xffffu) << 16);
#endif
    }
    static SWB_HD int lo(T v) { return (int)(int16_t)(v & 0xffffu); }
    static SWB_HD int hi(T v) { return (int)(int16_t)(v >> 16); }
    // LDW bytes of a profile row (4 rows per word): one LDS.32 / LDS.64 / LDS.128; the wide forms need rows that start
    // on LDW-byte boundaries (the SPLIT kernels stage theirs that way)
    template <int LDW> static SWB_HD void ld_prof(const uint32_t *p, uint32_t (&w)[LDW / 4])
    {
        if constexpr (LDW == 16) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else if constexpr (LDW == 8) {
            const uint2 v = *reinterpret_cast<const uint2 *>(p);
            w[0] = v.x; w[1] = v.y;
        } else {
            w[0] = *p;
        }
    }
    // profile byte I (0..3) of the A word and of the B word, sign-extended into one s16x2
    template <int I> static SWB_HD T pair(uint32_t wa, uint32_t wb) { return swb_prmt(wa, wb, 0xC480u + 0x1111u * I); }
    template <class BE> static SWB_HD T shfl_up(BE &be, T v, int d, int w) { return be.shfl_up(v, d, w); }
    template <class BE> static SWB_HD T shfl_xor(BE &be, T v, int m, int w) { return be.shfl_xor(v, m, w); }
    template <class BE> static SWB_HD T ld(BE &be, const T *p) { return be.ld_cg(p); }
    template <class BE> static SWB_HD void st(BE &be, T *p, T v) { be.st_cg(p, v); }
    template <class BE> static SWB_HD void ld4(BE &be, const T *p, T *o)
    {
        uint4 v = be.ld_cg4(reinterpret_cast<const uint4 *>(p));
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    template <class BE> static SWB_HD void st4(BE &be, T *p, const T *o)
    {
        be.st_cg4(reinterpret_cast<uint4 *>(p), make_uint4(o[0], o[1], o[2], o[3]));
    }
};

// ---------------------------------------------------------------------------------------------
// V16: plain signed domain. Stored per row: H(k, j-1) - g. Profile entry: S + g.
//   d = vadd2(diag-g, S+g) ; h = vimax3.relu(d, up-g, left-g) ; left' = vadd2(h, -g) ; best = max(best, d)
//   (SWB_V16_FORM 0: c = viaddmax.relu(diag-g, S+g, left-g) ; h = viaddmax(h, -g, c) ; left' = vadd2(h, -g))
struct V16 : V16Base {
    static const bool rebased = false;
    static const bool blocked = true;  // one-lane tiles walk in pass groups over column blocks (swb_run_tile)
    struct C { T negg; };
    static SWB_HD C consts(const SwbScoreParams &p) { C c; c.negg = splat(-p.gap); return c; }
    static SWB_HD T hzero(const C &) { return 0u; }
    static SWB_HD T lzero(const C &c) { return c.negg; }
    static SWB_HD int score_lo(T best, const C &) { return lo(best); }
    static SWB_HD int score_hi(T best, const C &) { return hi(best); }
    // one DB column against the K rows of this lane; returns the bottom H
    template <int K, int LDW>
    static SWB_HD T column(T up, T &diag0, T (&left)[K], T &best, const C &cst, uint32_t codeA, uint32_t codeB,
                           const int8_t *prow, uint32_t sstride)
    {
        const uint32_t *ra = reinterpret_cast<const uint32_t *>(prow + codeA * sstride);
        const uint32_t *rb = reinterpret_cast<const uint32_t *>(prow + codeB * sstride);
        T h = up;
        T dg = diag0;
        diag0 = add(up, cst.negg);
#if SWB_V16_FORM == 1
        T hg = diag0;  // H(row above, this column) - g
#endif
#pragma unroll
        for (int kw = 0; kw < K / 4; kw += LDW / 4) {
            uint32_t wva[LDW / 4], wvb[LDW / 4];
            ld_prof<LDW>(ra + kw, wva);
            ld_prof<LDW>(rb + kw, wvb);
#pragma unroll
            for (int j = 0; j < LDW / 4; ++j) {
                const int k4 = kw + j;
                const uint32_t wa = wva[j], wb = wvb[j];
                T c[4];
#if SWB_V16_FORM == 1
#define SWB_CELL(I)                                                        \
    {                                                                      \
        const T s = pair<I>(wa, wb);                                       \
        const T d = c[I] = add(dg, s);                                     \
        dg = left[4 * k4 + I];                                             \
        h = __vimax3_s16x2_relu(d, hg, dg);                                \
        left[4 * k4 + I] = hg = add(h, cst.negg);                          \
    }
#else
#define SWB_CELL(I)                                                        \
    {                                                                      \
        const T s = pair<I>(wa, wb);                                       \
        c[I] = __viaddmax_s16x2_relu(dg, s, left[4 * k4 + I]);             \
        dg = left[4 * k4 + I];                                             \
        h = __viaddmax_s16x2(h, cst.negg, c[I]);                           \
        left[4 * k4 + I] = add(h, cst.negg);                               \
    }
#endif
                SWB_CELL(0) SWB_CELL(1)
                best = __vimax3_s16x2(best, c[0], c[1]);
                SWB_CELL(2) SWB_CELL(3)
                best = __vimax3_s16x2(best, c[2], c[3]);
#undef SWB_CELL
            }
        }
        return h;
    }
};

// ---------------------------------------------------------------------------------------------
// V16R: s16x2 values relative to a per-sequence base that follows the data ("rebased").
// Neighbouring cells of the H matrix differ by at most maxS + g (H(i,j) >= H(i,j-1) - g by the gap term, and
// H(i,j) <= H(i,j-1) + maxS + g by induction over the four terms; the same along i), so all values of one pass
// (R = K * G rows) over one block of CB columns lie within (R + CB) * (maxS + g) of any one of them. With that span below
// 2^15 the block is computed exactly in wrapping 16-bit arithmetic relative to base = H(bottom row of the group's
// first lane, column before the block), whatever the absolute scores are. At the first column of a block a lane
// subtracts the new base from its row state (the amount travels down the lane group one column behind the data), folds
// its running maximum into an int32 and starts a new relative one. The host picks CB (SwbScoreParams::rebase_shift).
// The floor of the recurrence is not 0 in this domain, so the relu form of V16 does not apply; instead
//     c = viaddmax(diag - g, S + g, left) ; h = viaddmax(h, -g, c) ; left' = viaddmax(h, -g, floor)
// with floor = -g - base (clamped to -32000: far below every live value once base is large): the stored row state
// max(h - g, -g) equals H - g for the true H = max(h, 0), h itself may run g below zero along the chain, which never
// wins a later max (c >= left' >= -g). One viaddmax replaces V16's vadd2: 4.5 ALU-pipe instructions per cell pair
// in that form; SWB_V16_FORM 1 (the default) takes the sum out as a vadd2 like V16 and uses the floored row state of
// the row above as the up-term: prmt, vimax3, viaddmax, 1/2 vimax3 = 3.5.
// The boundary row between passes keeps relative values plus a log of the writer's base per block (SwbScoreParams::
// blog); the reader adds (writer's base - its own base) with one vadd2 per column.
struct V16R : V16Base {
    static const bool rebased = true;
    static const bool blocked = false;
    static const bool is16 = false;  // exact: never flags, counts as a recompute pass
    struct C {
        T negg, fl, zrel;   // -g | floor of the row state (-g absolute) | zero (absolute) in the current base
        int g;
        int baseA, baseB;   // current base of the two sequences (absolute H values)
        int bestA, bestB;   // running maxima of the blocks behind
    };
    static SWB_HD int clampf(int v) { return v < -32000 ? -32000 : v; }
    static SWB_HD T pack(int a, int b) { return ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16); }
    static SWB_HD void set_base(C &c, int a, int b)
    {
        c.baseA = a;
        c.baseB = b;
        c.fl = pack(clampf(-c.g - a), clampf(-c.g - b));
        c.zrel = pack(clampf(-a), clampf(-b));
    }
    static SWB_HD C consts(const SwbScoreParams &p)
    {
        C c;
        c.g = p.gap;
        c.negg = splat(-p.gap);
        c.bestA = c.bestB = 0;
        set_base(c, 0, 0);
        return c;
    }
    static SWB_HD T hzero(const C &c) { return c.zrel; }
    static SWB_HD T lzero(const C &c) { return c.negg; }  // only used at the start of a pass, where the base is 0
    static SWB_HD int score_lo(T, const C &c) { return c.bestA; }
    static SWB_HD int score_hi(T, const C &c) { return c.bestB; }
    static SWB_HD T sub(T a, T b)
    {
#ifdef __CUDA_ARCH__
        return __vsub2(a, b);
#else
        return ((a - b) & 0xffffu) | ((((a >> 16) - (b >> 16)) & 0xffffu) << 16);
#endif
    }
    // the running relative maximum joins the absolute one; a new relative maximum starts
    static SWB_HD void fold(T &best, C &c)
    {
        const int a = c.baseA + lo(best), b = c.baseB + hi(best);
        c.bestA = a > c.bestA ? a : c.bestA;
        c.bestB = b > c.bestB ? b : c.bestB;
        best = 0x80008000u;
    }
    // saturating per-half subtraction
    static SWB_HD T subs(T a, T b)
    {
#ifdef __CUDA_ARCH__
        return __vsubss2(a, b);
#else
        int l = lo(a) - lo(b), h = hi(a) - hi(b);
        l = l < -32768 ? -32768 : (l > 32767 ? 32767 : l);
        h = h < -32768 ? -32768 : (h > 32767 ? 32767 : h);
        return pack(l, h);
#endif
    }
    // new base = old base + d (d: relative value of the reference cell): all row state moves by -d. State that sits at
    // the clamped floor (-32000: lanes running on the padding behind the last column, fed with "zero" from above while
    // the base is beyond 32000) must stay there: a wrapping subtraction would turn it into a huge positive value that
    // ends up in the running maximum. Saturate, then floor again in the new base.
    template <int K> static SWB_HD void rebase(T d, T &diag0, T (&left)[K], T &best, C &c)
    {
        fold(best, c);
        set_base(c, c.baseA + lo(d), c.baseB + hi(d));
#pragma unroll
        for (int k = 0; k < K; ++k) left[k] = max2(subs(left[k], d), c.fl);
        diag0 = max2(subs(diag0, d), c.fl);
    }
    template <int K, int LDW>
    static SWB_HD T column(T up, T &diag0, T (&left)[K], T &best, const C &cst, uint32_t codeA, uint32_t codeB,
                           const int8_t *prow, uint32_t sstride)
    {
        const uint32_t *ra = reinterpret_cast<const uint32_t *>(prow + codeA * sstride);
        const uint32_t *rb = reinterpret_cast<const uint32_t *>(prow + codeB * sstride);
        T h = up;
        T dg = diag0;
        diag0 = __viaddmax_s16x2(up, cst.negg, cst.fl);
#if SWB_V16_FORM >= 1
        T hg = diag0;
#endif
#pragma unroll
        for (int kw = 0; kw < K / 4; kw += LDW / 4) {
            uint32_t wva[LDW / 4], wvb[LDW / 4];
            ld_prof<LDW>(ra + kw, wva);
            ld_prof<LDW>(rb + kw, wvb);
#pragma unroll
            for (int j = 0; j < LDW / 4; ++j) {
                const int k4 = kw + j;
                const uint32_t wa = wva[j], wb = wvb[j];
                T c[4];
#if SWB_V16_FORM >= 1
                // up-term = the floored row state of the row above (flooring it changes nothing: left >= floor)
#define SWB_CELL(I)                                                        \
    {                                                                      \
        const T s = pair<I>(wa, wb);                                       \
        c[I] = add(dg, s);                                                 \
        dg = left[4 * k4 + I];                                             \
        h = __vimax3_s16x2(c[I], hg, dg);                                  \
        left[4 * k4 + I] = hg = __viaddmax_s16x2(h, cst.negg, cst.fl);     \
    }
#else
#define SWB_CELL(I)                                                        \
    {                                                                      \
        const T s = pair<I>(wa, wb);                                       \
        c[I] = __viaddmax_s16x2(dg, s, left[4 * k4 + I]);                  \
        dg = left[4 * k4 + I];                                             \
        h = __viaddmax_s16x2(h, cst.negg, c[I]);                           \
        left[4 * k4 + I] = __viaddmax_s16x2(h, cst.negg, cst.fl);          \
    }
#endif
                SWB_CELL(0) SWB_CELL(1)
                best = __vimax3_s16x2(best, c[0], c[1]);
                SWB_CELL(2) SWB_CELL(3)
                best = __vimax3_s16x2(best, c[2], c[3]);
#undef SWB_CELL
            }
        }
        return h;
    }
};

// ---------------------------------------------------------------------------------------------
// V32: the same two sequences on two int32 lanes (no wrap for any realistic input).
struct V32 {
    static const bool rebased = false;
    static const bool blocked = false;
    struct T { int a, b; };
    struct C { int g; };
    static const bool is16 = false;
    static SWB_HD T mk(int a, int b) { T t; t.a = a; t.b = b; return t; }
    static SWB_HD int mx(int a, int b) { return a > b ? a : b; }
    static SWB_HD C consts(const SwbScoreParams &p) { C c; c.g = p.gap; return c; }
    static SWB_HD T hzero(const C &) { return mk(0, 0); }
    static SWB_HD T lzero(const C &c) { return mk(-c.g, -c.g); }
    static SWB_HD int score_lo(T best, const C &) { return best.a; }
    static SWB_HD int score_hi(T best, const C &) { return best.b; }
    static SWB_HD T max2(T a, T b) { return mk(mx(a.a, b.a), mx(a.b, b.b)); }
    template <int K, int LDW>  // LDW (bytes per profile load) is a hint of the s16 policies: this one loads words
    static SWB_HD T column(T up, T &diag0, T (&left)[K], T &best, const C &cst, uint32_t codeA, uint32_t codeB,
                           const int8_t *prow, uint32_t sstride)
    {
        const uint32_t *ra = reinterpret_cast<const uint32_t *>(prow + codeA * sstride);
        const uint32_t *rb = reinterpret_cast<const uint32_t *>(prow + codeB * sstride);
        T h = up;
        T dg = diag0;
        diag0 = mk(up.a - cst.g, up.b - cst.g);
#pragma unroll
        for (int k4 = 0; k4 < K / 4; ++k4) {
            const uint32_t wa = ra[k4];
            const uint32_t wb = rb[k4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // profile entry = S + g
                const int sa = (int)(int8_t)(wa >> (8 * i));
                const int sb = (int)(int8_t)(wb >> (8 * i));
                const T l = left[4 * k4 + i];
                const T c = mk(mx(mx(dg.a + sa, l.a), 0), mx(mx(dg.b + sb, l.b), 0));
                dg = l;
                h = mk(mx(h.a - cst.g, c.a), mx(h.b - cst.g, c.b));
                left[4 * k4 + i] = mk(h.a - cst.g, h.b - cst.g);
                best = max2(best, c);
            }
        }
        return h;
    }
    template <class BE> static SWB_HD T shfl_up(BE &be, T v, int d, int w)
    {
        return mk((int)be.shfl_up((uint32_t)v.a, d, w), (int)be.shfl_up((uint32_t)v.b, d, w));
    }
    template <class BE> static SWB_HD T shfl_xor(BE &be, T v, int m, int w)
    {
        return mk((int)be.shfl_xor((uint32_t)v.a, m, w), (int)be.shfl_xor((uint32_t)v.b, m, w));
    }
    template <class BE> static SWB_HD T ld(BE &be, const T *p)
    {
        uint2 v = be.ld_cg2(reinterpret_cast<const uint2 *>(p));
        return mk((int)v.x, (int)v.y);
    }
    template <class BE> static SWB_HD void st(BE &be, T *p, T v)
    {
        be.st_cg2(reinterpret_cast<uint2 *>(p), make_uint2((uint32_t)v.a, (uint32_t)v.b));
    }
    template <class BE> static SWB_HD void ld4(BE &be, const T *p, T *o)
    {
        uint4 v = be.ld_cg4(reinterpret_cast<const uint4 *>(p));
        uint4 w = be.ld_cg4(reinterpret_cast<const uint4 *>(p) + 1);
        o[0] = mk((int)v.x, (int)v.y); o[1] = mk((int)v.z, (int)v.w);
        o[2] = mk((int)w.x, (int)w.y); o[3] = mk((int)w.z, (int)w.w);
    }
    template <class BE> static SWB_HD void st4(BE &be, T *p, const T *o)
    {
        be.st_cg4(reinterpret_cast<uint4 *>(p), make_uint4(o[0].a, o[0].b, o[1].a, o[1].b));
        be.st_cg4(reinterpret_cast<uint4 *>(p) + 1, make_uint4(o[2].a, o[2].b, o[3].a, o[3].b));
    }
};

// ---------------------------------------------------------------------------------------------
// Affine gaps (Gotoh; SURVEY 8f rank 4, hinted at by the reference's "define affine penalty ?", SWSolver.cu:8):
//     E(i,j) = max(E(i,j-1) - ge, H(i,j-1) - go)      F(i,j) = max(F(i-1,j) - ge, H(i-1,j) - go)
//     H(i,j) = max(0, H(i-1,j-1) + S, E(i,j), F(i,j))           a gap of length L costs go + (L-1) * ge
// The element type carries two values: along the chain and across lanes / passes (h = H, f = F); per row, kept from
// the previous column (h = H - go, f = E). Profile entry: S + go. With go == ge this is the linear recurrence.
// V16A: two DB sequences per word (s16x2, 6.5 ALU-pipe instructions per cell pair); V32A: exact int32 recompute.
struct V16A {
    static const bool rebased = false;
    static const bool blocked = false;
    static const bool is16 = true;
    struct T { uint32_t h, f; };
    struct C { uint32_t neg_go, neg_ge; };
    static SWB_HD T mk(uint32_t h, uint32_t f) { T t; t.h = h; t.f = f; return t; }
    static SWB_HD C consts(const SwbScoreParams &p)
    {
        C c;
        c.neg_go = V16Base::splat(-p.gap_open);
        c.neg_ge = V16Base::splat(-p.gap_extend);
        return c;
    }
    static SWB_HD T hzero(const C &c) { return mk(0u, c.neg_go); }       // H = 0, F = "none" (anything <= -go)
    static SWB_HD T lzero(const C &c) { return mk(c.neg_go, c.neg_go); }  // H - go with H = 0, E = "none"
    static SWB_HD int score_lo(T best, const C &) { return V16Base::lo(best.h); }
    static SWB_HD int score_hi(T best, const C &) { return V16Base::hi(best.h); }
    static SWB_HD T max2(T a, T b) { return mk(V16Base::max2(a.h, b.h), a.f); }
    template <int K, int LDW>  // LDW (bytes per profile load) is a hint of the s16 policies: this one loads words
    static SWB_HD T column(T up, T &diag0, T (&left)[K], T &best, const C &cst, uint32_t codeA, uint32_t codeB,
                           const int8_t *prow, uint32_t sstride)
    {
        const uint32_t *ra = reinterpret_cast<const uint32_t *>(prow + codeA * sstride);
        const uint32_t *rb = reinterpret_cast<const uint32_t *>(prow + codeB * sstride);
        uint32_t hgo = V16::add(up.h, cst.neg_go);  // H(row above, j) - go
        uint32_t f = up.f;
        uint32_t dgo = diag0.h;
        uint32_t h = up.h;
        diag0.h = hgo;
#pragma unroll
        for (int k4 = 0; k4 < K / 4; ++k4) {
            const uint32_t wa = ra[k4];
            const uint32_t wb = rb[k4];
            uint32_t hh[4];
#if SWB_V16_FORM >= 1
#define SWB_CELL(I)                                                                          \
    {                                                                                        \
        const uint32_t s = V16Base::pair<I>(wa, wb);                                         \
        const uint32_t e = __viaddmax_s16x2(left[4 * k4 + I].f, cst.neg_ge, left[4 * k4 + I].h); \
        f = __viaddmax_s16x2(f, cst.neg_ge, hgo);                                            \
        const uint32_t d = V16::add(dgo, s);                                                 \
        h = __vimax3_s16x2_relu(d, e, f);                                                    \
        dgo = left[4 * k4 + I].h;                                                            \
        hgo = V16::add(h, cst.neg_go);                                                       \
        left[4 * k4 + I].h = hgo;                                                            \
        left[4 * k4 + I].f = e;                                                              \
        hh[I] = d;                                                                           \
    }
#else
#define SWB_CELL(I)                                                                          \
    {                                                                                        \
        const uint32_t s = V16Base::pair<I>(wa, wb);                                         \
        const uint32_t e = __viaddmax_s16x2(left[4 * k4 + I].f, cst.neg_ge, left[4 * k4 + I].h); \
        f = __viaddmax_s16x2(f, cst.neg_ge, hgo);                                            \
        h = __viaddmax_s16x2_relu(dgo, s, V16Base::max2(e, f));                              \
        dgo = left[4 * k4 + I].h;                                                            \
        hgo = V16::add(h, cst.neg_go);                                                       \
        left[4 * k4 + I].h = hgo;                                                            \
        left[4 * k4 + I].f = e;                                                              \
        hh[I] = h;                                                                           \
    }
#endif
            SWB_CELL(0) SWB_CELL(1)
            best.h = __vimax3_s16x2(best.h, hh[0], hh[1]);
            SWB_CELL(2) SWB_CELL(3)
            best.h = __vimax3_s16x2(best.h, hh[2], hh[3]);
#undef SWB_CELL
        }
        return mk(h, f);
    }
    template <class BE> static SWB_HD T shfl_up(BE &be, T v, int d, int w)
    {
        return mk(be.shfl_up(v.h, d, w), be.shfl_up(v.f, d, w));
    }
    template <class BE> static SWB_HD T shfl_xor(BE &be, T v, int m, int w)
    {
        return mk(be.shfl_xor(v.h, m, w), be.shfl_xor(v.f, m, w));
    }
    template <class BE> static SWB_HD T ld(BE &be, const T *p)
    {
        uint2 v = be.ld_cg2(reinterpret_cast<const uint2 *>(p));
        return mk(v.x, v.y);
    }
    template <class BE> static SWB_HD void st(BE &be, T *p, T v)
    {
        be.st_cg2(reinterpret_cast<uint2 *>(p), make_uint2(v.h, v.f));
    }
    template <class BE> static SWB_HD void ld4(BE &be, const T *p, T *o)
    {
        uint4 v = be.ld_cg4(reinterpret_cast<const uint4 *>(p));
        uint4 w = be.ld_cg4(reinterpret_cast<const uint4 *>(p) + 1);
        o[0] = mk(v.x, v.y); o[1] = mk(v.z, v.w);
        o[2] = mk(w.x, w.y); o[3] = mk(w.z, w.w);
    }
    template <class BE> static SWB_HD void st4(BE &be, T *p, const T *o)
    {
        be.st_cg4(reinterpret_cast<uint4 *>(p), make_uint4(o[0].h, o[0].f, o[1].h, o[1].f));
        be.st_cg4(reinterpret_cast<uint4 *>(p) + 1, make_uint4(o[2].h, o[2].f, o[3].h, o[3].f));
    }
};

struct V32A {
    static const bool rebased = false;
    static const bool blocked = false;
    static const bool is16 = false;
    struct T { int ha, hb, fa, fb; };
    struct C { int go, ge; };
    static SWB_HD T mk(int ha, int hb, int fa, int fb) { T t; t.ha = ha; t.hb = hb; t.fa = fa; t.fb = fb; return t; }
    static SWB_HD int mx(int a, int b) { return a > b ? a : b; }
    static SWB_HD C consts(const SwbScoreParams &p) { C c; c.go = p.gap_open; c.ge = p.gap_extend; return c; }
    static SWB_HD T hzero(const C &c) { return mk(0, 0, -c.go, -c.go); }
    static SWB_HD T lzero(const C &c) { return mk(-c.go, -c.go, -c.go, -c.go); }
    static SWB_HD int score_lo(T best, const C &) { return best.ha; }
    static SWB_HD int score_hi(T best, const C &) { return best.hb; }
    static SWB_HD T max2(T a, T b) { return mk(mx(a.ha, b.ha), mx(a.hb, b.hb), a.fa, a.fb); }
    template <int K, int LDW>  // LDW (bytes per profile load) is a hint of the s16 policies: this one loads words
    static SWB_HD T column(T up, T &diag0, T (&left)[K], T &best, const C &cst, uint32_t codeA, uint32_t codeB,
                           const int8_t *prow, uint32_t sstride)
    {
        const uint32_t *ra = reinterpret_cast<const uint32_t *>(prow + codeA * sstride);
        const uint32_t *rb = reinterpret_cast<const uint32_t *>(prow + codeB * sstride);
        int hgoa = up.ha - cst.go, hgob = up.hb - cst.go;
        int fa = up.fa, fb = up.fb;
        int dgoa = diag0.ha, dgob = diag0.hb;
        int ha = up.ha, hb = up.hb;
        diag0.ha = hgoa;
        diag0.hb = hgob;
#pragma unroll
        for (int k4 = 0; k4 < K / 4; ++k4) {
            const uint32_t wa = ra[k4];
            const uint32_t wb = rb[k4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                T &l = left[4 * k4 + i];
                const int sa = (int)(int8_t)(wa >> (8 * i));  // S + go
                const int sb = (int)(int8_t)(wb >> (8 * i));
                const int ea = mx(l.fa - cst.ge, l.ha), eb = mx(l.fb - cst.ge, l.hb);
                fa = mx(fa - cst.ge, hgoa);
                fb = mx(fb - cst.ge, hgob);
                ha = mx(mx(dgoa + sa, mx(ea, fa)), 0);
                hb = mx(mx(dgob + sb, mx(eb, fb)), 0);
                dgoa = l.ha;
                dgob = l.hb;
                hgoa = ha - cst.go;
                hgob = hb - cst.go;
                l = mk(hgoa, hgob, ea, eb);
                best.ha = mx(best.ha, ha);
                best.hb = mx(best.hb, hb);
            }
        }
        return mk(ha, hb, fa, fb);
    }
    template <class BE> static SWB_HD T shfl_up(BE &be, T v, int d, int w)
    {
        return mk((int)be.shfl_up((uint32_t)v.ha, d, w), (int)be.shfl_up((uint32_t)v.hb, d, w),
                  (int)be.shfl_up((uint32_t)v.fa, d, w), (int)be.shfl_up((uint32_t)v.fb, d, w));
    }
    template <class BE> static SWB_HD T shfl_xor(BE &be, T v, int m, int w)
    {
        return mk((int)be.shfl_xor((uint32_t)v.ha, m, w), (int)be.shfl_xor((uint32_t)v.hb, m, w), v.fa, v.fb);
    }
    template <class BE> static SWB_HD T ld(BE &be, const T *p)
    {
        uint4 v = be.ld_cg4(reinterpret_cast<const uint4 *>(p));
        return mk((int)v.x, (int)v.y, (int)v.z, (int)v.w);
    }
    template <class BE> static SWB_HD void st(BE &be, T *p, T v)
    {
        be.st_cg4(reinterpret_cast<uint4 *>(p), make_uint4(v.ha, v.hb, v.fa, v.fb));
    }
    template <class BE> static SWB_HD void ld4(BE &be, const T *p, T *o)
    {
#pragma unroll
        for (int u = 0; u < 4; ++u) o[u] = ld(be, p + u);
    }
    template <class BE> static SWB_HD void st4(BE &be, T *p, const T *o)
    {
#pragma unroll
        for (int u = 0; u < 4; ++u) st(be, p + u, o[u]);
    }
};

// ---------------------------------------------------------------------------------------------
// One tile, all query rows of the current chunk.
// Boundary scratch layout (elements of V::T, base tile.bnd_off), values in the policy's h domain:
//   G == 1 : [chunk c][lane][4 columns]   -> one 16/32-byte vector per lane and chunk
//   G  > 1 : [slot][column]               -> the first lane of a group reads 4 columns as one vector, the last lane
//                                            writes 4 columns as one vector (every lane lags one chunk behind the
//                                            lane above)
//
// SPLIT (lane-group tiles of long sequences): the passes of one tile are separate work items taken by different warps,
// which run as a pipeline over the columns. The warp of pass ss publishes how many columns of its bottom row are in the
// boundary scratch (prog[ss], every 64 columns, release store); the warp of pass ss+1 polls it (acquire load, back-off)
// before it reads them (a resume margin, SWB_SPLIT_HYST, was measured and is 0).
// Items are handed out by one counter, pass-major inside a lane-group class (pass 0 of every tile, then pass 1, ...): a
// waiting warp always waits for an item that was handed out before its own, i.e. that a resident warp owns or has
// finished -- no deadlock -- and only a few passes per tile are in flight at a time, so little of a launch's first wave
// is spent waiting for the pipeline of one tile to fill. Scores of the passes are combined with atomicMax.
template <int K, class V, bool GROUPED, bool SPLIT, class BE>
SWB_HD void swb_run_tile(BE &be, const SwbScoreParams &p, const SwbTile &tile, uint32_t tile_idx,
                         const int8_t *sprof, uint32_t sstride, uint32_t ss_begin = 0, uint32_t ss_count = 0xffffffffu,
                         uint32_t *prog = nullptr, uint32_t smem_ss0 = 0)
{
    typedef typename V::T T;
    typename V::C cst = V::consts(p);
    const T LZERO = V::lzero(cst);
    const int lane = be.lane();
    const int logG = GROUPED ? (int)tile.logG : 0;
    const int G = 1 << logG;
    const int g = lane & (G - 1);
    const int slot = lane >> logG;
    const int P = 32 >> logG;
    const uint32_t W = tile.width;
    const uint32_t nchunks = W >> 2;
    const bool lead = (g == 0);
    const bool tail = (g == G - 1);
    const uint32_t rows_per_super = (uint32_t)K << logG;
    const uint32_t nsuper = (p.rows + rows_per_super - 1) / rows_per_super;
    const uint8_t *res = p.residues + tile.res_off + (size_t)slot * 8u;
    const size_t res_stride = (size_t)P * 8u;
    T *bnd = reinterpret_cast<T *>(p.bnd) + tile.bnd_off;
    T best = V::hzero(cst);
    // V16R: base log of the boundary rows, [pass parity][slot][block] (see swb_blog_*)
    const uint32_t cbmask = V::rebased ? (1u << (p.rebase_shift - 2u)) - 1u : 0u;  // chunks per rebase block - 1
    const uint32_t blog_nb = V::rebased ? (W >> p.rebase_shift) + 2u : 0u;
    const size_t blog_par = ((size_t)W * (size_t)P >> 6) + 33u;
    uint2 *blog = V::rebased ? reinterpret_cast<uint2 *>(p.blog) + swb_blog_offset(tile.bnd_off, tile_idx) : nullptr;
    const uint32_t pass0 = V::rebased ? p.row0 / rows_per_super : 0u;  // passes of the query chunks before this launch

    const uint32_t ss_end = ss_count < nsuper - ss_begin ? ss_begin + ss_count : nsuper;
    if constexpr (!GROUPED) {
        // One lane per pair: no exchange between lanes. Residue codes of a chunk (4 columns x two sequences) come as byte
        // loads: zero-extended in registers, no ALU-pipe instruction is spent on unpacking them.
        //
        // Order of the work. The straight order -- pass after pass (K rows each) over the whole width -- streams the
        // tile's residues and its boundary row (bottom row of one pass = top row of the next, 4 B per pair-column)
        // through memory once per PASS, and with ~2,400 resident warps those streams are larger than the L2: ncu shows
        // 140 GB of DRAM traffic for one 5,478-row query against 0.2 GB of residues. With SWB_PASS_GROUP > 1 the work is
        // blocked twice instead: that many consecutive passes walk TOGETHER over column blocks of SWB_BLOCK_CHUNKS
        // chunks (block 0 by every pass of the group, then block 1, ...), the boundary row and the residues of a block
        // are reused from L1/L2 by the next pass a few microseconds later, and what a pass needs to continue in the next
        // block (its K row values and the diagonal element) is parked in a small per-warp scratch (colstate); only the
        // bottom row of a whole group still goes through memory. Measured on B200 (swb_types.h): half the DRAM traffic,
        // 3.5 % slower -- the ALU pipe is the bound, not memory -- so the product builds with groups of one.
        // Only the V16 policy can be blocked: the two-value affine state (64 registers of rows at K = 32) spilled in the
        // hot loop with the parking code (4,905 against 5,249 GCUPS).
        constexpr uint32_t PGV = V::blocked ? SWB_PASS_GROUP : 1u;
        T *const cstate = PGV > 1u ? reinterpret_cast<T *>(p.colstate) + (size_t)be.warp_slot() * swb_colstate_elems(K) : nullptr;
        const size_t cs_pass = (size_t)(K + 4) * 32u;  // elements of T per parked pass: [K/4 + 1][lane][4]
        for (uint32_t pg0 = ss_begin; pg0 < ss_end; pg0 += PGV) {
            const uint32_t pg1 = pg0 + PGV < ss_end ? pg0 + PGV : ss_end;
            const uint32_t cbc = PGV > 1u && pg1 - pg0 > 1u ? SWB_BLOCK_CHUNKS : nchunks;  // a lone pass runs straight through
            for (uint32_t cb0 = 0; cb0 < nchunks; cb0 += cbc) {
                const uint32_t cb1 = cb0 + cbc < nchunks ? cb0 + cbc : nchunks;
                for (uint32_t ss = pg0; ss < pg1; ++ss) {
                    const int8_t *prow = sprof + (size_t)((ss - smem_ss0) * (uint32_t)K);
                    const bool read_top = !(p.first_chunk && ss == 0);
                    const bool write_bot = !(p.last_chunk && ss + 1 == nsuper);
                    T *const park = cstate + (size_t)(ss - pg0) * cs_pass + (size_t)lane * 4u;
                    T left[K];
                    T diag0;
                    if (PGV == 1u || cb0 == 0) {
#pragma unroll
                        for (int k = 0; k < K; ++k) left[k] = LZERO;
                        diag0 = LZERO;
                    } else {
#pragma unroll
                        for (int k4 = 0; k4 < K / 4; ++k4) V::ld4(be, park + (size_t)k4 * 128u, &left[4 * k4]);
                        T d4[4];
                        V::ld4(be, park + (size_t)(K / 4) * 128u, d4);
                        diag0 = d4[0];
                    }
                    uint32_t ca[4], cb[4];
                    T bc[4];
                    {
                        const uint8_t *r0 = res + (size_t)cb0 * res_stride;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            ca[u] = be.ld_code(r0 + 2 * u);
                            cb[u] = be.ld_code(r0 + 2 * u + 1);
                            bc[u] = V::hzero(cst);
                        }
                        if (read_top) V::ld4(be, bnd + ((size_t)cb0 * 32u + lane) * 4u, bc);
                    }
                    for (uint32_t c = cb0; c < cb1; ++c) {
                        // prefetch the next chunk of residues and of the top boundary row
                        uint32_t na[4], nb[4];
                        T bn[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            na[u] = SWB_PAD;
                            nb[u] = SWB_PAD;
                            bn[u] = V::hzero(cst);
                        }
                        if (c + 1 < cb1) {
                            const uint8_t *rnext = res + (size_t)(c + 1) * res_stride;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                na[u] = be.ld_code(rnext + 2 * u);
                                nb[u] = be.ld_code(rnext + 2 * u + 1);
                            }
                            if (read_top) V::ld4(be, bnd + ((size_t)(c + 1) * 32u + lane) * 4u, bn);
                        }
                        T outb[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            outb[u] = V::template column<K, SWB_BULK_LDW>(bc[u], diag0, left, best, cst, ca[u], cb[u], prow, sstride);
                        if (write_bot) V::st4(be, bnd + ((size_t)c * 32u + lane) * 4u, outb);
                        // The loads above run one chunk ahead, which covers a cache hit but not a trip to HBM: the first
                        // pass of a group meets residues and a boundary row that nobody touched recently. Pull the
                        // lines of the chunk SWB_PF_CHUNKS ahead into L2; no register is tied up. HERE, behind the
                        // columns: at the top of the chunk the next instruction that reused a register of the
                        // prefetch's address pair waited on it for ~500 cycles per chunk (12.7 % of the stall samples,
                        // profiles/r2y_*; 9,458 -> 10,003 GCUPS by moving it, profiles/r2z_sweep_pf_mode.txt).
                        if (SWB_PF_CHUNKS > 0u && ss == pg0 && c + SWB_PF_CHUNKS < nchunks) {
                            be.prefetch_l2(res + (size_t)(c + SWB_PF_CHUNKS) * res_stride);
                            if (read_top) be.prefetch_l2(bnd + ((size_t)(c + SWB_PF_CHUNKS) * 32u + lane) * 4u);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            ca[u] = na[u];
                            cb[u] = nb[u];
                            bc[u] = bn[u];
                        }
                    }
                    if (PGV > 1u && cb1 < nchunks) {
#pragma unroll
                        for (int k4 = 0; k4 < K / 4; ++k4) V::st4(be, park + (size_t)k4 * 128u, &left[4 * k4]);
                        T d4[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) d4[u] = diag0;
                        V::st4(be, park + (size_t)(K / 4) * 128u, d4);
                    }
                }
            }
        }
    } else {
    for (uint32_t ss = ss_begin; ss < ss_end; ++ss) {
        // smem_ss0: the pass whose first row sits at row 0 of the staged profile (0 except in SPLIT launches)
        const int8_t *prow = sprof + (size_t)((((ss - smem_ss0) << logG) + (uint32_t)g) * (uint32_t)K);
        const bool read_top = !(p.first_chunk && ss == 0);
        const bool write_bot = !(p.last_chunk && ss + 1 == nsuper);
        const bool wait_top = SPLIT && ss > 0;  // the row above comes from another warp of this launch
        uint32_t avail = 0;                     // columns of that row known to be in the scratch
        T left[K];
#pragma unroll
        for (int k = 0; k < K; ++k) left[k] = LZERO;
        T diag0 = LZERO;
        uint32_t ca[4], cb[4];
        T bc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) bc[u] = V::hzero(cst);
        bool top_cur = nchunks > 0 && read_top;  // bc holds values of the row above (uniform over the warp)
        if (wait_top) {
            const uint32_t need = W < 4u ? W : 4u;
            avail = be.wait_progress(prog + ss - 1, need, W);
        }

        if constexpr (!SPLIT) {
            // Lane-group wavefront, one COLUMN of lag per lane: at column step t lane g works on column t - g and hands
            // its bottom H and the residue pair to lane g + 1 with shuffles; the first lane of a group fetches the
            // residue codes and the top boundary row, the last lane stores the bottom row (scalars: its columns lag
            // G - 1 behind, so they are not 4-aligned). Measured on B200 this form is the faster one for the tiles of
            // the bulk launches (a warp shares its scheduler with three others, the short lag keeps its lanes' chains
            // close together); the pipelined (SPLIT) launches below use one chunk of lag instead.
            const uint32_t nsteps4 = (W + (uint32_t)G - 1u + 3u) >> 2;
            T hprev = V::hzero(cst);
            uint32_t aprev = SWB_PAD, bprev = SWB_PAD;
            int colg = -g;  // V16R: this lane's column
            T dconv = T(), dcur = T();
            uint2 *blog_rd = nullptr, *blog_wr = nullptr;
            const uint32_t colmask = V::rebased ? (1u << p.rebase_shift) - 1u : 0u;
            if constexpr (V::rebased) {
                blog_wr = blog + (size_t)((pass0 + ss) & 1u) * blog_par + (size_t)slot * blog_nb;
                blog_rd = blog + (size_t)((pass0 + ss + 1u) & 1u) * blog_par + (size_t)slot * blog_nb;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ca[u] = SWB_PAD;
                cb[u] = SWB_PAD;
            }
            if (lead && nchunks > 0) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ca[u] = be.ld_code(res + 2 * u);
                    cb[u] = be.ld_code(res + 2 * u + 1);
                }
                if (read_top) V::ld4(be, bnd + (size_t)slot * W, bc);
            }
            for (uint32_t c = 0; c < nsteps4; ++c) {
                // prefetch the next chunk of residues and of the top boundary row
                uint32_t na[4], nb[4];
                T bn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    na[u] = SWB_PAD;
                    nb[u] = SWB_PAD;
                    bn[u] = V::hzero(cst);
                }
                const bool top_next = c + 1 < nchunks && read_top;
                if (lead && c + 1 < nchunks) {
                    const uint8_t *rnext = res + (size_t)(c + 1) * res_stride;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        na[u] = be.ld_code(rnext + 2 * u);
                        nb[u] = be.ld_code(rnext + 2 * u + 1);
                    }
                    if (read_top) V::ld4(be, bnd + (size_t)slot * W + (size_t)(c + 1) * 4u, bn);
                }
                T outb[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t a = ca[u], b = cb[u];
                    T up = bc[u];
                    const uint32_t a2 = be.shfl_up(aprev, 1, G);
                    const uint32_t b2 = be.shfl_up(bprev, 1, G);
                    const T u2 = V::shfl_up(be, hprev, 1, G);
                    if constexpr (V::rebased) {
                        // first column of a block: move to the new base. The first lane of the group takes it from its
                        // own bottom row (the column just behind), the others receive the amount from the lane above,
                        // which made the same move one column step ago.
                        const T d2 = V::shfl_up(be, dcur, 1, G);
                        if (colg > 0 && ((uint32_t)colg & colmask) == 0u) {
                            const T d = lead ? hprev : d2;
                            dcur = d;
                            V::template rebase<K>(d, diag0, left, best, cst);
                            const uint32_t blk = (uint32_t)colg >> p.rebase_shift;
                            if (lead) {
                                dconv = T();
                                if (read_top) {
                                    const uint2 bw = be.ld_cg2(blog_rd + blk);
                                    dconv = V::pack((int)bw.x - cst.baseA, (int)bw.y - cst.baseB);
                                }
                            }
                            if (tail && write_bot) be.st_cg2(blog_wr + blk, make_uint2((uint32_t)cst.baseA, (uint32_t)cst.baseB));
                        }
                        // the row above: relative to its writer's base -> to this lane's base; none: zero (absolute)
                        up = top_cur ? V::add(up, dconv) : V::hzero(cst);
                        ++colg;
                    }
                    if (!lead) { a = a2; b = b2; up = u2; }
                    const T h = V::template column<K, SWB_BULK_LDW>(up, diag0, left, best, cst, a, b, prow, sstride);
                    outb[u] = h;
                    hprev = h;
                    aprev = a;
                    bprev = b;
                }
                if (write_bot && tail) {
                    const int32_t col0 = (int32_t)(c * 4u) - (G - 1);  // column of outb[0]
                    T *dst = bnd + (size_t)slot * W + col0;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (col0 + u >= 0 && col0 + u < (int32_t)W) V::st(be, dst + u, outb[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ca[u] = na[u];
                    cb[u] = nb[u];
                    bc[u] = bn[u];
                }
                top_cur = top_next;
            }
        } else {
            // Pipelined launches: lane-group wavefront with one CHUNK (4 columns) of lag per lane: at step c lane g
            // works on chunk c - g. Every
            // lane loads the residue codes of its chunk itself (byte loads; the lane above touched the same line one
            // step earlier), receives the four bottom H of the lane above -- its outputs of the previous step, i.e. of
            // this very chunk -- with four shuffles, and the last lane stores its four bottom H as one vector. Lanes
            // whose chunk lies before the first or behind the last one run on padding codes, which cannot raise a score.
            constexpr int LDW = K >= 16 ? 16 : 8;  // bytes per profile load (the SPLIT kernels stage aligned code rows)
            const uint32_t nsteps = nchunks + (uint32_t)G - 1u;
            T hout[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) hout[u] = V::hzero(cst);
            // V16R: the base difference to the row above, the amount of this lane's latest rebase
            T dconv = T(), dcur = T();
            uint2 *blog_rd = nullptr, *blog_wr = nullptr;
            if constexpr (V::rebased) {
                blog_wr = blog + (size_t)((pass0 + ss) & 1u) * blog_par + (size_t)slot * blog_nb;
                blog_rd = blog + (size_t)((pass0 + ss + 1u) & 1u) * blog_par + (size_t)slot * blog_nb;
            }
            {
                const bool have = lead && nchunks > 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ca[u] = have ? be.ld_code(res + 2 * u) : (uint32_t)SWB_PAD;
                    cb[u] = have ? be.ld_code(res + 2 * u + 1) : (uint32_t)SWB_PAD;
                }
                if (have && read_top) V::ld4(be, bnd + (size_t)slot * W, bc);
            }
            for (uint32_t c = 0; c < nsteps; ++c) {
                const int32_t cg = (int32_t)c - g;  // this lane's chunk
                // prefetch: the codes of this lane's next chunk, the next chunk of the top boundary row (first lane)
                uint32_t na[4], nb[4];
                T bn[4];
                const bool top_next = c + 1 < nchunks && read_top;
                if (wait_top && c + 1 < nchunks) {
                    const uint32_t need = 4u * c + 8u < W ? 4u * c + 8u : W;
                    if (avail < need) avail = be.wait_progress(prog + ss - 1, need, W);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    na[u] = SWB_PAD;
                    nb[u] = SWB_PAD;
                    bn[u] = V::hzero(cst);
                }
                if ((uint32_t)(cg + 1) < nchunks) {
                    const uint8_t *rnext = res + (size_t)(uint32_t)(cg + 1) * res_stride;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        na[u] = be.ld_code(rnext + 2 * u);
                        nb[u] = be.ld_code(rnext + 2 * u + 1);
                    }
                }
                if (lead && top_next) V::ld4(be, bnd + (size_t)slot * W + (size_t)(c + 1) * 4u, bn);
                T up[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) up[u] = V::shfl_up(be, hout[u], 1, G);
                if constexpr (V::rebased) {
                    // first chunk of a block: move to the new base. The first lane of the group takes it from its own
                    // bottom row (the column just behind), the others receive the amount from the lane above, which
                    // made the same move one step ago (before it computed the values this lane has just received).
                    const T d2 = V::shfl_up(be, dcur, 1, G);
                    if (cg > 0 && ((uint32_t)cg & cbmask) == 0u) {
                        const T d = lead ? hout[3] : d2;
                        dcur = d;
                        V::template rebase<K>(d, diag0, left, best, cst);
                        const uint32_t blk = (uint32_t)cg >> (p.rebase_shift - 2u);
                        if (lead) {
                            dconv = T();
                            if (read_top) {
                                const uint2 bw = be.ld_cg2(blog_rd + blk);
                                dconv = V::pack((int)bw.x - cst.baseA, (int)bw.y - cst.baseB);
                            }
                        }
                        if (tail && write_bot) be.st_cg2(blog_wr + blk, make_uint2((uint32_t)cst.baseA, (uint32_t)cst.baseB));
                    }
                }
                if (lead) {
                    // the row above: relative to its writer's base -> to this lane's base (V16R); none: zero
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if constexpr (V::rebased) up[u] = top_cur ? V::add(bc[u], dconv) : V::hzero(cst);
                        else up[u] = bc[u];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    hout[u] = V::template column<K, LDW>(up[u], diag0, left, best, cst, ca[u], cb[u], prow, sstride);
                if (write_bot) {
                    if (tail && (uint32_t)cg < nchunks) V::st4(be, bnd + (size_t)slot * W + (size_t)cg * 4u, hout);
                    if (SPLIT && ((c & 15u) == 15u || c + 1 == nsteps)) {
                        // every slot's last lane has stored its chunk: one lane publishes for the warp
                        be.syncwarp();
                        if (lane == 31) {
                            const int32_t done = 4 * ((int32_t)c - (G - 1) + 1);
                            be.publish(prog + ss, (uint32_t)(done < 0 ? 0 : (done > (int32_t)W ? (int32_t)W : done)));
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ca[u] = na[u];
                    cb[u] = nb[u];
                    bc[u] = bn[u];
                }
                top_cur = top_next;
            }
        }
        if constexpr (V::rebased) {  // the next pass starts at base 0 again
            V::fold(best, cst);
            V::set_base(cst, 0, 0);
        }
        be.syncwarp();
    }
    }

    if constexpr (V::rebased) {
        for (int m = G >> 1; m >= 1; m >>= 1) {
            const int oa = (int)be.shfl_xor((uint32_t)cst.bestA, m, G), ob = (int)be.shfl_xor((uint32_t)cst.bestB, m, G);
            cst.bestA = oa > cst.bestA ? oa : cst.bestA;
            cst.bestB = ob > cst.bestB ? ob : cst.bestB;
        }
    } else if (GROUPED) {
        for (int m = G >> 1; m >= 1; m >>= 1) best = V::max2(best, V::shfl_xor(be, best, m, G));
    }
    bool flagged = false;
    if (lead && slot < (int)tile.npairs) {
        const size_t s0 = 2u * ((size_t)tile.first_pair + (size_t)slot);
        int a = V::score_lo(best, cst), b = V::score_hi(best, cst);
        if (SPLIT) {
            be.atomic_max(p.scores + s0, a);
            be.atomic_max(p.scores + s0 + 1, b);
        } else {
            if (!p.first_chunk) {
                const int pa = p.scores[s0], pb = p.scores[s0 + 1];
                a = a > pa ? a : pa;
                b = b > pb ? b : pb;
            }
            p.scores[s0] = a;
            p.scores[s0 + 1] = b;
        }
        flagged = V::is16 && ((a > b ? a : b) > p.ovf_thr);
    }
    if (V::is16) {
        if (be.any(flagged) && lane == 0) p.flags[tile_idx] = 1;
    } else if (p.recount && p.last_chunk && (!SPLIT || ss_end == nsuper) && lane == 0) {
        be.count(p.recount);
    }
}

// Per-warp loop over the dynamically scheduled tiles of one launch. A launch covers up to SWB_MAX_RANGES
// ranges of the tile array (the tiles of the group sizes that use this kernel's K for this query); the shared
// counter hands out positions of the concatenated ranges, longest tiles first. SPLIT launches hand out (tile, pass)
// items of the first range instead.
template <int K, class V, bool SPLIT, class BE>
SWB_HD void swb_warp_loop(BE &be, const SwbScoreParams &p, const int8_t *sprof, uint32_t sstride)
{
    // First wave: warp w of block b starts with work item b * (warps per block) + w, every later item comes from the
    // shared counter. Items are ordered longest first, so the warps of a block start on tiles of nearly the same length
    // and finish together: a block (and its shared memory) leaves the SM when its SLOWEST warp is done, and with a
    // first wave handed out in arrival order nearly every block held one of the longest tiles, i.e. seven waiting warps.
    bool first = !SPLIT && p.static_wave != 0u;
    for (;;) {
        uint32_t v;
        if (first) {
            first = false;
            v = be.static_item(p.warps_active);
        } else {
            v = be.next_tile(p.counter) + (SPLIT ? 0u : p.static_wave);
        }
        if (v >= p.ntiles) break;
        if (SPLIT) {
            // one warp per block; it stages only the K << logG profile rows of its pass (sstride = K * 32 + 4).
            // Item -> (tile, pass): the split set is a run of tiles per lane-group size (class j = 32 >> j lanes)
            int l = SWB_MAX_LOGG;
            uint32_t tile0 = p.split_tile_start[0], item0 = 0;
#pragma unroll
            for (int j = 0; j + 1 < SWB_MAX_LOGG; ++j)
                if (v >= p.split_item_end[j]) {
                    l = SWB_MAX_LOGG - 1 - j;
                    tile0 = p.split_tile_start[j + 1];
                    item0 = p.split_item_end[j];
                }
            const uint32_t rows_per_pass = (uint32_t)K << l;
            const uint32_t passes = (p.rows + rows_per_pass - 1u) / rows_per_pass;
            const uint32_t ntl = (p.split_item_end[SWB_MAX_LOGG - l] - item0) / passes;  // tiles of this class
            const uint32_t ss = (v - item0) / ntl, t = (v - item0) - ss * ntl;            // pass-major
            const uint32_t ti = tile0 + t;
            if (p.only_flagged && !be.ld_flag(p.flags + ti)) continue;  // every pass of the tile skips alike
            const SwbTile tile = be.ld_tile(p.tiles + ti);
            be.stage_rows(const_cast<int8_t *>(sprof), sstride, p.profile, p.prof_stride, p.row0 + ss * rows_per_pass,
                          rows_per_pass);
            swb_run_tile<K, V, true, true>(be, p, tile, ti, sprof, sstride, ss, 1u, p.prog + item0 + (size_t)t * passes, ss);
            continue;
        }
        const uint32_t vt = v;
        uint32_t ti = p.range_start[0] + vt;
#pragma unroll
        for (int r = 1; r < SWB_MAX_RANGES; ++r)
            if (vt >= p.range_cum[r - 1]) ti = p.range_start[r] + (vt - p.range_cum[r - 1]);
        if (p.only_flagged && !be.ld_flag(p.flags + ti)) continue;
        const SwbTile tile = be.ld_tile(p.tiles + ti);
        if constexpr (V::rebased) {  // rebasing lives in the lane-group program; one lane per pair is its G = 1 case
            swb_run_tile<K, V, true, false>(be, p, tile, ti, sprof, sstride);
        } else {
            if (tile.logG == 0)
                swb_run_tile<K, V, false, false>(be, p, tile, ti, sprof, sstride);
            else
                swb_run_tile<K, V, true, false>(be, p, tile, ti, sprof, sstride);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Packed residue word of a tile: 4 columns x (seqA, seqB) of one slot. raw = concatenated codes of
// the whole DB; seq_off / seq_len are indexed by SORTED position. Shared by the device pack kernel
// and the host emulation.
SWB_HD uint64_t swb_pack_word(const SwbTile &t, uint32_t c, uint32_t slot, const uint8_t *raw,
                              const uint64_t *seq_off, const uint32_t *seq_len, uint32_t nseq)
{
    uint64_t w = 0;
    for (int half = 0; half < 2; ++half) {
        const uint64_t s = 2ull * ((uint64_t)t.first_pair + slot) + half;
        const bool live = slot < t.npairs && s < nseq;
        const uint32_t len = live ? seq_len[s] : 0u;
        const uint8_t *src = live ? raw + seq_off[s] : raw;
        for (int u = 0; u < 4; ++u) {
            const uint32_t col = c * 4u + u;
            const uint32_t code = col < len ? (uint32_t)(src[col] & 31u) : (uint32_t)SWB_PAD;
            w |= (uint64_t)code << (8 * (2 * u + half));
        }
    }
    return w;
}
