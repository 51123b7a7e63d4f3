// Register-resident stripe kernels for affine_3 (src/algn.c:2411-2548 and its cost-only twin :2287-2403).
//
// A group of G lanes owns one pair.  Lane l keeps the DP state of 2K consecutive diagonals of the stripe in
// registers (4 ints each: close-block CB, extend-vertical EV, extend-horizontal EH, extend-block EB), the
// sweep advances one anti-diagonal per step, and a step updates the K diagonals of matching parity:
//
//   double step u, lane l:  rows  i0 - m          (i0 = u - l K,        m = 0..K-1)
//                           cols  j0 + m [+1]     (j0 = u + d0 + l K)
//     even step: cells (i0 - m, j0 + m)      on diagonals q = 2m      left <- q-1, up <- q+1, diag <- q
//     odd  step: cells (i0 - m, j0 + m + 1)  on diagonals q = 2m + 1
//
// so the only lane-to-lane traffic is two warp shuffles per step (the left neighbour of q = 0 comes from lane
// l-1, the upper neighbour of q = 2K-1 from lane l+1), and the per-row / per-column gap bookkeeping
// (cells.cuh AffRow / AffCol) slides through register windows, one new row and one new column per double step.
//
// Stripe placement: diagonal dhi + 1 -- the reference's "poisoned" cells (:2528-2536) -- always sits on the
// last diagonal of the last lane, so d0 = dhi + 2 - 2KG.  If that leaves diagonals below dlo (LOW = true), they
// follow the reference's left-edge rule (:2476-2494), which is what the cell on dlo must see to its left.
// Row 0 and column 0 (initialize_matrices_affine :2183-2248) are produced by the same sweep with their own
// formulas during a short boundary phase; the steady phase carries no per-cell position tests at all.
//
// Direction bytes go to HBM in anti-diagonal-major order: each lane stores its K bytes of a step as one
// 4- or 8-byte word, a group's lanes are contiguous (coalesced 32/64-byte segments).
#pragma once
#include <algorithm>
#include <type_traits>

#include "cells.cuh"
#include "classes.h"
#include "staging.cuh"

namespace poyb200 {

constexpr int STRIPE_WARPS = 4;
#ifndef STRIPE_MIN_BLOCKS
#define STRIPE_MIN_BLOCKS 2
#endif
constexpr int STRIPE_LUT_BYTES = 16 * 17 * 8;  // == 16 * LUT_ROW_BYTES
constexpr int STRIPE_TABLE_BYTES = STRIPE_LUT_BYTES + 64 * 4 + STRIPE_WARPS * 4 * STAGE_BAR_BYTES;

// All DP values are carried multiplied by 4, and each of the four states keeps a constant 2-bit tag in its low
// bits: EH 0, CB 1, EV 2, EB 3.  The tags are the tie-break priorities of the reference's traceback:
//   * ASSIGN_MINIMUM keeps every minimal state and backtrace_affine reads them H > A > V > D (:2006-2012), so
//     min(EH, CB, EV, EB) over the tagged values yields the mode in its low two bits;
//   * FILL_CLOSE_BLOCK_DIAGONAL keeps every minimal predecessor, read H > D > V > A (:2049-2051); the four
//     candidates inherit the tags of the states they start from (0, 3, 2, 1) and two of them are re-tagged by
//     +-2 (folded into constants), so the min yields that choice too.
// Comparisons the reference makes between two candidates (extend < open) always see equal tags, so they are
// exact.  HIGH_NUM becomes 4 000 000; nothing comes near 2^31.
constexpr int HIGH4 = 4 * HIGH_NUM;
constexpr int TAG_EH = 0, TAG_CB = 1, TAG_EV = 2, TAG_EB = 3;
constexpr int LUT_ROW_BYTES = 17 * 8;  // 16 int2 entries + 1 pad: spreads the rows over the shared-memory banks

struct AffWinRow {
    int vx4;       // 4 * si_vertical_extension
    int gopge4p1;  // 4 * (si_gap_opening + si_gap_extension) + 1   (CB tag 1 -> EV tag 2)
    int gop4;      // 4 * si_gap_opening
    int lut;       // byte offset of the row (si & 15) in the cost LUT
    int gf;        // 1 when si carries the gap bit, else 0 (a multiplier: the selects run on the FMA pipe as IMADs)
};
struct AffWinCol {
    int hx4;       // 4 * sj_horizontal_extension[j]
    int gopg4m1;   // 4 * (gap_open_prec[j] + gap_row[j]) - 1       (CB tag 1 -> EH tag 0)
    int gop4;      // 4 * gap_open_prec[j]
    int lut;       // byte offset of the column (sj & 15) inside a LUT row
    int gf;
};

// One interior cell on tagged, x4 values, for DNA matrices (lcm = 5, gap = 16 = TMPGAP, codes < 32), where "has
// the gap bit", "si_base != si_no_gap" and "si & TMPGAP" are the same predicate.  d = {4*cost, 4*cost - 2}.
// Returns the direction byte (common.cuh).
//
// NOEB: neither operand carries a gap bit beyond its leading element and gap_open > 0.  Then `both` is false in every
// cell, the block-diagonal state is >= HIGH_NUM everywhere except the origin, its only finite use -- the
// ALIGN_TO_DIAGONAL candidate of cell (1, 1), worth cost + gap_open against ALIGN_TO_ALIGN's cost -- loses strictly, and
// it can never be the minimal state of a cell the traceback visits: the whole EB state is dropped.
template <bool BT, bool NOEB>
__device__ __forceinline__ int aff_cell_dna(int ehl, int cbl, int evu, int cbu, int cbd, int evd, int ehd, int ebd,
                                            const AffWinRow &r, const AffWinCol &c, int2 d, int go8, int &cb, int &ev,
                                            int &eh, int &eb) {
    int byte = 0;
    // FILL_EXTEND_HORIZONTAL :1765-1787 -- extend wins only when strictly cheaper
    int x = ehl + c.hx4, y = cbl + c.gopg4m1;
    eh = min(x, y);
    if (BT) byte = (x < y) ? 0 : AB_ENDH;
    // FILL_EXTEND_VERTICAL :1813-1830
    x = evu + r.vx4;
    y = cbu + r.gopge4p1;
    ev = min(x, y);
    if (BT) byte |= (x < y) ? 0 : AB_ENDV;
    if (NOEB) {
        // "x & mask" written as a multiply-add: the INT32 ALU pipe is the bottleneck, the FMA pipe has room
        const int a2 = (r.gop4 * c.gf + ehd) + d.x;
        const int a1 = (c.gop4 * r.gf + evd) + d.x;
        const int a0 = cbd + d.x + 2;
        const int ck = min(a0, min(a1, a2));
        cb = (ck & ~3) | TAG_CB;
        eb = ebd;
        if (BT) {
            const int fk = min(eh, min(ev, cb));
            byte += (fk & 3) * 4 + (ck & 3) + AB_ENDB;
        }
        return byte;
    }
    // FILL_EXTEND_BLOCK_DIAGONAL :1861-1882 / _NOBT :1837-1854
    const int bothf = r.gf * c.gf;              // 1 when both carry the gap bit
    const int dg = HIGH4 - HIGH4 * bothf;       // 0 or 4*HIGH_NUM
    if (BT) {
        const int c2 = cbd + 2;        // CB tag 1 -> EB tag 3
        eb = min(ebd, c2) + dg;        // extend and open share the addend (:1871-1872)
        byte |= (ebd < c2) ? 0 : AB_ENDB;
    } else {
        const int odg = dg + go8 * bothf;  // both ? 2*go : HIGH_NUM (:1846; its flag2 can never hold with flag)
        eb = min(ebd + dg, cbd + odg + 2);
    }
    // FILL_CLOSE_BLOCK_DIAGONAL :1923-1977, candidates tagged H 0, D 1, V 2, A 3
    const int a2 = (r.gop4 * c.gf + ehd) + d.x;
    const int a3 = ebd + d.y + max(r.gop4, c.gop4);
    const int a1 = (c.gop4 * r.gf + evd) + d.x;
    const int a0 = cbd + d.x + 2;
    const int ck = min(min(a0, a1), min(a2, a3));
    cb = (ck & ~3) | TAG_CB;
    if (BT) {
        const int fk = min(min(eh, ev), min(eb, cb));  // ASSIGN_MINIMUM :2251-2280
        byte += (fk & 3) * 4 + (ck & 3);
    }
    return byte;
}

template <int K, int G, bool BT, bool LOW, bool NOEB>
struct AffStripe {
    static constexpr int Q = 2 * K;

    // per-lane state
    int cb[Q], ev[Q], eh[Q], eb[Q];
    AffWinRow R[K];
    AffWinCol C[K + 1];
    int last_ci, last_cj;  // codes of the newest row / column in the windows
    // per-pair constants
    const uint8_t *si, *sj;  // shared memory copies
    const uint8_t *lut;      // int2 entries, rows of LUT_ROW_BYTES
    const int *prep, *get;   // 4 * prepend[c], 4 * cost[c][gap]
    int nr, nc, go4, lane, qlow;

    __device__ __forceinline__ AffWinRow make_row(int i) {
        const int ii = min(max(i, 0), nr);
        const int ci = si[ii], pi = last_ci;
        last_ci = ci;
        AffWinRow r;
        const int ge4 = get[ci];
        r.gop4 = (!(pi & 16) && (ci & 16)) ? 0 : go4;  // HAS_GAP_OPENING :1730
        r.vx4 = (i > 1 && (pi & 16) && !(ci & 16)) ? r.gop4 + ge4 : ge4;  // :2483-2485
        r.gopge4p1 = r.gop4 + ge4 + 1;
        r.lut = (ci & 15) * LUT_ROW_BYTES;
        r.gf = (ci >> 4) & 1;
        return r;
    }
    __device__ __forceinline__ AffWinCol make_col(int j) {
        const int jj = min(max(j, 0), nc);
        const int cj = sj[jj], pj = last_cj;
        last_cj = cj;
        AffWinCol c;
        const int g4 = prep[cj];
        c.gop4 = (!(pj & 16) && (cj & 16)) ? 0 : go4;
        c.hx4 = ((pj & 16) && !(cj & 16) && j != 1) ? c.gop4 + g4 : g4;  // :2454-2458
        c.gopg4m1 = c.gop4 + g4 - 1;
        c.lut = (cj & 15) * 8;
        c.gf = (cj >> 4) & 1;
        return c;
    }

    // Windows for double step u: rows i0 - m, columns j0 + n.
    __device__ __forceinline__ void init_windows(int i0, int j0) {
        last_ci = si[min(max(i0 - K, 0), nr)];
#pragma unroll
        for (int m = K - 1; m >= 0; m--) R[m] = make_row(i0 - m);
        last_cj = sj[min(max(j0 - 1, 0), nc)];
#pragma unroll
        for (int n = 0; n <= K; n++) C[n] = make_col(j0 + n);
    }
    __device__ __forceinline__ void slide_windows(int i0_new, int j0_new) {
#pragma unroll
        for (int m = K - 1; m >= 1; m--) R[m] = R[m - 1];
        R[0] = make_row(i0_new);
#pragma unroll
        for (int n = 0; n < K; n++) C[n] = C[n + 1];
        C[K] = make_col(j0_new + K);
    }

    // Row 0 / column 0 / lower spare diagonals: replace what the interior recurrence produced.
    template <bool BOUNDARY>
    __device__ __forceinline__ void fixups(int q, int i, int j, int ehl, int evu, const AffWinRow &r, const AffWinCol &c,
                                           int &ncb, int &nev, int &neh, int &neb) {
        if (LOW) {
            if (q < qlow) {  // below the stripe: the left-edge rule (:2486-2494)
                ncb = HIGH4 + TAG_CB; neh = HIGH4 + TAG_EH; neb = HIGH4 + TAG_EB;
                nev = evu + r.vx4;
            }
        }
        if (BOUNDARY) {
            if (i == 0) {
                if (j == 0) {  // :2194-2198
                    ncb = TAG_CB; neb = TAG_EB; neh = go4 + TAG_EH; nev = go4 + TAG_EV;
                } else {       // :2212-2217
                    const int rr = ehl + (c.gopg4m1 + 1 - c.gop4);
                    neh = rr; ncb = rr + TAG_CB; nev = HIGH4 + TAG_EV; neb = HIGH4 + TAG_EB;
                }
            } else if (j == 0) {  // column 0 = the left-edge cells of rows 1..39 (:2486-2494)
                ncb = HIGH4 + TAG_CB; neh = HIGH4 + TAG_EH; neb = HIGH4 + TAG_EB;
                nev = evu + r.vx4;
            }
        }
    }

    // One double step; the K direction bytes of each step come back as two 32-bit words (bytes 0-3, 4-7).
    template <bool BOUNDARY>
    __device__ __forceinline__ void double_step(int i0, int j0, uint32_t (&de)[2], uint32_t (&dod)[2]) {
        const int go8 = 2 * go4;
        // ---- even step: q = 2m, cell (i0 - m, j0 + m)
        int in_eh = __shfl_up_sync(0xffffffffu, eh[Q - 1], 1, G);
        int in_cb = __shfl_up_sync(0xffffffffu, cb[Q - 1], 1, G);
        if (lane == 0) { in_eh = HIGH4 + TAG_EH; in_cb = HIGH4 + TAG_CB; }  // the left-edge cells (:2487, :2494)
        de[0] = de[1] = dod[0] = dod[1] = 0;
#pragma unroll
        for (int m = 0; m < K; m++) {
            const int q = 2 * m;
            const int ehl = (m == 0) ? in_eh : eh[q - 1], cbl = (m == 0) ? in_cb : cb[q - 1];
            const int evu = ev[q + 1], cbu = cb[q + 1];
            const int2 d = *reinterpret_cast<const int2 *>(lut + R[m].lut + C[m].lut);
            int ncb, nev, neh, neb;
            const int byte = aff_cell_dna<BT, NOEB>(ehl, cbl, evu, cbu, cb[q], ev[q], eh[q], eb[q], R[m], C[m], d, go8, ncb, nev,
                                              neh, neb);
            fixups<BOUNDARY>(q, i0 - m, j0 + m, ehl, evu, R[m], C[m], ncb, nev, neh, neb);
            cb[q] = ncb; ev[q] = nev; eh[q] = neh; eb[q] = neb;
            if (BT) de[m >> 2] |= (uint32_t) byte << (8 * (m & 3));
        }
        // ---- odd step: q = 2m + 1, cell (i0 - m, j0 + m + 1)
        const int in_ev = __shfl_down_sync(0xffffffffu, ev[0], 1, G);
        const int in_cbu = __shfl_down_sync(0xffffffffu, cb[0], 1, G);
#pragma unroll
        for (int m = 0; m < K; m++) {
            const int q = 2 * m + 1;
            const int ehl = eh[q - 1], cbl = cb[q - 1];
            const int evu = (m == K - 1) ? in_ev : ev[q + 1], cbu = (m == K - 1) ? in_cbu : cb[q + 1];
            const int2 d = *reinterpret_cast<const int2 *>(lut + R[m].lut + C[m + 1].lut);
            int ncb, nev, neh, neb;
            const int byte = aff_cell_dna<BT, NOEB>(ehl, cbl, evu, cbu, cb[q], ev[q], eh[q], eb[q], R[m], C[m + 1], d, go8, ncb,
                                              nev, neh, neb);
            if (m == K - 1) {
                if (lane == G - 1) {  // diagonal dhi + 1: poisoned (:2531-2535)
                    ncb = HIGH4 + TAG_CB; nev = HIGH4 + TAG_EV; neh = HIGH4 + TAG_EH; neb = HIGH4 + TAG_EB;
                }
            }
            fixups<BOUNDARY>(q, i0 - m, j0 + m + 1, ehl, evu, R[m], C[m + 1], ncb, nev, neh, neb);
            cb[q] = ncb; ev[q] = nev; eh[q] = neh; eb[q] = neb;
            if (BT) dod[m >> 2] |= (uint32_t) byte << (8 * (m & 3));
        }
    }
};

template <int BL>
__device__ __forceinline__ void store_dir(uint8_t *p, const uint32_t (&v)[2]) {
    if (BL == 4) *reinterpret_cast<uint32_t *>(p) = v[0];
    else *reinterpret_cast<uint2 *>(p) = make_uint2(v[0], v[1]);
}

// seq_bytes: shared-memory bytes reserved per operand (multiple of 16, >= the longest sequence of the launch);
// nslots: depth of the staging ring (staging.cuh).
template <int K, int G, bool BT>
__global__ void __launch_bounds__(STRIPE_WARPS * 32, STRIPE_MIN_BLOCKS) aff_stripe_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm,
                                                                       const uint8_t *__restrict__ pool,
                                                                       uint8_t *__restrict__ dir, int *__restrict__ out_cost,
                                                                       int seq_bytes, int nslots, int allow_noeb, int *work_counter,
                                                                       const int *__restrict__ batch_list,
                                                                       const int *__restrict__ batch_count) {
    // batch_list != nullptr: only the batches aff_fast_kernel declined (aff_fast_kernels.cuh), *batch_count of them
    if (batch_list != nullptr && *batch_count == 0) return;
    constexpr int GPW = 32 / G;  // groups (pairs) per warp
    constexpr int Q = 2 * K;
    constexpr int BL = (K <= 4) ? 4 : 8;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *s_lut = smem;  // 16 rows of LUT_ROW_BYTES: int2 {4*cost, 4*cost - 2} for cost[(a & 15) << lcm | (b & 15)]
    int *s_prep = reinterpret_cast<int *>(smem + STRIPE_LUT_BYTES);  // 32 ints: 4 * prepend[c]
    int *s_get = s_prep + 32;                                         // 32 ints: 4 * cost[c << lcm | gap]
    StageBars *s_bar = reinterpret_cast<StageBars *>(s_get + 32);     // one ring per group
    uint8_t *s_seq = reinterpret_cast<uint8_t *>(s_bar + STRIPE_WARPS * 4);
    if (threadIdx.x < STRIPE_WARPS * GPW) StageRing<G>::init_bars(&s_bar[threadIdx.x]);
    for (int k = threadIdx.x; k < 256; k += blockDim.x) {
        const int c4 = 4 * __ldg(cm.cost + ((k >> 4) << cm.lcm) + (k & 15));
        *reinterpret_cast<int2 *>(s_lut + (k >> 4) * LUT_ROW_BYTES + (k & 15) * 8) = make_int2(c4, c4 - 2);
    }
    for (int k = threadIdx.x; k < 32; k += blockDim.x) {
        s_prep[k] = 4 * __ldg(cm.prepend + k);
        s_get[k] = 4 * __ldg(cm.cost + (k << cm.lcm) + cm.gap);
    }
    __syncthreads();

    const int warp_in_block = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
    const int grp = lane32 / G, lane = lane32 % G;
    StageRing<G> ring;
    ring.attach(&s_bar[warp_in_block * GPW + grp], s_seq + (size_t) ((warp_in_block * GPW + grp) * 2 * nslots) * seq_bytes, seq_bytes,
                nslots, lane);
    const int nbatches = (ntasks + GPW - 1) / GPW;

    int slot = 0;
    int batch = fetch_batch(work_counter, nbatches, batch_list, batch_count);
    if (batch >= 0) ring.produce_task(0, tasks, ntasks, batch * GPW + grp, pool, 16);
    while (batch >= 0) {
        int next = -1;
        if (nslots == 2) {  // the operands of the next batch travel under this one
            next = fetch_batch(work_counter, nbatches, batch_list, batch_count);
            if (next >= 0) ring.produce_task(slot ^ 1, tasks, ntasks, next * GPW + grp, pool, 16);
        }
        const int ti = batch * GPW + grp;
        const bool valid = ti < ntasks;
        Task t;
        if (valid) t = tasks[ti];
        else { t = Task{}; t.lr = 1; t.lc = 1; t.dhi = -1; t.dlo = -39; }
        const int nr = t.lr - 1, nc = t.lc - 1;
        ring.wait_full(slot);
        const uint8_t *my_seq = ring.rows(slot);

        const int d0 = t.dhi + 2 - Q * G;
        const int u_first = (-d0) >> 1;  // first double step: t = 2u + d0 in {-1, 0}
        const int u_last = valid ? ((nr + nc - d0) >> 1) : (u_first - 1);
        int u_end = u_last;
#pragma unroll
        for (int o = G; o < 32; o <<= 1) u_end = max(u_end, __shfl_xor_sync(0xffffffffu, u_end, o));
        int u_begin = u_first;
#pragma unroll
        for (int o = G; o < 32; o <<= 1) u_begin = min(u_begin, __shfl_xor_sync(0xffffffffu, u_begin, o));
        const int qlow_all = t.dlo - d0;  // diagonals below the stripe, over the whole group
        const bool low = qlow_all > 0;
        uint8_t *dbase = dir + t.dir_off;
        const int dd_f = (nc - nr) - d0, lane_f = dd_f / Q, q_f = dd_f % Q;
        int result = 0;

        auto run = [&](auto lowtag, auto noebtag) {
            constexpr bool LOW = decltype(lowtag)::value;
            constexpr bool NOEB = decltype(noebtag)::value;
            AffStripe<K, G, BT, LOW, NOEB> S;
            S.si = my_seq; S.sj = my_seq + seq_bytes;
            S.lut = s_lut; S.prep = s_prep; S.get = s_get;
            S.nr = nr; S.nc = nc; S.go4 = 4 * cm.gap_open; S.lane = lane;
            S.qlow = min(max(qlow_all - lane * Q, 0), Q);
#pragma unroll
            for (int q = 0; q < Q; q++) {
                S.cb[q] = HIGH4 + TAG_CB; S.ev[q] = HIGH4 + TAG_EV; S.eh[q] = HIGH4 + TAG_EH; S.eb[q] = HIGH4 + TAG_EB;
            }
            int u = u_begin;
            int i0 = u - lane * K, j0 = u + d0 + lane * K;
            S.init_windows(i0, j0);
            int u_b = max(G * K, 1 - d0);  // from here on every lane has i >= 1 and j >= 1 (warp-uniform maximum)
#pragma unroll
            for (int o = G; o < 32; o <<= 1) u_b = max(u_b, __shfl_xor_sync(0xffffffffu, u_b, o));
            // chunk address of (step T, this lane): tiles of 8 steps, see dir_index (common.cuh)
            auto chunk = [&](int T) {
                const int S2 = T - t.tshift;
                return dbase + (((size_t) (S2 >> 3) * G + lane) * 8 + (S2 & 7)) * BL;
            };
            auto emit = [&](const uint32_t (&de)[2], const uint32_t (&dod)[2]) {
                if (BT) {
                    const int te = 2 * u + d0;
                    if (u >= u_first && u <= u_last) {
                        if (te >= 0) store_dir<BL>(chunk(te), de);
                        if (te + 1 <= nr + nc) store_dir<BL>(chunk(te + 1), dod);
                    }
                }
                if (u == u_last && lane == lane_f) {
                    int r = 0;
#pragma unroll
                    for (int q = 0; q < Q; q++)
                        if (q == q_f) r = (NOEB ? min(S.cb[q], min(S.ev[q], S.eh[q])) : min(min(S.cb[q], S.ev[q]), min(S.eh[q], S.eb[q]))) >> 2;
                    result = r;
                }
            };
            for (; u < min(u_b, u_end + 1); u++) {
                uint32_t de[2], dod[2];
                S.template double_step<true>(i0, j0, de, dod);
                emit(de, dod);
                i0++; j0++;
                S.slide_windows(i0, j0);
            }
            for (; u <= u_end; u++) {
                uint32_t de[2], dod[2];
                S.template double_step<false>(i0, j0, de, dod);
                emit(de, dod);
                i0++; j0++;
                S.slide_windows(i0, j0);
            }
        };
        // LOW must be uniform over the warp's control flow: take the slower variant if any group needs it
        const bool any_low = __any_sync(0xffffffffu, low);
        // gap bits beyond the leading element of either operand (scanned in shared memory, 4 bytes per load)
        int gapbits = 0;
        if (valid) {
            for (int k = lane * 4; k < t.lr; k += G * 4) {
                uint32_t w = *reinterpret_cast<const uint32_t *>(my_seq + k);
                if (k == 0) w &= 0xffffff00u;
                if (k + 4 > t.lr) w &= 0xffffffffu >> (8 * (k + 4 - t.lr));
                gapbits |= (int) (w & 0x10101010u);
            }
            for (int k = lane * 4; k < t.lc; k += G * 4) {
                uint32_t w = *reinterpret_cast<const uint32_t *>(my_seq + seq_bytes + k);
                if (k == 0) w &= 0xffffff00u;
                if (k + 4 > t.lc) w &= 0xffffffffu >> (8 * (k + 4 - t.lc));
                gapbits |= (int) (w & 0x10101010u);
            }
        }
        const bool noeb = !__any_sync(0xffffffffu, gapbits != 0) && cm.gap_open > 0 && allow_noeb;
        if (any_low) run(std::true_type{}, std::false_type{});
        else if (noeb) run(std::false_type{}, std::true_type{});
        else run(std::false_type{}, std::false_type{});

        if (valid && lane == lane_f) {
            if (BT && nr == 0 && nc == 0) result = 0;
            out_cost[t.pair] = result;
        }
        ring.release(slot);  // this lane's last read of the staged operands is behind it
        if (nslots == 1) {
            next = fetch_batch(work_counter, nbatches, batch_list, batch_count);
            if (next >= 0) ring.produce_task(0, tasks, ntasks, next * GPW + grp, pool, 16);
        } else {
            slot ^= 1;
        }
        batch = next;
    }
}

// ---- host side -----------------------------------------------------------------------------------------
#ifdef POYB200_DEFINE_AFF_STRIPE  // the translation unit that owns these kernels (k_aff_stripe.cu)

template <int K, int G>
static cudaError_t stripe_launch_shape(bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir,
                                       int *cost, int sm_count, int seq_bytes, int allow_noeb, int *work_counter,
                                       const int *batch_list, const int *batch_count, cudaStream_t stream) {
    constexpr int GPW = 32 / G;
    const int nbatches = (n + GPW - 1) / GPW;
    auto kern = bt ? aff_stripe_kernel<K, G, true> : aff_stripe_kernel<K, G, false>;
    size_t smem = 0;
    int nslots = 1, per_sm = 1;
    cudaError_t e = stage_ring_config(kern, STRIPE_TABLE_BYTES, (size_t) STRIPE_WARPS * GPW * 2 * seq_bytes, STRIPE_WARPS * 32, smem,
                                      nslots, per_sm);
    if (e != cudaSuccess) return e;
    int blocks = std::min((nbatches + STRIPE_WARPS - 1) / STRIPE_WARPS, sm_count * per_sm);
    if (blocks < 1) blocks = 1;
    kern<<<blocks, STRIPE_WARPS * 32, smem, stream>>>(d_tasks, n, cm, pool, dir, cost, seq_bytes, nslots, allow_noeb, work_counter,
                                                      batch_list, batch_count);
    return cudaGetLastError();
}

cudaError_t stripe_launch(uint32_t klass, bool affine, bool bt, const Task *d_tasks, int n, DevCM cm,
                                        const uint8_t *pool, uint8_t *dir, int *cost, int sm_count, int seq_bytes,
                                        int allow_noeb, int *work_counter, const int *batch_list, const int *batch_count,
                                        cudaStream_t stream) {
    if (!affine) return cudaErrorNotSupported;
    switch (aff_shape_of(klass)) {
        case 0: return stripe_launch_shape<5, 8>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, allow_noeb, work_counter, batch_list, batch_count, stream);
        case 1: return stripe_launch_shape<6, 8>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, allow_noeb, work_counter, batch_list, batch_count, stream);
        case 2: return stripe_launch_shape<4, 16>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, allow_noeb, work_counter, batch_list, batch_count, stream);
        case 3: return stripe_launch_shape<6, 16>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, allow_noeb, work_counter, batch_list, batch_count, stream);
        case 4: return stripe_launch_shape<4, 32>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, allow_noeb, work_counter, batch_list, batch_count, stream);
        case 5: return stripe_launch_shape<6, 32>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, allow_noeb, work_counter, batch_list, batch_count, stream);
        case 6: return stripe_launch_shape<8, 32>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, allow_noeb, work_counter, batch_list, batch_count, stream);
        default: return cudaErrorInvalidValue;
    }
}

#endif  // POYB200_DEFINE_AFF_STRIPE

}  // namespace poyb200
