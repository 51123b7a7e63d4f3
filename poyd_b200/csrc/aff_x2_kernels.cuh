// Two pairs per lane group: the affine_3 fill (src/algn.c:2411-2548) of aff_fast_kernel<5, 8, true, true> on 16-bit values,
// two alignments side by side in the halves of every register.
//
// aff_fast_kernel is bound by instruction issue (profiles/README.md): per cell three minima, their adds, two compares and
// selects for the END_* flags, a table load and the packing of the direction code.  Blackwell keeps Hopper's DPX
// instructions, which do the min / add-min / min3 of this recurrence on BOTH 16-bit halves of a register at once and --
// VIMNMX.S16x2 -- also say, in two predicate outputs, which operand won: exactly the "opening <= extension" test behind
// END_HORIZONTAL / END_VERTICAL.  4 * cost + tag of a pair whose path costs stay below 5 000 fits 16 bits (500 bp DNA under
// 1 / 2 / 3 costs: < 2 100), so a lane group that carries pair A in the low halves and pair B in the high halves fills
// two alignments with (almost) the instructions of one:
//
//   per TWO cells                       aff_fast_kernel (x 2)      here
//   extension states (H, V)             2 x 10                     H: VIMNMX.S16x2, add; V: 2 adds, VIMNMX.S16x2; + 4 predicated adds = 9
//   cost[si][sj]                        2 x 2                      2 loads, 2 address adds, 1 merge               =  5
//   close-block minimum, tag-1 value    2 x 5                      VIADDMNMX.S16x2, VIMNMX3.S16x2, add, LOP3, add =  5
//   ASSIGN_MINIMUM, choice bits         2 x 4                      VIMNMX3.S16x2, 2 masks, 1 multiply-add         =  4
//   6-bit codes into the band words     2 x 1                      (split by pair: the bands stay per pair)       ~  2.4
//
// Everything else is aff_fast_kernel's: the same stripe of 80 diagonals per pair (shape (5, 8)), the same ring windows and
// unrolled blocks, the same boundary phase (aff_stripe_kernel's sweep, run for pair A and then for pair B, states
// converted to 16 bits afterwards), the same direction band per pair -- five 6-bit codes per 32-bit word, TF_DIR6 -- so
// the traceback kernel reads it unchanged, and the same results bit for bit: the only difference in value is the
// stand-in for HIGH_NUM (24 576 instead of 4 000 000), which a guard keeps above every real value (below).
//
// Pairing: a "double batch" is two consecutive batches of aff_fast_kernel (batch 2b in the low halves, 2b + 1 in the high
// halves, group by group).  A double batch is declined -- both batches appended to the list aff_fast_kernel takes next --
// when any of its eight pairs has gap bits, spare diagonals below dlo, or could exceed the 16-bit range:
//     4 * (unit * (rows + columns + slack) + gap_open) < X2_REAL_MAX,   unit = the largest entry of the cost / prepend tables.
#pragma once
#include "aff_fast_kernels.cuh"

namespace poyb200 {

constexpr int X2_HIGH = 24576;      // "no such cell": a multiple of 4 above every real value, with room for a few costs below 2^15
constexpr int X2_REAL_MAX = 20000;  // every real 4 * cost + tag stays below
constexpr int X2_SCR_INTS = 2 * FAST_SCR_INTS;
constexpr int X2_TABLE_BYTES = FAST_TABLE_BYTES + STRIPE_WARPS * 4 * STAGE_BAR_BYTES + STRIPE_WARPS * 4 * FAST_SCR_INTS * 4;

__host__ __device__ constexpr uint32_t x2_dup(int v) { return (uint32_t) v * 0x10001u; }
// a 32-bit state of the boundary phase (4 * cost + tag, HIGH4-based when the cell does not exist) as a 16-bit half
__device__ __forceinline__ uint32_t x2_narrow(int v) {
    return (uint32_t) ((v >= HIGH4 / 2) ? X2_HIGH + min(v - HIGH4, 4000) : v);
}

// min of both halves; fwA += BITS when the low half of `a` won or tied, fwB += BITS when its high half did.  One
// VIMNMX.S16x2 with two predicate outputs and two predicated adds: the pattern of __vibmin_s16x2 (crt/device_functions.hpp),
// with the predicates consumed on the spot (as returned booleans they outlive the seven predicate registers, and the
// compiler parks them in a bit mask: two more instructions each).
// X2_FLAG_FMA: the predicated adds as multiply-adds by an opaque 1 (FMA pipe; the INT32 ALU pipe is the busier one)
#ifndef X2_FLAG_FMA
#define X2_FLAG_FMA 0
#endif
// double steps per unrolled block (a divisor of K + 1 = 6).  Measured on 524 288 pairs of configs[1], device-resident
// (gpurun_out/r02x2s_*.json): 6 -> 972 GCUPS, 3 -> 1 023, 2 (with X2_FLAG_FMA) -> 1 010; X2_FLAG_FMA changes nothing (3: 1 013)
#ifndef X2_UNROLL
#define X2_UNROLL 3
#endif
template <uint32_t BITS>
__device__ __forceinline__ uint32_t x2_min_flag(uint32_t a, uint32_t b, uint32_t &fwA, uint32_t &fwB, uint32_t one) {
    uint32_t m;
#if X2_FLAG_FMA
    asm("{\n\t.reg .pred pu, pv;\n\t.reg .u16 rs0, rs1, rs2, rs3;\n\t"
        "min.s16x2 %0, %3, %4;\n\t"
        "mov.b32 {rs0, rs1}, %0;\n\t"
        "mov.b32 {rs2, rs3}, %3;\n\t"
        "setp.eq.s16 pv, rs0, rs2;\n\t"
        "setp.eq.s16 pu, rs1, rs3;\n\t"
        "@pv mad.lo.u32 %1, %6, %5, %1;\n\t"
        "@pu mad.lo.u32 %2, %6, %5, %2;\n\t}"
        : "=r"(m), "+r"(fwA), "+r"(fwB)
        : "r"(a), "r"(b), "n"(BITS), "r"(one));
#else
    (void) one;
    asm("{\n\t.reg .pred pu, pv;\n\t.reg .u16 rs0, rs1, rs2, rs3;\n\t"
        "min.s16x2 %0, %3, %4;\n\t"
        "mov.b32 {rs0, rs1}, %0;\n\t"
        "mov.b32 {rs2, rs3}, %3;\n\t"
        "setp.eq.s16 pv, rs0, rs2;\n\t"
        "setp.eq.s16 pu, rs1, rs3;\n\t"
        "@pv add.u32 %1, %1, %5;\n\t"
        "@pu add.u32 %2, %2, %5;\n\t}"
        : "=r"(m), "+r"(fwA), "+r"(fwB)
        : "r"(a), "r"(b), "n"(BITS));
#endif
    return m;
}

template <int K, int G>
struct AffX2 {
    static constexpr int Q = 2 * K, P = K + 1, BL = 4;
    // double steps per unrolled block: a whole ring turn (P) needs no register moves but is 32 KB of code, and the kernel then
    // waits for instructions (ncu: "no instruction" 0.70 stalls per issue); a divisor of P rotates the windows at the end
    static constexpr int UN = (P % X2_UNROLL == 0) ? X2_UNROLL : P;
    static_assert(K == 5, "6-bit packing is for five codes per word");

    uint32_t cb[Q], ev[Q], eh[Q];  // low half: pair A, high half: pair B.  cb = 4 * CB + 4 * gap_open (tag 0), see FAST_CELL_V2
    uint32_t Rv[P], Cv[P];         // 4 * cost[si][gap], 4 * prepend[sj], both pairs
    uint32_t RlA[P], RlB[P], ClA[P], ClB[P];  // shared address of the LUT row / byte offset of the column, per pair
    uint32_t siA, sjA, siB, sjB, tabR, tabC;  // shared addresses
    int nrA, ncA, nrB, ncB, lane;
    uint32_t keep2;                 // 0xfffcfffc
    uint32_t one;                   // 1, opaque (X2_FLAG_FMA)
    uint32_t c_d2, c_o2, cb_high2;  // 3 - go4 (state -> tag A 3 candidate), go4 - 1 (tag-1 value -> state), X2_HIGH + go4: per half
    int go4;
    // per pair
    int u_lastA, u_lastB, lane_fA, lane_fB;
    int *scrA, *scrB;  // shared, Q ints each

    __device__ __forceinline__ void load_row(int slot, int i) {
        const int2 a = lds_v2(tabR + 8 * lds_u8_seq(siA + min(i, nrA)));
        const int2 b = lds_v2(tabR + 8 * lds_u8_seq(siB + min(i, nrB)));
        Rv[slot] = (uint32_t) (b.x * 65536 + a.x);
        RlA[slot] = (uint32_t) a.y;
        RlB[slot] = (uint32_t) b.y;
    }
    __device__ __forceinline__ void load_col(int slot, int jA, int jB) {
        const int2 a = lds_v2(tabC + 8 * lds_u8_seq(sjA + min(jA, ncA)));
        const int2 b = lds_v2(tabC + 8 * lds_u8_seq(sjB + min(jB, ncB)));
        Cv[slot] = (uint32_t) (b.x * 65536 + a.x);
        ClA[slot] = (uint32_t) a.y;
        ClB[slot] = (uint32_t) b.y;
    }
    __device__ __forceinline__ void init_windows(int i0, int j0A, int j0B) {
#pragma unroll
        for (int r = -K + 1; r <= 0; r++) load_row((r + P) % P, i0 + r);
#pragma unroll
        for (int n = 0; n <= K; n++) load_col(n % P, j0A + n, j0B + n);
    }

    // One interior cell of both pairs, cell M of its step; returns the two 4-bit choice codes (low / high half) and adds
    // the END_HORIZONTAL / END_VERTICAL bits ("the opening won", ties included) to the band words of the step.
    template <int M>
    __device__ __forceinline__ uint32_t cell(uint32_t ehl, uint32_t cbl, uint32_t evu, uint32_t cbu, int q, int rs, int cs, uint32_t &fwA,
                                             uint32_t &fwB) {
        // every half is non-negative and below 2^15: plain 32-bit adds of non-negative halves never carry across
        // FILL_EXTEND_HORIZONTAL :1765-1787: opening (cbl = 4 CB + 4 go, tag 0) against extension (ehl, tag 0), then the
        // column's gap cost -- both candidates take the same addend, so the comparison is made before it
        const uint32_t neh = x2_min_flag<((uint32_t) AB_ENDH) << (6 * M)>(cbl, ehl, fwA, fwB, one) + Cv[cs];
        const uint32_t yo = cbu + Rv[rs] + x2_dup(TAG_EV);       // open vertically: tag 2
        const uint32_t ye = evu + Rv[rs];
        const uint32_t nev = x2_min_flag<((uint32_t) AB_ENDV) << (6 * M)>(yo, ye, fwA, fwB, one);  // FILL_EXTEND_VERTICAL :1813-1830
        const uint32_t d = (uint32_t) (lds_s32(RlB[rs] + ClB[cs]) * 65536 + lds_s32(RlA[rs] + ClA[cs]));  // 4 * cost[si & 15][sj & 15]
        // FILL_CLOSE_BLOCK_DIAGONAL :1923-1977 (tags A 3, V 2, H 0); c_d2 is negative per half: only the SIMD add may take it
        const uint32_t m01 = __viaddmin_s16x2(cb[q], c_d2, ev[q]);
        const uint32_t ck = __vimin3_s16x2(m01, eh[q], eh[q]) + d;
        const uint32_t ncb1 = (ck & keep2) | x2_dup(TAG_CB);     // keep2 in a register: one LOP3
        const uint32_t fk = __vimin3_s16x2(neh, nev, ncb1);      // ASSIGN_MINIMUM :2251-2280
        cb[q] = ncb1 + c_o2; ev[q] = nev; eh[q] = neh;
        return (fk & x2_dup(3)) * 4u + (ck & x2_dup(3));
    }

    // K pairs of 4-bit codes + the flag bits collected by the cells -> the band word of pair A and of pair B (code m at
    // bits 6m, as AffFast::pack)
    __device__ __forceinline__ void pack(const uint32_t (&by)[K], uint32_t &wA, uint32_t &wB) {
        const uint32_t t01 = by[1] * 64u + by[0], t23 = by[3] * 64u + by[2];  // 10 bits per half
        wA += (by[4] & 0xffffu) * (1u << 24) + ((t23 & 0xffffu) * 4096u + (t01 & 0xffffu));
        wB += (by[4] >> 16) * (1u << 24) + ((t23 >> 16) * 4096u + (t01 >> 16));
    }

    template <class T>
    static __device__ __forceinline__ void rotate(T (&a)[P]) {  // a[s] <- a[(s + UN) % P]
        if (UN == P) return;
        T t[P];
#pragma unroll
        for (int s_ = 0; s_ < P; s_++) t[s_] = a[(s_ + UN) % P];
#pragma unroll
        for (int s_ = 0; s_ < P; s_++) a[s_] = t[s_];
    }

    // One block of UN double steps starting at double step u (every lane at rows and columns >= 1), lane origin
    // (i0, j0A / j0B).  dptrA / dptrB as AffFast::block's dptr, per pair.
    __device__ __forceinline__ void block(int u, int i0, int j0A, int j0B, uint8_t *dptrA, uint8_t *dptrB) {
#pragma unroll
        for (int p = 0; p < UN; p++) {
            uint32_t deA = 0, deB = 0, doA = 0, doB = 0;
            uint32_t by[K];
            // ---- even step: q = 2m, cell (i0 + p - m, j0 + p + m)
            uint32_t in_eh = __shfl_up_sync(0xffffffffu, eh[Q - 1], 1, G);
            uint32_t in_cb = __shfl_up_sync(0xffffffffu, cb[Q - 1], 1, G);
            if (lane == 0) { in_eh = x2_dup(X2_HIGH + TAG_EH); in_cb = cb_high2; }  // the left-edge cells (:2487, :2494)
            by[0] = cell<0>(in_eh, in_cb, ev[1], cb[1], 0, (p + P) % P, p % P, deA, deB);
            by[1] = cell<1>(eh[1], cb[1], ev[3], cb[3], 2, (p - 1 + P) % P, (p + 1) % P, deA, deB);
            by[2] = cell<2>(eh[3], cb[3], ev[5], cb[5], 4, (p - 2 + P) % P, (p + 2) % P, deA, deB);
            by[3] = cell<3>(eh[5], cb[5], ev[7], cb[7], 6, (p - 3 + P) % P, (p + 3) % P, deA, deB);
            by[4] = cell<4>(eh[7], cb[7], ev[9], cb[9], 8, (p - 4 + P) % P, (p + 4) % P, deA, deB);
            pack(by, deA, deB);
            // ---- odd step: q = 2m + 1, cell (i0 + p - m, j0 + p + m + 1)
            const uint32_t in_ev = __shfl_down_sync(0xffffffffu, ev[0], 1, G);
            const uint32_t in_cbu = __shfl_down_sync(0xffffffffu, cb[0], 1, G);
            by[0] = cell<0>(eh[0], cb[0], ev[2], cb[2], 1, (p + P) % P, (p + 1) % P, doA, doB);
            by[1] = cell<1>(eh[2], cb[2], ev[4], cb[4], 3, (p - 1 + P) % P, (p + 2) % P, doA, doB);
            by[2] = cell<2>(eh[4], cb[4], ev[6], cb[6], 5, (p - 2 + P) % P, (p + 3) % P, doA, doB);
            by[3] = cell<3>(eh[6], cb[6], ev[8], cb[8], 7, (p - 3 + P) % P, (p + 4) % P, doA, doB);
            by[4] = cell<4>(eh[8], cb[8], in_ev, in_cbu, 9, (p - 4 + P) % P, (p + 5) % P, doA, doB);
            if (lane == G - 1) {  // diagonal dhi + 1: poisoned (:2531-2535)
                cb[9] = cb_high2; ev[9] = x2_dup(X2_HIGH + TAG_EV); eh[9] = x2_dup(X2_HIGH + TAG_EH);
            }
            pack(by, doA, doB);
            // ---- direction codes: one 8-byte store per pair and double step (AffFast::block)
            const int uu = u + p;
            {
                constexpr int TILE = G * 8 * BL;
                const int s2 = 2 * uu;
                const ptrdiff_t off = (ptrdiff_t) (s2 >> 3) * TILE + (s2 & 7) * BL;
                if (uu <= u_lastA) *reinterpret_cast<uint2 *>(dptrA + off) = make_uint2(deA, doA);
                if (uu <= u_lastB) *reinterpret_cast<uint2 *>(dptrB + off) = make_uint2(deB, doB);
            }
            // ---- the final cell (nr, nc) of a pair belongs to its double step u_last: park the minima of the owning lane
            if (__any_sync(0xffffffffu, uu == u_lastA || uu == u_lastB)) {
                if (uu == u_lastA && lane == lane_fA) {
#pragma unroll
                    for (int q = 0; q < Q; q++)
                        scrA[q] = min(min((int) (cb[q] & 0xffffu) - go4, (int) (eh[q] & 0xffffu)), (int) (ev[q] & 0xffffu));
                }
                if (uu == u_lastB && lane == lane_fB) {
#pragma unroll
                    for (int q = 0; q < Q; q++)
                        scrB[q] = min(min((int) (cb[q] >> 16) - go4, (int) (eh[q] >> 16)), (int) (ev[q] >> 16));
                }
            }
            // ---- windows: row i0 + p + 1 and column j0 + p + K + 1 enter
            load_row((p + 1) % P, i0 + p + 1);
            load_col((p + K + 1) % P, j0A + p + K + 1, j0B + p + K + 1);
        }
        if (UN != P) {  // the next block expects its rows / columns in the slots this one started with
            rotate(Rv); rotate(Cv); rotate(RlA); rotate(RlB); rotate(ClA); rotate(ClB);
        }
    }
};

// As aff_fast_kernel; ntasks tasks = nbatches batches of GPW, taken two batches at a time.  max_unit4 = 4 * the largest entry
// of the cost rows / prepend / gap columns the fast path reads (host-computed).  slow_list / slow_count receive the BATCH
// indices (aff_fast_kernel's numbering) this kernel declines.
template <int K, int G>
__global__ void __launch_bounds__(STRIPE_WARPS * 32, FAST_MIN_BLOCKS)
    aff_x2_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm, const uint8_t *__restrict__ pool, uint8_t *__restrict__ dir,
                  int *__restrict__ out_cost, int seq_bytes, int nslots, int *work_counter, int *slow_list, int *slow_count,
                  int max_unit4, uint32_t keep2, uint32_t one) {
    // keep2 = 0xfffcfffc arrives as an argument so that it lives in a register (a LOP3 takes one immediate)
    constexpr int GPW = 32 / G;
    constexpr int Q = 2 * K;
    constexpr bool BT = true;
    using S_t = AffX2<K, G>;
    constexpr int BL = S_t::BL, P = S_t::P;
    extern __shared__ __align__(16) uint8_t smem[];
    int2 *s_tabR = reinterpret_cast<int2 *>(smem);
    int2 *s_tabC = s_tabR + 256;
    int *s_lut = reinterpret_cast<int *>(s_tabC + 256);
    // tables of the boundary phase (AffStripe, stripe_kernels.cuh)
    uint8_t *s_lut2 = smem + 2 * 256 * 8 + 16 * FAST_LUT_ROW;
    int *s_prep = reinterpret_cast<int *>(s_lut2 + STRIPE_LUT_BYTES);
    int *s_get = s_prep + 32;
    StageBars *s_bar = reinterpret_cast<StageBars *>(s_get + 32);        // two staging rings per group (pair A, pair B)
    int *s_scr = reinterpret_cast<int *>(s_bar + STRIPE_WARPS * 4 * 2);  // X2_SCR_INTS per group
    uint8_t *s_seq = reinterpret_cast<uint8_t *>(s_scr + STRIPE_WARPS * 4 * X2_SCR_INTS);
    if (threadIdx.x < STRIPE_WARPS * GPW * 2) StageRing<G>::init_bars(&s_bar[threadIdx.x]);
    for (int k = threadIdx.x; k < 256; k += blockDim.x) {
        const int c4 = 4 * __ldg(cm.cost + ((k >> 4) << cm.lcm) + (k & 15));
        s_lut[fast_lut_slot(k >> 4) * (FAST_LUT_ROW / 4) + fast_lut_slot(k)] = c4;
        *reinterpret_cast<int2 *>(s_lut2 + (k >> 4) * LUT_ROW_BYTES + (k & 15) * 8) = make_int2(c4, c4 - 2);
        s_tabR[k] = make_int2(4 * __ldg(cm.cost + ((k & 31) << cm.lcm) + cm.gap),
                              (int) (smem_u32(s_lut) + fast_lut_slot(k) * FAST_LUT_ROW));
        s_tabC[k] = make_int2(4 * __ldg(cm.prepend + (k & 31)), fast_lut_slot(k) * 4);
    }
    for (int k = threadIdx.x; k < 32; k += blockDim.x) {
        s_prep[k] = 4 * __ldg(cm.prepend + k);
        s_get[k] = 4 * __ldg(cm.cost + (k << cm.lcm) + cm.gap);
    }
    __syncthreads();

    const int warp_in_block = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
    const int grp = lane32 / G, lane = lane32 % G;
    const int op_stride = seq_bytes + fast_operand_pad(K, G);
    const int go4 = 4 * cm.gap_open;
    const int gidx = warp_in_block * GPW + grp;
    StageRing<G> ring[2];
#pragma unroll
    for (int h = 0; h < 2; h++)
        ring[h].attach(&s_bar[gidx * 2 + h], s_seq + (size_t) ((gidx * 2 + h) * 2 * nslots) * op_stride, op_stride, nslots, lane);
    int *my_scr = s_scr + gidx * X2_SCR_INTS;
    const int nbatches = (ntasks + GPW - 1) / GPW;
    const int ndouble = (nbatches + 1) / 2;

    int slot = 0;
    int batch = fetch_batch(work_counter, ndouble, nullptr, nullptr);
    if (batch >= 0) {
#pragma unroll
        for (int h = 0; h < 2; h++) ring[h].produce_task(0, tasks, ntasks, (2 * batch + h) * GPW + grp, pool, 16);
    }
    while (batch >= 0) {
        int next = -1;
        if (nslots == 2) {  // the operands of the next double batch travel under this one
            next = fetch_batch(work_counter, ndouble, nullptr, nullptr);
            if (next >= 0) {
#pragma unroll
                for (int h = 0; h < 2; h++) ring[h].produce_task(slot ^ 1, tasks, ntasks, (2 * next + h) * GPW + grp, pool, 16);
            }
        }
        Task t[2];
        bool valid[2];
        int nr[2], nc[2], d0[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int ti = (2 * batch + h) * GPW + grp;
            valid[h] = ti < ntasks;
            if (valid[h]) t[h] = tasks[ti];
            else { t[h] = Task{}; t[h].lr = 1; t[h].lc = 1; t[h].dhi = -1; t[h].dlo = -39; }
            nr[h] = t[h].lr - 1; nc[h] = t[h].lc - 1;
            d0[h] = t[h].dhi + 2 - Q * G;
        }
        ring[0].wait_full(slot);
        ring[1].wait_full(slot);
        uint8_t *seq[2] = {ring[0].rows(slot), ring[1].rows(slot)};
      do {  // one pass; `break` hands both batches to aff_fast_kernel
        bool decline = false;
        int gapbits = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            // spare diagonals below dlo need the left-edge rule inside the stripe: not here
            decline |= valid[h] && (t[h].dlo - d0[h] > 0);
            // 16-bit range: no cell of the stripe, real or past the ends of the operands, may reach X2_REAL_MAX
            decline |= valid[h] && (max_unit4 * (t[h].lr + t[h].lc + 2 * Q * G + 4 * P) + 2 * go4 + 64 >= X2_REAL_MAX);
            // gap bits beyond the leading element of either operand (scanned in shared memory, 4 bytes per load)
            if (valid[h]) {
                for (int k = lane * 4; k < t[h].lr; k += G * 4) {
                    uint32_t w = *reinterpret_cast<const volatile uint32_t *>(seq[h] + k);
                    if (k == 0) w &= 0xffffff00u;
                    if (k + 4 > t[h].lr) w &= 0xffffffffu >> (8 * (k + 4 - t[h].lr));
                    gapbits |= (int) (w & 0x10101010u);
                }
                for (int k = lane * 4; k < t[h].lc; k += G * 4) {
                    uint32_t w = *reinterpret_cast<const volatile uint32_t *>(seq[h] + op_stride + k);
                    if (k == 0) w &= 0xffffff00u;
                    if (k + 4 > t[h].lc) w &= 0xffffffffu >> (8 * (k + 4 - t[h].lc));
                    gapbits |= (int) (w & 0x10101010u);
                }
            }
        }
        if (__any_sync(0xffffffffu, decline || gapbits != 0)) {
            if (lane32 == 0) {
                const int n2 = (2 * batch + 1 < nbatches) ? 2 : 1;
                const int at = atomicAdd(slow_count, n2);
                slow_list[at] = 2 * batch;
                if (n2 == 2) slow_list[at + 1] = 2 * batch + 1;
            }
            break;
        }

        int u_first[2], u_last[2], sbase[2], dd_f[2], lane_f[2];
        uint8_t *dbase[2];
        int u_end, u_begin, u_b;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            u_first[h] = (-d0[h]) >> 1;  // first double step: t = 2u + d0 in {-1, 0}
            u_last[h] = valid[h] ? ((nr[h] + nc[h] - d0[h]) >> 1) : (u_first[h] - 1);
            sbase[h] = (2 * u_first[h]) & ~7;  // d0 + sbase == t.tshift
            dbase[h] = dir + t[h].dir_off;
            dd_f[h] = (nc[h] - nr[h]) - d0[h];
            lane_f[h] = dd_f[h] / Q;
        }
        u_end = max(u_last[0], u_last[1]);
        u_begin = min(u_first[0], u_first[1]);
        u_b = max(G * K, 1 - min(d0[0], d0[1]));  // from here on every lane has i >= 1 and j >= 1 in both pairs
#pragma unroll
        for (int o = G; o < 32; o <<= 1) {
            u_end = max(u_end, __shfl_xor_sync(0xffffffffu, u_end, o));
            u_begin = min(u_begin, __shfl_xor_sync(0xffffffffu, u_begin, o));
            u_b = max(u_b, __shfl_xor_sync(0xffffffffu, u_b, o));
        }
        S_t S;
        int u = u_begin;
        // ---- boundary phase: row 0, column 0 and the cells before them, with the stripe kernel's own sweep, pair A then B
        // (unrolled: with a rolled loop here the compiler no longer keeps a converged-warp copy of the blocks below, and every
        // shuffle of the main phase becomes a WARPSYNC.COLLECTIVE sequence of seven instructions)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int nrh = h ? nr[1] : nr[0], nch = h ? nc[1] : nc[0], d0h = h ? d0[1] : d0[0];
            const int u_firsth = h ? u_first[1] : u_first[0], u_lasth = h ? u_last[1] : u_last[0], sbaseh = h ? sbase[1] : sbase[0];
            const int lane_fh = h ? lane_f[1] : lane_f[0];
            uint8_t *dbaseh = h ? dbase[1] : dbase[0];
            uint8_t *seqh = h ? seq[1] : seq[0];
            int *scrh = my_scr + h * FAST_SCR_INTS;
            u = u_begin;
            int i0 = u - lane * K, j0 = u + d0h + lane * K;
            AffStripe<K, G, BT, false, true> A;
            A.si = seqh; A.sj = seqh + op_stride;
            A.lut = s_lut2; A.prep = s_prep; A.get = s_get;
            A.nr = nrh; A.nc = nch; A.go4 = go4; A.lane = lane; A.qlow = 0;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                A.cb[q] = HIGH4 + TAG_CB; A.ev[q] = HIGH4 + TAG_EV; A.eh[q] = HIGH4 + TAG_EH; A.eb[q] = HIGH4 + TAG_EB;
            }
            A.init_windows(i0, j0);
            for (; u < u_b && u <= u_end; u++) {
                uint32_t de[2], dod[2];
                A.template double_step<true>(i0, j0, de, dod);
                const int te = 2 * u + d0h, s2 = 2 * u - sbaseh;  // s2 = te - tshift
                if (u >= u_firsth && u <= u_lasth) {
                    uint32_t we = 0, wo = 0;  // the stripe sweep produced bytes: repack five codes into one word
#pragma unroll
                    for (int m = 0; m < K; m++) {
                        we |= ((de[m >> 2] >> (8 * (m & 3))) & 63u) << (6 * m);
                        wo |= ((dod[m >> 2] >> (8 * (m & 3))) & 63u) << (6 * m);
                    }
                    de[0] = we;
                    dod[0] = wo;
                    if (te >= 0) store_dir<BL>(dbaseh + (((size_t) (s2 >> 3) * G + lane) * 8 + (s2 & 7)) * BL, de);
                    if (te + 1 <= nrh + nch) store_dir<BL>(dbaseh + (((size_t) ((s2 + 1) >> 3) * G + lane) * 8 + ((s2 + 1) & 7)) * BL, dod);
                }
                if (u == u_lasth && lane == lane_fh) {
#pragma unroll
                    for (int q = 0; q < Q; q++) scrh[q] = min(min(A.cb[q], A.eh[q]), A.ev[q]);
                }
                i0++; j0++;
                A.slide_windows(i0, j0);
            }
            // 32-bit states -> this pair's halves (cb: tag 1 -> tag 0, plus the gap opening; see AffX2::cb)
#pragma unroll
            for (int q = 0; q < Q; q++) {
                const uint32_t c = x2_narrow(A.cb[q] - TAG_CB + go4), v = x2_narrow(A.ev[q]), e = x2_narrow(A.eh[q]);
                if (h == 0) { S.cb[q] = c; S.ev[q] = v; S.eh[q] = e; }
                else { S.cb[q] |= c << 16; S.ev[q] |= v << 16; S.eh[q] |= e << 16; }
            }
        }
        if (u <= u_end) {
            const int i0 = u - lane * K;
            const int j0A = u + d0[0] + lane * K, j0B = u + d0[1] + lane * K;
            S.siA = smem_u32(seq[0]); S.sjA = smem_u32(seq[0] + op_stride);
            S.siB = smem_u32(seq[1]); S.sjB = smem_u32(seq[1] + op_stride);
            S.tabR = smem_u32(s_tabR); S.tabC = smem_u32(s_tabC);
            S.nrA = nr[0]; S.ncA = nc[0]; S.nrB = nr[1]; S.ncB = nc[1]; S.lane = lane;
            S.go4 = go4; S.keep2 = keep2; S.one = one;
            S.c_d2 = x2_dup((3 - go4) & 0xffff); S.c_o2 = x2_dup(go4 - TAG_CB); S.cb_high2 = x2_dup(X2_HIGH + go4);
            S.u_lastA = u_last[0]; S.u_lastB = u_last[1]; S.lane_fA = lane_f[0]; S.lane_fB = lane_f[1];
            S.scrA = my_scr; S.scrB = my_scr + FAST_SCR_INTS;
            S.init_windows(i0, j0A, j0B);
            // local step of (double step u, even half) is 2u - sbase: fold the per-pair part into the pointer
            uint8_t *dptrA = dbase[0] + (ptrdiff_t) lane * 8 * BL - (ptrdiff_t) (sbase[0] >> 3) * (G * 8 * BL);
            uint8_t *dptrB = dbase[1] + (ptrdiff_t) lane * 8 * BL - (ptrdiff_t) (sbase[1] >> 3) * (G * 8 * BL);
            int i = i0, jA = j0A, jB = j0B;
            for (; u <= u_end; u += S_t::UN, i += S_t::UN, jA += S_t::UN, jB += S_t::UN) S.block(u, i, jA, jB, dptrA, dptrB);
        }

#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (valid[h] && lane == lane_f[h]) {
                int result = my_scr[h * FAST_SCR_INTS + dd_f[h] % Q] >> 2;  // parked by this very lane
                if (nr[h] == 0 && nc[h] == 0) result = 0;
                out_cost[t[h].pair] = result;
            }
        }
      } while (0);
        __syncwarp();
        ring[0].release(slot);  // this lane's last read of the staged operands is behind it
        ring[1].release(slot);
        if (nslots == 1) {
            next = fetch_batch(work_counter, ndouble, nullptr, nullptr);
            if (next >= 0) {
#pragma unroll
                for (int h = 0; h < 2; h++) ring[h].produce_task(0, tasks, ntasks, (2 * next + h) * GPW + grp, pool, 16);
            }
        } else {
            slot ^= 1;
        }
        batch = next;
    }
}

#ifdef POYB200_DEFINE_AFF_FAST  // the translation unit that owns these kernels (k_aff_fast.cu)
// True when one staging slot for the CTA's 2 x 16 pairs of operands fits the shared memory of an SM, and pairs that long
// could pass the 16-bit guard at all (api.cu asks before every launch: a chunk of long pairs skips this kernel).
bool x2_usable(int seq_bytes, int max_unit4, int gap_open) {
    constexpr int K = 5, G = 8, GPW = 32 / G;
    const size_t ring1 = (size_t) STRIPE_WARPS * GPW * 2 * 2 * (seq_bytes + fast_operand_pad(K, G));
    if (X2_TABLE_BYTES + ring1 > (size_t) 200 * 1024) return false;
    if ((long long) max_unit4 * 64 >= X2_REAL_MAX) return false;
    // a cell that does not exist is worth X2_HIGH (+ the gap opening, + one or two table entries, + a tag): that must stay
    // inside 15 bits, and inside the 4000 x2_narrow() keeps of a HIGH_NUM-based value of the boundary phase
    return gap_open >= 0 && 4ll * gap_open + 2ll * max_unit4 + 16 < 4000;
}

// Shape (5, 8) with the 6-bit band only (the tasks the planner flagged TF_DIR6).
cudaError_t x2_launch(const Task *d_tasks, int n, DevCM cm, int max_unit4, const uint8_t *pool, uint8_t *dir, int *cost, int sm_count,
                      int seq_bytes, int *work_counter, int *slow_list, int *slow_count, cudaStream_t stream) {
    constexpr int K = 5, G = 8, GPW = 32 / G;
    const int nbatches = (n + GPW - 1) / GPW, ndouble = (nbatches + 1) / 2;
    auto kern = aff_x2_kernel<K, G>;
    size_t smem = 0;
    int nslots = 1, per_sm = 1;
    cudaError_t e = stage_ring_config(kern, X2_TABLE_BYTES, (size_t) STRIPE_WARPS * GPW * 2 * 2 * (seq_bytes + fast_operand_pad(K, G)),
                                      STRIPE_WARPS * 32, smem, nslots, per_sm);
    if (e != cudaSuccess) return e;
    int blocks = std::min((ndouble + STRIPE_WARPS - 1) / STRIPE_WARPS, sm_count * per_sm);
    if (blocks < 1) blocks = 1;
    kern<<<blocks, STRIPE_WARPS * 32, smem, stream>>>(d_tasks, n, cm, pool, dir, cost, seq_bytes, nslots, work_counter, slow_list, slow_count,
                                                      max_unit4, 0xfffcfffcu, 1u);
    return cudaGetLastError();
}
#endif  // POYB200_DEFINE_AFF_FAST

}  // namespace poyb200
