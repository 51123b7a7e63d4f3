// Fast path of the affine stripe sweep (src/algn.c:2411-2548): pairs whose operands carry no gap bit beyond the leading
// element, gap_open > 0, stripe flush with dlo (no spare diagonals).  Same stripe formulation, same direction bytes and
// same costs as aff_stripe_kernel (stripe_kernels.cuh) -- that kernel's NOEB variant is the specification -- but built
// for instruction count:
//
//   * the block-diagonal state is gone (see the NOEB note in stripe_kernels.cuh) and with it every per-row / per-column
//     gap-opening special case: a row needs {4 * cost[si][gap], address of its LUT row}, a column
//     {4 * prepend[sj], byte offset inside a LUT row}, both fetched from 256-entry shared tables with one 8-byte load;
//   * the row / column windows do not slide.  They are rings of K + 1 slots and the sweep is unrolled in blocks of
//     K + 1 double steps, so every slot index is a compile-time constant and no register is ever copied;
//   * by the choice of Task::tshift the two steps of a double step are neighbours inside a tile of the direction band
//     and the tile phase depends on the double-step counter only: one 16-byte store per double step, at an offset
//     that is the same for the whole warp;
//   * the middle of the sweep (every lane inside the matrix, every pair of the warp still running) runs blocks without
//     any position test; the first and last few blocks run the same code with the tests switched on (EDGE).
//
// Batches that do not qualify are appended to a list and taken by aff_stripe_kernel right after, on the same stream.
#pragma once
#include "stripe_kernels.cuh"

namespace poyb200 {

constexpr int FAST_LUT_ROW = 20 * 4;                        // 16 ints + 4 pad
// LUT slot of a 4-bit code: A, C, G, T (1, 2, 4, 8) take slots 0..3, so with rows 20 words apart the 16 combinations
// of unambiguous bases sit in 16 different banks; the other codes follow.
__host__ __device__ constexpr int fast_lut_slot(int code) {
    // slots {4, 0, 1, 5, 2, 6, 7, 8, 3, 9, 10, 11, 12, 13, 14, 15} for codes 0..15, one nibble each
    return (int) ((0xfedcba9387625104ull >> (4 * (code & 15))) & 15);
}
// tabR, tabC, LUT of the ring blocks; LUT, prepend and gap tables of the boundary phase; one mbarrier per group
constexpr int FAST_SCR_INTS = 16;  // >= 2 K
constexpr int FAST_TABLE_BYTES =
    2 * 256 * 8 + 16 * FAST_LUT_ROW + STRIPE_LUT_BYTES + 64 * 4 + STRIPE_WARPS * 4 * STAGE_BAR_BYTES + STRIPE_WARPS * 4 * FAST_SCR_INTS * 4;
// The unchecked blocks read codes past an operand's end (rows up to Q G / 2 + D - 2 past it, columns up to G K + D - 1):
// every staged operand gets that much private slack, so the stray reads never touch another warp's buffers.
__host__ __device__ constexpr int fast_operand_pad(int K, int G) { return (K * G + 8 + 15) & ~15; }
#ifndef FAST_MIN_BLOCKS
#define FAST_MIN_BLOCKS 3
#endif

__device__ __forceinline__ int lds_s32(uint32_t a) {
    int v;
    asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int2 lds_v2(uint32_t a) {
    int2 v;
    asm("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
// a * one + b with `one` an opaque register holding 1: an add the assembler must issue on the FMA pipe (the INT32 ALU
// pipe is the busier one in the unchecked block)
__device__ __forceinline__ int fma_add(int a, int one, int b) {
#ifdef FAST_X_NOFMA
    (void) one;
    return a + b;
#else
    int v;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(v) : "r"(a), "r"(one), "r"(b));
    return v;
#endif
}
// x += bits when a == b: a compare and ONE predicated add (the bits are clear in x, so the add is an OR that may go to
// either pipe) instead of compare + select + combine
template <int BITS>
__device__ __forceinline__ void add_if_eq(int &x, int a, int b) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\t@p add.s32 %0, %0, %3;\n\t}" : "+r"(x) : "r"(a), "r"(b), "n"(BITS));
}
// min(a + b, c) in one instruction (VIADDMNMX)
__device__ __forceinline__ int add_min(int a, int b, int c) { return __viaddmin_s32(a, b, c); }
// FAST_CELL_V2 = 1: an A/B build of the cell with fewer instructions (20 instead of 24 per cell) but 14 of them on the
// INT32 ALU pipe against 11.7 -- both pipes issue one warp instruction every two cycles, so the busier pipe decides and the
// default stays 0 (see AffFast::cell)
#ifndef FAST_CELL_V2
#define FAST_CELL_V2 0
#endif
// staged operands change from pair to pair: volatile, so the load is neither hoisted nor merged across pairs
__device__ __forceinline__ int lds_u8_seq(uint32_t a) {
    int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// D6: the K = 5 direction codes of a step go into ONE 32-bit word, 6 bits each -- END_BLOCK, which is set in every cell of a
// pair without gap bits, is left out and put back by the reader (TF_DIR6) -- so the band is half as large.
template <int K, int G, bool BT, bool D6 = false>
struct AffFast {
    // P = ring slots = double steps per unrolled block.  K + 1 columns are live in a step and the next row / column
    // is fetched one step ahead into the slot that just died, so K + 1 slots do.  Instruction-cache footprint decides the
    // speed of this kernel: a block is 6 double steps x 10 cells x ~26 instructions x 16 B = 25 KB for K = 5, and
    // nothing else may be large -- a first version whose boundary phases were unrolled the same way (88 KB per block)
    // ran at half the speed with 57 % "no instruction" stalls (profiles/).
    static_assert(!D6 || K == 5, "6-bit packing is for five codes per word");
    static constexpr int Q = 2 * K, P = K + 1, BL = (K <= 4 || D6) ? 4 : 8;

    int cb[Q], ev[Q], eh[Q];
    int Rv[P], Cv[P];       // 4 * cost[si][gap], 4 * prepend[sj]
    uint32_t Rl[P], Cl[P];  // shared address of the LUT row, byte offset of the column
    uint32_t si, sj, tabR, tabC;  // shared addresses
    int nr, nc, lane, keep;
    int one;       // 1, opaque (fma_add)
    int c_h, c_v;  // go4 (CB tag 0 -> EH tag 0), go4 + 2 (-> EV tag 2): registers, so the adds stay two-input
    int c_d, c_o;  // FAST_CELL_V2: 3 - go4 (state -> tag A 3 candidate), go4 - 1 (tag-1 close-block value -> state)
    // per pair
    int u_last, lane_f;
    int *scr;  // shared, Q ints per group

    // the close-block state of a cell that does not exist (left edge :2487-2494, diagonal dhi + 1 :2531-2535)
    __device__ __forceinline__ int cb_high() const { return FAST_CELL_V2 ? HIGH4 + c_h : HIGH4; }
    // 4 * CB with tag 0 out of the state
    __device__ __forceinline__ int cb_plain(int q) const { return FAST_CELL_V2 ? cb[q] - c_h : cb[q]; }

    // Rows / columns past the end of an operand (a pair that finished while others of the warp still run, or the last
    // lanes of a stripe that overhangs the matrix) read the last code again: such cells are never used.
    __device__ __forceinline__ void load_row(int slot, int i) {
        const int2 e = lds_v2(tabR + 8 * lds_u8_seq(si + min(i, nr)));
        Rv[slot] = e.x;
        Rl[slot] = (uint32_t) e.y;
    }
    __device__ __forceinline__ void load_col(int slot, int j) {
        const int2 e = lds_v2(tabC + 8 * lds_u8_seq(sj + min(j, nc)));
        Cv[slot] = e.x;
        Cl[slot] = (uint32_t) e.y;
    }
    // Windows as block entry expects them: rows i0-K+1 .. i0 in slots r mod P, columns j0 .. j0+K in slots n mod P
    // (all rows and columns >= 1 here).
    __device__ __forceinline__ void init_windows(int i0, int j0) {
#pragma unroll
        for (int r = -K + 1; r <= 0; r++) load_row((r + P) % P, i0 + r);
#pragma unroll
        for (int n = 0; n <= K; n++) load_col(n % P, j0 + n);
    }

    // The close-block state.
    //   FAST_CELL_V2: cb[] = 4 * CB + 4 * gap_open, tag 0 -- the value BOTH gap openings start from -- so
    //     * an extension state is one add and one fused add-min: min(ehl, cb_l + go) + Cv = min(ehl + Cv, (cb_l + go) + Cv);
    //     * "the opening won" (END_HORIZONTAL / END_VERTICAL, ties included: `ehl < t` is false) is `new state == opening
    //       candidate`, one compare whose predicate guards one add into the direction code -- no select, no combine;
    //     * the tag-1 close-block value of ASSIGN_MINIMUM is (ck & ~3) | 1, one LOP3, and the next state one add from it.
    //   otherwise (round 1 / round 2a cell, kept for A/B builds): cb[] = 4 * CB with tag 0, its consumers add their tag with
    //   the constant they add anyway, and the choice bits of the close-block minimum come out as a subtraction.
    __device__ __forceinline__ int cell(int ehl, int cbl, int evu, int cbu, int q, int rs, int cs) {
#if FAST_CELL_V2
        const int xo = cbl + Cv[cs];                    // open horizontally: tag 0
        const int neh = add_min(ehl, Cv[cs], xo);       // FILL_EXTEND_HORIZONTAL :1765-1787
        const int yo = cbu + Rv[rs] + TAG_EV;           // open vertically: tag 2
        const int nev = add_min(evu, Rv[rs], yo);       // FILL_EXTEND_VERTICAL :1813-1830
        const int d = lds_s32(Rl[rs] + Cl[cs]);         // 4 * cost[si & 15][sj & 15]
        const int ck = min(min(cb[q] + c_d, ev[q]), eh[q]) + d;  // FILL_CLOSE_BLOCK_DIAGONAL :1923-1977 (tags A 3, V 2, H 0)
        const int ncb1 = (ck & keep) | TAG_CB;
        int byte = 0;
        if (BT) {
            const int fk = min(min(neh, nev), ncb1);    // ASSIGN_MINIMUM :2251-2280
            constexpr int EB = D6 ? 0 : AB_ENDB;
            byte = ((fk * 4 + (ck & 3)) & 15) | EB;
            add_if_eq<AB_ENDH>(byte, neh, xo);
            add_if_eq<AB_ENDV>(byte, nev, yo);
        }
        cb[q] = ncb1 + c_o; ev[q] = nev; eh[q] = neh;
        return byte;
#else
        const int t = cbl + c_h, t2 = cbu + c_v;
        const int neh = min(ehl, t) + Cv[cs];     // FILL_EXTEND_HORIZONTAL :1765-1787
        const int nev = min(evu, t2) + Rv[rs];    // FILL_EXTEND_VERTICAL :1813-1830
        const int d = lds_s32(Rl[rs] + Cl[cs]);   // 4 * cost[si & 15][sj & 15]
        const int ck = min(min(cb[q] + 3, ev[q]), eh[q]) + d;  // FILL_CLOSE_BLOCK_DIAGONAL :1923-1977 (tags A 3, V 2, H 0)
        const int ncb = ck & keep;
        int byte = 0;
        if (BT) {
            const int fk = min(min(neh, nev), fma_add(ncb, one, one));  // ASSIGN_MINIMUM :2251-2280 (one = TAG_CB)
            constexpr int EB = D6 ? 0 : AB_ENDB;
            const int flags = fma_add((ehl < t) ? EB : (EB | AB_ENDH), one, (evu < t2) ? 0 : AB_ENDV);
            byte = ((fk * 4 + (ck - ncb)) & 15) | flags;
        }
        cb[q] = ncb; ev[q] = nev; eh[q] = neh;
        return byte;
#endif
    }

    // K direction bytes -> two words, by multiply-adds (FMA pipe)
    __device__ __forceinline__ void pack(const int (&by)[K], uint32_t (&w)[2]) {
        if (D6) {
            int a = by[K - 1];
#pragma unroll
            for (int m = K - 2; m >= 0; m--) a = a * 64 + by[m];
            w[0] = (uint32_t) a;
            return;
        }
        constexpr int N0 = K < 4 ? K : 4;
        int a = by[N0 - 1];
#pragma unroll
        for (int m = N0 - 2; m >= 0; m--) a = a * 256 + by[m];
        w[0] = (uint32_t) a;
        if (K > 4) {
            int b = by[K - 1];
#pragma unroll
            for (int m = K - 2; m >= 4; m--) b = b * 256 + by[m];
            w[1] = (uint32_t) b;
        }
    }

    // One block of P double steps starting at double step u (every lane at rows and columns >= 1), lane origin
    // (i0, j0).  dptr = the address this lane's chunk of local step 0 has once the pair's tile origin is folded in.
    __device__ __forceinline__ void block(int u, int i0, int j0, uint8_t *dptr) {
#pragma unroll
        for (int p = 0; p < P; p++) {
            uint32_t de[2] = {0, 0}, dod[2] = {0, 0};
            int by[K];
            // ---- even step: q = 2m, cell (i0 + p - m, j0 + p + m)
            int in_eh = __shfl_up_sync(0xffffffffu, eh[Q - 1], 1, G);
            int in_cb = __shfl_up_sync(0xffffffffu, cb[Q - 1], 1, G);
            if (lane == 0) { in_eh = HIGH4 + TAG_EH; in_cb = cb_high(); }  // the left-edge cells (:2487, :2494)
#pragma unroll
            for (int m = 0; m < K; m++) {
                const int q = 2 * m;
                const int ehl = (m == 0) ? in_eh : eh[q - 1], cbl = (m == 0) ? in_cb : cb[q - 1];
                by[m] = cell(ehl, cbl, ev[q + 1], cb[q + 1], q, (p - m + P) % P, (p + m) % P);
            }
            if (BT) pack(by, de);
            // ---- odd step: q = 2m + 1, cell (i0 + p - m, j0 + p + m + 1)
            const int in_ev = __shfl_down_sync(0xffffffffu, ev[0], 1, G);
            const int in_cbu = __shfl_down_sync(0xffffffffu, cb[0], 1, G);
#pragma unroll
            for (int m = 0; m < K; m++) {
                const int q = 2 * m + 1;
                const int evu = (m == K - 1) ? in_ev : ev[q + 1], cbu = (m == K - 1) ? in_cbu : cb[q + 1];
                by[m] = cell(eh[q - 1], cb[q - 1], evu, cbu, q, (p - m + P) % P, (p + m + 1) % P);
                if (m == K - 1) {
                    if (lane == G - 1) {  // diagonal dhi + 1: poisoned (:2531-2535)
                        cb[q] = cb_high(); ev[q] = HIGH4 + TAG_EV; eh[q] = HIGH4 + TAG_EH;
                    }
                }
            }
            if (BT) pack(by, dod);
            // ---- direction bytes.  Steps 2(u+p) and 2(u+p)+1 are neighbours inside a tile (Task::tshift): one store for
            // both, at an offset that depends on u only.  A pair that has finished stores nothing; the odd half of a
            // pair's last double step may lie past its last anti-diagonal, still inside the band's last tile.
            const int uu = u + p;
            if (BT) {
                constexpr int TILE = G * 8 * BL;
                const int s2 = 2 * uu;
                uint8_t *dst = dptr + (ptrdiff_t) (s2 >> 3) * TILE + (s2 & 7) * BL;
                if (uu <= u_last) {
                    if (BL == 4) *reinterpret_cast<uint2 *>(dst) = make_uint2(de[0], dod[0]);
                    else *reinterpret_cast<uint4 *>(dst) = make_uint4(de[0], de[1], dod[0], dod[1]);
                }
            }
            // ---- the final cell (nr, nc) belongs to double step u_last: the cost of the alignment (:2540-2547).
            // The lane that owns its diagonal parks the minima of all its slots in shared memory and picks the slot
            // after the sweep (selecting the slot here, by index, would push the state arrays to local memory).
            if (__any_sync(0xffffffffu, uu == u_last)) {
                if (uu == u_last && lane == lane_f) {
#pragma unroll
                    for (int q = 0; q < Q; q++) scr[q] = min(min(cb_plain(q), eh[q]), ev[q]);
                }
            }
            // ---- windows: row i0 + p + 1 and column j0 + p + K + 1 enter
            load_row((p + 1) % P, i0 + p + 1);
            load_col((p + K + 1) % P, j0 + p + K + 1);
        }
    }
};

// seq_bytes as for aff_stripe_kernel.  slow_list / slow_count receive the batches this kernel declines.
template <int K, int G, bool BT, bool D6 = false>
__global__ void __launch_bounds__(STRIPE_WARPS * 32, FAST_MIN_BLOCKS)
    aff_fast_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm, const uint8_t *__restrict__ pool, uint8_t *__restrict__ dir,
                    int *__restrict__ out_cost, int seq_bytes, int nslots, int *work_counter, const int *__restrict__ batch_list,
                    const int *__restrict__ batch_count, int *slow_list, int *slow_count, int keep_mask, int one) {
    // keep_mask = ~3 and one = 1 arrive as arguments so that they live in registers (see AffFast::cell, fma_add)
    // batch_list / batch_count: the batches aff_x2_kernel declined (null: all batches)
    if (batch_list != nullptr && *batch_count == 0) return;
    constexpr int GPW = 32 / G;
    constexpr int Q = 2 * K;
    using S_t = AffFast<K, G, BT, D6>;
    constexpr int BL = S_t::BL, P = S_t::P;
    extern __shared__ __align__(16) uint8_t smem[];
    int2 *s_tabR = reinterpret_cast<int2 *>(smem);
    int2 *s_tabC = s_tabR + 256;
    int *s_lut = reinterpret_cast<int *>(s_tabC + 256);
    // tables of the boundary phase (AffStripe, stripe_kernels.cuh)
    uint8_t *s_lut2 = smem + 2 * 256 * 8 + 16 * FAST_LUT_ROW;
    int *s_prep = reinterpret_cast<int *>(s_lut2 + STRIPE_LUT_BYTES);
    int *s_get = s_prep + 32;
    StageBars *s_bar = reinterpret_cast<StageBars *>(s_get + 32);  // one staging ring per group (staging.cuh)
    int *s_scr = reinterpret_cast<int *>(s_bar + STRIPE_WARPS * 4);  // FAST_SCR_INTS per group
    uint8_t *s_seq = reinterpret_cast<uint8_t *>(s_scr + STRIPE_WARPS * 4 * FAST_SCR_INTS);
    if (threadIdx.x < STRIPE_WARPS * GPW) StageRing<G>::init_bars(&s_bar[threadIdx.x]);
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
    StageRing<G> ring;
    ring.attach(&s_bar[warp_in_block * GPW + grp], s_seq + (size_t) ((warp_in_block * GPW + grp) * 2 * nslots) * op_stride, op_stride,
                nslots, lane);
    int *my_scr = s_scr + (warp_in_block * GPW + grp) * FAST_SCR_INTS;
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
        const int d0 = t.dhi + 2 - Q * G;
        ring.wait_full(slot);
        uint8_t *my_seq = ring.rows(slot);
      do {  // one pass; `break` hands the batch to aff_stripe_kernel
        // spare diagonals below dlo need the left-edge rule inside the stripe: not here
        const bool low = valid && (t.dlo - d0 > 0);
        if (__any_sync(0xffffffffu, low)) {
            if (lane32 == 0) slow_list[atomicAdd(slow_count, 1)] = batch;
            break;
        }
        // gap bits beyond the leading element of either operand (scanned in shared memory, 4 bytes per load)
        int gapbits = 0;
        if (valid) {
            for (int k = lane * 4; k < t.lr; k += G * 4) {
                uint32_t w = *reinterpret_cast<const volatile uint32_t *>(my_seq + k);
                if (k == 0) w &= 0xffffff00u;
                if (k + 4 > t.lr) w &= 0xffffffffu >> (8 * (k + 4 - t.lr));
                gapbits |= (int) (w & 0x10101010u);
            }
            for (int k = lane * 4; k < t.lc; k += G * 4) {
                uint32_t w = *reinterpret_cast<const volatile uint32_t *>(my_seq + op_stride + k);
                if (k == 0) w &= 0xffffff00u;
                if (k + 4 > t.lc) w &= 0xffffffffu >> (8 * (k + 4 - t.lc));
                gapbits |= (int) (w & 0x10101010u);
            }
        }
        if (__any_sync(0xffffffffu, gapbits != 0)) {
            if (lane32 == 0) slow_list[atomicAdd(slow_count, 1)] = batch;
            break;
        }

        const int u_first = (-d0) >> 1;  // first double step: t = 2u + d0 in {-1, 0}
        const int u_last = valid ? ((nr + nc - d0) >> 1) : (u_first - 1);
        const int sbase = (2 * u_first) & ~7;  // d0 + sbase == t.tshift
        uint8_t *dbase = dir + t.dir_off;
        const int dd_f = (nc - nr) - d0;
        const int lane_f = dd_f / Q;
        int u_end = u_last, u_begin = u_first;
        int u_b = max(G * K, 1 - d0);  // from here on every lane has i >= 1 and j >= 1
#pragma unroll
        for (int o = G; o < 32; o <<= 1) {
            u_end = max(u_end, __shfl_xor_sync(0xffffffffu, u_end, o));
            u_begin = min(u_begin, __shfl_xor_sync(0xffffffffu, u_begin, o));
            u_b = max(u_b, __shfl_xor_sync(0xffffffffu, u_b, o));
        }
        S_t S;
        int u = u_begin;
        int i0 = u - lane * K, j0 = u + d0 + lane * K;
        {
            // ---- boundary phase: row 0, column 0 and the cells before them, with the stripe kernel's own sweep
            AffStripe<K, G, BT, false, true> A;
            A.si = my_seq; A.sj = my_seq + op_stride;
            A.lut = s_lut2; A.prep = s_prep; A.get = s_get;
            A.nr = nr; A.nc = nc; A.go4 = 4 * cm.gap_open; A.lane = lane; A.qlow = 0;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                A.cb[q] = HIGH4 + TAG_CB; A.ev[q] = HIGH4 + TAG_EV; A.eh[q] = HIGH4 + TAG_EH; A.eb[q] = HIGH4 + TAG_EB;
            }
            A.init_windows(i0, j0);
            for (; u < u_b && u <= u_end; u++) {
                uint32_t de[2], dod[2];
                A.template double_step<true>(i0, j0, de, dod);
                if (BT) {
                    const int te = 2 * u + d0, s2 = 2 * u - sbase;  // s2 = te - tshift
                    if (u >= u_first && u <= u_last) {
                        if (D6) {  // the stripe sweep produced bytes: repack five codes into one word
                            uint32_t we = 0, wo = 0;
#pragma unroll
                            for (int m = 0; m < K; m++) {
                                we |= ((de[m >> 2] >> (8 * (m & 3))) & 63u) << (6 * m);
                                wo |= ((dod[m >> 2] >> (8 * (m & 3))) & 63u) << (6 * m);
                            }
                            de[0] = we;
                            dod[0] = wo;
                        }
                        if (te >= 0) store_dir<BL>(dbase + (((size_t) (s2 >> 3) * G + lane) * 8 + (s2 & 7)) * BL, de);
                        if (te + 1 <= nr + nc) store_dir<BL>(dbase + (((size_t) ((s2 + 1) >> 3) * G + lane) * 8 + ((s2 + 1) & 7)) * BL, dod);
                    }
                }
                if (u == u_last && lane == lane_f) {
#pragma unroll
                    for (int q = 0; q < Q; q++) my_scr[q] = min(min(A.cb[q], A.eh[q]), A.ev[q]);
                }
                i0++; j0++;
                A.slide_windows(i0, j0);
            }
#pragma unroll
            for (int q = 0; q < Q; q++) {
                S.cb[q] = A.cb[q] - TAG_CB + (FAST_CELL_V2 ? 4 * cm.gap_open : 0);
                S.ev[q] = A.ev[q]; S.eh[q] = A.eh[q];
            }
        }
        if (u <= u_end) {
            S.si = smem_u32(my_seq); S.sj = smem_u32(my_seq + op_stride);
            S.tabR = smem_u32(s_tabR); S.tabC = smem_u32(s_tabC);
            S.nr = nr; S.nc = nc; S.lane = lane; S.keep = keep_mask; S.one = one;
            S.c_h = 4 * cm.gap_open; S.c_v = 4 * cm.gap_open + TAG_EV;
            S.c_d = 3 - 4 * cm.gap_open; S.c_o = 4 * cm.gap_open - TAG_CB;
            S.u_last = u_last; S.lane_f = lane_f; S.scr = my_scr;
            S.init_windows(i0, j0);
            // local step of (double step u, even half) is 2u - sbase: fold the per-pair part into the pointer
            uint8_t *dptr = dbase + (ptrdiff_t) lane * 8 * BL - (ptrdiff_t) (sbase >> 3) * (G * 8 * BL);
            for (; u <= u_end; u += P, i0 += P, j0 += P) S.block(u, i0, j0, dptr);
        }

        if (valid && lane == lane_f) {
            int result = my_scr[dd_f % Q] >> 2;  // parked by this very lane
            if (BT && nr == 0 && nc == 0) result = 0;
            out_cost[t.pair] = result;
        }
      } while (0);
        __syncwarp();
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

#ifdef POYB200_DEFINE_AFF_FAST  // the translation unit that owns these kernels (k_aff_fast.cu)
template <int K, int G, bool D6 = false>
static cudaError_t fast_launch_shape(bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir, int *cost,
                                     int sm_count, int seq_bytes, int *work_counter, const int *batch_list, const int *batch_count,
                                     int *slow_list, int *slow_count, cudaStream_t stream) {
    constexpr int GPW = 32 / G;
    const int nbatches = (n + GPW - 1) / GPW;
    auto kern = bt ? aff_fast_kernel<K, G, true, D6> : aff_fast_kernel<K, G, false, false>;
    size_t smem = 0;
    int nslots = 1, per_sm = 1;
    cudaError_t e = stage_ring_config(kern, FAST_TABLE_BYTES, (size_t) STRIPE_WARPS * GPW * 2 * (seq_bytes + fast_operand_pad(K, G)),
                                      STRIPE_WARPS * 32, smem, nslots, per_sm);
    if (e != cudaSuccess) return e;
    int blocks = std::min((nbatches + STRIPE_WARPS - 1) / STRIPE_WARPS, sm_count * per_sm);
    if (blocks < 1) blocks = 1;
    kern<<<blocks, STRIPE_WARPS * 32, smem, stream>>>(d_tasks, n, cm, pool, dir, cost, seq_bytes, nslots, work_counter, batch_list, batch_count,
                                                      slow_list, slow_count, ~3, 1);
    return cudaGetLastError();
}

cudaError_t fast_launch(uint32_t klass, bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir,
                                      int *cost, int sm_count, int seq_bytes, int *work_counter, const int *batch_list,
                                      const int *batch_count, int *slow_list, int *slow_count, bool dir6, cudaStream_t stream) {
    if (dir6 && klass == 1)  // tasks flagged TF_DIR6 by the planner: the (5, 8) shape with the traceback kernel as reader
        return fast_launch_shape<5, 8, true>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, batch_list, batch_count, slow_list, slow_count, stream);
    switch (klass - 1) {
        case 0: return fast_launch_shape<5, 8>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, batch_list, batch_count, slow_list, slow_count, stream);
        case 1: return fast_launch_shape<6, 8>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, batch_list, batch_count, slow_list, slow_count, stream);
        case 2: return fast_launch_shape<4, 16>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, batch_list, batch_count, slow_list, slow_count, stream);
        case 3: return fast_launch_shape<6, 16>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, batch_list, batch_count, slow_list, slow_count, stream);
        case 4: return fast_launch_shape<4, 32>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, batch_list, batch_count, slow_list, slow_count, stream);
        case 5: return fast_launch_shape<6, 32>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, batch_list, batch_count, slow_list, slow_count, stream);
        default: return cudaErrorInvalidValue;
    }
}

#endif  // POYB200_DEFINE_AFF_FAST

}  // namespace poyb200
