// Ring kernels: the affine_3 fill (src/algn.c:2411-2548) AND the traceback (backtrace_affine, :1983-2097) of the pairs
// whose stripe is flush with dlo (no spare diagonals), in one persistent kernel.
//
// Fill.  Same stripe formulation, same direction codes and same costs as aff_stripe_kernel (stripe_kernels.cuh), which
// stays the specification, but built for instruction count:
//   * the row / column windows do not slide.  They are rings of K + 1 slots and the sweep is unrolled in blocks of K + 1
//     double steps, so every slot index is a compile-time constant and no register is ever copied;
//   * a row / column entering a window is one or two shared loads from small class tables: what a row contributes to a
//     cell depends only on (its code, whether the previous row carried the gap bit), likewise for columns;
//   * by the choice of Task::tshift the two steps of a double step are neighbours inside a tile of the direction band
//     and the tile phase depends on the double-step counter only: one store per double step at a warp-uniform offset;
//   * the middle of the sweep runs blocks without any position test; the first double steps (row 0 / column 0 / rows
//     and columns 1, whose bookkeeping has special cases, :2454-2458, 2483-2485) run aff_stripe_kernel's own sweep.
//   EBF = false: neither operand carries a gap bit beyond its leading element and gap_open > 0 (leaf sequences).  The
//     block-diagonal state is dropped (NOEB argument in stripe_kernels.cuh), a row is {4 cost[si][gap], LUT row}, a column
//     {4 prepend[sj], LUT column}; with K = 5 a cell's direction code is 6 bits and a lane's step is ONE 32-bit word.
//   EBF = true: operands with gap bits (internal-node medians): the full cell, block-diagonal state included.
//   Batches an instance cannot take (gap bits under EBF = false) are appended to a list for the next launch.
//
// Walk.  A warp fills PEND pairs (GPW at a time) into its own slots of a scratch band, then its lanes walk one pair
// each (walk.cuh), then the slots are reused.  The walk is a chain of dependent loads; running it inside the fill kernel
// overlaps that latency with the ALU-bound fill of the SM's other warps without a second kernel taking registers from
// the fill, the band of a pair is read back by the warp that wrote it, and the band never has to exist for a whole
// chunk (PEND x warps slots instead of one band per pair of the chunk).
#pragma once
#include "stripe_kernels.cuh"
#include "walk.cuh"

namespace poyb200 {

constexpr int RING_LUT_ROW = 20 * 4;  // 16 ints + 4 pad
// LUT slot of a 4-bit code: A, C, G, T (1, 2, 4, 8) take slots 0..3, so with rows 20 words apart the 16 combinations
// of unambiguous bases sit in 16 different banks; the other codes follow.
__host__ __device__ constexpr int ring_lut_slot(int code) {
    // slots {4, 0, 1, 5, 2, 6, 7, 8, 3, 9, 10, 11, 12, 13, 14, 15} for codes 0..15, one nibble each
    return (int) ((0xfedcba9387625104ull >> (4 * (code & 15))) & 15);
}
constexpr int RING_SCR_INTS = 16;  // >= 2 K
constexpr int RING_PEND = 32;      // pairs a warp fills before its lanes walk them
// tabR, tabC (int4, 64 classes each: code | prev-gap << 5), LUT of the ring blocks; LUT, prepend and gap tables of the
// boundary phase; staging barriers; per-group scratch; per-warp pending lists
constexpr int RING_TABLE_BYTES = 2 * 64 * 16 + 16 * RING_LUT_ROW + STRIPE_LUT_BYTES + 64 * 4 + STRIPE_WARPS * 4 * STAGE_BAR_BYTES +
                                 STRIPE_WARPS * 4 * RING_SCR_INTS * 4 + STRIPE_WARPS * RING_PEND * 4 + 256;
// The unchecked blocks read codes past an operand's end (rows up to Q G / 2 + D - 2 past it, columns up to G K + D - 1):
// every staged operand gets that much private slack, so the stray reads never touch another warp's buffers.
__host__ __device__ constexpr int ring_operand_pad(int K, int G) { return (K * G + 8 + 15) & ~15; }
#ifndef RING_MIN_BLOCKS
#define RING_MIN_BLOCKS 3
#endif
#ifndef RING_MIN_BLOCKS_EB
#define RING_MIN_BLOCKS_EB 2
#endif
// Double steps per unrolled block of the full (block-diagonal) instance with K = 5.  A whole ring turn (6) needs no register
// moves but is 42 KB of code, and the kernel then waits for instructions more than for anything else (ncu: "no instruction"
// 0.59 stalls per issue, profiles/README.md); half a turn (3) halves the code and rotates the windows by three slots -- 30
// register moves per 30 cells -- at the end of the block.
#ifndef RING_EB_UNROLL
#define RING_EB_UNROLL 3
#endif

__device__ __forceinline__ int lds_s32(uint32_t a) {
    int v;
    asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int4 lds_v4(uint32_t a) {
    int4 v;
    asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
// a * b + c on the FMA pipe (the INT32 ALU pipe is the busier one in the unchecked block)
__device__ __forceinline__ int imad(int a, int b, int c) {
    int v;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(v) : "r"(a), "r"(b), "r"(c));
    return v;
}
// staged operands change from pair to pair: volatile, so the load is neither hoisted nor merged across pairs
__device__ __forceinline__ int lds_u8_seq(uint32_t a) {
    int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// Direction-band format of a ring instance: bytes (K <= 4: 4 per word; EBF or K = 6: 8-byte chunks) or, for K = 5 without
// the block-diagonal state, five 6-bit codes in one word (END_BLOCK is constant there).
template <int K, bool EBF>
struct RingFmt {
    static constexpr bool DIR6 = (!EBF && K == 5);
    static constexpr int BL = (K <= 4 || DIR6) ? 4 : 8;
};

// The band of one pair as a ring kernel wrote it, read back by the walk of the same warp.
//
// A walk is a chain of dependent loads: the next cell is known only when the current direction code has been decoded.
// What IS known in advance is the order of the tiles: a step moves one or two anti-diagonals back, so the walk passes
// through every tile of 8 anti-diagonals, in descending order, and inside a tile it cannot drift further than into the
// neighbouring lane chunk (a chunk spans 2 K >= 8 diagonals).  Each walker therefore keeps a private two-buffer window in
// shared memory: the three adjacent lane chunks (3 * 8 * BL contiguous bytes) of the current tile and of the next one,
// fetched with cp.async (16-byte copies, L2 path: the band was written by this SM's own stores) one tile ahead.  The
// per-step access is then a shared-memory load; a step that leaves the window (rare) reads global memory directly.
template <int K, int G, bool EBF>
struct BandRing {
    static constexpr int BL = RingFmt<K, EBF>::BL;
    static constexpr bool DIR6 = RingFmt<K, EBF>::DIR6;
    static constexpr int CH = 8 * BL, WIN = 3 * CH;  // bytes of one lane chunk of a tile / of the window
    static constexpr int NW = (G >= 3) ? 3 : G;      // lanes in the window
    const uint8_t *dbase;
    int dbase_d, tshift;
    uint32_t sbuf;   // shared address of this thread's 2 * WIN bytes
    int cur_tile, ws0, ws1;

    __device__ __forceinline__ void load_tile(int tile, int lane) {
        const int w0 = min(max(lane - 1, 0), G - NW);
        if (tile & 1) ws1 = w0; else ws0 = w0;
        const uint8_t *src = dbase + ((size_t) tile * G + w0) * CH;
        const uint32_t dst = sbuf + (tile & 1) * WIN;
#pragma unroll
        for (int o = 0; o < NW * CH; o += 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // Before the first fetch: the tile of the walk's first cell (nr, nc) and the one below it.
    __device__ __forceinline__ void start(int i, int j) {
        const uint32_t dd = (uint32_t) ((j - i) - dbase_d), T = (uint32_t) (i + j - tshift);
        const int lane = (int) (dd / (2 * K)), tile = (int) (T >> 3);
        ws0 = ws1 = 0;
        load_tile(tile, lane);
        if (tile > 0) load_tile(tile - 1, lane);
        cur_tile = tile;
        if (tile > 0) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __device__ __forceinline__ int fetch(int i, int j) {
        const uint32_t dd = (uint32_t) ((j - i) - dbase_d), T = (uint32_t) (i + j - tshift);
        const int lane = (int) (dd / (2 * K)), tile = (int) (T >> 3);
        const uint32_t m = (dd - (uint32_t) lane * (2 * K)) >> 1;
        if (tile != cur_tile) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if (tile == cur_tile - 1) {  // one tile down (a step moves at most two anti-diagonals): it was prefetched
                cur_tile = tile;
                if (tile > 0) load_tile(tile - 1, lane);  // into the buffer the walk has just left
            } else {  // the walk spent some steps outside the band (edge cells need no fetch) and skipped a tile: start over
                start(i, j);
            }
        }
        const int rel = lane - ((tile & 1) ? ws1 : ws0);
        const uint32_t in_chunk = (T & 7) * BL;
        if ((unsigned) rel < (unsigned) NW) {
            const uint32_t a = sbuf + (tile & 1) * WIN + rel * CH + in_chunk;
            if (DIR6) {
                uint32_t w;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(a));
                return (int) ((w >> (6 * m)) & 63u) | AB_ENDB;
            }
            uint32_t v;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a + m));
            return (int) v;
        }
        const uint8_t *g = dbase + ((size_t) tile * G + lane) * CH + in_chunk;
        if (DIR6) return (int) ((*reinterpret_cast<const volatile uint32_t *>(g) >> (6 * m)) & 63u) | AB_ENDB;
        return *reinterpret_cast<const volatile uint8_t *>(g + m);
    }
    __device__ __forceinline__ void finish() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
};
template <int K, bool EBF>
__host__ __device__ constexpr int ring_walk_window_bytes() { return 2 * 3 * 8 * RingFmt<K, EBF>::BL; }  // per thread

template <int K, int G, bool BT, bool EBF>
struct AffRing {
    // P = ring slots = double steps per unrolled block.  K + 1 columns are live in a step and the next row / column
    // is fetched one step ahead into the slot that just died, so K + 1 slots do.  Instruction-cache footprint decides the
    // speed of this kernel: a block is 6 double steps x 10 cells x ~26 instructions x 16 B = 25 KB for K = 5, and
    // nothing else may be large -- a first version whose boundary phases were unrolled the same way (88 KB per block)
    // ran at half the speed with 57 % "no instruction" stalls (profiles/).
    static constexpr int Q = 2 * K, P = K + 1, BL = RingFmt<K, EBF>::BL;
    static constexpr bool DIR6 = RingFmt<K, EBF>::DIR6;
    static constexpr int NEB = EBF ? Q : 1, NW = EBF ? P : 1;
    static constexpr int UN = (EBF && K == 5 && (P % RING_EB_UNROLL) == 0) ? RING_EB_UNROLL : P;  // double steps per block

    int cb[Q], ev[Q], eh[Q], eb[NEB];  // cb carries tag 0 here (4 * CB); ev, eh, eb their state tags
    // row window:  Rv = 4 * vertical extension, Rl = shared address of the LUT row;
    //   EBF: Rg = 4 * (gap opening + extension) + TAG_EV, Rgop = 4 * gap opening, Rgf = 1 when the row carries the gap bit
    // column window: Cv = 4 * horizontal extension, Cl = byte offset of the column in a LUT row;
    //   EBF: Cg = 4 * (gap opening + prepend), Cgop, Cgf
    int Rv[P], Cv[P];
    uint32_t Rl[P], Cl[P];
    int Rg[NW], Rgop[NW], Rgf[NW], Cg[NW], Cgop[NW], Cgf[NW];
    uint32_t si, sj, tabR, tabC;  // shared addresses
    int nr, nc, lane, keep;
    int one;       // 1, opaque (forces multiply-adds)
    int c_h, c_v;  // EBF = false: go4 (CB tag 0 -> EH tag 0), go4 + 2 (-> EV tag 2): registers, so the adds stay two-input
    int last_ci, last_cj;  // EBF: codes of the newest row / column (the class of the next one depends on their gap bit)
    int high4, go8;        // EBF: 4 * HIGH_NUM; 8 * gap_open (cost-only build, :1846)
    // per pair
    int u_last, lane_f;
    int *scr;  // shared, Q ints per group

    // Rows / columns past the end of an operand (a pair that finished while others of the warp still run, or the last
    // lanes of a stripe that overhangs the matrix) read the last code again: such cells are never used.
    __device__ __forceinline__ void load_row(int slot, int i) {
        const int ci = lds_u8_seq(si + min(i, nr));
        if (EBF) {
            const int4 e = lds_v4(tabR + 16 * (ci + 2 * (last_ci & 16)));
            last_ci = ci;
            Rv[slot] = e.x; Rg[slot] = e.y; Rgop[slot] = e.z & ~3; Rgf[slot] = e.z & 1; Rl[slot] = (uint32_t) e.w;
        } else {
            const int4 e = lds_v4(tabR + 16 * ci);
            Rv[slot] = e.x; Rl[slot] = (uint32_t) e.w;
        }
    }
    __device__ __forceinline__ void load_col(int slot, int j) {
        const int cj = lds_u8_seq(sj + min(j, nc));
        if (EBF) {
            const int4 e = lds_v4(tabC + 16 * (cj + 2 * (last_cj & 16)));
            last_cj = cj;
            Cv[slot] = e.x; Cg[slot] = e.y; Cgop[slot] = e.z & ~3; Cgf[slot] = e.z & 1; Cl[slot] = (uint32_t) e.w;
        } else {
            const int4 e = lds_v4(tabC + 16 * cj);
            Cv[slot] = e.x; Cl[slot] = (uint32_t) e.w;
        }
    }
    // Windows as block entry expects them: rows i0-K+1 .. i0 in slots r mod P, columns j0 .. j0+K in slots n mod P
    // (all rows and columns >= 2 here: the class tables do not know the special cases of row / column 1).
    __device__ __forceinline__ void init_windows(int i0, int j0) {
        if (EBF) {
            last_ci = lds_u8_seq(si + min(i0 - K, nr));
            last_cj = lds_u8_seq(sj + min(j0 - 1, nc));
        }
#pragma unroll
        for (int r = -K + 1; r <= 0; r++) load_row((r + P) % P, i0 + r);
#pragma unroll
        for (int n = 0; n <= K; n++) load_col(n % P, j0 + n);
    }

    // One interior cell.  ehl, cbl: left neighbour; evu, cbu: upper neighbour; the diagonal neighbour is this slot's own
    // state.  Returns the direction code (common.cuh; without END_BLOCK when DIR6).
    __device__ __forceinline__ int cell(int ehl, int cbl, int evu, int cbu, int q, int rs, int cs) {
        int neh, nev, fl_h, fl_v;
        if (EBF) {
            const int xo = cbl + Cg[cs], xe = ehl + Cv[cs];   // FILL_EXTEND_HORIZONTAL :1765-1787
            neh = min(xe, xo);
            fl_h = (xe < xo) ? 0 : AB_ENDH;
            const int yo = cbu + Rg[rs], ye = evu + Rv[rs];   // FILL_EXTEND_VERTICAL :1813-1830
            nev = min(ye, yo);
            fl_v = (ye < yo) ? 0 : AB_ENDV;
        } else {
            const int t = cbl + c_h, t2 = cbu + c_v;
            neh = min(ehl, t) + Cv[cs];
            nev = min(evu, t2) + Rv[rs];
            fl_h = (ehl < t) ? 0 : AB_ENDH;
            fl_v = (evu < t2) ? 0 : AB_ENDV;
        }
        const int d = lds_s32(Rl[rs] + Cl[cs]);   // 4 * cost[si & 15][sj & 15]
        int ck, neb = 0, fl_b = DIR6 ? 0 : AB_ENDB;
        if (EBF) {
            // FILL_CLOSE_BLOCK_DIAGONAL :1923-1977, candidates tagged H 0, D 1, V 2, A 3
            const int a0 = cb[q] + 3;
            const int a1 = imad(Cgop[cs], Rgf[rs], ev[q]);
            const int a2 = imad(Rgop[rs], Cgf[cs], eh[q]);
            const int a3 = eb[q] + max(Rgop[rs], Cgop[cs]) - 2;
            ck = min(min(a0, a1), min(a2, a3)) + d;
            // FILL_EXTEND_BLOCK_DIAGONAL :1861-1882 / _NOBT :1837-1854: a0 = close-block + TAG_EB as well
            const int both = Rgf[rs] * Cgf[cs];
            const int dg = imad(both, -high4, high4);  // 0 when both carry the gap bit, else 4 * HIGH_NUM
            if (BT) {
                neb = min(eb[q], a0) + dg;  // extend and open share the addend (:1871-1872)
                fl_b = (eb[q] < a0) ? 0 : AB_ENDB;
            } else {
                neb = min(eb[q], imad(both, go8, a0)) + dg;  // opening costs 2 * go when both carry a gap (:1846)
            }
        } else {
            ck = min(min(cb[q] + 3, ev[q]), eh[q]) + d;  // tags A 3, V 2, H 0
        }
        const int ncb = ck & keep;
        int byte = 0;
        if (BT) {
            int fk = min(min(neh, nev), imad(ncb, one, one));  // ASSIGN_MINIMUM :2251-2280 (one = TAG_CB)
            if (EBF) fk = min(fk, neb);
            const int flags = imad(fl_h + fl_b, one, fl_v);
            byte = ((fk * 4 + (ck - ncb)) & 15) | flags;
        }
        cb[q] = ncb; ev[q] = nev; eh[q] = neh;
        if (EBF) eb[q] = neb;
        return byte;
    }

    // K direction codes -> one or two words, by multiply-adds (FMA pipe)
    __device__ __forceinline__ void pack(const int (&by)[K], uint32_t (&w)[2]) {
        if (DIR6) {
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

    template <class T, int N>
    static __device__ __forceinline__ void rotate(T (&a)[N]) {  // a[s] <- a[(s + UN) % N]
        if (N != P || UN == P) return;
        T t[N];
#pragma unroll
        for (int s_ = 0; s_ < N; s_++) t[s_] = a[(s_ + UN) % N];
#pragma unroll
        for (int s_ = 0; s_ < N; s_++) a[s_] = t[s_];
    }

    // One block of UN double steps starting at double step u (every lane at rows and columns >= 2), lane origin
    // (i0, j0).  dptr = the address this lane's chunk of local step 0 has once the pair's tile origin is folded in.
    __device__ __forceinline__ void block(int u, int i0, int j0, uint8_t *dptr) {
#pragma unroll
        for (int p = 0; p < UN; p++) {
            uint32_t de[2] = {0, 0}, dod[2] = {0, 0};
            int by[K];
            // ---- even step: q = 2m, cell (i0 + p - m, j0 + p + m)
            int in_eh = __shfl_up_sync(0xffffffffu, eh[Q - 1], 1, G);
            int in_cb = __shfl_up_sync(0xffffffffu, cb[Q - 1], 1, G);
            if (lane == 0) { in_eh = HIGH4 + TAG_EH; in_cb = HIGH4; }  // the left-edge cells (:2487, :2494)
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
                        cb[q] = HIGH4; ev[q] = HIGH4 + TAG_EV; eh[q] = HIGH4 + TAG_EH;
                        if (EBF) eb[q] = HIGH4 + TAG_EB;
                    }
                }
            }
            if (BT) pack(by, dod);
            // ---- direction codes.  Steps 2(u+p) and 2(u+p)+1 are neighbours inside a tile (Task::tshift): one store for
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
                    for (int q = 0; q < Q; q++) {
                        int v = min(min(cb[q], eh[q]), ev[q]);
                        if (EBF) v = min(v, eb[q]);
                        scr[q] = v;
                    }
                }
            }
            // ---- windows: row i0 + p + 1 and column j0 + p + K + 1 enter
            load_row((p + 1) % P, i0 + p + 1);
            load_col((p + K + 1) % P, j0 + p + K + 1);
        }
        if (UN != P) {  // the next block expects its rows / columns in the slots this one started with
            rotate(Rv); rotate(Cv); rotate(Rl); rotate(Cl);
            rotate(Rg); rotate(Rgop); rotate(Rgf); rotate(Cg); rotate(Cgop); rotate(Cgf);
        }
    }
};

// seq_bytes / nslots as for aff_stripe_kernel.  batch_list / batch_count: the batches a previous instance declined (null:
// all batches).  slow_list / slow_count receive the batches THIS instance declines (EBF = false: operands with gap bits).
// scratch: pend_max slots of slot_bytes per warp of the grid (BT only); out: where the walk writes.
template <int K, int G, bool BT, bool EBF>
__global__ void __launch_bounds__(STRIPE_WARPS * 32, EBF ? RING_MIN_BLOCKS_EB : RING_MIN_BLOCKS)
    aff_ring_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm, const uint8_t *__restrict__ pool, uint8_t *scratch,
                    unsigned long long slot_bytes, int pend_max, OutPtrs out, int seq_bytes, int nslots, int *work_counter,
                    const int *__restrict__ batch_list, const int *__restrict__ batch_count, int *slow_list, int *slow_count,
                    int keep_mask, int one) {
    // keep_mask = ~3 and one = 1 arrive as arguments so that they live in registers (see AffRing::cell)
    if (batch_list != nullptr && *batch_count == 0) return;
    constexpr int GPW = 32 / G;
    constexpr int Q = 2 * K;
    using S_t = AffRing<K, G, BT, EBF>;
    constexpr int BL = S_t::BL, P = S_t::P;
    constexpr int BLB = (K <= 4) ? 4 : 8;  // chunk bytes of the boundary phase's own (byte) codes
    extern __shared__ __align__(16) uint8_t smem[];
    int4 *s_tabR = reinterpret_cast<int4 *>(smem);
    int4 *s_tabC = s_tabR + 64;
    int *s_lut = reinterpret_cast<int *>(s_tabC + 64);
    // tables of the boundary phase (AffStripe, stripe_kernels.cuh)
    uint8_t *s_lut2 = smem + 2 * 64 * 16 + 16 * RING_LUT_ROW;
    int *s_prep = reinterpret_cast<int *>(s_lut2 + STRIPE_LUT_BYTES);
    int *s_get = s_prep + 32;
    StageBars *s_bar = reinterpret_cast<StageBars *>(s_get + 32);  // one staging ring per group (staging.cuh)
    int *s_scr = reinterpret_cast<int *>(s_bar + STRIPE_WARPS * 4);  // RING_SCR_INTS per group
    int *s_pend = s_scr + STRIPE_WARPS * 4 * RING_SCR_INTS;          // RING_PEND task indices per warp
    uint8_t *s_med = reinterpret_cast<uint8_t *>(s_pend + STRIPE_WARPS * RING_PEND);  // 16 x 16 medians of 4-bit codes (the walk)
    uint8_t *s_win = s_med + 256;                                                     // per-thread band windows of the walk
    uint8_t *s_seq = s_win + (BT ? STRIPE_WARPS * 32 * ring_walk_window_bytes<K, EBF>() : 0);
    if (threadIdx.x < STRIPE_WARPS * GPW) StageRing<G>::init_bars(&s_bar[threadIdx.x]);
    const int go4 = 4 * cm.gap_open;
    for (int k = threadIdx.x; k < 256; k += blockDim.x) {
        const int c4 = 4 * __ldg(cm.cost + ((k >> 4) << cm.lcm) + (k & 15));
        s_lut[ring_lut_slot(k >> 4) * (RING_LUT_ROW / 4) + ring_lut_slot(k)] = c4;
        s_med[k] = __ldg(cm.median + ((k >> 4) << cm.lcm) + (k & 15));
        *reinterpret_cast<int2 *>(s_lut2 + (k >> 4) * LUT_ROW_BYTES + (k & 15) * 8) = make_int2(c4, c4 - 2);
    }
    for (int k = threadIdx.x; k < 64; k += blockDim.x) {
        // class k: code = k & 31, previous element carried the gap bit = k >> 5 (rows / columns >= 2: no special cases)
        const int code = k & 31, pg = k >> 5, cg = (code >> 4) & 1;
        const int gop = (!pg && cg) ? 0 : go4;                                   // HAS_GAP_OPENING :1730
        const int ge4 = 4 * __ldg(cm.cost + (code << cm.lcm) + cm.gap), g4 = 4 * __ldg(cm.prepend + code);
        const int lutrow = (int) (smem_u32(s_lut) + ring_lut_slot(code) * RING_LUT_ROW), lutcol = ring_lut_slot(code) * 4;
        if (EBF) {
            const int vx = (pg && !cg) ? gop + ge4 : ge4;                        // :2483-2485
            const int hx = (pg && !cg) ? gop + g4 : g4;                          // :2454-2457
            s_tabR[k] = make_int4(vx, gop + ge4 + TAG_EV, gop + cg, lutrow);
            s_tabC[k] = make_int4(hx, gop + g4, gop + cg, lutcol);
        } else {
            s_tabR[k] = make_int4(ge4, 0, 0, lutrow);
            s_tabC[k] = make_int4(g4, 0, 0, lutcol);
        }
    }
    for (int k = threadIdx.x; k < 32; k += blockDim.x) {
        s_prep[k] = 4 * __ldg(cm.prepend + k);
        s_get[k] = 4 * __ldg(cm.cost + (k << cm.lcm) + cm.gap);
    }
    __syncthreads();

    const int warp_in_block = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
    const int grp = lane32 / G, lane = lane32 % G;
    const int op_stride = seq_bytes + ring_operand_pad(K, G);
    StageRing<G> ring;
    ring.attach(&s_bar[warp_in_block * GPW + grp], s_seq + (size_t) ((warp_in_block * GPW + grp) * 2 * nslots) * op_stride, op_stride,
                nslots, lane);
    int *my_scr = s_scr + (warp_in_block * GPW + grp) * RING_SCR_INTS;
    int *my_pend = s_pend + warp_in_block * RING_PEND;
    uint8_t *my_scratch = scratch + (size_t) (blockIdx.x * STRIPE_WARPS + warp_in_block) * (size_t) pend_max * slot_bytes;
    const int nbatches = (ntasks + GPW - 1) / GPW;
    int npend = 0;  // pairs filled and not yet walked (warp-uniform)

    // The lanes of the warp walk the pending pairs, one each (backtrace_affine + outputs, walk.cuh).
    auto walk_pending = [&]() {
        if (!BT) return;
        __syncwarp();  // the band stores and the pending list of this warp are visible to all its lanes
        const int ti = (lane32 < npend) ? my_pend[lane32] : -1;
        if (ti >= 0) {
            const Task t = tasks[ti];
            BandRing<K, G, EBF> band;
            band.dbase = my_scratch + (size_t) lane32 * slot_bytes;
            band.dbase_d = t.dbase;
            band.tshift = t.tshift;
            band.sbuf = smem_u32(s_win + (size_t) threadIdx.x * ring_walk_window_bytes<K, EBF>());
            aff_walk_pair(t, pool, band, MedianShared{smem_u32(s_med)}, cm, out);
            if (out.walked) out.walked[t.pair] = 1;
        }
        __syncwarp();
        npend = 0;
    };

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
      do {  // one pass; `break` hands the batch to the next launch
        if (!EBF) {
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
        }

        const int u_first = (-d0) >> 1;  // first double step: t = 2u + d0 in {-1, 0}
        const int u_last = valid ? ((nr + nc - d0) >> 1) : (u_first - 1);
        const int sbase = (2 * u_first) & ~7;  // d0 + sbase == t.tshift
        uint8_t *dbase = my_scratch + (size_t) (npend + grp) * slot_bytes;
        const int dd_f = (nc - nr) - d0;
        const int lane_f = dd_f / Q;
        int u_end = u_last, u_begin = u_first;
        int u_b = max(G * K, 1 - d0) + 1;  // from here on every lane has i >= 2 and j >= 2
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
            // ---- boundary phase: row 0, column 0, rows / columns 1 and the cells before them, with the stripe kernel's sweep
            AffStripe<K, G, BT, false, !EBF> A;
            A.si = my_seq; A.sj = my_seq + op_stride;
            A.lut = s_lut2; A.prep = s_prep; A.get = s_get;
            A.nr = nr; A.nc = nc; A.go4 = go4; A.lane = lane; A.qlow = 0;
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
                        if (S_t::DIR6) {  // the stripe sweep produced bytes: repack five codes into one word
                            uint32_t we = 0, wo = 0;
#pragma unroll
                            for (int m = 0; m < K; m++) {
                                we |= ((de[m >> 2] >> (8 * (m & 3))) & 63u) << (6 * m);
                                wo |= ((dod[m >> 2] >> (8 * (m & 3))) & 63u) << (6 * m);
                            }
                            if (te >= 0) *reinterpret_cast<uint32_t *>(dbase + (((size_t) (s2 >> 3) * G + lane) * 8 + (s2 & 7)) * 4) = we;
                            if (te + 1 <= nr + nc)
                                *reinterpret_cast<uint32_t *>(dbase + (((size_t) ((s2 + 1) >> 3) * G + lane) * 8 + ((s2 + 1) & 7)) * 4) = wo;
                        } else {
                            if (te >= 0) store_dir<BLB>(dbase + (((size_t) (s2 >> 3) * G + lane) * 8 + (s2 & 7)) * BLB, de);
                            if (te + 1 <= nr + nc) store_dir<BLB>(dbase + (((size_t) ((s2 + 1) >> 3) * G + lane) * 8 + ((s2 + 1) & 7)) * BLB, dod);
                        }
                    }
                }
                if (u == u_last && lane == lane_f) {
#pragma unroll
                    for (int q = 0; q < Q; q++) {
                        int v = min(min(A.cb[q], A.eh[q]), A.ev[q]);
                        if (EBF) v = min(v, A.eb[q]);
                        my_scr[q] = v;
                    }
                }
                i0++; j0++;
                A.slide_windows(i0, j0);
            }
#pragma unroll
            for (int q = 0; q < Q; q++) {
                S.cb[q] = A.cb[q] - TAG_CB; S.ev[q] = A.ev[q]; S.eh[q] = A.eh[q];
                if (EBF) S.eb[q] = A.eb[q];
            }
        }
        if (u <= u_end) {
            S.si = smem_u32(my_seq); S.sj = smem_u32(my_seq + op_stride);
            S.tabR = smem_u32(s_tabR); S.tabC = smem_u32(s_tabC);
            S.nr = nr; S.nc = nc; S.lane = lane; S.keep = keep_mask; S.one = one;
            S.c_h = go4; S.c_v = go4 + TAG_EV;
            S.high4 = HIGH4 * one; S.go8 = 2 * go4;
            S.u_last = u_last; S.lane_f = lane_f; S.scr = my_scr;
            S.init_windows(i0, j0);
            // local step of (double step u, even half) is 2u - sbase: fold the per-pair part into the pointer
            uint8_t *dptr = dbase + (ptrdiff_t) lane * 8 * BL - (ptrdiff_t) (sbase >> 3) * (G * 8 * BL);
            for (; u <= u_end; u += S_t::UN, i0 += S_t::UN, j0 += S_t::UN) S.block(u, i0, j0, dptr);
        }

        if (valid && lane == lane_f) {
            int result = my_scr[dd_f % Q] >> 2;  // parked by this very lane
            if (BT && nr == 0 && nc == 0) result = 0;
            out.cost[t.pair] = result;
        }
        if (BT) {
            if (lane == 0) my_pend[npend + grp] = valid ? ti : -1;
            npend += GPW;
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
        if (BT && (npend + GPW > pend_max || batch < 0) && npend > 0) walk_pending();
    }
}

#ifdef POYB200_DEFINE_AFF_RING  // the translation unit that owns these kernels (k_aff_ring.cu)
struct RingLaunch {
    const Task *d_tasks; int n; DevCM cm; const uint8_t *pool; uint8_t *scratch; size_t scratch_bytes; size_t slot_bytes; OutPtrs out;
    int sm_count; int seq_bytes; int *work_counter; const int *batch_list; const int *batch_count; int *slow_list; int *slow_count;
    cudaStream_t stream;
};

template <int K, int G, bool EBF>
static cudaError_t ring_launch_shape(bool bt, const RingLaunch &a) {
    constexpr int GPW = 32 / G;
    const int nbatches = (a.n + GPW - 1) / GPW;
    auto kern = bt ? aff_ring_kernel<K, G, true, EBF> : aff_ring_kernel<K, G, false, EBF>;
    size_t smem = 0;
    int nslots = 1, per_sm = 1;
    const size_t fixed = RING_TABLE_BYTES + (bt ? (size_t) STRIPE_WARPS * 32 * ring_walk_window_bytes<K, EBF>() : 0);
    cudaError_t e = stage_ring_config(kern, fixed, (size_t) STRIPE_WARPS * GPW * 2 * (a.seq_bytes + ring_operand_pad(K, G)),
                                      STRIPE_WARPS * 32, smem, nslots, per_sm);
    if (e != cudaSuccess) return e;
    int blocks = std::min((nbatches + STRIPE_WARPS - 1) / STRIPE_WARPS, a.sm_count * per_sm);
    if (blocks < 1) blocks = 1;
    // pending pairs per warp: as many as the scratch holds for this grid, a multiple of GPW, at most RING_PEND
    int pend = RING_PEND;
    if (bt) {
        const size_t per_warp = a.scratch_bytes / ((size_t) blocks * STRIPE_WARPS);
        pend = (int) std::min<size_t>(RING_PEND, per_warp / std::max<size_t>(a.slot_bytes, 1));
        pend -= pend % GPW;
        if (pend < GPW) return cudaErrorMemoryAllocation;
    }
    kern<<<blocks, STRIPE_WARPS * 32, smem, a.stream>>>(a.d_tasks, a.n, a.cm, a.pool, a.scratch, (unsigned long long) a.slot_bytes, pend,
                                                        a.out, a.seq_bytes, nslots, a.work_counter, a.batch_list, a.batch_count,
                                                        a.slow_list, a.slow_count, ~3, 1);
    return cudaGetLastError();
}

// Bytes of scratch one resident grid of ring kernels wants for slots of slot_bytes (RING_PEND slots per warp, at most
// RING_MAX_CTAS_PER_SM CTAs per SM).
constexpr int RING_MAX_CTAS_PER_SM = 4;
size_t ring_scratch_bytes(int sm_count, size_t slot_bytes) {
    return (size_t) sm_count * RING_MAX_CTAS_PER_SM * STRIPE_WARPS * RING_PEND * slot_bytes;
}

cudaError_t ring_launch(uint32_t klass, bool bt, bool ebf, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *scratch,
                        size_t scratch_bytes, size_t slot_bytes, OutPtrs out, int sm_count, int seq_bytes, int *work_counter,
                        const int *batch_list, const int *batch_count, int *slow_list, int *slow_count, cudaStream_t stream) {
    const RingLaunch a{d_tasks, n, cm, pool, scratch, scratch_bytes, slot_bytes, out, sm_count, seq_bytes, work_counter,
                       batch_list, batch_count, slow_list, slow_count, stream};
#define RING_CASE(IDX, KK, GG) \
    case IDX: return ebf ? ring_launch_shape<KK, GG, true>(bt, a) : ring_launch_shape<KK, GG, false>(bt, a)
    switch (klass - 1) {
        RING_CASE(0, 5, 8);
        RING_CASE(1, 6, 8);
        RING_CASE(2, 4, 16);
        RING_CASE(3, 6, 16);
        RING_CASE(4, 4, 32);
        RING_CASE(5, 6, 32);
        default: return cudaErrorInvalidValue;
    }
#undef RING_CASE
}
#endif  // POYB200_DEFINE_AFF_RING

}  // namespace poyb200
