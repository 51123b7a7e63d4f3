// Kernel classes: which fill kernel a pair takes and how its direction band is addressed.  Host-side planner logic shared
// by api.cu and the kernel translation units; no device code.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace poyb200 {

constexpr uint32_t KLASS_GENERIC = 0;
// klass = 1 + index into this table (affine stripe shapes: a group of G lanes owns 2 K G diagonals) when the stripe is flush
// with dlo, KLASS_AFF_SPARE + index when the shape leaves spare diagonals below dlo (aff_stripe_kernel's LOW variant).
struct StripeShape {
    int K, G;
};
constexpr StripeShape AFF_SHAPES[] = {{5, 8}, {6, 8}, {4, 16}, {6, 16}, {4, 32}, {6, 32}, {8, 32}};
constexpr int N_AFF_SHAPES = sizeof(AFF_SHAPES) / sizeof(AFF_SHAPES[0]);
constexpr uint32_t KLASS_AFF_SPARE = 17;
// Longest operand a stripe kernel stages whole in shared memory: 4 warps x (32 / G) pairs x 2 operands must fit next to
// the tables, the walk windows and the staging barriers (one ring slot; ~170 KB of the 227 KB are left for operands).
// Past 2048 elements a CTA's operands cost resident CTAs -- occupancy falls with the length -- but every pair up to these
// caps still runs in a register kernel instead of the generic one (state in global memory, an order of magnitude slower).
__host__ __device__ constexpr int stripe_max_seq_bytes(int G) { return G >= 32 ? 16384 : (G >= 16 ? 10240 : 5120); }

struct LinShape {
    int K, G;
};
constexpr LinShape LIN_SHAPES[] = {{8, 8}, {10, 8}, {12, 8}, {8, 16}, {12, 16}, {8, 32}, {10, 32}, {16, 32}};
constexpr int N_LIN_SHAPES = sizeof(LIN_SHAPES) / sizeof(LIN_SHAPES[0]);
constexpr uint32_t KLASS_LIN_BASE = 32;  // klass = KLASS_LIN_BASE + shape index
constexpr int LIN_MAX_LCM = 6;
// Column-striped kernel for full linear matrices (lin_rows_kernels.cuh): a group of G lanes, C columns per lane.
struct LinRowShape {
    int C, G;
};
constexpr LinRowShape LIN_ROW_SHAPES[] = {{8, 8}, {12, 8}, {16, 8}, {10, 16}, {12, 16}, {16, 16}, {10, 32}, {12, 32}, {16, 32}};
constexpr int N_LIN_ROW_SHAPES = sizeof(LIN_ROW_SHAPES) / sizeof(LIN_ROW_SHAPES[0]);
constexpr uint32_t KLASS_LINROW_BASE = 48;  // klass = KLASS_LINROW_BASE + shape index (classes stay below 64)

// Chooses a stripe shape for an affine pair; returns false when the pair must take the generic kernel.
static inline bool stripe_choose(Task &t, bool affine, int W, const DevCM &cm) {
    if (!affine) return false;
    if (cm.lcm != 5 || cm.gap != 16) return false;  // aff_cell_dna assumes the nucleotide encoding
    for (int s = 0; s < N_AFF_SHAPES; s++) {
        const int K = AFF_SHAPES[s].K, G = AFF_SHAPES[s].G;
        if (2 * K * G >= W + 1) {
            if (std::max(t.lr, t.lc) > stripe_max_seq_bytes(G)) continue;  // a wider group stages fewer pairs per CTA
            t.G = G;
            t.twoK = 2 * K;
            t.BL = (K <= 4) ? 4 : 8;
            t.dbase = t.dhi + 2 - 2 * K * G;
            t.klass = (t.dlo - t.dbase > 0) ? KLASS_AFF_SPARE + s : 1 + s;
            // steps are counted from a multiple of 8 double-step halves: see dir_index and aff_fast_kernels.cuh
            t.tshift = t.dbase + ((2 * ((-t.dbase) >> 1)) & ~7);
            return true;
        }
    }
    return false;
}

// Shape index of an affine stripe class (either kind), or -1.
static inline int aff_shape_of(uint32_t klass) {
    if (klass >= 1 && klass < 1 + (uint32_t) N_AFF_SHAPES) return (int) klass - 1;
    if (klass >= KLASS_AFF_SPARE && klass < KLASS_AFF_SPARE + (uint32_t) N_AFF_SHAPES) return (int) (klass - KLASS_AFF_SPARE);
    return -1;
}
// True when the class has a ring kernel (fill + traceback in one kernel): no spare diagonals, rings of <= 7 slots.
static inline bool ring_has_shape(uint32_t klass) {
    const int s = (int) klass - 1;
    return s >= 0 && s < N_AFF_SHAPES && AFF_SHAPES[s].K <= 6;
}
// True when the class has a fast kernel (aff_fast_kernel, the ring kernels' predecessor).
static inline bool fast_has_shape(uint32_t klass) { return ring_has_shape(klass); }

// Full matrices whose shorter operand (the columns) fits one group: the narrowest shape that covers the columns.
static inline bool lin_rows_choose(Task &t, const DevCM &cm) {
    if (!(t.flags & TF_FULL) || cm.lcm > LIN_MAX_LCM) return false;
    for (int s = 0; s < N_LIN_ROW_SHAPES; s++) {
        const int C = LIN_ROW_SHAPES[s].C, G = LIN_ROW_SHAPES[s].G;
        if (C * G < t.lc) continue;
        if (std::max(t.lr, t.lc) > stripe_max_seq_bytes(G)) continue;
        t.klass = KLASS_LINROW_BASE + s;
        t.G = G;
        t.twoK = C;
        t.BL = 4;
        t.dbase = 0;
        t.flags |= TF_DIR2 | TF_ROWMAJ;
        return true;
    }
    return false;
}

static inline bool lin_stripe_choose(Task &t, int W, const DevCM &cm) {
    if (cm.lcm > LIN_MAX_LCM) return false;
    for (int s = 0; s < N_LIN_SHAPES; s++) {
        const int K = LIN_SHAPES[s].K, G = LIN_SHAPES[s].G;
        if (2 * K * G >= W) {
            if (std::max(t.lr, t.lc) > stripe_max_seq_bytes(G)) continue;
            const int d0 = t.dhi + 1 - 2 * K * G;
            t.klass = KLASS_LIN_BASE + s;
            t.G = G;
            t.twoK = 2 * K;
            t.BL = 4;
            t.dbase = d0;
            t.flags |= TF_DIR2;
            return true;
        }
    }
    return false;
}

}  // namespace poyb200
