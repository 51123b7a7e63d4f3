// Shared device/host definitions of the sm_100a alignment kernels.
//
// Geometry used by every kernel (DESIGN.md "Stripe formulation"): the cells the reference visits form a
// stripe of diagonals dlo <= j - i <= dhi of the (rows x cols) matrix, swept by anti-diagonals t = i + j.
// A cell on diagonal d reads its left neighbour from diagonal d-1 and its upper neighbour from d+1, both
// written at t-1, and its diagonal neighbour from its own diagonal, written at t-2 -- so one value per
// diagonal is all the DP state there is.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace poyb200 {

constexpr int HIGH_NUM = 1000000;    // src/algn.c:38
constexpr int LIN_INF = 0x30000000;  // "no neighbour" for the linear kernels (never wins a min, never overflows)
constexpr int TMPGAP = 16;           // src/algn.c:1711

// linear direction bits, src/matrices.h:22-27
constexpr int D_ALIGN = 1, D_INSERT = 2, D_DELETE = 4;

// affine direction byte (7 bits): what backtrace_affine (src/algn.c:1983-2097) would decide at this cell
//   bits 1:0  mode after an align step  : 0 horizontal, 1 diagonal, 2 vertical, 3 stay align (priority :2049-2051)
//   bits 3:2  mode entered from m_todo  : 0 horizontal, 1 align, 2 vertical, 3 diagonal     (priority :2006-2012)
//   bit 4 END_HORIZONTAL, bit 5 END_VERTICAL, bit 6 END_BLOCK
// The two 2-bit codes are the tie-break priorities themselves, so the stripe kernels get them for free as the
// low bits of a min over tagged keys (stripe_kernels.cuh).
constexpr int AN_H = 0, AN_D = 1, AN_V = 2, AN_A = 3;
constexpr int AM_H = 0, AM_A = 1, AM_V = 2, AM_D = 3;
constexpr int AB_ENDH = 16, AB_ENDV = 32, AB_ENDB = 64;
constexpr int AFF_LEFT_EDGE_BYTE = (AM_V << 2) | AB_ENDV;   // DO_VERTICAL | END_VERTICAL   (:2476)
constexpr int AFF_RIGHT_EDGE_BYTE = (AM_H << 2) | AB_ENDH;  // DO_HORIZONTAL | END_HORIZONTAL (:2530)

// task flags
constexpr uint32_t TF_ROWS_ARE_B = 1;  // operand b sits on the rows: swap the aligned outputs back
constexpr uint32_t TF_FULL = 2;        // linear: full matrix (algn_fill_plane), no edge rules
constexpr uint32_t TF_SWAPED = 4;      // linear traceback tie flag (backtrack_2d `swaped`)
constexpr uint32_t TF_DIR2 = 8;        // direction band holds 2-bit resolved moves (linear stripe kernels)
constexpr uint32_t TF_DIR6 = 32;       // affine band of aff_fast_kernel<5, 8, .., true>: five 6-bit codes per 32-bit word (no END_BLOCK bit: always set)
constexpr uint32_t TF_ROWMAJ = 16;     // direction band in row-major tiles: lane = j / twoK owns twoK columns (lin_rows_kernels.cuh)

struct Task {
    uint32_t off_r, off_c;  // pool offsets of the row / column sequence
    int32_t lr, lc;         // stored lengths (leading gap included)
    int32_t dlo, dhi;       // stripe of diagonals visited
    uint32_t flags;
    uint32_t pair;          // index in the caller's pair list
    uint64_t dir_off;       // byte offset of this pair's direction band
    uint32_t G, twoK, BL;   // direction addressing, see dir_index
    int32_t dbase;          // diagonal held by lane 0, slot 0 (= dlo for the generic kernels)
    uint32_t klass;         // kernel class chosen by the planner
    int32_t tshift;         // direction addressing counts steps from anti-diagonal `tshift` (0 unless a kernel asks)
};

// Byte index of cell (i, j) inside a pair's direction band: anti-diagonal major, then lane-group chunk.
// Stripe kernels write one BL-byte chunk per lane per step; the generic kernels use G = 1.
// Anti-diagonals are tiled by 8: the 8 chunks a lane writes during steps 8k..8k+7 are contiguous (one 64-byte
// line for BL = 8), so a traceback, which moves one or two anti-diagonals per step and drifts slowly across
// diagonals, stays inside a line for several steps.
// Steps are counted from anti-diagonal t.tshift (<= 0 for every cell of the matrix): the affine stripe kernels choose it
// so that a tile starts where their double-step counter is a multiple of 4, whatever the pair's first diagonal is,
// which lets an unrolled block of 8 double steps store at compile-time offsets.
__host__ __device__ __forceinline__ uint64_t dir_index(const Task &t, int i, int j) {
    const uint32_t dd = (uint32_t) ((j - i) - t.dbase), T = (uint32_t) (i + j - t.tshift);
    const uint32_t lane = dd / t.twoK, m = (dd - lane * t.twoK) >> 1;
    return ((((uint64_t) (T >> 3) * t.G + lane) << 3) + (T & 7)) * t.BL + m;
}
// Direction code of cell (i, j): a byte, or a 2-bit field of a 32-bit chunk (TF_DIR2).
__device__ __forceinline__ int dir_fetch(const Task &t, const uint8_t *dbase, int i, int j) {
    if (t.flags & TF_ROWMAJ) {  // one 32-bit word of 2-bit moves per lane and row, rows tiled by 8
        const uint32_t lane = (uint32_t) j / t.twoK, m = (uint32_t) j - lane * t.twoK;
        const uint64_t chunk = ((((uint64_t) ((uint32_t) i >> 3) * t.G + lane) << 3) + ((uint32_t) i & 7)) * 4;
        return (__ldg(dbase + chunk + (m >> 2)) >> ((m & 3) * 2)) & 3;
    }
    const uint32_t dd = (uint32_t) ((j - i) - t.dbase), T = (uint32_t) (i + j - t.tshift);
    const uint32_t lane = dd / t.twoK, m = (dd - lane * t.twoK) >> 1;
    const uint64_t chunk = ((((uint64_t) (T >> 3) * t.G + lane) << 3) + (T & 7)) * t.BL;
    if (t.flags & TF_DIR2) return (__ldg(dbase + chunk + (m >> 2)) >> ((m & 3) * 2)) & 3;
    return __ldg(dbase + chunk + m);
}
// Bytes of one pair's direction band (steps 0 .. lr + lc - 2 - tshift).
__host__ __device__ __forceinline__ uint64_t dir_bytes(const Task &t) {
    if (t.flags & TF_ROWMAJ) return (uint64_t) ((t.lr + 7) >> 3) * 8 * t.G * 4;
    return (uint64_t) ((t.lr + t.lc - 1 - t.tshift + 7) >> 3) * 8 * t.G * t.BL;
}

struct DevCM {
    int a_sz, lcm, gap, cost_model_type, combinations, gap_open;
    const int *cost;
    const uint8_t *median;
    const int *prepend, *tail;
    int all_elements;
    const int *worst;  // may be null (only algn_worst_2 reads it)
};

struct OutPtrs {
    int *cost;
    uint8_t *median, *medianwg, *al_a, *al_b;
    int *out_len;
    long long stride;
    uint32_t want;
    uint8_t *bits_a, *bits_b, *bits_wg;  // POYB200_WANT_BITSETS rows
    long long bstride;                   // bytes per bitset row (multiple of 4)
    uint8_t *walked;                     // per pair: 1 once a ring kernel has walked it (the traceback kernel skips it); may be null
};

__device__ __forceinline__ int cm_cost(const DevCM &c, int a, int b) { return __ldg(c.cost + (a << c.lcm) + b); }
__device__ __forceinline__ int cm_median(const DevCM &c, int a, int b) { return __ldg(c.median + (a << c.lcm) + b); }

// One column of Sequence.Align.closest (src/sequence.ml:986-1018): a = element of the aligned s1 (the parent), b = element
// of the aligned s2.  Combination alphabets: Cost_matrix.Two_D.get_closest (src/cost_matrix.ml:681-700) -- b loses its gap
// bit (or collapses to the gap when both carry it), then the lowest set bit of b with the strictly smallest cost a x wins.
// Other alphabets: b, unless it is the `all` code, then a (or 1 when a is `all` too).
__device__ __forceinline__ int closest_elem(const DevCM &c, int a, int b) {
    if (!c.combinations) return (b == c.all_elements) ? ((a == c.all_elements) ? 1 : a) : b;
    const int gap = c.gap;
    if (a != gap && b != gap) b = ((a & gap) && (b & gap)) ? gap : (b & ~gap);
    int best = a, cur = 0x7fffffff;
    for (int bit = 0; bit < c.lcm; bit++) {
        const int x = 1 << bit;
        if (b & x) {
            const int nc = cm_cost(c, a, x);
            if (nc < cur) { best = x; cur = nc; }
        }
    }
    return best;
}

}  // namespace poyb200
