// Traceback + median extraction, one thread per pair.  A traceback is a serial pointer chase through the
// pair's packed direction band; running 10^5..10^6 of them side by side turns that latency into throughput,
// and the kernel overlaps with the (ALU-bound) fill of the next chunk on another stream.
#pragma once
#include "cells.cuh"
#include "walk.cuh"

namespace poyb200 {

// backtrace_affine (src/algn.c:1983-2097): the standalone traceback of the affine classes whose fill kernel does not walk
// its own pairs (generic kernels, stripes with spare diagonals).  One thread per pair, walk.cuh.
__global__ void __launch_bounds__(128) aff_traceback_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm,
                                                            const uint8_t *__restrict__ pool,
                                                            const uint8_t *__restrict__ dir, OutPtrs out, int *work_counter,
                                                            int wpw) {
  // walkers take pairs from a shared counter, `wpw` per warp at a time: no wave-quantisation tail.  Large batches run 32
  // walkers per warp (throughput); small ones spread over more warps with fewer walkers each, because the walkers of a
  // warp diverge (five modes, different moves) and a lone walker steps several times faster than one of 32.
  for (;;) {
    int base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(work_counter, wpw);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= ntasks) break;
    const int ti = base + (threadIdx.x & 31);
    if ((int) (threadIdx.x & 31) < wpw && ti < ntasks) {
        const Task t = tasks[ti];
        if (!(out.walked && out.walked[t.pair]))  // (a ring kernel may have walked the pair already)
            aff_walk_pair(t, pool, BandBytes(t, dir + t.dir_off), MedianGlobal{cm.median, cm.lcm}, cm, out);
    }
    __syncwarp();
  }
}

// backtrack_2d, linear branch (src/algn.c:3606-3665), fused with algn_ancestor_2 (:4126-4147, the
// cost_model != affine branch, the only one a linear alignment can reach) and
// algn_get_median_2d_with_gaps (:4024-4035) -- what SeqCS.DOS.median asks for (src/seqCS.ml:757-766).
__global__ void __launch_bounds__(128) lin_traceback_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm,
                                                            const uint8_t *__restrict__ pool,
                                                            const uint8_t *__restrict__ dir, OutPtrs out, int *work_counter,
                                                            int wpw) {
  // walkers take pairs from a shared counter, `wpw` per warp at a time: no wave-quantisation tail.  Large batches run 32
  // walkers per warp (throughput); small ones spread over more warps with fewer walkers each, because the walkers of a
  // warp diverge (five modes, different moves) and a lone walker steps several times faster than one of 32.
  for (;;) {
    int base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(work_counter, wpw);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= ntasks) break;
    const int ti = base + (threadIdx.x & 31);
    if ((int) (threadIdx.x & 31) < wpw && ti < ntasks) {
        const Task t = tasks[ti];
        lin_walk_pair(t, pool, LinBand(t, dir + t.dir_off), cm, out);
    }
    __syncwarp();
  }
}

// Stand-alone medians of already aligned pairs: which 0 algn_ancestor_2 (:4126-4147, including
// algn_correct_blocks_affine :4080-4124 for combination alphabets under the affine model), 1
// algn_get_median_2d_with_gaps (:4024), 2 algn_get_median_2d_no_gaps (:4042).  Output right aligned.
__global__ void __launch_bounds__(128) median_2_kernel(int which, DevCM cm, const uint8_t *__restrict__ a,
                                                       const uint8_t *__restrict__ b, long long in_stride,
                                                       const int *__restrict__ len, int n, uint8_t *out,
                                                       long long out_stride, int *out_len) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint8_t *sa = a + (size_t) p * in_stride, *sb = b + (size_t) p * in_stride;
    uint8_t *row = out + (size_t) p * out_stride;
    const int L = len[p], cap = (int) out_stride, gap = cm.gap;
    int pos = cap;
    if (which == 1) {
        for (int i = L - 1; i >= 0; i--) row[--pos] = (uint8_t) cm_median(cm, sa[i], sb[i]);
    } else if (which == 2) {
        for (int i = L - 1; i >= 0; i--) {
            const int m = cm_median(cm, sa[i], sb[i]);
            if (m != gap) row[--pos] = (uint8_t) m;
        }
        row[--pos] = (uint8_t) gap;
    } else if (!cm.combinations || cm.cost_model_type != 1) {
        for (int i = L - 1; i >= 0; i--) {
            const int m = cm_median(cm, sa[i], sb[i]);
            if (m != gap) row[--pos] = (uint8_t) m;
        }
        row[--pos] = (uint8_t) gap;  // nothing equal to gap survived, so "first != gap" always holds
    } else {
        // left-to-right block correction into the left part of the row, then compaction to the right
        int extending_gap = 0, inside_block = 0, prev_block = 0;
        for (int i = 0; i < L; i++) {
            const int ab = sa[i], bb = sb[i];
            int sbv = cm_median(cm, ab, bb);
            if (!inside_block && (!(ab & gap) || !(bb & gap))) inside_block = 0;
            else if (inside_block && (!(ab & gap) || !(bb & gap))) inside_block = 0;
            else if (((ab & gap) || (bb & gap)) && ((ab != gap) || (bb != gap))) inside_block = 1;
            else inside_block = 0;
            if (((gap & ab) || (gap & bb)) && !(sbv & gap) && !extending_gap) {
                prev_block = inside_block;
                extending_gap = 1;
            } else if ((gap & ab) && (gap & bb) && (sbv & gap) && (sbv != gap) && extending_gap && inside_block &&
                       !prev_block) {
                sbv = (~gap) & sbv;
                prev_block = 0;
            } else if ((gap & ab) && (gap & bb) && (1 == extending_gap)) {
                prev_block = inside_block;
                extending_gap = 0;
            }
            row[i] = (uint8_t) sbv;
        }
        for (int i = L - 1; i >= 0; i--) {
            const int v = row[i];
            if (v != gap) row[--pos] = (uint8_t) v;  // pos > i always: cap >= L + 1
        }
        row[--pos] = (uint8_t) gap;
    }
    out_len[p] = cap - pos;
}

// algn_calculate_from_2_aligned (src/algn.c:3311-3371) over already aligned pairs: the sum of matrix[a][b] over the columns
// plus a gap opening whenever a block of gaps starts in either row, with the reference's three-state scan.  `matrix` is
// c->worst (algn_worst_2, :3373) or c->cost (algn_verify_2, :3378).  One thread per pair; rows LEFT aligned.
__global__ void __launch_bounds__(128) calc_aligned_2_kernel(const int *__restrict__ matrix, DevCM cm, const uint8_t *__restrict__ a,
                                                             const uint8_t *__restrict__ b, long long in_stride,
                                                             const int *__restrict__ len, int n, int *__restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint8_t *s1 = a + (size_t) p * in_stride, *s2 = b + (size_t) p * in_stride;
    const int L = len[p], gap = cm.gap, go = cm.gap_open;
    const bool comb = cm.combinations != 0;
    int res = 0, gap_row = 0, i = 0;
    if (L > 0) {
        const int x = s1[0], y = s2[0];
        if (comb ? ((gap & x) && (gap & y)) : (x == gap && y == gap)) i = 1;
    }
    for (; i < L; i++) {
        const int x = s1[i], y = s2[i];
        const bool g1 = comb ? ((x & gap) != 0) : (x == gap), g2 = comb ? ((y & gap) != 0) : (y == gap);
        if (gap_row == 0) {
            if (comb ? (g1 && !g2) : g1) { res += go; gap_row = 1; }
            else if (comb ? (g2 && !g1) : g2) { res += go; gap_row = 2; }
        } else if (gap_row == 1) {
            if (!g1) {
                if (comb ? (g2 && !g1) : g2) { res += go; gap_row = 2; }
                else gap_row = 0;
            }
        } else {
            if (!g2) {
                if (g1) { res += go; gap_row = 1; }
                else gap_row = 0;
            }
        }
        res += __ldg(matrix + (x << cm.lcm) + y);
    }
    out[p] = res;
}

// algn_get_median_3d (src/algn.c:4160-4173) as the reference executes it: the loop never moves its three pointers, so
// the result is len copies of median3[last a][last b][last c].  Rows LEFT aligned in, RIGHT aligned out.
__global__ void __launch_bounds__(128) median_3_kernel(const uint8_t *__restrict__ median3, int lcm, const uint8_t *__restrict__ a,
                                                       const uint8_t *__restrict__ b, const uint8_t *__restrict__ c, long long in_stride,
                                                       const int *__restrict__ len, int n, uint8_t *out, long long out_stride,
                                                       int *out_len) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int L = len[p];
    out_len[p] = L;
    if (L <= 0) return;
    const size_t o = (size_t) p * in_stride + (L - 1);
    const int m = median3[((((size_t) a[o] << lcm) + b[o]) << lcm) + c[o]];
    uint8_t *row = out + (size_t) p * out_stride + (out_stride - L);
    for (int k = 0; k < L; k++) row[k] = (uint8_t) m;
}

// ---- device-resident sequence store (store.cu) --------------------------------------------------------------------------
// Result rows of a batch (right aligned, stride bytes, length in outlen4[4 p]) -> per row {length, number of elements
// carrying the gap bit (seq_CAML_count, src/seq.c:570-582)}.  One warp per row.
__global__ void __launch_bounds__(128) store_row_stats_kernel(const uint8_t *__restrict__ rows, long long stride,
                                                              const int *__restrict__ outlen4, int n, int gap, int *__restrict__ stats) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n) return;
    const int L = outlen4[4 * (size_t) p];
    const uint8_t *src = rows + (size_t) p * stride + (stride - L);
    int c = 0;
    for (int k = lane; k < L; k += 32) c += (src[k] & gap) != 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) { stats[2 * p] = L; stats[2 * p + 1] = c; }
}
// Copies the result rows to their places in the store's pool.  One warp per row.
__global__ void __launch_bounds__(128) store_append_kernel(const uint8_t *__restrict__ rows, long long stride,
                                                           const int *__restrict__ outlen4, const long long *__restrict__ newoff, int n,
                                                           uint8_t *__restrict__ pool) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n) return;
    const int L = outlen4[4 * (size_t) p];
    const uint8_t *src = rows + (size_t) p * stride + (stride - L);
    uint8_t *dst = pool + newoff[p];
    for (int k = lane; k < L; k += 32) dst[k] = src[k];
}
// eq[p] = 1 when the two sequences of pair p have the same length and the same elements (`0 = compare s1 s2`).
// jobs[p] = {offset a, offset b, length a, length b}.
__global__ void __launch_bounds__(128) store_equal_kernel(const uint8_t *__restrict__ pool, const uint4 *__restrict__ jobs, int n,
                                                          uint8_t *__restrict__ eq) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n) return;
    const uint4 jb = jobs[p];
    int same = jb.z == jb.w;
    if (same) {
        const uint8_t *x = pool + jb.x, *y = pool + jb.y;
        for (int k = lane; k < (int) jb.z; k += 32) same &= (x[k] == y[k]);
    }
    same = __all_sync(0xffffffffu, same);
    if (lane == 0) eq[p] = (uint8_t) same;
}
// Sequence.Align.closest s1 s2 when s1 = s2 (src/sequence.ml:1000-1009, combination alphabets): the gap bits past the
// first element are cleared, every element x becomes get_closest x x, gaps are removed and one is prepended.  Writes a
// right-aligned row of `stride` bytes and outlen4[4 p].  One thread per sequence.
// jobs[p] = {offset, length}.
__global__ void __launch_bounds__(128) store_closest_same_kernel(DevCM cm, const uint8_t *__restrict__ pool,
                                                                 const uint2 *__restrict__ jobs, int n, uint8_t *rows, long long stride,
                                                                 int *outlen4) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int L = (int) jobs[p].y, gap = cm.gap;
    const uint8_t *s = pool + jobs[p].x;
    uint8_t *row = rows + (size_t) p * stride;
    int pos = (int) stride;
    for (int k = L - 1; k >= 0; k--) {
        int x = s[k];
        if (k > 0) x &= ~gap;
        const int sel = closest_elem(cm, x, x);
        if (sel != gap) row[--pos] = (uint8_t) sel;
    }
    row[--pos] = (uint8_t) gap;
    outlen4[4 * (size_t) p] = (int) stride - pos;
}

}  // namespace poyb200
