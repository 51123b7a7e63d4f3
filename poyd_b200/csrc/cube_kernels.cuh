// Three-sequence cube: algn_fill_cube / backtrack_3d / algn_get_median_3d (src/algn.c:2806-3055, 3829-3905,
// 4160-4173) AS THE REFERENCE EXECUTES THEM.
//
// The reference fill is defective: its neighbour-row pointers advance l2-1 rows per plane instead
// of l2, so row (i, j) of plane i >= 1 reads the linear rows
//     upper = 1 + (i-1)(l2-1) + (j-1),   diag = upper - 1,   prev = i(l2-1) + (j-1)        (row (i,0): diag = (i-1)(l2-1))
// Parity with the reference means reproducing exactly that, so this kernel does.  What the defect leaves intact is the
// structure the GPU needs: a row is an elementwise function of three EARLIER rows plus an in-row min-plus prefix pass
// (the gap-gap-s3 candidate), and every row it reads lies at most l1 + l2 rows back.
//
// One CTA per triple.  The last l1 + l2 + 2 rows live in a per-CTA ring in global memory that stays L2 resident
// (148 CTAs x 0.7 MB for 300^3), read with ld.global.cg because the ring is rewritten; a thread owns the cells k,
// k + blockDim, ...; the in-row pass is a block-wide exclusive prefix-min over value - prefix(gap costs).  Candidates
// carry x8 values with a 3-bit tag in the reference's order of precedence (a later candidate only wins if strictly
// cheaper: P3, P1, P2, S3, S1, S2, SS), so one min resolves value and direction together.
// The direction cube (one byte per cell, the reference's own layout and codes) streams to HBM.
#pragma once
#include "common.cuh"
#include "launch.h"

namespace poyb200 {

// 3-D direction codes, src/matrices.h:34-40, in tag order P3, P1, P2, S3, S1, S2, SS
__device__ __constant__ uint8_t CUBE_CODE[8] = {4, 1, 2, 32, 8, 16, 64, 0};
constexpr int T_P3 = 0, T_P1 = 1, T_P2 = 2, T_S3 = 3, T_S1 = 4, T_S2 = 5, T_SS = 6;
constexpr int CUBE_INF = 0x3fffffff;

// Exclusive prefix-min over the block, in thread order; `carry` (in/out, block-uniform) is the minimum of everything
// that came before this call's elements.  One __syncthreads: the warp totals go through a double-buffered array
// (s_warp[2][32], `parity` flips per call) and are combined with redux.sync.
__device__ __forceinline__ int block_excl_prefix_min(int v, int &carry, int (*s_warp)[32], int &parity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = min(inc, n);
    }
    int exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = CUBE_INF;
    int *buf = s_warp[parity];
    parity ^= 1;
    if (lane == 31) buf[warp] = inc;
    __syncthreads();
    const int mine = (lane < nwarp) ? buf[lane] : CUBE_INF;
    const int before = __reduce_min_sync(0xffffffffu, (lane < warp) ? mine : CUBE_INF);
    const int total = __reduce_min_sync(0xffffffffu, mine);
    const int res = min(exc, min(before, carry));
    carry = min(carry, total);
    return res;
}

__global__ void __launch_bounds__(CUBE_THREADS) cube_fill_kernel(const Task3 *__restrict__ tasks, int ntasks, DevCM3 cm,
                                                                 const uint8_t *__restrict__ pool, int *ring_all,
                                                                 size_t ring_ints, uint8_t *__restrict__ dir,
                                                                 int *__restrict__ out_cost, int want_dir) {
    __shared__ int s_warp[2][32];
    int scan_parity = 0;
    int *ring = ring_all + (size_t) blockIdx.x * ring_ints;
    const int lcm = cm.lcm, gap = cm.gap;
    for (int ti = blockIdx.x; ti < ntasks; ti += gridDim.x) {
        const Task3 t = tasks[ti];
        const uint8_t *s1 = pool + t.off1, *s2 = pool + t.off2, *s3 = pool + t.off3;
        const int l1 = t.l1, l2 = t.l2, l3 = t.l3;
        const int RW = l1 + l2 + 2;
        uint8_t *dcube = dir + t.dir_off;
        const int s3_0 = s3[0];
        auto cost3 = [&](int a, int b, int c) { return 8 * __ldg(cm.cost + ((((a << lcm) + b) << lcm) + c)); };
        int *prefG = ring + (size_t) RW * l3;  // 8 * sum_{1 <= t <= k} cost3[gap][gap][s3[t]], once per triple
        __syncthreads();
        {
            int run = 0;
            for (int k0 = 0; k0 < l3; k0 += blockDim.x) {
                const int k = k0 + threadIdx.x;
                int pg = (k < l3 && k >= 1) ? cost3(gap, gap, s3[k]) : 0;
                const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int n = __shfl_up_sync(0xffffffffu, pg, o);
                    if (lane >= o) pg += n;
                }
                if (lane == 31) s_warp[0][warp] = pg;
                __syncthreads();
                int before = 0, total = 0;
                for (int w = 0; w < (int) (blockDim.x >> 5); w++) {
                    if (w < warp) before += s_warp[0][w];
                    total += s_warp[0][w];
                }
                __syncthreads();
                if (k < l3) __stcg(prefG + k, pg + before + run);
                run += total;
            }
        }
        __syncthreads();
        if (l3 <= (int) blockDim.x) {
            // ---- one cell per thread, software-pipelined ------------------------------------------------------
            // From plane 2 on a row only reads rows written at least 3 rows earlier, so (a) the loads of row r+1 are
            // issued at the start of row r and overlap its whole computation, and (b) the only barrier of a row is the
            // one inside the prefix scan -- it also publishes the previous rows' stores.
            const int k = threadIdx.x;
            const bool live = k < l3;
            const int c3 = live ? s3[k] : s3_0;
            const int pg = live ? __ldcg(prefG + k) : 0;
            const int nrows = l1 * l2;
            auto slot_of = [&](long long x) { return (int) (x % RW); };
            auto inc = [&](int sl) { return (sl + 1 == RW) ? 0 : sl + 1; };
            // in[0..5] = U[k], P[k], D[k], U[k-1], P[k-1], D[k-1] (unused ones stay 0)
            auto issue = [&](int kind, int su, int sd, int sp, int (&in)[6]) {
                in[0] = in[1] = in[2] = in[3] = in[4] = in[5] = 0;
                if (!live) return;
                const int *U = ring + (size_t) su * l3, *D = ring + (size_t) sd * l3, *P = ring + (size_t) sp * l3;
                if (kind == 1) { in[1] = __ldcg(P + k); if (k >= 1) in[4] = __ldcg(P + k - 1); }
                else if (kind == 2) { in[2] = __ldcg(D + k); if (k >= 1) in[5] = __ldcg(D + k - 1); }
                else if (kind == 3) {
                    in[0] = __ldcg(U + k); in[1] = __ldcg(P + k); in[2] = __ldcg(D + k);
                    if (k >= 1) { in[3] = __ldcg(U + k - 1); in[4] = __ldcg(P + k - 1); in[5] = __ldcg(D + k - 1); }
                }
            };
            int i = 0, j = 0, slot_r = 0, slot_u = 0, slot_d = 0, slot_p = 0;
            int pre[6];
            bool have = false;
            int c_a_g_k = 0, c_a_g_0 = 0;  // per plane: cost3(a, gap, s3[k]), cost3(a, gap, s3[0])
            for (int r = 0; r < nrows; r++) {
                int *mm = ring + (size_t) slot_r * l3;
                const int a = s1[i], b = s2[j];
                int kind;
                if (i == 0) {
                    kind = (j == 0) ? 0 : 1;
                    slot_p = (slot_r == 0) ? RW - 1 : slot_r - 1;
                } else if (j == 0) {
                    kind = 2;
                    const long long u = 1 + (long long) (i - 1) * (l2 - 1);
                    slot_u = slot_of(u);
                    slot_d = slot_of(u - 1);
                    slot_p = slot_of((long long) i * (l2 - 1));
                    c_a_g_k = cost3(a, gap, c3);
                    c_a_g_0 = cost3(a, gap, s3_0);
                } else {
                    kind = 3;
                }
                int in[6];
                if (have) {
#pragma unroll
                    for (int q = 0; q < 6; q++) in[q] = pre[q];
                } else {
                    issue(kind, slot_u, slot_d, slot_p, in);
                }
                // prefetch the next row's inputs (both rows in planes >= 2)
                have = false;
                if (i >= 2 && r + 1 < nrows) {
                    if (j + 1 < l2) {
                        const bool adv = (kind == 3);
                        issue(3, adv ? inc(slot_u) : slot_u, adv ? inc(slot_d) : slot_d, adv ? inc(slot_p) : slot_p, pre);
                    } else {
                        issue(2, 0, slot_of((long long) i * (l2 - 1)), 0, pre);
                    }
                    have = true;
                }
                int base = CUBE_INF;
                if (live) {
                    if (kind == 0) {
                        base = (k == 0) ? (0 | T_S2) : CUBE_INF;
                    } else if (kind == 1) {
                        base = in[1] + cost3(gap, b, s3_0) + T_P1;
                        if (k >= 1) base = min(base, in[4] + cost3(gap, b, c3) + T_S1);
                    } else if (kind == 2) {
                        base = in[2] + c_a_g_0 + T_P3;
                        if (k >= 1) base = min(base, in[5] + c_a_g_k + T_S3);
                    } else {
                        base = in[0] + c_a_g_0 + T_P3;
                        base = min(base, in[1] + cost3(gap, b, s3_0) + T_P1);
                        base = min(base, in[2] + cost3(a, b, s3_0) + T_P2);
                        if (k >= 1) {
                            base = min(base, in[3] + c_a_g_k + T_S3);
                            base = min(base, in[4] + cost3(gap, b, c3) + T_S1);
                            base = min(base, in[5] + cost3(a, b, c3) + T_S2);
                        }
                    }
                }
                int carry = CUBE_INF;
                const int v = live ? ((base & ~7) - pg) : CUBE_INF;
                const int e = block_excl_prefix_min(v, carry, s_warp, scan_parity);
                if (live) {
                    int fin = base & ~7, tag = base & 7;
                    if (e < v) { fin = e + pg; tag = T_SS; }
                    __stcg(mm + k, fin);
                    if (want_dir) __stcs(dcube + (size_t) r * l3 + k, CUBE_CODE[tag]);
                    if (r == nrows - 1 && k == l3 - 1) out_cost[t.triple] = fin >> 3;
                }
                if (i < 2) __syncthreads();  // planes 0 and 1 read rows written one or two rows earlier
                if (kind == 3) { slot_u = inc(slot_u); slot_d = inc(slot_d); slot_p = inc(slot_p); }
                slot_r = inc(slot_r);
                if (++j == l2) { j = 0; i++; }
            }
            __syncthreads();
            continue;
        }
        // Row bookkeeping without 64-bit divisions: (i, j) and the ring slots of the current row and of its three
        // neighbour rows advance by one per row; slot(x) = x mod RW is kept incrementally.
        const int nrows = l1 * l2;  // <= 2^28
        auto slot_of = [&](long long x) { return (int) (x % RW); };
        int i = 0, j = 0, slot_r = 0;
        int slot_u = 0, slot_d = 0, slot_p = 0;  // valid from plane 1 on
        for (int r = 0; r < nrows; r++) {
            int *mm = ring + (size_t) slot_r * l3;
            // neighbour rows and per-row constants (a = s1[i], b = s2[j])
            const int a = s1[i], b = s2[j];
            const int *U = nullptr, *D = nullptr, *P = nullptr;
            int kind;
            if (i == 0) {
                kind = (j == 0) ? 0 : 1;
                if (j > 0) P = ring + (size_t) (slot_r == 0 ? RW - 1 : slot_r - 1) * l3;
            } else if (j == 0) {
                kind = 2;
                // start of plane i: upper = 1 + (i-1)(l2-1), diag = upper - 1, prev = i(l2-1)  (for its row j = 1)
                const long long u = 1 + (long long) (i - 1) * (l2 - 1);
                slot_u = slot_of(u);
                slot_d = slot_of(u - 1);
                slot_p = slot_of((long long) i * (l2 - 1));
                D = ring + (size_t) slot_d * l3;
            } else {
                kind = 3;
                U = ring + (size_t) slot_u * l3;
                D = ring + (size_t) slot_d * l3;
                P = ring + (size_t) slot_p * l3;
            }
            const int c_a_g_0 = cost3(a, gap, s3_0), c_g_b_0 = cost3(gap, b, s3_0), c_a_b_0 = cost3(a, b, s3_0);
            int carry = CUBE_INF;   // min over earlier cells of (clean value - prefix of gap-gap costs)
            for (int k0 = 0; k0 < l3; k0 += blockDim.x) {
                const int k = k0 + threadIdx.x;
                const bool live = k < l3;
                const int c3 = live ? s3[k] : s3_0;
                const int pg = live ? __ldcg(prefG + k) : 0;
                int base = CUBE_INF;  // tagged x8 value of the best non-SS candidate
                if (live) {
                    if (kind == 0) {
                        base = (k == 0) ? (0 | T_S2) : CUBE_INF;  // the row is a pure prefix of gap costs (:2925-2928)
                    } else if (kind == 1) {
                        base = __ldcg(P + k) + c_g_b_0 + T_P1;
                        if (k >= 1) base = min(base, __ldcg(P + k - 1) + cost3(gap, b, c3) + T_S1);
                    } else if (kind == 2) {
                        base = __ldcg(D + k) + c_a_g_0 + T_P3;
                        if (k >= 1) base = min(base, __ldcg(D + k - 1) + cost3(a, gap, c3) + T_S3);
                    } else {
                        base = __ldcg(U + k) + c_a_g_0 + T_P3;
                        base = min(base, __ldcg(P + k) + c_g_b_0 + T_P1);
                        base = min(base, __ldcg(D + k) + c_a_b_0 + T_P2);
                        if (k >= 1) {
                            base = min(base, __ldcg(U + k - 1) + cost3(a, gap, c3) + T_S3);
                            base = min(base, __ldcg(P + k - 1) + cost3(gap, b, c3) + T_S1);
                            base = min(base, __ldcg(D + k - 1) + cost3(a, b, c3) + T_S2);
                        }
                    }
                }
                // in-row pass: final[k] = min(base[k], final[k-1] + gg[k])  (:3039-3046), as a prefix-min
                const int v = live ? ((base & ~7) - pg) : CUBE_INF;
                const int e = block_excl_prefix_min(v, carry, s_warp, scan_parity);
                if (live) {
                    int fin = base & ~7, tag = base & 7;
                    if (e < v) {
                        fin = e + pg;
                        tag = T_SS;
                    }
                    __stcg(mm + k, fin);
                    if (want_dir) __stcs(dcube + (size_t) r * l3 + k, CUBE_CODE[tag]);
                    if (r == nrows - 1 && k == l3 - 1) out_cost[t.triple] = fin >> 3;
                }
            }
            __syncthreads();  // the row is complete and visible before any later row reads it
            if (kind == 3) {
                slot_u = (slot_u + 1 == RW) ? 0 : slot_u + 1;
                slot_d = (slot_d + 1 == RW) ? 0 : slot_d + 1;
                slot_p = (slot_p + 1 == RW) ? 0 : slot_p + 1;
            }
            slot_r = (slot_r + 1 == RW) ? 0 : slot_r + 1;
            if (++j == l2) { j = 0; i++; }
        }
    }
}

// backtrack_3d (:3829-3905) + algn_get_median_3d (:4160-4173), one thread per triple.  status = 1 (nothing produced)
// when the reference's walk would index a sequence below 0 -- the reference does not check and reads out of bounds.
__global__ void __launch_bounds__(128) cube_traceback_kernel(const Task3 *__restrict__ tasks, int ntasks, DevCM3 cm,
                                                             const uint8_t *__restrict__ pool,
                                                             const uint8_t *__restrict__ dir, Out3 out) {
    const int ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= ntasks) return;
    const Task3 t = tasks[ti];
    const uint8_t *s1 = pool + t.off1, *s2 = pool + t.off2, *s3 = pool + t.off3;
    const uint8_t *d = dir + t.dir_off;
    const int cap = (int) out.stride, gap = cm.gap;
    const size_t row = (size_t) t.triple * out.stride;
    uint8_t *r1 = out.r1 + row, *r2 = out.r2 + row, *r3 = out.r3 + row;
    long long p = (long long) t.l1 * t.l2 * t.l3 - 1;
    const long long plane = (long long) t.l2 * t.l3, line = t.l3;
    int i1 = t.l1 - 1, i2 = t.l2 - 1, i3 = t.l3 - 1, n = 0, status = 0;
    const int limit = t.l1 + t.l2 + t.l3;
    int last1 = gap, last2 = gap, last3 = gap;
    while (p > 0) {
        const int v = __ldg(d + p);
        int u1 = 0, u2 = 0, u3 = 0;
        if (v & 16) { u1 = u2 = u3 = 1; p -= plane + line + 1; }
        else if (v & 32) { u1 = u3 = 1; p -= plane + 1; }
        else if (v & 8) { u2 = u3 = 1; p -= line + 1; }
        else if (v & 4) { u1 = 1; p -= plane; }
        else if (v & 64) { u3 = 1; p -= 1; }
        else if (v & 1) { u2 = 1; p -= line; }
        else { u1 = u2 = 1; p -= plane + line; }
        if ((u1 && i1 < 0) || (u2 && i2 < 0) || (u3 && i3 < 0) || n >= limit) { status = 1; break; }
        const int e1 = u1 ? s1[i1--] : gap, e2 = u2 ? s2[i2--] : gap, e3 = u3 ? s3[i3--] : gap;
        if (n == 0) { last1 = e1; last2 = e2; last3 = e3; }
        n++;
        if (out.want & POYB200_WANT3_ALIGNED) { r1[cap - n] = (uint8_t) e1; r2[cap - n] = (uint8_t) e2; r3[cap - n] = (uint8_t) e3; }
    }
    if (status) n = 0;
    if ((out.want & POYB200_WANT3_MEDIAN) && n > 0) {
        // the reference's loop never moves its pointers: n copies of the median of the last column
        const int m = __ldg(cm.median + ((((last1 << cm.lcm) + last2) << cm.lcm) + last3));
        uint8_t *md = out.median + row;
        for (int k = 1; k <= n; k++) md[cap - k] = (uint8_t) m;
    }
    out.out_len[t.triple] = n;
    out.status[t.triple] = status;
}

}  // namespace poyb200
