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

namespace poyb200 {

struct Task3 {
    uint32_t off1, off2, off3;
    int32_t l1, l2, l3;      // stored lengths (leading gap included)
    uint32_t triple;         // index in the caller's list
    uint32_t pad;
    uint64_t dir_off;        // byte offset of this triple's direction cube
};

struct DevCM3 {
    int lcm, gap;
    const int *cost;         // (1 << lcm)^3
    const uint8_t *median;
};

struct Out3 {
    int *cost, *out_len, *status;
    uint8_t *r1, *r2, *r3, *median;
    long long stride;
    uint32_t want;
};

constexpr int CUBE_THREADS = 512;
// 3-D direction codes, src/matrices.h:34-40, in tag order P3, P1, P2, S3, S1, S2, SS
__device__ __constant__ uint8_t CUBE_CODE[8] = {4, 1, 2, 32, 8, 16, 64, 0};
constexpr int T_P3 = 0, T_P1 = 1, T_P2 = 2, T_S3 = 3, T_S1 = 4, T_S2 = 5, T_SS = 6;
constexpr int CUBE_INF = 0x3fffffff;

// Exclusive prefix-min over the block, in thread order; `carry` (in/out, block-uniform) is the minimum of everything
// that came before this call's elements.  Two __syncthreads.
__device__ __forceinline__ int block_excl_prefix_min(int v, int &carry, int *s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = min(inc, n);
    }
    int exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = CUBE_INF;
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int before = carry, total = carry;
    for (int w = 0; w < nwarp; w++) {
        const int t = s_warp[w];
        if (w < warp) before = min(before, t);
        total = min(total, t);
    }
    __syncthreads();
    carry = total;
    return min(exc, before);
}

__global__ void __launch_bounds__(CUBE_THREADS) cube_fill_kernel(const Task3 *__restrict__ tasks, int ntasks, DevCM3 cm,
                                                                 const uint8_t *__restrict__ pool, int *ring_all,
                                                                 size_t ring_ints, uint8_t *__restrict__ dir,
                                                                 int *__restrict__ out_cost, int want_dir) {
    __shared__ int s_warp[CUBE_THREADS / 32];
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
        auto rowp = [&](long long r) { return ring + (size_t) (r % RW) * l3; };
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
                if (lane == 31) s_warp[warp] = pg;
                __syncthreads();
                int before = 0, total = 0;
                for (int w = 0; w < (int) (blockDim.x >> 5); w++) {
                    if (w < warp) before += s_warp[w];
                    total += s_warp[w];
                }
                __syncthreads();
                if (k < l3) __stcg(prefG + k, pg + before + run);
                run += total;
            }
        }
        __syncthreads();
        for (long long r = 0; r < (long long) l1 * l2; r++) {
            const int i = (int) (r / l2), j = (int) (r % l2);
            int *mm = rowp(r);
            // neighbour rows and per-row constants (a = s1[i], b = s2[j])
            const int a = s1[i], b = s2[j];
            const int *U = nullptr, *D = nullptr, *P = nullptr;
            int kind;
            if (i == 0) {
                kind = (j == 0) ? 0 : 1;
                if (j > 0) P = rowp(r - 1);
            } else if (j == 0) {
                kind = 2;
                D = rowp((long long) (i - 1) * (l2 - 1));
            } else {
                kind = 3;
                const long long u = 1 + (long long) (i - 1) * (l2 - 1) + (j - 1);
                U = rowp(u);
                D = rowp(u - 1);
                P = rowp((long long) i * (l2 - 1) + (j - 1));
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
                const int e = block_excl_prefix_min(v, carry, s_warp);
                if (live) {
                    int fin = base & ~7, tag = base & 7;
                    if (e < v) {
                        fin = e + pg;
                        tag = T_SS;
                    }
                    __stcg(mm + k, fin);
                    if (want_dir) __stcs(dcube + (size_t) r * l3 + k, CUBE_CODE[tag]);
                    if (r == (long long) l1 * l2 - 1 && k == l3 - 1) out_cost[t.triple] = fin >> 3;
                }
            }
            __syncthreads();  // the row is complete and visible before any later row reads it
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
