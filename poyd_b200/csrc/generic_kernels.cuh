// Generic fill kernels: one warp per pair, DP state (one entry per diagonal of the stripe) in a per-warp
// global-memory scratch that stays L1/L2 resident.  They accept any stripe width and any sequence length, so
// they are the catch-all behind the register-resident stripe kernels (stripe_kernels.cuh) -- there is no CPU
// fallback anywhere in the product.
#pragma once
#include "cells.cuh"

namespace poyb200 {

// ceil(a / 2) for possibly negative a
__device__ __forceinline__ int ceil_half(int a) { return (a + 1) >> 1; }

// affine_3 fill.  state: int4 (cb, ev, eh, eb) per diagonal, index x = d - dlo + 1, with x = 0 the reference's
// left-edge cells (src/algn.c:2476-2494) and x = W + 1 its poisoned cells (:2528-2536).
template <bool BT>
__global__ void __launch_bounds__(128) aff_generic_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm,
                                                          const uint8_t *__restrict__ pool, int4 *state_all,
                                                          int state_stride, uint8_t *__restrict__ dir,
                                                          int *__restrict__ out_cost) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    int4 *S = state_all + (size_t) warp * state_stride;
    const int gap = cm.gap, go = cm.gap_open;
    for (int ti = warp; ti < ntasks; ti += nwarps) {
        const Task t = tasks[ti];
        const uint8_t *si = pool + t.off_r, *sj = pool + t.off_c;
        const int nr = t.lr - 1, nc = t.lc - 1, dlo = t.dlo, dhi = t.dhi;
        uint8_t *dbase = dir + t.dir_off;
        for (int T = 0; T <= nr + nc; T++) {
            int ilo = max(max(0, T - nc), ceil_half(T - dhi - 1));
            int ihi = min(min(nr, T), (T - dlo + 1) >> 1);
            for (int i = ilo + lane; i <= ihi; i += 32) {
                const int j = T - i, d = j - i, x = d - dlo + 1;
                int4 nv;
                if (i == 0 && j == 0) {
                    nv = make_int4(0, go, go, 0);  // :2194-2198
                } else if (i == 0) {
                    const int r = S[x - 1].z + __ldg(cm.prepend + sj[j]);  // :2212-2217
                    nv = make_int4(r, HIGH_NUM, r, HIGH_NUM);
                } else if (j == 0 || x == 0) {
                    const int ci = si[i], pi = si[i - 1];
                    const AffRow r = aff_make_row(ci, pi, i, gap, go, cm_cost(cm, ci, gap));
                    nv = make_int4(HIGH_NUM, S[x + 1].y + r.vx, HIGH_NUM, HIGH_NUM);  // :2486-2494
                } else if (d == dhi + 1) {
                    nv = make_int4(HIGH_NUM, HIGH_NUM, HIGH_NUM, HIGH_NUM);
                } else {
                    const int ci = si[i], pi = si[i - 1], cj = sj[j], pj = sj[j - 1];
                    const AffRow r = aff_make_row(ci, pi, i, gap, go, cm_cost(cm, ci, gap));
                    const AffCol c = aff_make_col(cj, pj, j, gap, go, __ldg(cm.prepend + cj));
                    const int4 L = S[x - 1], U = S[x + 1], D = S[x];
                    const int dcost = cm_cost(cm, r.lut, c.lut);
                    const int byte = aff_cell<BT>(L.z, L.x, U.y, U.x, D.x, D.y, D.z, D.w, r, c, dcost, go, nv.x, nv.y,
                                                  nv.z, nv.w);
                    if (BT) dbase[dir_index(t, i, j)] = (uint8_t) byte;
                }
                S[x] = nv;
                if (i == nr && j == nc) {
                    int res = min(min(nv.x, nv.y), min(nv.z, nv.w));
                    if (BT && nr == 0 && nc == 0) res = 0;  // final_cost_matrix[0] :2194
                    out_cost[t.pair] = res;
                }
            }
            __syncwarp();
        }
    }
}

// Linear-gap fill (algn_fill_plane / algn_fill_plane_2).  state: one int per diagonal, x = d - dlo + 1;
// x = 0 and x = W + 1 stay LIN_INF: "no cell there" for the Ukkonen edge cells (src/algn.c:461-524).
template <bool BT>
__global__ void __launch_bounds__(128) lin_generic_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm,
                                                          const uint8_t *__restrict__ pool, int *state_all,
                                                          int state_stride, uint8_t *__restrict__ dir,
                                                          int *__restrict__ out_cost) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    int *S = state_all + (size_t) warp * state_stride;
    const int gap = cm.gap;
    for (int ti = warp; ti < ntasks; ti += nwarps) {
        const Task t = tasks[ti];
        const uint8_t *s1 = pool + t.off_r, *s2 = pool + t.off_c;
        const int nr = t.lr - 1, nc = t.lc - 1, dlo = t.dlo, dhi = t.dhi, W = dhi - dlo + 1;
        const bool full = (t.flags & TF_FULL) != 0;
        uint8_t *dbase = dir + t.dir_off;
        if (lane == 0) {
            S[0] = LIN_INF;
            S[W + 1] = LIN_INF;
        }
        __syncwarp();
        for (int T = 0; T <= nr + nc; T++) {
            int ilo = max(max(0, T - nc), ceil_half(T - dhi));
            int ihi = min(min(nr, T), (T - dlo) >> 1);
            for (int i = ilo + lane; i <= ihi; i += 32) {
                const int j = T - i, d = j - i, x = d - dlo + 1;
                int v, mask;
                if (i == 0 && j == 0) {
                    v = 0;
                    mask = D_ALIGN;  // :587-588
                } else if (i == 0) {
                    v = S[x - 1] + __ldg(cm.prepend + s2[j]);  // :597-598
                    mask = D_INSERT;
                } else if (j == 0) {
                    const int a = s1[i];
                    v = S[x + 1] + (full ? cm_cost(cm, a, gap) : __ldg(cm.tail + a));  // :570 / :659
                    mask = D_DELETE;
                } else {
                    const int a = s1[i], b = s2[j];
                    const int mu = S[x + 1];
                    v = lin_cell(S[x - 1], mu, S[x], cm_cost(cm, a, b), cm_cost(cm, gap, b), cm_cost(cm, a, gap), mask);
                    if (j == nc && (full || d != dhi)) v = lin_last_column(v, mu, __ldg(cm.tail + a), mask);
                }
                S[x] = v;
                if (BT) dbase[dir_index(t, i, j)] = (uint8_t) mask;
                if (i == nr && j == nc) out_cost[t.pair] = v;
            }
            __syncwarp();
        }
    }
}

}  // namespace poyb200
