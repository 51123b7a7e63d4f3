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

namespace poyb200 {

constexpr uint32_t KLASS_GENERIC = 0;
// klass = 1 + index into this table (affine stripe shapes).  LOW is chosen at run time per pair.
struct StripeShape {
    int K, G;
};
constexpr StripeShape AFF_SHAPES[] = {{5, 8}, {6, 8}, {4, 16}, {6, 16}, {4, 32}, {6, 32}, {8, 32}};
constexpr int N_AFF_SHAPES = sizeof(AFF_SHAPES) / sizeof(AFF_SHAPES[0]);
constexpr int STRIPE_MAX_SEQ_BYTES = 2048;  // per operand, staged in shared memory

constexpr int STRIPE_WARPS = 4;

struct AffWinRow {
    int vx, gopge, gop, lut, gm;  // lut = (si & 15) << 4, gm = -(si has the gap bit)
};
struct AffWinCol {
    int hx, gopg, gop, lut, gm;  // lut = sj & 15
};

// Fast form of aff_cell for DNA matrices (lcm = 5, gap = 16 = TMPGAP, codes < 32), where "has the gap bit",
// "si_base != si_no_gap" and "si & TMPGAP" are the same predicate.
template <bool BT>
__device__ __forceinline__ int aff_cell_dna(int ehl, int cbl, int evu, int cbu, int cbd, int evd, int ehd, int ebd,
                                            const AffWinRow &r, const AffWinCol &c, int dcost, int go, int &cb, int &ev,
                                            int &eh, int &eb) {
    int byte;
    int x = ehl + c.hx, y = cbl + c.gopg;
    eh = min(x, y);
    byte = (x < y) ? 0 : AB_ENDH;
    x = evu + r.vx;
    y = cbu + r.gopge;
    ev = min(x, y);
    byte |= (x < y) ? 0 : AB_ENDV;
    const int bothm = r.gm & c.gm;          // -1 when both carry the gap bit
    const int dg = HIGH_NUM & ~bothm;       // 0 or HIGH_NUM
    if (BT) {
        eb = min(ebd, cbd) + dg;            // ext and open share the addend (:1871-1872)
        byte |= (ebd < cbd) ? 0 : AB_ENDB;
    } else {
        const int odg = dg | ((2 * go) & bothm);  // both ? 2*go : HIGH_NUM  (:1846)
        eb = min(ebd + dg, cbd + odg);
    }
    const int a0 = cbd + dcost;
    const int a1 = evd + dcost + (c.gop & r.gm);
    const int a2 = ehd + dcost + (r.gop & c.gm);
    const int a3 = ebd + dcost + max(r.gop, c.gop);
    cb = min(min(a0, a1), min(a2, a3));
    if (BT) {
        const int nxt = (a2 == cb) ? AN_H : (a3 == cb) ? AN_D : (a1 == cb) ? AN_V : AN_A;
        const int f = min(min(eh, ev), min(eb, cb));
        const int mode = (eh == f) ? AM_H : (cb == f) ? AM_A : (ev == f) ? AM_V : AM_D;
        byte |= mode | (nxt << 2);
    }
    return byte;
}

template <int K, int G, bool BT, bool LOW>
struct AffStripe {
    static constexpr int Q = 2 * K;
    static constexpr int BL = (K <= 4) ? 4 : 8;

    // per-lane state
    int cb[Q], ev[Q], eh[Q], eb[Q];
    AffWinRow R[K];
    AffWinCol C[K + 1];
    int last_ci, last_cj;  // codes of the newest row / column in the windows
    // per-pair constants
    const uint8_t *si, *sj;  // shared memory copies
    const int *lut, *prep, *get;
    int nr, nc, go, lane, qlow;
    unsigned gmask;

    __device__ __forceinline__ AffWinRow make_row(int i) {
        const int ii = min(max(i, 0), nr);
        const int ci = si[ii], pi = last_ci;
        last_ci = ci;
        AffWinRow r;
        const int ge = get[ci];
        r.gop = (!(pi & 16) && (ci & 16)) ? 0 : go;
        r.vx = (i > 1 && (pi & 16) && !(ci & 16)) ? r.gop + ge : ge;
        r.gopge = r.gop + ge;
        r.lut = (ci & 15) << 4;
        r.gm = -((ci >> 4) & 1);
        return r;
    }
    __device__ __forceinline__ AffWinCol make_col(int j) {
        const int jj = min(max(j, 0), nc);
        const int cj = sj[jj], pj = last_cj;
        last_cj = cj;
        AffWinCol c;
        const int g = prep[cj];
        c.gop = (!(pj & 16) && (cj & 16)) ? 0 : go;
        c.hx = ((pj & 16) && !(cj & 16) && j != 1) ? c.gop + g : g;
        c.gopg = c.gop + g;
        c.lut = cj & 15;
        c.gm = -((cj >> 4) & 1);
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
                ncb = HIGH_NUM; neh = HIGH_NUM; neb = HIGH_NUM;
                nev = evu + r.vx;
            }
        }
        if (BOUNDARY) {
            if (i == 0) {
                if (j == 0) {  // :2194-2198
                    ncb = 0; neb = 0; neh = go; nev = go;
                } else {       // :2212-2217
                    const int rr = ehl + (c.gopg - c.gop);
                    neh = rr; ncb = rr; nev = HIGH_NUM; neb = HIGH_NUM;
                }
            } else if (j == 0) {  // column 0 = the left-edge cells of rows 1..39 (:2486-2494)
                ncb = HIGH_NUM; neh = HIGH_NUM; neb = HIGH_NUM;
                nev = evu + r.vx;
            }
        }
    }

    // One double step.  Returns the packed direction bytes of the even step in lo and of the odd step in hi.
    template <bool BOUNDARY>
    __device__ __forceinline__ void double_step(int i0, int j0, unsigned long long &dir_even, unsigned long long &dir_odd) {
        // ---- even step: q = 2m, cell (i0 - m, j0 + m)
        int in_eh = __shfl_up_sync(gmask, eh[Q - 1], 1, G);
        int in_cb = __shfl_up_sync(gmask, cb[Q - 1], 1, G);
        if (lane == 0) { in_eh = HIGH_NUM; in_cb = HIGH_NUM; }  // the left-edge cells (:2487, :2494)
        unsigned long long de = 0, dod = 0;
#pragma unroll
        for (int m = 0; m < K; m++) {
            const int q = 2 * m;
            const int ehl = (m == 0) ? in_eh : eh[q - 1], cbl = (m == 0) ? in_cb : cb[q - 1];
            const int evu = ev[q + 1], cbu = cb[q + 1];
            const int dcost = lut[R[m].lut + C[m].lut];
            int ncb, nev, neh, neb;
            const int byte = aff_cell_dna<BT>(ehl, cbl, evu, cbu, cb[q], ev[q], eh[q], eb[q], R[m], C[m], dcost, go, ncb, nev,
                                              neh, neb);
            fixups<BOUNDARY>(q, i0 - m, j0 + m, ehl, evu, R[m], C[m], ncb, nev, neh, neb);
            cb[q] = ncb; ev[q] = nev; eh[q] = neh; eb[q] = neb;
            if (BT) de |= (unsigned long long) byte << (8 * m);
        }
        // ---- odd step: q = 2m + 1, cell (i0 - m, j0 + m + 1)
        int in_ev = __shfl_down_sync(gmask, ev[0], 1, G);
        int in_cbu = __shfl_down_sync(gmask, cb[0], 1, G);
#pragma unroll
        for (int m = 0; m < K; m++) {
            const int q = 2 * m + 1;
            const int ehl = eh[q - 1], cbl = cb[q - 1];
            const int evu = (m == K - 1) ? in_ev : ev[q + 1], cbu = (m == K - 1) ? in_cbu : cb[q + 1];
            const int dcost = lut[R[m].lut + C[m + 1].lut];
            int ncb, nev, neh, neb;
            const int byte = aff_cell_dna<BT>(ehl, cbl, evu, cbu, cb[q], ev[q], eh[q], eb[q], R[m], C[m + 1], dcost, go, ncb,
                                              nev, neh, neb);
            if (m == K - 1) {
                if (lane == G - 1) {  // diagonal dhi + 1: poisoned (:2531-2535)
                    ncb = HIGH_NUM; nev = HIGH_NUM; neh = HIGH_NUM; neb = HIGH_NUM;
                }
            }
            fixups<BOUNDARY>(q, i0 - m, j0 + m + 1, ehl, evu, R[m], C[m + 1], ncb, nev, neh, neb);
            cb[q] = ncb; ev[q] = nev; eh[q] = neh; eb[q] = neb;
            if (BT) dod |= (unsigned long long) byte << (8 * m);
        }
        dir_even = de;
        dir_odd = dod;
    }
};

template <int BL>
__device__ __forceinline__ void store_dir(uint8_t *p, unsigned long long v) {
    if (BL == 4) *reinterpret_cast<uint32_t *>(p) = (uint32_t) v;
    else *reinterpret_cast<unsigned long long *>(p) = v;
}

// seq_bytes: shared-memory bytes reserved per operand (multiple of 16, >= the longest sequence of the launch).
template <int K, int G, bool BT>
__global__ void __launch_bounds__(STRIPE_WARPS * 32) aff_stripe_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm,
                                                                       const uint8_t *__restrict__ pool,
                                                                       uint8_t *__restrict__ dir, int *__restrict__ out_cost,
                                                                       int seq_bytes) {
    constexpr int GPW = 32 / G;  // groups (pairs) per warp
    constexpr int Q = 2 * K;
    constexpr int BL = (K <= 4) ? 4 : 8;
    extern __shared__ __align__(16) uint8_t smem[];
    int *s_lut = reinterpret_cast<int *>(smem);  // 256 ints: cost[(a & 15) << lcm | (b & 15)]
    int *s_prep = s_lut + 256;                    // 32 ints
    int *s_get = s_prep + 32;                     // 32 ints: cost[c << lcm | gap]
    uint8_t *s_seq = reinterpret_cast<uint8_t *>(s_get + 32);
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_lut[k] = __ldg(cm.cost + ((k >> 4) << cm.lcm) + (k & 15));
    for (int k = threadIdx.x; k < 32; k += blockDim.x) {
        s_prep[k] = __ldg(cm.prepend + k);
        s_get[k] = __ldg(cm.cost + (k << cm.lcm) + cm.gap);
    }
    __syncthreads();

    const int warp_in_block = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
    const int grp = lane32 / G, lane = lane32 % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
    uint8_t *my_seq = s_seq + (size_t) ((warp_in_block * GPW + grp) * 2) * seq_bytes;
    const int warp_global = blockIdx.x * STRIPE_WARPS + warp_in_block;
    const int total_warps = gridDim.x * STRIPE_WARPS;

    for (int batch = warp_global; batch * GPW < ntasks; batch += total_warps) {
        const int ti = batch * GPW + grp;
        const bool valid = ti < ntasks;
        Task t;
        if (valid) t = tasks[ti];
        else { t = Task{}; t.lr = 1; t.lc = 1; t.dhi = -1; t.dlo = -39; }
        const int nr = t.lr - 1, nc = t.lc - 1;
        // stage both operands in shared memory
        {
            const uint8_t *gr = pool + t.off_r, *gc = pool + t.off_c;
            if (valid) {
                if ((((uintptr_t) gr | (uintptr_t) gc) & 15) == 0) {
                    for (int k = lane * 16; k < t.lr; k += G * 16)
                        *reinterpret_cast<uint4 *>(my_seq + k) = __ldg(reinterpret_cast<const uint4 *>(gr + k));
                    for (int k = lane * 16; k < t.lc; k += G * 16)
                        *reinterpret_cast<uint4 *>(my_seq + seq_bytes + k) = __ldg(reinterpret_cast<const uint4 *>(gc + k));
                } else {
                    for (int k = lane; k < t.lr; k += G) my_seq[k] = __ldg(gr + k);
                    for (int k = lane; k < t.lc; k += G) my_seq[seq_bytes + k] = __ldg(gc + k);
                }
            } else if (lane == 0) {
                my_seq[0] = 16;
                my_seq[seq_bytes] = 16;
            }
        }
        __syncwarp();

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

        auto run = [&](auto lowtag) {
            constexpr bool LOW = decltype(lowtag)::value;
            AffStripe<K, G, BT, LOW> S;
            S.si = my_seq; S.sj = my_seq + seq_bytes;
            S.lut = s_lut; S.prep = s_prep; S.get = s_get;
            S.nr = nr; S.nc = nc; S.go = cm.gap_open; S.lane = lane; S.gmask = gmask;
            S.qlow = min(max(qlow_all - lane * Q, 0), Q);
#pragma unroll
            for (int q = 0; q < Q; q++) { S.cb[q] = HIGH_NUM; S.ev[q] = HIGH_NUM; S.eh[q] = HIGH_NUM; S.eb[q] = HIGH_NUM; }
            int u = u_begin;
            int i0 = u - lane * K, j0 = u + d0 + lane * K;
            S.init_windows(i0, j0);
            int u_b = max(G * K, 1 - d0);  // from here on every lane has i >= 1 and j >= 1 (warp-uniform maximum)
#pragma unroll
            for (int o = G; o < 32; o <<= 1) u_b = max(u_b, __shfl_xor_sync(0xffffffffu, u_b, o));
            auto emit = [&](unsigned long long de, unsigned long long dod) {
                if (BT) {
                    const int te = 2 * u + d0;
                    if (u >= u_first && u <= u_last) {
                        if (te >= 0) store_dir<BL>(dbase + ((size_t) te * G + lane) * BL, de);
                        if (te + 1 <= nr + nc) store_dir<BL>(dbase + ((size_t) (te + 1) * G + lane) * BL, dod);
                    }
                }
                if (u == u_last && lane == lane_f) {
                    int r = 0;
#pragma unroll
                    for (int q = 0; q < Q; q++)
                        if (q == q_f) r = min(min(S.cb[q], S.ev[q]), min(S.eh[q], S.eb[q]));
                    result = r;
                }
            };
            for (; u < min(u_b, u_end + 1); u++) {
                unsigned long long de, dod;
                S.template double_step<true>(i0, j0, de, dod);
                emit(de, dod);
                i0++; j0++;
                S.slide_windows(i0, j0);
            }
            for (; u <= u_end; u++) {
                unsigned long long de, dod;
                S.template double_step<false>(i0, j0, de, dod);
                emit(de, dod);
                i0++; j0++;
                S.slide_windows(i0, j0);
            }
        };
        // LOW must be uniform over the warp's control flow: take the slower variant if any group needs it
        const bool any_low = __any_sync(0xffffffffu, low);
        if (any_low) run(std::true_type{});
        else run(std::false_type{});

        if (valid && lane == lane_f) {
            if (BT && nr == 0 && nc == 0) result = 0;
            out_cost[t.pair] = result;
        }
        __syncwarp();
    }
}

// ---- host side -----------------------------------------------------------------------------------------

// Chooses a stripe shape for an affine pair; returns false when the pair must take the generic kernel.
static inline bool stripe_choose(Task &t, bool affine, int W, const DevCM &cm) {
    if (!affine) return false;
    if (cm.lcm != 5 || cm.gap != 16) return false;  // aff_cell_dna assumes the nucleotide encoding
    if (t.lr > STRIPE_MAX_SEQ_BYTES || t.lc > STRIPE_MAX_SEQ_BYTES) return false;
    for (int s = 0; s < N_AFF_SHAPES; s++) {
        const int K = AFF_SHAPES[s].K, G = AFF_SHAPES[s].G;
        if (2 * K * G >= W + 1) {
            t.klass = 1 + s;
            t.G = G;
            t.twoK = 2 * K;
            t.BL = (K <= 4) ? 4 : 8;
            t.dbase = t.dhi + 2 - 2 * K * G;
            return true;
        }
    }
    return false;
}

template <int K, int G>
static cudaError_t stripe_launch_shape(bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir,
                                       int *cost, int sm_count, int seq_bytes, cudaStream_t stream) {
    constexpr int GPW = 32 / G;
    const size_t smem = (256 + 64) * sizeof(int) + (size_t) STRIPE_WARPS * GPW * 2 * seq_bytes;
    const int nbatches = (n + GPW - 1) / GPW;
    auto kern = bt ? aff_stripe_kernel<K, G, true> : aff_stripe_kernel<K, G, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, STRIPE_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int blocks = std::min((nbatches + STRIPE_WARPS - 1) / STRIPE_WARPS, sm_count * per_sm);
    if (blocks < 1) blocks = 1;
    kern<<<blocks, STRIPE_WARPS * 32, smem, stream>>>(d_tasks, n, cm, pool, dir, cost, seq_bytes);
    return cudaGetLastError();
}

static inline cudaError_t stripe_launch(uint32_t klass, bool affine, bool bt, const Task *d_tasks, int n, DevCM cm,
                                        const uint8_t *pool, uint8_t *dir, int *cost, int sm_count, int seq_bytes,
                                        cudaStream_t stream) {
    if (!affine) return cudaErrorNotSupported;
    switch (klass - 1) {
        case 0: return stripe_launch_shape<5, 8>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, stream);
        case 1: return stripe_launch_shape<6, 8>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, stream);
        case 2: return stripe_launch_shape<4, 16>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, stream);
        case 3: return stripe_launch_shape<6, 16>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, stream);
        case 4: return stripe_launch_shape<4, 32>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, stream);
        case 5: return stripe_launch_shape<6, 32>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, stream);
        case 6: return stripe_launch_shape<8, 32>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace poyb200
