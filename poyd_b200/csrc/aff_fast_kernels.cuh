// Fast path of the affine stripe sweep (src/algn.c:2411-2548): pairs whose operands carry no gap bit beyond the leading
// element, gap_open > 0, stripe flush with dlo (no spare diagonals).  Same stripe formulation, same direction bytes and
// same costs as aff_stripe_kernel (stripe_kernels.cuh) -- that kernel's NOEB variant is the specification -- but built
// for instruction count:
//
//   * the block-diagonal state is gone (see the NOEB note in stripe_kernels.cuh) and with it every per-row / per-column
//     gap-opening special case: a row needs {4 * cost[si][gap], address of its LUT row}, a column
//     {4 * prepend[sj], byte offset inside a LUT row}, both fetched from 256-entry shared tables with one 8-byte load;
//   * the row / column windows do not slide.  They are rings of 8 slots and the sweep is unrolled in blocks of 8
//     double steps, so every slot index is a compile-time constant and no register is ever copied;
//   * a block starts where the double-step counter is a multiple of 4, which by the choice of Task::tshift is a
//     tile boundary of the direction band for every pair of the warp: the 16 stores of a block go to compile-time
//     offsets from one running pointer;
//   * the middle of the sweep (every lane inside the matrix, every pair of the warp still running) runs blocks without
//     any position test; the first and last few blocks run the same code with the tests switched on (EDGE).
//
// Batches that do not qualify are appended to a list and taken by aff_stripe_kernel right after, on the same stream.
#pragma once
#include "stripe_kernels.cuh"

namespace poyb200 {

constexpr int FAST_P = 8;                                   // ring slots = double steps per block
constexpr int FAST_LUT_ROW = 17 * 4;                        // 16 ints + 1 pad
constexpr int FAST_TABLE_BYTES = 2 * 256 * 8 + 16 * FAST_LUT_ROW + STRIPE_WARPS * 4 * 8;
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
// staged operands change from pair to pair: volatile, so the load is neither hoisted nor merged across pairs
__device__ __forceinline__ int lds_u8_seq(uint32_t a) {
    int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

template <int K, int G, bool BT>
struct AffFast {
    static constexpr int Q = 2 * K, P = FAST_P, BL = (K <= 4) ? 4 : 8;
    static constexpr int D = (K <= 6) ? 2 : 1;  // windows are loaded D double steps ahead
    static_assert(K + 1 + D <= P + 1 && K + D <= P, "ring too small");

    int cb[Q], ev[Q], eh[Q];
    int Rv[P], Cv[P];       // 4 * cost[si][gap], 4 * prepend[sj]
    uint32_t Rl[P], Cl[P];  // shared address of the LUT row, byte offset of the column
    uint32_t si, sj, tabR, tabC;  // shared addresses
    int nr, nc, go4, lane, keep;
    // per pair
    int d0, sbase, u_first, u_last, lane_f, q_f, result;
    uint8_t *dbase;

    template <bool EDGE>
    __device__ __forceinline__ void load_row(int slot, int i) {
        if (EDGE) i = min(max(i, 0), nr);
        const int2 e = lds_v2(tabR + 8 * lds_u8_seq(si + i));
        Rv[slot] = e.x;
        Rl[slot] = (uint32_t) e.y;
    }
    template <bool EDGE>
    __device__ __forceinline__ void load_col(int slot, int j) {
        if (EDGE) j = min(max(j, 0), nc);
        const int2 e = lds_v2(tabC + 8 * lds_u8_seq(sj + j));
        Cv[slot] = e.x;
        Cl[slot] = (uint32_t) e.y;
    }
    // Windows as block entry expects them: rows i0-K+1 .. i0+D-1 in slots (r & 7), columns j0 .. j0+K+D-1 in slots n & 7.
    __device__ __forceinline__ void init_windows(int i0, int j0) {
#pragma unroll
        for (int r = -K + 1; r <= D - 1; r++) load_row<true>(r & (P - 1), i0 + r);
#pragma unroll
        for (int n = 0; n <= K + D - 1; n++) load_col<true>(n & (P - 1), j0 + n);
    }

    template <bool EDGE>
    __device__ __forceinline__ int cell(int i, int j, int ehl, int cbl, int evu, int cbu, int q, int rs, int cs) {
        const int c_h = go4 - 1, c_v = go4 + 1;  // CB tag 1 -> EH tag 0 / EV tag 2
        const int t = cbl + c_h, t2 = cbu + c_v;
        int neh = min(ehl, t) + Cv[cs];           // FILL_EXTEND_HORIZONTAL :1765-1787
        int nev = min(evu, t2) + Rv[rs];          // FILL_EXTEND_VERTICAL :1813-1830
        const int d = lds_s32(Rl[rs] + Cl[cs]);   // 4 * cost[si & 15][sj & 15]
        const int ck = min(min(cb[q] + 2, ev[q]), eh[q]) + d;  // FILL_CLOSE_BLOCK_DIAGONAL :1923-1977 (tags A 3, V 2, H 0)
        int ncb = (ck & keep) | TAG_CB;
        int byte = 0;
        if (BT) {
            byte = (ehl < t) ? AB_ENDB : (AB_ENDB | AB_ENDH);
            byte += (evu < t2) ? 0 : AB_ENDV;
            const int fk = min(min(neh, nev), ncb);  // ASSIGN_MINIMUM :2251-2280
            byte += (fk & 3) * 4 + (ck & 3);
        }
        if (EDGE) {
            if (i == 0) {
                if (j == 0) {  // :2194-2198
                    ncb = TAG_CB; neh = go4 + TAG_EH; nev = go4 + TAG_EV;
                } else {       // :2212-2217
                    const int rr = ehl + Cv[cs];
                    neh = rr; ncb = rr + TAG_CB; nev = HIGH4 + TAG_EV;
                }
            } else if (j == 0) {  // the left-edge cells of rows 1..39 (:2486-2494)
                ncb = HIGH4 + TAG_CB; neh = HIGH4 + TAG_EH;
                nev = evu + Rv[rs];
            }
        }
        if (EDGE) {
            // the final cell: the cost of the alignment (:2540-2547).  Unchecked blocks end before any pair's last step.
            if (i == nr && j == nc) result = min(ncb, min(nev, neh)) >> 2;
        }
        cb[q] = ncb; ev[q] = nev; eh[q] = neh;
        return byte;
    }

    // One block of P double steps starting at double step u (u % 4 == 0 when !EDGE), lane origin (i0, j0).
    // !EDGE: dptr = address of this lane's chunk for the block's first step.
    template <bool EDGE>
    __device__ __forceinline__ void block(int u, int i0, int j0, uint8_t *dptr) {
#pragma unroll
        for (int p = 0; p < P; p++) {
            uint32_t de[2] = {0, 0}, dod[2] = {0, 0};
            // ---- even step: q = 2m, cell (i0 + p - m, j0 + p + m)
            int in_eh = __shfl_up_sync(0xffffffffu, eh[Q - 1], 1, G);
            int in_cb = __shfl_up_sync(0xffffffffu, cb[Q - 1], 1, G);
            if (lane == 0) { in_eh = HIGH4 + TAG_EH; in_cb = HIGH4 + TAG_CB; }  // the left-edge cells (:2487, :2494)
#pragma unroll
            for (int m = 0; m < K; m++) {
                const int q = 2 * m;
                const int ehl = (m == 0) ? in_eh : eh[q - 1], cbl = (m == 0) ? in_cb : cb[q - 1];
                const int byte = cell<EDGE>(i0 + p - m, j0 + p + m, ehl, cbl, ev[q + 1], cb[q + 1], q, (p - m) & (P - 1),
                                            (p + m) & (P - 1));
                if (BT) de[m >> 2] |= (uint32_t) byte << (8 * (m & 3));
            }
            // ---- odd step: q = 2m + 1, cell (i0 + p - m, j0 + p + m + 1)
            const int in_ev = __shfl_down_sync(0xffffffffu, ev[0], 1, G);
            const int in_cbu = __shfl_down_sync(0xffffffffu, cb[0], 1, G);
#pragma unroll
            for (int m = 0; m < K; m++) {
                const int q = 2 * m + 1;
                const int evu = (m == K - 1) ? in_ev : ev[q + 1], cbu = (m == K - 1) ? in_cbu : cb[q + 1];
                const int byte = cell<EDGE>(i0 + p - m, j0 + p + m + 1, eh[q - 1], cb[q - 1], evu, cbu, q, (p - m) & (P - 1),
                                            (p + m + 1) & (P - 1));
                if (m == K - 1) {
                    if (lane == G - 1) {  // diagonal dhi + 1: poisoned (:2531-2535)
                        cb[q] = HIGH4 + TAG_CB; ev[q] = HIGH4 + TAG_EV; eh[q] = HIGH4 + TAG_EH;
                    }
                }
                if (BT) dod[m >> 2] |= (uint32_t) byte << (8 * (m & 3));
            }
            // ---- direction bytes
            if (!EDGE) {
                if (BT) {
                    constexpr int TILE = G * 8 * BL;
                    store_dir<BL>(dptr + ((2 * p) >> 3) * TILE + ((2 * p) & 7) * BL, de);
                    store_dir<BL>(dptr + ((2 * p + 1) >> 3) * TILE + ((2 * p + 1) & 7) * BL, dod);
                }
            } else {
                const int uu = u + p;
                if (BT && uu >= u_first && uu <= u_last) {
                    const int te = 2 * uu + d0, s = 2 * uu - sbase;  // s = te - tshift
                    if (te >= 0) store_dir<BL>(dbase + (((size_t) (s >> 3) * G + lane) * 8 + (s & 7)) * BL, de);
                    if (te + 1 <= nr + nc) store_dir<BL>(dbase + (((size_t) ((s + 1) >> 3) * G + lane) * 8 + ((s + 1) & 7)) * BL, dod);
                }
                (void) uu;
            }
            // ---- windows: row i0 + p + D and column j0 + p + K + D enter
            load_row<EDGE>((p + D) & (P - 1), i0 + p + D);
            load_col<EDGE>((p + K + D) & (P - 1), j0 + p + K + D);
        }
    }
};

// seq_bytes as for aff_stripe_kernel.  slow_list / slow_count receive the batches this kernel declines.
template <int K, int G, bool BT>
__global__ void __launch_bounds__(STRIPE_WARPS * 32, FAST_MIN_BLOCKS)
    aff_fast_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm, const uint8_t *__restrict__ pool, uint8_t *__restrict__ dir,
                    int *__restrict__ out_cost, int seq_bytes, int *work_counter, int *slow_list, int *slow_count, int keep_mask) {
    constexpr int GPW = 32 / G;
    constexpr int Q = 2 * K;
    using S_t = AffFast<K, G, BT>;
    constexpr int BL = S_t::BL;
    extern __shared__ __align__(16) uint8_t smem[];
    int2 *s_tabR = reinterpret_cast<int2 *>(smem);
    int2 *s_tabC = s_tabR + 256;
    int *s_lut = reinterpret_cast<int *>(s_tabC + 256);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + 2 * 256 * 8 + 16 * FAST_LUT_ROW);
    uint8_t *s_seq = reinterpret_cast<uint8_t *>(s_bar + STRIPE_WARPS * 4);
    if (threadIdx.x < STRIPE_WARPS * 4) mbar_init(&s_bar[threadIdx.x], 1);
    uint32_t bar_phase = 0;
    for (int k = threadIdx.x; k < 256; k += blockDim.x) {
        s_lut[(k >> 4) * 17 + (k & 15)] = 4 * __ldg(cm.cost + ((k >> 4) << cm.lcm) + (k & 15));
        s_tabR[k] = make_int2(4 * __ldg(cm.cost + ((k & 31) << cm.lcm) + cm.gap), (int) (smem_u32(s_lut) + (k & 15) * FAST_LUT_ROW));
        s_tabC[k] = make_int2(4 * __ldg(cm.prepend + (k & 31)), (k & 15) * 4);
    }
    __syncthreads();

    const int warp_in_block = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
    const int grp = lane32 / G, lane = lane32 % G;
    const int op_stride = seq_bytes + fast_operand_pad(K, G);
    uint8_t *my_seq = s_seq + (size_t) ((warp_in_block * GPW + grp) * 2) * op_stride;

    for (;;) {
        int batch = 0;
        if (lane32 == 0) batch = atomicAdd(work_counter, 1);
        batch = __shfl_sync(0xffffffffu, batch, 0);
        if (batch * GPW >= ntasks) break;
        const int ti = batch * GPW + grp;
        const bool valid = ti < ntasks;
        Task t;
        if (valid) t = tasks[ti];
        else { t = Task{}; t.lr = 1; t.lc = 1; t.dhi = -1; t.dlo = -39; }
        const int nr = t.lr - 1, nc = t.lc - 1;
        const int d0 = t.dhi + 2 - Q * G;
        // spare diagonals below dlo need the left-edge rule inside the stripe: not here
        const bool low = valid && (t.dlo - d0 > 0);
        if (__any_sync(0xffffffffu, low)) {
            if (lane32 == 0) slow_list[atomicAdd(slow_count, 1)] = batch;
            continue;
        }
        __syncwarp();
        stage_pair<G>(my_seq, my_seq + op_stride, pool + t.off_r, pool + t.off_c, t.lr, t.lc, lane, valid,
                      &s_bar[warp_in_block * GPW + grp], bar_phase, 16);
        __syncwarp();
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
            continue;
        }

        S_t S;
        S.si = smem_u32(my_seq); S.sj = smem_u32(my_seq + op_stride);
        S.tabR = smem_u32(s_tabR); S.tabC = smem_u32(s_tabC);
        S.nr = nr; S.nc = nc; S.go4 = 4 * cm.gap_open; S.lane = lane; S.keep = keep_mask;
        S.d0 = d0;
        S.u_first = (-d0) >> 1;  // first double step: t = 2u + d0 in {-1, 0}
        S.u_last = valid ? ((nr + nc - d0) >> 1) : (S.u_first - 1);
        S.sbase = (2 * S.u_first) & ~7;
        S.dbase = dir + t.dir_off;
        const int dd_f = (nc - nr) - d0;
        S.lane_f = dd_f / Q; S.q_f = dd_f % Q;
        S.result = 0;
        int u_end = S.u_last, u_safe = S.u_last, u_begin = S.u_first;
        int u_b = max(G * K, 1 - d0);  // from here on every lane has i >= 1 and j >= 1
#pragma unroll
        for (int o = G; o < 32; o <<= 1) {
            u_end = max(u_end, __shfl_xor_sync(0xffffffffu, u_end, o));
            u_safe = min(u_safe, __shfl_xor_sync(0xffffffffu, u_safe, o));
            u_begin = min(u_begin, __shfl_xor_sync(0xffffffffu, u_begin, o));
            u_b = max(u_b, __shfl_xor_sync(0xffffffffu, u_b, o));
        }
#pragma unroll
        for (int q = 0; q < Q; q++) { S.cb[q] = HIGH4 + TAG_CB; S.ev[q] = HIGH4 + TAG_EV; S.eh[q] = HIGH4 + TAG_EH; }
        int u = u_begin & ~3;
        int i0 = u - lane * K, j0 = u + d0 + lane * K;
        S.init_windows(i0, j0);
        for (; u < u_b && u <= u_end; u += FAST_P, i0 += FAST_P, j0 += FAST_P) S.template block<true>(u, i0, j0, nullptr);
        if (u + FAST_P <= u_safe) {
            const int s0 = 2 * u - S.sbase;
            uint8_t *dptr = S.dbase + ((size_t) (s0 >> 3) * G + lane) * 8 * BL;
            for (; u + FAST_P <= u_safe; u += FAST_P, i0 += FAST_P, j0 += FAST_P, dptr += 2 * G * 8 * BL)
                S.template block<false>(u, i0, j0, dptr);
        }
        for (; u <= u_end; u += FAST_P, i0 += FAST_P, j0 += FAST_P) S.template block<true>(u, i0, j0, nullptr);

        if (valid && lane == S.lane_f) {
            if (BT && nr == 0 && nc == 0) S.result = 0;
            out_cost[t.pair] = S.result;
        }
        __syncwarp();
    }
}

#ifndef POYB200_KERNELS_ONLY
template <int K, int G>
static cudaError_t fast_launch_shape(bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir, int *cost,
                                     int sm_count, int seq_bytes, int *work_counter, int *slow_list, int *slow_count,
                                     cudaStream_t stream) {
    constexpr int GPW = 32 / G;
    const size_t smem = FAST_TABLE_BYTES + (size_t) STRIPE_WARPS * GPW * 2 * (seq_bytes + fast_operand_pad(K, G));
    const int nbatches = (n + GPW - 1) / GPW;
    auto kern = bt ? aff_fast_kernel<K, G, true> : aff_fast_kernel<K, G, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, STRIPE_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int blocks = std::min((nbatches + STRIPE_WARPS - 1) / STRIPE_WARPS, sm_count * per_sm);
    if (blocks < 1) blocks = 1;
    kern<<<blocks, STRIPE_WARPS * 32, smem, stream>>>(d_tasks, n, cm, pool, dir, cost, seq_bytes, work_counter, slow_list, slow_count, ~3);
    return cudaGetLastError();
}

// True when the class has a fast kernel (ring of 8 slots: K <= 6).
static inline bool fast_has_shape(uint32_t klass) {
    const int s = (int) klass - 1;
    return s >= 0 && s < N_AFF_SHAPES && AFF_SHAPES[s].K <= 6;
}

static inline cudaError_t fast_launch(uint32_t klass, bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir,
                                      int *cost, int sm_count, int seq_bytes, int *work_counter, int *slow_list, int *slow_count,
                                      cudaStream_t stream) {
    switch (klass - 1) {
        case 0: return fast_launch_shape<5, 8>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, slow_list, slow_count, stream);
        case 1: return fast_launch_shape<6, 8>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, slow_list, slow_count, stream);
        case 2: return fast_launch_shape<4, 16>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, slow_list, slow_count, stream);
        case 3: return fast_launch_shape<6, 16>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, slow_list, slow_count, stream);
        case 4: return fast_launch_shape<4, 32>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, slow_list, slow_count, stream);
        case 5: return fast_launch_shape<6, 32>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, work_counter, slow_list, slow_count, stream);
        default: return cudaErrorInvalidValue;
    }
}

#endif  // POYB200_KERNELS_ONLY

}  // namespace poyb200
