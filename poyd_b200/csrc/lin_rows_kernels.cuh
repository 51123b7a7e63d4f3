// Column-striped kernel for the FULL linear-gap matrix: algn_fill_plane (src/algn.c:562-700, the no-band twin that
// algn_fill_plane_2 falls back to in its cases 1 and 3a, :893, :936) -- every cell of the (rows x cols) matrix.
//
// The diagonal stripes of lin_stripe_kernels.cuh sweep a parallelogram; wrapped around a full square matrix, more than
// half of its cells lie outside.  Here lane l of a group of G lanes owns the C columns j = l C .. l C + C - 1 for the
// whole pair and the group sweeps ROWS with a skew of one row per lane: at step t lane l computes row i = t - l.  A cell
// needs its left and diagonal neighbours from the lane to the left (one shuffle per step: the last column of the row that
// lane finished a step earlier, kept for one more step as the diagonal neighbour), everything else is the lane's own
// previous row in registers.  Waste is the skew (G - 1 steps) and the columns past the operand, not half of the sweep.
//
// Per column the lane keeps two constants for the whole pair (offset of b_j in the cost LUT, 4 cost(gap, b_j) + tag), so
// a cell is one shared-memory load, three adds, one three-way min, and the tag extraction.  Values x4 with the move in the
// low two bits exactly as in lin_stripe_kernels.cuh (0 ALIGN, then INSERT / DELETE in the order backtrack_2d :3606-3665
// tries them for this pair); a row of a lane is 2 C <= 32 direction bits, stored as one word in row-major tiles of 8 rows
// (TF_ROWMAJ, common.cuh).
#pragma once
#include "stripe_kernels.cuh"

namespace poyb200 {

constexpr int LIN_ROWS_MIN_BLOCKS = 4;

// Shared memory layout: the same tables as lin_stripe_kernel (LUT of 4 cost[a][b], gaprow, gapcol, prep4, tail4).
template <int C, int G, bool BT>
__global__ void __launch_bounds__(STRIPE_WARPS * 32, LIN_ROWS_MIN_BLOCKS) lin_rows_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm,
                                                                                     const uint8_t *__restrict__ pool,
                                                                                     uint8_t *__restrict__ dir, int *__restrict__ out_cost,
                                                                                     int seq_bytes, int nslots, int custom_tail,
                                                                                     int *work_counter) {
    static_assert(2 * C <= 32, "a lane's row of moves must fit one word");
    constexpr int GPW = 32 / G;
    extern __shared__ __align__(16) uint8_t smem[];
    const int dim = 1 << cm.lcm, row_ints = dim + 1;
    int *s_lut = reinterpret_cast<int *>(smem);
    int *s_gaprow = s_lut + dim * row_ints, *s_gapcol = s_gaprow + dim, *s_prep = s_gapcol + dim, *s_tail = s_prep + dim;
    StageBars *s_bar = reinterpret_cast<StageBars *>(s_tail + dim + (((dim * (row_ints + 4)) & 1) ? 1 : 0));  // 8-byte aligned
    uint8_t *s_seq = reinterpret_cast<uint8_t *>(s_bar + STRIPE_WARPS * 4);
    s_seq += (16 - ((uintptr_t) s_seq & 15)) & 15;
    if (threadIdx.x < STRIPE_WARPS * GPW) StageRing<G>::init_bars(&s_bar[threadIdx.x]);
    for (int k = threadIdx.x; k < dim * dim; k += blockDim.x)
        s_lut[(k >> cm.lcm) * row_ints + (k & (dim - 1))] = 4 * __ldg(cm.cost + k);
    for (int k = threadIdx.x; k < dim; k += blockDim.x) {
        s_gaprow[k] = 4 * __ldg(cm.cost + (k << cm.lcm) + cm.gap);
        s_gapcol[k] = 4 * __ldg(cm.cost + (cm.gap << cm.lcm) + k);
        s_prep[k] = 4 * __ldg(cm.prepend + k);
        s_tail[k] = 4 * __ldg(cm.tail + k);
    }
    __syncthreads();

    const int warp_in_block = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
    const int grp = lane32 / G, lane = lane32 % G;
    StageRing<G> ring;
    ring.attach(&s_bar[warp_in_block * GPW + grp], s_seq + (size_t) ((warp_in_block * GPW + grp) * 2 * nslots) * seq_bytes, seq_bytes,
                nslots, lane);
    const int nbatches = (ntasks + GPW - 1) / GPW;
    const uint8_t *lut = reinterpret_cast<const uint8_t *>(s_lut);
    const int lut_row_bytes = row_ints * 4;

    int slot = 0;
    int batch = fetch_batch(work_counter, nbatches, nullptr, nullptr);
    if (batch >= 0) ring.produce_task(0, tasks, ntasks, batch * GPW + grp, pool, cm.gap);
    while (batch >= 0) {
        int next = -1;
        if (nslots == 2) {  // the operands of the next batch travel under this one
            next = fetch_batch(work_counter, nbatches, nullptr, nullptr);
            if (next >= 0) ring.produce_task(slot ^ 1, tasks, ntasks, next * GPW + grp, pool, cm.gap);
        }
        const int ti = batch * GPW + grp;
        const bool valid = ti < ntasks;
        Task t;
        if (valid) t = tasks[ti];
        else { t = Task{}; t.lr = 1; t.lc = 1; }
        const int nr = t.lr - 1, nc = t.lc - 1;
        ring.wait_full(slot);
        const uint8_t *s1 = ring.rows(slot), *s2 = s1 + seq_bytes;

        const bool swaped = (t.flags & TF_SWAPED) != 0;
        const int ins_tag = swaped ? 1 : 2, del_tag = swaped ? 2 : 1;
        const int j_first = lane * C;
        int colp[C], cins[C], V[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int b = s2[min(j_first + c, nc)];
            colp[c] = b * 4;
            cins[c] = s_gapcol[b] + ins_tag;
            V[c] = LIN_INF;
        }
        int left_prev = LIN_INF;
        // default prepend / tail costs (bit 1 of custom_tail): row 0 and column 0 are what the ordinary cell computes from
        // "no neighbour" inputs, given a diagonal seed that makes cell (0, 0) come out as 0 (ALIGN): no boundary variant
        const bool natural = (custom_tail & 2) != 0;
        if (natural && lane == 0)
            left_prev = -*reinterpret_cast<const int *>(lut + s1[0] * lut_row_bytes + colp[0]);
        const int lane_f = nc / C, c_f = nc - lane_f * C;
        int t_end = valid ? nr + lane_f : -1;  // the lanes past column nc need not finish
#pragma unroll
        for (int o = G; o < 32; o <<= 1) t_end = max(t_end, __shfl_xor_sync(0xffffffffu, t_end, o));
        uint32_t *dwords = reinterpret_cast<uint32_t *>(dir + t.dir_off);
        const bool stores = BT && valid && j_first <= nc;
        int result = 0;

        // BND: some lane is on row 0 or has not started (t < G); ENDP: custom tail costs (algn_fill_last_column's extra
        // DELETE candidate on column nc, :548-560 -- with the default tail it equals the ordinary one)
        auto step = [&](int tt, auto bndtag, auto endtag) {
            constexpr bool BND = decltype(bndtag)::value, ENDP = decltype(endtag)::value;
            int left_in = __shfl_up_sync(0xffffffffu, V[C - 1], 1, G);
            if (lane == 0) left_in = LIN_INF;
            const int i = tt - lane;
            const int a = s1[min(max(i, 0), nr)];
            const uint8_t *rowp = lut + a * lut_row_bytes;
            const int cdel = s_gaprow[a] + del_tag;
            int ctail = 0;
            if (ENDP) ctail = s_tail[a] + del_tag;
            uint32_t dw = 0;
            int diag = left_prev, left = left_in;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const int up = V[c];
                const int c_al = *reinterpret_cast<const int *>(rowp + colp[c]);
                int v = min(diag + c_al, min(left + cins[c], up + cdel));  // :381-431, ties resolved by the tags
                if (ENDP) {
                    if (j_first + c == nc && nc > 0) v = min(v, up + ctail);  // `if (l > 0)` :551: not on a lone column 0
                }
                if (BND) {
                    if (i == 0) {
                        const int j = j_first + c;
                        if (j == 0) v = 0;                                           // ALIGN (:587-588)
                        else v = left + s_prep[s2[min(j, nc)]] + ins_tag;             // :597-598
                    } else if (i < 0) {
                        v = LIN_INF;
                    }
                }
                diag = up;
                left = v & ~3;
                V[c] = left;
                if (BT) dw |= (uint32_t) (v & 3) << (2 * c);
            }
            left_prev = left_in;
            if (stores && i >= 0 && i <= nr) dwords[(((size_t) (i >> 3) * G + lane) << 3) + (i & 7)] = dw;
            if (i == nr && lane == lane_f) {
                int r = 0;
#pragma unroll
                for (int c = 0; c < C; c++)
                    if (c == c_f) r = V[c] >> 2;
                result = r;
            }
        };
        int tt = 0;
        if (natural) {
            for (; tt + 1 <= t_end; tt += 2) {
                step(tt, std::false_type{}, std::false_type{});
                step(tt + 1, std::false_type{}, std::false_type{});
            }
            if (tt <= t_end) step(tt, std::false_type{}, std::false_type{});
        } else if (custom_tail & 1) {
            for (; tt < G && tt <= t_end; tt++) step(tt, std::true_type{}, std::true_type{});
            for (; tt <= t_end; tt++) step(tt, std::false_type{}, std::true_type{});
        } else {
            for (; tt < G && tt <= t_end; tt++) step(tt, std::true_type{}, std::false_type{});
            for (; tt + 1 <= t_end; tt += 2) {  // two rows per trip: the row registers alternate instead of being copied
                step(tt, std::false_type{}, std::false_type{});
                step(tt + 1, std::false_type{}, std::false_type{});
            }
            if (tt <= t_end) step(tt, std::false_type{}, std::false_type{});
        }

        if (valid && lane == lane_f) out_cost[t.pair] = result;
        ring.release(slot);  // this lane's last read of the staged operands is behind it
        if (nslots == 1) {
            next = fetch_batch(work_counter, nbatches, nullptr, nullptr);
            if (next >= 0) ring.produce_task(0, tasks, ntasks, next * GPW + grp, pool, cm.gap);
        } else {
            slot ^= 1;
        }
        batch = next;
    }
}

// ---- host side -----------------------------------------------------------------------------------------
#ifdef POYB200_DEFINE_LIN_ROWS  // the translation unit that owns these kernels (k_lin_rows.cu)
static inline size_t lin_rows_table_bytes(int lcm) {
    const size_t dim = (size_t) 1 << lcm;
    return (dim * (dim + 1) + 4 * dim + 1) * sizeof(int) + STRIPE_WARPS * 4 * STAGE_BAR_BYTES + 16;
}

template <int C, int G>
static cudaError_t lin_rows_launch_shape(bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir, int *cost,
                                         int sm_count, int seq_bytes, int custom_tail, int *work_counter, cudaStream_t stream) {
    constexpr int GPW = 32 / G;
    const int nbatches = (n + GPW - 1) / GPW;
    auto kern = bt ? lin_rows_kernel<C, G, true> : lin_rows_kernel<C, G, false>;
    size_t smem = 0;
    int nslots = 1, per_sm = 1;
    cudaError_t e = stage_ring_config(kern, lin_rows_table_bytes(cm.lcm), (size_t) STRIPE_WARPS * GPW * 2 * seq_bytes, STRIPE_WARPS * 32,
                                      smem, nslots, per_sm);
    if (e != cudaSuccess) return e;
    int blocks = std::min((nbatches + STRIPE_WARPS - 1) / STRIPE_WARPS, sm_count * per_sm);
    if (blocks < 1) blocks = 1;
    kern<<<blocks, STRIPE_WARPS * 32, smem, stream>>>(d_tasks, n, cm, pool, dir, cost, seq_bytes, nslots, custom_tail, work_counter);
    return cudaGetLastError();
}

cudaError_t lin_rows_launch(uint32_t klass, bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir, int *cost,
                            int sm_count, int seq_bytes, int custom_tail, int *work_counter, cudaStream_t stream) {
#define ROWS_CASE(IDX, CC, GG) \
    case IDX: return lin_rows_launch_shape<CC, GG>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, custom_tail, work_counter, stream)
    static_assert(N_LIN_ROW_SHAPES == 9, "keep the switch in step with LIN_ROW_SHAPES");
    switch (klass - KLASS_LINROW_BASE) {
        ROWS_CASE(0, 8, 8);
        ROWS_CASE(1, 12, 8);
        ROWS_CASE(2, 16, 8);
        ROWS_CASE(3, 10, 16);
        ROWS_CASE(4, 12, 16);
        ROWS_CASE(5, 16, 16);
        ROWS_CASE(6, 10, 32);
        ROWS_CASE(7, 12, 32);
        ROWS_CASE(8, 16, 32);
        default: return cudaErrorInvalidValue;
    }
#undef ROWS_CASE
}
#endif  // POYB200_DEFINE_LIN_ROWS

}  // namespace poyb200
