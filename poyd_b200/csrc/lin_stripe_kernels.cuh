// Register-resident stripe kernels for the linear-gap fill: algn_fill_plane_2 / algn_fill_plane
// (src/algn.c:375-968), same sweep as the affine stripe kernels (stripe_kernels.cuh) with one int of state per
// diagonal.
//
// Edge rules as missing neighbours: a cell on dhi has no upper neighbour (algn_fill_ukk_right_cell :461-479), a cell
// on dlo no left neighbour (algn_fill_ukk_left_cell :507-524) -- both are LIN_INF here, which also disables the
// last-column candidate on the right edge.  dhi sits on the last diagonal of the last lane (d0 = dhi + 1 - 2KG); the
// diagonals a shape leaves over below dlo are pinned to LIN_INF (LOW variant).
//
// Tagged keys: values are carried x4; the three candidates of a cell get the tags 0 (ALIGN), 1 and 2 in the order
// backtrack_2d (:3606-3665) would try them for this pair -- INSERT before DELETE when `swaped`, DELETE before INSERT
// otherwise -- so min3 yields the traceback's move in the low two bits.  The reference stores all minima and
// decides in the traceback; deciding in the fill is equivalent because `swaped` is known per pair, and it shrinks
// the direction band to 2 bits per cell (16 cells per 32-bit word per lane per step).
#pragma once
#include "stripe_kernels.cuh"

namespace poyb200 {

struct LinWinRow {
    int lut;    // byte offset of row a in the cost LUT
    int cdel;   // 4 * cost(a, gap) + tag of DELETE
};
struct LinWinCol {
    int lut;    // 4 * b
    int cins;   // 4 * cost(gap, b) + tag of INSERT
};

template <int K, int G, bool BT, bool LOW>
struct LinStripe {
    static constexpr int Q = 2 * K;
    int M[Q];
    int floorv[LOW ? Q : 1];  // LIN_INF on diagonals below dlo, INT_MIN elsewhere
    LinWinRow R[K];
    LinWinCol C[K + 1];
    const uint8_t *s1, *s2;   // shared memory copies (rows, columns)
    const uint8_t *lut;       // int entries 4*cost[a][b], rows of lut_row_bytes
    const int *gaprow;        // 4 * cost(a, gap)
    const int *gapcol;        // 4 * cost(gap, b)
    const int *prep4, *tail4; // 4 * prepend[b], 4 * tail[a]
    int lut_row_bytes, nr, nc, lane, ins_tag, del_tag, full;

    __device__ __forceinline__ LinWinRow make_row(int i) {
        const int a = s1[min(max(i, 0), nr)];
        LinWinRow r;
        r.lut = a * lut_row_bytes;
        r.cdel = gaprow[a] + del_tag;
        return r;
    }
    __device__ __forceinline__ LinWinCol make_col(int j) {
        const int b = s2[min(max(j, 0), nc)];
        LinWinCol c;
        c.lut = b * 4;
        c.cins = gapcol[b] + ins_tag;
        return c;
    }
    __device__ __forceinline__ void init_windows(int i0, int j0) {
#pragma unroll
        for (int m = 0; m < K; m++) R[m] = make_row(i0 - m);
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

    // BND: row 0 / column 0 may be present; ENDP: the last column may be present (only needed for custom tail costs:
    // with the default tail[a] = cost(a, gap) the extra candidate equals the DELETE candidate)
    template <bool BND, bool ENDP>
    __device__ __forceinline__ int cell(int q, int i, int j, int ml, int mu, int md, const LinWinRow &r, const LinWinCol &c) {
        const int c_al = *reinterpret_cast<const int *>(lut + r.lut + c.lut);
        int v = min(md + c_al, min(ml + c.cins, mu + r.cdel));  // :381-431, ties resolved by the tags
        if (ENDP) {
            // algn_fill_last_column (:548-560): one more DELETE candidate, priced with the tail cost
            if (j == nc) v = min(v, mu + tail4[s1[min(max(i, 0), nr)]] + del_tag);
        }
        if (BND) {
            if (i == 0) {
                if (j == 0) v = 0;  // ALIGN (:587-588)
                else v = ml + prep4[s2[min(max(j, 0), nc)]] + ins_tag;  // :597-598
            } else if (j == 0) {
                const int a = s1[min(max(i, 0), nr)];
                v = mu + (full ? gaprow[a] : tail4[a]) + del_tag;  // :570 / :659
            }
        }
        if (LOW) v = max(v, floorv[q]);
        return v;
    }

    // mid_q >= 0: slot mid_q receives mid_val between the two halves (the seed of cell (0, 0) when that cell belongs to the
    // odd half: placed any earlier it would leak into the cells outside the matrix that the even half computes)
    template <bool BND, bool ENDP>
    __device__ __forceinline__ void double_step(int i0, int j0, uint32_t &de, uint32_t &dod, int mid_q = -1, int mid_val = 0) {
        int in_l = __shfl_up_sync(0xffffffffu, M[Q - 1], 1, G);
        if (lane == 0) in_l = LIN_INF;
        de = 0;
        dod = 0;
#pragma unroll
        for (int m = 0; m < K; m++) {
            const int q = 2 * m;
            const int v = cell<BND, ENDP>(q, i0 - m, j0 + m, (m == 0) ? in_l : M[q - 1], M[q + 1], M[q], R[m], C[m]);
            M[q] = v & ~3;
            if (BT) de |= (uint32_t) (v & 3) << (2 * m);
        }
        if (mid_q >= 0) {
#pragma unroll
            for (int q = 1; q < Q; q += 2)
                if (q == mid_q) M[q] = mid_val;
        }
        int in_u = __shfl_down_sync(0xffffffffu, M[0], 1, G);
        if (lane == G - 1) in_u = LIN_INF;
#pragma unroll
        for (int m = 0; m < K; m++) {
            const int q = 2 * m + 1;
            const int v = cell<BND, ENDP>(q, i0 - m, j0 + m + 1, M[q - 1], (m == K - 1) ? in_u : M[q + 1], M[q], R[m], C[m + 1]);
            M[q] = v & ~3;
            if (BT) dod |= (uint32_t) (v & 3) << (2 * m);
        }
    }
};

// Shared memory: LUT (dim rows of dim + 1 ints), gaprow[dim], gapcol[dim], prep4[dim], tail4[dim], sequences.
#ifndef LIN_STRIPE_MIN_BLOCKS
#define LIN_STRIPE_MIN_BLOCKS STRIPE_MIN_BLOCKS
#endif
template <int K, int G, bool BT>
__global__ void __launch_bounds__(STRIPE_WARPS * 32, LIN_STRIPE_MIN_BLOCKS) lin_stripe_kernel(const Task *__restrict__ tasks, int ntasks, DevCM cm,
                                                                          const uint8_t *__restrict__ pool,
                                                                          uint8_t *__restrict__ dir, int *__restrict__ out_cost,
                                                                          int seq_bytes, int nslots, int custom_tail, int *work_counter) {
    constexpr int GPW = 32 / G;
    constexpr int Q = 2 * K;
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
        else { t = Task{}; t.lr = 1; t.lc = 1; t.dhi = 0; t.dlo = 0; }
        const int nr = t.lr - 1, nc = t.lc - 1;
        ring.wait_full(slot);
        const uint8_t *my_seq = ring.rows(slot);

        const int d0 = t.dhi + 1 - Q * G;
        const int u_first = (-d0) >> 1;
        const int u_last = valid ? ((nr + nc - d0) >> 1) : (u_first - 1);
        int u_end = u_last, u_begin = u_first;
#pragma unroll
        for (int o = G; o < 32; o <<= 1) {
            u_end = max(u_end, __shfl_xor_sync(0xffffffffu, u_end, o));
            u_begin = min(u_begin, __shfl_xor_sync(0xffffffffu, u_begin, o));
        }
        const int qlow_all = t.dlo - d0;
        const bool low = qlow_all > 0;
        uint8_t *dbase = dir + t.dir_off;
        const int dd_f = (nc - nr) - d0, lane_f = dd_f / Q, q_f = dd_f % Q;
        const bool swaped = (t.flags & TF_SWAPED) != 0;
        int result = 0;

        auto run = [&](auto lowtag) {
            constexpr bool LOW = decltype(lowtag)::value;
            LinStripe<K, G, BT, LOW> S;
            S.s1 = my_seq; S.s2 = my_seq + seq_bytes;
            S.lut = reinterpret_cast<const uint8_t *>(s_lut); S.lut_row_bytes = row_ints * 4;
            S.gaprow = s_gaprow; S.gapcol = s_gapcol; S.prep4 = s_prep; S.tail4 = s_tail;
            S.nr = nr; S.nc = nc; S.lane = lane; S.full = (t.flags & TF_FULL) != 0;
            S.ins_tag = swaped ? 1 : 2;
            S.del_tag = swaped ? 2 : 1;
            const int qlow = min(max(qlow_all - lane * Q, 0), Q);
#pragma unroll
            for (int q = 0; q < Q; q++) {
                S.M[q] = LIN_INF;
                if (LOW) S.floorv[q] = (q < qlow) ? LIN_INF : (int) 0x80000000;
            }
            int u = u_begin;
            int i0 = u - lane * K, j0 = u + d0 + lane * K;
            S.init_windows(i0, j0);
            // default prepend / tail costs (bit 1 of custom_tail): row 0 and column 0 are what the ordinary cell computes from
            // "no neighbour" inputs once cell (0, 0) is seeded through its diagonal input: no boundary phase at all
            const bool natural = (custom_tail & 2) != 0;
            const int dd0 = -d0, lane0 = dd0 / Q, q0 = dd0 - lane0 * Q;  // the slot of diagonal 0 (if dd0 >= 0)
            int seed = 0;
            if (natural) seed = -*reinterpret_cast<const int *>(S.lut + S.s1[0] * S.lut_row_bytes + S.s2[0] * 4);
            int u_b = natural ? (int) 0x80000000 : max(G * K, 1 - d0);  // from here on every lane has i >= 1 and j >= 1
            // first double step in which some lane can touch the last column: j0 + K >= nc for the last lane
            int u_e = (custom_tail & 1) ? (nc - d0 - G * K) : 0x7fffffff;
#pragma unroll
            for (int o = G; o < 32; o <<= 1) {
                u_b = max(u_b, __shfl_xor_sync(0xffffffffu, u_b, o));
                u_e = min(u_e, __shfl_xor_sync(0xffffffffu, u_e, o));
            }
            auto chunk = [&](int T) { return dbase + (((size_t) (T >> 3) * G + lane) * 8 + (T & 7)) * 4; };
            auto emit = [&](uint32_t de, uint32_t dod) {
                if (BT) {
                    const int te = 2 * u + d0;
                    if (u >= u_first && u <= u_last) {
                        if (te >= 0) *reinterpret_cast<uint32_t *>(chunk(te)) = de;
                        if (te + 1 <= nr + nc) *reinterpret_cast<uint32_t *>(chunk(te + 1)) = dod;
                    }
                }
                if (u == u_last && lane == lane_f) {
                    int r = 0;
#pragma unroll
                    for (int q = 0; q < Q; q++)
                        if (q == q_f) r = S.M[q] >> 2;
                    result = r;
                }
            };
            // u_b / u_e are warp-uniform (max / min over the groups of the warp), so the phases of different groups may
            // overlap: all four combinations exist
            for (; u <= u_end; u++) {
                uint32_t de, dod;
                const bool bnd = u < u_b, endp = u >= u_e;
                if (natural && __any_sync(0xffffffffu, u == u_first)) {  // (the groups of a warp may start at different steps)
                    // the double step that holds cell (0, 0): on its even half when d0 is even (seed placed before the step),
                    // on its odd half otherwise (seed placed between the halves)
                    const bool mine = u == u_first && dd0 >= 0 && lane == lane0;
                    if (mine && !(q0 & 1)) {
#pragma unroll
                        for (int q = 0; q < Q; q += 2)
                            if (q == q0) S.M[q] = seed;
                    }
                    const int mid_q = (mine && (q0 & 1)) ? q0 : -1;
                    if (endp) S.template double_step<false, true>(i0, j0, de, dod, mid_q, seed);
                    else S.template double_step<false, false>(i0, j0, de, dod, mid_q, seed);
                    emit(de, dod);
                    i0++; j0++;
                    S.slide_windows(i0, j0);
                    continue;
                }
                if (bnd && endp) S.template double_step<true, true>(i0, j0, de, dod);
                else if (bnd) S.template double_step<true, false>(i0, j0, de, dod);
                else if (endp) S.template double_step<false, true>(i0, j0, de, dod);
                else S.template double_step<false, false>(i0, j0, de, dod);
                emit(de, dod);
                i0++; j0++;
                S.slide_windows(i0, j0);
            }
        };
        const bool any_low = __any_sync(0xffffffffu, low);
        if (any_low) run(std::true_type{});
        else run(std::false_type{});

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
static inline size_t lin_table_bytes(int lcm) {
    const size_t dim = (size_t) 1 << lcm;
    return (dim * (dim + 1) + 4 * dim + 1) * sizeof(int) + STRIPE_WARPS * 4 * STAGE_BAR_BYTES + 16;
}

#ifdef POYB200_DEFINE_LIN_STRIPE  // the translation unit that owns these kernels (k_lin_stripe.cu)
template <int K, int G>
static cudaError_t lin_stripe_launch_shape(bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool, uint8_t *dir,
                                           int *cost, int sm_count, int seq_bytes, int custom_tail, int *work_counter, cudaStream_t stream) {
    constexpr int GPW = 32 / G;
    const int nbatches = (n + GPW - 1) / GPW;
    auto kern = bt ? lin_stripe_kernel<K, G, true> : lin_stripe_kernel<K, G, false>;
    size_t smem = 0;
    int nslots = 1, per_sm = 1;
    cudaError_t e = stage_ring_config(kern, lin_table_bytes(cm.lcm), (size_t) STRIPE_WARPS * GPW * 2 * seq_bytes, STRIPE_WARPS * 32, smem,
                                      nslots, per_sm);
    if (e != cudaSuccess) return e;
    int blocks = std::min((nbatches + STRIPE_WARPS - 1) / STRIPE_WARPS, sm_count * per_sm);
    if (blocks < 1) blocks = 1;
    kern<<<blocks, STRIPE_WARPS * 32, smem, stream>>>(d_tasks, n, cm, pool, dir, cost, seq_bytes, nslots, custom_tail, work_counter);
    return cudaGetLastError();
}

cudaError_t lin_stripe_launch(uint32_t klass, bool bt, const Task *d_tasks, int n, DevCM cm, const uint8_t *pool,
                                            uint8_t *dir, int *cost, int sm_count, int seq_bytes, int custom_tail,
                                            int *work_counter, cudaStream_t stream) {
#define LIN_CASE(IDX, KK, GG) \
    case IDX: return lin_stripe_launch_shape<KK, GG>(bt, d_tasks, n, cm, pool, dir, cost, sm_count, seq_bytes, custom_tail, work_counter, stream)
    switch (klass - KLASS_LIN_BASE) {
        LIN_CASE(0, 8, 8);
        LIN_CASE(1, 10, 8);
        LIN_CASE(2, 12, 8);
        LIN_CASE(3, 8, 16);
        LIN_CASE(4, 12, 16);
        LIN_CASE(5, 8, 32);
        LIN_CASE(6, 10, 32);
        LIN_CASE(7, 16, 32);
        default: return cudaErrorInvalidValue;
    }
#undef LIN_CASE
}
#endif  // POYB200_DEFINE_LIN_STRIPE

}  // namespace poyb200
