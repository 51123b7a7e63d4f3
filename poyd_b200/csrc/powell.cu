// Powell's three-sequence affine Ukkonen aligner on the device (powell_core.h) and its C ABI entry point,
// poyb200_batch_powell_3: replaces the external powell_3D_align (src/ukkCommon.c:110-145) as Sequence.Align.align_3_powell /
// align_3_powell_inter call it (src/sequence.ml:1075-1114) for readjust_3d (:1116-1139).
//
// One CTA per triple: the cells a top level newly asks for are listed (relaxation of the demand closure, a counting sort by
// cost) and computed by all threads, cost level by cost level, between block barriers; U, the demand state and the lists
// live in a per-CTA workspace in HBM (it is L2-resident for the usual sizes).  The diagonal box of a workspace has a radius
// R; a triple whose computation touches the border reports PW_EBOX and is run again in the next round with 2 R.
#include <algorithm>
#include <vector>

#include "ctx.h"

#define PW_HD __device__ __forceinline__
#define PW_TID ((int) threadIdx.x)
#define PW_NT ((int) blockDim.x)
#define PW_SYNC() __syncthreads()
#define PW_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define PW_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define PW_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#include "powell_core.h"

namespace poyb200 {
namespace powell {

// Launch shapes (measured, profiles/r02_powell_probe.txt): a cost level is a chain of dependent loads, so the triples of an SM
// overlap each other's latency -- 256 threads x 2 CTAs per SM is the best throughput when there are more triples than SMs;
// a small batch of heavy triples finishes sooner with 512 threads on each (128 x 4 was 2 - 3 x slower than either).
constexpr int PW_THREADS_MANY = 256, PW_CTAS_MANY = 2;
constexpr int PW_THREADS_FEW = 512, PW_CTAS_FEW = 1;

struct Job {
    uint32_t off[3];
    int32_t len[3];   // stored lengths (leading gap included)
    uint32_t triple;  // index in the caller's list
};

struct OutP {
    int *cost, *out_len, *status;  // out_len: 2 per triple (aligned length, median length)
    uint8_t *r[3], *median;
    long long stride;
    uint32_t want;
    const uint8_t *median3;        // 3-D median table (may be null)
    int lcm3, gap;
};

template <int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) powell_kernel(const Job *__restrict__ jobs, int njobs, const uint8_t *__restrict__ pool,
                                                           const Tables *__restrict__ tables, Work *works, uint8_t *seqbuf,
                                                           int seqcap, int seq_shared, OutP out, int *counter) {
    extern __shared__ __align__(16) uint8_t s_seq[];  // the three sequences of the triple, when they fit (seq_shared)
    __shared__ Tables s_tb;
    __shared__ int s_job;
    for (int k = threadIdx.x; k < (int) (sizeof(Tables) / sizeof(int)); k += blockDim.x)
        reinterpret_cast<int *>(&s_tb)[k] = reinterpret_cast<const int *>(tables)[k];
    Work *w = &works[blockIdx.x];
    uint8_t *sq[3];
    for (int k = 0; k < 3; k++) sq[k] = seq_shared ? s_seq + (size_t) k * seqcap : seqbuf + ((size_t) blockIdx.x * 3 + k) * seqcap;
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) s_job = atomicAdd(counter, 1);
        __syncthreads();
        const int j = s_job;
        __syncthreads();
        if (j >= njobs) break;
        const Job job = jobs[j];
        if (threadIdx.x == 0) {
            w->status = PW_OK;
            w->A = sq[0]; w->B = sq[1]; w->C = sq[2];
            w->Alen = job.len[0] - 1; w->Blen = job.len[1] - 1; w->Clen = job.len[2] - 1;
            w->cab = (w->Alen - w->Blen) / 2;
            w->cac = (w->Alen - w->Clen) / 2;
            w->ncalc = 0;
        }
        __syncthreads();
        // copySequence (src/ukkCommon.c:87-108): the lowest base of every code; a code without one is an error there too
        for (int k = 0; k < 3; k++) {
            const uint8_t *src = pool + job.off[k];
            for (int i = threadIdx.x; i < job.len[k]; i += blockDim.x) {
                int v = 0;
                if (i + 1 < job.len[k]) {
                    const int c = src[i + 1];
                    v = (c & 1) ? 1 : (c & 2) ? 2 : (c & 4) ? 4 : (c & 8) ? 8 : 0;
                    if (v == 0) w->status = PW_EINPUT;
                }
                sq[k][i] = (uint8_t) v;  // one readable element past the end, equal to no base
            }
        }
        const int nx = w->D * w->D * NS;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) { w->top[i] = NEGBIG; w->prev[i] = NEGBIG; }
        if (w->nextOffset > (1ll << 30)) {  // the 32-bit tags would come round: forget every cell
            const size_t ne = (size_t) nx * w->Wd;
            for (size_t i = threadIdx.x; i < ne; i += blockDim.x) w->U[i].tag = -1;
            __syncthreads();
            if (threadIdx.x == 0) w->nextOffset = 1;
        }
        __syncthreads();
        int cost = -1;
        if (w->status == PW_OK) {
            Engine e;
            e.w = w;
            e.tb = &s_tb;
            cost = e.run();
        }
        __syncthreads();
        const int st = w->status, n = w->nres + 1;
        if (threadIdx.x == 0) {
            w->nextOffset = w->costOffset + w->maxlevels + 2;
            out.cost[job.triple] = cost;
            out.status[job.triple] = st;
            out.out_len[2 * job.triple] = st ? 0 : n;
            out.out_len[2 * job.triple + 1] = 0;
        }
        if (st == PW_OK && (out.want & (POYB200_WANT3_ALIGNED | POYB200_WANT3_MEDIAN))) {
            // printTraceBack (:477-495): the rows, forward, behind one gap; right aligned like every output row
            const uint8_t *res[3] = {w->resA, w->resB, w->resC};
            if (out.want & POYB200_WANT3_ALIGNED)
                for (int k = 0; k < 3; k++) {
                    uint8_t *row = out.r[k] + (size_t) job.triple * out.stride + (out.stride - n);
                    for (int i = threadIdx.x; i < n; i += blockDim.x) {
                        int v = 16;
                        if (i > 0) { v = res[k][w->nres - i]; if (v == 0xff) v = 16; }
                        row[i] = (uint8_t) v;
                    }
                }
            if ((out.want & POYB200_WANT3_MEDIAN) && threadIdx.x == 0) {
                // align_3_powell_inter (src/sequence.ml:1103-1114): the 3-D median of every column, gaps dropped, one gap in front
                uint8_t *row = out.median + (size_t) job.triple * out.stride;
                long long pos = out.stride;
                for (int i = n - 1; i >= 0; i--) {
                    int v[3];
                    for (int k = 0; k < 3; k++) { v[k] = i > 0 ? res[k][w->nres - i] : 16; if (v[k] == 0xff) v[k] = 16; }
                    const int m = out.median3[(((size_t) v[0] << out.lcm3) + v[1] << out.lcm3) + v[2]];
                    if (m != out.gap) row[--pos] = (uint8_t) m;
                }
                row[--pos] = (uint8_t) out.gap;
                out.out_len[2 * job.triple + 1] = (int) (out.stride - pos);
            }
        }
        __syncthreads();
    }
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Layout {
    size_t nx, u, top, prev, snap, keycnt, list, res, stack, total;
    int Wd, maxlevels, rescap, listcap;
};
static Layout make_layout(int R, int Wd, int maxlevels, int rescap) {
    Layout l{};
    const size_t D = 2 * (size_t) R + 1;
    l.nx = D * D * NS;
    l.Wd = Wd; l.maxlevels = maxlevels; l.rescap = rescap;
    l.listcap = (int) std::min<size_t>(2 * l.nx, (size_t) 1 << 30);
    size_t off = 0;
    l.u = off; off += align_up(l.nx * Wd * sizeof(Entry), 256);
    l.top = off; off += align_up(l.nx * sizeof(int), 256);
    l.prev = off; off += align_up(l.nx * sizeof(int), 256);
    l.snap = off; off += align_up(l.nx * sizeof(int), 256);
    l.keycnt = off; off += align_up((2 * ((size_t) maxlevels + 1) + 1 + 1024) * sizeof(int), 256);
    l.list = off; off += align_up((size_t) l.listcap * 2 * sizeof(int), 256);
    l.res = off; off += align_up(3 * (size_t) rescap, 256);
    l.stack = off; off += align_up(256 * sizeof(poyb200::powell::Task), 256);
    l.total = off;
    return l;
}

}  // namespace powell
}  // namespace poyb200

using namespace poyb200::powell;

extern "C" int poyb200_batch_powell_3(poyb200_ctx *ctx, const poyb200_batch3 *b, int32_t mm, int32_t go, int32_t ge) {
    if (!ctx || !b) return POYB200_EINVAL;
    const int n = b->n_triples;
    if (n < 0 || b->n_seqs < 0) return fail(ctx, POYB200_EINVAL, "negative count");
    if (mm <= 0 || go < 0 || ge <= 0) return fail(ctx, POYB200_EINVAL, "powell_3D_align needs mismatch > 0, gap opening >= 0, gap extension > 0");
    if (n == 0) return POYB200_OK;
    if (!b->pool || !b->seq_off || !b->seq_len || !b->triples || !b->cost || !b->out_len || !b->status)
        return fail(ctx, POYB200_EINVAL, "NULL array (cost, out_len and status are required)");
    if (b->pool_bytes >= ((size_t) 1 << 32)) return fail(ctx, POYB200_EINVAL, "pool larger than 4 GiB");
    const bool want_al = (b->want & POYB200_WANT3_ALIGNED) != 0, want_med = (b->want & POYB200_WANT3_MEDIAN) != 0;
    if (want_al && (!b->aligned_1 || !b->aligned_2 || !b->aligned_3)) return fail(ctx, POYB200_EINVAL, "WANT3_ALIGNED without buffers");
    if (want_med && !b->median) return fail(ctx, POYB200_EINVAL, "WANT3_MEDIAN without buffer");
    if (want_med && !ctx->has_cm3) return fail(ctx, POYB200_ENOCM, "WANT3_MEDIAN needs the 3-D cost matrix (poyb200_set_cm_3d)");
    cudaSetDevice(ctx->device);
    for (int s = 0; s < b->n_seqs; s++) {
        if (b->seq_len[s] < 1 || b->seq_len[s] > POYB200_MAX_SEQ_LEN) return fail(ctx, POYB200_ESEQLEN, "sequence empty or longer than 16384");
        if (b->seq_off[s] < 0 || (size_t) (b->seq_off[s] + b->seq_len[s]) > b->pool_bytes) return fail(ctx, POYB200_EINVAL, "sequence outside the pool");
    }
    std::vector<Job> jobs((size_t) n);
    std::vector<int> needR((size_t) n);
    long long maxcap = 16;
    int maxlen = 1, maxsum = 3;
    for (int p = 0; p < n; p++) {
        Job j{};
        for (int k = 0; k < 3; k++) {
            const int idx = b->triples[3 * p + k];
            if (idx < 0 || idx >= b->n_seqs) return fail(ctx, POYB200_EINVAL, "triple index out of range");
            j.off[k] = (uint32_t) b->seq_off[idx];
            j.len[k] = b->seq_len[idx];
            maxlen = std::max(maxlen, j.len[k]);
        }
        j.triple = (uint32_t) p;
        jobs[p] = j;
        const int sum = j.len[0] + j.len[1] + j.len[2];
        maxsum = std::max(maxsum, sum);
        maxcap = std::max<long long>(maxcap, sum);
        // the box is centred between the first and the last diagonal and must hold both strictly inside
        const int fab = j.len[0] - j.len[1], fac = j.len[0] - j.len[2];
        needR[p] = std::max(std::abs(fab), std::abs(fac)) / 2 + 3;
    }
    if (b->out_stride < maxcap && (want_al || want_med)) return fail(ctx, POYB200_EINVAL, "out_stride smaller than l1 + l2 + l3");
    const long long dstride = (maxcap + 15) & ~15ll;
    Tables tb;
    make_tables(tb, mm, go, ge);
    int Wd = 16;
    while (Wd < 4 * tb.maxSingleStep + 2 * go + 8) Wd *= 2;
    // no alignment costs more than gapping every element once; the cost fields of a cell are 16 bits wide, so a triple whose
    // cost would pass 32 000 ends with status PW_ECAP (the reference's own INFINITY is 5 000, ukkCommon.h:41)
    const long long lv = 2ll * maxsum * std::max(ge, mm) + 6ll * go + 16;
    const int maxlevels = (int) std::min<long long>(lv, 32000), rescap = maxsum + 1, seqcap = (maxlen + 16) & ~15;

    CK(ctx->d_pool.reserve(b->pool_bytes + 64));
    CK(ctx->d_costs.reserve((size_t) n + 1));
    CK(ctx->d_outlen.reserve(2 * (size_t) n + 4));
    CK(ctx->d_status.reserve((size_t) n + 4));
    const size_t ob = (size_t) n * dstride + 16;
    if (want_al) { CK(ctx->d_out[0].reserve(ob)); CK(ctx->d_out[1].reserve(ob)); CK(ctx->d_out[2].reserve(ob)); }
    if (want_med) CK(ctx->d_out[3].reserve(ob));
    CK(ctx->d_counters.reserve(64));
    CK(cudaMemcpyAsync(ctx->d_pool.p, b->pool, b->pool_bytes, cudaMemcpyHostToDevice, ctx->stream));
    Tables *d_tb = nullptr;
    Job *d_jobs = nullptr;
    CK(cudaMalloc(&d_tb, sizeof(Tables)));
    struct Guard {  // this call's scratch allocations
        std::vector<void *> p;
        ~Guard() { for (void *q : p) cudaFree(q); }
    } guard;
    guard.p.push_back(d_tb);
    CK(cudaMalloc(&d_jobs, (size_t) n * sizeof(Job)));
    guard.p.push_back(d_jobs);
    CK(cudaMemcpyAsync(d_tb, &tb, sizeof tb, cudaMemcpyHostToDevice, ctx->stream));
    OutP out{ctx->d_costs.p, ctx->d_outlen.p, ctx->d_status.p, {ctx->d_out[0].p, ctx->d_out[1].p, ctx->d_out[2].p}, ctx->d_out[3].p,
             dstride, b->want, ctx->has_cm3 ? ctx->dcm3.median : nullptr, ctx->has_cm3 ? ctx->dcm3.lcm : 0, ctx->has_cm3 ? ctx->dcm3.gap : 16};

    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = std::min<size_t>(free_b / 2, (size_t) 64 << 30);
    std::vector<int> status((size_t) n, PW_EBOX);
    std::vector<int> pending((size_t) n);
    for (int p = 0; p < n; p++) pending[p] = p;
    std::vector<Job> round_jobs;
    for (int R = 16; !pending.empty(); R *= 2) {
        if (R > 2048) return fail(ctx, POYB200_EINVAL, "Powell kernel: diagonal box larger than 2048 needed");
        round_jobs.clear();
        std::vector<int> later;
        for (int p : pending) (needR[p] <= R ? (void) round_jobs.push_back(jobs[p]) : (void) later.push_back(p));
        if (round_jobs.empty()) { pending.swap(later); continue; }
        const Layout lay = make_layout(R, Wd, maxlevels, rescap);
        const size_t per_cta = lay.total + 3 * (size_t) seqcap + 256;
        const bool many = round_jobs.size() > (size_t) ctx->sm_count;
        const int ctas = many ? PW_CTAS_MANY : PW_CTAS_FEW;
        int grid = (int) std::min<size_t>(std::min<size_t>(round_jobs.size(), (size_t) ctx->sm_count * ctas), std::max<size_t>(1, budget / per_cta));
        if (per_cta > budget) return fail(ctx, POYB200_ENOMEM, "Powell kernel: one workspace exceeds the memory budget");
        // the workspaces stay with the context between calls (allocating tens of GB costs more than aligning the triples)
        CK(ctx->d_pw_arena.reserve(lay.total * (size_t) grid));
        CK(ctx->d_pw_seq.reserve(3 * (size_t) seqcap * grid + sizeof(Work) * (size_t) grid + 256));
        uint8_t *arena = ctx->d_pw_arena.p, *seqbuf = ctx->d_pw_seq.p;
        Work *d_works = reinterpret_cast<Work *>(ctx->d_pw_seq.p + ((3 * (size_t) seqcap * grid + 255) & ~(size_t) 255));
        std::vector<Work> hw((size_t) grid);
        for (int g = 0; g < grid; g++) {
            Work &w = hw[g];
            memset(&w, 0, sizeof w);
            uint8_t *base = arena + lay.total * (size_t) g;
            w.R = R; w.D = 2 * R + 1; w.Wd = Wd;
            w.U = reinterpret_cast<Entry *>(base + lay.u);
            w.top = reinterpret_cast<int *>(base + lay.top);
            w.prev = reinterpret_cast<int *>(base + lay.prev);
            w.snap = reinterpret_cast<int *>(base + lay.snap);
            w.keycnt = reinterpret_cast<int *>(base + lay.keycnt);
            w.list = reinterpret_cast<int *>(base + lay.list);
            w.listcap = lay.listcap; w.maxlevels = maxlevels; w.keycap = 2 * (maxlevels + 1) + 1;
            w.resA = base + lay.res; w.resB = w.resA + rescap; w.resC = w.resB + rescap; w.rescap = rescap;
            w.stack = reinterpret_cast<poyb200::powell::Task *>(base + lay.stack); w.stackcap = 256;
            w.nextOffset = 1;
        }
        CK(cudaMemcpyAsync(d_works, hw.data(), sizeof(Work) * (size_t) grid, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(arena, 0xff, lay.total * (size_t) grid, ctx->stream));  // every tag = -1: nothing computed
        CK(cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(int), ctx->stream));
        CK(cudaMemcpyAsync(d_jobs, round_jobs.data(), round_jobs.size() * sizeof(Job), cudaMemcpyHostToDevice, ctx->stream));
        const int seq_shared = 3 * (size_t) seqcap <= 40 * 1024;  // within the default dynamic shared memory limit
        const size_t smem = seq_shared ? 3 * (size_t) seqcap : 0;
        if (many)
            powell_kernel<PW_THREADS_MANY, PW_CTAS_MANY><<<grid, PW_THREADS_MANY, smem, ctx->stream>>>(
                d_jobs, (int) round_jobs.size(), ctx->d_pool.p, d_tb, d_works, seqbuf, seqcap, seq_shared, out, ctx->d_counters.p);
        else
            powell_kernel<PW_THREADS_FEW, PW_CTAS_FEW><<<grid, PW_THREADS_FEW, smem, ctx->stream>>>(
                d_jobs, (int) round_jobs.size(), ctx->d_pool.p, d_tb, d_works, seqbuf, seqcap, seq_shared, out, ctx->d_counters.p);
        CK(cudaGetLastError());
        ctx->launches++;
        CK(cudaMemcpyAsync(status.data(), ctx->d_status.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (const Job &j : round_jobs)
            if (status[j.triple] == PW_EBOX) later.push_back((int) j.triple);
        std::sort(later.begin(), later.end());
        for (int p : later) needR[p] = std::max(needR[p], R + 1);  // the next round that takes it is 2 R
        pending.swap(later);
    }
    CK(cudaMemcpyAsync(b->cost, ctx->d_costs.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(b->out_len, ctx->d_outlen.p, 2 * (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(b->status, ctx->d_status.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    uint8_t *dst[4] = {b->aligned_1, b->aligned_2, b->aligned_3, b->median};
    const bool need[4] = {want_al, want_al, want_al, want_med};
    const size_t wbytes = (size_t) std::min<long long>(dstride, b->out_stride);
    for (int k = 0; k < 4; k++)
        if (need[k])
            CK(cudaMemcpy2DAsync(dst[k] + (b->out_stride - wbytes), (size_t) b->out_stride, ctx->d_out[k].p + (dstride - wbytes), (size_t) dstride,
                                 wbytes, (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_pw_arena.cap > ((size_t) 48 << 30)) ctx->d_pw_arena.release();  // an unusually large box: do not sit on it
    ctx->staged = false;
    return POYB200_OK;
}
