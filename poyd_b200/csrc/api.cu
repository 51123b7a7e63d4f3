// Host side of the C ABI (include/poyb200.h): planning, device memory, launches.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <vector>

#include "ctx.h"

// ---------------------------------------------------------------------------------------------------------
// geometry: which cells the reference visits
// ---------------------------------------------------------------------------------------------------------
struct LinBand {
    bool full;
    int dlo, dhi;
};

// algn_nw_limit (src/algn.c:3265-3266) + algn_fill_plane_2 (:880-966): l1 >= l2 stored lengths.
static LinBand linear_band(int l1, int l2, int deltaw) {
    int width = 50 + deltaw, height = (l1 - l2) + 50 + deltaw;
    if (width > l2) width = l2;
    if (height > l1) height = l1;
    LinBand b{false, 0, width - 1};
    if ((float) l1 >= ((float) 3 / (float) 2) * (float) l2) b.full = true;  // Case 1 :893
    else if (2 * height < l1) b.dlo = 1 - height;                           // Case 2 :898
    else if (8 >= l1 - height) b.full = true;                               // Case 3a :936
    else b.dlo = -((l1 - l2) + width);                                      // Case 3b :938-965
    if (b.full) {
        b.dlo = -(l1 - 1);
        b.dhi = l2 - 1;
    }
    return b;
}

// algn_fill_plane_3_aff (:2435-2445, 2475): rows = shorter; nr, nc = lengths without the leading gap.
static void affine_band(int nr, int nc, int &dlo, int &dhi) {
    int e = nc - nr + 8;
    if (e < 40) e = 40;
    if (e > nc) e = nc;
    dlo = -39;
    dhi = e - 1;
}

extern "C" int64_t poyb200_cells_linear(int32_t l1, int32_t l2, int32_t deltaw) {
    if (l1 < l2) std::swap(l1, l2);
    LinBand b = linear_band(l1, l2, deltaw);
    if (b.full) return (int64_t) l1 * l2;
    int64_t n = 0;
    for (int i = 0; i < l1; i++) {
        int lo = std::max(0, i + b.dlo), hi = std::min(l2 - 1, i + b.dhi);
        if (hi >= lo) n += hi - lo + 1;
    }
    return n;
}

extern "C" int64_t poyb200_cells_affine(int32_t la, int32_t lb) {
    int nr = std::min(la, lb) - 1, nc = std::max(la, lb) - 1;
    int s = 1, e = nc - nr + 8;
    if (e < 40) e = 40;
    if (e > nc) e = nc;
    int64_t n = nc + 1;  // the initialised row
    for (int i = 1; i <= nr; i++) {
        if (i > 40) s++;
        n += e - s + 2;  // band cells + the left-edge cell
        if (e < nc) e++;
    }
    return n;
}

// ---------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------
extern "C" const char *poyb200_version(void) { return "poyb200 0.1 (sm_100a)"; }

extern "C" void poyb200_default_config(poyb200_config *cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_bytes = (uint32_t) sizeof *cfg;
    cfg->force_generic = 0;
    cfg->allow_fast = 1;
    cfg->allow_noeb = 1;
    cfg->use_ring = 2;
    cfg->overlap_traceback = 1;
    cfg->dir_buffers = 3;
    cfg->traceback_threads_per_sm = 256;  // measured with the round-2 kernels: 256 -> 799, 512 -> 768, 128 -> 765, 1024 -> 696 GCUPS (configs[1])
    cfg->traceback_block = 128;
    cfg->traceback_priority = 1;
    cfg->chunk_pairs = 1 << 16;
    cfg->host_threads = 0;
    cfg->timing = 1;
    cfg->trace = 0;
    cfg->dir_budget_bytes = 0;
    cfg->allow_rows = 1;
    cfg->small_ring_pairs = 0;
    cfg->dir6 = 1;
    cfg->pair2 = 1;
    cfg->pair2_min_pairs = 0;
}

extern "C" int poyb200_create_ex(int device, const poyb200_config *user, poyb200_ctx **out) {
    if (!out) return POYB200_EINVAL;
    *out = nullptr;
    poyb200_config cfg;
    poyb200_default_config(&cfg);
    if (user) {
        // a caller built against an older header passes a shorter struct: the fields it does not know keep their defaults
        if (user->struct_bytes < sizeof(uint32_t) || user->struct_bytes > sizeof cfg) return POYB200_EINVAL;
        memcpy(&cfg, user, user->struct_bytes);
        cfg.struct_bytes = (uint32_t) sizeof cfg;
    }
    if (cfg.dir_buffers != 2 && cfg.dir_buffers != 3) return POYB200_EINVAL;
    if (cfg.traceback_block != 32 && cfg.traceback_block != 64 && cfg.traceback_block != 128) return POYB200_EINVAL;
    if (cfg.traceback_threads_per_sm < cfg.traceback_block || cfg.chunk_pairs < 1 || cfg.host_threads < 0 || cfg.dir_budget_bytes < 0)
        return POYB200_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return POYB200_ECUDA;  // no CPU fallback
    poyb200_ctx *ctx = new poyb200_ctx();
    ctx->cfg = cfg;
    if (device >= 0) {
        if (cudaSetDevice(device) != cudaSuccess) {
            delete ctx;
            return POYB200_ECUDA;
        }
        ctx->device = device;
    } else {
        cudaGetDevice(&ctx->device);
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return POYB200_ECUDA;
    }
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->s_len, cudaStreamNonBlocking);
    {
        // The traceback stream outranks the compute stream: when a fill ends, the walkers of its chunk are placed before the
        // CTAs of the next (persistent, SM-filling) fill, instead of waiting for that fill to drain.
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&ctx->s_tb, cudaStreamNonBlocking, cfg.traceback_priority ? hi : lo) != cudaSuccess)
            cudaStreamCreateWithFlags(&ctx->s_tb, cudaStreamNonBlocking);
    }
    cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming);
    ctx->host_threads = cfg.host_threads > 0 ? cfg.host_threads : (int) std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    ctx->chunk_pairs = (size_t) cfg.chunk_pairs;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // direction bands of one chunk: a quarter of the free HBM, capped at 40 GB (up to three buffers of this size)
    ctx->dir_budget = cfg.dir_budget_bytes > 0 ? (size_t) cfg.dir_budget_bytes : std::min<size_t>(free_b / 4, (size_t) 40 << 30);
    *out = ctx;
    return POYB200_OK;
}

extern "C" int poyb200_create(int device, poyb200_ctx **out) { return poyb200_create_ex(device, nullptr, out); }

extern "C" void poyb200_destroy(poyb200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->d_cost.release(); ctx->d_prepend.release(); ctx->d_tail.release(); ctx->d_median.release(); ctx->d_worst.release();
    ctx->d_pool.release(); ctx->d_dir.release(); ctx->d_tasks.release(); ctx->d_costs.release();
    ctx->d_outlen.release(); ctx->d_lin_state.release(); ctx->d_aff_state.release(); ctx->d_counters.release(); ctx->d_slow_list.release(); ctx->d_slow_list2.release();
    ctx->tasks.release();
    ctx->tasks_tmp.release();
    ctx->d_cost3.release(); ctx->d_ring.release(); ctx->d_status.release(); ctx->d_median3.release(); ctx->d_tasks3.release();
    ctx->d_pw_arena.release(); ctx->d_pw_seq.release();
    for (auto &b : ctx->d_out) b.release();
    for (auto &b : ctx->d_bits) b.release();
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->chunk_ev) cudaEventDestroy(e);
    for (auto &e : ctx->ev_done) cudaEventDestroy(e);
    for (auto &e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->s_len) cudaStreamDestroy(ctx->s_len);
    if (ctx->s_tb) cudaStreamDestroy(ctx->s_tb);
    for (auto &e : ctx->ev_fill) cudaEventDestroy(e);
    for (auto &e : ctx->ev_tb) cudaEventDestroy(e);
    ctx->d_dir2.release();
    ctx->d_dir3.release();
    ctx->d_scratch.release();
    ctx->d_walked.release();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *poyb200_last_error(const poyb200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int64_t poyb200_launch_count(const poyb200_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" void *poyb200_stream(poyb200_ctx *ctx) { return ctx ? (void *) ctx->stream : nullptr; }

extern "C" void *poyb200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void poyb200_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

extern "C" int poyb200_set_cm(poyb200_ctx *ctx, const poyb200_cm *cm) {
    if (!ctx || !cm || !cm->cost || !cm->median || !cm->prepend_cost || !cm->tail_cost)
        return ctx ? fail(ctx, POYB200_EINVAL, "poyb200_set_cm: NULL table") : POYB200_EINVAL;
    if (cm->lcm < 1 || cm->lcm > 8) return fail(ctx, POYB200_EINVAL, "poyb200_set_cm: lcm out of range (SEQT is 8 bit)");
    cudaSetDevice(ctx->device);
    const size_t dim = (size_t) 1 << cm->lcm;
    CK(ctx->d_cost.reserve(dim * dim));
    CK(ctx->d_median.reserve(dim * dim));
    CK(ctx->d_prepend.reserve(dim));
    CK(ctx->d_tail.reserve(dim));
    CK(cudaMemcpyAsync(ctx->d_cost.p, cm->cost, dim * dim * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_median.p, cm->median, dim * dim, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_prepend.p, cm->prepend_cost, dim * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tail.p, cm->tail_cost, dim * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (cm->worst) {
        CK(ctx->d_worst.reserve(dim * dim));
        CK(cudaMemcpyAsync(ctx->d_worst.p, cm->worst, dim * dim * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->custom_tail = 0;
    for (size_t a2 = 0; a2 < dim; a2++)
        if (cm->tail_cost[a2] != cm->cost[(a2 << cm->lcm) + cm->gap]) ctx->custom_tail = 1;
    // default first-row / first-column costs (prepend[b] = cost(gap, b), tail[a] = cost(a, gap)): row 0 and column 0 then
    // come out of the ordinary recurrence, and the linear kernels skip their boundary phase (bit 1 of the flag they receive)
    ctx->lin_natural = ctx->custom_tail ? 0 : 1;
    for (size_t b2 = 0; b2 < dim; b2++)
        if (cm->prepend_cost[b2] != cm->cost[((size_t) cm->gap << cm->lcm) + b2]) ctx->lin_natural = 0;
    // aff_x2_kernel (16-bit halves): the largest value its tables can hold -- cost[a & 15][b & 15], cost[a][gap] and
    // prepend[a] for the 32 codes of a DNA matrix
    ctx->x2_unit4 = 1 << 30;
    if (cm->lcm >= 5 && dim >= 32) {
        long long mx = 0;
        bool ok = true;
        auto see = [&](long long v) { if (v < 0) ok = false; mx = std::max(mx, v); };
        for (size_t a2 = 0; a2 < 16; a2++)
            for (size_t b2 = 0; b2 < 16; b2++) see(cm->cost[(a2 << cm->lcm) + b2]);
        for (size_t a2 = 0; a2 < 32; a2++) { see(cm->cost[(a2 << cm->lcm) + cm->gap]); see(cm->prepend_cost[a2]); }
        if (ok && mx < 4096) ctx->x2_unit4 = (int) (4 * std::max<long long>(mx, 1));
    }
    ctx->hcm = *cm;
    ctx->hcm.cost = nullptr; ctx->hcm.median = nullptr; ctx->hcm.worst = nullptr;
    ctx->hcm.prepend_cost = nullptr; ctx->hcm.tail_cost = nullptr;
    ctx->dcm = DevCM{cm->a_sz, cm->lcm, cm->gap, cm->cost_model_type, cm->combinations, cm->gap_open,
                     ctx->d_cost.p, ctx->d_median.p, ctx->d_prepend.p, ctx->d_tail.p, cm->all_elements,
                     cm->worst ? ctx->d_worst.p : nullptr};
    ctx->has_cm = true;
    return POYB200_OK;
}


// ---------------------------------------------------------------------------------------------------------
// kernel classes
// ---------------------------------------------------------------------------------------------------------
static inline uint32_t round16(uint32_t v) { return (v + 15u) & ~15u; }

// Picks the fill kernel for one pair and fixes the layout of its direction band.
static void choose_class(Task &t, bool affine, bool bt, int W, const DevCM &cm, bool allow_stripe, bool allow_rows) {
    (void) bt;
    if (allow_stripe && affine && stripe_choose(t, affine, W, cm)) return;
    if (allow_stripe && allow_rows && !affine && lin_rows_choose(t, cm)) return;
    if (allow_stripe && !affine && lin_stripe_choose(t, W, cm)) return;
    t.klass = KLASS_GENERIC;
    t.dbase = t.dlo;
    t.G = 1;
    t.twoK = 0xFFFFu;
    t.BL = round16((uint32_t) (W + 2) / 2 + 1);
}

// Work counters: one zeroed int per persistent launch of a call (reset in bulk by reset_counters).
constexpr size_t MAX_COUNTERS = 1 << 16;
// A call that launches more than MAX_COUNTERS persistent kernels (tiny chunks) starts over on a re-zeroed array: the
// memset is ordered on the compute stream, which every launch of the call is ordered after (fills run on it, tracebacks
// wait for an event recorded on it), and a counter is only reused MAX_COUNTERS launches -- many chunks -- later.
static int *next_counter(poyb200_ctx *ctx) {
    if (ctx->counter_next == MAX_COUNTERS) {
        cudaStreamSynchronize(ctx->s_tb);
        cudaStreamSynchronize(ctx->stream);
        cudaMemsetAsync(ctx->d_counters.p, 0, MAX_COUNTERS * sizeof(int), ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        ctx->counter_next = 0;
    }
    return ctx->d_counters.p + ctx->counter_next++;
}
static int reset_counters(poyb200_ctx *ctx) {
    if (ctx->ring_mode == 2 && ctx->staged && (ctx->mode == MODE_ALIGN_AFF)) {
        CK(ctx->d_walked.reserve(ctx->tasks.size() + 16));
        CK(cudaMemsetAsync(ctx->d_walked.p, 0, ctx->tasks.size() + 16, ctx->stream));
    }
    CK(ctx->d_counters.reserve(MAX_COUNTERS));
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, MAX_COUNTERS * sizeof(int), ctx->stream));
    ctx->counter_next = 0;
    return POYB200_OK;
}

// True when the pairs of this class are filled AND walked by the ring kernels (no separate traceback launch).
static bool ring_class(const poyb200_ctx *ctx, uint32_t klass, bool affine) {
    return affine && ctx->ring_mode == 1 && ring_has_shape(klass);
}
// use_ring = 2: pairs WITHOUT gap bits take aff_fast_kernel + the traceback kernel, the batches it declines take the full
// ring instance (fill + walk); the traceback kernel skips what the ring instance walked (OutPtrs::walked).
static bool mixed_class(const poyb200_ctx *ctx, uint32_t klass, bool affine) {
    return affine && ctx->ring_mode == 2 && ring_has_shape(klass) && ctx->cfg.allow_fast && ctx->cfg.allow_noeb && ctx->dcm.gap_open > 0;
}

static int launch_fill(poyb200_ctx *ctx, uint32_t klass, bool affine, bool bt, const Task *d_tasks, int n, const OutPtrs &out,
                       int seq_bytes) {
    if (n <= 0) return POYB200_OK;
    if (ring_class(ctx, klass, affine)) {
        // pairs without gap bits take the instance without the block-diagonal state, which lists the batches it declines
        // for the full instance
        const int *list = nullptr, *count = nullptr;
        const size_t slot = std::max<size_t>(ctx->ring_slot_bytes, 128);
        if (ctx->cfg.allow_noeb && ctx->dcm.gap_open > 0) {
            CK(ctx->d_slow_list.reserve((size_t) n + 8));
            int *cnt = next_counter(ctx);
            CK(ring_launch(klass, bt, false, d_tasks, n, ctx->dcm, ctx->cur_pool, ctx->d_scratch.p, ctx->d_scratch.cap, slot, out,
                           ctx->sm_count, seq_bytes, next_counter(ctx), nullptr, nullptr, ctx->d_slow_list.p, cnt, ctx->stream));
            ctx->launches++;
            list = ctx->d_slow_list.p;
            count = cnt;
        }
        CK(ring_launch(klass, bt, true, d_tasks, n, ctx->dcm, ctx->cur_pool, ctx->d_scratch.p, ctx->d_scratch.cap, slot, out,
                       ctx->sm_count, seq_bytes, next_counter(ctx), list, count, nullptr, nullptr, ctx->stream));
        ctx->launches++;
        return POYB200_OK;
    }
    if (klass >= KLASS_LINROW_BASE) {
        cudaError_t e = lin_rows_launch(klass, bt, d_tasks, n, ctx->dcm, ctx->cur_pool, ctx->cur_dir, ctx->d_costs.p, ctx->sm_count,
                                        seq_bytes, ctx->custom_tail | (ctx->lin_natural << 1), next_counter(ctx), ctx->stream);
        ctx->launches++;
        CK(e);
        return POYB200_OK;
    }
    if (klass >= KLASS_LIN_BASE) {
        cudaError_t e = lin_stripe_launch(klass, bt, d_tasks, n, ctx->dcm, ctx->cur_pool, ctx->cur_dir, ctx->d_costs.p,
                                          ctx->sm_count, seq_bytes, ctx->custom_tail | (ctx->lin_natural << 1), next_counter(ctx), ctx->stream);
        ctx->launches++;
        CK(e);
        return POYB200_OK;
    }
    if (klass != KLASS_GENERIC) {
        // pairs without gap bits take aff_fast_kernel; the batches it declines are listed for aff_stripe_kernel
        const int *list = nullptr, *count = nullptr;
        if (affine && ctx->cfg.allow_fast && ctx->cfg.allow_noeb && ctx->dcm.gap_open > 0 && fast_has_shape(klass)) {
            CK(ctx->d_slow_list.reserve((size_t) n + 8));  // grows only (one entry per batch would do)
            const bool dir6 = bt && klass == 1 && ctx->cfg.dir6 && mixed_class(ctx, klass, affine);
            // shape (5, 8), 6-bit band: two pairs per lane group on 16-bit halves first; the batches it declines (gap bits,
            // spare diagonals, costs that could leave the 16-bit range) are listed for aff_fast_kernel
            const int *list0 = nullptr, *count0 = nullptr;
            if (dir6 && ctx->cfg.pair2 && n >= ctx->cfg.pair2_min_pairs && x2_usable(seq_bytes, ctx->x2_unit4, ctx->dcm.gap_open)) {
                CK(ctx->d_slow_list2.reserve((size_t) n + 8));
                int *cnt0 = next_counter(ctx);
                cudaError_t e0 = x2_launch(d_tasks, n, ctx->dcm, ctx->x2_unit4, ctx->cur_pool, ctx->cur_dir, ctx->d_costs.p, ctx->sm_count,
                                           seq_bytes, next_counter(ctx), ctx->d_slow_list2.p, cnt0, ctx->stream);
                ctx->launches++;
                CK(e0);
                list0 = ctx->d_slow_list2.p;
                count0 = cnt0;
            }
            int *cnt = next_counter(ctx);
            cudaError_t e = fast_launch(klass, bt, d_tasks, n, ctx->dcm, ctx->cur_pool, ctx->cur_dir, ctx->d_costs.p, ctx->sm_count,
                                        seq_bytes, next_counter(ctx), list0, count0, ctx->d_slow_list.p, cnt, dir6, ctx->stream);
            ctx->launches++;
            CK(e);
            list = ctx->d_slow_list.p;
            count = cnt;
            if (mixed_class(ctx, klass, affine)) {
                const size_t slot = std::max<size_t>(ctx->ring_slot_bytes, 128);
                CK(ring_launch(klass, bt, true, d_tasks, n, ctx->dcm, ctx->cur_pool, ctx->d_scratch.p, ctx->d_scratch.cap, slot, out,
                               ctx->sm_count, seq_bytes, next_counter(ctx), list, count, nullptr, nullptr, ctx->stream));
                ctx->launches++;
                return POYB200_OK;
            }
        }
        cudaError_t e = stripe_launch(klass, affine, bt, d_tasks, n, ctx->dcm, ctx->cur_pool, ctx->cur_dir, ctx->d_costs.p,
                                      ctx->sm_count, seq_bytes, ctx->cfg.allow_noeb, next_counter(ctx), list, count, ctx->stream);
        ctx->launches++;
        CK(e);
        return POYB200_OK;
    }
    const int warps_per_block = 4;
    int blocks = std::min((n + warps_per_block - 1) / warps_per_block, ctx->sm_count * 8);
    const size_t nwarps = (size_t) blocks * warps_per_block;
    if (affine) {
        CK(ctx->d_aff_state.reserve(nwarps * ctx->state_stride));
        CK(aff_generic_launch(bt, blocks, d_tasks, n, ctx->dcm, ctx->cur_pool, ctx->d_aff_state.p, ctx->state_stride, ctx->cur_dir,
                              ctx->d_costs.p, ctx->stream));
    } else {
        CK(ctx->d_lin_state.reserve(nwarps * ctx->state_stride));
        CK(lin_generic_launch(bt, blocks, d_tasks, n, ctx->dcm, ctx->cur_pool, ctx->d_lin_state.p, ctx->state_stride, ctx->cur_dir,
                              ctx->d_costs.p, ctx->stream));
    }
    ctx->launches++;
    return POYB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// planning
// ---------------------------------------------------------------------------------------------------------
// Persistent host workers for the planner: starting 32 std::threads costs more than a planning pass over 1 M pairs takes
// per thread, and a call makes four such passes.  Workers sleep on a condition variable between passes.
class WorkerPool {
public:
    static WorkerPool &get() {
        static WorkerPool p;
        return p;
    }
    // runs job(t) for t in [0, n): t = 0 on the calling thread, the others on workers
    void run(int n, const std::function<void(int)> &job) {
        std::unique_lock<std::mutex> call(call_m_);  // one parallel region at a time
        ensure(n - 1);
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = &job;
            want_ = n - 1;
            pending_ = n - 1;
            gen_++;
        }
        cv_.notify_all();
        job(0);
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [&] { return pending_ == 0; });
        job_ = nullptr;
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }

private:
    void ensure(int n) {
        while ((int) th_.size() < n) {
            const int id = (int) th_.size() + 1;
            std::unique_lock<std::mutex> g(m_);
            const uint64_t seen = gen_;
            g.unlock();
            th_.emplace_back([this, id, seen] { loop(id, seen); });
        }
    }
    void loop(int id, uint64_t seen) {
        std::unique_lock<std::mutex> g(m_);
        for (;;) {
            cv_.wait(g, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            if (id <= want_) {
                const std::function<void(int)> *job = job_;
                g.unlock();
                (*job)(id);
                g.lock();
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::mutex m_, call_m_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> th_;
    const std::function<void(int)> *job_ = nullptr;
    uint64_t gen_ = 0;
    int want_ = 0, pending_ = 0;
    bool stop_ = false;
};

// Runs fn(lo, hi, slot) over [0, n) on up to `ctx->host_threads` threads.
template <typename F>
static void parallel_for(int nthreads, size_t n, F fn, size_t grain = 32768) {
    if ((size_t) nthreads > n / grain) nthreads = (int) (n / grain);  // at least `grain` items per thread
    if (nthreads <= 1) {
        fn((size_t) 0, n, 0);
        return;
    }
    const std::function<void(int)> job = [&](int t) { fn(n * t / nthreads, n * (t + 1) / nthreads, t); };
    WorkerPool::get().run(nthreads, job);
}

static int plan(poyb200_ctx *ctx, int mode, const poyb200_batch *b) {
    const bool affine = (mode == MODE_COST_AFF || mode == MODE_ALIGN_AFF);
    const bool bt = (mode == MODE_ALIGN_2 || mode == MODE_ALIGN_AFF);
    if (!ctx->has_cm) return fail(ctx, POYB200_ENOCM, "no cost matrix loaded (poyb200_set_cm)");
    if (b->n_pairs < 0 || b->n_seqs < 0) return fail(ctx, POYB200_EINVAL, "negative count");
    if (b->n_pairs > 0 && (!b->pool || !b->seq_off || !b->seq_len || !b->pairs))
        return fail(ctx, POYB200_EINVAL, "NULL input array");
    if (!affine && b->n_pairs > 0 && !b->deltaw) return fail(ctx, POYB200_EINVAL, "linear entry points need deltaw[]");
    // Sequence.Align diverts Affine to affine_3 (src/sequence.ml:716-723, 851-858); No_Alignment matrices take
    // the legacy *_aff fill, which this library does not provide (closing note).
    if (!affine && ctx->hcm.cost_model_type != 0)
        return fail(ctx, POYB200_EMODEL, "linear entry point called with cost_model_type != 0");
    if (affine && ctx->hcm.lcm < 4)
        return fail(ctx, POYB200_EMODEL, "affine_3 indexes cost[(a&15) << lcm | (b&15)]: needs lcm >= 4");
    if (b->pool_bytes >= ((size_t) 1 << 32)) return fail(ctx, POYB200_EINVAL, "pool larger than 4 GiB");
    const int NT = ctx->host_threads;
    struct Part {
        int err = 0;
        long long maxcap = 16;
        int maxW = 1, max_stripe_len = 16;
        bool in_order_class = true;
        uint32_t klass_or = 0, klass_and = 0xffffffffu;
        size_t ring_slot = 0;
    };
    std::vector<Part> parts((size_t) std::max(1, NT));
    const bool view = ctx->view;
    const int64_t vlo = view ? ctx->view_lo : 0;
    if (!view)
        parallel_for(NT, (size_t) b->n_seqs, [&](size_t lo, size_t hi, int slot) {
            for (size_t s = lo; s < hi; s++) {
                if (b->seq_len[s] < 1 || b->seq_len[s] > POYB200_MAX_SEQ_LEN) parts[slot].err = POYB200_ESEQLEN;
                else if (b->seq_off[s] < 0 || (size_t) (b->seq_off[s] + b->seq_len[s]) > b->pool_bytes) parts[slot].err = POYB200_EINVAL;
            }
        });
    for (auto &pt : parts) {
        if (pt.err == POYB200_ESEQLEN) return fail(ctx, POYB200_ESEQLEN, "sequence empty (no leading gap) or longer than 16384");
        if (pt.err) return fail(ctx, POYB200_EINVAL, "sequence outside the pool");
    }
    if (!ctx->tasks.resize((size_t) b->n_pairs)) return fail(ctx, POYB200_ENOMEM, "pinned allocation of the task array failed");
    const DevCM dcm = ctx->dcm;
    const bool allow_stripe = (!ctx->cfg.force_generic);
    const bool allow_rows = ctx->cfg.allow_rows != 0;
    parallel_for(NT, (size_t) b->n_pairs, [&](size_t lo, size_t hi, int slot) {
        Part pt;  // thread-local: the parts[] entries share cache lines
        struct Commit {
            Part &src, &dst;
            ~Commit() { dst = src; }
        } commit{pt, parts[slot]};
        for (size_t p = lo; p < hi; p++) {
            const int a = b->pairs[2 * p], c = b->pairs[2 * p + 1];
            if (a < 0 || a >= b->n_seqs || c < 0 || c >= b->n_seqs) {
                pt.err = POYB200_EINVAL;
                return;
            }
            const int la = b->seq_len[a], lb = b->seq_len[c];
            if (view) {  // the shard's window of the pool: its own operands only
                const int64_t oa = b->seq_off[a] - vlo, oc = b->seq_off[c] - vlo;
                if (la < 1 || la > POYB200_MAX_SEQ_LEN || lb < 1 || lb > POYB200_MAX_SEQ_LEN || oa < 0 || oc < 0 ||
                    (size_t) (oa + la) > b->pool_bytes || (size_t) (oc + lb) > b->pool_bytes) {
                    pt.err = POYB200_EINVAL;
                    return;
                }
            }
            Task t{};
            // affine_3: shorter operand on the rows, ties keep a (src/algn.c:2595); linear: longer operand on the
            // rows, ties keep a (src/sequence.ml:709-714)
            const bool rows_b = affine ? (la > lb) : (la < lb);
            const int r = rows_b ? c : a, col = rows_b ? a : c;
            t.off_r = (uint32_t) (b->seq_off[r] - vlo);
            t.off_c = (uint32_t) (b->seq_off[col] - vlo);
            t.lr = b->seq_len[r];
            t.lc = b->seq_len[col];
            t.flags = rows_b ? TF_ROWS_ARE_B : 0;
            t.pair = (uint32_t) p;
            if (affine) {
                affine_band(t.lr - 1, t.lc - 1, t.dlo, t.dhi);
            } else {
                if (b->deltaw[p] < 0 || b->deltaw[p] > (1 << 20)) {
                    pt.err = POYB200_EINVAL;
                    return;
                }
                LinBand lb2 = linear_band(t.lr, t.lc, b->deltaw[p]);
                t.dlo = lb2.dlo;
                t.dhi = lb2.dhi;
                if (lb2.full) t.flags |= TF_FULL;
                const bool sw = b->swaped ? (b->swaped[p] != 0) : (la >= lb);  // src/sequence.ml:818
                if (sw) t.flags |= TF_SWAPED;
            }
            const int W = t.dhi - t.dlo + 1;
            choose_class(t, affine, bt, W, dcm, allow_stripe, allow_rows);
            if (t.klass != KLASS_GENERIC) pt.max_stripe_len = std::max(pt.max_stripe_len, std::max(t.lr, t.lc));
            if (bt && (ring_class(ctx, t.klass, affine) || mixed_class(ctx, t.klass, affine)))
                pt.ring_slot = std::max(pt.ring_slot, ((size_t) dir_bytes(t) + 127) & ~(size_t) 127);
            else pt.maxW = std::max(pt.maxW, W);
            // default mode, shape (5, 8): aff_fast_kernel writes five 6-bit codes per word for the traceback kernel (the pairs
            // it declines are walked by the ring kernel from its own scratch slots, sized above with 8-byte chunks)
            if (bt && t.klass == 1 && mixed_class(ctx, t.klass, affine) && ctx->cfg.dir6) {
                t.BL = 4;
                t.flags |= TF_DIR6;
            }
            pt.maxcap = std::max<long long>(pt.maxcap, (long long) la + lb + 2);
            pt.klass_or |= t.klass;
            pt.klass_and &= t.klass;
            ctx->tasks[p] = t;
        }
    });
    long long maxcap = 16;
    int maxW = 1, max_stripe_len = 16;
    uint32_t k_or = 0, k_and = 0xffffffffu;
    ctx->ring_slot_bytes = 0;
    for (auto &pt : parts) {
        ctx->ring_slot_bytes = std::max(ctx->ring_slot_bytes, pt.ring_slot);
        if (pt.err) return fail(ctx, POYB200_EINVAL, "pair index out of range, or deltaw negative / absurdly large");
        maxcap = std::max(maxcap, pt.maxcap);
        maxW = std::max(maxW, pt.maxW);
        max_stripe_len = std::max(max_stripe_len, pt.max_stripe_len);
        k_or |= pt.klass_or;
        k_and &= pt.klass_and;
    }
    if (bt && b->n_pairs > 0 && !ctx->device_store) {
        if ((b->want & (POYB200_WANT_MEDIAN | POYB200_WANT_CLOSEST)) && !b->median)
            return fail(ctx, POYB200_EINVAL, "WANT_MEDIAN / WANT_CLOSEST without the median buffer");
        if ((b->want & POYB200_WANT_MEDIAN) && (b->want & POYB200_WANT_CLOSEST))
            return fail(ctx, POYB200_EINVAL, "WANT_MEDIAN and WANT_CLOSEST share the median rows: ask for one");
        if ((b->want & POYB200_WANT_MEDIANWG) && !b->medianwg) return fail(ctx, POYB200_EINVAL, "WANT_MEDIANWG without buffer");
        if ((b->want & POYB200_WANT_ALIGNED) && (!b->aligned_a || !b->aligned_b))
            return fail(ctx, POYB200_EINVAL, "WANT_ALIGNED without buffers");
        if ((b->want & ~POYB200_WANT_BITSETS) && b->out_stride < maxcap)
            return fail(ctx, POYB200_EINVAL, "out_stride smaller than len a + len b + 2");
        if (b->want & POYB200_WANT_BITSETS) {
            if (!b->bits_a || !b->bits_b || !b->bits_wg) return fail(ctx, POYB200_EINVAL, "WANT_BITSETS without buffers");
            if ((b->bits_stride & 3) || b->bits_stride * 8 < maxcap)
                return fail(ctx, POYB200_EINVAL, "bits_stride must be a multiple of 4 and hold len a + len b + 2 bits");
        }
        if (b->want && !b->out_len) return fail(ctx, POYB200_EINVAL, "out_len is NULL");
    }
    ctx->dstride = (maxcap + 15) & ~15ll;
    ctx->bstride = (ctx->dstride / 8 + 3) & ~3ll;
    ctx->state_stride = maxW + 2;
    ctx->stripe_seq_bytes = (max_stripe_len + 15) & ~15;
    // Group by kernel class (stable: keeps the caller's order inside a class).  A batch of one class -- the usual
    // case -- stays in the caller's order, which lets results leave chunk by chunk while later chunks compute.
    // Chunks are contiguous ranges of the caller's pair list (so the results of a chunk are contiguous rows and can leave
    // while later chunks compute); inside a chunk the tasks are grouped by kernel class.
    ctx->in_order = true;
    const bool one_class = (b->n_pairs == 0) || (k_or == k_and);
    const size_t ntasks = ctx->tasks.size();
    // ring-class pairs keep their band in the ring kernels' per-warp scratch, not in the chunk's direction buffer
    auto band_bytes = [&](const Task &t) {
        return (bt && !ring_class(ctx, t.klass, affine)) ? (((size_t) dir_bytes(t) + 63) & ~(size_t) 63) : (size_t) 0;
    };
    // 1. cut by count; if some chunk's direction bands exceed the budget (long sequences) cut serially by bytes instead
    ctx->chunks.clear();
    const size_t CP = ctx->chunk_pairs, nch0 = (ntasks + CP - 1) / CP;
    ctx->chunks.resize(nch0);
    bool over = false;
    parallel_for(NT, nch0, [&](size_t lo, size_t hi, int) {
        for (size_t c = lo; c < hi; c++) {
            const size_t kb = c * CP, ke = std::min(ntasks, kb + CP);
            size_t tot = 0;
            for (size_t k = kb; k < ke; k++) tot += band_bytes(ctx->tasks[k]);
            ctx->chunks[c] = Chunk{kb, ke, tot};
            if (tot > ctx->dir_budget) over = true;  // benign race: only ever set to true
        }
    }, 1);
    if (over) {
        ctx->chunks.clear();
        size_t begin = 0, off = 0;
        for (size_t k = 0; k < ntasks; k++) {
            const size_t bytes = band_bytes(ctx->tasks[k]);
            if (k > begin && (off + bytes > ctx->dir_budget || k - begin >= CP)) {
                ctx->chunks.push_back(Chunk{begin, k, off});
                begin = k;
                off = 0;
            }
            off += bytes;
        }
        if (ntasks > begin) ctx->chunks.push_back(Chunk{begin, ntasks, off});
    }
    // Taper: the results of the last chunk cannot overlap any compute, and where the download runs about as fast as the
    // kernels (all four sequences: 6.4 ms against 7 ms per 131 072 pairs) nothing of it can be hidden later either.  The last
    // ~1.5 chunks are therefore re-cut into pieces that shrink towards the end (..., 65 536, 32 768, 16 384, 16 384), so that
    // what is still to be downloaded when the last kernel ends is one small piece.
    if (bt && !over && ctx->chunks.size() >= 2) {  // (chunks cut by the band budget are left alone)
        size_t begin = ctx->chunks.back().begin;
        ctx->chunks.pop_back();
        while (ctx->chunks.size() >= 2 && ntasks - begin < CP + CP / 2) {
            begin = ctx->chunks.back().begin;
            ctx->chunks.pop_back();
        }
        std::vector<Chunk> tail;
        size_t end = ntasks;
        const size_t piece[4] = {16384, 16384, 32768, 65536};
        for (int q = 0; q < 4 && end - begin >= 2 * piece[q]; q++) {
            tail.push_back(Chunk{end - piece[q], end, 0});
            end -= piece[q];
        }
        while (end > begin) {
            const size_t len = std::min(CP, end - begin);
            tail.push_back(Chunk{end - len, end, 0});
            end -= len;
        }
        for (size_t k = tail.size(); k-- > 0;) ctx->chunks.push_back(tail[k]);
    }
    // ... and the first chunk likewise (1/8, 1/8, 1/4, 1/2): nothing can be downloaded before the first chunk is traced back,
    // and where the download is the longer leg (all four sequences) every millisecond it starts earlier is a millisecond
    // off the call
    if (bt && ctx->chunks.size() >= 2) {
        const Chunk first = ctx->chunks.front();
        const size_t len = first.end - first.begin;
        if (len >= 8192) {
            const size_t cut[5] = {0, len / 8, len / 4, len / 2, len};
            std::vector<Chunk> head;
            for (int q = 0; q < 4; q++) head.push_back(Chunk{first.begin + cut[q], first.begin + cut[q + 1], 0});
            ctx->chunks.erase(ctx->chunks.begin());
            ctx->chunks.insert(ctx->chunks.begin(), head.begin(), head.end());
        }
    }
    // 2. per chunk, in parallel: stable grouping by class (counting sort into the second array) and band layout
    if (!one_class && !ctx->tasks_tmp.resize(ntasks)) return fail(ctx, POYB200_ENOMEM, "pinned allocation of the task array failed");
    parallel_for(NT, ctx->chunks.size(), [&](size_t lo, size_t hi, int) {
        constexpr int NK = 64;
        for (size_t c = lo; c < hi; c++) {
            Chunk &ch = ctx->chunks[c];
            Task *src = ctx->tasks.data();
            if (!one_class) {
                size_t hist[NK] = {0};
                for (size_t k = ch.begin; k < ch.end; k++) hist[src[k].klass & (NK - 1)]++;
                size_t run = ch.begin;
                for (int q = 0; q < NK; q++) {
                    const size_t v = hist[q];
                    hist[q] = run;
                    run += v;
                }
                Task *dst = ctx->tasks_tmp.data();
                for (size_t k = ch.begin; k < ch.end; k++) dst[hist[src[k].klass & (NK - 1)]++] = src[k];
                src = dst;
            }
            size_t off = 0;
            for (size_t k = ch.begin; k < ch.end; k++) {
                src[k].dir_off = off;
                off += band_bytes(src[k]);
            }
            ch.dir_bytes = off;
        }
    }, 1);
    if (!one_class) std::swap(ctx->tasks, ctx->tasks_tmp);
    return POYB200_OK;
}

static int stage_impl(poyb200_ctx *ctx, int mode, const poyb200_batch *b, bool upload) {
    if (!ctx || !b) return POYB200_EINVAL;
    if (mode < 0 || mode > 3) return fail(ctx, POYB200_EINVAL, "bad mode");
    cudaSetDevice(ctx->device);
    ctx->staged = false;
    // small calls (a dependency level of a tree build): the ring kernels walk each pair from shared memory right after its
    // fill instead of a second kernel chasing the band through L2 -- lower latency, lower throughput
    ctx->ring_mode = ctx->cfg.use_ring;
    if (ctx->cfg.use_ring == 2 && ctx->cfg.small_ring_pairs > 0 && b->n_pairs <= ctx->cfg.small_ring_pairs) ctx->ring_mode = 1;
    int rc = plan(ctx, mode, b);
    if (rc) return rc;
    ctx->mode = mode;
    ctx->hb = *b;
    const bool bt = (mode == MODE_ALIGN_2 || mode == MODE_ALIGN_AFF);
    const size_t n = ctx->tasks.size();
    if (!ctx->device_store) {
        CK(ctx->d_pool.reserve(b->pool_bytes + 64));
        ctx->cur_pool = ctx->d_pool.p;
    }
    CK(ctx->d_tasks.reserve(n + 1));
    CK(ctx->d_costs.reserve(n + 1));
    size_t maxdir = 16;
    for (auto &c : ctx->chunks) maxdir = std::max(maxdir, c.dir_bytes);
    if (bt && ctx->ring_slot_bytes) {
        // as many slots per warp as the ring kernels can use, within the direction budget (they run with fewer if need be)
        // never more than one slot group per batch of the call: small calls (a tree level) and long pairs stay cheap
        const size_t full = ring_scratch_bytes(ctx->sm_count, ctx->ring_slot_bytes);
        const size_t by_pairs = ((size_t) n / 4 + 1) * 32 * ctx->ring_slot_bytes;
        CK(ctx->d_scratch.reserve(std::min(std::min(full, by_pairs), ctx->dir_budget)));
    }
    if (bt) {
        CK(ctx->d_dir.reserve(maxdir));
        if (ctx->cfg.overlap_traceback && ctx->chunks.size() >= 2) CK(ctx->d_dir2.reserve(maxdir));
        if (ctx->cfg.overlap_traceback && ctx->chunks.size() >= 3 && ctx->cfg.dir_buffers >= 3) CK(ctx->d_dir3.reserve(maxdir));
        CK(ctx->d_outlen.reserve(4 * n + 4));
        const size_t ob = n * (size_t) ctx->dstride + 16;
        if (b->want & (POYB200_WANT_MEDIAN | POYB200_WANT_CLOSEST)) CK(ctx->d_out[0].reserve(ob));
        if (b->want & POYB200_WANT_MEDIANWG) CK(ctx->d_out[1].reserve(ob));
        if (b->want & POYB200_WANT_ALIGNED) {
            CK(ctx->d_out[2].reserve(ob));
            CK(ctx->d_out[3].reserve(ob));
        }
        if (b->want & POYB200_WANT_BITSETS)
            for (int k = 0; k < 3; k++) CK(ctx->d_bits[k].reserve(n * (size_t) ctx->bstride + 16));
    }
    if (n && upload) {
        CK(cudaMemcpyAsync(ctx->d_pool.p, b->pool, b->pool_bytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_tasks.p, ctx->tasks.data(), n * sizeof(Task), cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->staged = true;
    return POYB200_OK;
}

extern "C" int poyb200_stage(poyb200_ctx *ctx, int mode, const poyb200_batch *b) { return stage_impl(ctx, mode, b, true); }
int poyb200_stage_internal(poyb200_ctx *ctx, int mode, const poyb200_batch *b, bool upload) { return stage_impl(ctx, mode, b, upload); }
int poyb200_run_staged(poyb200_ctx *ctx) { return poyb200_run(ctx); }

// Fill of chunk ci on the compute stream, its traceback on the traceback stream.  With two direction buffers the
// (latency-bound, 128..512 walkers per SM) traceback of chunk ci runs underneath the (ALU-bound) fill of chunk ci+1.
static int run_chunk(poyb200_ctx *ctx, size_t ci) {
    const int mode = ctx->mode;
    const bool affine = (mode == MODE_COST_AFF || mode == MODE_ALIGN_AFF);
    const bool bt = (mode == MODE_ALIGN_2 || mode == MODE_ALIGN_AFF);
    const Chunk &ch = ctx->chunks[ci];
    const bool two = bt && ctx->cfg.overlap_traceback && ctx->chunks.size() >= 2;
    cudaStream_t s_tb = two ? ctx->s_tb : ctx->stream;
    const int nbuf = !two ? 1 : (ctx->cfg.dir_buffers >= 3 && ctx->chunks.size() >= 3) ? 3 : 2;
    uint8_t *const bufs[3] = {ctx->d_dir.p, ctx->d_dir2.p, ctx->d_dir3.p};
    ctx->cur_dir = bufs[ci % nbuf];
    OutPtrs out{ctx->d_costs.p, ctx->d_out[0].p, ctx->d_out[1].p, ctx->d_out[2].p, ctx->d_out[3].p,
                ctx->d_outlen.p, ctx->dstride, ctx->hb.want, ctx->d_bits[0].p, ctx->d_bits[1].p, ctx->d_bits[2].p, ctx->bstride,
                (bt && affine && ctx->ring_mode == 2) ? ctx->d_walked.p : nullptr};
    if (two && ci >= (size_t) nbuf) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_tb[ci - nbuf], 0));  // the buffer is free again
    if (ctx->cfg.timing) CK(cudaEventRecord(ctx->chunk_ev[4 * ci], ctx->stream));
    // one fill launch per kernel class present in the chunk; the classes whose fill kernel does not walk its own pairs
    // get a traceback launch over their task range afterwards
    struct Group { size_t begin, end; };
    std::vector<Group> walk_groups;
    size_t k = ch.begin;
    while (k < ch.end) {
        size_t e = k;
        const uint32_t klass = ctx->tasks[k].klass;
        while (e < ch.end && ctx->tasks[e].klass == klass) e++;
        // shared memory per staged operand: the longest sequence of THIS class group (a few long pairs of another class must
        // not cost the short ones their occupancy)
        int longest = 16;
        for (size_t q = k; q < e; q++) longest = std::max(longest, std::max(ctx->tasks[q].lr, ctx->tasks[q].lc));
        int rc = launch_fill(ctx, klass, affine, bt, ctx->d_tasks.p + k, (int) (e - k), out, (longest + 15) & ~15);
        if (rc) return rc;
        if (bt && !ring_class(ctx, klass, affine)) {
            if (!walk_groups.empty() && walk_groups.back().end == k) walk_groups.back().end = e;
            else walk_groups.push_back(Group{k, e});
        }
        k = e;
    }
    if (ctx->cfg.timing) CK(cudaEventRecord(ctx->chunk_ev[4 * ci + 1], ctx->stream));
    if (bt) {
        if (two) {
            CK(cudaEventRecord(ctx->ev_fill[ci], ctx->stream));
            CK(cudaStreamWaitEvent(s_tb, ctx->ev_fill[ci], 0));
        }
        if (ctx->cfg.timing) CK(cudaEventRecord(ctx->chunk_ev[4 * ci + 2], s_tb));
        for (const Group &g : walk_groups) {
            const int nt = (int) (g.end - g.begin);
            // persistent grid: the walkers' working set (direction line + 2 sequence lines + 4 output lines each)
            // has to stay inside L1 / L2, so only traceback_threads_per_sm of them run per SM at a time
            const int tb_block = ctx->cfg.traceback_block;
            const int wpb = tb_block / 32;
            const int max_blocks = ctx->sm_count * std::max(1, ctx->cfg.traceback_threads_per_sm / tb_block);
            // walkers per warp: 32 when the batch fills the grid, fewer (down to 1) when it does not
            int wpw = (nt + max_blocks * wpb - 1) / (max_blocks * wpb);
            wpw = std::min(32, std::max(1, wpw));
            const int blocks = std::min((nt + wpb * wpw - 1) / (wpb * wpw), max_blocks);
            CK(traceback_launch(affine, blocks, tb_block, ctx->d_tasks.p + g.begin, nt, ctx->dcm, ctx->cur_pool, ctx->cur_dir, out,
                                next_counter(ctx), wpw, s_tb));
            ctx->launches++;
        }
        if (ctx->cfg.timing) CK(cudaEventRecord(ctx->chunk_ev[4 * ci + 3], s_tb));
    }
    CK(cudaEventRecord(ctx->ev_tb[ci], s_tb));  // chunk ci is complete
    return POYB200_OK;
}

static int prepare_events(poyb200_ctx *ctx) {
    while (ctx->cfg.timing && ctx->chunk_ev.size() < 4 * ctx->chunks.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        ctx->chunk_ev.push_back(e);
    }
    while (ctx->ev_fill.size() < ctx->chunks.size()) {
        cudaEvent_t e, f;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&f, cudaEventDisableTiming));
        ctx->ev_fill.push_back(e);
        ctx->ev_tb.push_back(f);
    }
    ctx->timed_chunks = ctx->cfg.timing ? ctx->chunks.size() : 0;
    return POYB200_OK;
}

// Makes the compute stream wait for the tracebacks still running on the traceback stream.
static int join_streams(poyb200_ctx *ctx) {
    const size_t nch = ctx->chunks.size();
    for (size_t ci = (nch >= 2 ? nch - 2 : 0); ci < nch; ci++) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_tb[ci], 0));
    return POYB200_OK;
}

extern "C" int poyb200_run(poyb200_ctx *ctx) {
    if (!ctx) return POYB200_EINVAL;
    if (!ctx->staged) return fail(ctx, POYB200_EINVAL, "poyb200_run without poyb200_stage");
    cudaSetDevice(ctx->device);
    int rc = prepare_events(ctx);
    if (rc) return rc;
    rc = reset_counters(ctx);
    if (rc) return rc;
    for (size_t ci = 0; ci < ctx->chunks.size(); ci++) {
        rc = run_chunk(ctx, ci);
        if (rc) return rc;
    }
    return join_streams(ctx);
}

extern "C" int poyb200_sync(poyb200_ctx *ctx) {
    if (!ctx) return POYB200_EINVAL;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    return POYB200_OK;
}

extern "C" int poyb200_last_run_ms(poyb200_ctx *ctx, float ms[2]) {
    if (!ctx || !ms) return POYB200_EINVAL;
    cudaSetDevice(ctx->device);
    ms[0] = ms[1] = 0.f;
    for (size_t c = 0; c < ctx->timed_chunks; c++) {
        float a = 0.f, b = 0.f;
        const bool bt = (ctx->mode == MODE_ALIGN_2 || ctx->mode == MODE_ALIGN_AFF);
        CK(cudaEventSynchronize(ctx->chunk_ev[4 * c + 1]));
        CK(cudaEventElapsedTime(&a, ctx->chunk_ev[4 * c], ctx->chunk_ev[4 * c + 1]));
        if (bt) {
            CK(cudaEventSynchronize(ctx->chunk_ev[4 * c + 3]));
            CK(cudaEventElapsedTime(&b, ctx->chunk_ev[4 * c + 2], ctx->chunk_ev[4 * c + 3]));
        }
        ms[0] += a;
        ms[1] += b;
        if (ctx->cfg.trace >= 3) {
            // device timeline of the chunk relative to the first fill: [fill start, fill end] [traceback start, traceback end]
            float f0 = 0.f, f1 = 0.f, t0 = 0.f, t1 = 0.f;
            cudaEventElapsedTime(&f0, ctx->chunk_ev[0], ctx->chunk_ev[4 * c]);
            cudaEventElapsedTime(&f1, ctx->chunk_ev[0], ctx->chunk_ev[4 * c + 1]);
            if (bt) {
                cudaEventElapsedTime(&t0, ctx->chunk_ev[0], ctx->chunk_ev[4 * c + 2]);
                cudaEventElapsedTime(&t1, ctx->chunk_ev[0], ctx->chunk_ev[4 * c + 3]);
            }
            fprintf(stderr, "[poyb200]   chunk %zu: fill %.2f..%.2f  traceback %.2f..%.2f ms\n", c, f0, f1, t0, t1);
        }
    }
    return POYB200_OK;
}

// D2H of the results of pairs [lo, hi) on stream st (rows of one pair range are contiguous on both sides).
// lens_known: out_len of the range is already on the host; then only the columns that hold data are copied (rows are
// right aligned, typical alignments are about half as long as the worst case the rows are sized for).
static int fetch_range(poyb200_ctx *ctx, size_t lo, size_t hi, cudaStream_t st, bool lens_known = false) {
    const poyb200_batch &b = ctx->hb;
    const bool bt = (ctx->mode == MODE_ALIGN_2 || ctx->mode == MODE_ALIGN_AFF);
    const size_t n = hi - lo;
    if (n == 0) return POYB200_OK;
    if (b.cost) CK(cudaMemcpyAsync(b.cost + lo, ctx->d_costs.p + lo, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (bt && b.want) {
        if (!lens_known)
            CK(cudaMemcpyAsync(b.out_len + 4 * lo, ctx->d_outlen.p + 4 * lo, 4 * n * sizeof(int), cudaMemcpyDeviceToHost, st));
        uint8_t *dst[4] = {b.median, b.medianwg, b.aligned_a, b.aligned_b};
        const uint32_t need[4] = {POYB200_WANT_MEDIAN | POYB200_WANT_CLOSEST, POYB200_WANT_MEDIANWG, POYB200_WANT_ALIGNED, POYB200_WANT_ALIGNED};
        // right-aligned device rows -> right-aligned caller rows
        const size_t wfull = (size_t) std::min<long long>(ctx->dstride, b.out_stride);
        int wmax[4] = {0, 0, 0, 0};
        if (lens_known)
            for (size_t p = lo; p < hi; p++)
                for (int k = 0; k < 4; k++) wmax[k] = std::max(wmax[k], b.out_len[4 * p + k]);
        for (int k = 0; k < 4; k++) {
            if (!(b.want & need[k])) continue;
            uint8_t *d = dst[k] + lo * (size_t) b.out_stride;
            const uint8_t *src = ctx->d_out[k].p + lo * (size_t) ctx->dstride;
            const size_t w = lens_known ? std::min(wfull, ((size_t) wmax[k] + 31) & ~(size_t) 31) : wfull;
            if (w == 0) continue;
            if (ctx->dstride == b.out_stride && w == wfull)
                CK(cudaMemcpyAsync(d, src, n * (size_t) ctx->dstride, cudaMemcpyDeviceToHost, st));
            else
                CK(cudaMemcpy2DAsync(d + (b.out_stride - w), (size_t) b.out_stride, src + (ctx->dstride - w),
                                     (size_t) ctx->dstride, w, n, cudaMemcpyDeviceToHost, st));
        }
        if (b.want & POYB200_WANT_BITSETS) {
            // right-aligned bit rows: the trailing bytes that hold the longest alignment of the range
            uint8_t *bdst[3] = {b.bits_a, b.bits_b, b.bits_wg};
            const size_t bfull = (size_t) std::min<long long>(ctx->bstride, b.bits_stride);
            const size_t bw = lens_known ? std::min(bfull, ((size_t) (wmax[2] + 7) / 8 + 31) & ~(size_t) 31) : bfull;
            for (int k = 0; k < 3 && bw; k++) {
                uint8_t *d = bdst[k] + lo * (size_t) b.bits_stride;
                const uint8_t *src = ctx->d_bits[k].p + lo * (size_t) ctx->bstride;
                CK(cudaMemcpy2DAsync(d + (b.bits_stride - bw), (size_t) b.bits_stride, src + (ctx->bstride - bw),
                                     (size_t) ctx->bstride, bw, n, cudaMemcpyDeviceToHost, st));
            }
        }
    }
    return POYB200_OK;
}

extern "C" int poyb200_fetch(poyb200_ctx *ctx) {
    if (!ctx) return POYB200_EINVAL;
    if (!ctx->staged) return fail(ctx, POYB200_EINVAL, "poyb200_fetch without poyb200_stage");
    cudaSetDevice(ctx->device);
    int rc = fetch_range(ctx, 0, ctx->tasks.size(), ctx->stream);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return POYB200_OK;
}

// One-shot call, pipelined over chunks on three streams:
//   s_in    uploads the task array and then the pool in slices;
//   stream  runs fill + traceback of chunk k as soon as the slices its pairs reference have arrived;
//   s_out   downloads the results of chunk k while chunk k+1 computes (needs the caller's pair order, i.e. one
//           kernel class; otherwise one download at the end).
static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

static int one_shot_impl(poyb200_ctx *ctx, int mode, const poyb200_batch *b) {
    const double t0 = now_ms();
    if (!ctx || !b) return POYB200_EINVAL;
    const bool trace = ctx->cfg.trace >= 1;
    // The pool does not depend on the plan: its first slices travel while the (multi-threaded, ~5.5 ns per pair) planner runs,
    // so the first chunk finds its operands on the device when the plan is ready.  Only as many slices as the planner's
    // run time covers (~290 bytes per pair at the link's 52 GB/s) go early: copies queue FIFO on the copy engine and the task
    // array, which the plan produces, must not wait behind the whole pool (measured: +21 ms on 1 M pairs when it did).
    constexpr size_t SLICE = (size_t) 32 << 20;
    const size_t nslices = (b->pool && b->n_pairs > 0) ? (b->pool_bytes + SLICE - 1) / SLICE : 0;
    const size_t nearly = std::min(nslices, ((size_t) std::max(b->n_pairs, 0) * 290 + SLICE - 1) / SLICE);
    cudaSetDevice(ctx->device);
    auto upload_slices = [&](size_t k0, size_t k1) -> int {
        for (size_t k = k0; k < k1; k++) {
            const size_t lo = k * SLICE, len = std::min(SLICE, b->pool_bytes - lo);
            CK(cudaMemcpyAsync(ctx->d_pool.p + lo, b->pool + lo, len, cudaMemcpyHostToDevice, ctx->s_in));
            CK(cudaEventRecord(ctx->ev_pool[k], ctx->s_in));
        }
        return POYB200_OK;
    };
    if (nslices) {
        CK(ctx->d_pool.reserve(b->pool_bytes + 64));
        ctx->cur_pool = ctx->d_pool.p;
        while (ctx->ev_pool.size() < nslices) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->ev_pool.push_back(e);
        }
        if (int urc = upload_slices(0, nearly)) return urc;
    }
    int rc = stage_impl(ctx, mode, b, false);
    if (rc) return rc;
    const double t1 = now_ms();
    const size_t n = ctx->tasks.size(), nch = ctx->chunks.size();
    if (n == 0) {
        cudaStreamSynchronize(ctx->s_in);
        return POYB200_OK;
    }
    rc = prepare_events(ctx);
    if (rc) return rc;
    rc = reset_counters(ctx);
    if (rc) return rc;
    while (ctx->ev_done.size() < nch + nslices + 1) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_done.push_back(e);
    }
    cudaEvent_t *ev_slice = ctx->ev_pool.data();
    // last pool byte each chunk needs (prefix maximum: slices arrive in order)
    std::vector<size_t> need(nch, 0);
    parallel_for(ctx->host_threads, nch, [&](size_t lo, size_t hi, int) {
        for (size_t ci = lo; ci < hi; ci++) {
            size_t m = 0;
            for (size_t k = ctx->chunks[ci].begin; k < ctx->chunks[ci].end; k++) {
                const Task &t = ctx->tasks[k];
                m = std::max(m, std::max((size_t) t.off_r + t.lr, (size_t) t.off_c + t.lc));
            }
            need[ci] = m;
        }
    }, 1);
    for (size_t ci = 1; ci < nch; ci++) need[ci] = std::max(need[ci], need[ci - 1]);
    // the task array follows the early slices, the rest of the pool follows the task array
    CK(cudaMemcpyAsync(ctx->d_tasks.p, ctx->tasks.data(), n * sizeof(Task), cudaMemcpyHostToDevice, ctx->s_in));
    CK(cudaEventRecord(ctx->ev_done[nch + nslices], ctx->s_in));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_done[nch + nslices], 0));
    if (int urc = upload_slices(nearly, nslices)) return urc;
    size_t waited = 0;  // slices the compute stream already waits for
    for (size_t ci = 0; ci < nch; ci++) {
        const size_t upto = std::min(nslices, (need[ci] + SLICE - 1) / SLICE);
        if (upto > waited) {
            CK(cudaStreamWaitEvent(ctx->stream, ev_slice[upto - 1], 0));
            waited = upto;
        }
        rc = run_chunk(ctx, ci);
        if (rc) return rc;
    }
    rc = join_streams(ctx);
    if (rc) return rc;
    if (ctx->in_order) {
        // All compute is enqueued; now follow it chunk by chunk: fetch the lengths of a finished chunk on their own
        // stream, then copy only the columns in use while later chunks compute.
        const bool bt = (mode == MODE_ALIGN_2 || mode == MODE_ALIGN_AFF) && b->want;
        for (size_t ci = 0; ci < nch; ci++) {
            const Chunk &ch = ctx->chunks[ci];
            if (bt) {
                // the lengths travel on a stream of their own: behind the pool slices on s_in they would hold the download of
                // the first chunks back until the whole pool is up (measured: 21 of 77 ms per 1 M pairs)
                CK(cudaStreamWaitEvent(ctx->s_len, ctx->ev_tb[ci], 0));
                CK(cudaMemcpyAsync(b->out_len + 4 * ch.begin, ctx->d_outlen.p + 4 * ch.begin, 4 * (ch.end - ch.begin) * sizeof(int),
                                   cudaMemcpyDeviceToHost, ctx->s_len));
                CK(cudaStreamSynchronize(ctx->s_len));
                if (trace) fprintf(stderr, "[poyb200]   chunk %zu (%zu pairs) traced back at +%.1f ms\n", ci, ch.end - ch.begin, now_ms() - t0);
            } else {
                CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_tb[ci], 0));
            }
            rc = fetch_range(ctx, ch.begin, ch.end, ctx->s_out, bt);
            if (rc) return rc;
            if (ctx->cfg.trace >= 2) {
                cudaStreamSynchronize(ctx->s_out);
                fprintf(stderr, "[poyb200]   chunk %zu downloaded at +%.1f ms\n", ci, now_ms() - t0);
            }
        }
    }
    if (!ctx->in_order) {
        rc = fetch_range(ctx, 0, n, ctx->stream);
        if (rc) return rc;
    }
    const double t2 = now_ms();
    CK(cudaStreamSynchronize(ctx->s_in));
    const double t3 = now_ms();
    CK(cudaStreamSynchronize(ctx->stream));
    const double t4 = now_ms();
    CK(cudaStreamSynchronize(ctx->s_out));
    const double t5 = now_ms();
    if (trace)
        fprintf(stderr, "[poyb200] plan+reserve %.1f ms, enqueue %.1f ms, wait h2d %.1f, wait compute %.1f, wait d2h %.1f, total %.1f ms\n",
                t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0);
    return POYB200_OK;
}

// Whatever went wrong, no copy from the caller's pool or into the caller's buffers is in flight when the call returns:
// the caller frees them next.
static int one_shot(poyb200_ctx *ctx, int mode, const poyb200_batch *b) {
    const int rc = one_shot_impl(ctx, mode, b);
    if (rc != POYB200_OK && ctx) {
        const std::string keep = ctx->err;
        for (cudaStream_t st : {ctx->s_in, ctx->stream, ctx->s_tb, ctx->s_out, ctx->s_len})
            if (st) cudaStreamSynchronize(st);
        cudaGetLastError();
        ctx->err = keep;
    }
    return rc;
}

// A shard of a larger batch (multi.cu): `b->pool` points view_lo bytes into the caller's pool.
int poyb200_one_shot_view(poyb200_ctx *ctx, int mode, const poyb200_batch *b, int64_t view_lo) {
    if (!ctx) return POYB200_EINVAL;
    ctx->view = true;
    ctx->view_lo = view_lo;
    const int rc = one_shot(ctx, mode, b);
    ctx->view = false;
    ctx->view_lo = 0;
    return rc;
}

extern "C" int poyb200_batch_cost_2(poyb200_ctx *ctx, const poyb200_batch *b) { return one_shot(ctx, MODE_COST_2, b); }
extern "C" int poyb200_batch_align_2(poyb200_ctx *ctx, const poyb200_batch *b) { return one_shot(ctx, MODE_ALIGN_2, b); }
extern "C" int poyb200_batch_cost_affine_3(poyb200_ctx *ctx, const poyb200_batch *b) { return one_shot(ctx, MODE_COST_AFF, b); }
extern "C" int poyb200_batch_align_affine_3(poyb200_ctx *ctx, const poyb200_batch *b) { return one_shot(ctx, MODE_ALIGN_AFF, b); }

extern "C" int poyb200_batch_median_2(poyb200_ctx *ctx, int which, const uint8_t *a, const uint8_t *b, int64_t in_stride,
                                      const int32_t *len, int32_t n, uint8_t *out, int64_t out_stride, int32_t *out_len) {
    if (!ctx) return POYB200_EINVAL;
    if (!ctx->has_cm) return fail(ctx, POYB200_ENOCM, "no cost matrix loaded (poyb200_set_cm)");
    if (which < 0 || which > 2 || n < 0) return fail(ctx, POYB200_EINVAL, "bad argument");
    if (n == 0) return POYB200_OK;
    if (!a || !b || !len || !out || !out_len) return fail(ctx, POYB200_EINVAL, "NULL array");
    for (int p = 0; p < n; p++)
        if (len[p] < 0 || len[p] > in_stride || len[p] + 1 > out_stride)
            return fail(ctx, POYB200_EINVAL, "row length does not fit the strides");
    cudaSetDevice(ctx->device);
    const size_t ib = (size_t) n * in_stride, ob = (size_t) n * out_stride;
    CK(ctx->d_out[2].reserve(ib + 16));
    CK(ctx->d_out[3].reserve(ib + 16));
    CK(ctx->d_out[0].reserve(ob + 16));
    CK(ctx->d_outlen.reserve(2 * (size_t) n + 4));
    CK(cudaMemcpyAsync(ctx->d_out[2].p, a, ib, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_out[3].p, b, ib, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_outlen.p + n, len, (size_t) n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(median_2_launch(which, ctx->dcm, ctx->d_out[2].p, ctx->d_out[3].p, in_stride, ctx->d_outlen.p + n, n, ctx->d_out[0].p, out_stride,
                       ctx->d_outlen.p, ctx->stream));
    ctx->launches++;
    CK(cudaMemcpyAsync(out, ctx->d_out[0].p, ob, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_len, ctx->d_outlen.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged = false;
    return POYB200_OK;
}

extern "C" int poyb200_batch_worst_2(poyb200_ctx *ctx, int which, const uint8_t *a, const uint8_t *b, int64_t in_stride,
                                     const int32_t *len, int32_t n, int32_t *out) {
    if (!ctx) return POYB200_EINVAL;
    if (!ctx->has_cm) return fail(ctx, POYB200_ENOCM, "no cost matrix loaded (poyb200_set_cm)");
    if (which < 0 || which > 1 || n < 0) return fail(ctx, POYB200_EINVAL, "bad argument");
    if (which == 0 && !ctx->dcm.worst) return fail(ctx, POYB200_EINVAL, "poyb200_batch_worst_2: the loaded cost matrix has no worst table");
    if (n == 0) return POYB200_OK;
    if (!a || !b || !len || !out) return fail(ctx, POYB200_EINVAL, "NULL array");
    for (int p = 0; p < n; p++)
        if (len[p] < 0 || len[p] > in_stride) return fail(ctx, POYB200_EINVAL, "row length does not fit the stride");
    cudaSetDevice(ctx->device);
    const size_t ib = (size_t) n * in_stride;
    CK(ctx->d_out[2].reserve(ib + 16));
    CK(ctx->d_out[3].reserve(ib + 16));
    CK(ctx->d_outlen.reserve(2 * (size_t) n + 4));
    CK(cudaMemcpyAsync(ctx->d_out[2].p, a, ib, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_out[3].p, b, ib, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_outlen.p + n, len, (size_t) n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(calc_aligned_2_launch(which == 0 ? ctx->dcm.worst : ctx->dcm.cost, ctx->dcm, ctx->d_out[2].p, ctx->d_out[3].p, in_stride,
                             ctx->d_outlen.p + n, n, ctx->d_outlen.p, ctx->stream));
    ctx->launches++;
    CK(cudaMemcpyAsync(out, ctx->d_outlen.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged = false;
    return POYB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// INT32 roofline probe
// ---------------------------------------------------------------------------------------------------------
template <int KIND>
static int peak_one(poyb200_ctx *ctx, int *d_out, double ops_per_step, double *gops) {
    const int blocks = ctx->sm_count * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
        CK(int32_peak_launch(KIND, blocks, threads, d_out, rep + 1, ctx->stream));
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        CK(cudaEventSynchronize(ctx->ev[1]));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        if (rep > 0 && ms < best) best = ms;
        ctx->launches++;
    }
    const double ops = (double) blocks * threads * PEAK_ITERS * PEAK_CHAINS * ops_per_step;
    *gops = ops / (best * 1e-3) * 1e-9;
    return POYB200_OK;
}

extern "C" int poyb200_int32_peak(poyb200_ctx *ctx, double *gops_add, double *gops_minmax, double *gops_mix) {
    if (!ctx || !gops_add || !gops_minmax || !gops_mix) return POYB200_EINVAL;
    cudaSetDevice(ctx->device);
    CK(ctx->d_costs.reserve(64));
    int rc = peak_one<0>(ctx, ctx->d_costs.p, 2.0, gops_add);
    if (rc) return rc;
    rc = peak_one<1>(ctx, ctx->d_costs.p, 3.0, gops_minmax);
    if (rc) return rc;
    return peak_one<2>(ctx, ctx->d_costs.p, 3.0, gops_mix);
}

// ---------------------------------------------------------------------------------------------------------
// three sequences
// ---------------------------------------------------------------------------------------------------------
extern "C" int64_t poyb200_cells_3d(int32_t l1, int32_t l2, int32_t l3) { return (int64_t) l1 * l2 * l3; }

extern "C" int poyb200_set_cm_3d(poyb200_ctx *ctx, const poyb200_cm3 *cm) {
    if (!ctx || !cm || !cm->cost || !cm->median) return ctx ? fail(ctx, POYB200_EINVAL, "poyb200_set_cm_3d: NULL table") : POYB200_EINVAL;
    if (cm->lcm < 1 || cm->lcm > 6) return fail(ctx, POYB200_EINVAL, "poyb200_set_cm_3d: lcm out of range");
    cudaSetDevice(ctx->device);
    const size_t n = (size_t) 1 << (3 * cm->lcm);
    CK(ctx->d_cost3.reserve(n));
    CK(ctx->d_median3.reserve(n));
    CK(cudaMemcpyAsync(ctx->d_cost3.p, cm->cost, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_median3.p, cm->median, n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->dcm3 = DevCM3{cm->lcm, cm->gap, ctx->d_cost3.p, ctx->d_median3.p};
    ctx->has_cm3 = true;
    return POYB200_OK;
}

extern "C" int poyb200_batch_median_3(poyb200_ctx *ctx, const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t in_stride,
                                      const int32_t *len, int32_t n, uint8_t *out, int64_t out_stride, int32_t *out_len) {
    if (!ctx) return POYB200_EINVAL;
    if (!ctx->has_cm3) return fail(ctx, POYB200_ENOCM, "no 3-D cost matrix loaded (poyb200_set_cm_3d)");
    if (n < 0) return fail(ctx, POYB200_EINVAL, "negative count");
    if (n == 0) return POYB200_OK;
    if (!a || !b || !c || !len || !out || !out_len) return fail(ctx, POYB200_EINVAL, "NULL array");
    for (int p = 0; p < n; p++)
        if (len[p] < 0 || len[p] > in_stride || len[p] > out_stride) return fail(ctx, POYB200_EINVAL, "row length does not fit the strides");
    cudaSetDevice(ctx->device);
    const size_t ib = (size_t) n * in_stride, ob = (size_t) n * out_stride;
    CK(ctx->d_out[1].reserve(ib + 16));
    CK(ctx->d_out[2].reserve(ib + 16));
    CK(ctx->d_out[3].reserve(ib + 16));
    CK(ctx->d_out[0].reserve(ob + 16));
    CK(ctx->d_outlen.reserve(2 * (size_t) n + 4));
    CK(cudaMemcpyAsync(ctx->d_out[1].p, a, ib, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_out[2].p, b, ib, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_out[3].p, c, ib, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_outlen.p + n, len, (size_t) n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(median_3_launch(ctx->dcm3.median, ctx->dcm3.lcm, ctx->d_out[1].p, ctx->d_out[2].p, ctx->d_out[3].p, in_stride, ctx->d_outlen.p + n, n,
                       ctx->d_out[0].p, out_stride, ctx->d_outlen.p, ctx->stream));
    ctx->launches++;
    CK(cudaMemcpyAsync(out, ctx->d_out[0].p, ob, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_len, ctx->d_outlen.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged = false;
    return POYB200_OK;
}

extern "C" int poyb200_batch_align_3(poyb200_ctx *ctx, const poyb200_batch3 *b) {
    if (!ctx || !b) return POYB200_EINVAL;
    if (!ctx->has_cm3) return fail(ctx, POYB200_ENOCM, "no 3-D cost matrix loaded (poyb200_set_cm_3d)");
    const int n = b->n_triples;
    if (n < 0 || b->n_seqs < 0) return fail(ctx, POYB200_EINVAL, "negative count");
    if (n == 0) return POYB200_OK;
    if (!b->pool || !b->seq_off || !b->seq_len || !b->triples || !b->cost) return fail(ctx, POYB200_EINVAL, "NULL input array");
    if (b->pool_bytes >= ((size_t) 1 << 32)) return fail(ctx, POYB200_EINVAL, "pool larger than 4 GiB");
    const bool want_al = (b->want & POYB200_WANT3_ALIGNED) != 0, want_med = (b->want & POYB200_WANT3_MEDIAN) != 0;
    const bool bt = want_al || want_med;
    if (want_al && (!b->aligned_1 || !b->aligned_2 || !b->aligned_3)) return fail(ctx, POYB200_EINVAL, "WANT3_ALIGNED without buffers");
    if (want_med && !b->median) return fail(ctx, POYB200_EINVAL, "WANT3_MEDIAN without buffer");
    if (bt && (!b->out_len || !b->status)) return fail(ctx, POYB200_EINVAL, "out_len / status is NULL");
    cudaSetDevice(ctx->device);
    std::vector<Task3> tasks((size_t) n);
    size_t max_ring = 1;
    int max_l3 = 1;
    long long maxcap = 16;
    for (int s = 0; s < b->n_seqs; s++) {
        if (b->seq_len[s] < 1 || b->seq_len[s] > POYB200_MAX_SEQ_LEN) return fail(ctx, POYB200_ESEQLEN, "sequence empty or longer than 16384");
        if (b->seq_off[s] < 0 || (size_t) (b->seq_off[s] + b->seq_len[s]) > b->pool_bytes) return fail(ctx, POYB200_EINVAL, "sequence outside the pool");
    }
    for (int p = 0; p < n; p++) {
        Task3 t{};
        int idx[3];
        for (int k = 0; k < 3; k++) {
            idx[k] = b->triples[3 * p + k];
            if (idx[k] < 0 || idx[k] >= b->n_seqs) return fail(ctx, POYB200_EINVAL, "triple index out of range");
        }
        t.off1 = (uint32_t) b->seq_off[idx[0]]; t.off2 = (uint32_t) b->seq_off[idx[1]]; t.off3 = (uint32_t) b->seq_off[idx[2]];
        t.l1 = b->seq_len[idx[0]]; t.l2 = b->seq_len[idx[1]]; t.l3 = b->seq_len[idx[2]];
        t.triple = (uint32_t) p;
        max_ring = std::max(max_ring, (size_t) (t.l1 + t.l2 + 3) * t.l3);
        max_l3 = std::max(max_l3, t.l3);
        maxcap = std::max<long long>(maxcap, (long long) t.l1 + t.l2 + t.l3);
        tasks[p] = t;
    }
    if (bt && b->out_stride < maxcap) return fail(ctx, POYB200_EINVAL, "out_stride smaller than l1 + l2 + l3");
    const long long dstride = (maxcap + 15) & ~15ll;
    // chunks by direction-cube bytes
    struct C3 { size_t begin, end; };
    std::vector<C3> chunks;
    size_t begin = 0, off = 0, maxdir = 16;
    for (size_t k = 0; k < tasks.size(); k++) {
        size_t bytes = bt ? (((size_t) tasks[k].l1 * tasks[k].l2 * tasks[k].l3 + 63) & ~(size_t) 63) : 0;
        if (k > begin && off + bytes > ctx->dir_budget) {
            chunks.push_back(C3{begin, k});
            maxdir = std::max(maxdir, off);
            begin = k;
            off = 0;
        }
        tasks[k].dir_off = off;
        off += bytes;
    }
    chunks.push_back(C3{begin, tasks.size()});
    maxdir = std::max(maxdir, off);
    // CTAs: as many as keep all rings inside ~96 MB of L2, at least one per SM's worth of work
    const size_t ring_bytes = max_ring * sizeof(int);
    int per_sm = (int) std::max<size_t>(1, std::min<size_t>(4, ((size_t) 96 << 20) / (ring_bytes * ctx->sm_count + 1)));
    const int grid_max = ctx->sm_count * per_sm;
    CK(ctx->d_pool.reserve(b->pool_bytes + 64));
    CK(ctx->d_tasks3.reserve((size_t) n));
    CK(ctx->d_costs.reserve((size_t) n + 1));
    CK(ctx->d_ring.reserve(max_ring * (size_t) grid_max));
    if (bt && ctx->ring_slot_bytes) {
        // as many slots per warp as the ring kernels can use, within the direction budget (they run with fewer if need be)
        // never more than one slot group per batch of the call: small calls (a tree level) and long pairs stay cheap
        const size_t full = ring_scratch_bytes(ctx->sm_count, ctx->ring_slot_bytes);
        const size_t by_pairs = ((size_t) n / 4 + 1) * 32 * ctx->ring_slot_bytes;
        CK(ctx->d_scratch.reserve(std::min(std::min(full, by_pairs), ctx->dir_budget)));
    }
    if (bt) {
        CK(ctx->d_dir.reserve(maxdir));
        CK(ctx->d_outlen.reserve((size_t) n + 4));
        CK(ctx->d_status.reserve((size_t) n + 4));
        const size_t ob = (size_t) n * dstride + 16;
        if (want_al) { CK(ctx->d_out[0].reserve(ob)); CK(ctx->d_out[1].reserve(ob)); CK(ctx->d_out[2].reserve(ob)); }
        if (want_med) CK(ctx->d_out[3].reserve(ob));
    }
    CK(cudaMemcpyAsync(ctx->d_pool.p, b->pool, b->pool_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tasks3.p, tasks.data(), (size_t) n * sizeof(Task3), cudaMemcpyHostToDevice, ctx->stream));
    Out3 out{ctx->d_costs.p, ctx->d_outlen.p, ctx->d_status.p, ctx->d_out[0].p, ctx->d_out[1].p, ctx->d_out[2].p, ctx->d_out[3].p,
             dstride, b->want};
    for (const C3 &ch : chunks) {
        const int nt = (int) (ch.end - ch.begin);
        const int grid = std::min(nt, grid_max);
        const int cube_threads = std::min(CUBE_THREADS, std::max(64, (max_l3 + 31) & ~31));
        CK(cube_fill_launch(grid, cube_threads, ctx->d_tasks3.p + ch.begin, nt, ctx->dcm3, ctx->d_pool.p, ctx->d_ring.p, max_ring,
                            ctx->d_dir.p, ctx->d_costs.p, bt ? 1 : 0, ctx->stream));
        ctx->launches++;
        if (bt) {
            CK(cube_traceback_launch(ctx->d_tasks3.p + ch.begin, nt, ctx->dcm3, ctx->d_pool.p, ctx->d_dir.p, out, ctx->stream));
            ctx->launches++;
        }
    }
    CK(cudaMemcpyAsync(b->cost, ctx->d_costs.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (bt) {
        CK(cudaMemcpyAsync(b->out_len, ctx->d_outlen.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(b->status, ctx->d_status.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        uint8_t *dst[4] = {b->aligned_1, b->aligned_2, b->aligned_3, b->median};
        const bool need[4] = {want_al, want_al, want_al, want_med};
        const size_t w = (size_t) std::min<long long>(dstride, b->out_stride);
        for (int k = 0; k < 4; k++)
            if (need[k])
                CK(cudaMemcpy2DAsync(dst[k] + (b->out_stride - w), (size_t) b->out_stride, ctx->d_out[k].p + (dstride - w),
                                     (size_t) dstride, w, (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->staged = false;
    return POYB200_OK;
}
