// Device-resident sequence store (include/poyb200.h, "store" section): sequences live in HBM, batches name their operands
// by store id, and what a batch produces -- the medians of a downpass level, the single assignments of an uppass depth --
// is appended to the store ON THE DEVICE.  Per batch only the task array goes up (56 bytes per pair) and costs, lengths and
// gap counts come back (12 bytes per pair): a tree level never crosses the host link with its sequences.
//
// Reference semantics kept here, so that callers stay thin: SeqCS.DOS.median's empty-operand rule is the caller's
// (it needs no alignment); Sequence.Align.cost_2's deltaw (src/sequence.ml:691-714: `count gap` of both operands + the
// length rule) is computed from the gap counts the store keeps per sequence; Sequence.Align.closest's two early exits
// (src/sequence.ml:975-1009) are taken inside poyb200_store_closest.
#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "ctx.h"

struct poyb200_store {
    poyb200_ctx *ctx = nullptr;
    uint8_t *d_pool = nullptr;
    size_t cap = 0, used = 0;
    std::vector<int64_t> off;
    std::vector<int32_t> len, cnt;  // cnt: elements with (code & gap) != 0
    std::vector<uint8_t> empty;     // Sequence.is_empty: every element is the gap
    DevBuf<long long> d_newoff;
    DevBuf<int> d_stats;
    DevBuf<uint4> d_jobs;
    DevBuf<uint8_t> d_eq;
    std::vector<int32_t> h_stats, h_cost, dw;
    std::vector<long long> h_newoff;
    int64_t pairs_done = 0, calls = 0, cells = 0;
    std::unordered_map<uint64_t, int64_t> cells_memo;  // DP cells by (len a, len b[, deltaw]): the formulas walk the rows
};

static int store_reserve(poyb200_store *s, size_t need) {
    poyb200_ctx *ctx = s->ctx;
    if (need <= s->cap) return POYB200_OK;
    if (need >= ((size_t) 1 << 32)) return fail(ctx, POYB200_ENOMEM, "store larger than 4 GiB (task offsets are 32 bit)");
    size_t want = std::max<size_t>(need + need / 2, (size_t) 64 << 20);
    want = std::min<size_t>(want, ((size_t) 1 << 32) - 256);
    uint8_t *p = nullptr;
    CK(cudaMalloc(&p, want + 256));
    if (s->used) CK(cudaMemcpyAsync(p, s->d_pool, s->used, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (s->d_pool) cudaFree(s->d_pool);
    s->d_pool = p;
    s->cap = want;
    return POYB200_OK;
}

extern "C" int poyb200_store_create(poyb200_ctx *ctx, poyb200_store **out) {
    if (!ctx || !out) return POYB200_EINVAL;
    poyb200_store *s = new poyb200_store();
    s->ctx = ctx;
    *out = s;
    return POYB200_OK;
}

extern "C" void poyb200_store_destroy(poyb200_store *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->d_pool) cudaFree(s->d_pool);
    s->d_newoff.release(); s->d_stats.release(); s->d_jobs.release(); s->d_eq.release();
    delete s;
}

extern "C" int32_t poyb200_store_size(const poyb200_store *s) { return s ? (int32_t) s->len.size() : 0; }
extern "C" int64_t poyb200_store_bytes(const poyb200_store *s) { return s ? (int64_t) s->used : 0; }

extern "C" int poyb200_store_add(poyb200_store *s, const uint8_t *bytes, const int64_t *off, const int32_t *len, int32_t n,
                                 int32_t *first_id) {
    if (!s || n < 0 || (n > 0 && (!bytes || !off || !len))) return POYB200_EINVAL;
    poyb200_ctx *ctx = s->ctx;
    if (!ctx->has_cm) return fail(ctx, POYB200_ENOCM, "no cost matrix loaded (the store keeps gap counts: it needs the gap code)");
    cudaSetDevice(ctx->device);
    const int gap = ctx->hcm.gap;
    size_t total = 0;
    for (int k = 0; k < n; k++) {
        if (len[k] < 1 || len[k] > POYB200_MAX_SEQ_LEN) return fail(ctx, POYB200_ESEQLEN, "sequence empty (no leading gap) or longer than 16384");
        total += ((size_t) len[k] + 15) & ~(size_t) 15;
    }
    int rc = store_reserve(s, s->used + total + 64);
    if (rc) return rc;
    std::vector<uint8_t> stage(total + 16, 0);
    size_t o = 0;
    if (first_id) *first_id = (int32_t) s->len.size();
    for (int k = 0; k < n; k++) {
        const uint8_t *src = bytes + off[k];
        memcpy(stage.data() + o, src, (size_t) len[k]);
        int c = 0, all = 1;
        for (int q = 0; q < len[k]; q++) { c += (src[q] & gap) != 0; all &= (src[q] == gap); }
        s->off.push_back((int64_t) (s->used + o));
        s->len.push_back(len[k]);
        s->cnt.push_back(c);
        s->empty.push_back((uint8_t) all);
        o += ((size_t) len[k] + 15) & ~(size_t) 15;
    }
    if (total) CK(cudaMemcpyAsync(s->d_pool + s->used, stage.data(), total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    s->used += total;
    return POYB200_OK;
}

extern "C" int poyb200_store_info(const poyb200_store *s, int32_t id, int32_t *len, int32_t *empty, int32_t *gap_count) {
    if (!s || id < 0 || id >= (int32_t) s->len.size()) return POYB200_EINVAL;
    if (len) *len = s->len[id];
    if (empty) *empty = s->empty[id];
    if (gap_count) *gap_count = s->cnt[id];
    return POYB200_OK;
}

extern "C" int poyb200_store_get(poyb200_store *s, int32_t id, uint8_t *out) {
    if (!s || !out || id < 0 || id >= (int32_t) s->len.size()) return POYB200_EINVAL;
    poyb200_ctx *ctx = s->ctx;
    cudaSetDevice(ctx->device);
    CK(cudaMemcpyAsync(out, s->d_pool + s->off[id], (size_t) s->len[id], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return POYB200_OK;
}

// deltaw of Sequence.Align.cost_2 (src/sequence.ml:691-714) for the pair (a, b); hint < 0 = no ?deltaw argument.
static int32_t deltaw_for(const poyb200_store *s, int a, int b, int hint) {
    const int l1 = std::max(s->len[a], s->len[b]), l2 = std::min(s->len[a], s->len[b]);
    const int dif = l1 - l2, lower = (int) ((double) l1 * 0.10);
    int d;
    if (hint < 0) d = (dif < lower) ? lower / 2 : 2;
    else d = (dif < lower) ? lower : hint;
    return std::max(s->cnt[a], s->cnt[b]) + d;
}

// Runs one alignment batch on store ids; the rows of the result stay in ctx->d_out[0] / d_outlen.
static int store_run(poyb200_store *s, int mode, uint32_t want, const int32_t *pairs, int32_t n, const int32_t *hint, int32_t *cost) {
    poyb200_ctx *ctx = s->ctx;
    const bool affine = ctx->hcm.cost_model_type == 1;
    const int ns = (int) s->len.size();
    for (int p = 0; p < 2 * n; p++)
        if (pairs[p] < 0 || pairs[p] >= ns) return fail(ctx, POYB200_EINVAL, "store id out of range");
    poyb200_batch b;
    memset(&b, 0, sizeof b);
    b.pool = reinterpret_cast<const uint8_t *>(s->d_pool);  // a DEVICE pointer: never dereferenced on the host (device_store)
    b.pool_bytes = s->used;
    b.seq_off = s->off.data();
    b.seq_len = s->len.data();
    b.n_seqs = ns;
    b.pairs = pairs;
    b.n_pairs = n;
    b.want = want;
    if (!affine) {
        s->dw.resize((size_t) n);
        for (int p = 0; p < n; p++) s->dw[p] = deltaw_for(s, pairs[2 * p], pairs[2 * p + 1], hint ? hint[p] : -1);
        b.deltaw = s->dw.data();
    }
    ctx->device_store = true;
    ctx->cur_pool = s->d_pool;
    ctx->view = true;  // validate the operands pair by pair, not the whole (ever growing) store on every call
    ctx->view_lo = 0;
    int rc = poyb200_stage_internal(ctx, mode, &b, false);
    ctx->view = false;
    if (rc == POYB200_OK) {
        cudaError_t e = cudaMemcpyAsync(ctx->d_tasks.p, ctx->tasks.data(), (size_t) n * sizeof(Task), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, POYB200_ECUDA, cudaGetErrorString(e));
    }
    if (rc == POYB200_OK) rc = poyb200_run_staged(ctx);
    ctx->device_store = false;
    if (rc) return rc;
    if (cost) CK(cudaMemcpyAsync(cost, ctx->d_costs.p, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    s->pairs_done += n;
    s->calls++;
    for (int p = 0; p < n; p++) {
        const int la = s->len[pairs[2 * p]], lb = s->len[pairs[2 * p + 1]];
        const uint64_t key = ((uint64_t) (uint32_t) std::max(la, lb) << 40) | ((uint64_t) (uint32_t) std::min(la, lb) << 20) |
                             (uint64_t) (affine ? 0 : (s->dw[p] & 0xfffff));
        auto it = s->cells_memo.find(key);
        if (it == s->cells_memo.end())
            it = s->cells_memo.emplace(key, affine ? poyb200_cells_affine(la, lb)
                                                   : poyb200_cells_linear(std::max(la, lb), std::min(la, lb), s->dw[p])).first;
        s->cells += it->second;
    }
    return POYB200_OK;
}

// The n rows of ctx->d_out[0] (lengths in d_outlen[4 p]) become store sequences; new_id[p] receives their ids.
static int store_append(poyb200_store *s, int32_t n, int32_t *new_id) {
    poyb200_ctx *ctx = s->ctx;
    CK(s->d_stats.reserve(2 * (size_t) n));
    CK(store_row_stats_launch(ctx->d_out[0].p, ctx->dstride, ctx->d_outlen.p, n, ctx->hcm.gap, s->d_stats.p, ctx->stream));
    ctx->launches++;
    s->h_stats.resize(2 * (size_t) n);
    CK(cudaMemcpyAsync(s->h_stats.data(), s->d_stats.p, 2 * (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    size_t total = 0;
    s->h_newoff.resize((size_t) n);
    for (int p = 0; p < n; p++) {
        s->h_newoff[p] = (long long) (s->used + total);
        total += ((size_t) s->h_stats[2 * p] + 15) & ~(size_t) 15;
    }
    int rc = store_reserve(s, s->used + total + 64);
    if (rc) return rc;
    CK(s->d_newoff.reserve((size_t) n));
    CK(cudaMemcpyAsync(s->d_newoff.p, s->h_newoff.data(), (size_t) n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(s->d_pool + s->used, 0, total, ctx->stream));  // the padding between sequences reads as zero
    CK(store_append_launch(ctx->d_out[0].p, ctx->dstride, ctx->d_outlen.p, s->d_newoff.p, n, s->d_pool, ctx->stream));
    ctx->launches++;
    for (int p = 0; p < n; p++) {
        new_id[p] = (int32_t) s->len.size();
        s->off.push_back(s->h_newoff[p]);
        s->len.push_back(s->h_stats[2 * p]);
        s->cnt.push_back(s->h_stats[2 * p + 1]);
        // a produced sequence has its only pure gap in front (gap columns are dropped, a gap is prepended): empty <=> length 1
        s->empty.push_back((uint8_t) (s->h_stats[2 * p] <= 1));
    }
    s->used += total;
    return POYB200_OK;
}

extern "C" int poyb200_store_median(poyb200_store *s, const int32_t *pairs, int32_t n, int32_t *cost, int32_t *new_id) {
    if (!s || n < 0 || (n > 0 && (!pairs || !cost || !new_id))) return POYB200_EINVAL;
    if (n == 0) return POYB200_OK;
    poyb200_ctx *ctx = s->ctx;
    cudaSetDevice(ctx->device);
    const bool affine = ctx->hcm.cost_model_type == 1;
    int rc = store_run(s, affine ? 3 : 1, POYB200_WANT_MEDIAN, pairs, n, nullptr, cost);
    if (rc) return rc;
    return store_append(s, n, new_id);
}

extern "C" int poyb200_store_distance(poyb200_store *s, const int32_t *pairs, int32_t n, const int32_t *hint, int32_t *cost) {
    if (!s || n < 0 || (n > 0 && (!pairs || !cost))) return POYB200_EINVAL;
    if (n == 0) return POYB200_OK;
    poyb200_ctx *ctx = s->ctx;
    cudaSetDevice(ctx->device);
    const bool affine = ctx->hcm.cost_model_type == 1;
    int rc = store_run(s, affine ? 2 : 0, 0, pairs, n, hint, cost);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return POYB200_OK;
}

extern "C" int poyb200_store_closest(poyb200_store *s, const int32_t *pairs, int32_t n, int32_t *new_id) {
    if (!s || n < 0 || (n > 0 && (!pairs || !new_id))) return POYB200_EINVAL;
    if (n == 0) return POYB200_OK;
    poyb200_ctx *ctx = s->ctx;
    if (!ctx->has_cm) return fail(ctx, POYB200_ENOCM, "no cost matrix loaded (poyb200_set_cm)");
    cudaSetDevice(ctx->device);
    const bool affine = ctx->hcm.cost_model_type == 1, comb = ctx->hcm.combinations != 0;
    const int ns = (int) s->len.size();
    for (int p = 0; p < 2 * n; p++)
        if (pairs[p] < 0 || pairs[p] >= ns) return fail(ctx, POYB200_EINVAL, "store id out of range");
    // exit 1: s2 empty -> s2 itself (src/sequence.ml:975-977)
    std::vector<int> todo;
    for (int p = 0; p < n; p++) {
        if (s->empty[pairs[2 * p + 1]]) new_id[p] = pairs[2 * p + 1];
        else todo.push_back(p);
    }
    // exit 2 (combination alphabets): s1 = s2 element by element
    std::vector<int> same, align;
    if (comb && !todo.empty()) {
        std::vector<uint4> jobs(todo.size());
        for (size_t k = 0; k < todo.size(); k++) {
            const int a = pairs[2 * todo[k]], b = pairs[2 * todo[k] + 1];
            jobs[k] = make_uint4((uint32_t) s->off[a], (uint32_t) s->off[b], (uint32_t) s->len[a], (uint32_t) s->len[b]);
        }
        CK(s->d_jobs.reserve(jobs.size()));
        CK(s->d_eq.reserve(jobs.size()));
        CK(cudaMemcpyAsync(s->d_jobs.p, jobs.data(), jobs.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream));
        CK(store_equal_launch(s->d_pool, s->d_jobs.p, (int) jobs.size(), s->d_eq.p, ctx->stream));
        ctx->launches++;
        std::vector<uint8_t> eq(jobs.size());
        CK(cudaMemcpyAsync(eq.data(), s->d_eq.p, jobs.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (size_t k = 0; k < todo.size(); k++) (eq[k] ? same : align).push_back(todo[k]);
    } else {
        align = todo;
    }
    std::vector<int32_t> ids;
    if (!align.empty()) {
        std::vector<int32_t> pp(2 * align.size());
        for (size_t k = 0; k < align.size(); k++) { pp[2 * k] = pairs[2 * align[k]]; pp[2 * k + 1] = pairs[2 * align[k] + 1]; }
        int rc = store_run(s, affine ? 3 : 1, POYB200_WANT_CLOSEST, pp.data(), (int) align.size(), nullptr, nullptr);
        if (rc) return rc;
        ids.resize(align.size());
        rc = store_append(s, (int) align.size(), ids.data());
        if (rc) return rc;
        for (size_t k = 0; k < align.size(); k++) new_id[align[k]] = ids[k];
    }
    if (!same.empty()) {
        const int m = (int) same.size();
        int maxlen = 1;
        std::vector<uint2> jobs(same.size());
        for (int k = 0; k < m; k++) {
            const int b = pairs[2 * same[k] + 1];
            jobs[k] = make_uint2((uint32_t) s->off[b], (uint32_t) s->len[b]);
            maxlen = std::max(maxlen, s->len[b]);
        }
        ctx->dstride = ((long long) maxlen + 1 + 15) & ~15ll;
        CK(ctx->d_out[0].reserve((size_t) m * (size_t) ctx->dstride + 16));
        CK(ctx->d_outlen.reserve(4 * (size_t) m + 4));
        CK(s->d_jobs.reserve((size_t) m));
        CK(cudaMemcpyAsync(s->d_jobs.p, jobs.data(), (size_t) m * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
        CK(store_closest_same_launch(ctx->dcm, s->d_pool, reinterpret_cast<const uint2 *>(s->d_jobs.p), m, ctx->d_out[0].p, ctx->dstride,
                                     ctx->d_outlen.p, ctx->stream));
        ctx->launches++;
        ids.resize(same.size());
        int rc = store_append(s, m, ids.data());
        if (rc) return rc;
        for (int k = 0; k < m; k++) new_id[same[k]] = ids[k];
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return POYB200_OK;
}

extern "C" void poyb200_store_stats(const poyb200_store *s, int64_t *calls, int64_t *pairs, int64_t *cells) {
    if (!s) return;
    if (calls) *calls = s->calls;
    if (pairs) *pairs = s->pairs_done;
    if (cells) *cells = s->cells;
}
