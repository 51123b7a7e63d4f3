// One batch over several GPUs of the node (BASELINE.json north_star: "batches shard across the 8 B200s by pair index, with
// results gathered on the host").  One context and one host thread per device; the pair list is cut into
// contiguous ranges of about equal DP work; every shard uploads only the window of the pool its pairs reference and writes
// its results straight into the caller's buffers at its pairs' rows, so the gather costs nothing.  No collective: pairs are
// independent.
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/poyb200.h"

int poyb200_one_shot_view(poyb200_ctx *ctx, int mode, const poyb200_batch *b, int64_t view_lo);  // api.cu

struct poyb200_multi {
    std::vector<poyb200_ctx *> ctx;
    std::vector<int> devices;
    std::string err;
    std::vector<int64_t> shard_begin;  // pair ranges of the last call (introspection)
};

extern "C" int poyb200_multi_create(const int *devices, int n_devices, const poyb200_config *cfg, poyb200_multi **out) {
    if (!out || n_devices < 1 || n_devices > 64) return POYB200_EINVAL;
    *out = nullptr;
    poyb200_multi *m = new poyb200_multi();
    for (int k = 0; k < n_devices; k++) {
        poyb200_ctx *c = nullptr;
        const int dev = devices ? devices[k] : k;
        const int rc = poyb200_create_ex(dev, cfg, &c);
        if (rc != POYB200_OK) {
            for (auto *x : m->ctx) poyb200_destroy(x);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
        m->devices.push_back(dev);
    }
    *out = m;
    return POYB200_OK;
}

extern "C" void poyb200_multi_destroy(poyb200_multi *m) {
    if (!m) return;
    for (auto *c : m->ctx) poyb200_destroy(c);
    delete m;
}

extern "C" const char *poyb200_multi_last_error(const poyb200_multi *m) { return m ? m->err.c_str() : "null handle"; }
extern "C" int poyb200_multi_devices(const poyb200_multi *m) { return m ? (int) m->ctx.size() : 0; }
extern "C" poyb200_ctx *poyb200_multi_ctx(poyb200_multi *m, int k) {
    return (m && k >= 0 && k < (int) m->ctx.size()) ? m->ctx[k] : nullptr;
}

extern "C" int poyb200_multi_set_cm(poyb200_multi *m, const poyb200_cm *cm) {
    if (!m) return POYB200_EINVAL;
    for (auto *c : m->ctx) {
        const int rc = poyb200_set_cm(c, cm);
        if (rc != POYB200_OK) {
            m->err = poyb200_last_error(c);
            return rc;
        }
    }
    return POYB200_OK;
}

extern "C" int64_t poyb200_multi_launch_count(const poyb200_multi *m) {
    int64_t n = 0;
    if (m)
        for (auto *c : m->ctx) n += poyb200_launch_count(c);
    return n;
}

extern "C" int poyb200_multi_shards(const poyb200_multi *m, int64_t *begin, int cap) {
    if (!m) return 0;
    const int n = (int) m->shard_begin.size();
    for (int k = 0; k < n && k < cap; k++) begin[k] = m->shard_begin[k];
    return n;
}

// mode: 0 cost_2, 1 align_2, 2 cost_affine_3, 3 align_affine_3 (as poyb200_stage)
extern "C" int poyb200_multi_batch(poyb200_multi *m, int mode, const poyb200_batch *b) {
    if (!m || !b) return POYB200_EINVAL;
    if (mode < 0 || mode > 3) { m->err = "bad mode"; return POYB200_EINVAL; }
    const int G = (int) m->ctx.size();
    const int64_t n = b->n_pairs;
    if (n < 0 || b->n_seqs < 0) { m->err = "negative count"; return POYB200_EINVAL; }
    if (n > 0 && (!b->pool || !b->seq_off || !b->seq_len || !b->pairs)) { m->err = "NULL input array"; return POYB200_EINVAL; }
    const bool affine = mode >= 2;
    if (!affine && n > 0 && !b->deltaw) { m->err = "linear entry points need deltaw[]"; return POYB200_EINVAL; }
    for (int64_t p = 0; p < 2 * n; p++)
        if (b->pairs[p] < 0 || b->pairs[p] >= b->n_seqs) { m->err = "pair index out of range"; return POYB200_EINVAL; }
    // work of a pair ~ stripe width x shorter length (the cells the reference visits, to within the band's corners)
    std::vector<double> cum((size_t) n + 1, 0.0);
    {
        const int T = (int) std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        auto job = [&](int t) {
            for (int64_t p = n * t / T; p < n * (t + 1) / T; p++) {
                const int la = b->seq_len[b->pairs[2 * p]], lb = b->seq_len[b->pairs[2 * p + 1]];
                const int lo = std::min(la, lb), hi = std::max(la, lb);
                double w;
                if (affine) w = (double) std::min(hi, std::max(40, hi - lo + 8) + 40) * lo;
                else w = (double) std::min(hi, 2 * (50 + b->deltaw[p]) + (hi - lo)) * lo;
                cum[(size_t) p + 1] = w + 64.0;
            }
        };
        for (int t = 1; t < T; t++) th.emplace_back(job, t);
        job(0);
        for (auto &x : th) x.join();
        for (int64_t p = 0; p < n; p++) cum[(size_t) p + 1] += cum[(size_t) p];
    }
    std::vector<int64_t> cut((size_t) G + 1, n);
    cut[0] = 0;
    for (int k = 1; k < G; k++) {
        const double want = cum[(size_t) n] * k / G;
        cut[k] = (int64_t) (std::lower_bound(cum.begin(), cum.end(), want) - cum.begin());
        cut[k] = std::max(cut[k], cut[k - 1]);
    }
    m->shard_begin.assign(cut.begin(), cut.end());
    std::vector<int> rcs((size_t) G, POYB200_OK);
    std::vector<std::thread> th;
    auto shard = [&](int k) {
        const int64_t p0 = cut[k], p1 = cut[k + 1];
        if (p1 <= p0) return;
        // the window of the pool this shard needs
        int64_t lo = INT64_MAX, hi = 0;
        for (int64_t p = 2 * p0; p < 2 * p1; p++) {
            const int s = b->pairs[p];
            lo = std::min(lo, b->seq_off[s]);
            hi = std::max(hi, b->seq_off[s] + (int64_t) b->seq_len[s]);
        }
        if (lo < 0 || (size_t) hi > b->pool_bytes) { rcs[k] = POYB200_EINVAL; return; }
        lo &= ~(int64_t) 15;  // keeps 16-byte aligned sequence starts aligned on the device
        poyb200_batch sub = *b;
        sub.pool = b->pool + lo;
        sub.pool_bytes = (size_t) (hi - lo);
        sub.pairs = b->pairs + 2 * p0;
        sub.n_pairs = (int32_t) (p1 - p0);
        if (b->deltaw) sub.deltaw = b->deltaw + p0;
        if (b->swaped) sub.swaped = b->swaped + p0;
        if (b->cost) sub.cost = b->cost + p0;
        if (b->out_len) sub.out_len = b->out_len + 4 * p0;
        if (b->median) sub.median = b->median + p0 * b->out_stride;
        if (b->medianwg) sub.medianwg = b->medianwg + p0 * b->out_stride;
        if (b->aligned_a) sub.aligned_a = b->aligned_a + p0 * b->out_stride;
        if (b->aligned_b) sub.aligned_b = b->aligned_b + p0 * b->out_stride;
        if (b->bits_a) sub.bits_a = b->bits_a + p0 * b->bits_stride;
        if (b->bits_b) sub.bits_b = b->bits_b + p0 * b->bits_stride;
        if (b->bits_wg) sub.bits_wg = b->bits_wg + p0 * b->bits_stride;
        rcs[k] = poyb200_one_shot_view(m->ctx[k], mode, &sub, lo);
    };
    for (int k = 1; k < G; k++) th.emplace_back(shard, k);
    shard(0);
    for (auto &x : th) x.join();
    for (int k = 0; k < G; k++)
        if (rcs[k] != POYB200_OK) {
            m->err = "device " + std::to_string(m->devices[k]) + ": " + poyb200_last_error(m->ctx[k]);
            return rcs[k];
        }
    return POYB200_OK;
}
