// Native host driver of include/poyb200_tree.h: level-order batching of a tree's medians, single assignments and edge
// distances on the device-resident store (store.cu).  Host code only; the arithmetic is the kernels'.
// The Python mirror poyd_b200/tree.py is the readable statement of the same algorithm (and cites the reference line by
// line); this file follows it function by function.
#include <algorithm>
#include <array>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/poyb200_tree.h"

namespace {

using Key = uint64_t;
inline Key key2(int a, int b) { return ((Key) (uint32_t) a << 32) | (uint32_t) b; }

struct Topo {
    // nb[v] = {parent, c1, c2} (-1 padded); deg 0 = absent, 1 = leaf, 3 = interior
    std::vector<std::array<int, 3>> nb;
    std::vector<int8_t> deg;
    std::vector<int> ids;  // present vertices, ascending
    int handle = -1;
    bool is_leaf(int v) const { return deg[v] == 1; }
    // index of the directional node (a looking away from b) in a flat per-tree table of 3 * nb.size() entries
    size_t dix(int a, int b) const {
        const auto &n = nb[a];
        return (size_t) 3 * a + (n[0] == b ? 0 : (n[1] == b ? 1 : 2));
    }
    void other_two(int nbr, int v, int &x, int &y) const {  // src/tree.ml:989-1010
        const auto &n = nb[v];
        if (nbr == n[0]) { x = n[1]; y = n[2]; }
        else if (nbr == n[1]) { x = n[0]; y = n[2]; }
        else { x = n[0]; y = n[1]; }
    }
    void set(int v, int a, int b, int c) {
        if (v >= (int) nb.size()) { nb.resize(v + 1, {-1, -1, -1}); deg.resize(v + 1, 0); }
        nb[v] = {a, b, c};
        deg[v] = (b < 0) ? 1 : 3;
    }
    void finish() {
        ids.clear();
        for (int v = 0; v < (int) deg.size(); v++)
            if (deg[v]) ids.push_back(v);
    }
    // Tree.get_pre_order_edges handle (src/tree.ml:1356-1423): every edge once, oriented away from the handle
    std::vector<std::pair<int, int>> pre_order_edges() const {
        std::vector<std::pair<int, int>> out, stack;
        const int h = handle, first = nb[h][0];
        if (!is_leaf(h)) {
            int x, y;
            other_two(first, h, x, y);
            stack = {{h, y}, {h, x}, {h, first}};
        } else {
            stack = {{h, first}};
        }
        while (!stack.empty()) {
            auto [pred, v] = stack.back();
            stack.pop_back();
            out.push_back({pred, v});
            if (!is_leaf(v)) {
                int x, y;
                other_two(pred, v, x, y);
                stack.push_back({v, y});
                stack.push_back({v, x});
            }
        }
        return out;
    }
};

Topo topo_from(const poyb200_topology &t) {
    Topo r;
    for (int k = 0; k < t.n_nodes; k++) r.set(t.ids[k], t.nbr[3 * k], t.nbr[3 * k + 1], t.nbr[3 * k + 2]);
    r.handle = t.handle;
    r.finish();
    return r;
}
void topo_write(const Topo &t, int32_t *ids, int32_t *nbr, int32_t *n_nodes, int32_t *handle) {
    int k = 0;
    for (int v : t.ids) {
        ids[k] = v;
        for (int q = 0; q < 3; q++) nbr[3 * k + q] = t.nb[v][q];
        k++;
    }
    if (n_nodes) *n_nodes = k;
    if (handle) *handle = t.handle;
}

// signature of every directional node of one tree, by Topo::dix; -1 = not known yet
using SigMap = std::vector<int>;

struct NodeInfo {
    std::vector<int> idx;  // store id per locus
    int64_t cost = 0;      // total cost of the subtree
    int minc = 0;          // min_child_code
    bool ready = false;
};
struct Fresh { int sx, sy, level, minc; };
struct EdgeInfo { std::vector<int> idx; int64_t cost = 0; bool ready = false; };

}  // namespace

struct poyb200_tree {
    poyb200_ctx *ctx = nullptr;
    poyb200_store *store = nullptr;
    int n_loci = 1;
    std::string err;
    std::unordered_map<int, std::vector<int>> leaves;  // code -> store id per locus (-1 = not set)
    std::unordered_map<int, std::vector<std::vector<uint8_t>>> leaf_bytes;  // the leaves themselves (to rebuild the store)
    int64_t trims = 0;
    int64_t done_calls = 0, done_pairs = 0, done_cells = 0;  // of stores that were trimmed away
    // caches (Evaluator._reset)
    std::unordered_map<Key, int> sig_;       // (sig x, sig y) ordered -> sig id of the median node
    std::vector<NodeInfo> node_;             // by sig id
    std::unordered_map<int, Fresh> fresh_;   // registered, not yet computed
    std::unordered_map<Key, EdgeInfo> edge_;
    std::unordered_map<int, int> leafsig_;
    std::unordered_map<Key, int> clo_;       // closest by (parent, mine) store id
    std::unordered_map<Key, int> dist_;      // DOS.distance by store id pair
    int64_t n_medians = 0;
    // last evaluate: singles[tree][vertex] -> store ids per locus
    std::vector<std::unordered_map<int, std::vector<int>>> last_singles;

    int fail(int rc, const std::string &m) { err = m; return rc; }
    int lib(int rc, const char *what) {
        if (rc != POYB200_OK) err = std::string(what) + ": " + poyb200_last_error(ctx);
        return rc;
    }
    bool emp(int id) const {
        int32_t e = 0;
        poyb200_store_info(store, id, nullptr, &e, nullptr);
        return e != 0;
    }
    int len(int id) const {
        int32_t l = 0;
        poyb200_store_info(store, id, &l, nullptr, nullptr);
        return l;
    }
    void reset() {
        sig_.clear(); node_.clear(); fresh_.clear(); edge_.clear(); leafsig_.clear(); clo_.clear(); dist_.clear();
        sig_.reserve(1 << 18); edge_.reserve(1 << 18); dist_.reserve(1 << 18); clo_.reserve(1 << 14);
        n_medians = 0;
    }
    // The store only grows (every median ever computed stays, which is what makes the cache work); a long search trims it:
    // a fresh store with the leaves only, all caches dropped -- the next evaluation recomputes what the current tree needs.
    int trim() {
        int64_t c = 0, p = 0, ce = 0;
        poyb200_store_stats(store, &c, &p, &ce);
        done_calls += c; done_pairs += p; done_cells += ce;
        poyb200_store_destroy(store);
        store = nullptr;
        int rc = lib(poyb200_store_create(ctx, &store), "poyb200_store_create");
        if (rc) return rc;
        reset();
        for (auto &kv : leaf_bytes) {
            auto &ids = leaves[kv.first];
            for (size_t l = 0; l < kv.second.size(); l++) {
                int64_t off = 0;
                int32_t len = (int32_t) kv.second[l].size(), id = -1;
                rc = lib(poyb200_store_add(store, kv.second[l].data(), &off, &len, 1, &id), "poyb200_store_add");
                if (rc) return rc;
                ids[l] = id;
            }
        }
        trims++;
        return POYB200_OK;
    }
    int new_sig() { node_.emplace_back(); return (int) node_.size() - 1; }
    int minc(int s) const { return node_[s].ready ? node_[s].minc : fresh_.at(s).minc; }

    int leaf(int code, int &sig) {
        auto it = leafsig_.find(code);
        if (it != leafsig_.end()) { sig = it->second; return POYB200_OK; }
        auto lv = leaves.find(code);
        if (lv == leaves.end()) return fail(POYB200_EINVAL, "taxon " + std::to_string(code) + " has no sequences (poyb200_tree_set_leaf)");
        for (int id : lv->second)
            if (id < 0) return fail(POYB200_EINVAL, "taxon " + std::to_string(code) + " misses a locus");
        sig = new_sig();
        node_[sig].idx = lv->second;
        node_[sig].cost = 0;
        node_[sig].minc = code;
        node_[sig].ready = true;
        leafsig_[code] = sig;
        return POYB200_OK;
    }

    // SeqCS.DOS.median over a batch with the empty-operand rule (src/seqCS.ml:748-752): (store id, cost) per job
    int medians(const std::vector<std::pair<int, int>> &jobs, std::vector<std::pair<int, int>> &out) {
        out.assign(jobs.size(), {-1, 0});
        std::vector<int32_t> pp;
        std::vector<size_t> where;
        for (size_t k = 0; k < jobs.size(); k++) {
            if (emp(jobs[k].first)) out[k] = {jobs[k].second, 0};
            else if (emp(jobs[k].second)) out[k] = {jobs[k].first, 0};
            else { pp.push_back(jobs[k].first); pp.push_back(jobs[k].second); where.push_back(k); }
        }
        if (!where.empty()) {
            std::vector<int32_t> cost(where.size()), ids(where.size());
            int rc = lib(poyb200_store_median(store, pp.data(), (int) where.size(), cost.data(), ids.data()), "poyb200_store_median");
            if (rc) return rc;
            for (size_t q = 0; q < where.size(); q++) out[where[q]] = {ids[q], cost[q]};
        }
        n_medians += (int64_t) jobs.size();
        return POYB200_OK;
    }

    // Signature of every directional node (u, v) = u looking away from v; missing medians are only registered.
    int collect(const Topo &t, SigMap &sig) {
        sig.assign(3 * t.nb.size(), -1);
        std::vector<std::pair<int, int>> stack;
        for (int u : t.ids) {
            for (int q = 0; q < t.deg[u]; q++) {
                stack.push_back({u, t.nb[u][q]});
                while (!stack.empty()) {
                    auto [a, b] = stack.back();
                    if (sig[t.dix(a, b)] >= 0) { stack.pop_back(); continue; }
                    if (t.is_leaf(a)) {
                        int s;
                        int rc = leaf(a, s);
                        if (rc) return rc;
                        sig[t.dix(a, b)] = s;
                        stack.pop_back();
                        continue;
                    }
                    int x, y;
                    t.other_two(b, a, x, y);
                    const bool hx = sig[t.dix(x, a)] >= 0, hy = sig[t.dix(y, a)] >= 0;
                    if (!hx || !hy) {
                        if (!hx) stack.push_back({x, a});
                        if (!hy) stack.push_back({y, a});
                        continue;
                    }
                    int sx = sig[t.dix(x, a)], sy = sig[t.dix(y, a)];
                    // Node.cs_median: the operand with the smaller min_child_code first (src/node.ml:343-348)
                    if (!(minc(sx) < minc(sy))) std::swap(sx, sy);
                    auto it = sig_.find(key2(sx, sy));
                    int s;
                    if (it == sig_.end()) {
                        s = new_sig();
                        sig_[key2(sx, sy)] = s;
                        auto lvl = [&](int z) { auto f = fresh_.find(z); return f == fresh_.end() ? 0 : f->second.level; };
                        fresh_[s] = Fresh{sx, sy, 1 + std::max(lvl(sx), lvl(sy)), std::min(minc(sx), minc(sy))};
                    } else {
                        s = it->second;
                    }
                    sig[t.dix(a, b)] = s;
                    stack.pop_back();
                }
            }
        }
        return POYB200_OK;
    }

    // The registered medians, one batch per dependency level, whatever number of trees registered them.
    int flush() {
        if (fresh_.empty()) return POYB200_OK;
        int top = 0;
        for (auto &kv : fresh_) top = std::max(top, kv.second.level);
        std::vector<std::vector<int>> by_level(top + 1);
        for (auto &kv : fresh_) by_level[kv.second.level].push_back(kv.first);
        for (int lv = 1; lv <= top; lv++) {
            auto &keys = by_level[lv];
            std::sort(keys.begin(), keys.end());  // creation order: deterministic batches
            std::vector<std::pair<int, int>> jobs, res;
            for (int k : keys) {
                const Fresh &f = fresh_[k];
                for (int l = 0; l < n_loci; l++) jobs.push_back({node_[f.sx].idx[l], node_[f.sy].idx[l]});
            }
            int rc = medians(jobs, res);
            if (rc) return rc;
            for (size_t i = 0; i < keys.size(); i++) {
                const Fresh f = fresh_[keys[i]];
                NodeInfo &n = node_[keys[i]];
                n.idx.resize(n_loci);
                int64_t c = node_[f.sx].cost + node_[f.sy].cost;
                for (int l = 0; l < n_loci; l++) { n.idx[l] = res[i * n_loci + l].first; c += res[i * n_loci + l].second; }
                n.cost = c;
                n.minc = std::min(node_[f.sx].minc, node_[f.sy].minc);
                n.ready = true;
            }
        }
        fresh_.clear();
        return POYB200_OK;
    }

    // refresh_all_edges (src/allDirChar.ml:672-700) for several trees: the median across every edge and its root cost
    int edge_medians_many(const std::vector<const SigMap *> &sigs, const std::vector<const Topo *> &topos,
                          const std::vector<std::vector<std::pair<int, int>>> &edges_all, std::vector<std::vector<Key>> &keys) {
        std::vector<std::pair<int, int>> jobs, res;
        std::vector<Key> todo;
        keys.assign(sigs.size(), {});
        for (size_t ti = 0; ti < sigs.size(); ti++) {
            for (auto [a, b] : edges_all[ti]) {
                int sa = (*sigs[ti])[topos[ti]->dix(a, b)], sb = (*sigs[ti])[topos[ti]->dix(b, a)];
                if (!(node_[sa].minc < node_[sb].minc)) std::swap(sa, sb);
                const Key k = key2(sa, sb);
                keys[ti].push_back(k);
                if (!edge_.count(k)) {
                    edge_[k] = EdgeInfo{};
                    todo.push_back(k);
                    for (int l = 0; l < n_loci; l++) jobs.push_back({node_[sa].idx[l], node_[sb].idx[l]});
                }
            }
        }
        int rc = medians(jobs, res);
        if (rc) return rc;
        for (size_t i = 0; i < todo.size(); i++) {
            const int sa = (int) (todo[i] >> 32), sb = (int) (todo[i] & 0xffffffffu);
            EdgeInfo &e = edge_[todo[i]];
            e.idx.resize(n_loci);
            int64_t c = node_[sa].cost + node_[sb].cost;
            for (int l = 0; l < n_loci; l++) { e.idx[l] = res[i * n_loci + l].first; c += res[i * n_loci + l].second; }
            e.cost = c;
            e.ready = true;
        }
        return POYB200_OK;
    }

    // closest parent mine by store id, each distinct (parent, mine) computed once
    int closest_cached(const std::vector<std::pair<int, int>> &jobs, std::vector<int> &out) {
        std::vector<int32_t> pp;
        std::vector<Key> fresh;
        for (auto &j : jobs) {
            const Key k = key2(j.first, j.second);
            if (!clo_.count(k)) {
                clo_[k] = -1;
                fresh.push_back(k);
                pp.push_back(j.first);
                pp.push_back(j.second);
            }
        }
        if (!fresh.empty()) {
            std::vector<int32_t> ids(fresh.size());
            int rc = lib(poyb200_store_closest(store, pp.data(), (int) fresh.size(), ids.data()), "poyb200_store_closest");
            if (rc) return rc;
            for (size_t q = 0; q < fresh.size(); q++) clo_[fresh[q]] = ids[q];
        }
        out.resize(jobs.size());
        for (size_t q = 0; q < jobs.size(); q++) out[q] = clo_[key2(jobs[q].first, jobs[q].second)];
        return POYB200_OK;
    }

    // DOS.distance (src/seqCS.ml:856-866: cost_2 ~deltaw:(max 8 |la - lb|)) for store id pairs not yet known
    int distances(const std::vector<std::pair<int, int>> &jobs) {
        std::vector<int32_t> pp, hint;
        std::vector<Key> fresh;
        for (auto &j : jobs) {
            const Key k = key2(j.first, j.second);
            if (dist_.count(k)) continue;
            if (emp(j.first) || emp(j.second)) { dist_[k] = 0; continue; }  // missing_distance
            dist_[k] = -1;
            fresh.push_back(k);
            pp.push_back(j.first);
            pp.push_back(j.second);
            hint.push_back(std::max(std::abs(len(j.first) - len(j.second)), 8));
        }
        if (!fresh.empty()) {
            std::vector<int32_t> cost(fresh.size());
            int rc = lib(poyb200_store_distance(store, pp.data(), (int) fresh.size(), hint.data(), cost.data()), "poyb200_store_distance");
            if (rc) return rc;
            for (size_t q = 0; q < fresh.size(); q++) dist_[fresh[q]] = cost[q];
        }
        return POYB200_OK;
    }

    int nonempty_parent(int parent, int mine) const { return emp(parent) ? mine : parent; }  // DOS.to_single, src/seqCS.ml:734-739

    int evaluate_many(const std::vector<Topo> &topos, bool keep, std::vector<poyb200_tree_cost> &out) {
        if (!keep) reset();
        fresh_.clear();
        const size_t nt = topos.size();
        std::vector<SigMap> sigs(nt);
        for (size_t ti = 0; ti < nt; ti++) {
            int rc = collect(topos[ti], sigs[ti]);
            if (rc) return rc;
        }
        int rc = flush();
        if (rc) return rc;
        std::vector<std::vector<std::pair<int, int>>> edges_all(nt);
        std::vector<const SigMap *> sp(nt);
        std::vector<const Topo *> tp(nt);
        for (size_t ti = 0; ti < nt; ti++) { edges_all[ti] = topos[ti].pre_order_edges(); sp[ti] = &sigs[ti]; tp[ti] = &topos[ti]; }
        std::vector<std::vector<Key>> ekeys;
        rc = edge_medians_many(sp, tp, edges_all, ekeys);
        if (rc) return rc;
        // general_pick_best_root with blindly_trust_downpass, tree by tree
        out.assign(nt, poyb200_tree_cost{});
        std::vector<size_t> root_pos(nt);
        for (size_t ti = 0; ti < nt; ti++) {
            const Topo &t = topos[ti];
            const auto &edges = edges_all[ti];
            const int h = t.handle, hp = t.nb[h][0];
            size_t rp = 0;
            for (size_t q = 0; q < edges.size(); q++)
                if (edges[q].first == h && edges[q].second == hp) rp = q;  // create_root: the handle and its parent
            int64_t best = edge_[ekeys[ti][rp]].cost;
            std::vector<size_t> order(edges.size());
            for (size_t q = 0; q < order.size(); q++) order[q] = q;
            std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) {
                if (edges[x].first != edges[y].first) return edges[x].first > edges[y].first;
                return edges[x].second > edges[y].second;
            });
            for (size_t q : order) {
                const int64_t c = edge_[ekeys[ti][q]].cost;
                if (std::llabs(best) > std::llabs(c)) { best = c; rp = q; }
            }
            root_pos[ti] = rp;
            out[ti].unadjusted = best;
            out[ti].root_a = edges[rp].first;
            out[ti].root_b = edges[rp].second;
        }
        // assign_single (uppass): all trees advance one depth per batch
        last_singles.assign(nt, {});
        struct Front { int ti, parent, cur; std::vector<int> ps; };
        std::vector<Front> frontier;
        {
            std::vector<std::pair<int, int>> jobs;
            for (size_t ti = 0; ti < nt; ti++) {
                const int a = out[ti].root_a, b = out[ti].root_b;
                const EdgeInfo &E = edge_[ekeys[ti][root_pos[ti]]];
                const NodeInfo &mine = node_[sigs[ti][topos[ti].dix(a, b)]];
                for (int l = 0; l < n_loci; l++) jobs.push_back({nonempty_parent(E.idx[l], mine.idx[l]), mine.idx[l]});
            }
            std::vector<int> rs;
            rc = closest_cached(jobs, rs);
            if (rc) return rc;
            for (size_t ti = 0; ti < nt; ti++) {
                std::vector<int> r(rs.begin() + ti * n_loci, rs.begin() + (ti + 1) * n_loci);
                frontier.push_back({(int) ti, out[ti].root_b, out[ti].root_a, r});
                frontier.push_back({(int) ti, out[ti].root_a, out[ti].root_b, r});
            }
        }
        while (!frontier.empty()) {
            std::vector<std::pair<int, int>> jobs;
            for (auto &f : frontier) {
                const NodeInfo &mine = node_[sigs[f.ti][topos[f.ti].dix(f.cur, f.parent)]];
                for (int l = 0; l < n_loci; l++) jobs.push_back({nonempty_parent(f.ps[l], mine.idx[l]), mine.idx[l]});
            }
            std::vector<int> res;
            rc = closest_cached(jobs, res);
            if (rc) return rc;
            std::vector<Front> nxt;
            for (size_t k = 0; k < frontier.size(); k++) {
                const Front &f = frontier[k];
                std::vector<int> sg(res.begin() + k * n_loci, res.begin() + (k + 1) * n_loci);
                last_singles[f.ti][f.cur] = sg;
                if (!topos[f.ti].is_leaf(f.cur)) {
                    int x, y;
                    topos[f.ti].other_two(f.parent, f.cur, x, y);
                    nxt.push_back({f.ti, f.cur, x, sg});
                    nxt.push_back({f.ti, f.cur, y, sg});
                }
            }
            frontier.swap(nxt);
        }
        // check_cost: distances between single assignments along the edges, oriented away from the handle
        {
            std::vector<std::pair<int, int>> jobs;
            for (size_t ti = 0; ti < nt; ti++)
                for (auto [p, v] : edges_all[ti])
                    for (int l = 0; l < n_loci; l++) jobs.push_back({last_singles[ti][p][l], last_singles[ti][v][l]});
            rc = distances(jobs);
            if (rc) return rc;
            size_t q = 0;
            for (size_t ti = 0; ti < nt; ti++) {
                int64_t adj = 0;
                for (size_t e = 0; e < edges_all[ti].size(); e++)
                    for (int l = 0; l < n_loci; l++, q++) adj += dist_[key2(jobs[q].first, jobs[q].second)];
                out[ti].adjusted = adj;
            }
        }
        return POYB200_OK;
    }
};

// ---- C ABI ----------------------------------------------------------------------------------------------------------
extern "C" int poyb200_tree_create(poyb200_ctx *ctx, int32_t n_loci, poyb200_tree **out) {
    if (!ctx || !out || n_loci < 1) return POYB200_EINVAL;
    poyb200_tree *t = new poyb200_tree();
    t->ctx = ctx;
    t->n_loci = n_loci;
    int rc = poyb200_store_create(ctx, &t->store);
    if (rc) { delete t; return rc; }
    *out = t;
    return POYB200_OK;
}
extern "C" void poyb200_tree_destroy(poyb200_tree *t) {
    if (!t) return;
    poyb200_store_destroy(t->store);
    delete t;
}
extern "C" const char *poyb200_tree_last_error(const poyb200_tree *t) { return t ? t->err.c_str() : "null handle"; }
extern "C" void poyb200_tree_reset(poyb200_tree *t) { if (t) t->reset(); }

extern "C" int poyb200_tree_set_leaf(poyb200_tree *t, int32_t code, int32_t locus, const uint8_t *seq, int32_t len) {
    if (!t || !seq || locus < 0 || locus >= t->n_loci) return POYB200_EINVAL;
    int64_t off = 0;
    int32_t id = -1;
    int rc = t->lib(poyb200_store_add(t->store, seq, &off, &len, 1, &id), "poyb200_store_add");
    if (rc) return rc;
    auto &v = t->leaves[code];
    if (v.empty()) v.assign(t->n_loci, -1);
    v[locus] = id;
    auto &lb = t->leaf_bytes[code];
    if (lb.empty()) lb.resize(t->n_loci);
    lb[locus].assign(seq, seq + len);
    t->leafsig_.erase(code);
    return POYB200_OK;
}

extern "C" int poyb200_tree_evaluate(poyb200_tree *t, const poyb200_topology *topos, int32_t n, int32_t keep, poyb200_tree_cost *out) {
    if (!t || !topos || !out || n < 1) return POYB200_EINVAL;
    std::vector<Topo> tv;
    for (int k = 0; k < n; k++) tv.push_back(topo_from(topos[k]));
    std::vector<poyb200_tree_cost> res;
    int rc = t->evaluate_many(tv, keep != 0, res);
    if (rc) return rc;
    for (int k = 0; k < n; k++) out[k] = res[k];
    return POYB200_OK;
}

extern "C" int poyb200_tree_single(poyb200_tree *t, int32_t which, int32_t vertex, int32_t locus, uint8_t *out, int32_t cap, int32_t *len) {
    if (!t || which < 0 || which >= (int) t->last_singles.size() || locus < 0 || locus >= t->n_loci || !len) return POYB200_EINVAL;
    auto it = t->last_singles[which].find(vertex);
    if (it == t->last_singles[which].end()) return t->fail(POYB200_EINVAL, "no such vertex in the last evaluation");
    const int id = it->second[locus];
    *len = t->len(id);
    if (*len > cap || !out) return t->fail(POYB200_EINVAL, "buffer too small");
    return t->lib(poyb200_store_get(t->store, id, out), "poyb200_store_get");
}

extern "C" int poyb200_tree_wagner(poyb200_tree *t, const int32_t *order, int32_t n, int32_t *ids, int32_t *nbr, int32_t *n_nodes,
                                   int32_t *handle, int64_t *steps) {
    if (!t || !order || n < 2 || !ids || !nbr) return POYB200_EINVAL;
    t->reset();
    Topo topo;
    const int t1 = order[0], t2 = order[1];
    topo.set(t1, t2, -1, -1);
    topo.set(t2, t1, -1, -1);
    topo.handle = t1;
    topo.finish();
    int next_id = 0;
    for (auto &kv : t->leaves) next_id = std::max(next_id, kv.first);
    next_id++;
    for (int s = 2; s < n; s++) {
        const int c = order[s];
        SigMap sig;
        int rc = t->collect(topo, sig);
        if (rc) return rc;
        rc = t->flush();
        if (rc) return rc;
        auto edges = topo.pre_order_edges();
        std::vector<std::vector<Key>> ekeys;
        rc = t->edge_medians_many({&sig}, {&topo}, {edges}, ekeys);
        if (rc) return rc;
        int csig;
        rc = t->leaf(c, csig);
        if (rc) return rc;
        const std::vector<int> cl = t->node_[csig].idx;
        // AllDirChar.cost_fn on EVERY edge in one batch (the Wagner manager visits all edges, src/queues.ml:367-407)
        std::vector<std::pair<int, int>> jobs;
        std::vector<int> owner;
        for (size_t k = 0; k < edges.size(); k++) {
            const EdgeInfo &E = t->edge_[ekeys[0][k]];
            for (int l = 0; l < t->n_loci; l++)
                if (!(t->emp(cl[l]) || t->emp(E.idx[l]))) { jobs.push_back({cl[l], E.idx[l]}); owner.push_back((int) k); }
        }
        rc = t->distances(jobs);
        if (rc) return rc;
        std::vector<int64_t> delta(edges.size(), 0);
        for (size_t q = 0; q < jobs.size(); q++) delta[owner[q]] += t->dist_[key2(jobs[q].first, jobs[q].second)];
        size_t k = 0;
        for (size_t q = 1; q < edges.size(); q++)
            if (delta[q] < delta[k]) k = q;  // first minimum
        const int a = edges[k].first, b = edges[k].second, v = next_id++;
        auto repl = [](std::array<int, 3> x, int from, int to) { for (auto &z : x) if (z == from) z = to; return x; };
        topo.set(v, a, b, c);
        topo.nb[a] = repl(topo.nb[a], b, v);
        topo.nb[b] = repl(topo.nb[b], a, v);
        topo.set(c, v, -1, -1);
        topo.finish();
        if (steps) { steps[4 * (s - 2)] = c; steps[4 * (s - 2) + 1] = a; steps[4 * (s - 2) + 2] = b; steps[4 * (s - 2) + 3] = delta[k]; }
    }
    topo_write(topo, ids, nbr, n_nodes, handle);
    return POYB200_OK;
}

// ---- SPR (header comment of poyb200_tree.h) ------------------------------------------------------------------------------
namespace {
struct Break {
    int p, c, x, y;               // edge (p, c) broken: p leaves with the clade rooted at c; x - y are joined
    std::vector<uint8_t> pruned;  // by vertex id
    Topo rest;                    // the tree without the clade and without p
};
}  // namespace

// One round of the search on the breaks k with k % nshards == shard: the first candidate, in (break, join edge) order, whose
// exactly evaluated cost beats `best`.  found = 0 when none of this shard's candidates does.
static int spr_round(poyb200_tree *t, const Topo &cur, int64_t best, int shard, int nshards, int window, poyb200_spr_result *res,
                     int *found, int64_t *key, int64_t *cost, Topo *joined_out) {
    *found = 0;
    const int L = t->n_loci;
    if (poyb200_store_bytes(t->store) > ((int64_t) 1200 << 20)) {  // a round of a 500-taxon search adds ~2 GB: start it lean
        int rc0 = t->trim();
        if (rc0) return rc0;
    }
    // ---- the current tree's directional medians and edge medians (cached: free after the first round)
    SigMap sig;
    int rc = t->collect(cur, sig);
    if (rc) return rc;
    rc = t->flush();
    if (rc) return rc;
    auto cur_edges = cur.pre_order_edges();
    std::vector<std::vector<Key>> ck;
    rc = t->edge_medians_many({&sig}, {&cur}, {cur_edges}, ck);
    if (rc) return rc;
    std::unordered_map<Key, Key> edge_key;  // (a, b) either orientation -> edge_ key
    for (size_t q = 0; q < cur_edges.size(); q++) {
        edge_key[key2(cur_edges[q].first, cur_edges[q].second)] = ck[0][q];
        edge_key[key2(cur_edges[q].second, cur_edges[q].first)] = ck[0][q];
    }
    // ---- every break: interior p (not the handle), one of its neighbours c carries the clade that leaves with p
    std::vector<Break> breaks;
    std::vector<int64_t> ordinal;  // position of the break in the unsharded enumeration
    int64_t ord = 0;
    for (int p : cur.ids) {
        if (cur.deg[p] != 3 || p == cur.handle) continue;
        for (int q = 0; q < 3; q++) {
            const int c = cur.nb[p][q];
            Break b;
            b.p = p; b.c = c;
            cur.other_two(c, p, b.x, b.y);
            b.pruned.assign(cur.deg.size(), 0);
            b.pruned[p] = b.pruned[c] = 1;
            std::vector<int> stack{c};
            while (!stack.empty()) {
                const int v = stack.back();
                stack.pop_back();
                for (int z = 0; z < cur.deg[v]; z++) {
                    const int w = cur.nb[v][z];
                    if (!b.pruned[w]) { b.pruned[w] = 1; stack.push_back(w); }
                }
            }
            if (b.pruned[cur.handle]) continue;
            const int64_t my = ord++;
            if (my % nshards != shard) continue;
            for (int v : cur.ids) {
                if (b.pruned[v]) continue;
                auto n = cur.nb[v];
                for (auto &z : n) { if (z == p) z = (v == b.x) ? b.y : b.x; }
                b.rest.set(v, n[0], n[1], n[2]);
            }
            b.rest.handle = cur.handle;
            b.rest.finish();
            if (b.rest.ids.size() < 2) continue;
            breaks.push_back(std::move(b));
            ordinal.push_back(my);
        }
    }
    res->breaks += (int64_t) breaks.size();
    // ---- medians of the broken trees, all breaks at once
    std::vector<SigMap> bsig(breaks.size());
    for (size_t k = 0; k < breaks.size(); k++) {
        rc = t->collect(breaks[k].rest, bsig[k]);
        if (rc) return rc;
    }
    rc = t->flush();
    if (rc) return rc;
    std::vector<std::vector<std::pair<int, int>>> bedges(breaks.size());
    std::vector<const SigMap *> bsp(breaks.size());
    std::vector<const Topo *> btp(breaks.size());
    for (size_t k = 0; k < breaks.size(); k++) { bedges[k] = breaks[k].rest.pre_order_edges(); bsp[k] = &bsig[k]; btp[k] = &breaks[k].rest; }
    std::vector<std::vector<Key>> bkeys;
    rc = t->edge_medians_many(bsp, btp, bedges, bkeys);
    if (rc) return rc;
    // ---- the whole sweep: cost_fn = distance(clade root, median across the join edge), every break x every edge
    std::vector<std::pair<int, int>> jobs;
    for (size_t k = 0; k < breaks.size(); k++) {
        const NodeInfo &clade = t->node_[sig[cur.dix(breaks[k].c, breaks[k].p)]];
        for (size_t e = 0; e < bedges[k].size(); e++) {
            const EdgeInfo &E = t->edge_[bkeys[k][e]];
            for (int l = 0; l < L; l++) jobs.push_back({clade.idx[l], E.idx[l]});
        }
    }
    rc = t->distances(jobs);
    if (rc) return rc;
    res->joins_swept += (int64_t) (jobs.size() / (size_t) L);
    // ---- replay the first-best manager in order: candidates with cc < break delta
    struct Cand { size_t k, e; };
    std::vector<Cand> cands;
    size_t q = 0;
    for (size_t k = 0; k < breaks.size(); k++) {
        const Break &b = breaks[k];
        // break delta: what the median across the broken edge cost (prev root cost - the two sides' own costs)
        const int64_t b_delta = t->edge_[edge_key[key2(b.p, b.c)]].cost - t->node_[sig[cur.dix(b.c, b.p)]].cost -
                                t->node_[sig[cur.dix(b.p, b.c)]].cost;
        for (size_t e = 0; e < bedges[k].size(); e++) {
            int64_t cc = 0;
            for (int l = 0; l < L; l++, q++) cc += t->dist_[key2(jobs[q].first, jobs[q].second)];
            const int u = bedges[k][e].first, v = bedges[k][e].second;
            const bool same_place = (u == b.x && v == b.y) || (u == b.y && v == b.x);
            if (!same_place && cc < b_delta) cands.push_back({k, e});
        }
    }
    for (size_t lo = 0; lo < cands.size(); lo += (size_t) window) {
        const size_t hi = std::min(cands.size(), lo + (size_t) window);
        std::vector<Topo> joined;
        for (size_t z = lo; z < hi; z++) {
            const Break &b = breaks[cands[z].k];
            const int u = bedges[cands[z].k][cands[z].e].first, v = bedges[cands[z].k][cands[z].e].second;
            Topo n2 = cur;
            auto repl = [&](int w, int from, int to) { for (auto &zz : n2.nb[w]) if (zz == from) zz = to; };
            repl(b.x, b.p, b.y);
            repl(b.y, b.p, b.x);
            repl(u, v, b.p);
            repl(v, u, b.p);
            n2.nb[b.p] = {u, v, b.c};
            joined.push_back(std::move(n2));
        }
        std::vector<poyb200_tree_cost> jc;
        rc = t->evaluate_many(joined, true, jc);
        if (rc) return rc;
        res->exact_evaluated += (int64_t) joined.size();
        for (size_t z = 0; z < joined.size(); z++) {
            if (jc[z].adjusted < best) {  // cst < cur_best_cost: take it (Tree.Break)
                *found = 1;
                *key = (ordinal[cands[lo + z].k] << 24) | (int64_t) cands[lo + z].e;
                *cost = jc[z].adjusted;
                *joined_out = joined[z];
                return POYB200_OK;
            }
        }
    }
    return POYB200_OK;
}

extern "C" int poyb200_tree_spr_round(poyb200_tree *t, const poyb200_topology *cur, int64_t best_cost, int32_t shard, int32_t nshards,
                                      int32_t window, int32_t *found, int64_t *key, int64_t *cost, int32_t *ids, int32_t *nbr,
                                      int32_t *handle, poyb200_spr_result *res) {
    if (!t || !cur || !found || !key || !cost || !ids || !nbr || !res || nshards < 1 || shard < 0 || shard >= nshards) return POYB200_EINVAL;
    if (window < 1) window = 32;
    int f = 0;
    Topo joined;
    int rc = spr_round(t, topo_from(*cur), best_cost, shard, nshards, window, res, &f, key, cost, &joined);
    if (rc) return rc;
    *found = f;
    if (f) topo_write(joined, ids, nbr, nullptr, handle);
    return POYB200_OK;
}

extern "C" int poyb200_tree_spr(poyb200_tree *t, const poyb200_topology *start, int32_t max_rounds, int32_t window, int32_t *ids,
                                int32_t *nbr, int32_t *handle, poyb200_spr_result *res) {
    if (!t || !start || !ids || !nbr || !res) return POYB200_EINVAL;
    if (window < 1) window = 32;
    memset(res, 0, sizeof *res);
    Topo cur = topo_from(*start);
    std::vector<poyb200_tree_cost> tc;
    int rc = t->evaluate_many({cur}, true, tc);
    if (rc) return rc;
    int64_t best = tc[0].adjusted;
    res->start_cost = best;
    for (;;) {
        if (max_rounds > 0 && res->rounds >= max_rounds) break;
        int found = 0;
        int64_t key = 0, cost = 0;
        Topo joined;
        rc = spr_round(t, cur, best, 0, 1, window, res, &found, &key, &cost, &joined);
        if (rc) return rc;
        if (!found) break;
        best = cost;
        cur = joined;  // restart the search from the new tree
        res->rounds++;
    }
    res->final_cost = best;
    topo_write(cur, ids, nbr, nullptr, handle);
    return POYB200_OK;
}

extern "C" void poyb200_tree_stats(const poyb200_tree *t, int64_t *calls, int64_t *pairs, int64_t *cells, int64_t *medians,
                                   int64_t *sequences) {
    if (!t) return;
    int64_t c = 0, p = 0, ce = 0;
    poyb200_store_stats(t->store, &c, &p, &ce);
    if (calls) *calls = c + t->done_calls;
    if (pairs) *pairs = p + t->done_pairs;
    if (cells) *cells = ce + t->done_cells;
    if (medians) *medians = t->n_medians;
    if (sequences) *sequences = poyb200_store_size(t->store);
}
