// Powell's three-sequence affine-gap aligner (Ukkonen furthest-reaching formulation with check-pointing):
// src/ukk.checkp.c + src/ukkCommon.c of the reference, the aligner behind Sequence.Align.readjust_3d
// (src/sequence.ml:1075-1139).  Level-synchronous restatement: one cooperating group of threads per triple.
//
// The reference is a memoised recursion: U(ab, ac, d, s) = furthest position on sequence A reachable on diagonal
// (ab, ac) = (i - j, i - k) with cost d ending in state s (16 states over {match, delete, insert}^3, setup() :247-350),
// demanded top-down from (final diagonal, d, MMM) for d = 0, 1, 2, ... (doUkk :113-214), then the alignment is recovered by
// re-running the same computation between check-points (doUkkInLimits / getSplitRecurse :216-367) down to base cases
// that keep complete `from` pointers (traceBack :370-444).  Every U value and `from` record is a pure function of the
// pass parameters, so the order of evaluation is free -- except for one thing: the check-point of the FIRST pass is placed
// by `furthestReached`, the largest U over the cells the recursion happened to have computed so far (:174-180, :597), and a
// different check-point can give a different (equally optimal) alignment.  So this restatement reproduces exactly WHICH
// cells the recursion computes at every top-level cost T:
//   * the demand edges of calcUkk (:608-784) are data-independent: (x, d) asks for (y, d - transCost - contCost) on the
//     neighbour diagonal for all 16 from-states, for (x, d - 1), and MMM asks for the other states of its own cell;
//     a demand stops at cells outside withinMatrix (:545-576) and at memoised cells;
//   * hence the cells computed for a diagonal/state x are a contiguous range of costs lower(x) .. top(x), and top() is the
//     least fixpoint of  top(y) >= top(x) - w(x -> y)  over the demand edges, seeded with top(root) = T.  Going from T - 1
//     to T every active top grows by at least one; a relaxation sweep finds what else changed (relax());
//   * the new cells of a top level are computed in order of their cost (they depend on lower costs, MMM also on the other
//     states of its own cell), each exactly as calcUkk does (calc()).
// U lives in a window of Wd costs per (diagonal, state) (the reference keeps 2 maxSingleStep planes and recomputes what it
// lost, :136-141); every read checks the cost tag, a miss is reported as PW_EWINDOW, never papered over.
//
// The same source is compiled for the device (powell_kernel.cuh: PW_TID / PW_NT / PW_SYNC map to a CTA) and, single-threaded,
// for the host-side test harness (tests/powell_host.cpp) that checks it against the compiled reference without a GPU.
#pragma once
#include <stdint.h>
#include <string.h>

#ifndef PW_HD
#ifdef __CUDACC__
#define PW_HD __host__ __device__ __forceinline__
#else
#define PW_HD inline
#endif
#endif

namespace poyb200 {
namespace powell {

constexpr int NS = 16;          // states with at least one match and at most one insert (setup() :264-273)
constexpr int PW_INF = 5000;    // INFINITY = MAXINT / 2, MAXINT = 10000 (ukkCommon.h:41, ukk.checkp.c:31)
constexpr int NEGBIG = -(1 << 28);
constexpr int PW_STRIDE = 8;    // top levels advanced per closure once the check-point of the first pass is placed
enum { PW_OK = 0, PW_EBOX = 1, PW_EWINDOW = 2, PW_ELIST = 3, PW_ESTACK = 4, PW_EINPUT = 5, PW_ECAP = 6 };

struct Tables {
    int mis, go, ge, maxSingleStep;
    int da[NS], db[NS], dc[NS];   // which sequences the state's step consumes (neighbours[] + step(), :207-217, :282-293)
    int cont[NS], second[NS];     // contCost, secondCost (:296-313)
    int trans[NS][NS];            // transCost[from][to] (:320-337)
};

// setup() of ukkCommon.c for (misCost, startInsert = startDelete, continueInsert = continueDelete).
inline void make_tables(Tables &t, int mm, int go, int ge) {
    memset(&t, 0, sizeof t);
    t.mis = mm; t.go = go; t.ge = ge;
    int st[NS][3], ns = 0;
    for (int s = 0; s < 27; s++) {
        const int tr[3] = {s % 3, (s / 3) % 3, (s / 9) % 3};  // 0 match, 1 del, 2 ins
        int nm = 0, nd = 0, ni = 0;
        for (int i = 0; i < 3; i++) { nm += tr[i] == 0; nd += tr[i] == 1; ni += tr[i] == 2; }
        if (nm == 0 || ni > 1) continue;
        for (int i = 0; i < 3; i++) st[ns][i] = tr[i];
        const int sel = ni == 0 ? 0 : 2;  // no insert: the matching sequences move; else only the inserted one
        t.da[ns] = tr[0] == sel; t.db[ns] = tr[1] == sel; t.dc[ns] = tr[2] == sel;
        if (ni > 0) { t.cont[ns] = ge; t.second[ns] = 0; }
        else if (nm == 3) { t.cont[ns] = mm; t.second[ns] = 1; }
        else if (nd == 1) { t.cont[ns] = ge; t.second[ns] = 1; }
        else { t.cont[ns] = 2 * ge; t.second[ns] = 0; }
        ns++;
    }
    int maxc = 0;
    for (int s1 = 0; s1 < NS; s1++)
        for (int s2 = 0; s2 < NS; s2++) {
            int c = 0, nm = 0;
            for (int i = 0; i < 3; i++) {
                if (st[s2][i] != 0 && st[s2][i] != st[s1][i]) c += go;
                nm += st[s2][i] == 0;
            }
            t.trans[s1][s2] = c;
            const int step = c + t.cont[s2] + mm * (nm - 1);
            if (step > maxc) maxc = step;
        }
    t.maxSingleStep = maxc;
}

struct Entry {  // one U cell (U_cell_type :44) + the distance of its check-point cell (CP(...)->dist, :342)
    int32_t tag;                 // d + costOffset ("computed")
    int16_t dist, fdist;
    int16_t fab, fac, fcost, fstate;
};

struct Task {  // one doUkkInLimits call
    int sab, sac, sCost, sState, sDist, fab, fac, fCost, fState, fDist;
};

// Per-triple workspace (global memory on the device).  Box: diagonals |ab - cab| <= R, |ac - cac| <= R.
struct Work {
    const uint8_t *A, *B, *C;    // 0-based characters (any code; equal codes match)
    int Alen, Blen, Clen;
    int R, D, cab, cac, Wd;      // D = 2 R + 1; Wd a power of two
    Entry *U;                    // D * D * NS * Wd
    int *snap;                   // D * D * NS: top() saved before a strided step of the first pass
    int *top, *prev;             // D * D * NS: newest computed cost of the cell (NEGBIG: none), and the same one level earlier
    int *keycnt;                 // keycap = 2 * (maxlevels + 1) + 1 key counters, then one partial sum per thread (<= 1024)
    int keycap;
    int *list;                   // 2 ints per new cell: x, cost
    int listcap, maxlevels;
    uint8_t *resA, *resB, *resC; // alignment in reverse order ('-' = 0xff), capacity rescap
    int rescap;
    Task *stack;                 // capacity stackcap
    int stackcap;
    // --- scalars shared by the group (written by thread 0 between barriers, or by atomics)
    int status, nres, nstack;
    int changed, fr, lo_ab, hi_ab, lo_ac, hi_ac, nlist, win_k1, win_base;
    int s_lo_ab, s_hi_ab, s_lo_ac, s_hi_ac;
    long long costOffset;
    long long ncalc;             // cells computed (statistics)
    long long st_sweeps, st_sweep_cells, st_levels, st_tops;  // statistics (thread 0)
    long long nextOffset;        // first tag offset of the next triple run in this workspace (tags never repeat)
};

#ifndef PW_TRACE
#define PW_TRACE(...) ((void) 0)
#endif
#ifndef PW_TID
#define PW_TID 0
#define PW_NT 1
#define PW_SYNC() ((void) 0)
#define PW_ATOMIC_MAX(p, v) do { if (*(p) < (v)) *(p) = (v); } while (0)
#define PW_ATOMIC_MIN(p, v) do { if (*(p) > (v)) *(p) = (v); } while (0)
#define PW_ATOMIC_ADD(p, v) pw_host_fetch_add((p), (v))
inline int pw_host_fetch_add(int *p, int v) { const int o = *p; *p += v; return o; }
#endif

struct Engine {
    Work *w;
    const Tables *tb;
    // pass parameters (the globals of ukk.checkp.c :86-107)
    int sab, sac, sCost, sState;      // sabG, sacG, sCostG, sStateG
    int endA, endB, endC;
    int CPcost, CPwidth, completeFromInfo;
    int startx;                       // box index of the pass's start cell

    PW_HD bool in_box(int ab, int ac) const {
        return ab >= w->cab - w->R && ab <= w->cab + w->R && ac >= w->cac - w->R && ac <= w->cac + w->R;
    }
    PW_HD int xi(int ab, int ac, int s) const { return ((ab - w->cab + w->R) * w->D + (ac - w->cac + w->R)) * NS + s; }
    PW_HD Entry &cell(int x, int d) const { return w->U[(size_t) x * w->Wd + (d & (w->Wd - 1))]; }
    PW_HD static int iabs(int v) { return v < 0 ? -v : v; }

    // smallest cost at which withinMatrix(ab, ac, .) holds (:545-576)
    PW_HD int lower(int ab, int ac) const {
        int a0 = iabs(sab - ab), a1 = iabs(sac - ac), a2 = iabs((sac - sab) - (ac - ab));
        // g, h = the two smallest
        int g = a0 < a1 ? a0 : a1, mx = a0 < a1 ? a1 : a0;
        int h = mx < a2 ? mx : a2;
        if (a2 < g) { h = g; g = a2; }
        int cheapest;
        if (sState == 0) cheapest = (g == 0 ? 0 : tb->go + g * tb->ge) + (h == 0 ? 0 : tb->go + h * tb->ge);
        else cheapest = (g == 0 ? 0 : g * tb->ge) + (h == 0 ? 0 : h * tb->ge);
        const int lo = cheapest + sCost;
        return lo < 0 ? 0 : lo;
    }
    PW_HD bool within(int ab, int ac, int d) const { return d >= 0 && d >= lower(ab, ac); }
    PW_HD bool diag_ok(int ab, int ac) const { return ab >= -endB && ab <= endA && ac >= -endC && ac <= endA; }  // :644
    PW_HD static bool ok_index(int a, int da, int end) {  // okIndex, ukkCommon.c:187-193
        if (a < 0) return false;
        return da ? a < end : a <= end;
    }

    // Ukk(ab, ac, d, s) as a READ: the relaxation has made sure the cell exists whenever the reference would compute it.
    PW_HD int U(int ab, int ac, int d, int s) const {
        if (!within(ab, ac, d)) return -PW_INF;
        if (!in_box(ab, ac)) { w->status = PW_EBOX; return -PW_INF; }
        const Entry &e = cell(xi(ab, ac, s), d);
        if (e.tag != (int32_t) (d + w->costOffset)) { w->status = PW_EWINDOW; return -PW_INF; }
        return e.dist;
    }
    PW_HD void inherit(Entry &dst, int ab, int ac, int d, int s) const {  // from = U(ab, ac, d, s)->from
        const Entry &e = cell(xi(ab, ac, s), d);
        dst.fab = e.fab; dst.fac = e.fac; dst.fcost = e.fcost; dst.fstate = e.fstate; dst.fdist = e.fdist;
    }

    // Ukk(.., d, state) as a read of cell x (box index, valid when `inb`) whose diagonal enters the matrix at cost `lo`.  Errors
    // are collected in `err` and stored once per calc(): a store between the loads would serialise them.
    PW_HD int read(int x, int d, int lo, bool inb, int &err) const {
        if (d < lo) return -PW_INF;  // !withinMatrix (lo >= 0)
        if (!inb) { err |= 1 << PW_EBOX; return -PW_INF; }
        const Entry &e = cell(x, d);
        if (e.tag != (int32_t) (d + w->costOffset)) { err |= 1 << PW_EWINDOW; return -PW_INF; }
        return e.dist;
    }

    // calcUkk (:608-784) for cell x = (ab, ac, toState) at cost d.  The 16 from-states are independent of each other until
    // the final "first strict improvement wins" scan, so their reads are issued together (three rounds of loads instead of
    // sixteen chains of three).
    PW_HD void calc(int ab, int ac, int d, int toState) const {
        const uint8_t *A = w->A, *B = w->B, *C = w->C;
        Entry out;
        out.fab = 0; out.fac = 0; out.fcost = -1; out.fstate = 0; out.fdist = 0;
        int bestDist = -PW_INF, err = 0;
        const bool cpwin = d >= CPcost && d < CPcost + CPwidth;
        const bool cpinherit = !completeFromInfo && d >= CPcost + CPwidth;
        if (cpwin) { out.fab = (int16_t) ab; out.fac = (int16_t) ac; out.fcost = (int16_t) d; out.fstate = (int16_t) toState; }
        const int da = tb->da[toState], db = tb->db[toState], dc = tb->dc[toState];
        const int ab1 = ab - da + db, ac1 = ac - da + dc;
        int from_ab = 0, from_ac = 0, from_cost = 0, from_state = -1;  // the predecessor chosen so far (from_state < 0: none)
        if (diag_ok(ab1, ac1)) {
            const bool inb = in_box(ab1, ac1);
            const int lo1 = lower(ab1, ac1), xb = inb ? xi(ab1, ac1, 0) : 0;
            const int base = d - tb->cont[toState], mis = tb->mis;
            const bool second = tb->second[toState] != 0;
            int a1v[NS], a2v[NS];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int fs = 0; fs < NS; fs++) {
                const int cost = base - tb->trans[fs][toState];
                a1v[fs] = read(xb + fs, cost, lo1, inb, err);
                a2v[fs] = second ? read(xb + fs, cost - mis, lo1, inb, err) : -PW_INF;
            }
            unsigned firstmask = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int fs = 0; fs < NS; fs++) {
                const int a1 = a1v[fs];
                if (ok_index(a1, da, endA) && ok_index(a1 - ab1, db, endB) && ok_index(a1 - ac1, dc, endC)) {
                    // whichCharCost(...) == 1 (ukkCommon.c:155-184): not all three equal, but two of them are
                    const int ca = da ? A[a1] : 256, cb = db ? B[a1 - ab1] : 256, cc = dc ? C[a1 - ac1] : 256;
                    if (!(ca == cb && ca == cc) && (ca == cb || ca == cc || cb == cc)) firstmask |= 1u << fs;
                }
            }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int fs = 0; fs < NS; fs++) {  // in the reference's order: the first strict improvement wins (:679)
                const int cost = base - tb->trans[fs][toState];
                int fromCost = -PW_INF, dist = -PW_INF;
                if ((firstmask >> fs) & 1) {
                    fromCost = cost;
                    dist = a1v[fs] + da;
                } else if (second) {
                    const int a2 = a2v[fs];
                    if (ok_index(a2, da, endA) && ok_index(a2 - ab1, db, endB) && ok_index(a2 - ac1, dc, endC)) {
                        fromCost = cost - mis;
                        dist = a2 + da;
                    }
                }
                if (bestDist < dist) {
                    bestDist = dist;
                    from_ab = ab1; from_ac = ac1; from_cost = fromCost; from_state = fs;
                }
            }
        }
        const bool inb0 = in_box(ab, ac);  // true: the cell itself is in the box
        const int lo0 = lower(ab, ac), x0 = xi(ab, ac, 0);
        {   // what can be reached for AT MOST cost d (:693-711)
            const int dist = read(x0 + toState, d - 1, lo0, inb0, err);
            if (ok_index(dist, 0, endA) && ok_index(dist - ab, 0, endB) && ok_index(dist - ac, 0, endC) && bestDist < dist) {
                bestDist = dist;
                from_ab = ab; from_ac = ac; from_cost = d - 1; from_state = toState;
            }
        }
        if (toState == 0) {  // extend along a run of matches from the furthest state of this cell (:713-764)
            int sv[NS];
            sv[0] = bestDist;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int st = 1; st < NS; st++) sv[st] = read(x0 + st, d, lo0, inb0, err);
            int dist = -PW_INF, best_state = -1;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int st = 0; st < NS; st++)
                if (sv[st] > dist) { dist = sv[st]; best_state = st; }
            while (ok_index(dist, 1, endA) && ok_index(dist - ab, 1, endB) && ok_index(dist - ac, 1, endC) && A[dist] == B[dist - ab] &&
                   A[dist] == C[dist - ac])
                dist++;
            if (dist > bestDist) {
                bestDist = dist;
                if (best_state != 0) { from_ab = ab; from_ac = ac; from_cost = d; from_state = best_state; }
            }
        }
        if (from_state >= 0) {
            if (completeFromInfo) {
                out.fab = (int16_t) from_ab; out.fac = (int16_t) from_ac; out.fcost = (int16_t) from_cost; out.fstate = (int16_t) from_state;
            } else if (cpinherit) {
                inherit(out, from_ab, from_ac, from_cost, from_state);
            }
        }
        out.dist = (int16_t) bestDist;
        if (cpwin && out.fab == ab && out.fac == ac && out.fcost == d && out.fstate == toState) out.fdist = (int16_t) bestDist;  // CP(...)->dist (:592-595)
        out.tag = (int32_t) (d + w->costOffset);
        cell(x0 + toState, d) = out;
        if (bestDist > w->fr) PW_ATOMIC_MAX(&w->fr, bestDist);  // furthestReached (:597)
        if (err) w->status = (err & (1 << PW_EBOX)) ? PW_EBOX : PW_EWINDOW;
    }

    PW_HD void decode(int x, int &ab, int &ac, int &st) const {
        st = x % NS;
        const int q = x / NS;
        ab = q / w->D - w->R + w->cab;
        ac = q % w->D - w->R + w->cac;
    }
    // top of x as a source of demands: the preset start cell is a memo hit, never expanded (:242-243, :582)
    PW_HD int src_top(int x) const {
        const int t = w->top[x];
        return (x == startx && t <= sCost) ? NEGBIG : t;
    }

    // Top level T - 1 -> T, step 1: every cell that was demanded keeps being demanded one cost higher (the same path).
    PW_HD void bump(int inc) const {
        const int lo_ab = w->lo_ab, hi_ab = w->hi_ab, lo_ac = w->lo_ac, hi_ac = w->hi_ac;
        const int nac = hi_ac - lo_ac + 1, total = (hi_ab - lo_ab + 1) * nac * NS;
        for (int k = PW_TID; k < total; k += PW_NT) {
            const int st = k % NS, q = k / NS;
            const int x = xi(lo_ab + q / nac, lo_ac + q % nac, st);
            const int t = w->top[x];
            w->prev[x] = t;
            if (t != NEGBIG && !(x == startx && t <= sCost)) w->top[x] = t + inc;
        }
    }
    // Step 2, repeated until nothing changes: one pull sweep of the demand closure over the bounding box of the active cells
    // (+ 1 diagonal); r0 / r1 = the cells the top level asks for directly (box indices, -1 = none).  In place: the closure is
    // a monotone fixpoint, the order of the updates does not matter.
    PW_HD void sweep(int T, int r0, int r1) const {
        const int lo_ab = w->lo_ab - 1, hi_ab = w->hi_ab + 1, lo_ac = w->lo_ac - 1, hi_ac = w->hi_ac + 1;
        const int nac = hi_ac - lo_ac + 1, total = (hi_ab - lo_ab + 1) * nac * NS;
        for (int k = PW_TID; k < total; k += PW_NT) {
            const int st = k % NS, q = k / NS;
            const int ab = lo_ab + q / nac, ac = lo_ac + q % nac;
            if (!in_box(ab, ac)) continue;
            const int y = xi(ab, ac, st);
            const int cur = w->top[y];
            int cand = NEGBIG;
            if (y == r0 || y == r1) cand = T;
            if (diag_ok(ab, ac)) {
                for (int ts = 0; ts < NS; ts++) {  // the cells whose neighbour diagonal this is
                    const int xab = ab + tb->da[ts] - tb->db[ts], xac = ac + tb->da[ts] - tb->dc[ts];
                    if (!in_box(xab, xac)) continue;
                    const int tx = src_top(xi(xab, xac, ts));
                    if (tx == NEGBIG) continue;
                    const int c = tx - (tb->trans[st][ts] + tb->cont[ts]);
                    if (c > cand) cand = c;
                }
            }
            if (st != 0) {  // MMM asks for the other states of its own cell (:732-739)
                const int tx = src_top(xi(ab, ac, 0));
                if (tx > cand) cand = tx;
            }
            if (cand == NEGBIG || cand <= cur) continue;
            if (cur == NEGBIG && cand < lower(ab, ac)) continue;  // outside the matrix: the demand returns -INFINITY (:581)
            w->top[y] = cand;
            w->changed = 1;
            if (ab - w->cab == -w->R || ab - w->cab == w->R || ac - w->cac == -w->R || ac - w->cac == w->R) w->status = PW_EBOX;
            PW_ATOMIC_MIN(&w->lo_ab, ab); PW_ATOMIC_MAX(&w->hi_ab, ab);
            PW_ATOMIC_MIN(&w->lo_ac, ac); PW_ATOMIC_MAX(&w->hi_ac, ac);
        }
    }
    // The cells that became demanded at this top level, as (x, cost) sorted by (cost, MMM last).  count = false: scatter.
    PW_HD void collect(bool scatter, int k0, int k1, int base) const {
        const int lo_ab = w->lo_ab, hi_ab = w->hi_ab, lo_ac = w->lo_ac, hi_ac = w->hi_ac;
        const int nac = hi_ac - lo_ac + 1, total = (hi_ab - lo_ab + 1) * nac * NS;
        for (int k = PW_TID; k < total; k += PW_NT) {
            const int st = k % NS, q = k / NS;
            const int ab = lo_ab + q / nac, ac = lo_ac + q % nac;
            const int x = xi(ab, ac, st);
            const int hi = w->top[x];
            if (hi == NEGBIG) continue;
            const int pv = w->prev[x];
            int lo = (pv == NEGBIG) ? lower(ab, ac) : pv + 1;
            for (int c = lo; c <= hi; c++) {
                const int key = 2 * (c - sCost) + (st == 0 ? 1 : 0);
                if (key < k0 || key >= k1) continue;
                const int pos = PW_ATOMIC_ADD(&w->keycnt[key], 1) - base;
                if (scatter) { w->list[2 * pos] = x; w->list[2 * pos + 1] = c; }
            }
        }
    }

    // All cells the reference computes for the top-level calls Ukk(root, T, .): relax, list, compute in cost order.
    PW_HD void top_level(int T, int r0, int r1, int inc) const {
        relax(T, r0, r1, inc);
        if (w->status) return;
        compute_new(T);
    }
    // The demand closure of top level T, `inc` costs after the previous one.
    PW_HD void relax(int T, int r0, int r1, int inc) const {
        if (T - sCost >= w->maxlevels) { if (PW_TID == 0) w->status = PW_ECAP; PW_SYNC(); return; }
        if (PW_TID == 0) w->st_tops++;
        bump(inc);
        PW_SYNC();
        for (;;) {
            if (PW_TID == 0) {
                w->changed = 0;
                w->st_sweeps++;
                w->st_sweep_cells += (long long) (w->hi_ab - w->lo_ab + 3) * (w->hi_ac - w->lo_ac + 3) * NS;
            }
            PW_SYNC();
            sweep(T, r0, r1);
            PW_SYNC();
            if (!w->changed || w->status) break;
            PW_SYNC();
        }
    }
    // The cells between prev() and top(), computed in order of their cost.
    PW_HD void compute_new(int T) const {
        const int nkeys = 2 * (T - sCost + 1);
        for (int k = PW_TID; k < nkeys; k += PW_NT) w->keycnt[k] = 0;
        PW_SYNC();
        collect(false, 0, nkeys, 0);
        PW_SYNC();
        {   // exclusive prefix over the keys, all threads: a contiguous chunk each, partial sums behind the keys
            int *part = w->keycnt + w->keycap;
            const int chunk = (nkeys + PW_NT - 1) / PW_NT;
            const int klo = PW_TID * chunk < nkeys ? PW_TID * chunk : nkeys, khi = klo + chunk < nkeys ? klo + chunk : nkeys;
            int sum = 0;
            for (int k = klo; k < khi; k++) sum += w->keycnt[k];
            part[PW_TID] = sum;
            PW_SYNC();
            int base = 0, total = 0;
            for (int u = 0; u < PW_NT; u++) { const int v = part[u]; if (u < PW_TID) base += v; total += v; }
            for (int k = klo; k < khi; k++) { const int c = w->keycnt[k]; w->keycnt[k] = base; base += c; }
            if (PW_TID == 0) {
                w->nlist = total;
                w->ncalc += total;
            }
        }
        PW_SYNC();
        // The list holds listcap cells: the keys are taken in windows [k0, k1) that fit (one window when the top levels are
        // stepped one by one; a pass computed in one go needs several).
        for (int k0 = 0; k0 < nkeys && !w->status;) {
            if (PW_TID == 0) {
                const int base = w->keycnt[k0];  // exclusive offset of k0 (keys >= k0 still hold their start offsets)
                int k1 = k0 + 1;  // start(k) = keycnt[k] for k < nkeys, the total for k = nkeys
                if ((k1 < nkeys ? w->keycnt[k1] : w->nlist) - base > w->listcap) w->status = PW_ELIST;
                else
                    while (k1 < nkeys && (k1 + 1 < nkeys ? w->keycnt[k1 + 1] : w->nlist) - base <= w->listcap) k1++;
                w->win_k1 = k1;
                w->win_base = base;
            }
            PW_SYNC();
            const int k1 = w->win_k1, base = w->win_base;
            if (w->status) break;
            collect(true, k0, k1, base);
            PW_SYNC();
            // keycnt[k] is now the END of key k for k0 <= k < k1 (and for every earlier key)
            for (int k = k0; k < k1; k++) {
                const int begin = (k == 0 ? 0 : w->keycnt[k - 1]) - base, end = w->keycnt[k] - base;
                if (end == begin) continue;
                if (PW_TID == 0) w->st_levels++;
                for (int i = begin + PW_TID; i < end; i += PW_NT) {
                    int ab, ac, st;
                    decode(w->list[2 * i], ab, ac, st);
                    calc(ab, ac, w->list[2 * i + 1], st);
                }
                PW_SYNC();
            }
            k0 = k1;
        }
    }

    // Demand state saved / restored around a strided step of the first pass (see run()).
    PW_HD void snapshot_save() const {
        const int lo_ab = w->lo_ab, hi_ab = w->hi_ab, lo_ac = w->lo_ac, hi_ac = w->hi_ac;
        const int nac = hi_ac - lo_ac + 1, total = (hi_ab - lo_ab + 1) * nac * NS;
        for (int k = PW_TID; k < total; k += PW_NT) {
            const int st = k % NS, q = k / NS;
            const int x = xi(lo_ab + q / nac, lo_ac + q % nac, st);
            w->snap[x] = w->top[x];
        }
        if (PW_TID == 0) { w->s_lo_ab = lo_ab; w->s_hi_ab = hi_ab; w->s_lo_ac = lo_ac; w->s_hi_ac = hi_ac; }
        PW_SYNC();
    }
    PW_HD void snapshot_restore() const {
        const int lo_ab = w->lo_ab, hi_ab = w->hi_ab, lo_ac = w->lo_ac, hi_ac = w->hi_ac;  // the larger, current region
        const int nac = hi_ac - lo_ac + 1, total = (hi_ab - lo_ab + 1) * nac * NS;
        PW_SYNC();
        for (int k = PW_TID; k < total; k += PW_NT) {
            const int st = k % NS, q = k / NS;
            const int ab = lo_ab + q / nac, ac = lo_ac + q % nac;
            const int x = xi(ab, ac, st);
            const bool saved = ab >= w->s_lo_ab && ab <= w->s_hi_ab && ac >= w->s_lo_ac && ac <= w->s_hi_ac;
            w->top[x] = saved ? w->snap[x] : NEGBIG;
            w->prev[x] = w->top[x];
        }
        PW_SYNC();
        if (PW_TID == 0) { w->lo_ab = w->s_lo_ab; w->hi_ab = w->s_hi_ab; w->lo_ac = w->s_lo_ac; w->hi_ac = w->s_hi_ac; }
        PW_SYNC();
    }
    // Largest U among the cells between prev() and top() -- already computed (replay of a strided step) -- into w->fr.
    PW_HD void scan_new() const {
        const int lo_ab = w->lo_ab, hi_ab = w->hi_ab, lo_ac = w->lo_ac, hi_ac = w->hi_ac;
        const int nac = hi_ac - lo_ac + 1, total = (hi_ab - lo_ab + 1) * nac * NS;
        for (int k = PW_TID; k < total; k += PW_NT) {
            const int st = k % NS, q = k / NS;
            const int ab = lo_ab + q / nac, ac = lo_ac + q % nac;
            const int x = xi(ab, ac, st);
            const int hi = w->top[x];
            if (hi == NEGBIG) continue;
            const int pv = w->prev[x];
            const int lo = (pv == NEGBIG) ? lower(ab, ac) : pv + 1;
            if (lo > hi) continue;
            const Entry &e = cell(x, hi);  // U grows with the cost on a diagonal: the top cell of the range holds the maximum
            if (e.tag != (int32_t) (hi + w->costOffset)) { w->status = PW_EWINDOW; continue; }
            if (e.dist > w->fr) PW_ATOMIC_MAX(&w->fr, (int) e.dist);
        }
        PW_SYNC();
    }

    // Start of a pass: the preset start cell and an empty demand state.  Thread 0, between barriers.
    PW_HD void begin_pass(const Task &t) {
        sab = t.sab; sac = t.sac; sCost = t.sCost; sState = t.sState;
        startx = xi(sab, sac, sState);
        if (PW_TID == 0) {
            const int R1 = w->R - 1;  // strictly inside: a cell on the border cannot see its outer neighbours
            if (iabs(sab - w->cab) > R1 || iabs(sac - w->cac) > R1 || iabs(t.fab - w->cab) > R1 || iabs(t.fac - w->cac) > R1) w->status = PW_EBOX;
            else {
                Entry &e = cell(startx, sCost);
                e.dist = (int16_t) t.sDist;
                e.tag = (int32_t) (sCost + w->costOffset);
                w->top[startx] = sCost;
                w->prev[startx] = sCost;
                // sweep region: hull of the start and the final diagonal (the roots must be inside it to be asked at all)
                w->lo_ab = sab < t.fab ? sab : t.fab; w->hi_ab = sab < t.fab ? t.fab : sab;
                w->lo_ac = sac < t.fac ? sac : t.fac; w->hi_ac = sac < t.fac ? t.fac : sac;
            }
        }
        PW_SYNC();
    }
    // End of a pass: forget the demand state of the region it touched.
    PW_HD void end_pass() const {
        const int lo_ab = w->lo_ab, hi_ab = w->hi_ab, lo_ac = w->lo_ac, hi_ac = w->hi_ac;
        const int nac = hi_ac - lo_ac + 1, total = (hi_ab - lo_ab + 1) * nac * NS;
        PW_SYNC();
        for (int k = PW_TID; k < total; k += PW_NT) {
            const int st = k % NS, q = k / NS;
            const int x = xi(lo_ab + q / nac, lo_ac + q % nac, st);
            w->top[x] = NEGBIG;
            w->prev[x] = NEGBIG;
        }
        PW_SYNC();
    }
    PW_HD bool computed(int ab, int ac, int d, int st) const {
        return in_box(ab, ac) && cell(xi(ab, ac, st), d).tag == (int32_t) (d + w->costOffset);
    }
    // best() :508-527
    PW_HD int best(int ab, int ac, int d, bool wantState) const {
        int bst = -PW_INF, bs = -1;
        for (int st = 0; st < NS; st++)
            if (computed(ab, ac, d, st) && cell(xi(ab, ac, st), d).dist > bst) { bst = cell(xi(ab, ac, st), d).dist; bs = st; }
        return wantState ? bs : bst;
    }
    PW_HD void push_res(int a, int b, int c) const {
        if (w->nres >= w->rescap) { w->status = PW_ECAP; return; }
        w->resA[w->nres] = (uint8_t) a; w->resB[w->nres] = (uint8_t) b; w->resC[w->nres] = (uint8_t) c;
        w->nres++;
    }
    // traceBack :370-444 (thread 0)
    PW_HD void trace_back(const Task &t) const {
        int ab = t.fab, ac = t.fac, d = t.fCost, st = t.fState;
        int guard = 0;
        while (ab != t.sab || ac != t.sac || d != t.sCost || st != t.sState) {
            if (!computed(ab, ac, d, st) || ++guard > 4 * w->rescap + 64) { w->status = PW_EWINDOW; return; }
            const Entry e = cell(xi(ab, ac, st), d);
            int a = e.dist;
            const int nab = e.fab, nac = e.fac, nd = e.fcost, ns = e.fstate;
            if (nd < 0 || ns < 0 || !computed(nab, nac, nd, ns)) { w->status = PW_EWINDOW; return; }
            int b = a - ab, c = a - ac;
            const int a1 = cell(xi(nab, nac, ns), nd).dist;
            const int b1 = a1 - nab, c1 = a1 - nac;
            while (a > a1 && b > b1 && c > c1) {  // run of matches
                a--; b--; c--;
                push_res(w->A[a], w->B[b], w->C[c]);
            }
            if (a != a1 || b != b1 || c != c1) {  // the step (nab, nac, nd, ns) -> (ab, ac, d, st)
                const int ra = a > a1 ? w->A[--a] : 0xff, rb = b > b1 ? w->B[--b] : 0xff, rc = c > c1 ? w->C[--c] : 0xff;
                push_res(ra, rb, rc);
            }
            if (w->status) return;
            ab = nab; ac = nac; d = nd; st = ns;
        }
    }
    PW_HD void push_task(const Task &t) const {
        if (w->nstack >= w->stackcap) { w->status = PW_ESTACK; return; }
        w->stack[w->nstack++] = t;
    }
    // getSplitRecurse :324-367: second half first (popped first), then the first half
    PW_HD void split(const Task &t) const {
        const Entry e = cell(xi(t.fab, t.fac, t.fState), t.fCost);
        if (e.fcost < 0) { w->status = PW_EWINDOW; return; }
        Task first = t, second = t;
        first.fab = e.fab; first.fac = e.fac; first.fCost = e.fcost; first.fState = e.fstate; first.fDist = e.fdist;
        second.sab = e.fab; second.sac = e.fac; second.sCost = e.fcost; second.sState = e.fstate; second.sDist = e.fdist;
        push_task(first);
        push_task(second);
    }

    // doUkkInLimits :216-322
    PW_HD void in_limits(Task t, int finalCost) {
        endA = t.fDist; endB = t.fDist - t.fab; endC = t.fDist - t.fac;
        completeFromInfo = 0;
        if (PW_TID == 0) w->costOffset += finalCost + 1;
        PW_SYNC();
        begin_pass(t);
        PW_TRACE("in_limits s=(%d,%d,c%d,s%d,d%d) f=(%d,%d,c%d,s%d,d%d) status %d\n", t.sab, t.sac, t.sCost, t.sState, t.sDist, t.fab, t.fac,
                 t.fCost, t.fState, t.fDist, w->status);
        if (w->status) return;
        const bool base = t.fCost - t.sCost <= CPwidth;
        if (base) completeFromInfo = 1;
        else CPcost = (t.fCost + t.sCost - CPwidth + 1) / 2;
        const int rf = xi(t.fab, t.fac, t.fState), r0 = base ? -1 : xi(t.fab, t.fac, 0);
        // The reference steps i = sCost, sCost + 1, ... until Ukk(f, i, fState) reaches fDist, which happens at i = fCost
        // (:261-271, :305-316).  The cells it has computed by then are those of the last step alone (top() only grows), and
        // their values do not depend on the stepping, so the closure is taken once, at fCost, and the cells are computed in
        // one sweep over the cost levels: 2 (fCost - sCost) block-wide levels instead of one set of levels per step.
        int T = t.fCost;
        top_level(T, r0, rf, T - t.sCost);
        if (w->status) return;
        int reach = -1;
        {
            int c = T - w->Wd / 2;
            if (c < t.sCost) c = t.sCost;
            for (; c <= T; c++)
                if (U(t.fab, t.fac, c, t.fState) >= t.fDist) { reach = c; break; }
        }
        while (reach < 0 && !w->status) {  // not reached at fCost: go on like the reference would
            T++;
            top_level(T, r0, rf, 1);
            if (w->status) return;
            if (U(t.fab, t.fac, T, t.fState) >= t.fDist) reach = T;
        }
        if (w->status) return;
        PW_TRACE("  reached at %d (expected %d)\n", reach, t.fCost);
        t.fCost = reach;  // `if (i != fCost) ... fCost = i` (:267-271, :312-316)
        if (PW_TID == 0) {
            if (base) trace_back(t);
            else split(t);
        }
        end_pass();
    }

    // doUkk :113-214.  Returns the cost; the alignment is left reversed in resA / resB / resC ('-' = 0xff).
    PW_HD int run() {
        const int Alen = w->Alen, Blen = w->Blen, Clen = w->Clen;
        CPwidth = tb->maxSingleStep;
        CPcost = 0;
        completeFromInfo = 0;
        int startDist = 0;
        while (startDist < Alen && startDist < Blen && startDist < Clen && w->A[startDist] == w->B[startDist] &&
               w->A[startDist] == w->C[startDist])
            startDist++;
        // (the reference compares against the terminating 0 of the shorter strings: same stop)
        const int finalab = Alen - Blen, finalac = Alen - Clen;
        endA = Alen; endB = Blen; endC = Clen;
        if (PW_TID == 0) { w->costOffset = w->nextOffset; w->fr = -1; w->nres = 0; w->nstack = 0; }  // the reference starts at 1 (:86)
        PW_SYNC();
        Task t0{0, 0, 0, 0, startDist, finalab, finalac, 0, 0, Alen};
        begin_pass(t0);
        if (w->status) return -1;
        if (PW_TID == 0) {  // fresh `from` of the very first cell: calloc'ed zeros (:152-153)
            Entry &e = cell(startx, 0);
            e.fab = e.fac = e.fcost = e.fstate = e.fdist = 0;
        }
        PW_SYNC();
        bool CPonDist = true;
        CPcost = PW_INF;
        const int rf = xi(finalab, finalac, 0);
        // Until the check-point is placed the top levels are stepped one by one: furthestReached must be what the recursion
        // would have seen after each of them.  Afterwards only "which cost reaches the end first" matters, the values do not
        // depend on the stepping, and the closure is advanced PW_STRIDE costs at a time (a few cells past the final cost get
        // computed that the reference never asks for; nothing reads them).
        int d = -1;
        bool single = false;
        for (bool done = false; !done;) {
            // A strided step asks for a few cells the reference never computes; if that alone pushes the closure onto the
            // border of the diagonal box, the step is undone (only its relaxation has run) and one level is stepped instead.
            const int stride = single ? 1 : PW_STRIDE;
            const int d0 = d, dn = d + stride, fr0 = w->fr;
            if (stride > 1) snapshot_save();
            top_level(dn, rf, -1, stride);
            if (w->status == PW_EBOX && stride > 1) {
                snapshot_restore();
                if (PW_TID == 0) w->status = PW_OK;
                PW_SYNC();
                single = true;
                continue;
            }
            if (w->status) return -1;
            single = false;
            if (CPonDist) {
                // If furthestReached is still short of |A| / 2 after the step, it was short after every top level inside it as
                // well.  Otherwise the step is replayed level by level on the values just computed (relaxation only) to find
                // the level at which the reference places its check-point; the demand state is then exactly the one of that
                // level, and the cells above it are computed again with the check-point in force.
                bool hit = w->fr >= Alen / 2;  // ... or the end of A is reached inside the step (it can be without any computed
                for (int c = d0 + 1; c <= dn && !hit; c++)  // cell: identical sequences end on the preset start cell)
                    hit = best(finalab, finalac, c, false) >= Alen;
                if (!hit) { d = dn; continue; }
                if (stride == 1) {  // the reference's order per level: compute, place the check-point, test the end
                    d = dn;
                    if (w->fr >= Alen / 2) { CPcost = d + 1; CPonDist = false; }
                    done = best(finalab, finalac, d, false) >= Alen;
                    continue;
                }
                snapshot_restore();
                if (PW_TID == 0) w->fr = fr0;
                PW_SYNC();
                for (d = d0 + 1; d <= dn; d++) {
                    relax(d, rf, -1, 1);
                    if (w->status) return -1;
                    scan_new();
                    if (w->status) return -1;
                    if (w->fr >= Alen / 2) { CPcost = d + 1; CPonDist = false; }
                    done = best(finalab, finalac, d, false) >= Alen;
                    if (done || !CPonDist) break;
                }
                if (d > dn) { if (PW_TID == 0) w->status = PW_EWINDOW; PW_SYNC(); return -1; }  // cannot happen
                PW_TRACE("pass1 replay stopped at d=%d fr=%d CPcost=%d done=%d\n", d, w->fr, CPcost, (int) done);
            } else {
                for (int c = d0 + 1; c <= dn && !done; c++)
                    if (best(finalab, finalac, c, false) >= Alen) { d = c; done = true; }
                if (!done) d = dn;
            }
        }
        const int finalCost = d;
        const int fState = best(finalab, finalac, finalCost, true);
        Task whole{0, 0, 0, 0, startDist, finalab, finalac, finalCost, fState, Alen};
        const bool redo = cell(xi(finalab, finalac, fState), finalCost).fcost <= 0;  // check-pointed too late (:195-200)
        PW_SYNC();
        if (PW_TID == 0) {
            if (redo) push_task(whole);
            else split(whole);
        }
        end_pass();
        while (!w->status && w->nstack > 0) {
            const Task t = w->stack[w->nstack - 1];
            PW_SYNC();
            if (PW_TID == 0) w->nstack--;
            PW_SYNC();
            in_limits(t, finalCost);
        }
        if (w->status) return -1;
        if (PW_TID == 0)  // printTraceBack :456-475: the first run of matches, in reverse order like the rest
            for (int i = startDist - 1; i >= 0; i--) push_res(w->A[i], w->B[i], w->C[i]);
        PW_SYNC();
        return w->status ? -1 : finalCost;
    }
};

}  // namespace powell
}  // namespace poyb200
