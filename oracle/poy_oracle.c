/* TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Plain-C restatement ("port") of the reference's direct-optimisation alignment path, written from the
 * recurrences and NOT from its control flow: where the reference walks five row-phase helpers over rolling
 * rows, this file describes the visited region as a diagonal stripe dlo <= j-i <= dhi plus per-cell rules,
 * which is also how the CUDA kernels see it.  Each function cites the reference lines it restates.
 *
 * PARITY PINNED: tests/test_oracle.py checks every entry point here bit-for-bit (costs,
 * direction-dependent outputs, medians) against oracle/_ref/libpoyref.so, i.e. the unmodified
 * /root/reference/src/algn.c compiled in this container, on seeded random inputs; the committed
 * tests/golden/ vectors were produced by that compiled reference (tests/golden/make_golden.py).  The checkers are also
 * pinned against numbers the reference's authors recorded: driven by poyd_b200/tree.py they reproduce the 52 tree costs of
 * the reference's test/cost_tests (tests/golden/trees/tree_costs.npz, tests/test_tree.py).
 * The reference repository itself holds no per-pair golden vectors for this path.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define PO_HIGH 1000000 /* HIGH_NUM, src/algn.c:38 */

/* 2-D linear direction bits, src/matrices.h:22-27 */
#define PO_ALIGN 1
#define PO_INSERT 2
#define PO_DELETE 4

/* affine_3 direction bits, src/algn.c:1695-1708 */
#define A2A 1
#define A2V 2
#define A2H 4
#define A2D 8
#define BEG_B 16
#define END_B 32
#define BEG_V 64
#define END_V 128
#define BEG_H 256
#define END_H 512
#define DO_A 1024
#define DO_V 2048
#define DO_H 4096
#define DO_D 8192
#define TGAP 16 /* TMPGAP, src/algn.c:1711 */

typedef struct po_cm {
    int32_t a_sz, lcm, gap, cost_model_type, combinations, gap_open;
    const int32_t *cost;     /* (1<<lcm)^2, index (a<<lcm)+b   src/cm.c:545 */
    const uint8_t *median;   /* same indexing                 src/cm.c:529 */
    const int32_t *prepend;  /* 1<<lcm                         src/cm.c:704 */
    const int32_t *tail;     /* 1<<lcm                         src/cm.c:711 */
} po_cm;

static inline int po_cost(const po_cm *c, int a, int b) { return c->cost[(a << c->lcm) + b]; }
static inline int po_med(const po_cm *c, int a, int b) { return c->median[(a << c->lcm) + b]; }

/* ------------------------------------------------------------------------------------------------------
 * Linear gaps: region visited by algn_fill_plane_2 (src/algn.c:872-968).
 *   full != 0 : every cell (Case 1 `l1 >= 1.5 l2` :893, Case 3a `l1 - height <= 8` :936) -> algn_fill_plane
 *   else      : stripe dlo <= j-i <= dhi.  Case 2 (:898-929): dlo = 1-height, Case 3b (:938-965):
 *               dlo = -(l1-l2+width); dhi = width-1 in both.
 */
typedef struct po_band { int full, dlo, dhi; } po_band;

po_band po_linear_band(int l1, int l2, int deltaw) {
    po_band b;
    int width = 50 + deltaw, height = (l1 - l2) + 50 + deltaw; /* algn_nw_limit :3265-3266, plane_2 :880-883 */
    if (width > l2) width = l2;
    if (height > l1) height = l1;
    b.full = 0; b.dlo = 0; b.dhi = width - 1;
    if ((float) l1 >= ((float) 3 / (float) 2) * (float) l2) b.full = 1;
    else if (2 * height < l1) b.dlo = 1 - height;
    else if (8 >= l1 - height) b.full = 1;
    else b.dlo = -((l1 - l2) + width);
    if (b.full) { b.dlo = -(l1 - 1); b.dhi = l2 - 1; }
    return b;
}

/* Number of DP cells the reference writes for one linear pair (SURVEY.md 8d "cells()"). */
long long po_cells_linear(int l1, int l2, int deltaw) {
    po_band b = po_linear_band(l1, l2, deltaw);
    if (b.full) return (long long) l1 * l2;
    long long n = 0;
    for (int i = 0; i < l1; i++) {
        int lo = i + b.dlo, hi = i + b.dhi;
        if (lo < 0) lo = 0;
        if (hi > l2 - 1) hi = l2 - 1;
        if (hi >= lo) n += hi - lo + 1;
    }
    return n;
}

/* Banded / full linear-gap fill.  s1 is the longer sequence (rows), both include the leading gap.
 * dir: l1*l2 uint16 (row stride l2) or NULL.  Cell rules restate algn_fill_row (:375-431),
 * algn_fill_ukk_right_cell (:461-479), algn_fill_ukk_left_cell (:507-524), algn_fill_last_column (:548-560),
 * algn_fill_first_row (:584-607), algn_fill_first_cell (:610-612) and algn_fill_full_row (:567-581). */
int po_cost_2(const po_cm *c, const uint8_t *s1, int l1, const uint8_t *s2, int l2, int deltaw, uint16_t *dir) {
    po_band b = po_linear_band(l1, l2, deltaw);
    int gap = c->gap;
    int *prev = (int *) malloc(sizeof(int) * (size_t) l2), *cur = (int *) malloc(sizeof(int) * (size_t) l2);
    /* row 0: only `width` cells in banded mode (:899, :939), all of them in full mode (:831) */
    int row0_hi = b.full ? l2 - 1 : b.dhi;
    prev[0] = 0;
    if (dir) dir[0] = PO_ALIGN;
    for (int j = 1; j <= row0_hi; j++) {
        prev[j] = prev[j - 1] + c->prepend[s2[j]];
        if (dir) dir[j] = PO_INSERT;
    }
    for (int i = 1; i < l1; i++) {
        int a = s1[i];
        int del = po_cost(c, a, gap); /* const_val :653 */
        int lo = i + b.dlo, hi = i + b.dhi;
        if (lo < 0) lo = 0;
        if (hi > l2 - 1) hi = l2 - 1;
        for (int j = lo; j <= hi; j++) {
            int d = j - i, v, m;
            if (j == 0) {
                /* full rows add cost(a, gap) (:570); banded rows add alg_row[0] = tail[a] (:659, cm.c:711) */
                v = prev[0] + (b.full ? del : c->tail[a]);
                m = PO_DELETE;
            } else {
                int t_ins = 0, t_del = 0, t_al = prev[j - 1] + po_cost(c, a, s2[j]);
                int use_ins = 1, use_del = 1;
                if (!b.full && d == b.dhi) use_del = 0;      /* right edge: no cell above */
                else if (!b.full && d == b.dlo) use_ins = 0; /* left edge: no cell to the left */
                v = t_al;
                if (use_ins) { t_ins = cur[j - 1] + po_cost(c, gap, s2[j]); if (t_ins < v) v = t_ins; }
                if (use_del) { t_del = prev[j] + del; if (t_del < v) v = t_del; }
                m = (t_al == v) ? PO_ALIGN : 0;
                if (use_ins && t_ins == v) m |= PO_INSERT;
                if (use_del && t_del == v) m |= PO_DELETE;
                if (j == l2 - 1 && (b.full || d != b.dhi)) {
                    int cst = c->tail[a] + prev[j]; /* last column :548-560 */
                    if (cst < v) { v = cst; m = PO_DELETE; }
                    else if (cst == v) m |= PO_DELETE;
                }
            }
            cur[j] = v;
            if (dir) dir[(size_t) i * l2 + j] = (uint16_t) m;
        }
        int *t = prev; prev = cur; cur = t;
    }
    int res = prev[l2 - 1];
    free(prev); free(cur);
    return res;
}

/* backtrack_2d, linear branch (src/algn.c:3606-3665).  r1/r2: capacity l1+l2, filled from the right;
 * returns the aligned length, data left-aligned on return. */
int po_backtrack_2d(const po_cm *c, const uint8_t *s1, int l1, const uint8_t *s2, int l2, const uint16_t *dir,
                    int swaped, uint8_t *r1, uint8_t *r2) {
    int cap = l1 + l2, n = 0, i = l1 - 1, j = l2 - 1;
    long long pos = (long long) i * l2 + j;
    uint8_t gap = (uint8_t) c->gap;
    while (pos >= 0) {
        int m = dir[pos];
        int second = swaped ? PO_INSERT : PO_DELETE;
        int mv;
        if (m & PO_ALIGN) mv = PO_ALIGN;
        else if (m & second) mv = second;
        else mv = swaped ? PO_DELETE : PO_INSERT;
        n++;
        if (mv == PO_ALIGN) { r1[cap - n] = s1[i]; r2[cap - n] = s2[j]; i--; j--; pos -= l2 + 1; }
        else if (mv == PO_INSERT) { r1[cap - n] = gap; r2[cap - n] = s2[j]; j--; pos -= 1; }
        else { r1[cap - n] = s1[i]; r2[cap - n] = gap; i--; pos -= l2; }
    }
    memmove(r1, r1 + cap - n, n);
    memmove(r2, r2 + cap - n, n);
    return n;
}

/* algn_remove_gaps + the state machine of algn_correct_blocks_affine (src/algn.c:4058-4124). */
static int po_correct_blocks_affine(int gap, uint8_t *s, int len, const uint8_t *a, const uint8_t *b) {
    int extending_gap = 0, inside_block = 0, prev_block = 0;
    for (int i = 0; i < len; i++) {
        int ab = a[i], bb = b[i], sb = s[i];
        if (!inside_block && (!(ab & gap) || !(bb & gap))) inside_block = 0;
        else if (inside_block && (!(ab & gap) || !(bb & gap))) inside_block = 0;
        else if (((ab & gap) || (bb & gap)) && ((ab != gap) || (bb != gap))) inside_block = 1;
        else inside_block = 0;
        if (((gap & ab) || (gap & bb)) && !(sb & gap) && !extending_gap) {
            prev_block = inside_block;
            extending_gap = 1;
        } else if ((gap & ab) && (gap & bb) && (sb & gap) && (sb != gap) && extending_gap && inside_block &&
                   !prev_block) {
            sb = (~gap) & sb;
            prev_block = 0;
        } else if ((gap & ab) && (gap & bb) && (1 == extending_gap)) {
            prev_block = inside_block;
            extending_gap = 0;
        }
        s[i] = (uint8_t) sb;
    }
    int n = 0;
    for (int i = 0; i < len; i++) if (s[i] != gap) s[n++] = s[i];
    memmove(s + 1, s, n);
    s[0] = (uint8_t) gap;
    return n + 1;
}

/* which 0: algn_ancestor_2 (:4126-4147); 1: algn_get_median_2d_with_gaps (:4024-4035);
 * 2: algn_get_median_2d_no_gaps (:4042-4056).  out capacity len+1.  Returns the output length. */
int po_median_2(const po_cm *c, int which, const uint8_t *a, const uint8_t *b, int len, uint8_t *out) {
    int gap = c->gap, n = 0;
    if (which == 1) {
        for (int i = 0; i < len; i++) out[i] = (uint8_t) po_med(c, a[i], b[i]);
        return len;
    }
    if (which == 2) {
        out[n++] = (uint8_t) gap;
        for (int i = 0; i < len; i++) { int m = po_med(c, a[i], b[i]); if (m != gap) out[n++] = (uint8_t) m; }
        return n;
    }
    int drop = (!c->combinations) || (c->cost_model_type != 1);
    uint8_t *tmp = (uint8_t *) malloc((size_t) len + 2);
    for (int i = 0; i < len; i++) {
        int m = po_med(c, a[i], b[i]);
        if (!drop || m != gap) tmp[n++] = (uint8_t) m;
    }
    if (!c->combinations || (c->cost_model_type != 1 && (n == 0 || tmp[0] != gap))) {
        out[0] = (uint8_t) gap;
        memcpy(out + 1, tmp, n);
        n += 1;
    } else if (c->combinations) {
        /* combinations && affine: nothing was dropped, n == len */
        n = po_correct_blocks_affine(gap, tmp, n, a, b);
        memcpy(out, tmp, n);
    } else {
        memcpy(out, tmp, n);
    }
    free(tmp);
    return n;
}

/* ------------------------------------------------------------------------------------------------------
 * affine_3 (src/algn.c:1693-2680).  si = rows (the shorter operand), sj = columns; li/lj are the stored
 * lengths (leading gap included), leni = li-1, lenj = lj-1 as in algn_CAML_align_affine_3 (:2600).
 */
static inline int has_gap_opening(int prev, int cur, int gap, int go) { /* :1730-1733 */
    return (!(gap & prev) && (gap & cur)) ? 0 : go;
}

/* Cells visited per pair (band + left-edge cell per row + the init row), SURVEY.md 8d. */
long long po_cells_affine(int li, int lj) {
    int leni = li - 1, lenj = lj - 1;
    if (leni > lenj) { int t = leni; leni = lenj; lenj = t; }
    int s = 1, e = lenj - leni + 8;
    if (e < 40) e = 40;
    if (e > lenj) e = lenj;
    long long n = lenj + 1;
    for (int i = 1; i <= leni; i++) {
        if (i > 40) s++;
        n += e - s + 2;
        if (e < lenj) e++;
    }
    return n;
}

/* Shared fill.  bt != 0 restates initialize_matrices_affine + algn_fill_plane_3_aff (:2183-2248, :2411-2548)
 * and writes dir[(leni+1)*(lenj+1)]; bt == 0 restates the _nobt twins (:2124-2179, :2287-2403), whose
 * block-diagonal opening cost differs (:1848 vs :1872). */
static int po_affine_fill(const po_cm *c, const uint8_t *si, int li, const uint8_t *sj, int lj, int bt,
                          uint16_t *dir) {
    int leni = li - 1, lenj = lj - 1, gap = c->gap, go = c->gap_open, W = lenj + 1;
    size_t rowsz = (size_t) W;
    int *buf = (int *) malloc(sizeof(int) * rowsz * 11);
    int *EH[2] = {buf, buf + rowsz}, *EV[2] = {buf + 2 * rowsz, buf + 3 * rowsz};
    int *EB[2] = {buf + 4 * rowsz, buf + 5 * rowsz}, *CB[2] = {buf + 6 * rowsz, buf + 7 * rowsz};
    int *g = buf + 8 * rowsz, *gop = buf + 9 * rowsz, *hx = buf + 10 * rowsz;
    int F_last = 0; /* final_cost_matrix[lenj] */
    for (int j = 0; j <= lenj; j++) g[j] = c->prepend[sj[j]]; /* prec row 0, cm.c:703-704 */
    /* row 0 (:2194-2219) */
    CB[0][0] = 0; EB[0][0] = 0; EH[0][0] = go; EV[0][0] = go;
    if (dir) dir[0] = 0xFFFF;
    for (int j = 1; j <= lenj; j++) {
        int r = EH[0][j - 1] + g[j];
        EH[0][j] = r; CB[0][j] = r; EB[0][j] = PO_HIGH; EV[0][j] = PO_HIGH;
        if (dir) dir[j] = DO_H | END_H;
        F_last = r;
    }
    if (lenj == 0) F_last = 0;
    /* per-column gap opening / horizontal extension (:2452-2458) */
    for (int j = 1; j <= lenj; j++) {
        gop[j] = has_gap_opening(sj[j - 1], sj[j], gap, go);
        hx[j] = ((sj[j - 1] & gap) && !(sj[j] & gap)) ? gop[j] + g[j] : g[j];
    }
    if (lenj >= 1) hx[1] = g[1];
    int s = 1, e = lenj - leni + 8;
    if (e < 40) e = 40;
    if (e > lenj) e = lenj;
    for (int i = 1; i <= leni; i++) {
        int p = (i - 1) & 1, q = i & 1;
        uint16_t *drow = dir ? dir + (size_t) i * W : NULL;
        int ic = si[i], ip = si[i - 1];
        if (i > 40) s++;
        int ge = po_cost(c, ic, gap);             /* HAS_GAP_EXTENSION :1721 */
        int gop_i = has_gap_opening(ip, ic, gap, go);
        int vx = (i > 1 && (ip & gap) && !(ic & gap)) ? gop_i + ge : ge;
        int in_ = ic & 15;
        const int32_t *row = c->cost + (in_ << c->lcm);
        /* left edge cell (:2476-2494) */
        int r = EV[p][s - 1] + vx;
        EH[q][s - 1] = PO_HIGH; CB[q][s - 1] = PO_HIGH; EB[q][s - 1] = PO_HIGH; EV[q][s - 1] = r;
        if (drow) drow[s - 1] = DO_V | END_V;
        if (s - 1 == lenj) F_last = r;
        for (int j = s; j <= e; j++) {
            int jc = sj[j], jn = jc & 15, m = 0;
            /* extend horizontal (:1765-1787) */
            int ext = EH[q][j - 1] + hx[j], opn = CB[q][j - 1] + gop[j] + g[j];
            int eh;
            if (ext < opn) { m |= BEG_H; eh = ext; } else { m |= END_H; eh = opn; }
            /* extend vertical (:1813-1830) */
            ext = EV[p][j] + vx; opn = CB[p][j] + gop_i + ge;
            int ev;
            if (ext < opn) { m |= BEG_V; ev = ext; } else { m |= END_V; ev = opn; }
            /* extend block diagonal (:1861-1882 / :1837-1854) */
            int both = (TGAP & ic) && (TGAP & jc);
            int dg = both ? 0 : PO_HIGH, odg;
            if (bt) odg = dg;
            else odg = both ? ((!(TGAP & ip) && !(TGAP & jc)) ? 0 : 2 * go) : PO_HIGH;
            ext = EB[p][j - 1] + dg; opn = CB[p][j - 1] + odg;
            int eb;
            if (ext < opn) { m |= BEG_B; eb = ext; } else { m |= END_B; eb = opn; }
            /* close block diagonal (:1923-1977) */
            int d = row[jn];
            int xo = gop[j] < gop_i ? gop_i : gop[j];
            int al = CB[p][j - 1] + d;
            int fv = EV[p][j - 1] + d + ((ic == in_) ? 0 : gop[j]);
            int fh = EH[p][j - 1] + d + ((jc == jn) ? 0 : gop_i);
            int fd = EB[p][j - 1] + d + xo;
            int cb = al, mk = A2A;
            if (cb >= fv) { if (cb > fv) { cb = fv; mk = A2V; } else mk |= A2V; }
            if (cb >= fh) { if (cb > fh) { cb = fh; mk = A2H; } else mk |= A2H; }
            if (cb >= fd) { if (cb > fd) { cb = fd; mk = A2D; } else mk |= A2D; }
            m |= mk;
            /* ASSIGN_MINIMUM (:2251-2280) */
            int f = eh; mk = DO_H;
            if (f >= ev) { if (f > ev) { f = ev; mk = DO_V; } else mk |= DO_V; }
            if (f >= eb) { if (f > eb) { f = eb; mk = DO_D; } else mk |= DO_D; }
            if (f >= cb) { if (f > cb) { f = cb; mk = DO_A; } else mk |= DO_A; }
            m |= mk;
            EH[q][j] = eh; EV[q][j] = ev; EB[q][j] = eb; CB[q][j] = cb;
            if (drow) drow[j] = (uint16_t) m;
            if (j == lenj) F_last = f;
        }
        if (e < lenj) { /* widen and poison (:2528-2536) */
            e++;
            EH[q][e] = PO_HIGH; EV[q][e] = PO_HIGH; EB[q][e] = PO_HIGH; CB[q][e] = PO_HIGH;
            if (drow) drow[e] = DO_H | END_H;
        }
    }
    int res;
    if (bt) res = F_last; /* :2546 */
    else {                /* :2398-2402 */
        int q = leni & 1;
        res = EH[q][lenj];
        if (res > EV[q][lenj]) res = EV[q][lenj];
        if (res > EB[q][lenj]) res = EB[q][lenj];
        if (res > CB[q][lenj]) res = CB[q][lenj];
    }
    free(buf);
    return res;
}

/* backtrace_affine (src/algn.c:1983-2097).  Outputs: capacity li+lj+2 each, left-aligned on return;
 * lens[0..3] = median, medianwg, resi, resj. */
static void po_backtrace_affine(const po_cm *c, const uint16_t *dir, const uint8_t *si, int li, const uint8_t *sj,
                                int lj, uint8_t *med, uint8_t *medwg, uint8_t *resi, uint8_t *resj, int *lens) {
    int leni = li - 1, lenj = lj - 1, W = lenj + 1, cap = li + lj + 2;
    int i = leni, j = lenj, nm = 0, nw = 0, nr = 0;
    enum { TODO, VERT, HORI, DIAG, ALGN } mode = TODO;
#define EMIT_MED(v) (med[cap - (++nm)] = (uint8_t) (v))
#define EMIT_WG(v) (medwg[cap - (++nw)] = (uint8_t) (v))
#define EMIT_RES(a, b) (nr++, resi[cap - nr] = (uint8_t) (a), resj[cap - nr] = (uint8_t) (b))
    while (i != 0 && j != 0) {
        int m = dir[(size_t) i * W + j], ic = si[i], jc = sj[j];
        if (mode == TODO) {
            if (m & DO_H) mode = HORI;
            else if (m & DO_A) mode = ALGN;
            else if (m & DO_V) mode = VERT;
            else mode = DIAG;
        } else if (mode == VERT) {
            if (m & END_V) mode = TODO;
            if (!(ic & TGAP)) { EMIT_MED(ic | TGAP); EMIT_WG(ic | TGAP); } else EMIT_WG(TGAP);
            EMIT_RES(ic, TGAP);
            i--;
        } else if (mode == HORI) {
            if (m & END_H) mode = TODO;
            if (!(jc & TGAP)) { EMIT_MED(jc | TGAP); EMIT_WG(jc | TGAP); } else EMIT_WG(TGAP);
            EMIT_RES(TGAP, jc);
            j--;
        } else if (mode == DIAG) {
            if (m & END_B) mode = TODO;
            EMIT_RES(ic, jc);
            EMIT_WG(TGAP);
            i--; j--;
        } else {
            if (m & A2H) mode = HORI;
            else if (m & A2D) mode = DIAG;
            else if (m & A2V) mode = VERT;
            int p = po_med(c, ic & 15, jc & 15);
            EMIT_MED(p); EMIT_WG(p);
            EMIT_RES(ic, jc);
            i--; j--;
        }
    }
    while (i != 0) {
        int ic = si[i];
        if (!(ic & TGAP)) { EMIT_MED(ic | TGAP); EMIT_WG(ic | TGAP); } else EMIT_WG(TGAP);
        EMIT_RES(ic, TGAP);
        i--;
    }
    while (j != 0) {
        int jc = sj[j];
        if (!(jc & TGAP)) { EMIT_MED(jc | TGAP); EMIT_WG(jc | TGAP); } else EMIT_WG(TGAP);
        EMIT_RES(TGAP, jc);
        j--;
    }
    EMIT_RES(TGAP, TGAP);
    EMIT_WG(TGAP);
    if (nm == 0 || med[cap - nm] != TGAP) EMIT_MED(TGAP); /* :2093; an empty median reads as "not a gap" */
#undef EMIT_MED
#undef EMIT_WG
#undef EMIT_RES
    memmove(med, med + cap - nm, nm);
    memmove(medwg, medwg + cap - nw, nw);
    memmove(resi, resi + cap - nr, nr);
    memmove(resj, resj + cap - nr, nr);
    lens[0] = nm; lens[1] = nw; lens[2] = nr; lens[3] = nr;
}

/* algn_CAML_align_affine_3 (:2551-2619): the shorter operand (ties: the first) takes the rows; res*
 * follow the ORIGINAL operand order.  dir_out (optional) gets the (min+... ) direction matrix. */
int po_align_affine_3(const po_cm *c, const uint8_t *sa, int la, const uint8_t *sb, int lb, uint8_t *median,
                      uint8_t *medianwg, uint8_t *resa, uint8_t *resb, int *lens, uint16_t *dir_out) {
    const uint8_t *si = sa, *sj = sb;
    int li = la, lj = lb;
    uint8_t *ri = resa, *rj = resb;
    if (la > lb) { si = sb; li = lb; sj = sa; lj = la; ri = resb; rj = resa; }
    uint16_t *dir = dir_out ? dir_out : (uint16_t *) malloc(sizeof(uint16_t) * (size_t) li * lj);
    int res = po_affine_fill(c, si, li, sj, lj, 1, dir);
    po_backtrace_affine(c, dir, si, li, sj, lj, median, medianwg, ri, rj, lens);
    if (!dir_out) free(dir);
    return res;
}

/* algn_CAML_cost_affine_3 (:2628-2680) */
int po_cost_affine_3(const po_cm *c, const uint8_t *sa, int la, const uint8_t *sb, int lb) {
    if (la <= lb) return po_affine_fill(c, sa, la, sb, lb, 0, NULL);
    return po_affine_fill(c, sb, lb, sa, la, 0, NULL);
}

/* Sequence.Align.align_2 for a linear matrix (src/sequence.ml:813-823, 849-861): longer operand first,
 * swaped = (la >= lb); ra/rb follow the ORIGINAL operand order.  Capacity la+lb each. */
int po_align_2(const po_cm *c, const uint8_t *sa, int la, const uint8_t *sb, int lb, int deltaw, uint8_t *ra,
               uint8_t *rb, int *rlen) {
    int swaped = la >= lb;
    const uint8_t *s1 = swaped ? sa : sb, *s2 = swaped ? sb : sa;
    int l1 = swaped ? la : lb, l2 = swaped ? lb : la;
    uint16_t *dir = (uint16_t *) calloc((size_t) l1 * l2, sizeof(uint16_t));
    int res = po_cost_2(c, s1, l1, s2, l2, deltaw, dir);
    *rlen = po_backtrack_2d(c, s1, l1, s2, l2, dir, swaped, swaped ? ra : rb, swaped ? rb : ra);
    free(dir);
    return res;
}

/* ---- batch driver (same contract as ref_batch in ref_driver.c) ------------------------------------- */
typedef struct po_arg {
    const po_cm *c; int mode; const uint8_t *pool; const long long *off; const int *len; const int *pairs;
    const int *deltaw; int lo, hi; int *cost; uint8_t *median, *medianwg, *ra, *rb; int *lens; long long stride;
} po_arg;

static void *po_worker(void *p) {
    po_arg *a = (po_arg *) p;
    int maxcap = 16;
    for (int k = a->lo; k < a->hi; k++) {
        int cc = a->len[a->pairs[2 * k]] + a->len[a->pairs[2 * k + 1]] + 2;
        if (cc > maxcap) maxcap = cc;
    }
    uint8_t *t0 = malloc(maxcap), *t1 = malloc(maxcap), *t2 = malloc(maxcap), *t3 = malloc(maxcap);
    for (int k = a->lo; k < a->hi; k++) {
        int ia = a->pairs[2 * k], ib = a->pairs[2 * k + 1];
        const uint8_t *sa = a->pool + a->off[ia], *sb = a->pool + a->off[ib];
        int la = a->len[ia], lb = a->len[ib], cost = 0, lens[4] = {0, 0, 0, 0};
        if (a->mode == 0) {
            cost = (la >= lb) ? po_cost_2(a->c, sa, la, sb, lb, a->deltaw[k], NULL)
                              : po_cost_2(a->c, sb, lb, sa, la, a->deltaw[k], NULL);
        } else if (a->mode == 1) {
            int rl;
            cost = po_align_2(a->c, sa, la, sb, lb, a->deltaw[k], t2, t3, &rl);
            lens[2] = lens[3] = rl;
            lens[0] = po_median_2(a->c, 0, t2, t3, rl, t0);
            lens[1] = po_median_2(a->c, 1, t2, t3, rl, t1);
        } else if (a->mode == 2) {
            cost = po_cost_affine_3(a->c, sa, la, sb, lb);
        } else {
            cost = po_align_affine_3(a->c, sa, la, sb, lb, t0, t1, t2, t3, lens, NULL);
        }
        if (a->cost) a->cost[k] = cost;
        if (a->lens) memcpy(a->lens + 4 * (size_t) k, lens, sizeof(lens));
        if (a->mode == 1 || a->mode == 3) {
            size_t o = (size_t) k * a->stride;
            if (a->median) memcpy(a->median + o, t0, lens[0]);
            if (a->medianwg) memcpy(a->medianwg + o, t1, lens[1]);
            if (a->ra) memcpy(a->ra + o, t2, lens[2]);
            if (a->rb) memcpy(a->rb + o, t3, lens[3]);
        }
    }
    free(t0); free(t1); free(t2); free(t3);
    return NULL;
}

int po_batch(const po_cm *c, int mode, const uint8_t *pool, const long long *off, const int *len, const int *pairs,
             const int *deltaw, int n, int nthreads, int *cost, uint8_t *median, uint8_t *medianwg, uint8_t *ra,
             uint8_t *rb, int *lens, long long stride) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    po_arg *args = malloc(sizeof(po_arg) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        po_arg b = {c, mode, pool, off, len, pairs, deltaw, (int) ((long long) n * t / nthreads),
                    (int) ((long long) n * (t + 1) / nthreads), cost, median, medianwg, ra, rb, lens, stride};
        args[t] = b;
        if (nthreads == 1) po_worker(&args[t]);
        else pthread_create(&th[t], NULL, po_worker, &args[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(args);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------
 * 3-D: algn_fill_cube / backtrack_3d / algn_get_median_3d AS THE REFERENCE EXECUTES THEM.
 *
 * The cube is l1 planes x l2 rows x l3 cells, filled row by row.  Row (i, j) has linear index r = i*l2 + j.  The
 * reference sets its three neighbour-row pointers once before the plane loop and advances them l2-1 rows per plane
 * (src/algn.c:2978-2982, 3023-3025) instead of l2, so for i >= 1 row (i, j >= 1) reads
 *     upper = 1 + (i-1)(l2-1) + (j-1),   diag = upper - 1,   prev = i(l2-1) + (j-1)
 * and row (i, 0) reads diag = (i-1)(l2-1) -- not the rows (i-1, j), (i-1, j-1), (i, j-1) a correct recurrence
 * needs.  All of these are earlier rows, so the result is deterministic; it is simply not an optimal alignment
 * cost.  Direction codes (src/matrices.h:34-40): P1 1, P2 2, P3 4, S1 8, S2 16, S3 32, SS 64; a later candidate
 * replaces an earlier one only if strictly cheaper, order P3, P1, P2, S3, S1, S2, SS.
 */
typedef struct po_cm3 {
    int32_t lcm, gap;
    const int32_t *cost3;   /* (1<<lcm)^3, index ((a << lcm) + b) << lcm) + c   src/cm.c:509-523 */
    const uint8_t *median3;
} po_cm3;

static inline int po_cost3(const po_cm3 *c, int a, int b, int d) { return c->cost3[(((a << c->lcm) + b) << c->lcm) + d]; }

/* dir: l1*l2*l3 bytes (one code per cell) or NULL.  Returns mm[-1] of src/algn.c:3054. */
int po_cost_3(const po_cm3 *c, const uint8_t *s1, int l1, const uint8_t *s2, int l2, const uint8_t *s3, int l3,
              uint8_t *dir) {
    const int gap = c->gap;
    size_t nrows = (size_t) l1 * l2;
    int *M = (int *) malloc(sizeof(int) * nrows * (size_t) l3);
    int *gg = (int *) malloc(sizeof(int) * (size_t) l3);
    for (int k = 0; k < l3; k++) gg[k] = po_cost3(c, gap, gap, s3[k]);
#define ROW(r) (M + (size_t) (r) * l3)
#define DIR(r, k, v) do { if (dir) dir[(size_t) (r) * l3 + (k)] = (uint8_t) (v); } while (0)
    /* plane 0 (:2918-2962) */
    ROW(0)[0] = 0; DIR(0, 0, 16);
    for (int k = 1; k < l3; k++) { ROW(0)[k] = ROW(0)[k - 1] + gg[k]; DIR(0, k, 64); }
    for (int j = 1; j < l2; j++) {
        int *mm = ROW(j), *prev = ROW(j - 1), b = s2[j];
        int g0 = po_cost3(c, gap, b, s3[0]);
        mm[0] = prev[0] + g0; DIR(j, 0, 1);
        for (int k = 1; k < l3; k++) {
            int v = prev[k] + g0, d = 1, t = prev[k - 1] + po_cost3(c, gap, b, s3[k]);
            if (t < v) { v = t; d = 8; }
            t = mm[k - 1] + gg[k];
            if (t < v) { v = t; d = 64; }
            mm[k] = v; DIR(j, k, d);
        }
    }
    for (int i = 1; i < l1; i++) {
        int a = s1[i];
        int a0 = po_cost3(c, a, gap, s3[0]);
        /* first row of the plane (:2990-3013): the "diag" pointer stands for the row above */
        {
            size_t r = (size_t) i * l2;
            int *mm = ROW(r), *D = ROW((size_t) (i - 1) * (l2 - 1));
            mm[0] = D[0] + a0; DIR(r, 0, 4);
            for (int k = 1; k < l3; k++) {
                int v = D[k] + a0, d = 4, t = D[k - 1] + po_cost3(c, a, gap, s3[k]);
                if (t < v) { v = t; d = 32; }
                t = gg[k] + mm[k - 1];
                if (t < v) { v = t; d = 64; }
                mm[k] = v; DIR(r, k, d);
            }
        }
        for (int j = 1; j < l2; j++) {
            size_t r = (size_t) i * l2 + j;
            int b = s2[j];
            int *mm = ROW(r);
            const int *U = ROW((size_t) 1 + (size_t) (i - 1) * (l2 - 1) + (j - 1));
            const int *D = U - l3;
            const int *P = ROW((size_t) i * (l2 - 1) + (j - 1));
            int s1gg = a0, gs2g = po_cost3(c, gap, b, s3[0]), s1s2g = po_cost3(c, a, b, s3[0]);
            for (int k = 0; k < l3; k++) {
                /* fill_parallel (:2840-2857) */
                int v = U[k] + s1gg, d = 4, t = P[k] + gs2g;
                if (t < v) { v = t; d = 1; }
                t = D[k] + s1s2g;
                if (t < v) { v = t; d = 2; }
                if (k >= 1) {
                    /* fill_moved (:2812-2836) */
                    t = U[k - 1] + po_cost3(c, a, gap, s3[k]);
                    if (t < v) { v = t; d = 32; }
                    t = P[k - 1] + po_cost3(c, gap, b, s3[k]);
                    if (t < v) { v = t; d = 8; }
                    t = D[k - 1] + po_cost3(c, a, b, s3[k]);
                    if (t < v) { v = t; d = 16; }
                    /* the in-row pass (:3039-3046) */
                    t = mm[k - 1] + gg[k];
                    if (t < v) { v = t; d = 64; }
                }
                mm[k] = v; DIR(r, k, d);
            }
        }
    }
    int res = ROW(nrows - 1)[l3 - 1];
#undef ROW
#undef DIR
    free(M); free(gg);
    return res;
}

/* The recurrence algn_fill_cube INTENDS (SURVEY.md A12 and "3-D policy" (1)): the same seven candidates, in the same order
 * and with the same strict `<` (P3 4, P1 1, P2 2, S3 32, S1 8, S2 16, SS 64; src/algn.c:2812-2857, 3039-3046), but read
 * from the rows a three-sequence Needleman-Wunsch needs -- (i-1, j), (i-1, j-1), (i, j-1) -- instead of the lagging rows
 * the compiled code reads.  Plane 0 and the cost expressions are the reference's own (they are correct as written).
 * Checked against exhaustive enumeration and an independent memoised recursion in tests/test_cube_intended.py; the CUDA
 * cube kernel reproduces the EXECUTED form (po_cost_3), this function documents what a corrected reference would return. */
int po_cost_3_intended(const po_cm3 *c, const uint8_t *s1, int l1, const uint8_t *s2, int l2, const uint8_t *s3, int l3,
                       uint8_t *dir) {
    const int gap = c->gap;
    size_t nrows = (size_t) l1 * l2;
    int *M = (int *) malloc(sizeof(int) * nrows * (size_t) l3);
    int *gg = (int *) malloc(sizeof(int) * (size_t) l3);
    for (int k = 0; k < l3; k++) gg[k] = po_cost3(c, gap, gap, s3[k]);
#define ROW(r) (M + (size_t) (r) * l3)
#define DIR(r, k, v) do { if (dir) dir[(size_t) (r) * l3 + (k)] = (uint8_t) (v); } while (0)
    ROW(0)[0] = 0; DIR(0, 0, 16);
    for (int k = 1; k < l3; k++) { ROW(0)[k] = ROW(0)[k - 1] + gg[k]; DIR(0, k, 64); }
    for (int j = 1; j < l2; j++) {  /* plane 0 (:2918-2962) */
        int *mm = ROW(j), *prev = ROW(j - 1), b = s2[j];
        int g0 = po_cost3(c, gap, b, s3[0]);
        mm[0] = prev[0] + g0; DIR(j, 0, 1);
        for (int k = 1; k < l3; k++) {
            int v = prev[k] + g0, d = 1, t = prev[k - 1] + po_cost3(c, gap, b, s3[k]);
            if (t < v) { v = t; d = 8; }
            t = mm[k - 1] + gg[k];
            if (t < v) { v = t; d = 64; }
            mm[k] = v; DIR(j, k, d);
        }
    }
    for (int i = 1; i < l1; i++) {
        int a = s1[i];
        int a0 = po_cost3(c, a, gap, s3[0]);
        {   /* row (i, 0): only s1 and s3 can move (:2990-3013), from row (i-1, 0) */
            size_t r = (size_t) i * l2;
            int *mm = ROW(r), *U = ROW((size_t) (i - 1) * l2);
            mm[0] = U[0] + a0; DIR(r, 0, 4);
            for (int k = 1; k < l3; k++) {
                int v = U[k] + a0, d = 4, t = U[k - 1] + po_cost3(c, a, gap, s3[k]);
                if (t < v) { v = t; d = 32; }
                t = gg[k] + mm[k - 1];
                if (t < v) { v = t; d = 64; }
                mm[k] = v; DIR(r, k, d);
            }
        }
        for (int j = 1; j < l2; j++) {
            size_t r = (size_t) i * l2 + j;
            int b = s2[j];
            int *mm = ROW(r);
            const int *U = ROW((size_t) (i - 1) * l2 + j);        /* (i-1, j)   */
            const int *D = ROW((size_t) (i - 1) * l2 + (j - 1));  /* (i-1, j-1) */
            const int *P = ROW((size_t) i * l2 + (j - 1));        /* (i, j-1)   */
            int s1gg = a0, gs2g = po_cost3(c, gap, b, s3[0]), s1s2g = po_cost3(c, a, b, s3[0]);
            for (int k = 0; k < l3; k++) {
                int v = U[k] + s1gg, d = 4, t = P[k] + gs2g;
                if (t < v) { v = t; d = 1; }
                t = D[k] + s1s2g;
                if (t < v) { v = t; d = 2; }
                if (k >= 1) {
                    t = U[k - 1] + po_cost3(c, a, gap, s3[k]);
                    if (t < v) { v = t; d = 32; }
                    t = P[k - 1] + po_cost3(c, gap, b, s3[k]);
                    if (t < v) { v = t; d = 8; }
                    t = D[k - 1] + po_cost3(c, a, b, s3[k]);
                    if (t < v) { v = t; d = 16; }
                    t = mm[k - 1] + gg[k];
                    if (t < v) { v = t; d = 64; }
                }
                mm[k] = v; DIR(r, k, d);
            }
        }
    }
    int res = ROW(nrows - 1)[l3 - 1];
#undef ROW
#undef DIR
    free(M); free(gg);
    return res;
}

/* backtrack_3d's walk (:3829-3905, same test order) over a direction cube of po_cost_3_intended, and the median
 * algn_get_median_3d INTENDS (:4160-4173 with the pointers moving): one cm_get_median_3d per column.  Left aligned on return. */
int po_backtrack_3_intended(const po_cm3 *c, const uint8_t *dir, const uint8_t *s1, int l1, const uint8_t *s2, int l2,
                            const uint8_t *s3, int l3, uint8_t *r1, uint8_t *r2, uint8_t *r3, uint8_t *med, int *status) {
    int cap = l1 + l2 + l3, n = 0, i1 = l1 - 1, i2 = l2 - 1, i3 = l3 - 1;
    long long p = (long long) l1 * l2 * l3 - 1, plane = (long long) l2 * l3, line = l3;
    uint8_t gap = (uint8_t) c->gap;
    *status = 0;
    while (p > 0) {
        int v = dir[p], u1 = 0, u2 = 0, u3 = 0;
        if (v & 16) { u1 = u2 = u3 = 1; p -= plane + line + 1; }
        else if (v & 32) { u1 = u3 = 1; p -= plane + 1; }
        else if (v & 8) { u2 = u3 = 1; p -= line + 1; }
        else if (v & 4) { u1 = 1; p -= plane; }
        else if (v & 64) { u3 = 1; p -= 1; }
        else if (v & 1) { u2 = 1; p -= line; }
        else { u1 = u2 = 1; p -= plane + line; }
        if ((u1 && i1 < 1) || (u2 && i2 < 1) || (u3 && i3 < 1) || n >= cap) { *status = 1; return 0; }
        n++;
        r1[cap - n] = u1 ? s1[i1--] : gap;
        r2[cap - n] = u2 ? s2[i2--] : gap;
        r3[cap - n] = u3 ? s3[i3--] : gap;
    }
    if (i1 != 0 || i2 != 0 || i3 != 0) { *status = 1; return 0; }  /* every element but the leading gaps consumed */
    memmove(r1, r1 + cap - n, n);
    memmove(r2, r2 + cap - n, n);
    memmove(r3, r3 + cap - n, n);
    for (int k = 0; k < n; k++) med[k] = c->median3[(((r1[k] << c->lcm) + r2[k]) << c->lcm) + r3[k]];
    return n;
}

/* backtrack_3d (:3829-3905) + algn_get_median_3d (:4160-4173).  Outputs need l1+l2+l3 (+1 for med) bytes, left aligned
 * on return.  *status = 1 (and nothing is produced) when the reference's walk would index a sequence below 0, which
 * the reference does not check. */
int po_backtrack_3(const po_cm3 *c, const uint8_t *dir, const uint8_t *s1, int l1, const uint8_t *s2, int l2,
                   const uint8_t *s3, int l3, uint8_t *r1, uint8_t *r2, uint8_t *r3, uint8_t *med, int *medlen,
                   int *status) {
    int cap = l1 + l2 + l3, n = 0, i1 = l1 - 1, i2 = l2 - 1, i3 = l3 - 1;
    long long p = (long long) l1 * l2 * l3 - 1, plane = (long long) l2 * l3, line = l3;
    uint8_t gap = (uint8_t) c->gap;
    *status = 0;
    while (p > 0) {
        int v = dir[p], u1 = 0, u2 = 0, u3 = 0;
        if (v & 16) { u1 = u2 = u3 = 1; p -= plane + line + 1; }
        else if (v & 32) { u1 = u3 = 1; p -= plane + 1; }
        else if (v & 8) { u2 = u3 = 1; p -= line + 1; }
        else if (v & 4) { u1 = 1; p -= plane; }
        else if (v & 64) { u3 = 1; p -= 1; }
        else if (v & 1) { u2 = 1; p -= line; }
        else { u1 = u2 = 1; p -= plane + line; }
        if ((u1 && i1 < 0) || (u2 && i2 < 0) || (u3 && i3 < 0) || n >= cap) { *status = 1; *medlen = 0; return 0; }
        n++;
        r1[cap - n] = u1 ? s1[i1--] : gap;
        r2[cap - n] = u2 ? s2[i2--] : gap;
        r3[cap - n] = u3 ? s3[i3--] : gap;
    }
    /* the median loop never moves its pointers: n copies of the median of the LAST column */
    if (n > 0) {
        int m = c->median3[(((r1[cap - 1] << c->lcm) + r2[cap - 1]) << c->lcm) + r3[cap - 1]];
        memset(med, m, n);
    }
    *medlen = n;
    memmove(r1, r1 + cap - n, n);
    memmove(r2, r2 + cap - n, n);
    memmove(r3, r3 + cap - n, n);
    return n;
}
