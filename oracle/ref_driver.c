/* TEST INFRASTRUCTURE ONLY -- never linked into, or called from, the product.
 *
 * Plain-C driver around the UNMODIFIED reference translation unit
 * /root/reference/src/algn.c (compiled where it lies, see oracle/Makefile).
 * Nothing of the reference is copied here: this file only builds the
 * reference's own structs (`struct seq`, `struct cm`, `struct matrices`) from
 * flat buffers and calls the reference's own non-stub functions in the order
 * its OCaml stubs do:
 *
 *   algn_CAML_simple_2        src/algn.c:3409  -> mat_setup_size + algn_nw
 *   algn_CAML_backtrack_2d    src/algn.c:3908  -> backtrack_2d
 *   algn_CAML_ancestor_2      src/algn.c:4288  -> algn_ancestor_2
 *   algn_CAML_median_2_*      src/algn.c:4198,4211
 *   algn_CAML_align_affine_3  src/algn.c:2551  (body replicated: carving of the
 *                                               scratch block, precalc on the
 *                                               longer operand, init, fill,
 *                                               backtrace)
 *   algn_CAML_cost_affine_3   src/algn.c:2628
 *   algn_CAML_simple_3/backtrack_3d/median_3   src/algn.c:3458-4235
 *
 * Built into oracle/_ref/libpoyref.so.  It is the parity oracle of the tests
 * and the "reference" CPU baseline of bench.py.
 */
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include <caml/mlvalues.h>
#include "matrices.h"
#include "seq.h"
#include "cm.h"

/* The OCaml runtime symbols the reference TU leaves undefined live in caml_runtime.c (same library). */

/* ---- reference functions we call (defined in algn.o) --------------------- */
extern cmt cm_set_val(int a_sz, int combinations, int do_aff, int gap_open, int is_metric, int all_elements, cmt res);
extern int mat_setup_size(matricest m, int w, int d, int h, int k, int lcm);
extern int algn_nw(const seqt s1, const seqt s2, const cmt c, matricest m, int deltawh);
extern void backtrack_2d(const seqt s1, const seqt s2, seqt r1, seqt r2, const matricest m, const cmt c,
                         int st_s1, int st_s2, int algn_s1, int algn_s2, int swaped, value a, value b);
extern void algn_ancestor_2(seqt s1, seqt s2, cmt m, seqt sm);
extern void algn_get_median_2d_with_gaps(seqt s1, seqt s2, cmt m, seqt sm);
extern void algn_get_median_2d_no_gaps(seqt s1, seqt s2, cmt m, seqt sm);
extern void cm_precalc_4algn(const cmt c, matricest matrix, const seqt s);
extern void initialize_matrices_affine(int go, const seqt si, const seqt sj, const cmt c, int *cbd, int *ebd,
                                       int *ev, int *eh, int *fcm, DIRECTION_MATRIX *dm, const int *prec);
extern void initialize_matrices_affine_nobt(int go, const seqt si, const seqt sj, const cmt c, int *cbd, int *ebd,
                                            int *ev, int *eh, const int *prec);
extern int algn_fill_plane_3_aff(const seqt si, const seqt sj, int leni, int lenj, int *fcm, DIRECTION_MATRIX *dm,
                                 const cmt c, int *eh, int *ev, int *cbd, int *ebd, const int *prec, int *gop,
                                 int *sjhe);
extern int algn_fill_plane_3_aff_nobt(const seqt si, const seqt sj, int leni, int lenj, const cmt c, int *eh, int *ev,
                                      int *cbd, int *ebd, const int *prec, int *gop, int *sjhe);
extern void backtrace_affine(DIRECTION_MATRIX *dm, const seqt si, const seqt sj, seqt median, seqt medianwg,
                             seqt resi, seqt resj, const cmt c);
extern int algn_worst_2(seqt s1, seqt s2, cmt c);
extern int algn_verify_2(seqt s1, seqt s2, cmt c);

/* ---- flat-buffer helpers -------------------------------------------------- */
typedef struct ref_cm {
    struct cm c;
    int table_dim; /* 1 << lcm */
} ref_cm;

/* a_sz_in / combinations as given to cm_CAML_create (src/cm.c:1262 -> cm_set_val :319).  The tables are
 * flat (1<<lcm) x (1<<lcm) [cost, median, worst] and (1<<lcm) [prepend, tail], exactly the index
 * convention (a << lcm) + b of cm_calc_cost (src/cm.c:545). */
void *ref_cm_create(int a_sz_in, int combinations, int cost_model, int gap_open, int is_metric, int all_elements,
                    const int *cost, const unsigned char *median, const int *worst, const int *prepend,
                    const int *tail) {
    ref_cm *r = (ref_cm *) calloc(1, sizeof(ref_cm));
    cm_set_val(a_sz_in, combinations, cost_model, gap_open, is_metric, all_elements, &r->c);
    int dim = 1 << r->c.lcm;
    r->table_dim = dim;
    memcpy(r->c.cost, cost, sizeof(int) * dim * dim);
    memcpy(r->c.median, median, dim * dim);
    if (worst) memcpy(r->c.worst, worst, sizeof(int) * dim * dim);
    memcpy(r->c.prepend_cost, prepend, sizeof(int) * dim);
    memcpy(r->c.tail_cost, tail, sizeof(int) * dim);
    return r;
}
void ref_cm_free(void *h) {
    ref_cm *r = (ref_cm *) h;
    free(r->c.cost); free(r->c.median); free(r->c.worst); free(r->c.prepend_cost); free(r->c.tail_cost);
    free(r);
}
const struct cm *ref_cm_struct(void *h) { return &((ref_cm *) h)->c; }
int ref_cm_lcm(void *h) { return ((ref_cm *) h)->c.lcm; }
int ref_cm_gap(void *h) { return ((ref_cm *) h)->c.gap; }
int ref_cm_a_sz(void *h) { return ((ref_cm *) h)->c.a_sz; }

/* A `struct seq` followed by its storage, laid out like the OCaml custom block (src/seq.h:31-36).  One
 * zeroed guard byte follows the storage so that the reference's `seq_get (sm, 0)` on an EMPTY result
 * (src/algn.c:2093, :4141) reads a defined 0 instead of whatever follows the block on the OCaml heap. */
static seqt mk_seq(int cap) {
    seqt s = (seqt) calloc(1, sizeof(struct seq) + cap + 16);
    s->magic_number = POY_SEQ_MAGIC_NUMBER;
    s->cap = cap;
    s->len = 0;
    s->head = (SEQT *) (s + 1);
    s->end = s->head + cap - 1;
    s->begin = s->end + 1;
    return s;
}
static seqt mk_seq_from(const unsigned char *data, int len, int cap) {
    if (cap < len) cap = len;
    seqt s = mk_seq(cap);
    s->len = len;
    s->begin = s->end - len + 1;
    memcpy(s->begin, data, len);
    return s;
}
static int seq_out(seqt s, unsigned char *out) {
    memcpy(out, s->begin, s->len);
    return s->len;
}

typedef struct ref_ws {
    struct matrices m;
} ref_ws;
void *ref_ws_create(void) { return calloc(1, sizeof(ref_ws)); }
void ref_ws_free(void *h) {
    ref_ws *w = (ref_ws *) h;
    free(w->m.matrix); free(w->m.matrix_d); free(w->m.precalc); free(w->m.pointers_3d);
    free(w);
}

/* algn_CAML_simple_2 (src/algn.c:3409).  Requires l1 >= l2, as Sequence.Align.cost_2 guarantees
 * (src/sequence.ml:709-714). */
int ref_cost_2(void *cmh, void *wsh, const unsigned char *s1, int l1, const unsigned char *s2, int l2, int deltaw) {
    ref_cm *c = (ref_cm *) cmh; ref_ws *w = (ref_ws *) wsh;
    seqt a = mk_seq_from(s1, l1, l1), b = mk_seq_from(s2, l2, l2);
    mat_setup_size(&w->m, l1, l2, 0, 0, c->c.lcm);
    int res = algn_nw(a, b, &c->c, &w->m, deltaw);
    free(a); free(b);
    return res;
}

/* algn_CAML_align_2d (src/algn.c:3987) = simple_2 then backtrack_2d.  r1/r2 need l1+l2 bytes
 * (src/sequence.ml:816-817).  Returns the cost, *rlen = aligned length. */
int ref_align_2(void *cmh, void *wsh, const unsigned char *s1, int l1, const unsigned char *s2, int l2, int deltaw,
                int swaped, unsigned char *r1, unsigned char *r2, int *rlen) {
    ref_cm *c = (ref_cm *) cmh; ref_ws *w = (ref_ws *) wsh;
    seqt a = mk_seq_from(s1, l1, l1), b = mk_seq_from(s2, l2, l2);
    seqt ra = mk_seq(l1 + l2), rb = mk_seq(l1 + l2);
    mat_setup_size(&w->m, l1, l2, 0, 0, c->c.lcm);
    int res = algn_nw(a, b, &c->c, &w->m, deltaw);
    backtrack_2d(a, b, ra, rb, &w->m, &c->c, 0, 0, l1, l2, swaped, 0, 0);
    *rlen = seq_out(ra, r1);
    seq_out(rb, r2);
    free(a); free(b); free(ra); free(rb);
    return res;
}

/* Direction matrix of the last ref_cost_2/ref_align_2 on this workspace (row stride l2), for debugging. */
void ref_get_dir_2(void *wsh, int l1, int l2, unsigned short *out) {
    ref_ws *w = (ref_ws *) wsh;
    memcpy(out, w->m.matrix_d, sizeof(unsigned short) * (size_t) l1 * l2);
}
void ref_clear_dir(void *wsh) {
    ref_ws *w = (ref_ws *) wsh;
    if (w->m.matrix_d) memset(w->m.matrix_d, 0, sizeof(unsigned short) * (size_t) w->m.len);
}

/* which: 0 = algn_ancestor_2 (:4126), 1 = algn_get_median_2d_with_gaps (:4024), 2 = _no_gaps (:4042) */
int ref_median_2(void *cmh, int which, const unsigned char *a1, const unsigned char *a2, int len, unsigned char *out) {
    ref_cm *c = (ref_cm *) cmh;
    seqt a = mk_seq_from(a1, len, len), b = mk_seq_from(a2, len, len), sm = mk_seq(len + 1);
    if (which == 0) algn_ancestor_2(a, b, &c->c, sm);
    else if (which == 1) algn_get_median_2d_with_gaps(a, b, &c->c, sm);
    else algn_get_median_2d_no_gaps(a, b, &c->c, sm);
    int n = seq_out(sm, out);
    free(a); free(b); free(sm);
    return n;
}

int ref_worst_2(void *cmh, const unsigned char *a1, const unsigned char *a2, int len) {
    ref_cm *c = (ref_cm *) cmh;
    seqt a = mk_seq_from(a1, len, len), b = mk_seq_from(a2, len, len);
    int r = algn_worst_2(a, b, &c->c);
    free(a); free(b);
    return r;
}

/* Body of algn_CAML_align_affine_3 (src/algn.c:2579-2617) on flat buffers.  Outputs need li+lj+2 bytes each
 * (src/sequence.ml:470-474); lens[0..3] = median, medianwg, resi, resj lengths. */
int ref_align_affine_3(void *cmh, void *wsh, const unsigned char *si, int li, const unsigned char *sj, int lj,
                       unsigned char *median, unsigned char *medianwg, unsigned char *resi, unsigned char *resj,
                       int *lens, unsigned short *dir_out) {
    ref_cm *c = (ref_cm *) cmh; ref_ws *w = (ref_ws *) wsh;
    cmt ccm = &c->c; matricest cam = &w->m;
    int cap = li + lj + 2;
    seqt csi = mk_seq_from(si, li, li), csj = mk_seq_from(sj, lj, lj);
    seqt cmedian = mk_seq(cap), cmedianwg = mk_seq(cap), cresi = mk_seq(cap), cresj = mk_seq(cap);
    int leni = li, lenj = lj, largest = leni > lenj ? leni : lenj, res;
    mat_setup_size(cam, largest, largest, 0, 0, ccm->lcm);
    int *matrix = cam->matrix, *prec = cam->precalc;
    int *cbd = matrix, *ebd = matrix + 2 * largest, *ev = matrix + 4 * largest, *eh = matrix + 6 * largest;
    int *fcm = matrix + 8 * largest, *gop = matrix + 10 * largest, *she = matrix + 11 * largest;
    DIRECTION_MATRIX *dm = cam->matrix_d;
    if (leni <= lenj) {
        cm_precalc_4algn(ccm, cam, csj);
        initialize_matrices_affine(ccm->gap_open, csi, csj, ccm, cbd, ebd, ev, eh, fcm, dm, prec);
        res = algn_fill_plane_3_aff(csi, csj, leni - 1, lenj - 1, fcm, dm, ccm, eh, ev, cbd, ebd, prec, gop, she);
        backtrace_affine(dm, csi, csj, cmedian, cmedianwg, cresi, cresj, ccm);
    } else {
        cm_precalc_4algn(ccm, cam, csi);
        initialize_matrices_affine(ccm->gap_open, csj, csi, ccm, cbd, ebd, ev, eh, fcm, dm, prec);
        res = algn_fill_plane_3_aff(csj, csi, lenj - 1, leni - 1, fcm, dm, ccm, eh, ev, cbd, ebd, prec, gop, she);
        backtrace_affine(dm, csj, csi, cmedian, cmedianwg, cresj, cresi, ccm);
    }
    lens[0] = seq_out(cmedian, median);
    lens[1] = seq_out(cmedianwg, medianwg);
    lens[2] = seq_out(cresi, resi);
    lens[3] = seq_out(cresj, resj);
    if (dir_out) memcpy(dir_out, dm, sizeof(unsigned short) * (size_t) leni * lenj);
    free(csi); free(csj); free(cmedian); free(cmedianwg); free(cresi); free(cresj);
    return res;
}

/* Body of algn_CAML_cost_affine_3 (src/algn.c:2647-2679). */
int ref_cost_affine_3(void *cmh, void *wsh, const unsigned char *si, int li, const unsigned char *sj, int lj) {
    ref_cm *c = (ref_cm *) cmh; ref_ws *w = (ref_ws *) wsh;
    cmt ccm = &c->c; matricest cam = &w->m;
    seqt csi = mk_seq_from(si, li, li), csj = mk_seq_from(sj, lj, lj);
    int leni = li, lenj = lj, largest = leni > lenj ? leni : lenj, res;
    mat_setup_size(cam, largest, largest, 0, 0, ccm->lcm);
    int *matrix = cam->matrix, *prec = cam->precalc;
    int *cbd = matrix, *ebd = matrix + 2 * largest, *ev = matrix + 4 * largest, *eh = matrix + 6 * largest;
    int *gop = matrix + 10 * largest, *she = matrix + 11 * largest;
    if (leni <= lenj) {
        cm_precalc_4algn(ccm, cam, csj);
        initialize_matrices_affine_nobt(ccm->gap_open, csi, csj, ccm, cbd, ebd, ev, eh, prec);
        res = algn_fill_plane_3_aff_nobt(csi, csj, leni - 1, lenj - 1, ccm, eh, ev, cbd, ebd, prec, gop, she);
    } else {
        cm_precalc_4algn(ccm, cam, csi);
        initialize_matrices_affine_nobt(ccm->gap_open, csj, csi, ccm, cbd, ebd, ev, eh, prec);
        res = algn_fill_plane_3_aff_nobt(csj, csi, lenj - 1, leni - 1, ccm, eh, ev, cbd, ebd, prec, gop, she);
    }
    free(csi); free(csj);
    return res;
}

/* ---- batch drivers: the same calls in a loop over a pair list, sharded over `nthreads` workers with one
 * `struct matrices` each (the reference is single-threaded per process; poyd runs one servant process per
 * core, README:147-150 -- a worker here executes exactly a servant's inner loop).
 * Sequences: `pool` + off[]/len[] (len includes the leading gap); pairs: 2 ints (a, b) per pair.
 * mode 0: linear cost (Sequence.Align.cost_2 ordering, deltaw[] given)
 * mode 1: linear align_2 + ancestor_2 + median_2_with_gaps (DOS.median, src/seqCS.ml:757-766)
 * mode 2: affine cost (cost_affine_3)
 * mode 3: affine align (align_affine_3)
 * Outputs (any may be NULL): cost[n]; for modes 1/3 out_* rows of `stride` bytes, LEFT aligned, with
 * lens[4*n] = median, medianwg, ra(resi), rb(resj). */
typedef struct batch_arg {
    void *cmh; int mode; const unsigned char *pool; const long long *off; const int *len;
    const int *pairs; const int *deltaw; int lo, hi; int *cost;
    unsigned char *median, *medianwg, *ra, *rb; int *lens; long long stride;
} batch_arg;

static void *batch_worker(void *p) {
    batch_arg *a = (batch_arg *) p;
    void *ws = ref_ws_create();
    int maxcap = 0;
    for (int k = a->lo; k < a->hi; k++) {
        int c = a->len[a->pairs[2 * k]] + a->len[a->pairs[2 * k + 1]] + 2;
        if (c > maxcap) maxcap = c;
    }
    unsigned char *t0 = malloc(maxcap + 16), *t1 = malloc(maxcap + 16), *t2 = malloc(maxcap + 16),
                  *t3 = malloc(maxcap + 16);
    for (int k = a->lo; k < a->hi; k++) {
        int ia = a->pairs[2 * k], ib = a->pairs[2 * k + 1];
        const unsigned char *sa = a->pool + a->off[ia], *sb = a->pool + a->off[ib];
        int la = a->len[ia], lb = a->len[ib], cost = 0, lens[4] = {0, 0, 0, 0};
        if (a->mode == 0) {
            cost = (la >= lb) ? ref_cost_2(a->cmh, ws, sa, la, sb, lb, a->deltaw[k])
                              : ref_cost_2(a->cmh, ws, sb, lb, sa, la, a->deltaw[k]);
        } else if (a->mode == 1) {
            int rl;
            /* Sequence.Align.align_2 -> cost_2 + create_edited_2 (src/sequence.ml:813-823, 860-861) */
            if (la >= lb) cost = ref_align_2(a->cmh, ws, sa, la, sb, lb, a->deltaw[k], 1, t2, t3, &rl);
            else cost = ref_align_2(a->cmh, ws, sb, lb, sa, la, a->deltaw[k], 0, t3, t2, &rl);
            lens[2] = lens[3] = rl;
            lens[0] = ref_median_2(a->cmh, 0, t2, t3, rl, t0);
            lens[1] = ref_median_2(a->cmh, 1, t2, t3, rl, t1);
        } else if (a->mode == 2) {
            cost = ref_cost_affine_3(a->cmh, ws, sa, la, sb, lb);
        } else {
            cost = ref_align_affine_3(a->cmh, ws, sa, la, sb, lb, t0, t1, t2, t3, lens, NULL);
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
    ref_ws_free(ws);
    return NULL;
}

int ref_batch(void *cmh, int mode, const unsigned char *pool, const long long *off, const int *len, const int *pairs,
              const int *deltaw, int n, int nthreads, int *cost, unsigned char *median, unsigned char *medianwg,
              unsigned char *ra, unsigned char *rb, int *lens, long long stride) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    batch_arg *args = malloc(sizeof(batch_arg) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        batch_arg b = {cmh, mode, pool, off, len, pairs, deltaw, (int) ((long long) n * t / nthreads),
                       (int) ((long long) n * (t + 1) / nthreads), cost, median, medianwg, ra, rb, lens, stride};
        args[t] = b;
        if (nthreads == 1) batch_worker(&args[t]);
        else pthread_create(&th[t], NULL, batch_worker, &args[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(args);
    return 0;
}

/* ---- 3-D: algn_CAML_simple_3 / backtrack_3d / median_3 (src/algn.c:3458-3475, 3960-3985, 4225-4235) ------------
 * The reference's cube fill is defective; this driver runs it as it is. */
extern cm_3dt cm_set_val_3d(int a_sz, int combinations, int do_aff, int gap_open, int all_elements, cm_3dt res);
extern int algn_nw_3d(const seqt s1, const seqt s2, const seqt s3, const cm_3dt c, matricest m, int w);
extern void backtrack_3d(const seqt s1, const seqt s2, seqt s3, seqt r1, seqt r2, seqt r3, matricest m, const cm_3dt c);
extern void algn_get_median_3d(seqt s1, seqt s2, seqt s3, cm_3dt m, seqt sm);

typedef struct ref_cm3 { struct cm_3d c; } ref_cm3;

/* cost3 / median3: (1 << lcm)^3 entries indexed ((a << lcm) + b) << lcm) + c (src/cm.c:509-523). */
void *ref_cm3_create(int a_sz_in, int combinations, int cost_model, int gap_open, int all_elements, const int *cost3,
                     const unsigned char *median3) {
    ref_cm3 *r = (ref_cm3 *) calloc(1, sizeof(ref_cm3));
    cm_set_val_3d(a_sz_in, combinations, cost_model, gap_open, all_elements, &r->c);
    size_t n = (size_t) 1 << (3 * r->c.lcm);
    memcpy(r->c.cost, cost3, n * sizeof(int));
    memcpy(r->c.median, median3, n);
    return r;
}
const struct cm_3d *ref_cm3_struct(void *h) { return &((ref_cm3 *) h)->c; }
void ref_cm3_free(void *h) { ref_cm3 *r = (ref_cm3 *) h; free(r->c.cost); free(r->c.median); free(r); }

/* Returns the cost.  If r1 != NULL also runs backtrack_3d, but only after checking on the direction cube that the
 * reference's walk never indexes a sequence below 0 (it does not check, src/algn.c:3871-3903); *status = 1 and no
 * walk otherwise.  r1..r3 need l1+l2+l3 bytes, med l1+l2+l3+1. */
int ref_align_3(void *cmh, void *wsh, const unsigned char *s1, int l1, const unsigned char *s2, int l2,
                const unsigned char *s3, int l3, unsigned char *r1, unsigned char *r2, unsigned char *r3, int *rlen,
                unsigned char *med, int *medlen, int *status, unsigned short *dir_out) {
    ref_cm3 *c = (ref_cm3 *) cmh; ref_ws *w = (ref_ws *) wsh;
    seqt a = mk_seq_from(s1, l1, l1), b = mk_seq_from(s2, l2, l2), d = mk_seq_from(s3, l3, l3);
    int res = algn_nw_3d(a, b, d, &c->c, &w->m, 0);
    if (dir_out) memcpy(dir_out, w->m.cube_d, sizeof(unsigned short) * (size_t) l1 * l2 * l3);
    if (status) *status = 0;
    if (r1) {
        /* dry run of the walk (same order of tests as backtrack_3d) */
        long long p = (long long) l1 * l2 * l3 - 1, plane = (long long) l2 * l3, line = l3;
        int i1 = l1 - 1, i2 = l2 - 1, i3 = l3 - 1, ok = 1, n = 0;
        const unsigned short *dm = w->m.cube_d;
        while (p > 0) {
            int v = dm[p], u1 = 0, u2 = 0, u3 = 0;
            if (v & 16) { u1 = u2 = u3 = 1; p -= plane + line + 1; }
            else if (v & 32) { u1 = u3 = 1; p -= plane + 1; }
            else if (v & 8) { u2 = u3 = 1; p -= line + 1; }
            else if (v & 4) { u1 = 1; p -= plane; }
            else if (v & 64) { u3 = 1; p -= 1; }
            else if (v & 1) { u2 = 1; p -= line; }
            else if (v & 2) { u1 = u2 = 1; p -= plane + line; }
            else { ok = 0; break; }
            if ((u1 && i1 < 0) || (u2 && i2 < 0) || (u3 && i3 < 0)) { ok = 0; break; }
            i1 -= u1; i2 -= u2; i3 -= u3;
            if (++n > l1 + l2 + l3) { ok = 0; break; }
        }
        if (!ok) { if (status) *status = 1; *rlen = 0; *medlen = 0; }
        else {
            int cap = l1 + l2 + l3;
            seqt x = mk_seq(cap), y = mk_seq(cap), z = mk_seq(cap), sm = mk_seq(cap + 1);
            backtrack_3d(a, b, d, x, y, z, &w->m, &c->c);
            *rlen = seq_out(x, r1); seq_out(y, r2); seq_out(z, r3);
            algn_get_median_3d(x, y, z, &c->c, sm);
            *medlen = seq_out(sm, med);
            free(x); free(y); free(z); free(sm);
        }
    }
    free(a); free(b); free(d);
    return res;
}
