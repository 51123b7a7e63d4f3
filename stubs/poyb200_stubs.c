/* poyb200_stubs.c -- the reference's alignment externals (src/sequence.ml:453-762, 919) on libpoyb200.
 *
 * (1) DROP-IN SYMBOLS.  This file defines the SAME C symbols the reference's src/algn.c defines for its OCaml
 *     externals -- algn_CAML_simple_2, algn_CAML_backtrack_2d(_bc), algn_CAML_align_2d(_bc), algn_CAML_cost_affine_3,
 *     algn_CAML_align_affine_3(_bc), algn_CAML_median_2_no_gaps, algn_CAML_median_2_with_gaps, algn_CAML_ancestor_2,
 *     algn_CAML_worst_2, algn_CAML_verify_2, algn_CAML_simple_3(_bc), algn_CAML_backtrack_3d(_bc),
 *     algn_CAML_align_3d(_bc), algn_CAML_median_3, and powell_3D_align(_bc) of src/ukkCommon.c -- with the same argument lists and results, each running on the GPU
 *     as a batch of one.  No OCaml source changes: src/algn_b200.c renames the originals
 *     out of the way and this file takes their place at link time.
 * (2) BATCHED EXTERNALS (poyb200_CAML_batch_*) for the batching layer in seqCS.ml / allDirChar.ml
 *     (stubs/poyb200_batching.ml).
 *
 * The image has no OCaml toolchain; this file is compiled against stand-in <caml/...> headers (oracle/shim) into
 * oracle/_ref/libpoystubs.so and EXECUTED by tests/test_stubs.py next to the reference's own algn_CAML_* symbols
 * (oracle/_ref/libpoyref.so) on hand-built custom blocks.
 *
 * Memory discipline follows the reference's stubs: arguments are rooted with CAMLparam, `struct seq` pointers are
 * re-derived at entry (Seq_custom_val, src/seq.h:32-36) because the GC may have moved the blocks, nothing allocates on
 * the OCaml heap while raw pointers into it are live, and no OCaml value is ever copied into unrooted C storage.
 */
#include <assert.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <caml/mlvalues.h>
#include <caml/memory.h>
#include <caml/alloc.h>
#include <caml/fail.h>
#include "seq.h"
#include "cm.h"
#include "poyb200.h"

static poyb200_ctx *g_ctx = NULL;
static uint64_t g_cm_hash = 0, g_cm3_hash = 0;
static int g_cm_loaded = 0, g_cm3_loaded = 0;

static void fail_with_ctx(const char *what) {
    static char msg[512];
    snprintf(msg, sizeof msg, "%s: %s", what, g_ctx ? poyb200_last_error(g_ctx) : "no GPU context");
    failwith(msg); /* OCaml Failure, the reference's own error convention (src/matrices.c:103-117) */
}

static void *xmalloc(size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) failwith("poyb200: out of memory");
    return p;
}
static void *xpinned(size_t n) {
    void *p = poyb200_host_alloc(n ? n : 1);
    if (!p) failwith("poyb200: pinned host allocation failed");
    return p;
}

static uint64_t fnv(uint64_t h, const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *) p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

static poyb200_ctx *the_ctx(void) {
    if (!g_ctx && poyb200_create(-1, &g_ctx) != POYB200_OK) failwith("poyb200: no usable CUDA device");
    return g_ctx;
}

/* The tables live in malloc'd arrays owned by an OCaml custom block: the block can move, a new matrix can reuse an old
 * address, and cm_CAML_set_* mutate in place -- so the device copy is keyed on CONTENT (scalars + every table), a few
 * microseconds of hashing per call, never on addresses. */
static poyb200_ctx *ctx_for(const struct cm *c) {
    poyb200_ctx *ctx = the_ctx();
    const size_t dim = (size_t) 1 << c->lcm;
    uint64_t h = 1469598103934665603ull;
    h = fnv(h, c, 8 * sizeof(int)); /* a_sz .. all_elements */
    h = fnv(h, c->cost, dim * dim * sizeof(int));
    h = fnv(h, c->median, dim * dim * sizeof(SEQT));
    if (c->worst) h = fnv(h, c->worst, dim * dim * sizeof(int));
    h = fnv(h, c->prepend_cost, dim * sizeof(int));
    h = fnv(h, c->tail_cost, dim * sizeof(int));
    if (!g_cm_loaded || h != g_cm_hash) {
        poyb200_cm m = {c->a_sz, c->lcm, c->gap, c->cost_model_type, c->combinations, c->gap_open, c->is_metric,
                        c->all_elements, c->cost, c->median, c->worst, c->prepend_cost, c->tail_cost};
        if (poyb200_set_cm(ctx, &m) != POYB200_OK) fail_with_ctx("poyb200_set_cm");
        g_cm_hash = h;
        g_cm_loaded = 1;
    }
    return ctx;
}

static poyb200_ctx *ctx_for_3d(const struct cm_3d *c) {
    poyb200_ctx *ctx = the_ctx();
    const size_t n = (size_t) 1 << (3 * c->lcm);
    uint64_t h = 1469598103934665603ull;
    h = fnv(h, c, 7 * sizeof(int)); /* a_sz .. all_elements */
    h = fnv(h, c->cost, n * sizeof(int));
    h = fnv(h, c->median, n * sizeof(SEQT));
    if (!g_cm3_loaded || h != g_cm3_hash) {
        poyb200_cm3 m = {c->lcm, c->gap, c->cost, c->median};
        if (poyb200_set_cm_3d(ctx, &m) != POYB200_OK) fail_with_ctx("poyb200_set_cm_3d");
        g_cm3_hash = h;
        g_cm3_loaded = 1;
    }
    return ctx;
}

/* Copies a right-aligned result row into an empty OCaml sequence exactly as repeated seq_prepend would. */
static void fill_seq(seqt dst, const uint8_t *row_end, int len) {
    if (len > dst->cap) failwith("poyb200: result longer than the preallocated sequence");
    dst->begin = dst->end - len + 1;
    dst->len = len;
    memcpy(dst->begin, row_end - len, (size_t) len);
}

/* ---- hidden state of the split calls ------------------------------------------------------------------------------
 * The reference's algn_CAML_simple_2 leaves its direction matrix in `mat` and algn_CAML_backtrack_2d walks it
 * (src/algn.c:3409-3427, 3908-3924; likewise simple_3 / backtrack_3d).  Both always go through the single global
 * Matrix.default (src/matrix.ml:28), strictly one after the other.  Here simple_* remembers its operands and
 * backtrack_* replays the alignment with the traceback on; the operands backtrack_* is handed must be the remembered
 * ones, byte for byte (anything else would read a stale matrix in the reference too). */
typedef struct last_call {
    int valid, n, len[3], deltaw;
    uint8_t *seq[3];
} last_call;
static last_call g_last2 = {0}, g_last3 = {0};

static void remember(last_call *lc, int n, const seqt *s, int deltaw) {
    for (int k = 0; k < 3; k++) { free(lc->seq[k]); lc->seq[k] = NULL; lc->len[k] = 0; }
    for (int k = 0; k < n; k++) {
        lc->len[k] = s[k]->len;
        lc->seq[k] = (uint8_t *) xmalloc((size_t) s[k]->len + 1);
        memcpy(lc->seq[k], s[k]->begin, (size_t) s[k]->len);
    }
    lc->n = n; lc->deltaw = deltaw; lc->valid = 1;
}
static int remembered(const last_call *lc, int n, const seqt *s) {
    if (!lc->valid || lc->n != n) return 0;
    for (int k = 0; k < n; k++)
        if (lc->len[k] != s[k]->len || memcmp(lc->seq[k], s[k]->begin, (size_t) s[k]->len) != 0) return 0;
    return 1;
}

/* ---- two sequences, linear gaps ------------------------------------------------------------------------------------- */

/* One pair through poyb200_batch_cost_2 / poyb200_batch_align_2.  r1 / r2 may be NULL (cost only). */
static int pair_linear(const struct cm *c, seqt x, seqt y, int deltaw, int swaped, seqt r1, seqt r2) {
    poyb200_ctx *ctx = ctx_for(c);
    const int l1 = x->len, l2 = y->len, stride = (l1 + l2 + 2 + 15) & ~15;
    int cost = 0, olen[4] = {0, 0, 0, 0};
    uint8_t *buf = (uint8_t *) xmalloc((size_t) l1 + l2 + 48 + 2 * (size_t) stride);
    uint8_t *pool = buf + 2 * (size_t) stride;
    const int o2 = (l1 + 15) & ~15;
    memcpy(pool, x->begin, (size_t) l1);
    memcpy(pool + o2, y->begin, (size_t) l2);
    int64_t off[2] = {0, o2};
    int32_t len[2] = {l1, l2}, pairs[2] = {0, 1}, dw = deltaw;
    uint8_t sw = (uint8_t) (swaped != 0);
    poyb200_batch bt;
    memset(&bt, 0, sizeof bt);
    bt.pool = pool; bt.pool_bytes = (size_t) o2 + l2; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = 2;
    bt.pairs = pairs; bt.n_pairs = 1; bt.deltaw = &dw; bt.cost = &cost;
    int rc;
    if (r1) {
        bt.swaped = &sw;
        bt.want = POYB200_WANT_ALIGNED;
        bt.aligned_a = buf; bt.aligned_b = buf + stride; bt.out_stride = stride; bt.out_len = olen;
        rc = poyb200_batch_align_2(ctx, &bt);
        if (rc == POYB200_OK) {
            fill_seq(r1, buf + stride, olen[2]);
            fill_seq(r2, buf + 2 * stride, olen[3]);
        }
    } else {
        rc = poyb200_batch_cost_2(ctx, &bt);
    }
    free(buf);
    if (rc != POYB200_OK) fail_with_ctx(r1 ? "poyb200_batch_align_2" : "poyb200_batch_cost_2");
    return cost;
}

/* algn_CAML_simple_2 (src/algn.c:3409): s1 is the longer sequence (src/sequence.ml:709-714). */
value algn_CAML_simple_2(value s1, value s2, value c, value a, value deltawh) {
    CAMLparam5(s1, s2, c, a, deltawh);
    seqt s[2];
    Seq_custom_val(s[0], s1);
    Seq_custom_val(s[1], s2);
    const int cost = pair_linear(Cost_matrix_struct(c), s[0], s[1], Int_val(deltawh), 0, NULL, NULL);
    remember(&g_last2, 2, s, Int_val(deltawh));
    CAMLreturn(Val_int(cost));
}

/* algn_CAML_backtrack_2d (src/algn.c:3908): walks what the preceding simple_2 on these operands left behind. */
value algn_CAML_backtrack_2d(value s1, value s2, value s1p, value s2p, value a, value c, value swap) {
    CAMLparam5(s1, s2, s1p, s2p, a);
    CAMLxparam2(c, swap);
    seqt s[2], r1, r2;
    Seq_custom_val(s[0], s1);
    Seq_custom_val(s[1], s2);
    Seq_custom_val(r1, s1p);
    Seq_custom_val(r2, s2p);
    if (!remembered(&g_last2, 2, s)) failwith("algn_CAML_backtrack_2d: no algn_CAML_simple_2 on these sequences precedes it");
    pair_linear(Cost_matrix_struct(c), s[0], s[1], g_last2.deltaw, Bool_val(swap), r1, r2);
    CAMLreturn(Val_unit);
}
value algn_CAML_backtrack_2d_bc(value *argv, int argn) {
    (void) argn;
    return algn_CAML_backtrack_2d(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6]);
}

/* algn_CAML_align_2d (src/algn.c:3987) = simple_2 + backtrack_2d, one GPU call. */
value algn_CAML_align_2d(value s1, value s2, value c, value a, value s1p, value s2p, value deltawh, value swaped) {
    CAMLparam5(s1, s2, c, a, s1p);
    CAMLxparam3(s2p, deltawh, swaped);
    seqt s[2], r1, r2;
    Seq_custom_val(s[0], s1);
    Seq_custom_val(s[1], s2);
    Seq_custom_val(r1, s1p);
    Seq_custom_val(r2, s2p);
    const int cost = pair_linear(Cost_matrix_struct(c), s[0], s[1], Int_val(deltawh), Bool_val(swaped), r1, r2);
    remember(&g_last2, 2, s, Int_val(deltawh));
    CAMLreturn(Val_int(cost));
}
value algn_CAML_align_2d_bc(value *argv, int argn) {
    (void) argn;
    return algn_CAML_align_2d(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6], argv[7]);
}

/* ---- two sequences, affine gaps ------------------------------------------------------------------------------------- */

/* algn_CAML_cost_affine_3 (src/algn.c:2628) */
value algn_CAML_cost_affine_3(value si, value sj, value cm, value am) {
    CAMLparam4(si, sj, cm, am);
    seqt a, b;
    Seq_custom_val(a, si);
    Seq_custom_val(b, sj);
    poyb200_ctx *ctx = ctx_for(Cost_matrix_struct(cm));
    const int la = a->len, lb = b->len, ob = (la + 15) & ~15;
    int cost = 0;
    uint8_t *pool = (uint8_t *) xmalloc((size_t) ob + lb + 32);
    memcpy(pool, a->begin, (size_t) la);
    memcpy(pool + ob, b->begin, (size_t) lb);
    int64_t off[2] = {0, ob};
    int32_t len[2] = {la, lb}, pairs[2] = {0, 1};
    poyb200_batch bt;
    memset(&bt, 0, sizeof bt);
    bt.pool = pool; bt.pool_bytes = (size_t) ob + lb; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = 2;
    bt.pairs = pairs; bt.n_pairs = 1; bt.cost = &cost;
    const int rc = poyb200_batch_cost_affine_3(ctx, &bt);
    free(pool);
    if (rc != POYB200_OK) fail_with_ctx("poyb200_batch_cost_affine_3");
    CAMLreturn(Val_int(cost));
}

/* algn_CAML_align_affine_3 (src/algn.c:2551): fills resi, resj, median, medianwg; returns the cost */
value algn_CAML_align_affine_3(value si, value sj, value cm, value am, value resi, value resj, value median,
                               value medianwg) {
    CAMLparam4(si, sj, cm, am);
    CAMLxparam4(resi, resj, median, medianwg);
    seqt a, b, ri, rj, md, mw;
    Seq_custom_val(a, si);
    Seq_custom_val(b, sj);
    Seq_custom_val(ri, resi);
    Seq_custom_val(rj, resj);
    Seq_custom_val(md, median);
    Seq_custom_val(mw, medianwg);
    poyb200_ctx *ctx = ctx_for(Cost_matrix_struct(cm));
    const int la = a->len, lb = b->len, ob = (la + 15) & ~15, stride = (la + lb + 2 + 15) & ~15;
    int cost = 0, olen[4] = {0, 0, 0, 0};
    uint8_t *buf = (uint8_t *) xmalloc((size_t) ob + lb + 32 + 4 * (size_t) stride);
    uint8_t *pool = buf + 4 * (size_t) stride;
    memcpy(pool, a->begin, (size_t) la);
    memcpy(pool + ob, b->begin, (size_t) lb);
    int64_t off[2] = {0, ob};
    int32_t len[2] = {la, lb}, pairs[2] = {0, 1};
    poyb200_batch bt;
    memset(&bt, 0, sizeof bt);
    bt.pool = pool; bt.pool_bytes = (size_t) ob + lb; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = 2;
    bt.pairs = pairs; bt.n_pairs = 1; bt.cost = &cost;
    bt.want = POYB200_WANT_MEDIAN | POYB200_WANT_MEDIANWG | POYB200_WANT_ALIGNED;
    bt.median = buf; bt.medianwg = buf + stride; bt.aligned_a = buf + 2 * stride; bt.aligned_b = buf + 3 * stride;
    bt.out_stride = stride; bt.out_len = olen;
    const int rc = poyb200_batch_align_affine_3(ctx, &bt);
    if (rc == POYB200_OK) {
        fill_seq(md, buf + stride, olen[0]);
        fill_seq(mw, buf + 2 * stride, olen[1]);
        fill_seq(ri, buf + 3 * stride, olen[2]);
        fill_seq(rj, buf + 4 * stride, olen[3]);
    }
    free(buf);
    if (rc != POYB200_OK) fail_with_ctx("poyb200_batch_align_affine_3");
    CAMLreturn(Val_int(cost));
}
value algn_CAML_align_affine_3_bc(value *argv, int argn) {
    (void) argn;
    return algn_CAML_align_affine_3(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6], argv[7]);
}

/* ---- functions of an aligned pair ------------------------------------------------------------------------------------ */

/* algn_CAML_ancestor_2 / _median_2_with_gaps / _median_2_no_gaps (src/algn.c:4288, 4211, 4198) */
static value median_2_common(int which, value s1, value s2, value c, value sm) {
    CAMLparam4(s1, s2, c, sm);
    seqt x, y, m;
    Seq_custom_val(x, s1);
    Seq_custom_val(y, s2);
    Seq_custom_val(m, sm);
    poyb200_ctx *ctx = ctx_for(Cost_matrix_struct(c));
    int32_t len = x->len, olen = 0;
    const int64_t istride = (len + 15) & ~15, ostride = (len + 1 + 15) & ~15;
    uint8_t *buf = (uint8_t *) xmalloc((size_t) (2 * istride + ostride) + 16);
    memset(buf, 0, (size_t) (2 * istride + ostride));
    memcpy(buf, x->begin, (size_t) len);
    memcpy(buf + istride, y->begin, (size_t) len);
    const int rc = poyb200_batch_median_2(ctx, which, buf, buf + istride, istride, &len, 1, buf + 2 * istride, ostride, &olen);
    if (rc == POYB200_OK) fill_seq(m, buf + 2 * istride + ostride, olen);
    free(buf);
    if (rc != POYB200_OK) fail_with_ctx("poyb200_batch_median_2");
    CAMLreturn(Val_unit);
}
value algn_CAML_ancestor_2(value s1, value s2, value c, value sm) { return median_2_common(0, s1, s2, c, sm); }
value algn_CAML_median_2_with_gaps(value s1, value s2, value c, value sm) { return median_2_common(1, s1, s2, c, sm); }
value algn_CAML_median_2_no_gaps(value s1, value s2, value c, value sm) { return median_2_common(2, s1, s2, c, sm); }

/* algn_CAML_worst_2 / algn_CAML_verify_2 (src/algn.c:3382, 3395): algn_calculate_from_2_aligned over c->worst / c->cost */
static value calc_aligned_common(int which, value s1, value s2, value c) {
    CAMLparam3(s1, s2, c);
    seqt x, y;
    Seq_custom_val(x, s1);
    Seq_custom_val(y, s2);
    poyb200_ctx *ctx = ctx_for(Cost_matrix_struct(c));
    int32_t len = x->len, res = 0;
    const int64_t istride = (len + 15) & ~15;
    uint8_t *buf = (uint8_t *) xmalloc((size_t) (2 * istride) + 16);
    memset(buf, 0, (size_t) (2 * istride));
    memcpy(buf, x->begin, (size_t) len);
    memcpy(buf + istride, y->begin, (size_t) len);
    const int rc = poyb200_batch_worst_2(ctx, which, buf, buf + istride, istride, &len, 1, &res);
    free(buf);
    if (rc != POYB200_OK) fail_with_ctx("poyb200_batch_worst_2");
    CAMLreturn(Val_int(res));
}
value algn_CAML_worst_2(value s1, value s2, value c) { return calc_aligned_common(0, s1, s2, c); }
value algn_CAML_verify_2(value s1, value s2, value c) { return calc_aligned_common(1, s1, s2, c); }

/* ---- three sequences -------------------------------------------------------------------------------------------------- */

/* One triple through poyb200_batch_align_3; r[] may be NULL (cost only). */
static int triple(const struct cm_3d *c, const seqt *s, seqt *r) {
    poyb200_ctx *ctx = ctx_for_3d(c);
    int64_t off[3];
    int32_t len[3], tri[3] = {0, 1, 2}, cost = 0, olen = 0, status = 0;
    size_t total = 0;
    for (int k = 0; k < 3; k++) {
        off[k] = (int64_t) total;
        len[k] = s[k]->len;
        total += ((size_t) s[k]->len + 15) & ~(size_t) 15;
    }
    const int64_t stride = ((int64_t) len[0] + len[1] + len[2] + 15) & ~15ll;
    uint8_t *buf = (uint8_t *) xmalloc(total + 32 + 3 * (size_t) stride);
    uint8_t *pool = buf + 3 * (size_t) stride;
    for (int k = 0; k < 3; k++) memcpy(pool + off[k], s[k]->begin, (size_t) s[k]->len);
    poyb200_batch3 bt;
    memset(&bt, 0, sizeof bt);
    bt.pool = pool; bt.pool_bytes = total; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = 3;
    bt.triples = tri; bt.n_triples = 1; bt.cost = &cost;
    if (r) {
        bt.want = POYB200_WANT3_ALIGNED;
        bt.aligned_1 = buf; bt.aligned_2 = buf + stride; bt.aligned_3 = buf + 2 * stride;
        bt.out_stride = stride; bt.out_len = &olen; bt.status = &status;
    }
    const int rc = poyb200_batch_align_3(ctx, &bt);
    if (rc == POYB200_OK && r && status == 0)
        for (int k = 0; k < 3; k++) fill_seq(r[k], buf + (size_t) (k + 1) * stride, olen);
    free(buf);
    if (rc != POYB200_OK) fail_with_ctx("poyb200_batch_align_3");
    /* the reference's walk would index a sequence below its start here (it does not check, src/algn.c:3871-3903) */
    if (r && status != 0) failwith("algn_CAML_backtrack_3d: the reference's traceback leaves the sequences on this input");
    return cost;
}

/* algn_CAML_simple_3 (src/algn.c:3458); `uk` is ignored by the reference's fill too (algn_fill_cube, :2885) */
value algn_CAML_simple_3(value s1, value s2, value s3, value c, value a, value uk) {
    CAMLparam5(s1, s2, s3, c, a);
    CAMLxparam1(uk);
    seqt s[3];
    Seq_custom_val(s[0], s1);
    Seq_custom_val(s[1], s2);
    Seq_custom_val(s[2], s3);
    const int cost = triple(Cost_matrix_struct_3d(c), s, NULL);
    remember(&g_last3, 3, s, 0);
    CAMLreturn(Val_int(cost));
}
value algn_CAML_simple_3_bc(value *argv, int argn) {
    (void) argn;
    return algn_CAML_simple_3(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5]);
}

/* algn_CAML_backtrack_3d (src/algn.c:3961) */
value algn_CAML_backtrack_3d(value s1, value s2, value s3, value s1p, value s2p, value s3p, value a, value c) {
    CAMLparam5(s1, s2, s1p, s2p, a);
    CAMLxparam3(s3, s3p, c);
    seqt s[3], r[3];
    Seq_custom_val(s[0], s1);
    Seq_custom_val(s[1], s2);
    Seq_custom_val(s[2], s3);
    Seq_custom_val(r[0], s1p);
    Seq_custom_val(r[1], s2p);
    Seq_custom_val(r[2], s3p);
    if (!remembered(&g_last3, 3, s)) failwith("algn_CAML_backtrack_3d: no algn_CAML_simple_3 on these sequences precedes it");
    triple(Cost_matrix_struct_3d(c), s, r);
    CAMLreturn(Val_unit);
}
value algn_CAML_backtrack_3d_bc(value *argv, int argn) {
    (void) argn;
    return algn_CAML_backtrack_3d(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6], argv[7]);
}

/* algn_CAML_align_3d (src/algn.c:4005) = simple_3 + backtrack_3d, one GPU call */
value algn_CAML_align_3d(value s1, value s2, value s3, value c, value a, value s1p, value s2p, value s3p, value uk) {
    CAMLparam5(s1, s2, s3, c, a);
    CAMLxparam4(s1p, s2p, s3p, uk);
    seqt s[3], r[3];
    Seq_custom_val(s[0], s1);
    Seq_custom_val(s[1], s2);
    Seq_custom_val(s[2], s3);
    Seq_custom_val(r[0], s1p);
    Seq_custom_val(r[1], s2p);
    Seq_custom_val(r[2], s3p);
    const int cost = triple(Cost_matrix_struct_3d(c), s, r);
    remember(&g_last3, 3, s, 0);
    CAMLreturn(Val_int(cost));
}
value algn_CAML_align_3d_bc(value *argv, int argn) {
    (void) argn;
    return algn_CAML_align_3d(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6], argv[7], argv[8]);
}

/* algn_CAML_median_3 (src/algn.c:4224) */
value algn_CAML_median_3(value s1, value s2, value s3, value m, value sm) {
    CAMLparam5(s1, s2, s3, m, sm);
    seqt x, y, z, out;
    Seq_custom_val(x, s1);
    Seq_custom_val(y, s2);
    Seq_custom_val(z, s3);
    Seq_custom_val(out, sm);
    poyb200_ctx *ctx = ctx_for_3d(Cost_matrix_struct_3d(m));
    int32_t len = x->len, olen = 0;
    const int64_t stride = (len + 15) & ~15;
    uint8_t *buf = (uint8_t *) xmalloc((size_t) (4 * stride) + 16);
    memset(buf, 0, (size_t) (4 * stride));
    /* the reference reads the LAST element of each sequence (seq_get_end) and the length of s1: rows end at len - 1 */
    const seqt in[3] = {x, y, z};
    for (int k = 0; k < 3; k++) {
        const int lk = in[k]->len < len ? in[k]->len : len;
        memcpy(buf + k * stride + (len - lk), in[k]->end - lk + 1, (size_t) lk);
    }
    const int rc = poyb200_batch_median_3(ctx, buf, buf + stride, buf + 2 * stride, stride, &len, 1, buf + 3 * stride, stride, &olen);
    if (rc == POYB200_OK && olen > 0) fill_seq(out, buf + 4 * stride, olen);
    free(buf);
    if (rc != POYB200_OK) fail_with_ctx("poyb200_batch_median_3");
    CAMLreturn(Val_unit);
}

/* powell_3D_align (src/ukkCommon.c:110-145; the external behind Sequence.Align.align_3_powell, src/sequence.ml:1075-1087):
 * this symbol replaces ukkCommon.o + ukk.checkp.o at link time.  ra / rb / rc arrive empty with capacity |sa| + |sb| + |sc|. */
value powell_3D_align(value sa, value sb, value sc, value ra, value rb, value rc, value mm, value go, value ge) {
    CAMLparam5(sa, sb, sc, mm, go);
    CAMLxparam4(ge, ra, rb, rc);
    seqt s[3], r[3];
    Seq_custom_val(s[0], sa);
    Seq_custom_val(s[1], sb);
    Seq_custom_val(s[2], sc);
    Seq_custom_val(r[0], ra);
    Seq_custom_val(r[1], rb);
    Seq_custom_val(r[2], rc);
    poyb200_ctx *ctx = the_ctx();
    int64_t off[3];
    int32_t len[3], tri[3] = {0, 1, 2}, cost = 0, olen[2] = {0, 0}, status = 0;
    size_t total = 0;
    for (int k = 0; k < 3; k++) {
        off[k] = (int64_t) total;
        len[k] = s[k]->len;
        total += ((size_t) s[k]->len + 15) & ~(size_t) 15;
    }
    const int64_t stride = ((int64_t) len[0] + len[1] + len[2] + 15) & ~15ll;
    uint8_t *buf = (uint8_t *) xmalloc(total + 32 + 3 * (size_t) stride);
    uint8_t *pool = buf + 3 * (size_t) stride;
    for (int k = 0; k < 3; k++) memcpy(pool + off[k], s[k]->begin, (size_t) s[k]->len);
    poyb200_batch3 bt;
    memset(&bt, 0, sizeof bt);
    bt.pool = pool; bt.pool_bytes = total; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = 3;
    bt.triples = tri; bt.n_triples = 1; bt.cost = &cost;
    bt.want = POYB200_WANT3_ALIGNED;
    bt.aligned_1 = buf; bt.aligned_2 = buf + stride; bt.aligned_3 = buf + 2 * stride;
    bt.out_stride = stride; bt.out_len = olen; bt.status = &status;
    const int rc3 = poyb200_batch_powell_3(ctx, &bt, Int_val(mm), Int_val(go), Int_val(ge));
    if (rc3 == POYB200_OK && status == 0)
        for (int k = 0; k < 3; k++) fill_seq(r[k], buf + (size_t) (k + 1) * stride, olen[0]);
    free(buf);
    if (rc3 != POYB200_OK) fail_with_ctx("poyb200_batch_powell_3");
    if (status == 5) failwith("This is impossible!"); /* copySequence's own message (:104) */
    if (status != 0) failwith("poyb200_batch_powell_3: internal limit reached");
    CAMLreturn(Val_int(cost));
}
value powell_3D_align_bc(value *argv, int argn) {
    (void) argn;
    return powell_3D_align(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6], argv[7], argv[8]);
}

/* ---- (2) batched externals ---------------------------------------------------------------------------------------------- */

typedef struct gathered {
    int ns, n;
    int64_t *off;
    int32_t *len, *pr, *dw;
    uint8_t *pool;
    size_t total;
    int maxcap;
} gathered;

static void gathered_free(gathered *g) {
    if (g->pool) poyb200_host_free(g->pool);
    free(g->off); free(g->len); free(g->pr);
    memset(g, 0, sizeof *g);
}

/* Distinct operands -> a pinned pool (16-byte aligned starts); pair indices and deltaw -> C arrays, range-checked. */
static void gather(gathered *g, value seqs, value pairs, value deltaw) {
    memset(g, 0, sizeof *g);
    g->ns = (int) Wosize_val(seqs);
    g->n = (int) Wosize_val(pairs) / 2;
    if (deltaw != Val_unit && (int) Wosize_val(deltaw) < g->n) failwith("poyb200: fewer deltaw values than pairs");
    g->off = (int64_t *) xmalloc(sizeof(int64_t) * (size_t) (g->ns + 1));
    g->len = (int32_t *) xmalloc(sizeof(int32_t) * (size_t) (g->ns + 1));
    g->pr = (int32_t *) xmalloc(sizeof(int32_t) * 3 * (size_t) (g->n + 1));
    g->dw = g->pr + 2 * (size_t) (g->n + 1);
    for (int s = 0; s < g->ns; s++) {
        seqt q;
        Seq_custom_val(q, Field(seqs, s));
        g->off[s] = (int64_t) g->total;
        g->len[s] = q->len;
        g->total += ((size_t) q->len + 15) & ~(size_t) 15;
    }
    g->pool = (uint8_t *) poyb200_host_alloc(g->total + 16);
    if (!g->pool) { gathered_free(g); failwith("poyb200: pinned host allocation failed"); }
    for (int s = 0; s < g->ns; s++) {
        seqt q;
        Seq_custom_val(q, Field(seqs, s));
        memcpy(g->pool + g->off[s], q->begin, (size_t) q->len);
    }
    g->maxcap = 16;
    for (int p = 0; p < g->n; p++) {
        const int ia = Int_val(Field(pairs, 2 * p)), ib = Int_val(Field(pairs, 2 * p + 1));
        if (ia < 0 || ia >= g->ns || ib < 0 || ib >= g->ns) { gathered_free(g); failwith("poyb200: pair index out of range"); }
        g->pr[2 * p] = ia; g->pr[2 * p + 1] = ib;
        g->dw[p] = (deltaw != Val_unit) ? Int_val(Field(deltaw, p)) : 0;
        const int cap = g->len[ia] + g->len[ib] + 2;
        if (cap > g->maxcap) g->maxcap = cap;
    }
}

static void batch_of(poyb200_batch *bt, const gathered *g, int32_t *cst) {
    memset(bt, 0, sizeof *bt);
    bt->pool = g->pool; bt->pool_bytes = g->total; bt->seq_off = g->off; bt->seq_len = g->len; bt->n_seqs = g->ns;
    bt->pairs = g->pr; bt->n_pairs = g->n; bt->deltaw = g->dw; bt->cost = cst;
}

/* external batch_align_affine_3 : s array -> int array -> Cost_matrix.Two_D.m -> s array -> s array -> s array ->
 *                                  s array -> int array = "poyb200_CAML_batch_align_affine_3_bc" "poyb200_CAML_batch_align_affine_3"
 * seqs: the distinct sequences; pairs: 2n indices; the four result arrays hold n preallocated empty sequences of
 * capacity len a + len b + 2 (exactly what Sequence.Align.align_affine_3 allocates per pair, src/sequence.ml:470-474).
 * Returns the n costs. */
value poyb200_CAML_batch_align_affine_3(value seqs, value pairs, value cm, value resi, value resj, value median,
                                        value medianwg) {
    CAMLparam5(seqs, pairs, cm, resi, resj);
    CAMLxparam2(median, medianwg);
    CAMLlocal1(costs);
    poyb200_ctx *ctx = ctx_for(Cost_matrix_struct(cm));
    gathered g;
    gather(&g, seqs, pairs, Val_unit);
    const int n = g.n;
    const int64_t stride = (g.maxcap + 15) & ~15;
    uint8_t *out = (uint8_t *) xpinned(4 * (size_t) n * stride + 16);
    int32_t *cst = (int32_t *) xmalloc(sizeof(int32_t) * 5 * (size_t) (n + 1)), *olen = cst + (n + 1);
    poyb200_batch bt;
    batch_of(&bt, &g, cst);
    bt.want = POYB200_WANT_MEDIAN | POYB200_WANT_MEDIANWG | POYB200_WANT_ALIGNED;
    bt.median = out; bt.medianwg = out + (size_t) n * stride; bt.aligned_a = out + 2 * (size_t) n * stride;
    bt.aligned_b = out + 3 * (size_t) n * stride; bt.out_stride = stride; bt.out_len = olen;
    const int rc = poyb200_batch_align_affine_3(ctx, &bt);
    if (rc == POYB200_OK) {
        for (int k = 0; k < 4; k++)
            for (int p = 0; p < n; p++) {
                seqt q;  /* the destination array is read through its rooted parameter every time */
                Seq_custom_val(q, Field(k == 0 ? median : k == 1 ? medianwg : k == 2 ? resi : resj, p));
                fill_seq(q, out + ((size_t) k * n + p + 1) * stride, olen[4 * p + k]);
            }
    }
    poyb200_host_free(out);
    gathered_free(&g);
    if (rc != POYB200_OK) { free(cst); fail_with_ctx("poyb200_batch_align_affine_3"); }
    costs = caml_alloc_tuple(n); /* an int array is a block of immediates */
    for (int p = 0; p < n; p++) Store_field(costs, p, Val_int(cst[p]));
    free(cst);
    CAMLreturn(costs);
}
value poyb200_CAML_batch_align_affine_3_bc(value *argv, int argn) {
    (void) argn;
    return poyb200_CAML_batch_align_affine_3(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6]);
}

/* external batch_cost_2 : s array -> int array -> int array -> Cost_matrix.Two_D.m -> int array
 * pairs: 2n indices; deltaw: n values (ignored for affine matrices, as Sequence.Align.cost_2 does, sequence.ml:716-719) */
value poyb200_CAML_batch_cost_2(value seqs, value pairs, value deltaw, value cm) {
    CAMLparam4(seqs, pairs, deltaw, cm);
    CAMLlocal1(costs);
    struct cm *c = Cost_matrix_struct(cm);
    const int affine = (c->cost_model_type == 1);
    poyb200_ctx *ctx = ctx_for(c);
    gathered g;
    gather(&g, seqs, pairs, deltaw);
    const int n = g.n;
    int32_t *cst = (int32_t *) xmalloc(sizeof(int32_t) * (size_t) (n + 1));
    poyb200_batch bt;
    batch_of(&bt, &g, cst);
    const int rc = affine ? poyb200_batch_cost_affine_3(ctx, &bt) : poyb200_batch_cost_2(ctx, &bt);
    gathered_free(&g);
    if (rc != POYB200_OK) { free(cst); fail_with_ctx("poyb200_batch_cost_2"); }
    costs = caml_alloc_tuple(n);
    for (int p = 0; p < n; p++) Store_field(costs, p, Val_int(cst[p]));
    free(cst);
    CAMLreturn(costs);
}

/* A right-aligned, most-significant-bit-first row of the library -> the byte string of an extlib BitSet of n bits
 * (bit i lives in byte i / 8 at position i mod 8). */
static void bitset_of_row(const uint8_t *row, int64_t stride, int n, unsigned char *dst, int dst_bytes) {
    memset(dst, 0, (size_t) dst_bytes);
    const int64_t first = 8 * stride - n;
    for (int i = 0; i < n; i++) {
        const int64_t p = first + i;
        if ((row[p >> 3] >> (7 - (p & 7))) & 1) dst[i >> 3] |= (unsigned char) (1 << (i & 7));
    }
}

/* external batch_median : s array -> int array -> int array -> Cost_matrix.Two_D.m -> s array ->
 *                          (int array * int array * string array * string array * string array)
 *   = "poyb200_CAML_batch_median"
 * What SeqCS.DOS.median computes for every pair (src/seqCS.ml:747-776), both gap models: the cost, the median (into
 * the n preallocated sequences of `median`), and the three gap bitsets of tmpa, tmpb, seqmwg -- returned as the
 * alignment length and three byte strings per pair, ready for `Packed (len, {BitSet.data; len}, Raw operand)`.
 * deltaw: n values for linear matrices (Sequence.Align.cost_2's deltaw, :691-714), ignored for affine ones. */
value poyb200_CAML_batch_median(value seqs, value pairs, value deltaw, value cm, value median) {
    CAMLparam5(seqs, pairs, deltaw, cm, median);
    CAMLlocal5(res, costs, lens, ba, bb);
    CAMLlocal2(bm, str);
    struct cm *c = Cost_matrix_struct(cm);
    const int affine = (c->cost_model_type == 1);
    poyb200_ctx *ctx = ctx_for(c);
    gathered g;
    gather(&g, seqs, pairs, deltaw);
    const int n = g.n;
    const int64_t stride = (g.maxcap + 15) & ~15, bstride = ((stride / 8) + 3) & ~3;
    uint8_t *out = (uint8_t *) xpinned((size_t) n * (size_t) (stride + 3 * bstride) + 16);
    int32_t *cst = (int32_t *) xmalloc(sizeof(int32_t) * 5 * (size_t) (n + 1)), *olen = cst + (n + 1);
    poyb200_batch bt;
    batch_of(&bt, &g, cst);
    bt.want = POYB200_WANT_MEDIAN | POYB200_WANT_BITSETS;
    bt.median = out; bt.out_stride = stride; bt.out_len = olen;
    bt.bits_a = out + (size_t) n * stride; bt.bits_b = bt.bits_a + (size_t) n * bstride;
    bt.bits_wg = bt.bits_b + (size_t) n * bstride; bt.bits_stride = bstride;
    const int rc = affine ? poyb200_batch_align_affine_3(ctx, &bt) : poyb200_batch_align_2(ctx, &bt);
    if (rc == POYB200_OK)
        for (int p = 0; p < n; p++) {
            seqt q;
            Seq_custom_val(q, Field(median, p));
            fill_seq(q, out + ((size_t) p + 1) * stride, olen[4 * p]);
        }
    gathered_free(&g);
    if (rc != POYB200_OK) { poyb200_host_free(out); free(cst); fail_with_ctx("poyb200_batch_median"); }
    /* From here on the OCaml heap is allocated: no raw `struct seq` pointer is live any more, and every block is
     * reached through a CAMLlocal root at the moment it is written (an allocation may move all of them). */
    costs = caml_alloc_tuple(n);
    lens = caml_alloc_tuple(n);
    ba = caml_alloc_tuple(n);
    bb = caml_alloc_tuple(n);
    bm = caml_alloc_tuple(n);
    for (int p = 0; p < n; p++) {
        const int cols = olen[4 * p + 2], nb = (cols + 7) / 8;
        Store_field(costs, p, Val_int(cst[p]));
        Store_field(lens, p, Val_int(cols));
        for (int k = 0; k < 3; k++) {
            const uint8_t *row = (k == 0 ? bt.bits_a : k == 1 ? bt.bits_b : bt.bits_wg) + (size_t) p * bstride;
            str = caml_alloc_string(nb);
            bitset_of_row(row, bstride, cols, Bytes_val(str), nb);
            Store_field(k == 0 ? ba : k == 1 ? bb : bm, p, str);
        }
    }
    poyb200_host_free(out);
    free(cst);
    res = caml_alloc_tuple(5);
    Store_field(res, 0, costs); Store_field(res, 1, lens); Store_field(res, 2, ba); Store_field(res, 3, bb);
    Store_field(res, 4, bm);
    CAMLreturn(res);
}

/* external batch_closest : s array -> int array -> int array -> Cost_matrix.Two_D.m -> s array -> int array
 *   = "poyb200_CAML_batch_closest"
 * Sequence.Align.closest s1 s2 (src/sequence.ml:967-1033) for every pair (s1 = pairs.(2p), s2 = pairs.(2p+1)) whose
 * early exits (empty s2, s1 = s2) the caller has already taken: fills `out.(p)` (capacity >= len a + len b + 2) with the
 * closest sequence and returns the alignment costs (`cst` of align_2; the re-costing of :1026 is a batch_cost_2 call). */
value poyb200_CAML_batch_closest(value seqs, value pairs, value deltaw, value cm, value outv) {
    CAMLparam5(seqs, pairs, deltaw, cm, outv);
    CAMLlocal1(costs);
    struct cm *c = Cost_matrix_struct(cm);
    const int affine = (c->cost_model_type == 1);
    poyb200_ctx *ctx = ctx_for(c);
    gathered g;
    gather(&g, seqs, pairs, deltaw);
    const int n = g.n;
    const int64_t stride = (g.maxcap + 15) & ~15;
    uint8_t *out = (uint8_t *) xpinned((size_t) n * (size_t) stride + 16);
    int32_t *cst = (int32_t *) xmalloc(sizeof(int32_t) * 5 * (size_t) (n + 1)), *olen = cst + (n + 1);
    poyb200_batch bt;
    batch_of(&bt, &g, cst);
    bt.want = POYB200_WANT_CLOSEST;
    bt.median = out; bt.out_stride = stride; bt.out_len = olen;
    const int rc = affine ? poyb200_batch_align_affine_3(ctx, &bt) : poyb200_batch_align_2(ctx, &bt);
    if (rc == POYB200_OK)
        for (int p = 0; p < n; p++) {
            seqt q;
            Seq_custom_val(q, Field(outv, p));
            fill_seq(q, out + ((size_t) p + 1) * stride, olen[4 * p]);
        }
    poyb200_host_free(out);
    gathered_free(&g);
    if (rc != POYB200_OK) { free(cst); fail_with_ctx("poyb200_batch_closest"); }
    costs = caml_alloc_tuple(n);
    for (int p = 0; p < n; p++) Store_field(costs, p, Val_int(cst[p]));
    free(cst);
    CAMLreturn(costs);
}
