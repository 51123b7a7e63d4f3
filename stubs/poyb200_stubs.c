/* poyb200_stubs.c -- OCaml-side binding of libpoyb200 for POY / poyd.
 *
 * Drop this file into the reference's src/ (next to algn.c), add it to libpoycside.clib and link with -lpoyb200.
 * It provides
 *   (1) replacements for the alignment externals of src/sequence.ml that keep their names' semantics and argument
 *       lists but run on the GPU as a batch of one (poyb200_CAML_* -- switch an `external` to them by changing the
 *       quoted symbol only), and
 *   (2) the batched externals the new batching layer in seqCS.ml / allDirChar.ml calls.
 *
 * It is compiled in this repository only as a syntax/ABI check against stand-in OCaml headers (oracle/shim); the
 * image has no OCaml toolchain.  Memory discipline follows the reference's stubs: arguments are rooted with
 * CAMLparam, `struct seq` pointers are re-derived at entry (Seq_custom_val, src/seq.h:33-37) because the GC may have
 * moved the blocks, and nothing allocates on the OCaml heap while raw pointers are live.
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
static const struct cm *g_cm_loaded = NULL;
static int g_cm_model = -1, g_cm_go = -1;

static void fail_with_ctx(const char *what) {
    static char msg[512];
    snprintf(msg, sizeof msg, "%s: %s", what, g_ctx ? poyb200_last_error(g_ctx) : "no GPU context");
    failwith(msg); /* OCaml Failure, the reference's own error convention (src/matrices.c:103-117) */
}

static poyb200_ctx *ctx_for(const struct cm *c) {
    if (!g_ctx && poyb200_create(-1, &g_ctx) != POYB200_OK) failwith("poyb200: no usable CUDA device");
    /* cost matrices are mutable OCaml-side (set_affine clones first, src/data.ml:3283-3289): reload when the block
     * or its model changes */
    if (g_cm_loaded != c || g_cm_model != c->cost_model_type || g_cm_go != c->gap_open) {
        poyb200_cm m = {c->a_sz, c->lcm, c->gap, c->cost_model_type, c->combinations, c->gap_open, c->is_metric,
                        c->all_elements, c->cost, c->median, c->worst, c->prepend_cost, c->tail_cost};
        if (poyb200_set_cm(g_ctx, &m) != POYB200_OK) fail_with_ctx("poyb200_set_cm");
        g_cm_loaded = c; g_cm_model = c->cost_model_type; g_cm_go = c->gap_open;
    }
    return g_ctx;
}

/* Copies a right-aligned result row into an empty OCaml sequence exactly as repeated seq_prepend would. */
static void fill_seq(seqt dst, const uint8_t *row_end, int len) {
    if (len > dst->cap) failwith("poyb200: result longer than the preallocated sequence");
    dst->begin = dst->end - len + 1;
    dst->len = len;
    memcpy(dst->begin, row_end - len, (size_t) len);
}

/* ---- (1) single-call replacements ------------------------------------------------------------------------- */

/* replaces algn_CAML_cost_affine_3 (src/algn.c:2628) */
value poyb200_CAML_cost_affine_3(value si, value sj, value cm, value am) {
    CAMLparam4(si, sj, cm, am);
    seqt a, b;
    Seq_custom_val(a, si);
    Seq_custom_val(b, sj);
    poyb200_ctx *ctx = ctx_for(Cost_matrix_struct(cm));
    int la = a->len, lb = b->len, cost = 0;
    uint8_t *pool = (uint8_t *) malloc((size_t) la + lb + 32);
    memcpy(pool, a->begin, la);
    memcpy(pool + la, b->begin, lb);
    int64_t off[2] = {0, la};
    int32_t len[2] = {la, lb}, pairs[2] = {0, 1};
    poyb200_batch bt = {0};
    bt.pool = pool; bt.pool_bytes = (size_t) la + lb; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = 2;
    bt.pairs = pairs; bt.n_pairs = 1; bt.cost = &cost;
    int rc = poyb200_batch_cost_affine_3(ctx, &bt);
    free(pool);
    if (rc != POYB200_OK) fail_with_ctx("poyb200_batch_cost_affine_3");
    CAMLreturn(Val_int(cost));
}

/* replaces algn_CAML_align_affine_3 (src/algn.c:2551): fills resi, resj, median, medianwg; returns the cost */
value poyb200_CAML_align_affine_3(value si, value sj, value cm, value am, value resi, value resj, value median,
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
    int la = a->len, lb = b->len, cost = 0, stride = (la + lb + 2 + 15) & ~15, olen[4];
    uint8_t *buf = (uint8_t *) malloc((size_t) la + lb + 32 + 4 * (size_t) stride);
    uint8_t *pool = buf + 4 * (size_t) stride;
    memcpy(pool, a->begin, la);
    memcpy(pool + la, b->begin, lb);
    int64_t off[2] = {0, la};
    int32_t len[2] = {la, lb}, pairs[2] = {0, 1};
    poyb200_batch bt = {0};
    bt.pool = pool; bt.pool_bytes = (size_t) la + lb; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = 2;
    bt.pairs = pairs; bt.n_pairs = 1; bt.cost = &cost;
    bt.want = POYB200_WANT_MEDIAN | POYB200_WANT_MEDIANWG | POYB200_WANT_ALIGNED;
    bt.median = buf; bt.medianwg = buf + stride; bt.aligned_a = buf + 2 * stride; bt.aligned_b = buf + 3 * stride;
    bt.out_stride = stride; bt.out_len = olen;
    int rc = poyb200_batch_align_affine_3(ctx, &bt);
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
value poyb200_CAML_align_affine_3_bc(value *argv, int argn) {
    (void) argn;
    return poyb200_CAML_align_affine_3(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6], argv[7]);
}

/* replaces algn_CAML_align_2d (src/algn.c:3987) = algn_CAML_simple_2 + algn_CAML_backtrack_2d.  s1 is the longer
 * sequence, as Sequence.Align.cost_2 / create_edited_2 guarantee (src/sequence.ml:709-714, 818-822). */
value poyb200_CAML_align_2d(value s1, value s2, value c, value a, value s1p, value s2p, value deltawh, value swaped) {
    CAMLparam5(s1, s2, c, a, s1p);
    CAMLxparam3(s2p, deltawh, swaped);
    seqt x, y, xp, yp;
    Seq_custom_val(x, s1);
    Seq_custom_val(y, s2);
    Seq_custom_val(xp, s1p);
    Seq_custom_val(yp, s2p);
    poyb200_ctx *ctx = ctx_for(Cost_matrix_struct(c));
    int l1 = x->len, l2 = y->len, cost = 0, stride = (l1 + l2 + 2 + 15) & ~15, olen[4];
    uint8_t *buf = (uint8_t *) malloc((size_t) l1 + l2 + 32 + 2 * (size_t) stride);
    uint8_t *pool = buf + 2 * (size_t) stride;
    memcpy(pool, x->begin, l1);
    memcpy(pool + l1, y->begin, l2);
    int64_t off[2] = {0, l1};
    int32_t len[2] = {l1, l2}, pairs[2] = {0, 1}, dw = Int_val(deltawh);
    uint8_t sw = (uint8_t) Bool_val(swaped);
    poyb200_batch bt = {0};
    bt.pool = pool; bt.pool_bytes = (size_t) l1 + l2; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = 2;
    bt.pairs = pairs; bt.n_pairs = 1; bt.deltaw = &dw; bt.swaped = &sw; bt.cost = &cost;
    bt.want = POYB200_WANT_ALIGNED;
    bt.aligned_a = buf; bt.aligned_b = buf + stride; bt.out_stride = stride; bt.out_len = olen;
    int rc = poyb200_batch_align_2(ctx, &bt);
    if (rc == POYB200_OK) {
        fill_seq(xp, buf + stride, olen[2]);
        fill_seq(yp, buf + 2 * stride, olen[3]);
    }
    free(buf);
    if (rc != POYB200_OK) fail_with_ctx("poyb200_batch_align_2");
    CAMLreturn(Val_int(cost));
}
value poyb200_CAML_align_2d_bc(value *argv, int argn) {
    (void) argn;
    return poyb200_CAML_align_2d(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6], argv[7]);
}

/* replaces algn_CAML_simple_2 (src/algn.c:3409) */
value poyb200_CAML_simple_2(value s1, value s2, value c, value a, value deltawh) {
    CAMLparam5(s1, s2, c, a, deltawh);
    seqt x, y;
    Seq_custom_val(x, s1);
    Seq_custom_val(y, s2);
    poyb200_ctx *ctx = ctx_for(Cost_matrix_struct(c));
    int l1 = x->len, l2 = y->len, cost = 0;
    uint8_t *pool = (uint8_t *) malloc((size_t) l1 + l2 + 32);
    memcpy(pool, x->begin, l1);
    memcpy(pool + l1, y->begin, l2);
    int64_t off[2] = {0, l1};
    int32_t len[2] = {l1, l2}, pairs[2] = {0, 1}, dw = Int_val(deltawh);
    poyb200_batch bt = {0};
    bt.pool = pool; bt.pool_bytes = (size_t) l1 + l2; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = 2;
    bt.pairs = pairs; bt.n_pairs = 1; bt.deltaw = &dw; bt.cost = &cost;
    int rc = poyb200_batch_cost_2(ctx, &bt);
    free(pool);
    if (rc != POYB200_OK) fail_with_ctx("poyb200_batch_cost_2");
    CAMLreturn(Val_int(cost));
}

/* replaces algn_CAML_ancestor_2 / _median_2_with_gaps / _median_2_no_gaps (src/algn.c:4288, 4211, 4198) */
static value median_2_common(int which, value s1, value s2, value c, value sm) {
    CAMLparam4(s1, s2, c, sm);
    seqt x, y, m;
    Seq_custom_val(x, s1);
    Seq_custom_val(y, s2);
    Seq_custom_val(m, sm);
    poyb200_ctx *ctx = ctx_for(Cost_matrix_struct(c));
    int32_t len = x->len, olen = 0;
    int64_t istride = (len + 15) & ~15, ostride = (len + 1 + 15) & ~15;
    uint8_t *buf = (uint8_t *) calloc(1, (size_t) (2 * istride + ostride));
    memcpy(buf, x->begin, len);
    memcpy(buf + istride, y->begin, len);
    int rc = poyb200_batch_median_2(ctx, which, buf, buf + istride, istride, &len, 1, buf + 2 * istride, ostride, &olen);
    if (rc == POYB200_OK) fill_seq(m, buf + 2 * istride + ostride, olen);
    free(buf);
    if (rc != POYB200_OK) fail_with_ctx("poyb200_batch_median_2");
    CAMLreturn(Val_unit);
}
value poyb200_CAML_ancestor_2(value s1, value s2, value c, value sm) { return median_2_common(0, s1, s2, c, sm); }
value poyb200_CAML_median_2_with_gaps(value s1, value s2, value c, value sm) { return median_2_common(1, s1, s2, c, sm); }
value poyb200_CAML_median_2_no_gaps(value s1, value s2, value c, value sm) { return median_2_common(2, s1, s2, c, sm); }

/* ---- (2) batched externals ----------------------------------------------------------------------------------- */

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
    const int ns = (int) Wosize_val(seqs), n = (int) Wosize_val(pairs) / 2;
    poyb200_ctx *ctx = ctx_for(Cost_matrix_struct(cm));
    int64_t *off = (int64_t *) malloc(sizeof(int64_t) * (size_t) (ns + 1));
    int32_t *len = (int32_t *) malloc(sizeof(int32_t) * (size_t) (ns + 1));
    int32_t *pr = (int32_t *) malloc(sizeof(int32_t) * 2 * (size_t) (n + 1));
    size_t total = 0;
    for (int s = 0; s < ns; s++) {
        seqt q;
        Seq_custom_val(q, Field(seqs, s));
        off[s] = (int64_t) total;
        len[s] = q->len;
        total += ((size_t) q->len + 15) & ~(size_t) 15;
    }
    uint8_t *pool = (uint8_t *) poyb200_host_alloc(total + 16);
    for (int s = 0; s < ns; s++) {
        seqt q;
        Seq_custom_val(q, Field(seqs, s));
        memcpy(pool + off[s], q->begin, (size_t) q->len);
    }
    int maxcap = 16;
    for (int p = 0; p < 2 * n; p++) pr[p] = Int_val(Field(pairs, p));
    for (int p = 0; p < n; p++) {
        int cap = len[pr[2 * p]] + len[pr[2 * p + 1]] + 2;
        if (cap > maxcap) maxcap = cap;
    }
    const int64_t stride = (maxcap + 15) & ~15;
    uint8_t *out = (uint8_t *) poyb200_host_alloc(4 * (size_t) n * stride + 16);
    int32_t *cst = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n + 1));
    int32_t *olen = (int32_t *) malloc(sizeof(int32_t) * 4 * (size_t) (n + 1));
    poyb200_batch bt = {0};
    bt.pool = pool; bt.pool_bytes = total; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = ns;
    bt.pairs = pr; bt.n_pairs = n; bt.cost = cst;
    bt.want = POYB200_WANT_MEDIAN | POYB200_WANT_MEDIANWG | POYB200_WANT_ALIGNED;
    bt.median = out; bt.medianwg = out + (size_t) n * stride; bt.aligned_a = out + 2 * (size_t) n * stride;
    bt.aligned_b = out + 3 * (size_t) n * stride; bt.out_stride = stride; bt.out_len = olen;
    int rc = poyb200_batch_align_affine_3(ctx, &bt);
    if (rc == POYB200_OK) {
        value dst[4] = {median, medianwg, resi, resj};
        for (int k = 0; k < 4; k++)
            for (int p = 0; p < n; p++) {
                seqt q;
                Seq_custom_val(q, Field(dst[k], p));
                fill_seq(q, out + ((size_t) k * n + p + 1) * stride, olen[4 * p + k]);
            }
    }
    poyb200_host_free(pool);
    poyb200_host_free(out);
    free(off); free(len); free(pr); free(olen);
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
    const int ns = (int) Wosize_val(seqs), n = (int) Wosize_val(pairs) / 2;
    struct cm *c = Cost_matrix_struct(cm);
    poyb200_ctx *ctx = ctx_for(c);
    int64_t *off = (int64_t *) malloc(sizeof(int64_t) * (size_t) (ns + 1));
    int32_t *len = (int32_t *) malloc(sizeof(int32_t) * (size_t) (ns + 1));
    int32_t *pr = (int32_t *) malloc(sizeof(int32_t) * 3 * (size_t) (n + 1)), *dw = pr + 2 * (size_t) (n + 1);
    size_t total = 0;
    for (int s = 0; s < ns; s++) {
        seqt q;
        Seq_custom_val(q, Field(seqs, s));
        off[s] = (int64_t) total;
        len[s] = q->len;
        total += ((size_t) q->len + 15) & ~(size_t) 15;
    }
    uint8_t *pool = (uint8_t *) poyb200_host_alloc(total + 16);
    for (int s = 0; s < ns; s++) {
        seqt q;
        Seq_custom_val(q, Field(seqs, s));
        memcpy(pool + off[s], q->begin, (size_t) q->len);
    }
    for (int p = 0; p < 2 * n; p++) pr[p] = Int_val(Field(pairs, p));
    for (int p = 0; p < n; p++) dw[p] = Int_val(Field(deltaw, p));
    int32_t *cst = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n + 1));
    poyb200_batch bt = {0};
    bt.pool = pool; bt.pool_bytes = total; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = ns;
    bt.pairs = pr; bt.n_pairs = n; bt.deltaw = dw; bt.cost = cst;
    int rc = (c->cost_model_type == 1) ? poyb200_batch_cost_affine_3(ctx, &bt) : poyb200_batch_cost_2(ctx, &bt);
    poyb200_host_free(pool);
    free(off); free(len); free(pr);
    if (rc != POYB200_OK) { free(cst); fail_with_ctx("poyb200_batch_cost_2"); }
    costs = caml_alloc_tuple(n);
    for (int p = 0; p < n; p++) Store_field(costs, p, Val_int(cst[p]));
    free(cst);
    CAMLreturn(costs);
}

/* ---- (3) the DOS.median payload and the uppass ------------------------------------------------------------------ */

/* Gathers the distinct operands into a pinned pool (16-byte aligned starts).  Returns the pool; fills off / len. */
static uint8_t *pool_of(value seqs, int ns, int64_t *off, int32_t *len, size_t *total_out) {
    size_t total = 0;
    for (int s = 0; s < ns; s++) {
        seqt q;
        Seq_custom_val(q, Field(seqs, s));
        off[s] = (int64_t) total;
        len[s] = q->len;
        total += ((size_t) q->len + 15) & ~(size_t) 15;
    }
    uint8_t *pool = (uint8_t *) poyb200_host_alloc(total + 16);
    for (int s = 0; s < ns; s++) {
        seqt q;
        Seq_custom_val(q, Field(seqs, s));
        memcpy(pool + off[s], q->begin, (size_t) q->len);
    }
    *total_out = total;
    return pool;
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
    const int ns = (int) Wosize_val(seqs), n = (int) Wosize_val(pairs) / 2;
    struct cm *c = Cost_matrix_struct(cm);
    poyb200_ctx *ctx = ctx_for(c);
    int64_t *off = (int64_t *) malloc(sizeof(int64_t) * (size_t) (ns + 1));
    int32_t *len = (int32_t *) malloc(sizeof(int32_t) * (size_t) (ns + 1));
    int32_t *pr = (int32_t *) malloc(sizeof(int32_t) * 3 * (size_t) (n + 1)), *dw = pr + 2 * (size_t) (n + 1);
    size_t total = 0;
    uint8_t *pool = pool_of(seqs, ns, off, len, &total);
    int maxcap = 16;
    for (int p = 0; p < 2 * n; p++) pr[p] = Int_val(Field(pairs, p));
    for (int p = 0; p < n; p++) {
        dw[p] = Int_val(Field(deltaw, p));
        const int cap = len[pr[2 * p]] + len[pr[2 * p + 1]] + 2;
        if (cap > maxcap) maxcap = cap;
    }
    const int64_t stride = (maxcap + 15) & ~15, bstride = ((stride / 8) + 3) & ~3;
    uint8_t *out = (uint8_t *) poyb200_host_alloc((size_t) n * (size_t) (stride + 3 * bstride) + 16);
    int32_t *cst = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n + 1));
    int32_t *olen = (int32_t *) malloc(sizeof(int32_t) * 4 * (size_t) (n + 1));
    poyb200_batch bt = {0};
    bt.pool = pool; bt.pool_bytes = total; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = ns;
    bt.pairs = pr; bt.n_pairs = n; bt.deltaw = dw; bt.cost = cst;
    bt.want = POYB200_WANT_MEDIAN | POYB200_WANT_BITSETS;
    bt.median = out; bt.out_stride = stride; bt.out_len = olen;
    bt.bits_a = out + (size_t) n * stride; bt.bits_b = bt.bits_a + (size_t) n * bstride;
    bt.bits_wg = bt.bits_b + (size_t) n * bstride; bt.bits_stride = bstride;
    int rc = (c->cost_model_type == 1) ? poyb200_batch_align_affine_3(ctx, &bt) : poyb200_batch_align_2(ctx, &bt);
    if (rc == POYB200_OK)
        for (int p = 0; p < n; p++) {
            seqt q;
            Seq_custom_val(q, Field(median, p));
            fill_seq(q, out + ((size_t) p + 1) * stride, olen[4 * p]);
        }
    poyb200_host_free(pool);
    free(off); free(len); free(pr);
    if (rc != POYB200_OK) { poyb200_host_free(out); free(cst); free(olen); fail_with_ctx("poyb200_batch_median"); }
    /* from here on the OCaml heap is allocated; no raw `struct seq` pointer is live any more */
    costs = caml_alloc_tuple(n);
    lens = caml_alloc_tuple(n);
    ba = caml_alloc_tuple(n);
    bb = caml_alloc_tuple(n);
    bm = caml_alloc_tuple(n);
    for (int p = 0; p < n; p++) {
        const int cols = olen[4 * p + 2], nb = (cols + 7) / 8;
        Store_field(costs, p, Val_int(cst[p]));
        Store_field(lens, p, Val_int(cols));
        const uint8_t *rows[3] = {bt.bits_a + (size_t) p * bstride, bt.bits_b + (size_t) p * bstride,
                                  bt.bits_wg + (size_t) p * bstride};
        value dst[3] = {ba, bb, bm};
        for (int k = 0; k < 3; k++) {
            str = caml_alloc_string(nb);
            bitset_of_row(rows[k], bstride, cols, Bytes_val(str), nb);
            Store_field(dst[k], p, str);
        }
    }
    poyb200_host_free(out);
    free(cst); free(olen);
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
    const int ns = (int) Wosize_val(seqs), n = (int) Wosize_val(pairs) / 2;
    struct cm *c = Cost_matrix_struct(cm);
    poyb200_ctx *ctx = ctx_for(c);
    int64_t *off = (int64_t *) malloc(sizeof(int64_t) * (size_t) (ns + 1));
    int32_t *len = (int32_t *) malloc(sizeof(int32_t) * (size_t) (ns + 1));
    int32_t *pr = (int32_t *) malloc(sizeof(int32_t) * 3 * (size_t) (n + 1)), *dw = pr + 2 * (size_t) (n + 1);
    size_t total = 0;
    uint8_t *pool = pool_of(seqs, ns, off, len, &total);
    int maxcap = 16;
    for (int p = 0; p < 2 * n; p++) pr[p] = Int_val(Field(pairs, p));
    for (int p = 0; p < n; p++) {
        dw[p] = Int_val(Field(deltaw, p));
        const int cap = len[pr[2 * p]] + len[pr[2 * p + 1]] + 2;
        if (cap > maxcap) maxcap = cap;
    }
    const int64_t stride = (maxcap + 15) & ~15;
    uint8_t *out = (uint8_t *) poyb200_host_alloc((size_t) n * (size_t) stride + 16);
    int32_t *cst = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n + 1));
    int32_t *olen = (int32_t *) malloc(sizeof(int32_t) * 4 * (size_t) (n + 1));
    poyb200_batch bt = {0};
    bt.pool = pool; bt.pool_bytes = total; bt.seq_off = off; bt.seq_len = len; bt.n_seqs = ns;
    bt.pairs = pr; bt.n_pairs = n; bt.deltaw = dw; bt.cost = cst;
    bt.want = POYB200_WANT_CLOSEST;
    bt.median = out; bt.out_stride = stride; bt.out_len = olen;
    int rc = (c->cost_model_type == 1) ? poyb200_batch_align_affine_3(ctx, &bt) : poyb200_batch_align_2(ctx, &bt);
    if (rc == POYB200_OK)
        for (int p = 0; p < n; p++) {
            seqt q;
            Seq_custom_val(q, Field(outv, p));
            fill_seq(q, out + ((size_t) p + 1) * stride, olen[4 * p]);
        }
    poyb200_host_free(pool);
    poyb200_host_free(out);
    free(off); free(len); free(pr); free(olen);
    if (rc != POYB200_OK) { free(cst); fail_with_ctx("poyb200_batch_closest"); }
    costs = caml_alloc_tuple(n);
    for (int p = 0; p < n; p++) Store_field(costs, p, Val_int(cst[p]));
    free(cst);
    CAMLreturn(costs);
}
