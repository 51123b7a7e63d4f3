/* poyb200.h -- C ABI of the B200 (sm_100a) implementation of POY's direct-optimisation alignment path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / CUDA types.  Every batch entry point
 * replaces one OCaml `external` of the reference (declared in src/sequence.ml, implemented in src/algn.c) and
 * computes, for every pair of the batch, exactly what that external returns for the pair -- bit for bit,
 * including traceback tie-breaking.  INTEGRATION.md shows the OCaml-side stubs that bind to it.
 *
 *   reference external (file:line)                                    entry point here
 *   ---------------------------------------------------------------  -------------------------------
 *   algn_CAML_simple_2          src/algn.c:3409, sequence.ml:457      poyb200_batch_cost_2
 *   algn_CAML_align_2d          src/algn.c:3987, sequence.ml:744      poyb200_batch_align_2
 *     (= simple_2 + algn_CAML_backtrack_2d  src/algn.c:3908)
 *   algn_CAML_ancestor_2        src/algn.c:4288, sequence.ml:919      poyb200_batch_align_2 (WANT_MEDIAN) /
 *                                                                     poyb200_batch_median_2 (which = 0)
 *   algn_CAML_median_2_with_gaps src/algn.c:4211, sequence.ml:756     ... (WANT_MEDIANWG) / (which = 1)
 *   algn_CAML_median_2_no_gaps  src/algn.c:4198, sequence.ml:753      poyb200_batch_median_2 (which = 2)
 *   algn_CAML_cost_affine_3     src/algn.c:2628, sequence.ml:465      poyb200_batch_cost_affine_3
 *   algn_CAML_align_affine_3    src/algn.c:2551, sequence.ml:461      poyb200_batch_align_affine_3
 *   algn_CAML_worst_2 / verify_2 src/algn.c:3382, 3395                poyb200_batch_worst_2 (which = 0 / 1)
 *   algn_CAML_simple_3 / backtrack_3d / align_3d  src/algn.c:3458, 3961, 4005   poyb200_batch_align_3
 *   algn_CAML_median_3          src/algn.c:4224, sequence.ml:762      poyb200_batch_median_3
 * stubs/poyb200_stubs.c defines the algn_CAML_* symbols themselves on top of these entry points.
 *
 * Sequences are the reference's `struct seq` payloads (src/seq.h:48-58): one byte per element (SEQT =
 * unsigned char), the first element being the leading gap, so a sequence of n bases has length n + 1.
 * A batch names its operands through a pool: `pool` holds the bytes of all DISTINCT sequences,
 * `seq_off[s]` / `seq_len[s]` locate sequence s, and pair p aligns sequences pairs[2p] and pairs[2p+1]
 * (a candidate-edge sweep shares one operand between thousands of pairs; it is uploaded once).
 *
 * Operand order follows Sequence.Align (src/sequence.ml:691-723, 813-823, 849-869): the linear kernels put
 * the longer operand on the rows and break traceback ties with swaped = (len a >= len b); the affine kernels
 * put the shorter operand (ties: a) on the rows.  Outputs always come back in the caller's (a, b) order.
 *
 * Outputs are caller-allocated (the reference's convention, src/sequence.ml:470-474, 816-817): row p of an
 * output buffer spans [p * out_stride, (p+1) * out_stride) and holds the sequence RIGHT aligned -- the layout
 * of the reference's `struct seq`, which is filled back to front by seq_prepend (begin = end - len + 1,
 * src/seq.h:31-36, src/seq.c:147-153) -- and out_len[4p + k] gives its length (k = 0 median, 1 medianwg,
 * 2 aligned a, 3 aligned b).  out_stride must be >= the largest len a + len b + 2 of the batch.
 *
 * Errors: every function returns 0 on success or a negative POYB200_E* code; poyb200_last_error gives the
 * text (the reference raises OCaml `Failure` through failwith; the stubs in INTEGRATION.md turn a non-zero
 * return into exactly that).  There is no CPU fallback: without a usable CUDA device poyb200_create fails.
 */
#ifndef POYB200_H
#define POYB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POYB200_OK 0
#define POYB200_ECUDA (-1)    /* CUDA runtime error (text in poyb200_last_error) */
#define POYB200_EINVAL (-2)   /* bad argument (NULL where data is needed, stride too small, ...) */
#define POYB200_ENOCM (-3)    /* no cost matrix loaded */
#define POYB200_EMODEL (-4)   /* cost_model_type not supported by this entry point (SURVEY.md 8a note) */
#define POYB200_ENOMEM (-5)   /* device or pinned-host allocation failed */
#define POYB200_ESEQLEN (-6)  /* a sequence is empty or longer than POYB200_MAX_SEQ_LEN (src/seq.c:370) */

#define POYB200_MAX_SEQ_LEN 16384 /* SHORT_SEQUENCES cap of the reference, src/seq.c:359-377 */

/* what to return from the align entry points */
#define POYB200_WANT_MEDIAN 1u    /* ancestor_2 (linear) / `median` of align_affine_3 */
#define POYB200_WANT_MEDIANWG 2u  /* median_2_with_gaps (linear) / `medianwg` */
#define POYB200_WANT_ALIGNED 4u   /* the two aligned (edited) sequences */
#define POYB200_WANT_BITSETS 8u   /* the three gap bitsets SeqCS.DOS.median keeps instead of the aligned sequences:
                                     seq_to_bitset gap tmpa / tmpb / seqmwg (src/seqCS.ml:649-655, 769-771) */

/* Flat image of the reference's `struct cm` (src/cm.h:32-46).  cost/median/worst have (1<<lcm)*(1<<lcm)
 * entries indexed (a << lcm) + b (src/cm.c:501-504); prepend_cost/tail_cost have 1<<lcm entries. */
#define POYB200_WANT_CLOSEST 16u  /* instead of the median: the `median` rows receive what Sequence.Align.closest s1 s2
                                     (src/sequence.ml:967-1033, s1 = operand a, s2 = operand b) builds from the aligned
                                     pair -- get_closest column by column, gaps removed, a gap prepended --, its length in
                                     out_len[4 p]; `cost` stays the alignment cost.  The caller handles the two early exits
                                     of closest (empty s2; s1 = s2) itself.  Excludes POYB200_WANT_MEDIAN. */

typedef struct poyb200_cm {
    int32_t a_sz, lcm, gap, cost_model_type, combinations, gap_open, is_metric, all_elements;
    const int32_t *cost;
    const uint8_t *median;
    const int32_t *worst; /* may be NULL: not used by the alignment path */
    const int32_t *prepend_cost;
    const int32_t *tail_cost;
} poyb200_cm;

typedef struct poyb200_batch {
    /* inputs (host memory; pinned memory from poyb200_host_alloc makes the copies asynchronous) */
    const uint8_t *pool;
    size_t pool_bytes;
    const int64_t *seq_off;
    const int32_t *seq_len;
    int32_t n_seqs;
    const int32_t *pairs; /* 2 * n_pairs sequence indices */
    int32_t n_pairs;
    const int32_t *deltaw;       /* linear entry points: `deltawh` of algn_CAML_simple_2, one per pair */
    const uint8_t *swaped;       /* optional, linear align only: explicit traceback tie flag per pair
                                    (algn_CAML_backtrack_2d's `swap`); NULL = (len a >= len b) */
    uint32_t want;               /* POYB200_WANT_* */
    /* outputs (host memory), NULL when not wanted */
    int32_t *cost;
    uint8_t *median, *medianwg, *aligned_a, *aligned_b;
    int64_t out_stride;
    int32_t *out_len; /* 4 * n_pairs: lengths of median, medianwg, aligned a, aligned b */
    /* POYB200_WANT_BITSETS: rows of bits_stride bytes (a multiple of 4, >= out_stride / 8).  Each row holds one bit
     * per column of the alignment, 1 where the element differs from the gap code (BitSet.is_set of seq_to_bitset),
     * as a RIGHT-aligned bit string, most significant bit of a byte first: column c of an alignment of n columns
     * (n = out_len[4 p + 2]) is bit 8 * bits_stride - n + c, i.e. numpy.unpackbits(row)[-n:] is the bitset in column
     * order.  Bits in front of the string are unspecified.  With cost + median + these three rows a caller holds
     * everything DOS.median stores (the aligned sequences follow from bitset + operand, bitset_to_seq, :623-647). */
    uint8_t *bits_a, *bits_b, *bits_wg;
    int64_t bits_stride;
} poyb200_batch;

typedef struct poyb200_ctx poyb200_ctx;

/* Every tunable of a context.  Fill it with poyb200_default_config, change what you need, pass it to
 * poyb200_create_ex; the library reads no environment variables.  struct_bytes versions the struct: a caller built
 * against an older header passes a shorter one and the fields it does not know keep their defaults. */
typedef struct poyb200_config {
    uint32_t struct_bytes;             /* sizeof(poyb200_config) as the caller knows it */
    int32_t force_generic;             /* 1: every pair takes the generic (any-shape) kernels; default 0 */
    int32_t allow_fast;                /* 0: no aff_fast_kernel (pairs without gap bits take aff_stripe_kernel); default 1 */
    int32_t allow_noeb;                /* 0: no no-gap-bit variant of aff_stripe_kernel either; default 1 */
    int32_t overlap_traceback;         /* 0: traceback on the compute stream, one direction buffer; default 1 */
    int32_t dir_buffers;               /* 2 or 3 direction-band buffers; default 3 */
    int32_t traceback_threads_per_sm;  /* resident traceback walkers per SM; default 256 */
    int32_t traceback_block;           /* 32, 64 or 128 threads per traceback CTA; default 128 */
    int32_t traceback_priority;        /* 1: the traceback stream outranks the fill stream; default 1 */
    int32_t chunk_pairs;               /* pairs per chunk of a one-shot call (pipelining granularity); default 65536 */
    int32_t host_threads;              /* planner threads, 0 = min(16, hardware threads); default 0 */
    int32_t timing;                    /* 1: CUDA events per chunk for poyb200_last_run_ms; default 1 */
    int32_t trace;                     /* stderr timeline of every one-shot call: 0 off, 1 host, 2 + downloads, 3 + device */
    int64_t dir_budget_bytes;          /* size limit of one direction buffer, 0 = a quarter of the free HBM, <= 40 GB */
    int32_t use_ring;                  /* affine stripes without spare diagonals: 0 aff_fast_kernel / aff_stripe_kernel + the separate
                                          traceback kernel; 1 ring kernels (fill + traceback in one kernel) for every pair;
                                          2 aff_fast_kernel + traceback kernel for pairs without gap bits, the full ring instance
                                          for the others (default: the fastest combination measured, profiles/README.md) */
    int32_t allow_rows;                /* 0: full linear matrices take the diagonal-stripe kernels too (no lin_rows_kernel); default 1 */
    int32_t small_ring_pairs;          /* use_ring = 2 only: calls of at most this many pairs run as use_ring = 1 (latency); 0 = never */
    int32_t dir6;                      /* use_ring = 2, stripe shape (5, 8): direction band of five 6-bit codes per 32-bit word (half the
                                          bytes of the 8-byte chunks); default 1 */
    int32_t pair2;                     /* dir6 pairs without gap bits whose costs fit 16 bits: aff_x2_kernel, two pairs per lane group
                                          (DPX 16x2 instructions), before aff_fast_kernel; default 1 */
    int32_t pair2_min_pairs;           /* pair2 only for launches of at least this many pairs; default 0 */
} poyb200_config;
void poyb200_default_config(poyb200_config *cfg);

/* device < 0: keep the calling thread's current CUDA device.  poyb200_create = poyb200_create_ex with the defaults. */
int poyb200_create(int device, poyb200_ctx **out);
int poyb200_create_ex(int device, const poyb200_config *cfg, poyb200_ctx **out);
void poyb200_destroy(poyb200_ctx *ctx);
const char *poyb200_last_error(const poyb200_ctx *ctx);
const char *poyb200_version(void);

/* Copies the tables to the device; replaces cm_CAML_create + the cm_CAML_set_* calls as far as the
 * alignment path is concerned (src/cm.c:1262, cost_matrix.ml:44-75). */
int poyb200_set_cm(poyb200_ctx *ctx, const poyb200_cm *cm);

/* Pinned host memory for batch buffers. */
void *poyb200_host_alloc(size_t bytes);
void poyb200_host_free(void *p);

/* --- one-shot batch calls: H2D, kernels, D2H --------------------------------------------------------- */
int poyb200_batch_cost_2(poyb200_ctx *ctx, const poyb200_batch *b);          /* linear, cost only */
int poyb200_batch_align_2(poyb200_ctx *ctx, const poyb200_batch *b);         /* linear, cost + traceback + medians */
int poyb200_batch_cost_affine_3(poyb200_ctx *ctx, const poyb200_batch *b);   /* affine_3, cost only */
int poyb200_batch_align_affine_3(poyb200_ctx *ctx, const poyb200_batch *b);  /* affine_3, everything */

/* which: 0 algn_ancestor_2, 1 algn_get_median_2d_with_gaps, 2 algn_get_median_2d_no_gaps.
 * a/b: n rows of `in_stride` bytes holding LEFT-aligned aligned sequences of length len[p] (equal for both,
 * as Sequence.Align.median_2 demands, src/sequence.ml:909-916); out rows of `out_stride` >= len + 1 bytes,
 * RIGHT aligned, lengths in out_len[p]. */
int poyb200_batch_median_2(poyb200_ctx *ctx, int which, const uint8_t *a, const uint8_t *b, int64_t in_stride,
                           const int32_t *len, int32_t n, uint8_t *out, int64_t out_stride, int32_t *out_len);

/* algn_CAML_worst_2 (src/algn.c:3382, sequence.ml `max_cost_2`; which = 0, needs the `worst` table of the loaded matrix) and
 * algn_CAML_verify_2 (src/algn.c:3395; which = 1): algn_calculate_from_2_aligned (:3311-3371) over n already aligned pairs,
 * rows as for poyb200_batch_median_2; out[p] receives the sum. */
int poyb200_batch_worst_2(poyb200_ctx *ctx, int which, const uint8_t *a, const uint8_t *b, int64_t in_stride, const int32_t *len,
                          int32_t n, int32_t *out);

/* --- three sequences: algn_CAML_simple_3 / algn_CAML_align_3d / algn_CAML_median_3 (src/algn.c:3458-3475, 3960-3985,
 * 4225-4235; sequence.ml:727-762) ----------------------------------------------------------------------------------
 * The reference's cube fill is defective (its neighbour-row pointers lag by one row per plane) and its
 * median loop never advances (A14); "identical to the reference" therefore means reproducing those results, which is
 * what these entry points do.  status[t] = 1 marks triples whose traceback the reference would run off the start of a
 * sequence (it does not check, A13): nothing is returned for them beyond the cost. */
#define POYB200_WANT3_ALIGNED 1u
#define POYB200_WANT3_MEDIAN 2u

typedef struct poyb200_cm3 {
    int32_t lcm, gap;
    const int32_t *cost;   /* (1 << lcm)^3 entries indexed ((a << lcm) + b) << lcm) + c  (src/cm.c:509-523) */
    const uint8_t *median; /* same indexing */
} poyb200_cm3;

typedef struct poyb200_batch3 {
    const uint8_t *pool;
    size_t pool_bytes;
    const int64_t *seq_off;
    const int32_t *seq_len;
    int32_t n_seqs;
    const int32_t *triples; /* 3 * n_triples sequence indices (s1, s2, s3 of algn_nw_3d) */
    int32_t n_triples;
    uint32_t want;          /* POYB200_WANT3_* */
    int32_t *cost;          /* n_triples */
    uint8_t *aligned_1, *aligned_2, *aligned_3, *median; /* rows of out_stride >= l1 + l2 + l3 bytes, right aligned */
    int64_t out_stride;
    int32_t *out_len;       /* n_triples: aligned length (= median length) */
    int32_t *status;        /* n_triples */
} poyb200_batch3;

int poyb200_set_cm_3d(poyb200_ctx *ctx, const poyb200_cm3 *cm);
int poyb200_batch_align_3(poyb200_ctx *ctx, const poyb200_batch3 *b);
/* algn_CAML_median_3 (src/algn.c:4224, sequence.ml:762) over n already aligned triples: rows of in_stride bytes, LEFT aligned,
 * len[p] elements each; out rows RIGHT aligned.  Reproduces algn_get_median_3d as executed (a constant sequence). */
int poyb200_batch_median_3(poyb200_ctx *ctx, const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t in_stride,
                           const int32_t *len, int32_t n, uint8_t *out, int64_t out_stride, int32_t *out_len);
/* Powell's three-sequence aligner under affine gap costs (src/ukk.checkp.c + src/ukkCommon.c): the external powell_3D_align
 * (src/ukkCommon.c:110-145, sequence.ml:1075-1087) for every triple -- mismatch cost mm, gap opening go, gap extension ge as
 * Sequence.Align.align_3_powell_inter derives them from the 2-D matrix (src/sequence.ml:1089-1102).  Elements collapse to their
 * lowest base like copySequence does (:87-108); cost[t] = the edit cost, aligned_1..3 = the three rows behind one gap column
 * (POYB200_WANT3_ALIGNED), median = the 3-D median of every column with gaps dropped and one gap in front (align_3_powell_inter
 * :1103-1114; POYB200_WANT3_MEDIAN, needs poyb200_set_cm_3d).  out_len holds TWO ints per triple: aligned length, median length.
 * status[t]: 0, or 5 = an element without a base (the reference raises "This is impossible!"), other values = internal limits. */
int poyb200_batch_powell_3(poyb200_ctx *ctx, const poyb200_batch3 *b, int32_t mm, int32_t go, int32_t ge);
/* cells of the cube: l1 * l2 * l3 */
int64_t poyb200_cells_3d(int32_t l1, int32_t l2, int32_t l3);

/* --- split form, used to time the device-resident part on its own ------------------------------------ */
/* mode: 0 cost_2, 1 align_2, 2 cost_affine_3, 3 align_affine_3 */
int poyb200_stage(poyb200_ctx *ctx, int mode, const poyb200_batch *b); /* plan + H2D; keeps b for fetch */
int poyb200_run(poyb200_ctx *ctx);                                      /* kernels only, asynchronous */
int poyb200_sync(poyb200_ctx *ctx);                                     /* wait for the context's stream */
int poyb200_fetch(poyb200_ctx *ctx);                                    /* D2H into the staged batch's outputs */

/* --- device-resident sequence store ---------------------------------------------------------------------------
 * Sequences live in HBM; batches name their operands by store id and what they produce is appended to the store ON
 * THE DEVICE, so a tree level (the medians of one downpass depth, the single assignments of one uppass depth) never
 * crosses the host link with its sequences: per batch 56 bytes per pair go up, 12 come back.  This is what the tree
 * driver (poyb200_tree.h) runs on.  Reference semantics folded in: the deltaw of Sequence.Align.cost_2
 * (src/sequence.ml:691-714) is derived from the per-sequence gap counts the store keeps (seq_CAML_count,
 * src/seq.c:570-582); poyb200_store_closest takes both early exits of Sequence.Align.closest (src/sequence.ml:975-1009).
 * A store belongs to one context and its cost matrix; it is limited to 4 GiB. */
typedef struct poyb200_store poyb200_store;
int poyb200_store_create(poyb200_ctx *ctx, poyb200_store **out);
void poyb200_store_destroy(poyb200_store *s);
int32_t poyb200_store_size(const poyb200_store *s);   /* number of sequences */
int64_t poyb200_store_bytes(const poyb200_store *s);  /* bytes of HBM in use */
/* uploads n sequences (bytes + off[k], len[k] elements, leading gap included); ids first_id .. first_id + n - 1 */
int poyb200_store_add(poyb200_store *s, const uint8_t *bytes, const int64_t *off, const int32_t *len, int32_t n, int32_t *first_id);
int poyb200_store_info(const poyb200_store *s, int32_t id, int32_t *len, int32_t *empty, int32_t *gap_count);
int poyb200_store_get(poyb200_store *s, int32_t id, uint8_t *out); /* len bytes to the host */
/* SeqCS.DOS.median for n pairs of store ids (src/seqCS.ml:747-776; both operands non-empty -- the empty-operand rule,
 * :748-752, needs no alignment and is the caller's): cost[p] and the id of the new median new_id[p] */
int poyb200_store_median(poyb200_store *s, const int32_t *pairs, int32_t n, int32_t *cost, int32_t *new_id);
/* SeqCS.DOS.distance / Sequence.Align.cost_2 (src/seqCS.ml:819-867): cost only.  hint[p] = the ?deltaw argument
 * (DOS.distance passes max 8 |la - lb|), NULL = none; ignored for affine matrices. */
int poyb200_store_distance(poyb200_store *s, const int32_t *pairs, int32_t n, const int32_t *hint, int32_t *cost);
/* Sequence.Align.closest s1 s2 (src/sequence.ml:967-1033) for n pairs (s1, s2): new_id[p] = the closest sequence
 * (s2 itself when it is empty) */
int poyb200_store_closest(poyb200_store *s, const int32_t *pairs, int32_t n, int32_t *new_id);
void poyb200_store_stats(const poyb200_store *s, int64_t *calls, int64_t *pairs, int64_t *cells);

/* --- one batch over several GPUs of the node ----------------------------------------------------------------
 * BASELINE.json north_star: "batches shard across the 8 B200s by pair index, with results gathered on the host".
 * One context and one host thread per device.  The pair list is cut into contiguous ranges of about equal DP work;
 * each device receives only the window of the pool its pairs reference and writes its results straight into the caller's
 * buffers at its pairs' rows (the batch's buffers are shared by all shards: no gather copy).  No collective is involved:
 * pairs are independent.  devices = NULL means 0 .. n_devices-1.  The whole pool may exceed 4 GiB here;
 * each shard's window must not. */
typedef struct poyb200_multi poyb200_multi;
int poyb200_multi_create(const int *devices, int n_devices, const poyb200_config *cfg, poyb200_multi **out);
void poyb200_multi_destroy(poyb200_multi *m);
const char *poyb200_multi_last_error(const poyb200_multi *m);
int poyb200_multi_set_cm(poyb200_multi *m, const poyb200_cm *cm);
/* mode: 0 cost_2, 1 align_2, 2 cost_affine_3, 3 align_affine_3 -- the one-shot call of that mode, sharded */
int poyb200_multi_batch(poyb200_multi *m, int mode, const poyb200_batch *b);
int poyb200_multi_devices(const poyb200_multi *m);
poyb200_ctx *poyb200_multi_ctx(poyb200_multi *m, int k);          /* the context of device k (borrowed) */
int64_t poyb200_multi_launch_count(const poyb200_multi *m);
/* pair index at which each shard of the last call began (n_devices + 1 entries, the last one = n_pairs) */
int poyb200_multi_shards(const poyb200_multi *m, int64_t *begin, int cap);

/* --- introspection for benchmarks and tests ----------------------------------------------------------- */
/* Number of kernels launched by this context so far. */
int64_t poyb200_launch_count(const poyb200_ctx *ctx);
/* DP cells the reference visits for one pair: linear (l1 >= l2 stored lengths, deltaw) and
 * affine_3 (stored lengths, any order). */
int64_t poyb200_cells_linear(int32_t l1, int32_t l2, int32_t deltaw);
int64_t poyb200_cells_affine(int32_t la, int32_t lb);
/* Device time in milliseconds of the kernels of the last poyb200_run, by phase: [0] fill, [1] traceback.
 * Measured with CUDA events on the context's stream. */
int poyb200_last_run_ms(poyb200_ctx *ctx, float ms[2]);
/* Raw cudaStream_t of the context (as void*), so a caller can bracket poyb200_run with its own events. */
void *poyb200_stream(poyb200_ctx *ctx);
/* Measures the INT32 ALU throughput of the device (dependent-free IADD3/VIMNMX mix), in Gop/s. */
int poyb200_int32_peak(poyb200_ctx *ctx, double *gops_add, double *gops_minmax, double *gops_mix);

#ifdef __cplusplus
}
#endif
#endif
