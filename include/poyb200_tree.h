/* poyb200_tree.h -- native host driver for the tree-level callers of the alignment path, C ABI.
 *
 * The reference evaluates a tree by forcing one lazy median at a time (AllDirChar.internal_downpass,
 * src/allDirChar.ml:722-786), each force ending in one algn_CAML_* call.  This driver computes the same quantities in
 * level-order batches on a device-resident sequence store (poyb200_store_*, poyb200.h), so that only costs cross the
 * host link:
 *   - the three directional medians of every interior vertex (create_lazy_interior_down/up, src/allDirChar.ml:49-97;
 *     operand order by min_child_code as Node.cs_median, src/node.ml:343-348; each = SeqCS.DOS.median, src/seqCS.ml:747-776),
 *     one batch per dependency level, hash-consed by the subtree they summarise, shared between trees;
 *   - edge medians and root selection (refresh_all_edges :672-700, create_root :99-125, general_pick_best_root +
 *     blindly_trust_downpass :787-869);
 *   - single assignment (assign_single :283-399, SeqCS.DOS.to_single src/seqCS.ml:730-745, Sequence.Align.closest
 *     src/sequence.ml:967-1033), one batch per depth;
 *   - the adjusted cost (check_cost :179-211): DOS.distance (src/seqCS.ml:819-867) along the edges -- the number
 *     Ptree.get_cost `Adjusted returns and the reference's test/cost_tests pin (test/cc*.costs);
 *   - Wagner builds with the candidate-edge sweep of AllDirChar.cost_fn (:1279-1317) as one batch per taxon
 *     (Ptree.make_wagner_tree, src/ptree.ml:948-1060);
 *   - SPR search: Ptree.single_spr_round (src/ptree.ml:1120-1170) under the first-best manager
 *     (Queues.first_best_srch_mgr, src/queues.ml:421-560): for every break of the current tree the WHOLE sweep of join costs
 *     is evaluated in one batch, then the manager's decisions are replayed in order (cc < break delta -> exact cost of the
 *     joined tree, accepted when it beats the best so far; exact costs are evaluated for a window of candidates in lockstep).
 * Vertex codes follow Tree.convert_to (src/tree.ml:724-860) -- they decide ties -- and are the caller's (the Python mirror
 * poyd_b200/tree.py builds them; poyd_b200/tree_native.py binds this header).
 */
#ifndef POYB200_TREE_H
#define POYB200_TREE_H
#include "poyb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct poyb200_tree poyb200_tree;

/* A topology: n_nodes vertices; ids[k] is the code of vertex k and nbr[3k .. 3k+2] its neighbours in the order the
 * reference stores them (Leaf: parent, -1, -1; Interior: parent, child 1, child 2 -- src/tree.ml); `handle` a vertex code. */
typedef struct poyb200_topology {
    int32_t n_nodes;
    const int32_t *ids;
    const int32_t *nbr;
    int32_t handle;
} poyb200_topology;

typedef struct poyb200_tree_cost {
    int64_t adjusted;    /* Ptree.get_cost `Adjusted */
    int64_t unadjusted;  /* root cost of the chosen root median */
    int32_t root_a, root_b;
} poyb200_tree_cost;

/* An evaluator over one context (whose cost matrix must be loaded) for n_loci sequence characters. */
int poyb200_tree_create(poyb200_ctx *ctx, int32_t n_loci, poyb200_tree **out);
void poyb200_tree_destroy(poyb200_tree *t);
const char *poyb200_tree_last_error(const poyb200_tree *t);
/* The sequence of taxon `code` for locus `locus` (leading gap included).  All loci of a taxon must be set before use. */
int poyb200_tree_set_leaf(poyb200_tree *t, int32_t code, int32_t locus, const uint8_t *seq, int32_t len);
/* Forgets every cached median (keeps the leaves). */
void poyb200_tree_reset(poyb200_tree *t);

/* Downpass + uppass of n trees over the same leaves in lockstep (every batch holds the work of all trees; whatever the
 * trees share is computed once).  keep = 1 keeps the median cache of earlier calls. */
int poyb200_tree_evaluate(poyb200_tree *t, const poyb200_topology *topos, int32_t n, int32_t keep, poyb200_tree_cost *out);
/* The single assignment of vertex `vertex`, locus `locus` of tree `which` of the LAST evaluate call: copies it to `out`
 * (capacity cap) and returns its length in *len. */
int poyb200_tree_single(poyb200_tree *t, int32_t which, int32_t vertex, int32_t locus, uint8_t *out, int32_t cap, int32_t *len);

/* Wagner build: taxa in `order` (n codes), each joined to the edge of smallest cost_fn (first minimum in pre-order).
 * The result topology is written to ids / nbr (capacity 2 n vertices) and *handle; steps (may be NULL) receives
 * {taxon, edge a, edge b, delta} x (n - 2). */
int poyb200_tree_wagner(poyb200_tree *t, const int32_t *order, int32_t n, int32_t *ids, int32_t *nbr, int32_t *n_nodes, int32_t *handle,
                        int64_t *steps);

typedef struct poyb200_spr_result {
    int64_t start_cost, final_cost;
    int32_t rounds;            /* accepted rearrangements */
    int64_t breaks, joins_swept, exact_evaluated;
} poyb200_spr_result;
/* SPR hill climb from `start` (see the header comment): the final topology is written to ids / nbr (capacity
 * start->n_nodes) and *handle.  max_rounds bounds the accepted moves (0 = until no break improves); window = candidates
 * evaluated exactly per lockstep call. */
int poyb200_tree_spr(poyb200_tree *t, const poyb200_topology *start, int32_t max_rounds, int32_t window, int32_t *ids, int32_t *nbr,
                     int32_t *handle, poyb200_spr_result *res);

/* One round of the same search restricted to the breaks k with k % nshards == shard (several GPUs: one shard of the
 * tree's neighbourhood each).  *found = 1 when one of the shard's candidates, taken in order, has an exact
 * cost below best_cost: then *key orders it among all shards' finds (the smallest key is the candidate the unsharded
 * search would have taken), *cost is its cost and ids / nbr / *handle its topology.  The caller reduces over the shards. */
int poyb200_tree_spr_round(poyb200_tree *t, const poyb200_topology *cur, int64_t best_cost, int32_t shard, int32_t nshards, int32_t window,
                           int32_t *found, int64_t *key, int64_t *cost, int32_t *ids, int32_t *nbr, int32_t *handle,
                           poyb200_spr_result *res);

/* batches / pairs / DP cells sent to the GPU so far, medians computed, sequences in the store */
void poyb200_tree_stats(const poyb200_tree *t, int64_t *calls, int64_t *pairs, int64_t *cells, int64_t *medians, int64_t *sequences);

#ifdef __cplusplus
}
#endif
#endif
