/* algn_b200.c -- link-time replacement of src/algn.c in the reference tree.
 *
 * Put this file and poyb200_stubs.c into the reference's src/, list `algn_b200.o poyb200_stubs.o` instead of `algn.o` in
 * src/libpoycside.clib, link with -lpoyb200.  Not one line of OCaml changes: the externals of src/sequence.ml keep their
 * symbol names.  What happens:
 *
 *   - the reference's own algn.c is compiled as part of this translation unit with the alignment externals RENAMED
 *     (algn_CAML_simple_2 -> algn_CAML_simple_2_cpu, ...), so every other function of the file -- algn_CAML_union,
 *     algn_CAML_myers, the *_limit family, algn_CAML_create_backtrack, algn_CAML_print_bcktrck, and the C helpers other
 *     stubs call -- is still there, unchanged;
 *   - poyb200_stubs.c defines the renamed externals' ORIGINAL names on top of libpoyb200 (the GPU path).
 *
 * The renamed CPU externals are not called by anything; they stay linkable for side-by-side checks
 * (tests/test_stubs.py does exactly that comparison in this repository, against the unmodified algn.o).
 */
#define algn_CAML_simple_2 algn_CAML_simple_2_cpu
#define algn_CAML_backtrack_2d algn_CAML_backtrack_2d_cpu
#define algn_CAML_backtrack_2d_bc algn_CAML_backtrack_2d_bc_cpu
#define algn_CAML_align_2d algn_CAML_align_2d_cpu
#define algn_CAML_align_2d_bc algn_CAML_align_2d_bc_cpu
#define algn_CAML_cost_affine_3 algn_CAML_cost_affine_3_cpu
#define algn_CAML_align_affine_3 algn_CAML_align_affine_3_cpu
#define algn_CAML_align_affine_3_bc algn_CAML_align_affine_3_bc_cpu
#define algn_CAML_median_2_no_gaps algn_CAML_median_2_no_gaps_cpu
#define algn_CAML_median_2_with_gaps algn_CAML_median_2_with_gaps_cpu
#define algn_CAML_ancestor_2 algn_CAML_ancestor_2_cpu
#define algn_CAML_worst_2 algn_CAML_worst_2_cpu
#define algn_CAML_verify_2 algn_CAML_verify_2_cpu
#define algn_CAML_simple_3 algn_CAML_simple_3_cpu
#define algn_CAML_simple_3_bc algn_CAML_simple_3_bc_cpu
#define algn_CAML_backtrack_3d algn_CAML_backtrack_3d_cpu
#define algn_CAML_backtrack_3d_bc algn_CAML_backtrack_3d_bc_cpu
#define algn_CAML_align_3d algn_CAML_align_3d_cpu
#define algn_CAML_align_3d_bc algn_CAML_align_3d_bc_cpu
#define algn_CAML_median_3 algn_CAML_median_3_cpu
#include "algn.c"
