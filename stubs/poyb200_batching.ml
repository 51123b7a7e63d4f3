(* poyb200_batching.ml -- the batching layer for src/seqCS.ml / src/allDirChar.ml (SURVEY.md 8f #1, INTEGRATION.md 3).

   NOT COMPILED HERE: this image has no OCaml toolchain.  It is the OCaml transcription of the call sequence that
   poyd_b200/tree.py executes through the same C ABI -- and that file reproduces the 52 tree costs of the
   reference's test/cost_tests (tests/test_tree.py) -- written against the reference's own types so that a
   maintainer can paste the three pieces where the comments say.  C side: stubs/poyb200_stubs.c, section (3). *)

(* ------------------------------------------------------------------------------------------------------------
   1. src/sequence.ml, inside module Align, next to c_align_affine_3 (:461-478)
   ------------------------------------------------------------------------------------------------------------ *)
external c_batch_median :
  s array -> int array -> int array -> Cost_matrix.Two_D.m -> s array ->
  (int array * int array * string array * string array * string array)
  = "poyb200_CAML_batch_median"

external c_batch_closest :
  s array -> int array -> int array -> Cost_matrix.Two_D.m -> s array -> int array
  = "poyb200_CAML_batch_closest"

external c_batch_cost_2 : s array -> int array -> int array -> Cost_matrix.Two_D.m -> int array
  = "poyb200_CAML_batch_cost_2"

(* deltaw of cost_2 for one pair, exactly :691-714 (None = no hint; Some v = the hint of DOS.distance) *)
let batch_deltaw ?deltaw s1 s2 cm =
  let ls1 = length s1 and ls2 = length s2 in
  let calc big small =
    let dif = big - small and lower = int_of_float ((float_of_int big) *. 0.10) in
    match deltaw with
    | None -> if dif < lower then lower / 2 else 2
    | Some v -> if dif < lower then lower else v in
  let gaps = max (count_gaps s1 cm) (count_gaps s2 cm) in
  gaps + (if ls1 >= ls2 then calc ls1 ls2 else calc ls2 ls1)

(* seqs: the distinct operands; pairs.(2p), pairs.(2p+1): their indices for pair p *)
let batch_median seqs pairs cm =
  let n = Array.length pairs / 2 in
  let a p = seqs.(pairs.(2 * p)) and b p = seqs.(pairs.(2 * p + 1)) in
  let med = Array.init n (fun p -> create (length (a p) + length (b p) + 2)) in
  let dw = Array.init n (fun p -> batch_deltaw (a p) (b p) cm) in
  let costs, cols, ba, bb, bm = c_batch_median seqs pairs dw cm med in
  Array.init n (fun p -> med.(p), costs.(p), cols.(p), ba.(p), bb.(p), bm.(p))

(* ------------------------------------------------------------------------------------------------------------
   2. src/seqCS.ml, inside module DOS, after `median` (:747-776): the same record from a batch result
   ------------------------------------------------------------------------------------------------------------ *)
let bitset_of cols data =
  (* BitSet.t is abstract in extlib; the stub delivers the bits in BitSet's own order (bit i: byte i / 8, position i mod 8) *)
  let set = BitSet.create cols in
  for i = 0 to cols - 1 do
    if (Char.code data.[i lsr 3]) land (1 lsl (i land 7)) <> 0 then BitSet.set set i
  done;
  set

let packed_of cols data operand = Packed (cols, bitset_of cols data, Raw operand)

let median_batch h (jobs : (do_single_sequence * do_single_sequence) array) =
  let gap = Cost_matrix.Two_D.gap h.c2 in
  let res = Array.make (Array.length jobs) None in
  (* the two early exits of DOS.median (:749-752) never reach the GPU *)
  let todo = ref [] in
  Array.iteri (fun k (a, b) ->
    if Sequence.is_empty a.sequence gap then res.(k) <- Some (create b.sequence, 0)
    else if Sequence.is_empty b.sequence gap then res.(k) <- Some (create a.sequence, 0)
    else todo := k :: !todo) jobs;
  let todo = Array.of_list (List.rev !todo) in
  (* distinct operands once: physical equality is enough, medians of one level are fresh values *)
  let tbl = Hashtbl.create 1024 and seqs = ref [] and cnt = ref 0 in
  let index s =
    try Hashtbl.find tbl (Obj.repr s) with Not_found ->
      Hashtbl.add tbl (Obj.repr s) !cnt; seqs := s :: !seqs; incr cnt; !cnt - 1 in
  let pairs = Array.make (2 * Array.length todo) 0 in
  Array.iteri (fun q k ->
    let a, b = jobs.(k) in
    pairs.(2 * q) <- index a.sequence; pairs.(2 * q + 1) <- index b.sequence) todo;
  let seqs = Array.of_list (List.rev !seqs) in
  let out = Sequence.Align.batch_median seqs pairs h.c2 in
  Array.iteri (fun q k ->
    let a, b = jobs.(k) in
    let seqm, cost, cols, ba, bb, bm = out.(q) in
    res.(k) <- Some ({ sequence = seqm;
                       aligned_children = (packed_of cols ba a.sequence, packed_of cols bb b.sequence,
                                           packed_of cols bm seqm);
                       costs = make_cost cost; position = 0 }, cost)) todo;
  Array.map (function Some x -> x | None -> assert false) res

(* ------------------------------------------------------------------------------------------------------------
   3. src/allDirChar.ml, module M: forcing the lazy medians level by level instead of one by one
   ------------------------------------------------------------------------------------------------------------
   internal_downpass (:722-786) creates one thunk per directional node (create_lazy_node, :49-58) and leaves
   the forcing to whoever asks first.  With the batching layer the thunks of one dependency level are forced
   together: collect, for every node_dir whose two operands are already values, the pair of Node.node_data;
   hand the dynamic characters of all of them to SeqCS.DOS.median_batch in one call (character by character:
   SeqCS.median maps DOS.median over the characters, :1822-1851, so the job array is vertices x characters,
   operands ordered by min_child_code as Node.cs_median does, src/node.ml:343-348); build each vertex's
   Node.node_data from its slice of the result exactly as Node.median does (:638-678: children's total_cost +
   the characters' costs); store it with AllDirNode.lazy_from_val.  Repeat until no thunk is left, then
   refresh_all_edges (:672-700) as one more level (every edge median is one job), pick_best_root unchanged
   (:787-869), and assign_single (:283-399) with every depth of the uppass as one c_batch_closest call.

   let force_levels ptree =
     let pending = ref (all_lazy_node_dirs ptree) in
     while !pending <> [] do
       let ready, later = List.partition (operands_are_values ptree) !pending in
       let jobs = Array.concat (List.map (dynamic_character_pairs ptree) ready) in
       let medians = SeqCS.DOS.median_batch (heuristic_of ptree) jobs in
       store_node_data ptree ready medians;     (* Node.median's bookkeeping, no alignment calls *)
       pending := later
     done

   The candidate-edge sweep of the Wagner build and of SPR/TBR joins (cost_fn, :1279-1317; src/ptree.ml:995-1030)
   is one c_batch_cost_2 call per clade: seqs = clade :: edge medians, pairs = (0, k) for every edge k,
   deltaw = batch_deltaw ~deltaw:(max 8 |la - lb|) (src/seqCS.ml:856-866); the costs go to the search manager in
   the original edge order (src/queues.ml:367-407), so search trajectories are unchanged. *)
