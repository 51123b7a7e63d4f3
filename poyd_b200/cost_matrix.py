"""Host-side mirror of the reference's ``Cost_matrix.Two_D`` (src/cost_matrix.ml).

The alignment kernels only *read* the tables built here (``struct cm``,
src/cm.h:32-46): ``cost``/``median``/``worst`` indexed ``(a << lcm) + b``
(src/cm.c:501-504, 545-552) and ``prepend_cost``/``tail_cost`` indexed by code.
This module restates the OCaml builders so that a caller who used
``Cost_matrix.Two_D.of_list`` / ``of_transformations_and_gaps`` /
``set_affine`` / ``default`` / ``default_aminoacids`` gets the same tables:

* ``cm_set_val`` geometry (src/cm.c:319-361): with combinations the gap is
  ``1 << (a_sz-1)``, ``a_sz`` becomes ``2^a_sz - 1`` and ``lcm = a_sz``;
  without, ``gap = a_sz`` and ``lcm = ceil_log_2(a_sz + 1)`` (src/cm.c:36-43).
* ``fill_best_cost_and_median_for_all_combinations`` (cost_matrix.ml:323-354),
  its ``_bitwise`` twin (:270-321), ``fill_medians`` (:356-395) and
  ``fill_default_prepend_tail`` (:397-403).

``fill_medians`` breaks ties among equally good medians with ``Random.int``
(cost_matrix.ml:389); that stream (PoyRandom) is not reproducible outside the
OCaml program, so ties are broken here by a documented rule -- the LOWEST code
-- and the resulting table is treated as input data fed identically to the
oracle and to the GPU.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence as _Seq

import numpy as np

LINNEAR = 0  # spelling follows cost_matrix.ml:30
AFFINE = 1
NO_ALIGNMENT = 2

# "max_int" of cost_matrix.ml:39 (Int32.max_int lsr 1)
_MAX_INT = (2**31 - 1) >> 1


def ceil_log_2(v: int) -> int:
    """src/cm.c:36-43 (note: returns one more than the bit length)."""
    i = 0
    while v != 0:
        i += 1
        v >>= 1
    return i + 1


def _bits(v: int, a_sz: int) -> List[int]:
    """split_integer_in_list_of_bits (cost_matrix.ml:205-214): set bits of v below a_sz."""
    return [1 << b for b in range(a_sz) if (v >> b) & 1]


@dataclasses.dataclass
class CostMatrix:
    """Flat image of ``struct cm`` (src/cm.h:32-46)."""

    a_sz_in: int  # alphabet size as passed to cm_CAML_create (incl. gap)
    a_sz: int
    lcm: int
    gap: int
    cost_model_type: int
    combinations: int
    gap_open: int
    is_metric: int
    all_elements: int
    cost: np.ndarray  # int32 [dim, dim]
    median: np.ndarray  # uint8 [dim, dim]
    worst: np.ndarray  # int32 [dim, dim]
    prepend_cost: np.ndarray  # int32 [dim]
    tail_cost: np.ndarray  # int32 [dim]

    @property
    def dim(self) -> int:
        return 1 << self.lcm

    # --- accessors named after the OCaml externals (cost_matrix.ml:44-75) ---
    def alphabet_size(self) -> int:
        return self.a_sz

    def combine(self) -> int:
        return self.combinations

    def affine(self):
        """Cost_matrix.Two_D.affine (cost_matrix.ml:123-127)."""
        if self.cost_model_type == 0:
            return ("Linnear",)
        if self.cost_model_type == 1:
            return ("Affine", self.gap_open)
        return ("No_Alignment",)

    def clone(self) -> "CostMatrix":
        return dataclasses.replace(
            self,
            cost=self.cost.copy(),
            median=self.median.copy(),
            worst=self.worst.copy(),
            prepend_cost=self.prepend_cost.copy(),
            tail_cost=self.tail_cost.copy(),
        )

    # --- builders -------------------------------------------------------------
    def _cleanup(self, item: int) -> int:
        """cleanup (cost_matrix.ml:255-265)."""
        if self.combinations == 0 or self.cost_model_type != AFFINE:
            return item
        if item != self.gap and (item & self.gap) != 0:
            return self.gap
        return item

    def _fill_all_combinations(self, a_sz: int) -> None:
        """fill_best_cost_and_median_for_all_combinations (cost_matrix.ml:323-354) with
        test_combinations (:219-243)."""
        n = (1 << a_sz) - 1
        gap, go = self.gap, self.gap_open
        c = self.cost
        single = [1 << i for i in range(a_sz)]
        # cost a v and cost v b are read while entries are being overwritten: rows are visited in
        # increasing (i, j) order and reads hit (a, v)/(v, b) with a, b, v single bits, whose entries are
        # never written (the [_],[_] branch leaves the cost alone), so a snapshot is equivalent.
        for i in range(1, n + 1):
            li = _bits(i, a_sz)
            for j in range(1, n + 1):
                lj = _bits(j, a_sz)
                best, cst, worst = 0, _MAX_INT, 0
                for a in li:
                    for b in lj:
                        for v in single:
                            goa = go if (self.cost_model_type == AFFINE and v == gap and (a & gap) and (b & gap)) else 0
                            tc = int(c[a, v]) + int(c[v, b]) + goa
                            if tc < cst:
                                cst, best = tc, v
                            elif tc == cst:
                                best |= v
                        cab = int(c[a, b])
                        if cab > worst:
                            worst = cab
                if len(li) == 1 and len(lj) == 1:
                    self.median[i, j] = i | j
                else:
                    c[i, j] = cst
                    self.median[i, j] = self._cleanup(best)
                self.worst[i, j] = worst

    def _fill_all_combinations_bitwise(self, a_sz: int) -> None:
        """fill_best_cost_and_median_for_all_combinations_bitwise (cost_matrix.ml:270-321)."""
        n = (1 << a_sz) - 1
        c = self.cost

        def find_best(l1, l2):
            best, med, worst = _MAX_INT, 0, 0
            # process l1 l2 (process l2 l1 init): the l2 x l1 sweep comes first
            for (xs, ys) in ((l2, l1), (l1, l2)):
                for x in xs:
                    for y in ys:
                        mb1 = int(c[x, y])
                        if mb1 < best:
                            best, med = mb1, x | y
                        elif mb1 == best:
                            med = x | y | med
                        worst = max(mb1, worst)
            return best, med, worst

        for i in range(1, n + 1):
            li = _bits(i, a_sz)
            for j in range(1, n + 1):
                lj = _bits(j, a_sz)
                if len(li) == 1 and len(lj) == 1:
                    self.median[i, j] = i | j
                    self.worst[i, j] = c[i, j]
                elif i & j:
                    _, _, w = find_best(li, lj)
                    self.median[i, j] = i & j
                    c[i, j] = 0
                    self.worst[i, j] = w
                else:
                    b, m, w = find_best(li, lj)
                    self.median[i, j] = m
                    c[i, j] = b
                    self.worst[i, j] = w

    def _fill_medians(self, a_sz: int) -> None:
        """fill_medians (cost_matrix.ml:356-395).  Tie rule: lowest code (see module docstring)."""
        c = self.cost
        ae = self.all_elements
        for i in range(1, a_sz + 1):
            for j in range(1, a_sz + 1):
                if i == ae:
                    res = [j]
                elif j == ae:
                    res = [i]
                else:
                    best, res = _MAX_INT, []
                    for k in range(1, a_sz + 1):
                        if k == ae:
                            continue
                        cc = int(c[k, j]) + int(c[i, k])
                        if cc < best:
                            best, res = cc, [k]
                        elif cc == best:
                            res.append(k)
                self.median[i, j] = self._cleanup(min(res)) if res else 0

    def _fill_default_prepend_tail(self) -> None:
        """fill_default_prepend_tail (cost_matrix.ml:397-403)."""
        for i in range(1, self.a_sz + 1):
            self.tail_cost[i] = self.cost[i, self.gap]
            self.prepend_cost[i] = self.cost[self.gap, i]

    def set_affine(self, model) -> None:
        """Cost_matrix.Two_D.set_affine (cost_matrix.ml:405-416): model is ("Linnear",),
        ("Affine", go) or ("No_Alignment",).  Mutates, like the reference (callers clone first,
        src/data.ml:3283-3289)."""
        if model[0] == "No_Alignment":
            self.cost_model_type, self.gap_open = NO_ALIGNMENT, 0
        elif model[0] == "Linnear":
            self.cost_model_type, self.gap_open = LINNEAR, 0
        else:
            self.cost_model_type, self.gap_open = AFFINE, int(model[1])
        if self.combinations == 1:
            self._fill_all_combinations_bitwise(self.lcm)
        else:
            self._fill_medians(self.a_sz)
        self._fill_default_prepend_tail()


def create(a_sz: int, combine: bool, aff: int, go: int, all_elements: int) -> CostMatrix:
    """cm_CAML_create -> cm_set_val (src/cm.c:319-361)."""
    if a_sz > 255:
        raise ValueError("alphabet larger than 255 needs --enable-large-alphabets (src/cm.c:323)")
    if combine:
        gap, asz, lcm, comb = 1 << (a_sz - 1), (1 << a_sz) - 1, a_sz, 1
    else:
        gap, asz, lcm, comb = a_sz, a_sz, ceil_log_2(a_sz + 1), 0
    dim = 1 << lcm
    return CostMatrix(
        a_sz_in=a_sz, a_sz=asz, lcm=lcm, gap=gap, cost_model_type=aff, combinations=comb, gap_open=go,
        is_metric=0, all_elements=all_elements,
        cost=np.zeros((dim, dim), np.int32), median=np.zeros((dim, dim), np.uint8),
        worst=np.zeros((dim, dim), np.int32), prepend_cost=np.zeros(dim, np.int32),
        tail_cost=np.zeros(dim, np.int32),
    )


def _input_is_metric(arr: np.ndarray) -> bool:
    """input_is_metric (cost_matrix.ml:513-518)."""
    w = arr.shape[0]
    if (arr < 0).any() or not (arr == arr.T).all() or (np.diag(arr) != 0).any():
        return False
    for k in range(w):
        if (arr > arr[:, k : k + 1] + arr[k : k + 1, :]).any():
            return False
    return True


def fill_cost_matrix(rows: _Seq[_Seq[int]], all_elements: int, use_comb: bool = True) -> CostMatrix:
    """fill_cost_matrix (cost_matrix.ml:520-538); created Linnear with gap opening 0 (:78-80)."""
    arr = np.asarray(rows, dtype=np.int64)
    a_sz = arr.shape[0]
    assert arr.shape == (a_sz, a_sz)
    m = create(a_sz, use_comb, LINNEAR, 0, all_elements)
    if use_comb:
        # store_input_list_in_cost_matrix_all_combinations (:156-173)
        for e1 in range(a_sz):
            for e2 in range(a_sz):
                m.cost[1 << e1, 1 << e2] = arr[e1, e2]
        if _input_is_metric(arr):
            m.is_metric = 1
            m._fill_all_combinations(a_sz)
        else:
            m._fill_all_combinations_bitwise(a_sz)
    else:
        # store_input_list_in_cost_matrix_no_comb (:175-200): the all_elements code costs 0 against anything
        for e1 in range(1, a_sz + 1):
            for e2 in range(1, a_sz + 1):
                h = 0 if (e1 == all_elements or e2 == all_elements) else int(arr[e1 - 1, e2 - 1])
                m.cost[e1, e2] = h
                m.worst[e1, e2] = h
        m._fill_medians(a_sz)
    m._fill_default_prepend_tail()
    return m


def of_list(rows, all_elements: int, use_comb: bool = True) -> CostMatrix:
    """Cost_matrix.Two_D.of_list / of_list_nocomb (cost_matrix.ml:588-604)."""
    return fill_cost_matrix(rows, all_elements, use_comb)


def of_transformations_and_gaps(use_combinations: bool, alph_size: int, trans: int, gaps: int,
                                all_elements: int) -> CostMatrix:
    """cost_matrix.ml:645-657: 0 on the diagonal, `gaps` in the last row/column, `trans` elsewhere."""
    rows = [[0 if x == p else (gaps if (x == alph_size - 1 or p == alph_size - 1) else trans)
             for x in range(alph_size)] for p in range(alph_size)]
    return of_list(rows, all_elements, use_combinations)


def default_nucleotides() -> CostMatrix:
    """Cost_matrix.Two_D.default (cost_matrix.ml:606-615): substitution 1, indel 2, all_elements 31."""
    return of_transformations_and_gaps(True, 5, 1, 2, 31)


def default_aminoacids() -> CostMatrix:
    """Cost_matrix.Two_D.default_aminoacids (cost_matrix.ml:619-643): 22x22, 1/2, all_elements 21."""
    return of_transformations_and_gaps(False, 22, 1, 2, 21)


def nucleotides(trans: int = 1, gaps: int = 2, gap_opening: Optional[int] = None) -> CostMatrix:
    """What ``transform (tcm:(trans,gaps), gap_opening:go)`` yields for DNA (src/data.ml:3215-3222,
    3283-3289): of_transformations_and_gaps, then clone + set_affine when a gap opening is given."""
    m = of_transformations_and_gaps(True, 5, trans, gaps, 31)
    if gap_opening is not None:
        m = m.clone()
        m.set_affine(("Affine", gap_opening))
    return m


# ---- Cost_matrix.Three_D (src/cost_matrix.ml:703-891) ------------------------------------------------------------
@dataclasses.dataclass
class CostMatrix3D:
    """Flat image of ``struct cm_3d`` (src/cm.h:222-241): cost / median indexed ((a << lcm) + b) << lcm) + c."""

    a_sz_in: int
    a_sz: int
    lcm: int
    gap: int
    cost_model_type: int
    combinations: int
    gap_open: int
    all_elements: int
    cost: np.ndarray  # int32 [dim, dim, dim]
    median: np.ndarray  # uint8 [dim, dim, dim]


def of_two_dim(m: CostMatrix) -> CostMatrix3D:
    """Cost_matrix.Three_D.of_two_dim (cost_matrix.ml:853-862) = cm_CAML_clone_to_3d (src/cm.c:1435) followed by
    of_two_dim_comb (:803-851) or of_two_dim_no_comb (:785-801)."""
    dim = 1 << m.lcm
    nm = CostMatrix3D(a_sz_in=m.lcm if m.combinations else m.a_sz, a_sz=m.a_sz, lcm=m.lcm, gap=m.gap,
                      cost_model_type=m.cost_model_type, combinations=m.combinations, gap_open=m.gap_open,
                      all_elements=m.all_elements, cost=np.zeros((dim, dim, dim), np.int32),
                      median=np.zeros((dim, dim, dim), np.uint8))
    c2 = m.cost.astype(np.int64)
    if m.combinations:
        alph, gap, lcm = m.a_sz, m.gap, m.lcm
        rng = np.arange(1, alph + 1)
        best = np.full((alph, alph, alph), _MAX_INT, np.int64)
        med = np.zeros((alph, alph, alph), np.int64)
        for l in range(lcm):
            inter = 1 << l
            ci = c2[inter, 1:alph + 1]
            cost = ci[:, None, None] + ci[None, :, None] + ci[None, None, :]
            if not m.is_metric and inter == gap:
                sh = ((rng & inter) != 0).astype(np.int64)
                ok = (sh[:, None, None] + sh[None, :, None] + sh[None, None, :]) >= 2
                cost = np.where(ok, cost, _MAX_INT)
            lt, eq = cost < best, cost == best
            med = np.where(lt, inter, np.where(eq, med | inter, med))
            best = np.where(lt, cost, best)
        # "We pick only one median among all options": the lowest set bit (pick_bit, :808-817)
        low = med & (-med)
        nm.cost[1:alph + 1, 1:alph + 1, 1:alph + 1] = best
        nm.median[1:alph + 1, 1:alph + 1, 1:alph + 1] = low
    else:
        alph = m.a_sz - 1
        # of_two_dim_no_comb starts from a zeroed matrix and only lowers costs (`cost < old_cost`), so with
        # non-negative 2-D costs nothing is ever written: reproduce exactly that (a reference quirk)
        for l in range(1, alph + 1):
            cl = c2[l, 1:alph + 1]
            cost = cl[:, None, None] + cl[None, :, None] + cl[None, None, :]
            old = nm.cost[1:alph + 1, 1:alph + 1, 1:alph + 1].astype(np.int64)
            lt = cost < old
            nm.cost[1:alph + 1, 1:alph + 1, 1:alph + 1] = np.where(lt, cost, old)
            nm.median[1:alph + 1, 1:alph + 1, 1:alph + 1] = np.where(lt, l, nm.median[1:alph + 1, 1:alph + 1, 1:alph + 1])
    return nm
