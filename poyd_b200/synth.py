"""Seeded synthetic workloads for the tests and bench.py (SURVEY.md 8d "Concrete synthetic inputs").

DNA codes are the reference's bit sets A=1 C=2 G=4 T=8 gap=16 (src/alphabet.ml:67-71); protein codes are
sequential 1..20, X=21, gap=22 (src/alphabet.ml:77-99).  Every sequence carries its leading gap.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .sequence import SeqPool

DNA_GAP = 16
PROTEIN_GAP = 22


def _mutate_rows(rng, parent: np.ndarray, bases: np.ndarray, subst: float, indel: float) -> Tuple[np.ndarray, np.ndarray]:
    """Vectorised point mutation of the rows of `parent` [n, L]: substitutions with probability `subst`,
    single-element deletions / insertions with probability indel/2 each.  Returns (flat children, lengths)."""
    n, L = parent.shape
    r = rng.random((n, L), dtype=np.float32)
    counts = np.ones((n, L), dtype=np.int8)
    counts[r < indel / 2] = 0
    counts[(r >= indel / 2) & (r < indel)] = 2
    child = parent.copy()
    sub = rng.random((n, L), dtype=np.float32) < subst
    child[sub] = bases[rng.integers(0, len(bases), size=int(sub.sum()))]
    lens = counts.sum(axis=1, dtype=np.int64)
    flat_counts = counts.reshape(-1).astype(np.int64)
    src = np.repeat(np.arange(n * L, dtype=np.int64), flat_counts)
    out = child.reshape(-1)[src]
    # the second copy of a doubled position is the inserted element: draw it afresh
    dup = np.zeros(len(src), dtype=bool)
    dup[1:] = src[1:] == src[:-1]
    out[dup] = bases[rng.integers(0, len(bases), size=int(dup.sum()))]
    return out, lens


def pair_batch(n_pairs: int, length: int = 500, seed: int = 2, alphabet: str = "dna", subst: float = 0.10,
               indel: float = 0.02, min_len: int = 0, ambiguity: float = 0.0, gap_ambiguity: float = 0.0,
               stride: int = 0) -> Tuple[SeqPool, np.ndarray]:
    """`n_pairs` (parent, child) pairs: parent uniform over the alphabet with `length` elements, child = parent
    with `subst` substitutions and `indel` single-element indels, clamped to [min_len, length] elements.
    ambiguity: fraction of positions replaced by a random IUPAC set (DNA only); gap_ambiguity: fraction of
    positions OR-ed with the gap bit (codes the affine medians really contain, src/algn.c:2016, 2028).
    Returns (pool, pairs) with pair p = (2p, 2p+1)."""
    rng = np.random.default_rng(seed)
    if alphabet == "dna":
        bases, gap = np.array([1, 2, 4, 8], np.uint8), DNA_GAP
    else:
        bases, gap = np.arange(1, 21, dtype=np.uint8), PROTEIN_GAP
    stride = stride or ((length + 1 + 15) // 16 * 16)
    mat = np.zeros((2 * n_pairs, stride), np.uint8)
    lens = np.zeros(2 * n_pairs, np.int32)
    chunk = 50_000
    for lo in range(0, n_pairs, chunk):
        hi = min(n_pairs, lo + chunk)
        m = hi - lo
        parent = bases[rng.integers(0, len(bases), size=(m, length))]
        if alphabet == "dna" and ambiguity > 0:
            amb = rng.random((m, length)) < ambiguity
            parent[amb] = rng.integers(1, 16, size=int(amb.sum()), dtype=np.uint8)
        flat, clen = _mutate_rows(rng, parent, bases, subst, indel)
        starts = np.concatenate([[0], np.cumsum(clen)[:-1]])
        clen_c = np.clip(clen, min_len, length)  # clamp: truncate long children, (rare) short ones are kept as they are
        clen_c = np.minimum(clen_c, clen)
        rows_a = slice(2 * lo, 2 * hi, 2)
        mat[rows_a, 0] = gap
        mat[rows_a, 1:length + 1] = parent
        lens[rows_a] = length + 1
        # scatter the children row by row through a flat index
        idx_row = np.repeat(np.arange(m, dtype=np.int64), clen_c)
        within = np.arange(len(idx_row), dtype=np.int64) - np.repeat(np.concatenate([[0], np.cumsum(clen_c)[:-1]]), clen_c)
        srcpos = np.repeat(starts, clen_c) + within
        child_rows = 2 * lo + 1 + 2 * idx_row
        mat[child_rows, 1 + within] = flat[srcpos]
        mat[2 * lo + 1:2 * hi:2, 0] = gap
        lens[2 * lo + 1:2 * hi:2] = clen_c + 1
    if alphabet == "dna" and gap_ambiguity > 0:
        g = rng.random(mat.shape) < gap_ambiguity
        cols = np.arange(stride)[None, :]
        g &= (cols >= 1) & (cols < lens[:, None])
        mat[g] |= np.uint8(DNA_GAP)
    pool = SeqPool.from_matrix(mat, lens)
    pairs = np.arange(2 * n_pairs, dtype=np.int32).reshape(-1, 2)
    return pool, pairs


def triple_batch(n_triples: int, length: int = 300, seed: int = 4, subst: float = 0.10, indel: float = 0.02):
    """configs[3] (SURVEY.md 8d cfg 4): `n_triples` DNA triples, a parent of `length` bases and two children at `subst`
    substitutions / `indel` single-base indels from it, children truncated to `length`.  Returns (pool, triples) with
    triple t = (3t, 3t+1, 3t+2) = (parent, child, child)."""
    rng = np.random.default_rng(seed)
    bases = np.array([1, 2, 4, 8], np.uint8)
    seqs = []
    parent = bases[rng.integers(0, 4, size=(n_triples, length))]
    kids = []
    for _ in range(2):
        flat, clen = _mutate_rows(rng, parent, bases, subst, indel)
        starts = np.concatenate([[0], np.cumsum(clen)[:-1]])
        kids.append((flat, starts, np.minimum(clen, length)))
    g = np.array([DNA_GAP], np.uint8)
    for t in range(n_triples):
        seqs.append(np.concatenate([g, parent[t]]))
        for flat, starts, clen in kids:
            seqs.append(np.concatenate([g, flat[starts[t]:starts[t] + clen[t]]]))
    pool = SeqPool(seqs)
    return pool, np.arange(3 * n_triples, dtype=np.int32).reshape(-1, 3)


def ragged_batch(n_pairs: int, max_len: int = 300, seed: int = 7, alphabet: str = "dna", related: float = 0.6,
                 gap_ambiguity: float = 0.0, ambiguity: float = 0.05, min_len: int = 0):
    """Pairs of very different shapes (lengths 0..max_len, related or unrelated operands, either order),
    for parity tests of the edge cases: empty sequences, l1 >> l2 (full-matrix case), narrow and wide bands."""
    rng = np.random.default_rng(seed)
    if alphabet == "dna":
        bases, gap = np.array([1, 2, 4, 8], np.uint8), DNA_GAP
    else:
        bases, gap = np.arange(1, 22, dtype=np.uint8), PROTEIN_GAP
    seqs = []

    def rnd(n):
        s = bases[rng.integers(0, len(bases), size=n)]
        if alphabet == "dna":
            if ambiguity > 0 and n:
                m = rng.random(n) < ambiguity
                s[m] = rng.integers(1, 16, size=int(m.sum()), dtype=np.uint8)
            if gap_ambiguity > 0 and n:
                s[rng.random(n) < gap_ambiguity] |= np.uint8(DNA_GAP)
        return np.concatenate([[gap], s]).astype(np.uint8)

    for _ in range(n_pairs):
        a = rnd(int(rng.integers(min_len, max_len + 1)))
        if rng.random() < related and len(a) > 1:
            flat, ln = _mutate_rows(rng, a[None, 1:], bases, 0.1, float(rng.choice([0.0, 0.03, 0.2])))
            b = np.concatenate([[gap], flat[:ln[0]]]).astype(np.uint8)
        else:
            b = rnd(int(rng.integers(min_len, max_len + 1)))
        if rng.random() < 0.5:
            a, b = b, a
        seqs += [a, b]
    pool = SeqPool(seqs)
    pairs = np.arange(2 * n_pairs, dtype=np.int32).reshape(-1, 2)
    return pool, pairs


def taxa_on_random_tree(n_taxa: int, length: int = 1500, seed: int = 5, subst: float = 0.03, indel: float = 0.005,
                        alphabet: str = "dna"):
    """Leaf sequences of a random binary tree (configs[4] of BASELINE.json: "500-taxon x 1.5 kb unaligned DNA"): a
    uniform root sequence evolves down a random-join topology, every branch applying `subst` substitutions and `indel`
    single-base indels.  Returns {taxon code 1..n: [sequence with its leading gap]} (one locus per taxon)."""
    rng = np.random.default_rng(seed)
    bases, gap = (np.array([1, 2, 4, 8], np.uint8), DNA_GAP) if alphabet == "dna" else (np.arange(1, 21, dtype=np.uint8), PROTEIN_GAP)
    # random topology by successive splits: a list of current tips, each a sequence; pick one, replace by two children
    tips = [bases[rng.integers(0, len(bases), size=length)]]
    while len(tips) < n_taxa:
        k = int(rng.integers(0, len(tips)))
        parent = tips.pop(k)
        for _ in range(2):
            flat, lens = _mutate_rows(rng, parent[None, :], bases, subst, indel)
            tips.append(flat[:int(lens[0])])
    perm = rng.permutation(n_taxa)
    return {i + 1: [np.concatenate([[gap], tips[int(p)]]).astype(np.uint8)] for i, p in enumerate(perm)}
