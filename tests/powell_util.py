"""Shared by the Powell tests: the compiled reference aligner (oracle/_ref/libpoyref.so, camlrt_powell in
oracle/caml_runtime.c = powell_3D_align of src/ukkCommon.c:110-145 on fresh blocks) and seeded triples."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "..", "oracle", "_ref", "libpoyref.so")
u8 = C.POINTER(C.c_uint8)
COSTS = [(1, 3, 2), (1, 0, 1), (2, 1, 1), (1, 2, 1), (3, 5, 2)]


def reference():
    """The compiled reference, or None when oracle/_ref is absent or predates the Powell recipe."""
    from oracle import oracle

    oracle.build(ref=True)
    if not os.path.exists(REF_SO):
        return None
    lib = C.CDLL(REF_SO)
    return lib if hasattr(lib, "camlrt_powell") else None


def ref_powell(lib, a, b, c, mm, go, ge):
    cap = len(a) + len(b) + len(c)
    rows = [np.zeros(cap, np.uint8) for _ in range(3)]
    n = C.c_int(0)
    cost = lib.camlrt_powell(a.ctypes.data_as(u8), len(a), b.ctypes.data_as(u8), len(b), c.ctypes.data_as(u8), len(c), mm, go, ge,
                             *[r.ctypes.data_as(u8) for r in rows], C.byref(n))
    return cost, [r[: n.value] for r in rows]


def dna(rng, n, amb=0.0):
    s = rng.choice(np.array([1, 2, 4, 8], np.uint8), size=n)
    if amb:
        hit = rng.random(n) < amb
        s[hit] |= rng.choice(np.array([1, 2, 4, 8], np.uint8), size=int(hit.sum()))
    return np.concatenate([[16], s]).astype(np.uint8)


def mutate(rng, a, p):
    out = [16]
    for x in a[1:]:
        r = rng.random()
        if r < p / 3:
            continue
        if r < 2 * p / 3:
            out.append(int(rng.choice([1, 2, 4, 8])))
        if r < p:
            out.append(int(rng.choice([1, 2, 4, 8])))
            continue
        out.append(int(x))
    return np.array(out, np.uint8)


def triples(seed, count, max_len, rates=(0.0, 0.05, 0.15, 0.4)):
    """`count` (a, b, c) with b, c mutated copies of a (or unrelated), lengths 1 .. max_len, 1 % two-base codes."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        n = int(rng.integers(1, max_len + 1))
        p = float(rng.choice(rates))
        a = dna(rng, n, amb=0.01)
        b = mutate(rng, a, p) if rng.random() < 0.9 else dna(rng, int(rng.integers(1, max_len + 1)))
        c = mutate(rng, a, p)
        if len(b) < 2 or len(c) < 2:
            continue
        out.append((a, b, c))
    return out
