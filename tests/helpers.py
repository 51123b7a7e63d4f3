"""Shared comparison code of the parity tests (tests only)."""
import numpy as np


def right_rows_equal(right, left, lens) -> bool:
    """True when row p of `right` (right aligned) holds the lens[p] leading bytes of row p of `left`, for every p."""
    n = len(lens)
    if n == 0:
        return True
    w = int(np.max(lens))
    if w == 0:
        return True
    for lo in range(0, n, 16384):  # blocks: bounded temporaries on 125 k-pair slices
        L = np.asarray(lens[lo:lo + 16384], dtype=np.int64)
        g = right[lo:lo + 16384, right.shape[1] - w:]
        cols = np.arange(w)[None, :]
        src = cols - (w - L[:, None])
        valid = src >= 0
        o = np.zeros_like(g)
        rows = np.nonzero(valid)[0]
        o[valid] = left[lo:lo + 16384][rows, src[valid]]
        if not np.array_equal(np.where(valid, g, 0), o):
            return False
    return True


def assert_aligned_equal(gpu, ora, want_median=True, want_wg=True, want_al=True, label=""):
    """gpu: poyd_b200.sequence.Aligned (rows right aligned); ora: dict from oracle batch (rows left aligned)."""
    n = len(ora["cost"])
    bad = np.nonzero(gpu.cost != ora["cost"])[0]
    assert len(bad) == 0, f"{label}: {len(bad)} cost mismatches, first pair {bad[:5]}: gpu {gpu.cost[bad[:5]]} ref {ora['cost'][bad[:5]]}"
    checks = []
    if want_median:
        checks.append(("median", 0, gpu.median))
    if want_wg:
        checks.append(("medianwg", 1, gpu.medianwg))
    if want_al:
        checks.append(("ra", 2, gpu.aligned_a))
        checks.append(("rb", 3, gpu.aligned_b))
    for name, k, buf in checks:
        lg, lo = gpu.lens[:, k], ora["lens"][:, k]
        badl = np.nonzero(lg != lo)[0]
        assert len(badl) == 0, f"{label}: {name} length mismatch at pairs {badl[:5]}: gpu {lg[badl[:5]]} ref {lo[badl[:5]]}"
        stride = buf.shape[1]
        if right_rows_equal(buf, ora[name], lo):
            continue  # vectorised pass; the loop below only runs to name the first difference
        for p in range(n):
            L = int(lo[p])
            g = buf[p, stride - L:]
            o = ora[name][p, :L]
            if not np.array_equal(g, o):
                pos = int(np.nonzero(g != o)[0][0])
                raise AssertionError(f"{label}: {name} differs for pair {p} at position {pos} of {L}: gpu {g[pos]} ref {o[pos]}")
