"""Shared comparison code of the parity tests (tests only)."""
import numpy as np


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
        for p in range(n):
            L = int(lo[p])
            g = buf[p, stride - L:]
            o = ora[name][p, :L]
            if not np.array_equal(g, o):
                pos = int(np.nonzero(g != o)[0][0])
                raise AssertionError(f"{label}: {name} differs for pair {p} at position {pos} of {L}: gpu {g[pos]} ref {o[pos]}")
