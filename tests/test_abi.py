"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/poyb200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "poyb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(poyb200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from poyd_b200 import build, _lib

    so = build.build()
    L = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/poyb200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "poyd_b200/_lib.py EXPORTS out of sync with the header"
    # the tree driver's header (include/poyb200_tree.h)
    src = open(os.path.join(ROOT, "include", "poyb200_tree.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    tree_names = sorted(set(re.findall(r"\b(poyb200_tree_[a-z0-9_]+)\s*\(", src)))
    assert len(tree_names) >= 9
    for n in tree_names:
        assert hasattr(L, n), f"{n} declared in include/poyb200_tree.h but not exported"


def test_no_device_is_a_loud_error():
    """Without a CUDA device the product must fail, not fall back (this test runs on the CPU-only box)."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    from poyd_b200 import cost_matrix, sequence

    try:
        sequence.Align(cost_matrix.default_nucleotides())
    except sequence.PoyB200Error as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Align() succeeded without a GPU")


def test_cells_formula_matches_survey_check_values():
    """SURVEY.md 8d check values for cells(): measured there on the compiled reference."""
    from poyd_b200 import sequence as S

    assert S.cells_linear(501, 501, 26) == 69951
    assert S.cells_linear(501, 476, 26) == 78076
    assert S.cells_linear(501, 451, 3) == 67149
    assert S.cells_linear(301, 301, 16) == 35141
    assert S.cells_linear(301, 301, 271) == 90601
    assert S.cells_linear(1501, 1501, 76) == 361001
    assert S.cells_linear(1501, 1401, 76) == 476001
    assert S.cells_affine(501, 501) == 38941
    assert S.cells_affine(476, 501) == 37616
    assert S.cells_affine(451, 501) == 43793
    assert S.cells_affine(301, 301) == 22741
    assert S.cells_affine(1501, 1501) == 119941


def test_ocaml_stubs_compile_against_reference_headers():
    """stubs/poyb200_stubs.c (the OCaml-side binding of INTEGRATION.md) must compile against the reference's own
    seq.h / cm.h and the stand-in OCaml headers.  Needs /root/reference (this container only)."""
    import subprocess

    import pytest

    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference headers not present")
    cmd = ["gcc", "-std=gnu99", "-fgnu89-inline", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-variable",
           "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + ref, "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "stubs", "poyb200_stubs.c")]
    subprocess.check_call(cmd)


def test_config_struct_matches_the_header_and_its_defaults():
    """poyb200_config (include/poyb200.h) field by field against the ctypes mirror, and the documented defaults --
    poyb200_default_config needs no device.  A caller built against an older header passes a shorter struct_bytes."""
    from poyd_b200 import _lib

    src = open(os.path.join(ROOT, "include", "poyb200.h")).read()
    body = src[src.index("typedef struct poyb200_config {"):src.index("} poyb200_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:uint32_t|int32_t|int64_t)\s+([a-z0-9_]+)\s*;", body)
    assert fields == [n for n, _ in _lib.Config._fields_], "poyd_b200/_lib.py Config out of sync with include/poyb200.h"
    cfg = _lib.make_config(None)
    assert cfg.struct_bytes == ctypes.sizeof(_lib.Config)
    want = dict(force_generic=0, allow_fast=1, allow_noeb=1, use_ring=2, overlap_traceback=1, dir_buffers=3,
                traceback_threads_per_sm=256, traceback_block=128, chunk_pairs=1 << 16, allow_rows=1, dir6=1, pair2=1,
                pair2_min_pairs=0)
    for k, v in want.items():
        assert getattr(cfg, k) == v, (k, getattr(cfg, k), v)
    assert _lib.make_config({"pair2": 0}).pair2 == 0
    import pytest

    with pytest.raises(KeyError):
        _lib.make_config({"no_such_field": 1})
