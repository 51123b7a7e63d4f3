"""ctypes binding of include/poyb200.h.  There is no fallback: if libpoyb200.so is missing or no CUDA device is
usable, importing the binding or creating a context raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("POYB200_SO") or os.path.join(HERE, "libpoyb200.so")  # override: experiments only

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


class CM(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("a_sz", "lcm", "gap", "cost_model_type", "combinations", "gap_open",
                                          "is_metric", "all_elements")] + [
        ("cost", i32p), ("median", u8p), ("worst", i32p), ("prepend_cost", i32p), ("tail_cost", i32p)]


class Batch(C.Structure):
    _fields_ = [("pool", C.c_void_p), ("pool_bytes", C.c_size_t), ("seq_off", C.c_void_p), ("seq_len", C.c_void_p),
                ("n_seqs", C.c_int32), ("pairs", C.c_void_p), ("n_pairs", C.c_int32), ("deltaw", C.c_void_p),
                ("swaped", C.c_void_p), ("want", C.c_uint32), ("cost", C.c_void_p), ("median", C.c_void_p),
                ("medianwg", C.c_void_p), ("aligned_a", C.c_void_p), ("aligned_b", C.c_void_p),
                ("out_stride", C.c_int64), ("out_len", C.c_void_p), ("bits_a", C.c_void_p), ("bits_b", C.c_void_p),
                ("bits_wg", C.c_void_p), ("bits_stride", C.c_int64)]


class Config(C.Structure):
    """poyb200_config (include/poyb200.h): every tunable of a context."""
    _fields_ = [("struct_bytes", C.c_uint32)] + [(n, C.c_int32) for n in (
        "force_generic", "allow_fast", "allow_noeb", "overlap_traceback", "dir_buffers", "traceback_threads_per_sm",
        "traceback_block", "traceback_priority", "chunk_pairs", "host_threads", "timing", "trace")] + [
        ("dir_budget_bytes", C.c_int64), ("use_ring", C.c_int32), ("allow_rows", C.c_int32), ("small_ring_pairs", C.c_int32), ("dir6", C.c_int32),
        ("pair2", C.c_int32), ("pair2_min_pairs", C.c_int32)]


def make_config(overrides=None) -> "Config":
    """Defaults of the library, then the `key=value,...` pairs of the POYB200_CONFIG environment variable (a convenience of
    this Python binding for tools and experiments -- the C library itself reads no environment), then `overrides`."""
    cfg = Config()
    lib().poyb200_default_config(C.byref(cfg))
    names = {n for n, _ in Config._fields_} - {"struct_bytes"}
    items = {}
    for part in os.environ.get("POYB200_CONFIG", "").split(","):
        if "=" in part:
            k, v = part.split("=", 1)
            items[k.strip()] = int(v)
    items.update(overrides or {})
    for k, v in items.items():
        if k not in names:
            raise KeyError(f"poyb200_config has no field {k!r}")
        setattr(cfg, k, int(v))
    return cfg


class CM3(C.Structure):
    _fields_ = [("lcm", C.c_int32), ("gap", C.c_int32), ("cost", i32p), ("median", u8p)]


class Batch3(C.Structure):
    _fields_ = [("pool", C.c_void_p), ("pool_bytes", C.c_size_t), ("seq_off", C.c_void_p), ("seq_len", C.c_void_p),
                ("n_seqs", C.c_int32), ("triples", C.c_void_p), ("n_triples", C.c_int32), ("want", C.c_uint32),
                ("cost", C.c_void_p), ("aligned_1", C.c_void_p), ("aligned_2", C.c_void_p), ("aligned_3", C.c_void_p),
                ("median", C.c_void_p), ("out_stride", C.c_int64), ("out_len", C.c_void_p), ("status", C.c_void_p)]


EXPORTS = [
    "poyb200_create", "poyb200_create_ex", "poyb200_default_config", "poyb200_destroy", "poyb200_last_error", "poyb200_version", "poyb200_set_cm",
    "poyb200_host_alloc", "poyb200_host_free", "poyb200_batch_cost_2", "poyb200_batch_align_2",
    "poyb200_batch_cost_affine_3", "poyb200_batch_align_affine_3", "poyb200_batch_median_2", "poyb200_stage",
    "poyb200_run", "poyb200_sync", "poyb200_fetch", "poyb200_launch_count", "poyb200_cells_linear",
    "poyb200_cells_affine", "poyb200_last_run_ms", "poyb200_stream", "poyb200_int32_peak", "poyb200_set_cm_3d",
    "poyb200_batch_align_3", "poyb200_batch_powell_3", "poyb200_cells_3d", "poyb200_batch_worst_2", "poyb200_batch_median_3",
    "poyb200_multi_create", "poyb200_multi_destroy", "poyb200_multi_last_error", "poyb200_multi_set_cm", "poyb200_multi_batch",
    "poyb200_multi_devices", "poyb200_multi_ctx", "poyb200_multi_launch_count", "poyb200_multi_shards",
    "poyb200_store_create", "poyb200_store_destroy", "poyb200_store_size", "poyb200_store_bytes", "poyb200_store_add",
    "poyb200_store_info", "poyb200_store_get", "poyb200_store_median", "poyb200_store_distance", "poyb200_store_closest",
    "poyb200_store_stats",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise RuntimeError(
            f"{SO} is missing: build it with `python -m poyd_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(SO)
    L.poyb200_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.poyb200_create_ex.argtypes = [C.c_int, C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.poyb200_default_config.argtypes = [C.POINTER(Config)]
    L.poyb200_default_config.restype = None
    L.poyb200_destroy.argtypes = [C.c_void_p]
    L.poyb200_destroy.restype = None
    L.poyb200_last_error.argtypes = [C.c_void_p]
    L.poyb200_last_error.restype = C.c_char_p
    L.poyb200_version.restype = C.c_char_p
    L.poyb200_set_cm.argtypes = [C.c_void_p, C.POINTER(CM)]
    L.poyb200_host_alloc.argtypes = [C.c_size_t]
    L.poyb200_host_alloc.restype = C.c_void_p
    L.poyb200_host_free.argtypes = [C.c_void_p]
    L.poyb200_host_free.restype = None
    for f in ("poyb200_batch_cost_2", "poyb200_batch_align_2", "poyb200_batch_cost_affine_3",
              "poyb200_batch_align_affine_3"):
        getattr(L, f).argtypes = [C.c_void_p, C.POINTER(Batch)]
    L.poyb200_batch_median_2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_int64, C.c_void_p]
    L.poyb200_batch_worst_2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]
    L.poyb200_batch_median_3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_int64, C.c_void_p]
    L.poyb200_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.poyb200_multi_destroy.argtypes = [C.c_void_p]
    L.poyb200_multi_destroy.restype = None
    L.poyb200_multi_last_error.argtypes = [C.c_void_p]
    L.poyb200_multi_last_error.restype = C.c_char_p
    L.poyb200_multi_set_cm.argtypes = [C.c_void_p, C.POINTER(CM)]
    L.poyb200_multi_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(Batch)]
    L.poyb200_multi_devices.argtypes = [C.c_void_p]
    L.poyb200_multi_ctx.argtypes = [C.c_void_p, C.c_int]
    L.poyb200_multi_ctx.restype = C.c_void_p
    L.poyb200_multi_launch_count.argtypes = [C.c_void_p]
    L.poyb200_multi_launch_count.restype = C.c_int64
    L.poyb200_multi_shards.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
    L.poyb200_store_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.poyb200_store_destroy.argtypes = [C.c_void_p]
    L.poyb200_store_destroy.restype = None
    L.poyb200_store_size.argtypes = [C.c_void_p]
    L.poyb200_store_bytes.argtypes = [C.c_void_p]
    L.poyb200_store_bytes.restype = C.c_int64
    L.poyb200_store_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
    L.poyb200_store_info.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.poyb200_store_get.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    L.poyb200_store_median.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.poyb200_store_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.poyb200_store_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.poyb200_store_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.poyb200_store_stats.restype = None
    L.poyb200_stage.argtypes = [C.c_void_p, C.c_int, C.POINTER(Batch)]
    for f in ("poyb200_run", "poyb200_sync", "poyb200_fetch"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.poyb200_launch_count.argtypes = [C.c_void_p]
    L.poyb200_launch_count.restype = C.c_int64
    L.poyb200_cells_linear.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.poyb200_cells_linear.restype = C.c_int64
    L.poyb200_cells_affine.argtypes = [C.c_int32, C.c_int32]
    L.poyb200_cells_affine.restype = C.c_int64
    L.poyb200_last_run_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.poyb200_stream.argtypes = [C.c_void_p]
    L.poyb200_stream.restype = C.c_void_p
    L.poyb200_int32_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.poyb200_set_cm_3d.argtypes = [C.c_void_p, C.POINTER(CM3)]
    L.poyb200_batch_align_3.argtypes = [C.c_void_p, C.POINTER(Batch3)]
    L.poyb200_cells_3d.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.poyb200_cells_3d.restype = C.c_int64
    _lib = L
    return L
