"""ctypes binding of ``libxpore_b200.so`` (C ABI in ``include/xpore_b200.h``) and its in-tree build.

There is no CPU fallback: if the library is missing or has no CUDA device to run on, the
compute entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("XPB200_LIB") or os.path.join(HERE, "libxpore_b200.so")  # override: experiments only
SOURCES = ["xp_api.cu", "xp_ingest.cpp"]
HEADERS = ["xp_math.cuh", "xp_core.cuh", "xp_kernels.cuh", "xp_fit.cuh", "xp_fit_tm.cuh", "xp_tables.inc", os.path.join("..", "..", "include", "xpore_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"]


class XpBatch(C.Structure):
    _fields_ = [("n_positions", C.c_int32), ("n_groups", C.c_int32), ("n_conditions", C.c_int32),
                ("max_reads", C.c_int32), ("n_reads", C.c_int64), ("y", C.c_void_p), ("read_off", C.c_void_p),
                ("grp_len", C.c_void_p), ("kmer_mean", C.c_void_p), ("kmer_std", C.c_void_p),
                ("grp_cond", C.c_void_p), ("y_format", C.c_int32), ("reserved", C.c_int32),
                ("y_scale", C.c_double)]


class XpMethod(C.Structure):
    _fields_ = [("max_iters", C.c_int32), ("prefilter", C.c_int32), ("stopping_criteria", C.c_double),
                ("prefilter_threshold", C.c_double), ("dirichlet_conc", C.c_double)]


class XpResult(C.Structure):
    _fields_ = [("status", C.c_void_p), ("n_iter", C.c_void_p), ("pval", C.c_void_p), ("fit", C.c_void_p),
                ("stats", C.c_void_p), ("coef", C.c_void_p)]


class XpLaunchInfo(C.Structure):
    _fields_ = [("sm_count", C.c_int32), ("ctas_per_sm", C.c_int32), ("warps_per_cta", C.c_int32),
                ("smem_per_cta", C.c_int32), ("slab_reads", C.c_int32), ("regs_per_thread", C.c_int32),
                ("positions_per_warp", C.c_int32), ("tmem_columns", C.c_int32)]


EXPORTS = ["xpb200_fit_width", "xpb200_stats_width", "xpb200_workspace_bytes", "xpb200_fit_device",
           "xpb200_fit_host", "xpb200_fit_host_rows", "xpb200_fit_host_rows_async", "xpb200_pack_rows", "xpb200_release", "xpb200_prefilter_device", "xpb200_stats_device",
           "xpb200_posterior_device", "xpb200_fit_launch_info", "xpb200_deep_split", "xpb200_host_chunk_plan", "xpb200_launch_count", "xpb200_measure_fp64_peak", "xpb200_eval_special",
           "xpb200_set_stage_timing", "xpb200_get_stage_ms", "xpb200_pack_records",
           "xpb200_debug_fill_smem", "xpb200_last_error", "xpb200_version",
           "xpb200_ingest_open", "xpb200_ingest_genes", "xpb200_ingest_copy", "xpb200_ingest_close",
           "xpb200_ingest_encoding", "xpb200_ingest_copy_u16", "xpb200_host_alloc", "xpb200_host_free", "xpb200_set_device",
           "xpb200_ingest_last_error"]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libxpore_b200.so is not built — run `python -c 'import __graft_entry__ as g; "
                               "g.build()'` (there is no CPU fallback for the diffmod kernels)")
        L = C.CDLL(LIB_PATH)
        L.xpb200_fit_width.restype = C.c_int32
        L.xpb200_fit_width.argtypes = [C.c_int32]
        L.xpb200_stats_width.restype = C.c_int32
        L.xpb200_stats_width.argtypes = [C.c_int32, C.c_int32]
        L.xpb200_workspace_bytes.restype = C.c_size_t
        L.xpb200_workspace_bytes.argtypes = [C.c_int32]
        L.xpb200_fit_device.restype = C.c_int32
        L.xpb200_fit_device.argtypes = [C.POINTER(XpBatch), C.POINTER(XpMethod), C.POINTER(XpResult), C.c_void_p,
                                        C.c_size_t, C.c_void_p, C.c_void_p]
        L.xpb200_fit_host.restype = C.c_int32
        L.xpb200_fit_host.argtypes = [C.c_int32, C.POINTER(XpBatch), C.POINTER(XpMethod), C.POINTER(XpResult),
                                      C.POINTER(C.c_uint32)]
        L.xpb200_fit_host_rows.restype = C.c_int32
        L.xpb200_fit_host_rows.argtypes = [C.c_int32, C.POINTER(XpBatch), C.POINTER(XpMethod), C.c_void_p, C.c_int64,
                                           C.POINTER(C.c_int64), C.POINTER(C.c_uint32)]
        L.xpb200_fit_host_rows_async.restype = C.c_int32
        L.xpb200_fit_host_rows_async.argtypes = [C.c_int32, C.POINTER(XpBatch), C.POINTER(XpMethod), C.c_void_p, C.c_int64,
                                                 C.c_int64, C.c_void_p, C.c_void_p]
        L.xpb200_pack_rows.restype = C.c_int32
        L.xpb200_pack_rows.argtypes = [C.POINTER(XpResult), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        if hasattr(L, "xpb200_debug_fill_smem"):  # absent only in experiment builds (XPB200_LIB)
            L.xpb200_debug_fill_smem.restype = C.c_int32
            L.xpb200_debug_fill_smem.argtypes = [C.c_uint64, C.c_void_p]
        L.xpb200_ingest_open.restype = C.c_void_p
        L.xpb200_ingest_open.argtypes = [C.c_int32, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_char_p)]
        L.xpb200_ingest_genes.restype = C.c_int32
        L.xpb200_ingest_genes.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                          C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.xpb200_ingest_copy.restype = C.c_int32
        L.xpb200_ingest_copy.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.xpb200_ingest_encoding.restype = C.c_int32
        L.xpb200_ingest_encoding.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.xpb200_ingest_copy_u16.restype = C.c_int32
        L.xpb200_ingest_copy_u16.argtypes = [C.c_void_p, C.c_void_p]
        L.xpb200_host_alloc.restype = C.c_void_p
        L.xpb200_host_alloc.argtypes = [C.c_size_t]
        L.xpb200_host_free.restype = None
        L.xpb200_host_free.argtypes = [C.c_void_p]
        L.xpb200_set_device.restype = C.c_int32
        L.xpb200_set_device.argtypes = [C.c_int32]
        L.xpb200_ingest_close.restype = None
        L.xpb200_ingest_close.argtypes = [C.c_void_p]
        L.xpb200_ingest_last_error.restype = C.c_char_p
        L.xpb200_release.restype = None
        L.xpb200_prefilter_device.restype = C.c_int32
        L.xpb200_prefilter_device.argtypes = [C.POINTER(XpBatch), C.POINTER(XpMethod), C.c_void_p, C.c_void_p,
                                              C.c_void_p]
        L.xpb200_stats_device.restype = C.c_int32
        L.xpb200_stats_device.argtypes = [C.POINTER(XpBatch), C.POINTER(XpMethod), C.POINTER(XpResult), C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        L.xpb200_posterior_device.restype = C.c_int32
        L.xpb200_posterior_device.argtypes = [C.POINTER(XpBatch), C.POINTER(XpResult), C.c_void_p, C.c_void_p,
                                              C.c_void_p]
        L.xpb200_fit_launch_info.restype = C.c_int32
        L.xpb200_fit_launch_info.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(XpLaunchInfo)]
        L.xpb200_deep_split.restype = C.c_int32
        L.xpb200_deep_split.argtypes = [C.c_int32]
        L.xpb200_launch_count.restype = C.c_uint64
        L.xpb200_measure_fp64_peak.restype = C.c_int32
        L.xpb200_measure_fp64_peak.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_double)]
        L.xpb200_eval_special.restype = C.c_int32
        L.xpb200_eval_special.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.xpb200_set_stage_timing.restype = C.c_int32
        L.xpb200_set_stage_timing.argtypes = [C.c_int32]
        L.xpb200_get_stage_ms.restype = C.c_int32
        L.xpb200_get_stage_ms.argtypes = [C.POINTER(C.c_float)]
        L.xpb200_pack_records.restype = C.c_int32
        L.xpb200_pack_records.argtypes = [C.POINTER(XpResult), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_int64, C.c_void_p]
        L.xpb200_last_error.restype = C.c_char_p
        L.xpb200_version.restype = C.c_char_p
        _lib = L
    return _lib


_warm = {"thread": None, "seconds": None, "device": None}


def warm_up_async(device: int = 0):
    """Create the CUDA context of ``device`` on a background thread (ctypes releases the GIL), so that interpreter
    work - imports, reading ``data.index``, parsing the first chunk - overlaps the 0.3 - 1 s it takes.  No-op when the
    library is missing or no device is usable; the compute entry points report that themselves."""
    import threading
    import time

    if _warm["thread"] is not None or not os.path.exists(LIB_PATH):
        return

    def run():
        t0 = time.perf_counter()
        try:
            L = lib()
            if L.xpb200_set_device(int(device)) == 0:
                L.xpb200_host_free(L.xpb200_host_alloc(4096))
        except Exception:  # noqa: BLE001  (the real call sites raise)
            pass
        _warm["seconds"] = time.perf_counter() - t0

    _warm["device"] = int(device)
    _warm["thread"] = threading.Thread(target=run, name="xpb200-cuda-init", daemon=True)
    _warm["thread"].start()


def warm_up_wait():
    """Seconds the background context creation took (None if it was never started)."""
    if _warm["thread"] is not None:
        _warm["thread"].join()
    return _warm["seconds"]


def check(rc: int):
    if rc != 0:
        raise RuntimeError("xpore_b200: %s (code %d)" % (lib().xpb200_last_error().decode(), rc))
