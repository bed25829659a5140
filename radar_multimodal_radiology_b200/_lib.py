"""ctypes binding of ``csrc/libradar_retrieval.so`` (C ABI in ``include/radar_retrieval.h``).

There is no CPU fallback: if the shared library is missing, or a compute entry point reports an
error, a ``RuntimeError`` is raised.  PyTorch is only used for device memory and streams.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import shutil
import subprocess
from typing import Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG_DIR, "csrc")
LIB_PATH = os.path.join(_CSRC, "libradar_retrieval.so")
# RADAR_DEBUG flavour: additionally exports radar_debug_filter_keys and honours the RADAR_TC_* bring-up switches.
# Only tests and tools load it; the product path (index.py and everything above it) uses LIB_PATH.
DBG_LIB_PATH = os.path.join(_CSRC, "libradar_retrieval_dbg.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), "include", "radar_retrieval.h")

MODE_DPR, MODE_KL, MODE_HYBRID = 0, 1, 2
PREC_BF16, PREC_FP32 = 0, 1
ALGO_AUTO, ALGO_SIMT_EXACT, ALGO_TC_FILTER, ALGO_KL_STREAM = 0, 1, 2, 3
NUM_OBS, OBS_PAD, KLPACK, MAX_K = 14, 16, 32, 128
ABI_VERSION = 4

MODE_BY_NAME = {"dpr": MODE_DPR, "kl": MODE_KL, "hybrid": MODE_HYBRID}
PREC_BY_NAME = {"bf16": PREC_BF16, "fp32": PREC_FP32}
ALGO_BY_NAME = {"auto": ALGO_AUTO, "simt": ALGO_SIMT_EXACT, "tc": ALGO_TC_FILTER, "kl_stream": ALGO_KL_STREAM}
# filter arithmetic of the KL-only tensor-core paths (enum radar_kl_variant)
KL_VARIANT_BY_NAME = {"auto": 0, "bf16x3": 1, "f16x1": 2, "f16x2": 3}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
    "-Xcompiler", "-fPIC",
]


class CorpusStruct(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("d", C.c_int32), ("reserved0", C.c_int32),
        ("emb_f32", C.c_void_p), ("emb_bf16", C.c_void_p), ("logq16", C.c_void_p), ("klpack", C.c_void_p),
        ("emb_max_norm", C.c_float), ("logq_max_abs", C.c_float), ("idx_offset", C.c_int64),
        ("logq_col_max", C.c_float * 16), ("kl16", C.c_void_p),
    ]


class QueriesStruct(C.Structure):
    _fields_ = [("q", C.c_int64), ("emb_f32", C.c_void_p), ("p16", C.c_void_p), ("entropy", C.c_void_p),
                ("after_scores", C.c_void_p), ("after_idx", C.c_void_p)]


class SearchParams(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("precision", C.c_int32), ("algo", C.c_int32), ("k", C.c_int32),
        ("alpha", C.c_float), ("overfetch", C.c_int32), ("num_sms", C.c_int32), ("kl_variant", C.c_int32),
    ]


class SearchStats(C.Structure):
    _fields_ = [
        ("algo_used", C.c_int32), ("kernel_launches", C.c_int32), ("uncertified", C.c_int64),
        ("parts", C.c_int32), ("kprime", C.c_int32), ("filter_sm_mhz", C.c_float), ("reserved", C.c_int32),
    ]


EXPORTS = [
    "radar_last_error", "radar_abi_version", "radar_device_info", "radar_set_device",
    "radar_profile_enable", "radar_profile_kernel_ms", "radar_pack_embeddings",
    "radar_kl_prepare_corpus", "radar_kl_prepare_queries", "radar_search_workspace_bytes", "radar_search",
    "radar_merge_topk", "radar_merge_packed", "radar_rerank_overlap", "radar_gather_bits", "radar_project_normalize",
    "radar_get_device", "radar_project_workspace_bytes", "radar_project_normalize_tc",
]
DEBUG_EXPORTS = ["radar_debug_filter_keys"]


def sources():
    return [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC)) if f.endswith((".cu", ".cuh"))]


def _stale(path: str) -> bool:
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(s) > t for s in sources() + [HEADER_PATH])


def needs_build() -> bool:
    return _stale(LIB_PATH) or _stale(DBG_LIB_PATH)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension for sm_100a, in tree (nvcc cross-compiles without a GPU): the release library and
    its RADAR_DEBUG flavour, side by side."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    jobs = []
    for path, extra in ((LIB_PATH, []), (DBG_LIB_PATH, ["-DRADAR_DEBUG"])):
        if not force and not _stale(path):
            continue
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [
            "-o", path, os.path.join(_CSRC, "radar_retrieval.cu"), "-lcudart_static"]
        jobs.append((path, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for path, proc in jobs:
        out, err = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {os.path.basename(path)}:\n" + out + err)
        if verbose and path == LIB_PATH:
            print(err)
    return LIB_PATH


_lib = None
_dbg_lib = None


def _declare(l, debug: bool):
    l.radar_last_error.restype = C.c_char_p
    l.radar_abi_version.restype = C.c_int
    l.radar_search_workspace_bytes.restype = C.c_size_t
    l.radar_search_workspace_bytes.argtypes = [C.POINTER(CorpusStruct), C.c_int64, C.POINTER(SearchParams)]
    l.radar_search.argtypes = [C.POINTER(CorpusStruct), C.POINTER(QueriesStruct), C.POINTER(SearchParams),
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(SearchStats),
                               C.c_void_p]
    l.radar_pack_embeddings.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    l.radar_kl_prepare_corpus.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
    l.radar_kl_prepare_queries.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    l.radar_merge_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    l.radar_merge_packed.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    l.radar_rerank_overlap.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    l.radar_gather_bits.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int64,
                                    C.c_void_p, C.c_void_p]
    l.radar_project_normalize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p]
    l.radar_project_workspace_bytes.restype = C.c_size_t
    l.radar_project_workspace_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
    l.radar_project_normalize_tc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    l.radar_set_device.argtypes = [C.c_int]
    l.radar_get_device.argtypes = [C.POINTER(C.c_int)]
    l.radar_profile_enable.argtypes = [C.c_int]
    l.radar_profile_kernel_ms.argtypes = [C.POINTER(C.c_float)]
    l.radar_device_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    if debug:
        l.radar_debug_filter_keys.argtypes = [C.POINTER(CorpusStruct), C.POINTER(QueriesStruct),
                                              C.POINTER(SearchParams), C.c_void_p, C.c_void_p, C.c_size_t,
                                              C.c_void_p]
    for name in EXPORTS + (DEBUG_EXPORTS if debug else []):
        getattr(l, name)  # raises AttributeError if a declared symbol is not exported
    return l


def lib():
    """Load the extension (no compute is run).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback for the retrieval kernels.")
        _lib = _declare(C.CDLL(LIB_PATH), debug=False)
    return _lib


def debug_lib():
    """The RADAR_DEBUG flavour (tests / bring-up tools only)."""
    global _dbg_lib
    if _dbg_lib is None:
        if not os.path.exists(DBG_LIB_PATH):
            raise RuntimeError(f"{DBG_LIB_PATH} is missing: build it with _lib.build()")
        _dbg_lib = _declare(C.CDLL(DBG_LIB_PATH), debug=True)
    return _dbg_lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().radar_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None passes NULL)."""
    return None if t is None else t.data_ptr()


def current_stream_ptr(device) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def device_index(device) -> int:
    import torch
    dev = torch.device(device)
    return dev.index if dev.index is not None else torch.cuda.current_device()


@contextlib.contextmanager
def device_guard(device, library=None):
    """Make ``device`` current for the duration of a library call -- in torch's CUDA runtime and in the library's
    own (statically linked) one -- and restore what was current before, so that a search on cuda:1 issued from a
    thread whose current device is cuda:0 leaves that thread on cuda:0."""
    import torch
    l = library or lib()
    idx = device_index(device)
    prev = C.c_int(-1)
    check(l.radar_get_device(C.byref(prev)), "radar_get_device")
    with torch.cuda.device(idx):
        if prev.value != idx:
            check(l.radar_set_device(idx), "radar_set_device")
        try:
            yield
        finally:
            if prev.value != idx and prev.value >= 0:
                l.radar_set_device(prev.value)


def device_info():
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib().radar_device_info(C.byref(a), C.byref(b), C.byref(c)), "radar_device_info")
    return a.value, b.value, c.value
