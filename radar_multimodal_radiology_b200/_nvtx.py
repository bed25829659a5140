"""NVTX ranges around the host-side stages of the path (index build, search, exchange, merge) -- SURVEY.md section 5
"tracing / profiling": the reference has none.  Free when no profiler is attached; visible in nsys / ncu timelines as
``radar:<stage>``.  ``RADAR_NVTX=0`` switches them off."""
from __future__ import annotations

import contextlib
import functools
import os

_ENABLED = os.environ.get("RADAR_NVTX", "1") != "0"
try:  # torch.cuda.nvtx needs a CUDA build of torch, not a GPU
    from torch.cuda import nvtx as _nvtx
    _nvtx.range_push  # noqa: B018
except Exception:  # pragma: no cover - CPU-only torch
    _nvtx = None


@contextlib.contextmanager
def range_(name: str):
    """``with range_("search"):`` -> NVTX range ``radar:search`` on the calling thread."""
    on = _ENABLED and _nvtx is not None
    if on:
        try:
            _nvtx.range_push("radar:" + name)
        except Exception:  # libnvToolsExt missing: tracing is an aid, never a reason to fail a search
            on = False
    try:
        yield
    finally:
        if on:
            _nvtx.range_pop()


def annotate(name: str):
    """Decorator form of :func:`range_`."""
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*a, **k):
            with range_(name):
                return fn(*a, **k)
        return wrapped
    return deco
