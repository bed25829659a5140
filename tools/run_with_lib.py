#!/usr/bin/env python
"""Experiment aid: run bench.py (or any script) against a VARIANT build of the C-ABI library.

    python tools/run_with_lib.py build  OUT.so  [-DNAME=VALUE ...]      # nvcc, same flags as the product build
    python tools/run_with_lib.py run    LIB.so  bench.py --workload kl_377k ...

Used for geometry sweeps of the kernels (e.g. -DRADAR_KL_BLOCK_N=112 -DRADAR_KL_SETS=4); the product always loads
radar_multimodal_radiology_b200/csrc/libradar_retrieval.so.
"""
import os
import runpy
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from radar_multimodal_radiology_b200 import _lib
    cmd, path, rest = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3:]
    if cmd == "build":
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        full = [nvcc] + _lib.NVCC_FLAGS + rest + ["-o", path, os.path.join(_lib._CSRC, "radar_retrieval.cu"), "-lcudart_static"]
        raise SystemExit(subprocess.run(full).returncode)
    if cmd == "run":
        _lib.LIB_PATH = path
        _lib.build = lambda *a, **k: path
        _lib.needs_build = lambda: False
        sys.argv = rest
        runpy.run_path(os.path.join(ROOT, rest[0]) if not os.path.isabs(rest[0]) else rest[0], run_name="__main__")
        return
    raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
