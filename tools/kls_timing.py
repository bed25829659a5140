#!/usr/bin/env python
"""Experiment aid: per-CTA start / prologue / end times and rare-path visits of kl_stream_kernel (needs a library built with
-DRADAR_KLS_TIMING: python tools/run_with_lib.py build build_variants/kls_timing.so -DRADAR_KLS_TIMING).
    python tools/kls_timing.py build_variants/kls_timing.so [rows] [queries] [k]"""
import ctypes as C, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radar_multimodal_radiology_b200 import _lib
path = os.path.abspath(sys.argv[1])
_lib.LIB_PATH = path
_lib.build = lambda *a, **k: path
_lib.needs_build = lambda: False
from radar_multimodal_radiology_b200 import synthetic as syn
from radar_multimodal_radiology_b200.index import RadarIndex

n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
q = int(sys.argv[3]) if len(sys.argv) > 3 else 32
k = int(sys.argv[4]) if len(sys.argv) > 4 else 32
dev = torch.device("cuda:0")
idx = RadarIndex(512, device=dev, precision="fp32")
idx.add_observations(syn.observation_probs(n, syn.SEED_CORPUS_PROBS, dev))
qp = syn.observation_probs(q, syn.SEED_QUERY_PROBS, dev)
for it in range(4):
    idx.search(None, k, query_probs=qp, mode="kl", collect_stats=(it == 3))
torch.cuda.synchronize()
print(idx.last_stats)
buf = (C.c_ulonglong * (296 * 4))()
lib = _lib.lib()
assert lib.radar_debug_kls_timing(buf) == 0
t = np.array(buf, dtype=np.uint64).reshape(296, 4)[:148].astype(np.int64)
t0 = t[:, 0].min()
start, pro, end, rare = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0, t[:, 3]
print("start  ns: min %d max %d" % (start.min(), start.max()))
print("prologue ns (after - start): min %d avg %d max %d" % ((pro - start).min(), (pro - start).mean(), (pro - start).max()))
print("end    ns: min %d avg %d max %d" % (end.min(), end.mean(), end.max()))
print("body   ns (end - after prologue): min %d avg %d max %d" % ((end - pro).min(), (end - pro).mean(), (end - pro).max()))
print("rare visits per CTA: min %d avg %d max %d" % (rare.min(), rare.mean(), rare.max()))
print("corr(body, rare) = %.3f" % np.corrcoef((end - pro).astype(float), rare.astype(float))[0, 1])
order = np.argsort(end)
print("slowest CTAs (id, end, rare):", [(int(i), int(end[i]), int(rare[i])) for i in order[-8:]])
print("fastest CTAs (id, end, rare):", [(int(i), int(end[i]), int(rare[i])) for i in order[:8]])
