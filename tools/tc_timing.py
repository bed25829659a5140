#!/usr/bin/env python
"""Experiment aid: per-CTA start / end times of the real pass of tc_filter_kernel on the headline workload (needs a library built
with -DRADAR_TC_TIMING: python tools/run_with_lib.py build build_variants/tc_timing.so -DRADAR_TC_TIMING).
    python tools/tc_timing.py build_variants/tc_timing.so [rows] [queries]"""
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
q = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
dev = torch.device("cuda:0")
idx = RadarIndex(512, device=dev, precision="fp32")
idx.reserve(n)
for b in range(0, n, 1_250_000):
    m = min(1_250_000, n - b)
    idx.add(syn.embeddings(m, 512, syn.SEED_CORPUS_EMB + b, dev))
    idx.add_observations(syn.observation_probs(m, syn.SEED_CORPUS_PROBS + b, dev))
qe = syn.embeddings(q, 512, 4242, dev)
qp = syn.observation_probs(q, syn.SEED_QUERY_PROBS, dev)
for it in range(3):
    idx.search(qe, 10, query_probs=qp, mode="hybrid", collect_stats=(it == 2))
torch.cuda.synchronize()
print(idx.last_stats)
buf = (C.c_ulonglong * (296 * 2))()
assert _lib.lib().radar_debug_tc_timing(buf) == 0
t = np.array(buf, dtype=np.uint64).reshape(296, 2)[:148].astype(np.int64)
t0 = t[:, 0].min()
start, end = (t[:, 0] - t0) / 1e6, (t[:, 1] - t0) / 1e6
print("start ms: min %.3f max %.3f" % (start.min(), start.max()))
print("end   ms: min %.3f avg %.3f max %.3f" % (end.min(), end.mean(), end.max()))
pe = end[0::2]
order = np.argsort(pe)
print("pair end times ms (sorted):", np.round(pe[order], 2).tolist())
