#!/usr/bin/env python
"""KL search of q queries over the first n rows of a LARGER resident corpus (RadarIndex.view_rows) vs over an index of its own."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radar_multimodal_radiology_b200 import synthetic as syn
from radar_multimodal_radiology_b200.index import RadarIndex

dev = torch.device("cuda:0")
n_big, n, q, k = int(sys.argv[1]), int(sys.argv[2]), 65536, 10
big = RadarIndex(512, device=dev, precision="fp32")
big.add_observations(syn.observation_probs(n_big, syn.SEED_CORPUS_PROBS, dev))
if len(sys.argv) > 3 and sys.argv[3] == "emb":  # the bench's resident corpus also carries the embedding tables
    big.reserve(n_big, embeddings=True, observations=False)
    for b in range(0, n_big, 1000000):
        big.add(syn.embeddings(min(1000000, n_big - b), 512, syn.SEED_CORPUS_EMB + b, dev))
own = RadarIndex(512, device=dev, precision="fp32")
own.add_observations(syn.observation_probs(n_big, syn.SEED_CORPUS_PROBS, dev)[:n])
qp = syn.observation_probs(q, syn.SEED_QUERY_PROBS, dev)
for name, idx in (("own", own), ("view", big.view_rows(0, n)), ("view+offset", big.view_rows(0, n, idx_offset=n))):
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        s, i = idx.search(None, k, query_probs=qp, mode="kl", collect_stats=(it == 4))
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            print(f"{name}: n={n} of {n_big}: {e0.elapsed_time(e1):.3f} ms  stats={idx.last_stats if it == 4 else ''}", flush=True)
