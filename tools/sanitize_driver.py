#!/usr/bin/env python
"""Small-shape run of every kernel path, for compute-sanitizer (tools/sanitize.sh).  Each search is checked against the
exact scan, so a sanitizer-clean run is also a correct one."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radar_multimodal_radiology_b200 import synthetic as syn  # noqa: E402
from radar_multimodal_radiology_b200.index import RadarIndex, merge_packed  # noqa: E402


def check(idx, name, qe, qp, mask, k, mode, **kw):
    s, i, p = idx.search(qe, k, query_probs=qp, mask=mask, mode=mode, return_packed=True, collect_stats=True, **kw)
    es, ei = idx.search(qe, k, query_probs=qp, mask=mask, mode=mode, algo="simt", precision="fp32")
    torch.cuda.synchronize()
    ok = bool(torch.equal(i, ei) and torch.equal(s, es))
    print(f"{name}: algo {idx.last_stats.algo_used}, uncertified {idx.last_stats.uncertified}, equals exact scan: {ok}", flush=True)
    assert ok, name
    return p


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = torch.device("cuda:0")
    d = 64
    if which in ("all", "filter"):
        n, q = 6000, 300
        idx = RadarIndex(d, device=dev, precision="fp32")
        idx.add(syn.embeddings(n, d, syn.SEED_CORPUS_EMB, dev))
        idx.add_observations(syn.observation_probs(n, syn.SEED_CORPUS_PROBS, dev))
        qe = syn.embeddings(q, d, syn.SEED_QUERY_EMB, dev)
        qp = syn.observation_probs(q, syn.SEED_QUERY_PROBS, dev)
        mask = syn.observation_masks(q, 1, dev)
        for mode in ("dpr", "hybrid"):
            p = check(idx, f"tc_filter {mode}", qe, qp if mode != "dpr" else None, mask if mode != "dpr" else None, 10, mode, algo="tc")
        merge_packed(torch.stack([p, p]), 10, "hybrid")
        for v in ("bf16x3", "f16x1", "f16x2"):
            check(idx, f"kl_filter {v} (cold start)", None, qp, mask, 10, "kl", algo="tc", kl_variant=v)
    if which in ("all", "prepass"):
        n, q = 70000, 2100
        idx = RadarIndex(d, device=dev, precision="fp32")
        idx.add_observations(syn.observation_probs(n, syn.SEED_CORPUS_PROBS, dev))
        qp = syn.observation_probs(q, syn.SEED_QUERY_PROBS, dev)
        check(idx, "kl_filter bf16x3 (prepass)", None, qp, None, 10, "kl", algo="tc")
        check(idx, "kl_filter f16x2 (prepass)", None, qp, None, 10, "kl", algo="tc", kl_variant="f16x2")
    if which in ("all", "stream"):
        n, q = 70000, 9
        idx = RadarIndex(d, device=dev, precision="fp32")
        idx.add_observations(syn.observation_probs(n, syn.SEED_CORPUS_PROBS, dev))
        qp = syn.observation_probs(q, syn.SEED_QUERY_PROBS, dev)
        for v in ("bf16x3", "f16x2"):
            check(idx, f"kl_stream {v}", None, qp, None, 10, "kl", kl_variant=v)
    print("sanitize driver: all paths agree with the exact scan", flush=True)


if __name__ == "__main__":
    main()
