"""Seeded synthetic inputs of the shapes BASELINE.json names.

Everything is drawn with a ``torch.Generator`` on the requested device in fp32; pass the same seed to get
the same bits again (per device type).  Tests generate on the CPU and upload, so the oracle and the
kernels see identical inputs; the 10M-row benchmark corpus is generated on the GPU (a host draw of 5e9
normals plus a 20 GB upload would dominate the run) and the CPU baseline reads a slice back.
"""
from __future__ import annotations

import math

import torch

PREVALENCE = [.05, .15, .25, .03, .12, .06, .07, .20, .03, .18, .01, .02, .30, .35]  # CheXpert-14 order
SEED_CORPUS_PROBS, SEED_QUERY_PROBS = 1234, 4321
SEED_CORPUS_EMB, SEED_QUERY_EMB, SEED_NEAR = 2345, 5432, 777
SEED_MASK = 999


def _gen(device, seed: int) -> torch.Generator:
    return torch.Generator(device=device).manual_seed(seed)


def observation_probs(rows: int, seed: int, device="cpu", chunk: int = 1 << 20) -> torch.Tensor:
    """sigmoid(1.5*randn + logit(prevalence)) -- 14 independent sigmoid probabilities per row."""
    g = _gen(device, seed)
    prev = torch.tensor(PREVALENCE, dtype=torch.float32, device=device)
    b = torch.log(prev) - torch.log1p(-prev)
    out = torch.empty((rows, 14), dtype=torch.float32, device=device)
    for s in range(0, rows, chunk):
        e = min(rows, s + chunk)
        out[s:e] = torch.sigmoid(1.5 * torch.randn((e - s, 14), generator=g, device=device) + b)
    return out


def embeddings(rows: int, d: int, seed: int, device="cpu", chunk: int = 1 << 18) -> torch.Tensor:
    """L2-normalised randn rows."""
    g = _gen(device, seed)
    out = torch.empty((rows, d), dtype=torch.float32, device=device)
    for s in range(0, rows, chunk):
        e = min(rows, s + chunk)
        out[s:e] = torch.nn.functional.normalize(torch.randn((e - s, d), generator=g, device=device), dim=-1)
    return out


def query_embeddings(rows: int, corpus: torch.Tensor, seed: int = SEED_QUERY_EMB, near_frac: float = 0.1,
                     noise: float = 0.3) -> torch.Tensor:
    """Query embeddings; ``near_frac`` of them are normalize(c[j] + noise*randn/sqrt(d)*sqrt(d)...) copies of
    random corpus rows so the top-k is not structureless."""
    device, d = corpus.device, corpus.shape[1]
    q = embeddings(rows, d, seed, device)
    n_near = int(rows * near_frac)
    if n_near > 0:
        g = _gen(device, SEED_NEAR)
        sel = torch.randperm(rows, generator=g, device=device)[:n_near]
        src = torch.randint(0, corpus.shape[0], (n_near,), generator=g, device=device)
        pert = noise * torch.randn((n_near, d), generator=g, device=device) / math.sqrt(d)
        q[sel] = torch.nn.functional.normalize(corpus[src] + pert, dim=-1)
    return q


def observation_masks(rows: int, round_idx: int, device="cpu") -> torch.Tensor:
    """uint8[rows,14]: round r keeps each observation with probability [1.0, 0.5, 0.25][r], >= 1 bit set."""
    keep = [1.0, 0.5, 0.25][min(round_idx, 2)]
    g = _gen(device, SEED_MASK + round_idx)
    m = (torch.rand((rows, 14), generator=g, device=device) < keep)
    empty = ~m.any(dim=1)
    if empty.any():
        first = torch.randint(0, 14, (rows,), generator=g, device=device)
        m[empty, first[empty]] = True
    return m.to(torch.uint8)


def mask_to_bits(mask: torch.Tensor) -> torch.Tensor:
    """uint8[Q,14] -> int16[Q] (CheXpert-14 bit order)."""
    w = (1 << torch.arange(14, device=mask.device, dtype=torch.int64))
    return (mask.to(torch.int64) * w).sum(dim=1).to(torch.int16)
