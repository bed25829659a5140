"""The certificate of the tcgen05 filter rests on |canonical key - filter key| <= qerr ("Certificate").

This CPU test restates the filter arithmetic (bf16-rounded embedding operands, bf16 hi/lo split of the KL operands with
the lo*lo product dropped, exact accumulation) and the bound computed by ``query_pack_kernel`` (csrc/tc_filter.cuh) in
numpy, and checks the inequality against the CANONICAL keys of the C oracle for every (query, case) pair of seeded
problems, including adversarial ones (probabilities at the clamp, masked queries, non-unit embeddings).  The GPU suite
checks the same inequality's consequence (certified results are bit-identical); this one pins the error model itself.
"""
import numpy as np
import pytest
import torch

from conftest import make_problem
from oracle import c_oracle as co


def bf16(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).bfloat16().float().numpy()


def filter_keys_and_bound(mode, q_emb, p16, ent, c_emb, logq, alpha, col_max):
    """float64 emulation of the filter key of every pair + the fp32 bound of query_pack_kernel."""
    oma = np.float32(1.0) - np.float32(alpha)
    f = np.zeros((p16.shape[0] if p16 is not None else q_emb.shape[0], (c_emb if c_emb is not None else logq).shape[0]))
    ip_mag = kl_mag = shift = 0.0
    if mode != co.MODE_KL:
        a = (np.float32(alpha) if mode == co.MODE_HYBRID else np.float32(1.0)) * q_emb.astype(np.float32)
        f += bf16(a).astype(np.float64) @ bf16(c_emb).astype(np.float64).T
        ip_mag = np.sqrt((a.astype(np.float64) ** 2).sum(1)) * np.linalg.norm(c_emb.astype(np.float64), axis=1).max()
    if mode != co.MODE_DPR:
        v = (oma if mode == co.MODE_HYBRID else np.float32(1.0)) * p16.astype(np.float32)
        v_hi = bf16(v)
        v_lo = bf16(v - v_hi)
        l_hi = bf16(logq)
        l_lo = bf16(logq - l_hi)
        d = np.float64
        f += v_hi.astype(d) @ l_hi.astype(d).T + v_hi.astype(d) @ l_lo.astype(d).T + v_lo.astype(d) @ l_hi.astype(d).T
        shift = ((oma * ent) if mode == co.MODE_HYBRID else ent).astype(np.float64)
        f -= shift[:, None]
        kl_mag = (np.abs(v).astype(np.float64) * (col_max * 1.0001)[None, :]).sum(1)
    e = 0.00403 * ip_mag + 1.65e-4 * kl_mag + 1e-6 * (np.abs(shift) + ip_mag + kl_mag) + 1e-30
    return f, np.broadcast_to(e, (f.shape[0],))


def canonical_keys(mode, p, p16, ent, logq, alpha):
    n = p["c_emb"].shape[0]
    ids = np.tile(np.arange(n, dtype=np.int64), (p["q_emb"].shape[0], 1))
    s = co.score_pairs(mode, ids, q_emb=p["q_emb"], p16=p16, entropy=ent, c_emb=p["c_emb"], logq16=logq, alpha=alpha)
    return (-s if mode == co.MODE_KL else s).astype(np.float64)  # KL is reported as +KL; its key is -KL


@pytest.mark.parametrize("mode", [co.MODE_DPR, co.MODE_KL, co.MODE_HYBRID])
@pytest.mark.parametrize("alpha", [0.5, 0.05, 0.95])
@pytest.mark.parametrize("variant", ["plain", "clamped", "scaled"])
def test_filter_error_bound_holds_for_every_pair(mode, alpha, variant):
    p = make_problem(1500, 48, seed=61)
    if variant == "clamped":  # probabilities at both ends of the clamp range: |log q| up to 18.4
        p["c_pr"][::3, ::2] = 0.0
        p["c_pr"][1::3, 1::2] = 1.0
        p["q_pr"][::2, :5] = 1.0
    if variant == "scaled":  # the index does not require unit-norm embeddings
        p["c_emb"] *= np.linspace(0.2, 3.0, p["c_emb"].shape[0], dtype=np.float32)[:, None]
        p["q_emb"] *= 2.5
    logq = co.prepare_corpus(p["c_pr"])
    p16, ent = co.prepare_queries(p["q_pr"], p["mask"])
    col_max = np.abs(logq).max(axis=0).astype(np.float64)
    f, e = filter_keys_and_bound(mode, p["q_emb"], p16, ent, p["c_emb"], logq, alpha, col_max)
    c = canonical_keys(mode, p, p16, ent, logq, alpha)
    err = np.abs(c - f)
    worst = (err / e[:, None]).max()
    assert worst <= 1.0, f"bound violated: worst |canonical - filter| / qerr = {worst:.3f}"
    if mode != co.MODE_KL:
        assert worst >= 0.01  # the bound is not vacuous: bf16 rounding of 512-d operands really is of this order
