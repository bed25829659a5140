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


def f16(x):
    return np.ascontiguousarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def fp16_filter_keys_and_bound(fmt, p16, ent, logq, col_max):
    """float64 emulation of the KL-only fp16 filters (csrc/kl_filter.cuh: fmt 1 = one product fp16(2^13 v) . fp16(2^11 L),
    fmt 2 = query split hi/lo, two products) + the fp32 bound of klf_pack_kernel."""
    v = p16.astype(np.float32)
    vs = v * np.float32(8192.0)
    v_hi = f16(vs)
    v_lo = f16(vs - v_hi)
    l16 = f16(logq.astype(np.float32) * np.float32(2048.0))
    assert np.all((l16 == 0) | (np.abs(l16) >= 6.2e-5)) and np.all((v_hi == 0) | (v_hi >= 6.2e-5)), "subnormal operand"
    d = np.float64
    acc = v_hi.astype(d) @ l16.astype(d).T
    if fmt == 2:
        acc += v_lo.astype(d) @ l16.astype(d).T
    f = acc / 2.0 ** 24 - ent.astype(d)[:, None]
    kl_mag = (np.abs(v).astype(d) * (col_max * 1.0001)[None, :]).sum(1)
    sl = ((v != 0) * (col_max * 1.0001)[None, :]).sum(1)
    rel = 1.13e-3 if fmt == 1 else 6.3e-4
    e = rel * kl_mag + 8e-9 * sl + 1e-6 * (np.abs(ent.astype(d)) + kl_mag) + 1e-30
    return f, e


@pytest.mark.parametrize("fmt", [1, 2])
@pytest.mark.parametrize("variant", ["plain", "clamped", "near_one"])
def test_fp16_kl_filter_error_bound_holds_for_every_pair(fmt, variant):
    p = make_problem(1500, 48, seed=62)
    if variant == "clamped":
        p["c_pr"][::3, ::2] = 0.0
        p["c_pr"][1::3, 1::2] = 1.0
        p["q_pr"][::2, :5] = 1.0
        p["q_pr"][1::2, 5:9] = 0.0   # clamped to 1e-8: the smallest non-zero query weight
    if variant == "near_one":  # log q just below zero: the smallest non-zero table entries
        p["c_pr"][::2, :7] = np.float32(1.0) - np.float32(2.0 ** -24)
        p["c_pr"][1::2, 7:] = np.float32(1.0) - np.float32(2.0 ** -20)
    logq = co.prepare_corpus(p["c_pr"])
    p16, ent = co.prepare_queries(p["q_pr"], p["mask"])
    col_max = np.abs(logq).max(axis=0).astype(np.float64)
    f, e = fp16_filter_keys_and_bound(fmt, p16, ent, logq, col_max)
    c = canonical_keys(co.MODE_KL, p, p16, ent, logq, 0.5)
    worst = (np.abs(c - f) / e[:, None]).max()
    assert worst <= 1.0, f"bound violated: worst |canonical - filter| / qerr = {worst:.3f}"
    assert worst >= 0.01
