"""Generate tests/golden/oracle_golden.npz: small fixed inputs + float64-oracle outputs for every mode.

The reference ships no golden vectors for this path, so these are produced by
``oracle/retrieval_oracle.py``'s float64 direct-form evaluation -- the frozen specification -- and
committed; the CUDA path and the C oracle are both tested against them.

Inputs are bf16-representable (stored as uint16 bit patterns), so the tensor-core filter's bf16 cast is
exact on them and every precision mode must return the same ids.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import retrieval_oracle as ro  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz")


def bf16_round(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def main():
    rng = np.random.default_rng(424242)
    n, q, d = 300, 24, 512
    c_emb = rng.standard_normal((n, d)).astype(np.float32)
    c_emb = bf16_round(c_emb / np.linalg.norm(c_emb, axis=1, keepdims=True))
    q_emb = rng.standard_normal((q, d)).astype(np.float32)
    q_emb[:6] = c_emb[rng.integers(0, n, 6)] + 0.02 * rng.standard_normal((6, d)).astype(np.float32)
    q_emb = bf16_round(q_emb / np.linalg.norm(q_emb, axis=1, keepdims=True))
    prev = np.array([.05, .15, .25, .03, .12, .06, .07, .20, .03, .18, .01, .02, .30, .35])
    b = np.log(prev / (1 - prev))
    c_pr = (1 / (1 + np.exp(-(1.5 * rng.standard_normal((n, 14)) + b)))).astype(np.float32)
    q_pr = (1 / (1 + np.exp(-(1.5 * rng.standard_normal((q, 14)) + b)))).astype(np.float32)
    c_pr[0] = 0.0          # clamps to eps
    c_pr[0, 2] = 0.5       # keep the row sum non-zero for the normalize=True case
    c_pr[1] = 1.0          # log = 0
    q_pr[0] = c_pr[5]      # KL == 0 against row 5
    q_pr[1, :7] = 0.0
    mask = (rng.random((q, 14)) < 0.5)
    mask[:, 3] = True
    out = {
        "c_emb_bits": (c_emb.view(np.uint32) >> 16).astype(np.uint16),
        "q_emb_bits": (q_emb.view(np.uint32) >> 16).astype(np.uint16),
        "c_probs": c_pr, "q_probs": q_pr, "mask": mask.astype(np.uint8),
    }
    for k in (1, 10, 32):
        s, i = ro.search_fp64(ro.MODE_DPR, k, q_emb=q_emb, c_emb=c_emb)
        out[f"dpr_k{k}_s"], out[f"dpr_k{k}_i"] = s, i
        s, i = ro.search_fp64(ro.MODE_KL, k, q_probs=q_pr, c_probs=c_pr)
        out[f"kl_k{k}_s"], out[f"kl_k{k}_i"] = s, i
        s, i = ro.search_fp64(ro.MODE_KL, k, q_probs=q_pr, c_probs=c_pr, mask=mask)
        out[f"klmask_k{k}_s"], out[f"klmask_k{k}_i"] = s, i
        s, i = ro.search_fp64(ro.MODE_KL, k, q_probs=q_pr, c_probs=c_pr, normalize=True)
        out[f"klnorm_k{k}_s"], out[f"klnorm_k{k}_i"] = s, i
        for alpha in (0.0, 0.25, 0.5, 1.0):
            s, i = ro.search_fp64(ro.MODE_HYBRID, k, q_emb=q_emb, c_emb=c_emb, q_probs=q_pr, c_probs=c_pr,
                                  mask=mask, alpha=alpha)
            tag = f"hyb_a{int(alpha * 100):03d}_k{k}"
            out[tag + "_s"], out[tag + "_i"] = s, i
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
