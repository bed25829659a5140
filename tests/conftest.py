import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def bits_to_f32(bits) -> np.ndarray:
    """uint16 bf16 bit patterns -> float32."""
    return (np.asarray(bits, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


@pytest.fixture(scope="session")
def golden():
    g = dict(np.load(os.path.join(GOLDEN, "oracle_golden.npz")))
    g["c_emb"] = bits_to_f32(g["c_emb_bits"])
    g["q_emb"] = bits_to_f32(g["q_emb_bits"])
    return g


@pytest.fixture(scope="session")
def ref_fixtures():
    import json
    with open(os.path.join(GOLDEN, "reference_fixtures.json")) as fh:
        return json.load(fh)["cases"]


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and load the C-ABI library; no compute is run."""
    from radar_multimodal_radiology_b200 import _lib
    _lib.build()
    return _lib.lib()


def make_problem(n, q, d=512, seed=0, near=True):
    """Seeded synthetic inputs as numpy arrays (CPU), shapes of SURVEY.md section 8d."""
    import torch
    from radar_multimodal_radiology_b200 import synthetic as syn
    c_emb = syn.embeddings(n, d, syn.SEED_CORPUS_EMB + seed)
    q_emb = syn.query_embeddings(q, c_emb, syn.SEED_QUERY_EMB + seed) if near else syn.embeddings(q, d, 99 + seed)
    c_pr = syn.observation_probs(n, syn.SEED_CORPUS_PROBS + seed)
    q_pr = syn.observation_probs(q, syn.SEED_QUERY_PROBS + seed)
    mask = syn.observation_masks(q, 1)
    return {k: v.numpy() for k, v in dict(c_emb=c_emb, q_emb=q_emb, c_pr=c_pr, q_pr=q_pr, mask=mask).items()}


def topk_sets_match(ids_a, ids_b, scores_true_of_a, scores_true_of_b, tol):
    """Tolerance-band comparison of two top-k id lists (SURVEY.md 'hard part 5'): ids may differ only
    where the TRUE scores of the differing elements are within tol of each other."""
    bad = 0
    for ra, rb, sa, sb in zip(ids_a, ids_b, scores_true_of_a, scores_true_of_b):
        only_a = [s for i, s in zip(ra, sa) if i not in set(rb)]
        only_b = [s for i, s in zip(rb, sb) if i not in set(ra)]
        for x, y in zip(sorted(only_a), sorted(only_b)):
            if abs(x - y) > tol:
                bad += 1
    return bad


def assert_topk_within_ties(ids, want_ids, true_scores, want_scores, tol, what=""):
    """north_star: "top-k index lists identical except where tied scores fall within tolerance".

    ``true_scores`` is the float64 ground-truth score MATRIX [Q,N] (API sign), ``want_ids`` / ``want_scores`` the golden
    top-k of that matrix, ``tol`` a scalar or a [Q,k] array.  (1) every position where the returned id differs from the
    golden one must hold a case whose TRUE score is within tol of the golden score at that rank (a swap inside a tie
    band); (2) the two id SETS may differ only by cases whose true scores are within tol of each other
    (:func:`topk_sets_match`).  Nothing else is tolerated -- in particular not "97 % of the ids agree"."""
    import numpy as np
    ids, want_ids = np.asarray(ids), np.asarray(want_ids)
    tol_a = np.broadcast_to(np.asarray(tol, dtype=np.float64), ids.shape)
    got_true = np.take_along_axis(true_scores, ids, axis=1)
    mism = ids != want_ids
    worst = np.abs(got_true - want_scores)[mism]
    assert np.all(worst <= tol_a[mism] + 1e-12), f"{what}: {int((worst > tol_a[mism] + 1e-12).sum())} positions differ outside the tie band"
    want_true = np.take_along_axis(true_scores, want_ids, axis=1)
    bad = topk_sets_match(ids, want_ids, got_true, want_true, float(np.max(tol_a)))
    assert bad == 0, f"{what}: {bad} set differences outside the tie band"
    return int(mism.sum())
