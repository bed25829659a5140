"""CPU: the oracle against its golden vectors, against itself (fp64 / BLAS fp32 / canonical C), and the
size-independent properties of SURVEY.md section 4 (alpha=1 == DPR order, alpha=0 == KL order, row
permutation invariance, shard-merge == single shard)."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import retrieval_oracle as ro
from conftest import assert_topk_within_ties, make_problem


def _prep(p, masked=True):
    logq = co.prepare_corpus(p["c_pr"])
    p16, ent = co.prepare_queries(p["q_pr"], p["mask"] if masked else None)
    return logq, p16, ent


def test_c_and_numpy_preparation_agree():
    p = make_problem(500, 40)
    for normalize in (False, True):
        assert np.array_equal(co.prepare_corpus(p["c_pr"], normalize=normalize),
                              ro.prepare_corpus_logq(p["c_pr"], normalize=normalize))
        a, ha = co.prepare_queries(p["q_pr"], p["mask"], normalize=normalize)
        b, hb = ro.prepare_queries(p["q_pr"], p["mask"], normalize=normalize)
        assert np.array_equal(a, b) and np.array_equal(ha, hb)
    logq = co.prepare_corpus(p["c_pr"])
    assert logq.shape == (500, 16) and np.all(logq[:, 14:] == 0) and np.all(logq[:, :14] <= 0)
    ref = np.log(np.clip(p["c_pr"].astype(np.float64), 1e-8, 1))
    assert np.max(np.abs(logq[:, :14] - ref)) <= 2e-6


@pytest.mark.parametrize("k", [1, 10, 32])
def test_golden_vectors_fp64_oracle_is_frozen(golden, k):
    """Re-running the fp64 oracle reproduces the committed fixtures exactly."""
    g = golden
    s, i = ro.search_fp64(ro.MODE_DPR, k, q_emb=g["q_emb"], c_emb=g["c_emb"])
    assert np.array_equal(i, g[f"dpr_k{k}_i"]) and np.allclose(s, g[f"dpr_k{k}_s"], rtol=0, atol=1e-12)
    s, i = ro.search_fp64(ro.MODE_KL, k, q_probs=g["q_probs"], c_probs=g["c_probs"], mask=g["mask"])
    assert np.array_equal(i, g[f"klmask_k{k}_i"]) and np.allclose(s, g[f"klmask_k{k}_s"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("k", [1, 10, 32])
def test_golden_vectors_canonical_c_oracle(golden, k):
    """Canonical fp32 arithmetic (what the CUDA path reproduces bit for bit) vs the fp64 fixtures:
    ids identical (the fixtures have no near-ties at fp32 resolution), scores within 1e-5 of the operand scale."""
    g = golden
    logq = co.prepare_corpus(g["c_probs"])
    s, i = co.search(co.MODE_DPR, k, q_emb=g["q_emb"], c_emb=g["c_emb"])
    assert np.array_equal(i, g[f"dpr_k{k}_i"])
    assert np.max(np.abs(s - g[f"dpr_k{k}_s"])) <= 1e-5 * max(1.0, np.abs(g[f"dpr_k{k}_s"]).max())
    for tag, mask, normalize in (("kl", None, False), ("klmask", g["mask"], False), ("klnorm", None, True)):
        lq = co.prepare_corpus(g["c_probs"], normalize=normalize)
        p16, ent = co.prepare_queries(g["q_probs"], mask, normalize=normalize)
        s, i = co.search(co.MODE_KL, k, p16=p16, entropy=ent, logq16=lq)
        want_s, want_i = g[f"{tag}_k{k}_s"], g[f"{tag}_k{k}_i"]
        scale = ro.kl_operand_scale_fp64(g["q_probs"], g["c_probs"], mask, normalize=normalize)
        tol = 1e-5 * np.take_along_axis(scale, want_i, axis=1)
        assert np.all(np.abs(s - want_s) <= tol + 1e-12)
        # KL near 0 cancels: ids may swap only inside the tolerance band (tie-aware comparison, not an agreement rate)
        true = ro.kl_matrix_fp64(g["q_probs"], g["c_probs"], mask, normalize=normalize)
        assert_topk_within_ties(i, want_i, true, want_s, 2 * tol, f"{tag} k={k}")
    p16, ent = co.prepare_queries(g["q_probs"], g["mask"])
    for alpha in (0.0, 0.25, 0.5, 1.0):
        s, i = co.search(co.MODE_HYBRID, k, q_emb=g["q_emb"], p16=p16, entropy=ent, c_emb=g["c_emb"], logq16=logq,
                         alpha=alpha)
        tag = f"hyb_a{int(alpha * 100):03d}_k{k}"
        true = ro.score_matrix_fp64(ro.MODE_HYBRID, q_emb=g["q_emb"], c_emb=g["c_emb"], q_probs=g["q_probs"],
                                    c_probs=g["c_probs"], mask=g["mask"], alpha=alpha)
        assert_topk_within_ties(i, g[tag + "_i"], true, g[tag + "_s"], 2e-4, tag)
        assert np.max(np.abs(s - g[tag + "_s"])) <= 2e-4


def test_kl_zero_for_identical_distribution(golden):
    """The un-normalised sum p (log p - log q) is exactly 0 for q == p (it is NOT >= 0 in general: a case
    with all probabilities 1 scores sum p log p < 0 -- SURVEY.md section 8c adopts it knowingly)."""
    g = golden
    kl = ro.kl_matrix_fp64(g["q_probs"][:1], g["c_probs"])
    assert abs(kl[0, 5]) < 1e-15 and kl[0, 1] < 0
    logq = co.prepare_corpus(g["c_probs"])
    p16, ent = co.prepare_queries(g["q_probs"][:1])
    pair = co.score_pairs(co.MODE_KL, np.array([[5]]), p16=p16, entropy=ent, logq16=logq)
    assert pair[0, 0] == 0.0
    # with normalize=True it is a categorical KL: non-negative, minimised by the identical row
    s, i = ro.search_fp64(ro.MODE_KL, 1, q_probs=g["q_probs"][:1], c_probs=g["c_probs"], normalize=True)
    assert i[0, 0] == 5 and abs(s[0, 0]) < 1e-12


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_three_oracle_flavours_agree(mode):
    p = make_problem(4000, 64, seed=3)
    logq, p16, ent = _prep(p)
    k = 10
    s64, i64 = ro.search_fp64(mode, k, q_emb=p["q_emb"], c_emb=p["c_emb"], q_probs=p["q_pr"], c_probs=p["c_pr"],
                              mask=p["mask"], alpha=0.5)
    sc, ic = co.search(mode, k, q_emb=p["q_emb"], p16=p16, entropy=ent, c_emb=p["c_emb"], logq16=logq, alpha=0.5)
    sb, ib = ro.search_blas32(mode, k, q_emb=p["q_emb"], c_emb=p["c_emb"], q_p16=p16, q_entropy=ent,
                              c_logq16=logq, alpha=0.5)
    assert np.mean(ic == i64) > 0.995 and np.mean(ib == i64) > 0.995
    assert np.max(np.abs(sc - s64)) < 2e-5 and np.max(np.abs(sb - s64)) < 2e-5
    pairs = co.score_pairs(mode, ic, q_emb=p["q_emb"], p16=p16, entropy=ent, c_emb=p["c_emb"], logq16=logq, alpha=0.5)
    assert np.array_equal(pairs, sc)


def test_alpha_limits_reproduce_pure_orderings():
    p = make_problem(3000, 32, seed=5)
    logq, p16, ent = _prep(p)
    kw = dict(q_emb=p["q_emb"], p16=p16, entropy=ent, c_emb=p["c_emb"], logq16=logq)
    _, i_dpr = co.search(co.MODE_DPR, 10, **kw)
    s1, i1 = co.search(co.MODE_HYBRID, 10, alpha=1.0, **kw)
    assert np.array_equal(i1, i_dpr)
    s_kl, i_kl = co.search(co.MODE_KL, 10, **kw)
    s0, i0 = co.search(co.MODE_HYBRID, 10, alpha=0.0, **kw)
    assert np.array_equal(i0, i_kl) and np.array_equal(s0, -s_kl)


def test_row_permutation_invariance():
    p = make_problem(2500, 16, seed=7)
    logq, p16, ent = _prep(p)
    perm = np.random.default_rng(1).permutation(2500)
    s, i = co.search(co.MODE_HYBRID, 8, q_emb=p["q_emb"], p16=p16, entropy=ent, c_emb=p["c_emb"], logq16=logq)
    s2, i2 = co.search(co.MODE_HYBRID, 8, q_emb=p["q_emb"], p16=p16, entropy=ent, c_emb=p["c_emb"][perm],
                       logq16=logq[perm])
    assert np.array_equal(s, s2) and np.array_equal(perm[i2], i)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_shard_merge_equals_single_shard(mode, world):
    p = make_problem(1003, 20, seed=11)
    logq, p16, ent = _prep(p)
    k = 10
    full_s, full_i = co.search(mode, k, q_emb=p["q_emb"], p16=p16, entropy=ent, c_emb=p["c_emb"], logq16=logq)
    parts_s, parts_i = [], []
    for lo, hi in ro.shard_bounds(1003, world):
        s, i = co.search(mode, k, q_emb=p["q_emb"], p16=p16, entropy=ent, c_emb=p["c_emb"][lo:hi],
                         logq16=logq[lo:hi], idx_offset=lo)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = co.merge_topk(np.stack(parts_s), np.stack(parts_i), k, ascending=(mode == co.MODE_KL))
    assert np.array_equal(ms, full_s) and np.array_equal(mi, full_i)
    ms2, mi2 = ro.merge_topk(list(zip(parts_s, parts_i)), k, descending=(mode != co.MODE_KL))
    assert np.array_equal(mi2, full_i)


def test_ties_break_by_smaller_id():
    c = np.zeros((10, 8), dtype=np.float32)
    c[:, 0] = 1.0
    q = np.zeros((1, 8), dtype=np.float32)
    q[0, 0] = 1.0
    s, i = co.search(co.MODE_DPR, 4, q_emb=q, c_emb=c)
    assert i.tolist() == [[0, 1, 2, 3]] and np.all(s == 1.0)
    s, i = ro.search_fp64(ro.MODE_DPR, 4, q_emb=q, c_emb=c)
    assert i.tolist() == [[0, 1, 2, 3]]


def test_index_flat_ip_restatement_semantics():
    """faiss.IndexFlatIP members used at dpr.py:297-313."""
    rng = np.random.default_rng(0)
    idx = ro.IndexFlatIP(16)
    assert idx.ntotal == 0 and bool(idx)
    x = rng.standard_normal((7, 16)).astype(np.float32)
    idx.add(x[:3])
    idx.add(x[3:])
    assert idx.ntotal == 7
    d, i = idx.search(x[:2], 10)
    assert d.shape == (2, 10) and i.dtype == np.int64
    assert np.all(i[:, 7:] == -1) and np.all(np.diff(d[:, :7], axis=1) <= 0)
    assert i[0, 0] == 0 and i[1, 0] == 1


def test_rerank_bitmask_form_matches_string_form(ref_fixtures):
    for case in ref_fixtures["rank_retrieved_passages"]:
        want = [tuple(x) for x in case["ranked"]]
        got = ro.rerank_scores(case["passages"], set(case["missing"]))
        assert [g[0] for g in got] == [w[0] for w in want]
        assert np.allclose([g[1] for g in got], [w[1] for w in want], rtol=0, atol=0)


def test_retrieval_metrics_match_reference(ref_fixtures):
    for case in ref_fixtures["retrieval_metrics"]:
        got = ro.retrieval_metrics(case["retrieved"], case["relevant"])
        for key in ("mrr", "precision@1", "precision@5", "precision@10", "recall@5", "recall@10", "ndcg@5",
                    "ndcg@10", "accuracy@5", "accuracy@10"):
            assert got[key] == pytest.approx(case[key], abs=1e-12), key
