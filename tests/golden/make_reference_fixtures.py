"""Generate tests/golden/reference_fixtures.json by RUNNING THE REFERENCE ITSELF.

Run in the authoring container only (``python tests/golden/make_reference_fixtures.py``): it imports
/root/reference, which does not exist on the GPU box.  The committed JSON is what the tests read.

The reference's scoring operator is ``faiss.IndexFlatIP`` (modeling_dense_passage_retrieval.py:297-313),
a third-party package that is not installed here and not vendored by the reference.  To exercise the
reference's OWN code around that operator (k defaulting/clamping :306-308, list conversion :314, the
hard-negative split :320-331, the fallback :318) a stand-in ``faiss`` module is injected whose
``IndexFlatIP`` follows faiss's documented brute-force semantics, written independently with plain
numpy (argsort of the fp32 product, descending, stable => ties by insertion order).
Everything else (re-rank :127-152, query text :115-125, detector :38-61, consistency :70-92, the RAG
loop :198-275) is pure reference Python and runs unmodified.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.json")


class _StandInIndexFlatIP:
    def __init__(self, d):
        self.d = d
        self.x = np.zeros((0, d), dtype=np.float32)

    @property
    def ntotal(self):
        return self.x.shape[0]

    def add(self, x):
        self.x = np.vstack([self.x, np.asarray(x, dtype=np.float32)])

    def search(self, x, k):
        s = np.asarray(x, dtype=np.float32) @ self.x.T
        order = np.argsort(-s, axis=1, kind="stable")[:, :k]
        return np.take_along_axis(s, order, axis=1), order.astype(np.int64)


def main() -> None:
    faiss = types.ModuleType("faiss")
    faiss.IndexFlatIP = _StandInIndexFlatIP
    sys.modules["faiss"] = faiss
    sys.path.insert(0, REF)
    import torch
    from annotate_retrieve import modeling_dense_passage_retrieval as dpr
    from annotate_retrieve import modeling_iterative_rag as rag

    assert dpr.HAS_FAISS
    fx = {"generated_from": "MOsama10/radar-multimodal-radiology @ /root/reference", "cases": {}}

    # ---- D2 / D3: HybridRetriever wrapper behaviour around the index ---------------------------------
    rng = np.random.default_rng(20260101)
    n, d = 40, 512
    emb = rng.standard_normal((n, d)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    # bf16-representable values so that every precision mode of the GPU path sees identical inputs
    emb = (emb.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    passages = [f"passage {i}" for i in range(n)]

    class FakeEmbedder:
        def encode_text(self, texts):
            idx = [int(t.split()[1]) for t in texts]
            return torch.from_numpy(emb[idx])

    cfg = dpr.RetrievalConfig(device="cpu")
    hr = dpr.HybridRetriever(cfg, FakeEmbedder())
    hr.build_indices(passages, [[] for _ in passages])
    assert hr.semantic_index is not None and hr.semantic_index.ntotal == n
    queries = rng.standard_normal((6, d)).astype(np.float32)
    queries /= np.linalg.norm(queries, axis=1, keepdims=True)
    queries = (queries.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    runs = []
    for qi, k in [(0, None), (1, 5), (2, 8), (3, 1), (4, 40), (5, 1000)]:
        got_p, got_s = hr.retrieve(torch.from_numpy(queries[qi]), k)
        assert all(isinstance(p, str) for p in got_p) and all(isinstance(s, float) for s in got_s)
        runs.append({"query": qi, "k": k, "passages": got_p, "scores": got_s})
    hn = hr.retrieve_with_hard_negatives(torch.from_numpy(queries[0]))
    hn2 = hr.retrieve_with_hard_negatives(torch.from_numpy(queries[1]), k=4, num_negatives=2)
    empty = dpr.HybridRetriever(cfg, FakeEmbedder())
    empty.build_indices([], [])
    e_p, e_s = empty.retrieve(torch.from_numpy(queries[0]), 5)
    fx["cases"]["hybrid_retriever"] = {
        "n": n, "d": d,
        "embeddings_bf16_bits": (emb.view(np.uint32) >> 16).astype(np.uint16).tolist(),
        "queries_bf16_bits": (queries.view(np.uint32) >> 16).astype(np.uint16).tolist(),
        "config_defaults": {"embedding_dim": cfg.embedding_dim, "num_retrieved": cfg.num_retrieved,
                            "hybrid_alpha": cfg.hybrid_alpha, "device": dpr.RetrievalConfig().device},
        "retrieve": runs,
        "hard_negatives_default": hn,
        "hard_negatives_k4_n2": hn2,
        "empty_index": {"passages": e_p, "scores": e_s},
    }

    # ---- R2 and friends: pure-Python pieces of the iterative-RAG module --------------------------------
    rcfg = rag.IterativeRAGConfig(device="cpu")
    fx["cases"]["rag_config_defaults"] = {
        "num_iterations": rcfg.num_iterations, "max_new_tokens": rcfg.max_new_tokens, "top_k": rcfg.top_k,
        "temperature": rcfg.temperature, "consistency_threshold": rcfg.consistency_threshold,
        "observation_vocab": rcfg.observation_vocab, "device": rag.IterativeRAGConfig().device}
    det = rag.ObservationDetector(rcfg)
    fx["cases"]["default_vocab"] = det.observation_vocab
    texts = [
        "Mild cardiomegaly with small left pleural effusion.",
        "No finding. Support devices in place.",
        "pulmonary edema and rib fracture; atelectasis at the bases",
        "", "Lung opacity concerning for pneumonia or consolidation", "ENLARGED CARDIOMEDIASTINUM",
    ]
    fx["cases"]["detect_observations"] = [{"text": t, "found": sorted(det.detect_observations(t))} for t in texts]
    tr = rag.TargetedRetriever(rcfg)
    rr = []
    missing_sets = [["Cardiomegaly", "Atelectasis"], ["Pneumonia"], ["Edema", "Fracture", "Pleural Effusion"], []]
    pass_sets = [
        ["Report with cardiomegaly and atelectasis", "Report with Cardiomegaly only", "Unremarkable",
         "atelectasis, pneumonia", "Pulmonary edema and rib fracture noted"],
        ["pneumonia", "no finding", "PNEUMONIA bilateral", "edema"],
        ["edema", "pleural effusion with fracture and edema", "fracture", "nothing", "edema fracture"],
        ["a", "b"],
    ]
    for ms, ps in zip(missing_sets, pass_sets):
        ranked = tr.rank_retrieved_passages(ps, set(ms))
        rr.append({"missing": ms, "passages": ps, "ranked": [[p, s] for p, s in ranked]})
    rr.append({"missing": ["Edema"], "passages": [], "ranked": tr.rank_retrieved_passages([], {"Edema"})})
    fx["cases"]["rank_retrieved_passages"] = rr
    fx["cases"]["build_retrieval_query"] = [
        {"missing": [], "context": "", "query": tr.build_retrieval_query(set())},
        {"missing": ["Edema"], "context": "", "query": tr.build_retrieval_query({"Edema"})},
        {"missing": ["Edema"], "context": "PA view", "query": tr.build_retrieval_query({"Edema"}, "PA view")},
    ]
    cv = rag.ConsistencyVerifier(rcfg)
    gens = [["cardiomegaly and edema", "cardiomegaly", "cardiomegaly with pneumonia"], ["x"], ["a", "b"],
            ["edema", "edema"]]
    fx["cases"]["consistency"] = [{"generations": g, "score": cv.compute_consistency(g),
                                   "consistent": sorted(cv.find_consistent_observations(g))} for g in gens]

    # ---- R1: the loop's call pattern at the retrieval seam (module __main__ mocks, :329-341) ----------
    model = rag.create_iterative_rag_model()
    calls = []

    def mock_retrieval(query, k):
        calls.append([query, k])
        return [f"Report {i} about {query[:20]}" for i in range(k)], [0.9 - i * 0.05 for i in range(k)]

    def mock_generation(context):
        return f"Generated report: {context[:50]}..."

    res = model.generate_with_iterative_retrieval("Initial findings", mock_retrieval, mock_generation,
                                                  reference_text="Reference with Cardiomegaly and Atelectasis")
    fx["cases"]["rag_loop_with_reference"] = {
        "num_calls": len(calls), "ks": [c[1] for c in calls],
        "query_words": sorted(calls[0][0].replace("Cases with ", "").split(", ")),
        "iterations": res["iterations"], "num_generations": len(res["generations"]),
        "retrieved_scores": res["retrieved_scores"], "num_retrieved_passages": len(res["retrieved_passages"]),
        "final_consistency": res["final_consistency"],
        "cumulative_findings": sorted(res["cumulative_findings"]),
    }
    calls.clear()
    res2 = model.generate_with_iterative_retrieval("Initial findings", mock_retrieval, mock_generation)
    fx["cases"]["rag_loop_without_reference"] = {"num_calls": len(calls), "iterations": res2["iterations"],
                                                 "num_generations": len(res2["generations"]),
                                                 "final_text": res2["final_text"]}
    calls.clear()
    res3 = model.generate_with_iterative_retrieval("Initial", lambda q, k: ([], []), mock_generation,
                                                   reference_text="Edema")
    fx["cases"]["rag_loop_empty_retrieval"] = {"iterations": res3["iterations"],
                                               "num_generations": len(res3["generations"])}
    ver = model.generate_with_verification("Initial prompt", mock_generation, num_samples=3)
    fx["cases"]["verification"] = {"consistency_score": ver["consistency_score"],
                                   "num_generations": len(ver["generations"]),
                                   "best_generation": ver["best_generation"]}

    # ---- evaluate_retrieval_system.RetrievalMetrics (:137-188) ------------------------------------------
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_eval", os.path.join(REF, "evaluate_retrieval_system.py"))
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        RM = mod.RetrievalMetrics
        mcases = []
        for retrieved, relevant in [([3, 7, 1, 9, 4, 0, 2, 8, 6, 5], [9, 2]), ([1, 2, 3], [4]), ([5, 4, 3, 2, 1], [5, 4, 3]),
                                    (list(range(12)), [11, 0, 30])]:
            r_s, rel_s = [f"id{v:03d}" for v in retrieved], {f"id{v:03d}" for v in relevant}
            mcases.append({
                "retrieved": retrieved, "relevant": relevant,
                "mrr": RM.calculate_mrr(r_s, rel_s),
                "precision@1": RM.calculate_precision_at_k(r_s, rel_s, 1),
                "precision@5": RM.calculate_precision_at_k(r_s, rel_s, 5),
                "precision@10": RM.calculate_precision_at_k(r_s, rel_s, 10),
                "recall@5": RM.calculate_recall_at_k(r_s, rel_s, 5),
                "recall@10": RM.calculate_recall_at_k(r_s, rel_s, 10),
                "ndcg@5": float(RM.calculate_ndcg_at_k(r_s, rel_s, 5)),
                "ndcg@10": float(RM.calculate_ndcg_at_k(r_s, rel_s, 10)),
                "accuracy@5": RM.calculate_retrieval_accuracy_at_5(r_s, rel_s),
                "accuracy@10": RM.calculate_retrieval_accuracy_at_10(r_s, rel_s),
            })
        fx["cases"]["retrieval_metrics"] = mcases
    except Exception as e:  # the module mkdirs a Windows path at import; record why if it cannot load
        fx["cases"]["retrieval_metrics_unavailable"] = repr(e)

    with open(OUT, "w") as fh:
        json.dump(fx, fh, indent=1, sort_keys=True)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
