#!/usr/bin/env python
"""bench.py -- the retrieval hot path on N GPUs of one node, one JSON line on rank 0.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU path (oracle port) on the host cores, same metric

A "step" is one pass of the hot path over one batch of synthetic queries: query preparation, the fused
score + top-k kernels over this rank's corpus shard, and for N > 1 the all-gather + merge of the per-rank
top-k lists.  The HEADLINE workload is `hybrid_10m` -- the configuration BASELINE.json's metric "queries/sec at
top-k=10 over 10M-case corpus" is quoted on: hybrid KL+DPR, 10M cases row-sharded over the N GPUs, 16 384
queries per step, top-k = 10 -- in FP32-CERTIFIED precision (result bit-identical to the canonical fp32 definition,
the reference's arithmetic; dpr.py:312-313).  The default run then measures, on the same resident corpus, every
other BASELINE configuration as `secondary[...]` records (a few steps each):
    hybrid_10m_bf16        the headline with the bf16 filter only (no certificate)
    hybrid_10m_k32         configs[3] as written: top-k = 32
    rag_rounds             configs[4]: 3 masked re-retrieval rounds x 16 384 queries, top-k = 5
    kl_latency             KL-only, 10M cases, 32 queries, top-k = 32 (the HBM-bound regime, BASELINE.md row 4')
    dpr_latency            DPR, 10M cases, ONE query per call, top-k = 10: the reference's own call pattern
                           (IndexFlatIP.search with nq = 1, dpr.py:312-314) -- HBM-bound: 1 024 B of bf16 embedding per case
    dpr_377k / kl_377k     configs[2] / configs[1]: 377k cases, 65 536 queries, top-k = 10
    clustered_2m / adversarial_2m   structured corpora (clustered embeddings; rows sorted so that every query's
                           scores keep rising along the sweep) -- the threshold filter's rare path under timing
and a `parity` record per workload: >= 256 of the step's queries re-run by the exact CUDA-core scan (at N > 1:
per-rank exact scan, scores and ids all-gathered separately and merged with torch ops -- independent of the
packed-word exchange + merge kernel the timed path uses), compared with what the timed path returned.
`value` is timed with inputs resident in HBM; `e2e` times the same call from pinned HOST buffers
(H2D of the queries + D2H of scores/ids inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "hybrid_10m": dict(mode="hybrid", n=10_000_000, q=16384, k=10),
    "hybrid_10m_bf16": dict(mode="hybrid", n=10_000_000, q=16384, k=10, precision="bf16"),
    "hybrid_10m_k32": dict(mode="hybrid", n=10_000_000, q=16384, k=32),
    "dpr_10m": dict(mode="dpr", n=10_000_000, q=16384, k=10),
    "kl_377k": dict(mode="kl", n=377_000, q=65536, k=10),
    "dpr_377k": dict(mode="dpr", n=377_000, q=65536, k=10),
    "kl_latency": dict(mode="kl", n=10_000_000, q=32, k=32),
    "dpr_latency": dict(mode="dpr", n=10_000_000, q=1, k=10),
    "rag_rounds": dict(mode="hybrid", n=10_000_000, q=16384, k=5, masked=True, rounds=3),
    "clustered_2m": dict(mode="hybrid", n=2_000_000, q=16384, k=10, corpus="clustered"),
    "adversarial_2m": dict(mode="hybrid", n=2_000_000, q=16384, k=10, corpus="adversarial"),
    "smoke": dict(mode="hybrid", n=200_000, q=1024, k=10),
}
SECONDARY = ["hybrid_10m_bf16", "hybrid_10m_k32", "rag_rounds", "kl_latency", "dpr_latency", "dpr_377k", "kl_377k",
             "clustered_2m", "adversarial_2m"]
D = 512
ALPHA = 0.5
GEN_BLOCK = 1_250_000  # corpus rows are generated in fixed blocks so the corpus is identical for every N
KERNEL_NAMES = {1: "simt_scan_kernel", 2: "tc_filter_kernel", 3: "kl_stream_kernel"}
PARITY_QUERIES = 256


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="radar", choices=["radar", "reference"])
    ap.add_argument("--workload", default="hybrid_10m", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["bf16", "fp32"],
                    help="default: fp32 (certified) unless the workload names bf16")
    ap.add_argument("--algo", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--kl-variant", default="auto", choices=["auto", "bf16x3", "f16x1", "f16x2"],
                    help="filter arithmetic of the KL-only tensor-core paths")
    ap.add_argument("--overfetch", type=int, default=0, help="candidates kept per query by the filters (0 = automatic)")
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--n", type=int, default=0, help="override total corpus rows")
    ap.add_argument("--q", type=int, default=0, help="override queries per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every search eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline workload only")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--secondary", default="", help="comma-separated subset of the secondary workloads")
    ap.add_argument("--secondary-steps", type=int, default=0, help="timed steps per secondary workload (default min(steps,5))")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# algorithmic work per unit (stated in DESIGN.md)
# ---------------------------------------------------------------------------------------------------
def flops_per_pair(mode):
    return {"kl": 28, "dpr": 2 * D, "hybrid": 2 * D + 28}[mode]


def bytes_per_corpus_row(mode):
    # what one pass of the dominant kernel must read per corpus row: bf16 embedding (1024 B) and/or the
    # 14 log-probabilities (56 B algorithmic; stored as a 64 B [hi|lo] bf16 row)
    return {"kl": 56, "dpr": 2 * D, "hybrid": 2 * D + 56}[mode]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except ValueError:
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU leg: the oracle port (BLAS fp32 restatement of the faiss flat scan + KL / hybrid), bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_threads_setup():
    """Use every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently throttle the BLAS
    arm.  Must run before numpy is imported.  Returns the thread count asked for."""
    n = os.cpu_count() or 1
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(n)
    return n


def cpu_threads_report():
    """What the BLAS actually uses (threadpoolctl), next to os.cpu_count()."""
    info = {"os_cpu_count": os.cpu_count(), "cpu_model": cpu_model()}
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count())
        pools = threadpoolctl.threadpool_info()
        info["blas_threads"] = max([p.get("num_threads", 1) for p in pools if p.get("user_api") == "blas"] or [1])
        info["pools"] = [f"{p.get('internal_api')}:{p.get('num_threads')}" for p in pools]
    except Exception as exc:  # threadpoolctl missing: say so
        info["blas_threads"] = None
        info["pools"] = [f"threadpoolctl unavailable: {type(exc).__name__}"]
    return info


def cpu_sample_sizes(wl, n_total):
    if wl["mode"] == "kl":
        return min(n_total, 2_000_000), min(wl["q"], 2048)
    return min(n_total, 400_000), min(wl["q"], 1024)


def run_cpu_sample(mode, k, c_emb, c_logq, q_emb, q_p16, q_ent, n_total, repeats=1):
    """queries/sec of the batched CPU port on (n_s corpus rows, q_s queries), scaled linearly to n_total rows."""
    from oracle import retrieval_oracle as ro
    m = {"dpr": ro.MODE_DPR, "kl": ro.MODE_KL, "hybrid": ro.MODE_HYBRID}[mode]
    n_s = (c_emb if c_emb is not None else c_logq).shape[0]
    q_s = (q_emb if q_emb is not None else q_p16).shape[0]
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        ro.search_blas32(m, k, q_emb=q_emb, c_emb=c_emb, q_p16=q_p16, q_entropy=q_ent, c_logq16=c_logq, alpha=ALPHA)
        best = min(best, time.perf_counter() - t0)
    return (q_s / best) * (n_s / n_total), best, n_s, q_s


def run_cpu_nq1(mode, k, c_emb, c_logq, q_emb, q_p16, q_ent, n_total, n_queries=32, budget_s=8.0):
    """The reference-faithful call pattern (dpr.py:312-314): ONE query per search call, the whole corpus streamed per
    query.  Timed on a small subsample (bounded by `budget_s`), scaled linearly to n_total rows."""
    from oracle import retrieval_oracle as ro
    m = {"dpr": ro.MODE_DPR, "kl": ro.MODE_KL, "hybrid": ro.MODE_HYBRID}[mode]
    n_s = (c_emb if c_emb is not None else c_logq).shape[0]
    q_s = (q_emb if q_emb is not None else q_p16).shape[0]
    done, t0 = 0, time.perf_counter()
    for i in range(min(n_queries, q_s)):
        sl = slice(i, i + 1)
        ro.search_blas32(m, k, q_emb=None if q_emb is None else q_emb[sl], c_emb=c_emb,
                         q_p16=None if q_p16 is None else q_p16[sl], q_entropy=None if q_ent is None else q_ent[sl],
                         c_logq16=c_logq, alpha=ALPHA)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    t = time.perf_counter() - t0
    return {"value": (done / t) * (n_s / n_total), "unit": "queries/s", "queries_timed": done, "seconds": round(t, 2),
            "note": f"one query per call over {n_s} rows (the reference's nq=1 pattern, dpr.py:312-314), scaled by "
                    f"{n_s}/{n_total}"}


def cpu_inputs_from_seed(wl, n_s, q_s):
    """Host-side synthetic sample (used by --impl reference, which never touches a GPU)."""
    from oracle import c_oracle as co
    from radar_multimodal_radiology_b200 import synthetic as syn
    mode = wl["mode"]
    c_emb = q_emb = c_logq = q_p16 = q_ent = None
    if mode != "kl":
        c = syn.embeddings(n_s, D, syn.SEED_CORPUS_EMB)
        q_emb = syn.query_embeddings(q_s, c).numpy()
        c_emb = c.numpy()
    if mode != "dpr":
        c_logq = co.prepare_corpus(syn.observation_probs(n_s, syn.SEED_CORPUS_PROBS).numpy())
        q_p16, q_ent = co.prepare_queries(syn.observation_probs(q_s, syn.SEED_QUERY_PROBS).numpy())
    return c_emb, c_logq, q_emb, q_p16, q_ent


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def main_reference(args, wl, n_total, k):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    asked = cpu_threads_setup()
    import numpy as np  # noqa: F401  (imported after the thread variables are set)
    threads = cpu_threads_report()
    n_s, q_s = cpu_sample_sizes(wl, n_total)
    inputs = cpu_inputs_from_seed(wl, n_s, q_s)
    for _ in range(max(1, min(args.warmup, 1))):
        run_cpu_sample(wl["mode"], k, *inputs, n_total=n_total)
    times = []
    for _ in range(args.steps):
        _, t, _, _ = run_cpu_sample(wl["mode"], k, *inputs, n_total=n_total)
        times.append(t)
    total = sum(times)
    # one "step" of the GPU arm is wl["q"] queries over n_total rows; the CPU step is the bounded sample,
    # scaled linearly in corpus rows (the scan is linear in N) -- stated in `sample`
    qps = (q_s * args.steps / total) * (n_s / n_total)
    nq1 = run_cpu_nq1(wl["mode"], k, *inputs, n_total=n_total)
    sample = (f"{q_s} queries x {n_s} corpus rows per step (numpy/OpenBLAS fp32 batched Q@C.T + argpartition top-{k}, "
              f"{threads['blas_threads']} BLAS threads of {threads['os_cpu_count']} cores), q/s scaled by "
              f"{n_s}/{n_total} to the {n_total}-row corpus; {threads['cpu_model']}")
    line = {
        "impl": "reference", "metric": metric_name(n_total, k), "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.workload, wl, n_total, k, wl.get("precision") or args.precision or "fp32",
                                  "host BLAS calls"),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads["blas_threads"] or asked, "kind": "port",
                         "sample": sample, "threads": threads, "reference_faithful_nq1": nq1},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def metric_name(n_total, k):
    """BASELINE.json quotes "queries/sec at top-k=10 over 10M-case corpus"; other workloads report plain queries/sec."""
    return "queries/sec at top-k=10 over 10M-case corpus" if (n_total == 10_000_000 and k == 10) else "queries/sec"


def workload_config(args, name, wl, n_total, k, precision, launch):
    nq = args.q or wl["q"]
    return {"workload": f"{name}: {wl['mode']} retrieval, {n_total} cases x {D}-d bf16 embeddings + 14-obs "
                        f"log-probs, {nq} queries/step, top-k={k}"
                        + (f", {wl['rounds']} masked rounds/step" if wl.get("rounds") else "")
                        + (f", {wl['corpus']} corpus" if wl.get("corpus") else ""),
            "corpus_rows": n_total, "queries_per_step": nq, "top_k": k, "embedding_dim": D,
            "hybrid_alpha": ALPHA,
            "precision": precision + (" (bf16 tensor-core filter + certificate: result bit-identical to canonical fp32)"
                                      if precision == "fp32" else " (bf16 tensor-core filter, canonical fp32 re-score)"),
            "algo": args.algo,
            "parallelism": f"corpus row-sharded over {args.gpus} GPU(s); one all-gather of packed top-k words + merge",
            "launch": launch,
            "l2_policy": "inputs larger than L2" if n_total * bytes_per_corpus_row(wl["mode"]) / max(1, args.gpus) > 2.6e8
            else "L2 flushed between timed steps (256 MiB write)"}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.args = args
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a GPU: the retrieval kernels have no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        assert self.world == args.gpus or self.world == 1, f"--gpus {args.gpus} but WORLD_SIZE={self.world}"
        from radar_multimodal_radiology_b200 import _lib as L
        L.build()  # no-op when the in-tree .so files are current
        self.L = L
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        self.corpora = {}  # key -> (RadarIndex local shard, n_total)
        self.graphs = []   # every GraphedSearch made (released before the process group goes away)

    # ---- corpora -----------------------------------------------------------------------------------
    def corpus(self, kind, n_total, need_emb, need_probs):
        """Resident local shard of a synthetic corpus (generated on the device, identical for every N)."""
        torch = self.torch
        from radar_multimodal_radiology_b200 import synthetic as syn
        from radar_multimodal_radiology_b200.index import RadarIndex
        from radar_multimodal_radiology_b200.sharded import shard_bounds
        key = (kind, n_total)
        have = self.corpora.get(key)
        if have is not None and (not need_emb or have.emb_f32 is not None) and (not need_probs or have.logq16 is not None):
            return have
        lo, hi = shard_bounds(n_total, self.world, self.rank)
        index = RadarIndex(D, device=self.dev, idx_offset=lo, precision="fp32", algo=self.args.algo,
                           kl_variant=self.args.kl_variant, overfetch=self.args.overfetch)
        index.reserve(hi - lo, embeddings=need_emb, observations=need_probs)
        g = torch.Generator(device=self.dev).manual_seed(4242)
        centers = u = None
        if kind == "clustered":
            centers = torch.nn.functional.normalize(torch.randn((4096, D), generator=g, device=self.dev), dim=-1)
        if kind == "adversarial":
            u = torch.nn.functional.normalize(torch.randn((D,), generator=g, device=self.dev), dim=0)
        pos = (lo // GEN_BLOCK) * GEN_BLOCK
        while pos < hi:
            blk = pos // GEN_BLOCK
            rows = min(GEN_BLOCK, n_total - pos)
            s, e = max(pos, lo), min(pos + rows, hi)
            if need_emb:
                emb = syn.embeddings(rows, D, syn.SEED_CORPUS_EMB + blk, self.dev)
                if kind == "clustered":  # rows = unit vectors at cos ~ 0.89 of one of 4096 centres
                    cid = torch.randint(0, 4096, (rows,), generator=g, device=self.dev)
                    emb = torch.nn.functional.normalize(centers[cid] + 0.5 * emb, dim=-1)
                if kind == "adversarial":
                    # every query of this workload leans towards u; rows ordered by <row, u> ascending make the scores
                    # of ALL queries rise along the sweep (running thresholds are stale for as long as possible)
                    frac = (torch.arange(pos, pos + rows, device=self.dev, dtype=torch.float32) / n_total) - 0.5
                    emb = torch.nn.functional.normalize(emb + (0.6 * frac)[:, None] * u[None, :], dim=-1)
                index.add(emb[s - pos:e - pos])
                del emb
            if need_probs:
                index.add_observations(syn.observation_probs(rows, syn.SEED_CORPUS_PROBS + blk, self.dev)[s - pos:e - pos])
            pos += rows
        index.aux = {"centers": centers, "u": u}
        torch.cuda.synchronize()
        self.corpora[key] = index
        return index

    def drop_corpus(self, kind, n_total):
        self.corpora.pop((kind, n_total), None)
        gc.collect()
        self.torch.cuda.empty_cache()

    def sharded_for(self, wl, n_total):
        """ShardedRadarIndex for a workload: the resident 10M corpus, a row-view of it (377k configs), or a structured
        2M corpus of its own."""
        from radar_multimodal_radiology_b200.sharded import ShardedRadarIndex, shard_bounds
        mode = wl["mode"]
        kind = wl.get("corpus", "iid")
        if kind == "iid" and n_total < 10_000_000 and ("iid", 10_000_000) in self.corpora:
            base = self.corpora[("iid", 10_000_000)]
            lo, hi = shard_bounds(n_total, self.world, self.rank)
            ok = (mode == "kl" or base.emb_f32 is not None) and (mode == "dpr" or base.logq16 is not None)
            if ok and hi - lo <= base.ntotal:
                return ShardedRadarIndex.wrap(base.view_rows(0, hi - lo, idx_offset=lo), n_total)
        local = self.corpus(kind, n_total, mode != "kl", mode != "dpr")
        return ShardedRadarIndex.wrap(local, n_total)

    # ---- one workload -------------------------------------------------------------------------------
    def run(self, name, steps, warmup, headline=False):
        torch, dist, L, args = self.torch, self.dist, self.L, self.args
        from radar_multimodal_radiology_b200 import synthetic as syn
        from radar_multimodal_radiology_b200.index import GraphedSearch
        wl = dict(WORKLOADS[name])
        n_total = (args.n if headline and args.n else wl["n"])
        k = (args.k if headline and args.k else wl["k"])
        nq = (args.q if headline and args.q else wl["q"])
        mode, rounds = wl["mode"], wl.get("rounds", 1)
        precision = wl.get("precision") or args.precision or "fp32"
        dev, world, rank = self.dev, self.world, self.rank
        index = self.sharded_for(wl, n_total)
        ri = index.index
        n_local = ri.ntotal
        skw = dict(alpha=ALPHA, mode=mode, precision=precision)

        # ---- queries (same on every rank) -------------------------------------------------------------
        q_emb = syn.embeddings(nq, D, syn.SEED_QUERY_EMB, dev) if mode != "kl" else None
        if q_emb is not None:
            g = torch.Generator(device=dev).manual_seed(syn.SEED_NEAR)
            kind = wl.get("corpus", "iid")
            if kind == "clustered":   # every query sits in a cluster: ~490 near neighbours with almost equal scores
                cid = torch.randint(0, 4096, (nq,), generator=g, device=dev)
                q_emb = torch.nn.functional.normalize(ri.aux["centers"][cid] + 0.5 * q_emb, dim=-1)
            elif kind == "adversarial":
                q_emb = torch.nn.functional.normalize(q_emb + 0.5 * ri.aux["u"][None, :], dim=-1)
            else:
                # 10 % of the queries sit next to a corpus row; rank 0's rows are broadcast
                n_near = nq // 10
                sel = torch.randperm(nq, generator=g, device=dev)[:n_near]
                src = torch.randint(0, max(1, min(n_local, GEN_BLOCK)), (n_near,), generator=g, device=dev)
                near = torch.nn.functional.normalize(
                    ri.emb_f32[src] + 0.3 * torch.randn((n_near, D), generator=g, device=dev) / D ** 0.5, dim=-1)
                if world > 1:
                    dist.broadcast(near, src=0)
                q_emb[sel] = near
        q_pr = syn.observation_probs(nq, syn.SEED_QUERY_PROBS, dev) if mode != "dpr" else None
        masks = [syn.observation_masks(nq, r, dev) for r in range(rounds)] if wl.get("masked") else [None] * rounds

        # pinned host copies for the end-to-end leg
        pin = lambda t: None if t is None else t.cpu().pin_memory()
        h_q_emb, h_q_pr, h_masks = pin(q_emb), pin(q_pr), [pin(m) for m in masks]
        h_out_s = torch.empty((nq, k), dtype=torch.float32).pin_memory()
        h_out_i = torch.empty((nq, k), dtype=torch.int64).pin_memory()
        h2d_total = rounds * sum(t.numel() * t.element_size() for t in (h_q_emb, h_q_pr) if t is not None) + \
            sum(m.numel() for m in h_masks if m is not None)
        h2d_bytes = -(-h2d_total // world)  # every rank uploads its 1/N slice; the slices travel over NVLink
        d2h_bytes = rounds * (h_out_s.numel() * 4 + h_out_i.numel() * 8)

        need_flush = n_local * bytes_per_corpus_row(mode) <= 2.6e8
        launches_per_step = [0]

        def step_device(collect_stats=False):
            """one step; statistics (which force a stream sync inside the call) are only collected outside the timed region"""
            launches, out = 0, None
            for r in range(rounds):
                out = index.search(q_emb, k, query_probs=q_pr, mask=masks[r], collect_stats=collect_stats, **skw)
                if collect_stats:  # radar_search's kernels + query preparation + the merge kernel after the all-gather
                    launches += ri.last_stats.kernel_launches + (1 if mode != "dpr" else 0) + (1 if world > 1 else 0)
            if collect_stats:
                launches_per_step[0] = launches
            return out

        # ---- the same search chain (collective and merge included) captured once per round in a CUDA graph; eager
        # launches remain the fallback and are what the statistics / kernel-event passes use
        graphed, graph_note = None, "eager"
        if not args.no_graph:
            try:
                graphed = [GraphedSearch(index, q_emb, k, query_probs=q_pr, mask=masks[r], **skw) for r in range(rounds)]
                self.graphs.extend(graphed)
                graph_note = "cuda-graph replay" + (" (NCCL all-gather captured)" if world > 1 else "")
            except Exception as exc:  # capture unsupported on this driver / backend: say so, stay eager
                graphed, graph_note = None, f"eager (graph capture failed: {type(exc).__name__}: {str(exc)[:80]})"
        if world > 1:  # every rank must take the same path
            flag = torch.tensor([1 if graphed is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0 and graphed is not None:
                graphed, graph_note = None, "eager (graph capture failed on another rank)"

        def step_graph():
            out = None
            for gs in graphed:
                out = gs.replay()
            return out

        def step_e2e():
            out = None
            for r in range(rounds):
                if graphed is not None:  # host -> the graph's static input tensors, replay, results -> host
                    if h_q_emb is not None:
                        index.upload_queries(h_q_emb, out=q_emb)
                    if h_q_pr is not None:
                        index.upload_queries(h_q_pr, out=q_pr)
                    if h_masks[r] is not None:
                        index.upload_queries(h_masks[r], out=masks[r])
                    s, i = graphed[r].replay()
                else:
                    xe = None if h_q_emb is None else index.upload_queries(h_q_emb)
                    xp = None if h_q_pr is None else index.upload_queries(h_q_pr)
                    xm = None if h_masks[r] is None else index.upload_queries(h_masks[r])
                    s, i = index.search(xe, k, query_probs=xp, mask=xm, **skw)
                h_out_s.copy_(s, non_blocking=True)
                h_out_i.copy_(i, non_blocking=True)
                out = (s, i)
            return out

        def timed(fn, n_steps, kernel_events=False):
            """sum of per-step CUDA-event times (L2 flush between steps is outside the timed region), max over ranks"""
            total_ms, kern_ms = 0.0, 0.0
            each = []
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            for _ in range(n_steps):
                if need_flush:
                    self.flush_buf.fill_(1)
                if kernel_events:
                    L.check(L.lib().radar_profile_enable(1), "radar_profile_enable")
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                total_ms += e0.elapsed_time(e1)
                each.append(round(e0.elapsed_time(e1), 4))
                if kernel_events:
                    ms = ctypes.c_float()
                    L.check(L.lib().radar_profile_kernel_ms(ctypes.byref(ms)), "radar_profile_kernel_ms")
                    kern_ms += ms.value  # the (last round's) dominant kernel span of this step, CUDA events on its stream
            if kernel_events:
                L.check(L.lib().radar_profile_enable(0), "radar_profile_enable")
            if world > 1:
                t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                total_ms, kern_ms = t.tolist()
            self.last_each = each  # this rank's per-step times of the last timed region
            return total_ms, kern_ms

        # ---- warm-up, then the timed regions -----------------------------------------------------------------
        for _ in range(max(3, warmup)):
            step_device()
        step_device(collect_stats=True)
        torch.cuda.synchronize()
        sampler = None
        if rank == 0 and headline:
            sampler = ClockSampler(self.local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                                   os.environ["CUDA_VISIBLE_DEVICES"].split(",")[self.local_rank])
            sampler.start()
            time.sleep(0.3)
        if graphed is not None:
            for _ in range(2):
                step_graph()
            total_ms, _ = timed(step_graph, steps)
            each_ms = list(self.last_each)
            _, kern_ms = timed(step_device, steps, kernel_events=True)  # dominant kernel alone: eager launches
        else:
            total_ms, kern_ms = timed(step_device, steps, kernel_events=True)
            each_ms = list(self.last_each)
        clocks = sampler.stop() if sampler is not None else None
        out_s, out_i = step_device(collect_stats=True)  # in-kernel clock of a launch made while the device is still under load
        stats = ri.last_stats
        e2e = None
        if not args.no_e2e:
            for _ in range(2):
                step_e2e()
            e2e_ms, _ = timed(step_e2e, steps)
            e2e = {"value": nq * rounds * steps / (e2e_ms / 1e3), "unit": "queries/s",
                   "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                   "ms_per_step": e2e_ms / steps}
            if world > 1:
                e2e["note"] = (f"each rank uploads 1/{world} of the query batch over PCIe and all-gathers the slices "
                               f"over NVLink; h2d_bytes_per_step is per rank")
            # restore the device-resident queries (the e2e leg overwrote them with identical values; keep it exact)
        value = nq * rounds * steps / (total_ms / 1e3)
        kern_ms_avg = kern_ms / steps
        unc_all = [int(stats.uncertified)]
        if world > 1:  # queries re-run by the exact scan on EVERY rank (a rank with many re-runs paces the all-gather)
            t = torch.zeros((world,), dtype=torch.int64, device=dev)
            t[rank] = int(stats.uncertified)
            dist.all_reduce(t)
            unc_all = t.tolist()

        # ---- roofline of the dominant kernel (per launch, this rank's shard) -------------------------------------
        peaks = self.peaks
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tc_burst, tc_sust = peaks.get("bf16_tflops", 1590.0), peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        flops = flops_per_pair(mode) * float(nq) * n_local
        bytes_alg = bytes_per_corpus_row(mode) * float(n_local) + nq * (56 + 2 * D)
        roof = {}
        if bytes_alg / (hbm_peak * 1e9) >= flops / (tc_burst * 1e12):
            ach = bytes_alg / (kern_ms_avg * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "peak_source": peak_src + ", copy bandwidth"}
        else:
            # policy (B200_PROFILING.md): the sustained cuBLAS figure for a kernel that runs long enough to sit under
            # the power cap (>= 50 ms), the burst figure otherwise; BOTH fractions are always reported
            ach = flops / (kern_ms_avg * 1e-3) / 1e12
            long_kernel = kern_ms_avg >= 50.0
            peak = tc_sust if long_kernel else tc_burst
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "frac_of_burst_peak": ach / tc_burst, "frac_of_sustained_peak": ach / tc_sust,
                    "peak_source": peak_src + (", sustained (kernel >= 50 ms)" if long_kernel else ", burst (kernel < 50 ms)")}
        roof["traffic"] = None  # DRAM bytes of this kernel per launch from the committed ncu capture of the same workload
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name if name != "hybrid_10m_bf16" else "hybrid_10m")
            if tr and not (headline and (args.n or args.q or args.k)):
                if world == 1:
                    roof["traffic"] = tr["bytes"]
                    roof["traffic_source"] = tr["source"]
                else:
                    roof["traffic_source"] = f"not captured at N={world} (ncu runs on one GPU; 1-GPU figure: {tr['bytes']})"
        except (OSError, ValueError):
            pass
        roof["kernel"] = KERNEL_NAMES.get(stats.algo_used, str(stats.algo_used))
        if mode == "kl" and stats.algo_used == 2:
            roof["kernel"] = "klf_kernel"  # the dedicated many-queries KL filter (csrc/kl_filter.cuh)
        roof["kernel_ms"] = kern_ms_avg
        roof["kernel_span"] = ("prepass + threshold selection + filter"
                               if (stats.algo_used == 2 and (mode == "kl" or n_local >= (1 << 17))) else "one launch")
        roof["algorithmic_per_launch"] = {"flops": flops, "bytes": bytes_alg}
        if mode == "kl" and stats.algo_used == 2:
            # the many-queries KL filter is bound by looking at the keys, not by producing them: one FMNMX3 lane-op per two
            # keys on 128 lanes per SM and clock is the floor of ONE sweep (the step is a 1/3 prepass + the real pass)
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            clk_hz = (stats.filter_sm_mhz or 1900.0) * 1e6
            floor_ms = float(nq) * n_local / (sms * 128.0 * clk_hz) * 1e3
            roof["select_floor"] = {"ms_per_sweep": floor_ms, "sweeps_per_step": 1.0 + 1.0 / 3.0,
                                    "frac": floor_ms * (1.0 + 1.0 / 3.0) / kern_ms_avg,
                                    "note": "keys / (SMs x 128 lanes x in-kernel clock): the bound this kernel is measured against "
                                            "(DESIGN.md section 4.2); the tensor fraction above counts 28 flop per pair"}
        if mode == "kl" and stats.algo_used == 3:
            roof["streamed_bytes_per_case"] = 64 if precision == "fp32" else 32
            roof["note"] = ("achieved = 56 algorithmic bytes per case / kernel time; the kernel streams the 64-byte [hi|lo] bf16 row "
                            "in certified precision and the 32-byte fp16 row in filter-only precision (same speed: DESIGN.md 4.3)")

        rec = {
            "value": value, "unit": "queries/s", "ms_per_step": total_ms / steps, "steps": steps,
            "ms_each_step_rank0": each_ms,
            "config": workload_config(args, name, wl, n_total, k, precision, graph_note),
            "e2e": e2e, "gpu_launches": launches_per_step[0] * steps, "roofline": roof,
            "search_stats": {"algo_used": stats.algo_used, "parts": stats.parts, "kprime": stats.kprime,
                             "uncertified": stats.uncertified, "uncertified_per_rank": unc_all,
                             "kernel_launches_per_search": stats.kernel_launches,
                             "filter_sm_mhz": round(stats.filter_sm_mhz, 1)},
        }
        if clocks is not None:
            rec["clocks"] = clocks
        if not args.no_parity:
            rec["parity"] = self.parity(index, q_emb, q_pr, masks[rounds - 1], k, skw, out_s, out_i, stats)
        rec["_ctx"] = dict(wl=wl, n_total=n_total, k=k, nq=nq, mode=mode, ri=ri, q_emb=q_emb, q_pr=q_pr, masks=masks,
                           precision=precision)
        return rec

    # ---- parity of what was timed -------------------------------------------------------------------------
    def parity(self, index, q_emb, q_pr, mask, k, skw, out_s, out_i, stats):
        """>= 256 of the step's queries through the exact CUDA-core scan.  N > 1: each rank scans its shard exactly,
        scores and ids are all-gathered SEPARATELY and merged with torch ops (lexicographic sort on (score, id)) -- a
        path that shares neither the packed-word exchange nor the merge kernel with the timed one."""
        torch, dist = self.torch, self.dist
        ri = index.index
        nq = out_s.shape[0]
        m = min(PARITY_QUERIES, nq)
        sel = torch.linspace(0, nq - 1, m, device=self.dev).round().long().unique()
        m = int(sel.numel())
        kw = dict(skw, precision="fp32")
        xe = None if q_emb is None else q_emb[sel].contiguous()
        xp = None if q_pr is None else q_pr[sel].contiguous()
        xm = None if mask is None else mask[sel].contiguous()
        k_local = min(k, ri.ntotal)
        es, ei = ri.search(xe, k_local, query_probs=xp, mask=xm, algo="simt", **kw)
        ascending = skw["mode"] == "kl"
        if self.world > 1:
            pad_s = float("inf") if ascending else float("-inf")
            ls = torch.full((m, k), pad_s, dtype=torch.float32, device=self.dev)
            li = torch.full((m, k), -1, dtype=torch.int64, device=self.dev)
            ls[:, :k_local], li[:, :k_local] = es, ei
            gs = [torch.empty_like(ls) for _ in range(self.world)]
            gi = [torch.empty_like(li) for _ in range(self.world)]
            dist.all_gather(gs, ls)
            dist.all_gather(gi, li)
            cs, ci = torch.cat(gs, dim=1), torch.cat(gi, dim=1)
            ci_key = torch.where(ci < 0, torch.full_like(ci, 1 << 62), ci)
            o1 = torch.argsort(ci_key, dim=1, stable=True)                    # secondary key: smaller id first
            cs, ci = torch.gather(cs, 1, o1), torch.gather(ci, 1, o1)
            o2 = torch.argsort(cs, dim=1, descending=not ascending, stable=True)  # primary key: better score first
            es, ei = torch.gather(cs, 1, o2)[:, :k], torch.gather(ci, 1, o2)[:, :k]
        ts, ti = out_s[sel], out_i[sel]
        torch.cuda.synchronize()
        same_ids = (ti == ei).all(dim=1)
        same_scores = (ts == es).all(dim=1)
        inter = (ti.unsqueeze(2) == ei.unsqueeze(1)).any(dim=2).float().sum(dim=1)
        # every returned score must be the canonical fp32 score of the returned id: where ids agree, scores agree
        score_of_common = ((ti == ei) & (ts != es)).any().item()
        return {"queries_checked": m, "reference": "exact CUDA-core scan (canonical fp32)"
                + ("; per-rank scan, separate all-gathers of scores and ids, torch lexsort merge" if self.world > 1 else ""),
                "recall_at_k": float((inter / k).mean().item()),
                "ids_equal_exact_scan": int(same_ids.sum().item()), "scores_bit_equal": int(same_scores.sum().item()),
                "ids_equal_fp32_certified": bool(same_ids.all().item() and same_scores.all().item())
                if skw["precision"] == "fp32" else None,
                "canonical_score_mismatch_on_common_ids": bool(score_of_common),
                "uncertified_in_step": int(stats.uncertified)}

    def release(self):
        """graphs that captured NCCL collectives must be gone before the communicator is torn down"""
        torch = self.torch
        for g in self.graphs:
            g.graph = None
            g.out = None
        self.graphs.clear()
        gc.collect()
        torch.cuda.synchronize()


def cpu_baseline_record(bench, ctx):
    """The batched BLAS port and the reference-faithful nq=1 loop on a bounded sample of the headline workload
    (rank 0, N = 1 only)."""
    from radar_multimodal_radiology_b200.index import prepare_queries
    wl, n_total, k, mode, ri = ctx["wl"], ctx["n_total"], ctx["k"], ctx["mode"], ctx["ri"]
    threads = cpu_threads_report()
    n_s, q_s = cpu_sample_sizes(wl, n_total)
    n_s = min(n_s, ri.ntotal)
    c_e = None if ri.emb_f32 is None or mode == "kl" else ri.emb_f32[:n_s].cpu().numpy()
    c_l = None if mode == "dpr" else ri.logq16[:n_s].cpu().numpy()
    q_e = None if ctx["q_emb"] is None else ctx["q_emb"][:q_s].cpu().numpy()
    q_p16 = q_ent = None
    if mode != "dpr":
        m0 = ctx["masks"][0]
        a, b = prepare_queries(ctx["q_pr"][:q_s], m0[:q_s] if m0 is not None else None, bench.dev)
        q_p16, q_ent = a.cpu().numpy(), b.cpu().numpy()
    qps, t, n_s, q_s = run_cpu_sample(mode, k, c_e, c_l, q_e, q_p16, q_ent, n_total, repeats=2)
    nq1 = run_cpu_nq1(mode, k, c_e, c_l, q_e, q_p16, q_ent, n_total)
    return {"value": qps, "unit": "queries/s", "cores": threads["blas_threads"] or os.cpu_count(), "kind": "port",
            "sample": f"{q_s} of the step's queries x first {n_s} corpus rows in {t:.2f} s (numpy/OpenBLAS fp32 batched "
                      f"Q@C.T + argpartition top-{k}, {threads['blas_threads']} BLAS threads of "
                      f"{threads['os_cpu_count']} cores), q/s scaled by {n_s}/{n_total}; {threads['cpu_model']}",
            "threads": threads, "reference_faithful_nq1": nq1}


def main():
    args = parse_args()
    wl = dict(WORKLOADS[args.workload])
    n_total = args.n or wl["n"]
    k = args.k or wl["k"]
    if args.impl == "reference":
        return main_reference(args, wl, n_total, k)
    cpu_threads_setup()  # before numpy / torch load their BLAS: the cpu_baseline leg uses every core
    b = Bench(args)
    torch, dist = b.torch, b.dist
    head = b.run(args.workload, args.steps, args.warmup, headline=True)
    ctx = head.pop("_ctx")
    secondary = {}
    if args.workload == "hybrid_10m" and not args.no_secondary and not (args.n or args.q or args.k):
        names = [s for s in (args.secondary.split(",") if args.secondary else SECONDARY) if s]
        sec_steps = args.secondary_steps or max(1, min(args.steps, 5))
        for name in names:
            try:
                rec = b.run(name, sec_steps, 3)
                rec.pop("_ctx")
                secondary[name] = rec
            except Exception as exc:  # a secondary workload must never take the headline down with it
                secondary[name] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
                if b.world > 1:
                    raise
            kind = WORKLOADS[name].get("corpus")
            if kind:
                b.drop_corpus(kind, WORKLOADS[name]["n"])
    cpu = None
    if b.rank == 0 and b.world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_record(b, ctx)
    if b.rank == 0:
        line = {
            "metric": metric_name(ctx["n_total"], ctx["k"]), "value": head["value"], "unit": "queries/s",
            "n_gpus": b.world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (seeded torch.Generator on device; SURVEY.md section 8d distributions)",
            "config": head["config"], "clocks": head.get("clocks"), "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "cpu_baseline": cpu,
            "search_stats": head["search_stats"], "parity": head.get("parity"), "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    b.release()
    if b.world > 1:
        # the JSON line is out; a communicator teardown that stalls must not hold the launcher hostage
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.barrier()
        dist.destroy_process_group()
        watchdog.cancel()


if __name__ == "__main__":
    main()
