#!/usr/bin/env python
"""bench.py -- the retrieval hot path on N GPUs of one node, one JSON line on rank 0.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU path (oracle port) on the host cores, same metric

A "step" is one pass of the hot path over one batch of synthetic queries: query preparation, the fused
score + top-k kernels over this rank's corpus shard, and for N > 1 the all-gather + merge of the per-rank
top-k lists.  Workloads (BASELINE.json `configs`):
    hybrid_10m  (default)  hybrid KL+DPR, 10M-case corpus row-sharded over the N GPUs, 16 384 queries, top-k=10
                           -- the configuration the metric "queries/sec at top-k=10 over 10M-case corpus" is quoted on
    kl_377k / dpr_377k     configs[1] / configs[2]: 377k cases, 65 536 queries, top-k=10
    kl_latency             KL-only, 10M cases, 32 queries, top-k=32 (the HBM-bound regime, BASELINE.md row 4')
    rag_rounds             configs[4]: 3 masked re-retrieval rounds x 16 384 queries, 10M corpus
`value` is timed with inputs resident in HBM; `e2e` times the same call from pinned HOST buffers
(H2D of the queries + D2H of scores/ids inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    #               mode      n_total     queries  k   d    notes
    "hybrid_10m": dict(mode="hybrid", n=10_000_000, q=16384, k=10, masked=False),
    "hybrid_10m_k32": dict(mode="hybrid", n=10_000_000, q=16384, k=32, masked=False),
    "dpr_10m": dict(mode="dpr", n=10_000_000, q=16384, k=10, masked=False),
    "kl_377k": dict(mode="kl", n=377_000, q=65536, k=10, masked=False),
    "dpr_377k": dict(mode="dpr", n=377_000, q=65536, k=10, masked=False),
    "kl_latency": dict(mode="kl", n=10_000_000, q=32, k=32, masked=False),
    "rag_rounds": dict(mode="hybrid", n=10_000_000, q=16384, k=5, masked=True, rounds=3),
    "smoke": dict(mode="hybrid", n=200_000, q=1024, k=10, masked=False),
}
D = 512
ALPHA = 0.5
GEN_BLOCK = 1_250_000  # corpus rows are generated in fixed blocks so the corpus is identical for every N


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="radar", choices=["radar", "reference"])
    ap.add_argument("--workload", default="hybrid_10m", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--algo", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--n", type=int, default=0, help="override total corpus rows")
    ap.add_argument("--q", type=int, default=0, help="override queries per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every search eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-e2e", action="store_true")
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
        # median over the samples taken under load (upper half of the power readings)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU leg: the oracle port (BLAS fp32 restatement of the faiss flat scan + KL / hybrid), bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_sample_sizes(wl, n_total):
    mode = wl["mode"]
    if mode == "kl":
        return min(n_total, 2_000_000), min(wl["q"], 2048)
    return min(n_total, 400_000), min(wl["q"], 1024)


def run_cpu_sample(mode, k, c_emb, c_logq, q_emb, q_p16, q_ent, n_total, repeats=1):
    """queries/sec of the CPU port on (n_s corpus rows, q_s queries), scaled linearly to n_total rows."""
    import numpy as np
    from oracle import retrieval_oracle as ro
    m = {"dpr": ro.MODE_DPR, "kl": ro.MODE_KL, "hybrid": ro.MODE_HYBRID}[mode]
    n_s = (c_emb if c_emb is not None else c_logq).shape[0]
    q_s = (q_emb if q_emb is not None else q_p16).shape[0]
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        ro.search_blas32(m, k, q_emb=q_emb, c_emb=c_emb, q_p16=q_p16, q_entropy=q_ent, c_logq16=c_logq, alpha=ALPHA)
        best = min(best, time.perf_counter() - t0)
    qps_sample = q_s / best
    return qps_sample * (n_s / n_total), best, n_s, q_s


def cpu_inputs_from_seed(wl, n_s, q_s):
    """Host-side synthetic sample (used by --impl reference, which never touches a GPU)."""
    import numpy as np
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
    import numpy as np  # noqa: F401
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
    cores = os.cpu_count()
    sample = (f"{q_s} queries x {n_s} corpus rows per step (numpy/OpenBLAS fp32 Q@C.T + argpartition top-{k}), "
              f"q/s scaled by {n_s}/{n_total} to the {n_total}-row corpus; {cpu_model()}")
    line = {
        "impl": "reference", "metric": metric_name(n_total, k), "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl, n_total, k),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def metric_name(n_total, k):
    """BASELINE.json quotes "queries/sec at top-k=10 over 10M-case corpus"; other workloads report plain queries/sec."""
    return "queries/sec at top-k=10 over 10M-case corpus" if (n_total == 10_000_000 and k == 10) else "queries/sec"


def workload_config(args, wl, n_total, k):
    return {"workload": f"{args.workload}: {wl['mode']} retrieval, {n_total} cases x {D}-d bf16 embeddings + 14-obs "
                        f"log-probs, {args.q or wl['q']} queries/step, top-k={k}"
                        + (f", {wl['rounds']} masked rounds/step" if wl.get("rounds") else ""),
            "corpus_rows": n_total, "queries_per_step": args.q or wl["q"], "top_k": k, "embedding_dim": D,
            "hybrid_alpha": ALPHA, "precision": args.precision, "algo": args.algo,
            "parallelism": f"corpus row-sharded over {args.gpus} GPU(s); all-gather + merge of per-shard top-k",
            "launch": "host BLAS calls",
            "l2_policy": "inputs larger than L2" if n_total * bytes_per_corpus_row(wl["mode"]) / max(1, args.gpus) > 2.6e8
            else "L2 flushed between timed steps (256 MiB write)"}


def main():
    args = parse_args()
    wl = dict(WORKLOADS[args.workload])
    n_total = args.n or wl["n"]
    k = args.k or wl["k"]
    nq = args.q or wl["q"]
    mode = wl["mode"]
    rounds = wl.get("rounds", 1)
    if args.impl == "reference":
        return main_reference(args, wl, n_total, k)

    import numpy as np
    import torch
    import torch.distributed as dist
    from radar_multimodal_radiology_b200 import _lib as L
    from radar_multimodal_radiology_b200 import synthetic as syn
    from radar_multimodal_radiology_b200.index import prepare_queries
    from radar_multimodal_radiology_b200.sharded import ShardedRadarIndex, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the retrieval kernels have no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    L.build()  # no-op when the in-tree .so is current

    # ---- corpus shard, generated on the device in fixed blocks (identical corpus for every N) ----------
    lo, hi = shard_bounds(n_total, world, rank)
    n_local = hi - lo
    c_emb = torch.empty((n_local, D), dtype=torch.float32, device=dev) if mode != "kl" else None
    c_pr = torch.empty((n_local, 14), dtype=torch.float32, device=dev) if mode != "dpr" else None
    b0 = lo // GEN_BLOCK
    pos = b0 * GEN_BLOCK
    while pos < hi:
        blk = pos // GEN_BLOCK
        rows = min(GEN_BLOCK, n_total - pos)
        s, e = max(pos, lo), min(pos + rows, hi)
        if c_emb is not None:
            c_emb[s - lo:e - lo] = syn.embeddings(rows, D, syn.SEED_CORPUS_EMB + blk, dev)[s - pos:e - pos]
        if c_pr is not None:
            c_pr[s - lo:e - lo] = syn.observation_probs(rows, syn.SEED_CORPUS_PROBS + blk, dev)[s - pos:e - pos]
        pos += rows
    index = ShardedRadarIndex(D, device=dev, precision=args.precision, algo=args.algo).build(n_total, c_emb, c_pr)
    ri = index.index
    del c_pr
    torch.cuda.synchronize()

    # ---- queries (same on every rank) ---------------------------------------------------------------------
    q_emb = syn.embeddings(nq, D, syn.SEED_QUERY_EMB, dev) if mode != "kl" else None
    if q_emb is not None and ri.emb_f32 is not None and rank == 0:
        pass
    if q_emb is not None:
        # 10 % of the queries sit next to a corpus row; rank 0's rows are broadcast
        n_near = nq // 10
        g = torch.Generator(device=dev).manual_seed(syn.SEED_NEAR)
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
    h2d_bytes = rounds * sum(t.numel() * t.element_size() for t in (h_q_emb, h_q_pr) if t is not None) + \
        sum(m.numel() for m in h_masks if m is not None)
    d2h_bytes = rounds * (h_out_s.numel() * 4 + h_out_i.numel() * 8)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    need_flush = n_local * bytes_per_corpus_row(mode) <= 2.6e8
    launches_per_step = [0]

    def step_device(collect_stats=False):
        """one step; statistics (which force a stream sync inside the call) are only collected outside the timed region"""
        launches = 0
        out = None
        for r in range(rounds):
            out = index.search(q_emb, k, query_probs=q_pr, mask=masks[r], alpha=ALPHA, mode=mode,
                               collect_stats=collect_stats)
            if collect_stats:  # radar_search's own kernels + query preparation + the merge kernel after the all-gather
                launches += ri.last_stats.kernel_launches + (1 if mode != "dpr" else 0) + (1 if world > 1 else 0)
        if collect_stats:
            launches_per_step[0] = launches
        return out

    # ---- the same search chain captured once per round in a CUDA graph (GraphedSearch); eager launches remain the
    # fallback and are what the statistics / kernel-event passes use
    graphed, graph_note = None, "eager"
    if world > 1:
        # a graph that captured NCCL collectives keeps the communicator busy at teardown (destroy_process_group hung
        # for minutes on this stack), and with >= 15 ms of kernel per step the launch chain is hidden anyway
        graph_note = "eager (collectives are not captured in a graph)"
    elif not args.no_graph:
        try:
            from radar_multimodal_radiology_b200.index import GraphedSearch
            graphed = [GraphedSearch(index, q_emb, k, query_probs=q_pr, mask=masks[r], alpha=ALPHA, mode=mode)
                       for r in range(rounds)]
            graph_note = "cuda-graph replay"
        except Exception as exc:  # capture unsupported on this driver / backend: say so, stay eager
            graphed, graph_note = None, f"eager (graph capture failed: {type(exc).__name__})"

    def step_graph():
        out = None
        for g in graphed:
            out = g.replay()
        return out

    def step_e2e():
        out = None
        for r in range(rounds):
            if graphed is not None:  # host -> the graph's static input tensors, replay, results -> host
                if h_q_emb is not None:
                    q_emb.copy_(h_q_emb, non_blocking=True)
                if h_q_pr is not None:
                    q_pr.copy_(h_q_pr, non_blocking=True)
                if h_masks[r] is not None:
                    masks[r].copy_(h_masks[r], non_blocking=True)
                s, i = graphed[r].replay()
            else:
                xe = None if h_q_emb is None else h_q_emb.to(dev, non_blocking=True)
                xp = None if h_q_pr is None else h_q_pr.to(dev, non_blocking=True)
                xm = None if h_masks[r] is None else h_masks[r].to(dev, non_blocking=True)
                s, i = index.search(xe, k, query_probs=xp, mask=xm, alpha=ALPHA, mode=mode)
            h_out_s.copy_(s, non_blocking=True)
            h_out_i.copy_(i, non_blocking=True)
            out = (s, i)
        return out

    def timed(fn, steps, kernel_events=False):
        """sum of per-step CUDA-event times (L2 flush between steps is outside the timed region), max over ranks"""
        total_ms, kern_ms = 0.0, 0.0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for _ in range(steps):
            if need_flush:
                flush_buf.fill_(1)
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
            if kernel_events:
                ms = ctypes.c_float()
                L.check(L.lib().radar_profile_kernel_ms(ctypes.byref(ms)), "radar_profile_kernel_ms")
                kern_ms += ms.value  # the (last round's) dominant kernel of this step, CUDA events on its stream
        if world > 1:
            t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms, kern_ms = t.tolist()
        return total_ms, kern_ms

    # ---- warm-up, then the timed regions ---------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_device()
    step_device(collect_stats=True)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank])
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if graphed is not None:
        for _ in range(2):
            step_graph()
        total_ms, _ = timed(step_graph, args.steps)
        _, kern_ms = timed(step_device, args.steps, kernel_events=True)  # dominant kernel alone: eager launches
    else:
        total_ms, kern_ms = timed(step_device, args.steps, kernel_events=True)
    clocks = sampler.stop() if rank == 0 else None
    step_device(collect_stats=True)  # in-kernel clock of a launch made while the device is still under load
    stats = ri.last_stats
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        e2e_ms, _ = timed(step_e2e, args.steps)
        e2e = {"value": nq * rounds * args.steps / (e2e_ms / 1e3), "unit": "queries/s",
               "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
               "ms_per_step": e2e_ms / args.steps}

    value = nq * rounds * args.steps / (total_ms / 1e3)
    kern_ms_avg = kern_ms / args.steps
    # ---- roofline of the dominant kernel (per launch, this rank's shard) -----------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tc_peak = peaks.get("bf16_tflops_sustained" if kern_ms_avg > 50 else "bf16_tflops", 1590.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    flops = flops_per_pair(mode) * float(nq) * n_local
    bytes_alg = bytes_per_corpus_row(mode) * float(n_local) + nq * (56 + 2 * D)
    t_flops = flops / (tc_peak * 1e12)
    t_bytes = bytes_alg / (hbm_peak * 1e9)
    if t_bytes >= t_flops:
        roof = {"bound": "hbm", "achieved": bytes_alg / (kern_ms_avg * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": flops / (kern_ms_avg * 1e-3) / 1e12, "peak": tc_peak, "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = None  # DRAM bytes of this kernel per launch from the committed ncu capture of the same workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        if tr and world == 1 and not (args.n or args.q or args.k):
            roof["traffic"] = tr["bytes"]
            roof["traffic_source"] = tr["source"]
    except (OSError, ValueError):
        pass
    roof["kernel"] = "tc_filter_kernel" if stats.algo_used == 2 else "simt_scan_kernel"
    roof["kernel_ms"] = kern_ms_avg
    roof["peak_source"] = peak_src + (", sustained" if kern_ms_avg > 50 and roof["bound"] == "tensor" else ", burst")
    roof["algorithmic_per_launch"] = {"flops": flops, "bytes": bytes_alg}

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N = 1 only) ------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s, q_s = cpu_sample_sizes(wl, n_total)
        n_s = min(n_s, n_local)
        c_e = None if ri.emb_f32 is None or mode == "kl" else ri.emb_f32[:n_s].cpu().numpy()
        c_l = None if mode == "dpr" else ri.logq16[:n_s].cpu().numpy()
        q_e = None if q_emb is None else q_emb[:q_s].cpu().numpy()
        q_p16 = q_ent = None
        if mode != "dpr":
            a, b = prepare_queries(q_pr[:q_s], masks[0][:q_s] if masks[0] is not None else None, dev)
            q_p16, q_ent = a.cpu().numpy(), b.cpu().numpy()
        qps, t, n_s, q_s = run_cpu_sample(mode, k, c_e, c_l, q_e, q_p16, q_ent, n_total, repeats=2)
        cpu = {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{q_s} of the step's queries x first {n_s} corpus rows in {t:.2f} s (numpy/OpenBLAS fp32 "
                         f"Q@C.T + argpartition top-{k}), q/s scaled by {n_s}/{n_total}; {cpu_model()}"}

    if rank == 0:
        line = {
            "metric": metric_name(n_total, k), "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic (seeded torch.Generator on device; SURVEY.md section 8d distributions)",
            "config": dict(workload_config(args, wl, n_total, k), launch=graph_note),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step[0] * args.steps,
            "roofline": roof, "cpu_baseline": cpu,
            "search_stats": {"algo_used": stats.algo_used, "parts": stats.parts, "kprime": stats.kprime,
                             "uncertified": stats.uncertified, "kernel_launches_per_search": stats.kernel_launches,
                             "filter_sm_mhz": round(stats.filter_sm_mhz, 1)},
        }
        print(json.dumps(line), flush=True)
    graphed = None
    torch.cuda.synchronize()
    if world > 1:
        # the JSON line is out; a communicator teardown that stalls must not hold the launcher hostage
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.barrier()
        dist.destroy_process_group()
        watchdog.cancel()


if __name__ == "__main__":
    main()
