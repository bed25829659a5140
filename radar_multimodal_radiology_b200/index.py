"""GPU flat index: the operator that replaces ``faiss.IndexFlatIP`` underneath RADAR's retrieval API.

Reference call sites (annotate_retrieve/modeling_dense_passage_retrieval.py):
    ``faiss.IndexFlatIP(d)`` :297, ``index.add(float32[N,d])`` :298, ``index.ntotal`` :300,
    ``index.search(float32[nq,d], k) -> (D float32[nq,k] descending, I int64[nq,k])`` :313,
    truthiness of the index object :310.
``RadarIndex`` keeps those four members (``GpuIndexFlatIP`` is the numpy-in / numpy-out spelling a faiss
user expects) and extends ``search`` with the pieces north_star names: per-query observation
probabilities (KL), observation masks, and the ``hybrid_alpha`` fusion -- all executed by the sm_100a
kernels in ``csrc/`` through the C ABI of ``include/radar_retrieval.h``.  PyTorch owns every buffer.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib as L
from ._nvtx import annotate as _nvtx_annotate, range_ as _nvtx_range

ArrayLike = Union[np.ndarray, torch.Tensor]


@dataclass
class SearchStats:
    algo_used: int = 0
    kernel_launches: int = 0
    uncertified: int = 0
    parts: int = 0
    kprime: int = 0
    filter_sm_mhz: float = 0.0


def _as_device_f32(x: ArrayLike, device: torch.device) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x, dtype=torch.float32)
    return x.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def prepare_queries(probs: ArrayLike, mask: Optional[ArrayLike], device, eps: float = 1e-8,
                    normalize: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """(p16 float32[Q,16], entropy float32[Q]) on ``device`` -- ``radar_kl_prepare_queries``."""
    device = torch.device(device)
    p = _as_device_f32(probs, device)
    if p.dim() != 2 or p.shape[1] != L.NUM_OBS:
        raise ValueError(f"query_probs must be [Q,{L.NUM_OBS}], got {tuple(p.shape)}")
    m = None
    if mask is not None:
        if isinstance(mask, np.ndarray):
            mask = torch.from_numpy(np.ascontiguousarray(mask))
        m = mask.to(device=device).ne(0).to(torch.uint8).contiguous()
        if tuple(m.shape) != tuple(p.shape):
            raise ValueError(f"mask must be {tuple(p.shape)}, got {tuple(m.shape)}")
    q = p.shape[0]
    p16 = torch.empty((q, L.OBS_PAD), dtype=torch.float32, device=device)
    ent = torch.empty((q,), dtype=torch.float32, device=device)
    with L.device_guard(device):
        L.check(L.lib().radar_kl_prepare_queries(L.ptr(p), L.ptr(m), q, L.NUM_OBS, eps, int(normalize), L.ptr(p16),
                                                 L.ptr(ent), L.current_stream_ptr(device)),
                "radar_kl_prepare_queries")
    return p16, ent


class Workspace:
    """Growable device scratch buffer handed to ``radar_search`` (the library allocates nothing itself)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.buf: Optional[torch.Tensor] = None
        self.frozen = False  # set by GraphedSearch after capture: the buffer must never move again

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes + 256:
            if self.frozen:
                raise RuntimeError("a captured search needs a larger workspace than it was captured with")
            self.buf = torch.empty((nbytes + 256,), dtype=torch.uint8, device=self.device)
        return self.buf


class _RowStore:
    """Append-only [rows, cols] device tensor with geometric capacity growth.

    ``faiss.IndexFlatIP.add`` copies what it is given and amortises growth; a ``torch.cat`` per ``add`` call would
    re-copy the whole corpus every time (quadratic traffic, 2x transient HBM) and storing the caller's tensor by
    reference would let later writes to it desynchronise the fp32 / bf16 copies of the index."""

    def __init__(self, cols: int, dtype: torch.dtype, device: torch.device):
        self.cols, self.dtype, self.device = cols, dtype, device
        self.buf: Optional[torch.Tensor] = None
        self.rows = 0

    def reserve(self, rows: int) -> None:
        if self.buf is not None and self.buf.shape[0] >= rows:
            return
        cap = rows if self.buf is None else max(rows, int(self.buf.shape[0] * 1.5) + 1)
        new = torch.empty((cap, self.cols), dtype=self.dtype, device=self.device)
        if self.rows:
            new[:self.rows].copy_(self.buf[:self.rows])
        self.buf = new

    def append_slot(self, n: int) -> torch.Tensor:
        """Uninitialised view of the next ``n`` rows (the caller fills it)."""
        self.reserve(self.rows + n)
        view = self.buf[self.rows:self.rows + n]
        self.rows += n
        return view

    def view(self) -> Optional[torch.Tensor]:
        return None if self.buf is None else self.buf[:self.rows]

    @classmethod
    def adopt(cls, t: torch.Tensor) -> "_RowStore":
        st = cls(t.shape[1], t.dtype, t.device)
        st.buf, st.rows = t.contiguous(), t.shape[0]
        return st


_TABLES = ("emb_f32", "emb_bf16", "logq16", "klpack", "kl16")  # resident corpus tensors (all optional)


class RadarIndex:
    """Exact flat index over (embedding, observation-probability) rows resident in HBM.

    ``add`` / ``search`` / ``ntotal`` follow ``faiss.IndexFlatIP``; ``search`` additionally takes
    ``query_probs``, ``mask``, ``alpha`` and ``mode``.
    """

    def __init__(self, d: int = 512, device: Union[str, torch.device] = "cuda", precision: str = "bf16",
                 eps: float = 1e-8, normalize: bool = False, idx_offset: int = 0, algo: str = "auto",
                 overfetch: int = 0, num_sms: int = 0, keep_bf16: bool = True, kl_variant: str = "auto"):
        self.d = int(d)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RadarIndex runs on a CUDA device only (there is no CPU fallback)")
        if precision not in L.PREC_BY_NAME:
            raise ValueError(f"precision must be one of {sorted(L.PREC_BY_NAME)}")
        if algo not in L.ALGO_BY_NAME:
            raise ValueError(f"algo must be one of {sorted(L.ALGO_BY_NAME)}")
        self.precision, self.algo = precision, algo
        self.eps, self.normalize = float(eps), bool(normalize)
        self.idx_offset = int(idx_offset)
        self.overfetch, self.num_sms = int(overfetch), int(num_sms)
        self.keep_bf16 = keep_bf16
        if kl_variant not in L.KL_VARIANT_BY_NAME:
            raise ValueError(f"kl_variant must be one of {sorted(L.KL_VARIANT_BY_NAME)}")
        self.kl_variant = kl_variant
        self._stores = {}  # name -> _RowStore (emb_f32, emb_bf16, logq16, klpack)
        self.emb_max_norm = 0.0
        self.logq_col_max = [0.0] * L.OBS_PAD  # per observation: max |log q| over the rows added so far
        self._ws = Workspace(self.device)
        # bumped whenever a buffer a captured CUDA graph may point into is replaced (add / reset / load / workspace
        # growth); GraphedSearch refuses to replay across a change
        self.generation = 0
        self.last_stats = SearchStats()
        L.lib()  # fail loudly now if the extension is missing

    # resident tensors: views of the append-only stores (what the C ABI is handed)
    def _view(self, name: str) -> Optional[torch.Tensor]:
        st = self._stores.get(name)
        return None if st is None else st.view()

    def _set(self, name: str, t: Optional[torch.Tensor]) -> None:
        if t is None:
            self._stores.pop(name, None)
        else:
            self._stores[name] = _RowStore.adopt(t)
        self.generation += 1

    emb_f32 = property(lambda self: self._view("emb_f32"), lambda self, t: self._set("emb_f32", t))
    emb_bf16 = property(lambda self: self._view("emb_bf16"), lambda self, t: self._set("emb_bf16", t))
    logq16 = property(lambda self: self._view("logq16"), lambda self, t: self._set("logq16", t))
    klpack = property(lambda self: self._view("klpack"), lambda self, t: self._set("klpack", t))
    kl16 = property(lambda self: self._view("kl16"), lambda self, t: self._set("kl16", t))

    def _store(self, name: str, cols: int, dtype: torch.dtype) -> _RowStore:
        st = self._stores.get(name)
        if st is None:
            st = self._stores[name] = _RowStore(cols, dtype, self.device)
        return st

    def reserve(self, rows: int, embeddings: bool = True, observations: bool = True) -> None:
        """Pre-size the resident tensors for ``rows`` rows (optional; avoids growth copies during a bulk build)."""
        if embeddings:
            self._store("emb_f32", self.d, torch.float32).reserve(rows)
            if self.keep_bf16:
                self._store("emb_bf16", self.d, torch.bfloat16).reserve(rows)
        if observations:
            self._store("logq16", L.OBS_PAD, torch.float32).reserve(rows)
            self._store("klpack", L.KLPACK, torch.bfloat16).reserve(rows)
            self._store("kl16", L.OBS_PAD, torch.float16).reserve(rows)
        self.generation += 1

    # ---- faiss.IndexFlatIP surface ------------------------------------------------------------------
    def __bool__(self) -> bool:  # dpr.py:310 tests truthiness of the index object
        return True

    @property
    def ntotal(self) -> int:
        if self.emb_f32 is not None:
            return int(self.emb_f32.shape[0])
        if self.logq16 is not None:
            return int(self.logq16.shape[0])
        return 0

    @_nvtx_annotate("index.add")
    def add(self, x: ArrayLike) -> None:
        """Append embedding rows (float32[N,d]); ``faiss.IndexFlatIP.add`` (dpr.py:298).  Like faiss, the rows are
        COPIED into the index (the caller may reuse its buffer)."""
        x = _as_device_f32(x, self.device)
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError(f"expected [N,{self.d}] embeddings, got {tuple(x.shape)}")
        if self.d % 4 != 0:
            raise ValueError("embedding dim must be a multiple of 4")
        n = x.shape[0]
        if n == 0:
            return
        dst = self._store("emb_f32", self.d, torch.float32).append_slot(n)
        dst.copy_(x)
        bf16 = self._store("emb_bf16", self.d, torch.bfloat16).append_slot(n) if self.keep_bf16 else None
        mx = torch.zeros((1,), dtype=torch.float32, device=self.device)
        with L.device_guard(self.device):
            L.check(L.lib().radar_pack_embeddings(L.ptr(dst), n, self.d, L.ptr(bf16), L.ptr(mx),
                                                  L.current_stream_ptr(self.device)), "radar_pack_embeddings")
        self.emb_max_norm = max(self.emb_max_norm, float(mx.item()))
        self.generation += 1

    @_nvtx_annotate("index.add_observations")
    def add_observations(self, probs: ArrayLike) -> None:
        """Append observation-probability rows (float32[N,14], CheXpert-14 order) -- K1 corpus side."""
        p = _as_device_f32(probs, self.device)
        if p.dim() != 2 or p.shape[1] != L.NUM_OBS:
            raise ValueError(f"expected [N,{L.NUM_OBS}] probabilities, got {tuple(p.shape)}")
        n = p.shape[0]
        if n == 0:
            return
        logq = self._store("logq16", L.OBS_PAD, torch.float32).append_slot(n)
        pack = self._store("klpack", L.KLPACK, torch.bfloat16).append_slot(n)
        kl16 = self._store("kl16", L.OBS_PAD, torch.float16).append_slot(n)
        with L.device_guard(self.device):
            L.check(L.lib().radar_kl_prepare_corpus(L.ptr(p), n, L.NUM_OBS, self.eps, int(self.normalize),
                                                    L.ptr(logq), L.ptr(pack), L.ptr(kl16),
                                                    L.current_stream_ptr(self.device)),
                    "radar_kl_prepare_corpus")
        col = logq.abs().amax(dim=0).tolist()  # one sync per add; tightens the filter's error bound
        self.logq_col_max = [max(a, float(b)) for a, b in zip(self.logq_col_max, col)]
        self.generation += 1

    def reset(self) -> None:
        self._stores = {}
        self.emb_max_norm = 0.0
        self.logq_col_max = [0.0] * L.OBS_PAD
        self.generation += 1

    # ---- persistence (the reference rebuilds its index in RAM on every run) ----------
    @_nvtx_annotate("index.save")
    def save(self, path: str) -> None:
        """Write this (shard of the) index as ``<path>.safetensors`` + ``<path>.json``.

        The reference persists model weights with safetensors (train_expert_models.py:279-283) and never persists the
        retrieval index (dpr.py:278-303); here the resident tensors are stored as they sit in HBM -- canonical fp32
        embeddings, their bf16 copy, the fp32 log table, its bf16 [hi|lo] pack and its fp16 copy -- so loading is a plain
        copy."""
        import json
        from safetensors.torch import save_file
        tensors = {}
        for name in _TABLES:
            t = getattr(self, name)
            if t is not None:
                tensors[name] = t.detach().cpu().contiguous()
        save_file(tensors, path + ".safetensors")
        meta = dict(format="radar-index-v1", d=self.d, ntotal=self.ntotal, eps=self.eps, normalize=self.normalize,
                    idx_offset=self.idx_offset, emb_max_norm=self.emb_max_norm, precision=self.precision,
                    logq_col_max=self.logq_col_max, tensors=sorted(tensors))
        with open(path + ".json", "w") as fh:
            json.dump(meta, fh, indent=1)

    @classmethod
    def load(cls, path: str, device: Union[str, torch.device] = "cuda", **kw) -> "RadarIndex":
        """Inverse of :meth:`save`; ``kw`` overrides constructor arguments such as ``precision`` / ``algo``."""
        import json
        from safetensors.torch import load_file
        with open(path + ".json") as fh:
            meta = json.load(fh)
        if meta.get("format") != "radar-index-v1":
            raise ValueError(f"{path}.json is not a radar index (format={meta.get('format')!r})")
        args = dict(d=meta["d"], device=device, precision=meta["precision"], eps=meta["eps"],
                    normalize=meta["normalize"], idx_offset=meta["idx_offset"])
        args.update(kw)
        index = cls(**args)
        tensors = load_file(path + ".safetensors", device=str(index.device))
        for name in _TABLES:
            if name in tensors:
                setattr(index, name, tensors[name].contiguous())
        index.emb_max_norm = float(meta["emb_max_norm"])
        index.logq_col_max = [float(v) for v in meta.get("logq_col_max", [0.0] * L.OBS_PAD)]
        if index.ntotal != meta["ntotal"]:
            raise ValueError(f"{path}: expected {meta['ntotal']} rows, found {index.ntotal}")
        return index

    def view_rows(self, lo: int, hi: int, idx_offset: Optional[int] = None, **kw) -> "RadarIndex":
        """A second index over rows ``[lo, hi)`` of this one WITHOUT copying them (the views share storage; the
        corpus-wide norm / log-table maxima are inherited, which keeps every error bound valid).  Returned ids are
        ``idx_offset + local row`` (default: this index's offset + ``lo``).  ``kw`` overrides constructor arguments."""
        if not (0 <= lo <= hi <= self.ntotal):
            raise ValueError(f"rows [{lo},{hi}) out of range for an index of {self.ntotal} rows")
        args = dict(d=self.d, device=self.device, precision=self.precision, eps=self.eps, normalize=self.normalize,
                    idx_offset=self.idx_offset + lo if idx_offset is None else idx_offset, algo=self.algo,
                    overfetch=self.overfetch, num_sms=self.num_sms, keep_bf16=self.keep_bf16,
                    kl_variant=self.kl_variant)
        args.update(kw)
        v = RadarIndex(**args)
        for name in _TABLES:
            t = self._view(name)
            if t is not None:
                setattr(v, name, t[lo:hi])
        v.emb_max_norm, v.logq_col_max = self.emb_max_norm, list(self.logq_col_max)
        return v

    # ---- search -------------------------------------------------------------------------------------
    def _corpus_struct(self, mode: int) -> L.CorpusStruct:
        c = L.CorpusStruct()
        c.n = self.ntotal
        c.d = self.d
        if mode != L.MODE_KL:
            if self.emb_f32 is None:
                raise RuntimeError("index holds no embeddings (call add) but mode needs them")
            c.emb_f32 = L.ptr(self.emb_f32)
            c.emb_bf16 = L.ptr(self.emb_bf16)
        if mode != L.MODE_DPR:
            if self.logq16 is None:
                raise RuntimeError("index holds no observation probabilities (call add_observations)")
            if self.emb_f32 is not None and self.logq16.shape[0] != self.emb_f32.shape[0]:
                raise RuntimeError("embedding rows and observation rows differ in count")
            c.logq16 = L.ptr(self.logq16)
            c.klpack = L.ptr(self.klpack)
            c.kl16 = L.ptr(self.kl16)
        c.emb_max_norm = self.emb_max_norm
        c.logq_max_abs = abs(math.log(self.eps)) * 1.0001
        c.idx_offset = self.idx_offset
        for j in range(L.OBS_PAD):
            c.logq_col_max[j] = self.logq_col_max[j] if mode != L.MODE_DPR else 0.0
        return c

    def _get_workspace(self, nbytes: int, holder: Optional["Workspace"] = None) -> torch.Tensor:
        """Scratch buffer of one search call.  Eager searches share the index's own holder (one search at a time per
        index and stream); a GraphedSearch -- or a caller searching the same index from several streams -- passes a
        private :class:`Workspace`, so nothing a captured graph points into is ever freed or shared."""
        return (holder or self._ws).get(nbytes)

    def resolve_mode(self, mode: Optional[str], have_emb: bool, have_probs: bool) -> int:
        if mode is None:
            if have_emb and have_probs and self.logq16 is not None and self.emb_f32 is not None:
                return L.MODE_HYBRID
            if have_emb and self.emb_f32 is not None:
                return L.MODE_DPR
            if have_probs:
                return L.MODE_KL
            raise ValueError("search needs query embeddings and/or query_probs")
        if mode not in L.MODE_BY_NAME:
            raise ValueError(f"mode must be one of {sorted(L.MODE_BY_NAME)}")
        return L.MODE_BY_NAME[mode]

    @_nvtx_annotate("index.search")
    def search(self, x: Optional[ArrayLike], k: int, query_probs: Optional[ArrayLike] = None,
               mask: Optional[ArrayLike] = None, alpha: float = 0.5, mode: Optional[str] = None,
               precision: Optional[str] = None, algo: Optional[str] = None, collect_stats: bool = False,
               prepared: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
               workspace: Optional[Workspace] = None, return_packed: bool = False,
               kl_variant: Optional[str] = None, overfetch: Optional[int] = None):
        """(scores float32[Q,k], ids int64[Q,k]) as CUDA tensors, best first.

        DPR: inner products, descending (``IndexFlatIP.search``).  KL: KL(p_query||q_case), ascending.
        hybrid: ``alpha*ip - (1-alpha)*KL``, descending.  Ties: smaller id first.
        ``prepared`` = (p16, entropy) from :func:`prepare_queries` skips the query-side preparation.
        ``k`` may exceed ``RADAR_MAX_K`` (faiss accepts any k, dpr.py:313): the ranking is then paged through in
        exact ``RADAR_MAX_K``-sized calls (``radar_queries.after_*``).  ``return_packed`` additionally returns the
        result as sortable uint64 words (int64 tensor of bit patterns) -- the form row-sharded ranks exchange.
        ``kl_variant`` / ``overfetch`` override the index-wide settings for this call (filter arithmetic of the KL-only
        tensor-core paths; candidates kept per query by a filter).
        """
        m = self.resolve_mode(mode, x is not None, query_probs is not None or prepared is not None)
        n = self.ntotal
        if n == 0:
            raise RuntimeError("search on an empty index")
        if k < 1:
            raise ValueError(f"k must be >= 1, got {k}")
        if k > n:
            raise ValueError(f"k={k} exceeds ntotal={n}; clamp k in the caller (dpr.py:308)")
        xq = p16 = ent = None
        if m != L.MODE_KL:
            if x is None:
                raise ValueError("query embeddings are required for dpr/hybrid")
            xq = _as_device_f32(x, self.device)
            if xq.dim() == 1:
                xq = xq.unsqueeze(0)
            if xq.shape[1] != self.d:
                raise ValueError(f"query embeddings must be [Q,{self.d}], got {tuple(xq.shape)}")
        if m != L.MODE_DPR:
            if prepared is not None:
                p16, ent = prepared
            else:
                if query_probs is None:
                    raise ValueError("query_probs are required for kl/hybrid")
                p16, ent = prepare_queries(query_probs, mask, self.device, self.eps, self.normalize)
            if xq is not None and p16.shape[0] != xq.shape[0]:
                raise ValueError("query embeddings and query_probs differ in row count")
        q = int((xq if xq is not None else p16).shape[0])
        out_s = torch.empty((q, k), dtype=torch.float32, device=self.device)
        out_i = torch.empty((q, k), dtype=torch.int64, device=self.device)
        out_p = torch.empty((q, k), dtype=torch.int64, device=self.device) if return_packed else None
        if q == 0:
            return (out_s, out_i, out_p) if return_packed else (out_s, out_i)
        prec = L.PREC_BY_NAME[precision or self.precision]
        alg = L.ALGO_BY_NAME[algo or self.algo]
        tune = (L.KL_VARIANT_BY_NAME[kl_variant or self.kl_variant], self.overfetch if overfetch is None else int(overfetch))
        if k <= L.MAX_K:
            self._search_call(m, xq, p16, ent, k, alpha, prec, alg, out_s, out_i, out_p, None, collect_stats, workspace,
                              tune=tune)
        else:
            # page through the exact ranking, RADAR_MAX_K results at a time, each call continuing strictly after the
            # last (score, id) of the previous one
            if alg not in (L.ALGO_AUTO, L.ALGO_SIMT_EXACT):
                raise ValueError(f"k={k} > {L.MAX_K} is served by the exact scan only (algo 'auto' or 'simt')")
            done, after = 0, None
            while done < k:
                kk = min(L.MAX_K, k - done)
                ps = torch.empty((q, kk), dtype=torch.float32, device=self.device)
                pi = torch.empty((q, kk), dtype=torch.int64, device=self.device)
                pp = torch.empty((q, kk), dtype=torch.int64, device=self.device) if return_packed else None
                self._search_call(m, xq, p16, ent, kk, alpha, L.PREC_FP32, L.ALGO_SIMT_EXACT, ps, pi, pp, after,
                                  collect_stats, workspace)
                out_s[:, done:done + kk], out_i[:, done:done + kk] = ps, pi
                if return_packed:
                    out_p[:, done:done + kk] = pp
                after = (ps[:, -1].contiguous(), pi[:, -1].contiguous())
                done += kk
        return (out_s, out_i, out_p) if return_packed else (out_s, out_i)

    def _search_call(self, m, xq, p16, ent, k, alpha, prec, alg, out_s, out_i, out_p, after, collect_stats,
                     workspace, library=None, tune=None) -> None:
        """One ``radar_search`` call (k <= RADAR_MAX_K) on already validated device tensors."""
        qs = L.QueriesStruct()
        qs.q = int((xq if xq is not None else p16).shape[0])
        qs.emb_f32 = L.ptr(xq)
        qs.p16, qs.entropy = L.ptr(p16), L.ptr(ent)
        if after is not None:
            qs.after_scores, qs.after_idx = L.ptr(after[0]), L.ptr(after[1])
        cs = self._corpus_struct(m)
        sp = L.SearchParams()
        sp.mode, sp.k, sp.alpha = m, k, float(alpha)
        sp.precision, sp.algo = prec, alg
        sp.kl_variant, sp.overfetch = tune if tune is not None else (L.KL_VARIANT_BY_NAME[self.kl_variant], self.overfetch)
        sp.num_sms = self.num_sms
        lib = library or L.lib()
        with L.device_guard(self.device, lib):
            nbytes = lib.radar_search_workspace_bytes(C.byref(cs), qs.q, C.byref(sp))
            if nbytes == 0:
                L.check(1, "radar_search_workspace_bytes")
            ws = self._get_workspace(int(nbytes), workspace)
            ws_ptr = (ws.data_ptr() + 255) // 256 * 256
            stats = L.SearchStats() if collect_stats else None
            rc = lib.radar_search(C.byref(cs), C.byref(qs), C.byref(sp), L.ptr(out_s), L.ptr(out_i), L.ptr(out_p),
                                  ws_ptr, ws.numel() - (ws_ptr - ws.data_ptr()), C.byref(stats) if stats else None,
                                  L.current_stream_ptr(self.device))
            L.check(rc, "radar_search")
        if stats is not None:
            self.last_stats = SearchStats(stats.algo_used, stats.kernel_launches, int(stats.uncertified),
                                          stats.parts, stats.kprime, float(stats.filter_sm_mhz))

    def debug_filter_keys(self, x, query_probs=None, mask=None, alpha=0.5, mode=None) -> torch.Tensor:
        """Dense [Q,N] dump of the tensor-core filter keys -- bring-up / test aid served by the RADAR_DEBUG flavour
        of the library (``libradar_retrieval_dbg.so``); small problems only."""
        lib = L.debug_lib()
        m = self.resolve_mode(mode, x is not None, query_probs is not None)
        qs = L.QueriesStruct()
        keep = []
        if m != L.MODE_KL:
            xq = _as_device_f32(x, self.device)
            qs.q, qs.emb_f32 = xq.shape[0], L.ptr(xq)
            keep.append(xq)
        if m != L.MODE_DPR:
            p16, ent = prepare_queries(query_probs, mask, self.device, self.eps, self.normalize)
            qs.q, qs.p16, qs.entropy = p16.shape[0], L.ptr(p16), L.ptr(ent)
            keep += [p16, ent]
        q = int(qs.q)
        cs = self._corpus_struct(m)
        sp = L.SearchParams()
        sp.mode, sp.k, sp.alpha = m, min(10, self.ntotal), float(alpha)
        sp.precision, sp.algo = L.PREC_BF16, L.ALGO_TC_FILTER
        sp.num_sms = self.num_sms
        out = torch.full((q, self.ntotal), float("nan"), dtype=torch.float32, device=self.device)
        with L.device_guard(self.device, lib):
            nbytes = lib.radar_search_workspace_bytes(C.byref(cs), q, C.byref(sp))
            if nbytes == 0:
                L.check(1, "radar_search_workspace_bytes")
            ws = self._get_workspace(int(nbytes))
            ws_ptr = (ws.data_ptr() + 255) // 256 * 256
            rc = lib.radar_debug_filter_keys(C.byref(cs), C.byref(qs), C.byref(sp), L.ptr(out), ws_ptr,
                                             ws.numel() - (ws_ptr - ws.data_ptr()), L.current_stream_ptr(self.device))
            if rc != 0:
                raise RuntimeError(f"radar_debug_filter_keys failed (code {rc}): {lib.radar_last_error().decode()}")
        return out


class GpuIndexFlatIP(RadarIndex):
    """``faiss.IndexFlatIP`` spelling: numpy in, numpy out, missing results padded with -1 / -FLT_MAX."""

    def __init__(self, d: int, device: Union[str, torch.device] = "cuda", precision: str = "fp32"):
        super().__init__(d, device=device, precision=precision)

    def search(self, x, k: int, **kw):  # type: ignore[override]
        as_numpy = isinstance(x, np.ndarray)
        nq = x.shape[0] if getattr(x, "ndim", 2) == 2 else 1
        n = self.ntotal
        kk = min(k, n)
        dev_s = torch.full((nq, k), -torch.finfo(torch.float32).max, dtype=torch.float32, device=self.device)
        dev_i = torch.full((nq, k), -1, dtype=torch.int64, device=self.device)
        if kk > 0:
            s, i = super().search(x, kk, **kw)
            dev_s[:, :kk], dev_i[:, :kk] = s, i
        if as_numpy:
            return dev_s.cpu().numpy(), dev_i.cpu().numpy()
        return dev_s, dev_i


class GraphedSearch:
    """A fixed-shape ``search`` call captured once in a CUDA graph and replayed.

    A search is a short chain of kernel launches (query preparation, the fused score + top-k kernel, select /
    re-score / final, and for a sharded index the exchange + merge); for small batches the launch latency of
    that chain is comparable to the kernels themselves.  ``replay()`` re-runs the captured chain on the SAME input
    tensors (copy new queries into them first) and returns the same output tensors.  ``index`` is a
    :class:`RadarIndex` or a ``ShardedRadarIndex``; keyword arguments are those of ``search``.

    Lifetime: the graph holds raw device pointers, so this object (a) owns a PRIVATE workspace that is frozen after
    capture, (b) keeps references to every tensor the captured kernels read or write -- the corpus tensors, the
    inputs, the outputs -- and (c) remembers the index ``generation``: after ``add`` / ``reset`` / ``load`` replaced a
    corpus buffer, ``replay()`` raises instead of reading memory the index no longer owns."""

    def __init__(self, index, x, k: int, warmup: int = 2, **search_kw):
        dev = index.device
        search_kw.pop("collect_stats", None)  # statistics need a stream sync, which a capture cannot contain
        self._base = getattr(index, "index", None) or index  # the RadarIndex of a ShardedRadarIndex
        self._workspace = Workspace(dev)
        search_kw["workspace"] = self._workspace
        self._inputs = (x, search_kw.get("query_probs"), search_kw.get("mask"), search_kw.get("prepared"))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside the capture: workspace allocation, driver entry points
            for _ in range(max(1, warmup)):
                index.search(x, k, **search_kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._workspace.frozen = True
        self._corpus = tuple(self._base._view(nm) for nm in _TABLES)
        self._generation = self._base.generation
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = index.search(x, k, **search_kw)
        self.index = index

    def replay(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._base.generation != self._generation:
            raise RuntimeError("the index changed (add / reset / load) after this search was captured; "
                               "build a new GraphedSearch")
        self.graph.replay()
        return self.out


def merge_topk(scores: torch.Tensor, ids: torch.Tensor, k: int, ascending: bool
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k of the union of per-shard lists; ``scores``/``ids`` are [parts, Q, k_in] CUDA tensors (the
    layout an all-gather produces).  ``radar_merge_topk``."""
    if scores.dim() != 3 or scores.shape != ids.shape:
        raise ValueError("scores/ids must be [parts, Q, k_in] with equal shapes")
    if scores.device.type != "cuda":
        raise RuntimeError("merge_topk runs on a CUDA device only (there is no CPU fallback)")
    parts, q, k_in = scores.shape
    scores = scores.contiguous().float()
    ids = ids.contiguous().long()
    out_s = torch.empty((q, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((q, k), dtype=torch.int64, device=scores.device)
    with L.device_guard(scores.device):
        L.check(L.lib().radar_merge_topk(L.ptr(scores), L.ptr(ids), q, parts, k_in, k, int(ascending), L.ptr(out_s),
                                         L.ptr(out_i), L.current_stream_ptr(scores.device)), "radar_merge_topk")
    return out_s, out_i


def merge_packed(packed: torch.Tensor, k: int, mode: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k of the union of per-shard PACKED lists (``search(..., return_packed=True)``), ``packed`` int64
    [parts, Q, k_in] holding the uint64 words -- ``radar_merge_packed``."""
    if packed.dim() != 3 or packed.dtype != torch.int64:
        raise ValueError("packed must be int64 [parts, Q, k_in]")
    if packed.device.type != "cuda":
        raise RuntimeError("merge_packed runs on a CUDA device only (there is no CPU fallback)")
    parts, q, k_in = packed.shape
    packed = packed.contiguous()
    out_s = torch.empty((q, k), dtype=torch.float32, device=packed.device)
    out_i = torch.empty((q, k), dtype=torch.int64, device=packed.device)
    with L.device_guard(packed.device):
        L.check(L.lib().radar_merge_packed(L.ptr(packed), q, parts, k_in, k, L.MODE_BY_NAME[mode], L.ptr(out_s),
                                           L.ptr(out_i), L.current_stream_ptr(packed.device)), "radar_merge_packed")
    return out_s, out_i


def project_normalize(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], out: str = "fp32",
                      algo: str = "auto"):
    """``F.normalize(F.linear(x, weight, bias), dim=-1)`` (dpr.py:246 / :263) as a fused CUDA operator.

    ``algo`` "auto" / "tc": the tcgen05 GEMM (bf16 hi/lo split x 3 products, fp32 accumulation; components within 1e-5 of
    fp32) when the shape allows (out features == 512, in features % 32 == 0 -- BiomedCLIP's 768 -> 512); "simt": the
    CUDA-core kernel (any shape).  ``out``: "fp32" -> float32[B,512]; "bf16" -> the bfloat16 rows (the A-operand rows of
    the DPR filter); "both" -> (fp32, bf16)."""
    if x.device.type != "cuda":
        raise RuntimeError("project_normalize runs on a CUDA device only")
    if out not in ("fp32", "bf16", "both"):
        raise ValueError("out must be 'fp32', 'bf16' or 'both'")
    x = x.contiguous().float()
    w = weight.detach().contiguous().float()
    b = None if bias is None else bias.detach().contiguous().float()
    rows, in_dim, out_dim = x.shape[0], x.shape[1], w.shape[0]
    lib = L.lib()
    with L.device_guard(x.device):
        ws_bytes = int(lib.radar_project_workspace_bytes(rows, in_dim, out_dim)) if algo in ("auto", "tc") else 0
        if algo == "tc" and ws_bytes == 0:
            raise ValueError("the tensor-core projection needs 512 output features and in features % 32 == 0")
        if ws_bytes:
            y = torch.empty((rows, out_dim), dtype=torch.float32, device=x.device) if out != "bf16" else None
            yb = torch.empty((rows, out_dim), dtype=torch.bfloat16, device=x.device) if out != "fp32" else None
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
            L.check(lib.radar_project_normalize_tc(L.ptr(x), L.ptr(w), L.ptr(b), rows, in_dim, out_dim, L.ptr(y), L.ptr(yb),
                                                   L.ptr(ws), ws_bytes, L.current_stream_ptr(x.device)),
                    "radar_project_normalize_tc")
            ws.record_stream(torch.cuda.current_stream(x.device))
            return y if out == "fp32" else (yb if out == "bf16" else (y, yb))
        y = torch.empty((rows, out_dim), dtype=torch.float32, device=x.device)
        L.check(lib.radar_project_normalize(L.ptr(x), L.ptr(w), L.ptr(b), rows, in_dim, out_dim, L.ptr(y),
                                            L.current_stream_ptr(x.device)), "radar_project_normalize")
    return y if out == "fp32" else (y.bfloat16() if out == "bf16" else (y, y.bfloat16()))
