"""Row-sharded multi-GPU search (the reference has no distributed code).

One process per GPU (``torchrun``), ``torch.distributed`` for the plumbing.  Rank g holds corpus rows
``[g*ceil(N/G), min(N,(g+1)*ceil(N/G)))`` of both the embedding matrix and the log-probability table;
queries are replicated.  Every rank runs the local fused score + top-k kernels, which also emit the result as ONE
sortable 64-bit word per entry -- (orderable key bits << 32) | (0xFFFFFFFF - global id), ``radar_search``'s
``out_packed`` -- so the exchange is a single all-gather of ``Q*k*8`` bytes per rank over NVLink / NVSwitch (a few MB:
latency-bound, which is why plain NCCL is used for it) followed by one merge kernel that sorts the G*k words of a
query.  The (key, id) order is total, so in fp32 precision (and on the exact path) the result does not depend on G;
in bf16 precision every rank's list is its shard's bf16-filtered top-k (recall >= 0.999 against fp32 per shard, hence
for the union).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from ._nvtx import annotate as _nvtx_annotate, range_ as _nvtx_range


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    per = -(-n // world)
    return min(n, rank * per), min(n, (rank + 1) * per)


def mode_name(mode: Optional[str], have_emb: bool, have_probs: bool) -> str:
    if mode is not None:
        return mode
    if have_emb and have_probs:
        return "hybrid"
    return "dpr" if have_emb else "kl"


class ShardedRadarIndex:
    """A ``RadarIndex`` per rank + all-gather + merge.

    ``local_search`` / ``merge`` are injectable so that the host-side logic (partitioning, padding of
    short shards, gather layout, packed-word ordering) can be exercised on CPU with the ``gloo`` backend in
    tests; the defaults are the CUDA kernels and there is no CPU fallback in the product path.
    ``local_search(x, k, ..., return_packed=True) -> (scores, ids, packed int64[Q,k])``,
    ``merge(packed int64[G,Q,k], k, mode) -> (scores, ids)``."""

    def __init__(self, d: int = 512, device="cuda", group: Optional[dist.ProcessGroup] = None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None, **index_kw):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device)
        self.d = d
        self._index_kw = index_kw
        self.index = None
        self._local_search = local_search
        self._merge = merge
        self.n_total = 0
        self.lo = self.hi = 0

    def build(self, n_total: int, local_embeddings=None, local_probs=None) -> "ShardedRadarIndex":
        """Index this rank's rows (already sliced by the caller with :func:`shard_bounds`)."""
        self.n_total = int(n_total)
        self.lo, self.hi = shard_bounds(self.n_total, self.world, self.rank)
        if self._local_search is None:
            from .index import RadarIndex
            self.index = RadarIndex(self.d, device=self.device, idx_offset=self.lo, **self._index_kw)
            if local_embeddings is not None and self.hi > self.lo:
                self.index.add(local_embeddings)
            if local_probs is not None and self.hi > self.lo:
                self.index.add_observations(local_probs)
            if self.index.ntotal != self.hi - self.lo:
                raise RuntimeError(f"rank {self.rank}: indexed {self.index.ntotal} rows, expected {self.hi - self.lo}")
        return self

    @classmethod
    def wrap(cls, local_index, n_total: int, group: Optional[dist.ProcessGroup] = None) -> "ShardedRadarIndex":
        """Adopt an existing per-rank ``RadarIndex`` (e.g. ``RadarIndex.view_rows`` of a resident shard) whose rows
        are this rank's slice ``shard_bounds(n_total, world, rank)`` and whose ``idx_offset`` is that slice's start."""
        sh = cls(local_index.d, device=local_index.device, group=group)
        sh.n_total = int(n_total)
        sh.lo, sh.hi = shard_bounds(sh.n_total, sh.world, sh.rank)
        if local_index.ntotal != sh.hi - sh.lo or local_index.idx_offset != sh.lo:
            raise RuntimeError(f"rank {sh.rank}: local index has {local_index.ntotal} rows at offset "
                               f"{local_index.idx_offset}, expected {sh.hi - sh.lo} at {sh.lo}")
        sh.index = local_index
        return sh

    @property
    def ntotal(self) -> int:
        return self.n_total

    # ---- queries that arrive on the host: every rank uploads 1/G of the rows over its own PCIe link and the ranks
    # exchange the slices over NVLink, instead of G copies of the whole batch crossing PCIe
    @_nvtx_annotate("sharded.upload_queries")
    def upload_queries(self, host: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``host``: the same (pinned) [Q, ...] tensor on every rank.  Returns the [Q, ...] device tensor."""
        if self.world == 1:
            if out is None:
                return host.to(self.device, non_blocking=True)
            out.copy_(host, non_blocking=True)
            return out
        q = host.shape[0]
        per = -(-q // self.world)
        padded = torch.empty((per * self.world,) + tuple(host.shape[1:]), dtype=host.dtype, device=self.device)
        lo, hi = min(q, self.rank * per), min(q, (self.rank + 1) * per)
        mine = padded[self.rank * per:self.rank * per + per]
        if hi > lo:
            mine[:hi - lo].copy_(host[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(padded, mine, group=self.group)
        if out is None:
            return padded[:q]
        out.copy_(padded[:q])
        return out

    @_nvtx_annotate("sharded.search")
    def search(self, x, k: int, query_probs=None, mask=None, alpha: float = 0.5, mode: Optional[str] = None,
               **kw) -> Tuple[torch.Tensor, torch.Tensor]:
        if k > self.n_total:
            raise ValueError(f"k={k} exceeds ntotal={self.n_total}")
        n_local = self.hi - self.lo
        k_local = min(k, n_local)
        fn = self._local_search or self.index.search
        if self.world == 1:
            return fn(x, k, query_probs=query_probs, mask=mask, alpha=alpha, mode=mode, **kw)
        mname = mode_name(mode, x is not None, query_probs is not None or kw.get("prepared") is not None)
        nq = (x if x is not None else (query_probs if query_probs is not None else kw["prepared"][0])).shape[0]
        if k_local == k:
            _, _, packed = fn(x, k, query_probs=query_probs, mask=mask, alpha=alpha, mode=mode, return_packed=True, **kw)
        else:  # a shard with fewer than k rows pads its list with empty words
            packed = torch.zeros((nq, k), dtype=torch.int64, device=self.device)
            if k_local > 0:
                packed[:, :k_local] = fn(x, k_local, query_probs=query_probs, mask=mask, alpha=alpha, mode=mode,
                                         return_packed=True, **kw)[2]
        # concatenated-along-dim-0 output is the form both NCCL and gloo accept; viewed as [G, Q, k]
        gathered = torch.empty((self.world * nq, k), dtype=torch.int64, device=self.device)
        with _nvtx_range("sharded.all_gather"):
            dist.all_gather_into_tensor(gathered, packed, group=self.group)
        gathered = gathered.view(self.world, nq, k)
        with _nvtx_range("sharded.merge"):
            if self._merge is not None:
                return self._merge(gathered, k, mname)
            from .index import merge_packed
            return merge_packed(gathered, k, mname)
