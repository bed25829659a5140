"""CPU, world_size 2 over gloo: the multi-GPU host logic (row partition, short-shard padding, all-gather
layout of the packed result words, global ids, merge order, sliced query upload).  The local search and the merge are injected from the oracle -- in the
product they are the CUDA kernels (tests/test_gpu_parity.py covers those); nothing here is a product
fallback."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q, k, mode_name, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import make_problem
    from oracle import c_oracle as co
    from radar_multimodal_radiology_b200.sharded import ShardedRadarIndex, shard_bounds

    p = make_problem(n, q, seed=21)
    mode = {"dpr": 0, "kl": 1, "hybrid": 2}[mode_name]
    lo, hi = shard_bounds(n, world, rank)
    logq = co.prepare_corpus(p["c_pr"][lo:hi]) if hi > lo else np.zeros((0, 16), np.float32)

    def local_search(x, kk, query_probs=None, mask=None, alpha=0.5, mode=None, return_packed=False, **kw):
        p16, ent = co.prepare_queries(query_probs.numpy(), None if mask is None else mask.numpy())
        m = {"dpr": 0, "kl": 1, "hybrid": 2}[mode]
        s, i = co.search(m, kk, q_emb=None if x is None else x.numpy(), p16=p16, entropy=ent,
                         c_emb=p["c_emb"][lo:hi], logq16=logq, alpha=alpha, idx_offset=lo)
        out = (torch.from_numpy(s), torch.from_numpy(i))
        return out + (torch.from_numpy(co.pack_results(m, s, i)),) if return_packed else out

    def merge(packed, kk, mode):
        s, i = co.merge_packed(packed.numpy(), kk, {"dpr": 0, "kl": 1, "hybrid": 2}[mode])
        return torch.from_numpy(s), torch.from_numpy(i)

    idx = ShardedRadarIndex(512, device="cpu", local_search=local_search, merge=merge).build(n)
    assert (idx.lo, idx.hi) == (lo, hi) and idx.ntotal == n
    x = None if mode_name == "kl" else torch.from_numpy(p["q_emb"])
    s, i = idx.search(x, k, query_probs=torch.from_numpy(p["q_pr"]), mask=torch.from_numpy(p["mask"]), alpha=0.5,
                      mode=mode_name)
    # host-resident queries: every rank uploads its 1/G slice, the slices are all-gathered (ragged last slice)
    host = torch.arange(q * 3, dtype=torch.float32).reshape(q, 3)
    up = idx.upload_queries(host)
    assert up.shape == host.shape and torch.equal(up, host)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), s=s.numpy(), i=i.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode_name,n,k", [("hybrid", 1003, 10), ("kl", 1003, 32), ("dpr", 17, 10)])
def test_world2_sharded_search_equals_single_shard(tmp_path, mode_name, n, k):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import make_problem
    from oracle import c_oracle as co
    q = 12
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n, q, k, mode_name, str(tmp_path)), nprocs=2, join=True)
    p = make_problem(n, q, seed=21)
    logq = co.prepare_corpus(p["c_pr"])
    p16, ent = co.prepare_queries(p["q_pr"], p["mask"])
    mode = {"dpr": 0, "kl": 1, "hybrid": 2}[mode_name]
    want_s, want_i = co.search(mode, k, q_emb=p["q_emb"], p16=p16, entropy=ent, c_emb=p["c_emb"], logq16=logq,
                               alpha=0.5)
    for r in range(2):
        got = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(got["i"], want_i), f"rank {r} ids"
        assert np.array_equal(got["s"], want_s), f"rank {r} scores"
