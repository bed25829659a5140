#!/usr/bin/env python
"""Where does a sharded search step spend its time?  (torchrun, one rank per GPU)
    python -m torch.distributed.run --nproc-per-node 2 tools/time_sharded.py [kl|dpr] [n] [q] [k]
Times, with CUDA events on rank 0: the local search (packed output), the all-gather of the packed words, the merge kernel."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radar_multimodal_radiology_b200 import synthetic as syn  # noqa: E402
from radar_multimodal_radiology_b200.index import RadarIndex, merge_packed  # noqa: E402
from radar_multimodal_radiology_b200.sharded import shard_bounds  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "kl"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 377000
    q = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_bounds(n, world, rank)
    idx = RadarIndex(512, device=dev, idx_offset=lo, precision="fp32")
    if mode != "kl":
        idx.add(syn.embeddings(n, 512, syn.SEED_CORPUS_EMB, dev)[lo:hi])
    if mode != "dpr":
        idx.add_observations(syn.observation_probs(n, syn.SEED_CORPUS_PROBS, dev)[lo:hi])
    qe = syn.embeddings(q, 512, syn.SEED_QUERY_EMB, dev) if mode != "kl" else None
    qp = syn.observation_probs(q, syn.SEED_QUERY_PROBS, dev) if mode != "dpr" else None
    gathered = torch.empty((world * q, k), dtype=torch.int64, device=dev)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    for it in range(6):
        e = [ev() for _ in range(4)]
        dist.barrier()
        torch.cuda.synchronize()
        e[0].record()
        _, _, packed = idx.search(qe, k, query_probs=qp, mode=mode, return_packed=True)
        e[1].record()
        dist.all_gather_into_tensor(gathered, packed)
        e[2].record()
        merge_packed(gathered.view(world, q, k), k, mode)
        e[3].record()
        torch.cuda.synchronize()
        if rank == 0 and it >= 2:
            print(f"{mode} n={n} q={q} k={k} world={world}: search {e[0].elapsed_time(e[1]):.3f} ms, all-gather "
                  f"{e[1].elapsed_time(e[2]):.3f} ms, merge {e[2].elapsed_time(e[3]):.3f} ms", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
