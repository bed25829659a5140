"""GPU bring-up aid: dense dump of the tcgen05 filter keys vs bf16 products, per mode, with diagnostics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import make_problem
from radar_multimodal_radiology_b200.index import RadarIndex
from oracle import c_oracle as co

dev = torch.device("cuda:0")
n, q = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 150
p = make_problem(n, q, seed=10)
idx = RadarIndex(512, device=dev)
idx.add(p["c_emb"]); idx.add_observations(p["c_pr"])
bf = lambda a: torch.from_numpy(np.ascontiguousarray(a)).bfloat16().double().numpy()
ip = bf(p["q_emb"]) @ bf(p["c_emb"]).T
logq = co.prepare_corpus(p["c_pr"]).astype(np.float64)
p16, ent = co.prepare_queries(p["q_pr"], p["mask"])
kl_key = p16.astype(np.float64) @ logq.T - ent.astype(np.float64)[:, None]
hyb = (bf(0.5 * p["q_emb"]) @ bf(p["c_emb"]).T) + 0.5 * kl_key
for mode, want, kw in (("dpr", ip, {}), ("kl", kl_key, dict(query_probs=p["q_pr"], mask=p["mask"])),
                       ("hybrid", hyb, dict(query_probs=p["q_pr"], mask=p["mask"], alpha=0.5))):
    try:
        keys = idx.debug_filter_keys(None if mode == "kl" else p["q_emb"], mode=mode, **kw)
        torch.cuda.synchronize()
        keys = keys.cpu().numpy()
        err = np.abs(keys - want)
        print(f"{mode}: nan={np.isnan(keys).sum()} max_err={np.nanmax(err):.3e} mean_err={np.nanmean(err):.3e} "
              f"want_absmax={np.abs(want).max():.3f}")
        if np.nanmax(err) > 1e-3:
            bad = np.argwhere(err > 1e-3)
            print("  first bad (q,n):", bad[:8].tolist(), "rows bad:", np.unique(bad[:, 0])[:16], "cols bad:", np.unique(bad[:, 1])[:16])
            print("  got", keys[bad[0][0], bad[0][1]], "want", want[bad[0][0], bad[0][1]])
            print("  corr:", np.corrcoef(np.nan_to_num(keys).ravel(), want.ravel())[0, 1])
    except Exception as e:
        print(mode, "FAILED:", repr(e))
        break
