"""Rows a1/a2 (graph build) and a13 (Procedure.Test) of SURVEY.md section 8, timed through the public API."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from factors_of_serendipity_recommendation_b200 import Procedure, _lgx, dataloader, model, synth, world

out = {}
for name in ("gowalla", "amazon-book"):
    nu, mi, E, d = synth.SHAPES[name]
    u, i = synth.make_interactions(nu, mi, E, seed=2020)
    tu, ti = torch.from_numpy(u), torch.from_numpy(i)
    _lgx.Graph.build(nu, mi, tu.cuda(), ti.cuda())                      # warm-up (context, first-launch costs)
    torch.cuda.synchronize()
    ts_host, ts_dev = [], []
    for _ in range(5):
        t0 = time.perf_counter(); g = _lgx.Graph.build(nu, mi, tu, ti); torch.cuda.synchronize(); ts_host.append(time.perf_counter() - t0)
        du, di = tu.cuda(), ti.cuda(); torch.cuda.synchronize()
        t0 = time.perf_counter(); g = _lgx.Graph.build(nu, mi, du, di); torch.cuda.synchronize(); ts_dev.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    from oracle import lightgcn_oracle as O
    O.build_norm_adj(nu, mi, u, i)
    t_cpu = time.perf_counter() - t0
    out[name] = {"edges": E, "nnz": g.nnz, "build_ms_from_host_arrays": 1e3 * min(ts_host), "build_ms_from_device_arrays": 1e3 * min(ts_dev),
                 "cpu_numpy_restatement_ms": 1e3 * t_cpu}
    if name == "amazon-book":
        ue, ie = synth.make_embeddings(nu, mi, d, seed=2020, trained_like=True)
        test_dict = synth.make_test_dict(nu, mi, u, i, per_user=5)
        cfg = dict(world.config); cfg.update(pretrain=1, user_emb=ue.numpy(), item_emb=ie.numpy())
        ds = dataloader.InteractionDataset(nu, mi, u, i, test_dict=test_dict, device="cuda")
        m = model.LightGCN(cfg, ds).cuda()
        world.configure(topks=[20])
        for dev_metrics in (False, True):
            Procedure.Test(ds, m, 0, device_metrics=dev_metrics)
            ts = []
            for _ in range(3):
                m._eval_cache = None
                torch.cuda.synchronize(); t0 = time.perf_counter()
                res = Procedure.Test(ds, m, 0, device_metrics=dev_metrics)
                torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            out[name][f"Procedure.Test_s_{'device' if dev_metrics else 'numpy'}_metrics"] = min(ts)
            out[name][f"metrics_{'device' if dev_metrics else 'numpy'}"] = {k: float(v[0]) for k, v in res.items()}
print(json.dumps(out))
