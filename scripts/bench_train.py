"""configs[1]: LightGCN 3-layer d=64, Yelp2018 shape, one full BPR training epoch on 1 B200.

Every edge is a training interaction: S = trainDataSize triples sampled on the device,
batch 2048 -> 763 steps of bpr_loss + backward + Adam (Procedure.BPR_train_original).  Also times the
reference's CPU code path (oracle port, stageOne) on a bounded sample of steps.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="yelp2018")
ap.add_argument("--epochs", type=int, default=3)
ap.add_argument("--fused-adam", type=int, default=1)
ap.add_argument("--cpu-steps", type=int, default=3)
ap.add_argument("--graph", type=int, default=1)
args = ap.parse_args()

from factors_of_serendipity_recommendation_b200 import Procedure, dataloader, model, synth, utils, world
nu, mi, E, d = synth.SHAPES[args.workload]
u, i = synth.make_interactions(nu, mi, E, seed=2020)
cfg = dict(world.config)
cfg.update(lightGCN_n_layers=3, latent_dim_rec=d, fused_adam=bool(args.fused_adam))
world.configure(bpr_batch_size=2048, cuda_graph=bool(args.graph and args.fused_adam))
ds = dataloader.InteractionDataset(nu, mi, u, i, device="cuda")
torch.manual_seed(2020)
m = model.LightGCN(cfg, ds).cuda()
bpr = utils.BPRLoss(m, cfg)
g = ds.getGraphHandle()
np.random.seed(2020)
times, infos = [], []
for ep in range(args.epochs + 1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    infos.append(Procedure.BPR_train_original(ds, m, bpr, ep))
    torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
steps = E // 2048 + 1
best = min(times[1:])
layer_bytes = g.nnz * 8 + (g.n_rows + 1) * 4 + 2 * g.n_rows * d * 4
out = {"workload": args.workload, "steps_per_epoch": steps, "epoch_s": best, "epoch_s_all": times, "steps_per_s": steps / best,
       "ms_per_step": 1e3 * best / steps, "propagated_edges_per_s": 6 * g.nnz * steps / best,
       "algorithmic_bytes_per_step": 6 * layer_bytes + 7 * g.n_rows * d * 4, "fused_adam": bool(args.fused_adam), "cuda_graph": bool(args.graph and args.fused_adam),
       "loss_trace": infos}
if args.cpu_steps > 0:
    from oracle import lightgcn_oracle as O
    torch.set_num_threads(os.cpu_count())
    ref = O.OracleLightGCN(nu, mi, u, i, latent_dim=d, n_layers=3)
    opt = torch.optim.Adam([ref.user_w, ref.item_w], lr=cfg["lr"])
    S = g.sample_bpr(2048 * (args.cpu_steps + 1), seed=1).cpu()
    ts = []
    for k in range(args.cpu_steps + 1):
        b = S[k * 2048:(k + 1) * 2048]
        t0 = time.perf_counter()
        O.stage_one(ref, opt, b[:, 0], b[:, 1], b[:, 2], cfg["decay"])
        ts.append(time.perf_counter() - t0)
    out["cpu_reference"] = {"ms_per_step": 1e3 * min(ts[1:]), "cores": os.cpu_count(), "kind": "port",
                            "sample": f"{args.cpu_steps} stageOne steps (PT/utils.py:43-52), epoch would take {min(ts[1:]) * steps:.0f} s"}
print(json.dumps(out))
