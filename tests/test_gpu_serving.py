"""serving.HostPipeline: overlapped host -> device -> host requests return exactly what the synchronous
model.topk call returns for the same tables, in order, with no slot clobbered (GPU)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_matches_synchronous_topk():
    from factors_of_serendipity_recommendation_b200 import dataloader, model, serving, synth, world
    nu, mi, d, k = 700, 1500, 64, 20
    u, i = synth.make_interactions(nu, mi, 20000, seed=3)
    cfg = dict(world.config)
    cfg.update(lightGCN_n_layers=3, latent_dim_rec=d)
    ds = dataloader.InteractionDataset(nu, mi, u, i, device="cuda")
    m = model.LightGCN(cfg, ds).cuda().eval()
    users = torch.arange(nu, device="cuda")
    tables = [synth.make_embeddings(nu, mi, d, seed=s, trained_like=True) for s in range(5)]
    want = []
    for ue, ie in tables:                                   # synchronous path: load the tables, call topk
        m.embedding_user.weight.data.copy_(ue)
        m.embedding_item.weight.data.copy_(ie)
        m._eval_cache = None
        m._packed = {}
        idx, _ = m.topk(users, k, mode="bf16x3")
        want.append(idx.cpu().numpy())
    for depth in (1, 2, 3):
        pipe = serving.HostPipeline.for_model(m, users, k, mode="bf16x3", depth=depth)
        outs = [torch.empty(nu, k, dtype=torch.int64).pin_memory() for _ in tables]
        pinned = [(ue.pin_memory(), ie.pin_memory()) for ue, ie in tables]
        for (ue, ie), out in zip(pinned, outs):
            pipe.submit(ue, ie, out)
        pipe.wait()
        for got, ref in zip(outs, want):
            assert np.array_equal(got.numpy(), ref)


def test_pipeline_rejects_device_tensors_and_bad_shapes():
    from factors_of_serendipity_recommendation_b200 import dataloader, model, serving, synth, world
    nu, mi, d = 300, 500, 64
    u, i = synth.make_interactions(nu, mi, 6000, seed=1)
    cfg = dict(world.config)
    cfg.update(lightGCN_n_layers=2, latent_dim_rec=d)
    m = model.LightGCN(cfg, dataloader.InteractionDataset(nu, mi, u, i, device="cuda")).cuda().eval()
    pipe = serving.HostPipeline.for_model(m, torch.arange(nu, device="cuda"), 20)
    out = torch.empty(nu, 20, dtype=torch.int64)
    with pytest.raises(ValueError):
        pipe.submit(torch.zeros(nu, d, device="cuda"), torch.zeros(mi, d), out)
    with pytest.raises(ValueError):
        pipe.submit(torch.zeros(nu, d), torch.zeros(mi + 1, d), out)
