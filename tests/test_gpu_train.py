"""BPR loss / backward / Adam / sampler / metrics kernels vs the reference's golden training step (GPU)."""
import numpy as np
import pytest
import torch

from oracle import lightgcn_oracle as O

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def make(mlls, t, **cfg_over):
    from factors_of_serendipity_recommendation_b200 import dataloader, model, world
    cfg = dict(world.config)
    cfg.update(lightGCN_n_layers=int(t["n_layers"]), pretrain=1, user_emb=t["w0_user"], item_emb=t["w0_item"],
               lr=float(t["lr"]), decay=float(t["decay"]))
    cfg.update(cfg_over)
    ds = dataloader.InteractionDataset(mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"],
                                       test_dict=mlls["test_dict"], device="cuda")
    return model.LightGCN(cfg, ds).cuda().train(), ds, cfg


def batch(t):
    return tuple(torch.from_numpy(t[k]).long().cuda() for k in ("users", "pos", "neg"))


def test_bpr_loss_and_gradients_match_reference(mlls, train_step):
    t = train_step
    m, _, cfg = make(mlls, t)
    u, p, n = batch(t)
    loss, reg = m.bpr_loss(u, p, n)
    assert loss.dim() == 0 and reg.dim() == 0 and loss.requires_grad
    assert abs(loss.item() - float(t["loss"])) <= 1e-5 * abs(float(t["loss"]))
    assert abs(reg.item() - float(t["reg_loss"])) <= 1e-5 * abs(float(t["reg_loss"]))
    (loss + reg * cfg["decay"]).backward()
    assert rel_err(m.embedding_user.weight.grad.cpu().numpy(), t["grad_user"]) <= 1e-5
    assert rel_err(m.embedding_item.weight.grad.cpu().numpy(), t["grad_item"]) <= 1e-5


def test_reference_style_loss_on_top_of_computer(mlls, train_step):
    """The reference's own bpr_loss body (gathers + softplus in torch) on our computer(): autograd glue."""
    t = train_step
    m, _, cfg = make(mlls, t)
    u, p, n = batch(t)
    ue, pe, ne, u0, p0, n0 = m.getEmbedding(u, p, n)
    reg = 0.5 * (u0.norm(2).pow(2) + p0.norm(2).pow(2) + n0.norm(2).pow(2)) / float(len(u))
    loss = torch.mean(torch.nn.functional.softplus((ue * ne).sum(1) - (ue * pe).sum(1)))
    assert abs(loss.item() - float(t["loss"])) <= 1e-5
    (loss + reg * cfg["decay"]).backward()
    assert rel_err(m.embedding_user.weight.grad.cpu().numpy(), t["grad_user"]) <= 1e-5
    assert rel_err(m.embedding_item.weight.grad.cpu().numpy(), t["grad_item"]) <= 1e-5
    gamma = m.forward(u[:64], p[:64])
    assert gamma.shape == (64,)


@pytest.mark.parametrize("fused", [False, True])
def test_stage_one_step_matches_reference(mlls, train_step, fused):
    from factors_of_serendipity_recommendation_b200 import utils
    t = train_step
    m, _, cfg = make(mlls, t, fused_adam=fused)
    bpr = utils.BPRLoss(m, cfg)
    u, p, n = batch(t)
    cri = bpr.stageOne(u, p, n)
    assert isinstance(cri, float)
    assert abs(cri - (float(t["loss"]) + float(t["decay"]) * float(t["reg_loss"]))) <= 1e-5
    # Adam's first step moves every touched weight by ~lr: compare the update itself
    du = m.embedding_user.weight.detach().cpu().numpy() - t["w0_user"]
    di = m.embedding_item.weight.detach().cpu().numpy() - t["w0_item"]
    ru, ri = t["w1_user"] - t["w0_user"], t["w1_item"] - t["w0_item"]
    # (scatter-add order is not deterministic and Adam's first step is ~ lr * sign(g): 0.1% of lr)
    assert np.abs(du - ru).max() <= 1e-3 * float(t["lr"])
    assert np.abs(di - ri).max() <= 1e-3 * float(t["lr"])
    m.eval()
    with torch.no_grad():
        gamma = m.forward(u[:64], p[:64])
    assert np.abs(gamma.cpu().numpy() - t["gamma_after"]).max() <= 1e-5 * np.abs(t["gamma_after"]).max() + 1e-7
    # a second step runs and the eval cache saw the update
    c2 = bpr.stageOne(u, p, n, sync=False)
    assert torch.is_tensor(c2) and c2.item() < cri


def test_adam_kernel_matches_torch():
    from factors_of_serendipity_recommendation_b200 import _lgx
    torch.manual_seed(0)
    p = torch.randn(1000, 64, device="cuda")
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn_like(p)
        ref.grad = g.clone()
        opt.step()
        _lgx.adam_step(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, step)
        assert (p - ref.detach()).abs().max().item() <= 1e-6


def test_device_sampler(mlls):
    from factors_of_serendipity_recommendation_b200 import dataloader, utils
    ds = dataloader.InteractionDataset(mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"], device="cuda")
    S = utils.UniformSample_original(ds, seed=7)
    assert S.shape == (ds.trainDataSize, 3) and S.dtype == torch.int64
    S2 = utils.UniformSample_original(ds, seed=7)
    assert torch.equal(S, S2)                                       # counter-based: reproducible
    S = S.cpu().numpy()
    train = set((mlls["train_user"].astype(np.int64) * ds.m_items + mlls["train_item"]).tolist())
    assert all(int(u) * ds.m_items + int(p) in train for u, p, _ in S[:5000])
    assert not any(int(u) * ds.m_items + int(n) in train for u, _, n in S[:5000])
    assert S[:, 2].min() >= 0 and S[:, 2].max() < ds.m_items
    # user marginal is uniform (PT/utils.py:76): chi-square-ish bound
    cnt = np.bincount(S[:, 0], minlength=ds.n_users)
    assert abs(cnt.mean() - ds.trainDataSize / ds.n_users) < 1e-9 and cnt.std() < 3 * np.sqrt(cnt.mean())
    # sampling.cpp semantics: every user exactly per_user triples (PT/sources/sampling.cpp:29-42)
    P = ds.getGraphHandle().sample_bpr(0, per_user=3, seed=1).cpu().numpy()
    assert P.shape == (ds.n_users * 3, 3) and np.array_equal(P[:, 0], np.repeat(np.arange(ds.n_users), 3))


def test_device_sampler_vs_reference_native_sampler(mlls):
    """lgx_sample_bpr(per_user=...) against the reference's OWN compiled sampler (oracle/_ref, PT/sources/sampling.cpp):
    same contract, same layout, same pos / neg marginals (chi-square, 16 bins); and the Python-sampler mode against the
    oracle's restatement of PT/utils.py:67-99."""
    from factors_of_serendipity_recommendation_b200 import dataloader
    from oracle import lightgcn_oracle as O
    nu, mi = mlls["n_users"], mlls["m_items"]
    ds = dataloader.InteractionDataset(nu, mi, mlls["train_user"], mlls["train_item"], device="cuda")
    all_pos = [[] for _ in range(nu)]
    for u, i in zip(mlls["train_user"].tolist(), mlls["train_item"].tolist()):
        all_pos[u].append(i)

    def chi2(a, b):
        a, b = np.asarray(a, float), np.asarray(b, float)
        ka, kb = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
        keep = (a + b) > 0
        return float((((ka * a - kb * b) ** 2) / (a + b))[keep].sum())

    S_dev = ds.getGraphHandle().sample_bpr(0, per_user=40, seed=11).cpu().numpy()
    O.check_bpr_triples(S_dev, all_pos, mi, per_user=40)
    ref = O.load_reference_sampler()
    if ref is not None:
        ref.seed(2020)
        S_ref = np.asarray(ref.sample_negative(nu, mi, 40 * nu, all_pos, 1))
    else:                                                   # the restatement (itself checked against oracle/_ref on CPU)
        S_ref = O.sample_per_user(nu, mi, 40 * nu, all_pos, np.random.RandomState(3))
    assert S_dev.shape == S_ref.shape
    for h_dev, h_ref in zip(O.sampler_marginals(S_dev, all_pos, mi), O.sampler_marginals(S_ref, all_pos, mi)):
        assert chi2(h_dev, h_ref) < 37.7                    # dof 15, p = 0.999
    # Python-sampler semantics (users drawn with replacement)
    n = 40 * nu
    S_dev = ds.getGraphHandle().sample_bpr(n, per_user=0, seed=5).cpu().numpy()
    O.check_bpr_triples(S_dev, all_pos, mi)
    S_py = O.uniform_sample_python(nu, mi, n, all_pos, np.random.RandomState(1))
    for h_dev, h_ref in zip(O.sampler_marginals(S_dev, all_pos, mi), O.sampler_marginals(S_py, all_pos, mi)):
        assert chi2(h_dev, h_ref) < 37.7
    hu_dev = np.histogram(S_dev[:, 0], bins=16, range=(0, nu))[0]
    hu_py = np.histogram(S_py[:, 0], bins=16, range=(0, nu))[0]
    assert chi2(hu_dev, hu_py) < 37.7


def test_train_epoch_runs_and_learns(mlls, train_step):
    from factors_of_serendipity_recommendation_b200 import Procedure, utils, world
    t = train_step
    m, ds, cfg = make(mlls, t, lr=0.01)
    world.configure(bpr_batch_size=2048, topks=[20])
    bpr = utils.BPRLoss(m, cfg)
    np.random.seed(0)
    before = Procedure.Test(ds, m, 0)["recall"][0]
    for ep in range(3):
        info = Procedure.BPR_train_original(ds, m, bpr, ep)
    assert info.startswith("loss") and "Sample" in info
    after = Procedure.Test(ds, m, 3)["recall"][0]
    assert after > before + 0.02                                    # random init ~0.01 -> learns


def test_edge_dropout_matches_torch_on_the_same_dropped_graph(mlls):
    """--dropout 1 (PT/model.py:125-143): keep mask is evaluated inside the SpMM; with the exported mask the
    torch reference on the explicitly dropped COO graph gives the same forward and the same gradients."""
    from factors_of_serendipity_recommendation_b200 import _lgx, synth
    nu, mi = mlls["n_users"], mlls["m_items"]
    g = _lgx.Graph.build(nu, mi, torch.from_numpy(mlls["train_user"]).cuda(), torch.from_numpy(mlls["train_item"]).cuda(), chunk_nnz=64)
    g.enable_dropout()
    keep, seed, L, d = 0.6, 12345, 3, 64
    mask = g.dropout_mask(keep, seed).bool()
    frac = mask.float().mean().item()
    assert abs(frac - keep) < 0.01                                         # Bernoulli(keep_prob) per stored entry
    assert not torch.equal(mask, g.dropout_mask(keep, seed + 1).bool())    # new seed, new graph
    assert torch.equal(mask, g.dropout_mask(keep, seed).bool())            # same seed, same graph
    e = g.export()
    rows = torch.repeat_interleave(torch.arange(g.n_rows, device="cuda"), e["indptr"][1:] - e["indptr"][:-1])
    cols = e["indices"].long()
    # mirrored-entry mask == mask looked up at the transposed position
    tmask = g.dropout_mask(keep, seed, transpose=True).bool()
    dense_keep = torch.zeros(g.n_rows, g.n_rows, dtype=torch.bool, device="cuda")
    dense_keep[rows, cols] = mask
    assert torch.equal(tmask, dense_keep[cols, rows])
    A = torch.sparse_coo_tensor(torch.stack([rows[mask], cols[mask]]), e["values"][mask] / keep, (g.n_rows, g.n_rows)).coalesce()
    ue, ie = synth.make_embeddings(nu, mi, d, seed=9)
    E0 = torch.cat([ue, ie]).cuda().requires_grad_(True)
    x, embs = E0, [E0]
    for _ in range(L):
        x = torch.sparse.mm(A, x)
        embs.append(x)
    ref = torch.stack(embs, 1).mean(1)
    out = g.propagate_fwd(E0.detach(), L, dropout=(keep, seed))
    assert rel_err(out.cpu().numpy(), ref.detach().cpu().numpy()) <= 1e-5
    W = torch.randn_like(ref)
    (ref * W).sum().backward()
    dE0 = g.propagate_bwd((W / (L + 1)).contiguous(), L, dropout=(keep, seed))
    assert rel_err(dE0.cpu().numpy(), E0.grad.cpu().numpy()) <= 1e-5


def test_model_with_dropout_trains(mlls, train_step):
    from factors_of_serendipity_recommendation_b200 import utils
    t = train_step
    m, _, cfg = make(mlls, t, dropout=1, keep_prob=0.6)
    m.train()
    u, p, n = batch(t)
    l1, _ = m.bpr_loss(u, p, n)
    l2, _ = m.bpr_loss(u, p, n)
    assert abs(l1.item() - l2.item()) > 0                                   # a new dropped graph per call
    assert abs(l1.item() - float(t["loss"])) < 0.01                         # close to the undropped loss at init
    bpr = utils.BPRLoss(m, cfg)
    c0 = bpr.stageOne(u, p, n)
    for _ in range(5):
        c = bpr.stageOne(u, p, n)
    assert c < c0
    m.eval()
    with torch.no_grad():
        a = m.computer()[0].clone()
        b = m.computer()[0]
    assert torch.equal(a, b)                                                # eval: no dropout (PT/model.py:158-159)


def test_cuda_graph_step_equals_eager(mlls, train_step):
    """One captured whole-step graph replayed == the same steps run eagerly (same batches, fused Adam)."""
    from factors_of_serendipity_recommendation_b200 import utils
    t = train_step
    u, p, n = batch(t)
    outs = []
    for graphed in (False, True):
        m, _, cfg = make(mlls, t, fused_adam=True)
        bpr = utils.BPRLoss(m, cfg)
        losses = [bpr.stageOne(u, p, n, sync=False).item() for _ in range(2)]       # eager warm-up steps
        if graphed:
            gs = utils.GraphedStageOne(bpr, u, p, n)
            losses += [gs.run(u.flip(0), p.flip(0), n.flip(0)).item(), gs.run(u, p, n).item()]
        else:
            losses += [bpr.stageOne(u.flip(0), p.flip(0), n.flip(0), sync=False).item(), bpr.stageOne(u, p, n, sync=False).item()]
        outs.append((losses, m.embedding_user.weight.detach().clone()))
    assert np.allclose(outs[0][0], outs[1][0], rtol=1e-5)
    assert (outs[0][1] - outs[1][1]).abs().max().item() <= 1e-5
    assert outs[0][0][3] < outs[0][0][0]
