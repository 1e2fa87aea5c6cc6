"""computer() forward / backward on the device vs the oracle and the reference's golden outputs (GPU).
Tolerance: max|x - ref| / max|ref| <= 1e-5 (fp32 contract of BASELINE.json)."""
import numpy as np
import pytest
import torch

from oracle import lightgcn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def make_model(nu, mi, u, i, n_layers, user_emb=None, item_emb=None, d=64, seed=0):
    from factors_of_serendipity_recommendation_b200 import dataloader, model, world
    cfg = dict(world.config)
    cfg.update(lightGCN_n_layers=n_layers, latent_dim_rec=d)
    ds = dataloader.InteractionDataset(nu, mi, u, i, device="cuda")
    if user_emb is not None:
        cfg.update(pretrain=1, user_emb=np.asarray(user_emb), item_emb=np.asarray(item_emb))
    torch.manual_seed(seed)
    return model.LightGCN(cfg, ds).cuda(), ds


def test_kat_weights_four_layers(mlls, kat):
    m, _ = make_model(mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"], 4,
                      kat["emb_user"], kat["emb_item"])
    m.eval()
    with torch.no_grad():
        lu, li = m.computer()
    assert rel_err(lu.cpu().numpy(), kat["light_users"]) <= TOL          # vs the unmodified reference
    assert rel_err(li.cpu().numpy(), kat["light_items"]) <= TOL
    assert m._flat_if_fused() is not None                                # tables are views of one [N, d] buffer
    assert set(m.state_dict().keys()) == {"embedding_user.weight", "embedding_item.weight"}


def test_synth_small_reference_outputs(synth_small):
    s = synth_small
    m, _ = make_model(int(s["n_users"]), int(s["m_items"]), s["train_user"], s["train_item"], 3, s["w_user"], s["w_item"])
    m.eval()
    with torch.no_grad():
        lu, li = m.computer()
        r = m.getUsersRating(torch.from_numpy(s["rating_users"]).cuda())
    assert rel_err(lu.cpu().numpy(), s["light_users"]) <= TOL
    assert rel_err(li.cpu().numpy(), s["light_items"]) <= TOL
    assert np.abs(r.cpu().numpy() - s["rating"]).max() <= 2e-6
    assert torch.all(lu[17] == m.embedding_user.weight[17] / 4)          # isolated user: only E0 contributes


@pytest.mark.parametrize("d,n_layers,chunk", [(64, 3, 0), (64, 1, 16), (128, 2, 32), (256, 2, 16), (32, 3, 8),
                                              (16, 2, 0), (100, 2, 16), (50, 2, 16), (192, 1, 0), (64, 0, 0)])
def test_dims_layers_and_long_rows(d, n_layers, chunk):
    from factors_of_serendipity_recommendation_b200 import _lgx, synth
    nu, mi = 400, 300
    u, i = synth.make_interactions(nu, mi, 9000, seed=d + n_layers)
    g = _lgx.Graph.build(nu, mi, torch.from_numpy(u).cuda(), torch.from_numpy(i).cuda(), chunk_nnz=chunk)
    if chunk:
        assert g.n_long > 0
    ue, ie = synth.make_embeddings(nu, mi, d, seed=5)
    E0 = torch.cat([ue, ie]).cuda()
    layers = torch.empty(max(n_layers, 1), nu + mi, d, device="cuda")
    out = g.propagate_fwd(E0, n_layers, layers_out=layers if n_layers else None)
    ref = O.OracleLightGCN(nu, mi, u, i, latent_dim=d, n_layers=n_layers, user_emb=ue, item_emb=ie)
    with torch.no_grad():
        ru, ri = ref.computer()
    assert rel_err(out.cpu().numpy(), torch.cat([ru, ri]).numpy()) <= TOL
    if n_layers:
        with torch.no_grad():
            e1 = torch.sparse.mm(ref.Graph, torch.cat([ue, ie]))
        assert rel_err(layers[0].cpu().numpy(), e1.numpy()) <= TOL
    out2 = g.propagate_fwd(E0, n_layers)          # no layers_out: ping-pong workspace path
    assert torch.equal(out, out2)                 # deterministic, bit-identical across the two paths


def test_spmm_primitive_and_row_shard():
    from factors_of_serendipity_recommendation_b200 import _lgx, synth
    nu, mi, d = 500, 700, 64
    u, i = synth.make_interactions(nu, mi, 15000, seed=9)
    g = _lgx.Graph.build(nu, mi, torch.from_numpy(u).cuda(), torch.from_numpy(i).cuda())
    e = g.export()
    X = torch.randn(nu + mi, d, device="cuda")
    Y = torch.empty_like(X)
    g.spmm(X, Y=Y)
    ref = torch.sparse.mm(g.to_torch_coo().cpu(), X.cpu())
    assert rel_err(Y.cpu().numpy(), ref.numpy()) <= TOL
    lo, hi = 300, 1000                              # row shard = the reference's _split_A_hat fold
    s, t = int(e["indptr"][lo]), int(e["indptr"][hi])
    gs = _lgx.Graph.from_csr(e["indptr"][lo:hi + 1] - s, e["indices"][s:t], e["values"][s:t], n_cols=nu + mi, chunk_nnz=32)
    Ys = torch.empty(hi - lo, d, device="cuda")
    S = torch.randn(hi - lo, d, device="cuda")
    S_out = torch.empty_like(S)
    gs.spmm(X, S_in=S, Y=Ys, S_out=S_out, div=2.0)
    assert rel_err(Ys.cpu().numpy(), ref[lo:hi].numpy()) <= TOL
    assert rel_err(S_out.cpu().numpy(), ((S.cpu() + ref[lo:hi]) / 2).numpy()) <= TOL


def test_backward_matches_autograd(mlls):
    from factors_of_serendipity_recommendation_b200 import synth
    nu, mi = mlls["n_users"], mlls["m_items"]
    ue, ie = synth.make_embeddings(nu, mi, 64, seed=3)
    m, _ = make_model(nu, mi, mlls["train_user"], mlls["train_item"], 3, ue.numpy(), ie.numpy())
    m.train()
    lu, li = m.computer()
    wu, wi = torch.randn_like(lu), torch.randn_like(li)
    ((lu * wu).sum() + (li * wi).sum()).backward()
    ref = O.OracleLightGCN(nu, mi, mlls["train_user"], mlls["train_item"], n_layers=3, user_emb=ue, item_emb=ie)
    ru, ri = ref.computer()
    ((ru * wu.cpu()).sum() + (ri * wi.cpu()).sum()).backward()
    assert rel_err(m.embedding_user.weight.grad.cpu().numpy(), ref.user_w.grad.numpy()) <= TOL
    assert rel_err(m.embedding_item.weight.grad.cpu().numpy(), ref.item_w.grad.numpy()) <= TOL
    # only one output used -> the other gradient is None inside backward
    m.zero_grad()
    lu, _ = m.computer()
    lu.square().sum().backward()
    ref.user_w.grad = ref.item_w.grad = None
    ru, _ = ref.computer()
    ru.square().sum().backward()
    assert rel_err(m.embedding_item.weight.grad.cpu().numpy(), ref.item_w.grad.numpy()) <= TOL


def test_gowalla_shape_forward():
    """configs[0]: Gowalla-shape graph, 3 layers, d=64 vs the CPU oracle."""
    from factors_of_serendipity_recommendation_b200 import synth
    nu, mi, E, d = synth.SHAPES["gowalla"]
    u, i = synth.make_interactions(nu, mi, E, seed=2020)
    ue, ie = synth.make_embeddings(nu, mi, d, seed=2020)
    m, ds = make_model(nu, mi, u, i, 3, ue.numpy(), ie.numpy())
    m.eval()
    with torch.no_grad():
        lu, li = m.computer()
    g = ds.getGraphHandle()
    assert g.nnz == 2 * E and g.n_rows == nu + mi
    ref = O.OracleLightGCN(nu, mi, u, i, n_layers=3, user_emb=ue, item_emb=ie)
    e = g.export()
    assert np.array_equal(e["indptr"].cpu().numpy(), ref.indptr) and np.array_equal(e["indices"].cpu().numpy(), ref.indices)
    assert np.array_equal(e["values"].cpu().numpy().view(np.int32), ref.data.view(np.int32))
    with torch.no_grad():
        ru, ri = ref.computer()
    assert rel_err(lu.cpu().numpy(), ru.numpy()) <= TOL and rel_err(li.cpu().numpy(), ri.numpy()) <= TOL
    # size-independent property: linearity  computer(a X) == a computer(X)
    with torch.no_grad():
        E0 = torch.cat([ue, ie]).cuda()
        a = g.propagate_fwd(E0 * 2.0, 3)
    assert torch.equal(a, torch.cat([lu, li]) * 2.0)          # scaling by 2 is exact in fp32
