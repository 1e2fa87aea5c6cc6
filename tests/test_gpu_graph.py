"""Device graph build vs the oracle / the reference's shipped adjacency -- bit-exact (GPU)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import GOLD
from oracle import lightgcn_oracle as O

pytestmark = pytest.mark.gpu


def lgx():
    from factors_of_serendipity_recommendation_b200 import _lgx
    return _lgx


def build(nu, mi, u, i, chunk=0, on_device=True):
    L = lgx()
    tu, ti = torch.from_numpy(np.asarray(u, dtype=np.int32)), torch.from_numpy(np.asarray(i, dtype=np.int32))
    if on_device:
        tu, ti = tu.cuda(), ti.cuda()
    return L.Graph.build(nu, mi, tu, ti, chunk_nnz=chunk)


def assert_matches_oracle(g, nu, mi, u, i):
    indptr, indices, data, degree = O.build_norm_adj(nu, mi, u, i)
    e = {k: v.cpu().numpy() for k, v in g.export().items()}
    assert np.array_equal(e["indptr"], indptr)
    assert np.array_equal(e["indices"], indices)
    assert np.array_equal(e["values"].view(np.int32), data.view(np.int32))        # bit-exact fp32
    assert np.array_equal(e["degree"], degree)
    assert np.array_equal(e["dinv"].view(np.int32), O.correctly_rounded_dinv(degree).view(np.int32))
    assert np.array_equal(e["row_order"], O.degree_sorted_row_order(degree, indptr))
    return e


def test_mlls_matches_shipped_adjacency(mlls):
    g = build(mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"])
    gold = sp.load_npz(os.path.join(GOLD, "mlls_s_pre_adj_mat.npz")).tocsr()
    gold.sort_indices()
    e = g.export()
    assert g.nnz == 127374 and g.n_rows == 2728
    assert np.array_equal(e["indptr"].cpu().numpy(), gold.indptr)
    assert np.array_equal(e["indices"].cpu().numpy(), gold.indices)
    assert np.array_equal(e["values"].cpu().numpy().view(np.int32), gold.data.view(np.int32))
    assert_matches_oracle(g, mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"])


def test_host_and_device_inputs_agree(mlls):
    a = build(mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"], on_device=True).export()
    b = build(mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"], on_device=False).export()
    for k in a:
        assert torch.equal(a[k], b[k])


def test_duplicates_and_isolated_nodes(synth_small):
    s = synth_small
    nu, mi = int(s["n_users"]), int(s["m_items"])
    g = build(nu, mi, s["train_user"], s["train_item"])
    e = assert_matches_oracle(g, nu, mi, s["train_user"], s["train_item"])
    rows = np.repeat(np.arange(nu + mi), np.diff(e["indptr"]))
    assert np.array_equal(rows, s["graph_rows"]) and np.array_equal(e["indices"], s["graph_cols"])   # reference Loader
    assert e["degree"][17] == 0 and e["dinv"][17] == 0.0
    coo = g.to_torch_coo()
    assert coo.is_coalesced() and coo._nnz() == g.nnz and coo.dtype == torch.float32


@pytest.mark.parametrize("seed,chunk", [(0, 0), (1, 8), (2, 64)])
def test_random_graphs(seed, chunk):
    rng = np.random.default_rng(seed)
    nu, mi, E = int(rng.integers(5, 400)), int(rng.integers(5, 600)), int(rng.integers(1, 20000))
    u = rng.integers(0, nu, E)
    i = (rng.pareto(1.2, E) * 3).astype(np.int64) % mi          # heavy-tailed items, many duplicates
    g = build(nu, mi, u, i, chunk=chunk)
    assert_matches_oracle(g, nu, mi, u, i)
    if chunk:
        assert g.chunk_nnz == chunk
        lens = np.diff(g.export()["indptr"].cpu().numpy())
        assert g.n_long == int((lens > chunk).sum())
        assert g.n_partials == int(sum(-(-l // chunk) for l in lens if l > chunk))


def test_empty_and_invalid_inputs():
    L = lgx()
    g = build(3, 4, np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert g.nnz == 0 and g.n_rows == 7
    with pytest.raises(RuntimeError, match="outside"):
        build(3, 4, np.array([0, 3]), np.array([0, 1]))
    with pytest.raises(RuntimeError, match="outside"):
        build(3, 4, np.array([0, 1]), np.array([0, -1]))
    with pytest.raises(ValueError):
        L.Graph.build(3, 4, torch.zeros(2, dtype=torch.int32), torch.zeros(3, dtype=torch.int32))


def test_from_csr_roundtrip_and_row_shard(mlls):
    L = lgx()
    g = build(mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"])
    e = g.export()
    g2 = L.Graph.from_csr(e["indptr"], e["indices"], e["values"], n_cols=g.n_cols, n_users=g.n_users, m_items=g.m_items)
    e2 = g2.export()
    for k in ("indptr", "indices", "values", "row_order"):
        assert torch.equal(e[k], e2[k])
    # a row shard: rows [100, 900)
    lo, hi = 100, 900
    ip = e["indptr"][lo:hi + 1] - e["indptr"][lo]
    s, t = int(e["indptr"][lo]), int(e["indptr"][hi])
    gs = L.Graph.from_csr(ip, e["indices"][s:t], e["values"][s:t], n_cols=g.n_cols)
    assert gs.n_rows == hi - lo and gs.n_cols == g.n_cols and gs.nnz == t - s


def test_dataset_api(mlls, synth_small, tmp_path):
    from factors_of_serendipity_recommendation_b200 import dataloader, world
    s = synth_small
    ds = dataloader.InteractionDataset(int(s["n_users"]), int(s["m_items"]), s["train_user"], s["train_item"], device="cuda")
    assert np.array_equal(ds.users_D, s["users_D"]) and np.array_equal(ds.items_D, s["items_D"])
    ap = ds.allPos
    assert np.array_equal(np.array([len(x) for x in ap]), s["allpos_len"])
    assert np.array_equal(np.concatenate(ap), s["allpos_flat"])
    fb = ds.getUserItemFeedback([0, 0], [int(ap[0][0]), int(ap[0][0]) + 1 if int(ap[0][0]) + 1 not in ap[0] else 0])
    assert fb[0] == 1
    # Loader on the shipped dataset files + npz cache round trip (PT/dataloader.py:343,367)
    d = tmp_path / "mlls"
    d.mkdir()
    for fn in ("train.txt", "test.txt"):
        (d / fn).write_bytes(open(os.path.join(GOLD, "mlls_" + fn), "rb").read())
    ld = dataloader.Loader(config=dict(world.config), path=str(d), device="cuda")
    assert ld.n_users == 608 and ld.m_items == 2120 and ld.trainDataSize == 63687 and ld.testDataSize == 15922
    assert ld.testDict == mlls["test_dict"]
    g = ld.getGraphHandle()
    assert (d / "s_pre_adj_mat.npz").exists()
    cached = sp.load_npz(str(d / "s_pre_adj_mat.npz"))
    gold = sp.load_npz(os.path.join(GOLD, "mlls_s_pre_adj_mat.npz"))
    assert (cached != gold).nnz == 0
    ld2 = dataloader.Loader(config=dict(world.config), path=str(d), device="cuda")     # now reads the cache
    assert torch.equal(ld2.getGraphHandle().export()["values"], g.export()["values"])
    G = ld.getSparseGraph()
    assert G.is_sparse and G.shape == (2728, 2728) and G.device.type == "cuda"
