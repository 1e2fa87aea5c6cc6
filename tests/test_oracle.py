"""The oracle (oracle/lightgcn_oracle.py) against the reference's golden vectors (CPU only)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import lightgcn_oracle as O
from conftest import GOLD


def ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_adjacency_matches_shipped_npz_bit_exact(mlls):
    gold = sp.load_npz(os.path.join(GOLD, "mlls_s_pre_adj_mat.npz")).tocsr()
    gold.sort_indices()
    indptr, indices, data, degree = O.build_norm_adj(mlls["n_users"], mlls["m_items"],
                                                     mlls["train_user"], mlls["train_item"])
    assert gold.shape == (2728, 2728) and gold.nnz == 127374
    assert np.array_equal(indptr, gold.indptr)
    assert np.array_equal(indices, gold.indices)
    assert gold.data.dtype == np.float32
    assert np.array_equal(data.view(np.int32), gold.data.view(np.int32))      # all 127 374 values, bit for bit
    assert degree.min() >= 4 and degree.max() == 1255


def test_adjacency_vs_reference_run_here(kat, mlls):
    # the reference's own np.power in this container is <= 1 ulp off the shipped file
    _, _, data, _ = O.build_norm_adj(mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"])
    assert ulp_diff(data, kat["graph_vals"]).max() <= 3       # two dinv factors, each <= 1 ulp off


def test_synth_small_graph_duplicates_and_isolated(synth_small):
    s = synth_small
    nu, mi = int(s["n_users"]), int(s["m_items"])
    indptr, indices, data, degree = O.build_norm_adj(nu, mi, s["train_user"], s["train_item"])
    rows = np.repeat(np.arange(nu + mi), np.diff(indptr))
    assert np.array_equal(rows, s["graph_rows"])
    assert np.array_equal(indices, s["graph_cols"])
    assert ulp_diff(data, s["graph_vals"]).max() <= 4
    # reference users_D / items_D clamp 0 -> 1 (PT/dataloader.py:290-293)
    ud = degree[:nu].astype(np.float64).copy(); ud[ud == 0] = 1
    idg = degree[nu:].astype(np.float64).copy(); idg[idg == 0] = 1
    assert np.array_equal(ud, s["users_D"]) and np.array_equal(idg, s["items_D"])
    assert degree[17] == 0 and degree[nu + 3] == 0
    # allPos = unique train items per user, ascending
    lens = np.diff(indptr)[:nu]
    assert np.array_equal(lens, s["allpos_len"])
    assert np.array_equal(indices[: indptr[nu]].astype(np.int64) - nu, s["allpos_flat"])


def test_computer_and_rating_match_reference(synth_small):
    s = synth_small
    nu, mi = int(s["n_users"]), int(s["m_items"])
    m = O.OracleLightGCN(nu, mi, s["train_user"], s["train_item"], n_layers=3,
                         user_emb=s["w_user"], item_emb=s["w_item"])
    with torch.no_grad():
        lu, li = m.computer()
        r = m.getUsersRating(torch.from_numpy(s["rating_users"]))
    assert np.abs(lu.numpy() - s["light_users"]).max() <= 1e-6 * np.abs(s["light_users"]).max()
    assert np.abs(li.numpy() - s["light_items"]).max() <= 1e-6 * np.abs(s["light_items"]).max()
    assert np.abs(r.numpy() - s["rating"]).max() <= 1e-6


def test_known_answer_mlls(kat, mlls):
    """TF/output/mlls/LightGCN.result:8 -> recall 0.16075 precision 0.10197 ndcg 0.14813."""
    m = O.OracleLightGCN(mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"],
                         n_layers=4, user_emb=kat["emb_user"], item_emb=kat["emb_item"])
    with torch.no_grad():
        lu, li = m.computer()
    assert np.abs(lu.numpy() - kat["light_users"]).max() <= 1e-5 * np.abs(kat["light_users"]).max()
    assert np.abs(li.numpy() - kat["light_items"]).max() <= 1e-5 * np.abs(kat["light_items"]).max()
    res = O.test_procedure(m, mlls["test_dict"], topks=(20,))
    assert abs(res["recall"][0] - 0.16075) < 5e-6 + 5e-6
    assert abs(res["precision"][0] - 0.10197) < 5e-6 + 5e-6
    assert abs(res["ndcg"][0] - 0.14813) < 5e-6 + 5e-6
    assert np.allclose(res["recall"], kat["recall"], atol=1e-6)
    assert np.allclose(res["precision"], kat["precision"], atol=1e-6)
    assert np.allclose(res["ndcg"], kat["ndcg"], atol=1e-6)


def test_train_step_matches_reference(train_step, mlls):
    t = train_step
    m = O.OracleLightGCN(mlls["n_users"], mlls["m_items"], mlls["train_user"], mlls["train_item"],
                         n_layers=3, user_emb=t["w0_user"], item_emb=t["w0_item"])
    opt = torch.optim.Adam([m.user_w, m.item_w], lr=float(t["lr"]))
    u, p, n = (torch.from_numpy(t[k]).long() for k in ("users", "pos", "neg"))
    loss, reg = m.bpr_loss(u, p, n)
    assert abs(loss.item() - float(t["loss"])) < 1e-6 and abs(reg.item() - float(t["reg_loss"])) < 1e-6
    total = loss + reg * float(t["decay"])
    opt.zero_grad(); total.backward()
    assert np.abs(m.user_w.grad.numpy() - t["grad_user"]).max() <= 1e-5 * np.abs(t["grad_user"]).max()
    assert np.abs(m.item_w.grad.numpy() - t["grad_item"]).max() <= 1e-5 * np.abs(t["grad_item"]).max()
    opt.step()
    assert np.abs(m.user_w.detach().numpy() - t["w1_user"]).max() <= 2e-6
    assert np.abs(m.item_w.detach().numpy() - t["w1_item"]).max() <= 2e-6


def test_topk_validity_checker():
    r = np.array([0.1, 0.9, 0.5, 0.5, 0.3], dtype=np.float32)
    assert O.topk_is_valid(r, np.array([1, 2]), 2)
    assert O.topk_is_valid(r, np.array([1, 3]), 2)          # tie at rank 2
    assert not O.topk_is_valid(r, np.array([1, 4]), 2)
    assert not O.topk_is_valid(r, np.array([2, 3]), 2)      # misses the strict top-1


def _all_pos(mlls):
    pos = [[] for _ in range(mlls["n_users"])]
    for u, i in zip(mlls["train_user"].tolist(), mlls["train_item"].tolist()):
        pos[u].append(i)
    return pos


def _chi2_two_sample(a, b):
    """Chi-square statistic for 'two histograms come from one distribution' (dof = bins - 1)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    ka, kb = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
    keep = (a + b) > 0
    return float((((ka * a - kb * b) ** 2) / (a + b))[keep].sum())


def test_native_sampler_restatement_vs_compiled_reference(mlls):
    """oracle.sample_per_user restates PT/sources/sampling.cpp:27-56; the reference's own sampler compiled into
    oracle/_ref (oracle/Makefile) honours the same contract and has the same pos / neg marginals."""
    ref = O.load_reference_sampler()
    if ref is None:
        pytest.skip("oracle/_ref sampler not built (needs /root/reference; __graft_entry__.build() makes it)")
    all_pos = _all_pos(mlls)
    nu, mi = mlls["n_users"], mlls["m_items"]
    train_num = 40 * nu + 17                                   # 40 triples per user; the remainder is dropped (:29)
    ref.seed(2020)
    S_ref = np.asarray(ref.sample_negative(nu, mi, train_num, all_pos, 1))
    assert S_ref.dtype == np.int32 and S_ref.shape == (40 * nu, 3)
    O.check_bpr_triples(S_ref, all_pos, mi, per_user=40)
    S_own = O.sample_per_user(nu, mi, train_num, all_pos, np.random.RandomState(7))
    assert S_own.shape == S_ref.shape and S_own.dtype == S_ref.dtype
    O.check_bpr_triples(S_own, all_pos, mi, per_user=40)
    hp_r, hn_r = O.sampler_marginals(S_ref, all_pos, mi)
    hp_o, hn_o = O.sampler_marginals(S_own, all_pos, mi)
    # 16 bins -> dof 15: chi2 < 37.7 holds with p = 0.999 when the distributions agree
    assert _chi2_two_sample(hp_r, hp_o) < 37.7
    assert _chi2_two_sample(hn_r, hn_o) < 37.7


def test_bpr_triple_checker_rejects_bad_samples(mlls):
    all_pos = _all_pos(mlls)
    mi = mlls["m_items"]
    good = O.sample_per_user(mlls["n_users"], mi, 2 * mlls["n_users"], all_pos, np.random.RandomState(0))
    O.check_bpr_triples(good, all_pos, mi, per_user=2)
    bad = good.copy()
    bad[5, 2] = bad[5, 1]                                      # a train item as the negative
    with pytest.raises(AssertionError):
        O.check_bpr_triples(bad, all_pos, mi)
    bad = good.copy()
    bad[3, 0] = (bad[3, 0] + 1) % mlls["n_users"]              # breaks the per-user layout (and most likely the positive)
    with pytest.raises(AssertionError):
        O.check_bpr_triples(bad, all_pos, mi, per_user=2)
