"""The rest of the reference's model API on the GPU: PureMF (PT/model.py:41-84) through getUsersRating / topk /
Procedure.Test, and a foreign, reference-style BasicDataset (only getSparseGraph() as a torch COO tensor, no
lgx handle -- the adoption path INTEGRATION.md section 1 describes) driving b200.LightGCN."""
import numpy as np
import pytest
import torch

from oracle import lightgcn_oracle as O

pytestmark = pytest.mark.gpu


def _problem(seed=3, nu=257, mi=613, E=9000, d=64):
    from factors_of_serendipity_recommendation_b200 import synth
    u, i = synth.make_interactions(nu, mi, E, seed=seed)
    ue, ie = synth.make_embeddings(nu, mi, d, seed=seed, trained_like=True)
    test_dict = synth.make_test_dict(nu, mi, u, i, per_user=4, seed=seed)
    return nu, mi, u, i, ue, ie, test_dict


def test_puremf_rating_topk_and_test_procedure():
    from factors_of_serendipity_recommendation_b200 import Procedure, dataloader, register, world
    nu, mi, u, i, ue, ie, test_dict = _problem()
    cfg = dict(world.config)
    cfg.update(latent_dim_rec=64)
    ds = dataloader.InteractionDataset(nu, mi, u, i, test_dict=test_dict, device="cuda")
    m = register.MODELS["mf"](cfg, ds).cuda().eval()
    with torch.no_grad():
        m.embedding_user.weight.copy_(ue)
        m.embedding_item.weight.copy_(ie)
    users = np.arange(nu)
    # getUsersRating == sigmoid(U I^T)  (PT/model.py:63-69)
    ref_rating = torch.sigmoid(ue @ ie.t()).numpy()
    got = m.getUsersRating(torch.from_numpy(users).cuda()).cpu().numpy()
    assert np.abs(got - ref_rating).max() <= 2e-6
    # forward(users, items) == sigmoid(<u, i>)  (PT/model.py:78-84)
    fu, fi = torch.tensor([0, 5, 9]).cuda(), torch.tensor([1, 2, 3]).cuda()
    assert np.allclose(m(fu, fi).detach().cpu().numpy(), ref_rating[[0, 5, 9], [1, 2, 3]], atol=2e-6)
    # fused top-k with the train mask
    s = (ue.double() @ ie.double().t()).numpy()
    indptr = np.concatenate([[0], np.cumsum(np.bincount(u, minlength=nu))])
    for r in range(nu):
        s[r, i[indptr[r]:indptr[r + 1]]] = -np.inf
    scale = np.abs(s[np.isfinite(s)]).max()
    for mode, tol in (("fp32", 2e-6), ("bf16x3", 1e-5), ("bf16", 1e-2)):
        idx, val = m.topk(torch.from_numpy(users).cuda(), 20, mode=mode)
        idx = idx.cpu().numpy()
        assert all(O.topk_is_valid(s[r], idx[r], 20, tol=tol * scale) for r in range(nu)), mode
    # Procedure.Test works for the MF baseline (the reference's Test only needs getUsersRating)
    world.configure(topks=[10, 20], test_u_batch_size=100)
    res = Procedure.Test(ds, m, 0, None, 0, mode="fp32")

    class _MF:                                             # the oracle's Test loop over a plain MF scorer
        def all_pos(self, batch):
            return [i[indptr[b]:indptr[b + 1]].astype(np.int64) for b in batch]

        def getUsersRating(self, batch):
            return torch.sigmoid(ue[batch] @ ie.t())

    ref = O.test_procedure(_MF(), test_dict, topks=(10, 20), u_batch_size=100)
    for k in res:
        assert np.allclose(res[k], ref[k], atol=1e-9), k
    world.configure(topks=[20])
    # the training loss of the baseline is the reference's formula (PT/model.py:71-76)
    bu = torch.tensor([0, 1, 2, 3]).cuda()
    loss, reg = m.bpr_loss(bu, torch.tensor([1, 2, 3, 4]).cuda(), torch.tensor([5, 6, 7, 8]).cuda())
    pu, pp, pn = ue[[0, 1, 2, 3]], ie[[1, 2, 3, 4]], ie[[5, 6, 7, 8]]
    ref_loss = torch.nn.functional.softplus((pu * pn).sum(1) - (pu * pp).sum(1)).mean()
    ref_reg = 0.5 * (pu.norm(2) ** 2 + pp.norm(2) ** 2 + pn.norm(2) ** 2) / 4.0
    assert abs(loss.item() - ref_loss.item()) < 1e-5 and abs(reg.item() - ref_reg.item()) < 1e-4


class _ForeignDataset:
    """What a user of the reference already has: the BasicDataset surface (PT/dataloader.py:23-67) with
    getSparseGraph() returning the coalesced torch COO tensor of D^-1/2 A D^-1/2 -- and nothing from this package."""

    def __init__(self, nu, mi, u, i, test_dict):
        self._nu, self._mi = nu, mi
        self.trainUser, self.trainItem = u, i
        self._test = test_dict
        indptr, indices, data, _ = O.build_norm_adj(nu, mi, u, i)
        self._graph = O.csr_to_torch_coo(indptr, indices, data, nu + mi)
        self._indptr = np.concatenate([[0], np.cumsum(np.bincount(u, minlength=nu))])

    @property
    def n_users(self):
        return self._nu

    @property
    def m_items(self):
        return self._mi

    @property
    def trainDataSize(self):
        return len(self.trainUser)

    @property
    def testDict(self):
        return self._test

    @property
    def allPos(self):
        return self.getUserPosItems(list(range(self._nu)))

    def getUserPosItems(self, users):
        return [self.trainItem[self._indptr[x]:self._indptr[x + 1]] for x in users]

    def getSparseGraph(self):
        return self._graph


def test_foreign_reference_style_dataset_drives_lightgcn():
    from factors_of_serendipity_recommendation_b200 import Procedure, model, world
    nu, mi, u, i, ue, ie, test_dict = _problem(seed=5)
    ds = _ForeignDataset(nu, mi, u, i, test_dict)
    cfg = dict(world.config)
    cfg.update(lightGCN_n_layers=3, latent_dim_rec=64, pretrain=1, user_emb=ue.numpy(), item_emb=ie.numpy())
    m = model.LightGCN(cfg, ds).cuda().eval()
    ref = O.OracleLightGCN(nu, mi, u, i, n_layers=3, user_emb=ue, item_emb=ie)
    with torch.no_grad():
        au, ai = m.computer()
        ru, ri = ref.computer()
    scale = max(ru.abs().max().item(), ri.abs().max().item())
    assert (au.cpu() - ru).abs().max().item() <= 1e-5 * scale
    assert (ai.cpu() - ri).abs().max().item() <= 1e-5 * scale
    # the adopted graph's user rows serve as the train mask of the fused top-k
    users = np.arange(nu)
    s = (ru.double() @ ri.double().t()).detach().numpy()
    for r, items in enumerate(ref.all_pos(users)):
        s[r, items] = -np.inf
    idx, _ = m.topk(torch.from_numpy(users).cuda(), 20, mode="bf16x3")
    idx = idx.cpu().numpy()
    sc = np.abs(s[np.isfinite(s)]).max()
    assert all(O.topk_is_valid(s[r], idx[r], 20, tol=1e-5 * sc) for r in range(nu))
    world.configure(topks=[20], test_u_batch_size=100)
    res = Procedure.Test(ds, m, 0, None, 0, mode="fp32")
    want = O.test_procedure(ref, test_dict, topks=(20,), u_batch_size=100)
    for k in res:
        assert np.allclose(res[k], want[k], atol=1e-9), k
    # one BPR step through the foreign dataset's graph: loss / reg match the oracle
    bu, bp, bn = torch.tensor([0, 1, 2, 3]), torch.tensor([u_ for u_ in i[:4]]), torch.tensor([7, 8, 9, 10])
    m.train()
    loss, reg = m.bpr_loss(bu.cuda(), bp.cuda(), bn.cuda())
    rl, rr = ref.bpr_loss(bu, bp, bn)
    assert abs(loss.item() - rl.item()) <= 1e-5 * max(1.0, abs(rl.item()))
    assert abs(reg.item() - rr.item()) <= 1e-5 * max(1.0, abs(rr.item()))


def test_embedding_export_and_candidate_buckets(tmp_path):
    """PT/main.py:31-41 (.npy export of the raw tables) and the min / max / 10-bucket labelling of
    /root/reference/recommend.py:375-380, checked against the reference's numpy expressions on the same tables."""
    from factors_of_serendipity_recommendation_b200 import dataloader, model, utils, world
    nu, mi, u, i, ue, ie, _ = _problem(seed=8, nu=301, mi=517, E=8000)
    cfg = dict(world.config)
    cfg.update(lightGCN_n_layers=2, latent_dim_rec=64, pretrain=1, user_emb=ue.numpy(), item_emb=ie.numpy())
    ds = dataloader.InteractionDataset(nu, mi, u, i, device="cuda")
    m = model.LightGCN(cfg, ds).cuda().eval()
    pu, pi = utils.export_embeddings(m, "unit", out_dir=str(tmp_path))
    emb_user, emb_item = np.load(pu), np.load(pi)
    assert pu.endswith("emb_user_unit.npy") and pi.endswith("emb_item_unit.npy")
    assert np.array_equal(emb_user, ue.numpy()) and np.array_equal(emb_item, ie.numpy())
    # the reference, verbatim (recommend.py:375-380)
    num_fold, epsilon = 10, 1e-8
    mat_dis = np.dot(emb_user, emb_item.T).astype(np.float16)
    max_dis, min_dis = np.max(mat_dis) + epsilon, np.min(mat_dis)
    inter = (max_dis - min_dis) / num_fold
    mat_label = np.floor((mat_dis - min_dis) / inter).astype(np.int8)
    got_min, got_max, got_inter, batches = utils.candidate_buckets(emb_user, emb_item, num_fold, epsilon, user_batch=128)
    assert np.float16(got_min) == np.float16(min_dis) and np.float16(got_max) == np.float16(max_dis)
    assert np.float16(got_inter) == np.float16(inter)
    labels = np.concatenate([lab.cpu().numpy() for _, lab in batches])
    assert labels.shape == mat_label.shape and labels.dtype == np.int8
    diff = labels.astype(np.int32) - mat_label.astype(np.int32)
    # the fp32 dot products are summed in a different order than numpy's BLAS: a score that sits on an fp16 rounding
    # boundary may land in the neighbouring bucket
    assert np.abs(diff).max() <= 1 and (diff != 0).mean() < 2e-3
    assert labels.min() >= 0 and labels.max() <= num_fold
