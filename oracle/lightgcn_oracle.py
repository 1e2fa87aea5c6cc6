"""CPU oracle for the LightGCN hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference's PyTorch LightGCN
path (PT/ = /root/reference/lightGCN/LightGCN-PyTorch-master/code/).  It is the
checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it; the product package never does (it fails loudly when the CUDA
library is missing instead of falling back to this).

The arithmetic itself lives in third-party wheels that the reference pins
(``torch==2.1.0`` TOP/requirements.txt:162, ``scipy==1.10.1`` :140); this container
has torch 2.11 / scipy 1.18.  The same library calls are used here
(``torch.sparse.mm``, ``torch.matmul``, ``torch.topk``), so the oracle is the
reference's own code path minus its module-level globals (``world``) and file IO.

Parity is PINNED (not "parity unpinned") by:
  * tests/golden/mlls_s_pre_adj_mat.npz   -- the adjacency shipped with the reference
    (TF/Data/mlls/s_pre_adj_mat.npz): ``build_norm_adj`` reproduces indptr, indices
    and all 127 374 fp32 values bit-exactly (tests/test_oracle.py).
  * tests/golden/mlls_kat.npz             -- outputs of the UNMODIFIED reference model
    imported from /root/reference (oracle/make_golden.py): computer(), ratings,
    top-20, Test() metrics; matches TF/output/mlls/LightGCN.result:8.
  * tests/golden/mlls_train_step.npz      -- reference bpr_loss / backward / Adam step.
  * tests/golden/synth_small.npz          -- reference Loader + model on a synthetic graph
    with duplicate edges and isolated nodes.
  * oracle/_ref/sampling*.so              -- the reference's OWN native BPR sampler, compiled from its single
    source file where it lies under /root/reference (oracle/Makefile); ``sample_per_user`` and the device
    sampler are checked against it (contract + distribution; its rand() stream cannot be matched).
"""
from __future__ import annotations

import numpy as np
import torch

MASK_VALUE = -(1 << 10)  # PT/Procedure.py:134


# --------------------------------------------------------------------------- graph
def correctly_rounded_dinv(deg: np.ndarray) -> np.ndarray:
    """fp32 d^-1/2 with inf -> 0 (PT/dataloader.py:357-359).

    The reference computes ``np.power(rowsum_f32, -0.5)``; whether that is correctly
    rounded depends on the numpy build (numpy 2.3.5 here is 1 ulp off on ~1/3 of the
    entries).  The adjacency shipped with the reference equals the correctly rounded
    value, i.e. fp64 pow rounded to fp32, so that is the contract.
    """
    deg = np.asarray(deg, dtype=np.float64)
    with np.errstate(divide="ignore"):
        dinv = np.power(deg, -0.5)
    dinv[np.isinf(dinv)] = 0.0
    return dinv.astype(np.float32)


def build_user_item_net(n_users: int, m_items: int, train_user, train_item):
    """UserItemNet = csr(ones, (u, i)); duplicate pairs are SUMMED (PT/dataloader.py:288-289).

    Returns (indptr int64[n_users+1], items int32[nnz_unique], counts int32[nnz_unique]).
    """
    u = np.asarray(train_user, dtype=np.int64)
    i = np.asarray(train_item, dtype=np.int64)
    key = u * np.int64(m_items) + i
    uniq, cnt = np.unique(key, return_counts=True)
    uu = (uniq // m_items).astype(np.int64)
    ii = (uniq % m_items).astype(np.int32)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.add.at(indptr, uu + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, ii, cnt.astype(np.int32)


def build_norm_adj(n_users: int, m_items: int, train_user, train_item):
    """D^-1/2 A D^-1/2 with A = [[0, R], [R^T, 0]] as canonical CSR (PT/dataloader.py:349-364).

    Returns (indptr int64[N+1], indices int32[nnz], data float32[nnz], degree int64[N]).
    Rows ascending, columns ascending, duplicates merged with value = multiplicity.
    data[k] = fl32(fl32(dinv[row] * a) * dinv[col])  (d_mat.dot(adj).dot(d_mat), fp32).
    degree = row sums of A (multiplicities counted), NOT clamped to 1.
    """
    N = n_users + m_items
    u = np.asarray(train_user, dtype=np.int64)
    i = np.asarray(train_item, dtype=np.int64)
    rows = np.concatenate([u, i + n_users])
    cols = np.concatenate([i + n_users, u])
    key = rows * np.int64(N) + cols
    uniq, cnt = np.unique(key, return_counts=True)
    r = uniq // N
    c = (uniq % N).astype(np.int32)
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(indptr, r + 1, 1)
    indptr = np.cumsum(indptr)
    degree = np.zeros(N, dtype=np.int64)
    np.add.at(degree, r, cnt)
    dinv = correctly_rounded_dinv(degree)
    a = cnt.astype(np.float32)
    data = (dinv[r] * a).astype(np.float32) * dinv[c]
    return indptr, c, data.astype(np.float32), degree


def degree_sorted_row_order(degree: np.ndarray, indptr: np.ndarray) -> np.ndarray:
    """Stable descending order of rows by stored-nnz count (the engine's scheduling order)."""
    nnz_per_row = np.diff(indptr)
    return np.argsort(-nnz_per_row, kind="stable").astype(np.int32)


def csr_to_torch_coo(indptr, indices, data, N: int) -> torch.Tensor:
    """PT/dataloader.py:331-337 + :374 -- coalesced fp32 COO with int64 indices."""
    rows = np.repeat(np.arange(N, dtype=np.int64), np.diff(indptr))
    index = torch.from_numpy(np.stack([rows, np.asarray(indices, dtype=np.int64)]))
    val = torch.from_numpy(np.asarray(data, dtype=np.float32))
    return torch.sparse_coo_tensor(index, val, (N, N)).coalesce()


# --------------------------------------------------------------------------- model
def computer(graph: torch.Tensor, user_w: torch.Tensor, item_w: torch.Tensor, n_layers: int):
    """LightGCN.computer (PT/model.py:145-177), dropout off."""
    all_emb = torch.cat([user_w, item_w])
    embs = [all_emb]
    for _ in range(n_layers):
        all_emb = torch.sparse.mm(graph, all_emb)
        embs.append(all_emb)
    embs = torch.stack(embs, dim=1)
    light_out = torch.mean(embs, dim=1)
    return torch.split(light_out, [user_w.shape[0], item_w.shape[0]])


def users_rating(all_users: torch.Tensor, all_items: torch.Tensor, users: torch.Tensor):
    """LightGCN.getUsersRating (PT/model.py:179-184) given computer() output."""
    users_emb = all_users[users.long()]
    return torch.sigmoid(torch.matmul(users_emb, all_items.t()))


def mask_and_topk(rating: torch.Tensor, all_pos, k: int):
    """PT/Procedure.py:129-135: train items -> -1024 (after sigmoid), torch.topk."""
    exclude_index, exclude_items = [], []
    for range_i, items in enumerate(all_pos):
        exclude_index.extend([range_i] * len(items))
        exclude_items.extend(np.asarray(items).tolist())
    rating = rating.clone()
    if exclude_index:
        rating[exclude_index, exclude_items] = MASK_VALUE
    vals, idx = torch.topk(rating, k=k)
    return rating, vals, idx


def bpr_loss(all_users, all_items, user_w, item_w, users, pos, neg):
    """LightGCN.getEmbedding + bpr_loss (PT/model.py:186-209)."""
    users, pos, neg = users.long(), pos.long(), neg.long()
    users_emb, pos_emb, neg_emb = all_users[users], all_items[pos], all_items[neg]
    u0, p0, n0 = user_w[users], item_w[pos], item_w[neg]
    reg_loss = (1 / 2) * (u0.norm(2).pow(2) + p0.norm(2).pow(2) + n0.norm(2).pow(2)) / float(len(users))
    pos_scores = torch.sum(torch.mul(users_emb, pos_emb), dim=1)
    neg_scores = torch.sum(torch.mul(users_emb, neg_emb), dim=1)
    loss = torch.mean(torch.nn.functional.softplus(neg_scores - pos_scores))
    return loss, reg_loss


def forward_pairs(all_users, all_items, users, items):
    """LightGCN.forward (PT/model.py:211-220)."""
    return torch.sum(all_users[users.long()] * all_items[items.long()], dim=1)


class OracleLightGCN:
    """A CPU LightGCN with the reference's parameter names, used by bench.py's
    cpu_baseline / --impl reference legs and by the parity tests."""

    def __init__(self, n_users, m_items, train_user, train_item, latent_dim=64, n_layers=3,
                 user_emb=None, item_emb=None, seed=2020):
        self.n_users, self.m_items, self.n_layers = n_users, m_items, n_layers
        indptr, indices, data, degree = build_norm_adj(n_users, m_items, train_user, train_item)
        self.indptr, self.indices, self.data, self.degree = indptr, indices, data, degree
        self.Graph = csr_to_torch_coo(indptr, indices, data, n_users + m_items)
        if user_emb is None:
            g = torch.Generator().manual_seed(seed)
            user_emb = torch.empty(n_users, latent_dim).normal_(std=0.1, generator=g)  # PT/model.py:112-113
            item_emb = torch.empty(m_items, latent_dim).normal_(std=0.1, generator=g)
        self.user_w = torch.as_tensor(user_emb, dtype=torch.float32).clone().requires_grad_(True)
        self.item_w = torch.as_tensor(item_emb, dtype=torch.float32).clone().requires_grad_(True)

    def computer(self):
        return computer(self.Graph, self.user_w, self.item_w, self.n_layers)

    def getUsersRating(self, users):
        au, ai = self.computer()  # the reference recomputes the propagation per call (PT/model.py:180)
        return users_rating(au, ai, users)

    def bpr_loss(self, users, pos, neg):
        au, ai = self.computer()
        return bpr_loss(au, ai, self.user_w, self.item_w, users, pos, neg)

    def all_pos(self, users):
        """Loader.getUserPosItems (PT/dataloader.py:404-408) from the user rows of the adjacency."""
        out = []
        for u in users:
            s, e = self.indptr[u], self.indptr[u + 1]
            out.append(self.indices[s:e].astype(np.int64) - self.n_users)
        return out


def stage_one(model: OracleLightGCN, opt: torch.optim.Optimizer, users, pos, neg, decay: float) -> float:
    """utils.BPRLoss.stageOne (PT/utils.py:43-52)."""
    loss, reg = model.bpr_loss(users, pos, neg)
    loss = loss + reg * decay
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss.cpu().item()


# --------------------------------------------------------------------------- metrics
def get_label(test_data, pred_data):
    """utils.getLabel (PT/utils.py:277-285)."""
    r = []
    for i in range(len(test_data)):
        ground = test_data[i]
        pred = np.array([x in ground for x in pred_data[i]]).astype("float")
        r.append(pred)
    return np.array(r).astype("float")


def recall_precision_at_k(test_data, r, k):
    """utils.RecallPrecision_ATk (PT/utils.py:218-229)."""
    right_pred = r[:, :k].sum(1)
    recall_n = np.array([len(test_data[i]) for i in range(len(test_data))])
    return {"recall": np.sum(right_pred / recall_n), "precision": np.sum(right_pred) / k}


def ndcg_at_k(test_data, r, k):
    """utils.NDCGatK_r (PT/utils.py:243-262)."""
    pred_data = r[:, :k]
    test_matrix = np.zeros((len(pred_data), k))
    for i, items in enumerate(test_data):
        length = k if k <= len(items) else len(items)
        test_matrix[i, :length] = 1
    idcg = np.sum(test_matrix * 1.0 / np.log2(np.arange(2, k + 2)), axis=1)
    dcg = np.sum(pred_data * (1.0 / np.log2(np.arange(2, k + 2))), axis=1)
    idcg[idcg == 0.0] = 1.0
    ndcg = dcg / idcg
    ndcg[np.isnan(ndcg)] = 0.0
    return np.sum(ndcg)


def test_procedure(model: OracleLightGCN, test_dict: dict, topks=(20,), u_batch_size=100):
    """Procedure.Test (PT/Procedure.py:96-174), single core, no tensorboard."""
    max_k = max(topks)
    results = {m: np.zeros(len(topks)) for m in ("precision", "recall", "ndcg")}
    users = list(test_dict.keys())
    with torch.no_grad():
        for s in range(0, len(users), u_batch_size):
            batch = users[s:s + u_batch_size]
            all_pos = model.all_pos(batch)
            ground = [test_dict[u] for u in batch]
            rating = model.getUsersRating(torch.tensor(batch, dtype=torch.long))
            _, _, idx = mask_and_topk(rating, all_pos, max_k)
            r = get_label(ground, idx.numpy())
            for j, k in enumerate(topks):
                ret = recall_precision_at_k(ground, r, k)
                results["recall"][j] += ret["recall"]
                results["precision"][j] += ret["precision"]
                results["ndcg"][j] += ndcg_at_k(ground, r, k)
    for m in results:
        results[m] /= float(len(users))
    return results


# --------------------------------------------------------------------------- sampler
def uniform_sample_python(n_users, m_items, train_size, all_pos, rng: np.random.RandomState):
    """utils.UniformSample_original_python (PT/utils.py:67-99)."""
    users = rng.randint(0, n_users, train_size)
    S = []
    for user in users:
        pos_for_user = all_pos[user]
        if len(pos_for_user) == 0:
            continue
        positem = pos_for_user[rng.randint(0, len(pos_for_user))]
        while True:
            negitem = rng.randint(0, m_items)
            if negitem in pos_for_user:
                continue
            break
        S.append([user, positem, negitem])
    return np.array(S)


def sample_per_user(n_users, m_items, train_size, all_pos, rng: np.random.RandomState, neg_num: int = 1):
    """Semantics of the reference's native sampler, sample_negative (PT/sources/sampling.cpp:27-56): EVERY user gets
    exactly train_size // n_users rows [user, pos, neg...], pos uniform over the user's positives, each neg uniform
    over the items that are not (rejection, :47-51).  The reference draws with rand() % n; the stream cannot be
    matched, the distribution can (tests compare against the compiled reference in oracle/_ref when present)."""
    per_user = train_size // n_users
    S = np.empty((n_users * per_user, 2 + neg_num), dtype=np.int32)
    row = 0
    for user in range(n_users):
        pos_for_user = all_pos[user]
        for _ in range(per_user):
            S[row, 0] = user
            S[row, 1] = pos_for_user[rng.randint(0, len(pos_for_user))]
            for j in range(neg_num):
                while True:
                    negitem = rng.randint(0, m_items)
                    if negitem not in pos_for_user:
                        break
                S[row, 2 + j] = negitem
            row += 1
    return S


def load_reference_sampler():
    """The reference's own pybind11 sampler compiled by oracle/Makefile into oracle/_ref/ (built where
    /root/reference exists; the .so travels to the GPU box).  None when it has not been built."""
    import glob
    import importlib.util
    import os
    hits = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "sampling*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("sampling", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def check_bpr_triples(S: np.ndarray, all_pos, m_items: int, per_user: int | None = None) -> None:
    """Contract every BPR sampler on this path honours (PT/utils.py:76-95, PT/sources/sampling.cpp:29-53):
    pos is a train item of the user, neg is not, ids in range; per_user: the native sampler's exact user layout."""
    S = np.asarray(S)
    assert S.ndim == 2 and S.shape[1] >= 3
    pos_sets = [set(int(x) for x in p) for p in all_pos]
    assert S[:, 0].min() >= 0 and S[:, 0].max() < len(all_pos)
    assert S[:, 1:].min() >= 0 and S[:, 1:].max() < m_items
    for row in S:
        ps = pos_sets[int(row[0])]
        assert int(row[1]) in ps, "positive is not a train item of its user"
        assert all(int(n) not in ps for n in row[2:]), "negative is a train item of its user"
    if per_user is not None:
        assert np.array_equal(S[:, 0], np.repeat(np.arange(len(all_pos)), per_user)), "per-user layout"


def sampler_marginals(S: np.ndarray, all_pos, m_items: int, n_bins: int = 16):
    """Two distribution summaries that do not depend on the RNG stream:
    * histogram of the rank of pos inside its user's sorted positives, normalised to [0,1) -> uniform if pos ~ U(allPos[u]);
    * histogram of neg over equal-width item-id bins (uniform over the user's non-positives, pooled over users)."""
    S = np.asarray(S)
    sorted_pos = [np.sort(np.asarray(p)) for p in all_pos]
    frac = np.array([(np.searchsorted(sorted_pos[int(u)], int(p)) + 0.5) / len(sorted_pos[int(u)]) for u, p in S[:, :2]])
    h_pos = np.histogram(frac, bins=n_bins, range=(0.0, 1.0))[0]
    h_neg = np.histogram(S[:, 2], bins=n_bins, range=(0, m_items))[0]
    return h_pos, h_neg


# --------------------------------------------------------------------------- parity helpers
def topk_is_valid(ref_rating_row: np.ndarray, idx: np.ndarray, k: int, tol: float = 0.0) -> bool:
    """'identical up to ties' (SURVEY.md section 8c): every returned index scores at least the
    reference k-th value (minus tol) and every item strictly above it (plus tol) is returned."""
    kth = np.partition(ref_rating_row, -k)[-k]
    if len(set(idx.tolist())) != k:
        return False
    if np.any(ref_rating_row[idx] < kth - tol):
        return False
    must = np.nonzero(ref_rating_row > kth + tol)[0]
    return bool(np.isin(must, idx).all())
