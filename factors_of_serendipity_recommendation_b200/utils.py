"""Training utilities behind the reference's names (PT/utils.py): BPRLoss.stageOne, samplers,
minibatch / shuffle, timer, ranking metrics."""
from __future__ import annotations

import os
import time as _time

import numpy as np
import torch
from torch import optim

from . import _lgx, world


class BPRLoss:
    """PT/utils.py:34-52.  Adam over the model's parameters; stageOne = one optimisation step.

    ``sync=False`` returns the loss as a 0-d device tensor instead of a Python float, removing the
    per-mini-batch host sync of the reference (PT/utils.py:52)."""

    def __init__(self, recmodel, config: dict):
        self.model = recmodel
        self.weight_decay = config["decay"]
        self.lr = config["lr"]
        # the fused Adam needs the model's single [N, d] parameter buffer on a CUDA device (LightGCN after .cuda());
        # any other model (PureMF, a model still on the host) keeps torch.optim.Adam like the reference
        flat = recmodel._flat_if_fused() if getattr(recmodel, "_flat_if_fused", None) is not None else None
        self.fused = bool(config.get("fused_adam", False)) and flat is not None and flat.is_cuda
        if self.fused:
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
            self._state = torch.zeros(4, dtype=torch.int32, device=flat.device)   # device-side Adam step counter
            self.opt = None
        else:
            self.opt = optim.Adam(recmodel.parameters(), lr=self.lr)
        self._graphed = None

    def stageOne(self, users, pos, neg, sync: bool = True):
        loss, reg_loss = self.model.bpr_loss(users, pos, neg)
        loss = loss + reg_loss * self.weight_decay
        if self.fused:
            wu, wi = self.model.embedding_user.weight, self.model.embedding_item.weight
            wu.grad = wi.grad = None
            loss.backward()
            flat = self.model._flat_if_fused()
            gu, gi = wu.grad, wi.grad
            # the fused backward returns both gradients as views of ONE [N, d] buffer: same storage, adjacent, and the
            # storage really covers both (two separately allocated blocks can be adjacent by accident)
            st_u, st_i = gu.untyped_storage(), gi.untyped_storage()
            if (gu.is_contiguous() and gi.is_contiguous() and st_u.data_ptr() == st_i.data_ptr()
                    and gu.data_ptr() + gu.numel() * 4 == gi.data_ptr()
                    and gu.data_ptr() - st_u.data_ptr() + flat.numel() * 4 <= st_u.nbytes()):
                grad = torch.as_strided(gu, (flat.shape[0], flat.shape[1]), (flat.shape[1], 1))
            else:
                grad = torch.cat([gu, gi])
            _lgx.adam_step_dev(flat, grad, self._m, self._v, self.lr, 0.9, 0.999, 1e-8, self._state)
            self.model._eval_cache = None
            self.model._packed = {}
        else:
            self.opt.zero_grad()
            loss.backward()
            self.opt.step()
        return loss.cpu().item() if sync else loss.detach()


class GraphedStageOne:
    """Whole-step CUDA graph of BPRLoss.stageOne (bpr_loss forward, backward, fused Adam) for one batch size.
    Needs the fused Adam (no host-side step value) and no edge dropout (its seed is a host value per call).
    Call .run(users, pos, neg) -> 0-d device tensor with the step's loss."""

    def __init__(self, bpr: BPRLoss, users, pos, neg):
        if not bpr.fused:
            raise RuntimeError("CUDA-graph capture of the BPR step needs config['fused_adam'] = True")
        if bpr.model.config.get("dropout") and bpr.model.training:
            raise RuntimeError("edge dropout draws a new host-side seed per call: not graph-capturable")
        self.bpr = bpr
        self.u, self.p, self.n = users.clone(), pos.clone(), neg.clone()
        side = torch.cuda.Stream(device=users.device)
        side.wait_stream(torch.cuda.current_stream(users.device))
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(self.graph, stream=side):
                self.loss = bpr.stageOne(self.u, self.p, self.n, sync=False)
        torch.cuda.current_stream(users.device).wait_stream(side)

    def run(self, users, pos, neg):
        self.u.copy_(users)
        self.p.copy_(pos)
        self.n.copy_(neg)
        self.graph.replay()
        # the replayed Adam kernel writes the parameter buffer through a raw pointer (no autograd version bump): drop the
        # model's cached propagation / packed operands here, or computer() / topk() after graph steps would be stale
        self.bpr.model._eval_cache = None
        self.bpr.model._packed = {}
        return self.loss


def UniformSample_original(dataset, neg_ratio=1, seed=None):
    """PT/utils.py:55-64.  With a device graph the triples are drawn by lgx_sample_bpr (counter-based
    RNG, rejection by binary search in the sorted train row) with the Python sampler's distribution:
    trainDataSize users drawn uniformly with replacement.  Returns int64 [S, 3] (device tensor)."""
    if hasattr(dataset, "getGraphHandle"):
        g = dataset.getGraphHandle()
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        return g.sample_bpr(dataset.trainDataSize, per_user=0, seed=seed)
    return UniformSample_original_python(dataset)


def UniformSample_original_python(dataset):
    """PT/utils.py:67-99, host sampler with numpy's global RNG (same call order as the reference)."""
    users = np.random.randint(0, dataset.n_users, dataset.trainDataSize)
    all_pos = dataset.allPos
    triples = []
    for user in users:
        pos_for_user = all_pos[user]
        if len(pos_for_user) == 0:
            continue
        positem = pos_for_user[np.random.randint(0, len(pos_for_user))]
        negitem = np.random.randint(0, dataset.m_items)
        while negitem in pos_for_user:
            negitem = np.random.randint(0, dataset.m_items)
        triples.append([user, positem, negitem])
    return np.array(triples)


def set_seed(seed):
    np.random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.manual_seed(seed)


def getFileName():
    if world.model_name == "mf":
        file = f"mf-{world.dataset}-{world.config['latent_dim_rec']}.pth.tar"
    else:
        file = f"lgn-{world.dataset}-{world.config['lightGCN_n_layers']}-{world.config['latent_dim_rec']}.pth.tar"
    return os.path.join(world.PATH, file)


def minibatch(*tensors, **kwargs):
    """PT/utils.py:121-130."""
    batch_size = kwargs.get("batch_size", world.config["bpr_batch_size"])
    n = len(tensors[0])
    for i in range(0, n, batch_size):
        if len(tensors) == 1:
            yield tensors[0][i:i + batch_size]
        else:
            yield tuple(x[i:i + batch_size] for x in tensors)


def shuffle(*arrays, **kwargs):
    """PT/utils.py:133-151: one numpy permutation applied to every array."""
    if len(set(len(x) for x in arrays)) != 1:
        raise ValueError("All inputs to shuffle must have the same length.")
    perm = np.arange(len(arrays[0]))
    np.random.shuffle(perm)
    idx = perm
    if torch.is_tensor(arrays[0]):
        idx = torch.from_numpy(perm).to(arrays[0].device)
    result = arrays[0][idx] if len(arrays) == 1 else tuple(x[idx] for x in arrays)
    return (result, perm) if kwargs.get("indices", False) else result


class timer:
    """PT/utils.py:154-213: wall-clock context manager with a global named tape."""
    TAPE = [-1]
    NAMED_TAPE = {}

    @staticmethod
    def get():
        return timer.TAPE.pop() if len(timer.TAPE) > 1 else -1

    @staticmethod
    def dict(select_keys=None):
        keys = timer.NAMED_TAPE.keys() if select_keys is None else select_keys
        return "|" + "".join(f"{k}:{timer.NAMED_TAPE[k]:.2f}|" for k in keys)

    @staticmethod
    def zero(select_keys=None):
        for k in (timer.NAMED_TAPE.keys() if select_keys is None else select_keys):
            timer.NAMED_TAPE[k] = 0

    def __init__(self, tape=None, **kwargs):
        self.named = kwargs.get("name") or False
        if self.named:
            timer.NAMED_TAPE.setdefault(self.named, 0.0)
        else:
            self.tape = tape or timer.TAPE

    def __enter__(self):
        self.start = _time.time()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self.named:
            timer.NAMED_TAPE[self.named] += _time.time() - self.start
        else:
            self.tape.append(_time.time() - self.start)


# ------------------------------------------------------------------------------------ metrics
def RecallPrecision_ATk(test_data, r, k):
    """PT/utils.py:218-229: sums over the batch (the caller divides by the number of users)."""
    right_pred = r[:, :k].sum(1)
    recall_n = np.array([len(t) for t in test_data])
    return {"recall": np.sum(right_pred / recall_n), "precision": np.sum(right_pred) / k}


def MRRatK_r(r, k):
    """PT/utils.py:232-240 (kept as written there: divides by log2(1/rank))."""
    pred = r[:, :k] / np.log2(1.0 / np.arange(1, k + 1))
    return np.sum(pred.sum(1))


def NDCGatK_r(test_data, r, k):
    """PT/utils.py:243-262: binary gains, IDCG over min(k, |ground truth|)."""
    assert len(r) == len(test_data)
    discounts = 1.0 / np.log2(np.arange(2, k + 2))
    lengths = np.minimum(np.array([len(t) for t in test_data]), k)
    ideal = (np.arange(k)[None, :] < lengths[:, None]).astype(np.float64)
    idcg = (ideal * discounts).sum(1)
    dcg = (r[:, :k] * discounts).sum(1)
    idcg[idcg == 0.0] = 1.0
    ndcg = dcg / idcg
    ndcg[np.isnan(ndcg)] = 0.0
    return np.sum(ndcg)


def AUC(all_item_scores, dataset, test_data):
    """PT/utils.py:265-274."""
    from sklearn.metrics import roc_auc_score

    r_all = np.zeros((dataset.m_items,))
    r_all[test_data] = 1
    keep = all_item_scores >= 0
    return roc_auc_score(r_all[keep], all_item_scores[keep])


def getLabel(test_data, pred_data):
    """PT/utils.py:277-285: r[i, j] = 1 iff the j-th prediction of user i is in the ground truth."""
    out = np.zeros((len(test_data), np.asarray(pred_data).shape[1]), dtype="float")
    for i, truth in enumerate(test_data):
        out[i] = np.isin(np.asarray(pred_data[i]), np.asarray(truth))
    return out


# ------------------------------------------------------------------------------------------------
def export_embeddings(model, dataset_name: str | None = None, out_dir: str = "."):
    """PT/main.py:31-41 (the --load branch): write the two RAW embedding tables as emb_user_<dataset>.npy /
    emb_item_<dataset>.npy, the files the serendipity pipeline reads back (recommend.py:363-364).
    -> (path_user, path_item)"""
    import os
    name = dataset_name if dataset_name is not None else world.dataset
    pu = os.path.join(out_dir, f"emb_user_{name}.npy")
    pi = os.path.join(out_dir, f"emb_item_{name}.npy")
    np.save(pu, model.embedding_user.weight.detach().cpu().numpy())
    np.save(pi, model.embedding_item.weight.detach().cpu().numpy())
    return pu, pi


def candidate_buckets(emb_user, emb_item, num_fold: int = 10, epsilon: float = 1e-8, user_batch: int = 8192, device="cuda"):
    """recommend.py:375-380 on the GPU: yields (first_user, labels int8 [b, n_item]) per user batch plus the
    (min_dis, max_dis, inter) triple, without ever holding the [n_user, n_item] matrix.
    -> (min_dis, max_dis, inter, generator of (start, labels))"""
    from . import _lgx
    U = torch.as_tensor(np.asarray(emb_user), dtype=torch.float32).to(device).contiguous()
    I = torch.as_tensor(np.asarray(emb_item), dtype=torch.float32).to(device).contiguous()
    mm = _lgx.score_minmax(U, I).cpu().numpy().astype(np.float16)
    min_dis = mm[0]
    max_dis = mm[1] + epsilon                     # np.float16 + python float, as the reference writes it
    inter = (max_dis - min_dis) / num_fold

    def batches():
        for s in range(0, U.shape[0], user_batch):
            yield s, _lgx.score_bucket(U[s:s + user_batch].contiguous(), None, I, float(np.float16(min_dis)), float(np.float16(inter)))
    return min_dis, max_dis, inter, batches()
