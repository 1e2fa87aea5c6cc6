"""LightGCN / PureMF behind the reference's model API (PT/model.py), executed by liblgx on a B200.

Same constructor, attributes and methods as the reference:
    LightGCN(config, dataset); .embedding_user / .embedding_item (state_dict keys unchanged);
    computer() -> (users, items); getUsersRating(users) -> sigmoid(U I^T) [B, m_items];
    getEmbedding; bpr_loss(users, pos, neg) -> (loss, reg_loss); forward(users, items).
Additive fast path: topk(users, k) = fused score + train mask + top-K (PT/Procedure.py:127-135).
There is no CPU path: the model refuses to run off a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import _lgx, world
from .dataloader import BasicDataset


class BasicModel(nn.Module):
    def __init__(self):
        super().__init__()

    def getUsersRating(self, users):
        raise NotImplementedError


class PairWiseModel(BasicModel):
    def __init__(self):
        super().__init__()

    def bpr_loss(self, users, pos, neg):
        raise NotImplementedError


# --------------------------------------------------------------------------------- autograd glue
class _Propagate(torch.autograd.Function):
    """computer(): forward = L fused SpMM layers; backward = the same kernels on the gradient
    (A_hat symmetric, Horner form) -- replaces SparseAddmmBackward0 + the per-call COO re-sort."""

    @staticmethod
    def forward(ctx, flat_w, user_w, item_w, graph, n_layers, dropout=None):
        E0 = flat_w if flat_w is not None else torch.cat([user_w, item_w])
        out = graph.propagate_fwd(E0.detach(), n_layers, dropout=dropout)
        ctx.graph, ctx.n_layers, ctx.n_users, ctx.dropout = graph, n_layers, user_w.shape[0], dropout
        users, items = out[: ctx.n_users], out[ctx.n_users:]
        return users, items

    @staticmethod
    def backward(ctx, g_users, g_items):
        n_users = ctx.n_users
        d = (g_users if g_users is not None else g_items).shape[1]
        ref = g_users if g_users is not None else g_items
        g = torch.zeros(ctx.graph.n_rows, d, dtype=torch.float32, device=ref.device)
        scale = 1.0 / (ctx.n_layers + 1)                 # backward of torch.mean over the L+1 layers
        if g_users is not None:
            torch.mul(g_users, scale, out=g[:n_users])
        if g_items is not None:
            torch.mul(g_items, scale, out=g[n_users:])
        dE0 = ctx.graph.propagate_bwd(g, ctx.n_layers, dropout=ctx.dropout)
        return None, dE0[:n_users], dE0[n_users:], None, None, None


class _BprLoss(torch.autograd.Function):
    """bpr_loss fused: propagate + 6 gathers + dots + softplus + reg in two kernels; backward =
    scatter of the row gradients + Horner propagate + reg scatter, no host sync."""

    @staticmethod
    def forward(ctx, flat_w, user_w, item_w, graph, n_layers, users, pos, neg, dropout=None):
        E0 = (flat_w if flat_w is not None else torch.cat([user_w, item_w])).detach()
        n_users = user_w.shape[0]
        light = graph.propagate_fwd(E0, n_layers, dropout=dropout)
        out2, coef = _lgx.bpr_forward(light, E0, users, pos, neg, n_users)
        ctx.graph, ctx.n_layers, ctx.n_users, ctx.dropout = graph, n_layers, n_users, dropout
        ctx.save_for_backward(light, E0, users, pos, neg, coef)
        return out2[0], out2[1]

    @staticmethod
    def backward(ctx, g_loss, g_reg):
        light, E0, users, pos, neg, coef = ctx.saved_tensors
        n_users, L = ctx.n_users, ctx.n_layers
        dE0 = torch.zeros_like(E0)
        if g_loss is not None:
            G = torch.zeros_like(E0)
            _lgx.bpr_backward_light(light, users, pos, neg, coef, n_users, 1.0 / (L + 1),
                                    g_loss.to(torch.float32).contiguous(), G)
            ctx.graph.propagate_bwd(G, L, out=dE0, dropout=ctx.dropout)
        if g_reg is not None:
            _lgx.bpr_backward_reg(E0, users, pos, neg, n_users, 1.0, g_reg.to(torch.float32).contiguous(), dE0)
        return None, dE0[:n_users], dE0[n_users:], None, None, None, None, None, None


def _as_index(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t))
    return t.to(device=device, dtype=torch.int64).contiguous()


def tc_mode_for(mode_id: int, d: int, k: int) -> int:
    """The tcgen05 tile covers d % 64 == 0, k <= 32, d <= 256 (bf16) / 128 (bf16x3); every other shape runs the exact
    fp32 CUDA-core kernel (about 15x slower per score -- a warning is printed once per shape)."""
    if mode_id != _lgx.SCORE_FP32 and (d % 64 != 0 or k > 32 or (mode_id == _lgx.SCORE_BF16X3 and d > 128) or d > 256):
        key = (mode_id, d, k)
        if key not in _warned_fp32:
            _warned_fp32.add(key)
            world.cprint(f"[lgx] scoring d={d}, k={k} is outside the tensor-core tile (d % 64 == 0, k <= 32, d <= 256): "
                         "using the fp32 CUDA-core kernel (~15x slower per score)")
        return _lgx.SCORE_FP32
    return mode_id


_warned_fp32 = set()


def fused_topk(graph, all_users, all_items, users, k: int, mode_id: int, packed_items=None):
    """score + train mask (graph's user rows, or None) + top-k for the batch `users` -> (idx, val, packed item operand)."""
    d = all_items.shape[1]
    mode_id = tc_mode_for(mode_id, d, k)
    if mode_id == _lgx.SCORE_FP32:
        U_op, I_op = all_users.index_select(0, users), all_items
    else:
        I_op = packed_items if packed_items is not None else _lgx.pack_operand(all_items, None, mode_id, True)
        U_op = _lgx.pack_operand(all_users, users, mode_id, False)
    idx, val = _lgx.score_topk(graph, U_op, users, I_op, d, k, mode_id)
    return idx, val, (I_op if mode_id != _lgx.SCORE_FP32 else None), mode_id


class PureMF(BasicModel):
    """PT/model.py:41-84; scoring runs on the same kernels (no graph)."""

    def __init__(self, config: dict, dataset: BasicDataset):
        super().__init__()
        self.num_users = dataset.n_users
        self.num_items = dataset.m_items
        self.latent_dim = config["latent_dim_rec"]
        self.f = nn.Sigmoid()
        self.embedding_user = nn.Embedding(self.num_users, self.latent_dim)   # N(0,1) default init, PT/model.py:52-57
        self.embedding_item = nn.Embedding(self.num_items, self.latent_dim)
        object.__setattr__(self, "_dataset", dataset)      # not a submodule / not in the state_dict

    def getUsersRating(self, users):
        users = _as_index(users, self.embedding_user.weight.device)
        with torch.no_grad():
            return _lgx.score_dense(self.embedding_user.weight.detach().contiguous(), users,
                                    self.embedding_item.weight.detach().contiguous(), apply_sigmoid=True)

    def topk(self, users, k: int, exclude_train: bool = True, mode: str | None = None, sigmoid: bool = False):
        """Fused getUsersRating + train mask + torch.topk (PT/Procedure.py:127-135) so Procedure.Test serves the MF
        baseline too (the reference's Test only needs getUsersRating, PT/Procedure.py:126)."""
        dev = self.embedding_user.weight.device
        if dev.type != "cuda":
            raise RuntimeError("PureMF (B200 engine) must be on a CUDA device: there is no CPU path")
        graph = None
        if exclude_train:
            if self._dataset is None or not hasattr(self._dataset, "getGraphHandle"):
                raise RuntimeError("PureMF.topk(exclude_train=True) needs a dataset with getGraphHandle()")
            graph = self._dataset.getGraphHandle()
        with torch.no_grad():
            users = _as_index(users, dev)
            idx, val, _, _ = fused_topk(graph, self.embedding_user.weight.detach().contiguous(),
                                        self.embedding_item.weight.detach().contiguous(), users, k,
                                        _lgx.MODES[mode or "bf16x3"])
            if sigmoid:
                val = torch.where(val == -1024.0, val, torch.sigmoid(val))
            return idx, val

    def bpr_loss(self, users, pos, neg):
        users_emb = self.embedding_user(users.long())
        pos_emb = self.embedding_item(pos.long())
        neg_emb = self.embedding_item(neg.long())
        pos_scores = torch.sum(users_emb * pos_emb, dim=1)
        neg_scores = torch.sum(users_emb * neg_emb, dim=1)
        loss = torch.mean(nn.functional.softplus(neg_scores - pos_scores))
        reg_loss = (1 / 2) * (users_emb.norm(2).pow(2) + pos_emb.norm(2).pow(2) + neg_emb.norm(2).pow(2)) / float(len(users))
        return loss, reg_loss

    def forward(self, users, items):
        users_emb = self.embedding_user(users.long())
        items_emb = self.embedding_item(items.long())
        return self.f(torch.sum(users_emb * items_emb, dim=1))


class LightGCN(BasicModel):
    """PT/model.py:87-220."""

    def __init__(self, config: dict, dataset: BasicDataset):
        super().__init__()
        self.config = config
        self.dataset = dataset
        self.__init_weight()

    def __init_weight(self):
        self.num_users = self.dataset.n_users
        self.num_items = self.dataset.m_items
        self.latent_dim = self.config["latent_dim_rec"]
        self.n_layers = self.config["lightGCN_n_layers"]
        self.keep_prob = self.config["keep_prob"]
        self.A_split = self.config["A_split"]
        if self.A_split:
            raise NotImplementedError("A_split folds are replaced by row-sharding across GPUs (parallel.py); "
                                      "the reference hard-disables them too (PT/world.py:49)")
        self.embedding_user = nn.Embedding(self.num_users, self.latent_dim)
        self.embedding_item = nn.Embedding(self.num_items, self.latent_dim)
        if self.config["pretrain"] == 0:
            nn.init.normal_(self.embedding_user.weight, std=0.1)              # PT/model.py:112-113
            nn.init.normal_(self.embedding_item.weight, std=0.1)
        else:
            self.embedding_user.weight.data.copy_(torch.from_numpy(np.asarray(self.config["user_emb"])))
            self.embedding_item.weight.data.copy_(torch.from_numpy(np.asarray(self.config["item_emb"])))
        self.f = nn.Sigmoid()
        self._drop_calls = 0       # one fresh edge-dropout seed per computer() call in training mode
        self._flat = None          # one contiguous [N, d] buffer; the two weights are views of it
        self._graph = None
        self._eval_cache = None
        self._packed = {}
        self._fuse_parameters()

    # ---- storage: users and items tables live in one [N, d] buffer (no torch.cat per call, PT/model.py:151)
    def _fuse_parameters(self):
        wu, wi = self.embedding_user.weight, self.embedding_item.weight
        flat = torch.empty(self.num_users + self.num_items, self.latent_dim, dtype=wu.dtype, device=wu.device)
        flat[: self.num_users].copy_(wu.data)
        flat[self.num_users:].copy_(wi.data)
        wu.data = flat[: self.num_users]
        wi.data = flat[self.num_users:]
        self._flat = flat

    def _flat_if_fused(self):
        wu, wi, f = self.embedding_user.weight, self.embedding_item.weight, self._flat
        if (f is not None and f.device == wu.device and wu.data_ptr() == f.data_ptr()
                and wi.data_ptr() == f.data_ptr() + self.num_users * self.latent_dim * f.element_size()
                and wu.is_contiguous() and wi.is_contiguous()):
            return f
        return None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._fuse_parameters()
        self._eval_cache = None
        self._packed = {}
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._eval_cache = None
        self._packed = {}
        return out

    # ---- graph
    @property
    def graph(self) -> "_lgx.Graph":
        if self._graph is None:
            dev = self.embedding_user.weight.device
            if dev.type != "cuda":
                raise RuntimeError("LightGCN (B200 engine) must be on a CUDA device: there is no CPU path")
            if hasattr(self.dataset, "getGraphHandle"):
                self._graph = self.dataset.getGraphHandle()
            else:                                   # foreign BasicDataset: adopt its getSparseGraph() tensor
                coo = self.dataset.getSparseGraph().coalesce().to(dev)
                n = coo.shape[0]
                rows, cols = coo.indices()
                indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
                indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
                self._graph = _lgx.Graph.from_csr(indptr, cols.to(torch.int32), coo.values().float(), n_cols=coo.shape[1],
                                                  n_users=self.num_users, m_items=self.num_items)
        return self._graph

    @property
    def Graph(self):
        """The reference attribute (a torch sparse tensor, PT/model.py:120), materialised on demand."""
        return self.dataset.getSparseGraph()

    # ---- propagation
    def _weights_key(self):
        wu, wi = self.embedding_user.weight, self.embedding_item.weight
        return (wu.data_ptr(), wi.data_ptr(), wu._version, wi._version, self.n_layers)

    def _dropout_spec(self):
        """(keep_prob, seed) for this call when --dropout 1 and training (PT/model.py:154-159), else None.
        The reference draws torch.rand on the CPU and rebuilds the COO tensor per call; here the keep mask is a
        counter-based hash evaluated inside the SpMM, one new seed per call (statistical parity)."""
        if not (self.config["dropout"] and self.training):
            return None
        self.graph.enable_dropout()
        self._drop_calls += 1
        return (float(self.keep_prob), (int(self.config.get("seed", world.seed)) << 32) + self._drop_calls)

    def computer(self):
        """PT/model.py:145-177."""
        wu, wi = self.embedding_user.weight, self.embedding_item.weight
        drop = self._dropout_spec()
        need_grad = torch.is_grad_enabled() and (wu.requires_grad or wi.requires_grad)
        if not need_grad:
            key = self._weights_key()
            if drop is None and self._eval_cache is not None and self._eval_cache[0] == key:
                return self._eval_cache[1]
            flat = self._flat_if_fused()
            E0 = flat if flat is not None else torch.cat([wu.detach(), wi.detach()])
            out = self.graph.propagate_fwd(E0.detach(), self.n_layers, dropout=drop)
            res = (out[: self.num_users], out[self.num_users:])
            if drop is None:
                self._eval_cache = (key, res)
            return res
        return _Propagate.apply(self._flat_if_fused(), wu, wi, self.graph, self.n_layers, drop)

    # ---- scoring
    def getUsersRating(self, users):
        """PT/model.py:179-184: sigmoid(U_B I^T), fp32 [B, m_items]."""
        with torch.no_grad():
            all_users, all_items = self.computer()
            users = _as_index(users, all_users.device)
            return _lgx.score_dense(all_users, users, all_items, apply_sigmoid=True)

    def topk(self, users, k: int, exclude_train: bool = True, mode: str | None = None, sigmoid: bool = False):
        """Fused getUsersRating + train mask + torch.topk (PT/Procedure.py:127-135).
        -> (idx int64 [B,k], score fp32 [B,k]); scores are raw dot products unless sigmoid=True."""
        mode_id = _lgx.MODES[mode or self.config.get("score_mode", "bf16x3")]
        with torch.no_grad():
            all_users, all_items = self.computer()
            users = _as_index(users, all_users.device)
            mode_id = tc_mode_for(mode_id, self.latent_dim, k)
            key = (self._weights_key(), mode_id)
            idx, val, packed, _ = fused_topk(self.graph if exclude_train else None, all_users, all_items, users, k, mode_id,
                                             packed_items=self._packed.get(key))
            if packed is not None and key not in self._packed:
                self._packed = {key: packed}
            if sigmoid:
                val = torch.where(val == -1024.0, val, torch.sigmoid(val))
            return idx, val

    # ---- training
    def getEmbedding(self, users, pos_items, neg_items):
        """PT/model.py:186-194."""
        all_users, all_items = self.computer()
        users_emb = all_users[users]
        pos_emb = all_items[pos_items]
        neg_emb = all_items[neg_items]
        return (users_emb, pos_emb, neg_emb, self.embedding_user(users), self.embedding_item(pos_items),
                self.embedding_item(neg_items))

    def bpr_loss(self, users, pos, neg):
        """PT/model.py:196-209 -> (loss, reg_loss), differentiable w.r.t. the two embedding tables."""
        dev = self.embedding_user.weight.device
        users, pos, neg = _as_index(users, dev), _as_index(pos, dev), _as_index(neg, dev)
        return _BprLoss.apply(self._flat_if_fused(), self.embedding_user.weight, self.embedding_item.weight,
                              self.graph, self.n_layers, users, pos, neg, self._dropout_spec())

    def forward(self, users, items):
        """PT/model.py:211-220."""
        all_users, all_items = self.computer()
        users_emb = all_users[users]
        items_emb = all_items[items]
        return torch.sum(torch.mul(users_emb, items_emb), dim=1)
