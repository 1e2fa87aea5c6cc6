"""ctypes binding of liblgx.so (include/lgx.h).  PyTorch supplies device memory and streams only.

There is no fallback: if the library is missing or the device is not sm_100 every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LGX_LIB_PATH", os.path.join(_PKG, "liblgx.so"))   # override: A/B builds in experiments

SCORE_FP32, SCORE_BF16, SCORE_BF16X3 = 0, 1, 2
MODES = {"fp32": SCORE_FP32, "bf16": SCORE_BF16, "bf16x3": SCORE_BF16X3}

_lib = None

_P = C.c_void_p
_SIGS = {
    "lgx_last_error": (C.c_char_p, []),
    "lgx_version": (C.c_int, []),
    "lgx_device_check": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "lgx_graph_build": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, _P, _P, C.c_int32, _P, C.POINTER(_P)]),
    "lgx_graph_build_host": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, _P, _P, C.c_int32, _P, C.POINTER(_P)]),
    "lgx_graph_from_csr": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P,
                                     C.POINTER(_P)]),
    "lgx_graph_info": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "lgx_graph_export": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "lgx_graph_pointers": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "lgx_graph_get_flags": (C.c_int, [_P]),
    "lgx_graph_set_flags": (C.c_int, [_P, C.c_int32]),
    "lgx_graph_destroy": (C.c_int, [_P]),
    "lgx_spmm_workspace_bytes": (C.c_size_t, [_P, C.c_int32]),
    "lgx_spmm": (C.c_int, [_P, _P, _P, _P, _P, C.c_float, C.c_int32, _P, _P]),
    "lgx_spmm_peers": (C.c_int, [_P, _P, _P, C.POINTER(_P), C.c_int32, C.c_int64, C.c_int32, _P, C.c_float, C.c_int32, _P, _P]),
    "lgx_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P), C.c_char_p]),
    "lgx_peer_open": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "lgx_peer_close": (C.c_int, [_P]),
    "lgx_peer_copy": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32, C.c_size_t, C.c_size_t, _P]),
    "lgx_peer_free": (C.c_int, [_P]),
    "lgx_peer_barrier": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32, C.c_uint32, _P]),
    "lgx_propagate_workspace_bytes": (C.c_size_t, [_P, C.c_int32, C.c_int32]),
    "lgx_propagate_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "lgx_propagate_bwd": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "lgx_graph_enable_dropout": (C.c_int, [_P, _P]),
    "lgx_dropout_mask": (C.c_int, [_P, C.c_float, C.c_uint64, C.c_int32, _P, _P]),
    "lgx_propagate_fwd_dropout": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_float, C.c_uint64, _P, _P]),
    "lgx_propagate_bwd_dropout": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_float, C.c_uint64, _P, _P]),
    "lgx_score_dense": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, C.c_int32, _P, C.c_int32, _P]),
    "lgx_score_minmax": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "lgx_score_bucket": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P]),
    "lgx_pack_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "lgx_pack_operand": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "lgx_score_topk_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "lgx_score_plan": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "lgx_score_topk": (C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                 _P, _P, _P, C.c_size_t, _P]),
    "lgx_topk_merge": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "lgx_bpr_forward": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "lgx_bpr_backward_light": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P, _P]),
    "lgx_bpr_backward_reg": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P, _P]),
    "lgx_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, _P]),
    "lgx_adam_step_dev": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, _P, _P]),
    "lgx_sample_bpr": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_uint64, _P, _P]),
    "lgx_rank_metrics": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
}
EXPORTED = tuple(_SIGS)


def lib():
    """Load liblgx.so once.  Missing library is a hard error (no CPU / eager fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m factors_of_serendipity_recommendation_b200.build` "
                "(there is no fallback path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().lgx_last_error().decode()
        raise RuntimeError(f"liblgx error {rc}: {msg}")


def ptr(t):
    """Device (or host) pointer of a tensor, None -> NULL."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("liblgx needs CUDA tensors: there is no CPU path in this engine")
        if t is not None and not t.is_contiguous():
            raise RuntimeError("liblgx needs contiguous tensors")


class Graph:
    """Owner of an lgx_graph handle (canonical CSR of D^-1/2 A D^-1/2 + SpMM schedule)."""

    def __init__(self, handle: int, device: torch.device):
        self.handle = C.c_void_p(handle)
        self.device = device
        info = (C.c_int64 * 10)()
        check(lib().lgx_graph_info(self.handle, info))
        (self.n_rows, self.n_cols, self.nnz, self.n_users, self.m_items, self.n_work, self.n_long,
         self.max_row_nnz, self.n_partials, self.chunk_nnz) = [int(x) for x in info]
        self._ws = {}

    # ---- construction
    @staticmethod
    def build(n_users: int, m_items: int, users: torch.Tensor, items: torch.Tensor, chunk_nnz: int = 0) -> "Graph":
        """users/items int32: CUDA tensors (device build) or CPU tensors (copied inside)."""
        users = users.to(torch.int32).contiguous()
        items = items.to(torch.int32).contiguous()
        if users.numel() != items.numel():
            raise ValueError("users and items must have the same length")
        out = C.c_void_p()
        if users.is_cuda:
            dev = users.device
            with torch.cuda.device(dev):
                check(lib().lgx_graph_build(n_users, m_items, users.numel(), ptr(users), ptr(items), chunk_nnz,
                                            stream(), C.byref(out)))
        else:
            if not torch.cuda.is_available():
                raise RuntimeError("liblgx needs a CUDA device (B200, sm_100): there is no CPU path")
            dev = torch.device("cuda", torch.cuda.current_device())
            check(lib().lgx_graph_build_host(n_users, m_items, users.numel(), ptr(users), ptr(items), chunk_nnz,
                                             stream(), C.byref(out)))
        return Graph(out.value, dev)

    @staticmethod
    def from_csr(indptr: torch.Tensor, indices: torch.Tensor, values: torch.Tensor, n_cols: int,
                 n_users: int = 0, m_items: int = 0, chunk_nnz: int = 0) -> "Graph":
        indptr = indptr.to(torch.int64).contiguous()
        indices = indices.to(torch.int32).contiguous()
        values = values.to(torch.float32).contiguous()
        require_cuda(indptr, indices, values)
        out = C.c_void_p()
        with torch.cuda.device(indptr.device):
            check(lib().lgx_graph_from_csr(indptr.numel() - 1, n_cols, indices.numel(), ptr(indptr), ptr(indices),
                                           ptr(values), n_users, m_items, chunk_nnz, stream(), C.byref(out)))
        return Graph(out.value, indptr.device)

    @property
    def normalized(self) -> bool:
        """every value == dinv[row] * dinv[col] (graphs built from unique pairs)"""
        return bool(lib().lgx_graph_get_flags(self.handle) & 1)

    def assume_normalized(self, flag: bool = True):
        """declare that an adopted CSR (from_csr) is a row block of a normalised adjacency"""
        check(lib().lgx_graph_set_flags(self.handle, 1 if flag else 0))

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                lib().lgx_graph_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- export
    def export(self):
        """-> dict of device tensors: indptr int64, indices int32, values f32, degree int32, dinv f32, row_order int32."""
        dev = self.device
        out = dict(
            indptr=torch.empty(self.n_rows + 1, dtype=torch.int64, device=dev),
            indices=torch.empty(self.nnz, dtype=torch.int32, device=dev),
            values=torch.empty(self.nnz, dtype=torch.float32, device=dev),
            degree=torch.empty(self.n_rows, dtype=torch.int32, device=dev),
            dinv=torch.empty(self.n_rows, dtype=torch.float32, device=dev),
            row_order=torch.empty(self.n_rows, dtype=torch.int32, device=dev),
        )
        with torch.cuda.device(dev):
            check(lib().lgx_graph_export(self.handle, ptr(out["indptr"]), ptr(out["indices"]), ptr(out["values"]),
                                         ptr(out["degree"]), ptr(out["dinv"]), ptr(out["row_order"]), stream()))
        return out

    def to_torch_coo(self) -> torch.Tensor:
        """The reference's getSparseGraph() contract: coalesced fp32 COO, int64 indices (PT/dataloader.py:331-337,374)."""
        e = self.export()
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=self.device), e["indptr"][1:] - e["indptr"][:-1])
        idx = torch.stack([rows, e["indices"].long()])
        return torch.sparse_coo_tensor(idx, e["values"], (self.n_rows, self.n_cols), is_coalesced=True)

    # ---- workspaces (cached per (kind, d, L))
    def workspace(self, nbytes: int, key) -> torch.Tensor:
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    # ---- propagation
    def spmm(self, X, S_in=None, Y=None, S_out=None, div: float = 1.0):
        require_cuda(X, S_in, Y, S_out)
        d = X.shape[1]
        ws = self.workspace(lib().lgx_spmm_workspace_bytes(self.handle, d), ("spmm", d))
        with torch.cuda.device(self.device):
            check(lib().lgx_spmm(self.handle, ptr(X), ptr(S_in), ptr(Y), ptr(S_out), div, d, ptr(ws), stream()))

    def spmm_peers(self, X, peer_ptrs, row_offset: int, S_in=None, S_out=None, div: float = 1.0, store_mean: bool = False):
        """One layer whose output rows go straight into every rank's gathered buffer (peer_ptrs: list of ints)."""
        require_cuda(X, S_in, S_out)
        d = X.shape[1]
        ws = self.workspace(lib().lgx_spmm_workspace_bytes(self.handle, d), ("spmm", d))
        arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        with torch.cuda.device(self.device):
            check(lib().lgx_spmm_peers(self.handle, ptr(X), ptr(S_in), arr, len(peer_ptrs), row_offset, int(store_mean),
                                       ptr(S_out), div, d, ptr(ws), stream()))

    def enable_dropout(self):
        with torch.cuda.device(self.device):
            check(lib().lgx_graph_enable_dropout(self.handle, stream()))

    def dropout_mask(self, keep_prob: float, seed: int, transpose: bool = False) -> torch.Tensor:
        out = torch.empty(self.nnz, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().lgx_dropout_mask(self.handle, keep_prob, seed, int(transpose), ptr(out), stream()))
        return out

    def propagate_fwd(self, E0: torch.Tensor, n_layers: int, out=None, layers_out=None, dropout=None) -> torch.Tensor:
        """dropout = (keep_prob, seed): propagate through the edge-dropped graph of that seed."""
        require_cuda(E0, out, layers_out)
        if E0.dtype != torch.float32 or E0.shape[0] != self.n_rows:
            raise ValueError("E0 must be fp32 [n_rows, d]")
        d = E0.shape[1]
        if out is None:
            out = torch.empty_like(E0)
        ws = self.workspace(lib().lgx_propagate_workspace_bytes(self.handle, d, n_layers), ("prop", d))
        with torch.cuda.device(self.device):
            if dropout is not None:
                check(lib().lgx_propagate_fwd_dropout(self.handle, ptr(E0), ptr(out), n_layers, d, float(dropout[0]),
                                                      int(dropout[1]), ptr(ws), stream()))
            else:
                check(lib().lgx_propagate_fwd(self.handle, ptr(E0), ptr(out), ptr(layers_out), n_layers, d, ptr(ws), stream()))
        return out

    def propagate_bwd(self, g_scaled: torch.Tensor, n_layers: int, out=None, dropout=None) -> torch.Tensor:
        require_cuda(g_scaled, out)
        d = g_scaled.shape[1]
        if out is None:
            out = torch.empty_like(g_scaled)
        ws = self.workspace(lib().lgx_propagate_workspace_bytes(self.handle, d, n_layers), ("prop", d))
        with torch.cuda.device(self.device):
            if dropout is not None:
                check(lib().lgx_propagate_bwd_dropout(self.handle, ptr(g_scaled), ptr(out), n_layers, d, float(dropout[0]),
                                                      int(dropout[1]), ptr(ws), stream()))
            else:
                check(lib().lgx_propagate_bwd(self.handle, ptr(g_scaled), ptr(out), n_layers, d, ptr(ws), stream()))
        return out

    # ---- sampler
    def sample_bpr(self, n_samples: int, per_user: int = 0, seed: int = 2020) -> torch.Tensor:
        n = self.n_users * per_user if per_user > 0 else n_samples
        out = torch.empty(n, 3, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().lgx_sample_bpr(self.handle, n_samples, per_user, seed, ptr(out), stream()))
        return out


class PeerBuffer:
    """A device buffer the other ranks of the node can map (CUDA IPC): backs the gathered layers of the
    fused SpMM + all-gather.  ``tensor`` is a zero-copy torch view of the local allocation."""

    class _CAI:
        def __init__(self, p, shape):
            self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (p, False), "version": 3,
                                             "strides": None}

    def __init__(self, shape, device):
        self.shape, self.device = tuple(shape), device
        nbytes = 4
        for x in self.shape:
            nbytes *= int(x)
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        with torch.cuda.device(device):
            check(lib().lgx_peer_alloc(nbytes, C.byref(p), h))
        self.ptr, self.handle = int(p.value), h.raw
        self.tensor = torch.as_tensor(PeerBuffer._CAI(self.ptr, self.shape), device=device)
        self.opened = []

    def open_peer(self, handle: bytes) -> int:
        p = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().lgx_peer_open(handle, C.byref(p)))
        self.opened.append(int(p.value))
        return int(p.value)

    def close(self):
        for p in self.opened:
            lib().lgx_peer_close(C.c_void_p(p))
        self.opened = []
        if self.ptr:
            self.tensor = None
            lib().lgx_peer_free(C.c_void_p(self.ptr))
            self.ptr = 0


def peer_barrier(flag_ptrs, self_rank: int, epoch: int, device):
    """Device-side cross-rank barrier over peer-mapped flag arrays (lgx_peer_barrier) on the current stream."""
    arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
    with torch.cuda.device(device):
        check(lib().lgx_peer_barrier(arr, len(flag_ptrs), self_rank, epoch & 0xFFFFFFFF or 1, stream()))


def peer_copy(peer_ptrs, self_rank: int, offset_bytes: int, nbytes: int, cuda_stream, device):
    """P2P DMA of one byte range of my buffer into every other rank's buffer (one async copy per peer)."""
    arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
    with torch.cuda.device(device):
        check(lib().lgx_peer_copy(arr, len(peer_ptrs), self_rank, offset_bytes, nbytes, C.c_void_p(cuda_stream.cuda_stream)))


# ------------------------------------------------------------------------------------- free functions
def device_check():
    sm, l2 = C.c_int(), C.c_int64()
    check(lib().lgx_device_check(C.byref(sm), C.byref(l2)))
    return sm.value, l2.value


def score_dense(U, users, I, apply_sigmoid=True, out=None):
    require_cuda(U, users, I, out)
    B = U.shape[0] if users is None else users.numel()
    M, d = I.shape
    if out is None:
        out = torch.empty(B, M, dtype=torch.float32, device=I.device)
    with torch.cuda.device(I.device):
        check(lib().lgx_score_dense(ptr(U), ptr(users), B, ptr(I), M, d, ptr(out), int(apply_sigmoid), stream()))
    return out


def score_minmax(U, I):
    """(min, max) of fp16(U I^T) as a device float[2] tensor -- recommend.py:377 without the [n_user, n_item] matrix."""
    require_cuda(U, I)
    out = torch.empty(2, dtype=torch.float32, device=I.device)
    ws = torch.empty(2, dtype=torch.int32, device=I.device)
    with torch.cuda.device(I.device):
        check(lib().lgx_score_minmax(ptr(U), U.shape[0], ptr(I), I.shape[0], I.shape[1], ptr(out), ptr(ws), stream()))
    return out


def score_bucket(U, users, I, min_dis: float, inter: float, out=None):
    """int8 labels [B, M] = floor((fp16(score) - min_dis) / inter) in fp16 arithmetic (recommend.py:380)."""
    require_cuda(U, users, I, out)
    B = U.shape[0] if users is None else users.numel()
    M, d = I.shape
    if out is None:
        out = torch.empty(B, M, dtype=torch.int8, device=I.device)
    with torch.cuda.device(I.device):
        check(lib().lgx_score_bucket(ptr(U), ptr(users), B, ptr(I), M, d, float(min_dis), float(inter), ptr(out), stream()))
    return out


def pack_operand(src, row_ids, mode: int, is_items: bool):
    require_cuda(src, row_ids)
    rows = src.shape[0] if row_ids is None else row_ids.numel()
    d = src.shape[1]
    ktot = 3 * d if mode == SCORE_BF16X3 else d
    dst = torch.empty(rows, ktot, dtype=torch.bfloat16, device=src.device)
    with torch.cuda.device(src.device):
        check(lib().lgx_pack_operand(ptr(src), ptr(row_ids), rows, d, mode, int(is_items), ptr(dst), stream()))
    return dst


def score_plan(B: int, M: int, d: int, k: int, mode: int, sms: int = 0) -> dict:
    """Host-only: the (user tile, item split) decomposition lgx_score_topk uses for this shape."""
    out = (C.c_int32 * 4)()
    check(lib().lgx_score_plan(B, M, d, k, mode, sms, out))
    return {"user_tiles": out[0], "item_tiles": out[1], "splits": out[2], "tiles_per_split": out[3]}


_topk_ws = {}


def score_topk(graph, U_op, users, I_op, d: int, k: int, mode: int, item_offset: int = 0):
    """-> (idx int64 [B,k], raw score fp32 [B,k]).  U_op/I_op: fp32 (FP32 mode) or packed bf16 operands."""
    require_cuda(U_op, users, I_op)
    B, M = U_op.shape[0], I_op.shape[0]
    dev = I_op.device
    nbytes = lib().lgx_score_topk_workspace_bytes(B, M, d, k, mode)
    ws = _topk_ws.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _topk_ws[dev] = ws
    idx = torch.empty(B, k, dtype=torch.int64, device=dev)
    val = torch.empty(B, k, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().lgx_score_topk(graph.handle if graph is not None else None, ptr(U_op), ptr(users), B, ptr(I_op), M,
                                   d, k, mode, item_offset, ptr(idx), ptr(val), ptr(ws), ws.numel(), stream()))
    return idx, val


def topk_merge(cand_idx, cand_val):
    """[P, B, k] candidates -> (idx [B,k], val [B,k])."""
    require_cuda(cand_idx, cand_val)
    P, B, k = cand_idx.shape
    idx = torch.empty(B, k, dtype=torch.int64, device=cand_idx.device)
    val = torch.empty(B, k, dtype=torch.float32, device=cand_idx.device)
    with torch.cuda.device(cand_idx.device):
        check(lib().lgx_topk_merge(ptr(cand_idx), ptr(cand_val), P, B, k, ptr(idx), ptr(val), stream()))
    return idx, val


def bpr_forward(light, E0, users, pos, neg, n_users: int):
    require_cuda(light, E0, users, pos, neg)
    B, d = users.numel(), light.shape[1]
    out2 = torch.empty(2, dtype=torch.float32, device=light.device)
    coef = torch.empty(3 * B, dtype=torch.float32, device=light.device)
    with torch.cuda.device(light.device):
        check(lib().lgx_bpr_forward(ptr(light), ptr(E0), ptr(users), ptr(pos), ptr(neg), B, n_users, d, ptr(out2),
                                    ptr(coef), stream()))
    return out2, coef


def bpr_backward_light(light, users, pos, neg, coef, n_users: int, grad_scale: float, grad_scale_dev, G):
    require_cuda(light, users, pos, neg, coef, grad_scale_dev, G)
    B, d = users.numel(), light.shape[1]
    with torch.cuda.device(light.device):
        check(lib().lgx_bpr_backward_light(ptr(light), ptr(users), ptr(pos), ptr(neg), ptr(coef), B, n_users, d,
                                           grad_scale, ptr(grad_scale_dev), ptr(G), stream()))


def bpr_backward_reg(E0, users, pos, neg, n_users: int, grad_scale: float, grad_scale_dev, dE0):
    require_cuda(E0, users, pos, neg, grad_scale_dev, dE0)
    B, d = users.numel(), E0.shape[1]
    with torch.cuda.device(E0.device):
        check(lib().lgx_bpr_backward_reg(ptr(E0), ptr(users), ptr(pos), ptr(neg), B, n_users, d, grad_scale,
                                         ptr(grad_scale_dev), ptr(dE0), stream()))


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step: int):
    require_cuda(param, grad, exp_avg, exp_avg_sq)
    with torch.cuda.device(param.device):
        check(lib().lgx_adam_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), lr, beta1, beta2,
                                  eps, step, stream()))


def adam_step_dev(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, state4):
    require_cuda(param, grad, exp_avg, exp_avg_sq, state4)
    with torch.cuda.device(param.device):
        check(lib().lgx_adam_step_dev(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), lr, beta1, beta2,
                                      eps, ptr(state4), stream()))


def rank_metrics(topk_idx, k: int, gt_ptr, gt_items, sums3):
    require_cuda(topk_idx, gt_ptr, gt_items, sums3)
    B, k_stride = topk_idx.shape
    with torch.cuda.device(topk_idx.device):
        check(lib().lgx_rank_metrics(ptr(topk_idx), B, k_stride, k, ptr(gt_ptr), ptr(gt_items), ptr(sums3), stream()))
