"""Host-facing serving loop for the hot path: embedding tables arrive in (pinned) host memory, the
top-K item ids go back to host memory, and the three legs of consecutive requests overlap.

The reference keeps everything on one stream and pays the PCIe copies serially around the kernels
(PT/Procedure.py:124-136 moves a 36 MB score matrix to the host per batch).  Here a request is
    upload stream   : H2D of the [n_users + m_items, d] fp32 tables into one of ``depth`` device slots
    compute stream  : propagation (L SpMM layers + layer mean) -> operand packing -> fused score/mask/top-K
    download stream : D2H of the [B, k] int64 ids into the caller's host buffer
chained by CUDA events, so request i's kernels run while request i+1 uploads and request i-1 downloads.
Per-request work is unchanged; only the copy latency leaves the critical path.
"""
from __future__ import annotations

import torch

from . import _lgx


def upload_slices(rank: int, world: int, n_rows: int, n_users: int):
    """Rows of the stacked [users; items] table that ``rank`` uploads: chunk = ceil(n_rows / world) rows from
    rank * chunk.  -> (chunk, user_part, item_part) where each part is (dst_lo, dst_hi, src_lo, src_hi) — rows
    [dst_lo, dst_hi) of the rank's local chunk come from rows [src_lo, src_hi) of the host user / item table — or
    None when the rank's range does not touch that table.  Pure index arithmetic (tests/test_parallel_cpu.py)."""
    chunk = (n_rows + world - 1) // world
    r0 = rank * chunk
    r1 = min(n_rows, r0 + chunk)
    user_part = item_part = None
    if r0 < min(r1, n_users):
        user_part = (0, min(r1, n_users) - r0, r0, min(r1, n_users))
    if max(r0, n_users) < r1:
        lo = max(r0, n_users)
        item_part = (lo - r0, r1 - r0, lo - n_users, r1 - n_users)
    return chunk, user_part, item_part


class HostPipeline:
    """Double-buffered full-catalogue top-K over host-resident embedding tables.

    ``step_fn(E0, light) -> (idx, val)`` runs one hot-path step on the device tables ``E0`` ([N, d] fp32;
    ``light`` is a same-shaped scratch buffer for the propagated embeddings).  ``for_model`` builds it for a
    LightGCN model on one GPU, ``for_engine`` for a ``parallel.ShardedEngine`` rank.
    """

    def __init__(self, step_fn, n_rows: int, dim: int, device, depth: int = 2, flush_l2: torch.Tensor | None = None,
                 upload_group=None):
        """``upload_group``: a torch.distributed process group used ONLY for uploads.  With it every rank copies
        1/world of the table rows from the host and the ranks all-gather them over NVLink, instead of each rank
        pulling the whole table over its own PCIe link (at 8 GPUs the full upload, not the kernels, bounded the
        step).  It must be a dedicated group: collectives of one group are serialised in issue order, so sharing
        the compute path's group would chain request i+1's upload behind request i's kernels."""
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.step_fn = step_fn
        self.dev = torch.device(device)
        self.depth = depth
        self.flush_l2 = flush_l2
        self.n_rows = n_rows
        self.group = upload_group
        if upload_group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(upload_group), dist.get_rank(upload_group)
            self.chunk = (n_rows + self.world - 1) // self.world
            self.padded = [torch.empty(self.chunk * self.world, dim, dtype=torch.float32, device=self.dev)
                           for _ in range(depth)]
            self.local = [torch.empty(self.chunk, dim, dtype=torch.float32, device=self.dev) for _ in range(depth)]
            self.tables = [t[:n_rows] for t in self.padded]
        else:
            self.world, self.rank = 1, 0
            self.tables = [torch.empty(n_rows, dim, dtype=torch.float32, device=self.dev) for _ in range(depth)]
        self.light = [torch.empty(n_rows, dim, dtype=torch.float32, device=self.dev) for _ in range(depth)]
        self.s_up = torch.cuda.Stream(device=self.dev)
        self.s_comp = torch.cuda.Stream(device=self.dev)
        self.s_down = torch.cuda.Stream(device=self.dev)
        for st in (self.s_up, self.s_comp, self.s_down):   # whatever the caller prepared on its stream is visible
            st.wait_stream(torch.cuda.current_stream(self.dev))
        self.idx = [None] * depth               # per-slot [B, k] result buffers, allocated by the first request
        self.uploaded = [torch.cuda.Event() for _ in range(depth)]
        self.computed = [None] * depth          # recorded once the slot has been used
        self.downloaded = [None] * depth
        self.n_submitted = 0
        self.last_download = None

    # ---- constructors for the two product paths
    @classmethod
    def for_model(cls, model, users, k: int, mode: str = "bf16", depth: int = 2, flush_l2=None):
        g, L, d = model.graph, model.n_layers, model.latent_dim
        nu = model.num_users
        mode_id = _lgx.MODES[mode]
        users = users.to(model.embedding_user.weight.device)
        # every user in order = the identity batch (users=None): the scoring call then keeps the graph's train mask
        # in tile-bucketed form instead of bucketing the batch's rows on every request (checked once, here)
        if users.numel() == nu and bool((users == torch.arange(nu, device=users.device)).all()):
            users = None

        def step(E0, light):
            g.propagate_fwd(E0, L, out=light)
            au, ai = light[:nu], light[nu:]
            if mode_id == _lgx.SCORE_FP32:
                return _lgx.score_topk(g, au if users is None else au.index_select(0, users), users, ai, d, k, mode_id)
            I_op = _lgx.pack_operand(ai, None, mode_id, True)
            U_op = _lgx.pack_operand(au, users, mode_id, False)
            return _lgx.score_topk(g, U_op, users, I_op, d, k, mode_id)

        dev = model.embedding_user.weight.device
        return cls(step, nu + model.num_items, d, dev, depth=depth, flush_l2=flush_l2)

    @classmethod
    def for_engine(cls, engine, n_rows: int, dim: int, users, k: int, mode_id: int, shard: str = "auto", depth: int = 2,
                   flush_l2=None, upload_group=None):
        def step(E0, light):
            return engine.step(E0, users, k, mode_id, shard=shard)

        return cls(step, n_rows, dim, engine.dev, depth=depth, flush_l2=flush_l2, upload_group=upload_group)

    # ---- one request
    def submit(self, host_user: torch.Tensor, host_item: torch.Tensor, host_out: torch.Tensor) -> None:
        """Enqueue one request.  ``host_user`` [n_users, d] and ``host_item`` [m_items, d] are fp32 host tensors
        (pinned for asynchronous copies); ``host_out`` [B, k] int64 receives the ids.  Returns immediately;
        ``host_out`` is valid after ``wait()``.  The caller must not overwrite the host tables of a request
        until it has been uploaded (``wait()`` or ``depth`` later submits)."""
        if host_user.is_cuda or host_item.is_cuda or host_out.is_cuda:
            raise ValueError("HostPipeline takes host tensors; use model.topk for device-resident tables")
        b = self.n_submitted % self.depth
        nu = host_user.shape[0]
        tables = self.tables[b]
        if nu + host_item.shape[0] != tables.shape[0] or host_user.shape[1] != tables.shape[1]:
            raise ValueError("host tables do not match the pipeline's [n_users + m_items, d] shape")
        with torch.cuda.stream(self.s_up):
            if self.computed[b] is not None:
                self.s_up.wait_event(self.computed[b])        # the slot's previous request has consumed its tables
            if self.group is None:
                tables[:nu].copy_(host_user, non_blocking=True)
                tables[nu:].copy_(host_item, non_blocking=True)
            else:
                import torch.distributed as dist
                local = self.local[b]                            # my 1/world of the stacked [users; items] rows
                _, user_part, item_part = upload_slices(self.rank, self.world, self.n_rows, nu)
                for part, host in ((user_part, host_user), (item_part, host_item)):
                    if part is not None:
                        local[part[0]:part[1]].copy_(host[part[2]:part[3]], non_blocking=True)
                dist.all_gather_into_tensor(self.padded[b], local, group=self.group)
            self.uploaded[b].record(self.s_up)
        with torch.cuda.stream(self.s_comp):
            self.s_comp.wait_event(self.uploaded[b])
            if self.downloaded[b] is not None:
                self.s_comp.wait_event(self.downloaded[b])    # the slot's previous result has left the device
            if self.flush_l2 is not None:
                self.flush_l2.fill_(1)                        # benchmarking: evict L2 between requests
            idx, _ = self.step_fn(tables, self.light[b])
            # Results leave through a per-slot buffer: every temporary of the step lives and dies on the compute
            # stream, so the caching allocator reuses it in stream order however far the host runs ahead (a
            # cross-stream record_stream() lifetime forced fresh cudaMallocs -- sporadic 3x stalls on B200).
            if self.idx[b] is None or self.idx[b].shape != idx.shape:
                self.idx[b] = torch.empty_like(idx)
            self.idx[b].copy_(idx)
            done = torch.cuda.Event()
            done.record(self.s_comp)
            self.computed[b] = done
        with torch.cuda.stream(self.s_down):
            self.s_down.wait_event(done)
            host_out.copy_(self.idx[b], non_blocking=True)
            self.downloaded[b] = torch.cuda.Event()
            self.downloaded[b].record(self.s_down)
            self.last_download = torch.cuda.Event(enable_timing=True)
            self.last_download.record(self.s_down)
        self.n_submitted += 1

    def start_event(self) -> torch.cuda.Event:
        """Timing event recorded on the upload stream after everything already enqueued on the three streams."""
        self.s_up.wait_stream(self.s_comp)
        self.s_up.wait_stream(self.s_down)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(self.s_up)
        return ev

    def wait(self) -> None:
        """Block until every submitted request's ids are in its host buffer."""
        self.s_up.synchronize()
        self.s_comp.synchronize()
        self.s_down.synchronize()
