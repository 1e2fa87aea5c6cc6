"""Dataset + graph build behind the reference's BasicDataset / Loader interface (PT/dataloader.py).

Same attributes and methods as the reference (n_users, m_items, trainDataSize, testDict, allPos,
getUserPosItems, getUserItemFeedback, getSparseGraph, trainUser/trainItem, users_D/items_D); the
adjacency is built on the device by liblgx (lgx_graph_build) instead of scipy lil/dok assembly.
"""
from __future__ import annotations

import os
from time import time

import numpy as np
import torch

from . import _lgx, world


class BasicDataset:
    """PT/dataloader.py:24-69."""

    @property
    def n_users(self):
        raise NotImplementedError

    @property
    def m_items(self):
        raise NotImplementedError

    @property
    def trainDataSize(self):
        raise NotImplementedError

    @property
    def testDict(self):
        raise NotImplementedError

    @property
    def allPos(self):
        raise NotImplementedError

    def getUserItemFeedback(self, users, items):
        raise NotImplementedError

    def getUserPosItems(self, users):
        raise NotImplementedError

    def getUserNegItems(self, users):
        raise NotImplementedError

    def getSparseGraph(self):
        raise NotImplementedError


def parse_interactions(path: str):
    """'uid item item ...' per line (PT/dataloader.py:247-262) -> (unique users, users[E], items[E]).

    Lines with a user id and no items are skipped (the reference crashes on them; the TF loader
    skips them, TF/utility/load_data.py:42-45 -- amazon-book/test.txt has 4 such lines).

    The reference parses line by line in Python (one list comprehension + two ``extend`` per user).  Here the whole
    file is converted in one C pass (np.fromstring, whitespace-separated integers) and the line structure comes from
    the byte array: token starts by vectorised compares, tokens per line by a binary search of the newline offsets,
    the user column by np.repeat.  Same output (tests/test_abi.py), about 2x faster on an Amazon-Book-sized train.txt."""
    import warnings
    raw = np.fromfile(path, dtype=np.uint8)
    empty = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64))
    if raw.size == 0:
        return empty
    is_nl = raw == 10
    is_space = is_nl | (raw == 32) | (raw == 13) | (raw == 9)
    starts = np.flatnonzero(is_space[:-1] & ~is_space[1:]) + 1       # first byte of every token
    if not is_space[0]:
        starts = np.concatenate(([0], starts))
    if starts.size == 0:
        return empty
    with warnings.catch_warnings():
        warnings.simplefilter("error")                               # numpy only WARNS when it stops at a bad token
        try:
            values = np.fromstring(raw.tobytes(), dtype=np.int64, sep=" ")
        except (DeprecationWarning, ValueError) as e:
            raise ValueError(f"{path}: not a whitespace-separated integer file ({e})") from None
    if values.size != starts.size:
        raise ValueError(f"{path}: {starts.size} tokens but {values.size} integers parsed")
    # tokens per line: tokens that start before each newline (and the unterminated last line)
    line_ends = np.concatenate((np.flatnonzero(is_nl), [raw.size]))
    upto = np.searchsorted(starts, line_ends, side="left")
    per_line = np.diff(np.concatenate(([0], upto)))
    first_tok_of_line = upto - per_line
    keep_line = per_line >= 2                                         # a user id and at least one item
    uniq = values[first_tok_of_line[keep_line]]
    drop = np.zeros(values.size, dtype=bool)                          # user-id tokens and id-only lines
    drop[first_tok_of_line[per_line >= 1]] = True
    items = values[~drop]
    return uniq, np.repeat(uniq, per_line[keep_line] - 1), items


def parse_interactions_lines(path: str):
    """The reference's line loop (PT/dataloader.py:247-262), kept as the checker for parse_interactions."""
    uniq, us, its = [], [], []
    with open(path) as f:
        for line in f:
            parts = line.strip("\n").strip(" ").split(" ")
            if len(parts) < 2 or parts[0] == "":
                continue
            row = np.array(parts[1:], dtype=np.int64)
            uid = int(parts[0])
            uniq.append(uid)
            us.append(np.full(row.size, uid, dtype=np.int64))
            its.append(row)
    if not us:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.array(uniq, dtype=np.int64), np.concatenate(us), np.concatenate(its)


class InteractionDataset(BasicDataset):
    """A dataset from in-memory interaction arrays (synthetic graphs, tests, bench)."""

    def __init__(self, n_users: int, m_items: int, train_user, train_item, test_dict=None, config=None,
                 device=None, path=None, graph=None, train_size=None):
        """``graph``: a prebuilt _lgx.Graph (graphs generated on the device never visit the host; then
        train_user / train_item may be None and train_size gives trainDataSize)."""
        self.config = world.config if config is None else config
        self.device = torch.device(device) if device is not None else world.device
        self.n_user, self.m_item = int(n_users), int(m_items)
        self.trainUser = np.asarray(train_user) if train_user is not None else None
        self.trainItem = np.asarray(train_item) if train_item is not None else None
        self.traindataSize = int(train_size if train_size is not None else self.trainUser.size)
        self.__testDict = dict(test_dict) if test_dict is not None else {}
        self.path = path
        self.Graph = None
        self._graph = graph
        self._allPos = None
        self._csr_host = None

    # ---- reference properties
    @property
    def n_users(self):
        return self.n_user

    @property
    def m_items(self):
        return self.m_item

    @property
    def trainDataSize(self):
        return self.traindataSize

    @property
    def testDict(self):
        return self.__testDict

    @property
    def allPos(self):
        if self._allPos is None:
            self._allPos = self.getUserPosItems(list(range(self.n_user)))
        return self._allPos

    # ---- engine handle
    def getGraphHandle(self) -> "_lgx.Graph":
        """The device CSR of D^-1/2 A D^-1/2 (built once).  Reads <path>/s_pre_adj_mat.npz when it exists,
        like PT/dataloader.py:343, otherwise builds on the device and (with a path) writes the cache (:367)."""
        if self._graph is None:
            if self.device.type != "cuda":
                raise RuntimeError("the B200 LightGCN engine needs a CUDA device: there is no CPU path")
            cache = os.path.join(self.path, "s_pre_adj_mat.npz") if self.path else None
            with torch.cuda.device(self.device):
                if cache and os.path.exists(cache):
                    import scipy.sparse as sp

                    m = sp.load_npz(cache).tocsr()
                    m.sort_indices()
                    self._graph = _lgx.Graph.from_csr(
                        torch.from_numpy(m.indptr.astype(np.int64)).to(self.device),
                        torch.from_numpy(m.indices.astype(np.int32)).to(self.device),
                        torch.from_numpy(m.data.astype(np.float32)).to(self.device),
                        n_cols=m.shape[1], n_users=self.n_user, m_items=self.m_item)
                else:
                    s = time()
                    self._graph = _lgx.Graph.build(self.n_user, self.m_item,
                                                   torch.from_numpy(self.trainUser.astype(np.int32)),
                                                   torch.from_numpy(self.trainItem.astype(np.int32)))
                    self.build_seconds = time() - s
                    if cache:
                        self.save_npz(cache)
        return self._graph

    def _host_csr(self):
        if self._csr_host is None:
            e = self.getGraphHandle().export()
            self._csr_host = {k: v.cpu().numpy() for k, v in e.items()}
        return self._csr_host

    def save_npz(self, file: str):
        """scipy.sparse.save_npz-compatible cache of the normalised adjacency (PT/dataloader.py:367)."""
        import scipy.sparse as sp

        c = self._host_csr()
        N = self.n_user + self.m_item
        m = sp.csr_matrix((c["values"], c["indices"], c["indptr"].astype(np.int32)), shape=(N, N))
        sp.save_npz(file, m)

    # ---- reference methods
    @property
    def users_D(self):
        d = self._host_csr()["degree"][: self.n_user].astype(np.float64)
        d[d == 0.0] = 1.0                              # PT/dataloader.py:290-291
        return d

    @property
    def items_D(self):
        d = self._host_csr()["degree"][self.n_user:].astype(np.float64)
        d[d == 0.0] = 1.0                              # PT/dataloader.py:292-293
        return d

    def getSparseGraph(self):
        """Coalesced fp32 COO on the device, int64 indices: the reference's contract
        (PT/dataloader.py:373-374).  The engine itself uses getGraphHandle() and never needs this copy."""
        if self.Graph is None:
            self.Graph = self.getGraphHandle().to_torch_coo()
        return self.Graph

    def getUserPosItems(self, users):
        """Train items of each user, ascending (PT/dataloader.py:404-408: UserItemNet[u].nonzero()[1])."""
        c = self._host_csr()
        indptr, indices = c["indptr"], c["indices"]
        return [indices[indptr[u]:indptr[u + 1]].astype(np.int64) - self.n_user for u in users]

    def getUserItemFeedback(self, users, items):
        """PT/dataloader.py:392-402: 1 where (user, item) is a train interaction."""
        c = self._host_csr()
        indptr, indices = c["indptr"], c["indices"]
        out = np.zeros(len(users), dtype="uint8")
        for k, (u, i) in enumerate(zip(users, items)):
            row = indices[indptr[u]:indptr[u + 1]]
            j = np.searchsorted(row, i + self.n_user)
            out[k] = j < row.size and row[j] == i + self.n_user
        return out


class Loader(InteractionDataset):
    """PT/dataloader.py:223-414: reads <path>/train.txt and test.txt."""

    def __init__(self, config=None, path="../data/gowalla", device=None):
        config = world.config if config is None else config
        self.split = config["A_split"]
        self.folds = config["A_n_fold"]
        self.mode_dict = {"train": 0, "test": 1}
        self.mode = self.mode_dict["train"]
        uniq_tr, tu, ti = parse_interactions(path + "/train.txt")
        uniq_te, eu, ei = parse_interactions(path + "/test.txt")
        n_user = int(max(tu.max(initial=0), eu.max(initial=0))) + 1
        m_item = int(max(ti.max(initial=0), ei.max(initial=0))) + 1
        test_dict = {}
        for u, i in zip(eu.tolist(), ei.tolist()):      # __build_test, PT/dataloader.py:378-390
            test_dict.setdefault(u, []).append(i)
        super().__init__(n_user, m_item, tu, ti, test_dict, config=config, device=device, path=path)
        self.trainUniqueUsers, self.testUniqueUsers = uniq_tr, uniq_te
        self.testUser, self.testItem = eu, ei
        self.testDataSize = int(eu.size)
