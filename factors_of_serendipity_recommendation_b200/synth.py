"""Seeded synthetic bipartite interaction graphs of the BASELINE.json shapes.

There is no network in the build or bench environment and the reference's gowalla /
yelp2018 / amazon-book ``train.txt`` files are not shipped
(/root/reference/.MISSING_LARGE_BLOBS), so every benchmark shape is generated:
unique (user, item) pairs, every user and item has degree >= 1, item popularity
~ rank^-0.8 (ids shuffled), user activity ~ log-normal(sigma=1)  (SURVEY.md section 8d).
"""
from __future__ import annotations

import numpy as np

# name -> (n_users, m_items, n_edges, latent_dim)
SHAPES = {
    "tiny": (300, 500, 6_000, 64),
    "mlls-shape": (608, 2_120, 63_687, 64),
    "gowalla": (29_858, 40_981, 1_027_370, 64),
    "yelp2018": (31_668, 38_048, 1_561_406, 64),
    "amazon-book": (52_643, 91_599, 2_984_108, 64),
    "synth-1b": (10_000_000, 2_000_000, 1_000_000_000, 128),
}


def make_interactions(n_users: int, m_items: int, n_edges: int, seed: int = 2020):
    """Return (users int32[E], items int32[E]) sorted by (user, item), all pairs unique."""
    if n_edges < max(n_users, m_items):
        raise ValueError("n_edges must cover every user and item at least once")
    if n_edges > n_users * m_items:
        raise ValueError("n_edges exceeds the dense size")
    rng = np.random.default_rng(seed)
    pop = np.arange(1, m_items + 1, dtype=np.float64) ** -0.8
    pop = pop[rng.permutation(m_items)]
    pop_cdf = np.cumsum(pop / pop.sum())
    act = rng.lognormal(mean=0.0, sigma=1.0, size=n_users)
    act_cdf = np.cumsum(act / act.sum())

    def draw_items(k):
        return np.minimum(np.searchsorted(pop_cdf, rng.random(k)), m_items - 1).astype(np.int64)

    def draw_users(k):
        return np.minimum(np.searchsorted(act_cdf, rng.random(k)), n_users - 1).astype(np.int64)

    # coverage edges: one per user (popularity-sampled item) and one per item (activity-sampled user)
    cov_u = np.concatenate([np.arange(n_users, dtype=np.int64), draw_users(m_items)])
    cov_i = np.concatenate([draw_items(n_users), np.arange(m_items, dtype=np.int64)])
    cov = np.unique(cov_u * m_items + cov_i)
    keys = cov
    while keys.size < n_edges:
        need = n_edges - keys.size
        k = int(need * 1.3) + 1024
        extra = draw_users(k) * m_items + draw_items(k)
        keys = np.union1d(keys, extra)
    if keys.size > n_edges:
        is_cov = np.isin(keys, cov, assume_unique=True)
        removable = np.nonzero(~is_cov)[0]
        drop = rng.choice(removable, size=keys.size - n_edges, replace=False)
        keep = np.ones(keys.size, dtype=bool)
        keep[drop] = False
        keys = keys[keep]
    users = (keys // m_items).astype(np.int32)
    items = (keys % m_items).astype(np.int32)
    return users, items


def make_embeddings(n_users: int, m_items: int, dim: int, seed: int = 2020, trained_like: bool = False):
    """N(0, 0.1^2) fp32 tables as PT/model.py:112-113.  ``trained_like`` rescales rows so that
    raw scores span roughly +-10 like the shipped mlls weights (non-degenerate rankings)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    u = torch.empty(n_users, dim).normal_(std=0.1, generator=g)
    i = torch.empty(m_items, dim).normal_(std=0.1, generator=g)
    if trained_like:
        u = u * torch.empty(n_users, 1).uniform_(1.0, 6.0, generator=g)
        i = i * torch.empty(m_items, 1).uniform_(1.0, 6.0, generator=g)
    return u, i


def make_test_dict(n_users: int, m_items: int, users, items, per_user: int = 5, seed: int = 2021):
    """Held-out items per user (not in train) -> {user: [items]} like Loader.testDict."""
    rng = np.random.default_rng(seed)
    train = set((np.asarray(users, dtype=np.int64) * m_items + np.asarray(items, dtype=np.int64)).tolist())
    out = {}
    for u in range(n_users):
        got = []
        while len(got) < per_user:
            c = int(rng.integers(0, m_items))
            if u * m_items + c not in train and c not in got:
                got.append(c)
        out[u] = got
    return out
