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
    "synth-10m": (100_000, 20_000, 10_000_000, 128),
    "synth-100m": (1_000_000, 200_000, 100_000_000, 128),
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


def make_interactions_device(n_users: int, m_items: int, n_edges: int, seed: int = 2020, device="cuda"):
    """Same distributional contract as make_interactions (unique pairs, every user and item covered,
    item popularity ~ rank^-0.8, user activity ~ log-normal(1)), generated with torch ops on the GPU for
    the shapes numpy cannot build in reasonable time (>= 10^7 edges).  Deterministic per (seed, shape).
    Returns int32 device tensors sorted by (user, item)."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    pop = torch.arange(1, m_items + 1, device=device, dtype=torch.float64) ** -0.8
    pop = pop[torch.randperm(m_items, device=device, generator=g)]
    pop_cdf = torch.cumsum(pop / pop.sum(), 0)
    act = torch.exp(torch.randn(n_users, device=device, generator=g, dtype=torch.float64))
    act_cdf = torch.cumsum(act / act.sum(), 0)

    def draw(cdf, k, hi):
        r = torch.rand(k, device=device, generator=g, dtype=torch.float64)
        return torch.searchsorted(cdf, r).clamp_(max=hi - 1)

    cov = torch.cat([torch.arange(n_users, device=device) * m_items + draw(pop_cdf, n_users, m_items),
                     draw(act_cdf, m_items, n_users) * m_items + torch.arange(m_items, device=device)])
    keys = torch.unique(cov)
    n_cov_unique = keys.numel()
    is_cov_sorted = None
    while keys.numel() < n_edges:
        need = n_edges - keys.numel()
        k = int(need * 1.25) + 4096
        extra = draw(act_cdf, k, n_users) * m_items + draw(pop_cdf, k, m_items)
        keys = torch.unique(torch.cat([keys, extra]))
    if keys.numel() > n_edges:
        cov_u = torch.unique(cov)
        pos = torch.searchsorted(cov_u, keys).clamp_(max=cov_u.numel() - 1)
        removable = torch.nonzero(cov_u[pos] != keys).squeeze(1)
        drop = removable[torch.randperm(removable.numel(), device=device, generator=g)[: keys.numel() - n_edges]]
        keep = torch.ones(keys.numel(), dtype=torch.bool, device=device)
        keep[drop] = False
        keys = keys[keep]
    return (keys // m_items).to(torch.int32), (keys % m_items).to(torch.int32)


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
