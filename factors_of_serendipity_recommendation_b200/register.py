"""Plugin registry with the reference's shape (PT/register.py:25-28): MODELS[name](config, dataset).

The reference loads its dataset at import time from ``../data/<name>``; here ``load_dataset`` does it
on request.  A reference checkout can swap its own entry with
``register.MODELS['lgn'] = factors_of_serendipity_recommendation_b200.model.LightGCN``."""
from __future__ import annotations

from . import dataloader, model, world

MODELS = {
    "mf": model.PureMF,
    "lgn": model.LightGCN,
}

dataset = None


def load_dataset(path=None, device=None):
    global dataset
    path = path or ("../data/" + world.dataset)
    dataset = dataloader.Loader(config=world.config, path=path, device=device)
    return dataset
