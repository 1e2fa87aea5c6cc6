"""Configuration shim with the reference's keys and defaults (PT/world.py:32-71, PT/parse.py:8-46).

The reference parses sys.argv at import time and keeps everything in module globals; here the same
names exist but nothing is parsed on import -- call ``configure(...)`` (or edit ``config``) instead.
"""
from __future__ import annotations

import multiprocessing

import torch

config = {
    "bpr_batch_size": 2048,       # --bpr_batch
    "latent_dim_rec": 64,         # --recdim
    "lightGCN_n_layers": 3,       # --layer
    "dropout": 0,                 # --dropout
    "keep_prob": 0.6,             # --keepprob
    "A_n_fold": 100,              # --a_fold
    "test_u_batch_size": 100,     # --testbatch
    "multicore": 0,
    "lr": 0.001,
    "decay": 1e-4,
    "pretrain": 0,
    "A_split": False,             # PT/world.py:49
    "bigdata": False,
    # engine-only keys (absent from the reference)
    "score_mode": "bf16x3",       # fp32 | bf16 | bf16x3 for the fused top-K
    "fused_adam": True,           # BPRLoss uses lgx_adam_step_dev (one fused pass, device-side step counter) instead of
                                  # torch.optim.Adam; parity-tested against the reference's first step
    "cuda_graph": True,           # BPR_train_original replays one CUDA graph per full batch (needs fused_adam; off
                                  # automatically with edge dropout, whose seed is a host value per call)
}
GPU = torch.cuda.is_available()
device = torch.device("cuda" if GPU else "cpu")
CORES = multiprocessing.cpu_count() // 2
seed = 2020
dataset = "gowalla"
model_name = "lgn"
TRAIN_epochs = 1000
TRAIN_patience = 5
LOAD = 0
PATH = "./checkpoints"
topks = [20]
tensorboard = 0
comment = "lgn"


def configure(**kw):
    """Set config keys / module globals by name, e.g. configure(lightGCN_n_layers=4, topks=[20, 50])."""
    g = globals()
    for k, v in kw.items():
        if k in config:
            config[k] = v
        elif k in g:
            g[k] = v
        else:
            raise KeyError(f"unknown option {k}")


def cprint(words: str):
    print(f"\033[0;30;43m{words}\033[0m")
