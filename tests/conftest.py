import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLD = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def read_interactions(path):
    """'uid item item ...' lines -> (users, items) int32 arrays, file order (PT/dataloader.py:247-262)."""
    us, its = [], []
    with open(path) as f:
        for line in f:
            parts = line.strip("\n").split(" ")
            if len(parts) < 2 or parts[1] == "":
                continue
            items = [int(x) for x in parts[1:]]
            us.extend([int(parts[0])] * len(items))
            its.extend(items)
    return np.array(us, dtype=np.int32), np.array(its, dtype=np.int32)


@pytest.fixture(scope="session")
def mlls():
    tu, ti = read_interactions(os.path.join(GOLD, "mlls_train.txt"))
    eu, ei = read_interactions(os.path.join(GOLD, "mlls_test.txt"))
    n_users = int(max(tu.max(), eu.max())) + 1
    m_items = int(max(ti.max(), ei.max())) + 1
    test_dict = {}
    for u, i in zip(eu.tolist(), ei.tolist()):
        test_dict.setdefault(u, []).append(i)
    return dict(train_user=tu, train_item=ti, n_users=n_users, m_items=m_items, test_dict=test_dict)


@pytest.fixture(scope="session")
def kat():
    return dict(np.load(os.path.join(GOLD, "mlls_kat.npz")))


@pytest.fixture(scope="session")
def train_step():
    return dict(np.load(os.path.join(GOLD, "mlls_train_step.npz")))


@pytest.fixture(scope="session")
def synth_small():
    return dict(np.load(os.path.join(GOLD, "synth_small.npz")))
