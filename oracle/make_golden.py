"""Generate tests/golden/* by running the UNMODIFIED reference from /root/reference.

Run once in the build container (the GPU box has no /root/reference):

    python oracle/make_golden.py

It stages the reference's PyTorch LightGCN sources (PT/ = lightGCN/LightGCN-PyTorch-master/code)
in a temp dir (they are imported from there, never copied into this repo), imports
``world, utils, register, model, Procedure`` with the bootstrap from SURVEY.md appendix A,
and records reference outputs as small fixtures:

  mlls_train.txt / mlls_test.txt    dataset shipped with the reference (TF/Data/mlls)
  mlls_s_pre_adj_mat.npz            adjacency shipped with the reference (golden CSR)
  mlls_kat.npz                      shipped weights -> computer() / ratings / top-20 / Test()
  mlls_train_step.npz               one bpr_loss + backward + Adam step of the reference
  synth_small.npz                   reference Loader + model on a synthetic graph with
                                    duplicate edges and isolated nodes
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PT = f"{REF}/lightGCN/LightGCN-PyTorch-master/code"
MLLS = f"{REF}/LightGCN-tf/Data/mlls"
WEIGHTS = f"{REF}/LightGCN-tf/weights/mlls/LightGCN/64-64-64-64/l0.01_r1e-05-1e-05-0.01"
GOLD = f"{REPO}/tests/golden"


def write_interactions(path, n_users, users, items):
    rows = [[] for _ in range(n_users)]
    for u, i in zip(users, items):
        rows[u].append(int(i))
    with open(path, "w") as f:
        for u, r in enumerate(rows):
            if r:
                f.write(" ".join([str(u)] + [str(x) for x in r]) + "\n")


def main():
    sys.path.insert(0, REPO)
    from factors_of_serendipity_recommendation_b200 import synth

    os.makedirs(GOLD, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="lgn_ref_")
    os.makedirs(f"{tmp}/code/sources")
    os.makedirs(f"{tmp}/data/mlls")
    for fn in os.listdir(PT):
        if fn.endswith(".py"):
            shutil.copy(f"{PT}/{fn}", f"{tmp}/code/{fn}")
    for fn in ("train.txt", "test.txt"):
        shutil.copy(f"{MLLS}/{fn}", f"{tmp}/data/mlls/{fn}")
        shutil.copy(f"{MLLS}/{fn}", f"{GOLD}/mlls_{fn}")
    shutil.copy(f"{MLLS}/s_pre_adj_mat.npz", f"{GOLD}/mlls_s_pre_adj_mat.npz")

    # synthetic small graph with duplicates and isolated nodes -> its own dataset dir
    su, si = synth.make_interactions(300, 500, 6000, seed=7)
    keep = (su != 17) & (su != 299) & (si != 3) & (si != 499)      # isolated user 17, item 3; max ids still present in test
    su, si = su[keep], si[keep]
    dup = np.arange(0, su.size, 97)                                  # duplicate some edges (summed by the reference)
    su = np.concatenate([su, su[dup], su[dup[:5]]]).astype(np.int32)  # a few edges three times
    si = np.concatenate([si, si[dup], si[dup[:5]]]).astype(np.int32)
    os.makedirs(f"{tmp}/data/synth_small")
    write_interactions(f"{tmp}/data/synth_small/train.txt", 300, su, si)
    tu = np.array([0, 5, 17, 299], dtype=np.int32)                   # test file pins n_user=300, m_item=500
    ti = np.array([1, 2, 3, 499], dtype=np.int32)
    write_interactions(f"{tmp}/data/synth_small/test.txt", 300, tu, ti)

    os.chdir(f"{tmp}/code")
    sys.path.insert(0, f"{tmp}/code")
    sys.argv = ["x", "--dataset", "mlls", "--tensorboard", "0", "--load", "0", "--topks", "[20]",
                "--layer", "4", "--pretrain", "1"]
    import torch
    import world
    world.config["user_emb"] = np.load(f"{WEIGHTS}/emb_user.npy")
    world.config["item_emb"] = np.load(f"{WEIGHTS}/emb_item.npy")
    world.device = torch.device("cpu")
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        import utils
        import register
        import model
        import Procedure
    ds = register.dataset

    # ---- KAT: shipped weights, 4 layers (TF/output/mlls/LightGCN.result:8)
    with contextlib.redirect_stdout(sink):
        m = model.LightGCN(world.config, ds).eval()
        res = Procedure.Test(ds, m, 0, None, 0)
    with torch.no_grad():
        lu, li = m.computer()
        users = list(ds.testDict.keys())
        ut = torch.tensor(users, dtype=torch.long)
        rating = m.getUsersRating(ut)
        all_pos = ds.getUserPosItems(users)
        ex_i, ex_j = [], []
        for r, items in enumerate(all_pos):
            ex_i.extend([r] * len(items))
            ex_j.extend(items)
        rating[ex_i, ex_j] = -(1 << 10)
        vals, idx = torch.topk(rating, k=20)
    print("KAT", res)
    g = m.Graph
    np.savez_compressed(
        f"{GOLD}/mlls_kat.npz",
        emb_user=world.config["user_emb"], emb_item=world.config["item_emb"],
        light_users=lu.numpy(), light_items=li.numpy(),
        test_users=np.array(users, dtype=np.int64), topk_idx=idx.numpy(), topk_val=vals.numpy(),
        rating_first8=rating[:8].numpy(),
        precision=res["precision"], recall=res["recall"], ndcg=res["ndcg"],
        graph_vals=g.values().numpy(),   # values as THIS container's numpy computes them (<= 1 ulp from the shipped npz)
        n_layers=np.int64(4))

    # ---- one training step, 3 layers, N(0, 0.1) init, seed 2020
    world.config["pretrain"] = 0
    world.config["lightGCN_n_layers"] = 3
    utils.set_seed(2020)
    with contextlib.redirect_stdout(sink):
        m3 = model.LightGCN(world.config, ds)
    w0u = m3.embedding_user.weight.detach().clone().numpy()
    w0i = m3.embedding_item.weight.detach().clone().numpy()
    bpr = utils.BPRLoss(m3, world.config)
    np.random.seed(2020)
    S = utils.UniformSample_original_python(ds)[:2048]
    bu, bp, bn = (torch.tensor(S[:, k]).long() for k in range(3))
    m3.train()
    loss, reg = m3.bpr_loss(bu, bp, bn)
    total = loss + reg * world.config["decay"]
    bpr.opt.zero_grad()
    total.backward()
    gu = m3.embedding_user.weight.grad.detach().clone().numpy()
    gi = m3.embedding_item.weight.grad.detach().clone().numpy()
    bpr.opt.step()
    with torch.no_grad():
        lu3, li3 = m3.computer()
        gamma = m3.forward(bu[:64], bp[:64])
    np.savez_compressed(
        f"{GOLD}/mlls_train_step.npz",
        w0_user=w0u, w0_item=w0i, users=S[:, 0], pos=S[:, 1], neg=S[:, 2],
        loss=loss.item(), reg_loss=reg.item(), decay=world.config["decay"], lr=world.config["lr"],
        grad_user=gu, grad_item=gi,
        w1_user=m3.embedding_user.weight.detach().numpy(), w1_item=m3.embedding_item.weight.detach().numpy(),
        gamma_after=gamma.numpy(),
        n_layers=np.int64(3))
    print("train step loss", loss.item(), "reg", reg.item())

    # ---- synthetic small graph through the reference Loader (duplicates, isolated nodes)
    import dataloader
    with contextlib.redirect_stdout(sink):
        ds2 = dataloader.Loader(path=f"{tmp}/data/synth_small")
        utils.set_seed(11)
        m2 = model.LightGCN(world.config, ds2).eval()
    assert ds2.n_users == 300 and ds2.m_items == 500, (ds2.n_users, ds2.m_items)
    g2 = m2.Graph
    with torch.no_grad():
        lu2, li2 = m2.computer()
        some = torch.tensor([0, 1, 5, 17, 100, 299], dtype=torch.long)
        r2 = m2.getUsersRating(some)
    np.savez_compressed(
        f"{GOLD}/synth_small.npz",
        train_user=su, train_item=si, n_users=np.int64(300), m_items=np.int64(500),
        graph_rows=g2.indices()[0].numpy(), graph_cols=g2.indices()[1].numpy(), graph_vals=g2.values().numpy(),
        users_D=ds2.users_D, items_D=ds2.items_D,
        allpos_len=np.array([len(x) for x in ds2.allPos], dtype=np.int64),
        allpos_flat=np.concatenate(ds2.allPos).astype(np.int64),
        w_user=m2.embedding_user.weight.detach().numpy(), w_item=m2.embedding_item.weight.detach().numpy(),
        light_users=lu2.numpy(), light_items=li2.numpy(),
        rating_users=some.numpy(), rating=r2.numpy(), n_layers=np.int64(3))
    print("synth_small nnz", g2._nnz())
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
