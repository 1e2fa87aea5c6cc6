"""Epoch loops behind the reference's names (PT/Procedure.py): BPR_train_original and Test."""
from __future__ import annotations

import numpy as np
import torch

from . import _lgx, utils, world
from .utils import timer


def BPR_train_original(dataset, recommend_model, loss_class, epoch, neg_k=1, w=None):
    """PT/Procedure.py:26-57.  Samples on the device when the dataset has a graph handle; the loss is
    accumulated on the device and read back once per epoch (the reference syncs every mini-batch)."""
    Recmodel = recommend_model
    Recmodel.train()
    bpr = loss_class
    with timer(name="Sample"):
        S = utils.UniformSample_original(dataset)
    dev = next(Recmodel.parameters()).device
    if not torch.is_tensor(S):
        S = torch.from_numpy(np.asarray(S)).long()
    S = S.to(dev)
    users, posItems, negItems = S[:, 0].contiguous(), S[:, 1].contiguous(), S[:, 2].contiguous()
    users, posItems, negItems = utils.shuffle(users, posItems, negItems)
    bs = world.config["bpr_batch_size"]
    total_batch = len(users) // bs + 1
    aver_loss = torch.zeros((), device=dev)
    # config['cuda_graph']: after 3 eager steps (lazy initialisation happens there) the full-size batches
    # replay one captured graph of the whole step; the last, shorter batch runs eagerly
    use_graph = (bool(world.config.get("cuda_graph", False)) and getattr(bpr, "fused", False) and dev.type == "cuda"
                 and not (getattr(Recmodel, "config", {}).get("dropout") and Recmodel.training))
    graphed = getattr(bpr, "_graphed", None)
    for batch_i, (bu, bp, bn) in enumerate(utils.minibatch(users, posItems, negItems, batch_size=bs)):
        if use_graph and len(bu) == bs and (graphed is not None or batch_i >= 3):
            if graphed is None:
                graphed = bpr._graphed = utils.GraphedStageOne(bpr, bu, bp, bn)
            cri = graphed.run(bu, bp, bn)
        else:
            cri = bpr.stageOne(bu, bp, bn, sync=False)
        aver_loss += cri
        if world.tensorboard and w is not None:
            w.add_scalar("BPRLoss/BPR", cri.item(), epoch * int(len(users) / bs) + batch_i)
    Recmodel._eval_cache = None
    Recmodel._packed = {}
    aver_loss = aver_loss.item() / total_batch
    time_info = timer.dict()
    timer.zero()
    return f"loss{aver_loss:.3f}-{time_info}"


def test_one_batch(X):
    """PT/Procedure.py:60-72."""
    sorted_items = X[0].numpy() if torch.is_tensor(X[0]) else np.asarray(X[0])
    groundTrue = X[1]
    r = utils.getLabel(groundTrue, sorted_items)
    pre, recall, ndcg = [], [], []
    for k in world.topks:
        ret = utils.RecallPrecision_ATk(groundTrue, r, k)
        pre.append(ret["precision"])
        recall.append(ret["recall"])
        ndcg.append(utils.NDCGatK_r(groundTrue, r, k))
    return {"recall": np.array(recall), "precision": np.array(pre), "ndcg": np.array(ndcg)}


def early_stopping(metrics_best, metric_test):
    """PT/Procedure.py:74-94: keep per-metric bests, stop when nothing improved."""
    need_stop = True
    for metric, best in metrics_best.items():
        if metric not in metric_test:
            continue
        for idx, (b, t) in enumerate(zip(best, metric_test[metric])):
            if b < t:
                need_stop = False
                metrics_best[metric][idx] = t
    return metrics_best, need_stop


def Test(dataset, Recmodel, epoch, w=None, multicore=0, device_metrics=False, mode=None):
    """PT/Procedure.py:96-174 -> {'precision','recall','ndcg'}: np.ndarray[len(topks)].

    Same result contract; the work is reorganised for the device: one propagation for the whole
    pass (the reference recomputes computer() per 100-user batch), fused score + mask + top-K per
    batch of ``test_u_batch_size`` users so that only [B, max_K] indices leave the GPU, metrics
    with the reference's numpy formulas (or lgx_rank_metrics when device_metrics=True)."""
    u_batch_size = world.config["test_u_batch_size"]
    testDict = dataset.testDict
    Recmodel = Recmodel.eval()
    max_K = max(world.topks)
    results = {m: np.zeros(len(world.topks)) for m in ("precision", "recall", "ndcg")}
    users = list(testDict.keys())
    if not users:
        return results
    dev = next(Recmodel.parameters()).device
    big = max(u_batch_size, 16384)                       # device batches; results are batch-size independent
    with torch.no_grad():
        rating_list, ground_list = [], []
        for start in range(0, len(users), big):
            batch_users = users[start:start + big]
            bu = torch.as_tensor(batch_users, dtype=torch.int64, device=dev)
            idx, _ = Recmodel.topk(bu, max_K, exclude_train=True, mode=mode)
            rating_list.append(idx)
            ground_list.append([testDict[u] for u in batch_users])
        if device_metrics:
            for idx, ground in zip(rating_list, ground_list):
                lens = np.array([len(g) for g in ground], dtype=np.int64)
                gt_ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)])).to(dev)
                gt_items = torch.from_numpy(np.concatenate([np.asarray(g, dtype=np.int64) for g in ground])).to(dev)
                for j, k in enumerate(world.topks):
                    sums = torch.zeros(3, dtype=torch.float64, device=dev)
                    _lgx.rank_metrics(idx, k, gt_ptr, gt_items, sums)
                    s = sums.cpu().numpy()
                    results["recall"][j] += s[0]
                    results["precision"][j] += s[1] / k
                    results["ndcg"][j] += s[2]
        else:
            for idx, ground in zip(rating_list, ground_list):
                part = test_one_batch((idx.cpu(), ground))
                for m in results:
                    results[m] += part[m]
    for m in results:
        results[m] /= float(len(users))
    if world.tensorboard and w is not None:
        for name, key in (("Recall", "recall"), ("Precision", "precision"), ("NDCG", "ndcg")):
            w.add_scalars(f"Test/{name}@{world.topks}",
                          {str(world.topks[i]): results[key][i] for i in range(len(world.topks))}, epoch)
    return results
