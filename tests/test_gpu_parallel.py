"""ShardedEngine (row-sharded SpMM + NCCL all-gather, item-sharded top-K + merge) vs the 1-GPU path.
Runs with as many ranks as there are GPUs on the box (1 on the default GPU test box, 2+ under
`gpurun --gpus N`); the world-size-2 plumbing is also covered on CPU by tests/test_parallel_cpu.py."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from factors_of_serendipity_recommendation_b200 import _lgx, dataloader, model, parallel, synth, world as W
    nu, mi, d, L, k = 1501, 2222, 64, 3, 20
    u, i = synth.make_interactions(nu, mi, 40000, seed=5)
    ue, ie = synth.make_embeddings(nu, mi, d, seed=5, trained_like=True)
    cfg = dict(W.config)
    cfg.update(pretrain=1, user_emb=ue.numpy(), item_emb=ie.numpy(), lightGCN_n_layers=L)
    ds = dataloader.InteractionDataset(nu, mi, u, i, device=dev)
    m = model.LightGCN(cfg, ds).to(dev).eval()
    g = ds.getGraphHandle()
    E0 = m._flat_if_fused()
    users = torch.arange(nu, device=dev)
    with torch.no_grad():
        lu, li = m.computer()
        ref = torch.cat([lu, li])
        for mode in ("overlap", "fused", "replicated", "allgather"):          # every exchange variant gives the 1-GPU result
            eng = parallel.ShardedEngine(g, nu, mi, d, L, rank, world, dev, propagate=mode)
            assert eng.mode == mode
            for rep in range(2):                                    # twice: buffers are reused across calls
                light = eng.propagate(E0)
                assert (light - ref).abs().max().item() <= 1e-5 * ref.abs().max().item(), (mode, rep)
            if mode != "allgather":
                eng.close()
        # sharded graph build: every rank starts from its slice of the interaction list, never sees the full graph
        ut, it_ = torch.from_numpy(u), torch.from_numpy(i)
        for mode in ("fused", "allgather"):
            eng2 = parallel.ShardedEngine.from_edge_partition(nu, mi, ut[rank::world], it_[rank::world], d, L, rank, world,
                                                              dev, propagate=mode)
            for rep in range(2):
                light = eng2.propagate(E0)
                assert (light - ref).abs().max().item() <= 1e-5 * ref.abs().max().item(), ("edge partition", mode, rep)
            eng2.close()
        for mode in ("fp32", "bf16x3"):
            idx1, val1 = m.topk(users, k, mode=mode)
            for shard in ("items", "users"):
                idxN, valN = eng.score(ref, users, k, _lgx.MODES[mode], shard=shard)   # same embeddings -> identical lists
                assert torch.equal(idx1, idxN) and torch.equal(val1, valN), (mode, shard)
        idx, val = eng.step(E0, users, k, _lgx.MODES["bf16x3"])
        same = (idx == m.topk(users, k, mode="bf16x3")[0]).float().mean().item()
        assert same > 0.999                                               # propagate differs in the last ulp only
        # host pipeline over the engine: sharded upload (1/world of the rows per rank + all-gather on a dedicated
        # group) and full upload return the same ids for every request, in order
        from factors_of_serendipity_recommendation_b200 import serving
        reqs = [tuple(t.pin_memory() for t in synth.make_embeddings(nu, mi, d, seed=s, trained_like=True)) for s in range(4)]
        got = {}
        for name, group in (("full", None), ("sharded", dist.new_group(backend="nccl"))):
            pipe = serving.HostPipeline.for_engine(eng, nu + mi, d, users, k, _lgx.MODES["bf16x3"], upload_group=group)
            outs = [torch.empty(nu, k, dtype=torch.int64).pin_memory() for _ in reqs]
            for (hu, hi), out in zip(reqs, outs):
                pipe.submit(hu, hi, out)
            pipe.wait()
            got[name] = [o.clone() for o in outs]
        for a, b in zip(got["full"], got["sharded"]):
            assert torch.equal(a, b)
        assert not torch.equal(got["full"][0], got["full"][1])            # the requests really differ
    torch.cuda.synchronize()
    eng.close()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_engine_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    world = min(8, torch.cuda.device_count())
    port = 29600 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
