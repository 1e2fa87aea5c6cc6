"""N > 1 plumbing on CPU: world_size-2 gloo processes run the row partition, the per-layer all-gather
and the item-sharded top-K exchange of parallel.py.  The embedding arithmetic is done here with torch
CPU ops (a test double for the CUDA kernels); what is under test is the sharding/relabelling/exchange."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from factors_of_serendipity_recommendation_b200 import parallel, synth
    from oracle import lightgcn_oracle as O

    nu, mi, d, L, k = 151, 222, 16, 3, 10                     # N = 373: odd -> one padding row at world 2
    u, i = synth.make_interactions(nu, mi, 4000, seed=5)
    ue, ie = synth.make_embeddings(nu, mi, d, seed=5, trained_like=True)
    ref = O.OracleLightGCN(nu, mi, u, i, latent_dim=d, n_layers=L, user_emb=ue, item_emb=ie)
    indptr = torch.from_numpy(ref.indptr)
    indices = torch.from_numpy(ref.indices)
    values = torch.from_numpy(ref.data)
    order = torch.from_numpy(O.degree_sorted_row_order(ref.degree, ref.indptr))
    N = nu + mi

    ptr, cols, vals, n_local, new_id, old_of_new = parallel.shard_csr(indptr, indices, values, order, rank, world)
    # partition invariants
    assert n_local == (N + world - 1) // world
    assert torch.equal(old_of_new[new_id], torch.arange(N))
    mine = old_of_new[rank * n_local:(rank + 1) * n_local]
    assert torch.equal(mine[mine >= 0], order.long()[rank::world])
    nnz_all = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(nnz_all, torch.tensor([cols.numel()]))
    assert sum(int(x) for x in nnz_all) == indices.numel()
    assert max(int(x) for x in nnz_all) <= 1.15 * indices.numel() / world       # degree-cyclic deal balances nnz

    # SHARDED BUILD (no rank holds the whole graph): every rank starts from an arbitrary half of the interaction list and
    # must end up with exactly the row block shard_csr cuts out of the canonical CSR -- bit for bit
    sel = torch.from_numpy(np.random.default_rng(9).permutation(len(u)))[rank::world]
    ptr2, cols2, vals2, n_local2, new_id2, old2, deg2 = parallel.build_local_csr(
        nu, mi, torch.from_numpy(u)[sel], torch.from_numpy(i)[sel], rank, world)
    assert n_local2 == n_local and torch.equal(new_id2, new_id) and torch.equal(old2, old_of_new)
    assert torch.equal(deg2, torch.from_numpy(ref.degree).long())
    assert torch.equal(ptr2, ptr) and torch.equal(cols2, cols) and torch.equal(vals2, vals)

    A_local = torch.sparse_csr_tensor(ptr, cols.long(), vals, (n_local, world * n_local))
    E0 = torch.cat([ue, ie])
    X = torch.zeros(world * n_local, d)
    X[new_id] = E0
    S = X[rank * n_local:(rank + 1) * n_local].clone()
    for l in range(L):
        Y = A_local @ X
        S = S + Y
        parts = [torch.empty_like(Y) for _ in range(world)]
        dist.all_gather(parts, Y)
        X = torch.cat(parts)
    S = S / (L + 1)
    parts = [torch.empty_like(S) for _ in range(world)]
    dist.all_gather(parts, S)
    light = torch.cat(parts)[new_id]
    with torch.no_grad():
        ru, ri = ref.computer()
    full = torch.cat([ru, ri])
    assert (light - full).abs().max() <= 1e-5 * full.abs().max()

    # item-sharded scoring + candidate exchange
    lo, hi = parallel.item_shard_bounds(mi, rank, world)
    users = torch.arange(nu)
    score = (light[:nu] @ light[nu:][lo:hi].t())
    for r, items in enumerate(ref.all_pos(users.numpy())):
        sel = items[(items >= lo) & (items < hi)] - lo
        score[r, sel] = float("-inf")
    v, ix = torch.topk(score, k)
    # the product's exchange: ids and scores packed into ONE int32 collective (parallel.gather_packed)
    all_i, all_v = parallel.gather_packed(ix + lo, v, world)
    assert all_i.dtype == torch.int64 and all_v.dtype == torch.float32 and all_i.shape == (world,) + tuple(ix.shape)
    assert torch.equal(all_i[rank], ix + lo) and torch.equal(all_v[rank], v)          # bit-exact round trip
    midx, mval = parallel.merge_candidates_reference(all_i, all_v, k)
    s_full = (full[:nu].double() @ full[nu:].double().t()).numpy()
    for r, items in enumerate(ref.all_pos(users.numpy())):
        s_full[r, items] = -np.inf
    scale = np.abs(s_full[np.isfinite(s_full)]).max()
    assert all(O.topk_is_valid(s_full[r], midx[r].numpy(), k, tol=1e-5 * scale) for r in range(nu))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_row_sharded_propagation_and_item_sharded_topk_gloo(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_partition_and_merge_single_process():
    sys.path.insert(0, REPO)
    from factors_of_serendipity_recommendation_b200 import parallel
    order = torch.tensor([4, 0, 2, 1, 3], dtype=torch.int32)
    n_local, new_id, old = parallel.partition_rows(order, 2)
    assert n_local == 3 and old.tolist() == [4, 2, 3, 0, 1, -1] and new_id.tolist() == [3, 4, 1, 2, 0]
    assert parallel.item_shard_bounds(10, 0, 4) == (0, 3) and parallel.item_shard_bounds(10, 3, 4) == (9, 10)
    ci = torch.tensor([[[5, 1]], [[7, 9]]])
    cv = torch.tensor([[[2.0, 1.0]], [[2.0, 0.5]]])
    idx, val = parallel.merge_candidates_reference(ci, cv, 3)
    assert idx.tolist() == [[5, 7, 1]] and val.tolist() == [[2.0, 2.0, 1.0]]
    # sharded build at world 1 == the oracle's canonical CSR, duplicates merged into multiplicities (value 2)
    from oracle import lightgcn_oracle as O
    uu = np.array([0, 0, 1, 2, 2, 2, 3], dtype=np.int32)
    ii = np.array([1, 1, 0, 2, 0, 1, 2], dtype=np.int32)      # (0, 1) appears twice
    indptr, indices, data, degree = O.build_norm_adj(4, 3, uu, ii)
    ptr, cols, vals, n_local, new_id, old, deg = parallel.build_local_csr(4, 3, torch.from_numpy(uu), torch.from_numpy(ii), 0, 1)
    order = torch.from_numpy(O.degree_sorted_row_order(degree, indptr))
    # world 1: local row l is global row old[l]; columns come back relabelled through new_id
    assert torch.equal(deg, torch.from_numpy(degree).long())
    for l in range(n_local):
        g_row = int(old[l])
        a, b = int(ptr[l]), int(ptr[l + 1])
        assert old[cols[a:b].long()].tolist() == indices[indptr[g_row]:indptr[g_row + 1]].tolist()
        assert np.array_equal(vals[a:b].numpy(), data[indptr[g_row]:indptr[g_row + 1]])
    # packing keeps every bit of the scores (incl. -inf padding and the -1 "no item" id)
    pi = torch.tensor([[3, -1, 2_000_000_000]]); pv = torch.tensor([[1.5, float("-inf"), -1024.0]])
    ui, uv = parallel.unpack_candidates(parallel.pack_candidates(pi, pv))
    assert torch.equal(ui, pi) and torch.equal(uv, pv)


def test_upload_slices_tile_the_stacked_table():
    """serving.upload_slices (sharded host upload at N > 1): over all ranks the slices reproduce the stacked
    [users; items] table exactly once, in the padded all-gather layout rank * chunk + local row."""
    from factors_of_serendipity_recommendation_b200.serving import upload_slices
    rng = np.random.default_rng(0)
    cases = [(52643, 91599, 8), (10, 3, 4), (3, 10, 4), (1, 1, 2), (5, 0, 3), (7, 9, 1), (100, 100, 7)]
    cases += [(int(rng.integers(1, 500)), int(rng.integers(0, 500)), int(rng.integers(1, 17))) for _ in range(200)]
    for nu, mi, world in cases:
        n_rows = nu + mi
        users = np.arange(nu) + 1_000_000
        items = np.arange(mi) + 2_000_000
        stacked = np.concatenate([users, items])
        chunk = None
        gathered = None
        for rank in range(world):
            c, up, ip = upload_slices(rank, world, n_rows, nu)
            if chunk is None:
                chunk = c
                gathered = np.full(chunk * world, -1, dtype=np.int64)
            assert c == chunk and chunk * world >= n_rows
            local = np.full(chunk, -1, dtype=np.int64)
            for part, host in ((up, users), (ip, items)):
                if part is not None:
                    d0, d1, s0, s1 = part
                    assert 0 <= d0 < d1 <= chunk and 0 <= s0 < s1 <= len(host) and d1 - d0 == s1 - s0
                    assert np.all(local[d0:d1] == -1)                   # user and item parts do not overlap
                    local[d0:d1] = host[s0:s1]
            gathered[rank * chunk:(rank + 1) * chunk] = local
        assert np.array_equal(gathered[:n_rows], stacked)
        assert np.all(gathered[n_rows:] == -1)                          # only padding is left unwritten
