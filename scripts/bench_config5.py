"""BASELINE.json configs[4] for real: full user x item scoring + top-K sweep, d = 64 / 128 / 256, 4096-user batches,
2 M items ITEM-SHARDED across the N GPUs of the box (one process per GPU, torchrun):

    every rank scores the batch against its slice of the catalogue (lgx_score_topk with item_offset),
    the [B, K] candidate lists are exchanged in ONE packed all-gather (parallel.gather_packed) and merged
    (lgx_topk_merge) -- SURVEY.md section 8(e) row 4.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/bench_config5.py [--items 2000000] [--batch 4096] [--batches 8]

Per d and mode it prints one JSON line: per-batch ms (max over ranks, CUDA events) split into scoring / exchange+merge,
users/s, per-GPU TFLOP/s and its fraction of the measured bf16 peak, and a parity verdict on sampled rows against fp64
scores over the WHOLE catalogue (tolerance 1e-2 bf16 / 1e-5 bf16x3 of the score scale)."""
import argparse
import datetime
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from factors_of_serendipity_recommendation_b200 import _lgx, parallel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=2_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--batches", type=int, default=8)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--dims", type=int, nargs="*", default=[64, 128, 256])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops": 1590.0}
    B, M, K = args.batch, args.items, args.k
    lo, hi = parallel.item_shard_bounds(M, rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for d in args.dims:
        g = torch.Generator(device=dev).manual_seed(1000 + d)          # same tables on every rank
        U = torch.empty(B * args.batches, d, device=dev).normal_(std=0.1, generator=g)
        I = torch.empty(M, d, device=dev).normal_(std=0.1, generator=g)
        I *= torch.empty(M, 1, device=dev).uniform_(0.5, 2.0, generator=g)
        shard = I[lo:hi].contiguous()
        for mode in ("bf16", "bf16x3"):
            mid = _lgx.MODES[mode]
            if mode == "bf16x3" and d > 128:
                continue                                   # 3d = 768 columns do not fit the kernel's resident user tile
            Io = _lgx.pack_operand(shard, None, mid, True)

            def one_batch(b, ev=None):
                Ub = U[b * B:(b + 1) * B]
                Uo = _lgx.pack_operand(Ub, None, mid, False)
                idx, val = _lgx.score_topk(None, Uo, None, Io, d, K, mid, item_offset=lo)
                if ev is not None:
                    ev[1].record()
                if world > 1:
                    ai, av = parallel.gather_packed(idx, val, world)
                    idx, val = _lgx.topk_merge(ai, av)
                return idx, val

            for b in range(min(2, args.batches)):
                one_batch(b)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t_tot, t_score, last = [], [], None
            for b in range(args.batches):
                flush.fill_(1)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
                last = one_batch(b, ev)
                ev[2].record()
                torch.cuda.synchronize()
                t_tot.append(ev[0].elapsed_time(ev[2]))
                t_score.append(ev[0].elapsed_time(ev[1]))
            ms, ms_score = statistics.median(t_tot), statistics.median(t_score)
            if world > 1:
                tt = torch.tensor([ms, ms_score], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms, ms_score = tt.tolist()
            # parity: 48 rows of the last batch against fp64 scores over the whole catalogue (rank 0)
            ok = None
            if rank == 0:
                idx, val = last
                rows = torch.arange(0, B, B // 48, device=dev)[:48]
                Ub = U[(args.batches - 1) * B:][rows].double()
                tol = (1e-2 if mode == "bf16" else 1e-5)
                ok = True
                for c in range(0, rows.numel(), 16):                      # 16 x 2 M fp64 scores at a time
                    s = Ub[c:c + 16] @ I.double().t()
                    scale = s.abs().max().item()
                    kth = torch.topk(s, K).values[:, -1]
                    got = torch.gather(s, 1, idx[rows[c:c + 16]])
                    ok &= bool((got >= kth[:, None] - tol * scale).all().item())
                    must = (s > (kth[:, None] + tol * scale)).sum(1)
                    ok &= bool(((got > (kth[:, None] + tol * scale)).sum(1) == must).all().item())
                    ok &= bool((idx[rows[c:c + 16]].sort(1).values.diff(dim=1) != 0).all().item())
                tf = 2.0 * B * (hi - lo) * d / (ms_score * 1e-3) / 1e12
                print(json.dumps({"config": "configs[4]", "d": d, "mode": mode, "n_gpus": world, "batch_users": B, "items": M,
                                  "items_per_rank": hi - lo, "k": K, "ms_per_batch": round(ms, 4),
                                  "ms_scoring": round(ms_score, 4), "ms_exchange_merge": round(ms - ms_score, 4),
                                  "users_per_s": round(B / (ms * 1e-3)), "tflops_per_gpu": round(tf, 1),
                                  "frac_of_bf16_peak_per_gpu": round(tf / peaks["bf16_tflops"], 4),
                                  "exchange_bytes_per_rank": B * K * 8, "topk_valid_vs_fp64_full_catalogue": ok}), flush=True)
        del U, I, shard
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
