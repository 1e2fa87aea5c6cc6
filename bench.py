#!/usr/bin/env python
"""bench.py -- LightGCN forward (3-layer propagation) + full-catalogue top-20 scoring.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over the whole synthetic graph of the named shape:
computer() (L fused SpMM layers) followed by fused score + train-mask + top-20 for EVERY user.
Prints ONE JSON line (see the round prompt for the contract).  metric = users scored top-20 / s
over the whole step (propagation included); the SpMM propagated-edges/s + HBM GB/s and the scoring
TFLOP/s are in the `spmm` / `scoring` / `roofline*` objects of the same line.

--impl reference times the reference's CPU code path (oracle port: the same torch.sparse.mm /
matmul / topk calls, PT/Procedure.py:121-135 as written) on the host cores, one 100-user Test batch
per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

K_TOP = 20
N_LAYERS = 3


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sustained=j.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_traffic(workload, mode):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/traffic.json, written by scripts/summarize_profile.py); None when no capture exists."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    t = json.load(open(p)).get(f"{workload}:{mode}", {})
    return t.get("spmm_layer_bytes"), t.get("score_bytes")


def spmm_algorithmic_bytes(n_nodes, nnz, d, n_layers):
    """per layer nnz*(4+4) + (N+1)*4 + 2*N*d*4; forward adds one write of the mean."""
    layer = nnz * 8 + (n_nodes + 1) * 4 + 2 * n_nodes * d * 4
    return layer, n_layers * layer + n_nodes * d * 4


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args, shape):
    import torch
    from factors_of_serendipity_recommendation_b200 import synth
    from oracle import lightgcn_oracle as O

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nu, mi, E, d = shape
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    u, i = synth.make_interactions(nu, mi, E, seed=2020)
    ue, ie = synth.make_embeddings(nu, mi, d, seed=2020)
    ref = O.OracleLightGCN(nu, mi, u, i, latent_dim=d, n_layers=N_LAYERS, user_emb=ue, item_emb=ie)
    batch = 100                                             # --testbatch default, PT/parse.py:26
    rng = np.random.default_rng(0)
    times = []
    with torch.no_grad():
        for s in range(args.warmup + args.steps):
            users = rng.choice(nu, size=batch, replace=False)
            t0 = time.perf_counter()
            rating = ref.getUsersRating(torch.from_numpy(users))          # recomputes computer(), PT/model.py:180
            O.mask_and_topk(rating, ref.all_pos(users), K_TOP)            # PT/Procedure.py:129-135
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = batch * len(times) / total
    line = {
        "impl": "reference", "metric": "users scored top-20/sec (3-layer propagation + full-catalogue scoring)",
        "value": value, "unit": "users/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "n_users": nu, "m_items": mi, "edges": E, "d": d, "layers": N_LAYERS, "k": K_TOP},
        "cpu_baseline": {"value": value, "unit": "users/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} Test batches of {batch} users as written in PT/Procedure.py:121-135 "
                                   "(getUsersRating recomputes computer() per batch), torch CPU ops, all host threads"},
        "e2e": {"value": value, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(shape, u, i, ue, ie, budget_s=20.0):
    """Oracle port on the host cores, bounded: one computer() + as many 100-user batches as fit ~budget."""
    import torch
    from oracle import lightgcn_oracle as O

    nu, mi, E, d = shape
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = O.OracleLightGCN(nu, mi, u, i, latent_dim=d, n_layers=N_LAYERS, user_emb=ue, item_emb=ie)
    rng = np.random.default_rng(0)
    with torch.no_grad():
        t0 = time.perf_counter()
        au, ai = ref.computer()
        t_prop = time.perf_counter() - t0
        n_b, t_batches, t_score_only = 0, 0.0, 0.0
        while n_b < 2 or (t_batches < budget_s and n_b < 12):
            users = rng.choice(nu, size=100, replace=False)
            t0 = time.perf_counter()
            rating = ref.getUsersRating(torch.from_numpy(users))
            O.mask_and_topk(rating, ref.all_pos(users), K_TOP)
            t_batches += time.perf_counter() - t0
            t0 = time.perf_counter()
            rating = O.users_rating(au, ai, torch.from_numpy(users))
            O.mask_and_topk(rating, ref.all_pos(users), K_TOP)
            t_score_only += time.perf_counter() - t0
            n_b += 1
    as_written = 100 * n_b / t_batches
    hoisted = nu / (t_prop + (nu / 100.0) * (t_score_only / n_b))
    return {"value": as_written, "unit": "users/s", "cores": cores, "kind": "port",
            "sample": f"{n_b} Test batches of 100 users as written (PT/Procedure.py:121-135: computer() recomputed per "
                      f"batch); computer() alone {t_prop * 1e3:.0f} ms = {N_LAYERS * 2 * E / t_prop / 1e6:.1f} M edges/s; "
                      f"with computer() hoisted out of the loop the same code would give {hoisted:.0f} users/s",
            "propagate_ms": t_prop * 1e3, "edges_per_s": N_LAYERS * 2 * E / t_prop, "hoisted_users_per_s": hoisted}


def check_timed_launch(idx, emb, nu, mi, u, i, mode, n_sample=512):
    """Check sampled users of the step's result: fp64 CPU scores from the same propagated embeddings, train items
    masked, tests/-style 'identical up to ties' (tolerance 1e-2 bf16, 1e-5 bf16x3, 2e-6 fp32)."""
    import torch
    from oracle import lightgcn_oracle as O
    tol = {"fp32": 2e-6, "bf16x3": 1e-5, "bf16": 1e-2}[mode]
    rng = np.random.default_rng(11)
    rows = np.sort(rng.choice(nu, size=min(n_sample, nu), replace=False))
    au = emb[:nu][torch.from_numpy(rows).to(emb.device)].double().cpu().numpy()
    ai = emb[nu:].double().cpu().numpy()
    s = au @ ai.T
    indptr = np.concatenate([[0], np.cumsum(np.bincount(u, minlength=nu))])
    for r, uid in enumerate(rows):
        s[r, i[indptr[uid]:indptr[uid + 1]]] = -np.inf
    scale = float(np.abs(s[np.isfinite(s)]).max())
    got = idx[torch.from_numpy(rows).to(idx.device)].cpu().numpy()
    bad = sum(0 if O.topk_is_valid(s[r], got[r], K_TOP, tol=tol * scale) else 1 for r in range(len(rows)))
    return {"ok": bad == 0, "users_checked": int(len(rows)), "violations": int(bad), "tolerance_rel": tol,
            "against": "fp64 CPU scores of the same propagated embeddings, train items masked"}


def check_layer_rows(graph, X, n_rows=64, seed=5):
    """One SpMM layer of `graph` on X, sampled output rows against an fp64 gather of the exported CSR rows (all on the
    device, independent of the SpMM kernel).  -> dict for the JSON line."""
    import torch
    Y = torch.empty(graph.n_rows, X.shape[1], device=X.device)
    graph.spmm(X, Y=Y)
    e = graph.export()
    gen = torch.Generator(device="cpu").manual_seed(seed)
    rows = torch.randint(0, graph.n_rows, (n_rows,), generator=gen).to(X.device)
    rows = torch.cat([rows, e["row_order"][:2].long(), e["row_order"][-2:].long()])       # + the longest and shortest rows
    worst = 0.0
    for r in rows.tolist():
        a, b = int(e["indptr"][r]), int(e["indptr"][r + 1])
        terms = e["values"][a:b].double()[:, None] * X[e["indices"][a:b].long()].double()
        ref = terms.sum(0)
        scale = terms.abs().sum(0).max().item() + 1e-30
        worst = max(worst, (Y[r].double() - ref).abs().max().item() / scale)
    del e, Y
    return {"ok": worst <= 1e-5, "rows_checked": int(rows.numel()), "max_rel_err": worst, "tolerance": 1e-5,
            "against": "fp64 gather of the exported CSR rows (torch, on the device)"}


def north_star_scale_leg(args, rank, world_size, dev, dist):
    """configs[3]: 10 M users / 2 M items / 1e9 edges, d = 128, 3 layers; row-sharded with the fused peer-store
    exchange at N > 1.  Returns the dict for the JSON line (or {'error': ...})."""
    import torch
    from factors_of_serendipity_recommendation_b200 import _lgx, synth
    try:
        nu, mi, E, d = synth.SHAPES[args.scale_workload]
        t0 = time.perf_counter()
        u_d, i_d = synth.make_interactions_device(nu, mi, E, seed=2020, device=dev)
        g = _lgx.Graph.build(nu, mi, u_d, i_d, chunk_nnz=args.chunk)
        del u_d, i_d
        torch.cuda.empty_cache()
        build_s = time.perf_counter() - t0
        gen = torch.Generator(device=dev).manual_seed(2020)
        E0 = torch.empty(nu + mi, d, device=dev).normal_(std=0.1, generator=gen)
        N, nnz = g.n_rows, g.nnz
        if world_size > 1:
            from factors_of_serendipity_recommendation_b200 import parallel
            engine = parallel.ShardedEngine(g, nu, mi, d, N_LAYERS, rank, world_size, dev, propagate=args.propagate)
            del g
            torch.cuda.empty_cache()
            engine.propagate(E0)                      # fills engine.X[0] with the relabelled layer-0 embeddings
            run = lambda: engine.propagate(E0)
            mode = engine.mode
        else:
            engine = None
            out = torch.empty_like(E0)
            run = lambda: g.propagate_fwd(E0, N_LAYERS, out=out)
            mode = "single GPU"
        # parity at this shape: one layer of the graph this rank propagates (its row shard at N > 1), sampled rows
        if engine is None:
            parity = check_layer_rows(g, E0)
        elif getattr(engine, "local", None) is not None:
            parity = check_layer_rows(engine.local, engine.X[0])
        else:
            parity = check_layer_rows(engine.g_full, E0)
        torch.cuda.empty_cache()
        for _ in range(2):
            run()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        reps = 3
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            run()
        b.record()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        if dist is not None:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = tt.item()
        layer_bytes, _ = spmm_algorithmic_bytes(N, nnz, d, N_LAYERS)
        res = {"workload": args.scale_workload, "n_users": nu, "m_items": mi, "edges": E, "d": d, "layers": N_LAYERS,
               "n_gpus": world_size, "mode": mode, "ms_propagate": ms, "edges_per_s": N_LAYERS * nnz / (ms * 1e-3),
               "hbm_gbs_per_gpu_algorithmic": layer_bytes / world_size / (ms / N_LAYERS * 1e-3) / 1e9,
               "parity_rank0": parity, "graph_build_s_per_rank": build_s, "timing": "CUDA events, max over ranks, 3 repetitions after 2 warm-ups; "
               "the graph (8 GB of indices) is far larger than L2"}
        if engine is not None:
            engine.close()
        return res
    except Exception as e:       # the headline line must not be lost to the extra leg
        return {"error": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, shape):
    import torch
    import torch.distributed as dist
    from factors_of_serendipity_recommendation_b200 import build as _build

    rank = int(os.environ.get("RANK", "0"))
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        _build.build(verbose=False)          # no-op when liblgx.so is newer than its sources
    from factors_of_serendipity_recommendation_b200 import _lgx, dataloader, model, synth, world

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size != args.gpus and world_size > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world_size}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        import datetime
        # a bounded collective timeout: a rank that dies must not hold the other GPUs for the default 10 minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
        dist.barrier()                        # ranks > 0 wait for rank 0's build check
    peaks = load_peaks()
    nu, mi, E, d = shape
    mode = args.mode
    mode_id = _lgx.MODES[mode]

    big = E > 20_000_000                      # generated on the device, never visits the host
    huge = E >= 500_000_000                   # tables of several GB: no host copies, no e2e leg
    cfg = dict(world.config)
    cfg.update(lightGCN_n_layers=N_LAYERS, latent_dim_rec=d, score_mode=mode)
    if big:
        u_d, i_d = synth.make_interactions_device(nu, mi, E, seed=2020, device=dev)
        g = _lgx.Graph.build(nu, mi, u_d, i_d, chunk_nnz=args.chunk)
        del u_d, i_d
        torch.cuda.empty_cache()
        ds = dataloader.InteractionDataset(nu, mi, None, None, device=dev, graph=g, train_size=E)
        u = i = None
    else:
        u, i = synth.make_interactions(nu, mi, E, seed=2020)
        ds = dataloader.InteractionDataset(nu, mi, u, i, device=dev)
    if huge:
        m, ue, ie = None, None, None
        gen = torch.Generator(device=dev).manual_seed(2020)
        E0_huge = torch.empty(nu + mi, d, device=dev).normal_(std=0.1, generator=gen)     # PT/model.py:112-113
    else:
        ue, ie = synth.make_embeddings(nu, mi, d, seed=2020)
        cfg.update(pretrain=1, user_emb=ue.numpy(), item_emb=ie.numpy())
        m = model.LightGCN(cfg, ds).to(dev).eval()
    g = ds.getGraphHandle()
    N, nnz = g.n_rows, g.nnz
    n_score = nu if nu <= 200_000 else 65_536      # huge graphs: score a 65 536-user batch per step (configs[4] style)
    all_users = torch.arange(n_score, dtype=torch.int64, device=dev)
    # the whole user range in order is the identity batch: passed as users=None, so the scoring call can keep the
    # graph's train mask in its tile-bucketed form (built once per graph, like the reference's allPos)
    users_arg = None if n_score == nu else all_users
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    if world_size > 1:
        from factors_of_serendipity_recommendation_b200 import parallel
        engine = parallel.ShardedEngine(g, nu, mi, d, N_LAYERS, rank, world_size, dev, propagate=args.propagate)
        torch.cuda.empty_cache()
    else:
        engine = None

    E0 = E0_huge if huge else m._flat_if_fused()
    assert E0 is not None
    light = torch.empty_like(E0)

    def step_resident():
        """inputs resident in HBM; returns event triplet timings handled by the caller"""
        if engine is not None:
            return engine.step(E0, all_users, K_TOP, mode_id, shard=args.shard)
        g.propagate_fwd(E0, N_LAYERS, out=light)
        au, ai = light[:nu], light[nu:]
        if mode_id == _lgx.SCORE_FP32:
            U_op, I_op = au[:n_score], ai
        else:
            I_op = _lgx.pack_operand(ai, None, mode_id, True)
            U_op = _lgx.pack_operand(au, users_arg, mode_id, False)
        return _lgx.score_topk(g, U_op, users_arg, I_op, d, K_TOP, mode_id)

    def timed_step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        flush.fill_(1)                                        # evict L2 between timed iterations
        if engine is not None:
            ev[0].record()
            engine.step(E0, all_users, K_TOP, mode_id, events=ev, shard=args.shard)
            ev[3].record()
            return ev
        ev[0].record()
        g.propagate_fwd(E0, N_LAYERS, out=light)
        ev[1].record()
        au, ai = light[:nu], light[nu:]
        if mode_id == _lgx.SCORE_FP32:
            U_op, I_op = au[:n_score], ai
        else:
            I_op = _lgx.pack_operand(ai, None, mode_id, True)
            U_op = _lgx.pack_operand(au, users_arg, mode_id, False)
        ev[2].record()
        _lgx.score_topk(g, U_op, users_arg, I_op, d, K_TOP, mode_id)
        ev[3].record()
        return ev

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the launch that is about to be timed is checked first: sampled users against fp64 CPU scores of the same
    # propagated embeddings with the users' train items masked ("identical up to ties", tolerance per mode)
    parity = None
    if not huge and u is not None and not args.no_check:
        # every rank runs the step (it contains collectives at N > 1); rank 0 checks its copy of the result
        if engine is None:
            chk_idx, chk_val = step_resident()
            chk_emb = light
        else:
            chk_idx, chk_val = engine.step(E0, all_users, K_TOP, mode_id, shard=args.shard)
            chk_emb = engine.last_light()
        torch.cuda.synchronize()
        if rank == 0:
            parity = check_timed_launch(chk_idx, chk_emb, nu, mi, u, i, mode)
            if not parity["ok"]:
                print(f"bench.py: the timed launch fails its parity check: {parity}", file=sys.stderr, flush=True)
        del chk_idx, chk_val
    barrier()

    for _ in range(max(args.warmup, 3)):
        timed_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [timed_step() for _ in range(args.steps)]
    barrier()
    # ---- sustained leg: the same step repeated back to back for >= --min-seconds so that clocks settle under load
    # (the K-step burst above lasts tens of ms at boost clock); reported beside the burst number
    sustained = None
    if args.min_seconds > 0:
        n_rep = max(args.steps, int(args.min_seconds * 1e3 / max(1e-3, sum(e[0].elapsed_time(e[3]) for e in evs) / len(evs))))
        if world_size > 1:
            tt = torch.tensor([n_rep], device=dev, dtype=torch.int64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            n_rep = int(tt.item())
        sus_sampler = ClockSampler(local_rank)
        if rank == 0:
            sus_sampler.start()
        s_evs = [timed_step() for _ in range(n_rep)]
        barrier()
        sus_clocks = sus_sampler.stop() if rank == 0 else None
        sus_ms = sum(e[0].elapsed_time(e[3]) for e in s_evs) / n_rep
        sus_score = sum(e[2].elapsed_time(e[3]) for e in s_evs) / n_rep
        sus_prop = sum(e[0].elapsed_time(e[1]) for e in s_evs) / n_rep
        if world_size > 1:
            tt = torch.tensor([sus_ms, sus_score, sus_prop], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sus_ms, sus_score, sus_prop = tt.tolist()
        sustained = {"steps": n_rep, "ms_per_step": sus_ms, "value": n_score / (sus_ms * 1e-3), "scoring_ms": sus_score,
                     "propagate_ms": sus_prop, "clocks": sus_clocks,
                     "note": "same step, L2 flushed before every step, back to back for >= --min-seconds"}
    t_step = [e[0].elapsed_time(e[3]) for e in evs]
    t_prop = [e[0].elapsed_time(e[1]) for e in evs]
    t_pack = [e[1].elapsed_time(e[2]) for e in evs]
    t_score = [e[2].elapsed_time(e[3]) for e in evs]
    total_ms = sum(t_step)
    if world_size > 1:
        tt = torch.tensor([total_ms, sum(t_prop), sum(t_pack), sum(t_score)], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)             # max over ranks, device-timed
        total_ms, sp, spk, ss = tt.tolist()
        t_prop_mean, t_pack_mean, t_score_mean = sp / args.steps, spk / args.steps, ss / args.steps
    else:
        t_prop_mean, t_pack_mean, t_score_mean = (statistics.mean(x) for x in (t_prop, t_pack, t_score))
    ms_per_step = total_ms / args.steps
    value = n_score / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with HOST buffers (H2D of the tables, D2H of the top-20)
    e2e_total = None
    if not huge:
        host_u, host_i = ue.clone().pin_memory(), ie.clone().pin_memory()
        host_out = torch.empty(n_score, K_TOP, dtype=torch.int64).pin_memory()

        # the public host-facing call: serving.HostPipeline.submit(host tables, host result buffer).  Each step's
        # H2D (both tables), L2 flush, propagation + scoring and D2H are inside the timed region; consecutive steps
        # overlap their copies with the neighbours' kernels (depth-2 device slots).
        from factors_of_serendipity_recommendation_b200 import serving
        if engine is not None:
            up_group = dist.new_group(backend="nccl") if args.e2e_upload == "sharded" else None
            pipe = serving.HostPipeline.for_engine(engine, N, d, all_users, K_TOP, mode_id, shard=args.shard,
                                                   depth=args.e2e_depth, flush_l2=flush, upload_group=up_group)
        else:
            pipe = serving.HostPipeline.for_model(m, all_users, K_TOP, mode=mode, depth=args.e2e_depth, flush_l2=flush)
        host_outs = [host_out] + [torch.empty_like(host_out).pin_memory() for _ in range(args.e2e_depth)]

        for step_no in range(3):
            pipe.submit(host_u, host_i, host_outs[step_no % len(host_outs)])
        pipe.wait()
        barrier()
        t_a = pipe.start_event()
        for step_no in range(args.steps):
            pipe.submit(host_u, host_i, host_outs[step_no % len(host_outs)])
        t_b = pipe.last_download
        pipe.wait()
        barrier()
        e2e_total = t_a.elapsed_time(t_b)
        if world_size > 1:
            tt = torch.tensor([e2e_total], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_total = tt.item()
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = n_score / (e2e_total / args.steps * 1e-3) if e2e_total else None

    if rank == 0:
        layer_bytes, fwd_bytes = spmm_algorithmic_bytes(N, nnz, d, N_LAYERS)
        spmm_gbs = layer_bytes / (t_prop_mean / N_LAYERS * 1e-3) / 1e9       # per layer launch (dominant SpMM kernel)
        flops = 2.0 * n_score * mi * d
        # per GPU: at N > 1 every rank scores 1/N of the (user x item) pairs in t_score_mean (max over ranks)
        score_tf = flops / world_size / (t_score_mean * 1e-3) / 1e12
        # replicated propagation: every rank does the whole layer; sharded modes: 1/N of the rows per rank
        prop_share = 1.0 if (engine is None or engine.mode == "replicated") else 1.0 / world_size
        spmm_gbs = spmm_gbs * prop_share
        tr_spmm, tr_score = measured_traffic(args.workload, mode) if world_size == 1 else (None, None)
        roof_spmm = {"bound": "hbm", "achieved": spmm_gbs, "peak": peaks["hbm"], "unit": "GB/s",
                     "frac": spmm_gbs / peaks["hbm"], "traffic": tr_spmm, "peak_source": peaks["src"],
                     "kernel": "k_spmm_fixed x3 (+ k_spmm_long); achieved/traffic are per layer launch: "
                               "algorithmic bytes of one layer / (propagation time / layers)",
                     "launch_ms": t_prop_mean / N_LAYERS, "algorithmic_bytes": layer_bytes,
                     "gather_ceiling_note": "random row gathers top out at 17-18.5 TB/s from L2 / 7.1 TB/s from HBM "
                                            "(profiles/r1_l2_gather_probe.txt)",
                     "no_reuse_gather_bytes": N_LAYERS * (nnz * (8 + 4 * d) + N * d * 4)}
        roof_score = {"bound": "tensor", "achieved": score_tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                      "frac": score_tf / peaks["tf_burst"], "traffic": tr_score, "peak_source": peaks["src"] + " (burst)",
                      "kernel": (("k_score_topk_gq + k_rescore_topk (k_mask_buckets: once per graph for the identity batch)"
                                  if (engine is None and users_arg is None) else "k_mask_buckets + k_score_topk_gq + k_rescore_topk")
                                 if mode_id else "k_score_topk_fp32 + merge"),
                      "launch_ms": t_score_mean, "algorithmic_flops": flops / world_size, "per_gpu": True}
        if sustained is not None and peaks["tf_sustained"]:
            s_tf = flops / world_size / (sustained["scoring_ms"] * 1e-3) / 1e12
            sustained["scoring_tflops_per_gpu"] = s_tf
            sustained["scoring_frac_of_sustained_peak"] = s_tf / peaks["tf_sustained"]
            sustained["spmm_gbs_per_gpu"] = layer_bytes * prop_share / (sustained["propagate_ms"] / N_LAYERS * 1e-3) / 1e9
            sustained["spmm_frac"] = sustained["spmm_gbs_per_gpu"] / peaks["hbm"]
        dominant = roof_score if t_score_mean >= t_prop_mean else roof_spmm
        # our kernels per step: L x (SpMM [+ k_spmm_long]) + 2 x k_pack + [k_mask_buckets] + k_score_topk_gq + k_rescore_topk
        # (k_mask_buckets only for explicit user batches: the identity batch's buckets are built once, before the timed
        # region, and kept with the graph; fp32 mode: k_score_topk_fp32 + k_topk_merge)
        buckets_per_step = 0 if (engine is None and users_arg is None) else 1
        launches_per_step = N_LAYERS * (1 + (1 if g.n_long > 0 else 0)) + ((2 + 2 + buckets_per_step) if mode_id else 2)
        line = {
            "metric": "users scored top-20/sec (3-layer propagation + full-catalogue scoring)",
            "value": value, "unit": "users/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 propagation, " + {"fp32": "f32", "bf16": "bf16", "bf16x3": "bf16x3 (hi/lo split)"}[mode] + " scoring",
            "data": "synthetic",
            "config": {"workload": args.workload, "n_users": nu, "m_items": mi, "edges": E, "nnz": nnz, "d": d,
                       "layers": N_LAYERS, "k": K_TOP, "users_scored_per_step": n_score, "score_mode": mode, "l2": "flushed between steps (512 MB fill)",
                       "parallelism": "1 GPU" if world_size == 1 else
                       f"propagation {engine.mode} (fused = SpMM epilogue stores into peer memory over NVLink; allgather = NCCL per layer), "
                       f"scoring sharded by {args.shard} x{world_size}"},
            "spmm": {"propagated_edges_per_s": N_LAYERS * nnz / (t_prop_mean * 1e-3), "hbm_gbs": spmm_gbs,
                     "ms": t_prop_mean, "layers": N_LAYERS},
            "scoring": {"users_per_s": n_score / (t_score_mean * 1e-3), "tflops": score_tf, "ms": t_score_mean,
                        "pack_ms": t_pack_mean},
            "roofline": dominant, "roofline_spmm": roof_spmm, "roofline_scoring": roof_score,
            "e2e": {"value": e2e_value, "unit": "users/s", "h2d_bytes_per_step": int((nu + mi) * d * 4),
                    "d2h_bytes_per_step": int(n_score * K_TOP * 8),
                    "ms_per_step": e2e_total / args.steps if e2e_total else None,
                    "api": "serving.HostPipeline.submit(host_user_emb, host_item_emb, host_out)",
                    "pipeline_depth": args.e2e_depth, "l2_flush_inside": True,
                    "upload": (args.e2e_upload if world_size > 1 else "full")},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "sustained": sustained,
            "parity_check": parity,
        }
        if world_size == 1 and not args.no_cpu and not big:
            line["cpu_baseline"] = cpu_baseline_sample(shape, u, i, ue, ie)
    # ---- north-star scaling leg (BASELINE.json configs[3]): propagation on the 1B-edge graph at the same N
    scale_leg = None
    if args.scale_leg != "off" and args.workload == "amazon-book":
        if engine is not None:
            engine.close()
            engine = None
        del flush
        torch.cuda.empty_cache()
        scale_leg = north_star_scale_leg(args, rank, world_size, dev, dist if world_size > 1 else None)
    if rank == 0:
        line["north_star_scale"] = scale_leg
        print(json.dumps(line), flush=True)
    if world_size > 1:
        if engine is not None:
            engine.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="amazon-book")
    ap.add_argument("--mode", default="bf16", choices=["fp32", "bf16", "bf16x3"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-upload", default="sharded", choices=["sharded", "full"],
                    help="N > 1: each rank uploads 1/N of the table rows and the ranks all-gather over NVLink (sharded), "
                         "or every rank uploads the whole tables over its own PCIe link (full)")
    ap.add_argument("--e2e-depth", type=int, default=2,
                    help="device slots of the host pipeline in the e2e leg (1 = copies and kernels strictly serial)")
    ap.add_argument("--propagate", default="auto", choices=["auto", "overlap", "fused", "allgather", "replicated"],
                    help="propagation exchange at N > 1 (parallel.ShardedEngine)")
    ap.add_argument("--chunk", type=int, default=0, help="long-row split size for the graph build (0 = default 256)")
    ap.add_argument("--shard", default="auto", choices=["auto", "items", "users"], help="scoring split at N > 1")
    ap.add_argument("--min-seconds", type=float, default=2.0,
                    help="sustained leg: repeat the timed step back to back for at least this long (0 = off)")
    ap.add_argument("--no-check", action="store_true", help="skip the parity check of the timed launch")
    ap.add_argument("--scale-leg", default="auto", choices=["auto", "off"],
                    help="after the headline leg, time 3-layer propagation on the 1B-edge synthetic graph (configs[3]) "
                         "at the same N and report it as north_star_scale")
    ap.add_argument("--scale-workload", default="synth-1b")
    args = ap.parse_args()
    from factors_of_serendipity_recommendation_b200 import synth
    shape = synth.SHAPES[args.workload]
    if args.impl == "reference":
        run_reference(args, shape)
    else:
        run_ours(args, shape)


if __name__ == "__main__":
    main()
