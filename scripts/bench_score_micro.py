"""Scoring-kernel experiments at the Amazon-Book shape (one process, variants switched through the environment).

    python scripts/bench_score_micro.py [--d 64] [--mode bf16] [--variants name=ENV1:V1,ENV2:V2 ...]

Times lgx_score_topk (CUDA events, L2 flushed, median of N) for each variant; LGX_GQ_DEBUG is re-read on every call,
the other switches only at process start, so variants that need them are run in a child process."""
import argparse, json, os, statistics, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_variant(args):
    import torch
    from factors_of_serendipity_recommendation_b200 import _lgx, synth
    nu, mi, E, _ = synth.SHAPES[args.workload]
    d = args.d
    mid = _lgx.MODES[args.mode]
    g = None
    if not args.nomask:
        u, i = synth.make_interactions(nu, mi, E, seed=2020)
        g = _lgx.Graph.build(nu, mi, torch.from_numpy(u), torch.from_numpy(i))
    gen = torch.Generator(device="cuda").manual_seed(1)
    U = torch.empty(nu, d, device="cuda").normal_(std=0.1, generator=gen)
    I = torch.empty(mi, d, device="cuda").normal_(std=0.1, generator=gen)
    if args.users:
        nu_s = min(nu, args.users)
        U = U[:nu_s].contiguous()
    else:
        nu_s = nu
    users = torch.arange(nu_s, device="cuda")
    Uo = _lgx.pack_operand(U, None, mid, False)
    Io = _lgx.pack_operand(I, None, mid, True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    sweep = [("dbg", x) for x in args.dbg] if args.splits is None else [("splits", x) for x in args.splits]
    for kind, dbg in sweep:
        if kind == "dbg":
            os.environ["LGX_GQ_DEBUG"] = str(dbg)
        else:
            os.environ["LGX_SCORE_SPLITS"] = str(dbg)
            if dbg == 0:
                os.environ.pop("LGX_SCORE_SPLITS")
        for _ in range(3):
            _lgx.score_topk(g, Uo, users, Io, d, args.k, mid)
        ts = []
        for _ in range(args.iters):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); _lgx.score_topk(g, Uo, users, Io, d, args.k, mid); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out[f"{kind}{dbg}"] = round(statistics.median(ts), 4)
    plan = _lgx.score_plan(nu_s, mi, d, args.k, mid)
    print(json.dumps({"variant": args.name, "d": d, "mode": args.mode, "nomask": args.nomask, "users": nu_s, "planner": plan, "ms": out}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="amazon-book")
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--nomask", action="store_true")
    ap.add_argument("--dbg", type=int, nargs="*", default=[0])
    ap.add_argument("--users", type=int, default=0, help="score only the first N users (0 = all)")
    ap.add_argument("--splits", type=int, nargs="*", default=None, help="sweep LGX_SCORE_SPLITS over these values (0 = planner)")
    ap.add_argument("--name", default="default")
    ap.add_argument("--variants", nargs="*", default=None, help="name=ENV:VAL,ENV:VAL ... each run in a child process")
    args = ap.parse_args()
    if args.variants is None:
        return run_variant(args)
    for v in args.variants:
        name, _, envs = v.partition("=")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, val = kv.partition(":")
            env[k] = val
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--d", str(args.d), "--k", str(args.k),
               "--mode", args.mode, "--iters", str(args.iters), "--name", name, "--users", str(args.users), "--dbg", *map(str, args.dbg)]
        if args.splits is not None:
            cmd += ["--splits", *map(str, args.splits)]
        if args.nomask:
            cmd.append("--nomask")
        subprocess.run(cmd, env=env, check=False, timeout=600)


if __name__ == "__main__":
    main()
