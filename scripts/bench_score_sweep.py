"""configs[4]: full user x item scoring + top-K sweep, d = 64/128/256, 4096-user batches, 2M items sharded over
8 B200 -> every rank scores the batch against its 250K-item shard (lgx_score_topk with item_offset) and the
[B, K] candidate lists are merged (lgx_topk_merge).  This script times ONE rank's shard on one GPU, all modes,
and checks the result against an fp32 torch reference on a slice of the batch."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factors_of_serendipity_recommendation_b200 import _lgx

B, M_TOTAL, P, K = 4096, 2_000_000, 8, 20
M = M_TOTAL // P
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops": 1590.0}
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
rows = []
for d in (64, 128, 256):
    g = torch.Generator(device="cuda").manual_seed(d)
    U = torch.randn(B, d, device="cuda", generator=g) * 0.3
    I = torch.randn(M, d, device="cuda", generator=g) * 0.3
    ref = (U[:64].double() @ I.double().t())
    ref_top = torch.topk(ref, K).values[:, -1]
    for mode in ("bf16", "bf16x3", "fp32"):
        mid = _lgx.MODES[mode]
        if mode == "bf16x3" and d > 128:
            continue                                   # user tile (3d) + 2 item stages exceed 227 KB of shared memory
        def run():
            if mid == 0:
                return _lgx.score_topk(None, U, None, I, d, K, mid, item_offset=3 * M)
            Uo = _lgx.pack_operand(U, None, mid, False)
            Io = _lgx.pack_operand(I, None, mid, True)
            return _lgx.score_topk(None, Uo, None, Io, d, K, mid, item_offset=3 * M)
        for _ in range(3):
            idx, val = run()
        ts = []
        for _ in range(10 if mid else 3):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); idx, val = run(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        got = ref[torch.arange(64, device="cuda")[:, None], idx[:64] - 3 * M]
        tol = {"bf16": 1e-2, "bf16x3": 1e-5, "fp32": 2e-6}[mode] * ref.abs().max().item()
        ok = bool((got >= ref_top[:, None] - tol).all())
        tf = 2.0 * B * M * d / (ms * 1e-3) / 1e12
        rows.append({"d": d, "mode": mode, "ms_per_batch_shard": round(ms, 4), "users_per_s_per_gpu_shard": round(B / (ms * 1e-3)),
                     "tflops": round(tf, 1), "frac_of_bf16_peak": round(tf / peaks["bf16_tflops"], 4), "topk_valid": ok})
        print(json.dumps(rows[-1]), flush=True)
print(json.dumps({"config": f"B={B} users x {M} items (1/8 of {M_TOTAL}), K={K}, one B200 = one rank's shard", "rows": rows}))
