"""Time lgx_propagate_fwd (3 layers) for the current LGX_SPMM_VARIANT on a named synthetic shape."""
import os, sys, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factors_of_serendipity_recommendation_b200 import _lgx, synth

name = sys.argv[1] if len(sys.argv) > 1 else "amazon-book"
chunk = int(os.environ.get("LGX_CHUNK", "0"))
nu, mi, E, d = synth.SHAPES[name]
u, i = synth.make_interactions(nu, mi, E, seed=2020)
ue, ie = synth.make_embeddings(nu, mi, d, seed=2020)
g = _lgx.Graph.build(nu, mi, torch.from_numpy(u).cuda(), torch.from_numpy(i).cuda(), chunk_nnz=chunk)
E0 = torch.cat([ue, ie]).cuda()
out = torch.empty_like(E0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

def run(cold, reps=20):
    ts = []
    for _ in range(reps):
        if cold:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.propagate_fwd(E0, 3, out=out); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)

run(False, 5)
cold = run(True)
warm = run(False)
layer_bytes = g.nnz * 8 + (g.n_rows + 1) * 4 + 2 * g.n_rows * d * 4
fwd_bytes = 3 * layer_bytes + g.n_rows * d * 4
print(json.dumps({"variant": os.environ.get("LGX_SPMM_VARIANT", "0") + "/" + os.environ.get("LGX_SPMM_VARIANT128", "0"), "hot": os.environ.get("LGX_SPMM_HOT_DEGREE", "-"),
                  "chunk": g.chunk_nnz, "n_long": g.n_long, "shape": name,
                  "cold_ms_med": round(cold[0], 4), "cold_ms_min": round(cold[1], 4), "warm_ms_med": round(warm[0], 4),
                  "cold_gbs": round(fwd_bytes / cold[0] / 1e6, 1), "warm_gbs": round(fwd_bytes / warm[0] / 1e6, 1),
                  "cold_gedges": round(3 * g.nnz / cold[0] / 1e6, 2), "checksum": float(out.double().sum())}))
