"""Timeline of one CTA of the scoring kernel (diagnostic -DLGX_GQ_PROF build through LGX_LIB_PATH): per tile, when each
role waited / issued / released, in cycles relative to the first traced event."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from factors_of_serendipity_recommendation_b200 import _lgx, synth

lib = ctypes.CDLL(_lgx.LIB_PATH)
nu, mi, E, d = synth.SHAPES["amazon-book"]
u, i = synth.make_interactions(nu, mi, E, seed=2020)
g = _lgx.Graph.build(nu, mi, torch.from_numpy(u), torch.from_numpy(i))
gen = torch.Generator(device="cuda").manual_seed(1)
U = torch.empty(nu, d, device="cuda").normal_(std=0.1, generator=gen)
I = torch.empty(mi, d, device="cuda").normal_(std=0.1, generator=gen)
users = torch.arange(nu, device="cuda")
Uo = _lgx.pack_operand(U, None, _lgx.SCORE_BF16, False)
Io = _lgx.pack_operand(I, None, _lgx.SCORE_BF16, True)
for _ in range(3):
    _lgx.score_topk(g, Uo, users, Io, d, 20, _lgx.SCORE_BF16)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (12 * 64 * 6))()
assert lib.lgx_debug_gq_trace(buf) == 0
T = np.array(buf, dtype=np.int64).reshape(12, 64, 6)
np.save("gpurun_out/gq_trace.npy", T)
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 20, int(sys.argv[2]) if len(sys.argv) > 2 else 32
t0 = T[0, lo, 0]
rel = lambda x: int(x - t0) if x else -1
print("tile | MMA: wait_start tempty_ok mfull_ok full_ok real_start real_end | builder: reclaim_start reclaim_ok published | "
      "epilogue (min..max over 8 warps): tfull_seen release finish")
for t in range(lo, hi):
    m = [rel(x) for x in T[0, t]]
    b = [rel(x) for x in T[1 + (t & 1), t, :3]]
    e = T[4:12, t, :4]
    seen, relz, fin, wst = e[:, 1] - t0, e[:, 2] - t0, e[:, 3] - t0, e[:, 0] - t0
    print(f"{t:3d} | {m} | {b} | wait_from {wst.min()}..{wst.max()} seen {seen.min()}..{seen.max()} release {relz.min()}..{relz.max()} "
          f"finish {fin.min()}..{fin.max()}")
