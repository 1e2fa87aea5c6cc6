"""Per-CTA wall-clock phases of the scoring kernel (diagnostic -DLGX_GQ_PROF build through LGX_LIB_PATH)."""
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
Uo = _lgx.pack_operand(U, None, _lgx.SCORE_BF16, False)
Io = _lgx.pack_operand(I, None, _lgx.SCORE_BF16, True)
for _ in range(3):
    _lgx.score_topk(g, Uo, None, Io, d, 20, _lgx.SCORE_BF16)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (1024 * 8))()
assert lib.lgx_debug_gq_cta(buf) == 0
T = np.array(buf, dtype=np.int64).reshape(1024, 8)[:412]
np.save("gpurun_out/gq_cta.npy", T)
t0 = T[:, 1].min()
start, loop0, loop1, epi, end = [(T[:, k] - t0) / 1000.0 for k in range(1, 6)]   # us
print(f"kernel span {end.max():.1f} us; CTA duration mean {np.mean(end - start):.1f} min {np.min(end - start):.1f} max {np.max(end - start):.1f} us")
print(f"prologue (entry -> MMA loop) mean {np.mean(loop0 - start):.2f} max {np.max(loop0 - start):.2f} us; MMA loop mean {np.mean(loop1 - loop0):.1f} us; "
      f"loop end -> epilogue done mean {np.mean(epi - loop1):.2f} us; epilogue done -> exit mean {np.mean(end - epi):.2f} max {np.max(end - epi):.2f} us")
# per SM: CTAs in order, gaps between one CTA's exit and the next one's entry
gaps, per_sm_end = [], []
for sm in np.unique(T[:, 0]):
    idx = np.where(T[:, 0] == sm)[0]
    idx = idx[np.argsort(start[idx])]
    per_sm_end.append(end[idx[-1]])
    for a, b in zip(idx[:-1], idx[1:]):
        gaps.append(start[b] - end[a])
gaps = np.array(gaps)
print(f"SMs used {len(per_sm_end)}; gap exit -> next entry mean {gaps.mean():.2f} max {gaps.max():.2f} us; first-wave start spread {start[np.argsort(start)[:148]].max():.2f} us")
per_sm_end = np.array(per_sm_end)
print(f"per-SM finish: min {per_sm_end.min():.1f} median {np.median(per_sm_end):.1f} max {per_sm_end.max():.1f} us; CTAs per SM: {np.bincount(np.unique(T[:,0], return_counts=True)[1])}")
dur = end - start
mhz = (T[:, 7] - T[:, 6]) / np.maximum(dur, 1e-9)          # SM cycles per microsecond over the CTA's life
wave = np.argsort(np.argsort(start)) // 148
for w in range(3):
    sel = wave == w
    print(f"wave {w}: CTAs {sel.sum()} duration mean {dur[sel].mean():.1f} us, cycles mean {(T[sel, 7] - T[sel, 6]).mean():.0f}, effective SM clock mean {mhz[sel].mean():.0f} MHz (min {mhz[sel].min():.0f}, max {mhz[sel].max():.0f})")
order = np.argsort(dur)
print("slowest CTAs (tile, us):", [(int(k), round(float(dur[k]), 1)) for k in order[-5:]], " fastest:", [(int(k), round(float(dur[k]), 1)) for k in order[:3]])
