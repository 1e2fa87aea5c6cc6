"""Cycle accounting of the scoring kernel's roles (diagnostic build of lgx_score_gq.cu with -DLGX_GQ_PROF, loaded
through LGX_LIB_PATH; never the shipped library).  Prints per-tile averages for a list of LGX_GQ_DEBUG settings."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from factors_of_serendipity_recommendation_b200 import _lgx, synth

lib = ctypes.CDLL(_lgx.LIB_PATH)
nu, mi, E, d = synth.SHAPES["amazon-book"]
d = int(os.environ.get("PROF_D", d))
u, i = synth.make_interactions(nu, mi, E, seed=2020)
g = _lgx.Graph.build(nu, mi, torch.from_numpy(u), torch.from_numpy(i))
gen = torch.Generator(device="cuda").manual_seed(1)
U = torch.empty(nu, d, device="cuda").normal_(std=0.1, generator=gen)
I = torch.empty(mi, d, device="cuda").normal_(std=0.1, generator=gen)
users = torch.arange(nu, device="cuda")
Uo = _lgx.pack_operand(U, None, _lgx.SCORE_BF16, False)
Io = _lgx.pack_operand(I, None, _lgx.SCORE_BF16, True)
out = (ctypes.c_ulonglong * 32)()
N = 5
for dbg in [int(x) for x in sys.argv[1:]] or [0]:
    os.environ["LGX_GQ_DEBUG"] = str(dbg)
    _lgx.score_topk(g, Uo, users, Io, d, 20, _lgx.SCORE_BF16)
    torch.cuda.synchronize()
    lib.lgx_debug_gq_prof(out, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(N):
        _lgx.score_topk(g, Uo, users, Io, d, 20, _lgx.SCORE_BF16)
    b.record()
    torch.cuda.synchronize()
    lib.lgx_debug_gq_prof(out, 1)
    c = [int(x) for x in out]
    tiles = max(1, c[6])
    r = lambda x: round(x, 1)
    print(json.dumps({"dbg": dbg, "d": d, "ms_per_call": r(a.elapsed_time(b) / N * 1000) / 1000, "tiles_per_call": tiles // N,
                      "mma": {"loop": r(c[1] / tiles), "wait3": r(c[0] / tiles), "real_issue": r(c[2] / tiles),
                              "mask_part": r((c[1] - c[0] - c[2]) / tiles),
                              "wait_each": {"tempty": r(c[3] / tiles), "mfull": r(c[4] / tiles), "full": r(c[5] / tiles)}},
                      "epi": {"wait_tfull": r(c[8] / 8 / tiles), "tfull_to_release": r(c[9] / 8 / tiles), "after_release": r(c[10] / 8 / tiles)},
                      "builder": {"busy_per_built_tile": r((c[13] - c[12]) / tiles), "reclaim_wait_per_built_tile": r(c[12] / tiles)},
                      "tma": {"wait_empty": r(c[16] / tiles)}}), flush=True)
