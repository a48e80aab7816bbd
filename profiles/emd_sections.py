"""Per-section clock64 totals of the EMD kernel for single LPs of different sizes (setup, phase init, argmin, settle,
augment, relax, phases, waves).

Needs the profiling build: make -C <package>/csrc EXTRA=-DMARSB200_EMD_PROFILE (touch emd.cu first), which adds
marsb200_debug_emd_profile(); rebuild without EXTRA afterwards."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, marsb200
from marsb200 import ops, _lib
raw = ctypes.CDLL(_lib.lib._name)
dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
E = 1
eps = [marsb200.make_episode(shape, 40 + i, dev) for i in range(E)]
b = marsb200.stack_episodes(eps)
n, g = shape.N, shape.g
fs = ops.normalize_rows(b["feat_s"].reshape(E, n, shape.C)); fq = ops.normalize_rows(b["feat_q"])
row_fg = ops.pool_mask(b["support_mask"], g).reshape(E, n)
cost = ops.sim_contract(fs, fq, n, n, shape.C, want_sim=False, want_cost=True)["cost"]
bits = ops.pack_masks(b["masks"]); pooled, area, cnt = ops.pool_packed(bits, shape.H, shape.W, g)
T = int(row_fg.sum())
order = torch.argsort(cnt[0], descending=True)
names = ["setup", "phase init", "argmin", "settle", "augment", "-", "relax", "rest", "phases", "waves"]
for rank in (0, 64, 128, 255):
    p = int(order[rank])
    one = pooled[0:1, p:p + 1].contiguous()
    ops.emd_scores(cost[0:1], row_fg[0:1], one, t_cap=T)
    buf = (ctypes.c_longlong * 16)()
    raw.marsb200_debug_emd_profile(buf, 1)
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.emd_scores(cost[0:1], row_fg[0:1], one, t_cap=T); c.record(); torch.cuda.synchronize()
    raw.marsb200_debug_emd_profile(buf, 1)
    print(f"T={T} M={int(cnt[0, p])}: {a.elapsed_time(c):.2f} ms | " + " ".join(f"{n}={buf[i]}" for i, n in enumerate(names)))
