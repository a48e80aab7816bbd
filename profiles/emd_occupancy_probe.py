"""Does the EMD solver's throughput follow the number of problems in flight per SM?  Small problems (T, M <= 256) so that
the shared-memory state is 22 KB: the 256-thread build keeps 4 CTAs per SM (register file), a 128-thread build 8.
Build variants: make -C <pkg>/csrc EXTRA=-DMARSB200_EMD_THREADS=128 (touch emd.cu first)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200
from marsb200 import ops
from marsb200.synthetic import random_masks
dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
E, g, n = 8, shape.g, shape.N
gen = torch.Generator(device=dev).manual_seed(11)
b = marsb200.stack_episodes([marsb200.make_episode(shape, 50 + i, dev) for i in range(E)])
fs = ops.normalize_rows(b["feat_s"].reshape(E, n, shape.C)); fq = ops.normalize_rows(b["feat_q"])
cost = ops.sim_contract(fs, fq, n, n, shape.C, want_sim=False, want_cost=True)["cost"]
support = torch.stack([random_masks(1, shape.H, shape.W, gen, dev, 0.08, 0.14, dup_frac=0.0) for _ in range(E)])
row_fg = ops.pool_mask(support, g).reshape(E, n)
masks = torch.stack([random_masks(shape.P, shape.H, shape.W, gen, dev, 0.01, 0.12, dtype=torch.uint8) for _ in range(E)])
pooled, area, cnt = ops.pool_packed(ops.pack_masks(masks), shape.H, shape.W, g)
T, M = int(row_fg.sum(1).max()), int(cnt.max())
print("max T", T, "max M", M, "mean M", float(cnt.float().mean()))
tc, mc = (T + 31) // 32 * 32, (M + 31) // 32 * 32
out = ops.emd_scores(cost, row_fg, pooled, t_cap=tc, m_cap=mc)
torch.cuda.synchronize()
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    out = ops.emd_scores(cost, row_fg, pooled, t_cap=tc, m_cap=mc, check=False)
c.record(); torch.cuda.synchronize()
ms = a.elapsed_time(c) / 3
print(f"caps ({tc}, {mc}): {ms:.2f} ms for {E * shape.P} LPs = {E * shape.P / ms:.1f} k LP/s, checksum {float(out.sum()):.9f}")
