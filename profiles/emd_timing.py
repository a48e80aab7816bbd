"""Device EMD at BASELINE c2 shapes: P = 256 transport LPs per episode (T fg support patches x M_p proposal patches)."""
import sys, os, statistics, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, marsb200
from marsb200 import ops
dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
E = int(sys.argv[1]) if len(sys.argv) > 1 else 2
eps = [marsb200.make_episode(shape, 40 + i, dev) for i in range(E)]
b = marsb200.stack_episodes(eps)
n, g = shape.N, shape.g
fs = ops.normalize_rows(b["feat_s"].reshape(E, n, shape.C)); fq = ops.normalize_rows(b["feat_q"])
row_fg = ops.pool_mask(b["support_mask"], g).reshape(E, n)
cost = ops.sim_contract(fs, fq, n, n, shape.C, want_sim=False, want_cost=True)["cost"]
bits = ops.pack_masks(b["masks"]); pooled, area, cnt = ops.pool_packed(bits, shape.H, shape.W, g)
T = row_fg.sum(1).cpu().tolist(); print("T per episode", T, "M_p mean/max", float(cnt.float().mean()), int(cnt.max()))
out = ops.emd_scores(cost, row_fg, pooled, pooled_count=cnt)
ts = []
for _ in range(3):
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = ops.emd_scores(cost, row_fg, pooled, t_cap=max(T), m_cap=int(cnt.max())); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
ms = statistics.median(ts)
print(f"E={E}: {ms:.1f} ms for {E * shape.P} LPs -> {ms / E:.1f} ms/episode, {E * shape.P / ms * 1e3:.0f} LP/s")
# host reference for a few LPs (HiGHS exact LP; POT is not installed)
from oracle import mars_oracle as orc
sup = row_fg[0].cpu().bool(); cm = cost[0].cpu()
pm = ((pooled[0][:, :, None] >> torch.arange(32, device=dev, dtype=torch.int32)) & 1).reshape(shape.P, -1)[:, :n].bool().cpu()
idx = [0, 1, 2]
t0 = time.perf_counter(); ref = [orc.emd_score(sup, pm[i], cm) for i in idx]; t1 = time.perf_counter()
print("host exact LP:", (t1 - t0) / len(idx) * 1e3, "ms per LP; max |diff|", max(abs(out[0, i].item() - ref[k]) for k, i in enumerate(idx)))
# per-LP latency by size: the launch is bounded by its largest problem
order = torch.argsort(cnt[0], descending=True)
for rank in (0, 64, 128, 255):
    p = int(order[rank])
    one = pooled[0:1, p:p + 1].contiguous()
    ops.emd_scores(cost[0:1], row_fg[0:1], one)
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.emd_scores(cost[0:1], row_fg[0:1], one, t_cap=T[0]); c.record(); torch.cuda.synchronize()
    print(f"single LP T={T[0]} M={int(cnt[0, p])}: {a.elapsed_time(c):.2f} ms")
