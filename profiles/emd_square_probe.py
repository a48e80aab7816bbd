"""How fast does the transport-LP kernel solve a SQUARE problem (= an assignment problem) against the
   Jonker-Volgenant kernel and scipy?  T = M = n patches of a c2 episode's cost matrix."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import marsb200
from marsb200 import ops
from scipy.optimize import linear_sum_assignment

dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
b = marsb200.stack_episodes([marsb200.make_episode(shape, 40, dev)])
n, C = shape.N, shape.C
fs = ops.normalize_rows(b["feat_s"].reshape(1, n, C)); fq = ops.normalize_rows(b["feat_q"])
cost = ops.sim_contract(fs, fq, n, n, C, want_sim=False, want_cost=True)["cost"]  # [1, n, n]
npw = (n + 31) // 32
for size in (256, 512, 1024, 1369):
    row_fg = torch.zeros(1, n, dtype=torch.uint8, device=dev); row_fg[0, :size] = 1
    pooled = torch.zeros(1, 1, npw, dtype=torch.int32, device=dev)
    bits = np.zeros(npw * 32, dtype=np.uint8); bits[:size] = 1
    pooled[0, 0] = torch.from_numpy(np.packbits(bits, bitorder="little").view(np.int32)).to(dev)
    ops.emd_scores(cost, row_fg, pooled, t_cap=size, m_cap=size)
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = ops.emd_scores(cost, row_fg, pooled, t_cap=size, m_cap=size); c.record(); torch.cuda.synchronize()
    sub = cost[0, :size, :size].double().cpu().numpy()
    t0 = time.perf_counter(); r, cc = linear_sum_assignment(sub); t_sp = time.perf_counter() - t0
    ref = 1.0 - sub[r, cc].sum() / size
    rows = torch.arange(size, device=dev, dtype=torch.int32)[None]
    try:
        ops.lsap(cost[:, :size, :size].contiguous(), maximize=False)
        a2, c2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2.record(); ops.lsap(cost[:, :size, :size].contiguous(), maximize=False); c2.record(); torch.cuda.synchronize()
        t_jv = a2.elapsed_time(c2)
    except Exception as ex:
        t_jv = float("nan"); print("lsap:", ex)
    print(f"n={size}: transport kernel {a.elapsed_time(c):.2f} ms (|diff| {abs(float(out[0, 0]) - ref):.1e}), JV kernel {t_jv:.2f} ms, scipy {t_sp * 1e3:.1f} ms", flush=True)
