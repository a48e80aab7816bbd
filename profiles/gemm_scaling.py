"""Contraction time vs K (fixed 1369 x 1369 output, 8 episodes): separates the per-k-block cost from the per-tile cost."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marsb200
from marsb200 import ops

dev = torch.device("cuda:0")
E, M, N = 8, 1369, 1369
for K in (256, 1024, 2048, 4096):
    a = torch.randn(E, M, K, device=dev)
    b = torch.randn(E, N, K, device=dev)
    fa, fb = ops.normalize_rows(a), ops.normalize_rows(b)
    row_fg = (torch.rand(E, M, device=dev) < 0.3).to(torch.uint8)
    out = {}
    for mode, kw in (("colstats", dict(want_sim=False, row_fg=row_fg)), ("store S", dict(want_sim=True))):
        res = ops.sim_contract(fa, fb, M, N, K, out=out, **kw)
        out = {k: v for k, v in res.items() if v is not None}
        ts = []
        for _ in range(10):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); ops.sim_contract(fa, fb, M, N, K, out=out, **kw); t.record()
            torch.cuda.synchronize(); ts.append(s.elapsed_time(t))
        ms = statistics.median(ts)
        flops = 3 * 2 * E * 1408 * 1408 * K
        print(f"K={K:5d} {mode:9s} {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s(tf32 issued)  per k-block per tile-wave: {ms*1e3/(K/32)/ (968/148):6.2f} us")
