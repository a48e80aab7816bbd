"""Timing of the device assignment solver at the Matcher shapes: 1-shot forward (325 x 1369) and the near-square 5-shot
forward problem (1374 x 1369), against scipy on the host (objective must agree)."""
import os, sys, time
import numpy as np
import torch
from scipy.optimize import linear_sum_assignment
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200
from marsb200 import ops
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(7)
for r, c in ((325, 1369), (1374, 1369), (1369, 6845)):
    protos = torch.randn(8, 256, generator=gen)
    fa = torch.nn.functional.normalize(0.6 * protos[torch.randint(0, 8, (r,), generator=gen)] + 0.8 * torch.randn(r, 256, generator=gen), dim=1)
    fb = torch.nn.functional.normalize(0.6 * protos[torch.randint(0, 8, (c,), generator=gen)] + 0.8 * torch.randn(c, 256, generator=gen), dim=1)
    S = (fa @ fb.T).contiguous()
    Sd = S.to(dev)
    for _ in range(2):
        r2c, obj = ops.lsap(Sd, None, None, maximize=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r2c, obj = ops.lsap(Sd, None, None, maximize=True)
    b.record(); torch.cuda.synchronize()
    t0 = time.perf_counter(); rr, cc = linear_sum_assignment(S.numpy(), maximize=True); t1 = time.perf_counter()
    want = float(S.numpy()[rr, cc].astype(np.float64).sum())
    print(f"{r} x {c}: device {a.elapsed_time(b):.2f} ms, scipy {1e3 * (t1 - t0):.2f} ms, objective diff {abs(float(obj[0]) - want):.2e}")
