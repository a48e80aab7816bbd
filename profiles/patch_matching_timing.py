"""Matcher.patch_level_matching on the device at the 5-shot shape (5 x 1369 support patches, 1369 query patches) with T >= N masked
support patches: forward (T x 1369) and reverse (1369 x 6845) assignment back to back against side by side on two streams."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200

dev = torch.device("cuda:0")
g, ns, c = 37, 5, 256
n = g * g
gen = torch.Generator().manual_seed(3)
protos = torch.randn(8, c, generator=gen)
fs = torch.nn.functional.normalize(0.6 * protos[torch.randint(0, 8, (ns * n,), generator=gen)] + 0.8 * torch.randn(ns * n, c, generator=gen), dim=1).to(dev)
fq = torch.nn.functional.normalize(0.6 * protos[torch.randint(0, 8, (n,), generator=gen)] + 0.8 * torch.randn(n, c, generator=gen), dim=1).to(dev)
for t_target in (1374, 2000, 1000):
    pool = torch.zeros(ns * n)
    pool[torch.randperm(ns * n, generator=gen)[:t_target]] = 1
    pool = pool.to(dev)
    ref = None
    for cr in (False, True):
        pm = marsb200.PatchMatcher(g, 14, (518, 518), dev, concurrent_reverse=cr)
        pm.match(fs, fq, pool)
        times = []
        for rep in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = pm.match(fs, fq, pool)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        sig = sorted(map(tuple, res["points"].cpu().tolist()))
        ref = ref or sig
        print(f"T = {t_target}, N = {n}: concurrent_reverse={cr}: {min(times):.2f} ms per match (host clock, sync both sides; "
              f"five runs {['%.1f' % t for t in times]}), same points {sig == ref}", flush=True)
