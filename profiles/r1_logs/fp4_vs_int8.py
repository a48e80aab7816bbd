import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, marsb200
from marsb200 import ops, _lib
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
for (e, p, h, w, dens) in ((1, 24, 160, 160, 0.3), (2, 77, 140, 200, 0.5), (1, 128, 518, 518, 0.3), (2, 200, 518, 518, 0.9), (1, 256, 1024, 1024, 1.0), (2, 256, 1024, 1024, 0.4)):
    m = (torch.rand(e, p, h, w, device=dev, generator=g) < dens).to(torch.uint8)
    bits = ops.pack_masks(m)
    ref = ops.pairwise_inter(bits, backend=_lib.PAIR_MMA)
    got = ops.pairwise_inter(bits, backend=_lib.PAIR_FP4)
    torch.cuda.synchronize()
    ok = torch.equal(ref, got)
    print(f"E={e} P={p} {h}x{w} density {dens}: exact={ok} max|diff|={(ref.long() - got.long()).abs().max().item()} max count {ref.max().item()}", flush=True)
E = 16
shape = marsb200.CONFIGS["c2"]
bits = ops.pack_masks(torch.stack([marsb200.make_episode(shape, i, dev)["masks"] for i in range(E)]))
inter = torch.empty(E, shape.P, shape.P, dtype=torch.int32, device=dev)
ref = ops.pairwise_inter(bits, backend=_lib.PAIR_MMA)
got = ops.pairwise_inter(bits, backend=_lib.PAIR_FP4)
print("c2 batch exact:", torch.equal(ref, got))
for name, be in (("i8", _lib.PAIR_MMA), ("fp4", _lib.PAIR_FP4)):
    for _ in range(3): ops.pairwise_inter(bits, backend=be, out=inter)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ops.pairwise_inter(bits, backend=be, out=inter)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 20
    print(f"{name}: {t*1e3:.0f} us for {E} episodes = {E * shape.P * (shape.P + 1) / 2 / t / 1e6:.2f} G unordered pairs/s")
