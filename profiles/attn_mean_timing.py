"""HBM rate of the attention-mean kernel at DINOv2 ViT-L/14-reg4 shapes (24 layers x 16 heads x 1374^2)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, marsb200
from marsb200 import ops
dev = torch.device("cuda:0")
for dtype in (torch.float32, torch.float16):
    L, h, T, skip = 24, 16, 1374, 5
    maps = [torch.rand(1, h, T, T, device=dev, dtype=dtype) for _ in range(L)]
    out = ops.attn_mean(maps, skip)
    ref = torch.stack([m[0, :, skip:, skip:] for m in maps]).mean(dim=(0, 1)).float()
    err = (out - ref).abs().max().item()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.attn_mean(maps, skip, out=out); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    nbytes = L * h * (T - skip) * (T - skip) * maps[0].element_size()
    print(f"{dtype}: {ms:.3f} ms, {nbytes / ms / 1e6:.0f} GB/s of useful bytes, max |err| {err:.2e}")
    del maps
