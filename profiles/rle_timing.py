"""Timing of the RLE ingest (rle_fill + bit_transpose) at c2, 16 episodes (4096 masks of 1024 x 1024)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200
from marsb200 import ops
dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
E = 16
gen = torch.Generator(device=dev).manual_seed(3)
from marsb200.synthetic import random_masks
masks = torch.stack([random_masks(shape.P, shape.H, shape.W, gen, dev, dtype=torch.uint8) for _ in range(E)])
bits_ref = ops.pack_masks(masks)
counts, offsets = marsb200.masks_to_rle(masks.reshape(-1, shape.H, shape.W))
counts, offsets = counts.to(dev), offsets.to(dev)
ws = torch.empty(int(ops.lib.marsb200_rle_workspace_bytes(E * shape.P, shape.H, shape.W)), device=dev, dtype=torch.uint8)
out = torch.empty_like(bits_ref).view(E * shape.P, -1)
for _ in range(3):
    ops.rle_decode(counts, offsets, shape.H, shape.W, out=out, workspace=ws)
torch.cuda.synchronize()
assert torch.equal(out.view_as(bits_ref), bits_ref)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.rle_decode(counts, offsets, shape.H, shape.W, out=out, workspace=ws, check_status=False)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"rle_decode: {ms:.3f} ms per {E * shape.P} masks ({counts.numel() * 4 / 1e6:.1f} MB of counts -> {out.numel() * 4 / 1e6:.0f} MB of bits), "
      f"{out.numel() * 4 * 3 / ms / 1e6:.0f} GB/s of packed traffic (memset + fill + transpose read/write)")
