"""Synthetic episodes of the shapes BASELINE.json names.

Backbone outputs are replaced by seeded tensors of the right shape and
statistics: prototype-mixed DINOv2 patch features, softmax attention means,
normalised AlphaCLIP features, and proposal masks that are unions of random
rectangles / ellipses with exact and 1-pixel-shifted duplicates (so the NMS and
the tie rule are exercised).  The generator is plain torch and runs on any
device; the same code builds the CPU copies the oracle consumes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class EpisodeShape:
    ns: int = 1          # shots
    g: int = 37          # DINOv2 patch grid side (518 / 14)
    C: int = 1024        # DINOv2 ViT-L/14 width
    P: int = 256         # proposals
    H: int = 1024        # proposal resolution
    W: int = 1024
    gt: int = 33         # CLIP ViT-B/16 grid for the vta map
    D: int = 768         # AlphaCLIP embedding width

    @property
    def N(self):
        return self.g * self.g


# the configurations of BASELINE.json (c5 = many c2 episodes)
CONFIGS = {
    "c1": EpisodeShape(ns=1, g=37, C=1024, P=128, H=518, W=518),
    "c2": EpisodeShape(ns=1, g=37, C=1024, P=256, H=1024, W=1024),
    "c3": EpisodeShape(ns=5, g=37, C=1024, P=512, H=518, W=518),
    "c4": EpisodeShape(ns=1, g=37, C=1024, P=1000, H=1024, W=1024),
}


def _shapes(n, h, w, gen, device, min_frac, max_frac):
    """n random boxes as (y0, x0, bh, bw, is_ellipse) tensors on `device`."""
    frac = torch.empty(n, device=device).uniform_(min_frac, max_frac, generator=gen)
    aspect = torch.empty(n, device=device).uniform_(0.5, 2.0, generator=gen)
    bh = (frac * h * w * aspect).sqrt().round().clamp(1, h)
    bw = (frac * h * w / aspect).sqrt().round().clamp(1, w)
    y0 = (torch.rand(n, device=device, generator=gen) * (h - bh + 1)).floor()
    x0 = (torch.rand(n, device=device, generator=gen) * (w - bw + 1)).floor()
    ell = torch.rand(n, device=device, generator=gen) < 0.5
    return y0, x0, bh, bw, ell


def random_masks(n, h, w, gen, device, min_frac=0.005, max_frac=0.4, dup_frac=0.1, chunk=32, dtype=torch.float32):
    """[n, h, w] 0/1 masks: unions of 1-3 shapes; ~dup_frac exact and ~dup_frac shifted duplicates; never empty."""
    out = torch.empty((n, h, w), device=device, dtype=dtype)
    yy = torch.arange(h, device=device, dtype=torch.float32)[None, :, None]
    xx = torch.arange(w, device=device, dtype=torch.float32)[None, None, :]
    n_shapes = torch.randint(1, 4, (n,), device=device, generator=gen)
    params = [_shapes(n, h, w, gen, device, min_frac / 2, max_frac / 2) for _ in range(3)]
    for s in range(0, n, chunk):
        e = min(s + chunk, n)
        m = torch.zeros((e - s, h, w), device=device, dtype=torch.bool)
        for k in range(3):
            y0, x0, bh, bw, ell = [p[s:e, None, None] for p in params[k]]
            rect = (yy >= y0) & (yy < y0 + bh) & (xx >= x0) & (xx < x0 + bw)
            cy, cx = y0 + bh / 2, x0 + bw / 2
            inside = ((yy - cy) / (bh / 2 + 1e-6)) ** 2 + ((xx - cx) / (bw / 2 + 1e-6)) ** 2 <= 1.0
            shape = torch.where(ell, inside, rect)
            m |= shape & (n_shapes[s:e, None, None] > k)
        empty = ~m.flatten(1).any(1)
        m[empty, h // 2, w // 2] = True
        out[s:e] = m.to(dtype)
    if dup_frac > 0 and n >= 8:
        n_dup = max(1, int(n * dup_frac))
        src = torch.randint(0, n // 2, (2 * n_dup,), device=device, generator=gen)
        dst = torch.arange(n - 2 * n_dup, n, device=device)
        out[dst[:n_dup]] = out[src[:n_dup]]
        out[dst[n_dup:]] = torch.roll(out[src[n_dup:]], 1, dims=2)
    return out


def make_episode(shape: EpisodeShape, seed: int, device="cpu", mask_dtype=torch.float32, with_emd=True) -> dict:
    """One synthetic episode (dict of tensors on `device`), deterministic for a (seed, device type)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + seed)
    s = shape
    n, nt = s.N, s.gt * s.gt
    protos = torch.randn(8, s.C, device=device, generator=gen)

    def feats(rows):
        k = torch.randint(0, 8, (rows,), device=device, generator=gen)
        return 0.6 * protos[k] + 0.8 * torch.randn(rows, s.C, device=device, generator=gen)

    ep = dict(
        feat_s=feats(s.ns * n).reshape(s.ns, n, s.C),
        feat_q=feats(n),
        support_mask=random_masks(s.ns, s.H, s.W, gen, device, 0.05, 0.3, dup_frac=0.0, dtype=torch.float32),
        attn_vva=torch.softmax(2.0 * torch.randn(n, n, device=device, generator=gen), dim=-1),
        vta_raw=torch.rand(s.gt, s.gt, device=device, generator=gen),
        attn_vta=torch.softmax(2.0 * torch.randn(nt, nt, device=device, generator=gen), dim=-1),
        masks=random_masks(s.P, s.H, s.W, gen, device, dtype=mask_dtype),
        clip_img=torch.nn.functional.normalize(torch.randn(s.P, s.D, device=device, generator=gen), dim=1),
        clip_txt=torch.nn.functional.normalize(torch.randn(s.D, device=device, generator=gen), dim=0),
    )
    if with_emd:
        ep["emd"] = torch.rand(s.P, device=device, generator=gen, dtype=torch.float64)
    return ep


def stack_episodes(episodes) -> dict:
    """List of episode dicts -> one batch dict with a leading E dimension."""
    return {k: torch.stack([e[k] for e in episodes]) for k in episodes[0]}


def to_device(batch: dict, device, non_blocking=False) -> dict:
    return {k: v.to(device, non_blocking=non_blocking) for k, v in batch.items()}


def masks_to_rle(masks: torch.Tensor):
    """Uncompressed COCO RLE of [n, H, W] masks (column-major runs, first run = zeros) as (counts int32 [total],
    offsets int64 [n + 1]) CPU tensors: SAM's wire format (segment_anything/utils/amg.py:107-135).  Host-side
    input preparation for tests and benchmarks, not part of the ranking path."""
    import numpy as np

    m = (masks.detach().cpu().numpy() > 0)
    counts, offsets = [], [0]
    for one in m:
        flat = one.T.reshape(-1)
        change = np.nonzero(flat[1:] ^ flat[:-1])[0] + 1
        idx = np.concatenate([[0], change, [flat.size]])
        c = (idx[1:] - idx[:-1]).astype(np.int32)
        if flat[0]:
            c = np.concatenate([np.zeros(1, np.int32), c])
        counts.append(c)
        offsets.append(offsets[-1] + len(c))
    return torch.from_numpy(np.concatenate(counts)), torch.tensor(offsets, dtype=torch.int64)
