"""Episode-level comparison of a device result with ``mars_oracle.run_episode``.

TEST INFRASTRUCTURE ONLY (same rules as ``mars_oracle.py``): used by ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline leg of ``bench.py`` to check what the
CUDA path produced.  Bars are the north-star's: integer work bit-exact, float scores
within ``rtol`` relative, ranking identical except for swaps whose reference scores tie
within that tolerance.

When such a tolerated swap occurs the device's NMS / selection / merge ran on ITS order, so
the expectation is recomputed by the oracle on the device's order (greedy NMS depends on
the visiting order) instead of being skipped.
"""
from __future__ import annotations

import numpy as np
import torch

from . import mars_oracle as orc

RTOL = 1e-4


def oracle_config(cfg, g: int) -> dict:
    """RankingConfig (or any object with the same attributes) -> the dict ``run_episode`` takes."""
    return dict(g=g, vva_box_threshold=cfg.vva_box_threshold, vta_box_threshold=cfg.vta_box_threshold,
                alpha=cfg.alpha, static_threshold=cfg.static_threshold, dynamic_threshold=cfg.dynamic_threshold,
                nms_iou_threshold=cfg.nms_iou_threshold)


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def unpack_bits(words, n: int) -> np.ndarray:
    """[rows, words] uint32/int32 bit rows -> bool [rows, n]."""
    w = _np(words)
    w = w.view(np.uint32).reshape(w.shape[0], -1)
    bits = np.unpackbits(w.view(np.uint8).reshape(w.shape[0], -1), axis=-1, bitorder="little")
    return bits[:, :n].astype(bool)


def compare_episode(ref: dict, got: dict, masks: torch.Tensor, cfg: dict, rtol: float = RTOL,
                    check_maps: bool = True) -> dict:
    """``ref`` = ``run_episode`` output; ``got`` = one episode's device outputs (tensors / arrays: order, scores,
    flags, and optionally row_fg, prior, vva, vta, pooled, clip, inter, area, merged [H*W] or merged_bits).
    Returns {"ok": bool, "failures": [...], "tie_swaps": n}.  Never raises: callers decide."""
    fails = []
    P = len(ref["order"])

    def close(name, a, b, atol):
        a, b = _np(a).reshape(-1).astype(np.float64), _np(b).reshape(-1).astype(np.float64)
        bad = np.abs(a - b) > atol + rtol * np.abs(b)
        if bad.any():
            i = int(np.argmax(np.abs(a - b) - rtol * np.abs(b)))
            fails.append(f"{name}: {int(bad.sum())} of {a.size} outside rtol {rtol} (worst {a[i]!r} vs {b[i]!r})")

    def equal(name, a, b):
        a, b = _np(a), _np(b)
        if a.shape != b.shape or not np.array_equal(a, b):
            fails.append(f"{name}: not bit-exact ({int((a.reshape(-1) != b.reshape(-1)).sum()) if a.size == b.size else 'shape'} differ)")

    if check_maps:
        if got.get("row_fg") is not None:
            equal("support bits", _np(got["row_fg"]).reshape(-1) > 0, _np(ref["support_bits"]).reshape(-1) > 0)
        for k, atol in (("prior", 2e-5), ("vva", 5e-5), ("vta", 5e-5)):
            if got.get(k) is not None:
                close(k, got[k], ref[k], atol)
        if got.get("pooled") is not None:
            n = _np(ref["pooled"]).reshape(P, -1).shape[1]
            equal("pooled bitmaps", unpack_bits(got["pooled"], n), _np(ref["pooled"]).reshape(P, -1) > 0)
        if got.get("clip") is not None:
            close("clip", got["clip"], ref["clip"], 1e-6)
    nms = cfg.get("nms_iou_threshold")
    if nms is not None and got.get("inter") is not None:
        equal("intersections", got["inter"], ref["inter"])
        if got.get("area") is not None:
            equal("areas", got["area"], ref["area"])
    close("scores", got["scores"], ref["scores"], 1e-6)
    order_g, order_r = _np(got["order"]).astype(np.int64), np.asarray(ref["order"]).astype(np.int64)
    swaps = np.nonzero(order_g != order_r)[0]
    for r in swaps:
        a, b = ref["scores"][order_g[r]], ref["scores"][order_r[r]]
        if abs(a - b) > rtol * max(abs(a), abs(b)) + 1e-9:
            fails.append(f"order: rank {int(r)} holds {int(order_g[r])} instead of {int(order_r[r])} outside the tie tolerance")
            break
    if sorted(order_g.tolist()) != list(range(P)):
        fails.append("order: not a permutation")
        return {"ok": False, "failures": fails, "tie_swaps": int(len(swaps))}

    # keep-set / selection / merged mask on the device's own order (equal to the oracle's unless a tie swapped)
    flags = _np(got["flags"]).astype(np.uint8)
    scores_r = np.asarray(ref["scores"], dtype=np.float64)
    keep = np.ones(P, dtype=bool)
    if nms is not None:
        keep = orc.mask_nms(order_g, ref["inter"], ref["area"], nms)
        equal("NMS keep-set", (flags & 1).astype(bool), keep)
    sel_ranked = orc.merge_select(scores_r[order_g], cfg["static_threshold"], cfg["dynamic_threshold"]) & keep[order_g]
    sel = np.zeros(P, dtype=bool)
    sel[order_g[sel_ranked]] = True
    sel_g = (flags & 2) > 0
    top = scores_r[order_g[0]]
    bound = cfg["dynamic_threshold"] * top if top < cfg["static_threshold"] else cfg["static_threshold"]
    diff = np.nonzero(sel_g != sel)[0]
    # a threshold flip is only tolerated for a score within the tolerance of the bound
    bad = [int(i) for i in diff if abs(scores_r[i] - bound) > rtol * abs(bound) + 1e-9]
    if bad:
        fails.append(f"selection: proposals {bad[:8]} differ and are not within tolerance of the bound {bound!r}")
    merged_g = None
    if got.get("merged") is not None:
        merged_g = _np(got["merged"]).reshape(-1) > 0
    elif got.get("merged_bits") is not None:
        merged_g = unpack_bits(_np(got["merged_bits"]).reshape(1, -1), masks[0].numel())[0]
    if merged_g is not None and not bad:
        want = _np(orc.merge_masks(masks, np.nonzero(sel_g)[0])).reshape(-1) > 0  # on the device's (tolerated) selection
        equal("merged mask", merged_g, want)
    return {"ok": not fails, "failures": fails, "tie_swaps": int(len(swaps))}
