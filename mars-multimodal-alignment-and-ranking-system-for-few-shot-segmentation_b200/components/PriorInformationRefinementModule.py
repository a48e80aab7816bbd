"""Drop-in for mars/components/PriorInformationRefinementModule.py (same class, ctor and compute signature).

The arithmetic after the backbone has produced the attention maps runs in
libmarsb200: attention mean (PriorInformationRefinementModule.py:31-45), box
mask from 8-connected components (:56-63, :91-122; replaces the OpenCV round
trip through host memory), Sinkhorn-1 normalisation, R = max(D, D D^T) on the
tcgen05 contraction and the refinement (:70-87).
"""
import torch

from .. import ops


class PriorInformationRefinementModule:
    def __init__(self, box_threshold: float, last_n_attention_maps_for_refinement: int, device, num_regs: int = 0):
        self.threshold = box_threshold
        self.last_n_attention_maps_for_refinement = last_n_attention_maps_for_refinement
        self.device = device
        self.num_regs = num_regs

    def compute(self, prior: torch.Tensor, attn_maps: list) -> torch.Tensor:
        shape = prior.shape
        g = shape[-1]
        if shape[-2] != g:
            raise ValueError("PIR expects a square prior")
        maps = list(attn_maps)[-self.last_n_attention_maps_for_refinement:]
        maps = [a.to(self.device) for a in maps]
        attn = ops.attn_mean(maps, skip=1 + self.num_regs)
        if attn.shape[0] != g * g:
            raise ValueError(f"attention has {attn.shape[0]} patch tokens, prior has {g * g}")
        out = ops.pir_refine(prior.to(self.device).float().reshape(1, g * g), attn[None], g, self.threshold)
        return out.reshape(shape)

    def _scoremap2bbox(self, scoremap, multi_contour_eval: bool = False):
        """`(boxes ndarray [k, 4] of (x0, y0, x1, y1), k)` like the reference (:91-122): one box per 8-connected
        component of `uint8(scoremap * 255) > int(threshold * max)`, with the reference's clip `x1 = min(x + w, W - 1)`;
        `[[0, 0, 0, 0]], 1` when nothing passes the threshold.  Computed by the same kernel that builds the box mask
        inside `compute`.  Two documented differences to OpenCV's contour list: hole contours (nested in their
        component's box, so they never change the mask) are not listed, and with `multi_contour_eval=False` the
        component with the largest BOX area stands in for `cv2.contourArea` (the pipeline always passes True, :53-56)."""
        import numpy as np

        prior = torch.as_tensor(np.asarray(scoremap, dtype=np.float32), device=self.device)
        g = prior.shape[-1]
        boxes, count = ops.scoremap_boxes(prior.reshape(1, -1), g, self.threshold)
        k = int(count[0])
        if k == 0:
            return np.asarray([[0, 0, 0, 0]]), 1
        out = boxes[0, :k].cpu().numpy().astype(np.int64)
        if not multi_contour_eval:
            area = (out[:, 2] - out[:, 0] + 1) * (out[:, 3] - out[:, 1] + 1)
            return out[int(np.argmax(area))][None], 1
        return out, k
