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
        """Box mask the kernel builds for `scoremap` (uint8 [g,g]); the reference returns the box list instead."""
        prior = torch.as_tensor(scoremap, dtype=torch.float32, device=self.device)
        g = prior.shape[-1]
        eye = torch.eye(g * g, device=self.device)[None]
        _, box = ops.pir_refine(prior.reshape(1, -1), eye, g, self.threshold, want_box=True)
        return box.reshape(g, g)
