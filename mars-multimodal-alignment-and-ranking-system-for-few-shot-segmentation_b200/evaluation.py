"""Drop-in for mars/utils/evaluation.py: prediction-vs-ground-truth areas on the device."""
import torch

from . import ops


class Evaluator:
    r"""Computes intersection and union between prediction and ground-truth (Evaluator.classify_prediction)."""

    @classmethod
    def initialize(cls):
        cls.ignore_index = 255

    @classmethod
    def classify_prediction(cls, pred_mask: torch.Tensor, batch: dict):
        """pred_mask [B,H,W]; batch['query_mask'] [B,H,W]; optional batch['query_ignore_idx'].

        Returns (area_inter [2,B], area_union [2,B]) float tensors like the reference (mars/utils/evaluation.py:12-38);
        unlike the reference the inputs are not modified in place.
        """
        gt = batch.get("query_mask")
        ignore = batch.get("query_ignore_idx")
        if ignore is not None:
            assert torch.logical_and(ignore, gt).sum() == 0
        dev = pred_mask.device if pred_mask.is_cuda else torch.device("cuda")
        out = ops.eval_areas(pred_mask.to(dev).float(), gt.to(dev).float(),
                             None if ignore is None else ignore.to(dev).float())
        out = out.float()
        return out[:, :2].t().contiguous(), out[:, 2:].t().contiguous()
