"""Drop-in for mars/components/FilteringMergingModule.py (same class, ctor and method signatures).

AlphaCLIP stays a PyTorch input producer.  Everything else runs in
libmarsb200: proposals are bit-packed once (128-bit loads), pooled to the
patch grid, scored against the vva / vta maps, fused with the EMD and
AlphaCLIP scores, ranked (stable, like Python's sorted) and OR-merged.
The Python loop over proposals of the reference (FilteringMergingModule.py:103-123)
and its 3*P device->host copies disappear.

EMD (`ot.emd2`, an exact transport LP per proposal, FilteringMergingModule.py:142-169)
is solved exactly on the device for all proposals at once (`ops.emd_scores`,
primal-dual method on integer flows).  `emd_fn=` (a host solver taking the
cost sub-matrix, e.g. a POT wrapper) or precomputed `emd_scores=` override it;
`_compute_emd` keeps the reference's per-proposal host signature for callers that
use it directly.

AlphaCLIP features in float16 (what the reference's own producer code returns on a
GPU) select the float16 score sequence of the reference: float16 dot products,
float16 min-max (with NumPy's weak-scalar rule for the 1e-7), float16 first addition,
float64 afterwards; float32 features select the float32 sequence.
"""
from typing import Callable, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import ops


class FilteringMergingModule:
    def __init__(self, alpha_clip_model: nn.Module, img_transforms, mask_transforms, alpha: float,
                 static_threshold: float, dynamic_threshold: float, device,
                 nms_iou_threshold: Optional[float] = None, emd_fn: Optional[Callable] = None):
        self.alpha_clip_model = alpha_clip_model
        self.img_transforms = img_transforms
        self.mask_transforms = mask_transforms
        self.alpha = alpha
        self.static_threshold = static_threshold
        self.dynamic_threshold = dynamic_threshold
        self.device = device
        self.alpha_clip_batch_size = 128
        # extensions (off by default = reference behaviour)
        self.nms_iou_threshold = nms_iou_threshold
        self.emd_fn = emd_fn
        self.last = None  # device-side results of the last call (scores, order, flags, inter, ...)

    # ------------------------------------------------------------------ public API (reference names)
    def compute(self, query_img, mask_proposals, support_mask, cost_matrix, patch_features_spatial_dimension,
                vva, vta, text, emd_scores=None, alphaclip_feats=None) -> torch.Tensor:
        self._rank(query_img, mask_proposals, support_mask, cost_matrix, patch_features_spatial_dimension, vva, vta,
                   text, emd_scores, alphaclip_feats)
        h, w = mask_proposals.shape[-2:]
        _, merged = ops.merge_masks(self.last["bits"], self.last["flags"], h * w)
        return merged.reshape(h, w)

    def _score_proposals(self, query_img, mask_proposals, support_mask, cost_matrix,
                         patch_features_spatial_dimension, vva, vta, text, emd_scores=None, alphaclip_feats=None):
        """List of (mask, score) sorted by descending score (FilteringMergingModule.py:59-140)."""
        self._rank(query_img, mask_proposals, support_mask, cost_matrix, patch_features_spatial_dimension, vva, vta,
                   text, emd_scores, alphaclip_feats)
        order = self.last["order"][0].cpu().numpy()
        scores = self.last["scores"][0].cpu().numpy()
        return [(mask_proposals[i], scores[i]) for i in order]

    def _merge_masks(self, ranked_masks) -> torch.Tensor:
        """Threshold selection + OR merge of an already ranked list (FilteringMergingModule.py:209-221)."""
        top = ranked_masks[0][1]
        bound = self.dynamic_threshold * top if top < self.static_threshold else self.static_threshold
        chosen = torch.stack([m for m, s in ranked_masks if s >= bound]).to(self.device)
        bits = ops.pack_masks(chosen)[None]
        flags = torch.full((1, chosen.shape[0]), 3, dtype=torch.uint8, device=self.device)
        h, w = chosen.shape[-2:]
        _, merged = ops.merge_masks(bits, flags, h * w)
        return merged.reshape(h, w)

    def _compute_emd(self, support_mask, mask_proposal, cost_matrix) -> float:
        """1 - EMD between the pooled support mask and a pooled proposal (FilteringMergingModule.py:142-169)."""
        sub = cost_matrix[support_mask.flatten().bool().to(cost_matrix.device), :][
            :, mask_proposal.flatten().bool().to(cost_matrix.device)]
        sub = sub.detach().cpu().numpy()
        if self.emd_fn is not None:
            return 1 - self.emd_fn(sub)
        try:
            import ot
        except ImportError as exc:  # POT is the reference's dependency (requirements.txt:23)
            raise RuntimeError("EMD needs POT (`ot`) or an `emd_fn`; or pass emd_scores=...") from exc
        t, m = sub.shape
        return 1 - ot.emd2(a=[1.0 / t] * t, b=[1.0 / m] * m, M=sub)

    # ------------------------------------------------------------------ AlphaCLIP producer side
    def _compute_alphaclip_text_feats(self, text):
        from alpha_clip import tokenize as alpha_clip_tokenizer

        tokens = alpha_clip_tokenizer(text).to(self.device)
        with torch.no_grad():
            feats = self.alpha_clip_model.encode_text(tokens)
            feats = feats / feats.norm(dim=-1, keepdim=True)
        return feats

    def _compute_alphaclip_vis_feats(self, image, masks):
        image_in = self.img_transforms(image.permute(1, 2, 0).cpu().numpy()).unsqueeze(0).half().to(self.device)
        feats = []
        for i in range(0, masks.shape[0], self.alpha_clip_batch_size):
            chunk = masks[i:i + self.alpha_clip_batch_size]
            alpha = torch.stack([self.mask_transforms((m.cpu().numpy() * 255).astype(np.uint8)) for m in chunk])
            alpha = alpha.half().to(self.device)
            with torch.no_grad():
                f = self.alpha_clip_model.visual(image_in.repeat(alpha.shape[0], 1, 1, 1), alpha)
            feats.append(f / f.norm(dim=-1, keepdim=True))
        return torch.cat(feats, dim=0)

    # ------------------------------------------------------------------ the device path
    def _rank(self, query_img, mask_proposals, support_mask, cost_matrix, g, vva, vta, text, emd_scores,
              alphaclip_feats):
        dev = self.device
        p, h, w = mask_proposals.shape
        n = g * g
        if not mask_proposals.is_cuda and mask_proposals.dtype in (torch.float32, torch.uint8, torch.bool):
            # the reference's proposals are CPU tensors (main_MARS.py:62-69): host threads pack them and only the bits cross
            # PCIe (1/32 of the float32 bytes; ops.host_pack_masks writes the device kernel's layout bit for bit)
            bits = ops.host_pack_masks(mask_proposals).to(dev)[None]
        else:
            bits = ops.pack_masks(mask_proposals.to(dev))[None]
        pooled, area, cnt = ops.pool_packed(bits, h, w, g)
        sv, st, uc = ops.region_sums(pooled, vva.to(dev).reshape(1, n), vta.to(dev).reshape(1, n))
        if alphaclip_feats is None:
            txt = self._compute_alphaclip_text_feats(text)
            img = self._compute_alphaclip_vis_feats(query_img[0], mask_proposals)
        else:
            img, txt = alphaclip_feats
        # AlphaCLIP runs in half precision on a GPU (FilteringMergingModule.py:189,195): with float16 features the reference's
        # dot products, their min-max and the first addition of the fusion are float16 arithmetic (:97,126-136); the
        # kernels reproduce that sequence.  float32 features take the float32 sequence.
        clip_f16 = img.dtype == torch.float16 and txt.dtype == torch.float16
        if not clip_f16:
            img, txt = img.float(), txt.float()
        clip = ops.clip_scores(img.to(dev)[None], txt.to(dev).reshape(1, -1))
        if emd_scores is None:
            sup = ops.pool_mask(support_mask.to(dev).permute(1, 0, 2, 3), g).reshape(-1)
            if self.emd_fn is None:
                emd = ops.emd_scores(cost_matrix.to(dev).float()[None], sup[None], pooled, pooled_count=cnt)
            else:
                pooled_np = self._unpack_pooled(pooled[0], n)
                emd_scores = [self._compute_emd(sup, pooled_np[i], cost_matrix) for i in range(p)]
        if emd_scores is not None:
            emd = torch.as_tensor(np.asarray(emd_scores, dtype=np.float64), device=dev).reshape(1, p)
        inter = ops.pairwise_inter(bits) if self.nms_iou_threshold is not None else None
        res = ops.fuse_rank(emd, clip, cnt, sv, st, uc, inter, self.alpha, self.static_threshold,
                            self.dynamic_threshold, self.nms_iou_threshold, clip_f16=clip_f16)
        res.update(bits=bits, pooled=pooled, area=area, pooled_count=cnt, inter=inter, clip=clip)
        self.last = res

    @staticmethod
    def _unpack_pooled(pooled: torch.Tensor, n: int) -> torch.Tensor:
        """[P, npw] packed pooled bitmaps -> bool [P, n] on the same device."""
        shifts = torch.arange(32, device=pooled.device, dtype=torch.int32)
        return ((pooled[:, :, None] >> shifts) & 1).reshape(pooled.shape[0], -1)[:, :n].bool()
