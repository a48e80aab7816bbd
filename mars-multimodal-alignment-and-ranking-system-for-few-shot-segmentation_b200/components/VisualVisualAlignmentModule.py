"""Drop-in for mars/components/VisualVisualAlignmentModule.py (same class, ctor, compute, clear, attributes).

The DINOv2 backbone stays a PyTorch input producer; everything after
`forward_features` / `get_last_self_attention` runs in libmarsb200: row
normalisation fused with the TF32 split (VisualVisualAlignmentModule.py:124-125),
S = Fs Fq^T with the fg/bg column max / mean fused into the contraction epilogue
(:69, :76-102 - the reference recomputes the products twice), min-max, PIR.
`similarity_matrix` / `cost_matrix` keep their names and shapes but stay on the
device (the reference moves them to the CPU, :69).
"""
import torch
import torch.nn as nn

from .. import ops
from .PriorInformationRefinementModule import PriorInformationRefinementModule


class VisualVisualAlignmentModule:
    def __init__(self, model: nn.Module, model_transforms, model_patch_size: int,
                 model_embedding_spatial_dimensions: int, model_num_regs: int, vva_refinement_box_threshold: float,
                 last_n_attention_maps_for_refinement: int, device, matrices_on_cpu: bool = False):
        self.model = model
        self.model_transforms = model_transforms
        self.model_patch_size = model_patch_size
        self.model_embedding_spatial_dimensions = model_embedding_spatial_dimensions
        self.model_num_regs = model_num_regs
        self.device = device
        self.pir = PriorInformationRefinementModule(
            box_threshold=vva_refinement_box_threshold,
            last_n_attention_maps_for_refinement=last_n_attention_maps_for_refinement,
            device=device, num_regs=model_num_regs)
        # the reference moves `similarity_matrix` / `cost_matrix` to the CPU (VisualVisualAlignmentModule.py:69-70) because its
        # EMD runs on the host; here the consumer (FilteringMergingModule) is on the device, so they stay there by default.
        # `matrices_on_cpu=True` restores the reference's placement for callers that index them with CPU tensors.
        self.matrices_on_cpu = matrices_on_cpu
        self.similarity_matrix = None
        self.cost_matrix = None

    def compute(self, support_imgs: torch.Tensor, support_masks: torch.Tensor, query_img: torch.Tensor) -> torch.Tensor:
        g = self.model_embedding_spatial_dimensions
        n = g * g
        support_imgs = support_imgs.to(self.device)
        support_masks = support_masks.to(self.device)
        query_img = query_img.to(self.device)

        fs_raw = self._backbone_patch_tokens(support_imgs[0])
        fq_raw = self._backbone_patch_tokens(query_img)
        attn_maps = list(self.model.get_last_self_attention(
            self.model_transforms(query_img[0]).unsqueeze(0).to(self.device)))
        m, c = fs_raw.shape
        fs = ops.normalize_rows(fs_raw)
        fq = ops.normalize_rows(fq_raw)
        row_fg = ops.pool_mask(support_masks.permute(1, 0, 2, 3), g).reshape(1, m)
        if not bool(row_fg.any()):
            # the reference fails here too: max over an empty foreground (VisualVisualAlignmentModule.py:82)
            raise RuntimeError("empty pooled support mask: no foreground support patch")
        res = ops.sim_contract(fs, fq, m, n, c, want_sim=True, want_cost=True, row_fg=row_fg)
        self.similarity_matrix = res["sim"][0].cpu() if self.matrices_on_cpu else res["sim"][0]
        self.cost_matrix = res["cost"][0].cpu() if self.matrices_on_cpu else res["cost"][0]
        if not bool((row_fg == 0).any()):
            print("[VVA] - No background VVA computed, only foreground VVA.")
        prior = ops.vva_finalize(res["colstats"], row_fg, m, n).reshape(g, g)
        refined = self.pir.compute(prior=prior, attn_maps=attn_maps)
        lo, hi = refined.min(), refined.max()
        return (refined - lo) / (1e-7 + hi - lo)

    def _backbone_patch_tokens(self, imgs) -> torch.Tensor:
        imgs = torch.cat([self.model_transforms(i).unsqueeze(0).to(self.device) for i in imgs], dim=0)
        with torch.no_grad():
            feats = self.model.forward_features(imgs)["x_prenorm"][:, 1 + self.model_num_regs:]
        return feats.reshape(-1, self.model.embed_dim).float()

    def _extract_patch_features(self, imgs) -> torch.Tensor:
        """L2-normalised patch features, as the reference method returns them (:113-127)."""
        padded, _ = ops.normalize_rows(self._backbone_patch_tokens(imgs))
        raw_rows = padded.shape[1]
        feats = padded[0]
        rows = sum(1 for _ in imgs) * self.model_embedding_spatial_dimensions ** 2
        return feats[:min(rows, raw_rows), :self.model.embed_dim]

    def clear(self):
        self.similarity_matrix = None
        self.cost_matrix = None
