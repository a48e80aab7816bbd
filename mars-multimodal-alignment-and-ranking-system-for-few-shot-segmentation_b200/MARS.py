"""Drop-in for mars/MARS.py: same class, `predict` / `clear` signatures and timing attributes.

The text retriever (ViP-LLaVA) and the Grad-CAM producer of the VTA map are
PyTorch components the caller supplies, exactly as in the reference; the
visual-visual alignment and filtering/merging components are the B200-native
ones of this package.
"""
import time
from typing import Optional

import torch

from . import ops
from .components import FilteringMergingModule, VisualVisualAlignmentModule


class MARS:
    def __init__(self, text_retriever_component, visual_text_alignment_component,
                 visual_visual_alignment_component: VisualVisualAlignmentModule,
                 filtering_merging_component: FilteringMergingModule, mask_generator=None):
        self.text_retriever_component = text_retriever_component
        self.visual_text_alignment_component = visual_text_alignment_component
        self.visual_visual_alignment_component = visual_visual_alignment_component
        self.filtering_merging_component = filtering_merging_component
        self.mask_generator = mask_generator
        self.time_start_ranking = None
        self.time_start_ranking_after_text_extraction = None
        self.time_end_ranking = None

    def predict(self, support_images: torch.Tensor, support_masks: torch.Tensor, query_image: torch.Tensor,
                mask_proposals: Optional[torch.Tensor] = None):
        self.time_start_ranking = time.time()
        assert (mask_proposals is not None or self.mask_generator is not None)  # mars/MARS.py:43
        if self.mask_generator is not None:
            mask_proposals = self.mask_generator.generate(support_images, support_masks, query_image)

        name, description = self.text_retriever_component.get_conceptual_information(
            support_images=support_images, support_masks=support_masks)
        self.time_start_ranking_after_text_extraction = time.time()

        vva_comp = self.visual_visual_alignment_component
        vva = vva_comp.compute(support_imgs=support_images, support_masks=support_masks, query_img=query_image)
        vta = self.visual_text_alignment_component.compute(query_image=query_image, fg_label=name, bg_labels=[])
        # nearest resize to the vva grid + min-max (mars/MARS.py:77-82), one kernel
        g = vva.shape[-1]
        vta = ops.resize_minmax(vta.to(vva.device).float()[None], g).reshape(g, g)

        text = [f"a {name}."] if description == "" else [f"a {name}, {description}."]
        predicted = self.filtering_merging_component.compute(
            query_img=query_image, mask_proposals=mask_proposals, support_mask=support_masks,
            cost_matrix=vva_comp.cost_matrix,
            patch_features_spatial_dimension=vva_comp.model_embedding_spatial_dimensions,
            vva=vva, vta=vta, text=text)
        torch.cuda.synchronize()  # the stamps below are host clocks; make them cover the device work
        self.time_end_ranking = time.time()
        return predicted

    def clear(self):
        self.visual_visual_alignment_component.clear()


def build_MARS_fss(args, text_retriever_component=None, visual_text_alignment_component=None, dino_model=None,
                   dino_transforms=None, alpha_clip_model=None, alpha_clip_transforms=None, mask_generator=None):
    """Assemble MARS (mars/MARS.py:110-116).

    `build_MARS_fss(args)` - the reference's one-argument form: the PyTorch producers (ViP-LLaVA text retriever,
    CLIP Grad-CAM component, DINOv2, AlphaCLIP) are loaded with the reference's own loaders, which must be importable
    (the reference checkout on `sys.path`); the ranking components on top are this package's.  Any producer passed
    explicitly is used as given, so already-loaded models are never loaded twice."""
    import os

    need_ref = [n for n, v in (("text retriever", text_retriever_component), ("visual-text component", visual_text_alignment_component),
                               ("DINOv2", dino_model), ("AlphaCLIP", alpha_clip_model)) if v is None]
    if need_ref:
        try:  # mars/MARS.py:9-10, VisualVisualAlignmentModule.py:133-154, FilteringMergingModule.py:222-230
            if text_retriever_component is None:
                from mars.components.TextRetrieverModule import build_text_retriever_component
                text_retriever_component = build_text_retriever_component(args=args)
            if visual_text_alignment_component is None:
                from mars.components.VisualTextAlignmentModule import build_visual_text_alignment_component
                visual_text_alignment_component = build_visual_text_alignment_component(args=args)
            if dino_model is None or alpha_clip_model is None:
                from utils.backbone_loader import BackboneLoader
            if dino_model is None:
                weights = 'dinov2_vitl14_reg4_pretrain.pth' if args.num_regs == 4 else 'dinov2_vitl14_pretrain.pth'
                dino_model, dino_transforms = BackboneLoader.load_backbone(
                    backbone_name='dinov2', backbone_size=args.dino_backbone, device=args.device,
                    backbone_weights_path=os.path.join(args.models_path, weights),
                    encoder_kwargs=dict(img_size=args.input_size, patch_size=14, init_values=1e-5, ffn_layer='mlp',
                                        block_chunks=0, num_register_tokens=args.num_regs, qkv_bias=True, proj_bias=True,
                                        ffn_bias=True))
            if alpha_clip_model is None:
                alpha_clip_model, alpha_clip_transforms = BackboneLoader.load_backbone(
                    backbone_name='alphaclip', backbone_size='ViT-L/14@336px', device=args.device,
                    backbone_weights_path=os.path.join(args.models_path, 'clip_l14_336_grit_20m_4xe.pth'),
                    encoder_kwargs={'device': args.device, 'download_root': args.models_path})
        except ImportError as ex:
            raise ImportError(f"build_MARS_fss(args): the {', '.join(need_ref)} producer(s) were not passed and the reference's "
                              f"loaders are not importable ({ex}); put the reference checkout on sys.path or pass the "
                              f"loaded PyTorch producers explicitly") from ex
    vva = VisualVisualAlignmentModule(
        model=dino_model, model_transforms=dino_transforms, model_patch_size=14,
        model_embedding_spatial_dimensions=args.input_size // 14, model_num_regs=args.num_regs,
        vva_refinement_box_threshold=args.vva_refinement_box_threshold,
        last_n_attention_maps_for_refinement=args.last_n_attn_for_vva_refinement, device=args.device)
    fm = FilteringMergingModule(
        alpha_clip_model=alpha_clip_model, img_transforms=alpha_clip_transforms[0],
        mask_transforms=alpha_clip_transforms[1], alpha=args.alpha_coverage,
        static_threshold=args.static_threshold, dynamic_threshold=args.dynamic_threshold, device=args.device)
    return MARS(text_retriever_component, visual_text_alignment_component, vva, fm, mask_generator)
