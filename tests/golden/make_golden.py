"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own modules.

Run in the build container only (the reference checkout does not exist on the
GPU box):

    python tests/golden/make_golden.py [/root/reference]

The reference's hot-path modules import three things that are not installed
here (POT ``ot``, ``alpha_clip``, ``utils.backbone_loader`` -> loralib); they
are shimmed in ``sys.modules`` exactly as SURVEY.md A.2 describes.  ``ot.emd2``
is replaced by the exact transport LP of ``oracle/mars_oracle.emd_exact``
(POT itself is unavailable: EMD parity is unpinned).  Backbones are fakes that
return seeded tensors, so the vectors pin everything the reference computes
*after* the backbones.

Outputs (small, committed): ``vva_*.npz``, ``pir_*.npz``, ``fm_*.npz``, ``eval_*.npz``, ``amg_*.npz``, ``diag_*.npz``, ``matcher_*.npz``, ``mars_*.npz``.
Inputs that would be large are regenerated from the recorded seed by
``tests/golden/cases.py`` and guarded by a checksum stored in the fixture.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True

from oracle import mars_oracle as orc  # noqa: E402
import cases  # noqa: E402


def install_shims(ref_root):
    sys.path.insert(0, ref_root)
    ot = types.ModuleType("ot")
    ot.emd2 = lambda a, b, M: np.float64(orc.emd_exact(np.asarray(M)))
    ac = types.ModuleType("alpha_clip")
    ac.tokenize = lambda t: torch.zeros(len(t), 77, dtype=torch.long)
    ub = types.ModuleType("utils.backbone_loader")
    ub.BackboneLoader = object
    sys.modules.update({"ot": ot, "alpha_clip": ac, "utils.backbone_loader": ub})


class FakeDino(torch.nn.Module):
    """Returns the seeded x_prenorm tensors in call order (support first, then query)."""

    def __init__(self, feats_in_call_order, attn_maps, embed_dim):
        super().__init__()
        self._feats = list(feats_in_call_order)
        self._attn = attn_maps
        self.embed_dim = embed_dim

    def forward_features(self, imgs):
        return {"x_prenorm": self._feats.pop(0)}

    def get_last_self_attention(self, img):
        return tuple(self._attn)


class FakeAlphaClip:
    def __init__(self, img_feats, text_feat):
        self._img = img_feats
        self._txt = text_feat
        self._cursor = 0
        self.visual = self._visual

    def _visual(self, image, alpha):
        n = alpha.shape[0]
        out = self._img[self._cursor:self._cursor + n]
        self._cursor += n
        return out

    def encode_text(self, tok):
        return self._txt.reshape(1, -1)


def gen_vva(name, spec):
    from mars.components.VisualVisualAlignmentModule import VisualVisualAlignmentModule

    c = cases.vva_inputs(spec)
    regs = spec["regs"]
    ns, n, cdim = c["feat_s"].shape
    pad = lambda f: torch.cat([torch.full((f.shape[0], 1 + regs, cdim), 7.0), f], dim=1)
    model = FakeDino([pad(c["feat_s"]), pad(c["feat_q"][None])], c["attn_maps"], cdim)
    mod = VisualVisualAlignmentModule(
        model=model, model_transforms=lambda x: x, model_patch_size=14,
        model_embedding_spatial_dimensions=spec["g"], model_num_regs=regs,
        vva_refinement_box_threshold=spec["thr"],
        last_n_attention_maps_for_refinement=spec["last_n"], device="cpu")
    h = spec["H"]
    out = mod.compute(torch.zeros(1, ns, 3, h, h), c["support_mask"][None], torch.zeros(1, 3, h, h))
    st = 1 if n <= 200 else 11  # keep the fixture small: strided sample of the big matrices
    np.savez_compressed(
        os.path.join(HERE, f"vva_{name}.npz"), spec=np.asarray(repr(spec)), stride=st,
        checksum=cases.checksum(c), sim=mod.similarity_matrix.numpy()[::st, ::st],
        cost=mod.cost_matrix.numpy()[::st, ::st], vva_refined=out.numpy())
    print("vva", name, out.shape, float(out.max()))


def gen_pir(name, spec):
    from mars.components.PriorInformationRefinementModule import PriorInformationRefinementModule

    c = cases.pir_inputs(spec)
    mod = PriorInformationRefinementModule(box_threshold=spec["thr"],
                                           last_n_attention_maps_for_refinement=spec["last_n"],
                                           device="cpu", num_regs=spec["regs"])
    out = mod.compute(prior=c["prior"], attn_maps=c["attn_maps"])
    box, cnt = mod._scoremap2bbox(c["prior"].numpy(), multi_contour_eval=True)
    np.savez_compressed(os.path.join(HERE, f"pir_{name}.npz"), spec=np.asarray(repr(spec)),
                        checksum=cases.checksum(c), refined=out.numpy(), boxes=np.asarray(box), cnt=cnt)
    print("pir", name, out.shape, cnt)


def gen_fm(name, spec):
    from mars.components.FilteringMergingModule import FilteringMergingModule

    c = cases.fm_inputs(spec)
    if spec.get("clip_fp16"):
        # POT's ot.emd2 returns a Python float for NumPy inputs; with float16 AlphaCLIP scores that matters: NumPy adds a
        # (weak) Python float to a float16 array IN float16, an np.float64 would promote the sum
        sys.modules["ot"].emd2 = lambda a, b, M: float(orc.emd_exact(np.asarray(M)))
    model = FakeAlphaClip(c["clip_img"], c["clip_txt"])
    mod = FilteringMergingModule(
        alpha_clip_model=model, img_transforms=lambda x: torch.zeros(3, 8, 8),
        mask_transforms=lambda m: torch.from_numpy(m)[None].float(),
        alpha=spec["alpha"], static_threshold=spec["static"], dynamic_threshold=spec["dynamic"], device="cpu")
    h = spec["H"]
    ranked = mod._score_proposals(
        query_img=torch.zeros(1, 3, h, h), mask_proposals=c["masks"], support_mask=c["support_mask"][None],
        cost_matrix=c["cost"], patch_features_spatial_dimension=spec["g"], vva=c["vva"], vta=c["vta"], text=["a thing."])
    merged = mod._merge_masks(ranked)
    order = []
    for m, _ in ranked:
        order.append([i for i in range(c["masks"].shape[0]) if m.data_ptr() == c["masks"][i].data_ptr()][0])
    scores = np.asarray([float(np.asarray(s).reshape(-1)[0]) for _, s in ranked], dtype=np.float64)
    # the emd scores the reference saw (1 - emd), recomputed through its own method
    pooled_sup = torch.nn.functional.adaptive_max_pool2d(c["support_mask"][:, None].float(), (spec["g"], spec["g"]))
    emd = []
    for m_p in c["masks"]:
        pm = torch.nn.functional.adaptive_max_pool2d(m_p[None].float(), (spec["g"], spec["g"]))[0]
        emd.append(mod._compute_emd(pooled_sup, pm, c["cost"]))
    np.savez_compressed(os.path.join(HERE, f"fm_{name}.npz"), spec=np.asarray(repr(spec)),
                        checksum=cases.checksum(c), order=np.asarray(order), scores=scores,
                        emd=np.asarray(emd, dtype=np.float64),
                        merged_bits=np.packbits(merged.numpy() > 0), merged_shape=np.asarray(merged.shape))
    print("fm", name, order[:8], scores[:4])
    if spec.get("clip_fp16"):
        sys.modules["ot"].emd2 = lambda a, b, M: np.float64(orc.emd_exact(np.asarray(M)))


def gen_eval(name, spec):
    """Evaluator.classify_prediction + AverageMeter.update / compute_iou of the reference, one episode per update
    like main_MARS.py:72-73.  mars/utils/logger.py imports comet_ml and tensorboardX (not installed): shimmed."""
    cm = types.ModuleType("comet_ml")
    cm.Experiment = object
    tb = types.ModuleType("tensorboardX")
    tb.SummaryWriter = object
    sys.modules.update({"comet_ml": cm, "tensorboardX": tb})
    from mars.utils.evaluation import Evaluator
    from mars.utils.logger import AverageMeter

    c = cases.eval_inputs(spec)
    Evaluator.initialize()
    ds = types.SimpleNamespace(benchmark=spec["benchmark"], class_ids=spec["class_ids"])
    meter = AverageMeter(ds, device="cpu")
    inters, unions = [], []
    for i in range(spec["n"]):
        batch = dict(query_mask=c["gt"][i:i + 1].clone())
        if c["ignore"] is not None:
            batch["query_ignore_idx"] = c["ignore"][i:i + 1].clone()
        ai, au = Evaluator.classify_prediction(c["pred"][i:i + 1].clone(), batch)
        meter.update(ai, au, c["class_id"][i:i + 1], loss=None)
        inters.append(ai[:, 0].numpy())
        unions.append(au[:, 0].numpy())
    miou, fb_iou, cats = meter.compute_iou()
    np.savez_compressed(os.path.join(HERE, f"eval_{name}.npz"), spec=np.asarray(repr(spec)),
                        area_inter=np.stack(inters), area_union=np.stack(unions),
                        intersection_buf=meter.intersection_buf.numpy(), union_buf=meter.union_buf.numpy(),
                        miou=np.float64(miou), fb_iou=np.float64(fb_iou), cats_iou=cats.numpy())


def gen_amg(name, spec, ref_root):
    """mask_to_rle_pytorch / rle_to_mask / batched_mask_to_box / calculate_stability_score of the reference's
    segment_anything/utils/amg.py (loaded by path: the package __init__ needs matplotlib) and torchvision's nms,
    which is what batched_nms with a single category reduces to (automatic_mask_generator.py:370-375)."""
    import importlib.util

    import torchvision

    sp = importlib.util.spec_from_file_location("ref_amg", os.path.join(ref_root, "segment_anything", "utils", "amg.py"))
    amg = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(amg)
    c = cases.amg_inputs(spec)
    rles = amg.mask_to_rle_pytorch(c["masks"])
    for r, m in zip(rles, c["masks"]):
        assert np.array_equal(amg.rle_to_mask(r), m.numpy())
    counts = np.concatenate([np.asarray(r["counts"], dtype=np.int32) for r in rles])
    offsets = np.cumsum([0] + [len(r["counts"]) for r in rles]).astype(np.int64)
    boxes = amg.batched_mask_to_box(c["masks"]).numpy()
    stab = amg.calculate_stability_score(c["logits"], 0.0, 1.0).numpy()
    keep = torchvision.ops.nms(c["boxes"], c["scores"], 0.5).numpy()
    np.savez_compressed(os.path.join(HERE, f"amg_{name}.npz"), spec=np.asarray(repr(spec)), counts=counts,
                        offsets=offsets, boxes=boxes, stability=stab, nms_keep=keep,
                        areas=np.asarray([amg.area_from_rle(r) for r in rles]))


def gen_diag(name, spec, ref_root):
    """get_ref_to_target_similarity / get_aposteriori_statistics: matcher/Matcher.py cannot be imported here
    (matplotlib, timm, POT), so the two method definitions are cut out of the file with `ast` and executed as
    they are on a stand-in object that carries the attributes they read."""
    import ast
    import types as _t

    import torch.nn.functional as F

    src = open(os.path.join(ref_root, "matcher", "Matcher.py")).read()
    tree = ast.parse(src)
    wanted = {"get_ref_to_target_similarity", "get_aposteriori_statistics"}
    funcs = [n for cls in tree.body if isinstance(cls, ast.ClassDef) for n in cls.body
             if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {f.name for f in funcs} == wanted
    ns_ = {"torch": torch, "F": F, "np": np}
    exec(compile(ast.Module(body=funcs, type_ignores=[]), "Matcher.py[extract]", "exec"), ns_)
    c = cases.diag_inputs(spec)
    g, nshot = spec["g"], spec["ns"]
    ref_n, tar_n = orc.normalize_rows(c["ref_raw"]), orc.normalize_rows(c["tar_raw"])
    S = ref_n @ tar_n.t()
    me = _t.SimpleNamespace(nshot=nshot, encoder=_t.SimpleNamespace(patch_size=1),
                            generator=_t.SimpleNamespace(predictor=_t.SimpleNamespace(model=_t.SimpleNamespace(mask_threshold=0.0))),
                            ref_masks_pool=c["ref_mask"].reshape(-1), S=S,
                            unnormalized_ref_feats=c["ref_raw"], unnormalized_tar_feat=c["tar_raw"])
    r2t = ns_["get_ref_to_target_similarity"](me, ref_n, tar_n, c["ref_mask"])
    stats = ns_["get_aposteriori_statistics"](me, c["tar_mask"].reshape(1, 1, g, g).float())
    np.savez_compressed(os.path.join(HERE, f"diag_{name}.npz"), spec=np.asarray(repr(spec)),
                        ref_to_target=r2t.reshape(-1).numpy(),
                        stats=np.asarray([stats["aposteriori_similarity_mean"], stats["aposteriori_similarity_max"],
                                          stats["aposteriori_similarity_std"], stats["embeddings_euclidean_distance"]]))


def load_reference_matcher(ref_root):
    """`Matcher` and `RobustPromptSampler` of matcher/Matcher.py with every method body as it is in the reference.  The
    module itself cannot be imported here (matplotlib, timm, POT, segment_anything), so the two class definitions are
    cut out with `ast` (minus the plotting methods) and executed in a namespace that supplies their imports; `ot.emd2`
    is the exact transport LP of the oracle (POT is not installed)."""
    import ast
    import random

    import cv2
    import torch.nn.functional as F
    from scipy.optimize import linear_sum_assignment

    tree = ast.parse(open(os.path.join(ref_root, "matcher", "Matcher.py")).read())
    skip = {"visualize_internal_state", "set_visualizer_parameters", "positive_prompts_experiment",
            "negative_prompts_from_discarded_experiment", "negative_prompts_from_cost_experiment"}
    classes = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("Matcher", "RobustPromptSampler"):
            node.body = [b for b in node.body if not (isinstance(b, ast.FunctionDef) and b.name in skip)]
            classes.append(node)
    assert len(classes) == 2
    ot = types.SimpleNamespace(emd2=lambda a, b, M: np.float64(orc.emd_exact(np.asarray(M))))
    ns_ = {"torch": torch, "F": F, "np": np, "cv2": cv2, "ot": ot, "random": random,
           "linear_sum_assignment": linear_sum_assignment, "SamAutomaticMaskGenerator": object,
           "kmeans_pp": None, "__name__": "reference_matcher"}
    exec(compile(ast.Module(body=classes, type_ignores=[]), "Matcher.py[extract]", "exec"), ns_)
    return ns_["Matcher"], ns_["RobustPromptSampler"]


def gen_matcher(name, spec, ref_root):
    """Matcher.predict of the reference, method bodies unchanged, on a fake encoder / SAM generator."""
    import random

    RefMatcher, _ = load_reference_matcher(ref_root)
    c = cases.matcher_inputs(spec)
    enc = cases.FakePatchEncoder(c["ref_raw"], c["tar_raw"], spec["ns"], spec["ps"], spec["C"])
    gen = cases.FakeSamGenerator(c["proposals"], c["point_coords"])
    size = spec["g"] * spec["ps"]
    m = RefMatcher(encoder=enc, encoder_transforms=lambda x: x, generator=gen, input_size=size,
                   sample_range=spec["sample_range"], max_sample_iterations=spec["max_iter"], alpha=spec["alpha"],
                   beta=spec["beta"], exp=spec["exp"], score_filter_cfg=dict(spec["cfg"]),
                   num_merging_mask=spec["num_merging_mask"],
                   use_negative_priors_from_discarded=spec["neg_discarded"],
                   use_negative_priors_from_cost=spec["neg_cost"], device=torch.device("cpu"))
    m.set_reference(c["ref_imgs"], c["ref_masks"].clone())
    m.set_target(c["tar_img"])
    ref_feats, tar_feat = m.extract_img_feats()
    pts, neg, box, S, C, reduced, reduced_neg = m.patch_level_matching(ref_feats, tar_feat)
    m.set_rps()
    per_mask = []
    inner = m.rps.get_mask_scores

    def recording(**kw):
        out = inner(**kw)
        per_mask.append((float(out[0][0]), float(out[1][0]), float(out[2])))
        return out

    m.rps.get_mask_scores = recording
    random.seed(spec["seed"])
    neg_for_generation = neg
    if isinstance(neg, list):  # see cases.py: the reference cannot run mask_generation with negative priors enabled
        m.use_negative_priors_from_discarded = m.use_negative_priors_from_cost = False
        neg_for_generation = neg[0]
    merged, final = m.mask_generation(m.tar_img_np, pts, box, pts, m.ref_masks_pool, C, neg_for_generation)
    # prompt sampling on its own, from a known seed and point set
    random.seed(spec["seed"] + 1)
    demo_pts = np.arange(2 * 11).reshape(11, 2)
    s_many, l_many = m.rps.sample_points(demo_pts, negative_points=demo_pts[:5] + 100)
    s_few, l_few = m.rps.sample_points(demo_pts[:5])

    def rows(a):
        a = np.asarray(a).reshape(-1, 2)
        return a[np.lexsort((a[:, 1], a[:, 0]))].astype(np.int64)

    neg_rows = [rows(x) for x in neg] if isinstance(neg, list) else [rows(neg)]
    stats = {**m.get_patch_matching_statistics(), **m.get_mask_generation_statistics()}
    np.savez_compressed(
        os.path.join(HERE, f"matcher_{name}.npz"), spec=np.asarray(repr(spec)),
        ref_masks_pool=m.ref_masks_pool.numpy(), points=rows(pts), n_neg_sets=len(neg_rows),
        **{f"neg{i}": r for i, r in enumerate(neg_rows)}, reduced=int(reduced),
        reduced_neg=np.asarray([(-1 if r is None else r) for r in reduced_neg], dtype=np.int64),
        sim=S.numpy(), per_mask=np.asarray(per_mask, dtype=np.float64), merged=merged.numpy() > 0,
        final=float(final), merged_count=int(m.number_of_merged_masks),
        masks_to_merge=(m.masks_to_merge.reshape(-1, size, size).numpy() > 0),
        stats_keys=np.asarray(sorted(stats)), stats_vals=np.asarray([float(stats[k]) for k in sorted(stats)]),
        sample_many=np.concatenate([x.reshape(-1) for x in s_many]), label_many=np.concatenate([x.reshape(-1) for x in l_many]),
        sample_few=np.concatenate([x.reshape(-1) for x in s_few]), label_few=np.concatenate([x.reshape(-1) for x in l_few]),
        combos_5_3=np.asarray(m.rps.combinations(5, 3)))


def gen_mars(name, spec, ref_root):
    """MARS.predict (mars/MARS.py:33-104, class body cut out with `ast`: the module imports nltk / CLIP / Grad-CAM) with
    the reference's own VisualVisualAlignmentModule and FilteringMergingModule underneath, fake backbones, a fake text
    retriever and a fake visual-text alignment component.  `ot.emd2` = the exact LP of the oracle (install_shims)."""
    import ast
    import time
    from typing import Optional

    import torch.nn.functional as F
    from mars.components.FilteringMergingModule import FilteringMergingModule
    from mars.components.VisualVisualAlignmentModule import VisualVisualAlignmentModule

    tree = ast.parse(open(os.path.join(ref_root, "mars", "MARS.py")).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "MARS"]
    ns_ = {"torch": torch, "F": F, "time": time, "Optional": Optional, "TextRetrieverModule": object,
           "VisualTextAlignmentModule": object, "VisualVisualAlignmentModule": VisualVisualAlignmentModule,
           "FilteringMergingModule": FilteringMergingModule}
    exec(compile(ast.Module(body=cls, type_ignores=[]), "MARS.py[extract]", "exec"), ns_)
    v = cases.VVA_CASES[spec["vva"]]
    c = cases.mars_inputs(spec)
    regs, cdim, h = v["regs"], v["C"], v["H"]
    ns = c["feat_s"].shape[0]
    pad = lambda f: torch.cat([torch.full((f.shape[0], 1 + regs, cdim), 7.0), f], dim=1)
    vva_mod = VisualVisualAlignmentModule(
        model=FakeDino([pad(c["feat_s"]), pad(c["feat_q"][None])], c["attn_maps"], cdim), model_transforms=lambda x: x,
        model_patch_size=14, model_embedding_spatial_dimensions=v["g"], model_num_regs=regs,
        vva_refinement_box_threshold=v["thr"], last_n_attention_maps_for_refinement=v["last_n"], device="cpu")
    fm = FilteringMergingModule(
        alpha_clip_model=FakeAlphaClip(c["clip_img"], c["clip_txt"]), img_transforms=lambda x: torch.zeros(3, 8, 8),
        mask_transforms=lambda m: torch.from_numpy(m)[None].float(), alpha=spec["alpha"],
        static_threshold=spec["static"], dynamic_threshold=spec["dynamic"], device="cpu")
    seen = {}
    inner = fm.compute

    def recording(**kw):
        seen.update(vva=kw["vva"].clone(), vta=kw["vta"].clone(), text=list(kw["text"]), g=kw["patch_features_spatial_dimension"])
        return inner(**kw)

    fm.compute = recording
    text = types.SimpleNamespace(get_conceptual_information=lambda support_images, support_masks: ("thing", spec["description"]))
    vta = types.SimpleNamespace(compute=lambda query_image, fg_label, bg_labels: c["vta_raw"])
    mars = ns_["MARS"](text, vta, vva_mod, fm)
    pred = mars.predict(torch.zeros(1, ns, 3, h, h), c["support_mask"][None], torch.zeros(1, 3, h, h), c["masks"])
    np.savez_compressed(os.path.join(HERE, f"mars_{name}.npz"), spec=np.asarray(repr(spec)), merged=pred.numpy() > 0,
                        vva=seen["vva"].numpy(), vta=seen["vta"].numpy(), text=np.asarray(seen["text"][0]), g=seen["g"])
    print("mars", name, pred.shape, int((pred > 0).sum()), seen["text"])


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    install_shims(ref_root)
    torch.set_num_threads(1)  # deterministic reductions in the fixtures
    for name, spec in cases.VVA_CASES.items():
        gen_vva(name, spec)
    for name, spec in cases.PIR_CASES.items():
        gen_pir(name, spec)
    for name, spec in cases.FM_CASES.items():
        gen_fm(name, spec)
    for name, spec in cases.EVAL_CASES.items():
        gen_eval(name, spec)
    for name, spec in cases.AMG_CASES.items():
        gen_amg(name, spec, ref_root)
    for name, spec in cases.DIAG_CASES.items():
        gen_diag(name, spec, ref_root)
    for name, spec in cases.MATCHER_CASES.items():
        gen_matcher(name, spec, ref_root)
    for name, spec in cases.MARS_CASES.items():
        gen_mars(name, spec, ref_root)


if __name__ == "__main__":
    main()
