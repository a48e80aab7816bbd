"""Seeded inputs for the golden fixtures (shared by make_golden.py and the tests).

Everything is generated with CPU generators of fixed seed so the tests can
rebuild the exact inputs the reference saw; each fixture stores a checksum of
its inputs so a generator drift is detected instead of silently compared.
Do not edit a case after its fixture has been generated - add a new one.
"""
import hashlib

import numpy as np
import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def checksum(case: dict) -> np.ndarray:
    h = hashlib.sha256()
    for k in sorted(case):
        v = case[k]
        if isinstance(v, (list, tuple)):
            for t in v:
                h.update(t.contiguous().numpy().tobytes())
        else:
            h.update(v.contiguous().numpy().tobytes())
    return np.asarray(h.hexdigest())


def blob_masks(n, h, w, seed, min_frac=0.01, max_frac=0.35, dup_every=0):
    """``[n,h,w]`` float32 0/1 masks: unions of 1-3 rectangles / ellipses; never empty."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    out = np.zeros((n, h, w), dtype=np.float32)
    for i in range(n):
        if dup_every and i and i % dup_every == 0:
            src = out[rs.randint(0, i)]
            out[i] = np.roll(src, 1, axis=1) if (i // dup_every) % 2 else src
            continue
        m = np.zeros((h, w), dtype=bool)
        for _ in range(rs.randint(1, 4)):
            frac = rs.uniform(min_frac, max_frac)
            aspect = rs.uniform(0.5, 2.0)
            bh = max(1, min(h, int(round(np.sqrt(frac * h * w * aspect)))))
            bw = max(1, min(w, int(round(np.sqrt(frac * h * w / aspect)))))
            y0 = rs.randint(0, h - bh + 1)
            x0 = rs.randint(0, w - bw + 1)
            if rs.rand() < 0.5:
                m[y0:y0 + bh, x0:x0 + bw] = True
            else:
                cy, cx = y0 + bh / 2.0, x0 + bw / 2.0
                m |= ((yy - cy) / (bh / 2.0 + 1e-6)) ** 2 + ((xx - cx) / (bw / 2.0 + 1e-6)) ** 2 <= 1.0
        if not m.any():
            m[h // 2, w // 2] = True
        out[i] = m
    return torch.from_numpy(out)


def proto_features(rows, c, seed, n_proto=8):
    """Features with shared prototypes so cosine similarities span a wide range."""
    g = _gen(seed)
    protos = torch.randn(n_proto, c, generator=_gen(seed + 7919))
    k = torch.randint(0, n_proto, (rows,), generator=g)
    return 0.6 * protos[k] + 0.8 * torch.randn(rows, c, generator=g)


def attention_stack(layers, heads, tokens, seed, four_d=True, sharp=2.0):
    """`sharp` scales the logits: 2.0 = diffuse attention, 12.0 = near-one-hot rows (peaky)."""
    g = _gen(seed)
    maps = []
    for _ in range(layers):
        a = torch.softmax(sharp * torch.randn(heads, tokens, tokens, generator=g), dim=-1)
        maps.append(a[None] if four_d else a)
    return maps


# ---------------------------------------------------------------- VVA cases
VVA_CASES = {
    # g=10 (H=140): tiny, 2-shot, register tokens, only the last 2 of 3 attention layers used
    "g10_2shot": dict(seed=11, g=10, H=140, ns=2, C=32, regs=4, layers=3, heads=2, last_n=2, thr=0.8),
    # native geometry 518/14 = 37, 1-shot
    "g37_1shot": dict(seed=12, g=37, H=518, ns=1, C=64, regs=4, layers=2, heads=2, last_n=24, thr=0.8),
    # support mask covering every patch: the reference skips the background term
    "g10_allfg": dict(seed=13, g=10, H=140, ns=1, C=32, regs=0, layers=1, heads=1, last_n=1, thr=0.5, all_fg=True),
}


def vva_inputs(spec):
    g, h, ns, c = spec["g"], spec["H"], spec["ns"], spec["C"]
    n = g * g
    seed = spec["seed"]
    feats = proto_features((ns + 1) * n, c, seed)
    support = blob_masks(ns, h, h, seed + 1, 0.05, 0.3)
    if spec.get("all_fg"):
        support = torch.ones_like(support)
    tokens = 1 + spec["regs"] + n
    return dict(feat_s=feats[:ns * n].reshape(ns, n, c).contiguous(), feat_q=feats[ns * n:].contiguous(),
                support_mask=support, attn_maps=attention_stack(spec["layers"], spec["heads"], tokens, seed + 2))


# ---------------------------------------------------------------- PIR cases
PIR_CASES = {
    "g12_multi": dict(seed=21, g=12, regs=0, layers=2, heads=3, last_n=8, thr=0.4, kind="blobs", four_d=True),
    "g33_border": dict(seed=22, g=33, regs=0, layers=1, heads=2, last_n=8, thr=0.4, kind="border", four_d=False),
    "g9_flat": dict(seed=23, g=9, regs=2, layers=1, heads=1, last_n=1, thr=0.8, kind="flat", four_d=True),
    "g16_rand": dict(seed=24, g=16, regs=1, layers=3, heads=2, last_n=2, thr=0.6, kind="rand", four_d=True),
    # adversarial for the R (R (B*p)) evaluation order: near-one-hot attention rows and a prior whose range is
    # 2 % of its mean, so the final min-max amplifies any rounding difference ~50x
    "g14_peaky_lowrange": dict(seed=25, g=14, regs=0, layers=2, heads=2, last_n=2, thr=0.4, kind="lowrange", four_d=True,
                               sharp=12.0),
    # the same attention at the native grid with a random prior
    "g37_peaky": dict(seed=26, g=37, regs=4, layers=1, heads=2, last_n=1, thr=0.8, kind="rand", four_d=True, sharp=10.0),
}


def pir_inputs(spec):
    g = spec["g"]
    seed = spec["seed"]
    rs = np.random.RandomState(seed)
    if spec["kind"] == "blobs":
        prior = 0.2 * rs.rand(g, g)
        prior[1:4, 1:5] += 0.7
        prior[7:11, 6:12] += 0.6      # touches the right border
        prior[5, 0] = 0.95            # isolated single pixel on the left border
        prior[9:12, 0:2] += 0.65      # touches the bottom border
    elif spec["kind"] == "border":
        prior = 0.1 * rs.rand(g, g)
        prior[g - 6:, g - 6:] = 0.9   # component touching the bottom-right corner
        prior[0:3, 0:3] = 0.8
        prior[10:20, 10:20] = 0.7
        prior[13:17, 13:17] = 0.05    # a hole inside a component
        prior[14:16, 14:16] = 0.75    # an island inside the hole
    elif spec["kind"] == "lowrange":
        prior = 0.6 + 0.012 * rs.rand(g, g)  # every pixel above the threshold: one component, box clipped at the border
    elif spec["kind"] == "flat":
        prior = np.zeros((g, g))      # no pixel above the threshold -> no contour -> B == 0
    else:
        prior = rs.rand(g, g)
    prior = torch.from_numpy(np.clip(prior, 0, 1).astype(np.float32))
    tokens = 1 + spec["regs"] + g * g
    return dict(prior=prior, attn_maps=attention_stack(spec["layers"], spec["heads"], tokens, seed + 2, spec["four_d"],
                                                       spec.get("sharp", 2.0)))


# ---------------------------------------------------------------- FM cases
FM_CASES = {
    # tiny geometry with overlapping pooling bins (H=100, g=7: 100/7 is not an integer)
    "g7_overlap": dict(seed=31, g=7, H=100, ns=1, P=9, D=32, alpha=0.85, static=0.55, dynamic=0.95, dup_every=4),
    # native geometry, 2-shot support (shot-major cost rows)
    "g37_native": dict(seed=32, g=37, H=518, ns=2, P=6, D=48, alpha=0.85, static=0.55, dynamic=0.95, dup_every=0, small_support=True),
    # high static threshold -> the dynamic branch of the merge is taken
    "g10_dynamic": dict(seed=33, g=10, H=140, ns=1, P=12, D=16, alpha=0.7, static=0.99, dynamic=0.8, dup_every=5),
    # every proposal is the same mask: all four scores tie, min-max gives 0 / 1e-7, the stable rule keeps index order
    "g8_identical": dict(seed=34, g=8, H=112, ns=1, P=6, D=16, alpha=0.85, static=0.55, dynamic=0.95, dup_every=0,
                         identical=True),
    # 1024 -> 37 pooling (28/29-pixel bins that overlap by one pixel): proposal 0 is the single pixel (27, 27), which lies
    # in the overlap of bins 0 and 1 on both axes (4 patches); proposal 1 is the single pixel (28, 28) (1 patch)
    "g37_onepixel_1024": dict(seed=35, g=37, H=1024, ns=1, P=5, D=16, alpha=0.85, static=0.55, dynamic=0.95, dup_every=0,
                              onepixel=True),
    # AlphaCLIP features in float16, as the reference runs them on a GPU (FilteringMergingModule.py:97,126-136,189,195):
    # the dot products, their min-max and the first addition of the fusion are float16 arithmetic in NumPy
    "g10_clipfp16": dict(seed=36, g=10, H=140, ns=1, P=14, D=64, alpha=0.85, static=0.55, dynamic=0.95, dup_every=0,
                         clip_fp16=True),
}


def fm_inputs(spec):
    g, h, ns, p, d = spec["g"], spec["H"], spec["ns"], spec["P"], spec["D"]
    n = g * g
    seed = spec["seed"]
    gen = _gen(seed)
    masks = blob_masks(p, h, h, seed + 1, 0.01, 0.2, spec["dup_every"])
    if spec.get("small_support"):
        support = blob_masks(ns, h, h, seed + 2, 0.01, 0.03)
    else:
        support = blob_masks(ns, h, h, seed + 2, 0.05, 0.25)
    if spec.get("identical"):
        masks = masks[:1].repeat(p, 1, 1).contiguous()
    if spec.get("onepixel"):
        masks[0] = 0
        masks[0, 27, 27] = 1
        masks[1] = 0
        masks[1, 28, 28] = 1
    fs = torch.nn.functional.normalize(proto_features(ns * n, 24, seed + 3), dim=1)
    fq = torch.nn.functional.normalize(proto_features(n, 24, seed + 4), dim=1)
    cost = (1 - fs @ fq.T) / 2
    img = torch.nn.functional.normalize(torch.randn(p, d, generator=gen), dim=1)
    txt = torch.nn.functional.normalize(torch.randn(d, generator=gen), dim=0)
    vva = torch.rand(g, g, generator=gen)
    vta = torch.rand(g, g, generator=gen)
    if spec.get("clip_fp16"):
        img, txt = img.half(), txt.half()
    return dict(masks=masks, support_mask=support, cost=cost.contiguous(), clip_img=img, clip_txt=txt, vva=vva, vta=vta)


def alphaclip_as_seen(c):
    """The AlphaCLIP (image, text) features as the reference's dot product sees them: its producer code re-normalises
    what the model returns in the feature dtype (FilteringMergingModule.py:179,201).  The float32 fixtures hold already
    normalised features (the step is an identity up to 1 ulp and was generated without it); for float16 features the
    half-precision normalisation is part of the input the fixture pins."""
    img, txt = c["clip_img"], c["clip_txt"]
    if img.dtype == torch.float16:
        img = img / img.norm(dim=-1, keepdim=True)
        txt = txt.reshape(1, -1)
        txt = (txt / txt.norm(dim=-1, keepdim=True)).reshape(-1)
    return img, txt


# ---- evaluator / AverageMeter: episodes of (prediction, ground truth, ignore band, class id)
EVAL_CASES = {
    "coco_like": dict(benchmark="coco", class_ids=list(range(0, 80, 4)), n=24, H=96, W=128, seed=801, ignore=False),
    "pascal_ignore": dict(benchmark="pascal5i", class_ids=[1, 2, 3, 4, 5], n=15, H=80, W=80, seed=802, ignore=True),
}


def eval_inputs(spec):
    """pred / gt [n,H,W] float32 0/1, ignore [n,H,W] (disjoint from gt) or None, class_id [n] int64 (0-based)."""
    n, h, w = spec["n"], spec["H"], spec["W"]
    gt = blob_masks(n, h, w, spec["seed"], 0.02, 0.4)
    noise = blob_masks(n, h, w, spec["seed"] + 1, 0.01, 0.2)
    pred = ((gt + torch.roll(noise, 3, dims=2)) > 0).float() * (torch.roll(gt, 5, dims=1) + noise > 0).float()
    pred[0] = 0  # an empty prediction
    pred[1] = gt[1]  # a perfect one
    ignore = None
    if spec["ignore"]:
        band = blob_masks(n, h, w, spec["seed"] + 2, 0.01, 0.1)
        ignore = band * (1 - gt)
    rs = np.random.RandomState(spec["seed"] + 3)
    ids = spec["class_ids"]
    zero_based = [i - 1 for i in ids] if spec["benchmark"] == "pascal5i" else ids
    class_id = torch.from_numpy(rs.choice(zero_based, size=n)).long()
    return dict(pred=pred, gt=gt, ignore=ignore, class_id=class_id)


# ---- SAM-AMG post-processing
AMG_CASES = {
    "blobs_64x96": dict(n=10, H=64, W=96, seed=901),
    "blobs_128": dict(n=6, H=128, W=128, seed=902),
}


def amg_inputs(spec):
    """masks [n,H,W] bool (one empty, one full, one single pixel), logits [n,H,W] fp32, boxes [k,4] fp32 + scores."""
    n, h, w = spec["n"], spec["H"], spec["W"]
    masks = blob_masks(n, h, w, spec["seed"], 0.02, 0.4) > 0
    masks[0] = False
    masks[1] = True
    masks[2] = False
    masks[2, h - 1, w - 1] = True
    rs = np.random.RandomState(spec["seed"] + 1)
    logits = torch.from_numpy(rs.randn(n, h, w).astype(np.float32)) * 2 + (masks.float() * 4 - 2)
    k = 40
    xy = rs.rand(k, 2) * np.array([w * 0.6, h * 0.6])
    wh = rs.rand(k, 2) * np.array([w * 0.4, h * 0.4]) + 2
    boxes = np.concatenate([xy, xy + wh], axis=1).astype(np.float32)
    boxes[k // 2:] = boxes[:k - k // 2] + rs.randn(k - k // 2, 4).astype(np.float32)  # near duplicates
    scores = rs.rand(k).astype(np.float32)
    scores[5] = scores[7]  # a tie
    return dict(masks=masks, logits=logits, boxes=torch.from_numpy(boxes), scores=torch.from_numpy(scores))


# ---- Matcher diagnostics
DIAG_CASES = {"g12_2shot": dict(g=12, ns=2, C=48, seed=951), "g16_1shot": dict(g=16, ns=1, C=64, seed=952)}


def diag_inputs(spec):
    g, ns, c = spec["g"], spec["ns"], spec["C"]
    n = g * g
    ref_raw = proto_features(ns * n, c, spec["seed"]) * 3.0 + 0.5
    tar_raw = proto_features(n, c, spec["seed"] + 1) * 3.0 + 0.5
    ref_mask = blob_masks(ns, g, g, spec["seed"] + 2, 0.1, 0.4).reshape(ns, n)
    tar_mask = blob_masks(1, g, g, spec["seed"] + 3, 0.1, 0.4).reshape(n)
    return dict(ref_raw=ref_raw, tar_raw=tar_raw, ref_mask=ref_mask, tar_mask=tar_mask)


# ---- Matcher.predict through the reference's own methods (matcher/Matcher.py), fake encoder / SAM generator --------
# All cases use the score-filter merge rule: the reference's top-k branch cannot run (it indexes a [1,H,W] array with
# (y, x) at Matcher.py:822-827 and raises IndexError), so the top-k rule is checked against the oracle restatement only.
# With negative priors enabled mask_generation raises as well (:782 iterates a list of point arrays), so those cases
# switch the flags off after patch_level_matching and hand the first negative set on as plain discarded points.
MATCHER_CASES = {
    "g10_1shot_basic": dict(g=10, ps=14, ns=1, C=32, n_masks=14, seed=1201, alpha=1.0, beta=0.0, exp=0.0,
                           num_merging_mask=5, sample_range=(2, 4), max_iter=6, neg_discarded=False, neg_cost=False,
                           cfg=dict(emd=0.0, purity=0.0, coverage=0.0, score_filter=True, score=0.5, score_norm=0.05,
                                    topk_scores_threshold=0.0)),
    "g12_2shot_filters": dict(g=12, ps=14, ns=2, C=48, n_masks=18, seed=1202, alpha=0.8, beta=0.2, exp=1.0,
                              num_merging_mask=6, sample_range=(1, 3), max_iter=5, neg_discarded=True, neg_cost=False,
                              cfg=dict(emd=0.6, purity=0.02, coverage=0.05, score_filter=True, score=0.5,
                                       score_norm=0.5, topk_scores_threshold=0.0)),
    "g10_1shot_negcost": dict(g=10, ps=14, ns=1, C=32, n_masks=12, seed=1203, alpha=1.0, beta=0.5, exp=0.5,
                              num_merging_mask=4, sample_range=(2, 3), max_iter=4, neg_discarded=False, neg_cost=True,
                              cfg=dict(emd=0.0, purity=0.0, coverage=0.2, score_filter=True, score=0.45,
                                       score_norm=0.8, topk_scores_threshold=0.0)),
}


def matcher_inputs(spec):
    """Seeded inputs of a Matcher case: raw patch features, support masks, SAM-style proposals with the prompt points
    the fake generator reports for them (independent of the sampled prompts, so both sides see the same proposals)."""
    g, ps, ns, c = spec["g"], spec["ps"], spec["ns"], spec["C"]
    n, size = g * g, g * ps
    rs = np.random.RandomState(spec["seed"] + 7)
    ref_raw = proto_features(ns * n, c, spec["seed"]) * 2.0 + 0.25
    tar_raw = proto_features(n, c, spec["seed"] + 1) * 2.0 + 0.25
    ref_masks = blob_masks(ns, size, size, spec["seed"] + 2, 0.08, 0.3).reshape(1, ns, size, size)
    proposals = blob_masks(spec["n_masks"], size, size, spec["seed"] + 3, 0.01, 0.3).numpy() > 0
    proposals[1] = False  # an empty proposal: the reference scores it against every patch (Matcher.py:1181-1185)
    point_coords = [[[int(rs.randint(0, size)), int(rs.randint(0, size))] for _ in range(rs.randint(1, 4))]
                    for _ in range(spec["n_masks"])]
    ref_imgs = torch.zeros(1, ns, 3, size, size)
    tar_img = torch.zeros(1, 3, size, size)
    return dict(ref_raw=ref_raw, tar_raw=tar_raw, ref_masks=ref_masks, proposals=proposals, point_coords=point_coords,
                ref_imgs=ref_imgs, tar_img=tar_img)


class FakeSamGenerator:
    """Stands in for SamAutomaticMaskGenerator: fixed proposals, records the prompts it was given."""

    class _Model:
        mask_threshold = 0.0

    class _Predictor:
        pass

    def __init__(self, proposals, point_coords):
        self.proposals, self.point_coords = proposals, point_coords
        self.predictor = self._Predictor()
        self.predictor.model = self._Model()
        self.calls, self.resets = [], 0

    def generate(self, image, select_point_coords=None, select_point_labels=None, select_box=None,
                 select_mask_input=None):
        self.calls.append(dict(coords=select_point_coords, labels=select_point_labels, box=select_box))
        return [dict(segmentation=m, point_coords=pc) for m, pc in zip(self.proposals, self.point_coords)]

    def reset_stored_features(self):
        self.resets += 1


class FakePatchEncoder(torch.nn.Module):
    """DINOv2-shaped encoder stub: x_prenorm = [cls | patches], support call first, then the query."""
    family = "vits"
    num_register_tokens = 0

    def __init__(self, ref_raw, tar_raw, ns, patch_size, embed_dim):
        super().__init__()
        n = tar_raw.shape[0]
        self._feats = [torch.cat([torch.zeros(ns, 1, embed_dim), ref_raw.reshape(ns, n, embed_dim)], dim=1),
                       torch.cat([torch.zeros(1, 1, embed_dim), tar_raw.reshape(1, n, embed_dim)], dim=1)]
        self._i = 0
        self.patch_size, self.embed_dim = patch_size, embed_dim

    def forward_features(self, imgs):
        out = self._feats[self._i % 2].to(imgs.device)
        self._i += 1
        return {"x_prenorm": out}


# ---- MARS.predict end to end through the reference's own MARS / VisualVisualAlignmentModule / FilteringMergingModule ----
MARS_CASES = {
    "g10_2shot": dict(vva="g10_2shot", P=14, gt=8, D=24, seed=2101, alpha=0.85, static=0.55, dynamic=0.95, description=""),
    "g10_allfg": dict(vva="g10_allfg", P=9, gt=7, D=16, seed=2102, alpha=0.7, static=0.4, dynamic=0.9,
                      description="seen from above"),
}


def mars_inputs(spec):
    v = VVA_CASES[spec["vva"]]
    c = vva_inputs(v)
    gen = _gen(spec["seed"])
    masks = blob_masks(spec["P"], v["H"], v["H"], seed=spec["seed"] + 1, min_frac=0.02, max_frac=0.3)
    vta_raw = torch.rand(spec["gt"], spec["gt"], generator=gen)
    clip_img = torch.nn.functional.normalize(torch.randn(spec["P"], spec["D"], generator=gen), dim=1)
    clip_txt = torch.nn.functional.normalize(torch.randn(spec["D"], generator=gen), dim=0)
    return dict(c, masks=masks, vta_raw=vta_raw, clip_img=clip_img, clip_txt=clip_txt)
