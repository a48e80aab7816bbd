"""CPU oracle for the MARS proposal scoring / ranking / merging stage.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file, and only as the checker or the
timed CPU baseline.  The shipped path (the ``marsb200`` package) never imports
it and fails loudly when its CUDA library is missing.

Every function restates, in plain torch/numpy on the CPU, one piece of the
reference's algorithm and cites the reference ``file:line`` it follows
(paths relative to the reference checkout).  Where the reference has no code
(mask-vs-mask intersection/union, mask NMS — SURVEY.md §0 D2) the function is
a builder-defined specification and says so.

Pinning: ``tests/golden/make_golden.py`` runs the *reference's own modules*
(imported from the reference checkout with three import shims) on seeded
inputs and commits inputs+outputs under ``tests/golden/``; the CPU test-suite
checks this oracle against those vectors (alignment prior, prior refinement,
proposal scoring / merging, evaluator + AverageMeter, SAM-AMG RLE / boxes /
stability score, torchvision box NMS, and the whole Matcher ancestry path -
``set_reference`` -> ``patch_level_matching`` -> ``mask_generation`` with its ``RobustPromptSampler`` - by executing the
reference's own method bodies, cut out of matcher/Matcher.py with ``ast`` - and ``MARS.predict`` end to end through the
reference's ``MARS``, ``VisualVisualAlignmentModule`` and ``FilteringMergingModule`` classes).  The builder-defined pieces
(`pairwise_intersections`, `mask_nms`), the exact-EMD stand-in (POT 0.9.4 is not
installed anywhere we can run; the optimum of the LP is solver-independent and four
independent exact solvers agree on it: HiGHS here, a network simplex - POT's algorithm
class - from networkx in `emd_network_simplex`, a from-scratch C network simplex
(`oracle/emd_netsimplex.c`, bound by `oracle/emd_c.py`; also the host EMD baseline of
bench.py) and the lcm-expanded assignment in the tests) and the
Matcher assignment matching (scipy's
LSAP tie-breaking is implementation-defined; compared by objective value) have
no reference output to pin against: **parity unpinned** for those four, pinned
for everything else.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-7  # the reference's min-max / ratio epsilon (FilteringMergingModule.py:108-132)


# --------------------------------------------------------------------------
# A1 / A2 / A3: features -> similarity -> visual-visual prior
# --------------------------------------------------------------------------
def normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """Row-wise L2 normalisation, ``x / max(||x||, 1e-12)``.

    Follows mars/components/VisualVisualAlignmentModule.py:124-125 and
    matcher/Matcher.py:297-298.
    """
    x = x.reshape(-1, x.shape[-1]).float()
    return F.normalize(x, p=2, dim=1)


def similarity_and_cost(fs_n: torch.Tensor, fq_n: torch.Tensor):
    """``S = Fs @ Fq^T`` and the transport cost ``(1 - S) / 2``.

    Follows VisualVisualAlignmentModule.py:69-70 and matcher/Matcher.py:437-440.
    """
    sim = torch.matmul(fs_n, fq_n.T)
    return sim, (1 - sim) / 2


def pool_mask(mask: torch.Tensor, g: int) -> torch.Tensor:
    """Adaptive max pool of ``[..., H, W]`` masks to ``[..., g, g]`` booleans.

    Bin ``i`` spans ``[floor(i*H/g), ceil((i+1)*H/g))`` (PyTorch's adaptive
    pooling rule).  Follows FilteringMergingModule.py:73-76,
    104-107 and VisualVisualAlignmentModule.py:72-75.  The reference tests the
    pooled value with ``> 0`` (proposals) or ``.bool()`` (support mask); both
    agree for the non-negative masks the pipeline carries.
    """
    lead = mask.shape[:-2]
    m = mask.reshape(-1, 1, mask.shape[-2], mask.shape[-1]).float()
    pooled = F.adaptive_max_pool2d(m, (g, g))
    return (pooled > 0).reshape(*lead, g, g)


def minmax(x):
    """``(x - min) / (1e-7 + max - min)`` (VisualVisualAlignmentModule.py:102, mars/MARS.py:82)."""
    return (x - x.min()) / (EPS + x.max() - x.min())


def vva_prior(fs_n: torch.Tensor, fq_n: torch.Tensor, support_bits: torch.Tensor, g: int) -> torch.Tensor:
    """Foreground-minus-background ``mean * max`` prior over support rows, min-max scaled.

    ``support_bits`` is the shot-major flattened pooled support mask
    ``[ns*N]`` (bool).  Follows VisualVisualAlignmentModule.py:76-102: the
    products are recomputed on the fg / bg row subsets exactly as the reference
    does, the bg term is skipped when there are no bg rows.
    """
    sel = support_bits.reshape(-1).bool()
    s_fg = torch.matmul(fs_n[sel], fq_n.T)
    s_bg = torch.matmul(fs_n[~sel], fq_n.T)
    vva = s_fg.mean(dim=0) * s_fg.max(dim=0).values
    if s_bg.shape[0] != 0:
        vva = vva - s_bg.mean(dim=0) * s_bg.max(dim=0).values
    vva = vva.reshape(g, g)
    return minmax(vva)


# --------------------------------------------------------------------------
# A4: prior-information refinement (PIR)
# --------------------------------------------------------------------------
def attention_mean(attn_maps: Sequence[torch.Tensor], last_n: int, num_regs: int) -> torch.Tensor:
    """Mean over the last ``last_n`` layers and all heads of the patch-to-patch attention.

    Accepts ``[1,h,T,T]`` or ``[h,T,T]`` maps; drops the cls + register rows
    and columns.  Follows PriorInformationRefinementModule.py:31-45.
    """
    r = 1 + num_regs
    if attn_maps[0].dim() == 4:
        stack = torch.stack([a[0, :, r:, r:] for a in attn_maps], dim=0)[-last_n:]
    else:
        stack = torch.stack([a[:, r:, r:] for a in attn_maps], dim=0)[-last_n:]
    return torch.mean(stack, dim=(0, 1)).float()


def _components_8(fg: np.ndarray):
    """Bounding boxes (x, y, w, h) of the 8-connected components of a boolean image.

    Pure-numpy flood fill; stands in for cv2.findContours + cv2.boundingRect
    (PriorInformationRefinementModule.py:103-116).  Hole contours returned by
    RETR_TREE lie inside their component's box and never change the box union
    (SURVEY.md A.1-4), so components are sufficient.
    """
    h, w = fg.shape
    seen = np.zeros_like(fg, dtype=bool)
    boxes = []
    for y in range(h):
        for x in range(w):
            if not fg[y, x] or seen[y, x]:
                continue
            stack = [(y, x)]
            seen[y, x] = True
            x0 = x1 = x
            y0 = y1 = y
            while stack:
                cy, cx = stack.pop()
                x0, x1 = min(x0, cx), max(x1, cx)
                y0, y1 = min(y0, cy), max(y1, cy)
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        ny, nx = cy + dy, cx + dx
                        if 0 <= ny < h and 0 <= nx < w and fg[ny, nx] and not seen[ny, nx]:
                            seen[ny, nx] = True
                            stack.append((ny, nx))
            boxes.append((x0, y0, x1 - x0 + 1, y1 - y0 + 1))
    return boxes


def box_mask(prior: np.ndarray, threshold: float, use_cv2: bool = False) -> np.ndarray:
    """The 0/1 box mask ``B`` of PIR, ``[h, w]`` float32.

    ``img = uint8(prior * 255)`` (truncation), ``thr = int(threshold * max(img))``,
    keep ``img > thr``, one box per connected component with the reference's
    clipping quirk ``x1 = min(x + w, W - 1)`` and *exclusive* fill
    (PriorInformationRefinementModule.py:56-63, 91-122).  ``use_cv2=True`` runs
    the same OpenCV calls as the reference (for cross-checking the flood fill).
    """
    height, width = prior.shape
    img = (prior * 255).astype(np.uint8)
    thr = int(threshold * np.max(img))
    fg = img > thr
    if use_cv2:
        import cv2

        _, binary = cv2.threshold(src=np.expand_dims(img, 2), thresh=thr, maxval=255, type=cv2.THRESH_BINARY)
        contours = cv2.findContours(image=binary, mode=cv2.RETR_TREE, method=cv2.CHAIN_APPROX_SIMPLE)[0]
        rects = [cv2.boundingRect(c) for c in contours]
    else:
        rects = _components_8(fg)
    out = np.zeros((height, width), dtype=np.float32)
    for (x, y, w, h) in rects:
        x1 = min(x + w, width - 1)
        y1 = min(y + h, height - 1)
        out[y:y1, x:x1] = 1
    return out


def pir_refine(prior: torch.Tensor, attn_mean: torch.Tensor, threshold: float, use_cv2: bool = False) -> torch.Tensor:
    """Refine a ``[g,g]`` prior with the attention-derived matrix ``R``.

    ``D = A / colsum; D = D / rowsum; R = max(D, D D^T); R = R R;
    out = (R * B) @ prior``.  Follows PriorInformationRefinementModule.py:47-89.
    """
    shape = prior.shape
    p_np = prior.detach().cpu().numpy().astype(np.float32)
    b = torch.from_numpy(box_mask(p_np, threshold, use_cv2=use_cv2)).reshape(1, -1)
    a = attn_mean.float()
    d = a / torch.sum(a, dim=0, keepdim=True)
    d = d / torch.sum(d, dim=1, keepdim=True)
    r = torch.max(d, d @ d.t())
    r = torch.matmul(r, r)
    out = torch.matmul(r * b, torch.from_numpy(p_np).reshape(-1, 1))
    return out.reshape(shape)


def nearest_resize(x: torch.Tensor, size) -> torch.Tensor:
    """``F.interpolate(mode='nearest')`` of a 2-D map (mars/MARS.py:77-81)."""
    return F.interpolate(x[None, None].float(), size, mode="nearest")[0, 0]


# --------------------------------------------------------------------------
# A7: exact EMD (stand-in for POT's ot.emd2; parity unpinned)
# --------------------------------------------------------------------------
def emd_exact(cost: np.ndarray) -> float:
    """Exact optimal-transport cost with uniform marginals ``1/T`` and ``1/M``.

    The reference calls ``ot.emd2`` (POT 0.9.4, network simplex) at
    FilteringMergingModule.py:160-166 / matcher/Matcher.py:1187-1193; POT is
    not installed here, so this solves the same transportation LP with HiGHS.
    An empty marginal (all-zero proposal) is defined as cost 0.
    """
    from scipy.optimize import linprog
    from scipy.sparse import lil_matrix

    c = np.asarray(cost, dtype=np.float64)
    t, m = c.shape
    if t == 0 or m == 0:
        return 0.0
    a_eq = lil_matrix((t + m, t * m))
    for i in range(t):
        a_eq[i, i * m:(i + 1) * m] = 1
    for j in range(m):
        a_eq[t + j, j::m] = 1
    b_eq = np.concatenate([np.full(t, 1.0 / t), np.full(m, 1.0 / m)])
    res = linprog(c.reshape(-1), A_eq=a_eq.tocsr(), b_eq=b_eq, bounds=(0, None), method="highs")
    if res.status != 0:
        raise RuntimeError(f"transport LP failed: {res.message}")
    return float(res.fun)


def emd_network_simplex(cost: np.ndarray) -> float:
    """The same transport LP solved by a NETWORK SIMPLEX (networkx 3.x), the algorithm class POT's ``ot.emd2`` uses
    (LEMON network simplex; FilteringMergingModule.py:160-166).  Integer arithmetic throughout: float32 costs in
    [2^-5, 1) are exact multiples of 2^-29, supplies are M units and demands T units of mass 1/(T M), so the optimum
    comes out exact.  Second, independent exact solver beside ``emd_exact`` (HiGHS); used by the tests only."""
    import networkx as nx

    c = np.asarray(cost, dtype=np.float64)
    t, m = c.shape
    if t == 0 or m == 0:
        return 0.0
    scale = float(1 << 29)
    ci = c * scale
    if not np.all(ci == np.rint(ci)):
        raise ValueError("costs must be float32 values in [2^-5, 1) (exact multiples of 2^-29)")
    ci = ci.astype(np.int64)
    graph = nx.DiGraph()
    graph.add_nodes_from(((0, i) for i in range(t)), demand=-m)
    graph.add_nodes_from(((1, j) for j in range(m)), demand=t)
    graph.add_weighted_edges_from(((0, i), (1, j), int(ci[i, j])) for i in range(t) for j in range(m))
    total, _ = nx.network_simplex(graph)
    return float(total) / scale / (t * m)


def emd_score(pooled_support: torch.Tensor, pooled_proposal: torch.Tensor, cost_matrix: torch.Tensor) -> float:
    """``1 - emd`` on the cost sub-matrix (FilteringMergingModule.py:142-169)."""
    sub = cost_matrix[pooled_support.reshape(-1).bool(), :][:, pooled_proposal.reshape(-1).bool()]
    return 1.0 - emd_exact(sub.cpu().numpy())


# --------------------------------------------------------------------------
# A6 / A8 / A11: per-proposal region scores, fusion, ranking, merge
# --------------------------------------------------------------------------
def region_scores(masks: torch.Tensor, vva: np.ndarray, vta: np.ndarray, g: int):
    """Per-proposal pooled bitmap, coverage and the two alignment means.

    Loops over the proposals exactly like the reference hot loop
    (FilteringMergingModule.py:77-81, 103-110); quotients come out float64 as
    in the reference (numpy float32 / float64).
    Returns ``pooled [P,g,g] bool, coverage [P], pvv_align [P], pvt_align [P]``.
    """
    union = (torch.sum(masks, dim=0) > 0).float()
    pooled_union = F.adaptive_max_pool2d(union.unsqueeze(0), (g, g)).squeeze(0).numpy() > 0
    pooled, cov, avv, avt = [], [], [], []
    for m_p in masks:
        pm = F.adaptive_max_pool2d(m_p.unsqueeze(0).float(), (g, g)).squeeze(0).numpy() > 0
        n = np.sum(pm)
        cov.append(n / (EPS + np.sum(pooled_union)))
        avv.append(np.sum(vva[pm]) / (EPS + n))
        avt.append(np.sum(vta[pm]) / (EPS + n))
        pooled.append(pm)
    return np.stack(pooled), np.asarray(cov, dtype=np.float64), np.asarray(avv, dtype=np.float64), np.asarray(avt, dtype=np.float64)


def fuse_scores(emd_scores, clip_scores, coverage, pvv_align, pvt_align, alpha: float) -> np.ndarray:
    """Fused proposal score, ``[P]`` float64.

    ``pvv = a*align + (1-a)*coverage`` (FilteringMergingModule.py:118-119),
    min-max of the EMD and AlphaCLIP scores over the proposals (:126-132; the
    AlphaCLIP min-max is evaluated in the feature dtype), mean of the four (:136).

    float32 AlphaCLIP scores (the primary oracle): float32 min-max, float64 sum.
    float16 AlphaCLIP scores (the reference as run on a GPU: features are ``.half()``, :189,195, and
    ``(img_feats @ text_feats.T).cpu().numpy()`` is a float16 array, :97): the expressions of :126-136 are
    evaluated with the reference's own operand types - Python floats for the EMD (POT returns a Python float), NumPy
    float16 for AlphaCLIP, np.float64 for pvv / pvt - so NumPy's promotion rules (a Python float is weak: ``1e-7 + max``
    and ``emd_n + clip_n`` are float16 operations) apply exactly as they do in the reference.
    """
    pvv = alpha * np.asarray(pvv_align) + (1 - alpha) * np.asarray(coverage)
    pvt = alpha * np.asarray(pvt_align) + (1 - alpha) * np.asarray(coverage)
    if np.asarray(clip_scores).dtype == np.float16:
        emd = [float(v) for v in np.asarray(emd_scores, dtype=np.float64)]
        clip = [np.asarray([v], dtype=np.float16) for v in np.asarray(clip_scores).reshape(-1)]  # shape-(1,) arrays, :97
        min_emd, max_emd = min(emd), max(emd)
        min_c, max_c = min(clip), max(clip)
        emd_n = [(v - min_emd) / (1e-7 + max_emd - min_emd) for v in emd]
        clip_n = [(v - min_c) / (1e-7 + max_c - min_c) for v in clip]
        out = [(emd_n[i] + clip_n[i] + np.float64(pvv[i]) + np.float64(pvt[i])) / 4 for i in range(len(emd))]
        return np.asarray([float(np.asarray(v).reshape(-1)[0]) for v in out], dtype=np.float64)
    emd = np.asarray(emd_scores, dtype=np.float64)
    clip = np.asarray(clip_scores, dtype=np.float32).reshape(-1)
    emd_n = (emd - emd.min()) / (EPS + emd.max() - emd.min())
    clip_n = (clip - clip.min()) / (np.float32(EPS) + clip.max() - clip.min())
    return (emd_n + clip_n.astype(np.float64) + pvv + pvt) / 4


def stable_rank(scores: np.ndarray) -> np.ndarray:
    """Descending order that keeps the original order among equal scores.

    Python's ``sorted(key=score, reverse=True)`` (FilteringMergingModule.py:138)
    is stable, SURVEY.md A.1-7.
    """
    return np.argsort(-np.asarray(scores, dtype=np.float64), kind="stable")


def merge_select(ranked_scores: np.ndarray, static_threshold: float, dynamic_threshold: float) -> np.ndarray:
    """Boolean selection over the ranked list (FilteringMergingModule.py:213-217)."""
    top = ranked_scores[0]
    if top < static_threshold:
        return ranked_scores >= dynamic_threshold * top
    return ranked_scores >= static_threshold


def merge_masks(masks: torch.Tensor, selected_indices) -> torch.Tensor:
    """OR of the selected masks as a float32 ``[H,W]`` map (FilteringMergingModule.py:219)."""
    sel = torch.as_tensor(np.asarray(selected_indices), dtype=torch.long)
    return (torch.sum(masks[sel].float(), dim=0) > 0).float()


def clip_scores(img_feats: torch.Tensor, text_feat: torch.Tensor) -> np.ndarray:
    """``img_feats @ text_feats.T`` (FilteringMergingModule.py:97), ``[P]``."""
    return (img_feats @ text_feat.reshape(1, -1).T).reshape(-1).numpy()


# --------------------------------------------------------------------------
# A9 / A10: builder-defined (no reference code; parity unpinned)
# --------------------------------------------------------------------------
def pairwise_intersections(masks: torch.Tensor):
    """``inter[i,j] = sum(m_i & m_j)``, ``area_i = inter[i,i]`` as int32.

    Builder-defined.  Union follows the formula the
    reference's evaluator uses, ``area_a + area_b - inter``
    (mars/utils/evaluation.py:36).  Computed blockwise as an exact integer
    contraction.
    """
    mb = (masks.reshape(masks.shape[0], -1) > 0)
    p, hw = mb.shape
    inter = torch.zeros((p, p), dtype=torch.float64)
    step = 1 << 18
    for s in range(0, hw, step):
        blk = mb[:, s:s + step].float()
        inter += (blk @ blk.T).double()  # exact: each block sum < 2**24
    inter = inter.to(torch.int32)
    area = torch.diagonal(inter).clone()
    return inter, area


def iou_matrix(inter: torch.Tensor, area: torch.Tensor) -> torch.Tensor:
    """float32 IoU with ``union == 0 -> 0``."""
    union = area[:, None] + area[None, :] - inter
    iou = inter.float() / union.float()
    return torch.where(union > 0, iou, torch.zeros_like(iou))


def mask_nms(order: np.ndarray, inter: torch.Tensor, area: torch.Tensor, iou_threshold: float) -> np.ndarray:
    """Greedy score-ordered suppression; returns a keep flag per proposal index.

    Builder-defined, with the semantics of torchvision's ``nms`` as used at
    segment_anything/automatic_mask_generator.py:370-376: walk the proposals in
    rank order, drop one iff its IoU with an already-kept higher-ranked
    proposal is ``> iou_threshold``.
    """
    iou = iou_matrix(inter, area).numpy()
    p = len(order)
    keep = np.zeros(p, dtype=bool)
    kept = []
    for idx in order:
        if all(not (iou[idx, k] > np.float32(iou_threshold)) for k in kept):
            keep[idx] = True
            kept.append(idx)
    return keep


# --------------------------------------------------------------------------
# A12: Matcher-derived scoring and merging
# --------------------------------------------------------------------------
def matcher_patch_matching(ref_feats: torch.Tensor, tar_feat: torch.Tensor, ref_masks_pool: torch.Tensor, g: int,
                           patch_size: int, input_size):
    """Forward / reverse LSAP matching, retain rule, half selection, de-duplication and patch-centre coordinates.

    Follows matcher/Matcher.py:436-547 with scipy's linear_sum_assignment (the reference's solver).  Returns the
    matched and discarded points as sorted lists of (x, y) and the number of points kept by the half rule.
    """
    from scipy.optimize import linear_sum_assignment

    sim = ref_feats @ tar_feat.t()
    mask = ref_masks_pool.flatten().bool()
    s_fwd = sim[mask]
    fr, fc = linear_sum_assignment(s_fwd.numpy(), maximize=True)
    sim_f = s_fwd[fr, fc]
    idx_mask = mask.nonzero()[:, 0]
    s_rev = sim.t()[fc]
    rr, rc = linear_sum_assignment(s_rev.numpy(), maximize=True)
    retain = torch.isin(torch.as_tensor(rc), idx_mask)
    fc_t = torch.as_tensor(fc)
    if not (retain == False).all().item():  # noqa: E712  (mirrors the reference's test)
        pos, neg, sim_pos = fc_t[retain], fc_t[~retain], sim_f[retain]
    else:
        pos, neg, sim_pos = fc_t, fc_t, sim_f
    reduced = len(sim_pos) // 2 if len(sim_pos) > 40 else len(sim_pos)
    order = torch.sort(sim_pos, descending=True)[1][:reduced]

    def centres(idx):
        out = set()
        for p in set(idx.tolist()):
            x, y = (p % g) * patch_size + patch_size // 2, (p // g) * patch_size + patch_size // 2
            if x < input_size[1] and y < input_size[0]:
                out.add((int(x), int(y)))
        return sorted(out)

    return centres(pos[order]), centres(neg), reduced


def matcher_mask_scores(masks: np.ndarray, all_points: np.ndarray, g: int):
    """Purity and coverage of every mask from the matched points.

    ``masks`` is ``[n,H,W]`` bool, ``all_points`` ``[K,2]`` (x, y) pixel
    coordinates.  Follows matcher/Matcher.py:1163-1209: pooled area by
    any-pixel pooling (``cv2.resize(INTER_AREA) > 0`` is the same predicate at
    the native 518/14 geometry, SURVEY.md A.2 probe4), ``purity = in/max(area,1)
    + 1e-6``, ``coverage = in/K + 1e-6`` as float32 tensors.
    """
    n, h, w = masks.shape
    pts = all_points.astype(np.int64)[:, ::-1]
    pts = np.clip(pts, 0, [h - 1, w - 1])
    pooled = pool_mask(torch.from_numpy(masks), g).numpy()
    purity = torch.zeros(n)
    coverage = torch.zeros(n)
    for i in range(n):
        inside = int(masks[i][pts[:, 0], pts[:, 1]].sum())
        area = max(float(pooled[i].sum()), 1.0)
        purity[i] = (torch.tensor([inside / area]) + 1e-6)[0]
        coverage[i] = (torch.tensor([inside / all_points.shape[0]]) + 1e-6)[0]
    return purity, coverage


def matcher_fuse(emd: torch.Tensor, purity: torch.Tensor, coverage: torch.Tensor, alpha: float, beta: float, exp: float):
    """``alpha*emd + beta*purity*coverage**exp`` (matcher/Matcher.py:719-720)."""
    return alpha * emd + beta * purity * coverage ** exp


def matcher_metric_filter(scores, metrics: dict, cfg: dict):
    """Index set surviving the coverage/emd/purity filters (matcher/Matcher.py:732-746)."""
    idx_all = torch.arange(scores.shape[0])
    metrics = dict(metrics)
    for metric in ["coverage", "emd", "purity"]:
        if cfg.get(metric, 0) > 0:
            thres = min(cfg[metric], metrics[metric].max())
            idx = torch.where(metrics[metric] >= thres)[0]
            scores = scores[idx]
            idx_all = idx_all[idx]
            for key in metrics:
                metrics[key] = metrics[key][idx]
    return scores, idx_all


def matcher_merge_topk(scores: torch.Tensor, num_merging_mask: int, topk_scores_threshold: float):
    """Top-k merge selection (matcher/Matcher.py:788-832): indices into ``scores`` and the mean score."""
    topk = min(num_merging_mask, scores.size(0))
    topk_idx = scores.topk(topk)[1]
    topk_scores = scores[topk_idx].numpy()
    if topk_scores_threshold > 0:
        topk_scores = topk_scores / topk_scores.max()
    sel = topk_scores > topk_scores_threshold
    return topk_idx.numpy()[sel], topk_scores[sel].mean()


def matcher_merge_score_filter(scores: torch.Tensor, num_merging_mask: int, score: float, score_norm: float):
    """Score-filter merge selection (matcher/Matcher.py:749-787): indices and the mean score."""
    distances = 1 - scores
    distances, rank = torch.sort(distances, descending=False)
    distances_norm = distances - distances.min()
    distances_norm = distances_norm / (distances.max() + 1e-6)
    keep = distances < score
    keep[..., 0] = True
    keep = keep * (distances_norm < score_norm)
    idx = rank[keep][:num_merging_mask]
    return idx.numpy(), scores[idx].mean()


def _patch_centres(idx, g: int, patch_size: int, input_size):
    """Unique patch indices -> sorted (x, y) patch-centre pixels inside the image (matcher/Matcher.py:519-543)."""
    out = set()
    for p in set(int(i) for i in idx):
        x, y = (p % g) * patch_size + patch_size // 2, (p // g) * patch_size + patch_size // 2
        if x < input_size[1] and y < input_size[0]:
            out.add((x, y))
    return sorted(out)


def matcher_negative_priors(ref_feats: torch.Tensor, tar_feat: torch.Tensor, ref_masks_pool: torch.Tensor, g: int,
                            patch_size: int, input_size, source: str):
    """Negative point priors of Matcher (scipy LSAP, the reference's solver).

    ``source="discarded"``: forward matches whose reverse match left the support mask, least similar half
    (matcher/Matcher.py:304-348).  ``source="cost"``: the bidirectional matching run on ``C = (1 - S) / 2`` with
    ``maximize=True``, keeping pairs whose reverse match lies outside the mask (:350-417) - including the reference's
    indexing of the *unfiltered* forward columns with positions of the filtered, sorted costs (:386-398).
    Returns (sorted (x, y) list or None, reduced_points_num or None).
    """
    from scipy.optimize import linear_sum_assignment

    sim = ref_feats @ tar_feat.t()
    mask = ref_masks_pool.flatten().bool()
    idx_mask = mask.nonzero()[:, 0]
    if source == "discarded":
        s_fwd = sim[mask]
        fr, fc = linear_sum_assignment(s_fwd.numpy(), maximize=True)
        sim_f = s_fwd[fr, fc]
        rr, rc = linear_sum_assignment(sim.t()[fc].numpy(), maximize=True)
        discarded = torch.isin(torch.as_tensor(rc), idx_mask, invert=True)
        if (discarded == False).all().item():  # noqa: E712
            return None, None
        cols, sims = torch.as_tensor(fc)[discarded], sim_f[discarded]
        k = len(sims) // 2 if len(sims) > 40 else len(sims)
        order = torch.sort(sims, descending=False)[1][:k]
        return _patch_centres(cols[order].tolist(), g, patch_size, input_size), k
    cost = (1 - sim) / 2
    fr, fc = linear_sum_assignment(cost.numpy(), maximize=True)
    cost_f = cost[fr, fc]
    rr, rc = linear_sum_assignment(cost.t()[fc].numpy(), maximize=True)
    retain = torch.isin(torch.as_tensor(rc), idx_mask, invert=True)
    cost_kept = cost_f[retain] if not (retain == False).all().item() else cost_f  # noqa: E712
    k = len(cost_kept) // 2 if len(cost_kept) > 40 else len(cost_kept)
    pos = torch.sort(cost_kept, descending=True)[1][:k]
    return _patch_centres(torch.as_tensor(fc)[pos].tolist(), g, patch_size, input_size), k


def matcher_combinations(n: int, k: int):
    """k-subsets of range(n) ordered by largest element, then recursively (matcher/Matcher.py:1212-1224)."""
    if k > n:
        return []
    if k == 0:
        return [[]]
    return [c + [i] for i in range(n) for c in matcher_combinations(i, k - 1)]


def matcher_sample_points(points, sample_range, max_iterations: int, negative_points=None, rng=None):
    """Prompt subsets handed to SAM (matcher/Matcher.py:1226-1296); ``rng`` defaults to the ``random`` module, whose
    calls are made in the reference's order."""
    import random as _random

    rng = rng or _random
    samples, labels = [], []
    n = len(points)
    for size in range(min(sample_range[0], n), min(sample_range[1], n) + 1):
        if n > 8:
            index = [rng.sample(range(n), size) for _ in range(max_iterations)]
        else:
            index = matcher_combinations(n, size)
        pos = np.take(points, index, axis=0)
        samples.append(pos)
        labels.append(np.ones((pos.shape[0], size)))
        if negative_points is not None:
            m = len(negative_points)
            if n > 8 and m > 8:
                index_neg = [rng.sample(range(m), size) for _ in range(max_iterations)]
            else:
                index_neg = [rng.choices(range(m), k=size) for _ in range(len(index))]
            neg = np.take(negative_points, index_neg, axis=0)
            samples.append(neg)
            labels.append(np.zeros((neg.shape[0], size)))
    if negative_points is None:
        return samples, labels
    return ([np.hstack((samples[i], samples[i + 1])) for i in range(0, len(samples), 2)],
            [np.hstack((labels[i], labels[i + 1])) for i in range(0, len(labels), 2)])


def matcher_generate_and_merge(masks: np.ndarray, all_points: np.ndarray, cost: torch.Tensor,
                               ref_masks_pool: torch.Tensor, g: int, alpha: float, beta: float, exp: float, cfg: dict,
                               num_merging_mask: int) -> dict:
    """Scores, filters and merges SAM proposals like Matcher.mask_generation (matcher/Matcher.py:676-834).

    ``masks`` bool ``[n,H,W]``.  EMD per mask on ``cost[support fg][:, pooled mask]`` with the empty-mask rule of
    :1181-1185 (an empty pooled mask is scored against every patch, area = g*g); returns per-mask (purity, coverage,
    emd), the merge order (mask indices), the merged mask and the final score."""
    n = masks.shape[0]
    pooled = pool_mask(torch.from_numpy(masks), g).reshape(n, -1).bool()
    pooled[~pooled.any(dim=1)] = True
    emd = torch.tensor([emd_score(ref_masks_pool, pooled[i], cost) for i in range(n)], dtype=torch.float32)
    purity, coverage = matcher_mask_scores(masks, all_points, g)
    for i in range(n):  # the empty-mask rule also changes the purity denominator
        if not masks[i].any():
            purity[i] = torch.tensor([0.0 / max(float(pooled[i].sum()), 1.0)])[0] + 1e-6
    scores = matcher_fuse(emd, purity, coverage, alpha, beta, exp)
    kept_scores, idx = matcher_metric_filter(scores, dict(purity=purity, coverage=coverage, emd=emd), cfg)
    if cfg["score_filter"]:
        chosen, final = matcher_merge_score_filter(kept_scores, num_merging_mask, cfg["score"], cfg["score_norm"])
    else:
        chosen, final = matcher_merge_topk(kept_scores, num_merging_mask, cfg["topk_scores_threshold"])
    order = idx.numpy()[np.atleast_1d(chosen)]
    merged = masks[order].sum(0) > 0
    return dict(purity=purity, coverage=coverage, emd=emd, scores=scores, order=order, merged=merged, final=float(final))


# --------------------------------------------------------------------------
# 8f-3: evaluator (the step after the path)
# --------------------------------------------------------------------------
def evaluator_areas(pred_mask: torch.Tensor, gt_mask: torch.Tensor, ignore: Optional[torch.Tensor] = None):
    """Per-sample ``[bg, fg]`` intersection and union areas (mars/utils/evaluation.py:12-38)."""
    pred_mask = pred_mask.clone().float()
    gt_mask = gt_mask.clone().float()
    if ignore is not None:
        gt_mask = gt_mask + ignore.float() * 255
        pred_mask[gt_mask == 255] = 255
    inter, pred_a, gt_a = [], [], []
    for p, g_ in zip(pred_mask, gt_mask):
        same = p[p == g_]
        inter.append(torch.histc(same, bins=2, min=0, max=1) if same.numel() else torch.zeros(2))
        pred_a.append(torch.histc(p, bins=2, min=0, max=1))
        gt_a.append(torch.histc(g_, bins=2, min=0, max=1))
    inter = torch.stack(inter).t()
    union = torch.stack(pred_a).t() + torch.stack(gt_a).t() - inter
    return inter, union


def average_meter(area_inter, area_union, class_id, nclass: int, class_ids_interest):
    """AverageMeter.update over all samples then compute_iou (mars/utils/logger.py:61-78), float32 like the reference.

    area_inter / area_union ``[2, n]``; returns ``(intersection_buf, union_buf, miou, fb_iou, iou_fg)``.
    """
    inter_buf = torch.zeros(2, nclass)
    union_buf = torch.zeros(2, nclass)
    for i in range(area_inter.shape[1]):  # one episode per update, like main_MARS.py:72-73
        inter_buf.index_add_(1, class_id[i:i + 1], area_inter[:, i:i + 1].float())
        union_buf.index_add_(1, class_id[i:i + 1], area_union[:, i:i + 1].float())
    interest = torch.as_tensor(class_ids_interest)
    iou = inter_buf / torch.max(torch.stack([union_buf, torch.ones_like(union_buf)]), dim=0)[0]
    iou = iou.index_select(1, interest)
    miou = iou[1].mean() * 100
    fb_iou = (inter_buf.index_select(1, interest).sum(dim=1) / union_buf.index_select(1, interest).sum(dim=1)).mean() * 100
    return inter_buf, union_buf, float(miou), float(fb_iou), iou[1]


# --------------------------------------------------------------------------
# A13: Matcher diagnostics
# --------------------------------------------------------------------------
def ref_to_target_similarity(ref_feats: torch.Tensor, tar_feat: torch.Tensor, ref_masks_pool: torch.Tensor) -> torch.Tensor:
    """Mean over the masked support patches of ``Fq @ Fs_masked^T`` -> ``[N]`` (matcher/Matcher.py:593-611)."""
    masked = ref_feats[ref_masks_pool.reshape(-1).bool()]
    return (tar_feat @ masked.t()).mean(dim=-1)


def aposteriori_statistics(S: torch.Tensor, ref_mask: torch.Tensor, tar_mask: torch.Tensor, unnorm_ref: torch.Tensor,
                           unnorm_tar: torch.Tensor) -> dict:
    """Similarity statistics of the masked sub-matrix and the prototype distance (matcher/Matcher.py:1069-1089)."""
    ref_mask, tar_mask = ref_mask.reshape(-1).bool(), tar_mask.reshape(-1).bool()
    sub = S[ref_mask, :][:, tar_mask]
    ref_proto = unnorm_ref[ref_mask].mean(dim=0)
    tar_proto = unnorm_tar[tar_mask].mean(dim=0)
    return dict(aposteriori_similarity_mean=sub.mean().item(),
                aposteriori_similarity_max=sub.max().item() if sub.numel() > 0 else 0,
                aposteriori_similarity_std=sub.std().item(),
                embeddings_euclidean_distance=torch.norm(ref_proto - tar_proto, p=2).item())


# --------------------------------------------------------------------------
# SAM automatic-mask-generator post-processing
# --------------------------------------------------------------------------
def mask_to_rle(mask: np.ndarray):
    """Uncompressed COCO RLE counts of one ``[H, W]`` boolean mask (segment_anything/utils/amg.py:107-135)."""
    flat = np.asarray(mask, dtype=bool).T.reshape(-1)  # Fortran order
    change = np.nonzero(flat[1:] ^ flat[:-1])[0] + 1
    idx = np.concatenate([[0], change, [flat.size]])
    counts = (idx[1:] - idx[:-1]).tolist()
    return counts if not flat[0] else [0] + counts


def rle_to_mask(counts, h: int, w: int) -> np.ndarray:
    """Inverse (amg.py:138-149)."""
    mask = np.empty(h * w, dtype=bool)
    idx, parity = 0, False
    for c in counts:
        mask[idx:idx + c] = parity
        idx += c
        parity ^= True
    return mask.reshape(w, h).transpose()


def mask_boxes(masks: torch.Tensor) -> torch.Tensor:
    """XYXY boxes, ``[0,0,0,0]`` for empty masks (batched_mask_to_box, amg.py:310-353)."""
    masks = masks > 0
    out = torch.zeros((masks.shape[0], 4), dtype=torch.int64)
    for i, m in enumerate(masks):
        ys, xs = torch.nonzero(m.any(1)).flatten(), torch.nonzero(m.any(0)).flatten()
        if ys.numel():
            out[i] = torch.tensor([xs.min(), ys.min(), xs.max(), ys.max()])
    return out


def stability_score(logits: torch.Tensor, mask_threshold: float, offset: float) -> torch.Tensor:
    """calculate_stability_score (amg.py:156-176)."""
    inter = (logits > (mask_threshold + offset)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    union = (logits > (mask_threshold - offset)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    return inter / union


def box_nms(boxes: torch.Tensor, scores: torch.Tensor, thr: float) -> torch.Tensor:
    """Greedy NMS with torchvision semantics (automatic_mask_generator.py:370-375); ties: lower index first."""
    order = sorted(range(len(scores)), key=lambda i: (-float(scores[i]), i))
    b = boxes.float()
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    keep = []
    for i in order:
        ok = True
        for k in keep:
            iw = torch.minimum(b[i, 2], b[k, 2]) - torch.maximum(b[i, 0], b[k, 0])
            ih = torch.minimum(b[i, 3], b[k, 3]) - torch.maximum(b[i, 1], b[k, 1])
            inter = iw.clamp(min=0) * ih.clamp(min=0)
            if inter / (area[i] + area[k] - inter) > thr:
                ok = False
                break
        if ok:
            keep.append(i)
    return torch.tensor(keep, dtype=torch.int64)


# --------------------------------------------------------------------------
# Whole-episode restatement (what bench.py times as the CPU baseline)
# --------------------------------------------------------------------------
def run_episode(ep: dict, cfg: dict, emd_fn: Optional[Callable] = None) -> dict:
    """The full ranking stage on one episode of already-computed backbone tensors.

    ``ep`` keys: feat_s [ns,N,C], feat_q [N,C], support_mask [ns,H,W],
    attn_vva [N,N], vta_raw [gt,gt], attn_vta [Nt,Nt], masks [P,H,W],
    clip_img [P,D], clip_txt [D], emd [P] (used when ``emd_fn`` is None).
    Order of operations follows mars/MARS.py:62-101.  Returns every
    intermediate the parity tests compare.
    """
    g = cfg["g"]
    fs = normalize_rows(ep["feat_s"])
    fq = normalize_rows(ep["feat_q"])
    sim, cost = similarity_and_cost(fs, fq)
    sup_bits = pool_mask(ep["support_mask"], g).reshape(-1)
    prior = vva_prior(fs, fq, sup_bits, g)
    vva = minmax(pir_refine(prior, ep["attn_vva"], cfg["vva_box_threshold"]))
    vta = pir_refine(ep["vta_raw"], ep["attn_vta"], cfg["vta_box_threshold"])
    vta = minmax(nearest_resize(vta, (g, g)))

    masks = ep["masks"]
    pooled, cov, avv, avt = region_scores(masks, vva.numpy(), vta.numpy(), g)
    if emd_fn is not None:
        emd = np.asarray([emd_fn(sup_bits, torch.from_numpy(pm), cost) for pm in pooled])
    else:
        emd = ep["emd"].numpy().astype(np.float64)
    clip = clip_scores(ep["clip_img"], ep["clip_txt"])
    scores = fuse_scores(emd, clip, cov, avv, avt, cfg["alpha"])
    order = stable_rank(scores)
    out = dict(sim=sim, cost=cost, support_bits=sup_bits, prior=prior, vva=vva, vta=vta,
               pooled=pooled, coverage=cov, pvv_align=avv, pvt_align=avt, clip=clip, emd=emd,
               scores=scores, order=order)
    nms_thr = cfg.get("nms_iou_threshold")
    sel = merge_select(scores[order], cfg["static_threshold"], cfg["dynamic_threshold"])
    if nms_thr is not None:
        inter, area = pairwise_intersections(masks)
        keep = mask_nms(order, inter, area, nms_thr)
        sel = sel & keep[order]
        out.update(inter=inter, area=area, keep=keep)
    out["selected"] = order[sel]
    out["merged"] = merge_masks(masks, order[sel])
    return out
