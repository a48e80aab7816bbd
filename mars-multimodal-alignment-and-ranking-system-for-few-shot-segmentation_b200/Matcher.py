"""Drop-in `Matcher` and `RobustPromptSampler` (matcher/Matcher.py of the reference) on the marsb200 kernels.

Call signatures, return types and the diagnostic attributes follow the reference class (constructor :33-126,
set_reference :128-185, set_target :187-205, predict :216-249, extract_img_feats :251-302, patch_level_matching
:419-577, mask_generation :619-834, getters :1039-1095, clear :1097-1134; RobustPromptSampler :1140-1296).
The encoder and the SAM generator stay PyTorch input producers.  Everything between them runs on the device:

* S = ref @ tar^T and C = (1 - S) / 2          -> tcgen05 contraction (`ops.sim_contract`)
* forward / reverse assignment (scipy LSAP)     -> `ops.lsap`, S never leaves the device
* per-mask `ot.emd2`, purity, coverage          -> ONE batch over all generated masks: `ops.pack_masks`, `ops.pool_packed`,
                                                   `ops.emd_scores`, `ops.points_in_masks`, `ops.matcher_scores`
* metric filters, top-k / score-filter merge    -> `MatcherScorer` (OR of packed rows)

Deliberate differences, all outside the arithmetic: the matched points come back sorted by patch index (the reference
iterates a Python `set`), LSAP ties may resolve differently than scipy (same objective), k-means clustering of the
points (`use_points_or_centers=False`, matcher/k_means.py) is not part of this stage and must be supplied as
`clustering_fn`.
"""
from __future__ import annotations

import random
from itertools import combinations as _combinations
from typing import Callable, Optional

import numpy as np
import torch

from . import ops
from .matcher_scoring import MatcherScorer, PatchMatcher


class RobustPromptSampler:
    """matcher/Matcher.py:1140-1296."""

    def __init__(self, encoder_feat_size, sample_range, max_iterations, device="cuda"):
        self.encoder_feat_size = encoder_feat_size
        self.sample_range = sample_range
        self.max_iterations = max_iterations
        self.device = torch.device(device)

    # ------------------------------------------------------------------ scores of generated masks
    def batch_mask_scores(self, masks, all_points, emd_cost: torch.Tensor, ref_masks_pool: torch.Tensor,
                          alpha: float = 1.0, beta: float = 0.0, exp: float = 0.0) -> dict:
        """All masks of a target at once: masks [n,H,W] (bool / uint8 / float, numpy or torch).

        Returns dict(bits, purity [n], coverage [n], emd [n] (= 1 - emd2), scores [n] = alpha*emd + beta*purity*coverage^exp
        (:719-720), pooled_count, shape); the reference computes the three metrics one mask at a time (:1176-1210).  A mask whose pooled bitmap is empty is scored against all
        patches, as the reference's `thres = masks.max() - 1e-6` branch does (:1181-1185).
        """
        dev = self.device
        masks = torch.as_tensor(np.ascontiguousarray(masks) if isinstance(masks, np.ndarray) else masks)
        if masks.dtype == torch.bool:
            masks = masks.to(torch.uint8)
        masks = masks.to(dev)
        n, h, w = masks.shape
        g = self.encoder_feat_size
        bits = ops.pack_masks(masks)
        pooled, _, cnt = ops.pool_packed(bits, h, w, g)
        empty = cnt == 0
        if bool(empty.any()):  # every patch becomes foreground (all g*g bits set, tail bits of the last word clear)
            full = torch.full((pooled.shape[-1],), -1, dtype=torch.int32, device=dev)
            tail = g * g - 32 * (pooled.shape[-1] - 1)
            if tail < 32:
                full[-1] = (1 << tail) - 1
            pooled = torch.where(empty[:, None], full[None, :], pooled)
            cnt = torch.where(empty, torch.full_like(cnt, g * g), cnt)
        row_fg = (ref_masks_pool.to(dev).flatten() != 0).to(torch.uint8)
        emd = ops.emd_scores(emd_cost.to(dev).float()[None], row_fg[None], pooled[None].contiguous(),
                             pooled_count=cnt)[0].float()
        pts = torch.as_tensor(np.asarray(all_points), dtype=torch.int32, device=dev).reshape(-1, 2)
        inside = ops.points_in_masks(bits, h, w, pts)
        purity, coverage, scores = ops.matcher_scores(inside, cnt, emd, pts.shape[0], alpha, beta, exp)
        return dict(bits=bits, purity=purity, coverage=coverage, emd=emd, scores=scores, pooled_count=cnt, shape=(h, w))

    def get_mask_scores(self, points, masks, all_points, emd_cost, ref_masks_pool):
        """One mask `[1,H,W]` (numpy bool), same 6-tuple as the reference (:1152-1210)."""
        assert all_points is not None
        res = self.batch_mask_scores(np.asarray(masks)[:1], all_points, emd_cost, ref_masks_pool)
        labels = np.ones((np.asarray(points).shape[0],))
        return res["purity"].cpu(), res["coverage"].cpu(), float(res["emd"][0]), points, labels, masks

    # ------------------------------------------------------------------ prompt sampling (host, no arithmetic)
    def combinations(self, n, k):
        """k-subsets of range(n) in the reference's order: increasing largest element, recursively (:1212-1224)."""
        return sorted((list(c) for c in _combinations(range(n), k)), key=lambda c: c[::-1]) if k <= n else []

    def sample_points(self, points, negative_points=None):
        """Prompt subsets of sizes sample_range[0]..sample_range[1]: random draws when there are more than 8 points,
        every combination otherwise; with negative points each positive set is paired with as many negatives (:1226-1296).
        Consumes `random` in the reference's call order."""
        sample_list, label_list = [], []
        n = len(points)
        lo, hi = min(self.sample_range[0], n), min(self.sample_range[1], n)
        for size in range(lo, hi + 1):
            if n > 8:
                index = [random.sample(range(n), size) for _ in range(self.max_iterations)]
            else:
                index = self.combinations(n, size)
            sample = np.take(points, index, axis=0)
            sample_list.append(sample)
            label_list.append(np.ones((sample.shape[0], size)))
            if negative_points is not None:
                m = len(negative_points)
                if n > 8 and m > 8:
                    index_neg = [random.sample(range(m), size) for _ in range(self.max_iterations)]
                else:
                    index_neg = [random.choices(range(m), k=size) for _ in range(len(index))]
                sample_neg = np.take(negative_points, index_neg, axis=0)
                assert sample.shape[0] == sample_neg.shape[0]
                sample_list.append(sample_neg)
                label_list.append(np.zeros((sample_neg.shape[0], size)))
        if negative_points is None:
            return sample_list, label_list
        pts = [np.hstack((sample_list[i], sample_list[i + 1])) for i in range(0, len(sample_list), 2)]
        lbl = [np.hstack((label_list[i], label_list[i + 1])) for i in range(0, len(label_list), 2)]
        return pts, lbl


class Matcher:
    """matcher/Matcher.py:32-1138 (visualisation and logging helpers excluded)."""

    def __init__(self, encoder, encoder_transforms, use_encoder_registers=False, generator=None, input_size=518,
                 num_centers=8, use_box=False, use_points_or_centers=True, sample_range=(4, 6),
                 max_sample_iterations=30, alpha=1., beta=0., exp=0., score_filter_cfg=None, num_merging_mask=10,
                 use_negative_priors_from_discarded=False, use_negative_priors_from_cost=False,
                 merge_prompt_types=False, visualize=False, device=None,
                 clustering_fn: Optional[Callable] = None):
        if device is None:
            device = torch.device("cuda:0")
        if torch.device(device).type != "cuda":
            raise RuntimeError("marsb200.Matcher needs a CUDA device (there is no CPU path)")
        if visualize:
            raise NotImplementedError("visualize_internal_state is plotting, not part of the ranking stage")
        self.encoder = encoder
        self.generator = generator
        self.rps = None
        self.input_size = input_size if isinstance(input_size, tuple) else (input_size, input_size)
        self.encoder_transform = encoder_transforms
        self.use_encoder_registers = use_encoder_registers
        self.num_centers = num_centers
        self.use_box = use_box
        self.use_points_or_centers = use_points_or_centers
        self.sample_range = sample_range
        self.max_sample_iterations = max_sample_iterations
        self.alpha, self.beta, self.exp = alpha, beta, exp
        assert score_filter_cfg is not None
        self.score_filter_cfg = score_filter_cfg
        self.num_merging_mask = num_merging_mask
        self.use_negative_priors_from_discarded = use_negative_priors_from_discarded
        self.use_negative_priors_from_cost = use_negative_priors_from_cost
        self.merge_prompt_types = merge_prompt_types
        self.visualize = False
        self.visualization_parameters = None
        self.logger = None
        self.device = torch.device(device)
        self.clustering_fn = clustering_fn
        self._reset_state()

    def _reset_state(self):
        self.tar_img = self.tar_img_np = None
        self.ref_imgs = self.ref_masks = self.ref_masks_pool = self.nshot = None
        self.encoder_img_size = self.encoder_feat_size = None
        self.unnormalized_ref_feats = self.unnormalized_tar_feat = None
        self.stored_ref_feats = self.stored_tar_feat = None
        self.S = self.S_forward = self.S_reverse = None
        self.sim_scores_after_forward_matching = self.sim_scores_after_backward_matching = None
        self.sim_discarded_patches = None
        self.number_support_patches_forward_matching = self.number_query_patches_forward_matching = None
        self.number_support_patches_backward_matching = self.number_query_patches_backward_matching = None
        self.number_of_merged_masks = self.number_of_masks_before_score_filtering = None
        self.number_of_points_used_for_prediction = self.number_of_points_usable_for_prediction = None
        self.positive_points_inside_mask = self.negative_points_inside_mask = None
        self.masks_to_merge = self.unfiltered_generated_masks = None
        self.metric_filters = {}

    # ------------------------------------------------------------------ inputs
    def _mask_threshold(self) -> float:
        try:
            return float(self.generator.predictor.model.mask_threshold)
        except AttributeError:
            return 0.0  # SAM's class constant

    def _pool_to_patches(self, masks: torch.Tensor) -> torch.Tensor:
        """avg_pool2d(mask, patch) > mask_threshold (:172-179).  With the threshold 0 of SAM and non-negative masks this is
        "any pixel of the patch set": the adaptive max pool kernel (row A3)."""
        ps = self.encoder.patch_size
        n, _, h, w = masks.shape
        if self._mask_threshold() == 0.0 and h % ps == 0 and w % ps == 0 and h == w and bool((masks >= 0).all()):
            return ops.pool_mask(masks.reshape(n, h, w).to(self.device).float(), h // ps).reshape(-1).float()
        import torch.nn.functional as F

        return (F.avg_pool2d(masks.to(self.device).float(), (ps, ps)) > self._mask_threshold()).float().reshape(-1)

    def set_reference(self, imgs, masks):
        """imgs [1,ns,3,h,w], masks [1,ns,h,w] (:128-185); an all-zero mask set gets the 14x14 centre square."""
        if masks.sum() == 0:
            _, _, sh, sw = masks.shape
            masks[..., (sh // 2 - 7):(sh // 2 + 7), (sw // 2 - 7):(sw // 2 + 7)] = 1
        imgs = imgs.flatten(0, 1)
        img_size = imgs.shape[-1]
        assert img_size == self.input_size[-1]
        self.encoder_img_size = img_size
        self.encoder_feat_size = img_size // self.encoder.patch_size
        masks = masks.permute(1, 0, 2, 3)  # ns, 1, h, w
        self.ref_masks_pool = self._pool_to_patches(masks)
        self.nshot = masks.shape[0]
        self.ref_imgs = imgs
        self.ref_masks = masks

    def set_target(self, img):
        img_h, img_w = img.shape[-2:]
        assert img_h == self.input_size[0] and img_w == self.input_size[1]
        self.tar_img = img
        self.tar_img_np = img.mul(255).byte().squeeze(0).permute(1, 2, 0).cpu().numpy()

    def set_rps(self):
        if self.rps is None:
            assert self.encoder_feat_size is not None
            self.rps = RobustPromptSampler(encoder_feat_size=self.encoder_feat_size, sample_range=self.sample_range,
                                           max_iterations=self.max_sample_iterations, device=self.device)

    def set_logger(self, logger):
        self.logger = logger

    # ------------------------------------------------------------------ the stage
    def predict(self, target_mask=None):
        ref_feats, tar_feat = self.extract_img_feats()
        all_points, negative_points, box, S, C, _, _ = self.patch_level_matching(ref_feats=ref_feats, tar_feat=tar_feat)
        if self.use_points_or_centers:
            points = all_points
        else:
            points = self.clustering(all_points)
        self.set_rps()
        return self.mask_generation(self.tar_img_np, points, box, all_points, self.ref_masks_pool, C, negative_points,
                                    target_mask=target_mask)

    def extract_img_feats(self):
        """Encoder forward (PyTorch producer) and row L2-normalisation on the device (:251-302)."""
        if self.stored_ref_feats is not None and self.stored_tar_feat is not None:
            return self.stored_ref_feats, self.stored_tar_feat
        ref_imgs = torch.cat([self.encoder_transform(r)[None, ...] for r in self.ref_imgs], dim=0).to(self.device)
        tar_img = torch.cat([self.encoder_transform(t)[None, ...] for t in self.tar_img], dim=0).to(self.device)
        family = getattr(self.encoder, "family", "vits")
        with torch.no_grad():
            if family.startswith("vits"):
                skip = 1 + (self.encoder.num_register_tokens if self.use_encoder_registers else 0)
                ref_feats = self.encoder.forward_features(ref_imgs)["x_prenorm"][:, skip:]
                tar_feat = self.encoder.forward_features(tar_img)["x_prenorm"][:, skip:]
            else:
                ref_feats, tar_feat = self.encoder(ref_imgs), self.encoder(tar_img)
        c = self.encoder.embed_dim
        ref_feats = ref_feats.reshape(-1, c).float().contiguous()
        tar_feat = tar_feat.reshape(-1, c).float().contiguous()
        self.unnormalized_ref_feats, self.unnormalized_tar_feat = ref_feats.clone(), tar_feat.clone()
        self.stored_ref_feats = ops.normalize_rows(ref_feats)[0][0, :ref_feats.shape[0], :c].contiguous()
        self.stored_tar_feat = ops.normalize_rows(tar_feat)[0][0, :tar_feat.shape[0], :c].contiguous()
        return self.stored_ref_feats, self.stored_tar_feat

    def _points(self, patch_idx: torch.Tensor) -> np.ndarray:
        """Unique patch indices -> patch-centre (x, y) pixel coordinates inside the image (:519-543)."""
        g, ps = self.encoder_feat_size, self.encoder.patch_size
        idx = torch.unique(patch_idx)
        x = (idx % g) * ps + ps // 2
        y = (idx // g) * ps + ps // 2
        ok = (x < self.input_size[1]) & (y < self.input_size[0])
        return torch.stack([x[ok], y[ok]], dim=1).cpu().numpy().astype(np.int64)

    def patch_level_matching(self, ref_feats, tar_feat):
        """-> (points, negative_priors | points_discarded, box, S, C, reduced_points_num, reduced_points_num_neg) (:419-577)."""
        pm = PatchMatcher(self.encoder_feat_size, self.encoder.patch_size, self.input_size, self.device)
        res = pm.match(ref_feats, tar_feat, self.ref_masks_pool)
        self._match = res
        self.S, C = res["S"], res["C"]
        fwd_rows, fwd_cols = res["indices_forward"]
        retain = res["retain"]
        mask = self.ref_masks_pool.to(self.device).flatten() != 0
        self.S_forward = self.S[mask]
        self.S_reverse = self.S.t()[fwd_cols]
        sim_f = self.S[fwd_rows, fwd_cols]
        self.sim_scores_after_forward_matching = sim_f
        self.number_support_patches_forward_matching = int(fwd_rows.numel())
        self.number_query_patches_forward_matching = int(fwd_cols.numel())
        any_kept = bool(retain.any())
        kept = int(retain.sum()) if any_kept else int(fwd_rows.numel())
        self.number_support_patches_backward_matching = kept
        self.number_query_patches_backward_matching = kept
        self.sim_scores_after_backward_matching = res["sim_matched"]
        self.sim_discarded_patches = sim_f[~retain] if any_kept else sim_f.clone()
        points = res["points"].cpu().numpy().astype(np.int64)
        points_discarded = res["points_discarded"].cpu().numpy().astype(np.int64)

        negative_priors, reduced_points_num_neg = [], []
        if self.use_negative_priors_from_discarded:
            neg, k = self.sample_negative_points_from_discarded((fwd_rows, fwd_cols), sim_f, retain)
            negative_priors.append(neg)
            reduced_points_num_neg.append(k)
        if self.use_negative_priors_from_cost:
            neg, k = self.sample_negative_points_from_cost(C)
            negative_priors.append(neg)
            reduced_points_num_neg.append(k)
        if self.use_box:
            box = np.array([max(points[:, 0].min(), 0), max(points[:, 1].min(), 0),
                            min(points[:, 0].max(), self.input_size[1] - 1),
                            min(points[:, 1].max(), self.input_size[0] - 1)])
        else:
            box = None
        return (points, negative_priors if len(negative_priors) > 0 else points_discarded, box, self.S, C,
                res["reduced_points_num"], reduced_points_num_neg)

    def sample_negative_points_from_discarded(self, idxs_forward, sim_scores_forward, retain):
        """Forward matches whose reverse match left the support mask, least similar half first (:304-348).
        `retain` is the reverse-match-inside-mask flag per forward pair (the reference recomputes it from the reverse
        assignment and the mask indices)."""
        discarded = ~retain
        if not bool(discarded.any()):
            return None, None
        cols, sims = idxs_forward[1][discarded], sim_scores_forward[discarded]
        k = len(sims) // 2 if len(sims) > 40 else len(sims)
        order = torch.sort(sims, descending=False)[1][:k]
        return self._points(cols[order]), k

    def sample_negative_points_from_cost(self, C):
        """The same bidirectional matching run on the cost matrix (most dissimilar pairs), keeping query patches whose
        reverse match lies OUTSIDE the support mask (:350-417)."""
        dev = self.device
        m, n = C.shape
        mask = self.ref_masks_pool.to(dev).flatten() != 0
        if m >= n:  # assignment of min(m, n) pairs: solve with the shorter side as rows
            q2s, _ = ops.lsap(C.t().contiguous())
            cols = torch.arange(n, device=dev)
            rows = q2s[0][:n].long()
            order = torch.argsort(rows)  # scipy returns the pairs sorted by row of C
            rows, cols = rows[order], cols[order]
        else:
            r2c, _ = ops.lsap(C)
            rows = torch.arange(m, device=dev)
            cols = r2c[0][:m].long()
        cost_f = C[rows, cols]
        sel = torch.zeros(n, dtype=torch.uint8, device=dev)
        sel[cols] = 1
        q2s, _ = ops.lsap(C.t().contiguous(), row_sel=sel)
        rev = q2s[0][cols].long()
        retain = ~mask[rev.clamp(min=0)]
        cost_kept = cost_f[retain] if bool(retain.any()) else cost_f
        k = len(cost_kept) // 2 if len(cost_kept) > 40 else len(cost_kept)
        # the reference indexes the UNFILTERED forward columns with positions of the filtered, sorted costs (:386-398)
        pos = torch.sort(cost_kept, descending=True)[1][:k]
        return self._points(cols[pos]), k

    def clustering(self, points):
        if self.clustering_fn is None:
            raise NotImplementedError("k-means++ prompt clustering (matcher/k_means.py) is outside the ranking stage: "
                                      "pass clustering_fn or keep use_points_or_centers=True")
        return np.array(self.clustering_fn(points, min(self.num_centers, len(points)))).astype(np.int64)

    def mask_generation(self, tar_img_np, points, box, all_ponits, ref_masks_pool, C, negative_points=None,
                        target_mask: torch.Tensor = None):
        """Prompts -> SAM proposals (producer) -> batch scoring, filtering and merging on the device (:619-834).
        Returns (merged mask float32 [1,H,W] on the device, final score)."""
        samples_list, label_list = [], []
        if self.use_negative_priors_from_discarded or self.use_negative_priors_from_cost:
            for neg in negative_points:
                if neg is not None and len(neg) > 0:
                    s, l = self.rps.sample_points(points, negative_points=neg)
                else:
                    s, l = self.rps.sample_points(points)
                samples_list.extend(s)
                label_list.extend(l)
            if self.merge_prompt_types:
                s, l = self.rps.sample_points(points)
                samples_list.extend(s)
                label_list.extend(l)
        else:
            samples_list, label_list = self.rps.sample_points(points)

        proposals = self.generator.generate(
            tar_img_np, select_point_coords=samples_list, select_point_labels=label_list,
            select_box=[box] if self.use_box else None,
            select_mask_input=target_mask.cpu().numpy() if target_mask is not None else None)
        masks = torch.stack([torch.as_tensor(np.asarray(q["segmentation"])) for q in proposals]).to(self.device)
        masks = (masks > 0).to(torch.uint8)
        point_coords = [q["point_coords"] for q in proposals]

        res = self.rps.batch_mask_scores(masks, all_ponits, C, ref_masks_pool, self.alpha, self.beta, self.exp)
        scorer = MatcherScorer(self.encoder_feat_size, self.alpha, self.beta, self.exp, self.num_merging_mask,
                               self.score_filter_cfg, self.device)
        self.unfiltered_generated_masks = masks.float()
        sel = scorer.select(res)
        self.metric_filters = sel["metric_filters"]
        chosen = sel["chosen_global"]
        self.number_of_masks_before_score_filtering = sel["before_score_filtering"]
        self.number_of_merged_masks = int(chosen.numel())
        self.masks_to_merge = masks[chosen].float()
        used = set(tuple(p) for i in chosen.tolist() for p in point_coords[i])
        self.number_of_points_used_for_prediction = len(used)
        self.number_of_points_usable_for_prediction = len(all_ponits)
        merged_bits, merged = scorer.merge_selected(res, chosen)

        def inside(pts):
            pts = np.asarray(pts).reshape(-1, 2) if pts is not None and len(pts) > 0 else np.zeros((0, 2))
            if pts.shape[0] == 0:
                return 0
            t = torch.as_tensor(pts, dtype=torch.int32, device=self.device)
            return int(ops.points_in_masks(merged_bits, res["shape"][0], res["shape"][1], t)[0])

        self.positive_points_inside_mask = inside(all_ponits)
        if isinstance(negative_points, list):  # one array per negative-prior source
            self.negative_points_inside_mask = sum(inside(p) for p in negative_points)
        else:
            self.negative_points_inside_mask = inside(negative_points)
        return merged, sel["final_score"]

    # ------------------------------------------------------------------ getters (:1039-1095)
    def get_ref_to_target_similarity(self, ref_feats, tar_feat, ref_masks_pool):
        pm = PatchMatcher(self.encoder_feat_size, self.encoder.patch_size, self.input_size, self.device)
        return pm.get_ref_to_target_similarity(ref_feats, tar_feat, ref_masks_pool)

    def get_negative_point_priors(self, similarity):
        pass

    def get_purity_filter(self):
        return self.metric_filters["purity"]

    def get_similarities(self):
        return (self.S_forward, self.S_reverse, self.sim_scores_after_forward_matching,
                self.sim_scores_after_backward_matching, self.sim_discarded_patches)

    def get_patch_matching_statistics(self) -> dict:
        f_s, b_s = self.number_support_patches_forward_matching, self.number_support_patches_backward_matching
        f_q, b_q = self.number_query_patches_forward_matching, self.number_query_patches_backward_matching
        return {"number_support_patches_forward_matching": f_s, "number_support_patches_backward_matching": b_s,
                "number_discarded_support_patches": f_s - b_s, "number_query_patches_forward_matching": f_q,
                "number_query_patches_backward_matching": b_q, "number_discarded_query_patches": f_q - b_q}

    def get_mask_generation_statistics(self) -> dict:
        usable, used = self.number_of_points_usable_for_prediction, self.number_of_points_used_for_prediction
        pos, neg = self.positive_points_inside_mask, self.negative_points_inside_mask
        return {"number_of_merged_masks": self.number_of_merged_masks,
                "number_of_masks_before_score_filtering": self.number_of_masks_before_score_filtering,
                "ratio_of_merged_masks": self.number_of_merged_masks / self.number_of_masks_before_score_filtering,
                "number_of_points_usable_for_prediction": usable, "number_of_points_used_for_prediction": used,
                "ratio_points_used_vs_usable": used / usable, "positive_points_inside_mask": pos,
                "negative_points_inside_mask": neg,
                "ratio_negative_vs_positive_points_inside_mask": neg / max(1, pos),
                "ratio_positive_points_inside_mask_vs_usable_points": pos / usable}

    def get_aposteriori_statistics(self, mask: torch.Tensor):
        pm = PatchMatcher(self.encoder_feat_size, self.encoder.patch_size, self.input_size, self.device)
        pooled = self._pool_to_patches(mask.reshape(1, 1, *mask.shape[-2:]))
        return pm.get_aposteriori_statistics(self.S, self.ref_masks_pool, pooled, self.unnormalized_ref_feats,
                                             self.unnormalized_tar_feat)

    def get_masks_to_merge(self):
        return self.masks_to_merge

    def get_unfiltered_generated_masks(self):
        return self.unfiltered_generated_masks

    def clear(self):
        self._reset_state()
        if self.generator is not None and hasattr(self.generator, "reset_stored_features"):
            self.generator.reset_stored_features()
