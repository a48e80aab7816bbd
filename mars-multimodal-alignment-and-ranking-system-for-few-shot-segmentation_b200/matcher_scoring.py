"""Matcher-derived scoring and merging on packed masks.

Mirrors the arithmetic of `RobustPromptSampler.get_mask_scores` (matcher/Matcher.py:1152-1210), the score
fusion (:719-720), the metric filters (:732-746) and the two merge rules (:749-787 score filter, :788-832
top-k) for ALL masks of a target at once: the masks are bit-packed once, purity / coverage come from point
lookups and the pooled bitmap counts, and the selected masks are OR-merged from the packed rows.
LSAP matching, prompt sampling and the SAM calls of Matcher stay outside.
The selection rules act on a few hundred scores and use torch sort / topk on the device, exactly the calls
the reference makes (its tie order is unspecified there too).
"""
from typing import Dict, Optional, Tuple

import torch

from . import ops


class MatcherScorer:
    def __init__(self, encoder_feat_size: int, alpha: float = 1.0, beta: float = 0.0, exp: float = 0.0,
                 num_merging_mask: int = 10, score_filter_cfg: Optional[Dict] = None, device="cuda"):
        self.encoder_feat_size = encoder_feat_size
        self.alpha, self.beta, self.exp = alpha, beta, exp
        self.num_merging_mask = num_merging_mask
        self.score_filter_cfg = dict(emd=0.0, purity=0.0, coverage=0.0, score_filter=False, score=0.33,
                                     score_norm=0.1, topk_scores_threshold=0.0)
        if score_filter_cfg:
            self.score_filter_cfg.update(score_filter_cfg)
        self.device = torch.device(device)

    # ------------------------------------------------------------------ get_mask_scores for all masks
    def mask_scores(self, masks: torch.Tensor, all_points, emd: torch.Tensor):
        """masks [n,H,W] (bool / uint8 / float), all_points [K,2] (x, y), emd [n] (= 1 - emd2, host solver).

        Returns dict(bits, purity, coverage, scores) with the reference's `+ 1e-6` offsets (Matcher.py:1206-1207).
        """
        masks = masks.to(self.device)
        n, h, w = masks.shape
        bits = ops.pack_masks(masks)
        _, _, pooled_count = ops.pool_packed(bits, h, w, self.encoder_feat_size)
        pts = torch.as_tensor(all_points, dtype=torch.int32, device=self.device).reshape(-1, 2)
        inside = ops.points_in_masks(bits, h, w, pts)
        purity, coverage, scores = ops.matcher_scores(inside, pooled_count, emd.to(self.device).float(), pts.shape[0],
                                                      self.alpha, self.beta, self.exp)
        return dict(bits=bits, purity=purity, coverage=coverage, emd=emd.to(self.device).float(), scores=scores,
                    shape=(h, w))

    # ------------------------------------------------------------------ metric filters (Matcher.py:732-746)
    def metric_filter(self, res: dict, record: Optional[dict] = None) -> torch.Tensor:
        idx = torch.arange(res["scores"].numel(), device=self.device)
        metrics = {k: res[k] for k in ("purity", "coverage", "emd")}
        for metric in ("coverage", "emd", "purity"):
            thr = self.score_filter_cfg[metric]
            if thr > 0:
                t = min(thr, float(metrics[metric].max()))
                keep = torch.where(metrics[metric] >= t)[0]
                if record is not None:
                    record[metric] = keep
                idx = idx[keep]
                metrics = {k: v[keep] for k, v in metrics.items()}
        return idx

    # ------------------------------------------------------------------ merge rules
    def select(self, res: dict) -> dict:
        """Which masks get merged: metric filters, then the score-filter (:749-787) or the top-k rule (:788-832).

        Returns dict(idx (survivors of the metric filters), chosen (positions in idx), chosen_global (mask indices, in
        merge order), final_score (numpy float32 for top-k, 0-dim CPU tensor for the score filter, as in the reference),
        metric_filters, before_score_filtering)."""
        filters = {}
        idx = self.metric_filter(res, filters)
        scores = res["scores"][idx]
        cfg = self.score_filter_cfg
        if cfg["score_filter"]:
            distances, rank = torch.sort(1 - scores, descending=False)
            norm = (distances - distances.min()) / (distances.max() + 1e-6)
            keep = distances < cfg["score"]
            keep[..., 0] = True
            keep = keep & (norm < cfg["score_norm"])
            chosen = rank[keep][: self.num_merging_mask]
            final = scores[chosen].mean().cpu()
            before = min(int(scores.numel()), self.num_merging_mask)
        else:
            topk = min(self.num_merging_mask, scores.numel())
            top_idx = scores.topk(topk)[1]
            top_scores = scores[top_idx]
            if cfg["topk_scores_threshold"] > 0:
                top_scores = top_scores / top_scores.max()
            sel = top_scores > cfg["topk_scores_threshold"]
            chosen = top_idx[sel]
            final = top_scores[sel].cpu().numpy().mean()
            before = topk
        return dict(idx=idx, chosen=chosen, chosen_global=idx[chosen], final_score=final, metric_filters=filters,
                    before_score_filtering=before)

    def merge_selected(self, res: dict, chosen_global: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """OR of the packed rows of the chosen masks -> (merged bits [1, wpm], merged float32 [1,H,W])."""
        h, w = res["shape"]
        flags = torch.zeros((1, res["bits"].shape[0]), dtype=torch.uint8, device=self.device)
        flags[0, chosen_global] = 3
        merged_bits, merged = ops.merge_masks(res["bits"][None], flags, h * w, want_bits=True)
        return merged_bits, merged.reshape(1, h, w)

    def merge(self, res: dict) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (merged mask float32 [1,H,W], mean score) like Matcher.mask_generation (:834)."""
        sel = self.select(res)
        _, merged = self.merge_selected(res, sel["chosen_global"])
        final = sel["final_score"]
        return merged, torch.as_tensor(final, device=self.device) if not torch.is_tensor(final) else final.to(self.device)


_SIDE_STREAMS: Dict[torch.device, torch.cuda.Stream] = {}


def _side_stream(device: torch.device) -> torch.cuda.Stream:
    """One persistent side stream per device for the reverse assignment (Matcher builds a PatchMatcher per call)."""
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if device not in _SIDE_STREAMS:
        _SIDE_STREAMS[device] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[device]


class PatchMatcher:
    """Bidirectional patch matching of Matcher on the device.

    Follows `Matcher.patch_level_matching` (matcher/Matcher.py:436-547): S = ref @ tar^T and C = (1 - S) / 2 from the
    tcgen05 contraction, forward assignment of the masked support patches to query patches, reverse assignment of
    the matched query patches to support patches (both exact LSAP, `ops.lsap`), retain the pairs whose reverse match
    lies in the mask, keep the better half when more than 40 survive, de-duplicate and convert to patch-centre pixel
    coordinates.  Prompt sampling / negative priors / SAM stay outside.
    """

    def __init__(self, encoder_feat_size: int, patch_size: int, input_size, device="cuda", concurrent_reverse: bool = True):
        self.encoder_feat_size = encoder_feat_size
        self.patch_size = patch_size
        self.input_size = tuple(input_size)  # (H, W)
        self.device = torch.device(device)
        # With at least as many masked support patches as query patches (the multi-shot case) the forward assignment
        # matches EVERY query patch, so the reverse problem - all query patches against all support patches - does not
        # depend on the forward result: the two single-CTA solvers then run side by side on two streams.
        self.concurrent_reverse = concurrent_reverse

    def match(self, ref_feats: torch.Tensor, tar_feat: torch.Tensor, ref_masks_pool: torch.Tensor):
        """ref_feats [ns*N, C] and tar_feat [N, C] (raw or normalised rows), ref_masks_pool [ns*N] (0/1).

        Returns dict(points [K,2] (x, y) int64, points_discarded [K',2], S, C, reduced_points_num, sim_matched,
        indices_forward (support rows, query patches), retain).
        """
        dev = self.device
        ref = ops.normalize_rows(ref_feats.to(dev).float())
        tar = ops.normalize_rows(tar_feat.to(dev).float())
        m, c = ref_feats.shape
        n = tar_feat.shape[0]
        res = ops.sim_contract(ref, tar, m, n, c, want_sim=True, want_cost=True)
        S, C = res["sim"][0], res["cost"][0]
        mask = ref_masks_pool.to(dev).flatten() != 0
        idx_mask = torch.nonzero(mask).flatten()
        q2s = None
        if self.concurrent_reverse and idx_mask.numel() >= n:
            main, side = torch.cuda.current_stream(), _side_stream(dev)
            St = S.t().contiguous()
            # every buffer of the side launch comes from the main stream's pool (nothing is allocated under `side`)
            rev_status = torch.zeros(1, device=dev, dtype=torch.int32)
            rev_out = (torch.empty((1, n), device=dev, dtype=torch.int32), torch.empty((1,), device=dev, dtype=torch.float64))
            side.wait_stream(main)
            with torch.cuda.stream(side):
                q2s, _ = ops.lsap(St, status=rev_status, out=rev_out)  # no read-back: enqueued, then the forward problem
        # forward: masked support rows -> query patches
        r2c, _ = ops.lsap(S, row_sel=mask.to(torch.uint8))
        fwd_rows = idx_mask[r2c[0][idx_mask] >= 0]
        fwd_cols = r2c[0][fwd_rows].long()
        sim_f = S[fwd_rows, fwd_cols]
        # reverse: matched query patches -> all support rows
        if q2s is not None:
            main.wait_stream(side)
            ops.raise_on_lsap_status(rev_status)
            if fwd_cols.numel() != n:  # cannot happen for an exact solver; never trust it silently
                raise ops.MarsB200Error("forward assignment left query patches unmatched although T >= N")
        else:
            sel = torch.zeros(n, dtype=torch.uint8, device=dev)
            sel[fwd_cols] = 1
            q2s, _ = ops.lsap(S.t().contiguous(), row_sel=sel)
        rev_rows = q2s[0][fwd_cols].long()
        retain = mask[rev_rows.clamp(min=0)] & (rev_rows >= 0)
        if bool(retain.any()):
            pos_cols, neg_cols, sim_pos = fwd_cols[retain], fwd_cols[~retain], sim_f[retain]
        else:  # the reference keeps everything when nothing survives (Matcher.py:483-497)
            pos_cols, neg_cols, sim_pos = fwd_cols, fwd_cols, sim_f
        reduced = len(sim_pos) // 2 if len(sim_pos) > 40 else len(sim_pos)
        order = torch.sort(sim_pos, descending=True)[1][:reduced]
        matched = torch.unique(pos_cols[order])
        unmatched = torch.unique(neg_cols)
        return dict(points=self._centres(matched), points_discarded=self._centres(unmatched), S=S, C=C,
                    reduced_points_num=reduced, sim_matched=sim_pos, indices_forward=(fwd_rows, fwd_cols),
                    retain=retain)

    # ---- diagnostics of the Matcher ancestry
    def get_ref_to_target_similarity(self, ref_feats, tar_feat, ref_masks_pool) -> torch.Tensor:
        """Mean over the masked support patches of `Fq Fs_masked^T` -> [1, N] (matcher/Matcher.py:593-611)."""
        dev = self.device
        ref = ops.normalize_rows(ref_feats.to(dev).float(), normalize=False)
        tar = ops.normalize_rows(tar_feat.to(dev).float(), normalize=False)
        m, c = ref_feats.shape
        n = tar_feat.shape[0]
        S = ops.sim_contract(ref, tar, m, n, c, want_sim=True)["sim"]
        return ops.masked_row_mean(S, ref_masks_pool.to(dev).reshape(1, m) != 0)

    def get_aposteriori_statistics(self, S, ref_masks_pool, target_mask_pooled, unnormalized_ref_feats,
                                   unnormalized_tar_feat) -> dict:
        """Statistics of S[ref mask][:, target mask] and the distance between the two mask-pooled feature
        prototypes (matcher/Matcher.py:1069-1089).  The prototypes are one masks-by-features contraction."""
        dev = self.device
        rm = (ref_masks_pool.to(dev).reshape(1, -1) != 0)
        tm = (target_mask_pooled.to(dev).reshape(1, -1) != 0)
        st = ops.masked_sim_stats(S.to(dev).float(), rm, tm)[0]

        def proto(feats, mask):
            n = feats.shape[0]
            bits = ops.pack_masks(mask.reshape(1, 1, n).to(torch.uint8))  # one "mask" of n patches -> packed words
            return ops.masked_feature_means(bits[:, :(n + 31) // 32].contiguous()[None], feats.to(dev).float()[None])[0, 0]

        pr, pt = proto(unnormalized_ref_feats, rm), proto(unnormalized_tar_feat, tm)
        return dict(aposteriori_similarity_mean=float(st[0]), aposteriori_similarity_max=float(st[1]),
                    aposteriori_similarity_std=float(st[2]),
                    embeddings_euclidean_distance=float(torch.norm(pr - pt, p=2)))

    def _centres(self, patch_idx: torch.Tensor) -> torch.Tensor:
        g, ps = self.encoder_feat_size, self.patch_size
        x = (patch_idx % g) * ps + ps // 2
        y = (patch_idx // g) * ps + ps // 2
        ok = (x < self.input_size[1]) & (y < self.input_size[0])
        return torch.stack([x[ok], y[ok]], dim=1)
