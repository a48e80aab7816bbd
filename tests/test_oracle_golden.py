"""The CPU oracle against the golden vectors produced by the reference's own modules."""
import ast
import os

import numpy as np
import pytest
import torch

import cases
from oracle import mars_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(kind, name):
    z = np.load(os.path.join(GOLD, f"{kind}_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    return z, spec


@pytest.mark.parametrize("name", list(cases.VVA_CASES))
def test_vva_matches_reference(name):
    z, spec = load("vva", name)
    c = cases.vva_inputs(spec)
    assert str(cases.checksum(c)) == str(z["checksum"]), "input generator drifted"
    g = spec["g"]
    fs, fq = orc.normalize_rows(c["feat_s"]), orc.normalize_rows(c["feat_q"])
    sim, cost = orc.similarity_and_cost(fs, fq)
    st = int(z["stride"])
    np.testing.assert_allclose(sim.numpy()[::st, ::st], z["sim"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(cost.numpy()[::st, ::st], z["cost"], rtol=0, atol=2e-6)
    bits = orc.pool_mask(c["support_mask"], g).reshape(-1)
    prior = orc.vva_prior(fs, fq, bits, g)
    a = orc.attention_mean(c["attn_maps"], spec["last_n"], spec["regs"])
    out = orc.minmax(orc.pir_refine(prior, a, spec["thr"]))
    np.testing.assert_allclose(out.numpy(), z["vva_refined"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", list(cases.PIR_CASES))
@pytest.mark.parametrize("use_cv2", [False, True])
def test_pir_matches_reference(name, use_cv2):
    z, spec = load("pir", name)
    c = cases.pir_inputs(spec)
    assert str(cases.checksum(c)) == str(z["checksum"])
    a = orc.attention_mean(c["attn_maps"], spec["last_n"], spec["regs"])
    out = orc.pir_refine(c["prior"], a, spec["thr"], use_cv2=use_cv2)
    np.testing.assert_allclose(out.numpy(), z["refined"], rtol=1e-5, atol=1e-7)
    # the box union the reference built
    b_ref = np.zeros((spec["g"], spec["g"]), dtype=np.float32)
    for x0, y0, x1, y1 in z["boxes"][: int(z["cnt"])]:
        b_ref[y0:y1, x0:x1] = 1
    np.testing.assert_array_equal(orc.box_mask(c["prior"].numpy(), spec["thr"], use_cv2=use_cv2), b_ref)


@pytest.mark.parametrize("name", list(cases.FM_CASES))
def test_filtering_merging_matches_reference(name):
    z, spec = load("fm", name)
    c = cases.fm_inputs(spec)
    assert str(cases.checksum(c)) == str(z["checksum"])
    g = spec["g"]
    pooled, cov, avv, avt = orc.region_scores(c["masks"], c["vva"].numpy(), c["vta"].numpy(), g)
    sup = orc.pool_mask(c["support_mask"], g).reshape(-1)
    emd = np.asarray([orc.emd_score(sup, torch.from_numpy(pm), c["cost"]) for pm in pooled])
    np.testing.assert_allclose(emd, z["emd"], rtol=0, atol=1e-9)
    clip = orc.clip_scores(*cases.alphaclip_as_seen(c))
    scores = orc.fuse_scores(z["emd"], clip, cov, avv, avt, spec["alpha"])
    order = orc.stable_rank(scores)
    np.testing.assert_array_equal(order, z["order"])
    np.testing.assert_allclose(scores[order], z["scores"], rtol=1e-6, atol=1e-7)
    sel = orc.merge_select(scores[order], spec["static"], spec["dynamic"])
    merged = orc.merge_masks(c["masks"], order[sel])
    ref = np.unpackbits(z["merged_bits"])[: merged.numel()].reshape(tuple(z["merged_shape"]))
    np.testing.assert_array_equal(merged.numpy() > 0, ref > 0)


def test_run_episode_consistency():
    """run_episode chains the same pieces (used as the checker for the CUDA episode path)."""
    spec = cases.FM_CASES["g10_dynamic"]
    c = cases.fm_inputs(spec)
    vs = cases.VVA_CASES["g10_2shot"]
    v = cases.vva_inputs(vs)
    g = 10
    ep = dict(feat_s=v["feat_s"][:1], feat_q=v["feat_q"], support_mask=c["support_mask"],
              attn_vva=orc.attention_mean(v["attn_maps"], 2, 4), vta_raw=torch.rand(8, 8),
              attn_vta=torch.softmax(torch.randn(64, 64), -1), masks=c["masks"],
              clip_img=c["clip_img"], clip_txt=c["clip_txt"], emd=torch.rand(spec["P"]))
    cfg = dict(g=g, vva_box_threshold=0.8, vta_box_threshold=0.4, alpha=0.85, static_threshold=0.55,
               dynamic_threshold=0.95, nms_iou_threshold=0.7)
    out = orc.run_episode(ep, cfg)
    assert out["merged"].shape == (140, 140)
    assert out["keep"][out["order"][0]]
    assert set(out["selected"]) <= set(np.nonzero(out["keep"])[0])
    inter, area = out["inter"], out["area"]
    mb = c["masks"].reshape(spec["P"], -1) > 0
    assert torch.equal(area, mb.sum(1).to(torch.int32))
    assert torch.equal(inter, (mb.float() @ mb.float().T).to(torch.int32))


@pytest.mark.parametrize("name", list(cases.EVAL_CASES))
def test_evaluator_and_average_meter_match_reference(name):
    """Evaluator.classify_prediction + AverageMeter of the reference (golden) vs the oracle restatements."""
    z = np.load(os.path.join(GOLD, f"eval_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.eval_inputs(spec)
    inter, union = orc.evaluator_areas(c["pred"], c["gt"], c["ignore"])
    np.testing.assert_array_equal(inter.t().numpy(), z["area_inter"])
    np.testing.assert_array_equal(union.t().numpy(), z["area_union"])
    nclass = z["intersection_buf"].shape[1]
    ids = spec["class_ids"]
    interest = [i - 1 for i in ids] if spec["benchmark"] == "pascal5i" else ids
    ib, ub, miou, fb, cats = orc.average_meter(inter, union, c["class_id"], nclass, interest)
    np.testing.assert_array_equal(ib.numpy(), z["intersection_buf"])
    np.testing.assert_array_equal(ub.numpy(), z["union_buf"])
    np.testing.assert_allclose(miou, float(z["miou"]), rtol=1e-6)
    np.testing.assert_allclose(fb, float(z["fb_iou"]), rtol=1e-6)
    np.testing.assert_allclose(cats.numpy()[:20], z["cats_iou"], rtol=1e-6)


@pytest.mark.parametrize("name", list(cases.AMG_CASES))
def test_amg_postprocessing_matches_reference(name):
    """RLE, boxes, stability score of the reference's amg.py and torchvision's nms (golden) vs the oracle."""
    z = np.load(os.path.join(GOLD, f"amg_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.amg_inputs(spec)
    h, w = spec["H"], spec["W"]
    for i, m in enumerate(c["masks"]):
        counts = orc.mask_to_rle(m.numpy())
        np.testing.assert_array_equal(counts, z["counts"][z["offsets"][i]:z["offsets"][i + 1]])
        np.testing.assert_array_equal(orc.rle_to_mask(counts, h, w), m.numpy())
        assert sum(counts[1::2]) == int(z["areas"][i])
    np.testing.assert_array_equal(orc.mask_boxes(c["masks"]).numpy(), z["boxes"])
    np.testing.assert_array_equal(orc.stability_score(c["logits"], 0.0, 1.0).numpy(), z["stability"])
    np.testing.assert_array_equal(orc.box_nms(c["boxes"], c["scores"], 0.5).numpy(), z["nms_keep"])


@pytest.mark.parametrize("name", list(cases.DIAG_CASES))
def test_matcher_diagnostics_match_reference(name):
    """get_ref_to_target_similarity / get_aposteriori_statistics (the reference's method bodies, golden) vs the oracle."""
    z = np.load(os.path.join(GOLD, f"diag_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.diag_inputs(spec)
    ref_n, tar_n = orc.normalize_rows(c["ref_raw"]), orc.normalize_rows(c["tar_raw"])
    np.testing.assert_allclose(orc.ref_to_target_similarity(ref_n, tar_n, c["ref_mask"]).numpy(), z["ref_to_target"],
                               rtol=1e-6, atol=1e-7)
    st = orc.aposteriori_statistics(ref_n @ tar_n.t(), c["ref_mask"], c["tar_mask"], c["ref_raw"], c["tar_raw"])
    np.testing.assert_allclose([st["aposteriori_similarity_mean"], st["aposteriori_similarity_max"],
                                st["aposteriori_similarity_std"], st["embeddings_euclidean_distance"]], z["stats"],
                               rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", list(cases.MATCHER_CASES))
def test_matcher_pipeline_matches_reference(name):
    """Matcher.set_reference / patch_level_matching / mask_generation / RobustPromptSampler (the reference's method
    bodies, golden) vs the oracle restatement."""
    import random

    z = np.load(os.path.join(GOLD, f"matcher_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.matcher_inputs(spec)
    g, ps, ns = spec["g"], spec["ps"], spec["ns"]
    size = g * ps
    pooled = orc.pool_mask(c["ref_masks"][0], g).reshape(-1).float()
    np.testing.assert_array_equal(pooled.numpy(), z["ref_masks_pool"])
    ref_n, tar_n = orc.normalize_rows(c["ref_raw"]), orc.normalize_rows(c["tar_raw"])
    sim, cost = orc.similarity_and_cost(ref_n, tar_n)
    np.testing.assert_allclose(sim.numpy(), z["sim"], rtol=1e-5, atol=1e-6)
    pts, discarded, reduced = orc.matcher_patch_matching(ref_n, tar_n, pooled, g, ps, (size, size))
    np.testing.assert_array_equal(np.asarray(pts), z["points"])
    assert reduced == int(z["reduced"])
    if spec["neg_discarded"] or spec["neg_cost"]:
        neg, k = orc.matcher_negative_priors(ref_n, tar_n, pooled, g, ps, (size, size),
                                             "discarded" if spec["neg_discarded"] else "cost")
        np.testing.assert_array_equal(np.asarray(neg), z["neg0"])
        assert k == int(z["reduced_neg"][0])
    else:
        np.testing.assert_array_equal(np.asarray(discarded), z["neg0"])
    out = orc.matcher_generate_and_merge(c["proposals"], z["points"], cost, pooled, g, spec["alpha"], spec["beta"],
                                         spec["exp"], spec["cfg"], spec["num_merging_mask"])
    np.testing.assert_allclose(out["purity"].numpy(), z["per_mask"][:, 0], rtol=1e-6)
    np.testing.assert_allclose(out["coverage"].numpy(), z["per_mask"][:, 1], rtol=1e-6)
    np.testing.assert_allclose(out["emd"].numpy(), z["per_mask"][:, 2], rtol=1e-6)
    np.testing.assert_array_equal(c["proposals"][out["order"]], z["masks_to_merge"])
    np.testing.assert_array_equal(out["merged"], z["merged"][0])
    assert abs(out["final"] - float(z["final"])) < 1e-6
    assert len(out["order"]) == int(z["merged_count"])
    # prompt sampling: same `random` stream, same subsets
    np.testing.assert_array_equal(np.asarray(orc.matcher_combinations(5, 3)), z["combos_5_3"])
    random.seed(spec["seed"] + 1)
    demo = np.arange(22).reshape(11, 2)
    s_many, l_many = orc.matcher_sample_points(demo, spec["sample_range"], spec["max_iter"], negative_points=demo[:5] + 100)
    s_few, l_few = orc.matcher_sample_points(demo[:5], spec["sample_range"], spec["max_iter"])
    np.testing.assert_array_equal(np.concatenate([x.reshape(-1) for x in s_many]), z["sample_many"])
    np.testing.assert_array_equal(np.concatenate([x.reshape(-1) for x in l_many]), z["label_many"])
    np.testing.assert_array_equal(np.concatenate([x.reshape(-1) for x in s_few]), z["sample_few"])
    np.testing.assert_array_equal(np.concatenate([x.reshape(-1) for x in l_few]), z["label_few"])


@pytest.mark.parametrize("name", list(cases.MARS_CASES))
def test_mars_predict_end_to_end_matches_reference(name):
    """MARS.predict of the reference (its own MARS / VisualVisualAlignmentModule / FilteringMergingModule classes on
    fake backbones, exact LP for ot.emd2) against the oracle's composition of the same stage, end to end: refined vva,
    resized + min-max vta, the AlphaCLIP text and the merged mask."""
    z = np.load(os.path.join(GOLD, f"mars_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    v = cases.VVA_CASES[spec["vva"]]
    c = cases.mars_inputs(spec)
    g = v["g"]
    assert int(z["g"]) == g
    fs, fq = orc.normalize_rows(c["feat_s"]), orc.normalize_rows(c["feat_q"])
    _, cost = orc.similarity_and_cost(fs.reshape(-1, fs.shape[-1]), fq)
    bits = orc.pool_mask(c["support_mask"], g).reshape(-1)
    prior = orc.vva_prior(fs, fq, bits, g)
    attn = orc.attention_mean(c["attn_maps"], v["last_n"], v["regs"])
    vva = orc.minmax(orc.pir_refine(prior, attn, v["thr"]))
    vta = orc.minmax(orc.nearest_resize(c["vta_raw"], (g, g)))
    np.testing.assert_allclose(vva.numpy(), z["vva"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(vta.numpy(), z["vta"], rtol=1e-6, atol=1e-7)
    text = "a thing." if spec["description"] == "" else f"a thing, {spec['description']}."
    assert str(z["text"]) == text
    pooled, cov, avv, avt = orc.region_scores(c["masks"], vva.numpy(), vta.numpy(), g)
    emd = np.asarray([orc.emd_score(bits, torch.from_numpy(pm), cost) for pm in pooled])
    scores = orc.fuse_scores(emd, orc.clip_scores(c["clip_img"], c["clip_txt"]), cov, avv, avt, spec["alpha"])
    order = orc.stable_rank(scores)
    sel = orc.merge_select(scores[order], spec["static"], spec["dynamic"])
    merged = orc.merge_masks(c["masks"], order[sel])
    np.testing.assert_array_equal(merged.numpy() > 0, z["merged"])


def _emd_by_assignment(cost: np.ndarray) -> float:
    """Independent exact solver for the uniform-marginal transport LP: expand the T x M problem to an L x L assignment
    (L = lcm(T, M); source i appears L/T times, sink j L/M times, every unit carries mass 1/L) and solve it with scipy's
    LSAP.  The expansion is exact: a transport plan with rational entries of denominator L is a sum of L unit matchings."""
    import math

    from scipy.optimize import linear_sum_assignment

    t, m = cost.shape
    big = math.lcm(t, m)
    exp = np.repeat(np.repeat(np.asarray(cost, dtype=np.float64), big // t, axis=0), big // m, axis=1)
    r, c = linear_sum_assignment(exp)
    return float(exp[r, c].sum() / big)


@pytest.mark.parametrize("t,m", [(6, 4), (4, 6), (9, 6), (10, 4), (7, 5), (12, 8), (15, 10), (8, 12), (5, 5), (1, 7), (13, 1)])
def test_exact_emd_rectangular_cross_check(t, m):
    """The oracle's HiGHS transport LP (stand-in for ot.emd2, FilteringMergingModule.py:162-166) against an independent
    exact solver on RECTANGULAR problems (only the square case was cross-checked)."""
    rs = np.random.RandomState(100 * t + m)
    cost = ((1 - rs.uniform(-0.2, 0.9, size=(t, m))) / 2).astype(np.float32)
    assert abs(orc.emd_exact(cost) - _emd_by_assignment(cost)) < 1e-12


@pytest.mark.parametrize("t,m", [(40, 27), (64, 50), (120, 77), (90, 90), (33, 208), (245, 218)])
def test_exact_emd_network_simplex_cross_check(t, m):
    """HiGHS LP against a network simplex (the algorithm class of POT's ot.emd2, FilteringMergingModule.py:162-166; POT
    itself cannot be installed here) at the sizes the c2 episodes produce: T ~ 245-402 support patches, M up to 650."""
    pytest.importorskip("networkx")
    rs = np.random.RandomState(100 * t + m)
    cost = ((1 - rs.uniform(-0.2, 0.9, size=(t, m))) / 2).astype(np.float32)
    assert abs(orc.emd_exact(cost) - orc.emd_network_simplex(cost)) < 1e-12
