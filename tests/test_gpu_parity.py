"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden vectors.

Bars (BASELINE.json north_star): integer work (packed bits, pooled bitmaps, areas, intersections,
NMS keep-sets, merged masks) bit-exact; float scores within 1e-4 relative; ranking order identical
except for ties inside that tolerance.
"""
import ast
import os

import numpy as np
import pytest
import torch

import cases
from oracle import compare
from oracle import mars_oracle as orc

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-4  # the north-star tolerance for floating-point scores


@pytest.fixture(scope="module")
def mb():
    import marsb200

    return marsb200


def dev():
    return torch.device("cuda:0")


def np_pack(masks: np.ndarray, wpm: int) -> np.ndarray:
    """Reference packing: bit k of word w = pixel 32*w + k, zero padded to wpm words."""
    n = masks.shape[0]
    flat = (masks.reshape(n, -1) > 0)
    padded = np.zeros((n, wpm * 32), dtype=bool)
    padded[:, : flat.shape[1]] = flat
    return np.packbits(padded.reshape(n, wpm, 32), axis=-1, bitorder="little").view(np.uint32).reshape(n, wpm)


def unpack_pooled(pooled: torch.Tensor, n: int) -> np.ndarray:
    p = pooled.cpu().numpy().view(np.uint32)
    bits = np.unpackbits(p.view(np.uint8).reshape(p.shape[0], -1), axis=-1, bitorder="little")
    return bits[:, :n].astype(bool)


def assert_order_matches(order_gpu, scores_gpu, order_ref, scores_ref, rtol=RTOL):
    """Identical order except where the swapped proposals' reference scores tie within the tolerance."""
    order_gpu, order_ref = np.asarray(order_gpu), np.asarray(order_ref)
    np.testing.assert_allclose(scores_gpu, scores_ref, rtol=rtol, atol=1e-6)
    for r in np.nonzero(order_gpu != order_ref)[0]:
        a, b = scores_ref[order_gpu[r]], scores_ref[order_ref[r]]
        assert abs(a - b) <= rtol * max(abs(a), abs(b)) + 1e-9, f"rank {r}: {order_gpu[r]} vs {order_ref[r]}"


# ------------------------------------------------------------------------------------------ ingest
@pytest.mark.parametrize("shape,dtype", [
    ((6, 518, 518), torch.float32),   # vector path (HW % 4 == 0)
    ((6, 518, 518), torch.uint8),     # scalar path (HW % 16 != 0)
    ((3, 1024, 1024), torch.float32),
    ((3, 1024, 1024), torch.uint8),   # vector path
    ((3, 1024, 1024), torch.bool),
    ((5, 37, 41), torch.float32),     # odd size -> scalar path
    ((1, 1, 1), torch.float32),
])
def test_pack_masks_bit_exact(mb, shape, dtype):
    n, h, w = shape
    g = torch.Generator().manual_seed(5)
    m = (torch.rand(shape, generator=g) < 0.3)
    m[0] = False            # empty mask
    if n > 1:
        m[1] = True         # full mask
    bits = mb.ops.pack_masks(m.to(dtype).to(dev()))
    wpm = mb.ops.words_per_mask(h * w)
    assert bits.shape == (n, wpm)
    np.testing.assert_array_equal(bits.cpu().numpy().view(np.uint32), np_pack(m.numpy(), wpm))


@pytest.mark.parametrize("h,w,g,dtype", [
    (1024, 1024, 37, torch.float32),   # c2 / c4 geometry: overlapping 28/29-pixel bins
    (1024, 1024, 37, torch.uint8),
    (96, 64, 5, torch.float32),        # rows shorter than a warp's 1024 pixels: a warp covers many rows
    (100, 128, 7, torch.uint8),
    (33, 32, 4, torch.float32),        # one word per row
    (518, 544, 37, torch.float32),     # W % 32 == 0 but H*W not a multiple of the block size
    (518, 518, 37, torch.float32),     # W % 32 != 0: the call falls back to the two kernels
    (140, 140, 10, torch.uint8),
])
def test_fused_pack_pool_equals_pack_then_pool(mb, h, w, g, dtype):
    """ops.pack_pool (one pass: packed bits + pooled bitmaps from the words in registers) against pack_masks followed by
    pool_packed, bit for bit, and the pooled bitmaps against adaptive_max_pool2d (FilteringMergingModule.py:103-107)."""
    n = 9
    masks = cases.blob_masks(n, h, w, seed=h + w + g, min_frac=0.002, max_frac=0.3)
    masks[0] = 0                       # empty mask
    masks[1] = 0
    masks[1, h - 1, w - 1] = 1         # last pixel only
    masks[2] = 1                       # full mask
    x = masks.to(dev()).to(dtype)
    bits_ref = mb.ops.pack_masks(x)
    pooled_ref, area_ref, cnt_ref = mb.ops.pool_packed(bits_ref, h, w, g)
    bits, (pooled, area, cnt) = mb.ops.pack_pool(x, g)
    assert torch.equal(bits, bits_ref) and torch.equal(pooled, pooled_ref)
    assert torch.equal(area, area_ref) and torch.equal(cnt, cnt_ref)
    want = torch.nn.functional.adaptive_max_pool2d(masks[:, None], (g, g)).flatten(1) > 0
    np.testing.assert_array_equal(unpack_pooled(pooled, g * g), want.numpy())
    # a second call into the same output buffers (the engine reuses them every step) gives the same result
    bits2, (pooled2, area2, cnt2) = mb.ops.pack_pool(x, g, out_bits=bits, out_pool=(pooled, area, cnt))
    assert torch.equal(pooled2, pooled_ref) and torch.equal(area2, area_ref) and torch.equal(cnt2, cnt_ref)


def test_pack_nonbinary_values(mb):
    """Pixels are set iff value > 0 (the reference pools and thresholds with `> 0`)."""
    m = torch.tensor([[[0.0, 0.5, -1.0, 2.0, 1e-30, -0.0, 255.0, 0.0]]])
    bits = mb.ops.pack_masks(m.to(dev()))
    assert int(bits[0, 0].item()) & 0xFF == 0b01011010


@pytest.mark.parametrize("h,g", [(100, 7), (140, 10), (518, 37), (1024, 37), (64, 64)])
def test_pooling_matches_adaptive_max_pool(mb, h, g):
    masks = cases.blob_masks(7, h, h, seed=h + g, min_frac=0.0005, max_frac=0.2)
    masks[0] = 0
    masks[0, h - 1, h - 1] = 1  # single pixel in the last (possibly shared) bin
    masks[1] = 0                # empty
    ref = orc.pool_mask(masks, g).reshape(7, -1).numpy()
    d = masks.to(dev())
    np.testing.assert_array_equal(mb.ops.pool_mask(d, g).cpu().numpy() > 0, ref)
    bits = mb.ops.pack_masks(d)
    pooled, area, cnt = mb.ops.pool_packed(bits, h, h, g)
    np.testing.assert_array_equal(unpack_pooled(pooled, g * g), ref)
    np.testing.assert_array_equal(area.cpu().numpy(), (masks > 0).flatten(1).sum(1).numpy())
    np.testing.assert_array_equal(cnt.cpu().numpy(), ref.sum(1))


@pytest.mark.parametrize("p,h", [(9, 100), (64, 140), (130, 200), (256, 72), (257, 64), (300, 96), (600, 64)])
@pytest.mark.parametrize("backend", [0, 1, 2, 3], ids=["popc", "mma", "fp4", "auto"])
def test_pairwise_intersections_bit_exact(mb, p, h, backend):
    masks = cases.blob_masks(p, h, h, seed=p, min_frac=0.01, max_frac=0.3, dup_every=7)
    inter_ref, area_ref = orc.pairwise_intersections(masks)
    bits = mb.ops.pack_masks(masks.to(dev()))[None]
    if backend == mb.ops.PAIR_FP4 and p > 256:  # one 256-row block only: refused, never silently something else
        with pytest.raises(mb.MarsB200Error, match="256"):
            mb.ops.pairwise_inter(bits, backend=backend)
        return
    inter = mb.ops.pairwise_inter(bits, backend=backend)[0].cpu()
    assert torch.equal(inter, inter_ref)
    assert torch.equal(torch.diagonal(inter), area_ref)


def test_pairwise_fp4_counts_are_exact_at_full_size(mb):
    """Block-scaled FP4 products of 0/1 values accumulate in fp32: exact up to 2^24 pixels.  Dense and all-ones masks at
    1024 x 1024 (every count = 2^20) against the int8 tensor-core kernel; sparse structured masks against popcount."""
    d = dev()
    g = torch.Generator(device=d).manual_seed(5)
    m = (torch.rand(2, 256, 1024, 1024, device=d, generator=g) < 0.6).to(torch.uint8)
    m[1, :40] = 1
    m[0, 7] = 0
    bits = mb.ops.pack_masks(m)
    del m
    fp4 = mb.ops.pairwise_inter(bits, backend=mb.ops.PAIR_FP4)
    assert torch.equal(fp4, mb.ops.pairwise_inter(bits, backend=mb.ops.PAIR_MMA))
    assert int(fp4[1, 3, 17]) == 1024 * 1024 and int(fp4[0, 7].abs().sum()) == 0
    assert torch.equal(fp4[:, :32, :32], mb.ops.pairwise_inter(bits[:, :32].contiguous(), backend=mb.ops.PAIR_POPC))


@pytest.mark.parametrize("backend", [0, 1, 2], ids=["popc", "mma", "fp4"])
def test_pairwise_batched_episodes(mb, backend):
    masks = cases.blob_masks(3 * 20, 96, 96, seed=77).reshape(3, 20, 96, 96)
    bits = mb.ops.pack_masks(masks.to(dev()))
    inter = mb.ops.pairwise_inter(bits, backend=backend).cpu()
    for e in range(3):
        assert torch.equal(inter[e], orc.pairwise_intersections(masks[e])[0])


@pytest.mark.parametrize("e,p,h,dtype", [(1, 9, 100, torch.float32), (2, 200, 96, torch.float32),
                                           (3, 256, 64, torch.float32), (1, 130, 518, torch.float32),
                                           (1, 300, 64, torch.float32),   # P > 256 -> two-kernel path
                                           (2, 64, 96, torch.uint8)])     # u8 -> two-kernel path
def test_fused_pack_pairwise(mb, e, p, h, dtype):
    masks = cases.blob_masks(e * p, h, h, seed=3 * p + e, min_frac=0.01, max_frac=0.3, dup_every=5).reshape(e, p, h, h)
    bits, inter = mb.ops.pack_pairwise(masks.to(dtype).to(dev()), backend=mb.ops.PAIR_MMA)
    wpm = mb.ops.words_per_mask(h * h)
    np.testing.assert_array_equal(bits.cpu().numpy().view(np.uint32).reshape(e * p, wpm),
                                  np_pack(masks.reshape(e * p, h, h).numpy(), wpm))
    for i in range(e):
        assert torch.equal(inter[i].cpu(), orc.pairwise_intersections(masks[i])[0])


# ------------------------------------------------------------------------------------------ VVA
def _backends(mb):
    return [mb.ops.GEMM_SIMT, mb.ops.GEMM_TCGEN05]


@pytest.mark.parametrize("name", list(cases.VVA_CASES))
@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
def test_similarity_and_prior(mb, name, backend):
    z = np.load(os.path.join(GOLD, f"vva_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.vva_inputs(spec)
    g, ns = spec["g"], spec["ns"]
    n, m, k = g * g, ns * g * g, spec["C"]
    fs = mb.ops.normalize_rows(c["feat_s"].reshape(1, m, k).to(dev()))
    fq = mb.ops.normalize_rows(c["feat_q"][None].to(dev()))
    # A1: the normalised rows in the zero-padded operand layout
    fs_ref = orc.normalize_rows(c["feat_s"])
    xn, lo = fs
    np.testing.assert_allclose(xn[0, :m, :k].cpu().numpy(), fs_ref.numpy(), rtol=0, atol=3e-7)
    assert float(xn[0, m:].abs().max() if xn.shape[1] > m else 0.0) == 0.0
    assert float(xn[0, :, k:].abs().max() if xn.shape[2] > k else 0.0) == 0.0
    # the residual is exactly what the top 19 bits of the fp32 word leave out
    hi = (xn.view(torch.int32) & ~0x1fff).view(torch.float32)
    assert torch.equal(hi + lo, xn) and float(lo.abs().max()) < 2.0 ** -10
    row_fg = mb.ops.pool_mask(c["support_mask"].to(dev()), g).reshape(1, m)
    res = mb.ops.sim_contract(fs, fq, m, n, k, want_sim=True, want_cost=True, row_fg=row_fg, backend=backend)
    st = int(z["stride"])
    np.testing.assert_allclose(res["sim"][0].cpu().numpy()[::st, ::st], z["sim"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(res["cost"][0].cpu().numpy()[::st, ::st], z["cost"], rtol=0, atol=3e-6)
    prior = mb.ops.vva_finalize(res["colstats"], row_fg, m, n)[0].cpu()
    fq_ref = orc.normalize_rows(c["feat_q"])
    prior_ref = orc.vva_prior(fs_ref, fq_ref, orc.pool_mask(c["support_mask"], g).reshape(-1), g).reshape(-1)
    np.testing.assert_allclose(prior.numpy(), prior_ref.numpy(), rtol=RTOL, atol=2e-5)


@pytest.mark.parametrize("name", list(cases.VVA_CASES))
def test_vva_module_matches_reference(mb, name):
    """The drop-in class with a fake backbone against the reference module's golden output."""
    z = np.load(os.path.join(GOLD, f"vva_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.vva_inputs(spec)
    regs, cdim, ns, h = spec["regs"], spec["C"], spec["ns"], spec["H"]

    class FakeDino(torch.nn.Module):
        embed_dim = cdim

        def __init__(self):
            super().__init__()
            pad = lambda f: torch.cat([torch.full((f.shape[0], 1 + regs, cdim), 7.0), f], dim=1).to(dev())
            self.feats = [pad(c["feat_s"]), pad(c["feat_q"][None])]

        def forward_features(self, imgs):
            return {"x_prenorm": self.feats.pop(0)}

        def get_last_self_attention(self, img):
            return tuple(a.to(dev()) for a in c["attn_maps"])

    mod = mb.VisualVisualAlignmentModule(
        model=FakeDino(), model_transforms=lambda x: x, model_patch_size=14,
        model_embedding_spatial_dimensions=spec["g"], model_num_regs=regs,
        vva_refinement_box_threshold=spec["thr"], last_n_attention_maps_for_refinement=spec["last_n"], device=dev())
    out = mod.compute(torch.zeros(1, ns, 3, h, h), c["support_mask"][None], torch.zeros(1, 3, h, h))
    np.testing.assert_allclose(out.cpu().numpy(), z["vva_refined"], rtol=RTOL, atol=2e-5)
    st = int(z["stride"])
    np.testing.assert_allclose(mod.cost_matrix.cpu().numpy()[::st, ::st], z["cost"], rtol=0, atol=3e-6)
    mod.clear()
    assert mod.cost_matrix is None and mod.similarity_matrix is None


def test_vva_empty_support_raises(mb):
    spec = cases.VVA_CASES["g10_allfg"]
    c = cases.vva_inputs(spec)

    class Fake(torch.nn.Module):
        embed_dim = spec["C"]

        def forward_features(self, imgs):
            return {"x_prenorm": torch.randn(imgs.shape[0], 1 + 100, spec["C"], device=dev())}

        def get_last_self_attention(self, img):
            return tuple(a.to(dev()) for a in c["attn_maps"])

    mod = mb.VisualVisualAlignmentModule(Fake(), lambda x: x, 14, 10, 0, 0.5, 1, dev())
    with pytest.raises(RuntimeError):
        mod.compute(torch.zeros(1, 1, 3, 140, 140), torch.zeros(1, 1, 140, 140), torch.zeros(1, 3, 140, 140))


# ------------------------------------------------------------------------------------------ PIR
@pytest.mark.parametrize("name", list(cases.PIR_CASES))
@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
def test_pir_matches_reference(mb, name, backend):
    z = np.load(os.path.join(GOLD, f"pir_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.pir_inputs(spec)
    g = spec["g"]
    maps = [a.to(dev()) for a in c["attn_maps"]][-spec["last_n"]:]
    attn = mb.ops.attn_mean(maps, skip=1 + spec["regs"])
    a_ref = orc.attention_mean(c["attn_maps"], spec["last_n"], spec["regs"])
    np.testing.assert_allclose(attn.cpu().numpy(), a_ref.numpy(), rtol=2e-6, atol=1e-9)
    out, box = mb.ops.pir_refine(c["prior"].reshape(1, -1).to(dev()), attn[None], g, spec["thr"], want_box=True,
                                 backend=backend)
    b_ref = np.zeros((g, g), dtype=np.uint8)
    for x0, y0, x1, y1 in z["boxes"][: int(z["cnt"])]:
        b_ref[y0:y1, x0:x1] = 1
    np.testing.assert_array_equal(box.reshape(g, g).cpu().numpy(), b_ref)
    np.testing.assert_allclose(out.reshape(g, g).cpu().numpy(), z["refined"], rtol=RTOL, atol=1e-7)
    # what the pipeline consumes is the min-max of this map (VisualVisualAlignmentModule.py:102, mars/MARS.py:82): the
    # R (R (B*p)) evaluation order must survive that amplification too - the peaky / low-range cases are built for it
    mm = mb.ops.pir_refine(c["prior"].reshape(1, -1).to(dev()), attn[None], g, spec["thr"], apply_minmax=True, backend=backend)
    ref = z["refined"].astype(np.float32)
    ref_mm = (ref - ref.min()) / (np.float32(1e-7) + ref.max() - ref.min())
    np.testing.assert_allclose(mm.reshape(g, g).cpu().numpy(), ref_mm, rtol=RTOL, atol=2e-5)


@pytest.mark.parametrize("name", list(cases.PIR_CASES))
def test_scoremap2bbox_returns_the_reference_boxes(mb, name):
    """`_scoremap2bbox` keeps the reference's return type (boxes ndarray, count): every box is one of OpenCV's boxes of
    the golden run (which may list hole contours in addition) and together they fill the same mask."""
    z = np.load(os.path.join(GOLD, f"pir_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.pir_inputs(spec)
    g = spec["g"]
    pir = mb.PriorInformationRefinementModule(spec["thr"], spec["last_n"], dev(), num_regs=spec["regs"])
    boxes, cnt = pir._scoremap2bbox(scoremap=c["prior"].numpy(), multi_contour_eval=True)
    assert isinstance(boxes, np.ndarray) and boxes.shape == (cnt, 4)
    gold = {tuple(int(v) for v in b) for b in z["boxes"][: int(z["cnt"])]}
    assert {tuple(int(v) for v in b) for b in boxes} <= gold
    b_ref, b_got = np.zeros((g, g), dtype=np.uint8), np.zeros((g, g), dtype=np.uint8)
    for x0, y0, x1, y1 in z["boxes"][: int(z["cnt"])]:
        b_ref[y0:y1, x0:x1] = 1
    for x0, y0, x1, y1 in boxes:
        b_got[y0:y1, x0:x1] = 1
    np.testing.assert_array_equal(b_got, b_ref)
    one, k = pir._scoremap2bbox(scoremap=c["prior"].numpy(), multi_contour_eval=False)
    assert k == 1 and one.shape == (1, 4)
    none, k0 = mb.PriorInformationRefinementModule(1.0, 1, dev())._scoremap2bbox(np.zeros((g, g), np.float32), True)
    assert k0 == 1 and none.tolist() == [[0, 0, 0, 0]]


def test_pir_module_fp16_and_3d_maps(mb):
    spec = cases.PIR_CASES["g33_border"]
    c = cases.pir_inputs(spec)
    maps16 = [a.half() for a in c["attn_maps"]]
    ref = orc.pir_refine(c["prior"], orc.attention_mean(maps16, spec["last_n"], 0), spec["thr"])
    mod = mb.PriorInformationRefinementModule(spec["thr"], spec["last_n"], dev(), num_regs=0)
    out = mod.compute(c["prior"], [a.to(dev()) for a in maps16])
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("tokens", [38, 37, 70])
@pytest.mark.parametrize("regs", [0, 3, 4])
def test_attn_mean_fp16_layouts(mb, tokens, regs):
    """fp16 attention maps: the paired-column kernel (even token count) and the scalar one (odd) against the oracle,
    for even and odd numbers of skipped cls / register tokens."""
    g = torch.Generator().manual_seed(tokens * 10 + regs)
    maps = [torch.softmax(2 * torch.randn(1, 3, tokens, tokens, generator=g), dim=-1).half() for _ in range(4)]
    ref = orc.attention_mean(maps, 3, regs)
    got = mb.ops.attn_mean([a.to(dev()) for a in maps[-3:]], skip=1 + regs)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=1e-3, atol=1e-6)  # one fp16 rounding of the mean
    # a view that starts at an odd element: not 4-byte aligned, takes the scalar kernel, same numbers
    flat = torch.zeros(3 * tokens * tokens + 1, dtype=torch.float16, device=dev())
    shifted = []
    for a in maps[-3:]:
        buf = flat.clone()
        buf[1:] = a.reshape(-1).to(dev())
        shifted.append(buf[1:].view(3, tokens, tokens))
    again = mb.ops.attn_mean(shifted, skip=1 + regs)
    assert torch.equal(again, got)


def test_resize_minmax(mb):
    x = torch.rand(3, 33, 33, generator=torch.Generator().manual_seed(3))
    out = mb.ops.resize_minmax(x.to(dev()), 37).reshape(3, 37, 37).cpu()
    for e in range(3):
        ref = orc.minmax(orc.nearest_resize(x[e], (37, 37)))
        np.testing.assert_allclose(out[e].numpy(), ref.numpy(), rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------------------ fuse / rank / merge
@pytest.mark.parametrize("name", list(cases.FM_CASES))
def test_filtering_merging_module_matches_reference(mb, name):
    z = np.load(os.path.join(GOLD, f"fm_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.fm_inputs(spec)
    h = spec["H"]
    mod = mb.FilteringMergingModule(None, None, None, alpha=spec["alpha"], static_threshold=spec["static"],
                                    dynamic_threshold=spec["dynamic"], device=dev())
    kw = dict(query_img=torch.zeros(1, 3, h, h), mask_proposals=c["masks"], support_mask=c["support_mask"][None],
              cost_matrix=c["cost"].to(dev()), patch_features_spatial_dimension=spec["g"], vva=c["vva"], vta=c["vta"],
              text=["a thing."], emd_scores=z["emd"], alphaclip_feats=cases.alphaclip_as_seen(c))
    ranked = mod._score_proposals(**kw)
    order = [[i for i in range(spec["P"]) if m.data_ptr() == c["masks"][i].data_ptr()][0] for m, _ in ranked]
    scores_by_index_ref = np.empty(spec["P"])
    scores_by_index_ref[z["order"]] = z["scores"]
    scores_by_index = np.empty(spec["P"])
    scores_by_index[order] = [s for _, s in ranked]
    assert_order_matches(order, scores_by_index, z["order"], scores_by_index_ref)
    merged_ref = np.unpackbits(z["merged_bits"])[: h * h].reshape(h, h) > 0
    merged = mod.compute(**kw)
    np.testing.assert_array_equal(merged.cpu().numpy() > 0, merged_ref)
    assert merged.dtype == torch.float32 and tuple(merged.shape) == (h, h)
    np.testing.assert_array_equal(mod._merge_masks(ranked).cpu().numpy() > 0, merged_ref)


def test_filtering_merging_host_emd_path(mb):
    """Without precomputed EMD the module gathers the cost sub-matrix and calls the host solver, like the reference."""
    z = np.load(os.path.join(GOLD, "fm_g7_overlap.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.fm_inputs(spec)
    mod = mb.FilteringMergingModule(None, None, None, spec["alpha"], spec["static"], spec["dynamic"], dev(),
                                    emd_fn=orc.emd_exact)
    ranked = mod._score_proposals(torch.zeros(1, 3, 100, 100), c["masks"], c["support_mask"][None], c["cost"].to(dev()),
                                  spec["g"], c["vva"], c["vta"], ["x"], alphaclip_feats=(c["clip_img"], c["clip_txt"]))
    np.testing.assert_allclose([s for _, s in ranked], z["scores"], rtol=RTOL)


@pytest.mark.parametrize("precomputed", [False, True])
@pytest.mark.parametrize("p,thr", [(12, 0.5), (64, 0.7), (200, 0.3), (1, 0.7), (33, 0.0), (700, 0.5), (1100, 0.5)])
def test_nms_keep_set_bit_exact(mb, p, thr, precomputed):
    """precomputed: the suppression relation comes from marsb200_nms_bitmask (the engine's path) instead of being built
    inside the ranking kernel; both must give the oracle's keep-set."""
    h = 96
    masks = cases.blob_masks(p, h, h, seed=100 + p, min_frac=0.02, max_frac=0.3, dup_every=3)
    rs = np.random.RandomState(p)
    scores = rs.rand(p)
    if p > 4:
        scores[3] = scores[1]  # exact tie: the stable rule keeps index order
    inter_ref, area_ref = orc.pairwise_intersections(masks)
    order_ref = orc.stable_rank(scores)
    keep_ref = orc.mask_nms(order_ref, inter_ref, area_ref, thr)
    # drive fuse_rank so that its fused score equals `scores`: emd/clip constant, pvv = pvt = 2*score
    d = dev()
    bits = mb.ops.pack_masks(masks.to(d))[None]
    inter = mb.ops.pairwise_inter(bits)
    cnt = torch.ones((1, p), dtype=torch.int32, device=d)
    sv = torch.as_tensor(2 * scores, dtype=torch.float32, device=d).reshape(1, p)
    uc = torch.ones((1,), dtype=torch.int32, device=d)
    zero = torch.zeros((1, p), device=d)
    rel = mb.ops.nms_bitmask(inter, thr) if precomputed else None  # (P > 1024: the ranking kernel ignores it and walks the ranks itself)
    if precomputed:
        # the relation itself against its definition (float32 quotient, strict comparison, no self-suppression)
        want = (orc.iou_matrix(inter_ref, area_ref).numpy() > np.float32(thr)) & ~np.eye(p, dtype=bool)
        words = rel[0].cpu().numpy().view(np.uint32)
        got = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(p, -1)[:, :p].astype(bool)
        np.testing.assert_array_equal(got, want)
    res = mb.ops.fuse_rank(zero.double(), zero, cnt, sv, sv, uc, inter, alpha=1.0, static_threshold=0.0,
                           dynamic_threshold=0.0, nms_iou_threshold=thr, nms_bits=rel)
    order = res["order"][0].cpu().numpy()
    s32 = (2 * scores).astype(np.float32).astype(np.float64) / (1e-7 + 1)
    order_expected = orc.stable_rank((s32 + s32) / 4)
    np.testing.assert_array_equal(order, order_expected)
    keep_expected = orc.mask_nms(order_expected, inter_ref, area_ref, thr)
    flags = res["flags"][0].cpu().numpy()
    np.testing.assert_array_equal((flags & 1).astype(bool), keep_expected)
    assert int(res["summary"][0, 0]) == int(keep_expected.sum())
    if np.array_equal(order_expected, order_ref):
        np.testing.assert_array_equal(keep_expected, keep_ref)


# ------------------------------------------------------------------------------------------ whole episodes
def _run_engine(mb, shape, n_ep, cfg, mask_dtype=torch.float32, seed0=0):
    eps = [mb.make_episode(shape, seed0 + i, "cpu", mask_dtype) for i in range(n_ep)]
    eng = mb.RankingEngine(shape, n_ep, cfg, dev(), mask_dtype)
    out = eng.run(mb.to_device(mb.stack_episodes(eps), dev()))
    torch.cuda.synchronize()
    return eps, eng, {k: (v.cpu() if v is not None else None) for k, v in out.items()}


def _check_episode(mb, shape, cfg, ep, out, e, emd_fn=None):
    """One episode of a device batch against `orc.run_episode` (oracle/compare.py): maps and scores within RTOL,
    integer outputs bit-exact; the NMS keep-set, the selection and the merged mask are ALWAYS checked - on the device's
    own order when a tolerated tie swapped two ranks."""
    ocfg = compare.oracle_config(cfg, shape.g)
    ep = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in ep.items()}
    ep["masks"] = ep["masks"].float()
    ref = orc.run_episode(ep, ocfg, emd_fn=emd_fn)
    got = {k: (v[e] if v is not None else None) for k, v in out.items()}
    res = compare.compare_episode(ref, got, ep["masks"], ocfg, rtol=RTOL)
    assert res["ok"], res["failures"]
    return ref, res


@pytest.mark.parametrize("nms", [None, 0.7])
@pytest.mark.parametrize("mask_dtype", [torch.float32, torch.uint8])
def test_engine_small_episodes(mb, nms, mask_dtype):
    shape = mb.EpisodeShape(ns=2, g=10, C=64, P=24, H=140, W=140, gt=8, D=32)
    cfg = mb.RankingConfig(nms_iou_threshold=nms, want_sim=True, want_cost=True)
    eps, eng, out = _run_engine(mb, shape, 3, cfg, mask_dtype)
    for e in range(3):
        _check_episode(mb, shape, cfg, eps[e], out, e)


@pytest.mark.parametrize("episodes", [1, 3])
def test_engine_stream_options_are_bit_identical(mb, episodes):
    """The one-timeline schedule's stream options (high-priority alignment streams, hoisted vva contraction, no side streams
    at all) only move kernels between streams: every output must be bit-identical."""
    shape = mb.EpisodeShape(ns=1, g=14, gt=12, C=64, D=32, P=40, H=96, W=96)
    batch = mb.stack_episodes([mb.make_episode(shape, 70 + i, dev()) for i in range(episodes)])
    outs = []
    for kw in (dict(overlap_streams=False), dict(priority_streams=False, hoist_vva_contraction=False),
               dict(priority_streams=True, hoist_vva_contraction=False), dict(priority_streams=True, hoist_vva_contraction=True),
               dict(latency_contraction_sms=0), dict(latency_contraction_sms=1), dict(latency_contraction_sms=48)):
        eng = mb.RankingEngine(shape, episodes, mb.RankingConfig(nms_iou_threshold=0.6, **kw), dev())
        o = eng.run(batch)
        torch.cuda.synchronize()
        outs.append({k: o[k].clone() for k in ("vva", "vta", "inter", "scores", "order", "flags", "merged_bits", "clip", "row_fg")})
    for o in outs[1:]:
        for k, v in o.items():
            assert torch.equal(v, outs[0][k]), k


@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8, torch.bool])
def test_pack_slices_tile_the_mask(mb, dtype):
    """marsb200_pack_masks_slice: slices of whole 512-word blocks, in any order, land exactly the bits of the one-launch pack;
    slices that are not whole blocks or leave the mask are refused."""
    ops = mb.ops
    n, h, w = 7, 256, 320  # 81920 pixels = 2560 words = 5 blocks of 512
    masks = cases.blob_masks(n, h, w, seed=11).to(dtype).to(dev())
    want = ops.pack_masks(masks)
    got = torch.full_like(want, -1)
    for k in (3, 0, 4, 1):
        ops.pack_masks_slice(masks, 512 * k, 512, out=got)
    assert not torch.equal(got, want)
    ops.pack_masks_slice(masks, 1024, 512, out=got)
    assert torch.equal(got, want)
    got2 = torch.full_like(want, -1)
    ops.pack_masks_slice(masks, 0, 1024, out=got2)
    ops.pack_masks_slice(masks, 1024, 1536, out=got2)
    assert torch.equal(got2, want)
    for wb, wc in ((0, 256), (256, 512), (2048, 1024), (0, 0)):
        with pytest.raises(mb.MarsB200Error):
            ops.pack_masks_slice(masks, wb, wc, out=got)


@pytest.mark.parametrize("p,backend", [(40, "PAIR_FP4"), (256, "PAIR_FP4"), (40, "PAIR_MMA"), (300, "PAIR_MMA"), (300, "PAIR_AUTO")])
def test_pairwise_slices_sum_to_the_whole(mb, p, backend):
    """marsb200_pairwise_inter_slice: the contributions of pixel slices that tile the mask add up (integer atomics) to the
    one-launch intersections, bit for bit, whatever the order; the popcount back end has no slices."""
    ops = mb.ops
    backend = getattr(ops, backend)
    e, h, w = 2, 128, 192  # 24576 pixels = 768 words
    bits = ops.pack_masks(cases.blob_masks(e * p, h, w, seed=p).reshape(e, p, h, w).to(dev()))
    want = ops.pairwise_inter(bits, backend=backend)
    assert torch.equal(want, ops.pairwise_inter(bits, backend=ops.PAIR_POPC))
    got = torch.full_like(want, 12345)
    for i, (wb, wc) in enumerate(((512, 256), (0, 8), (8, 248), (256, 256))):
        ops.pairwise_inter_slice(bits, wb, wc, i > 0, out=got, backend=backend)
    assert torch.equal(got, want)
    with pytest.raises(mb.MarsB200Error):
        ops.pairwise_inter_slice(bits, 4, 8, False, out=got, backend=backend)
    with pytest.raises(mb.MarsB200Error):
        ops.pairwise_inter_slice(bits, 0, 776, False, out=got, backend=backend)
    with pytest.raises(mb.MarsB200Error):
        ops.pairwise_inter_slice(bits, 0, 256, False, out=got, backend=ops.PAIR_POPC)


@pytest.mark.parametrize("episodes", [1, 2])
@pytest.mark.parametrize("mask_dtype", [torch.float32, torch.uint8])
def test_engine_sliced_ingest_is_bit_identical(mb, episodes, mask_dtype):
    """Optional schedule: the masks are packed in pixel slices and the intersections of one slice counted beside the read of
    the next.  Every output equals the unsliced schedule, eagerly and as a CUDA-graph replay; masks that do not tile into
    512-word blocks keep one slice."""
    shape = mb.EpisodeShape(ns=1, g=14, gt=12, C=64, D=32, P=40, H=256, W=256)
    batch = mb.stack_episodes([mb.make_episode(shape, 90 + i, dev(), mask_dtype) for i in range(episodes)])
    keys = ("vva", "vta", "inter", "scores", "order", "flags", "merged_bits", "clip", "pooled", "area")
    outs = []
    for k in (1, 2, 4):
        eng = mb.RankingEngine(shape, episodes, mb.RankingConfig(nms_iou_threshold=0.6, latency_ingest_slices=k), dev(), mask_dtype)
        assert eng._slices == k
        o = eng.run(batch)
        torch.cuda.synchronize()
        outs.append({kk: o[kk].clone() for kk in keys})
        eng.capture(batch)
        eng.replay()
        o = eng.replay()
        torch.cuda.synchronize()
        outs.append({kk: o[kk].clone() for kk in keys})
    for o in outs[1:]:
        for kk, v in o.items():
            assert torch.equal(v, outs[0][kk]), kk
    odd = mb.EpisodeShape(ns=1, g=14, gt=12, C=64, D=32, P=40, H=96, W=96)
    assert mb.RankingEngine(odd, 1, mb.RankingConfig(nms_iou_threshold=0.6, latency_ingest_slices=4), dev())._slices == 1
    assert mb.RankingEngine(shape, 1, mb.RankingConfig(nms_iou_threshold=0.6), dev())._slices == 1  # off by default (measured slower)
    assert mb.RankingEngine(shape, 1, mb.RankingConfig(latency_ingest_slices=4), dev())._slices == 1  # no NMS, no intersections


@pytest.mark.parametrize("raw_fraction", [0.0, 0.3, 1.0])
@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
def test_host_mask_ingest_equals_device_packing(mb, raw_fraction, dtype):
    """HostMaskIngest (a share of the proposals raw over PCIe + device kernel, the rest packed by host threads) lands the
    same bits as packing the whole batch on the device, for every split."""
    e, p, h, w = 2, 10, 96, 160
    masks = cases.blob_masks(e * p, h, w, seed=5).reshape(e, p, h, w).to(dtype)
    host = masks.pin_memory()
    ing = mb.HostMaskIngest(e, p, h, w, dev(), mask_dtype=dtype, raw_fraction=raw_fraction, threads=3)
    st = torch.cuda.Stream(device=dev())
    for _ in range(2):  # the second upload reuses the pinned buffer
        bits = ing.upload(host, st)
        st.synchronize()
    want = mb.ops.pack_masks(masks.to(dev()))
    assert torch.equal(bits, want.reshape(bits.shape))
    assert ing.h2d_bytes(masks.element_size()) == e * ing.p_raw * h * w * masks.element_size() + e * (p - ing.p_raw) * bits.shape[-1] * 4


def test_engine_c1_shape_against_oracle(mb):
    """BASELINE config 1 (the reference's CPU-runnable case): N=1369, C=1024, P=128 at 518x518."""
    shape = mb.CONFIGS["c1"]
    cfg = mb.RankingConfig(nms_iou_threshold=0.7)
    eps, eng, out = _run_engine(mb, shape, 1, cfg)
    _check_episode(mb, shape, cfg, eps[0], out, 0)


def test_engine_graph_replay_matches_eager(mb):
    shape = mb.EpisodeShape(ns=1, g=10, C=64, P=16, H=140, W=140, gt=8, D=32)
    cfg = mb.RankingConfig(nms_iou_threshold=0.7)
    batch = mb.to_device(mb.stack_episodes([mb.make_episode(shape, i) for i in range(2)]), dev())
    eng = mb.RankingEngine(shape, 2, cfg, dev())
    eager = {k: v.clone() for k, v in eng.run(batch).items() if v is not None}
    eng.capture(batch)
    replay = eng.replay(batch)
    torch.cuda.synchronize()
    for k, v in eager.items():
        assert torch.equal(v, replay[k]), k
    # the result records are written by the fuse / rank kernel itself (no concatenation): every field equals the outputs
    rec = mb.decode_records(eng.records().cpu(), shape.P)
    assert torch.equal(rec["order"], replay["order"].cpu())
    assert torch.equal(rec["scores"], replay["scores"].float().cpu())
    assert torch.equal(rec["flags"], replay["flags"].cpu())
    assert torch.equal(rec["summary"], replay["summary"].cpu())
    assert eng.records().shape == (2, mb.ops.record_bytes(shape.P))


def test_full_size_properties_c2(mb):
    """Size-independent properties at BASELINE config 2 (P=256 at 1024x1024), where the oracle is too slow to loop."""
    shape = mb.CONFIGS["c2"]
    d = dev()
    gen = torch.Generator(device=d).manual_seed(9)
    from marsb200.synthetic import random_masks

    masks = random_masks(shape.P, shape.H, shape.W, gen, d)
    bits = mb.ops.pack_masks(masks)
    # round trip: merging a single packed row and expanding it gives the mask back
    for i in (0, 17, 255):
        flags = torch.zeros((1, shape.P), dtype=torch.uint8, device=d)
        flags[0, i] = 3
        _, back = mb.ops.merge_masks(bits[None], flags, shape.H * shape.W)
        assert torch.equal(back.reshape(shape.H, shape.W), masks[i])
    # u8 ingest packs to the same bits
    assert torch.equal(mb.ops.pack_masks(masks.to(torch.uint8)), bits)
    inter = mb.ops.pairwise_inter(bits[None])[0]
    area = masks.flatten(1).sum(1).to(torch.int32)
    assert torch.equal(torch.diagonal(inter), area)
    assert torch.equal(inter, inter.T)
    assert bool((inter <= torch.minimum(area[:, None], area[None, :])).all())
    # spot rows against an exact torch contraction on the device
    flat = masks.flatten(1)
    for i in (3, 200):
        assert torch.equal(inter[i], (flat * flat[i]).sum(1).to(torch.int32))
    # duplicates injected by the generator have IoU 1 with their source
    pooled, area2, cnt = mb.ops.pool_packed(bits, shape.H, shape.W, shape.g)
    assert torch.equal(area2, area)
    ref_pool = torch.nn.functional.adaptive_max_pool2d(masks[:8, None], (shape.g, shape.g)).flatten(1) > 0
    np.testing.assert_array_equal(unpack_pooled(pooled[:8], shape.N), ref_pool.cpu().numpy())


# ------------------------------------------------------------------------------------------ BASELINE configs at FULL size
def _cpu_out(out, keys=("row_fg", "prior", "vva", "vta", "pooled", "clip", "inter", "area", "scores", "order", "flags",
                         "merged_bits", "emd")):
    return {k: (out[k].cpu() if out.get(k) is not None else None) for k in keys}


@pytest.mark.parametrize("mask_dtype", [torch.float32, torch.uint8])
def test_engine_c2_full_against_oracle(mb, mask_dtype):
    """BASELINE config 2 at full size (1-shot, N=1369, C=1024, P=256 at 1024x1024, IoU-NMS 0.7): two whole episodes
    through RankingEngine against orc.run_episode - maps, pooled bitmaps, all 65 536 intersections, scores, order,
    keep-set, selection and the merged mask (mars/MARS.py:62-101, FilteringMergingModule.py:59-140)."""
    shape = mb.CONFIGS["c2"]
    cfg = mb.RankingConfig(nms_iou_threshold=0.7, want_merged_f32=False)
    batch = mb.stack_episodes([mb.make_episode(shape, 9000 + i, dev(), mask_dtype) for i in range(2)])
    eng = mb.RankingEngine(shape, 2, cfg, dev(), mask_dtype)
    out = eng.run(batch)
    torch.cuda.synchronize()
    out = _cpu_out(out)
    for e in range(2):
        _check_episode(mb, shape, cfg, {k: v[e] for k, v in batch.items()}, out, e)


def test_engine_c2_full_pipelined_bench_schedule(mb):
    """The exact schedule bench.py times: PipelinedRanking, 16 c2 episodes per step on two SM partitions, two buffer
    sets, two resident batches alternating.  After three steps the records of steps 1 and 2 (first and last episode of
    each) equal the oracle's results for those episodes."""
    shape = mb.CONFIGS["c2"]
    E = 16
    cfg = mb.RankingConfig(nms_iou_threshold=0.7, tensor_partition_sms=64, partition_vta_on_hbm=False)
    batches = [mb.stack_episodes([mb.make_episode(shape, 9200 + b * E + i, dev()) for i in range(E)]) for b in range(2)]
    pipe = mb.PipelinedRanking(shape, E, cfg, dev(), depth=2)
    try:
        tickets = [pipe.submit(batches[0]), pipe.submit(batches[1])]
        pipe.result(tickets[0])
        tickets.append(pipe.submit(batches[0]))  # reuses buffer set 0 after its first step drained
        for t in tickets[1:]:
            out = pipe.result(t)
            torch.cuda.synchronize()
            rec = mb.decode_records(pipe.engine(t).records().cpu(), shape.P)
            full = _cpu_out(out)
            assert torch.equal(rec["order"], full["order"]) and torch.equal(rec["flags"], full["flags"])
            for e in (0, E - 1):
                _check_episode(mb, shape, cfg, {k: v[e] for k, v in batches[t % 2].items()}, full, e)
    finally:
        pipe.close()


def test_engine_c3_full_against_oracle(mb):
    """BASELINE config 3 at full size: 5-shot (6845 support rows), P=512 at 518x518 - one whole episode against the
    oracle (P > 256: the intersections take the multi-block tensor-core path), plus the transport LPs of 16 proposals
    solved on the device against the oracle's exact LP (FilteringMergingModule.py:142-169)."""
    shape = mb.CONFIGS["c3"]
    cfg = mb.RankingConfig(nms_iou_threshold=0.7, want_cost=True, want_merged_f32=False)
    eps, eng, out = _run_engine(mb, shape, 1, cfg, seed0=9100)
    ref, _ = _check_episode(mb, shape, cfg, eps[0], out, 0)
    # 16 proposals spread over the size range (every 32nd of the proposals sorted by pooled patch count)
    counts = ref["pooled"].reshape(shape.P, -1).sum(1)
    pick = np.argsort(counts, kind="stable")[15::32][:16]
    d = dev()
    pooled = eng.pool_out[0][0][torch.as_tensor(pick, device=d)][None].contiguous()
    t_fg = int(ref["support_bits"].sum())
    emd = mb.ops.emd_scores(eng.gemm_out["cost"], eng.row_fg.reshape(1, -1), pooled, t_cap=(t_fg + 63) // 64 * 64)
    torch.cuda.synchronize()
    want = np.asarray([orc.emd_score(ref["support_bits"], torch.from_numpy(ref["pooled"][i]), ref["cost"]) for i in pick])
    np.testing.assert_allclose(emd[0].cpu().numpy(), want, rtol=0, atol=2e-5)  # costs differ by ~1e-6 (3xTF32 vs MKL)


def test_mask_chain_c4_full_against_oracle(mb):
    """BASELINE config 4 at full size: P=1000 at 1024x1024 - all 10^6 intersections of the tensor-core kernel against
    the oracle's blockwise exact contraction, then the NMS keep-set, selection and merged mask."""
    shape = mb.CONFIGS["c4"]
    d = dev()
    gen = torch.Generator(device=d).manual_seed(4004)
    from marsb200.synthetic import random_masks

    masks = random_masks(shape.P, shape.H, shape.W, gen, d, dtype=torch.uint8)
    bits = mb.ops.pack_masks(masks)[None]
    inter = mb.ops.pairwise_inter(bits)
    masks_cpu = masks.cpu()
    inter_ref, area_ref = orc.pairwise_intersections(masks_cpu)
    assert torch.equal(inter[0].cpu(), inter_ref)
    p = shape.P
    rs = np.random.RandomState(11)
    scores = rs.rand(p)
    cnt = torch.ones((1, p), dtype=torch.int32, device=d)
    sv = torch.as_tensor(2 * scores, dtype=torch.float32, device=d).reshape(1, p)
    uc = torch.ones((1,), dtype=torch.int32, device=d)
    zero = torch.zeros((1, p), device=d)
    res = mb.ops.fuse_rank(zero.double(), zero, cnt, sv, sv, uc, inter, alpha=1.0, static_threshold=0.9,
                           dynamic_threshold=0.5, nms_iou_threshold=0.7)
    s32 = (2 * scores).astype(np.float32).astype(np.float64) / (1e-7 + 1)
    fused = (s32 + s32) / 4
    order_expected = orc.stable_rank(fused)
    np.testing.assert_array_equal(res["order"][0].cpu().numpy(), order_expected)
    keep_expected = orc.mask_nms(order_expected, inter_ref, area_ref, 0.7)
    flags = res["flags"][0].cpu().numpy()
    np.testing.assert_array_equal((flags & 1).astype(bool), keep_expected)
    sel_expected = np.zeros(p, dtype=bool)
    sel_ranked = orc.merge_select(fused[order_expected], 0.9, 0.5) & keep_expected[order_expected]
    sel_expected[order_expected[sel_ranked]] = True
    np.testing.assert_array_equal((flags & 2) > 0, sel_expected)
    merged_bits, _ = mb.ops.merge_masks(bits, res["flags"], shape.H * shape.W, want_bits=True, want_f32=False)
    ref = orc.merge_masks(masks_cpu, np.nonzero(sel_expected)[0])
    got = compare.unpack_bits(merged_bits.cpu().reshape(1, -1), shape.H * shape.W)[0]
    np.testing.assert_array_equal(got, ref.reshape(-1).numpy() > 0)


# ------------------------------------------------------------------------------------------ Matcher / evaluator
def test_matcher_scoring(mb):
    h, g, n, k = 140, 10, 14, 60
    masks = cases.blob_masks(n, h, h, seed=41, min_frac=0.01, max_frac=0.2)
    rs = np.random.RandomState(4)
    pts = np.stack([rs.randint(-3, h + 3, k), rs.randint(-3, h + 3, k)], axis=1)  # some outside: clipped
    purity_ref, cov_ref = orc.matcher_mask_scores(masks.numpy() > 0, pts, g)
    emd = torch.rand(n)
    ref_scores = orc.matcher_fuse(emd, purity_ref, cov_ref, 1.0, 0.5, 2.0)
    bits = mb.ops.pack_masks(masks.to(dev()))
    pin = mb.ops.points_in_masks(bits, h, h, torch.as_tensor(pts, dtype=torch.int32, device=dev()))
    _, _, cnt = mb.ops.pool_packed(bits, h, h, g)
    purity, cov, scores = mb.ops.matcher_scores(pin, cnt, emd.to(dev()), k, 1.0, 0.5, 2.0)
    np.testing.assert_allclose(purity.cpu().numpy(), purity_ref.numpy(), rtol=1e-6)
    np.testing.assert_allclose(cov.cpu().numpy(), cov_ref.numpy(), rtol=1e-6)
    np.testing.assert_allclose(scores.cpu().numpy(), ref_scores.numpy(), rtol=1e-5)


def test_evaluator_areas(mb):
    g = torch.Generator().manual_seed(8)
    pred = (torch.rand(3, 90, 70, generator=g) < 0.4).float()
    gt = (torch.rand(3, 90, 70, generator=g) < 0.5).float()
    ignore = ((torch.rand(3, 90, 70, generator=g) < 0.1) & (gt == 0)).float()
    for ig in (None, ignore):
        inter_ref, union_ref = orc.evaluator_areas(pred, gt, ig)
        out = mb.ops.eval_areas(pred.to(dev()), gt.to(dev()), None if ig is None else ig.to(dev())).cpu()
        np.testing.assert_array_equal(out[:, :2].numpy(), inter_ref.t().numpy().astype(np.int32))
        np.testing.assert_array_equal(out[:, 2:].numpy(), union_ref.t().numpy().astype(np.int32))


# ------------------------------------------------------------------------------------------ other BASELINE configs
def test_engine_5shot_c3_like(mb):
    """5-shot episodes (BASELINE config 3 geometry at a reduced size): shot-major support rows, T > N possible."""
    shape = mb.EpisodeShape(ns=5, g=12, C=96, P=40, H=168, W=168, gt=9, D=48)
    cfg = mb.RankingConfig(nms_iou_threshold=0.6)
    eps, eng, out = _run_engine(mb, shape, 2, cfg)
    for e in range(2):
        _check_episode(mb, shape, cfg, eps[e], out, e)


def test_high_proposal_c4_like(mb):
    """P = 1000 (BASELINE config 4) at a reduced resolution: multi-block tensor-core pairwise + bitmask NMS."""
    p, h = 1000, 128
    masks = cases.blob_masks(p, h, h, seed=4242, min_frac=0.01, max_frac=0.25, dup_every=9)
    inter_ref, area_ref = orc.pairwise_intersections(masks)
    d = dev()
    bits = mb.ops.pack_masks(masks.to(d))[None]
    for backend in (mb.ops.PAIR_MMA, mb.ops.PAIR_POPC):
        assert torch.equal(mb.ops.pairwise_inter(bits, backend=backend)[0].cpu(), inter_ref)
    inter = mb.ops.pairwise_inter(bits)
    rs = np.random.RandomState(7)
    scores = rs.rand(p)
    cnt = torch.ones((1, p), dtype=torch.int32, device=d)
    sv = torch.as_tensor(2 * scores, dtype=torch.float32, device=d).reshape(1, p)
    uc = torch.ones((1,), dtype=torch.int32, device=d)
    zero = torch.zeros((1, p), device=d)
    res = mb.ops.fuse_rank(zero.double(), zero, cnt, sv, sv, uc, inter, alpha=1.0, static_threshold=0.9,
                           dynamic_threshold=0.5, nms_iou_threshold=0.7)
    s32 = (2 * scores).astype(np.float32).astype(np.float64) / (1e-7 + 1)
    fused = (s32 + s32) / 4
    order_expected = orc.stable_rank(fused)
    np.testing.assert_array_equal(res["order"][0].cpu().numpy(), order_expected)
    keep_expected = orc.mask_nms(order_expected, inter_ref, area_ref, 0.7)
    flags = res["flags"][0].cpu().numpy()
    np.testing.assert_array_equal((flags & 1).astype(bool), keep_expected)
    sel_expected = np.zeros(p, dtype=bool)
    sel_ranked = orc.merge_select(fused[order_expected], 0.9, 0.5) & keep_expected[order_expected]
    sel_expected[order_expected[sel_ranked]] = True
    np.testing.assert_array_equal((flags & 2) > 0, sel_expected)
    _, merged = mb.ops.merge_masks(bits, res["flags"], h * h)
    ref = orc.merge_masks(masks, np.nonzero(sel_expected)[0])
    np.testing.assert_array_equal(merged.reshape(h, h).cpu().numpy() > 0, ref.numpy() > 0)


def test_mars_predict_dropin(mb):
    """MARS.predict with fake PyTorch producers: same composition as mars/MARS.py:33-104, checked against the oracle."""
    spec = cases.VVA_CASES["g10_2shot"]
    c = cases.vva_inputs(spec)
    g, h, ns, cdim, regs = spec["g"], spec["H"], spec["ns"], spec["C"], spec["regs"]
    p = 14
    masks = cases.blob_masks(p, h, h, seed=91, min_frac=0.02, max_frac=0.3)
    gen = torch.Generator().manual_seed(17)
    vta_raw = torch.rand(8, 8, generator=gen)
    img = torch.nn.functional.normalize(torch.randn(p, 24, generator=gen), dim=1)
    txt = torch.nn.functional.normalize(torch.randn(24, generator=gen), dim=0)
    emd = torch.rand(p, generator=gen, dtype=torch.float64)

    class FakeDino(torch.nn.Module):
        embed_dim = cdim

        def __init__(self):
            super().__init__()
            pad = lambda f: torch.cat([torch.zeros(f.shape[0], 1 + regs, cdim), f], dim=1).to(dev())
            self.feats = [pad(c["feat_s"]), pad(c["feat_q"][None])]

        def forward_features(self, imgs):
            return {"x_prenorm": self.feats.pop(0)}

        def get_last_self_attention(self, img):
            return tuple(a.to(dev()) for a in c["attn_maps"])

    class FakeText:
        def get_conceptual_information(self, support_images, support_masks):
            return "thing", ""

    class FakeVTA:
        def compute(self, query_image, fg_label, bg_labels):
            return vta_raw

    vva_mod = mb.VisualVisualAlignmentModule(FakeDino(), lambda x: x, 14, g, regs, spec["thr"], spec["last_n"], dev())
    fm = mb.FilteringMergingModule(None, None, None, alpha=0.85, static_threshold=0.55, dynamic_threshold=0.95,
                                   device=dev())
    orig = fm.compute
    fm.compute = lambda **kw: orig(emd_scores=emd.numpy(), alphaclip_feats=(img, txt), **kw)
    mars = mb.MARS(FakeText(), FakeVTA(), vva_mod, fm)
    pred = mars.predict(torch.zeros(1, ns, 3, h, h), c["support_mask"][None], torch.zeros(1, 3, h, h), masks)
    assert mars.time_start_ranking <= mars.time_start_ranking_after_text_extraction <= mars.time_end_ranking
    # oracle composition of the same stage
    fs, fq = orc.normalize_rows(c["feat_s"]), orc.normalize_rows(c["feat_q"])
    bits = orc.pool_mask(c["support_mask"], g).reshape(-1)
    prior = orc.vva_prior(fs, fq, bits, g)
    a = orc.attention_mean(c["attn_maps"], spec["last_n"], regs)
    vva = orc.minmax(orc.pir_refine(prior, a, spec["thr"]))
    vta = orc.minmax(orc.nearest_resize(vta_raw, (g, g)))
    pooled, cov, avv, avt = orc.region_scores(masks, vva.numpy(), vta.numpy(), g)
    scores = orc.fuse_scores(emd.numpy(), orc.clip_scores(img, txt), cov, avv, avt, 0.85)
    order = orc.stable_rank(scores)
    sel = orc.merge_select(scores[order], 0.55, 0.95)
    ref = orc.merge_masks(masks, order[sel])
    assert_order_matches(fm.last["order"][0].cpu().numpy(), fm.last["scores"][0].cpu().numpy(), order, scores)
    np.testing.assert_array_equal(pred.cpu().numpy() > 0, ref.numpy() > 0)
    mars.clear()
    assert vva_mod.cost_matrix is None


@pytest.mark.parametrize("name", list(cases.MARS_CASES))
def test_mars_predict_matches_reference_end_to_end(mb, name):
    """The drop-in MARS / VisualVisualAlignmentModule / FilteringMergingModule (EMD solved on the device) against the
    golden output of the reference's own classes run end to end on the same fake backbones (tests/golden/mars_*.npz)."""
    z = np.load(os.path.join(GOLD, f"mars_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    v = cases.VVA_CASES[spec["vva"]]
    c = cases.mars_inputs(spec)
    g, h, cdim, regs = v["g"], v["H"], v["C"], v["regs"]
    ns = c["feat_s"].shape[0]

    class FakeDino(torch.nn.Module):
        embed_dim = cdim

        def __init__(self):
            super().__init__()
            pad = lambda f: torch.cat([torch.full((f.shape[0], 1 + regs, cdim), 7.0), f], dim=1).to(dev())
            self.feats = [pad(c["feat_s"]), pad(c["feat_q"][None])]

        def forward_features(self, imgs):
            return {"x_prenorm": self.feats.pop(0)}

        def get_last_self_attention(self, img):
            return tuple(a.to(dev()) for a in c["attn_maps"])

    class FakeText:
        def get_conceptual_information(self, support_images, support_masks):
            return "thing", spec["description"]

    class FakeVTA:
        def compute(self, query_image, fg_label, bg_labels):
            return c["vta_raw"]

    vva_mod = mb.VisualVisualAlignmentModule(FakeDino(), lambda x: x, 14, g, regs, v["thr"], v["last_n"], dev())
    fm = mb.FilteringMergingModule(None, None, None, alpha=spec["alpha"], static_threshold=spec["static"],
                                   dynamic_threshold=spec["dynamic"], device=dev())
    seen = {}
    orig = fm.compute

    def compute(**kw):
        seen.update(vva=kw["vva"], vta=kw["vta"], text=kw["text"])
        return orig(alphaclip_feats=(c["clip_img"], c["clip_txt"]), **kw)

    fm.compute = compute
    mars = mb.MARS(FakeText(), FakeVTA(), vva_mod, fm)
    pred = mars.predict(torch.zeros(1, ns, 3, h, h), c["support_mask"][None], torch.zeros(1, 3, h, h), c["masks"])
    np.testing.assert_allclose(seen["vva"].reshape(g, g).cpu().numpy(), z["vva"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(seen["vta"].reshape(g, g).cpu().numpy(), z["vta"], rtol=RTOL, atol=1e-6)
    assert seen["text"] == [str(z["text"])]
    assert pred.shape == (h, h) and pred.dtype == torch.float32
    np.testing.assert_array_equal(pred.cpu().numpy() > 0, z["merged"])


@pytest.mark.parametrize("mode", ["topk", "score_filter", "metric_filter"])
def test_matcher_scorer_merge(mb, mode):
    """MatcherScorer against the oracle restatement of matcher/Matcher.py:719-834."""
    h, g, n, k = 140, 10, 18, 80
    masks = cases.blob_masks(n, h, h, seed=55, min_frac=0.01, max_frac=0.15)
    rs = np.random.RandomState(5)
    pts = np.stack([rs.randint(0, h, k), rs.randint(0, h, k)], axis=1)
    emd = torch.rand(n, generator=torch.Generator().manual_seed(6))
    cfg = dict(topk=dict(topk_scores_threshold=0.6), score_filter=dict(score_filter=True, score=0.6, score_norm=0.5),
               metric_filter=dict(coverage=0.05, emd=0.3, purity=0.01))[mode]
    scorer = mb.MatcherScorer(g, alpha=1.0, beta=0.5, exp=1.0, num_merging_mask=6, score_filter_cfg=cfg, device=dev())
    res = scorer.mask_scores(masks, pts, emd)
    merged, score = scorer.merge(res)
    purity, coverage = orc.matcher_mask_scores(masks.numpy() > 0, pts, g)
    scores = orc.matcher_fuse(emd, purity, coverage, 1.0, 0.5, 1.0)
    np.testing.assert_allclose(res["scores"].cpu().numpy(), scores.numpy(), rtol=1e-5)
    full = dict(emd=0.0, purity=0.0, coverage=0.0)
    full.update({k_: v for k_, v in cfg.items() if k_ in full})
    fscores, idx = orc.matcher_metric_filter(scores, dict(purity=purity, coverage=coverage, emd=emd), full)
    if cfg.get("score_filter"):
        chosen, ref_score = orc.matcher_merge_score_filter(fscores, 6, cfg["score"], cfg["score_norm"])
    else:
        chosen, ref_score = orc.matcher_merge_topk(fscores, 6, cfg.get("topk_scores_threshold", 0.0))
    ref = orc.merge_masks(masks, idx.numpy()[chosen])
    np.testing.assert_array_equal(merged[0].cpu().numpy() > 0, ref.numpy() > 0)
    np.testing.assert_allclose(float(score), float(ref_score), rtol=1e-5)
    assert merged.shape == (1, h, h) and merged.dtype == torch.float32


def test_evaluator_dropin(mb):
    g = torch.Generator().manual_seed(18)
    pred = (torch.rand(2, 64, 80, generator=g) < 0.4).float()
    gt = (torch.rand(2, 64, 80, generator=g) < 0.5).float()
    ignore = ((torch.rand(2, 64, 80, generator=g) < 0.1) & (gt == 0)).float()
    mb.Evaluator.initialize()
    inter, union = mb.Evaluator.classify_prediction(pred.to(dev()), dict(query_mask=gt, query_ignore_idx=ignore))
    inter_ref, union_ref = orc.evaluator_areas(pred, gt, ignore)
    np.testing.assert_array_equal(inter.cpu().numpy(), inter_ref.numpy())
    np.testing.assert_array_equal(union.cpu().numpy(), union_ref.numpy())


@pytest.mark.parametrize("name", list(cases.EVAL_CASES))
def test_average_meter_matches_reference(mb, name):
    """Evaluator + AverageMeter drop-ins (device buffers, exact int64 counts) against the reference's golden run;
    two meters fed with disjoint shards sum to the same buffers (what all_reduce does across ranks)."""
    import types

    z = np.load(os.path.join(GOLD, f"eval_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.eval_inputs(spec)
    d = dev()
    ds = types.SimpleNamespace(benchmark=spec["benchmark"], class_ids=spec["class_ids"])
    mb.Evaluator.initialize()
    meter = mb.AverageMeter(ds, device=d)
    shard = [mb.AverageMeter(ds, device=d), mb.AverageMeter(ds, device=d)]
    for i in range(spec["n"]):
        batch = dict(query_mask=c["gt"][i:i + 1])
        if c["ignore"] is not None:
            batch["query_ignore_idx"] = c["ignore"][i:i + 1]
        ai, au = mb.Evaluator.classify_prediction(c["pred"][i:i + 1].to(d), batch)
        np.testing.assert_array_equal(ai[:, 0].cpu().numpy(), z["area_inter"][i])
        np.testing.assert_array_equal(au[:, 0].cpu().numpy(), z["area_union"][i])
        meter.update(ai, au, c["class_id"][i:i + 1], loss=None)
        shard[i % 2].update(ai, au, c["class_id"][i:i + 1], loss=None)
    np.testing.assert_array_equal(meter.intersection_buf.cpu().numpy(), z["intersection_buf"].astype(np.int64))
    np.testing.assert_array_equal(meter.union_buf.cpu().numpy(), z["union_buf"].astype(np.int64))
    assert torch.equal(shard[0].intersection_buf + shard[1].intersection_buf, meter.intersection_buf)
    assert torch.equal(shard[0].union_buf + shard[1].union_buf, meter.union_buf)
    miou, fb, cats = meter.compute_iou()
    np.testing.assert_allclose(float(miou), float(z["miou"]), rtol=1e-5)   # the reference divides in float32
    np.testing.assert_allclose(float(fb), float(z["fb_iou"]), rtol=1e-5)
    np.testing.assert_allclose(cats.cpu().numpy(), z["cats_iou"], rtol=1e-5, atol=1e-7)
    # batched path: all samples in one call
    areas = mb.ops.eval_areas(c["pred"].to(d), c["gt"].to(d), None if c["ignore"] is None else c["ignore"].to(d))
    m2 = mb.AverageMeter(ds, device=d)
    m2.update_areas(areas, c["class_id"])
    assert torch.equal(m2.intersection_buf, meter.intersection_buf) and torch.equal(m2.union_buf, meter.union_buf)
    with pytest.raises(mb.MarsB200Error):
        mb.ops.eval_accumulate(areas, torch.full((spec["n"],), 5000, device=d), m2.intersection_buf, m2.union_buf, check_status=True)


# ------------------------------------------------------------------------------------------ Matcher diagnostics (A13)
@pytest.mark.parametrize("name", list(cases.DIAG_CASES))
def test_matcher_diagnostics(mb, name):
    z = np.load(os.path.join(GOLD, f"diag_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.diag_inputs(spec)
    g = spec["g"]
    pm = mb.PatchMatcher(g, 14, (g * 14, g * 14), dev())
    ref_n, tar_n = orc.normalize_rows(c["ref_raw"]), orc.normalize_rows(c["tar_raw"])
    r2t = pm.get_ref_to_target_similarity(ref_n, tar_n, c["ref_mask"])
    np.testing.assert_allclose(r2t[0].cpu().numpy(), z["ref_to_target"], rtol=RTOL, atol=1e-6)
    S = (ref_n @ tar_n.t())
    st = pm.get_aposteriori_statistics(S, c["ref_mask"], c["tar_mask"], c["ref_raw"], c["tar_raw"])
    got = [st["aposteriori_similarity_mean"], st["aposteriori_similarity_max"], st["aposteriori_similarity_std"],
           st["embeddings_euclidean_distance"]]
    np.testing.assert_allclose(got, z["stats"], rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("backend", [1, 0], ids=["simt", "tcgen05"])
def test_masked_feature_means_all_proposals(mb, backend):
    """Prototypes of every proposal at once ([P, N] x [N, C] on the tensor cores) vs feats[mask].mean(0)."""
    g, p, cdim, h = 14, 20, 96, 196
    masks = cases.blob_masks(p, h, h, 41, 0.01, 0.3)
    masks[3] = 0
    feats = cases.proto_features(g * g, cdim, 42) * 2.0 - 0.3
    d = dev()
    pooled, _, cnt = mb.ops.pool_packed(mb.ops.pack_masks(masks.to(d)), h, h, g)
    got = mb.ops.masked_feature_means(pooled[None], feats.to(d)[None], backend=backend)[0].cpu()
    pm = orc.pool_mask(masks, g).reshape(p, -1)
    for i in range(p):
        if i == 3:
            assert bool(torch.isnan(got[i]).all())
            continue
        np.testing.assert_allclose(got[i].numpy(), feats[pm[i]].mean(dim=0).numpy(), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------ wire format / AMG post-processing
@pytest.mark.parametrize("name", list(cases.AMG_CASES))
def test_amg_postprocessing(mb, name):
    """RLE decode -> packed bits == pack(mask), boxes, stability score and box NMS against the reference's golden."""
    z = np.load(os.path.join(GOLD, f"amg_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.amg_inputs(spec)
    h, w, d = spec["H"], spec["W"], dev()
    bits_ref = mb.ops.pack_masks(c["masks"].to(d))
    bits = mb.ops.rle_decode(torch.from_numpy(z["counts"]).to(d), torch.from_numpy(z["offsets"]).to(d), h, w)
    assert torch.equal(bits, bits_ref)
    np.testing.assert_array_equal(mb.ops.mask_boxes(bits, h, w).cpu().numpy(), z["boxes"])
    score, counts = mb.ops.stability_score(c["logits"].to(d), 0.0, 1.0)
    np.testing.assert_array_equal(score.cpu().numpy(), z["stability"])
    np.testing.assert_array_equal(counts[:, 0].cpu().numpy(), (c["logits"] > 1.0).flatten(1).sum(1).numpy())
    keep, order, mask = mb.ops.box_nms(c["boxes"].to(d), c["scores"].to(d), 0.5)
    np.testing.assert_array_equal(keep.cpu().numpy(), z["nms_keep"])
    with pytest.raises(mb.MarsB200Error):  # counts that do not cover the image
        bad = torch.from_numpy(z["counts"]).clone()
        bad[0] += 1
        mb.ops.rle_decode(bad.to(d), torch.from_numpy(z["offsets"]).to(d), h, w)


def test_rle_ingest_full_size(mb):
    """1024 x 1024 proposals: RLE -> bits equals the float32 ingest bit for bit; boxes from bits match torch."""
    shape = mb.CONFIGS["c2"]
    masks = mb.synthetic.random_masks(24, shape.H, shape.W, torch.Generator(device=dev()).manual_seed(5), dev())
    mcpu = masks.cpu().numpy() > 0
    rles = [orc.mask_to_rle(m) for m in mcpu]
    counts = torch.from_numpy(np.concatenate([np.asarray(r, dtype=np.int32) for r in rles]))
    offsets = torch.from_numpy(np.cumsum([0] + [len(r) for r in rles]).astype(np.int64))
    bits = mb.ops.rle_decode(counts.to(dev()), offsets.to(dev()), shape.H, shape.W)
    assert torch.equal(bits, mb.ops.pack_masks(masks))
    np.testing.assert_array_equal(mb.ops.mask_boxes(bits, shape.H, shape.W).cpu().numpy(), orc.mask_boxes(masks.cpu()).numpy())
    # a larger NMS problem against the oracle
    gen = torch.Generator().manual_seed(9)
    xy = torch.rand(300, 2, generator=gen) * 800
    boxes = torch.cat([xy, xy + torch.rand(300, 2, generator=gen) * 200 + 4], dim=1)
    scores = torch.rand(300, generator=gen)
    keep, _, _ = mb.ops.box_nms(boxes.to(dev()), scores.to(dev()), 0.7)
    np.testing.assert_array_equal(keep.cpu().numpy(), orc.box_nms(boxes, scores, 0.7).numpy())


# ------------------------------------------------------------------------------------------ exact EMD on the device
def _emd_inputs(ns, g, p, h, seed):
    n = g * g
    fs = torch.nn.functional.normalize(cases.proto_features(ns * n, 24, seed), dim=1)
    fq = torch.nn.functional.normalize(cases.proto_features(n, 24, seed + 1), dim=1)
    cost = ((1 - fs @ fq.T) / 2).contiguous()
    support = cases.blob_masks(ns, h, h, seed + 2, 0.03, 0.2)
    masks = cases.blob_masks(p, h, h, seed + 3, 0.01, 0.25)
    return cost, support, masks


@pytest.mark.parametrize("ns,g,p,h", [(1, 7, 9, 100), (2, 10, 12, 140), (1, 12, 6, 168)])
def test_emd_scores_match_exact_lp(mb, ns, g, p, h):
    """Device EMD against the oracle's exact LP (HiGHS) on the same cost sub-matrices."""
    cost, support, masks = _emd_inputs(ns, g, p, h, seed=300 + g)
    masks[p - 1] = 0  # empty proposal: defined as zero transport cost
    d = dev()
    row_fg = mb.ops.pool_mask(support.to(d), g).reshape(1, -1)
    bits = mb.ops.pack_masks(masks.to(d))
    pooled, _, _ = mb.ops.pool_packed(bits, h, h, g)
    got = mb.ops.emd_scores(cost.to(d)[None], row_fg, pooled[None])[0].cpu().numpy()
    sup = orc.pool_mask(support, g).reshape(-1)
    pm = orc.pool_mask(masks, g).reshape(p, -1)
    want = np.asarray([orc.emd_score(sup, pm[i], cost) for i in range(p)])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)
    assert got[p - 1] == 1.0


def test_emd_caps_and_larger_problems(mb):
    """m_cap / t_cap: a tight cap gives the same scores, a cap that is too small is reported; medium-size LPs
    (hundreds of phases, many tied waves) against the oracle's exact LP."""
    ns, g, p, h = 1, 24, 10, 336
    cost, support, masks = _emd_inputs(ns, g, p, h, seed=911)
    d = dev()
    row_fg = mb.ops.pool_mask(support.to(d), g).reshape(1, -1)
    bits = mb.ops.pack_masks(masks.to(d))
    pooled, _, cnt = mb.ops.pool_packed(bits, h, h, g)
    full = mb.ops.emd_scores(cost.to(d)[None], row_fg, pooled[None])[0]
    tight = mb.ops.emd_scores(cost.to(d)[None], row_fg, pooled[None], pooled_count=cnt)[0]
    assert torch.equal(full, tight)
    sup = orc.pool_mask(support, g).reshape(-1)
    pm = orc.pool_mask(masks, g).reshape(p, -1)
    big = sorted(range(p), key=lambda i: -int(cnt[i]))[:3]
    want = np.asarray([orc.emd_score(sup, pm[i], cost) for i in big])
    np.testing.assert_allclose(full.cpu().numpy()[big], want, rtol=0, atol=1e-9)
    # caps size the shared-memory fast path only: problems beyond them are solved by the global-state launch, bit for bit
    # the same values (same solver, other memory), never an error (the reference's ot.emd2 has no capacity limit)
    small_m = mb.ops.emd_scores(cost.to(d)[None], row_fg, pooled[None], m_cap=int(cnt.max()) // 2)[0]
    small_t = mb.ops.emd_scores(cost.to(d)[None], row_fg, pooled[None], t_cap=int(row_fg.sum()) // 3)[0]
    assert torch.equal(small_m, full) and torch.equal(small_t, full)


def test_emd_5shot_large_support_masks_take_the_global_state_path(mb):
    """a 5-shot episode whose objects cover most of the image has more foreground support rows (here ~4700)
    than the shared-memory state can hold (~2400): the drop-in must still score it, like ot.emd2 does
    (FilteringMergingModule.py:142-169).  Values against the oracle's exact LP on two proposals."""
    ns, g, h, p = 5, 37, 148, 6
    gen = torch.Generator().manual_seed(77)
    n = g * g
    fs = orc.normalize_rows(torch.randn(ns * n, 32, generator=gen))
    fq = orc.normalize_rows(torch.randn(n, 32, generator=gen))
    _, cost = orc.similarity_and_cost(fs, fq)
    support = torch.zeros(ns, h, h)
    support[:, 10:140, 8:132] = 1.0  # ~70 % of every shot
    masks = cases.blob_masks(p, h, h, seed=5, min_frac=0.002, max_frac=0.02)
    d = dev()
    row_fg = mb.ops.pool_mask(support.to(d), g).reshape(1, -1)
    t_fg = int(row_fg.sum())
    assert t_fg > 2600
    bits = mb.ops.pack_masks(masks.to(d))
    pooled, _, cnt = mb.ops.pool_packed(bits, h, h, g)
    got = mb.ops.emd_scores(cost.to(d)[None], row_fg, pooled[None], pooled_count=cnt)[0].cpu().numpy()
    assert np.isfinite(got).all()
    sup = orc.pool_mask(support, g).reshape(-1)
    pm = orc.pool_mask(masks, g).reshape(p, -1)
    small = sorted(range(p), key=lambda i: int(cnt[i]))[:2]
    want = np.asarray([orc.emd_score(sup, pm[i], cost) for i in small])
    np.testing.assert_allclose(got[small], want, rtol=0, atol=1e-9)
    # the same through the drop-in module (it used to raise 't_cap + N too large for the shared-memory state')
    fm = mb.FilteringMergingModule(alpha_clip_model=None, img_transforms=None, mask_transforms=None, alpha=0.85,
                                   static_threshold=0.55, dynamic_threshold=0.95, device=d)
    gen2 = torch.Generator().manual_seed(3)
    img = torch.nn.functional.normalize(torch.randn(p, 16, generator=gen2), dim=1)
    txt = torch.nn.functional.normalize(torch.randn(16, generator=gen2), dim=0)
    ranked = fm._score_proposals(torch.zeros(1, 3, h, h), masks, support[None], cost, g, torch.rand(g, g, generator=gen2),
                                 torch.rand(g, g, generator=gen2), ["x"], alphaclip_feats=(img, txt))
    assert len(ranked) == p and all(np.isfinite(float(sc)) for _, sc in ranked)


def test_emd_tight_caps_more_sources_than_sinks(mb):
    """5-shot-like shape: more support rows than proposal patches, caps read from the data (t_cap > m_cap)."""
    ns, g, p, h = 3, 12, 8, 168
    cost, support, masks = _emd_inputs(ns, g, p, h, seed=77)
    support = cases.blob_masks(ns, h, h, 78, 0.25, 0.45)   # large support masks: T around 200 of 432 rows
    d = dev()
    row_fg = mb.ops.pool_mask(support.to(d), g).reshape(1, -1)
    bits = mb.ops.pack_masks(masks.to(d))
    pooled, _, cnt = mb.ops.pool_packed(bits, h, h, g)
    assert int(row_fg.sum()) > int(cnt.max())
    got = mb.ops.emd_scores(cost.to(d)[None], row_fg, pooled[None], pooled_count=cnt)[0].cpu().numpy()
    sup = orc.pool_mask(support, g).reshape(-1)
    pm = orc.pool_mask(masks, g).reshape(p, -1)
    want = np.asarray([orc.emd_score(sup, pm[i], cost) for i in range(p)])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)


def test_emd_duplicate_proposals_share_one_lp(mb):
    """Proposals of an episode with the same pooled bitmap (exact copies, one-pixel shifts, two empty masks) are one LP:
    solved once, copied to the others - same values as solving each, per episode (a copy in ANOTHER episode has another
    cost matrix and is not linked)."""
    ns, g, p, h = 1, 12, 12, 168
    d = dev()
    costs, fgs, pooleds, wants = [], [], [], []
    for e in range(2):
        cost, support, masks = _emd_inputs(ns, g, p, h, seed=640 + e)
        masks[5] = masks[2]
        masks[9] = masks[2]
        masks[7] = torch.roll(masks[4], 1, dims=1)  # usually the same pooled bitmap, sometimes not: both are fine
        masks[3] = 0
        masks[11] = 0
        sup = orc.pool_mask(support, g).reshape(-1)
        pm = orc.pool_mask(masks, g).reshape(p, -1)
        wants.append([1.0 if not pm[i].any() else orc.emd_score(sup, pm[i], cost) for i in range(p)])
        costs.append(cost)
        fgs.append(mb.ops.pool_mask(support.to(d), g).reshape(-1))
        pooleds.append(mb.ops.pool_packed(mb.ops.pack_masks(masks.to(d)), h, h, g)[0])
    got = mb.ops.emd_scores(torch.stack(costs).to(d), torch.stack(fgs), torch.stack(pooleds)).cpu().numpy()
    np.testing.assert_allclose(got, np.asarray(wants), rtol=0, atol=1e-9)
    for e in range(2):
        assert got[e, 5] == got[e, 2] == got[e, 9] and got[e, 3] == got[e, 11] == 1.0
    assert got[0, 2] != got[1, 2]


@pytest.mark.parametrize("t,m", [(6, 4), (4, 6), (9, 6), (10, 4), (7, 5), (12, 8), (15, 10), (8, 12), (20, 3), (1, 7)])
def test_emd_rectangular_equals_expanded_assignment(mb, t, m):
    """Device EMD against an INDEPENDENT exact solver on rectangular problems: the T x M uniform-marginal LP expanded to
    an lcm(T, M)^2 assignment and solved by scipy's LSAP (no LP solver, no shared code with the oracle's HiGHS path)."""
    import math

    from scipy.optimize import linear_sum_assignment

    g = 5
    n = g * g
    rs = np.random.RandomState(1000 * t + m)
    cost = ((1 - rs.uniform(-0.2, 0.9, size=(n, n))) / 2).astype(np.float32)
    rows = np.sort(rs.choice(n, t, replace=False))
    cols = np.sort(rs.choice(n, m, replace=False))
    row_fg = torch.zeros(1, n, dtype=torch.uint8)
    row_fg[0, rows] = 1
    word = np.zeros(1, dtype=np.uint32)
    for c in cols:
        word[0] |= np.uint32(1) << np.uint32(c)
    pooled = torch.from_numpy(word.view(np.int32)).reshape(1, 1, 1)
    got = float(mb.ops.emd_scores(torch.from_numpy(cost).to(dev())[None], row_fg.to(dev()), pooled.to(dev()))[0, 0])
    big = math.lcm(t, m)
    sub = cost[rows][:, cols].astype(np.float64)
    exp = np.repeat(np.repeat(sub, big // t, axis=0), big // m, axis=1)
    r, c = linear_sum_assignment(exp)
    want = 1.0 - exp[r, c].sum() / big
    assert abs(got - want) < 1e-12


def test_emd_c2_full_size_against_network_simplex(mb):
    """Device EMD of a whole c2 episode (P = 256 LPs, T ~ 400 support patches) against a NETWORK SIMPLEX - the algorithm
    class of the reference's ot.emd2 (FilteringMergingModule.py:162-166) - on the largest, the median and the smallest
    non-empty proposals; integer arithmetic on the host side, so the comparison is exact up to float64 rounding."""
    pytest.importorskip("networkx")
    shape = mb.CONFIGS["c2"]
    b = mb.stack_episodes([mb.make_episode(shape, 40, dev())])
    n, g = shape.N, shape.g
    fs = mb.ops.normalize_rows(b["feat_s"].reshape(1, n, shape.C))
    fq = mb.ops.normalize_rows(b["feat_q"])
    row_fg = mb.ops.pool_mask(b["support_mask"], g).reshape(1, n)
    cost = mb.ops.sim_contract(fs, fq, n, n, shape.C, want_sim=False, want_cost=True)["cost"]
    pooled, _, cnt = mb.ops.pool_packed(mb.ops.pack_masks(b["masks"]), shape.H, shape.W, g)
    got = mb.ops.emd_scores(cost, row_fg, pooled)[0].cpu().numpy()
    c_host = cost[0].cpu().numpy()
    rows = row_fg[0].cpu().numpy().astype(bool)
    order = np.argsort(-cnt[0].cpu().numpy(), kind="stable")
    nonempty = [int(p) for p in order if int(cnt[0, p]) > 0]
    picks = [nonempty[0], nonempty[1], nonempty[len(nonempty) // 2], nonempty[-1]]
    words = pooled[0].cpu().numpy().view(np.uint32)
    for p in picks:
        cols = np.array([(words[p, j >> 5] >> np.uint32(j & 31)) & 1 for j in range(n)], dtype=bool)
        sub = c_host[rows][:, cols]
        assert sub.min() >= 2.0 ** -5 and sub.max() < 1.0
        want = 1.0 - orc.emd_network_simplex(sub)
        assert abs(got[p] - want) < 1e-9, (p, sub.shape, got[p], want)


def test_emd_square_case_equals_assignment(mb):
    """T == M: the transport LP is an assignment problem; compare with scipy's exact LSAP at a larger size."""
    from scipy.optimize import linear_sum_assignment

    g, n_sel = 20, 180
    n = g * g
    gen = torch.Generator().manual_seed(77)
    cost = torch.rand(n, n, generator=gen)
    rows = torch.randperm(n, generator=gen)[:n_sel].sort().values
    cols = torch.randperm(n, generator=gen)[:n_sel].sort().values
    row_fg = torch.zeros(1, n, dtype=torch.uint8)
    row_fg[0, rows] = 1
    pooled_bits = np.zeros((1, (n + 31) // 32 * 32), dtype=bool)
    pooled_bits[0, cols.numpy()] = True
    pooled = torch.from_numpy(np.packbits(pooled_bits.reshape(1, -1, 32), axis=-1, bitorder="little").view(np.int32).reshape(1, 1, -1))
    got = float(mb.ops.emd_scores(cost.to(dev())[None], row_fg.to(dev()), pooled.to(dev()))[0, 0])
    sub = cost[rows][:, cols].double().numpy()
    r, c = linear_sum_assignment(sub)
    want = 1.0 - sub[r, c].sum() / n_sel
    assert abs(got - want) < 1e-9


def test_emd_feeds_the_fused_ranking(mb):
    """Full FilteringMerging drop-in with the device EMD against the reference's golden scores (host LP there)."""
    z = np.load(os.path.join(GOLD, "fm_g10_dynamic.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.fm_inputs(spec)
    d = dev()
    g, h = spec["g"], spec["H"]
    row_fg = mb.ops.pool_mask(c["support_mask"].to(d), g).reshape(1, -1)
    bits = mb.ops.pack_masks(c["masks"].to(d))
    pooled, _, _ = mb.ops.pool_packed(bits, h, h, g)
    emd = mb.ops.emd_scores(c["cost"].to(d)[None], row_fg, pooled[None])[0]
    np.testing.assert_allclose(emd.cpu().numpy(), z["emd"], rtol=0, atol=1e-9)
    mod = mb.FilteringMergingModule(None, None, None, spec["alpha"], spec["static"], spec["dynamic"], d)
    ranked = mod._score_proposals(torch.zeros(1, 3, h, h), c["masks"], c["support_mask"][None], c["cost"].to(d), g,
                                  c["vva"], c["vta"], ["x"], emd_scores=emd.cpu().numpy(),
                                  alphaclip_feats=(c["clip_img"], c["clip_txt"]))
    np.testing.assert_allclose([s for _, s in ranked], z["scores"], rtol=RTOL)


def test_engine_with_device_emd(mb):
    """Whole episode with the EMD solved on the device, against the oracle running its exact LP per proposal."""
    shape = mb.EpisodeShape(ns=1, g=8, C=48, P=10, H=112, W=112, gt=6, D=24)
    cfg = mb.RankingConfig(nms_iou_threshold=0.7, emd_on_device=True)
    eps = [mb.make_episode(shape, 70 + i, "cpu", with_emd=False) for i in range(2)]
    eng = mb.RankingEngine(shape, 2, cfg, dev())
    out = eng.run(mb.to_device(mb.stack_episodes(eps), dev()))
    torch.cuda.synchronize()
    ocfg = dict(g=shape.g, vva_box_threshold=cfg.vva_box_threshold, vta_box_threshold=cfg.vta_box_threshold,
                alpha=cfg.alpha, static_threshold=cfg.static_threshold, dynamic_threshold=cfg.dynamic_threshold,
                nms_iou_threshold=cfg.nms_iou_threshold)
    for e, ep in enumerate(eps):
        ref = orc.run_episode(ep, ocfg, emd_fn=orc.emd_score)
        assert_order_matches(out["order"][e].cpu().numpy(), out["scores"][e].cpu().numpy(), ref["order"], ref["scores"])


def test_filtering_merging_default_device_emd(mb):
    """The drop-in's default EMD path (no POT, no emd_fn) reproduces the reference's golden ranking."""
    z = np.load(os.path.join(GOLD, "fm_g7_overlap.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.fm_inputs(spec)
    mod = mb.FilteringMergingModule(None, None, None, spec["alpha"], spec["static"], spec["dynamic"], dev())
    ranked = mod._score_proposals(torch.zeros(1, 3, 100, 100), c["masks"], c["support_mask"][None], c["cost"].to(dev()),
                                  spec["g"], c["vva"], c["vta"], ["x"], alphaclip_feats=(c["clip_img"], c["clip_txt"]))
    np.testing.assert_allclose([s for _, s in ranked], z["scores"], rtol=RTOL)


def test_match_argmax_topk_and_mutual(mb):
    """Warp-shuffle row top-k / column arg-max against torch, and the mutual-NN retain rule."""
    gen = torch.Generator().manual_seed(21)
    e, m, n, k = 2, 150, 137, 5
    sim = torch.rand(e, m, n, generator=gen)
    sim[0, 3, 10] = sim[0, 3, 50] = 2.0       # tie in a row -> lowest column first
    sim[1, 7, 20] = sim[1, 90, 20] = 3.0      # tie in a column -> lowest row
    mask = (torch.rand(e, m, generator=gen) < 0.4).to(torch.uint8)
    res = mb.ops.match_argmax(sim.to(dev()), k=k, row_mask=mask.to(dev()))
    tv, ti = torch.sort(sim, dim=2, descending=True, stable=True)
    np.testing.assert_array_equal(res["row_idx"].cpu().numpy(), ti[:, :, :k].numpy())
    np.testing.assert_array_equal(res["row_vals"].cpu().numpy(), tv[:, :, :k].numpy())
    masked = sim.masked_fill(mask[:, :, None] == 0, float("-inf"))
    cv = masked.max(dim=1).values
    ci = (masked == cv[:, None, :]).float().argmax(dim=1)  # first row attaining the max
    np.testing.assert_array_equal(res["col_vals"].cpu().numpy(), cv.numpy())
    np.testing.assert_array_equal(res["col_idx"].cpu().numpy(), ci.numpy())
    rows, q, val = mb.ops.mutual_matches(sim[0].to(dev()), mask[0].to(dev()))
    fg = torch.nonzero(mask[0]).flatten()
    fwd = sim[0].argmax(dim=1)
    back = sim[0].argmax(dim=0)
    keep = mask[0][back[fwd[fg]]] != 0
    np.testing.assert_array_equal(rows.cpu().numpy(), fg[keep].numpy())
    np.testing.assert_array_equal(q.cpu().numpy(), fwd[fg][keep].numpy())


@pytest.mark.parametrize("r,c,maximize", [(40, 90, True), (90, 40, True), (64, 64, False), (300, 1369, True), (1, 5, True)])
def test_lsap_matches_scipy(mb, r, c, maximize):
    """Exact assignment against scipy.optimize.linear_sum_assignment (the solver Matcher calls, Matcher.py:449,471)."""
    from scipy.optimize import linear_sum_assignment

    gen = torch.Generator().manual_seed(r * 7 + c)
    sim = torch.rand(2, r, c, generator=gen)
    row_sel = (torch.rand(2, r, generator=gen) < 0.7).to(torch.uint8)
    row_sel[:, 0] = 1
    r2c, obj = mb.ops.lsap(sim.to(dev()), row_sel=row_sel.to(dev()), maximize=maximize)
    r2c, obj = r2c.cpu().numpy(), obj.cpu().numpy()
    for e in range(2):
        rows = np.nonzero(row_sel[e].numpy())[0]
        sub = sim[e][rows].double().numpy()
        ri, ci = linear_sum_assignment(sub, maximize=maximize)
        assert abs(obj[e] - sub[ri, ci].sum()) < 1e-9
        got = r2c[e]
        assigned = got[got >= 0]
        assert len(set(assigned.tolist())) == len(assigned) == min(len(rows), c)
        assert (got[np.setdiff1d(np.arange(r), rows)] == -1).all()
        expect = np.full(r, -1)
        expect[rows[ri]] = ci
        np.testing.assert_array_equal(got, expect)  # random costs: the optimum is unique


@pytest.mark.parametrize("r,c,maximize,sel", [(130, 128, True, False), (128, 131, False, False), (300, 300, True, False),
                                               (64, 64, True, False), (70, 66, True, True), (1374, 1369, True, False)])
def test_lsap_near_square_matches_scipy(mb, r, c, maximize, sel):
    """Near-square problems take the multi-source phase solver on the zero-padded square (lsap_square_kernel; the 5-shot
    forward matching of Matcher.py:449 is 1374 x 1369): assignment and objective against scipy."""
    from scipy.optimize import linear_sum_assignment

    gen = torch.Generator().manual_seed(r * 11 + c)
    sim = torch.rand(2, r, c, generator=gen)
    row_sel = torch.ones(2, r, dtype=torch.uint8)
    col_sel = torch.ones(2, c, dtype=torch.uint8)
    if sel:
        row_sel[:, 5] = 0
        col_sel[:, 7] = 0
        col_sel[1, 9] = 0
    r2c, obj = mb.ops.lsap(sim.to(dev()), row_sel=row_sel.to(dev()), col_sel=col_sel.to(dev()), maximize=maximize)
    r2c, obj = r2c.cpu().numpy(), obj.cpu().numpy()
    for e in range(2):
        rows = np.nonzero(row_sel[e].numpy())[0]
        cols = np.nonzero(col_sel[e].numpy())[0]
        sub = sim[e][rows][:, cols].double().numpy()
        ri, ci = linear_sum_assignment(sub, maximize=maximize)
        assert abs(obj[e] - sub[ri, ci].sum()) < 1e-9
        expect = np.full(r, -1)
        expect[rows[ri]] = cols[ci]
        np.testing.assert_array_equal(r2c[e], expect)  # random costs: the optimum is unique


@pytest.mark.parametrize("kind", ["all_equal", "few_levels", "duplicate_rows"])
def test_lsap_near_square_with_ties(mb, kind):
    """Ties everywhere (equal-distance waves settle many sinks at once, augmentation order is arbitrary): the near-square
    solver must still return A valid optimal assignment - compared with scipy by objective, like any LSAP with ties."""
    from scipy.optimize import linear_sum_assignment

    r, c = 150, 146
    gen = torch.Generator().manual_seed(3)
    if kind == "all_equal":
        sim = torch.full((1, r, c), 0.25)
    elif kind == "few_levels":
        sim = torch.randint(0, 4, (1, r, c), generator=gen).float() / 4
    else:
        base = torch.rand(1, 10, c, generator=gen)
        sim = base[:, torch.randint(0, 10, (r,), generator=gen)]
    r2c, obj = mb.ops.lsap(sim.to(dev()), maximize=True)
    got = r2c[0].cpu().numpy()
    assigned = got[got >= 0]
    assert len(assigned) == min(r, c) and len(set(assigned.tolist())) == len(assigned)
    ri, ci = linear_sum_assignment(sim[0].double().numpy(), maximize=True)
    want = sim[0].double().numpy()[ri, ci].sum()
    rows = np.nonzero(got >= 0)[0]
    assert abs(sim[0].double().numpy()[rows, got[rows]].sum() - want) < 1e-9
    assert abs(float(obj[0]) - want) < 1e-9


def test_bidirectional_lsap_matching(mb):
    """Forward + reverse assignment and the retain rule of Matcher.patch_level_matching (Matcher.py:443-477)."""
    from scipy.optimize import linear_sum_assignment

    spec = cases.VVA_CASES["g10_2shot"]
    c = cases.vva_inputs(spec)
    g = spec["g"]
    fs, fq = orc.normalize_rows(c["feat_s"]), orc.normalize_rows(c["feat_q"])
    sim = fs @ fq.T                                   # [ns*N, N]
    mask = orc.pool_mask(c["support_mask"], g).reshape(-1)
    # reference restatement with scipy
    s_fwd = sim[mask]
    fr, fc = linear_sum_assignment(s_fwd.numpy(), maximize=True)
    s_rev = sim.t()[fc]
    rr, rc = linear_sum_assignment(s_rev.numpy(), maximize=True)
    mask_idx = torch.nonzero(mask).flatten().numpy()
    retain_ref = np.isin(rc, mask_idx)
    # device
    d = dev()
    sim_d = sim.to(d)
    r2c, _ = mb.ops.lsap(sim_d, row_sel=mask.to(torch.uint8).to(d))
    fwd_cols = r2c[0][torch.nonzero(mask.to(d)).flatten()].long()
    np.testing.assert_array_equal(fwd_cols.cpu().numpy(), fc)
    sel = torch.zeros(sim.shape[1], dtype=torch.uint8, device=d)
    sel[fwd_cols] = 1
    q2s, _ = mb.ops.lsap(sim_d.t().contiguous(), row_sel=sel)
    # scipy's reverse problem has rows in forward-match order; map by query patch
    ref_by_patch = {int(fc[k]): int(rc[k]) for k in range(len(fc))}
    got_by_patch = {int(p): int(q2s[0][p]) for p in fc}
    assert ref_by_patch == got_by_patch
    retain = np.isin(np.asarray([got_by_patch[int(p)] for p in fc]), mask_idx)
    np.testing.assert_array_equal(retain, retain_ref)


@pytest.mark.parametrize("name", ["g10_2shot", "g37_1shot"])
def test_patch_matcher_against_scipy_restatement(mb, name):
    """PatchMatcher (device LSAP) against the oracle restatement of Matcher.patch_level_matching."""
    spec = cases.VVA_CASES[name]
    c = cases.vva_inputs(spec)
    g, h = spec["g"], spec["H"]
    fs, fq = orc.normalize_rows(c["feat_s"]), orc.normalize_rows(c["feat_q"])
    pool = orc.pool_mask(c["support_mask"], g).reshape(-1).float()
    pts_ref, neg_ref, reduced_ref = orc.matcher_patch_matching(fs, fq, pool, g, 14, (h, h))
    pm = mb.PatchMatcher(g, 14, (h, h), dev())
    res = pm.match(fs, fq, pool)
    assert res["reduced_points_num"] == reduced_ref
    got = sorted(map(tuple, res["points"].cpu().tolist()))
    got_neg = sorted(map(tuple, res["points_discarded"].cpu().tolist()))
    assert got == pts_ref
    assert got_neg == neg_ref
    np.testing.assert_allclose(res["C"].cpu().numpy(), ((1 - fs @ fq.T) / 2).numpy(), atol=3e-6)


def test_engine_wire_formats(mb):
    """The engine ranks identically from float32 masks, uncompressed RLE and pre-packed bits."""
    shape = mb.EpisodeShape(ns=1, g=12, C=64, P=16, H=160, W=160, gt=9, D=32)
    cfg = mb.RankingConfig(nms_iou_threshold=0.7)
    eps = [mb.make_episode(shape, 90 + i) for i in range(2)]
    batch = mb.to_device(mb.stack_episodes(eps), dev())
    eng = mb.RankingEngine(shape, 2, cfg, dev())
    ref = {k: v.clone() for k, v in eng.run(batch).items() if k in ("order", "scores", "flags", "merged_bits", "inter")}
    counts, offsets = mb.masks_to_rle(batch["masks"].reshape(-1, shape.H, shape.W))
    rle = {k: v for k, v in batch.items() if k != "masks"}
    rle["mask_rle_counts"], rle["mask_rle_offsets"] = counts.to(dev()), offsets.to(dev())
    packed = {k: v for k, v in batch.items() if k != "masks"}
    packed["mask_bits"] = mb.ops.pack_masks(batch["masks"])
    for other in (rle, packed):
        out = eng.run(other)
        for k, v in ref.items():
            assert torch.equal(out[k], v), k


@pytest.mark.parametrize("case", ["single", "all_equal", "empty_proposal", "all_empty"])
def test_fuse_rank_degenerate_inputs(mb, case):
    """Edge cases the reference leaves to arithmetic: P = 1 (both min-max terms vanish), all scores
    equal (order = index order), an all-zero proposal (0 / 1e-7 = 0 alignment), no proposal pixel at all."""
    g, h = 8, 64
    n = g * g
    p = 1 if case == "single" else 6
    masks = cases.blob_masks(p, h, h, 11, 0.05, 0.3)
    if case == "all_equal":
        masks[:] = masks[0]
    if case == "empty_proposal":
        masks[2] = 0
    if case == "all_empty":
        masks[:] = 0
    rs = np.random.RandomState(3)
    vva = rs.rand(g, g).astype(np.float32)
    vta = rs.rand(g, g).astype(np.float32)
    emd = np.full(p, 0.4) if case in ("all_equal", "all_empty") else rs.rand(p)
    clip = np.full(p, 0.2, dtype=np.float32) if case in ("all_equal", "all_empty") else rs.rand(p).astype(np.float32)
    pooled_ref, cov, avv, avt = orc.region_scores(masks, vva, vta, g)
    want = orc.fuse_scores(emd, clip, cov, avv, avt, 0.85)
    d = dev()
    bits = mb.ops.pack_masks(masks.to(d))[None]
    pooled, area, cnt = mb.ops.pool_packed(bits, h, h, g)
    sv, st, uc = mb.ops.region_sums(pooled, torch.from_numpy(vva).to(d).reshape(1, n), torch.from_numpy(vta).to(d).reshape(1, n))
    res = mb.ops.fuse_rank(torch.from_numpy(emd).to(d).reshape(1, p), torch.from_numpy(clip).to(d).reshape(1, p), cnt, sv, st,
                           uc, None, 0.85, 0.55, 0.95, None)
    np.testing.assert_allclose(res["scores"][0].cpu().numpy(), want, rtol=RTOL, atol=1e-9)
    assert_order_matches(res["order"][0].cpu().numpy(), res["scores"][0].cpu().numpy(), orc.stable_rank(want), want)
    if case in ("all_equal", "all_empty"):
        np.testing.assert_array_equal(res["order"][0].cpu().numpy(), np.arange(p))


def test_emd_repeatable_and_orientation_free(mb):
    """The solver's atomics reorder work between runs, the optimum must not move (a race would): three runs of
    64 LPs agree to 1e-12, and the transposed problem (roles of support rows and proposal patches exchanged, which
    flips the solver's source / sink orientation) has the same value."""
    ns, g, p, h = 1, 24, 64, 336
    cost, support, masks = _emd_inputs(ns, g, p, h, seed=4242)
    d = dev()
    n = g * g
    row_fg = mb.ops.pool_mask(support.to(d), g).reshape(1, -1)
    pooled, _, cnt = mb.ops.pool_packed(mb.ops.pack_masks(masks.to(d)), h, h, g)
    runs = [mb.ops.emd_scores(cost.to(d)[None], row_fg, pooled[None], pooled_count=cnt)[0] for _ in range(3)]
    assert float((runs[0] - runs[1]).abs().max()) < 1e-12 and float((runs[0] - runs[2]).abs().max()) < 1e-12
    # transposed: cost^T, "support" = patches of proposal i, "proposal" = the support rows
    sup_bits = mb.ops.pack_masks(row_fg.reshape(1, 1, n))[:, :(n + 31) // 32].contiguous()[None]   # [1, 1, npw]
    for i in sorted(range(p), key=lambda k: -int(cnt[k]))[:6] + sorted(range(p), key=lambda k: int(cnt[k]))[:6]:
        rows_i = ((pooled[i][:, None] >> torch.arange(32, device=d, dtype=torch.int32)) & 1).reshape(-1)[:n].to(torch.uint8)
        flipped = mb.ops.emd_scores(cost.t().contiguous().to(d)[None], rows_i[None], sup_bits)[0, 0]
        assert abs(float(flipped) - float(runs[0][i])) < 1e-12


def test_patch_matcher_5shot_full_size(mb):
    """c3 shape: 5 x 1369 support patches against 1369 query patches.  The reverse assignment is 1369 x 6845, the
    largest problem the solver's shared-memory state has to hold; compared with the scipy restatement."""
    shape = mb.EpisodeShape(ns=5, g=37, C=256, P=4, H=518, W=518)
    ep = mb.make_episode(shape, 21)
    n = shape.N
    fs = orc.normalize_rows(ep["feat_s"].reshape(5 * n, shape.C))
    fq = orc.normalize_rows(ep["feat_q"])
    pool = orc.pool_mask(ep["support_mask"], shape.g).reshape(-1).float()
    pts_ref, neg_ref, reduced_ref = orc.matcher_patch_matching(fs, fq, pool, shape.g, 14, (518, 518))
    res = mb.PatchMatcher(shape.g, 14, (518, 518), dev()).match(fs, fq, pool)
    assert res["reduced_points_num"] == reduced_ref
    assert sorted(map(tuple, res["points"].cpu().tolist())) == pts_ref
    assert sorted(map(tuple, res["points_discarded"].cpu().tolist())) == neg_ref


@pytest.mark.parametrize("cover", [0.2, 0.6, 1.0])
def test_patch_matcher_concurrent_reverse(mb, cover):
    """With T >= N masked support patches the forward assignment matches every query patch, so PatchMatcher solves the reverse
    problem beside the forward one on a second stream: same points as the sequential schedule and as the scipy restatement.
    cover = 0.2 gives T < N (sequential either way)."""
    g, ns, c = 12, 3, 48
    n = g * g
    gen = torch.Generator().manual_seed(int(cover * 100))
    protos = torch.randn(6, c, generator=gen)
    fs = orc.normalize_rows(0.6 * protos[torch.randint(0, 6, (ns * n,), generator=gen)] + 0.8 * torch.randn(ns * n, c, generator=gen))
    fq = orc.normalize_rows(0.6 * protos[torch.randint(0, 6, (n,), generator=gen)] + 0.8 * torch.randn(n, c, generator=gen))
    pool = (torch.rand(ns * n, generator=gen) < cover).float() if cover < 1.0 else torch.ones(ns * n)
    t = int(pool.sum())
    assert (t >= n) == (cover > 0.2)
    pts_ref, neg_ref, reduced_ref = orc.matcher_patch_matching(fs, fq, pool, g, 14, (g * 14, g * 14))
    res = [mb.PatchMatcher(g, 14, (g * 14, g * 14), dev(), concurrent_reverse=cr).match(fs, fq, pool) for cr in (True, False)]
    for r in res:
        assert r["reduced_points_num"] == reduced_ref
        assert sorted(map(tuple, r["points"].cpu().tolist())) == pts_ref
        assert sorted(map(tuple, r["points_discarded"].cpu().tolist())) == neg_ref
    assert torch.equal(res[0]["retain"], res[1]["retain"])
    assert torch.equal(res[0]["indices_forward"][1], res[1]["indices_forward"][1])


def test_engine_check_status_raises_the_library_error(mb):
    """RankingEngine.check_status reports non-finite fused scores as MarsB200Error (ops re-exports the error type)."""
    assert mb.ops.MarsB200Error is mb.MarsB200Error
    shape = mb.EpisodeShape(ns=1, g=10, C=32, P=8, H=64, W=64, gt=8, D=16)
    batch = mb.stack_episodes([mb.make_episode(shape, 5, dev())])
    batch["emd"] = batch["emd"].clone()
    batch["emd"][0, 3] = float("nan")
    eng = mb.RankingEngine(shape, 1, mb.RankingConfig(), dev())
    eng.run(batch)
    with pytest.raises(mb.MarsB200Error):
        eng.check_status()


# ------------------------------------------------------------------ SM partitions (CUDA green contexts)
def test_stream_sm_count_follows_the_partition(mb):
    """Persistent kernels size their grids from the stream: whole device on a plain stream, the partition's SMs on a
    green-context stream."""
    from marsb200.partition import SmPartition, stream_sm_count

    total = torch.cuda.get_device_properties(dev()).multi_processor_count
    assert stream_sm_count(torch.cuda.current_stream()) == total
    assert stream_sm_count(torch.cuda.Stream()) == total
    part = SmPartition(dev(), 56)
    try:
        assert part.tensor_sms >= 56 and part.tensor_sms + part.hbm_sms <= total
        assert stream_sm_count(part.tensor_stream) == part.tensor_sms
        assert stream_sm_count(part.hbm_stream) == part.hbm_sms
        assert stream_sm_count(part.extra_stream("tensor")) == part.tensor_sms
    finally:
        part.close()


def test_stream_sm_cap(mb):
    """marsb200_stream_set_sm_cap: a capped stream reports the cap (never more than the device), other streams are not
    affected, 0 removes it; the engine leaves no cap behind on the (pooled) stream handles it used."""
    from marsb200.partition import set_stream_sm_cap, stream_sm_count

    total = torch.cuda.get_device_properties(dev()).multi_processor_count
    a, b = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
    set_stream_sm_cap(a, 48)
    try:
        assert stream_sm_count(a) == 48 and stream_sm_count(b) == total
        set_stream_sm_cap(a, 32)
        assert stream_sm_count(a) == 32
        set_stream_sm_cap(a, 10 * total)
        assert stream_sm_count(a) == total
    finally:
        set_stream_sm_cap(a, 0)
    assert stream_sm_count(a) == total
    with pytest.raises(mb.MarsB200Error):
        set_stream_sm_cap(a, -1)
    shape = mb.EpisodeShape(ns=1, g=14, gt=12, C=64, D=32, P=40, H=96, W=96)
    eng = mb.RankingEngine(shape, 1, mb.RankingConfig(nms_iou_threshold=0.6), dev())
    assert eng._contraction_sms == 40
    eng.run(mb.stack_episodes([mb.make_episode(shape, 3, dev())]))
    torch.cuda.synchronize()
    for st in (eng._hi, eng._side3, eng._side4):
        assert stream_sm_count(st) == total
    assert mb.RankingEngine(shape, 4, mb.RankingConfig(), dev())._contraction_sms == 0


@pytest.mark.parametrize("chunks,vta_on_hbm,tail,extra", [
    (1, True, 0, {}), (2, False, 1, {}), (3, True, 2, {}),
    (2, False, 0, dict(partition_prep_on_hbm=True)),          # streaming preparation (staged PIR) on the hbm partition
    (3, True, 0, dict(partition_prep_on_hbm=True)),
    (2, False, 0, dict(partition_pool_side_stream=True)),     # pooling beside the next chunk's ingest
    (2, False, 0, dict(partition_pool_on_tensor=True)),
    (2, False, 0, dict(fused_pool=True)),                     # one-pass pack + pool in the hbm partition
])
def test_engine_partitioned_matches_one_timeline(mb, chunks, vta_on_hbm, tail, extra):
    """The two-partition schedule (ingest beside the contractions) and every optional placement of its pieces are the same
    arithmetic: every output bit for bit."""
    shape = mb.EpisodeShape(ns=1, g=12, C=64, P=40, H=160, W=160, gt=9, D=32)
    eps = [mb.make_episode(shape, 300 + i) for i in range(5)]
    batch = mb.to_device(mb.stack_episodes(eps), dev())
    base = mb.RankingEngine(shape, 5, mb.RankingConfig(nms_iou_threshold=0.7), dev())
    ref = {k: v.clone() for k, v in base.run(batch).items() if v is not None}
    cfg = mb.RankingConfig(nms_iou_threshold=0.7, tensor_partition_sms=56, partition_chunks=chunks,
                           partition_vta_on_hbm=vta_on_hbm, partition_pairwise_tail=tail, **extra)
    eng = mb.RankingEngine(shape, 5, cfg, dev())
    try:
        for _ in range(2):
            out = eng.run(batch)
            torch.cuda.synchronize()
            for k, v in ref.items():
                assert torch.equal(out[k], v), k
    finally:
        eng._part.close()


def test_pir_stages_equal_the_single_call(mb):
    """marsb200_pir_stages: normalise / contract / apply run separately (in that order, on one workspace) give the same
    bits as marsb200_pir_refine."""
    g, e = 12, 3
    n = g * g
    gen = torch.Generator().manual_seed(12)
    attn = torch.softmax(2.0 * torch.randn(e, n, n, generator=gen), -1).to(dev())
    prior = torch.rand(e, n, generator=gen).to(dev())
    whole = mb.ops.pir_refine(prior, attn, g, 0.5, apply_minmax=True)
    ws = mb.ops.pir_workspace(e, n, dev())
    mb.ops.pir_refine(None, attn, g, 0.5, workspace=ws, stages=mb.ops.PIR_NORMALISE)
    mb.ops.pir_refine(None, None, g, 0.5, workspace=ws, stages=mb.ops.PIR_CONTRACT, episodes=e)
    staged = mb.ops.pir_refine(prior, None, g, 0.5, apply_minmax=True, workspace=ws, stages=mb.ops.PIR_APPLY)
    assert torch.equal(whole, staged)


def test_pack_inside_a_partition_is_bit_exact(mb):
    """pack_masks / pairwise_inter on a green-context stream (fewer SMs, other grid sizes) against the oracle."""
    from marsb200.partition import SmPartition

    g = torch.Generator().manual_seed(11)
    masks = (torch.rand(37, 200, 312, generator=g) < 0.35).float()
    ref_bits = np_pack(masks.numpy() > 0, mb.ops.words_per_mask(200 * 312))
    ref_inter = orc.pairwise_intersections(masks)
    part = SmPartition(dev(), 64)
    try:
        x = masks.to(dev())
        torch.cuda.synchronize()
        with torch.cuda.stream(part.hbm_stream):
            bits = mb.ops.pack_masks(x)
        with torch.cuda.stream(part.tensor_stream):
            part.tensor_stream.wait_stream(part.hbm_stream)
            inter = mb.ops.pairwise_inter(bits)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(bits.cpu().numpy().view(np.uint32), ref_bits)
        got = ref_inter[0] if isinstance(ref_inter, tuple) else ref_inter
        np.testing.assert_array_equal(inter[0].cpu().numpy(), got.numpy())
    finally:
        part.close()


# ------------------------------------------------------------------ Matcher drop-in against the reference's own methods
@pytest.mark.parametrize("name", list(cases.MATCHER_CASES))
def test_matcher_dropin_matches_reference(mb, name):
    """marsb200.Matcher (same constructor and method signatures) against golden vectors produced by the reference's
    Matcher / RobustPromptSampler method bodies on the same fake encoder and SAM generator."""
    import random

    z = np.load(os.path.join(GOLD, f"matcher_{name}.npz"))
    spec = ast.literal_eval(str(z["spec"]))
    c = cases.matcher_inputs(spec)
    g, ps, size = spec["g"], spec["ps"], spec["g"] * spec["ps"]
    enc = cases.FakePatchEncoder(c["ref_raw"], c["tar_raw"], spec["ns"], ps, spec["C"])
    gen = cases.FakeSamGenerator(c["proposals"], c["point_coords"])
    m = mb.Matcher(encoder=enc, encoder_transforms=lambda x: x, generator=gen, input_size=size,
                   sample_range=spec["sample_range"], max_sample_iterations=spec["max_iter"], alpha=spec["alpha"],
                   beta=spec["beta"], exp=spec["exp"], score_filter_cfg=dict(spec["cfg"]),
                   num_merging_mask=spec["num_merging_mask"], use_negative_priors_from_discarded=spec["neg_discarded"],
                   use_negative_priors_from_cost=spec["neg_cost"], device=dev())
    m.set_reference(c["ref_imgs"], c["ref_masks"].clone())
    m.set_target(c["tar_img"])
    np.testing.assert_array_equal(m.ref_masks_pool.cpu().numpy(), z["ref_masks_pool"])
    ref_feats, tar_feat = m.extract_img_feats()
    pts, neg, box, S, C, reduced, reduced_neg = m.patch_level_matching(ref_feats, tar_feat)
    np.testing.assert_allclose(S.cpu().numpy(), z["sim"], rtol=RTOL, atol=3e-6)
    np.testing.assert_allclose(C.cpu().numpy(), (1 - z["sim"]) / 2, rtol=RTOL, atol=3e-6)

    def rows(a):
        a = np.asarray(a).reshape(-1, 2)
        return a[np.lexsort((a[:, 1], a[:, 0]))].astype(np.int64)

    np.testing.assert_array_equal(rows(pts), z["points"])
    assert reduced == int(z["reduced"]) and box is None
    neg_sets = neg if isinstance(neg, list) else [neg]
    assert len(neg_sets) == int(z["n_neg_sets"])
    np.testing.assert_array_equal(rows(neg_sets[0]), z["neg0"])
    assert [(-1 if r is None else r) for r in reduced_neg] == z["reduced_neg"].tolist()
    stats = dict(zip([str(k) for k in z["stats_keys"]], z["stats_vals"]))
    for k, v in m.get_patch_matching_statistics().items():
        assert v == stats[k], k

    # mask generation: the reference cannot run it with negative priors switched on (cases.py), mirror the fixture
    m.set_rps()
    neg_for_generation = neg
    if isinstance(neg, list):
        m.use_negative_priors_from_discarded = m.use_negative_priors_from_cost = False
        neg_for_generation = neg[0]
    random.seed(spec["seed"])
    merged, final = m.mask_generation(m.tar_img_np, pts, box, pts, m.ref_masks_pool, C, neg_for_generation)
    assert merged.shape == (1, size, size) and merged.dtype == torch.float32 and merged.is_cuda
    res = m.rps.batch_mask_scores(c["proposals"], pts, C, m.ref_masks_pool, spec["alpha"], spec["beta"], spec["exp"])
    np.testing.assert_allclose(res["purity"].cpu().numpy(), z["per_mask"][:, 0], rtol=RTOL)
    np.testing.assert_allclose(res["coverage"].cpu().numpy(), z["per_mask"][:, 1], rtol=RTOL)
    np.testing.assert_allclose(res["emd"].cpu().numpy(), z["per_mask"][:, 2], rtol=RTOL)
    np.testing.assert_array_equal(merged.cpu().numpy() > 0, z["merged"])
    np.testing.assert_array_equal(m.get_masks_to_merge().cpu().numpy() > 0, z["masks_to_merge"])
    assert abs(float(final) - float(z["final"])) <= RTOL * abs(float(z["final"]))
    assert m.number_of_merged_masks == int(z["merged_count"])
    got = m.get_mask_generation_statistics()
    for k in ("number_of_masks_before_score_filtering", "number_of_points_usable_for_prediction",
              "number_of_points_used_for_prediction", "positive_points_inside_mask", "negative_points_inside_mask",
              "ratio_points_used_vs_usable", "ratio_negative_vs_positive_points_inside_mask"):
        assert got[k] == pytest.approx(stats[k]), k
    assert m.get_unfiltered_generated_masks().shape == (spec["n_masks"], size, size)
    # the single-mask entry point of the sampler returns the reference's 6-tuple
    pur, cov, emd_score, p_out, labels, ori = m.rps.get_mask_scores(points=pts, masks=c["proposals"][3][None],
                                                                   all_points=pts, emd_cost=C,
                                                                   ref_masks_pool=m.ref_masks_pool)
    assert abs(float(pur[0]) - z["per_mask"][3, 0]) <= RTOL * z["per_mask"][3, 0] + 1e-9
    assert abs(emd_score - z["per_mask"][3, 2]) <= RTOL and labels.shape == (len(pts),)
    # diagnostics getters of the class (row A13) on the state this run left behind
    ref_n, tar_n = orc.normalize_rows(c["ref_raw"]), orc.normalize_rows(c["tar_raw"])
    pooled_ref = torch.from_numpy(z["ref_masks_pool"])
    r2t = m.get_ref_to_target_similarity(ref_feats, tar_feat, m.ref_masks_pool)
    np.testing.assert_allclose(r2t.reshape(-1).cpu().numpy(),
                               orc.ref_to_target_similarity(ref_n, tar_n, pooled_ref.reshape(spec["ns"], -1)).reshape(-1).numpy(),
                               rtol=RTOL, atol=1e-6)
    st = m.get_aposteriori_statistics(merged)
    tar_pool = orc.pool_mask(torch.from_numpy(z["merged"]).float(), g).reshape(-1)
    want = orc.aposteriori_statistics(ref_n @ tar_n.t(), pooled_ref, tar_pool, c["ref_raw"], c["tar_raw"])
    for k, v in want.items():
        assert st[k] == pytest.approx(v, rel=1e-3, abs=1e-5), k
    assert m.get_similarities()[0].shape[0] == int(pooled_ref.sum())
    # predict() = the same stages end to end; clear() resets the state and the generator
    m2 = mb.Matcher(encoder=cases.FakePatchEncoder(c["ref_raw"], c["tar_raw"], spec["ns"], ps, spec["C"]),
                    encoder_transforms=lambda x: x, generator=gen, input_size=size, sample_range=spec["sample_range"],
                    max_sample_iterations=spec["max_iter"], alpha=spec["alpha"], beta=spec["beta"], exp=spec["exp"],
                    score_filter_cfg=dict(spec["cfg"]), num_merging_mask=spec["num_merging_mask"], device=dev())
    m2.set_reference(c["ref_imgs"], c["ref_masks"].clone())
    m2.set_target(c["tar_img"])
    pred, score = m2.predict()
    np.testing.assert_array_equal(pred.cpu().numpy() > 0, z["merged"])
    assert abs(float(score) - float(z["final"])) <= RTOL * abs(float(z["final"]))
    m2.clear()
    assert m2.S is None and m2.ref_masks_pool is None and gen.resets == 1


def test_matcher_dropin_topk_rule_and_empty_reference(mb):
    """The top-k merge rule (the reference's own branch raises IndexError in its diagnostics, Matcher.py:822-827) against
    the oracle restatement, and the all-zero support mask rule of set_reference (:150-154)."""
    spec = dict(cases.MATCHER_CASES["g10_1shot_basic"])
    spec["cfg"] = dict(spec["cfg"], score_filter=False, topk_scores_threshold=0.99)
    c = cases.matcher_inputs(spec)
    g, ps, size = spec["g"], spec["ps"], spec["g"] * spec["ps"]
    gen = cases.FakeSamGenerator(c["proposals"], c["point_coords"])
    m = mb.Matcher(encoder=cases.FakePatchEncoder(c["ref_raw"], c["tar_raw"], 1, ps, spec["C"]),
                   encoder_transforms=lambda x: x, generator=gen, input_size=size, sample_range=spec["sample_range"],
                   max_sample_iterations=spec["max_iter"], score_filter_cfg=spec["cfg"], num_merging_mask=5, device=dev())
    m.set_reference(c["ref_imgs"], torch.zeros_like(c["ref_masks"]))
    pool = m.ref_masks_pool.cpu().reshape(g, g)
    assert pool.sum() == 4 and pool[g // 2 - 1:g // 2 + 1, g // 2 - 1:g // 2 + 1].all()  # rows 63..76 touch 2 x 2 patches
    m.set_reference(c["ref_imgs"], c["ref_masks"].clone())
    m.set_target(c["tar_img"])
    pred, score = m.predict()
    ref_n, tar_n = orc.normalize_rows(c["ref_raw"]), orc.normalize_rows(c["tar_raw"])
    pooled = orc.pool_mask(c["ref_masks"][0], g).reshape(-1).float()
    _, cost = orc.similarity_and_cost(ref_n, tar_n)
    pts, _, _ = orc.matcher_patch_matching(ref_n, tar_n, pooled, g, ps, (size, size))
    out = orc.matcher_generate_and_merge(c["proposals"], np.asarray(pts), cost, pooled, g, 1.0, 0.0, 0.0, spec["cfg"], 5)
    np.testing.assert_array_equal(pred.cpu().numpy()[0] > 0, out["merged"])
    assert abs(float(score) - out["final"]) <= RTOL * abs(out["final"])
    assert m.number_of_merged_masks == len(out["order"])


def test_pipelined_steps_match_one_timeline(mb):
    """PipelinedRanking: two buffer sets on one SM partition, step i + 1 enqueued before step i is joined - every
    step's outputs still equal the one-timeline engine's, bit for bit, and a buffer set is only reused after its
    previous step has drained."""
    shape = mb.EpisodeShape(ns=1, g=12, C=64, P=40, H=160, W=160, gt=9, D=32)
    batches = [mb.to_device(mb.stack_episodes([mb.make_episode(shape, 500 + 4 * b + i) for i in range(4)]), dev())
               for b in range(3)]
    base = mb.RankingEngine(shape, 4, mb.RankingConfig(nms_iou_threshold=0.7), dev())
    refs = [{k: v.clone() for k, v in base.run(b).items() if v is not None} for b in batches]
    cfg = mb.RankingConfig(nms_iou_threshold=0.7, tensor_partition_sms=64, partition_chunks=2, partition_vta_on_hbm=False)
    pipe = mb.PipelinedRanking(shape, 4, cfg, dev(), depth=2)
    try:
        prev = None
        for i in range(7):
            t = pipe.submit(batches[i % 3])
            if prev is not None:
                out = pipe.result(prev)
                torch.cuda.synchronize()
                for k, v in refs[prev % 3].items():
                    assert torch.equal(out[k], v), (prev, k)
                rec = mb.decode_records(pipe.engine(prev).records().cpu(), shape.P)
                assert torch.equal(rec["order"], refs[prev % 3]["order"].cpu())
            prev = t
        out = pipe.result(prev)
        torch.cuda.synchronize()
        for k, v in refs[prev % 3].items():
            assert torch.equal(out[k], v), (prev, k)
    finally:
        pipe.close()


def test_interleaved_engines_match_one_timeline(mb):
    """InterleavedRanking (two whole-device engines on two streams, the schedule of the uint8 / packed / RLE variants):
    step i + 1 is enqueued before step i is joined; every step's outputs equal the single engine's, bit for bit."""
    shape = mb.EpisodeShape(ns=1, g=12, C=64, P=40, H=160, W=160, gt=9, D=32)
    cfg = mb.RankingConfig(nms_iou_threshold=0.7)
    batches = [mb.to_device(mb.stack_episodes([mb.make_episode(shape, 700 + 4 * b + i, "cpu", torch.uint8) for i in range(4)]), dev())
               for b in range(3)]
    base = mb.RankingEngine(shape, 4, cfg, dev(), torch.uint8)
    refs = [{k: v.clone() for k, v in base.run(b).items() if v is not None} for b in batches]
    inter = mb.InterleavedRanking(shape, 4, cfg, dev(), torch.uint8, depth=2)
    prev = None
    for i in range(7):
        t = inter.submit(batches[i % 3])
        if prev is not None:
            out = inter.result(prev)
            torch.cuda.synchronize()
            for k, v in refs[prev % 3].items():
                assert torch.equal(out[k], v), (prev, k)
        prev = t
    out = inter.result(prev)
    torch.cuda.synchronize()
    for k, v in refs[prev % 3].items():
        assert torch.equal(out[k], v), (prev, k)
