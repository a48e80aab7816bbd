"""Property tests (hypothesis) of the rules the kernels implement, run without a GPU.

SURVEY.md section 4 (ii): the reference has no tests, so the semantics of the path are pinned by properties stated directly
from the reference lines: the adaptive-pooling bin rule, intersections as popcounts of the packed words against the
`mb @ mb^T` contraction, greedy mask NMS against its defining invariants, the stable descending rank of Python's `sorted`,
the merge thresholds, and the RLE wire format.  The oracle is the subject here; where a product function runs on the host
(the packed bit layout written by `marsb200_host_pack_masks`, the RLE encoder of the synthetic generator, the shard ranges)
it is checked against the same property.
"""
import math

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import mars_oracle as orc

SETTINGS = dict(max_examples=40, deadline=None, derandomize=True, database=None)  # the same examples on every run


def _masks(seed: int, n: int, h: int, w: int, density: float) -> torch.Tensor:
    gen = torch.Generator().manual_seed(seed)
    return (torch.rand(n, h, w, generator=gen) < density).float()


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), g=st.integers(1, 12), dh=st.integers(0, 40), dw=st.integers(0, 40),
       density=st.sampled_from([0.0, 0.002, 0.05, 0.5]))
def test_pooling_bin_rule(seed, g, dh, dw, density):
    """Bin i of the pooled map covers [floor(i H / g), ceil((i + 1) H / g)) - F.adaptive_max_pool2d's rule, which is what
    FilteringMergingModule.py:104-107 applies to every proposal; neighbouring bins overlap whenever H % g != 0."""
    h, w = g + dh, g + dw
    m = _masks(seed, 2, h, w, density)
    got = orc.pool_mask(m, g).numpy()
    want = np.zeros((2, g, g), dtype=bool)
    mn = m.numpy() > 0
    for i in range(g):
        r0, r1 = (i * h) // g, -((-(i + 1) * h) // g)
        for j in range(g):
            c0, c1 = (j * w) // g, -((-(j + 1) * w) // g)
            want[:, i, j] = mn[:, r0:r1, c0:c1].any(axis=(1, 2))
    assert np.array_equal(got, want)


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), p=st.integers(1, 9), h=st.integers(1, 40), w=st.integers(1, 45),
       density=st.sampled_from([0.0, 0.1, 0.5, 1.0]), dtype=st.sampled_from(["f32", "u8", "bool"]))
def test_intersections_are_popcounts_of_the_packed_words(seed, p, h, w, density, dtype):
    """inter[i, j] = |m_i & m_j|: the oracle's blockwise `mb @ mb^T` equals AND + popcount over the words that the product's
    host packer writes (the layout every device kernel reads: bit k of word w = pixel 32 w + k, zero padding), for ragged
    shapes whose last word is partial."""
    from marsb200 import ops

    m = _masks(seed, p, h, w, density)
    inter, area = orc.pairwise_intersections(m)
    x = {"f32": m, "u8": (m * 255).to(torch.uint8), "bool": m.bool()}[dtype]
    bits = ops.host_pack_masks(x, threads=2).numpy().view(np.uint32)
    assert bits.shape == (p, ops.words_per_mask(h * w))
    anded = bits[:, None, :] & bits[None, :, :]
    pop = np.unpackbits(anded.view(np.uint8), axis=-1).sum(-1)
    assert np.array_equal(pop, inter.numpy())
    assert np.array_equal(area.numpy(), m.reshape(p, -1).sum(1).numpy().astype(np.int64))
    # the padding of every row is zero: a popcount over the whole row is the area
    assert np.array_equal(np.unpackbits(bits.view(np.uint8), axis=-1).sum(-1), area.numpy())


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), p=st.integers(1, 14), thr=st.sampled_from([0.0, 0.3, 0.5, 0.7, 1.0]))
def test_mask_nms_invariants(seed, p, thr):
    """Greedy suppression in rank order (torchvision `nms` semantics, automatic_mask_generator.py:370-376): the first ranked
    proposal is kept, no two kept proposals overlap by more than the threshold, and every dropped proposal has a kept,
    higher-ranked one that does.  These three invariants determine the keep-set."""
    rng = np.random.default_rng(seed)
    base = _masks(seed, max(1, p // 2), 12, 12, 0.4)
    m = base[rng.integers(0, base.shape[0], size=p)].clone()  # duplicates and near-duplicates
    flip = torch.from_numpy(rng.random((p, 12, 12)) < 0.05)
    m = torch.where(flip, 1 - m, m)
    inter, area = orc.pairwise_intersections(m)
    order = rng.permutation(p)
    keep = orc.mask_nms(order, inter, area, thr)
    iou = orc.iou_matrix(inter, area).numpy()
    rank = np.empty(p, dtype=int)
    rank[order] = np.arange(p)
    assert keep[order[0]]
    kept = np.nonzero(keep)[0]
    for a in kept:
        for b in kept:
            assert a == b or not iou[a, b] > np.float32(thr)
    for d in np.nonzero(~keep)[0]:
        assert any(rank[k] < rank[d] and iou[d, k] > np.float32(thr) for k in kept)
    if thr >= 1.0:
        assert keep.all()  # IoU never exceeds 1


@settings(**SETTINGS)
@given(scores=st.lists(st.sampled_from([0.0, 0.25, 0.5, 0.5000001, 0.75, 1.0, -1.0]), min_size=1, max_size=30))
def test_stable_rank_is_pythons_sorted(scores):
    """FilteringMergingModule.py:138 ranks with `sorted(..., key=score, reverse=True)`: stable, so equal scores keep the
    ascending proposal index."""
    want = [i for i, _ in sorted(enumerate(scores), key=lambda t: t[1], reverse=True)]
    assert orc.stable_rank(np.asarray(scores)).tolist() == want


@settings(**SETTINGS)
@given(scores=st.lists(st.floats(0.0, 1.0, allow_nan=False, width=32), min_size=1, max_size=20),
       static=st.sampled_from([0.3, 0.55, 0.9]), dynamic=st.sampled_from([0.5, 0.95, 1.0]))
def test_merge_thresholds(scores, static, dynamic):
    """FilteringMergingModule.py:213-217: everything >= static_threshold when the best proposal reaches it, otherwise
    everything >= dynamic_threshold * best; the best proposal is always selected."""
    ranked = np.sort(np.asarray(scores, dtype=np.float64))[::-1]
    sel = orc.merge_select(ranked, static, dynamic)
    top = ranked[0]
    want = [(s >= static) if top >= static else (s >= dynamic * top) for s in ranked]
    assert sel.tolist() == want
    assert sel[0]
    assert not (~sel[:-1] & sel[1:]).any()  # a prefix of the ranking


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 4), h=st.integers(1, 20), w=st.integers(1, 20),
       density=st.sampled_from([0.0, 0.1, 0.5, 1.0]))
def test_rle_wire_format_round_trip(seed, n, h, w, density):
    """Uncompressed COCO RLE (amg.py:107-149): column-major runs, the first run counts zeros (so a mask that starts with a set
    pixel has a leading 0), the counts add up to H W, decoding restores the mask; the generator's batched encoder (the
    engine's RLE ingest format) writes the same counts."""
    import marsb200

    m = _masks(seed, n, h, w, density)
    counts, offsets = marsb200.masks_to_rle(m)
    assert offsets[0] == 0 and offsets[-1] == counts.numel()
    for i in range(n):
        c = orc.mask_to_rle(m[i].numpy() > 0)
        assert sum(c) == h * w and all(x > 0 for x in c[1:])
        assert np.array_equal(orc.rle_to_mask(c, h, w), m[i].numpy() > 0)
        assert counts[offsets[i]:offsets[i + 1]].tolist() == c


@settings(**SETTINGS)
@given(num=st.integers(0, 5000), world=st.integers(1, 16))
def test_shard_ranges_partition_the_episodes(num, world):
    """Contiguous blocks of ceil(E / W) episodes per rank: disjoint, in rank order, covering every episode; only the
    last ranks can hold fewer (or none)."""
    import marsb200

    spans = [marsb200.shard_range(num, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == num
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    per = math.ceil(num / world)
    assert all(0 <= n <= per for n in sizes) and sizes == sorted(sizes, reverse=True)
    assert sum(1 for n in sizes if 0 < n < per) <= 1


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), p=st.integers(1, 12))
def test_fusion_is_invariant_to_proposal_order(seed, p):
    """FilteringMergingModule.py:118-138: the min-max terms are taken over the whole proposal set, so permuting the proposals
    permutes the fused scores and nothing else."""
    rng = np.random.default_rng(seed)
    emd, clip = rng.random(p), rng.random(p).astype(np.float32)
    cov, a_vv, a_vt = rng.random(p).astype(np.float32), rng.random(p).astype(np.float32), rng.random(p).astype(np.float32)
    base = np.asarray(orc.fuse_scores(emd, clip, cov, a_vv, a_vt, 0.85), dtype=np.float64)
    perm = rng.permutation(p)
    again = np.asarray(orc.fuse_scores(emd[perm], clip[perm], cov[perm], a_vv[perm], a_vt[perm], 0.85), dtype=np.float64)
    np.testing.assert_allclose(again, base[perm], rtol=0, atol=1e-12)
    if p == 1:  # both min-max terms vanish for a single proposal
        pvv, pvt = 0.85 * a_vv[0] + 0.15 * cov[0], 0.85 * a_vt[0] + 0.15 * cov[0]
        assert abs(base[0] - (pvv + pvt) / 4) < 1e-6
