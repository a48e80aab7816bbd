"""Small invocations of every kernel family, run under compute-sanitizer (profiles/run_sanitizer.sh).

memcheck: out-of-bounds / misaligned global + shared accesses; racecheck: shared-memory hazards (the EMD and
pairwise kernels keep cross-warp lists in shared memory; the tensor-core kernels hand tiles over through mbarriers).
Sizes are tiny: the sanitizer slows kernels down by 10-100x.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
import marsb200  # noqa: E402
from marsb200 import ops  # noqa: E402

dev = torch.device("cuda:0")
if which in ("all", "smoke"):
    entry.smoke()  # whole episode incl. device EMD, contractions, PIR, pairwise, fuse_rank, merge
if which in ("all", "pairwise"):
    g = torch.Generator().manual_seed(3)
    masks = (torch.rand(2, 300, 96, 96, generator=g) < 0.3).to(dev)
    bits = ops.pack_masks(masks)
    ref = None
    for be in (ops.PAIR_POPC, ops.PAIR_MMA):
        out = ops.pairwise_inter(bits, backend=be)
        ref = out if ref is None else ref
        assert torch.equal(out, ref)
    out = ops.pairwise_inter(bits[:, :200].contiguous(), backend=ops.PAIR_FP4)
    assert torch.equal(out, ref[:, :200, :200])
if which in ("all", "emd"):
    shape = marsb200.EpisodeShape(ns=2, g=14, C=64, P=48, H=196, W=196, gt=9, D=32)
    cfg = marsb200.RankingConfig(nms_iou_threshold=0.7, emd_on_device=True)
    eng = marsb200.RankingEngine(shape, 2, cfg, dev)
    eng.run(marsb200.to_device(marsb200.stack_episodes([marsb200.make_episode(shape, 70 + i, with_emd=False) for i in range(2)]), dev))
if which in ("all", "lsap"):
    g = torch.Generator().manual_seed(5)
    S = torch.rand(1, 90, 140, generator=g).to(dev)
    ops.lsap(S, torch.ones(1, 90, dtype=torch.uint8, device=dev), None, maximize=True)
torch.cuda.synchronize()
print("sanitize target ok:", which)
