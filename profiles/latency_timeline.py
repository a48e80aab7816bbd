"""Single-episode (c2) step as a timeline: CUDA events around every op of RankingEngine.run, per stream, relative to the
   start of the step (eager launches; the graph replay removes the launch gaps, the order of the chains stays)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200
from marsb200 import ops

dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
cfg = marsb200.RankingConfig(nms_iou_threshold=0.7)
eng = marsb200.RankingEngine(shape, 1, cfg, dev)
one = [marsb200.stack_episodes([marsb200.make_episode(shape, i, dev)]) for i in range(2)]
for i in range(4):
    eng.run(one[i % 2])
torch.cuda.synchronize()

names = ["normalize_rows", "pool_mask", "sim_contract", "vva_finalize", "pir_refine", "resize_minmax", "clip_scores",
         "pack_masks", "pool_packed", "pairwise_inter", "region_sums", "fuse_rank", "merge_masks"]
log = []
orig = {}
for nme in names:
    fn = getattr(ops, nme)
    orig[nme] = fn

    def wrap(*a, _fn=fn, _n=nme, **k):
        st = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        r = _fn(*a, **k)
        e1.record(st)
        log.append((_n, st.cuda_stream, e0, e1))
        return r

    setattr(ops, nme, wrap)

for rep in range(3):
    log.clear()
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record()
    eng.run(one[rep % 2])
    t1 = torch.cuda.Event(enable_timing=True)
    t1.record()
    torch.cuda.synchronize()
print(f"step (eager, instrumented): {t0.elapsed_time(t1) * 1e3:.1f} us")
ids = {}
for nme, sid, e0, e1 in sorted(log, key=lambda r: t0.elapsed_time(r[2])):
    k = ids.setdefault(sid, len(ids))
    print(f"stream {k}  {nme:16s} start {t0.elapsed_time(e0) * 1e3:7.1f}  end {t0.elapsed_time(e1) * 1e3:7.1f}  ({e0.elapsed_time(e1) * 1e3:6.1f} us)")
