"""Packed / uint8 proposals, device-resident: one timeline vs two interleaved whole-device engines vs two engines on two
SM halves (SpatialRanking)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200
from marsb200 import ops
dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
E = 16
cfg = marsb200.RankingConfig(nms_iou_threshold=0.7)
batches = [marsb200.stack_episodes([marsb200.make_episode(shape, b * E + i, dev, torch.uint8) for i in range(E)]) for b in range(2)]
packed = [{k: v for k, v in b.items() if k != "masks"} for b in batches]
for a, b in zip(packed, batches):
    a["mask_bits"] = ops.pack_masks(b["masks"])
def tm(obj, data, iters=20):
    def loop(n):
        if isinstance(obj, marsb200.RankingEngine):
            for i in range(n): obj.run(data[i % 2])
            return
        prev = None
        for i in range(n):
            t = obj.submit(data[i % 2])
            if prev is not None: obj.result(prev)
            prev = t
        obj.result(prev)
    loop(4); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); loop(iters); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
base = marsb200.RankingEngine(shape, E, cfg, dev, torch.uint8)
ref = {k: v.clone() for k, v in base.run(packed[0]).items() if v is not None and k in ("order", "scores", "flags", "inter", "merged_bits", "vva", "vta")}
for name, data, dt in (("packed", packed, torch.uint8), ("u8", batches, torch.uint8)):
    print(name, "one timeline", round(tm(marsb200.RankingEngine(shape, E, cfg, dev, dt), data), 3), "ms", flush=True)
    print(name, "interleaved", round(tm(marsb200.InterleavedRanking(shape, E, cfg, dev, dt), data), 3), "ms", flush=True)
    for sms in (72, 64, 80):
        sp = marsb200.SpatialRanking(shape, E, cfg, dev, dt, first_sms=sms)
        out = sp.result(sp.submit(data[0])); torch.cuda.synchronize()
        bad = [k for k, v in ref.items() if not torch.equal(out[k], v)]
        print(name, f"spatial {sp._part.tensor_sms}/{sp._part.hbm_sms}", round(tm(sp, data), 3), "ms  mismatches:", bad, flush=True)
        sp.close()
