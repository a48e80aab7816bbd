"""One-timeline throughput (c2, 16 episodes per step, float32 / uint8 / packed proposals) with and without high-priority
   alignment streams."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200
from marsb200 import ops

dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
E = 16
batches = [marsb200.stack_episodes([marsb200.make_episode(shape, b * E + i, dev) for i in range(E)]) for b in range(2)]
u8 = [dict(b, masks=b["masks"].to(torch.uint8)) for b in batches]
packed = []
for b in batches:
    d = {k: v for k, v in b.items() if k != "masks"}
    d["mask_bits"] = ops.pack_masks(b["masks"])
    packed.append(d)
for name, data, md in (("f32", batches, torch.float32), ("u8", u8, torch.uint8), ("packed", packed, torch.float32)):
    for prio, hoist in ((False, False), (True, False), (True, True), (True, False), (True, True)):
        eng = marsb200.RankingEngine(shape, E, marsb200.RankingConfig(nms_iou_threshold=0.7, priority_streams=prio, hoist_vva_contraction=hoist), dev, md)
        for i in range(4):
            eng.run(data[i % 2])
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            eng.run(data[i % 2])
        b_.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b_) / 20
        print(f"{name:7s} priority_streams={prio} hoist={hoist}: {ms:.3f} ms per step = {E / ms * 1e3:.0f} episodes/s", flush=True)
        del eng
