"""Single-episode (c2) ranking, eager launches: the target of the per-kernel launch list behind the latency numbers.
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lat.csv python profiles/latency_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200

dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
cfg = marsb200.RankingConfig(nms_iou_threshold=0.7)
eng = marsb200.RankingEngine(shape, 1, cfg, dev)
one = [marsb200.stack_episodes([marsb200.make_episode(shape, i, dev)]) for i in range(2)]
for i in range(4):
    eng.run(one[i % 2])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(10):
    eng.run(one[i % 2])
b.record()
torch.cuda.synchronize()
print(f"eager: {a.elapsed_time(b) / 10:.4f} ms per episode")
eng.capture(one[0])
for i in range(3):
    eng.replay()
a.record()
for i in range(10):
    eng.replay()
b.record()
torch.cuda.synchronize()
print(f"graph: {a.elapsed_time(b) / 10:.4f} ms per episode")
