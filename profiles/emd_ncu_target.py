"""One full-scoring step at c2 (8 episodes, 2048 transport LPs on the device) - the target of the emd_kernel ncu capture:
   ncu --set full --clock-control none --import-source on -k regex:emd_kernel -c 1 -o gpurun_out/emd python profiles/emd_ncu_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200
from marsb200 import ops

dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
E = 8
batch = marsb200.stack_episodes([marsb200.make_episode(shape, i, dev) for i in range(E)])
m_cap = int(ops.pool_packed(ops.pack_masks(batch["masks"]), shape.H, shape.W, shape.g)[2].max())
t_cap = int(ops.pool_mask(batch["support_mask"], shape.g).reshape(E, -1).sum(1).max())
cfg = marsb200.RankingConfig(nms_iou_threshold=0.7, emd_on_device=True, emd_m_cap=(m_cap + 63) // 64 * 64,
                             emd_t_cap=(t_cap + 63) // 64 * 64)
eng = marsb200.RankingEngine(shape, E, cfg, dev)
eng.run(batch)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.run(batch)
b.record()
torch.cuda.synchronize()
eng.check_status()
print(f"full-scoring step: {a.elapsed_time(b):.2f} ms for {E} episodes ({E * shape.P} LPs), t_cap={cfg.emd_t_cap} m_cap={cfg.emd_m_cap}")
