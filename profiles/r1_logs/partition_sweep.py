import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, marsb200
dev = torch.device("cuda:0")
E = 16
shape = marsb200.CONFIGS["c2"]
batches = []
for b in range(2):
    eps = [marsb200.make_episode(shape, b * E + i, dev, torch.float32) for i in range(E)]
    batches.append(marsb200.stack_episodes(eps)); del eps
def tm(eng, iters=10):
    for i in range(3): eng.run(batches[i % 2])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): eng.run(batches[i % 2])
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
keys = ("order", "scores", "flags", "summary", "inter", "bits", "pooled", "area", "vva", "vta", "merged_bits", "merged")
base = marsb200.RankingEngine(shape, E, marsb200.RankingConfig(nms_iou_threshold=0.7), dev)
ref = {k: v.clone() for k, v in base.run(batches[1]).items() if k in keys and v is not None}
print(f"one timeline: {tm(base):.3f} ms per {E} episodes", flush=True)
for sms in (40, 48, 56, 64):
    for chunks, vh, tail in ((4, True, 0), (4, False, 0), (8, True, 0), (2, True, 0)):
        cfg = marsb200.RankingConfig(nms_iou_threshold=0.7, tensor_partition_sms=sms, partition_chunks=chunks, partition_vta_on_hbm=vh, partition_pairwise_tail=tail)
        eng = marsb200.RankingEngine(shape, E, cfg, dev)
        out = eng.run(batches[1]); torch.cuda.synchronize()
        bad = [k for k in ref if not torch.equal(out[k], ref[k])]
        t = tm(eng)
        print(f"tensor {eng._part.tensor_sms} / hbm {eng._part.hbm_sms} SMs, {chunks} chunks vta_on_hbm={vh} tail={tail}: {t:.3f} ms  mismatches: {bad}", flush=True)
        eng._part.close(); del eng
