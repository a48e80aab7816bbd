"""Single-episode (c2) latency on two SM partitions (green contexts) against the one-timeline engine.
   python profiles/latency_partitioned.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200

dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
one = [marsb200.stack_episodes([marsb200.make_episode(shape, i, dev)]) for i in range(2)]


def timeit(eng, label, iters=20):
    for i in range(4):
        eng.run(one[i % 2])
    torch.cuda.synchronize()
    # latency = one step at a time: the host waits for each result before it submits the next episode
    t0 = time.perf_counter()
    for i in range(iters):
        eng.run(one[i % 2])
        torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / iters
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        eng.run(one[i % 2])
    b.record()
    torch.cuda.synchronize()
    print(f"{label:40s} back-to-back {a.elapsed_time(b) / iters:.4f} ms   one at a time (host clock) {wall * 1e3:.4f} ms", flush=True)


eng = marsb200.RankingEngine(shape, 1, marsb200.RankingConfig(nms_iou_threshold=0.7), dev)
timeit(eng, "one timeline, eager")
eng.capture(one[0])
for i in range(3):
    eng.replay()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    eng.replay()
    torch.cuda.synchronize()
print(f"{'one timeline, graph replay':40s} one at a time (host clock) {(time.perf_counter() - t0) / 20 * 1e3:.4f} ms", flush=True)
for sms in (40, 56, 64, 80, 96):
    for vta_on_hbm in (True, False):
        for tail in (0, 1):
            cfg = marsb200.RankingConfig(nms_iou_threshold=0.7, tensor_partition_sms=sms, partition_chunks=1,
                                         partition_vta_on_hbm=vta_on_hbm, partition_pairwise_tail=tail)
            try:
                eng = marsb200.RankingEngine(shape, 1, cfg, dev)
                timeit(eng, f"partition tensor={sms} vta_on_hbm={vta_on_hbm} pair_on_hbm={tail}")
                eng.close()
            except Exception as ex:  # green contexts unavailable
                print(f"partition {sms}: {type(ex).__name__}: {ex}")
                break
