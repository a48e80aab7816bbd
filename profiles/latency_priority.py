"""Single-episode (c2) latency: alignment chains on high-priority streams against the default streams (graph replay)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200

dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
one = [marsb200.stack_episodes([marsb200.make_episode(shape, i, dev)]) for i in range(2)]
ref = None
import itertools
for prio, hoist in ((False, False), (True, False), (True, True), (True, False), (True, True)):
    eng = marsb200.RankingEngine(shape, 1, marsb200.RankingConfig(nms_iou_threshold=0.7, priority_streams=prio, hoist_vva_contraction=hoist), dev)
    for i in range(4):
        eng.run(one[i % 2])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20):
        eng.run(one[i % 2])
    b.record(); torch.cuda.synchronize()
    eager = a.elapsed_time(b) / 20
    eng.capture(one[0])
    for i in range(3):
        eng.replay()
    a.record()
    for i in range(20):
        eng.replay()
    b.record(); torch.cuda.synchronize()
    out = eng.outputs()
    sig = (out["order"].clone(), out["flags"].clone(), out["merged_bits"].clone())
    if ref is None:
        ref = sig
    same = all(torch.equal(x, y) for x, y in zip(sig, ref))
    t0 = time.perf_counter()
    for i in range(20):
        eng.replay(); torch.cuda.synchronize()
    print(f"priority_streams={prio} hoist={hoist}: eager {eager:.4f} ms, graph {a.elapsed_time(b) / 20:.4f} ms, graph one at a time (host clock) {(time.perf_counter() - t0) / 20 * 1e3:.4f} ms, same outputs {same}", flush=True)
