"""Single-episode (c2) latency against RankingConfig.latency_ingest_slices (pixel slices of the ingest, the intersections of
slice k counted beside the read of slice k + 1): graph replay, 3 rounds per setting."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200

dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
dtypes = {"f32": torch.float32, "u8": torch.uint8}
for name, md in dtypes.items():
    one = [marsb200.stack_episodes([marsb200.make_episode(shape, i, dev, md)]) for i in range(2)]
    ref = None
    for k in (1, 2, 4, 8, 16, None, 1):
        eng = marsb200.RankingEngine(shape, 1, marsb200.RankingConfig(nms_iou_threshold=0.7, latency_ingest_slices=k), dev, md)
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
        for i in range(5):
            eng.replay()
        rounds = []
        for r in range(3):
            a.record()
            for i in range(40):
                eng.replay()
            b.record(); torch.cuda.synchronize()
            rounds.append(a.elapsed_time(b) / 40)
        out = eng.outputs()
        sig = (out["order"].clone(), out["flags"].clone(), out["merged_bits"].clone(), out["scores"].clone(), out["inter"].clone())
        if ref is None:
            ref = sig
        same = all(torch.equal(x, y) for x, y in zip(sig, ref))
        print(f"{name} masks, latency_ingest_slices={k}: eager {eager:.4f} ms, graph {min(rounds):.4f} ms (rounds {['%.4f' % x for x in rounds]}), same outputs {same}", flush=True)
