"""Single-episode (c2) latency against RankingConfig.latency_contraction_sms (CTAs the contractions of the alignment streams
may start, marsb200_stream_set_sm_cap): graph replay, 3 rounds per setting."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marsb200

dev = torch.device("cuda:0")
shape = marsb200.CONFIGS["c2"]
one = [marsb200.stack_episodes([marsb200.make_episode(shape, i, dev)]) for i in range(2)]
ref = None
for cap in (0, 96, 64, 48, 40, 32, 24, None, 0):
    eng = marsb200.RankingEngine(shape, 1, marsb200.RankingConfig(nms_iou_threshold=0.7, latency_contraction_sms=cap), dev)
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
    sig = (out["order"].clone(), out["flags"].clone(), out["merged_bits"].clone(), out["scores"].clone())
    if ref is None:
        ref = sig
    same = all(torch.equal(x, y) for x, y in zip(sig, ref))
    print(f"latency_contraction_sms={cap}: eager {eager:.4f} ms, graph {min(rounds):.4f} ms (rounds {['%.4f' % x for x in rounds]}), same outputs {same}", flush=True)
