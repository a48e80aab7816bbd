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
keys = ("order", "scores", "flags", "summary", "inter", "bits", "pooled", "area", "vva", "vta", "merged_bits", "merged")
base = marsb200.RankingEngine(shape, E, marsb200.RankingConfig(nms_iou_threshold=0.7), dev)
refs = [{k: v.clone() for k, v in base.run(batches[b]).items() if k in keys and v is not None} for b in range(2)]
del base
def tm(pipe, iters=12):
    def loop(n):
        prev = None
        for i in range(n):
            t = pipe.submit(batches[i % 2])
            if prev is not None: pipe.result(prev)
            prev = t
        pipe.result(prev)
    loop(4); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); loop(iters); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for sms, chunks, vh, pt in ((64, 4, False, False), (64, 4, False, True), (56, 4, False, True), (56, 2, False, True), (56, 8, False, True),
                            (48, 4, False, True), (56, 4, True, True), (64, 2, False, True)):
    if True:
        cfg = marsb200.RankingConfig(nms_iou_threshold=0.7, tensor_partition_sms=sms, partition_chunks=chunks, partition_vta_on_hbm=vh,
                                     partition_prep_on_hbm=pt)
        pipe = marsb200.PipelinedRanking(shape, E, cfg, dev)
        # correctness: 5 overlapping steps, every result checked after its join
        tickets = []
        bad = []
        prev = None
        for i in range(5):
            t = pipe.submit(batches[i % 2])
            if prev is not None:
                out = pipe.result(prev); torch.cuda.synchronize()
                bad += [(prev, k) for k in refs[prev % 2] if not torch.equal(out[k], refs[prev % 2][k])]
            prev = t
        out = pipe.result(prev); torch.cuda.synchronize()
        bad += [(prev, k) for k in refs[prev % 2] if not torch.equal(out[k], refs[prev % 2][k])]
        t = tm(pipe)
        print(f"pipelined, tensor {pipe.engines[0]._part.tensor_sms} / hbm {pipe.engines[0]._part.hbm_sms} SMs, {chunks} chunks, vta_on_hbm={vh}, prep_on_hbm={pt}: {t:.3f} ms per step  mismatches: {bad}", flush=True)
        pipe.close(); del pipe
