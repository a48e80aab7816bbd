"""Per-op CUDA-event timings of one RankingEngine step (warm, in isolation), to see where a step goes."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marsb200
from marsb200 import ops

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8
shape = marsb200.CONFIGS["c2"]
dev = torch.device("cuda:0")
cfg = marsb200.RankingConfig(nms_iou_threshold=0.7, overlap_streams=False)
eng = marsb200.RankingEngine(shape, E, cfg, dev)
batches = [marsb200.stack_episodes([marsb200.make_episode(shape, b * E + i, dev) for i in range(E)]) for b in range(2)]
s, e = shape, E
n, m = s.N, s.ns * s.N
for b in batches:
    eng.run(b)
torch.cuda.synchronize()

def t(name, fn, iters=10):
    evs = []
    for i in range(iters):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(batches[i % 2]); b_.record(); evs.append((a, b_))
    torch.cuda.synchronize()
    ms = statistics.median(a.elapsed_time(b_) for a, b_ in evs)
    print(f"{name:28s} {ms*1e3:9.1f} us")
    return ms

tot = 0
tot += t("normalize_rows x2", lambda b: (ops.normalize_rows(b["feat_s"].reshape(e, m, s.C), True, out=eng.fs), ops.normalize_rows(b["feat_q"].reshape(e, n, s.C), True, out=eng.fq)))
tot += t("pool_mask", lambda b: ops.pool_mask(b["support_mask"], s.g, out=eng.row_fg))
tot += t("sim_contract", lambda b: ops.sim_contract(eng.fs, eng.fq, m, n, s.C, want_sim=False, row_fg=eng.row_fg, out=eng.gemm_out))
tot += t("vva_finalize", lambda b: ops.vva_finalize(eng.gemm_out["colstats"], eng.row_fg, m, n, out=eng.prior))
tot += t("pir vva (N=1369)", lambda b: ops.pir_refine(eng.prior, b["attn_vva"], s.g, 0.8, apply_minmax=True, workspace=eng.pir_ws, out=eng.vva))
tot += t("pir vta (N=1089)", lambda b: ops.pir_refine(b["vta_raw"], b["attn_vta"], s.gt, 0.4, workspace=eng.pir_ws, out=eng.vta_ref))
tot += t("resize_minmax", lambda b: ops.resize_minmax(eng.vta_ref.reshape(e, s.gt, s.gt), s.g, True, out=eng.vta))
tot += t("clip_scores", lambda b: ops.clip_scores(b["clip_img"], b["clip_txt"], out=eng.clip))
chain = tot
print(f"{'== alignment chain':28s} {chain*1e3:9.1f} us")
side = 0
side += t("pack_masks", lambda b: ops.pack_masks(b["masks"], out=eng.bits))
side += t("pool_packed", lambda b: ops.pool_packed(eng.bits, s.H, s.W, s.g, out=eng.pool_out))
side += t("pairwise_inter (mma)", lambda b: ops.pairwise_inter(eng.bits, out=eng.inter))
t("pairwise_inter (popc)", lambda b: ops.pairwise_inter(eng.bits, backend=ops.PAIR_POPC, out=eng.inter))
t("pack_pairwise fused", lambda b: ops.pack_pairwise(b["masks"], out=(eng.bits, eng.inter)))
print(f"{'== mask chain':28s} {side*1e3:9.1f} us")
tail = 0
tail += t("region_sums", lambda b: ops.region_sums(eng.pool_out[0], eng.vva, eng.vta, out=eng.region_out))
tail += t("fuse_rank", lambda b: ops.fuse_rank(b["emd"], eng.clip, eng.pool_out[2], eng.region_out[0], eng.region_out[1], eng.region_out[2], eng.inter, 0.85, 0.55, 0.95, 0.7, out=eng.rank_out))
tail += t("merge_masks", lambda b: ops.merge_masks(eng.bits, eng.rank_out["flags"], s.H * s.W, want_bits=True, want_f32=True, out=eng.merge_out))
print(f"{'== tail':28s} {tail*1e3:9.1f} us")
for name, c in (("serial", marsb200.RankingConfig(nms_iou_threshold=0.7, overlap_streams=False)),
                ("overlap", marsb200.RankingConfig(nms_iou_threshold=0.7)),
                ("fused", marsb200.RankingConfig(nms_iou_threshold=0.7, fused_ingest=True))):
    en = marsb200.RankingEngine(shape, E, c, dev)
    t(f"whole step [{name}]", lambda b: en.run(b), iters=20)
    del en
