import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, marsb200
from marsb200 import ops
from marsb200.partition import SmPartition, stream_sm_count
dev = torch.device("cuda:0")
E = 16
shape = marsb200.CONFIGS["c2"]
masks = [torch.stack([marsb200.make_episode(shape, b * E + i, dev)["masks"] for i in range(E)]) for b in range(2)]
bits = ops.pack_masks(masks[0]); ref_bits = bits.clone()
M = N = 1369; K = 1024
a = torch.randn(E, M, K, device=dev); b = torch.randn(E, N, K, device=dev)
fa, fb = ops.normalize_rows(a), ops.normalize_rows(b)
row_fg = (torch.rand(E, M, device=dev) < 0.2).to(torch.uint8)
out = ops.sim_contract(fa, fb, M, N, K, want_sim=False, row_fg=row_fg)
ref_cs = out["colstats"].clone()
inter = ops.pairwise_inter(ref_bits); ref_inter = inter.clone()
gemm = lambda: ops.sim_contract(fa, fb, M, N, K, want_sim=False, row_fg=row_fg, out=out)
pair = lambda: ops.pairwise_inter(ref_bits, out=inter)
def tensor_work():
    for _ in range(3): gemm()
    pair()
def tm(fn, stream, iters=6):
    with torch.cuda.stream(stream):
        for i in range(2): fn(i)
        stream.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(stream)
        for i in range(iters): fn(i)
        b_.record(stream); stream.synchronize()
    return a_.elapsed_time(b_) / iters * 1e3
main = torch.cuda.current_stream()
print(f"whole device ({stream_sm_count(main)} SMs): pack {tm(lambda i: ops.pack_masks(masks[i%2], out=bits), main):.0f} us; gemm x3 + pairwise {tm(lambda i: tensor_work(), main):.0f} us", flush=True)
for x in (32, 48, 56, 64, 72, 80):
    try:
        part = SmPartition(dev, x)
    except Exception as ex:
        print("split", x, "failed:", ex); continue
    ts, hs = part.tensor_stream, part.hbm_stream
    assert stream_sm_count(ts) == part.tensor_sms and stream_sm_count(hs) == part.hbm_sms
    t_pack = tm(lambda i: ops.pack_masks(masks[i%2], out=bits), hs)
    t_tens = tm(lambda i: tensor_work(), ts)
    # both at once
    torch.cuda.synchronize()
    def both(iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(main); ts.wait_event(e0); hs.wait_event(e0)
        with torch.cuda.stream(hs):
            for i in range(iters): ops.pack_masks(masks[i%2], out=bits)
            e1.record(hs)
        with torch.cuda.stream(ts):
            for i in range(iters): tensor_work()
            e2.record(ts)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3, e0.elapsed_time(e2) / iters * 1e3
    both(2)
    p, t = both(6)
    ok = torch.equal(bits, ops.pack_masks(masks[1])) and torch.equal(inter, ref_inter) and torch.equal(out["colstats"], ref_cs)
    print(f"tensor {part.tensor_sms} SMs / hbm {part.hbm_sms} SMs: alone pack {t_pack:.0f} us, tensor {t_tens:.0f} us; together pack ends {p:.0f} us, tensor ends {t:.0f} us; exact={ok}", flush=True)
    part.close()
