"""clock64 timeline of CTA 0 of the symmetric D D^T contraction inside marsb200_pir_refine (c2: 16 episodes, N = 1369):
where the MMA thread waits (accumulator not yet drained = epilogue-bound, operands not yet landed = TMA-bound).
Needs the profiling build (see gemm_timeline.py)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, marsb200
from marsb200 import ops, _lib
raw = ctypes.CDLL(_lib.lib._name)
dev = torch.device("cuda:0")
for E, g in ((16, 37), (16, 33)):
    n = g * g
    gen = torch.Generator(device=dev).manual_seed(1)
    attn = torch.softmax(2.0 * torch.randn(E, n, n, device=dev, generator=gen), -1)
    prior = torch.rand(E, n, device=dev, generator=gen)
    ws = ops.pir_workspace(E, n, dev)
    out = torch.empty(E, n, device=dev)
    for _ in range(2):
        ops.pir_refine(prior, attn, g, 0.5, workspace=ws, out=out)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 128)()
    raw.marsb200_debug_gemm_profile(buf, 1)
    ops.pir_refine(prior, attn, g, 0.5, workspace=ws, out=out)
    torch.cuda.synchronize()
    raw.marsb200_debug_gemm_profile(buf, 1)
    t0 = buf[0]
    print(f"N={n} E={E}: CTA 0 timeline (clks from the first stamp)")
    print("  tile: mma_start mma_acc_free mma_issued | epi_wait epi_acc_full epi_tmem_released epi_loop_end epi_done")
    for tl in range(8):
        v = [buf[tl * 8 + i] - t0 for i in range(8)]
        if buf[tl * 8 + 2] == 0:
            break
        print(f"  {tl}: {v[0]:9d} {v[1]:9d} {v[2]:9d} | {v[3]:9d} {v[4]:9d} {v[5]:9d} {v[7]:9d} {v[6]:9d}   "
              f"(mma waited {v[1] - v[0]} clks for the accumulator, issued for {v[2] - v[1]}; epilogue {v[6] - v[4]} clks)")
