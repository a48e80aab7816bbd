"""clock64 timeline of CTA 0 of the contraction kernel (MMA thread and epilogue warp 0, per tile).

Needs the profiling build: make -C <package>/csrc EXTRA=-DMARSB200_GEMM_PROFILE (touch gemm_tc.cu first), which adds
marsb200_debug_gemm_profile(); rebuild without EXTRA afterwards.  This is how the k-loop cost (clks per k-block), the
issue-blocking of tcgen05.mma and the fp64 epilogue slowdown in DESIGN.md section 4 were measured."""
import sys, os, statistics, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, marsb200
from marsb200 import ops
from marsb200 import _lib
raw = ctypes.CDLL(_lib.lib._name)
dev = torch.device("cuda:0")
E, M, N = 8, 1369, 1369
for K, stats in ((1024, True),):
    a = torch.randn(E, M, K, device=dev); b = torch.randn(E, N, K, device=dev)
    fa, fb = ops.normalize_rows(a), ops.normalize_rows(b)
    out = {}
    row_fg = (torch.rand(E, M, device=dev) < 0.2).to(torch.uint8) if stats else None
    res = ops.sim_contract(fa, fb, M, N, K, out=out, want_sim=not stats, row_fg=row_fg); out = {k: v for k, v in res.items() if v is not None}
    buf = (ctypes.c_longlong * 128)()
    raw.marsb200_debug_gemm_profile(buf, 1)
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.sim_contract(fa, fb, M, N, K, out=out, want_sim=not stats, row_fg=row_fg); t.record(); torch.cuda.synchronize()
    raw.marsb200_debug_gemm_profile(buf, 1)
    print(f"K={K} stats={stats}: {s.elapsed_time(t)*1e3:.1f} us = {s.elapsed_time(t)*1e3*1965:.0f} clks; CTA 0 timeline (clks from first stamp):")
    t0 = buf[0]
    print("  tile: mma_start  mma_accfree  mma_issued | epi_ready_to_wait  epi_acc_full  epi_tmem_released  epi_done")
    for tl in range(7):
        v = [buf[tl * 8 + i] - t0 for i in range(7)]
        print(f"  {tl}: {v[0]:9d} {v[1]:9d} {v[2]:9d} | {v[3]:9d} {v[4]:9d} {v[5]:9d} loop_end {buf[tl*8+7]-t0:9d} bar1 {buf[64+tl*2]-t0:9d} combine {buf[64+tl*2+1]-t0:9d} done {v[6]:9d}")
