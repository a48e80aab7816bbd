"""Per-pivot work of the oracle's C network simplex on the transport LPs of synthetic c2 episodes (host only): pivots, cycle length,
size of the subtree whose duals change, arcs priced.  The numbers a CTA-parallel network simplex would be designed against
(EMD)."""
import ctypes, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location("syn", os.path.join(ROOT, "mars-multimodal-alignment-and-ranking-system-for-few-shot-segmentation_b200", "synthetic.py"))
syn = importlib.util.module_from_spec(spec); sys.modules["syn"] = syn; spec.loader.exec_module(syn)
from oracle import mars_oracle as orc

so = os.path.join(tempfile.mkdtemp(), "libstats.so")
subprocess.run(["gcc", "-O2", "-std=gnu11", "-DMARS_ORACLE_STATS", "-shared", "-fPIC", "-o", so,
                os.path.join(ROOT, "oracle", "emd_netsimplex.c"), "-lm"], check=True)
lib = ctypes.CDLL(so)
lib.mars_oracle_emd_netsimplex.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
stats = (ctypes.c_longlong * 8).in_dll(lib, "mars_oracle_stats")
shape = syn.CONFIGS["c2"]
for seed in (40, 41):
    ep = syn.make_episode(shape, seed)
    fs, fq = orc.normalize_rows(ep["feat_s"].reshape(-1, shape.C)), orc.normalize_rows(ep["feat_q"])
    _, cost = orc.similarity_and_cost(fs, fq)
    sup = orc.pool_mask(ep["support_mask"].float(), shape.g).reshape(-1)
    pm = orc.pool_mask(ep["masks"].float(), shape.g).reshape(shape.P, -1)
    C = cost[sup.bool()].numpy().astype(np.float64)
    for i in range(8):
        stats[i] = 0
    lps, nodes = 0, 0
    for k in range(shape.P):
        sub = np.ascontiguousarray(C[:, pm[k].numpy()])
        if sub.shape[1] == 0:
            continue
        obj = ctypes.c_double()
        assert lib.mars_oracle_emd_netsimplex(sub.ctypes.data, sub.shape[0], sub.shape[1], ctypes.byref(obj), None, None) == 0
        lps += 1; nodes += sum(sub.shape)
    p = stats[0]
    print(f"episode {seed}: T = {C.shape[0]}, {lps} LPs, mean nodes {nodes / lps:.0f}; pivots per LP {p / lps:.0f}; per pivot: cycle "
          f"{stats[1] / p:.1f} arcs (max {stats[2]}), cut depth {stats[5] / p:.1f}, smaller side of the cut {stats[3] / p:.1f} nodes, "
          f"arcs priced {stats[4] / p:.0f}")
