"""CPU-only checks: the C-ABI library loads and exports every declared symbol; host-side logic."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import marsb200
    from marsb200 import _lib

    header = open(os.path.join(ROOT, "include", "marsb200.h")).read()
    declared = set(re.findall(r"\b(marsb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(_lib.lib, name), f"{name} declared in include/marsb200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.lib.marsb200_version() >= 100
    assert marsb200.ops.words_per_mask(518 * 518) == 8416
    assert marsb200.ops.words_per_mask(1024 * 1024) == 32768
    assert marsb200.ops.pad_rows(1369) == 1408 and marsb200.ops.pad_k(1369) == 1376


def test_ops_refuse_cpu_tensors():
    import marsb200

    with pytest.raises(marsb200.MarsB200Error):
        marsb200.ops.pack_masks(torch.zeros(2, 8, 8))
    with pytest.raises(RuntimeError):
        marsb200.RankingEngine(marsb200.CONFIGS["c1"], 1, marsb200.RankingConfig(), "cpu")


def test_shard_ranges_cover_everything():
    from marsb200 import shard_range

    for n in (1, 7, 4096, 4097):
        for w in (1, 2, 4, 8):
            got = []
            for r in range(w):
                lo, hi = shard_range(n, r, w)
                got.extend(range(lo, hi))
            assert got == list(range(n))


def test_record_roundtrip():
    from marsb200 import decode_records

    p, e = 5, 3  # the flags section is padded to a multiple of 4 bytes (marsb200_record_bytes)
    order = torch.arange(e * p, dtype=torch.int32).reshape(e, p)
    scores = torch.rand(e, p)
    flags = torch.randint(0, 4, (e, p), dtype=torch.uint8)
    summary = torch.arange(e * 4, dtype=torch.int32).reshape(e, 4)
    pad = torch.zeros((e, (p + 3) // 4 * 4 - p), dtype=torch.uint8)
    rec = torch.cat([order.view(torch.uint8).reshape(e, -1), scores.view(torch.uint8).reshape(e, -1), flags, pad,
                     summary.view(torch.uint8).reshape(e, -1)], dim=1)
    from marsb200 import ops
    assert rec.shape[1] == ops.record_bytes(p)
    d = decode_records(rec, p)
    assert torch.equal(d["order"], order) and torch.equal(d["scores"], scores)
    assert torch.equal(d["flags"], flags) and torch.equal(d["summary"], summary)


def _gather_worker(rank, world, port, n_ep, q):
    import torch.distributed as dist

    from marsb200 import gather_records, shard_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_ep, rank, world)
    local = torch.arange(lo, hi, dtype=torch.uint8)[:, None].repeat(1, 6)
    out = gather_records(local, n_ep)
    if rank == 0:
        q.put(out.numpy())
    dist.destroy_process_group()


def _fake_records(rank, e, p):
    """What fuse_rank writes for `e` episodes of rank `rank` (layout of marsb200_record_bytes), built with numpy."""
    rs = np.random.RandomState(100 + rank)
    fl = (p + 3) // 4 * 4
    rec = np.zeros((e, 8 * p + fl + 16), dtype=np.uint8)
    fields = []
    for i in range(e):
        order = rs.permutation(p).astype(np.int32)
        score = rs.rand(p).astype(np.float32)
        flags = rs.randint(0, 4, p).astype(np.uint8)
        summary = np.asarray([int((flags & 1).sum()), int((flags & 2).sum() // 2), int(order[0]), 0], dtype=np.int32)
        rec[i, :4 * p] = order.view(np.uint8)
        rec[i, 4 * p:8 * p] = score.view(np.uint8)
        rec[i, 8 * p:9 * p] = flags
        rec[i, 8 * p + fl:] = summary.view(np.uint8)
        fields.append((order, score, flags, summary))
    return rec, fields


def _inplace_gather_worker(rank, world, port, e, p, q):
    """The engine's pattern (RankingEngine.attach_gather_table / gather): every rank's kernels write its slice of ONE
    table, the all-gather runs in place on it."""
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec, _ = _fake_records(rank, e, p)
    table = torch.zeros((world * e, rec.shape[1]), dtype=torch.uint8)
    mine = table[rank * e:(rank + 1) * e]
    mine.copy_(torch.from_numpy(rec))  # stands in for fuse_rank writing the records
    dist.all_gather_into_tensor(table, mine)
    q.put((rank, table.numpy().copy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("p", [12, 10])  # 10: the flags section is padded to a multiple of 4 bytes
def test_two_rank_inplace_gather_of_result_records(p):
    """world_size-2 gloo run: the gathered table on EVERY rank equals the concatenation of the rank-local records, and
    decode_records splits it back into the fields each rank produced."""
    import torch.multiprocessing as mp

    import marsb200

    e, world = 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + p
    procs = [ctx.Process(target=_inplace_gather_worker, args=(r, world, port, e, p, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    want = np.concatenate([_fake_records(r, e, p)[0] for r in range(world)])
    assert want.shape[1] == marsb200.ops.record_bytes(p)
    for r in range(world):
        np.testing.assert_array_equal(got[r], want)
    d = marsb200.decode_records(torch.from_numpy(want), p)
    for r in range(world):
        for i, (order, score, flags, summary) in enumerate(_fake_records(r, e, p)[1]):
            k = r * e + i
            np.testing.assert_array_equal(d["order"][k].numpy(), order)
            np.testing.assert_array_equal(d["scores"][k].numpy(), score)
            np.testing.assert_array_equal(d["flags"][k].numpy(), flags)
            np.testing.assert_array_equal(d["summary"][k].numpy(), summary)


@pytest.mark.parametrize("n_ep", [8, 7])
def test_two_rank_gather_equals_single_rank(n_ep):
    """world_size-2 gloo run of the sharded path: concatenated records equal the 1-rank table."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_ep
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_ep, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(out, np.arange(n_ep, dtype=np.uint8)[:, None].repeat(6, 1))


def test_synthetic_episode_shapes():
    import marsb200

    shape = marsb200.EpisodeShape(ns=2, g=6, C=16, P=10, H=60, W=60, gt=4, D=8)
    ep = marsb200.make_episode(shape, 3)
    assert ep["feat_s"].shape == (2, 36, 16) and ep["masks"].shape == (10, 60, 60)
    assert bool((ep["masks"].flatten(1).sum(1) > 0).all())
    again = marsb200.make_episode(shape, 3)
    assert all(torch.equal(ep[k], again[k]) for k in ep)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs) prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "episodes_per_sec" and d["unit"] == "episodes/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_header_is_plain_c():
    """include/marsb200.h must parse as C99 and as C++ (the C-ABI boundary: no torch or CUDA types in signatures)."""
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "marsb200.h")
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    subprocess.run(["gcc", "-x", "c", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", hdr], check=True)
    subprocess.run(["g++", "-x", "c++", "-fsyntax-only", "-Wall", "-Werror", hdr], check=True)
    import re

    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)  # declarations only, comments stripped
    for banned in ("torch", "at::", "Tensor", "cudaStream_t", "std::"):
        assert banned not in code, banned


def test_prompt_sampler_reproduces_reference_stream():
    """RobustPromptSampler.combinations / sample_points of the drop-in against the golden vectors made by the reference's
    own class with the same `random` seed (host-side logic, no device needed)."""
    import ast
    import random
    import sys

    import marsb200

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases

    for name in cases.MATCHER_CASES:
        z = np.load(os.path.join(ROOT, "tests", "golden", f"matcher_{name}.npz"))
        spec = ast.literal_eval(str(z["spec"]))
        rps = marsb200.RobustPromptSampler(spec["g"], spec["sample_range"], spec["max_iter"], device="cpu")
        np.testing.assert_array_equal(np.asarray(rps.combinations(5, 3)), z["combos_5_3"])
        assert rps.combinations(3, 4) == [] and rps.combinations(4, 0) == [[]]
        random.seed(spec["seed"] + 1)
        demo = np.arange(22).reshape(11, 2)
        s_many, l_many = rps.sample_points(demo, negative_points=demo[:5] + 100)
        s_few, l_few = rps.sample_points(demo[:5])
        np.testing.assert_array_equal(np.concatenate([x.reshape(-1) for x in s_many]), z["sample_many"])
        np.testing.assert_array_equal(np.concatenate([x.reshape(-1) for x in l_many]), z["label_many"])
        np.testing.assert_array_equal(np.concatenate([x.reshape(-1) for x in s_few]), z["sample_few"])
        np.testing.assert_array_equal(np.concatenate([x.reshape(-1) for x in l_few]), z["label_few"])


def test_launch_count_follows_the_schedule():
    """bench.py reports gpu_launches from this count: chunked pack / pool / pairwise on SM partitions, six EMD kernels."""
    import marsb200
    from marsb200 import RankingConfig, kernel_launches_per_run

    base = kernel_launches_per_run(RankingConfig(nms_iou_threshold=0.7))
    assert kernel_launches_per_run(RankingConfig(nms_iou_threshold=0.7, tensor_partition_sms=56), 16) == base + 3 * 4
    assert kernel_launches_per_run(RankingConfig(nms_iou_threshold=0.7, tensor_partition_sms=56, partition_chunks=8), 2) == base + 4
    assert kernel_launches_per_run(RankingConfig(nms_iou_threshold=None, tensor_partition_sms=56), 16) == \
        kernel_launches_per_run(RankingConfig(nms_iou_threshold=None)) + 2 * 3
    assert kernel_launches_per_run(RankingConfig(nms_iou_threshold=0.7, emd_on_device=True)) == base + 6


def test_build_mars_fss_one_argument_form_needs_the_reference_loaders():
    """`build_MARS_fss(args)` (the reference's signature, mars/MARS.py:110-116) loads the PyTorch producers with the
    reference's own loaders; without the reference checkout on sys.path it must say so instead of failing obscurely."""
    import types

    import marsb200

    args = types.SimpleNamespace(input_size=518, num_regs=4, device="cuda:0", models_path="/nonexistent",
                                 dino_backbone="vitl14", vva_refinement_box_threshold=0.8,
                                 last_n_attn_for_vva_refinement=24, alpha_coverage=0.85, static_threshold=0.55,
                                 dynamic_threshold=0.95)
    with pytest.raises(ImportError, match="reference"):
        marsb200.build_MARS_fss(args)


def test_record_layout_matches_the_header():
    """marsb200_record_bytes: order int32[P] | score float32[P] | flags uint8[P rounded up to 4] | summary int32[4]."""
    from marsb200 import ops

    for p in (1, 3, 4, 10, 256, 1000):
        assert ops.record_bytes(p) == 8 * p + (p + 3) // 4 * 4 + 16
        assert ops.record_bytes(p) % 4 == 0


@pytest.mark.parametrize("n,h,w,dtype", [(5, 37, 41, "f32"), (3, 64, 64, "u8"), (2, 256, 512, "f32"), (4, 7, 5, "bool"), (1, 1, 1, "f32")])
@pytest.mark.parametrize("threads", [1, 3, 0])
def test_host_pack_masks_matches_the_packed_layout(n, h, w, dtype, threads):
    """marsb200_host_pack_masks (host threads, no CUDA call) writes exactly the bit layout of the device ingest kernel:
    bit k of word w = pixel 32 w + k is > 0 (NaN and negatives clear), rows padded with zero words to a multiple of 32."""
    import numpy as np
    import torch

    from marsb200 import ops

    gen = torch.Generator().manual_seed(100 * n + h)
    m = torch.rand(n, h, w, generator=gen) < 0.3
    if dtype == "f32":
        x = m.float() * (0.1 + torch.rand(n, h, w, generator=gen))
        x[0, 0, 0] = float("nan")
        if w > 1:
            x[0, 0, 1] = -1.0
    else:
        x = m.to(torch.uint8 if dtype == "u8" else torch.bool)
        if dtype == "u8":
            x = x * 255
    bits = ops.host_pack_masks(x, threads=threads)
    hw = h * w
    wpm = ops.words_per_mask(hw)
    assert tuple(bits.shape) == (n, wpm) and wpm % 32 == 0
    ref = np.zeros((n, wpm * 32), dtype=np.uint8)
    with np.errstate(invalid="ignore"):
        ref[:, :hw] = x.reshape(n, -1).numpy() > 0
    want = np.packbits(ref.reshape(n, wpm, 32), axis=-1, bitorder="little").view(np.uint32).reshape(n, wpm)
    np.testing.assert_array_equal(bits.numpy().view(np.uint32), want)


def test_stream_sm_cap_table_is_host_side():
    """marsb200_stream_set_sm_cap only edits a host table (no CUDA call): it accepts any handle, replaces and removes
    entries, and refuses negative caps - the GPU suite checks what marsb200_stream_sm_count then reports."""
    import ctypes

    from marsb200 import _lib

    h = ctypes.c_void_p(0x1234)
    assert _lib.lib.marsb200_stream_set_sm_cap(h, 48) == 0
    assert _lib.lib.marsb200_stream_set_sm_cap(h, 32) == 0
    assert _lib.lib.marsb200_stream_set_sm_cap(h, 0) == 0
    assert _lib.lib.marsb200_stream_set_sm_cap(h, 0) == 0  # removing a missing entry is not an error
    assert _lib.lib.marsb200_stream_set_sm_cap(h, -1) != 0
    assert b"cap" in _lib.lib.marsb200_last_error()


@pytest.mark.parametrize("t,n,ns", [(150, 144, 3), (144, 144, 2), (400, 144, 3)])
def test_reverse_assignment_is_independent_of_the_forward_one_when_t_ge_n(t, n, ns):
    """The claim behind PatchMatcher.concurrent_reverse, pinned on the reference's own solver (scipy, Matcher.py:449,470):
    with T >= N masked support patches the forward assignment matches EVERY query patch, so `S.t()[indices_forward[1]]` is a
    row permutation of S.t() and its optimal assignment is the identity-order one gathered through that permutation."""
    import numpy as np
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(t * 1000 + n)
    S = rng.standard_normal((ns * n, n)).astype(np.float32)  # tie-free with probability one
    rows = np.sort(rng.choice(ns * n, size=t, replace=False))
    fr, fc = linear_sum_assignment(S[rows], maximize=True)
    assert len(fc) == n and sorted(fc.tolist()) == list(range(n))
    rr, rc = linear_sum_assignment(S.T[fc], maximize=True)          # the reference's order of rows
    ir, ic = linear_sum_assignment(S.T, maximize=True)              # identity order, solvable before the forward problem
    assert np.array_equal(rr, np.arange(n)) and np.array_equal(ir, np.arange(n))
    assert np.array_equal(rc, ic[fc])


def test_roofline_traffic_is_tied_to_the_kernel_source(monkeypatch):
    """bench.py reports `roofline.traffic` from the committed ncu capture only while the sha256 of the ingest kernel's source
    still matches the one recorded with the capture: the committed capture is current, and any other
    hash - a changed kernel - yields null instead of a stale number."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    got = bench.ncu_traffic_bytes("pack_f32_vec_kernel", "c2", 16, "f32")
    algorithmic = 16 * 256 * (4 * 1024 * 1024 + 4 * 32768)
    assert got is not None and 0.95 * algorithmic < got < 1.05 * algorithmic  # no re-reads: traffic = algorithmic bytes
    assert bench.ncu_traffic_bytes("pack_f32_vec_kernel", "c2", 8, "f32") is None  # no capture of that launch shape
    monkeypatch.setattr(bench, "kernel_source_sha16", lambda *a, **k: "0" * 16)
    assert bench.ncu_traffic_bytes("pack_f32_vec_kernel", "c2", 16, "f32") is None


def test_bench_cpu_emd_leg_runs_both_host_solvers():
    """bench.py's CPU EMD baseline: the sampled HiGHS LPs and every LP of the episode through the C network simplex (all host
    threads), which must agree; the full-scoring host figure is built from the faster one."""
    import importlib.util

    import marsb200

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test_emd", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    shape = marsb200.EpisodeShape(ns=1, g=10, C=32, P=9, H=80, W=80, gt=8, D=16)
    batch = marsb200.stack_episodes([marsb200.make_episode(shape, 3)])
    res = bench.cpu_emd_sample(shape, batch, 2)
    assert res["lps_per_s"] > 0 and res["cores"] == 1
    ns = res["network_simplex"]
    assert ns["lps"] == 9 and ns["lps_per_s"] > 0 and ns["agrees_with_highs"] is True


def _meter_all_reduce_worker(rank, world, port, nclass, q):
    """A rank of a sharded evaluation: int64 area buffers filled from this rank's episodes, then AverageMeter.all_reduce."""
    import types

    import torch.distributed as dist

    from marsb200.evaluation import AverageMeter

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(7 + rank)
    bufs = [torch.from_numpy(rs.randint(0, 2 ** 40, size=(2, nclass)).astype(np.int64)) for _ in range(4)]
    meter = types.SimpleNamespace(intersection_buf=bufs[0], union_buf=bufs[1], intersection_buf_known_bad=bufs[2],
                                  union_buf_known_bad=bufs[3])
    AverageMeter.all_reduce(meter)  # the method only touches the four buffers; the class itself refuses CPU devices
    q.put((rank, [b.numpy().copy() for b in bufs]))
    dist.destroy_process_group()


def test_two_rank_all_reduce_of_the_evaluation_buffers():
    """world_size-2 gloo run of the sharded evaluation: after AverageMeter.all_reduce every rank holds the
    exact int64 sums of both ranks' intersection / union buffers (counts above 2^32: no float32 rounding as in logger.py)."""
    import torch.multiprocessing as mp

    nclass, world = 20, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 27500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_meter_all_reduce_worker, args=(r, world, port, nclass, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    want = [sum(np.random.RandomState(7 + r).randint(0, 2 ** 40, size=(4, 2, nclass)).astype(np.int64)[k] for r in range(world))
            for k in range(4)]
    for r in range(world):
        for k in range(4):
            np.testing.assert_array_equal(got[r][k], want[k])
