#!/usr/bin/env python
"""Benchmark of the MARS ranking stage (BASELINE.json metric: episodes/s, config c2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A *step* is one pass of the whole hot path over one batch of `--episodes-per-step` synthetic
episodes per GPU (inputs resident in HBM for `value`; pinned host buffers + H2D/D2H inside the
timed region for `e2e`).  Under torchrun (N > 1) every rank processes its own shard of episodes and
the per-episode result records are all-gathered over NCCL each step; the time is the max over
ranks, measured with CUDA events between barrier + synchronize.

`--impl reference` times the reference's algorithm on the host CPU (the oracle port: the reference
is Python and cannot travel to the GPU box), on the same workload, one episode per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--episodes-per-step", type=int, default=16)
    ap.add_argument("--e2e-episodes-per-step", type=int, default=2)
    ap.add_argument("--mask-dtype", default="f32", choices=["f32", "u8"])
    ap.add_argument("--nms", type=float, default=0.7)
    ap.add_argument("--fused-ingest", action="store_true", help="one-pass pack + pairwise kernel")
    ap.add_argument("--tensor-partition-sms", type=int, default=-1,
                    help="SMs of the tensor green-context partition for the device-resident loop (0 = one whole-device "
                         "timeline; -1 = 56 for float32 masks at c2, else 0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify-gather", action="store_true", help="skip the multi-GPU check of the gathered record table")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short c1 / c3 / c4 runs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-emd", action="store_true", help="skip the full-scoring (device EMD) variant")
    ap.add_argument("--cpu-emd-lps", type=int, default=4, help="transport LPs timed on the host for the EMD baseline")
    ap.add_argument("--cpu-sample-episodes", type=int, default=2)
    return ap.parse_args()


def kernel_source_sha16(kernel="pack_f32_vec_kernel", source="masks.cu"):
    """sha256 (first 16 hex digits) of the CUDA source of ONE kernel (its `__global__` definition up to the closing brace
    at column 0, plus the tuning constants above it): ties an ncu capture to the code it measured without going stale when
    an unrelated kernel of the same file changes."""
    import hashlib
    import re

    path = os.path.join(ROOT, "mars-multimodal-alignment-and-ranking-system-for-few-shot-segmentation_b200", "csrc", source)
    with open(path) as f:
        text = f.read()
    m = re.search(r"__global__[^;{]*\b" + re.escape(kernel) + r"\s*\(.*?\n}\n", text, re.S)
    consts = "".join(re.findall(r"^constexpr int PACK_\w+ = [^;]+;$", text, re.M))
    return hashlib.sha256(((m.group(0) if m else text) + consts).encode()).hexdigest()[:16]


def ncu_traffic_bytes(kernel, workload, episodes, mask_dtype):
    """dram read+write bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/ncu_traffic.json),
    or None when no capture of THIS version of the kernel source exists (a stale capture says nothing about new code)."""
    try:
        sha = kernel_source_sha16(kernel)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            for row in json.load(f):
                if (row["kernel"], row["workload"], row["episodes_per_launch"], row["mask_dtype"]) == \
                        (kernel, workload, episodes, mask_dtype) and row.get("source_sha16") == sha:
                    return row["dram_read_bytes"] + row["dram_write_bytes"]
    except Exception:
        pass
    return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML in a thread; nvidia-smi as fallback)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.windows = []  # [t0, t1] perf_counter intervals of the timed regions
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _nvml_loop(self):
        nv, h = self._nvml
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                t = time.perf_counter()
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.sm.append((t, mhz, [name for name, bit in bits.items() if r & bit]))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; map through the UUID torch reports
            props = torch.cuda.get_device_properties(self.index)
            uuid = "GPU-" + str(props.uuid) if not str(props.uuid).startswith("GPU-") else str(props.uuid)
            try:
                h = nv.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._nvml = (nv, h)
            self._thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self._thread.start()
            return
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self._thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self._nvml is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            inside = [x for x in self.sm if any(a <= x[0] <= b for a, b in self.windows)] or self.sm
            reasons = sorted({r for x in inside for r in x[2]})
            return {"sm_mhz": statistics.median(x[1] for x in inside) if inside else None, "sm_max_mhz": self.max_mhz,
                    "samples": len(inside), "samples_total": len(self.sm), "reasons": reasons, "source": "nvml",
                    "note": "samples taken inside the timed regions (device-resident loop and e2e loop)"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source available"]}
        time.sleep(0.15)
        self.proc.terminate()
        self._thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def workload_description(shape, args):
    return (f"{args.workload}: {shape.ns}-shot episode, N={shape.N} patches x C={shape.C}, P={shape.P} proposals at "
            f"{shape.H}x{shape.W} ({args.mask_dtype} masks), vva+PIR, vta PIR, region scores, AlphaCLIP cosine, "
            f"pairwise intersections + IoU-NMS {args.nms}, fuse/rank, merge; EMD scores are an input")


def workload_config(shape, args):
    """The `config` object of the JSON line: the workload only, identical for both arms (how each arm schedules it is
    reported under `schedule`)."""
    return {"workload": workload_description(shape, args), "shots": shape.ns, "patches": shape.N, "feature_width": shape.C,
            "proposals": shape.P, "proposal_resolution": [shape.H, shape.W], "mask_format": args.mask_dtype,
            "nms_iou_threshold": args.nms, "emd": "input vector",
            "l2": "inputs larger than L2: two resident batches alternate, each step reads far more than the 126 MB L2"}


def load_synthetic():
    """The synthetic-episode generator (pure torch) loaded by path: the reference arm must not import the product
    package (importing `marsb200` maps libmarsb200.so)."""
    import importlib.util

    path = os.path.join(ROOT, "mars-multimodal-alignment-and-ranking-system-for-few-shot-segmentation_b200", "synthetic.py")
    spec = importlib.util.spec_from_file_location("marsb200_synthetic_standalone", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def oracle_cfg(shape, args):
    return dict(g=shape.g, vva_box_threshold=0.8, vta_box_threshold=0.4, alpha=0.85, static_threshold=0.55,
                dynamic_threshold=0.95, nms_iou_threshold=args.nms)


# ------------------------------------------------------------------------------------------------
def cpu_reference_episodes(shape, args, episodes):
    """Run the oracle port on the given CPU episodes (dicts of CPU tensors) with all host threads; returns
    (seconds per episode, oracle results)."""
    from oracle import mars_oracle as orc

    torch.set_num_threads(os.cpu_count() or 1)
    try:  # torchrun exports OMP_NUM_THREADS=1: the host baseline gets every core in numpy's BLAS / OpenMP pools as well
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    cfg = oracle_cfg(shape, args)
    times, results = [], []
    for ep in episodes:
        ep = dict(ep)
        ep["masks"] = ep["masks"].float()  # the reference's wire format
        t0 = time.perf_counter()
        results.append(orc.run_episode(ep, cfg))
        times.append(time.perf_counter() - t0)
    return times, results


def cpu_emd_sample(shape, batch, n_lps):
    """Host time of the oracle's exact transport LP (HiGHS; POT is not installed) on a few proposals of episode 0."""
    from oracle import mars_oracle as orc

    ep = {k: v[0].cpu() for k, v in batch.items()}
    fs, fq = orc.normalize_rows(ep["feat_s"].reshape(-1, shape.C)), orc.normalize_rows(ep["feat_q"])
    _, cost = orc.similarity_and_cost(fs, fq)
    sup = orc.pool_mask(ep["support_mask"].float(), shape.g).reshape(-1)
    pm = orc.pool_mask(ep["masks"][:n_lps].float(), shape.g).reshape(n_lps, -1)
    t0 = time.perf_counter()
    for i in range(n_lps):
        orc.emd_score(sup, pm[i], cost)
    dt = time.perf_counter() - t0
    out = {"lps_per_s": n_lps / dt, "cores": 1, "kind": "port (HiGHS exact LP; POT 0.9.4 is not installed)",
           "sample": f"{n_lps} LPs of episode 0 (T={int(sup.sum())} support patches)"}
    try:
        # the algorithm class POT's ot.emd2 uses: a network simplex (oracle/emd_netsimplex.c), every LP of episode 0, one LP
        # per host thread at a time (the C call releases the GIL)
        from concurrent.futures import ThreadPoolExecutor

        from oracle.emd_c import emd_network_simplex_c

        pm_all = orc.pool_mask(ep["masks"].float(), shape.g).reshape(ep["masks"].shape[0], -1)
        rows = cost[sup.bool()].numpy().astype("float64")
        subs = [rows[:, pm_all[i].numpy()].copy() for i in range(pm_all.shape[0])]
        cores = os.cpu_count() or 1
        with ThreadPoolExecutor(cores) as pool:
            t0 = time.perf_counter()
            vals = list(pool.map(emd_network_simplex_c, subs))
            dt = time.perf_counter() - t0
        out["network_simplex"] = {"lps_per_s": len(subs) / dt, "cores": cores, "lps": len(subs), "seconds": dt,
                                  "kind": "port (network simplex in C, the algorithm class of POT's ot.emd2; oracle/emd_netsimplex.c)",
                                  "agrees_with_highs": bool(all(abs((1.0 - orc.emd_score(sup, pm[i], cost)) - vals[i]) < 1e-9
                                                                for i in range(min(n_lps, 2))))}
    except Exception as ex:  # the C oracle did not build on this box: the HiGHS figure above stands alone
        out["network_simplex"] = {"unavailable": repr(ex)}
    return out


def run_reference(args):
    """The reference arm: the reference's algorithm (oracle port) on the host cores only.  Nothing of the product is
    imported and CUDA is never initialised: episodes are generated on the CPU (two distinct ones, alternated)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    syn = load_synthetic()
    shape = syn.CONFIGS[args.workload]
    md = torch.float32 if args.mask_dtype == "f32" else torch.uint8
    eps = [syn.make_episode(shape, 10_000 + i, "cpu", md) for i in range(2)]
    cpu_reference_episodes(shape, args, eps[:1])  # one warm-up episode is enough on the CPU
    times, _ = cpu_reference_episodes(shape, args, [eps[i % 2] for i in range(args.steps)])
    total = sum(times)
    value = len(times) / total
    cores = os.cpu_count() or 1
    sample = f"{len(times)} episodes of {args.workload}, one per step, torch CPU with {cores} threads, EMD excluded"
    emit(json.dumps({
        "impl": "reference", "metric": "episodes_per_sec", "value": value, "unit": "episodes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(shape, args),
        "schedule": {"episodes_per_step": 1, "device": "host CPU", "threads": cores},
        "cpu_baseline": {"value": value, "unit": "episodes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    import marsb200
    from marsb200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (marsb200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    shape = marsb200.CONFIGS[args.workload]
    md = torch.float32 if args.mask_dtype == "f32" else torch.uint8
    cfg = marsb200.RankingConfig(nms_iou_threshold=args.nms, fused_ingest=args.fused_ingest)
    E = args.episodes_per_step
    part_sms = args.tensor_partition_sms
    if part_sms < 0:
        part_sms = 64 if (md == torch.float32 and args.workload == "c2" and not args.fused_ingest and E >= 8) else 0
    # the headline loop runs the ingest and the contractions on disjoint SM partitions (same kernels, same results) and
    # two buffer sets take the steps in turn, so the ingest of step i + 1 overlaps the scoring tail of step i; the
    # small-batch variants below (e2e, latency, full scoring) stay on one whole-device timeline
    cfg_main = marsb200.RankingConfig(nms_iou_threshold=args.nms, fused_ingest=args.fused_ingest,
                                      tensor_partition_sms=part_sms or None, partition_vta_on_hbm=False)
    pipe = None
    if part_sms:
        try:
            pipe = marsb200.PipelinedRanking(shape, E, cfg_main, dev, md, depth=2)
        except Exception as ex:  # no green contexts on this driver / cuda-python: same kernels on one timeline
            print(f"[bench] SM partitions unavailable ({ex!r}): running the one-timeline schedule", file=sys.stderr)
            part_sms, cfg_main = 0, cfg
    eng = pipe.engines[0] if pipe is not None else marsb200.RankingEngine(shape, E, cfg_main, dev, md)

    # two distinct resident batches, alternated: every step reads inputs far larger than the 126 MB L2
    n_batches = 2
    batches = []
    for b in range(n_batches):
        eps = [marsb200.make_episode(shape, (rank * n_batches + b) * E + i, dev, md) for i in range(E)]
        batches.append(marsb200.stack_episodes(eps))
        del eps
    bytes_per_step = sum(v.numel() * v.element_size() for v in batches[0].values())
    total_episodes = world * E

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:  # every engine's fuse / rank kernel writes its records into its rank's slice of an all-gather table
        for en in (pipe.engines if pipe is not None else [eng]):
            en.attach_gather_table(world, rank)

    def finish(ticket):
        # results of a pipelined step: join it on this stream, then (multi-GPU) all-gather its records, in place
        if pipe is not None:
            pipe.result(ticket)
        if world > 1:
            (pipe.engine(ticket) if pipe is not None else eng).gather()

    def run_steps(n):
        """n steps back to back; with the pipeline the results of step i are collected after step i + 1 is enqueued,
        the last one before returning (so every step's work and gather lie inside the caller's timed region)."""
        prev = None
        for i in range(n):
            if pipe is not None:
                ticket = pipe.submit(batches[i % n_batches])
                if prev is not None:
                    finish(prev)
                prev = ticket
            else:
                eng.run(batches[i % n_batches])
                finish(None)
        if prev is not None:
            finish(prev)

    run_steps(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    start.record()
    run_steps(args.steps)
    stop.record()
    barrier()
    sampler.windows.append((w0, time.perf_counter()))
    elapsed_ms = start.elapsed_time(stop)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = total_episodes * args.steps / (elapsed_ms / 1e3)

    # ---- the same device-resident loop fed with lighter proposal formats (uint8 masks, packed bits)
    def value_variant(kind, interleaved=False):
        inter, eng_v = None, None
        if kind == "one_timeline":
            alt, dt = batches, md
        elif kind == "u8":
            alt, dt = [dict(b, masks=(b["masks"] > 0).to(torch.uint8)) for b in batches], torch.uint8
        else:
            alt, dt = [{k: v for k, v in b.items() if k != "masks"} for b in batches], md
            for a, b in zip(alt, batches):
                a["mask_bits"] = ops.pack_masks(b["masks"])
        if interleaved:
            inter = marsb200.InterleavedRanking(shape, E, cfg, dev, dt, depth=2)
        else:
            eng_v = marsb200.RankingEngine(shape, E, cfg, dev, dt)
        def loop(n):
            if inter is None:
                for i in range(n):
                    eng_v.run(alt[i % n_batches])
                return
            prev = None
            for i in range(n):
                t = inter.submit(alt[i % n_batches])
                if prev is not None:
                    inter.result(prev)
                prev = t
            inter.result(prev)

        loop(args.warmup)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        loop(args.steps)
        a1.record()
        barrier()
        ms = a0.elapsed_time(a1) / args.steps
        return {"value": world * E / (ms / 1e3), "unit": "episodes/s", "ms_per_step": ms,
                "schedule": "one timeline" if inter is None else "two whole-device engines on two streams, steps interleaved"}

    value_variants = None
    if md == torch.float32 and not args.no_e2e:
        value_variants = {"u8_masks": value_variant("u8", True), "packed_masks": value_variant("bits", True),
                          "u8_masks_one_timeline": value_variant("u8"), "packed_masks_one_timeline": value_variant("bits")}
    if part_sms:
        value_variants = dict(value_variants or {}, one_timeline=value_variant("one_timeline"))

    # ---- per-kernel timing of the dominant kernels with CUDA events on the launching stream
    def time_kernel(fn, iters):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for i, (a, b) in enumerate(evs):
            a.record()
            fn(i)
            b.record()
        torch.cuda.synchronize()
        return statistics.mean(a.elapsed_time(b) for a, b in evs)

    iters = max(4, min(args.steps, 20))
    pack_ms = time_kernel(lambda i: ops.pack_masks(batches[i % n_batches]["masks"], out=eng.bits), iters)
    # the one-pass variant (packed bits + pooled bitmaps from the words in registers), kept as an op, not used by the engine
    pack_pool_ms = time_kernel(lambda i: ops.pack_pool(batches[i % n_batches]["masks"], shape.g, out_bits=eng.bits,
                                                       out_pool=eng.pool_out), iters)
    pool_ms = time_kernel(lambda i: ops.pool_packed(eng.bits, shape.H, shape.W, shape.g, out=eng.pool_out), iters)
    pair_ms = time_kernel(lambda i: ops.pairwise_inter(eng.bits, backend=cfg.pair_backend, out=eng.inter), iters)
    # the one-pass kernel exists for the int8 back end only (it owns all of TMEM); asked for explicitly here
    fused_ms = time_kernel(lambda i: ops.pack_pairwise(batches[i % n_batches]["masks"], backend=ops.PAIR_MMA,
                                                       out=(eng.bits, eng.inter)), iters)
    hw = shape.H * shape.W
    wpm = ops.words_per_mask(hw)
    pack_bytes = E * shape.P * (hw * (4 if md == torch.float32 else 1) + wpm * 4)
    peak, peak_kind = measured_peak_gbs()
    achieved = pack_bytes / (pack_ms / 1e3) / 1e9
    pairs = E * shape.P * (shape.P + 1) // 2
    roofline = {"kernel": "pack_masks (mask ingest -> packed bits)", "bound": "hbm", "achieved": achieved,
                "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_bytes("pack_f32_vec_kernel" if md == torch.float32 else "pack_u8_vec_kernel",
                                             args.workload, E, args.mask_dtype),
                "algorithmic_bytes_per_launch": pack_bytes, "ms_per_launch": pack_ms}
    step_ms = elapsed_ms / args.steps
    whole_step = {"input_bytes_per_step_per_gpu": bytes_per_step, "ms_per_step": step_ms,
                  "input_gbs": bytes_per_step / (step_ms / 1e3) / 1e9, "frac_of_hbm_peak": bytes_per_step / (step_ms / 1e3) / 1e9 / peak,
                  "note": "inputs only; the kernels of a step move ~1.22x the input bytes (packed bits, operands with their TF32 "
                          "residuals, attention, G: profiles/r2_ncu_full_top_kernels.csv), so the step as a whole is HBM-bound"}
    pairwise = {"kernel": "pairwise_inter", "unordered_pairs_per_s": pairs / (pair_ms / 1e3), "ms_per_launch": pair_ms,
                "word_ops_per_s": pairs * wpm / (pair_ms / 1e3)}
    fused_pool = {"kernel": "pack_pool (one pass: packed bits + pooled bitmaps)", "ms_per_launch": pack_pool_ms,
                  "two_kernels_ms": pack_ms + pool_ms, "note": "not used by the engine: slower than pack + pool_packed"}
    fused = {"kernel": "pack_pairwise (one pass: packed bits + intersections)", "ms_per_launch": fused_ms,
             "achieved_gbs": pack_bytes / (fused_ms / 1e3) / 1e9, "frac_of_hbm_peak": pack_bytes / (fused_ms / 1e3) / 1e9 / peak,
             "unordered_pairs_per_s": pairs / (fused_ms / 1e3)}

    # ---- end to end: pinned host inputs -> H2D -> ranking -> D2H of the result records, every step.
    # Two buffer sets (pinned host records, device inputs, engine) take the steps in turn: the H2D copy of step i + 1 runs on
    # a copy stream while step i computes, the D2H of step i's records runs behind its compute on a third stream and the host
    # reads them one step later.  Every byte of every step's inputs crosses PCIe inside the timed region (or, for the
    # `backbone_resident` variants, every byte of the PROPOSALS: in the reference pipeline the DINOv2 / CLIP tensors are
    # produced on the GPU and only the proposals are loaded from disk, main_MARS.py:62).
    rle_cache = {}

    # host threads of the ingest: the ranks of one node share its cores (and its DRAM, which both lanes read the proposals from)
    host_threads = max(1, (os.cpu_count() or 16) // world)

    host_cache = {}

    def run_e2e(e2e_dtype, wire="dense", episodes=None, resident_backbone=False, host_pack=None, steps=None, warmup=None):
        """host_pack = share of every episode's proposals that crosses PCIe raw (packed by the device kernel); the rest is
        packed by host threads in front of the copy (marsb200.HostMaskIngest).  None = everything raw (the plain path)."""
        Ee = min(episodes or args.e2e_episodes_per_step, E)
        n_steps, n_warm = steps or args.steps, warmup if warmup is not None else max(2, args.warmup)
        engs = [marsb200.RankingEngine(shape, Ee, cfg, dev, e2e_dtype) for _ in range(2)]
        ckey = (e2e_dtype, wire, Ee, resident_backbone)
        if ckey not in host_cache:  # the pinned host copy of the batch is the same for every ingest variant
            hb = {k: v[:Ee].cpu() for k, v in batches[0].items()}
            if wire == "rle":  # SAM's own output format: uncompressed COCO RLE, decoded on the device
                if Ee not in rle_cache:
                    rle_cache[Ee] = marsb200.masks_to_rle(hb["masks"].reshape(-1, shape.H, shape.W))
                hb.pop("masks")
                hb["mask_rle_counts"], hb["mask_rle_offsets"] = rle_cache[Ee]
            elif wire == "bits":  # proposals kept bit-packed by the producer
                hb["mask_bits"] = ops.pack_masks(batches[0]["masks"][:Ee]).cpu()
                hb.pop("masks")
            else:
                hb["masks"] = hb["masks"].to(e2e_dtype)
            mask_keys = [k for k in hb if k.startswith("mask")]
            moved = mask_keys if resident_backbone else list(hb)
            on_dev = {k: hb[k].to(dev) for k in hb if k not in moved}  # produced on the device by the backbones
            host_cache[ckey] = ({k: hb[k].pin_memory() for k in moved}, on_dev)
        host, fixed = dict(host_cache[ckey][0]), host_cache[ckey][1]
        ingests, host_masks = None, None
        if host_pack is not None and wire == "dense":
            host_masks = host.pop("masks")  # stays in pinned host memory; HostMaskIngest moves it in two lanes
            ingests = [marsb200.HostMaskIngest(Ee, shape.P, shape.H, shape.W, dev, mask_dtype=e2e_dtype, raw_fraction=host_pack,
                                               threads=host_threads) for _ in range(2)]
        dev_in = [dict({k: torch.empty_like(v, device=dev) for k, v in host.items()}, **fixed) for _ in range(2)]
        rec_host = [torch.empty((Ee, engs[0].record_bytes()), dtype=torch.uint8).pin_memory() for _ in range(2)]
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        if ingests:
            h2d += ingests[0].h2d_bytes(host_masks.element_size())
        d2h = rec_host[0].numel()
        main = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        ev_in = [torch.cuda.Event() for _ in range(2)]     # inputs of the set have landed
        ev_done = [torch.cuda.Event() for _ in range(2)]   # the step that used the set has been computed
        ev_out = [torch.cuda.Event() for _ in range(2)]    # its records are in host memory

        def upload(i):
            b = i % 2
            s_in.wait_event(ev_done[b])  # the set's previous step no longer reads these inputs
            with torch.cuda.stream(s_in):
                for k in host:
                    dev_in[b][k].copy_(host[k], non_blocking=True)
            if ingests:  # raw lane + device packing enqueued on s_in, the other lane packed by host threads meanwhile
                dev_in[b]["mask_bits"] = ingests[b].upload(host_masks, s_in)
            ev_in[b].record(s_in)

        def loop(n):
            for b in range(2):
                ev_done[b].record(main)
            upload(0)
            for i in range(n):
                b = i % 2
                main.wait_event(ev_in[b])
                main.wait_event(ev_out[b])  # the records of the set's previous step have left the device
                engs[b].run(dev_in[b])
                rec = engs[b].records()
                ev_done[b].record(main)
                s_out.wait_event(ev_done[b])
                with torch.cuda.stream(s_out):
                    rec_host[b].copy_(rec, non_blocking=True)
                    ev_out[b].record(s_out)
                rec.record_stream(s_out)
                if i + 1 < n:
                    upload(i + 1)  # the copies (and the host-side packing) of step i + 1 run while the device ranks step i
                if i >= 1:
                    ev_out[(i - 1) % 2].synchronize()  # the caller reads step i - 1's result on the host
            ev_out[(n - 1) % 2].synchronize()

        for b in range(2):
            ev_out[b].record(s_out)
        loop(n_warm)
        barrier()
        s2, t2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        s2.record()
        loop(n_steps)
        main.wait_stream(s_out)
        t2.record()
        barrier()
        sampler.windows.append((w0, time.perf_counter()))
        ms2 = s2.elapsed_time(t2)
        if world > 1:
            t = torch.tensor([ms2], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        # what the loop produced against the device-resident engine on the same episodes (records: order, scores, flags)
        out = {}
        if e2e_dtype == md and wire == "dense":
            engs[0].run({k: v[:Ee] for k, v in batches[0].items()})
            out["records_equal_device_resident_run"] = bool(torch.equal(rec_host[(n_steps - 1) % 2], engs[0].records().cpu()))
        if ingests:
            out.update({"host_pack": {"raw_fraction": host_pack, "proposals_raw_over_pcie": ingests[0].p_raw,
                                      "proposals_packed_by_host_threads": shape.P - ingests[0].p_raw,
                                      "host_threads": host_threads or os.cpu_count()}})
        out.update({"value": world * Ee * n_steps / (ms2 / 1e3), "unit": "episodes/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "episodes_per_step": Ee, "ms_per_step": ms2 / n_steps,
                "pcie_gbs": h2d * n_steps / (ms2 / 1e3) / 1e9,
                "inputs_over_pcie": "proposals only (backbone tensors produced on the device)" if resident_backbone
                                    else "every input of the step",
                "host_mask_format": {"dense": "f32" if e2e_dtype == torch.float32 else "u8", "rle": "uncompressed COCO RLE",
                                     "bits": "packed bits"}[wire]})
        return out

    e2e, e2e_variants = None, None
    if not args.no_e2e:
        # Host buffers in, records out.  A PCIe 5 link carries the reference's float32 proposals (1.07 GB per c2 episode) at
        # 55 GB/s = 50 episodes/s; the host's cores read them at ~116 GB/s.  The ingest therefore runs two lanes at once
        # (marsb200.HostMaskIngest): a share of every episode's proposals crosses PCIe raw and is packed by the device
        # kernel, the rest is packed by host threads and only its bits are copied.  The share balances the two lanes.
        # which share goes raw depends on the box (cores, DRAM, how many ranks share them): a short calibration picks it
        # before the timed loop (one rank: the optimum is flat around 0.25-0.3, profiles/r2_logs/e2e_host_pack_sweep.log;
        # from ~4 ranks on the ranks' PCIe links alone use all the host's DRAM delivers and everything goes raw)
        calib = {c: run_e2e(md, host_pack=c if c < 1.0 else None, steps=4, warmup=2)["value"] for c in (0.15, 0.3, 0.45, 0.7, 1.0)}
        raw_share = max(calib, key=calib.get)
        e2e = run_e2e(md, host_pack=raw_share if raw_share < 1.0 else None)
        e2e["raw_share_calibration"] = {"episodes_per_s_by_raw_share": calib, "chosen": raw_share,
                                        "how": "4 timed steps per candidate before the timed loop (max over ranks)"}
        e2e["note"] = ("PCIe-bound: the reference's float32 wire format is 1.1 GB per episode and this many ranks' PCIe links already "
                       "draw all the host's DRAM delivers, so every proposal crosses PCIe raw (single-lane ingest)") if raw_share >= 1.0 else \
                      ("two-lane ingest of the reference's float32 host proposals: `proposals_raw_over_pcie` of every episode cross "
                       "PCIe as float32 and are packed by the device kernel while host threads pack the others (a format "
                       "conversion in front of the copy, bit-identical; nothing is scored on the host); every byte of the step's "
                       "inputs still starts in pinned host memory inside the timed region.  `e2e_variants."
                       "f32_host_masks_all_raw_over_pcie` is the single-lane path (PCIe-bound at ~55 GB/s)")
        if md == torch.float32:  # the same call with lighter proposal wire formats (PCIe carries far fewer bytes)
            Ev = min(E, 16)
            e2e_variants = {"f32_host_masks_all_raw_over_pcie": run_e2e(md),
                            "f32_host_masks_all_packed_by_host_threads": run_e2e(md, host_pack=0.0),
                            "u8_host_masks": run_e2e(torch.uint8, episodes=min(4, E)),
                            "u8_host_masks_two_lane_ingest": run_e2e(torch.uint8, episodes=min(4, E),
                                                                     host_pack=raw_share if raw_share < 1.0 else None),
                            "packed_host_masks": run_e2e(md, "bits", episodes=Ev),
                            "packed_host_masks_backbone_resident": run_e2e(md, "bits", episodes=Ev, resident_backbone=True)}
            if shape.H % 32 == 0 and shape.W % 32 == 0:  # the device RLE decoder needs word-aligned rows and columns
                e2e_variants["rle_host_masks"] = run_e2e(md, "rle", episodes=Ev)
                e2e_variants["rle_host_masks_backbone_resident"] = run_e2e(md, "rle", episodes=Ev, resident_backbone=True)

    # ---- single-episode latency (the reference ranks one episode at a time, main_MARS.py:54-94): eager launches
    # and one CUDA-graph replay of the same kernel sequence
    def latency():
        eng1 = marsb200.RankingEngine(shape, 1, cfg, dev, md)
        one = [{k: v[i:i + 1].contiguous() for k, v in batches[0].items()} for i in range(2)]
        for i in range(3):
            eng1.run(one[i % 2])
        eager = time_kernel(lambda i: eng1.run(one[i % 2]), 10)
        eng1.capture(one[0])
        for i in range(3):
            eng1.replay()
        graph = time_kernel(lambda i: eng1.replay(), 10)
        return {"episodes": 1, "eager_ms": eager, "cuda_graph_ms": graph,
                "note": "inputs resident in HBM; one episode = 1.07 GB of float32 masks (163 us at the HBM peak)"}

    lat = latency() if not args.no_e2e else None

    # ---- full scoring: the P transport LPs per episode solved on the device as well
    full = None
    if not args.no_emd:
        Ef = min(E, 8)
        sub = [{k: v[:Ef] for k, v in b.items()} for b in batches]
        # a deployment sizes the solver's state for the largest proposal it admits; here: the largest pooled
        # proposal of the synthetic set, rounded up (read once, outside the timed region)
        m_cap = max(int(ops.pool_packed(ops.pack_masks(b["masks"]), shape.H, shape.W, shape.g)[2].max()) for b in sub)
        m_cap = min(shape.N, (m_cap + 63) // 64 * 64)
        t_cap = max(int(ops.pool_mask(b["support_mask"], shape.g).reshape(Ef, -1).sum(1).max()) for b in sub)
        t_cap = min(shape.ns * shape.N, (t_cap + 63) // 64 * 64)
        cfg_f = marsb200.RankingConfig(nms_iou_threshold=args.nms, emd_on_device=True, emd_m_cap=m_cap, emd_t_cap=t_cap)
        eng_f = marsb200.RankingEngine(shape, Ef, cfg_f, dev, md)
        for i in range(2):
            eng_f.run(sub[i % n_batches])
        barrier()
        s3, t3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_full = max(2, min(args.steps, 4))
        w0 = time.perf_counter()
        s3.record()
        for i in range(n_full):
            eng_f.run(sub[i % n_batches])
        t3.record()
        barrier()
        sampler.windows.append((w0, time.perf_counter()))
        ms3 = s3.elapsed_time(t3) / n_full
        emd_each = []
        for b in range(n_batches):  # the LPs of BOTH resident batches (the timed loop above alternates between them)
            eng_f.run(sub[b])
            emd_each.append(time_kernel(lambda i: ops.emd_scores(eng_f.gemm_out["cost"], eng_f.row_fg.reshape(Ef, -1), eng_f.pool_out[0],
                                                                 t_cap=eng_f.emd_t_cap, m_cap=eng_f.emd_m_cap, workspace=eng_f.emd_ws,
                                                                 out=eng_f.emd_out, check=False), 3))
        emd_ms = statistics.mean(emd_each)
        full = {"what": "the same step with the P transport LPs per episode (ot.emd2) solved exactly on the device",
                "episodes_per_step": Ef, "ms_per_step": ms3, "value": world * Ef / (ms3 / 1e3), "unit": "episodes/s",
                "emd_kernel_ms": emd_ms, "emd_lps_per_s": Ef * shape.P / (emd_ms / 1e3), "emd_m_cap": m_cap, "emd_t_cap": t_cap}
        del eng_f

    # ---- the gathered record table equals the rank-local results (multi-GPU only): rank 0 regenerates every other rank's
    # batch 0 from its seeds, ranks it locally and compares with that rank's slice of the all-gathered table
    gather_check = None
    if world > 1 and not args.no_verify_gather:
        if pipe is not None:
            out = pipe.result(pipe.submit(batches[0]))
            en = pipe.engine(pipe._next - 1)
        else:
            eng.run(batches[0])
            en = eng
        local = en.records().clone()
        table = en.gather().clone()  # the same in-place collective the timed loop runs
        torch.cuda.synchronize()
        ok_local = bool(torch.equal(table[rank * E:(rank + 1) * E], local))
        bad = []
        if rank == 0:
            chk = marsb200.RankingEngine(shape, E, cfg, dev, md)
            for r in range(1, world):
                eps = [marsb200.make_episode(shape, (r * n_batches + 0) * E + i, dev, md) for i in range(E)]
                chk.run(marsb200.stack_episodes(eps))
                del eps
                if not torch.equal(table[r * E:(r + 1) * E], chk.records()):
                    bad.append(r)
            del chk
        flag = torch.tensor([0 if ok_local else 1], device=dev)
        dist.all_reduce(flag)
        gather_check = {"ranks": world, "episodes": total_episodes, "ok": bool(flag.item() == 0) and not bad,
                        "ranks_with_wrong_slices": bad,
                        "how": "every rank: own slice of the gathered table == local records; rank 0: recomputed every other "
                               "rank's episodes from their seeds on its own GPU, bit-equal records"}

    # ---- the other BASELINE configs, device-resident, one timeline, a few steps each (driver-visible)
    def other_config(name, episodes, mask_dt):
        shp = marsb200.CONFIGS[name]
        dt = torch.float32 if mask_dt == "f32" else torch.uint8
        bs = [marsb200.stack_episodes([marsb200.make_episode(shp, 777 + (rank * 2 + b) * episodes + i, dev, dt)
                                       for i in range(episodes)]) for b in range(2)]
        en = marsb200.RankingEngine(shp, episodes, cfg, dev, dt)
        for i in range(3):
            en.run(bs[i % 2])
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(3):
            en.run(bs[i % 2])
        a1.record()
        barrier()
        ms = a0.elapsed_time(a1) / 3
        pk = time_kernel(lambda i: ops.pack_masks(bs[i % 2]["masks"], out=en.bits), 4)
        pr = time_kernel(lambda i: ops.pairwise_inter(en.bits, backend=cfg.pair_backend, out=en.inter), 4)
        hw_, wpm_ = shp.H * shp.W, ops.words_per_mask(shp.H * shp.W)
        pack_b = episodes * shp.P * (hw_ * (4 if dt == torch.float32 else 1) + wpm_ * 4)
        prs = episodes * shp.P * (shp.P + 1) // 2
        res = {"episodes_per_step": episodes, "mask_format": mask_dt, "value": world * episodes / (ms / 1e3),
               "unit": "episodes/s", "ms_per_step": ms, "proposals": shp.P, "shots": shp.ns,
               "resolution": [shp.H, shp.W],
               "pack_kernel": {"ms": pk, "gbs": pack_b / (pk / 1e3) / 1e9, "frac_of_hbm_peak": pack_b / (pk / 1e3) / 1e9 / peak},
               "pairwise_kernel": {"ms": pr, "unordered_pairs_per_s": prs / (pr / 1e3),
                                   "backend": "kind::mxf4" if shp.P <= 256 else "kind::i8"},
               "dominant_kernel": "pack_masks" if pk >= pr else "pairwise_inter"}
        del en, bs
        torch.cuda.empty_cache()
        return res

    other = None
    if args.workload == "c2" and not args.no_other_configs:
        other = {"c1": other_config("c1", 16, "f32"), "c3": other_config("c3", 8, "f32"),
                 "c4": other_config("c4", 2, "f32")}

    clocks = sampler.stop()
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the CPU baseline runs on the SAME episodes the timed loop ranked (the first ones of resident batch 0) and
        # doubles as the checker of what that loop produced: one more step of the timed schedule, then
        # order / scores / intersections / keep-set / selection / merged mask against the oracle
        n_chk = min(args.cpu_sample_episodes, E)
        if pipe is not None:
            out = pipe.result(pipe.submit(batches[0]))
        else:
            out = eng.run(batches[0])
        torch.cuda.synchronize()
        keys = ("row_fg", "prior", "vva", "vta", "pooled", "clip", "inter", "area", "scores", "order", "flags", "merged_bits")
        got = {k: out[k][:n_chk].cpu() for k in keys if out.get(k) is not None}
        eps_cpu = [{k: v[i].cpu() for k, v in batches[0].items()} for i in range(n_chk)]
        times, refs = cpu_reference_episodes(shape, args, eps_cpu)
        cores = os.cpu_count() or 1
        cpu_baseline = {"value": len(times) / sum(times), "unit": "episodes/s", "cores": cores, "kind": "port",
                        "sample": f"{len(times)} episodes of {args.workload} through oracle.run_episode "
                                  f"(torch CPU, {cores} threads, EMD excluded): the first episodes of the timed loop's "
                                  f"resident batch 0"}
        from oracle import compare

        fails, swaps = [], 0
        for i in range(n_chk):
            res = compare.compare_episode(refs[i], {k: v[i] for k, v in got.items()}, eps_cpu[i]["masks"].float(),
                                          oracle_cfg(shape, args))
            fails += [f"episode {i}: {f}" for f in res["failures"]]
            swaps += res["tie_swaps"]
        parity = {"episodes": n_chk, "ok": not fails, "tie_swaps_within_tolerance": swaps, "failures": fails[:8],
                  "checked": "support bits, prior, vva, vta, pooled bitmaps, clip, intersections, areas (bit-exact), "
                             "scores (1e-4 rel), order, NMS keep-set, selection, merged mask (bit-exact) of the schedule "
                             "the headline times"}
        del got, eps_cpu, refs
        if full is not None:
            try:  # a host-side side figure must never cost the run its JSON line
                full["cpu_emd"] = cpu_emd_sample(shape, batches[0], args.cpu_emd_lps)
                ns = full["cpu_emd"].get("network_simplex", {})
                if "lps_per_s" in ns:
                    per_episode = 1.0 / cpu_baseline["value"] + shape.P / ns["lps_per_s"]
                    how = ("host time of one episode without EMD (all cores) plus its P transport LPs at the measured all-core rate "
                           "of the C network simplex (every LP of episode 0 was solved)")
                else:
                    per_episode = 1.0 / cpu_baseline["value"] + shape.P / full["cpu_emd"]["lps_per_s"]
                    how = ("extrapolated: host time of one episode without EMD (all cores) plus P transport LPs at the sampled "
                           "one-core HiGHS rate")
                full["cpu_full_scoring"] = {"value": 1.0 / per_episode, "unit": "episodes/s", "how": how}
            except Exception as ex:
                full["cpu_emd"] = {"error": repr(ex)}

    if rank == 0:
        emit(json.dumps({
            "metric": "episodes_per_sec", "value": value, "unit": "episodes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+fp32(3xtf32)",
            "data": "synthetic",
            "config": workload_config(shape, args),
            "schedule": {"episodes_per_step_per_gpu": E, "input_bytes_per_step_per_gpu": bytes_per_step,
                       "gemm_backend": "tcgen05" if ops.DEFAULT_GEMM == ops.GEMM_TCGEN05 else "simt",
                       "pair_backend": {ops.PAIR_POPC: "popc", ops.PAIR_MMA: "mma (kind::i8)", ops.PAIR_FP4: "fp4 (kind::mxf4)",
                                        ops.PAIR_AUTO: "auto: kind::mxf4 for P <= 256, kind::i8 above"}[ops.DEFAULT_PAIR],
                       "fused_ingest": bool(args.fused_ingest),
                       "sm_partition": ({"tensor_sms": eng._part.tensor_sms, "hbm_sms": eng._part.hbm_sms,
                                         "chunks": len(eng._chunks), "pipeline_depth": len(pipe.engines)}
                                        if pipe is not None else None)},
            "clocks": clocks, "value_variants": value_variants, "e2e": e2e, "e2e_variants": e2e_variants,
            "gpu_launches": marsb200.kernel_launches_per_run(cfg_main, E) * args.steps * world,
            "roofline": roofline, "whole_step": whole_step, "pairwise": pairwise, "fused_ingest": fused, "fused_pool": fused_pool, "full_scoring": full,
            "single_episode_latency": lat, "cpu_baseline": cpu_baseline, "parity_checked": parity,
            "gather_verified": gather_check, "other_configs": other,
        }))
    if world > 1:
        dist.destroy_process_group()


_RESULT_STREAM = None


def emit(line: str) -> None:
    """The one JSON line of the contract, on the process's original stdout."""
    out = _RESULT_STREAM or sys.stdout
    out.write(line + "\n")
    out.flush()


if __name__ == "__main__":
    # stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL prints its version there) goes to
    # stderr for the rest of the run
    sys.stdout.flush()
    _RESULT_STREAM = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
