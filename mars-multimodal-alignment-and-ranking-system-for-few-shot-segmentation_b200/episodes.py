"""Batched episode engine: the whole ranking stage for E episodes per launch sequence.

One `RankingEngine` owns every intermediate buffer for a fixed episode shape
(nothing is allocated inside `run`), enqueues the kernel sequence on the
current stream without a single host synchronisation, and can therefore be
captured into a CUDA graph (`capture`).  Episodes are independent
(main_MARS.py:54-94 of the reference processes them one by one), so multi-GPU
execution shards the episode list across ranks and all-gathers the fixed-size
result records.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from .partition import set_stream_sm_cap
from .synthetic import EpisodeShape


@dataclass
class RankingConfig:
    vva_box_threshold: float = 0.8     # main_MARS.py:151
    vta_box_threshold: float = 0.4     # main_MARS.py:144
    alpha: float = 0.85                # main_MARS.py:157 (alpha_coverage)
    static_threshold: float = 0.55     # main_MARS.py:155
    dynamic_threshold: float = 0.95    # main_MARS.py:156
    nms_iou_threshold: Optional[float] = None  # None = the reference's behaviour (no mask NMS)
    want_sim: bool = False             # export S like VisualVisualAlignmentModule.similarity_matrix
    want_cost: bool = False            # export (1 - S) / 2 (the EMD cost)
    want_merged_f32: bool = True       # the float32 [H, W] map the reference returns
    overlap_streams: bool = True       # mask chain (HBM-bound) on a second stream beside the contractions
    # the alignment chains on high-priority streams: their CTAs are placed before the pending CTAs of the ingest kernel
    # (whose grid alone fills the device for the whole read of the masks): single-episode graph latency 0.43 -> 0.34 ms,
    # one-timeline throughput +1.5 % (float32) ... +5 % (packed proposals)
    priority_streams: bool = True
    # one-timeline schedule: the prior-independent half of the vva refinement (attention normalisation, D D^T) runs on its
    # own stream beside normalise -> S -> prior instead of behind it: the alignment chain is the critical path of a small
    # batch (graph latency 0.33 -> 0.305 ms); at 16 episodes per step it is neutral to -3 %.  None = on for <= 2 episodes
    hoist_vva_contraction: Optional[bool] = None
    # one-timeline schedule: CTAs the persistent contractions of the alignment streams may start (marsb200_stream_set_sm_cap).
    # The ingest of ONE episode needs ~100 SMs to read its masks at the HBM rate; contractions that take the whole device
    # stall it (graph latency 0.308 -> 0.294 ms at 40, profiles/r2_logs/latency_contraction_cap.log).  None = 40 for
    # <= 2 episodes per step of DENSE masks (packed / RLE proposals have no such read: never capped), no cap above; 0 = never
    latency_contraction_sms: Optional[int] = None
    # one-timeline schedule: pixel slices the ingest of a step is cut into so that the intersections of slice k run beside the
    # read of slice k + 1 (marsb200_pack_masks_slice / marsb200_pairwise_inter_slice; masks whose pixels fill whole 512-word
    # blocks).  Exact, but measured SLOWER on one c2 episode (graph latency 0.292 -> 0.317 / 0.373 ms with 2 / 4 slices; at best
    # equal with the intersections capped at 64 SMs, profiles/r2_logs/latency_ingest_slices.log): an ingest CTA and a
    # 197 KB tensor-core CTA do not share an SM, so the slices serialise and each pays the intersections' fixed cost.  Off.
    latency_ingest_slices: int = 1
    fused_ingest: bool = False         # one-pass pack + pairwise kernel (owns all TMEM: cannot overlap the contractions)
    fused_pool: bool = False           # one-pass pack + pooled bitmaps (ops.pack_pool); measured slower than the two kernels
    emd_on_device: bool = False        # solve the P transport LPs per episode on the device instead of taking batch["emd"]
    # fast-path sizing of the EMD solver (state in shared memory); problems beyond it take its global-state launch, so
    # these are performance hints, never capacity limits
    emd_t_cap: Optional[int] = None    # fg support rows (default min(ns*N, 2048))
    emd_m_cap: Optional[int] = None    # pooled patches of a proposal (default N); smaller -> more LPs per SM
    gemm_backend: Optional[int] = None
    pair_backend: Optional[int] = None
    # SMs of the `tensor` green-context partition (partition.py); the mask ingest gets the rest of the device and runs
    # beside the contractions instead of before them.  None = one whole-device timeline (the streams above).
    tensor_partition_sms: Optional[int] = None
    partition_chunks: int = 4          # episode chunks pack -> pairwise are pipelined in across the two partitions
    partition_pairwise_tail: int = 0   # intersections of the last chunks run on the `hbm` partition after the ingest
    partition_vta_on_hbm: bool = True  # vta refinement on the `hbm` partition after the ingest (else beside the vva chain)
    partition_pool_on_tensor: bool = False  # pooled bitmaps of a chunk on the `tensor` partition (the `hbm` one only packs)
    # the streaming, prior-independent preparation (row normalisation of the features, Sinkhorn normalisation of both
    # attention matrices) runs on the `hbm` partition in front of the ingest instead of on the `tensor` partition
    partition_prep_on_hbm: bool = False
    partition_pool_side_stream: bool = False  # pooled bitmaps on a second stream of the `hbm` partition, beside the next chunk's pack


def kernel_launches_per_run(cfg: RankingConfig, episodes_per_batch: Optional[int] = None) -> int:
    """How many of our kernels one `RankingEngine.run` launches (counted from the sequence below)."""
    n = 0
    if cfg.tensor_partition_sms and not cfg.fused_ingest:
        # pack, pool_packed and pairwise run once per episode chunk instead of once per batch
        k = max(1, cfg.partition_chunks if episodes_per_batch is None else min(cfg.partition_chunks, episodes_per_batch))
        if episodes_per_batch is not None:
            per = (episodes_per_batch + k - 1) // k
            k = (episodes_per_batch + per - 1) // per
        n += (k - 1) * (2 + (2 if cfg.nms_iou_threshold is not None else 0))  # pack, pool_packed, pairwise + suppression relation
    n += 2                     # normalize_rows x2
    n += 1                     # pool_mask
    n += 1                     # sim_contract
    n += 1                     # vva_finalize
    n += 2 * (1 + 3 + 1 + 2)   # two PIR passes: box mask, colsum x2 + rownorm, contraction, two mat-vecs
    n += 1                     # min-max of the refined vva
    n += 1                     # resize + min-max of the vta
    n += 1                     # pack (memset not counted)
    if cfg.nms_iou_threshold is not None and not cfg.fused_ingest:
        n += 1                 # pairwise intersections (part of the pack kernel with fused_ingest)
    if cfg.nms_iou_threshold is not None:
        n += 1                 # suppression relation from the intersections (P <= 1024; larger P builds it while ranking)
    n += 1                     # pool_packed
    n += 1                     # region sums + union count (one launch)
    if cfg.emd_on_device:
        n += 6                 # exact EMD: problem sizes, duplicate links + marks, processing order, the solver, copies
    n += 1                     # clip scores
    n += 1                     # fuse / rank / nms / select
    n += 1                     # merge
    return n


class RankingEngine:
    def __init__(self, shape: EpisodeShape, episodes_per_batch: int, cfg: RankingConfig, device,
                 mask_dtype=torch.float32, partition=None):
        if torch.device(device).type != "cuda":
            raise RuntimeError("RankingEngine needs a CUDA device (marsb200 has no CPU path)")
        self.shape, self.E, self.cfg, self.device = shape, episodes_per_batch, cfg, torch.device(device)
        self.mask_dtype = mask_dtype
        s, e = shape, episodes_per_batch
        n, m, nt = s.N, s.ns * s.N, s.gt * s.gt
        dev = self.device
        f32, i32, u8 = torch.float32, torch.int32, torch.uint8
        new = lambda shp, dt: torch.empty(shp, device=dev, dtype=dt)
        self.fs = (new((e, ops.pad_rows(m), ops.pad_k(s.C)), f32), new((e, ops.pad_rows(m), ops.pad_k(s.C)), f32))
        self.fq = (new((e, ops.pad_rows(n), ops.pad_k(s.C)), f32), new((e, ops.pad_rows(n), ops.pad_k(s.C)), f32))
        self.row_fg = new((e, s.ns, n), u8)
        self.gemm_out = dict(colstats=new((e, ops.pad_rows(m) // 128, 4, n), f32))
        if cfg.want_sim:
            self.gemm_out["sim"] = new((e, m, n), f32)
        if cfg.want_cost or cfg.emd_on_device:
            self.gemm_out["cost"] = new((e, m, n), f32)
        self.emd_ws = None
        if cfg.emd_on_device:
            self.emd_t_cap = cfg.emd_t_cap or min(m, 2048)
            self.emd_m_cap = min(cfg.emd_m_cap or n, n)
            nbytes = int(ops.lib.marsb200_emd_workspace_bytes(e, s.P, n, m, self.emd_t_cap, self.emd_m_cap))
            self.emd_ws = new((nbytes,), u8)
            self.emd_out = new((e, s.P), torch.float64)
            self.emd_status = torch.zeros(1, device=dev, dtype=i32)  # persistent: read by check_status()
        self.prior = new((e, n), f32)
        self.vva = new((e, n), f32)
        self.vta_ref = new((e, nt), f32)
        self.vta = new((e, n), f32)
        self.pir_ws = ops.pir_workspace(e, n, dev)
        self.pir_ws_vta = ops.pir_workspace(e, nt, dev)  # own workspace: the two refinements run concurrently
        wpm = ops.words_per_mask(s.H * s.W)
        npw = (n + 31) // 32
        self.bits = new((e, s.P, wpm), i32)
        self.pool_out = (new((e, s.P, npw), i32), new((e, s.P), i32), new((e, s.P), i32))
        self.region_out = (new((e, s.P), f32), new((e, s.P), f32), new((e,), i32))
        self.inter = new((e, s.P, s.P), i32) if cfg.nms_iou_threshold is not None else None
        # pairwise suppression relation (proposal space), computed right behind the intersections: the ranking kernel only scans it
        self.nms_bits = new((e, s.P, (s.P + 31) // 32), i32) if cfg.nms_iou_threshold is not None and s.P <= 1024 else None
        self.clip = new((e, s.P), f32)
        self.rank_out = dict(scores=new((e, s.P), torch.float64), order=new((e, s.P), i32),
                             flags=new((e, s.P), u8), summary=new((e, 4), i32))
        # result records [E, record_bytes], written by fuse_rank itself; `attach_gather_table` re-points this at the rank's
        # slice of the all-gather table so that nothing is copied before the collective
        self.record_buf = new((e, ops.record_bytes(s.P)), u8)
        self._gather_table = None
        self.merge_out = dict(bits=new((e, wpm), i32))
        if cfg.want_merged_f32:
            self.merge_out["f32"] = new((e, s.H * s.W), f32)
        self._graph = None
        self._static = None
        self._rle_ws = None
        self._side = torch.cuda.Stream(device=dev) if cfg.overlap_streams else None
        self._side2 = torch.cuda.Stream(device=dev) if cfg.overlap_streams else None
        hi = dict(priority=-1) if cfg.priority_streams else {}
        self._side3 = torch.cuda.Stream(device=dev, **hi) if cfg.overlap_streams else None
        self._hi = torch.cuda.Stream(device=dev, priority=-1) if cfg.overlap_streams and cfg.priority_streams else None
        self._ev_hi = torch.cuda.Event()
        self._ev_rowfg = torch.cuda.Event()
        hoist = cfg.hoist_vva_contraction if cfg.hoist_vva_contraction is not None else e <= 2
        self._side4 = torch.cuda.Stream(device=dev, **hi) if cfg.overlap_streams and hoist else None
        cap = cfg.latency_contraction_sms if cfg.latency_contraction_sms is not None else (40 if e <= 2 else 0)
        self._contraction_sms = cap if (cfg.overlap_streams and cfg.priority_streams) else 0
        # pixel-sliced ingest of a small batch (see _mask_chain): slices of whole 512-word blocks that tile the mask exactly
        k = max(1, int(cfg.latency_ingest_slices))
        sliceable = (cfg.overlap_streams and cfg.nms_iou_threshold is not None and not cfg.fused_ingest and not cfg.fused_pool
                     and k > 1 and wpm * 32 == s.H * s.W and wpm % (512 * k) == 0 and s.H * s.W % 16 == 0
                     and (cfg.pair_backend if cfg.pair_backend is not None else ops.DEFAULT_PAIR) != ops.PAIR_POPC)
        self._slices = k if sliceable else 1
        self._pair = torch.cuda.Stream(device=dev, **hi) if sliceable else None
        self._ev_slice = [torch.cuda.Event() for _ in range(self._slices)]
        self._ev_pair = torch.cuda.Event()
        self._ev_vva_g = torch.cuda.Event()
        self._ev_vta = torch.cuda.Event()
        self._ev_fork = torch.cuda.Event()
        self._ev_join = torch.cuda.Event()
        self._ev_pack = torch.cuda.Event()
        self._ev_pool = torch.cuda.Event()
        self._ev_prep = torch.cuda.Event()
        self._part = None
        self._capturing = False
        self._pool_ready = False
        if cfg.tensor_partition_sms:
            from .partition import SmPartition

            # engines of one pipeline (PipelinedRanking) share the partition and its streams
            self._part = partition if partition is not None else SmPartition(dev, cfg.tensor_partition_sms)
            if not hasattr(self._part, "side_stream"):
                self._part.side_stream = self._part.extra_stream("tensor")
            self._part_side = self._part.side_stream
            if cfg.partition_pool_side_stream and not hasattr(self._part, "hbm_side_stream"):
                self._part.hbm_side_stream = self._part.extra_stream("hbm")
            self._pending = False  # a partitioned step has been enqueued and not yet joined
            k = max(1, min(cfg.partition_chunks, e))
            per = (e + k - 1) // k
            self._chunks = [(lo, min(lo + per, e)) for lo in range(0, e, per)]
            self._ev_chunk = [torch.cuda.Event() for _ in self._chunks]

    # ------------------------------------------------------------------ the kernel sequence
    def _ingest(self, batch: dict):
        """Proposals -> packed bits.  Wire formats: `masks` [E,P,H,W] float32 / uint8 / bool (the reference's tensors),
        `mask_rle_counts` + `mask_rle_offsets` (uncompressed COCO RLE of all E*P masks, SAM's output format), or
        `mask_bits` [E,P,wpm] (already packed)."""
        s = self.shape
        if "mask_rle_counts" in batch:
            if self._rle_ws is None:
                self._rle_ws = torch.empty(int(ops.lib.marsb200_rle_workspace_bytes(self.E * s.P, s.H, s.W)),
                                           device=self.device, dtype=torch.uint8)
            ops.rle_decode(batch["mask_rle_counts"], batch["mask_rle_offsets"], s.H, s.W,
                           out=self.bits.view(self.E * s.P, -1), check_status=False, workspace=self._rle_ws)
        elif "mask_bits" in batch:
            self.bits.copy_(batch["mask_bits"].view(self.bits.dtype).reshape(self.bits.shape))
        else:
            ops.pack_masks(batch["masks"], out=self.bits)

    def _pairwise(self, lo: int = 0, hi: Optional[int] = None):
        """Intersections of the episodes [lo, hi) and, from them, the suppression relation the ranking kernel scans."""
        hi = self.E if hi is None else hi
        ops.pairwise_inter(self.bits[lo:hi], backend=self.cfg.pair_backend, out=self.inter[lo:hi])
        self._relation(lo, hi)

    def _relation(self, lo: int = 0, hi: Optional[int] = None):
        if self.nms_bits is not None:
            hi = self.E if hi is None else hi
            ops.nms_bitmask(self.inter[lo:hi], self.cfg.nms_iou_threshold, out=self.nms_bits[lo:hi])

    def _mask_chain(self, batch: dict):
        """Ingest side: pack -> pooled bitmaps / areas -> pairwise intersections (depends on the masks only)."""
        s, cfg = self.shape, self.cfg
        self._pool_ready = False
        if self.inter is not None and cfg.fused_ingest and "masks" in batch:
            # one pass over the masks: packed bits + intersections (falls back to two kernels when not fusable)
            ops.pack_pairwise(batch["masks"], backend=cfg.pair_backend, out=(self.bits, self.inter))
            self._relation()
            ops.pool_packed(self.bits, s.H, s.W, s.g, out=self.pool_out)
            return
        if "masks" in batch and cfg.fused_pool:
            # dense masks: packed bits and pooled bitmaps out of ONE pass over the masks (ops.pack_pool).  Measured slower
            # than the two kernels on B200 (the pooling epilogue costs the ingest kernel its load rate), so off
            # by default
            ops.pack_pool(batch["masks"], s.g, out_bits=self.bits, out_pool=self.pool_out)
            if self.inter is not None:
                self._pairwise()
            return
        if "masks" in batch and self.inter is not None and self._pair is not None:
            # small batch: the masks are packed one pixel slice after the other and the intersections of slice k are counted
            # (high-priority stream, tensor cores) while slice k + 1 streams in from HBM; integer sums, any order
            cur = torch.cuda.current_stream()
            per = self.bits.shape[-1] // self._slices
            for k in range(self._slices):
                ops.pack_masks_slice(batch["masks"], k * per, per, out=self.bits)
                self._ev_slice[k].record(cur)
                self._pair.wait_event(self._ev_slice[k])
                with torch.cuda.stream(self._pair):
                    ops.pairwise_inter_slice(self.bits, k * per, per, k > 0, out=self.inter, backend=cfg.pair_backend)
            with torch.cuda.stream(self._pair):
                self._relation()
                self._ev_pair.record(self._pair)
            ops.pool_packed(self.bits, s.H, s.W, s.g, out=self.pool_out)
            self._ev_pool.record(cur)
            cur.wait_event(self._ev_pair)
            self._pool_ready = True
            return
        self._ingest(batch)
        if self.inter is None:
            ops.pool_packed(self.bits, s.H, s.W, s.g, out=self.pool_out)
        elif self._side2 is None:
            ops.pool_packed(self.bits, s.H, s.W, s.g, out=self.pool_out)
            self._pairwise()
        else:
            # pooling (re-reads the bits, HBM/issue-bound) runs beside the tensor-core pairwise kernel
            cur = torch.cuda.current_stream()
            self._ev_pack.record(cur)
            self._side2.wait_event(self._ev_pack)
            with torch.cuda.stream(self._side2):
                ops.pool_packed(self.bits, s.H, s.W, s.g, out=self.pool_out)
                self._ev_pool.record(self._side2)
            self._pairwise()
            cur.wait_event(self._ev_pool)
            self._pool_ready = True  # _ev_pool marks the pooled bitmaps: the region sums need not wait for the intersections

    def _run_partitioned(self, batch: dict, wait: bool = True) -> dict:
        """The same kernel sequence on two disjoint SM sets.  `hbm` partition: pack (+ pooled bitmaps) of one episode
        chunk after the other, at the HBM roofline.  `tensor` partition: normalise -> S -> vva / vta refinement first
        (they do not depend on the masks), then the tensor-core pairwise kernel chunk by chunk as the packed bits
        arrive, then scoring, ranking and merging.  Both sets are busy for the whole step; on one timeline the ingest
        and the tensor kernels add up."""
        s, e, cfg, part = self.shape, self.E, self.cfg, self._part
        n, m = s.N, s.ns * s.N
        main = torch.cuda.current_stream()
        hbm, ten, side = part.hbm_stream, part.tensor_stream, self._part_side
        streams = (hbm, ten, side) + ((part.hbm_side_stream,) if cfg.partition_pool_side_stream else ())
        if self._pending:  # this buffer set is reused: the previous step that ran in it must have drained
            for st in streams:
                st.wait_event(self._ev_join)
        self._ev_fork.record(main)
        for st in streams:
            st.wait_event(self._ev_fork)
        masks = batch["masks"]
        prep = cfg.partition_prep_on_hbm
        with torch.cuda.stream(hbm):
            if prep:
                # memory-streaming work that depends on the inputs only: it belongs with the ingest, not with the
                # tensor-core chain (which then starts with the contractions)
                ops.normalize_rows(batch["feat_s"].reshape(e, m, s.C), True, out=self.fs)
                ops.normalize_rows(batch["feat_q"].reshape(e, n, s.C), True, out=self.fq)
                ops.pir_refine(None, batch["attn_vva"], s.g, cfg.vva_box_threshold, workspace=self.pir_ws,
                               stages=ops.PIR_NORMALISE)
                ops.pir_refine(None, batch["attn_vta"], s.gt, cfg.vta_box_threshold, workspace=self.pir_ws_vta,
                               stages=ops.PIR_NORMALISE)
                self._ev_prep.record(hbm)
            for (lo, hi), ev in zip(self._chunks, self._ev_chunk):
                if cfg.fused_pool and not cfg.partition_pool_on_tensor:
                    ops.pack_pool(masks[lo:hi], s.g, out_bits=self.bits[lo:hi], out_pool=tuple(t[lo:hi] for t in self.pool_out))
                    ev.record(hbm)
                    continue
                ops.pack_masks(masks[lo:hi], out=self.bits[lo:hi])
                ev.record(hbm)
                if cfg.partition_pool_side_stream and not cfg.partition_pool_on_tensor:
                    # pooling is issue-bound, the ingest load-path-bound: the pooled bitmaps of this chunk are computed
                    # on a second stream of the same partition while the next chunk is being packed
                    hs = part.hbm_side_stream
                    hs.wait_event(ev)
                    with torch.cuda.stream(hs):
                        ops.pool_packed(self.bits[lo:hi], s.H, s.W, s.g, out=tuple(t[lo:hi] for t in self.pool_out))
                elif not cfg.partition_pool_on_tensor:
                    ops.pool_packed(self.bits[lo:hi], s.H, s.W, s.g, out=tuple(t[lo:hi] for t in self.pool_out))
            # the last chunks' intersections stay on this partition: their bits only exist when the ingest is over and
            # the tensor partition still has its own queue to drain
            tail = self._chunks[len(self._chunks) - cfg.partition_pairwise_tail:] if cfg.partition_pairwise_tail else []
            if self.inter is not None:
                for lo, hi in tail:
                    self._pairwise(lo, hi)
            if cfg.partition_pool_side_stream and not cfg.partition_pool_on_tensor:
                self._ev_pack.record(part.hbm_side_stream)
                hbm.wait_event(self._ev_pack)  # the pooled bitmaps of every chunk are done
            self._ev_pool.record(hbm)
        # the vta refinement is independent of everything else: it fills the `hbm` partition once the ingest is done
        # (the tensor partition is the longer chain) or runs beside the vva chain inside the tensor partition
        vta_stream = hbm if cfg.partition_vta_on_hbm else side
        vta_stages = (ops.PIR_CONTRACT | ops.PIR_APPLY) if prep else ops.PIR_ALL
        with torch.cuda.stream(vta_stream):
            if prep:
                vta_stream.wait_event(self._ev_prep)
            ops.pir_refine(batch["vta_raw"], batch["attn_vta"], s.gt, cfg.vta_box_threshold, apply_minmax=False,
                           backend=cfg.gemm_backend, workspace=self.pir_ws_vta, out=self.vta_ref, stages=vta_stages)
            self._ev_vta.record(vta_stream)
        with torch.cuda.stream(ten):
            if prep:
                ten.wait_event(self._ev_prep)
            else:
                ops.normalize_rows(batch["feat_s"].reshape(e, m, s.C), True, out=self.fs)
                ops.normalize_rows(batch["feat_q"].reshape(e, n, s.C), True, out=self.fq)
            ops.pool_mask(batch["support_mask"], s.g, out=self.row_fg)
            ops.sim_contract(self.fs, self.fq, m, n, s.C, want_sim=cfg.want_sim,
                             want_cost=cfg.want_cost or cfg.emd_on_device, row_fg=self.row_fg, backend=cfg.gemm_backend,
                             out=self.gemm_out)
            ops.vva_finalize(self.gemm_out["colstats"], self.row_fg, m, n, out=self.prior)
            ops.pir_refine(self.prior, batch["attn_vva"], s.g, cfg.vva_box_threshold, apply_minmax=True,
                           backend=cfg.gemm_backend, workspace=self.pir_ws, out=self.vva,
                           stages=(ops.PIR_CONTRACT | ops.PIR_APPLY) if prep else ops.PIR_ALL)
            ops.clip_scores(batch["clip_img"], batch["clip_txt"], out=self.clip)
            n_front = len(self._chunks) - len(tail)
            for k, ((lo, hi), ev) in enumerate(zip(self._chunks, self._ev_chunk)):
                on_tensor = self.inter is not None and k < n_front
                if on_tensor or cfg.partition_pool_on_tensor:
                    ten.wait_event(ev)
                if cfg.partition_pool_on_tensor:
                    ops.pool_packed(self.bits[lo:hi], s.H, s.W, s.g, out=tuple(t[lo:hi] for t in self.pool_out))
                if on_tensor:
                    self._pairwise(lo, hi)
            ten.wait_event(self._ev_vta)
            ops.resize_minmax(self.vta_ref.reshape(e, s.gt, s.gt), s.g, True, out=self.vta)
            ten.wait_event(self._ev_pool)
            ops.region_sums(self.pool_out[0], self.vva, self.vta, out=self.region_out)
            emd = batch.get("emd")
            if cfg.emd_on_device:
                emd = ops.emd_scores(self.gemm_out["cost"], self.row_fg.reshape(e, m), self.pool_out[0],
                                     t_cap=self.emd_t_cap, m_cap=self.emd_m_cap, workspace=self.emd_ws, out=self.emd_out,
                                     check=False, status=self.emd_status)
            ops.fuse_rank(emd, self.clip, self.pool_out[2], self.region_out[0], self.region_out[1],
                          self.region_out[2], self.inter, cfg.alpha, cfg.static_threshold, cfg.dynamic_threshold,
                          cfg.nms_iou_threshold, out=self.rank_out, record=self.record_buf, nms_bits=self.nms_bits)
            ops.merge_masks(self.bits, self.rank_out["flags"], s.H * s.W, want_bits=True,
                            want_f32=cfg.want_merged_f32, out=self.merge_out)
            self._ev_join.record(ten)
        self._pending = True
        if wait:
            self.join()
        return self.outputs()

    def close(self) -> None:
        """Releases the SM partition (green contexts and their streams) of an engine that owns one."""
        if self._part is not None:
            self._part.close()
            self._part = None

    def join(self) -> dict:
        """Orders the current stream after the step enqueued by `run(..., wait=False)`; returns its outputs."""
        if self._part is not None and self._pending:
            torch.cuda.current_stream().wait_event(self._ev_join)
        return self.outputs()

    def run(self, batch: dict, wait: bool = True) -> dict:
        """Enqueues one step.  With `wait=False` (partitioned schedule only) the current stream is NOT ordered after the
        step - `join()` does that - so the next step's ingest can start while this one's tail is still running."""
        s, e, cfg = self.shape, self.E, self.cfg
        n, m = s.N, s.ns * s.N
        if self._part is not None and "masks" in batch and not cfg.fused_ingest and not self._capturing:
            return self._run_partitioned(batch, wait)
        # the cap protects the dense-mask ingest; packed / RLE proposals have no such read and keep the whole device
        capped = ([st for st in (self._hi, self._side3, self._side4) if st is not None]
                  if self._contraction_sms and "masks" in batch else [])
        for st in capped:  # host-side launch parameter: read when the contractions are enqueued below
            set_stream_sm_cap(st, self._contraction_sms)
        try:
            return self._run_one_timeline(batch)
        finally:
            for st in capped:  # torch hands out pooled stream handles: never leave a cap on one
                set_stream_sm_cap(st, 0)

    def _run_one_timeline(self, batch: dict) -> dict:
        s, e, cfg = self.shape, self.E, self.cfg
        n, m = s.N, s.ns * s.N
        main = torch.cuda.current_stream()
        if self._side is not None:
            # the mask chain is HBM-bound and the alignment chain tensor/L2-bound: let them share the SMs
            self._ev_fork.record(main)
            self._side.wait_event(self._ev_fork)
            with torch.cuda.stream(self._side):
                self._mask_chain(batch)
                self._ev_join.record(self._side)
        if self._side3 is not None:
            # the vta refinement depends on nothing of the vva chain: its small kernels (box mask, column sums, mat-vecs)
            # fill the gaps of the other chain's contractions
            # ... and so do the pooled support mask and the AlphaCLIP scores: they ride on this (shorter) chain instead of
            # lengthening the vva chain, which is on the critical path of a small batch
            self._side3.wait_event(self._ev_fork)
            with torch.cuda.stream(self._side3):
                ops.pool_mask(batch["support_mask"], s.g, out=self.row_fg)
                self._ev_rowfg.record(self._side3)
                ops.pir_refine(batch["vta_raw"], batch["attn_vta"], s.gt, cfg.vta_box_threshold, apply_minmax=False,
                               backend=cfg.gemm_backend, workspace=self.pir_ws_vta, out=self.vta_ref)
                ops.clip_scores(batch["clip_img"], batch["clip_txt"], out=self.clip)
                self._ev_vta.record(self._side3)
        import contextlib

        if self._side4 is not None:
            self._side4.wait_event(self._ev_fork)
            with torch.cuda.stream(self._side4):
                ops.pir_refine(None, batch["attn_vva"], s.g, cfg.vva_box_threshold, backend=cfg.gemm_backend,
                               workspace=self.pir_ws, stages=ops.PIR_NORMALISE | ops.PIR_CONTRACT)
                self._ev_vva_g.record(self._side4)
        if self._hi is not None:
            self._hi.wait_event(self._ev_fork)
        with (torch.cuda.stream(self._hi) if self._hi is not None else contextlib.nullcontext()):
            chain = torch.cuda.current_stream()
            ops.normalize_rows(batch["feat_s"].reshape(e, m, s.C), True, out=self.fs)
            ops.normalize_rows(batch["feat_q"].reshape(e, n, s.C), True, out=self.fq)
            if self._side3 is not None:
                chain.wait_event(self._ev_rowfg)
            else:
                ops.pool_mask(batch["support_mask"], s.g, out=self.row_fg)
            ops.sim_contract(self.fs, self.fq, m, n, s.C, want_sim=cfg.want_sim, want_cost=cfg.want_cost or cfg.emd_on_device,
                             row_fg=self.row_fg, backend=cfg.gemm_backend, out=self.gemm_out)
            ops.vva_finalize(self.gemm_out["colstats"], self.row_fg, m, n, out=self.prior)
            if self._side4 is not None:
                chain.wait_event(self._ev_vva_g)
            ops.pir_refine(self.prior, batch["attn_vva"], s.g, cfg.vva_box_threshold, apply_minmax=True,
                           backend=cfg.gemm_backend, workspace=self.pir_ws, out=self.vva,
                           stages=ops.PIR_APPLY if self._side4 is not None else ops.PIR_ALL)
            if self._side3 is not None:
                chain.wait_event(self._ev_vta)
            else:
                ops.pir_refine(batch["vta_raw"], batch["attn_vta"], s.gt, cfg.vta_box_threshold, apply_minmax=False,
                               backend=cfg.gemm_backend, workspace=self.pir_ws_vta, out=self.vta_ref)
            ops.resize_minmax(self.vta_ref.reshape(e, s.gt, s.gt), s.g, True, out=self.vta)
            if self._side3 is None:
                ops.clip_scores(batch["clip_img"], batch["clip_txt"], out=self.clip)
            if self._hi is not None:
                self._ev_hi.record(self._hi)
        if self._hi is not None:
            main.wait_event(self._ev_hi)
        if self._side is not None and self._pool_ready:
            # the region sums only need the pooled bitmaps; the intersections and the suppression relation are first
            # read by the ranking kernel
            main.wait_event(self._ev_pool)
            ops.region_sums(self.pool_out[0], self.vva, self.vta, out=self.region_out)
            main.wait_event(self._ev_join)
        else:
            if self._side is not None:
                main.wait_event(self._ev_join)
            else:
                self._mask_chain(batch)
            ops.region_sums(self.pool_out[0], self.vva, self.vta, out=self.region_out)
        emd = batch.get("emd")
        if cfg.emd_on_device:
            emd = ops.emd_scores(self.gemm_out["cost"], self.row_fg.reshape(e, m), self.pool_out[0], t_cap=self.emd_t_cap,
                                 m_cap=self.emd_m_cap, workspace=self.emd_ws, out=self.emd_out, check=False,
                                 status=self.emd_status)
        ops.fuse_rank(emd, self.clip, self.pool_out[2], self.region_out[0], self.region_out[1],
                      self.region_out[2], self.inter, cfg.alpha, cfg.static_threshold, cfg.dynamic_threshold,
                      cfg.nms_iou_threshold, out=self.rank_out, record=self.record_buf, nms_bits=self.nms_bits)
        ops.merge_masks(self.bits, self.rank_out["flags"], s.H * s.W, want_bits=True,
                        want_f32=cfg.want_merged_f32, out=self.merge_out)
        return self.outputs()

    def outputs(self) -> dict:
        out = dict(row_fg=self.row_fg, prior=self.prior, vva=self.vva, vta=self.vta, bits=self.bits,
                   pooled=self.pool_out[0], area=self.pool_out[1], pooled_count=self.pool_out[2],
                   sum_vva=self.region_out[0], sum_vta=self.region_out[1], union_count=self.region_out[2],
                   inter=self.inter, clip=self.clip, merged_bits=self.merge_out["bits"],
                   merged=self.merge_out.get("f32"), sim=self.gemm_out.get("sim"), cost=self.gemm_out.get("cost"),
                   emd=self.emd_out if self.cfg.emd_on_device else None,
                   emd_status=self.emd_status if self.cfg.emd_on_device else None)
        out.update(self.rank_out)
        return out

    def check_status(self) -> None:
        """Host-side check of the last step (one small device->host read, so it is not part of `run`): raises if the EMD
        solver reported a fault or if a fused score came out non-finite (`summary[:, 3]`, set by fuse_rank)."""
        if self.cfg.emd_on_device:
            ops.raise_on_emd_status(int(self.emd_status.item()))
        bad = torch.nonzero(self.rank_out["summary"][:, 3]).reshape(-1)
        if bad.numel():
            raise ops.MarsB200Error(f"non-finite fused scores in episodes {bad.tolist()} of the batch")

    # ------------------------------------------------------------------ CUDA graph replay
    def capture(self, batch: dict):
        """Capture `run` on static copies of `batch`; `replay(batch)` then refreshes the inputs and replays."""
        self._static = {k: v.clone() for k, v in batch.items()}
        self._capturing = True  # a graph is one whole-device timeline: green-context streams are not captured
        stream = torch.cuda.Stream(device=self.device)
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            self.run(self._static)  # warm-up outside the capture (function attributes, lazy module load)
        torch.cuda.current_stream().wait_stream(stream)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.run(self._static)
        self._capturing = False
        return self

    def replay(self, batch: Optional[dict] = None) -> dict:
        if batch is not None:
            for k, v in batch.items():
                self._static[k].copy_(v, non_blocking=True)
        self._graph.replay()
        return self.outputs()

    # ------------------------------------------------------------------ result records
    def record_bytes(self) -> int:
        return ops.record_bytes(self.shape.P)

    def records(self) -> torch.Tensor:
        """Fixed-size per-episode result records [E, record_bytes] uint8 (order int32[P] | score float32[P] | flags
        uint8[pad4(P)] | summary int32[4]), written by the fuse / rank kernel of the last step: no copy, no concatenation.
        The tensor is the engine's own buffer (or its slice of the gather table) - clone it to keep it across steps."""
        return self.record_buf

    def attach_gather_table(self, world_size: int, rank: int) -> torch.Tensor:
        """Allocates the all-gather table [world_size * E, record_bytes] and makes this rank's slice the buffer the fuse /
        rank kernel writes the records into; `gather()` then runs the collective in place."""
        if self._graph is not None:
            raise RuntimeError("attach_gather_table must precede capture(): the graph holds the record pointer")
        e = self.E
        self._gather_table = torch.empty((world_size * e, self.record_bytes()), device=self.device, dtype=torch.uint8)
        self.record_buf = self._gather_table[rank * e:(rank + 1) * e]
        return self._gather_table

    def gather(self, group=None) -> torch.Tensor:
        """The path's only collective: all-gather of the result records, in place in the table this engine's kernels wrote
        their slice of (NCCL over NVLink on GPUs)."""
        import torch.distributed as dist

        if self._gather_table is None:
            raise RuntimeError("call attach_gather_table(world_size, rank) first")
        dist.all_gather_into_tensor(self._gather_table, self.record_buf, group=group)
        return self._gather_table


class PipelinedRanking:
    """Software pipeline over steps: `depth` buffer sets (RankingEngines) share one SM partition and take the steps in
    turn, so the ingest of step i + 1 runs on the `hbm` partition while the tensor partition is still scoring step i.
    `submit(batch)` enqueues a step and returns its ticket; `result(ticket)` orders the current stream after that step
    and returns its outputs (valid until the same buffer set is submitted again, `depth` steps later)."""

    def __init__(self, shape: EpisodeShape, episodes_per_batch: int, cfg: RankingConfig, device, mask_dtype=torch.float32,
                 depth: int = 2):
        if not cfg.tensor_partition_sms:
            raise ValueError("PipelinedRanking needs cfg.tensor_partition_sms (the two-partition schedule)")
        first = RankingEngine(shape, episodes_per_batch, cfg, device, mask_dtype)
        self.engines = [first] + [RankingEngine(shape, episodes_per_batch, cfg, device, mask_dtype, partition=first._part)
                                  for _ in range(depth - 1)]
        self._next = 0

    def submit(self, batch: dict) -> int:
        ticket = self._next
        self.engines[ticket % len(self.engines)].run(batch, wait=False)
        self._next += 1
        return ticket

    def engine(self, ticket: int) -> "RankingEngine":
        return self.engines[ticket % len(self.engines)]

    def result(self, ticket: int) -> dict:
        return self.engine(ticket).join()

    def close(self):
        self.engines[0]._part.close()


class InterleavedRanking:
    """Two (or more) one-timeline engines on their own streams taking the steps in turn: the low-occupancy tail of step i
    (one-CTA-per-episode kernels: box mask, fuse / rank / NMS, selection) overlaps the contractions and the ingest of
    step i + 1.  The schedule for proposal formats whose ingest is cheap (uint8 masks, packed bits, RLE), where no SM
    partition pays; same interface as `PipelinedRanking`."""

    def __init__(self, shape: EpisodeShape, episodes_per_batch: int, cfg: RankingConfig, device, mask_dtype=torch.float32,
                 depth: int = 2):
        if cfg.tensor_partition_sms:
            raise ValueError("InterleavedRanking runs whole-device engines (use PipelinedRanking with SM partitions)")
        self.engines = [RankingEngine(shape, episodes_per_batch, cfg, device, mask_dtype) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(device=device) for _ in range(depth)]
        self._fork = [torch.cuda.Event() for _ in range(depth)]
        self._done = [torch.cuda.Event() for _ in range(depth)]
        self._next = 0

    def submit(self, batch: dict) -> int:
        ticket = self._next
        k = ticket % len(self.engines)
        main = torch.cuda.current_stream()
        self._fork[k].record(main)           # the batch (and whatever produced it) is ready on the caller's stream
        self.streams[k].wait_event(self._fork[k])
        with torch.cuda.stream(self.streams[k]):
            self.engines[k].run(batch)
            self._done[k].record(self.streams[k])
        self._next += 1
        return ticket

    def engine(self, ticket: int) -> "RankingEngine":
        return self.engines[ticket % len(self.engines)]

    def result(self, ticket: int) -> dict:
        torch.cuda.current_stream().wait_event(self._done[ticket % len(self.engines)])
        return self.engine(ticket).outputs()

    def close(self):
        pass


def decode_records(records: torch.Tensor, p: int) -> dict:
    """Splits result records [n, record_bytes] (CPU or CUDA) into order / scores / flags / summary."""
    n = records.shape[0]
    fl = (p + 3) // 4 * 4
    order = records[:, :4 * p].contiguous().view(torch.int32).reshape(n, p)
    scores = records[:, 4 * p:8 * p].contiguous().view(torch.float32).reshape(n, p)
    flags = records[:, 8 * p:9 * p]
    summary = records[:, 8 * p + fl:8 * p + fl + 16].contiguous().view(torch.int32).reshape(n, 4)
    return dict(order=order, scores=scores, flags=flags, summary=summary)


def shard_range(num_episodes: int, rank: int, world_size: int):
    """Contiguous block of episode ids owned by `rank`."""
    per = (num_episodes + world_size - 1) // world_size
    lo = min(rank * per, num_episodes)
    return lo, min(lo + per, num_episodes)


def gather_records(local: torch.Tensor, num_episodes: int, group=None) -> torch.Tensor:
    """All-gather per-rank record blocks into the global [num_episodes, record_bytes] table.

    The only collective of the path: one all_gather over NCCL (NVLink/NVSwitch) on GPUs, gloo in the
    CPU tests.  Ranks own contiguous, equally sized blocks (the last one may be padded).
    """
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = (num_episodes + world - 1) // world
    if local.shape[0] < per:
        pad = torch.zeros((per - local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * per, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:num_episodes]
