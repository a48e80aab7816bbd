"""Thin torch wrappers over the C ABI: one function per entry point of include/marsb200.h.

PyTorch is used only for device memory and streams.  Every wrapper takes CUDA
tensors, allocates its outputs with torch (so the caching allocator owns
them), and enqueues on the current torch stream without synchronising.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import GEMM_SIMT, GEMM_TCGEN05, MASK_F32, MASK_U8, PAIR_AUTO, PAIR_FP4, PAIR_MMA, PAIR_POPC, check, lib

MarsB200Error = _lib.MarsB200Error

__all__ = [
    "words_per_mask", "pad_rows", "pad_k", "normalize_rows", "pool_mask", "sim_contract", "match_argmax", "mutual_matches", "lsap", "vva_finalize",
    "attn_mean", "pir_refine", "resize_minmax", "pack_masks", "pack_pairwise", "pool_packed", "region_sums", "pairwise_inter",
    "emd_scores", "clip_scores", "fuse_rank", "merge_masks", "points_in_masks", "matcher_scores", "eval_areas", "eval_accumulate", "eval_iou", "rle_decode", "mask_boxes", "stability_score", "box_nms", "masked_feature_means", "masked_sim_stats", "masked_row_mean",
    "GEMM_TCGEN05", "GEMM_SIMT", "PAIR_POPC", "PAIR_MMA", "PAIR_FP4", "PAIR_AUTO",
]

# default back ends (module-level so tests can pin either one)
# (MARSB200_GEMM=simt / MARSB200_PAIR=popc|mma select the validation kernels, for debugging only)
DEFAULT_GEMM = GEMM_SIMT if os.environ.get("MARSB200_GEMM", "") == "simt" else GEMM_TCGEN05
DEFAULT_PAIR = {"popc": PAIR_POPC, "mma": PAIR_MMA, "fp4": PAIR_FP4}.get(os.environ.get("MARSB200_PAIR", ""), PAIR_AUTO)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _cuda(t: torch.Tensor, dtype=None, name="tensor") -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.MarsB200Error(f"{name} must be a CUDA tensor (marsb200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _mask_tensor(m: torch.Tensor):
    """Masks as the ingest kernels read them: float32 as is, bool/uint8 as bytes, anything else -> float32."""
    if not m.is_cuda:
        raise _lib.MarsB200Error("masks must be CUDA tensors (marsb200 has no CPU path)")
    if m.dtype == torch.float32:
        return m.contiguous(), MASK_F32
    if m.dtype == torch.bool:
        return m.contiguous().view(torch.uint8), MASK_U8
    if m.dtype == torch.uint8:
        return m.contiguous(), MASK_U8
    return m.float().contiguous(), MASK_F32


def words_per_mask(hw: int) -> int:
    return int(lib.marsb200_words_per_mask(hw))


def pad_rows(rows: int) -> int:
    return int(lib.marsb200_pad_rows(rows))


def pad_k(k: int) -> int:
    return int(lib.marsb200_pad_k(k))


# ----------------------------------------------------------------------------- A1
def normalize_rows(x: torch.Tensor, normalize: bool = True, out=None):
    """x [E, rows, k] -> (xn, lo), both [E, pad_rows, pad_k] fp32: the (normalised) rows in the contraction's
    zero-padded layout and their tf32 residuals (xn minus its top 19 bits), the operand pair of `sim_contract`."""
    x = _cuda(x, torch.float32, "x")
    if x.dim() == 2:
        x = x[None]
    e, rows, k = x.shape
    if out is None:
        xn = torch.empty((e, pad_rows(rows), pad_k(k)), device=x.device, dtype=torch.float32)
        lo = torch.empty_like(xn)
    else:
        xn, lo = out
    check(lib.marsb200_normalize_rows(x.data_ptr(), k, e, rows, k, int(normalize), xn.data_ptr(), lo.data_ptr(), _stream()))
    return xn, lo


# ----------------------------------------------------------------------------- A3 / A6
def pool_mask(masks: torch.Tensor, g: int, out=None) -> torch.Tensor:
    """masks [..., H, W] -> uint8 [..., g*g] (adaptive max pool > 0)."""
    m, dt = _mask_tensor(masks)
    h, w = m.shape[-2:]
    n = m.numel() // (h * w)
    if out is None:
        out = torch.empty(tuple(m.shape[:-2]) + (g * g,), device=m.device, dtype=torch.uint8)
    check(lib.marsb200_pool_mask(m.data_ptr(), dt, n, h, w, g, out.data_ptr(), _stream()))
    return out


# ----------------------------------------------------------------------------- A2 + A3
def sim_contract(a, b, m: int, n: int, k: int, want_sim=True, want_cost=False, row_fg=None, backend=None, out=None):
    """S = A B^T on operands from normalize_rows.  Returns dict(sim, cost, colstats)."""
    (a, a_lo), (b, b_lo) = a, b
    e = a.shape[0]
    dev = a.device
    out = out or {}
    sim = out.get("sim") if want_sim else None
    cost = out.get("cost") if want_cost else None
    if want_sim and sim is None:
        sim = torch.empty((e, m, n), device=dev, dtype=torch.float32)
    if want_cost and cost is None:
        cost = torch.empty((e, m, n), device=dev, dtype=torch.float32)
    colstats = None
    if row_fg is not None:
        row_fg = _cuda(row_fg, torch.uint8, "row_fg").reshape(e, m)
        colstats = out.get("colstats")
        if colstats is None:
            colstats = torch.empty((e, pad_rows(m) // 128, 4, n), device=dev, dtype=torch.float32)
    check(lib.marsb200_sim_contract(a.data_ptr(), a_lo.data_ptr(), b.data_ptr(), b_lo.data_ptr(), e, m, n, k,
                                    _ptr(sim), _ptr(cost), _ptr(row_fg), _ptr(colstats),
                                    DEFAULT_GEMM if backend is None else backend, _stream()))
    return dict(sim=sim, cost=cost, colstats=colstats)


def match_argmax(sim: torch.Tensor, k: int = 1, row_mask: Optional[torch.Tensor] = None, rows=True, cols=True):
    """sim [E, M, N] -> dict(row_vals/row_idx [E, M, k], col_vals/col_idx [E, N]); ties -> lowest index."""
    sim = _cuda(sim, torch.float32, "sim")
    if sim.dim() == 2:
        sim = sim[None]
    e, m, n = sim.shape
    dev = sim.device
    rv = torch.empty((e, m, k), device=dev, dtype=torch.float32) if rows else None
    ri = torch.empty((e, m, k), device=dev, dtype=torch.int32) if rows else None
    cv = torch.empty((e, n), device=dev, dtype=torch.float32) if cols else None
    ci = torch.empty((e, n), device=dev, dtype=torch.int32) if cols else None
    mk = None if row_mask is None else _cuda(row_mask, torch.uint8, "row_mask").reshape(e, m)
    check(lib.marsb200_match_argmax(sim.data_ptr(), _ptr(mk), e, m, n, k, _ptr(rv), _ptr(ri), _ptr(cv), _ptr(ci), _stream()))
    return dict(row_vals=rv, row_idx=ri, col_vals=cv, col_idx=ci)


def mutual_matches(sim: torch.Tensor, row_fg: torch.Tensor):
    """Bidirectional arg-max matching for one episode: fg support rows whose best query patch points back into the mask.

    Forward: every fg support row takes its best query patch.  Reverse: that query patch takes its best support row
    over ALL rows; the pair is kept when the reverse match is a fg row (the arg-max analogue of the retain rule at
    matcher/Matcher.py:468-477).  Returns (support_rows, query_patches, similarity) of the kept pairs.
    """
    sim = sim.reshape(1, *sim.shape[-2:])
    fg = _cuda(row_fg, torch.uint8, "row_fg").reshape(1, -1)
    fwd = match_argmax(sim, k=1, cols=False)
    rev = match_argmax(sim, rows=False)
    rows = torch.nonzero(fg[0]).flatten()
    q = fwd["row_idx"][0, rows, 0].long()
    back = rev["col_idx"][0, q].long()
    keep = fg[0, back] != 0
    return rows[keep], q[keep], fwd["row_vals"][0, rows, 0][keep]


def lsap(sim: torch.Tensor, row_sel: Optional[torch.Tensor] = None, col_sel: Optional[torch.Tensor] = None,
         maximize: bool = True, check_status: bool = True, status: Optional[torch.Tensor] = None, out=None):
    """Exact assignment on sim [E, R, C] restricted to the selected rows / columns.

    Returns (row_to_col [E, R] int32 with -1 for unassigned rows, objective [E] float64).  With a caller-owned `status`
    ([1] int32, zeroed) nothing is read back here - the caller checks it (`raise_on_lsap_status`) once it has joined the
    stream, so the launch does not synchronise.  `out` = (row_to_col, objective) buffers to fill instead of new ones.
    """
    sim = _cuda(sim, torch.float32, "sim")
    if sim.dim() == 2:
        sim = sim[None]
    e, r, c = sim.shape
    rs = None if row_sel is None else _cuda(row_sel, torch.uint8, "row_sel").reshape(e, r)
    cs = None if col_sel is None else _cuda(col_sel, torch.uint8, "col_sel").reshape(e, c)
    if out is None:
        r2c = torch.empty((e, r), device=sim.device, dtype=torch.int32)
        obj = torch.empty((e,), device=sim.device, dtype=torch.float64)
    else:
        r2c, obj = out
    own_status = status is None
    if own_status:
        status = torch.zeros(1, device=sim.device, dtype=torch.int32)
    # the solver's state is sized for the selected rows / columns (one small sync, like the reference's host call)
    nr = r if rs is None else max(1, int((rs != 0).sum(dim=1).max().item()))
    nc = c if cs is None else max(1, int((cs != 0).sum(dim=1).max().item()))
    check(lib.marsb200_lsap(sim.data_ptr(), _ptr(rs), _ptr(cs), e, r, c, int(maximize), min(nr, nc), max(nr, nc),
                            r2c.data_ptr(), obj.data_ptr(), status.data_ptr(), _stream()))
    if check_status and own_status:
        raise_on_lsap_status(status)
    return r2c, obj


def raise_on_lsap_status(status: torch.Tensor) -> None:
    code = int(status.item())
    if code:
        raise _lib.MarsB200Error(f"lsap: a problem of size {code} exceeds the shared-memory state")


def vva_finalize(colstats: torch.Tensor, row_fg: torch.Tensor, m: int, n: int, out=None) -> torch.Tensor:
    e = colstats.shape[0]
    row_fg = _cuda(row_fg, torch.uint8, "row_fg").reshape(e, m)
    if out is None:
        out = torch.empty((e, n), device=colstats.device, dtype=torch.float32)
    check(lib.marsb200_vva_finalize(colstats.data_ptr(), row_fg.data_ptr(), e, m, n, out.data_ptr(), _stream()))
    return out


# ----------------------------------------------------------------------------- A4
def attn_mean(maps: Sequence[torch.Tensor], skip: int, out=None) -> torch.Tensor:
    """Mean over layers and heads of [heads, T, T] (or [1, heads, T, T]) maps -> [N, N] with N = T - skip."""
    prepared = []
    for a in maps:
        if a.dim() == 4:
            a = a[0]
        if a.dtype not in (torch.float32, torch.float16):
            a = a.float()
        prepared.append(_cuda(a, None, "attention map"))
    dt = prepared[0].dtype
    if any(p.dtype != dt or p.shape != prepared[0].shape for p in prepared):
        raise _lib.MarsB200Error("attention maps must share dtype and shape")
    heads, t, _ = prepared[0].shape
    n = t - skip
    if out is None:
        out = torch.empty((n, n), device=prepared[0].device, dtype=torch.float32)
    arr = (ctypes.c_void_p * len(prepared))(*[p.data_ptr() for p in prepared])
    check(lib.marsb200_attn_mean(arr, len(prepared), 0 if dt == torch.float32 else 1, heads, t, skip,
                                 out.data_ptr(), out.stride(0), _stream()))
    return out


def pir_workspace(e: int, n: int, device) -> torch.Tensor:
    nbytes = int(lib.marsb200_pir_workspace_bytes(e, n))
    return torch.empty((nbytes + 255) // 256 * 256, device=device, dtype=torch.uint8)


PIR_NORMALISE, PIR_CONTRACT, PIR_APPLY, PIR_ALL = 1, 2, 4, 7


def pir_refine(prior: Optional[torch.Tensor], attn: Optional[torch.Tensor], g: int, box_threshold: float, apply_minmax=False,
               want_box=False, backend=None, workspace=None, out=None, stages: int = PIR_ALL, episodes: Optional[int] = None):
    """prior [E, g*g], attn [E, N, N] -> refined [E, N] (and the uint8 box mask when asked).

    `stages` selects parts of the refinement (include/marsb200.h): PIR_NORMALISE (attention -> operands in `workspace`,
    needs only `attn`), PIR_CONTRACT (G = D D^T, needs only `workspace`), PIR_APPLY (box mask, mat-vecs, min-max; needs
    `prior`).  A scheduler can run the first two before the prior exists; they must share `workspace` and keep the order."""
    n = g * g
    if prior is not None:
        prior = _cuda(prior, torch.float32, "prior").reshape(-1, n)
        e = prior.shape[0]
    elif attn is not None:
        e = attn.numel() // (attn.shape[-2] * attn.shape[-1])
    else:
        e = int(episodes)
    ld = n
    if attn is not None:
        attn = _cuda(attn, torch.float32, "attn").reshape(e, n, -1)
        ld = attn.stride(1)
    dev = prior.device if prior is not None else (attn.device if attn is not None else workspace.device)
    if workspace is None:
        workspace = pir_workspace(e, n, dev)
    if out is None and (stages & PIR_APPLY):
        out = torch.empty((e, n), device=dev, dtype=torch.float32)
    box = torch.empty((e, n), device=dev, dtype=torch.uint8) if want_box else None
    check(lib.marsb200_pir_stages(_ptr(prior), _ptr(attn), ld, e, g, float(box_threshold),
                                  int(apply_minmax), _ptr(out), _ptr(box), workspace.data_ptr(),
                                  workspace.numel(), DEFAULT_GEMM if backend is None else backend, int(stages), _stream()))
    return (out, box) if want_box else out


def scoremap_boxes(prior: torch.Tensor, g: int, box_threshold: float):
    """prior [E, g*g] -> (boxes int32 [E, g*g, 4] as (x0, y0, x1, y1), count int32 [E]): the per-component boxes behind
    the PIR box mask (PriorInformationRefinementModule.py:91-122)."""
    prior = _cuda(prior, torch.float32, "prior").reshape(-1, g * g)
    e, n = prior.shape
    boxes = torch.zeros((e, n, 4), device=prior.device, dtype=torch.int32)
    count = torch.zeros((e,), device=prior.device, dtype=torch.int32)
    check(lib.marsb200_scoremap_boxes(prior.data_ptr(), e, g, float(box_threshold), boxes.data_ptr(), count.data_ptr(),
                                      _stream()))
    return boxes, count


def resize_minmax(src: torch.Tensor, gd: int, apply_minmax=True, out=None) -> torch.Tensor:
    src = _cuda(src, torch.float32, "src")
    gs = src.shape[-1]
    e = src.numel() // (gs * gs)
    if out is None:
        out = torch.empty((e, gd * gd), device=src.device, dtype=torch.float32)
    check(lib.marsb200_resize_minmax(src.data_ptr(), e, gs, gd, int(apply_minmax), out.data_ptr(), _stream()))
    return out


# ----------------------------------------------------------------------------- A6 / A9
def pack_masks(masks: torch.Tensor, out=None) -> torch.Tensor:
    """masks [..., H, W] -> packed bits [..., words_per_mask(H*W)] (int32 storage of uint32 words)."""
    m, dt = _mask_tensor(masks)
    h, w = m.shape[-2:]
    n = m.numel() // (h * w)
    wpm = words_per_mask(h * w)
    if out is None:
        out = torch.empty(tuple(m.shape[:-2]) + (wpm,), device=m.device, dtype=torch.int32)
    check(lib.marsb200_pack_masks(m.data_ptr(), dt, n, h * w, out.data_ptr(), _stream()))
    return out


def pack_masks_slice(masks: torch.Tensor, word_begin: int, word_count: int, out: torch.Tensor) -> torch.Tensor:
    """Packed words [word_begin, word_begin + word_count) of every mask into `out` [..., wpm] (whole 512-word blocks)."""
    m, dt = _mask_tensor(masks)
    h, w = m.shape[-2:]
    check(lib.marsb200_pack_masks_slice(m.data_ptr(), dt, m.numel() // (h * w), h * w, word_begin, word_count, out.data_ptr(),
                                        _stream()))
    return out


def pairwise_inter_slice(bits: torch.Tensor, word_begin: int, word_count: int, accumulate: bool, out: torch.Tensor,
                         backend=None) -> torch.Tensor:
    """Adds (or, with accumulate=False, writes) the intersections counted over one pixel slice of bits [E, P, wpm]."""
    if bits.dim() == 2:
        bits = bits[None]
    e, p, wpm = bits.shape
    check(lib.marsb200_pairwise_inter_slice(bits.data_ptr(), e, p, wpm, word_begin, word_count, int(bool(accumulate)),
                                            out.data_ptr(), DEFAULT_PAIR if backend is None else backend, _stream()))
    return out


def pool_packed(bits: torch.Tensor, h: int, w: int, g: int, out=None):
    """bits [..., wpm] -> (pooled [..., ceil(g*g/32)] int32, area [...], pooled_count [...])."""
    lead = tuple(bits.shape[:-1])
    n = bits.numel() // bits.shape[-1]
    npw = (g * g + 31) // 32
    if out is None:
        pooled = torch.empty(lead + (npw,), device=bits.device, dtype=torch.int32)
        area = torch.empty(lead, device=bits.device, dtype=torch.int32)
        cnt = torch.empty(lead, device=bits.device, dtype=torch.int32)
    else:
        pooled, area, cnt = out
    check(lib.marsb200_pool_packed(bits.data_ptr(), n, h, w, g, pooled.data_ptr(), area.data_ptr(), cnt.data_ptr(), _stream()))
    return pooled, area, cnt


def pack_pool(masks: torch.Tensor, g: int, out_bits=None, out_pool=None):
    """masks [..., H, W] -> (bits [..., wpm], (pooled, area, pooled_count)) in one pass over the masks: `pack_masks`
    followed by `pool_packed`, with the pooling done on the packed words while they are in registers (one kernel when
    W % 32 == 0, else the two kernels - identical results either way)."""
    m, dt = _mask_tensor(masks)
    h, w = m.shape[-2:]
    lead = tuple(m.shape[:-2])
    n = m.numel() // (h * w)
    wpm = words_per_mask(h * w)
    npw = (g * g + 31) // 32
    if out_bits is None:
        out_bits = torch.empty(lead + (wpm,), device=m.device, dtype=torch.int32)
    if out_pool is None:
        out_pool = (torch.empty(lead + (npw,), device=m.device, dtype=torch.int32),
                    torch.empty(lead, device=m.device, dtype=torch.int32),
                    torch.empty(lead, device=m.device, dtype=torch.int32))
    pooled, area, cnt = out_pool
    check(lib.marsb200_pack_pool_masks(m.data_ptr(), dt, n, h, w, g, out_bits.data_ptr(), pooled.data_ptr(), area.data_ptr(),
                                       cnt.data_ptr(), _stream()))
    return out_bits, (pooled, area, cnt)


def region_sums(pooled: torch.Tensor, vva: torch.Tensor, vta: torch.Tensor, out=None):
    """pooled [E, P, npw]; vva/vta [E, N] -> (sum_vva [E,P], sum_vta [E,P], union_count [E])."""
    e, p, _ = pooled.shape
    n = vva.shape[-1]
    vva = _cuda(vva, torch.float32, "vva").reshape(e, n)
    vta = _cuda(vta, torch.float32, "vta").reshape(e, n)
    if out is None:
        sv = torch.empty((e, p), device=pooled.device, dtype=torch.float32)
        st = torch.empty_like(sv)
        uc = torch.empty((e,), device=pooled.device, dtype=torch.int32)
    else:
        sv, st, uc = out
    check(lib.marsb200_region_sums(pooled.data_ptr(), e, p, n, vva.data_ptr(), vta.data_ptr(), sv.data_ptr(),
                                   st.data_ptr(), uc.data_ptr(), _stream()))
    return sv, st, uc


def pairwise_inter(bits: torch.Tensor, backend=None, out=None) -> torch.Tensor:
    """bits [E, P, wpm] -> int32 intersections [E, P, P] (diagonal = area)."""
    if bits.dim() == 2:
        bits = bits[None]
    e, p, wpm = bits.shape
    if out is None:
        out = torch.empty((e, p, p), device=bits.device, dtype=torch.int32)
    check(lib.marsb200_pairwise_inter(bits.data_ptr(), e, p, wpm, out.data_ptr(),
                                      DEFAULT_PAIR if backend is None else backend, _stream()))
    return out


def pack_pairwise(masks: torch.Tensor, backend=None, out=None):
    """masks [E, P, H, W] -> (bits [E, P, wpm], inter [E, P, P]) in one pass over the masks when possible."""
    m, dt = _mask_tensor(masks)
    if m.dim() == 3:
        m = m[None]
    e, p, h, w = m.shape
    wpm = words_per_mask(h * w)
    if out is None:
        bits = torch.empty((e, p, wpm), device=m.device, dtype=torch.int32)
        inter = torch.empty((e, p, p), device=m.device, dtype=torch.int32)
    else:
        bits, inter = out
    check(lib.marsb200_pack_pairwise(m.data_ptr(), dt, e, p, h * w, bits.data_ptr(), inter.data_ptr(),
                                     DEFAULT_PAIR if backend is None else backend, _stream()))
    return bits, inter


# ----------------------------------------------------------------------------- A7
def emd_scores(cost: torch.Tensor, row_fg: torch.Tensor, pooled: torch.Tensor, t_cap: Optional[int] = None,
               m_cap: Optional[int] = None, pooled_count: Optional[torch.Tensor] = None, workspace=None, out=None,
               check=True, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Exact `1 - emd2` of every proposal: cost [E, m_rows, N] fp32, row_fg [E, m_rows] u8, pooled [E, P, npw].

    `t_cap` (fg support rows) and `m_cap` (pooled patches of a proposal) size the solver's shared-memory fast path; a
    problem beyond them is solved by the global-state launch (slower, any size), so they are performance hints, not
    limits.  When omitted they are read from `row_fg` / `pooled_count` - one small device->host sync, like the
    reference's own host EMD (`m_cap` falls back to N without `pooled_count`).  `status` (int32 [1], optional) receives
    the solver's status word; with `check` it is read back and an exception raised on a fault.
    """
    cost = _cuda(cost, torch.float32, "cost")
    if cost.dim() == 2:
        cost = cost[None]
    e, m_rows, n = cost.shape
    row_fg = _cuda(row_fg, torch.uint8, "row_fg").reshape(e, m_rows)
    pooled = pooled.reshape(e, -1, pooled.shape[-1])
    p = pooled.shape[1]
    if t_cap is None:
        t_cap = max(1, int((row_fg != 0).sum(dim=1).max().item()))
    if m_cap is None:
        m_cap = n if pooled_count is None else max(1, int(pooled_count.max().item()))
    m_cap = min(int(m_cap), n)
    nbytes = int(lib.marsb200_emd_workspace_bytes(e, p, n, m_rows, t_cap, m_cap))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, device=cost.device, dtype=torch.uint8)
    if out is None:
        out = torch.empty((e, p), device=cost.device, dtype=torch.float64)
    if status is None:
        status = torch.zeros(1, device=cost.device, dtype=torch.int32)
    check_rc = lib.marsb200_emd_scores(cost.data_ptr(), row_fg.data_ptr(), pooled.data_ptr(), e, p, m_rows, n, t_cap,
                                       m_cap, workspace.data_ptr(), workspace.numel(), out.data_ptr(),
                                       status.data_ptr(), _stream())
    _lib.check(check_rc)
    if check:
        raise_on_emd_status(int(status.item()))
    return out


def raise_on_emd_status(need: int) -> None:
    """Turns the EMD solver's status word (include/marsb200.h) into an exception."""
    if need < 0:
        raise _lib.MarsB200Error("emd_scores: internal flow-node pool exhausted")
    if need >= (1 << 24):
        raise _lib.MarsB200Error(f"emd_scores: a proposal covers {need - (1 << 24)} patches, beyond the 16-bit index range")
    if need:
        raise _lib.MarsB200Error(f"emd_scores: an episode has {need} foreground support rows, beyond the 16-bit index range")


# ----------------------------------------------------------------------------- A8 / A10 / A11
def clip_scores(img: torch.Tensor, txt: torch.Tensor, out=None) -> torch.Tensor:
    """img [E, P, D] . txt [E, D] -> [E, P] float32.  float16 features (the reference's AlphaCLIP on a GPU) give the
    float16-rounded dot products of a half-precision matmul, stored as float32 - pass `clip_f16=True` to `fuse_rank`."""
    half = img.dtype == torch.float16
    img = _cuda(img, torch.float16 if half else torch.float32, "img")
    if img.dim() == 2:
        img = img[None]
    e, p, d = img.shape
    txt = _cuda(txt, torch.float16 if half else torch.float32, "txt").reshape(e, d)
    if out is None:
        out = torch.empty((e, p), device=img.device, dtype=torch.float32)
    fn = lib.marsb200_clip_scores_f16 if half else lib.marsb200_clip_scores
    check(fn(img.data_ptr(), txt.data_ptr(), e, p, d, out.data_ptr(), _stream()))
    return out


def host_pack_masks(masks: torch.Tensor, out: Optional[torch.Tensor] = None, threads: int = 0) -> torch.Tensor:
    """HOST tensors in, HOST tensor out: masks [..., H, W] float32 / uint8 / bool (CPU, contiguous; pinned or not) -> packed
    bits [..., words_per_mask(H * W)] int32 in the layout of `pack_masks`, by a team of host threads.  The conversion in
    front of the host->device copy (1/32 of the bytes); the only op of this module that takes CPU tensors."""
    if masks.is_cuda:
        raise ValueError("host_pack_masks packs HOST masks; device masks go through pack_masks")
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    if masks.dtype not in (torch.float32, torch.uint8):
        raise TypeError(f"masks must be float32, uint8 or bool, got {masks.dtype}")
    if not masks.is_contiguous():
        masks = masks.contiguous()
    hw = masks.shape[-1] * masks.shape[-2]
    n = masks.numel() // hw
    wpm = words_per_mask(hw)
    shape = tuple(masks.shape[:-2]) + (wpm,)
    if out is None:
        out = torch.empty(shape, dtype=torch.int32)
    if out.is_cuda or out.dtype != torch.int32 or out.numel() != n * wpm or not out.is_contiguous():
        raise ValueError("out must be a contiguous CPU int32 tensor of n * words_per_mask(H * W) elements")
    check(lib.marsb200_host_pack_masks(masks.data_ptr(), MASK_U8 if masks.dtype == torch.uint8 else MASK_F32, n, hw,
                                       out.data_ptr(), int(threads)))
    return out


def nms_bitmask(inter: torch.Tensor, nms_iou_threshold: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pairwise suppression relation [E, P, ceil(P/32)] int32 (bit j of row i: IoU(i, j) > threshold, j != i) from the
    intersections [E, P, P]; independent of the ranking, so it can run as soon as `inter` exists (fuse_rank `nms_bits`)."""
    inter = _cuda(inter, torch.int32, "inter")
    if inter.dim() != 3 or inter.shape[1] != inter.shape[2]:
        raise ValueError("inter must be [E, P, P]")
    e, p, _ = inter.shape
    if out is None:
        out = torch.empty((e, p, (p + 31) // 32), device=inter.device, dtype=torch.int32)
    elif not out.is_cuda or out.dtype != torch.int32 or not out.is_contiguous() or out.numel() != e * p * ((p + 31) // 32):
        raise ValueError("out must be a contiguous CUDA int32 tensor [E, P, ceil(P / 32)]")
    check(lib.marsb200_nms_bitmask(inter.data_ptr(), e, p, float(nms_iou_threshold), out.data_ptr(), _stream()))
    return out


def fuse_rank(emd, clip, pooled_count, sum_vva, sum_vta, union_count, inter, alpha, static_threshold,
              dynamic_threshold, nms_iou_threshold=None, out=None, clip_f16=False, record=None, nms_bits=None):
    """Returns dict(scores [E,P] f64, order [E,P] i32, flags [E,P] u8, summary [E,4] i32).  `clip_f16`: `clip` holds
    float16 values and the fusion follows NumPy's float16 sequence (include/marsb200.h).  `record`: optional uint8
    [E, >= record_bytes(P)] rows (may be a slice of a larger table) the kernel also writes the result records into.
    `nms_bits`: the relation from `nms_bitmask(inter, nms_iou_threshold)` (else it is built inside the ranking kernel)."""
    e, p = clip.shape
    dev = clip.device
    emd = _cuda(emd, torch.float64, "emd").reshape(e, p)
    if out is None:
        out = dict(scores=torch.empty((e, p), device=dev, dtype=torch.float64),
                   order=torch.empty((e, p), device=dev, dtype=torch.int32),
                   flags=torch.empty((e, p), device=dev, dtype=torch.uint8),
                   summary=torch.empty((e, 4), device=dev, dtype=torch.int32))
    use_nms = inter is not None and nms_iou_threshold is not None and nms_iou_threshold >= 0
    check(lib.marsb200_fuse_rank(emd.data_ptr(), clip.data_ptr(), pooled_count.data_ptr(), sum_vva.data_ptr(),
                                 sum_vta.data_ptr(), union_count.data_ptr(), _ptr(inter) if use_nms else None,
                                 _ptr(nms_bits) if use_nms and p <= 1024 else None, e, p,
                                 float(alpha), float(static_threshold), float(dynamic_threshold),
                                 float(nms_iou_threshold) if use_nms else -1.0, int(bool(clip_f16)), out["scores"].data_ptr(),
                                 out["order"].data_ptr(), out["flags"].data_ptr(), out["summary"].data_ptr(),
                                 _ptr(record), 0 if record is None else record.stride(0), _stream()))
    return out


def record_bytes(p: int) -> int:
    return int(lib.marsb200_record_bytes(int(p)))


def merge_masks(bits: torch.Tensor, flags: torch.Tensor, hw: int, want_bits=False, want_f32=True, out=None):
    """OR of the rows whose flag has bit 1 set.  Returns (merged_bits [E,wpm] | None, merged_f32 [E,HW] | None)."""
    e, p, wpm = bits.shape
    out = out or {}
    mb = out.get("bits")
    mf = out.get("f32")
    if want_bits and mb is None:
        mb = torch.empty((e, wpm), device=bits.device, dtype=torch.int32)
    if want_f32 and mf is None:
        mf = torch.empty((e, hw), device=bits.device, dtype=torch.float32)
    check(lib.marsb200_merge_masks(bits.data_ptr(), flags.data_ptr(), e, p, hw, _ptr(mb) if want_bits else None,
                                   _ptr(mf) if want_f32 else None, _stream()))
    return (mb if want_bits else None), (mf if want_f32 else None)


# ----------------------------------------------------------------------------- A12 / 8f-3
def points_in_masks(bits: torch.Tensor, h: int, w: int, points: torch.Tensor) -> torch.Tensor:
    n = bits.numel() // bits.shape[-1]
    points = _cuda(points, torch.int32, "points").reshape(-1, 2)
    out = torch.empty((n,), device=bits.device, dtype=torch.int32)
    check(lib.marsb200_points_in_masks(bits.data_ptr(), n, h, w, points.data_ptr(), points.shape[0], out.data_ptr(), _stream()))
    return out


def matcher_scores(points_in, pooled_count, emd, k: int, alpha: float, beta: float, exp: float):
    n = points_in.numel()
    emd = _cuda(emd, torch.float32, "emd")
    purity = torch.empty((n,), device=emd.device, dtype=torch.float32)
    coverage = torch.empty_like(purity)
    scores = torch.empty_like(purity)
    check(lib.marsb200_matcher_scores(points_in.data_ptr(), pooled_count.data_ptr(), emd.data_ptr(), n, k, alpha, beta,
                                      exp, purity.data_ptr(), coverage.data_ptr(), scores.data_ptr(), _stream()))
    return purity, coverage, scores


def eval_areas(pred: torch.Tensor, gt: torch.Tensor, ignore: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pred, gt, ignore [n, H, W] -> int32 [n, 4] = {inter_bg, inter_fg, union_bg, union_fg}."""
    pred = _cuda(pred, torch.float32, "pred")
    gt = _cuda(gt, torch.float32, "gt")
    n = pred.shape[0]
    hw = pred.numel() // n
    ig = None if ignore is None else _cuda(ignore, torch.float32, "ignore")
    out = torch.empty((n, 4), device=pred.device, dtype=torch.int32)
    check(lib.marsb200_eval_areas(pred.data_ptr(), gt.data_ptr(), _ptr(ig), n, hw, out.data_ptr(), _stream()))
    return out


def eval_accumulate(areas: torch.Tensor, class_id: torch.Tensor, inter_buf: torch.Tensor, union_buf: torch.Tensor,
                    check_status: bool = False) -> None:
    """inter_buf / union_buf [2, nclass] int64 += the areas [n, 4] of the samples' classes (AverageMeter.update)."""
    areas = _cuda(areas, torch.int32, "areas")
    class_id = _cuda(class_id.reshape(-1), torch.int64, "class_id")
    n = areas.shape[0]
    assert class_id.numel() == n and inter_buf.dtype == torch.int64 and union_buf.dtype == torch.int64
    status = torch.zeros(1, device=areas.device, dtype=torch.int32)
    check(lib.marsb200_eval_accumulate(areas.data_ptr(), class_id.data_ptr(), n, inter_buf.shape[1], inter_buf.data_ptr(),
                                       union_buf.data_ptr(), status.data_ptr(), _stream()))
    if check_status and int(status.item()):
        raise _lib.MarsB200Error("eval_accumulate: class id outside [0, nclass)")


def eval_iou(inter_buf: torch.Tensor, union_buf: torch.Tensor, interest: torch.Tensor) -> torch.Tensor:
    """-> float64 [2 + k] = {mIoU, FB-IoU, fg IoU per class of interest} (AverageMeter.compute_iou)."""
    interest = _cuda(interest.reshape(-1), torch.int64, "interest")
    out = torch.empty(2 + interest.numel(), device=inter_buf.device, dtype=torch.float64)
    check(lib.marsb200_eval_iou(inter_buf.data_ptr(), union_buf.data_ptr(), inter_buf.shape[1], interest.data_ptr(),
                                interest.numel(), out.data_ptr(), _stream()))
    return out


# ----------------------------------------------------------------------------- 8f-4: wire format / AMG post-processing
def rle_decode(counts: torch.Tensor, offsets: torch.Tensor, h: int, w: int, out=None, check_status: bool = True,
               workspace=None):
    """Uncompressed COCO RLE -> packed bits [n, words_per_mask(h*w)].

    counts int32 [total] (all masks concatenated, each starting with a run of zeros), offsets int64 [n + 1].
    """
    counts = _cuda(counts, torch.int32, "counts")
    offsets = _cuda(offsets, torch.int64, "offsets")
    n = offsets.numel() - 1
    if out is None:
        out = torch.empty((n, words_per_mask(h * w)), device=counts.device, dtype=torch.int32)
    nbytes = int(lib.marsb200_rle_workspace_bytes(n, h, w))
    ws = workspace if workspace is not None and workspace.numel() >= nbytes else \
        torch.empty(nbytes, device=counts.device, dtype=torch.uint8)
    status = torch.zeros(1, device=counts.device, dtype=torch.int32)
    check(lib.marsb200_rle_decode(counts.data_ptr(), offsets.data_ptr(), n, h, w, out.data_ptr(), ws.data_ptr(),
                                  ws.numel(), status.data_ptr(), _stream()))
    if check_status and int(status.item()):
        raise _lib.MarsB200Error("rle_decode: malformed RLE (counts of a mask do not sum to H*W)")
    return out


def mask_boxes(bits: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """packed bits [..., wpm] -> XYXY boxes int32 [..., 4] ([0,0,0,0] for empty masks)."""
    lead = tuple(bits.shape[:-1])
    n = bits.numel() // bits.shape[-1]
    out = torch.empty(lead + (4,), device=bits.device, dtype=torch.int32)
    check(lib.marsb200_mask_boxes(bits.data_ptr(), n, h, w, out.data_ptr(), _stream()))
    return out


def stability_score(logits: torch.Tensor, mask_threshold: float, threshold_offset: float):
    """logits [n, H, W] fp32 -> (score fp32 [n], counts int32 [n, 2] = pixels above t + o / above t - o)."""
    logits = _cuda(logits, torch.float32, "logits")
    n = logits.shape[0]
    out = torch.empty(n, device=logits.device, dtype=torch.float32)
    counts = torch.empty((n, 2), device=logits.device, dtype=torch.int32)
    check(lib.marsb200_stability_score(logits.data_ptr(), n, logits.numel() // n, float(mask_threshold),
                                       float(threshold_offset), out.data_ptr(), counts.data_ptr(), _stream()))
    return out, counts


def box_nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float):
    """torchvision.ops.nms semantics.  Returns (keep_idx int64 sorted by decreasing score, order, keep mask)."""
    boxes = _cuda(boxes, torch.float32, "boxes")
    scores = _cuda(scores, torch.float32, "scores")
    n = boxes.shape[0]
    order = torch.empty(n, device=boxes.device, dtype=torch.int32)
    keep = torch.empty(n, device=boxes.device, dtype=torch.uint8)
    n_keep = torch.zeros(1, device=boxes.device, dtype=torch.int32)
    ws = torch.empty(int(lib.marsb200_box_nms_workspace_bytes(n)), device=boxes.device, dtype=torch.uint8)
    check(lib.marsb200_box_nms(boxes.data_ptr(), scores.data_ptr(), n, float(iou_threshold), order.data_ptr(),
                               keep.data_ptr(), n_keep.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    o = order.long()
    return o[keep[o] != 0], order, keep


# ----------------------------------------------------------------------------- A13: mask-pooled features / statistics
def masked_feature_means(pooled: torch.Tensor, feats: torch.Tensor, backend=None) -> torch.Tensor:
    """pooled bitmaps [E, P, npw] x feats [E, N, C] -> prototypes [E, P, C] (mean of the features under each mask)."""
    feats = _cuda(feats, torch.float32, "feats")
    if feats.dim() == 2:
        feats = feats[None]
    e, n, c = feats.shape
    pooled = pooled.reshape(e, -1, pooled.shape[-1]).contiguous()
    p = pooled.shape[1]
    out = torch.empty((e, p, c), device=feats.device, dtype=torch.float32)
    ws = torch.empty(int(lib.marsb200_masked_feature_means_workspace_bytes(e, p, n, c)), device=feats.device, dtype=torch.uint8)
    check(lib.marsb200_masked_feature_means(pooled.data_ptr(), feats.data_ptr(), e, p, n, c, out.data_ptr(), ws.data_ptr(),
                                            ws.numel(), DEFAULT_GEMM if backend is None else backend, _stream()))
    return out


def masked_sim_stats(sim: torch.Tensor, row_mask: torch.Tensor, col_mask: torch.Tensor) -> torch.Tensor:
    """sim [E, M, N] -> float64 [E, 4] = {mean, max, unbiased std, count} over sim[row_mask][:, col_mask]."""
    sim = _cuda(sim, torch.float32, "sim")
    if sim.dim() == 2:
        sim = sim[None]
    e, m, n = sim.shape
    rm = _cuda(row_mask, torch.uint8, "row_mask").reshape(e, m)
    cm = _cuda(col_mask, torch.uint8, "col_mask").reshape(e, n)
    out = torch.empty((e, 4), device=sim.device, dtype=torch.float64)
    check(lib.marsb200_masked_sim_stats(sim.data_ptr(), rm.data_ptr(), cm.data_ptr(), e, m, n, out.data_ptr(), _stream()))
    return out


def masked_row_mean(sim: torch.Tensor, row_mask: torch.Tensor) -> torch.Tensor:
    """sim [E, M, N] -> [E, N]: mean over the selected rows."""
    sim = _cuda(sim, torch.float32, "sim")
    if sim.dim() == 2:
        sim = sim[None]
    e, m, n = sim.shape
    rm = _cuda(row_mask, torch.uint8, "row_mask").reshape(e, m)
    out = torch.empty((e, n), device=sim.device, dtype=torch.float32)
    check(lib.marsb200_masked_row_mean(sim.data_ptr(), rm.data_ptr(), e, m, n, out.data_ptr(), _stream()))
    return out
