"""Spatial partition of one GPU's SMs into two CUDA green contexts.

The ranking stage has two kinds of work per step: the mask ingest streams the proposals once at the HBM roofline
(pack_masks), everything that follows it on the alignment side is tensor-core or latency bound.  Time-sliced on the
same SMs they serialise (a step is the sum of its kernels); on disjoint SM sets they run side by side.
`SmPartition` carves the device into a `tensor` partition of about `tensor_sms` SMs and an `hbm` partition with the
rest, and hands out one CUDA stream per partition as `torch.cuda.ExternalStream`s.  Kernels launched on such a stream
only occupy their partition; libmarsb200's persistent kernels size their grids from `marsb200_stream_sm_count`.

No reference counterpart (the reference is single-stream Python, main_MARS.py:54-94).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _ok(ret):
    from cuda.bindings import driver as drv

    err = ret[0]
    if err != drv.CUresult.CUDA_SUCCESS:
        raise _lib.MarsB200Error(f"CUDA driver call failed: {err!r}")
    rest = ret[1:]
    return rest[0] if len(rest) == 1 else rest


def stream_sm_count(stream: torch.cuda.Stream) -> int:
    n = ctypes.c_int(0)
    _lib.check(_lib.lib.marsb200_stream_sm_count(ctypes.c_void_p(stream.cuda_stream), ctypes.byref(n)))
    return n.value


def set_stream_sm_cap(stream: torch.cuda.Stream, cap: int) -> None:
    """Caps the SM count libmarsb200's persistent kernels see for `stream` (0 removes the cap)."""
    _lib.check(_lib.lib.marsb200_stream_set_sm_cap(ctypes.c_void_p(stream.cuda_stream), int(cap)))


class SmPartition:
    """Two disjoint SM sets of `device`: `.tensor_stream` (>= tensor_sms SMs) and `.hbm_stream` (the remainder)."""

    def __init__(self, device, tensor_sms: int):
        from cuda.bindings import driver as drv

        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SmPartition needs a CUDA device")
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()  # primary context is live
            dev = _ok(drv.cuDeviceGet(self.device.index or 0))
            sm_type = drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM
            whole = _ok(drv.cuDeviceGetDevResource(dev, sm_type))
            groups, n_groups, remaining = _ok(drv.cuDevSmResourceSplitByCount(1, whole, 0, int(tensor_sms)))
            if n_groups < 1 or remaining.sm.smCount == 0:
                raise _lib.MarsB200Error(f"cannot split {whole.sm.smCount} SMs into {tensor_sms} + rest")
            self._ctx, self._streams = [], []
            flags = drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM
            for res in (groups[0], remaining):
                desc = _ok(drv.cuDevResourceGenerateDesc([res], 1))
                ctx = _ok(drv.cuGreenCtxCreate(desc, dev, int(flags)))
                stream = _ok(drv.cuGreenCtxStreamCreate(ctx, int(drv.CUstream_flags.CU_STREAM_NON_BLOCKING), 0))
                self._ctx.append(ctx)
                self._streams.append(stream)
            self.tensor_stream = torch.cuda.ExternalStream(int(self._streams[0]), device=self.device)
            self.hbm_stream = torch.cuda.ExternalStream(int(self._streams[1]), device=self.device)
            self.tensor_sms = int(groups[0].sm.smCount)
            self.hbm_sms = int(remaining.sm.smCount)

    def extra_stream(self, which: str) -> torch.cuda.ExternalStream:
        """Another stream inside the `tensor` or `hbm` partition."""
        from cuda.bindings import driver as drv

        ctx = self._ctx[0 if which == "tensor" else 1]
        stream = _ok(drv.cuGreenCtxStreamCreate(ctx, int(drv.CUstream_flags.CU_STREAM_NON_BLOCKING), 0))
        self._streams.append(stream)
        return torch.cuda.ExternalStream(int(stream), device=self.device)

    def close(self):
        from cuda.bindings import driver as drv

        torch.cuda.synchronize(self.device)
        for s in self._streams:
            drv.cuStreamDestroy(s)
        for c in self._ctx:
            drv.cuGreenCtxDestroy(c)
        self._streams, self._ctx = [], []
