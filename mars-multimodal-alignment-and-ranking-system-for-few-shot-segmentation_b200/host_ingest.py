"""Proposals that are still in HOST memory -> packed bits in HBM.

The reference hands MARS.predict CPU tensors (`main_MARS.py:62-69`: the SAM proposals are loaded from disk) and moves
them to the GPU as float32 - 1.07 GB per c2 episode, which a PCIe 5 link carries at 55 GB/s = 50 episodes/s, far below
what the device ranks.  The ingest of a host batch therefore has two lanes that run at the same time:

* a `raw_fraction` of every episode's proposals crosses PCIe as it is and is packed by the device kernel
  (`marsb200_pack_masks`), and
* the rest is packed by a team of host threads (`marsb200_host_pack_masks`, AVX2) into a pinned buffer and only its
  bits (1/32 of the bytes) are copied.

Nothing is scored on the host: this is a format conversion in front of the copy, bit-identical to the device kernel.
"""
from __future__ import annotations

import torch

from . import ops


class HostMaskIngest:
    def __init__(self, episodes: int, proposals: int, height: int, width: int, device, mask_dtype=torch.float32,
                 raw_fraction: float = 0.3, threads: int = 0):
        if torch.device(device).type != "cuda":
            raise RuntimeError("HostMaskIngest feeds a CUDA device (marsb200 has no CPU path)")
        if not 0.0 <= raw_fraction <= 1.0:
            raise ValueError("raw_fraction must lie in [0, 1]")
        self.E, self.P, self.H, self.W = episodes, proposals, height, width
        self.device = torch.device(device)
        self.threads = int(threads)
        self.p_raw = min(proposals, int(round(proposals * raw_fraction)))
        wpm = ops.words_per_mask(height * width)
        self.bits_dev = torch.empty((episodes, proposals, wpm), dtype=torch.int32, device=self.device)
        self.bits_host = torch.empty((episodes, proposals - self.p_raw, wpm), dtype=torch.int32).pin_memory() \
            if self.p_raw < proposals else None
        self.raw_dev = torch.empty((episodes, self.p_raw, height, width), dtype=mask_dtype, device=self.device) \
            if self.p_raw else None
        self._copied = torch.cuda.Event()  # the pinned bit buffer has been read by the last upload's copies
        self._copied.record(torch.cuda.current_stream(self.device))

    def h2d_bytes(self, mask_itemsize: int = 4) -> int:
        """Bytes one upload moves over PCIe."""
        raw = self.E * self.p_raw * self.H * self.W * mask_itemsize
        return raw + (0 if self.bits_host is None else self.bits_host.numel() * 4)

    def upload(self, host_masks: torch.Tensor, stream: torch.cuda.Stream) -> torch.Tensor:
        """host_masks [E, P, H, W] float32 / uint8 in (pinned) host memory.  Enqueues the copies and the device-side packing
        on `stream` and packs the other lane on the host while they run (the call blocks for that long); returns the
        device tensor [E, P, wpm] the bits land in - valid once `stream` has drained."""
        if host_masks.is_cuda or tuple(host_masks.shape) != (self.E, self.P, self.H, self.W):
            raise ValueError(f"host_masks must be a CPU tensor of shape {(self.E, self.P, self.H, self.W)}")
        if self.raw_dev is not None and host_masks.dtype != self.raw_dev.dtype:
            raise TypeError(f"host_masks are {host_masks.dtype}, the ingest was built for {self.raw_dev.dtype}")
        p_raw = self.p_raw
        with torch.cuda.stream(stream):
            for e in range(self.E if p_raw else 0):  # the DMA of the raw lane runs while the host threads pack
                self.raw_dev[e].copy_(host_masks[e, :p_raw], non_blocking=True)
                ops.pack_masks(self.raw_dev[e], out=self.bits_dev[e, :p_raw])
        if self.bits_host is not None:
            self._copied.synchronize()  # the previous upload's copies out of the pinned buffer are done
            for e in range(self.E):
                ops.host_pack_masks(host_masks[e, p_raw:], out=self.bits_host[e], threads=self.threads)
            with torch.cuda.stream(stream):
                for e in range(self.E):
                    self.bits_dev[e, p_raw:].copy_(self.bits_host[e], non_blocking=True)
                self._copied.record(stream)
        return self.bits_dev
