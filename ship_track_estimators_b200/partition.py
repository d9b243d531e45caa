"""Two CUDA streams on disjoint sets of SMs, for running the two passes of different tiles side by side.

The forward filter is bound by the FP64 pipe and leaves DRAM idle; the tape smoother is bound by DRAM
latency / bandwidth and leaves the FP64 pipe idle.  On the SAME SMs their two instruction streams do not
fit the instruction cache together (``ukf_roles_kernel``: slower than two launches).  On
DISJOINT SMs each kernel keeps its own registers, occupancy and instruction cache, and the smoother
confined to a third of the SMs loses less than proportionally (it is short of loads in flight, not of
SMs): measured on one B200 (``tools/green_ctx_probe.py``, GPU calls 57 / 58), smoother of tile i on 48
SMs beside filter of tile i+1 on 100 SMs takes 11.50 ms against 12.16 ms for the two launches back to
back on all 148 SMs (151 552 x 512 tiles).

The partition is a pair of CUDA *green contexts* (driver API, CUDA >= 12.4, reached through
``cuda.bindings.driver`` of the ``cuda-python`` package): the device's SM resource split in two, one
stream created on each.  Memory is the primary context's, so every tensor is visible from both streams,
and the streams order against ordinary ones with events as usual.  The reference has no counterpart
(it is a sequential Python loop over ships, ``examples/example_ukf_rts_smoother_batch.py:19-90``).
"""
from __future__ import annotations

from typing import Optional

import torch


class SmPartition:
    """``smoother_stream`` on at least ``smoother_sms`` SMs (rounded up to the hardware's granularity,
    8 on a B200), ``filter_stream`` on the remaining ones.  Raises ``RuntimeError`` when green contexts
    are not available (no ``cuda-python``, an old driver); nothing falls back silently."""

    def __init__(self, device: Optional[torch.device] = None, smoother_sms: int = 48):
        try:
            from cuda.bindings import driver as cu
        except Exception as exc:  # pragma: no cover - the image ships cuda-python
            raise RuntimeError("SmPartition needs the cuda-python package (cuda.bindings.driver)") from exc
        self._cu = cu
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(self.device):
            torch.cuda.init()
            torch.zeros(1, device=self.device)   # the primary context exists and is current
            dev = self._ck(cu.cuDeviceGet(self.device.index or 0))
            sm_all = self._ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
            self.total_sms = int(sm_all.sm.smCount)
            if not 0 < smoother_sms < self.total_sms:
                raise ValueError(f"smoother_sms must lie in (0, {self.total_sms})")
            groups, _, rest = self._ck(cu.cuDevSmResourceSplitByCount(1, sm_all, 0, smoother_sms))
            self._ctx, streams, counts = [], [], []
            for res in (groups[0], rest):
                desc = self._ck(cu.cuDevResourceGenerateDesc([res], 1))
                ctx = self._ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
                raw = self._ck(cu.cuGreenCtxStreamCreate(ctx, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
                self._ctx.append((ctx, raw))
                streams.append(torch.cuda.ExternalStream(int(raw), device=self.device))
                counts.append(int(res.sm.smCount))
        self.smoother_stream, self.filter_stream = streams
        self.smoother_sms, self.filter_sms = counts
        if self.filter_sms <= 0:
            raise RuntimeError("the split left no SMs for the filter")

    @staticmethod
    def _ck(ret):
        err, rest = ret[0], ret[1:]
        if int(err) != 0:
            raise RuntimeError(f"CUDA driver call failed: {err}")
        return rest[0] if len(rest) == 1 else rest

    def close(self) -> None:
        """Destroy the two streams and green contexts (after synchronising them)."""
        cu = self._cu
        for ctx, raw in self._ctx:
            cu.cuStreamSynchronize(raw)
            cu.cuStreamDestroy(raw)
            cu.cuGreenCtxDestroy(ctx)
        self._ctx = []

    def __enter__(self) -> "SmPartition":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def __repr__(self) -> str:
        return f"SmPartition(filter_sms={self.filter_sms}, smoother_sms={self.smoother_sms}, device={self.device})"
