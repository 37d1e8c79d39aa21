"""Host-side plans: one `DecoderPlan` = one TTSZipformer over a fixed (N rows, T frames) shape,
holding the PyTorch-owned workspace and the opaque C-ABI plan handle."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .weights import PackedZipformer


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class DecoderPlan:
    def __init__(self, packed: PackedZipformer, N: int, T: int):
        if packed.device.type != "cuda":
            raise _lib.ZvbError("zipvoice_b200 runs on a B200 only: weights must live on a CUDA device")
        self.lib = _lib.load()
        self.packed = packed
        self.N, self.T = int(N), int(T)
        self.model, self._keep = packed.model_struct(self.T)
        nbytes = C.c_size_t(0)
        _lib.check(self.lib.zvb_plan_workspace_bytes(C.byref(self.model), self.N, self.T, C.byref(nbytes)))
        self.workspace_bytes = int(nbytes.value)
        with torch.cuda.device(packed.device):
            self.workspace = torch.zeros(self.workspace_bytes, dtype=torch.uint8, device=packed.device)
            handle = C.c_void_p()
            _lib.check(self.lib.zvb_plan_create(C.byref(self.model), self.N, self.T, self.workspace.data_ptr(),
                                                self.workspace_bytes, C.byref(handle)))
        self.handle = handle
        io = _lib.zvb_io()
        _lib.check(self.lib.zvb_plan_io(self.handle, C.byref(io)))
        self.io = io
        self.out_dim = packed.out_dim
        self.in_dim = packed.in_dim

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.zvb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _view(self, ptr: int, nbytes: int) -> torch.Tensor:
        off = ptr - self.workspace.data_ptr()
        return self.workspace[off: off + nbytes]

    def out_view(self) -> torch.Tensor:
        n = self.N * self.T * self.out_dim
        return self._view(self.io.out, n * 4).view(torch.float32).view(self.N, self.T, self.out_dim)

    def forward_f32(self, x: torch.Tensor, t: Optional[torch.Tensor], mask: torch.Tensor,
                    g: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Seam 1: fm_decoder(x, t, padding_mask, guidance_scale) (reference: zipformer.py:242-293)."""
        assert x.shape == (self.N, self.T, self.in_dim), (x.shape, (self.N, self.T, self.in_dim))
        x = x.contiguous().float()
        mask8 = mask.contiguous().to(torch.uint8)
        t32 = t.contiguous().float() if t is not None else None
        g32 = g.contiguous().float() if g is not None else None
        out = torch.empty(self.N, self.T, self.out_dim, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(self.lib.zvb_decoder_forward_f32(
                self.handle, x.data_ptr(), t32.data_ptr() if t32 is not None else None, mask8.data_ptr(),
                g32.data_ptr() if g32 is not None else None, out.data_ptr(), _stream_ptr()))
        return out

    def profile(self, shapes: bool = False, with_bytes: bool = False):
        """One forward over the resident buffers with a CUDA event per kernel (synchronises).
        Returns a list of (category name, milliseconds, algorithmic work[, algorithmic HBM bytes]
        [, (rows, cols, K, tile N)]); work = FLOPs for the tensor-core kernels, bytes for the others."""
        import numpy as np
        cap = 4096
        ms = np.zeros(cap, dtype=np.float32)
        cat = np.zeros(cap, dtype=np.int32)
        work = np.zeros(cap, dtype=np.float64)
        nbytes = np.zeros(cap, dtype=np.float64)
        shp = np.zeros(cap * 4, dtype=np.int32)
        n = C.c_int(0)
        with torch.cuda.device(self.packed.device):
            _lib.check(self.lib.zvb_decoder_profile(self.handle, _stream_ptr(), cap, ms.ctypes.data, cat.ctypes.data,
                                                    work.ctypes.data, nbytes.ctypes.data, shp.ctypes.data, C.byref(n)))
        out = []
        for i in range(n.value):
            row = [_lib.CATEGORIES[int(cat[i])], float(ms[i]), float(work[i])]
            if with_bytes:
                row.append(float(nbytes[i]))
            if shapes:
                row.append(tuple(int(x) for x in shp[4 * i: 4 * i + 4]))
            out.append(tuple(row))
        return out

    def sample(self, x: torch.Tensor, text: torch.Tensor, speech: torch.Tensor, mask8: torch.Tensor,
               guidance: Optional[torch.Tensor], ts_dev: torch.Tensor, ts_host: torch.Tensor, num_step: int,
               mode: int, vrec: Optional[torch.Tensor] = None) -> None:
        """Seam 2 inner loop, in place on `x` (all tensors contiguous, on the plan's device)."""
        B, T, F = x.shape
        Ft = text.shape[2]
        assert ts_host.dtype == torch.float32 and ts_host.device.type == "cpu" and ts_host.numel() == num_step + 1
        with torch.cuda.device(x.device):
            _lib.check(self.lib.zvb_sample(
                self.handle, x.data_ptr(), text.data_ptr(), speech.data_ptr(), mask8.data_ptr(),
                guidance.data_ptr() if guidance is not None else None, ts_dev.data_ptr(), ts_host.data_ptr(),
                int(num_step), int(mode), int(B), int(F), int(Ft),
                vrec.data_ptr() if vrec is not None else None, _stream_ptr()))


class PlanCache:
    """(N, T) -> DecoderPlan for one packed network."""

    def __init__(self, packed: PackedZipformer, max_plans: int = 8):
        self.packed = packed
        self.max_plans = max_plans
        self._plans: Dict[Tuple[int, int], DecoderPlan] = {}

    def get(self, N: int, T: int) -> DecoderPlan:
        key = (int(N), int(T))
        if key not in self._plans:
            if len(self._plans) >= self.max_plans:
                self._plans.pop(next(iter(self._plans)))
            self._plans[key] = DecoderPlan(self.packed, N, T)
        return self._plans[key]
