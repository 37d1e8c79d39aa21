"""Host-side plans: one `DecoderPlan` = one TTSZipformer over a fixed (N rows, T frames) shape,
holding the PyTorch-owned workspace and the opaque C-ABI plan handle."""
from __future__ import annotations

import collections
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
        # CUDA graphs captured over this plan's workspace (model.B200EulerSolver) live and die with the plan:
        # evicting the plan drops them too, so no captured graph can outlive the buffers it points into
        self.graphs: Dict[tuple, dict] = {}
        self._sat_counter: Optional[torch.Tensor] = None

    @property
    def nbytes(self) -> int:
        return self.workspace_bytes + sum(st.get("nbytes", 0) for st in self.graphs.values())

    def enable_saturation_check(self) -> None:
        """Every forward of this plan also counts, kernel by kernel, fp16 outputs that hit the largest finite
        fp16 magnitude (parity aid, include/zipvoice_b200.h: zvb_plan_set_saturation_counter)."""
        if self._sat_counter is None:
            self._sat_counter = torch.zeros(1, dtype=torch.int64, device=self.packed.device)
            _lib.check(self.lib.zvb_plan_set_saturation_counter(self.handle, self._sat_counter.data_ptr()))
            self.graphs.clear()             # graphs captured before the switch do not contain the scans

    def saturated(self) -> int:
        return int(self._sat_counter.item()) if self._sat_counter is not None else 0

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


def round_up(x: int, step: int) -> int:
    return (int(x) + step - 1) // step * step if step and step > 1 else int(x)


class PlanCache:
    """(rows, frames) -> DecoderPlan for one packed network: shape-bucketed, least-recently-used, bounded by
    plan count AND bytes (a plan for N=128, T=1219 holds a 4.5 GB workspace).

    `frame_bucket` / `row_bucket` > 1 round the requested shape up (reference counterpart: one TensorRT engine
    serves N 1-4 / T 100-3000, zipvoice/bin/tensorrt_export.py:112-131); the caller pads its inputs and masks
    the extra frames / rows, which is exactly the reference batching a short utterance with a longer one.
    0 / 1 = exact shapes."""

    def __init__(self, packed: PackedZipformer, max_plans: int = 12, max_bytes: Optional[int] = None,
                 frame_bucket: int = 0, row_bucket: int = 0, factory=None):
        self.packed = packed
        self.factory = factory or DecoderPlan
        self.max_plans = max_plans
        self.max_bytes = max_bytes
        self.frame_bucket = frame_bucket
        self.row_bucket = row_bucket
        self.check_saturation = False
        self.created = 0
        self._plans: "collections.OrderedDict[Tuple[int, int], DecoderPlan]" = collections.OrderedDict()

    def shape_for(self, N: int, T: int) -> Tuple[int, int]:
        return round_up(N, self.row_bucket), round_up(T, self.frame_bucket)

    def _budget(self) -> int:
        if self.max_bytes is not None:
            return self.max_bytes
        if self.packed.device.type == "cuda":
            return int(0.6 * torch.cuda.get_device_properties(self.packed.device).total_memory)
        return 1 << 62

    def get(self, N: int, T: int) -> DecoderPlan:
        """The plan covering (N, T): its N / T are the bucketed sizes (>= the request)."""
        key = self.shape_for(N, T)
        plan = self._plans.get(key)
        if plan is None:
            plan = self.factory(self.packed, *key)
            if self.check_saturation:
                plan.enable_saturation_check()
            self._plans[key] = plan
            self.created += 1
        self._plans.move_to_end(key)
        budget = self._budget()
        while len(self._plans) > 1 and (len(self._plans) > self.max_plans or
                                        sum(p.nbytes for p in self._plans.values()) > budget):
            self._plans.popitem(last=False)        # least recently used; its graphs go with it
        return plan

    def __len__(self) -> int:
        return len(self._plans)

    def plans(self):
        return list(self._plans.values())
