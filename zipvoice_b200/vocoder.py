"""The step right after the sampler: mel -> waveform with the Vocos vocoder on the GPU.

The reference calls the external `vocos` package (vocos 0.1.0, uv.lock:1239-1241; `Vocos.from_pretrained(
"charactr/vocos-mel-24khz")`, reference: zipvoice/bin/infer_zipvoice.py:301-312) as
    wav = vocoder.decode(pred_features).squeeze(1).clamp(-1, 1)          # infer_zipvoice.py:409, 594-598
with `pred_features` = (B, 100, T) log-mel (already divided by feat_scale).  This class keeps that surface
(`decode`, state_dict keys of the published checkpoint: `backbone.*`, `head.*`) and runs the model -- VocosBackbone
(Conv1d embed, LayerNorm, 8 ConvNeXt blocks, LayerNorm) + ISTFTHead(padding="center") -- through the C ABI:
tensor-core GEMMs with GELU / layer-scale / residual epilogues, the depthwise convolution kernel, LayerNorm, and an
fp32 FFT inverse STFT (csrc/audio.cuh, csrc/engine.cu: build_vocoder).  `decode_batch` decodes a padded batch with
per-utterance lengths exactly as if every utterance were decoded alone (what the reference's batch loop does,
infer_zipvoice.py:590-603).  No CPU path."""
from __future__ import annotations

import collections
import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib


def _ceil8(x: int) -> int:
    return (x + 7) // 8 * 8


def synth_vocos_state_dict(seed: int = 0, dim: int = 512, intermediate: int = 1536, n_layers: int = 8, n_mels: int = 100,
                           n_fft: int = 1024) -> Dict[str, torch.Tensor]:
    """Random weights with the key set / shapes of the published vocos-mel-24khz checkpoint (no checkpoint offline).
    Magnitudes follow vocos' own initialisation (trunc_normal std 0.02 for conv / linear weights, layer scale
    1 / n_layers) with gains so that the head produces log-magnitudes of order 1 and phases over several radians."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    sd = {
        "backbone.embed.weight": rn(dim, n_mels, 7, std=0.05), "backbone.embed.bias": rn(dim, std=0.05),
        "backbone.norm.weight": 1.0 + rn(dim, std=0.1), "backbone.norm.bias": rn(dim, std=0.1),
        "backbone.final_layer_norm.weight": 1.0 + rn(dim, std=0.1), "backbone.final_layer_norm.bias": rn(dim, std=0.1),
        "head.out.weight": rn(n_fft + 2, dim, std=0.06), "head.out.bias": rn(n_fft + 2, std=0.3),
        "head.istft.window": torch.hann_window(n_fft),
    }
    for i in range(n_layers):
        p = f"backbone.convnext.{i}."
        sd[p + "dwconv.weight"] = rn(dim, 1, 7, std=0.3)
        sd[p + "dwconv.bias"] = rn(dim, std=0.1)
        sd[p + "norm.weight"] = 1.0 + rn(dim, std=0.1)
        sd[p + "norm.bias"] = rn(dim, std=0.1)
        sd[p + "pwconv1.weight"] = rn(intermediate, dim, std=0.06)
        sd[p + "pwconv1.bias"] = rn(intermediate, std=0.1)
        sd[p + "pwconv2.weight"] = rn(dim, intermediate, std=0.06)
        sd[p + "pwconv2.bias"] = rn(dim, std=0.1)
        sd[p + "gamma"] = 0.3 + rn(dim, std=0.05)
    return sd


class PackedVocos:
    """Device-resident, kernel-layout weights of one Vocos model + the `zvb_vocoder` struct."""

    def __init__(self, sd: Dict[str, torch.Tensor], device, hop_length: int = 256):
        self.device = torch.device(device)
        self._keep: List[torch.Tensor] = []
        f = lambda k: sd[k].detach().float().cpu()
        emb = f("backbone.embed.weight")                                    # (dim, n_mels, 7)
        self.dim, self.n_mels, self.kernel = emb.shape
        self.n_layers = len([k for k in sd if k.startswith("backbone.convnext.") and k.endswith(".gamma")])
        self.intermediate = sd["backbone.convnext.0.pwconv1.weight"].shape[0]
        self.n_fft = sd["head.out.weight"].shape[0] - 2
        self.hop = int(hop_length)
        assert self.n_layers <= _lib.ZVB_VOC_MAX_LAYERS
        v = _lib.zvb_vocoder()
        v.abi_version = _lib.ZVB_ABI_VERSION
        v.dim, v.intermediate, v.n_layers, v.n_mels = self.dim, self.intermediate, self.n_layers, self.n_mels
        v.n_fft, v.hop, v.kernel = self.n_fft, self.hop, self.kernel
        # Conv1d(n_mels, dim, 7) as a linear over the window operand A[(n,t)][k * n_mels + c] (csrc/audio.cuh)
        v.embed = self._linear(emb.permute(0, 2, 1).reshape(self.dim, self.kernel * self.n_mels), f("backbone.embed.bias"))
        v.norm_w, v.norm_b = self._ptr(f("backbone.norm.weight")), self._ptr(f("backbone.norm.bias"))
        for i in range(self.n_layers):
            p = f"backbone.convnext.{i}."
            ly = v.layers[i]
            ly.dw_w = self._ptr(f(p + "dwconv.weight").reshape(self.dim, self.kernel).t())       # [7][dim]
            ly.dw_b = self._ptr(f(p + "dwconv.bias"))
            ly.ln_w, ly.ln_b = self._ptr(f(p + "norm.weight")), self._ptr(f(p + "norm.bias"))
            ly.pw1 = self._linear(f(p + "pwconv1.weight"), f(p + "pwconv1.bias"))
            gamma = f(p + "gamma")                                         # layer scale folded into pwconv2
            ly.pw2 = self._linear(f(p + "pwconv2.weight") * gamma[:, None], f(p + "pwconv2.bias") * gamma)
        v.final_w = self._ptr(f("backbone.final_layer_norm.weight"))
        v.final_b = self._ptr(f("backbone.final_layer_norm.bias"))
        v.head = self._linear(f("head.out.weight"), f("head.out.bias"))
        win = f("head.istft.window") if "head.istft.window" in sd else torch.hann_window(self.n_fft)
        v.window = self._ptr(win)
        self.struct = v

    def _ptr(self, t: torch.Tensor) -> int:
        t = t.contiguous().float().to(self.device)
        self._keep.append(t)
        return t.data_ptr()

    def _linear(self, w: torch.Tensor, b: Optional[torch.Tensor]) -> _lib.zvb_linear:
        out_f, in_f = w.shape
        buf = torch.zeros(out_f, _ceil8(in_f), dtype=torch.float16)
        buf[:, :in_f] = w.clamp(-65504.0, 65504.0).to(torch.float16)
        buf = buf.to(self.device)
        self._keep.append(buf)
        return _lib.zvb_linear(buf.data_ptr(), self._ptr(b) if b is not None else None, out_f, in_f, buf.shape[1], out_f)


class VocoderPlan:
    def __init__(self, packed: PackedVocos, N: int, T: int):
        self.lib = _lib.load()
        self.packed, self.N, self.T = packed, int(N), int(T)
        n = C.c_size_t(0)
        _lib.check(self.lib.zvb_vocoder_workspace_bytes(C.byref(packed.struct), self.N, self.T, C.byref(n)))
        self.workspace_bytes = int(n.value)
        with torch.cuda.device(packed.device):
            self.workspace = torch.zeros(self.workspace_bytes, dtype=torch.uint8, device=packed.device)
            h = C.c_void_p()
            _lib.check(self.lib.zvb_vocoder_create(C.byref(packed.struct), self.N, self.T, self.workspace.data_ptr(),
                                                   self.workspace_bytes, C.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.zvb_vocoder_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def decode(self, mel: torch.Tensor, lens: torch.Tensor, scale: float, clamp: bool) -> torch.Tensor:
        assert mel.shape == (self.N, self.T, self.packed.n_mels) and mel.dtype == torch.float32 and mel.is_contiguous()
        assert lens.dtype == torch.int32 and lens.numel() == self.N
        wav = torch.empty(self.N, self.packed.hop * (self.T - 1), dtype=torch.float32, device=mel.device)
        with torch.cuda.device(mel.device):
            _lib.check(self.lib.zvb_vocoder_decode(self.handle, mel.data_ptr(), lens.data_ptr(), float(scale), int(clamp),
                                                   wav.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return wav

    def profile(self):
        import numpy as np
        cap = 256
        ms, cat = np.zeros(cap, np.float32), np.zeros(cap, np.int32)
        work, nbytes = np.zeros(cap, np.float64), np.zeros(cap, np.float64)
        n = C.c_int(0)
        with torch.cuda.device(self.packed.device):
            _lib.check(self.lib.zvb_vocoder_profile(self.handle, torch.cuda.current_stream().cuda_stream, cap, ms.ctypes.data,
                                                    cat.ctypes.data, work.ctypes.data, nbytes.ctypes.data, C.byref(n)))
        return [(_lib.CATEGORIES[int(cat[i])], float(ms[i]), float(work[i])) for i in range(n.value)]


class Vocos:
    """`vocoder` object of the reference's inference scripts (reference: infer_zipvoice.py:301-312, 409, 594)."""

    def __init__(self, hop_length: int = 256, frame_bucket: int = 64, row_bucket: int = 1, max_plans: int = 8):
        self.hop_length = hop_length
        self.frame_bucket, self.row_bucket, self.max_plans = frame_bucket, row_bucket, max_plans
        self.device = torch.device("cpu")
        self._sd: Optional[Dict[str, torch.Tensor]] = None
        self.packed: Optional[PackedVocos] = None
        self._plans: "collections.OrderedDict[Tuple[int, int], VocoderPlan]" = collections.OrderedDict()

    # nn.Module-like surface (`vocoder.load_state_dict(...)`, `.to(device).eval()`, infer_zipvoice.py:303-311, 833)
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        self._sd = {k: v.detach() for k, v in sd.items()}
        if self.device.type == "cuda":
            self._materialise()
        return self

    def to(self, device):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if self._sd is not None and self.device.type == "cuda":
            self._materialise()
        return self

    def eval(self):
        return self

    def _materialise(self):
        self.packed = PackedVocos(self._sd, self.device, self.hop_length)
        self._plans.clear()

    def _plan(self, N: int, T: int) -> VocoderPlan:
        if self.packed is None:
            raise _lib.ZvbError("vocoder is not on a CUDA device: call load_state_dict(...) and .to('cuda') "
                                "(zipvoice_b200 has no CPU path)")
        rb, fb = max(1, self.row_bucket), max(1, self.frame_bucket)
        key = ((N + rb - 1) // rb * rb, max(2, (T + fb - 1) // fb * fb))
        p = self._plans.get(key)
        if p is None:
            p = VocoderPlan(self.packed, *key)
            self._plans[key] = p
        self._plans.move_to_end(key)
        while len(self._plans) > self.max_plans:
            self._plans.popitem(last=False)
        return p

    @torch.inference_mode()
    def decode_batch(self, mel: torch.Tensor, lens: torch.Tensor, scale: float = 1.0, clamp: bool = False
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
        """mel (B, T, n_mels) time-major rows as `model.sample` returns them, lens (B,) frames; every utterance is
        decoded as if alone.  Returns (wav (B, hop * (max len - 1)) zero padded, samples per utterance)."""
        B, T, _ = mel.shape
        lens = lens.to(self.device)
        assert int(lens.min()) >= 1 and int(lens.max()) <= T
        plan = self._plan(B, T)
        m = mel.to(self.device, torch.float32)
        if (plan.N, plan.T) != (B, T):
            mp = torch.zeros(plan.N, plan.T, m.shape[2], dtype=torch.float32, device=self.device)
            mp[:B, :T] = m
            lp = torch.ones(plan.N, dtype=torch.int32, device=self.device)
            lp[:B] = lens.to(torch.int32)
        else:
            mp, lp = m.contiguous(), lens.to(torch.int32).contiguous()
        wav = plan.decode(mp, lp, scale, clamp)
        n_out = self.hop_length * (int(lens.max()) - 1)
        return wav[:B, :n_out], (lens - 1) * self.hop_length

    @torch.inference_mode()
    def decode(self, features_input: torch.Tensor, **kwargs) -> torch.Tensor:
        """vocos' signature: (B, C, T) or (C, T) mel -> (B, hop * (T - 1)) audio; all rows have T frames."""
        if features_input.dim() == 2:
            features_input = features_input.unsqueeze(0)
        B, _, T = features_input.shape
        mel = features_input.to(self.device, torch.float32).permute(0, 2, 1).contiguous()
        wav, _ = self.decode_batch(mel, torch.full((B,), T, dtype=torch.int64, device=self.device))
        return wav
