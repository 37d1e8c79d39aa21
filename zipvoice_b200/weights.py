"""Weight packer: reference-format fp32 `state_dict` -> device buffers laid out for the kernels.

* every nn.Linear weight becomes fp16 row-major (out, k_pitch) with k_pitch = ceil8(in), zero
  padded (TMA needs 16-byte row pitches; out-of-range rows/cols are zero-filled by TMA);
* projections followed by a gate are re-ordered per 256-row tile as [128 | 128] so that both
  gate operands of an output column land in one accumulator tile:
    NonlinAttention.in_proj rows (s | x)  (reference: modules/zipformer.py:1516-1525),
    ConvolutionModule.in_proj rows (x | s) (reference: modules/zipformer.py:1657-1661);
* the depthwise kernel (D,1,K) is transposed to fp32 [K][D]; softmax(downsample.bias) and the
  per-layer rel-pos table E = linear_pos(pos_emb) (reference: modules/zipformer.py:995-1056,
  1215-1219) are folded on the host.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import Dict, List, Optional

import torch

from . import _lib
from .config import ZipformerConfig


def _ceil8(x: int) -> int:
    return (x + 7) // 8 * 8


POS_PAD = 128        # zero entries before / after the 2L-1 offsets (a tile's 255-offset window never leaves the table)


def pack_pos_table(e: torch.Tensor) -> torch.Tensor:
    """E[h][r][d] = linear_pos(pos_emb) fp32 (H, 2L-1, 4) -> the attention kernel's table (uint8 tensor):
    [H][2L-1 + 2*POS_PAD] entries of 16 bytes = fp16 column PAIRS {log2e*E[r][d], log2e*E[r+1][d]}, d = 0..3
    (entry index = r + POS_PAD, zeros outside the table: one contiguous 4080-byte bulk copy stages the
    window of a score tile, one 16-byte read serves two adjacent key columns), followed by [H] fp32:
    max_r |E[h][r]|_2, the bound of the softmax shift (reference: modules/zipformer.py:1215-1248)."""
    H, R, D = e.shape
    assert D == 4
    emax = e.float().norm(dim=2).amax(dim=1).contiguous()
    pad = torch.zeros(H, R + 2 * POS_PAD + 1, 4, dtype=torch.float32, device=e.device)
    pad[:, POS_PAD:POS_PAD + R] = e.float() * 1.4426950408889634
    pairs = torch.stack([pad[:, :-1], pad[:, 1:]], dim=-1).to(torch.float16).contiguous()   # (H, R+2P, 4, 2)
    return torch.cat([pairs.view(torch.uint8).reshape(-1), emax.view(torch.uint8).reshape(-1)]).contiguous()


TC_PADZ = 128        # zero entries in front of E in the tensor-core table (csrc/attn3.cuh: A3_PADZ)


def tc_table_entries(L: int) -> int:
    """Entries per copy of the tensor-core table (csrc/engine.cu: attn_tc_lz)."""
    return (2 * L + 264 + 1) // 2 * 2


def pack_pos_table_tc(e: torch.Tensor) -> torch.Tensor:
    """E[h][r][d] fp32 (H, 2L-1, 4) -> the table of the tensor-core attention kernel (csrc/attn3.cuh), uint8 tensor:
    [H][2][LZ] entries of 4 fp16: copy 0 holds log2e*E[h][r][0..3] at entry TC_PADZ + r and zeros elsewhere, copy 1 is
    copy 0 shifted by one entry (so that a window starting at an odd entry is 16-byte aligned in one of the two);
    followed by [H] fp32 max_r |log2e*E[h][r]|_2, the bound of the cheap softmax shift."""
    H, R, D = e.shape
    assert D == 4 and R % 2 == 1
    L = (R + 1) // 2
    LZ = tc_table_entries(L)
    es = e.float() * 1.4426950408889634
    emax = es.norm(dim=2).amax(dim=1).contiguous()
    z = torch.zeros(H, 2, LZ + 1, 4, dtype=torch.float32, device=e.device)
    z[:, 0, TC_PADZ:TC_PADZ + R] = es
    z[:, 1, :LZ] = z[:, 0, 1:LZ + 1]
    tab = z[:, :, :LZ].to(torch.float16).contiguous()
    return torch.cat([tab.view(torch.uint8).reshape(-1), emax.view(torch.uint8).reshape(-1)]).contiguous()


def rel_pos_embedding(L: int, pos_dim: int, device) -> torch.Tensor:
    """CompactRelPositionalEncoding rows for offsets -(L-1)..(L-1) (reference:
    modules/zipformer.py:995-1056), shape (2L-1, pos_dim) fp32."""
    x = torch.arange(-(L - 1), L, device=device, dtype=torch.float32).unsqueeze(1)
    freqs = 1 + torch.arange(pos_dim // 2, device=device)
    cl = pos_dim ** 0.5
    xc = cl * x.sign() * ((x.abs() + cl).log() - math.log(cl))
    xa = (xc / (pos_dim / (2.0 * math.pi))).atan()
    pe = torch.zeros(x.shape[0], pos_dim, device=device)
    pe[:, 0::2] = (xa * freqs).cos()
    pe[:, 1::2] = (xa * freqs).sin()
    pe[:, -1] = 1.0
    return pe


class _PosTables(list):
    """Per-layer rel-pos tables of one length (a list that can be weakly referenced)."""


class PackedZipformer:
    """Device-resident weights of one TTSZipformer plus builders for the `zvb_model` struct."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, cfg: ZipformerConfig, device,
                 stream_index: int = 0):
        self.cfg = cfg
        self.device = torch.device(device)
        self.prefix = prefix
        self._keep: List[torch.Tensor] = []
        self._sd = sd
        c = cfg
        if len(c.in_dims) == 1:
            pin, pout = prefix + "in_proj", prefix + "out_proj"
        else:
            pin, pout = prefix + f"in_proj.{stream_index}", prefix + f"out_proj.{stream_index}"
        self.in_dim = c.in_dims[stream_index]
        self.out_dim = c.out_dims[stream_index]
        self.in_proj = self._linear(pin)
        self.out_proj = self._linear(pout)
        self.time = None
        if c.time_embed_dim != -1:
            self.time = dict(w0=self._f32(prefix + "time_embed.0.weight"), b0=self._f32(prefix + "time_embed.0.bias"),
                             w2=self._f32(prefix + "time_embed.2.weight"), b2=self._f32(prefix + "time_embed.2.bias"),
                             g=self._f32(prefix + "guidance_scale_embed.weight") if c.use_guidance_scale_embed else None)
        # guidance_scale_embed maps its OWN embedding width (192 in the reference, zipformer.py:128,233-238) to
        # time_embed_dim: the width is taken from the weight, not assumed equal to time_embed_dim
        self.guidance_dim = 0
        if self.time is not None:
            te = c.time_embed_dim
            assert tuple(self.time["w0"].shape) == (2 * te, te) and tuple(self.time["w2"].shape) == (te, 2 * te), \
                "time_embed weights do not match time_embed_dim"
            if self.time["g"] is not None:
                assert self.time["g"].dim() == 2 and self.time["g"].shape[0] == te and self.time["g"].shape[1] % 2 == 0, \
                    f"guidance_scale_embed.weight {tuple(self.time['g'].shape)} does not map an even width to {te}"
                self.guidance_dim = int(self.time["g"].shape[1])
        self.stacks = []
        self.layers = []
        self.linear_pos = []            # per layer (H*4, pos_dim) fp32 on device
        for s, (ds, nl, k) in enumerate(zip(c.downsampling_factor, c.num_layers, c.cnn_kernel)):
            sp = prefix + f"encoders.{s}."
            st = dict(ds=ds, nl=nl, k=k, first=len(self.layers), dsw=[1.0, 0.0, 0.0, 0.0], comb=None,
                      tw=None, tb=None)
            if ds != 1:
                w = torch.softmax(sd[sp + "downsample.bias"].float().cpu(), dim=0).tolist()
                st["dsw"] = (w + [0.0] * 4)[:4]
                st["comb"] = self._f32(sp + "out_combiner.bypass_scale")
                sp = sp + "encoder."
            if c.time_embed_dim != -1:
                st["tw"] = self._f32(sp + "time_emb.1.weight")
                st["tb"] = self._f32(sp + "time_emb.1.bias")
            for j in range(nl):
                self.layers.append(self._layer(sp + f"layers.{j}.", k))
            self.stacks.append(st)
        self._sd = None
        # length -> tables, alive only while a plan of that length holds them (plans keep strong references)
        self._pos_cache: "weakref.WeakValueDictionary[int, _PosTables]" = weakref.WeakValueDictionary()

    # ------------------------------------------------------------------ tensor helpers
    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        t = t.contiguous().to(self.device)
        self._keep.append(t)
        return t

    def _f32(self, key: str) -> torch.Tensor:
        return self._dev(self._sd[key].detach().float())

    def _pack_w(self, w: torch.Tensor) -> torch.Tensor:
        out_f, in_f = w.shape
        kp = _ceil8(in_f)
        buf = torch.zeros(out_f, kp, dtype=torch.float16)
        buf[:, :in_f] = w.clamp(-65504.0, 65504.0).to(torch.float16)
        return self._dev(buf)

    def _mk(self, w_dev: torch.Tensor, b_dev: Optional[torch.Tensor], out_f: int, in_f: int):
        return dict(w=w_dev, b=b_dev, out=out_f, inf=in_f, kp=w_dev.shape[1], rows=w_dev.shape[0])

    def _linear(self, p: str, rows: Optional[slice] = None):
        w = self._sd[p + ".weight"].detach().float().cpu()
        b = self._sd.get(p + ".bias")
        if rows is not None:
            w = w[rows]
            b = b[rows] if b is not None else None
        bd = self._dev(b.detach().float()) if b is not None else None
        return self._mk(self._pack_w(w), bd, w.shape[0], w.shape[1])

    def _stacked(self, ps: List[str]):
        """Several nn.Linear with the same input as ONE weight: rows (and biases) concatenated in the given order."""
        w = torch.cat([self._sd[q + ".weight"].detach().float().cpu() for q in ps], dim=0)
        b = torch.cat([self._sd[q + ".bias"].detach().float().cpu() for q in ps], dim=0)
        return self._mk(self._pack_w(w), self._dev(b), w.shape[0], w.shape[1])

    def _gated(self, p: str, a: slice, b: slice):
        w = self._sd[p + ".weight"].detach().float().cpu()
        bias = self._sd[p + ".bias"].detach().float().cpu()
        wa, wb, ba, bb = w[a], w[b], bias[a], bias[b]
        n, in_f = wa.shape
        tiles = (n + 127) // 128
        W = torch.zeros(tiles * 256, in_f)
        Bv = torch.zeros(tiles * 256)
        for t in range(tiles):
            r = min(128, n - t * 128)
            W[t * 256: t * 256 + r] = wa[t * 128: t * 128 + r]
            W[t * 256 + 128: t * 256 + 128 + r] = wb[t * 128: t * 128 + r]
            Bv[t * 256: t * 256 + r] = ba[t * 128: t * 128 + r]
            Bv[t * 256 + 128: t * 256 + 128 + r] = bb[t * 128: t * 128 + r]
        return self._mk(self._pack_w(W), self._dev(Bv), n, in_f)

    def _layer(self, p: str, k: int):
        c = self.cfg
        D, nah = c.dim, c.na_hidden
        ly = dict(
            attn_in=self._linear(p + "self_attn_weights.in_proj"),
            ff_in=[self._linear(p + f"feed_forward{i}.in_proj") for i in (1, 2, 3)],
            ff_out=[self._linear(p + f"feed_forward{i}.out_proj") for i in (1, 2, 3)],
            ff1_attn=self._stacked([p + "feed_forward1.in_proj", p + "self_attn_weights.in_proj"]),
            na_sx=self._gated(p + "nonlin_attention.in_proj", slice(0, nah), slice(nah, 2 * nah)),
            na_y=self._linear(p + "nonlin_attention.in_proj", slice(2 * nah, 3 * nah)),
            na_out=self._linear(p + "nonlin_attention.out_proj"),
            sa_in=[self._linear(p + f"self_attn{i}.in_proj") for i in (1, 2)],
            sa_out=[self._linear(p + f"self_attn{i}.out_proj") for i in (1, 2)],
            conv_in=[self._gated(p + f"conv_module{i}.in_proj", slice(0, D), slice(D, 2 * D)) for i in (1, 2)],
            dw_w=[self._dev(self._sd[p + f"conv_module{i}.depthwise_conv.weight"].detach().float()
                            .reshape(D, k).t()) for i in (1, 2)],
            dw_b=[self._f32(p + f"conv_module{i}.depthwise_conv.bias") for i in (1, 2)],
            conv_out=[self._linear(p + f"conv_module{i}.out_proj") for i in (1, 2)],
            norm_bias=self._f32(p + "norm.bias"),
            norm_log_scale=self._dev(self._sd[p + "norm.log_scale"].detach().float().reshape(1)),
            bypass=self._f32(p + "bypass.bypass_scale"),
            bypass_mid=self._f32(p + "bypass_mid.bypass_scale"),
        )
        self.linear_pos.append(self._f32(p + "self_attn_weights.linear_pos.weight"))
        return ly

    # ------------------------------------------------------------------ struct builders
    def pos_tables(self, L: int) -> List[torch.Tensor]:
        """Per-layer rel-pos table in the kernel's layout (see pack_pos_table), one byte tensor per layer."""
        tabs = self._pos_cache.get(L)
        if tabs is None:
            c = self.cfg
            pe = rel_pos_embedding(L, c.pos_dim, self.device)
            tabs = _PosTables()
            for wp in self.linear_pos:
                e = (pe @ wp.t()).reshape(2 * L - 1, c.num_heads, c.pos_head_dim).permute(1, 0, 2).contiguous()
                tabs.append((pack_pos_table(e), pack_pos_table_tc(e)))
            self._pos_cache[L] = tabs
        return tabs

    @staticmethod
    def _lin_struct(d) -> _lib.zvb_linear:
        return _lib.zvb_linear(d["w"].data_ptr(), d["b"].data_ptr() if d["b"] is not None else None,
                               d["out"], d["inf"], d["kp"], d["rows"])

    def model_struct(self, T: int):
        """Returns (zvb_model, keepalive) for a plan over T frames."""
        c = self.cfg
        assert c.query_head_dim == 32 and c.pos_head_dim == 4, "kernels are built for q/k dim 32, pos dim 4"
        ptr = lambda t: t.data_ptr() if t is not None else None
        nl = len(self.layers)
        arr = (_lib.zvb_layer * nl)()
        keep = [arr]
        li = 0
        m = _lib.zvb_model()
        for s, st in enumerate(self.stacks):
            L = (T + st["ds"] - 1) // st["ds"]
            tabs = self.pos_tables(L)
            keep.append(tabs)
            for j in range(st["nl"]):
                ly, z = self.layers[li], arr[li]
                z.attn_in = self._lin_struct(ly["attn_in"])
                z.pos_table = tabs[li][0].data_ptr()
                z.pos_table_tc = tabs[li][1].data_ptr()
                for i in range(3):
                    z.ff_in[i] = self._lin_struct(ly["ff_in"][i])
                    z.ff_out[i] = self._lin_struct(ly["ff_out"][i])
                z.ff1_attn = self._lin_struct(ly["ff1_attn"])
                z.na_sx = self._lin_struct(ly["na_sx"])
                z.na_y = self._lin_struct(ly["na_y"])
                z.na_out = self._lin_struct(ly["na_out"])
                for i in range(2):
                    z.sa_in[i] = self._lin_struct(ly["sa_in"][i])
                    z.sa_out[i] = self._lin_struct(ly["sa_out"][i])
                    z.conv_in[i] = self._lin_struct(ly["conv_in"][i])
                    z.dw_w[i] = ly["dw_w"][i].data_ptr()
                    z.dw_b[i] = ly["dw_b"][i].data_ptr()
                    z.conv_out[i] = self._lin_struct(ly["conv_out"][i])
                z.norm_bias = ly["norm_bias"].data_ptr()
                z.norm_log_scale = ly["norm_log_scale"].data_ptr()
                z.bypass_scale = ly["bypass"].data_ptr()
                z.bypass_mid_scale = ly["bypass_mid"].data_ptr()
                li += 1
            zs = m.stacks[s]
            zs.downsample, zs.num_layers, zs.conv_kernel, zs.first_layer = st["ds"], st["nl"], st["k"], st["first"]
            for i in range(4):
                zs.ds_weights[i] = st["dsw"][i]
            zs.out_combiner_scale = ptr(st["comb"])
            zs.time_w, zs.time_b = ptr(st["tw"]), ptr(st["tb"])
        m.abi_version = _lib.ZVB_ABI_VERSION
        m.dim, m.num_heads, m.value_head_dim = c.dim, c.num_heads, c.value_head_dim
        m.in_dim, m.out_dim = self.in_dim, self.out_dim
        for i in range(3):
            m.ff_dims[i] = c.ff_dims[i]
        m.na_hidden = c.na_hidden
        m.time_dim = c.time_embed_dim if c.time_embed_dim != -1 else 0
        m.use_guidance_embed = 1 if c.use_guidance_scale_embed else 0
        m.guidance_dim = self.guidance_dim
        m.num_stacks, m.num_layers = len(self.stacks), nl
        m.in_proj, m.out_proj = self._lin_struct(self.in_proj), self._lin_struct(self.out_proj)
        if self.time is not None:
            m.time0_w, m.time0_b = self.time["w0"].data_ptr(), self.time["b0"].data_ptr()
            m.time2_w, m.time2_b = self.time["w2"].data_ptr(), self.time["b2"].data_ptr()
            m.guidance_w = ptr(self.time["g"])
        m.layers = C.cast(arr, C.POINTER(_lib.zvb_layer))
        return m, keep
