"""Seeded synthetic weights and inputs (no checkpoints or audio are available offline).

`synth_state_dict` produces a state_dict with exactly the keys and shapes of the reference
models (reference: SURVEY.md Appendix B; zipvoice/models/zipvoice.py:95-133,
modules/zipformer.py:179-240,370-404) from a CPU `torch.Generator`, so that the same weights
can be rebuilt bit-for-bit on any machine without the reference being present.  Magnitudes
follow the reference initialisers (nn.Linear kaiming-uniform, `ScaledLinear` initial_scale)
with two deliberate deviations that make the parity tests stronger than a fresh init would:
the attention projections are scaled up so that the softmax is peaky and the rel-pos bias
matters, and the trivially-initialised parameters (BiasNorm bias, bypass scales, downsample
bias) get non-trivial values as they have after training.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

from .config import ZipVoiceConfig, ZipformerConfig


# Gains over the reference initialisers (see module docstring): chosen so that attention scores
# have a spread of a few units (peaky but not arg-max) while the reference's own bf16-autocast
# deviation from fp32 on these weights stays below 1e-2 rel-L2 per velocity (tools/make_golden.py
# prints it), i.e. the synthetic network is as well conditioned as a trained one.
GAINS = dict(attn=2.0, pos=6.0, sa_out=3.0, ff_in=1.0, ff_out=2.0, na_in=1.0, na_out=3.0, conv_in=1.0,
             conv_out=3.0)


class _Gen:
    def __init__(self, seed: int):
        self.g = torch.Generator(device="cpu")
        self.g.manual_seed(seed)

    def uniform(self, shape, bound: float) -> torch.Tensor:
        return (torch.rand(shape, generator=self.g, dtype=torch.float32) * 2 - 1) * bound

    def normal(self, shape, std: float, mean: float = 0.0) -> torch.Tensor:
        return torch.randn(shape, generator=self.g, dtype=torch.float32) * std + mean


def _linear(sd, g: _Gen, name: str, out_f: int, in_f: int, scale: float = 1.0, bias: bool = True,
            scaled: bool = False, gain: float = 1.0):
    bound = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = g.uniform((out_f, in_f), bound) * (scale * gain)
    if bias:
        # ScaledLinear re-draws the bias in +-0.1*scale (reference: modules/scaling.py:503-507)
        sd[name + ".bias"] = g.uniform((out_f,), 0.1 * scale if scaled else bound)


def _layer(sd, g: _Gen, p: str, c: ZipformerConfig, kernel: int):
    D, H = c.dim, c.num_heads
    sd[p + "bypass.bypass_scale"] = g.uniform((D,), 0.3) + 0.6
    sd[p + "bypass_mid.bypass_scale"] = g.uniform((D,), 0.3) + 0.6
    # attention: initial_scale = query_head_dim**-0.25 (reference: zipformer.py:1108-1113);
    # the gain gives a score spread of a few units instead of ~0.05 at a fresh init.
    _linear(sd, g, p + "self_attn_weights.in_proj", c.attn_in_dim, D,
            scale=c.query_head_dim ** -0.25, scaled=True, gain=GAINS["attn"])
    _linear(sd, g, p + "self_attn_weights.linear_pos", H * c.pos_head_dim, c.pos_dim,
            scale=0.05, bias=False, gain=GAINS["pos"])
    for i in (1, 2):
        _linear(sd, g, p + f"self_attn{i}.in_proj", H * c.value_head_dim, D)
        _linear(sd, g, p + f"self_attn{i}.out_proj", D, H * c.value_head_dim, scale=0.05,
                scaled=True, gain=GAINS["sa_out"])
    for i, f in zip((1, 2, 3), c.ff_dims):
        _linear(sd, g, p + f"feed_forward{i}.in_proj", f, D, gain=GAINS["ff_in"])
        _linear(sd, g, p + f"feed_forward{i}.out_proj", D, f, scale=0.1, scaled=True,
                gain=GAINS["ff_out"])
    _linear(sd, g, p + "nonlin_attention.in_proj", 3 * c.na_hidden, D, gain=GAINS["na_in"])
    _linear(sd, g, p + "nonlin_attention.out_proj", D, c.na_hidden, scale=0.05, scaled=True,
            gain=GAINS["na_out"])
    for i in (1, 2):
        _linear(sd, g, p + f"conv_module{i}.in_proj", 2 * D, D, gain=GAINS["conv_in"])
        kb = 1.0 / math.sqrt(kernel)
        sd[p + f"conv_module{i}.depthwise_conv.weight"] = g.uniform((D, 1, kernel), kb)
        sd[p + f"conv_module{i}.depthwise_conv.bias"] = g.uniform((D,), kb)
        _linear(sd, g, p + f"conv_module{i}.out_proj", D, D, scale=0.05, scaled=True,
                gain=GAINS["conv_out"])
    sd[p + "norm.log_scale"] = g.normal((), 0.2, 0.5)
    sd[p + "norm.bias"] = g.normal((D,), 0.1)


def _zipformer(sd, g: _Gen, p: str, c: ZipformerConfig):
    if len(c.in_dims) == 1:
        _linear(sd, g, p + "in_proj", c.dim, c.in_dims[0])
        _linear(sd, g, p + "out_proj", c.out_dims[0], c.dim)
    else:  # two-stream (reference: modules/zipformer_two_stream.py:160-167)
        for i, (di, do) in enumerate(zip(c.in_dims, c.out_dims)):
            _linear(sd, g, p + f"in_proj.{i}", c.dim, di)
        for i, (di, do) in enumerate(zip(c.in_dims, c.out_dims)):
            _linear(sd, g, p + f"out_proj.{i}", do, c.dim)
    for s, (ds, nl, k) in enumerate(zip(c.downsampling_factor, c.num_layers, c.cnn_kernel)):
        sp = p + f"encoders.{s}."
        if ds != 1:
            sd[sp + "downsample.bias"] = g.normal((ds,), 0.3)
            sd[sp + "out_combiner.bypass_scale"] = g.uniform((c.dim,), 0.3) + 0.6
            sp = sp + "encoder."
        if c.time_embed_dim != -1:
            _linear(sd, g, sp + "time_emb.1", c.dim, c.time_embed_dim)
        for j in range(nl):
            _layer(sd, g, sp + f"layers.{j}.", c, k)
    if c.time_embed_dim != -1:
        te = c.time_embed_dim
        _linear(sd, g, p + "time_embed.0", te * 2, te)
        _linear(sd, g, p + "time_embed.2", te, te * 2)
    if c.use_guidance_scale_embed:
        _linear(sd, g, p + "guidance_scale_embed", c.time_embed_dim, c.time_embed_dim, scale=0.1,
                bias=False)


def synth_state_dict(cfg: ZipVoiceConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """fp32 CPU state_dict with the reference's key set for `cfg.variant`."""
    g = _Gen(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _zipformer(sd, g, "fm_decoder.", cfg.fm_decoder())
    _zipformer(sd, g, "text_encoder.", cfg.text_encoder())
    sd["embed.weight"] = g.normal((cfg.vocab_size, cfg.text_embed_dim), 1.0)
    if cfg.is_dialog:
        sd["spk_embed.weight"] = g.normal((2, cfg.feat_dim), 0.1)
    return sd


def synth_utterances(cfg: ZipVoiceConfig, batch: int, prompt_frames: int, target_frames,
                     prompt_tokens: int = 45, tokens: int = 150, seed: int = 666,
                     ragged: bool = False):
    """Synthetic sampler inputs of the shapes in SURVEY.md §8(d).

    Returns dict(tokens, prompt_tokens, prompt_features (B,Tp,F), prompt_features_lens (B,),
    features_lens (B,) total frames, x0 (B,T,F)).  `duration="real"` callers pass
    `features_lens - prompt_features_lens`... (see ZipVoice.sample).  Prompt mel follows
    N(0,1)*0.3-0.5 in feat-scaled units; noise is drawn from a CPU generator so both sides of
    a parity test consume the identical x0.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    F = cfg.feat_dim * (2 if cfg.is_stereo else 1)
    hi = cfg.vocab_size - (3 if cfg.is_dialog else 1)
    if isinstance(target_frames, int):
        if ragged:
            lo = max(1, int(target_frames * 0.64))
            tgt = torch.randint(lo, target_frames + 1, (batch,), generator=g)
            tgt[0] = target_frames
            tgt = torch.sort(tgt, descending=True).values
        else:
            tgt = torch.full((batch,), target_frames, dtype=torch.int64)
    else:
        tgt = torch.as_tensor(target_frames, dtype=torch.int64)
    toks: List[List[int]] = []
    ptoks: List[List[int]] = []
    for b in range(batch):
        t = torch.randint(1, hi, (tokens,), generator=g).tolist()
        pt = torch.randint(1, hi, (prompt_tokens,), generator=g).tolist()
        if cfg.is_dialog:
            # speaker-turn markers [S1]/[S2] (reference: zipvoice_dialog.py:118-125)
            pt[0] = cfg.spk_a_id
            for i in range(0, tokens, 25):
                t[i] = cfg.spk_b_id if (i // 25) % 2 == 0 else cfg.spk_a_id
        toks.append(t)
        ptoks.append(pt)
    if ragged:
        plen = torch.randint(max(1, int(prompt_frames * 0.7)), prompt_frames + 1, (batch,),
                             generator=g)
        plen[0] = prompt_frames
    else:
        plen = torch.full((batch,), prompt_frames, dtype=torch.int64)
    pf = torch.randn(batch, prompt_frames, F, generator=g) * 0.3 - 0.5
    pf = pf * (torch.arange(prompt_frames)[None, :, None] < plen[:, None, None])
    total = plen + tgt
    T = int(total.max())
    x0 = torch.randn(batch, T, F, generator=g)
    return dict(tokens=toks, prompt_tokens=ptoks, prompt_features=pf,
                prompt_features_lens=plen, target_lens=tgt, features_lens=total, x0=x0)
