"""ORACLE — test infrastructure only (not the product, never imported by `zipvoice_b200`).

A CPU, fp32, plain-PyTorch restatement of the reference's inference hot path
(`ZipVoice.sample` / `sample_intermediate` -> Euler solver -> CFG -> `TTSZipformer`), written
as free functions over a reference-format `state_dict`, each citing the reference file:line
it follows (paths relative to the reference repo root).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` leg may use it.

Pinning: the reference holds no golden vectors for this path, so this
file is pinned against the reference ITSELF: `tools/make_golden.py` imports the reference
from /root/reference (possible only in the build container), loads the same synthetic
weights, runs `model.solver.sample` / `model.fm_decoder` / `model.text_encoder` and stores the
outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this restatement against
those fixtures (<= 2e-5 rel-L2, i.e. fp32 re-association noise), and
`tests/test_oracle_vs_reference.py` runs the live comparison whenever /root/reference exists.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------- scaling.py ops
def swoosh_l(x: Tensor) -> Tensor:
    # modules/scaling.py:1189-1195 (fused act+linear path: log(1+exp(x-4)) with inf guard)
    xo = x - 4.0
    ls = (1.0 + xo.exp()).log()
    ls = torch.where(ls == float("inf"), xo, ls)
    return ls - 0.08 * x - 0.035


def swoosh_r(x: Tensor) -> Tensor:
    # modules/scaling.py:1200-1206 ; standalone module uses logaddexp (:1131), same value
    xo = x - 1.0
    ls = (1.0 + xo.exp()).log()
    ls = torch.where(ls == float("inf"), xo, ls)
    return ls - 0.08 * x - 0.313261687


def swoosh_r_module(x: Tensor) -> Tensor:
    # modules/scaling.py:1131 (SwooshR module inside nn.Sequential time embeddings)
    return torch.logaddexp(torch.zeros((), dtype=x.dtype), x - 1.0) - 0.08 * x - 0.313261687


def bias_norm(x: Tensor, bias: Tensor, log_scale: Tensor) -> Tensor:
    # modules/scaling.py:358-363 (no epsilon)
    scales = torch.mean((x - bias) ** 2, dim=-1, keepdim=True) ** -0.5 * log_scale.exp()
    return x * scales


def linear(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


# ----------------------------------------------------------------------------- zipformer.py
def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    # modules/zipformer.py:47-69 ; t: (N,) -> (N, dim) = [cos | sin]
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[..., None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def rel_pos_table(L: int, pos_dim: int) -> Tensor:
    # modules/zipformer.py:995-1056 ; rows <-> offsets -(L-1)..(L-1), shape (2L-1, pos_dim)
    x = torch.arange(-(L - 1), L).to(torch.float32).unsqueeze(1)
    freqs = 1 + torch.arange(pos_dim // 2)
    cl = pos_dim ** 0.5
    xc = cl * x.sign() * ((x.abs() + cl).log() - math.log(cl))
    length_scale = pos_dim / (2.0 * math.pi)
    xa = (xc / length_scale).atan()
    pe = torch.zeros(x.shape[0], pos_dim)
    pe[:, 0::2] = (xa * freqs).cos()
    pe[:, 1::2] = (xa * freqs).sin()
    pe[:, -1] = 1.0
    return pe


def attention_weights(sd: SD, p: str, x: Tensor, pos_emb: Tensor, key_padding_mask: Optional[Tensor],
                      H: int, dq: int, dp: int) -> Tensor:
    """modules/zipformer.py:1149-1306 ; x (L,N,D) -> (H,N,L,L) fp32 softmax weights."""
    L, N, _ = x.shape
    qkp = linear(sd, p + "in_proj", x)
    qd = dq * H
    q = qkp[..., :qd].reshape(L, N, H, dq).permute(2, 1, 0, 3)
    k = qkp[..., qd:2 * qd].reshape(L, N, H, dq).permute(2, 1, 3, 0)
    pp = qkp[..., 2 * qd:].reshape(L, N, H, dp).permute(2, 1, 0, 3)
    scores = torch.matmul(q, k)                                        # :1205
    pe = F.linear(pos_emb, sd[p + "linear_pos.weight"])                 # (2L-1, H*dp) :1215
    pe = pe.reshape(1, 2 * L - 1, H, dp).permute(2, 0, 3, 1)           # (H,1,dp,2L-1)
    pos = torch.matmul(pp, pe)                                         # (H,N,L,2L-1) :1224
    # skew, :1239-1248 : pos_abs[i, j] = pos_rel[i, (L-1) - i + j]
    idx = (L - 1) - torch.arange(L)[:, None] + torch.arange(L)[None, :]
    pos = torch.gather(pos, 3, idx.expand(H, N, L, L))
    scores = scores + pos
    if key_padding_mask is not None:
        scores = scores.masked_fill(key_padding_mask.unsqueeze(1), -1000)  # :1286-1289
    return scores.softmax(dim=-1)


def self_attention(sd: SD, p: str, x: Tensor, w: Tensor) -> Tensor:
    # modules/zipformer.py:1359-1396
    L, N, _ = x.shape
    H = w.shape[0]
    v = linear(sd, p + "in_proj", x).reshape(L, N, H, -1).permute(2, 1, 0, 3)
    v = torch.matmul(w, v).permute(2, 1, 0, 3).reshape(L, N, -1)
    return linear(sd, p + "out_proj", v)


def nonlin_attention(sd: SD, p: str, x: Tensor, w0: Tensor) -> Tensor:
    # modules/zipformer.py:1499-1544 ; w0 = attn_weights[0:1] -> one head over all channels
    L, N, _ = x.shape
    s, v, y = linear(sd, p + "in_proj", x).chunk(3, dim=2)
    v = v * torch.tanh(s)
    v = v.reshape(L, N, 1, -1).permute(2, 1, 0, 3)
    v = torch.matmul(w0, v).permute(2, 1, 0, 3).reshape(L, N, -1)
    return linear(sd, p + "out_proj", v * y)


def feedforward(sd: SD, p: str, x: Tensor) -> Tensor:
    # modules/zipformer.py:1433-1439 + scaling.py:1334-1349 (SwooshL then linear)
    return linear(sd, p + "out_proj", swoosh_l(linear(sd, p + "in_proj", x)))


def conv_module(sd: SD, p: str, x: Tensor, key_padding_mask: Optional[Tensor]) -> Tensor:
    # modules/zipformer.py:1638-1680
    x = linear(sd, p + "in_proj", x)
    x, s = x.chunk(2, dim=2)
    x = (x * torch.sigmoid(s)).permute(1, 2, 0)                        # (N,C,L)
    if key_padding_mask is not None:
        x = x.masked_fill(key_padding_mask.unsqueeze(1).expand_as(x), 0.0)
    w = sd[p + "depthwise_conv.weight"]
    x = F.conv1d(x, w, sd[p + "depthwise_conv.bias"], padding=w.shape[-1] // 2, groups=w.shape[0])
    x = x.permute(2, 0, 1)
    return linear(sd, p + "out_proj", swoosh_r(x))


def bypass(scale: Tensor, orig: Tensor, x: Tensor) -> Tensor:
    # modules/zipformer.py:775-776,803-804 (eval: raw parameter)
    return orig + (x - orig) * scale


def encoder_layer(sd: SD, p: str, src: Tensor, pos_emb: Tensor, time_emb: Optional[Tensor],
                  mask: Optional[Tensor], H: int, dq: int, dp: int) -> Tensor:
    """modules/zipformer.py:489-642, eval path (every balancer/whiten/dropout is identity)."""
    orig = src
    w = attention_weights(sd, p + "self_attn_weights.", src, pos_emb, mask, H, dq, dp)
    if time_emb is not None:
        src = src + time_emb
    src = src + feedforward(sd, p + "feed_forward1.", src)
    src = src + nonlin_attention(sd, p + "nonlin_attention.", src, w[0:1])
    src = src + self_attention(sd, p + "self_attn1.", src, w)
    if time_emb is not None:
        src = src + time_emb
    src = src + conv_module(sd, p + "conv_module1.", src, mask)
    src = src + feedforward(sd, p + "feed_forward2.", src)
    src = bypass(sd[p + "bypass_mid.bypass_scale"], orig, src)
    src = src + self_attention(sd, p + "self_attn2.", src, w)
    if time_emb is not None:
        src = src + time_emb
    src = src + conv_module(sd, p + "conv_module2.", src, mask)
    src = src + feedforward(sd, p + "feed_forward3.", src)
    src = bias_norm(src, sd[p + "norm.bias"], sd[p + "norm.log_scale"])
    return bypass(sd[p + "bypass.bypass_scale"], orig, src)


def encoder_stack(sd: SD, p: str, src: Tensor, time_emb: Optional[Tensor], mask: Optional[Tensor],
                  n_layers: int, c) -> Tensor:
    # modules/zipformer.py:702-744
    pos_emb = rel_pos_table(src.shape[0], c["pos_dim"])
    if time_emb is not None:
        time_emb = F.linear(swoosh_r_module(time_emb), sd[p + "time_emb.1.weight"],
                            sd[p + "time_emb.1.bias"])                 # :676-680,727-729
    for j in range(n_layers):
        src = encoder_layer(sd, p + f"layers.{j}.", src, pos_emb, time_emb, mask,
                            c["num_heads"], c["query_head_dim"], c["pos_head_dim"])
    return src


def downsample(bias: Tensor, src: Tensor, ds: int) -> Tensor:
    # modules/zipformer.py:887-913 (pads with the last row of the batch-padded tensor)
    L, N, C = src.shape
    d = (L + ds - 1) // ds
    pad = d * ds - L
    src = torch.cat((src, src[L - 1:].expand(pad, N, C)), dim=0).reshape(d, ds, N, C)
    wts = bias.softmax(dim=0)[:, None, None]
    return (src * wts).sum(dim=1)


def upsample(src: Tensor, ds: int) -> Tensor:
    # modules/zipformer.py:925-935
    L, N, C = src.shape
    return src.unsqueeze(1).expand(L, ds, N, C).reshape(L * ds, N, C)


def tts_zipformer(sd: SD, p: str, c: dict, x: Tensor, t: Optional[Tensor], padding_mask: Optional[Tensor],
                  guidance_scale: Optional[Tensor] = None) -> Tensor:
    """modules/zipformer.py:242-293 and zipformer_two_stream.py:219-264.

    `c` = dict(dim, downsampling_factor, num_layers, num_heads, query_head_dim, pos_head_dim,
    pos_dim, time_embed_dim, in_dims, use_guidance_scale_embed).  x (N,T,Cin) -> (N,T,Cout).
    """
    if len(c["in_dims"]) == 2:  # two-stream: projection chosen by the input width
        idx = 0 if x.size(2) == c["in_dims"][0] else 1
        pin, pout = p + f"in_proj.{idx}", p + f"out_proj.{idx}"
    else:
        pin, pout = p + "in_proj", p + "out_proj"
    x = linear(sd, pin, x.permute(1, 0, 2))
    time_emb = None
    if t is not None:
        time_emb = timestep_embedding(t, c["time_embed_dim"])
        if guidance_scale is not None:
            time_emb = time_emb + F.linear(timestep_embedding(guidance_scale, c["time_embed_dim"]),
                                           sd[p + "guidance_scale_embed.weight"])
        time_emb = linear(sd, p + "time_embed.2",
                          swoosh_r_module(linear(sd, p + "time_embed.0", time_emb)))
    for s, (ds, nl) in enumerate(zip(c["downsampling_factor"], c["num_layers"])):
        sp = p + f"encoders.{s}."
        if ds == 1:
            x = encoder_stack(sd, sp, x, time_emb, padding_mask, nl, c)
        else:  # modules/zipformer.py:850-870
            orig = x
            y = downsample(sd[sp + "downsample.bias"], x, ds)
            m = padding_mask[..., ::ds] if padding_mask is not None else None
            y = encoder_stack(sd, sp + "encoder.", y, time_emb, m, nl, c)
            y = upsample(y, ds)[: orig.shape[0]]
            x = bypass(sd[sp + "out_combiner.bypass_scale"], orig, y)
    return linear(sd, pout, x).permute(1, 0, 2)


def zipformer_cfg_dict(zc) -> dict:
    """Accepts a zipvoice_b200.config.ZipformerConfig-like object (duck-typed)."""
    return dict(dim=zc.dim, downsampling_factor=tuple(zc.downsampling_factor),
                num_layers=tuple(zc.num_layers), num_heads=zc.num_heads,
                query_head_dim=zc.query_head_dim, pos_head_dim=zc.pos_head_dim, pos_dim=zc.pos_dim,
                time_embed_dim=zc.time_embed_dim, in_dims=tuple(zc.in_dims),
                use_guidance_scale_embed=zc.use_guidance_scale_embed)


# ----------------------------------------------------------------------------- zipvoice.py / solver.py
def forward_fm_decoder(sd: SD, c: dict, t: Tensor, xt: Tensor, text_condition: Tensor,
                       speech_condition: Tensor, padding_mask: Tensor,
                       guidance_scale: Optional[Tensor] = None) -> Tensor:
    # models/zipvoice.py:135-185
    xt = torch.cat([xt, text_condition, speech_condition], dim=2)
    while t.dim() > 1 and t.size(-1) == 1:
        t = t.squeeze(-1)
    if t.dim() == 0:
        t = t.repeat(xt.shape[0])
    if guidance_scale is not None:
        while guidance_scale.dim() > 1 and guidance_scale.size(-1) == 1:
            guidance_scale = guidance_scale.squeeze(-1)
        if guidance_scale.dim() == 0:
            guidance_scale = guidance_scale.repeat(xt.shape[0])
    return tts_zipformer(sd, "fm_decoder.", c, xt, t, padding_mask, guidance_scale)


def cfg_velocity(sd: SD, c: dict, t: Tensor, x: Tensor, text_condition: Tensor, speech_condition: Tensor,
                 padding_mask: Tensor, guidance_scale: Union[float, Tensor], distill: bool) -> Tensor:
    # modules/solver.py:40-110 (DiffusionModel) and :113-165 (DistillDiffusionModel)
    if not torch.is_tensor(guidance_scale):
        guidance_scale = torch.tensor(guidance_scale, dtype=t.dtype)
    if distill:
        return forward_fm_decoder(sd, c, t, x, text_condition, speech_condition, padding_mask,
                                  guidance_scale)
    if (guidance_scale == 0.0).all():
        return forward_fm_decoder(sd, c, t, x, text_condition, speech_condition, padding_mask)
    assert t.dim() == 0
    x2 = torch.cat([x] * 2, dim=0)
    m2 = torch.cat([padding_mask] * 2, dim=0)
    text2 = torch.cat([torch.zeros_like(text_condition), text_condition], dim=0)
    if t > 0.5:
        sp2 = torch.cat([torch.zeros_like(speech_condition), speech_condition], dim=0)
    else:
        guidance_scale = guidance_scale * 2
        sp2 = torch.cat([speech_condition, speech_condition], dim=0)
    u, cnd = forward_fm_decoder(sd, c, t, x2, text2, sp2, m2).chunk(2, dim=0)
    return (1 + guidance_scale) * cnd - guidance_scale * u


def get_time_steps(t_start: float, t_end: float, num_step: int, t_shift: float) -> Tensor:
    # modules/solver.py:256-281
    ts = torch.linspace(t_start, t_end, num_step + 1)
    return t_shift * ts / (1 + (t_shift - 1) * ts)


def euler_sample(sd: SD, c: dict, x: Tensor, text_condition: Tensor, speech_condition: Tensor,
                 padding_mask: Tensor, num_step: int = 10, guidance_scale: Union[float, Tensor] = 0.0,
                 t_start: float = 0.0, t_end: float = 1.0, t_shift: float = 1.0,
                 distill: bool = False, record: Optional[List[Tensor]] = None) -> Tensor:
    # modules/solver.py:182-240 ; `record`, if given, receives every step's velocity
    ts = get_time_steps(t_start, t_end, num_step, t_shift)
    for i in range(num_step):
        v = cfg_velocity(sd, c, ts[i], x, text_condition, speech_condition, padding_mask,
                         guidance_scale, distill)
        if record is not None:
            record.append(v)
        x = x + v * (ts[i + 1] - ts[i])
    return x


# ----------------------------------------------------------------------------- host prelude
def pad_labels(y: Sequence[Sequence[int]], pad_id: int) -> Tensor:
    # utils/common.py:261-274 (one pad appended to every sequence, then pad to max)
    y = [list(t) + [pad_id] for t in y]
    n = max(len(t) for t in y)
    return torch.tensor([t + [pad_id] * (n - len(t)) for t in y], dtype=torch.int64)


def make_pad_mask(lengths: Tensor, max_len: int = 0) -> Tensor:
    # utils/common.py:401-426
    max_len = max(max_len, int(lengths.max()))
    return torch.arange(max_len)[None, :].expand(lengths.size(0), max_len) >= lengths[:, None]


def tokens_index(features_lens: Tensor, tokens_lens: Tensor, num_frames: int) -> Tensor:
    # utils/common.py:252-258 + :277-301 (remaining frames point at the appended pad token)
    B = len(features_lens)
    ans = torch.zeros(B, num_frames, dtype=torch.int64)
    for b in range(B):
        n_tok = int(tokens_lens[b])
        d = int(features_lens[b]) // n_tok
        durs = [d] * n_tok
        durs.append(num_frames - sum(durs))
        cur = 0
        for i, dd in enumerate(durs):
            ans[b, cur:cur + dd] = i
            cur += dd
        assert cur == num_frames
    return ans


def forward_text_embed(sd: SD, tc: dict, tokens: Sequence[Sequence[int]], pad_id: int,
                       dialog: Optional[Tuple[int, int]] = None) -> Tuple[Tensor, Tensor]:
    # models/zipvoice.py:187-212 ; dialog override models/zipvoice_dialog.py:118-159
    padded = pad_labels(tokens, pad_id)
    embed = F.embedding(padded, sd["embed.weight"])
    lens = torch.tensor([len(t) for t in tokens], dtype=torch.int64)
    mask = make_pad_mask(lens, embed.shape[1])
    embed = tts_zipformer(sd, "text_encoder.", tc, embed, None, mask)
    if dialog is not None:
        a, b = dialog
        turn = ((padded == a) | (padded == b)).long().cumsum(dim=1) % 2
        turn = torch.where(padded == pad_id, -1, turn)
        embed = embed.clone()
        embed[turn == 0] += sd["spk_embed.weight"][0]
        embed[turn == 1] += sd["spk_embed.weight"][1]
    return embed, lens


def forward_text_condition(embed: Tensor, tokens_lens: Tensor, features_lens: Tensor):
    # models/zipvoice.py:214-251
    T = int(features_lens.max())
    mask = make_pad_mask(features_lens, T)
    idx = tokens_index(features_lens, tokens_lens, T)
    cond = torch.gather(embed, 1, idx.unsqueeze(-1).expand(embed.size(0), T, embed.size(-1)))
    return cond, mask


def predict_features_lens(prompt_features_lens: Tensor, prompt_tokens_lens: Tensor,
                          tokens_lens: Tensor, speed: float) -> Tensor:
    # models/zipvoice.py:323-325 (float32 arithmetic, ceil, int64)
    return prompt_features_lens + torch.ceil(
        prompt_features_lens / prompt_tokens_lens * tokens_lens / speed).to(torch.int64)


class OracleModel:
    """Bundles a state_dict + config; mirrors `ZipVoice.sample` / `sample_intermediate`."""

    def __init__(self, cfg, sd: SD):
        self.cfg = cfg
        self.sd = {k: v.detach().to(torch.float32).cpu() for k, v in sd.items()}
        self.fc = zipformer_cfg_dict(cfg.fm_decoder())
        self.tc = zipformer_cfg_dict(cfg.text_encoder())
        self.dialog = (cfg.spk_a_id, cfg.spk_b_id) if cfg.is_dialog else None

    def text_embed(self, tokens):
        return forward_text_embed(self.sd, self.tc, tokens, self.cfg.pad_id, self.dialog)

    def prelude(self, tokens, prompt_tokens, prompt_features, prompt_features_lens,
                features_lens=None, speed=1.0, duration="predict"):
        # models/zipvoice.py:419-451
        cat = [list(p) + list(t) for p, t in zip(prompt_tokens, tokens)]
        embed, cat_lens = self.text_embed(cat)
        if duration == "predict":
            pl = torch.tensor([len(t) for t in prompt_tokens], dtype=torch.int64)
            tl = torch.tensor([len(t) for t in tokens], dtype=torch.int64)
            fl = predict_features_lens(prompt_features_lens, pl, tl, speed)
        else:
            fl = prompt_features_lens + features_lens
        text_condition, padding_mask = forward_text_condition(embed, cat_lens, fl)
        T = text_condition.shape[1]
        speech = F.pad(prompt_features, (0, 0, 0, T - prompt_features.size(1)))
        sc_mask = make_pad_mask(prompt_features_lens, T)
        speech = torch.where(sc_mask.unsqueeze(-1), torch.zeros_like(speech), speech)
        return text_condition, speech, padding_mask

    def solve(self, x0, text_condition, speech_condition, padding_mask, record=None, **kw):
        return euler_sample(self.sd, self.fc, x0, text_condition, speech_condition, padding_mask,
                            distill=self.cfg.is_distill, record=record, **kw)

    def split(self, x1, padding_mask, prompt_features_lens):
        # models/zipvoice.py:469-486
        lens = (~padding_mask).sum(-1) - prompt_features_lens
        out = torch.zeros(x1.size(0), int(lens.max()), x1.size(2))
        pr = torch.zeros(x1.size(0), int(prompt_features_lens.max()), x1.size(2))
        for i in range(x1.size(0)):
            pl, gl = int(prompt_features_lens[i]), int(lens[i])
            out[i, :gl] = x1[i, pl:pl + gl]
            pr[i, :pl] = x1[i, :pl]
        return out, lens, pr, prompt_features_lens

    def sample(self, tokens, prompt_tokens, prompt_features, prompt_features_lens,
               features_lens=None, speed=1.0, t_shift=1.0, duration="predict", num_step=5,
               guidance_scale=0.5, x0=None, record=None):
        """models/zipvoice.py:388-486; `x0` replaces the device RNG draw at :453."""
        tc, sc, pm = self.prelude(tokens, prompt_tokens, prompt_features, prompt_features_lens,
                                  features_lens, speed, duration)
        if x0 is None:
            x0 = torch.randn(tc.shape[0], tc.shape[1], prompt_features.size(-1))
        x1 = self.solve(x0[:, : tc.shape[1]], tc, sc, pm, record=record, num_step=num_step,
                        guidance_scale=guidance_scale, t_shift=t_shift)
        return self.split(x1, pm, prompt_features_lens)

    def sample_intermediate(self, tokens, features, features_lens, noise, speech_condition_mask,
                            t_start, t_end, num_step=1, guidance_scale=None, record=None):
        # models/zipvoice.py:488-534
        embed, lens = self.text_embed(tokens)
        tc, pm = forward_text_condition(embed, lens, features_lens)
        sc = torch.where(speech_condition_mask.unsqueeze(-1), torch.zeros(()), features)
        x = self.solve(noise, tc, sc, pm, record=record, num_step=num_step,
                       guidance_scale=guidance_scale, t_start=t_start, t_end=t_end)
        return x, (~pm).sum(-1)
