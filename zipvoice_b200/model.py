"""Host-side mirror of the reference model API for the sampler hot path.

`ZipVoice`, `ZipVoiceDistill`, `ZipVoiceDialog`, `ZipVoiceDialogStereo` keep the reference's
constructor kwargs, `state_dict` keys and method names/signatures (reference:
zipvoice/models/zipvoice.py:35-534, zipvoice_distill.py:27-94, zipvoice_dialog.py:29-256) so that
`zipvoice.bin.infer_zipvoice` style callers can use them unchanged:

    model.sample(tokens, prompt_tokens, prompt_features, prompt_features_lens, ...)   -> 4-tuple
    model.sample_intermediate(...)                                                    -> (x, lens)
    model.fm_decoder(x=, t=, padding_mask=, guidance_scale=)     seam 1 (tensorrt.py:69-126)
    model.solver.sample(x=, text_condition=, ...)                seam 2 (solver.py:182-240)

All arithmetic of the Euler loop, CFG and both Zipformers runs in the sm_100a kernels behind the
C ABI; this file only does list/length bookkeeping (the reference's own host prelude,
zipvoice.py:187-330, utils/common.py:252-301, restated without Python loops over frames).
`accelerate(ref_model)` patches an instance of the *reference* classes in place instead.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Tuple, Union

import torch

from . import _lib
from .config import ZipVoiceConfig
from .engine import DecoderPlan, PlanCache
from .weights import PackedZipformer


# --------------------------------------------------------------------------------- host helpers
def pad_labels(y: List[List[int]], pad_id: int, device) -> torch.Tensor:
    """One pad appended to every sequence, then padded to the max (reference: common.py:261-274).
    One flat host tensor and one masked scatter instead of a Python loop over utterances."""
    lens = torch.tensor([len(t) for t in y], dtype=torch.int64)
    n = int(lens.max()) + 1 if len(y) else 1
    out = torch.full((len(y), n), pad_id, dtype=torch.int64)
    flat = torch.tensor(list(itertools.chain.from_iterable(y)), dtype=torch.int64)
    out[torch.arange(n)[None, :] < lens[:, None]] = flat
    return out.to(device)


def make_pad_mask(lengths: torch.Tensor, max_len: int = 0) -> torch.Tensor:
    """True at padded positions (reference: common.py:401-426)."""
    max_len = max(int(max_len), int(lengths.max()))
    return torch.arange(max_len, device=lengths.device)[None, :] >= lengths[:, None]


def tokens_index(features_lens: torch.Tensor, tokens_lens: torch.Tensor, num_frames: int) -> torch.Tensor:
    """Frame -> token position: each real token gets features_len // tokens_len frames, every
    remaining frame (and the batch padding) points at the appended pad token, index tokens_len
    (reference: common.py:252-258, 277-301).  Vectorised."""
    d = torch.div(features_lens, tokens_lens, rounding_mode="floor").clamp_min(1)
    frames = torch.arange(num_frames, device=features_lens.device)[None, :]
    idx = torch.div(frames, d[:, None], rounding_mode="floor")
    zero_d = (torch.div(features_lens, tokens_lens, rounding_mode="floor") == 0)[:, None]
    idx = torch.where(zero_d, tokens_lens[:, None].expand_as(idx), idx)
    return torch.minimum(idx, tokens_lens[:, None])


def get_time_steps(t_start: float, t_end: float, num_step: int, t_shift: float) -> torch.Tensor:
    """CPU fp32 time grid (reference: modules/solver.py:256-281)."""
    ts = torch.linspace(t_start, t_end, num_step + 1)
    return t_shift * ts / (1 + (t_shift - 1) * ts)


# --------------------------------------------------------------------------------- seam objects
class B200Zipformer:
    """Callable replacement of a `TTSZipformer` (seam 1).  Shapes are served by a bucketed LRU plan cache;
    when the plan is larger than the call, inputs are zero padded and the extra frames / rows masked."""

    def __init__(self, packed: PackedZipformer, frame_bucket: int = 0, row_bucket: int = 0):
        self.packed = packed
        self.plans = PlanCache(packed, frame_bucket=frame_bucket, row_bucket=row_bucket)

    def __call__(self, x: torch.Tensor, t: Optional[torch.Tensor] = None, padding_mask: Optional[torch.Tensor] = None,
                 guidance_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
        N, T, _ = x.shape
        if padding_mask is None:
            padding_mask = torch.zeros(N, T, dtype=torch.bool, device=x.device)
        if t is not None and t.dim() != 1:
            raise NotImplementedError("per-frame t (N,T) is a training-only input of the reference")
        plan = self.plans.get(N, T)
        if (plan.N, plan.T) == (N, T):
            return plan.forward_f32(x, t, padding_mask, guidance_scale)
        xp = torch.zeros(plan.N, plan.T, x.shape[2], dtype=torch.float32, device=x.device)
        xp[:N, :T] = x
        mp = torch.ones(plan.N, plan.T, dtype=torch.bool, device=x.device)
        mp[:N, :T] = padding_mask
        pad1 = lambda v: None if v is None else torch.nn.functional.pad(v.float(), (0, plan.N - N))
        return plan.forward_f32(xp, pad1(t), mp, pad1(guidance_scale))[:N, :T].contiguous()


class B200EulerSolver:
    """`model.solver` replacement (seam 2): Euler ODE + classifier-free guidance + decoder in one
    launch sequence, optionally replayed from a CUDA graph (reference: modules/solver.py:167-240).
    Graphs and their static input buffers are owned by the plan they were captured over."""

    def __init__(self, decoders: Dict[int, B200Zipformer], distill: bool, use_cuda_graph: bool = True):
        self.decoders = decoders          # feature width F -> decoder (stereo models have two)
        self.distill = distill
        self.use_cuda_graph = use_cuda_graph
        self.last_velocities: Optional[torch.Tensor] = None
        self.record_velocities = False

    @property
    def check_saturation(self) -> bool:
        return all(d.plans.check_saturation for d in self.decoders.values())

    @check_saturation.setter
    def check_saturation(self, on: bool) -> None:
        for d in self.decoders.values():
            d.plans.check_saturation = bool(on)
            if on:
                for p in d.plans.plans():
                    p.enable_saturation_check()

    def count_saturated(self) -> int:
        """fp16 values at the saturation limit seen so far by the decoder plans (0 unless check_saturation)."""
        return sum(p.saturated() for d in self.decoders.values() for p in d.plans.plans())

    def _mode_and_guidance(self, guidance_scale, B: int, device):
        if torch.is_tensor(guidance_scale):
            g = guidance_scale.detach().to(device=device, dtype=torch.float32).reshape(-1)
            if g.numel() == 1:
                g = g.expand(B)
            g = g.contiguous()
            all_zero = bool((g == 0).all())              # same host sync as solver.py:71
        else:
            g = torch.full((B,), float(guidance_scale), dtype=torch.float32, device=device)
            all_zero = float(guidance_scale) == 0.0
        if self.distill:
            return 2, g
        return (0, None) if all_zero else (1, g)

    def sample(self, x: torch.Tensor, text_condition: torch.Tensor, speech_condition: torch.Tensor,
               padding_mask: torch.Tensor, num_step: int = 10, guidance_scale: Union[float, torch.Tensor] = 0.0,
               t_start: float = 0.0, t_end: float = 1.0, t_shift: float = 1.0, **kwargs) -> torch.Tensor:
        assert isinstance(t_start, float) and isinstance(t_end, float)
        device = x.device
        B, T, F = x.shape
        mode, g = self._mode_and_guidance(guidance_scale, B, device)
        dec = self.decoders[F]
        Bp, Tp = dec.plans.shape_for(B, T)
        plan = dec.plans.get(2 * Bp if mode == 1 else Bp, Tp)
        if mode == 1 and plan.N != 2 * Bp:               # odd row bucket: keep the CFG halves equal
            Bp = plan.N // 2
        ts_host = get_time_steps(t_start, t_end, num_step, t_shift).contiguous()
        record = self.record_velocities
        key = (Bp, num_step, mode, tuple(bool(v > 0.5) for v in ts_host[:-1].tolist()), record)
        st = plan.graphs.get(key) if self.use_cuda_graph else None
        if st is None:
            st = dict(
                x=torch.zeros(Bp, Tp, F, dtype=torch.float32, device=device),
                text=torch.zeros(Bp, Tp, text_condition.shape[2], dtype=torch.float32, device=device),
                speech=torch.zeros(Bp, Tp, F, dtype=torch.float32, device=device),
                mask=torch.ones(Bp, Tp, dtype=torch.uint8, device=device),
                g=torch.zeros(Bp, dtype=torch.float32, device=device),
                ts=torch.empty(num_step + 1, dtype=torch.float32, device=device),
                vrec=torch.empty(num_step, Bp, Tp, F, dtype=torch.float32, device=device) if record else None,
                graph=None)
            st["nbytes"] = sum(v.numel() * v.element_size() for v in st.values() if torch.is_tensor(v))
            if self.use_cuda_graph:
                plan.graphs[key] = st
        if (Bp, Tp) != (B, T):
            # the ODE state is updated in place and an earlier, longer call may have left its conditions behind:
            # the padding is cleared on every call so that a result never depends on the call history
            st["x"].zero_()
            st["text"].zero_()
            st["speech"].zero_()
            st["mask"].fill_(1)
            st["g"].zero_()
        st["x"][:B, :T].copy_(x)
        st["text"][:B, :T].copy_(text_condition)
        st["speech"][:B, :T].copy_(speech_condition)
        st["mask"][:B, :T].copy_(padding_mask)
        if g is not None:
            st["g"][:B].copy_(g)
        # pageable source: the copy is staged before it returns, so `ts_host` may be reused at once and an
        # in-flight earlier sample never sees this call's time grid (stream order)
        st["ts"].copy_(ts_host)

        def run():
            plan.sample(st["x"], st["text"], st["speech"], st["mask"], st["g"] if mode != 0 else None, st["ts"],
                        ts_host, num_step, mode, st["vrec"])

        if not self.use_cuda_graph:
            run()
        elif st["graph"] is None:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run()
            st["graph"] = graph
            graph.replay()
        else:
            st["graph"].replay()
        if record:
            self.last_velocities = st["vrec"][:, :B, :T].clone()
        return st["x"][:B, :T].clone()


# --------------------------------------------------------------------------------- models
# The text encoder has no down-sampling stack: padded tokens are masked keys and zeroed convolution inputs, so
# running it on a longer, masked token axis leaves every valid position bit-identical.
TEXT_FRAME_BUCKET = 32


class ZipVoice:
    """The ZipVoice model on B200 (reference: zipvoice/models/zipvoice.py:35-133)."""

    variant = "zipvoice"

    def __init__(self, use_cuda_graph: bool = True, frame_bucket: int = 0, row_bucket: int = 0, **kwargs):
        """`frame_bucket` / `row_bucket`: round the decoder's (rows, frames) up to multiples of these before
        choosing a plan / CUDA graph (0 = exact shapes; serving with ragged traffic wants e.g. 64 / 8, see
        engine.PlanCache).  The text encoder always buckets its token axis by 32 (no effect on its result)."""
        self.cfg = ZipVoiceConfig(variant=self.variant, **kwargs)
        self.frame_bucket, self.row_bucket = int(frame_bucket), int(row_bucket)
        self.feat_dim = self.cfg.feat_dim
        self.text_embed_dim = self.cfg.text_embed_dim
        self.pad_id = self.cfg.pad_id
        self.use_cuda_graph = use_cuda_graph
        self.device = torch.device("cpu")
        self._sd: Optional[Dict[str, torch.Tensor]] = None
        self.fm_decoder = None
        self.text_encoder = None
        self.solver = None
        self.embed_weight = None
        self.spk_embed_weight = None
        self.training = False

    # nn.Module-like surface used by the reference callers (infer_zipvoice.py:811-827)
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
        sd, cfg, dev = self._sd, self.cfg, self.device
        fc = cfg.fm_decoder()
        decs = {}
        for i in range(len(fc.in_dims)):
            dz = B200Zipformer(PackedZipformer(sd, "fm_decoder.", fc, dev, stream_index=i),
                               frame_bucket=self.frame_bucket, row_bucket=self.row_bucket)
            decs[fc.out_dims[i]] = dz
        self._decoders = decs
        self.fm_decoder = _WidthDispatch(decs, fc) if len(decs) > 1 else next(iter(decs.values()))
        self.text_encoder = B200Zipformer(PackedZipformer(sd, "text_encoder.", cfg.text_encoder(), dev),
                                          frame_bucket=TEXT_FRAME_BUCKET, row_bucket=self.row_bucket)
        self.solver = B200EulerSolver(decs, distill=cfg.is_distill, use_cuda_graph=self.use_cuda_graph)
        self.embed_weight = sd["embed.weight"].float().to(dev)
        if cfg.is_dialog:
            self.spk_embed_weight = sd["spk_embed.weight"].float().to(dev)

    def _need(self):
        if self.solver is None:
            raise _lib.ZvbError("model is not on a CUDA device: call load_state_dict(...) and .to('cuda') "
                                "(zipvoice_b200 has no CPU path)")

    # ------------------------------------------------------------------ text side
    def forward_text_embed(self, tokens: List[List[int]]):
        """reference: zipvoice.py:187-212 (dialog: zipvoice_dialog.py:127-159)"""
        self._need()
        dev = self.device
        padded = pad_labels(tokens, self.pad_id, dev)
        embed = torch.nn.functional.embedding(padded, self.embed_weight)
        tokens_lens = torch.tensor([len(t) for t in tokens], dtype=torch.int64, device=dev)
        mask = make_pad_mask(tokens_lens, embed.shape[1])
        embed = self.text_encoder(x=embed, t=None, padding_mask=mask)
        if self.cfg.is_dialog:
            turn = ((padded == self.cfg.spk_a_id) | (padded == self.cfg.spk_b_id)).long().cumsum(dim=1) % 2
            turn = torch.where(padded == self.pad_id, -1, turn)
            embed = embed + (turn == 0).unsqueeze(-1) * self.spk_embed_weight[0] \
                + (turn == 1).unsqueeze(-1) * self.spk_embed_weight[1]
        return embed, tokens_lens

    def forward_text_condition(self, embed, tokens_lens, features_lens):
        """reference: zipvoice.py:214-251"""
        num_frames = int(features_lens.max())
        padding_mask = make_pad_mask(features_lens, num_frames)
        idx = tokens_index(features_lens, tokens_lens, num_frames)
        text_condition = torch.gather(embed, 1, idx.unsqueeze(-1).expand(embed.size(0), num_frames, embed.size(-1)))
        return text_condition, padding_mask

    def forward_text_train(self, tokens, features_lens):
        embed, tokens_lens = self.forward_text_embed(tokens)
        return self.forward_text_condition(embed, tokens_lens, features_lens.to(self.device))

    def forward_text_inference_gt_duration(self, tokens, features_lens, prompt_tokens, prompt_features_lens):
        tokens = [p + t for p, t in zip(prompt_tokens, tokens)]
        features_lens = prompt_features_lens.to(self.device) + features_lens.to(self.device)
        embed, tokens_lens = self.forward_text_embed(tokens)
        return self.forward_text_condition(embed, tokens_lens, features_lens)

    def forward_text_inference_ratio_duration(self, tokens, prompt_tokens, prompt_features_lens, speed):
        """reference: zipvoice.py:290-330"""
        dev = self.device
        cat_tokens = [p + t for p, t in zip(prompt_tokens, tokens)]
        pl = torch.tensor([len(t) for t in prompt_tokens], dtype=torch.int64, device=dev)
        tl = torch.tensor([len(t) for t in tokens], dtype=torch.int64, device=dev)
        embed, cat_lens = self.forward_text_embed(cat_tokens)
        pfl = prompt_features_lens.to(dev)
        features_lens = pfl + torch.ceil(pfl / pl * tl / speed).to(dtype=torch.int64)
        return self.forward_text_condition(embed, cat_lens, features_lens)

    # ------------------------------------------------------------------ decoder side
    def forward_fm_decoder(self, t, xt, text_condition, speech_condition, padding_mask=None, guidance_scale=None):
        """reference: zipvoice.py:135-185"""
        self._need()
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
            return self.fm_decoder(x=xt, t=t, padding_mask=padding_mask, guidance_scale=guidance_scale)
        return self.fm_decoder(x=xt, t=t, padding_mask=padding_mask)

    @torch.inference_mode()
    def sample(self, tokens, prompt_tokens, prompt_features, prompt_features_lens, features_lens=None,
               speed: float = 1.0, t_shift: float = 1.0, duration: str = "predict", num_step: int = 5,
               guidance_scale: float = 0.5, x0: Optional[torch.Tensor] = None):
        """reference: zipvoice.py:388-486.  `x0` (optional, not in the reference signature) injects
        the initial noise instead of drawing it with the device RNG (zipvoice.py:453)."""
        self._need()
        assert duration in ["real", "predict"]
        dev = self.device
        prompt_features = prompt_features.to(dev)
        prompt_features_lens = prompt_features_lens.to(dev)
        if duration == "predict":
            text_condition, padding_mask = self.forward_text_inference_ratio_duration(
                tokens=tokens, prompt_tokens=prompt_tokens, prompt_features_lens=prompt_features_lens, speed=speed)
        else:
            assert features_lens is not None
            text_condition, padding_mask = self.forward_text_inference_gt_duration(
                tokens=tokens, features_lens=features_lens, prompt_tokens=prompt_tokens,
                prompt_features_lens=prompt_features_lens)
        batch_size, num_frames, _ = text_condition.shape
        speech_condition = torch.nn.functional.pad(prompt_features, (0, 0, 0, num_frames - prompt_features.size(1)))
        sc_mask = make_pad_mask(prompt_features_lens, num_frames)
        speech_condition = torch.where(sc_mask.unsqueeze(-1), torch.zeros_like(speech_condition), speech_condition)
        if x0 is None:
            x0 = torch.randn(batch_size, num_frames, prompt_features.size(-1), device=dev)
        else:
            x0 = x0.to(dev)[:, :num_frames]
        x1 = self.solver.sample(x=x0, text_condition=text_condition, speech_condition=speech_condition,
                                padding_mask=padding_mask, num_step=num_step, guidance_scale=guidance_scale,
                                t_shift=t_shift)
        # split prompt / generated part (reference: zipvoice.py:469-486), vectorised
        lens = (~padding_mask).sum(-1) - prompt_features_lens
        max_p, max_g = int(prompt_features_lens.max()), int(lens.max())
        fr = torch.arange(max(max_p, max_g), device=dev)
        gi = (prompt_features_lens[:, None] + fr[None, :max_g]).clamp_max(num_frames - 1)
        x1_wo_prompt = torch.gather(x1, 1, gi.unsqueeze(-1).expand(-1, -1, x1.size(2)))
        x1_wo_prompt = x1_wo_prompt * (fr[None, :max_g] < lens[:, None]).unsqueeze(-1)
        x1_prompt = x1[:, :max_p] * (fr[None, :max_p] < prompt_features_lens[:, None]).unsqueeze(-1)
        return x1_wo_prompt, lens, x1_prompt, prompt_features_lens

    @torch.inference_mode()
    def sample_intermediate(self, tokens, features, features_lens, noise, speech_condition_mask, t_start: float,
                            t_end: float, num_step: int = 1, guidance_scale: torch.Tensor = None):
        """reference: zipvoice.py:488-534"""
        self._need()
        dev = self.device
        text_condition, padding_mask = self.forward_text_train(tokens=tokens, features_lens=features_lens)
        features = features.to(dev)
        speech_condition = torch.where(speech_condition_mask.to(dev).unsqueeze(-1), torch.zeros((), device=dev), features)
        x = self.solver.sample(x=noise.to(dev), text_condition=text_condition, speech_condition=speech_condition,
                               padding_mask=padding_mask, num_step=num_step, guidance_scale=guidance_scale,
                               t_start=t_start, t_end=t_end)
        return x, (~padding_mask).sum(-1)


class _WidthDispatch:
    """Two-stream decoder: projection pair chosen by input width (reference:
    modules/zipformer_two_stream.py:236-262)."""

    def __init__(self, decs: Dict[int, B200Zipformer], fc):
        self.by_in = {fc.in_dims[i]: decs[fc.out_dims[i]] for i in range(len(fc.in_dims))}

    def __call__(self, x, t=None, padding_mask=None, guidance_scale=None):
        assert x.size(2) in self.by_in, f"{x.size(2)} in {tuple(self.by_in)}"
        return self.by_in[x.size(2)](x=x, t=t, padding_mask=padding_mask, guidance_scale=guidance_scale)


class ZipVoiceDistill(ZipVoice):
    variant = "zipvoice_distill"


class ZipVoiceDialog(ZipVoice):
    variant = "zipvoice_dialog"


class ZipVoiceDialogStereo(ZipVoice):
    variant = "zipvoice_dialog_stereo"


MODEL_CLASSES = {c.variant: c for c in (ZipVoice, ZipVoiceDistill, ZipVoiceDialog, ZipVoiceDialogStereo)}


def build_model(cfg: ZipVoiceConfig, sd: Dict[str, torch.Tensor], device="cuda", use_cuda_graph: bool = True,
                frame_bucket: int = 0, row_bucket: int = 0) -> ZipVoice:
    kw = cfg.model_kwargs()
    if not cfg.is_dialog:
        kw.pop("spk_a_id", None)
        kw.pop("spk_b_id", None)
    m = MODEL_CLASSES[cfg.variant](use_cuda_graph=use_cuda_graph, frame_bucket=frame_bucket, row_bucket=row_bucket, **kw)
    m.load_state_dict(sd)
    return m.to(device)


def config_from_reference(ref_model) -> ZipVoiceConfig:
    """The hyper-parameters of an instance of the reference's `zipvoice.models.*` classes, read from the module
    attributes and the weight shapes (reference: zipvoice/models/zipvoice.py:38-133, modules/zipformer.py:109-240)."""
    name = type(ref_model).__name__
    variant = {"ZipVoice": "zipvoice", "ZipVoiceDistill": "zipvoice_distill", "ZipVoiceDialog": "zipvoice_dialog",
               "ZipVoiceDialogStereo": "zipvoice_dialog_stereo"}[name]
    sd = ref_model.state_dict()
    fm, te = ref_model.fm_decoder, ref_model.text_encoder
    fm0 = "fm_decoder.encoders.0.layers.0."
    H = fm.num_heads
    return ZipVoiceConfig(
        variant=variant, fm_decoder_downsampling_factor=list(fm.downsampling_factor),
        fm_decoder_num_layers=list(fm.num_encoder_layers), fm_decoder_cnn_module_kernel=list(fm.cnn_module_kernel),
        fm_decoder_feedforward_dim=sd[fm0 + "feed_forward2.in_proj.weight"].shape[0],
        fm_decoder_num_heads=H, fm_decoder_dim=fm.encoder_dim,
        text_encoder_num_layers=te.num_encoder_layers[0],
        text_encoder_feedforward_dim=sd["text_encoder.encoders.0.layers.0.feed_forward2.in_proj.weight"].shape[0],
        text_encoder_cnn_module_kernel=te.cnn_module_kernel[0],
        text_encoder_num_heads=te.num_heads, text_encoder_dim=te.encoder_dim,
        time_embed_dim=fm.time_embed_dim, text_embed_dim=ref_model.text_embed_dim,
        query_head_dim=fm.query_head_dim, value_head_dim=fm.value_head_dim,
        pos_head_dim=sd[fm0 + "self_attn_weights.linear_pos.weight"].shape[0] // H,
        pos_dim=sd[fm0 + "self_attn_weights.linear_pos.weight"].shape[1],
        feat_dim=ref_model.feat_dim, vocab_size=sd["embed.weight"].shape[0], pad_id=ref_model.pad_id,
        spk_a_id=getattr(ref_model, "spk_a_id", 360), spk_b_id=getattr(ref_model, "spk_b_id", 361))


def accelerate(ref_model, use_cuda_graph: bool = True, frame_bucket: int = 0, row_bucket: int = 0):
    """Patch an instance of the *reference* `zipvoice.models.*` classes in place, the way
    `load_trt` does (reference: zipvoice/utils/tensorrt.py:128-143): `fm_decoder`, `text_encoder`
    and `solver` are replaced; `sample`/`sample_intermediate` keep running the reference code."""
    cfg = config_from_reference(ref_model)
    dev = next(ref_model.parameters()).device
    if dev.type != "cuda":
        raise _lib.ZvbError("accelerate(): move the reference model to a CUDA device first (zipvoice_b200 has no CPU path)")
    shadow = build_model(cfg, ref_model.state_dict(), dev, use_cuda_graph, frame_bucket=frame_bucket,
                         row_bucket=row_bucket)
    del ref_model.fm_decoder            # nn.Module child -> plain attribute, as load_trt does
    del ref_model.text_encoder
    ref_model.fm_decoder = shadow.fm_decoder
    ref_model.text_encoder = shadow.text_encoder
    ref_model.solver = shadow.solver
    ref_model._zipvoice_b200 = shadow
    return ref_model
