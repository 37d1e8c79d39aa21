"""Model hyper-parameters of the ZipVoice hot path.

Mirrors the JSON the reference CLI loads (reference: egs/zipvoice/conf/zipvoice_base.json,
zipvoice/bin/infer_zipvoice.py:796-809) plus the tokenizer-derived `vocab_size`/`pad_id`
and the per-variant switches (reference: zipvoice/models/zipvoice.py:38-60,
zipvoice_distill.py:52-69, zipvoice_dialog.py:31-54,241-256).
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import List, Tuple

VARIANTS = ("zipvoice", "zipvoice_distill", "zipvoice_dialog", "zipvoice_dialog_stereo")


@dataclass
class ZipformerConfig:
    """One TTSZipformer (reference: modules/zipformer.py:109-240)."""

    in_dims: Tuple[int, ...]          # one entry, or two for the stereo two-stream decoder
    out_dims: Tuple[int, ...]
    dim: int
    downsampling_factor: Tuple[int, ...]
    num_layers: Tuple[int, ...]
    cnn_kernel: Tuple[int, ...]
    feedforward_dim: int
    num_heads: int
    query_head_dim: int
    value_head_dim: int
    pos_head_dim: int
    pos_dim: int
    time_embed_dim: int               # -1: no time embedding (text encoder)
    use_guidance_scale_embed: bool = False

    @property
    def ff_dims(self) -> Tuple[int, int, int]:
        # reference: modules/zipformer.py:383-391
        f = self.feedforward_dim
        return ((f * 3) // 4, f, (f * 5) // 4)

    @property
    def na_hidden(self) -> int:
        # reference: modules/zipformer.py:393-395
        return 3 * self.dim // 4

    @property
    def attn_in_dim(self) -> int:
        return (2 * self.query_head_dim + self.pos_head_dim) * self.num_heads


@dataclass
class ZipVoiceConfig:
    variant: str = "zipvoice"
    fm_decoder_downsampling_factor: List[int] = field(default_factory=lambda: [1, 2, 4, 2, 1])
    fm_decoder_num_layers: List[int] = field(default_factory=lambda: [2, 2, 4, 4, 4])
    fm_decoder_cnn_module_kernel: List[int] = field(default_factory=lambda: [31, 15, 7, 15, 31])
    fm_decoder_feedforward_dim: int = 1536
    fm_decoder_num_heads: int = 4
    fm_decoder_dim: int = 512
    text_encoder_num_layers: int = 4
    text_encoder_feedforward_dim: int = 512
    text_encoder_cnn_module_kernel: int = 9
    text_encoder_num_heads: int = 4
    text_encoder_dim: int = 192
    time_embed_dim: int = 192
    text_embed_dim: int = 192
    query_head_dim: int = 32
    value_head_dim: int = 12
    pos_head_dim: int = 4
    pos_dim: int = 48
    feat_dim: int = 100
    vocab_size: int = 360
    pad_id: int = 0
    spk_a_id: int = 360
    spk_b_id: int = 361

    def __post_init__(self):
        assert self.variant in VARIANTS, self.variant

    @property
    def is_dialog(self) -> bool:
        return self.variant in ("zipvoice_dialog", "zipvoice_dialog_stereo")

    @property
    def is_stereo(self) -> bool:
        return self.variant == "zipvoice_dialog_stereo"

    @property
    def is_distill(self) -> bool:
        return self.variant == "zipvoice_distill"

    def model_kwargs(self) -> dict:
        """kwargs accepted by the reference constructors (for the fixture generator)."""
        d = asdict(self)
        d.pop("variant")
        if not self.is_dialog:
            d.pop("spk_a_id")
            d.pop("spk_b_id")
        return d

    def fm_decoder(self) -> ZipformerConfig:
        f = self.feat_dim
        if self.is_stereo:
            in_dims, out_dims = (f * 5, f * 3), (f * 2, f)
        else:
            in_dims, out_dims = (f * 3,), (f,)
        return ZipformerConfig(
            in_dims=in_dims, out_dims=out_dims, dim=self.fm_decoder_dim,
            downsampling_factor=tuple(self.fm_decoder_downsampling_factor),
            num_layers=tuple(self.fm_decoder_num_layers),
            cnn_kernel=tuple(self.fm_decoder_cnn_module_kernel),
            feedforward_dim=self.fm_decoder_feedforward_dim,
            num_heads=self.fm_decoder_num_heads, query_head_dim=self.query_head_dim,
            value_head_dim=self.value_head_dim, pos_head_dim=self.pos_head_dim,
            pos_dim=self.pos_dim, time_embed_dim=self.time_embed_dim,
            use_guidance_scale_embed=self.is_distill)

    def text_encoder(self) -> ZipformerConfig:
        return ZipformerConfig(
            in_dims=(self.text_embed_dim,), out_dims=(self.feat_dim,), dim=self.text_encoder_dim,
            downsampling_factor=(1,), num_layers=(self.text_encoder_num_layers,),
            cnn_kernel=(self.text_encoder_cnn_module_kernel,),
            feedforward_dim=self.text_encoder_feedforward_dim,
            num_heads=self.text_encoder_num_heads, query_head_dim=self.query_head_dim,
            value_head_dim=self.value_head_dim, pos_head_dim=self.pos_head_dim,
            pos_dim=self.pos_dim, time_embed_dim=-1)


def tiny_config(variant: str = "zipvoice") -> ZipVoiceConfig:
    """A shrunken model with the same structure (U-Net 1/2/4/2/1) for second-scale CPU tests."""
    return ZipVoiceConfig(
        variant=variant,
        fm_decoder_downsampling_factor=[1, 2, 4, 2, 1],
        fm_decoder_num_layers=[1, 1, 1, 1, 1],
        fm_decoder_cnn_module_kernel=[31, 15, 7, 15, 31],
        fm_decoder_feedforward_dim=256, fm_decoder_dim=128,
        text_encoder_num_layers=1, text_encoder_feedforward_dim=128, text_encoder_dim=64,
        time_embed_dim=192, text_embed_dim=64,
        vocab_size=362 if "dialog" in variant else 360)
