"""zipvoice_b200 — B200-native (sm_100a) implementation of the ZipVoice sampler hot path:
`ZipVoice.sample` / `sample_intermediate` -> Euler ODE + classifier-free guidance ->
TTSZipformer fm_decoder (and the one-shot text_encoder), behind the reference's own model API."""
from .config import ZipVoiceConfig, ZipformerConfig, tiny_config, VARIANTS  # noqa: F401
from .synth import synth_state_dict, synth_utterances  # noqa: F401


def __getattr__(name):
    # model/engine import torch + ctypes lazily so that config/synth stay usable without CUDA
    if name in ("ZipVoice", "ZipVoiceDistill", "ZipVoiceDialog", "ZipVoiceDialogStereo", "build_model",
                "accelerate", "B200EulerSolver", "B200Zipformer", "MODEL_CLASSES"):
        from . import model
        return getattr(model, name)
    raise AttributeError(name)
