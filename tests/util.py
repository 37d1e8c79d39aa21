import os

import torch

from zipvoice_b200.config import ZipVoiceConfig, tiny_config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASE_CFG = {
    "tiny_zipvoice_cfg": lambda: tiny_config("zipvoice"),
    "tiny_zipvoice_g0": lambda: tiny_config("zipvoice"),
    "tiny_distill": lambda: tiny_config("zipvoice_distill"),
    "tiny_dialog": lambda: tiny_config("zipvoice_dialog"),
    "tiny_stereo": lambda: tiny_config("zipvoice_dialog_stereo"),
    "base_zipvoice_cfg": lambda: ZipVoiceConfig("zipvoice"),
}


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())
