"""CPU: the C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol that
include/zipvoice_b200.h declares; host-only entry points work; compute entry points fail loudly."""
import ctypes as C
import os
import re

import pytest
import torch

import __graft_entry__ as entry
from zipvoice_b200 import _lib
from zipvoice_b200.config import ZipVoiceConfig, tiny_config
from zipvoice_b200.synth import synth_state_dict
from zipvoice_b200.weights import PackedZipformer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "zipvoice_b200.h")).read()
    declared = set(re.findall(r"\b(zvb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    raw = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.zvb_abi_version() == _lib.ZVB_ABI_VERSION


def test_library_is_sm100a_tcgen05_tma():
    sass = os.popen(f"cuobjdump -sass {_lib.LIB_PATH} 2>/dev/null").read()
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):     # tcgen05.mma / TMA load / tcgen05.ld
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass                      # no legacy mma.sync path


def test_workspace_sizing_is_host_only(lib):
    cfg = ZipVoiceConfig()
    pk = PackedZipformer(synth_state_dict(tiny_config()), "fm_decoder.", tiny_config().fm_decoder(), "cpu")
    m, keep = pk.model_struct(200)
    n = C.c_size_t()
    assert lib.zvb_plan_workspace_bytes(C.byref(m), 4, 200, C.byref(n)) == 0
    small = n.value
    assert lib.zvb_plan_workspace_bytes(C.byref(m), 8, 200, C.byref(n)) == 0
    assert n.value > small > 0
    assert lib.zvb_plan_workspace_bytes(C.byref(m), 0, 200, C.byref(n)) == -1  # ZVB_ERR_INVALID
    assert b"positive" in lib.zvb_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from zipvoice_b200.model import build_model
    cfg = tiny_config()
    pk = PackedZipformer(synth_state_dict(cfg), "fm_decoder.", cfg.fm_decoder(), "cpu")
    m, keep = pk.model_struct(64)
    ws = torch.zeros(1 << 20, dtype=torch.uint8)
    h = C.c_void_p()
    rc = lib.zvb_plan_create(C.byref(m), 2, 64, ws.data_ptr(), ws.numel(), C.byref(h))
    assert rc in (-4, -2), rc                           # no device (or workspace) -- never a silent CPU path
    model = build_model(cfg, synth_state_dict(cfg), "cpu")
    with pytest.raises(_lib.ZvbError):
        model.sample([[1, 2]], [[3]], torch.zeros(1, 4, 100), torch.tensor([4]), num_step=1)


def test_vocoder_workspace_sizing_and_errors_are_host_only(lib):
    """zvb_vocoder_workspace_bytes needs no device; bad descriptions are rejected with a message."""
    from zipvoice_b200.vocoder import PackedVocos, synth_vocos_state_dict
    pk = PackedVocos(synth_vocos_state_dict(0, dim=256, intermediate=512, n_layers=2), "cpu")
    n = C.c_size_t()
    assert lib.zvb_vocoder_workspace_bytes(C.byref(pk.struct), 2, 100, C.byref(n)) == 0
    small = n.value
    assert lib.zvb_vocoder_workspace_bytes(C.byref(pk.struct), 4, 100, C.byref(n)) == 0
    assert n.value > small > 2 * 100 * 1024 * 4                       # at least the windowed time frames
    assert lib.zvb_vocoder_workspace_bytes(C.byref(pk.struct), 2, 1, C.byref(n)) == -1
    assert b"T at least 2" in lib.zvb_last_error()
    bad = _lib.zvb_vocoder.from_buffer_copy(pk.struct)
    bad.hop = 300                                                     # does not divide n_fft
    assert lib.zvb_vocoder_workspace_bytes(C.byref(bad), 2, 100, C.byref(n)) == -1
    assert b"hop" in lib.zvb_last_error()
    bad = _lib.zvb_vocoder.from_buffer_copy(pk.struct)
    bad.abi_version = 1
    assert lib.zvb_vocoder_workspace_bytes(C.byref(bad), 2, 100, C.byref(n)) == -1
    assert b"ABI" in lib.zvb_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_audio_stages_have_no_cpu_fallback(lib):
    from zipvoice_b200.frontend import VocosFbank
    from zipvoice_b200.vocoder import Vocos, synth_vocos_state_dict
    with pytest.raises(_lib.ZvbError):
        VocosFbank(device="cpu")
    voc = Vocos().load_state_dict(synth_vocos_state_dict(0, dim=256, intermediate=512, n_layers=1))
    with pytest.raises(_lib.ZvbError):
        voc.decode(torch.zeros(1, 100, 10))
    w = torch.zeros(1, 2048)
    out = torch.zeros(1, 8, 100)
    rc = lib.zvb_fbank(w.data_ptr(), None, 1, 2048, None, None, None, 100, 256, 1.0, out.data_ptr(), 8, None)
    assert rc in (-4, -1)                                             # no device (or null argument): never computed on the host


def _shape(lib, rows, n_out, k, lean_kind, attn_ctas=1, q_tiles=1, num_sms=148):
    bn, pair, cs = C.c_int(), C.c_int(), C.c_int()
    assert lib.zvb_debug_launch_shape(rows, n_out, k, lean_kind, num_sms, attn_ctas, q_tiles, C.byref(bn), C.byref(pair),
                                      C.byref(cs)) == 0
    return bn.value, pair.value, cs.value


def test_launch_shapes_are_host_only_and_follow_the_measured_rules(lib):
    """The launch-shape rules of DESIGN.md §3 (tile widths of small problems on the lean epilogues, CTA pairs only from
    2 x SMs m-tiles on, key split of the attention weights on small grids and wave tails), queried without a device."""
    one = 2 * 1218                                   # a single utterance: two CFG rows
    # residual-stream GEMMs (N = 512): 128-column lean tiles in one wave, no pairs; K = 1920 likewise
    assert _shape(lib, one, 512, 512, 2)[:2] == (128, 0)
    assert _shape(lib, one, 512, 1920, 2)[:2] == (128, 0)
    # feed-forward input GEMMs: multiples of 64 (lean plain epilogue), no pairs
    for n_out in (1424, 1536, 1920):
        bn, pair, _ = _shape(lib, one, n_out, 512, 1)
        assert bn % 64 == 0 and pair == 0, (n_out, bn, pair)
    # an op that can only take the generic epilogue may use any multiple of 16
    assert _shape(lib, one, 512, 512, 0)[0] % 16 == 0
    # the batch-64 workload: 256-column tiles as CTA pairs
    big = 128 * 1219
    assert _shape(lib, big, 512, 1536, 2)[:2] == (256, 1)
    assert _shape(lib, big, 1536, 512, 1)[:2] == (256, 1)
    # pairs start at 2 x SMs m-tiles (8 utterances: 153 m-tiles, 16: 305)
    assert _shape(lib, 16 * 1219, 512, 1536, 2)[1] == 0
    assert _shape(lib, 32 * 1219, 512, 1536, 2)[1] == 1
    # attention key split: 80 CTAs of 10 key tiles -> pairs; 40 of 5 -> four CTAs; a 60 s dialog (416 CTAs of 52) -> pairs
    # (tail of the last wave); batch 64 (5120 CTAs) and the stereo batch (2432) -> none
    assert _shape(lib, one, 512, 512, 2, attn_ctas=80, q_tiles=10)[2] == 2
    assert _shape(lib, one, 512, 512, 2, attn_ctas=40, q_tiles=5)[2] == 4
    assert _shape(lib, one, 512, 512, 2, attn_ctas=416, q_tiles=52)[2] == 2
    assert _shape(lib, one, 512, 512, 2, attn_ctas=5120, q_tiles=10)[2] == 1
    assert _shape(lib, one, 512, 512, 2, attn_ctas=2432, q_tiles=19)[2] == 1
    # errors are reported without a device as well
    assert lib.zvb_debug_launch_shape(0, 512, 512, 0, 148, 1, 1, None, None, None) == -1
    assert b"positive" in lib.zvb_last_error()
