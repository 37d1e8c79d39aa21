"""-m gpu: the whole CUDA path (text encoder, fm_decoder seam, Euler+CFG solver seam) against the
fixtures the reference produced; bf16 thresholds from BASELINE.md §5."""
import pytest

import model_checks as mc
from util import CASE_CFG

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASE_CFG))
def test_against_reference_fixture(name):
    mc.assert_case(name, mc.run_case(name))


def test_cuda_graph_replay_matches_eager():
    a = mc.run_case("tiny_zipvoice_cfg", use_cuda_graph=False)
    b = mc.run_case("tiny_zipvoice_cfg", use_cuda_graph=True)
    assert abs(a["x_rel"] - b["x_rel"]) < 1e-6, (a, b)


@pytest.mark.parametrize("switch", ["ZVB_NO_FUSED_PROLOGUE", "ZVB_NO_MERGE", "ZVB_NO_FAST_EPI", "ZVB_NO_PDL"])
def test_plan_switches_keep_parity(switch):
    """The A/B switches of csrc/engine.cu (`load_switches`) are read once per process: each legacy path is run in its own
    interpreter and must stay inside the same thresholds as the default plan."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **{switch: "1"})
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_check.py"), "model:tiny_zipvoice_cfg",
                          "model:base_zipvoice_cfg"], env=env, capture_output=True, text=True, timeout=600)
    lines = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0 and len(lines) == 2 and all(r["ok"] for r in lines), (out.stdout[-2000:], out.stderr[-2000:])
