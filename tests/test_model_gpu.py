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
