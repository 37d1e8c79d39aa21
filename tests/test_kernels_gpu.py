"""-m gpu: every hand-written kernel against a PyTorch fp32 computation of the same op, called
through the C ABI test entry points (tolerances in tests/kernel_checks.py)."""
import pytest

import kernel_checks as kc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(kc.ALL))
def test_kernel(name):
    kc.assert_ok(name, kc.ALL[name]())
