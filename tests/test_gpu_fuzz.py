"""Randomised differential test (tools/fuzz_parity.py): oracle vs tensor-core vs sequential-FMA
vs sparse kernels on random integer models, shapes, rules, site orders; all bit-identical."""
import os
import sys

import pytest

from conftest import ROOT, has_cuda

sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_random_models_all_kernels_agree(seed):
    if not has_cuda():
        pytest.skip("needs a CUDA device")
    import fuzz_parity
    assert fuzz_parity.run(seed, 10, verbose=False) == 0
