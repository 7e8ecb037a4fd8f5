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


@pytest.mark.parametrize("seed", [21, 22])
def test_random_small_stacked_and_group_models_agree(seed):
    """tools/fuzz_small_groups.py: stacked small models against the oracle (replay) and against
    the same model alone on the SIMT kernel; group models on the partitioned kernel, the
    one-warp-per-word kernel and the sparse kernel."""
    if not has_cuda():
        pytest.skip("needs a CUDA device")
    import fuzz_small_groups
    assert fuzz_small_groups.run(seed, 10) == 0
