"""Host logic of the BatchProcessor mirror (reference annealing/batch_processor.py:22-43, 180-288):
configuration validation, how models are grouped into stacked launches, loud failure without a GPU."""
import pytest
import torch

from conftest import has_cuda
from spin_glass_anneal_rl_b200.annealing.batch_processor import BatchConfig, BatchProcessor, plan_stacks


def test_batch_config_validation_matches_reference():
    c = BatchConfig()
    assert (c.batch_size, c.max_memory_usage, c.prefetch_batches, c.memory_optimization_level,
            c.checkpoint_interval) == (32, 0.8, 2, 1, 100)
    with pytest.raises(ValueError):
        BatchConfig(batch_size=0)
    with pytest.raises(ValueError):
        BatchConfig(max_memory_usage=1.5)
    with pytest.raises(ValueError):
        BatchConfig(memory_optimization_level=3)


def test_plan_stacks_groups_equal_sizes_in_order():
    sizes = [10, 20, 10, 300, 10, 20, 10]
    ok = [True, True, True, False, True, True, True]
    assert plan_stacks(sizes, ok, 2) == [[0, 2], [1, 5], [3], [4, 6]]
    assert plan_stacks(sizes, ok, 32) == [[0, 2, 4, 6], [1, 5], [3]]
    assert plan_stacks([5, 5, 5], [False, False, False], 8) == [[0], [1], [2]]
    assert plan_stacks([], [], 4) == []
    flat = sorted(i for g in plan_stacks(sizes, ok, 3) for i in g)
    assert flat == list(range(len(sizes)))


@pytest.mark.skipif(has_cuda(), reason="CPU-only behaviour")
def test_batch_processor_without_gpu_fails_loudly():
    import spin_glass_anneal_rl_b200 as sg
    from spin_glass_anneal_rl_b200.utils.exceptions import DeviceError
    bp = BatchProcessor(BatchConfig(batch_size=4), sg.GPUAnnealerConfig(n_sweeps=5, n_replicas=2))
    models = []
    for _ in range(3):
        m = sg.IsingModel(sg.IsingModelConfig(n_spins=12, use_sparse=False))
        m.set_couplings_from_matrix(torch.ones(12, 12) - torch.eye(12))
        models.append(m)
    with pytest.raises((DeviceError, RuntimeError, OSError)):
        bp.process_models_batch(models)
    assert bp.get_processing_stats()["processed_batches"] == 0
