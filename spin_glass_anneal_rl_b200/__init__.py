"""B200-native Ising annealing engine: the batched Monte Carlo sweep path of
spin-glass-anneal-rl behind the reference's own Python API.

    from spin_glass_anneal_rl_b200 import (IsingModel, IsingModelConfig, GPUAnnealer,
                                           GPUAnnealerConfig, ParallelTempering,
                                           ParallelTemperingConfig, UpdateRule, ScheduleType)

``install_as_spin_glass_rl()`` registers the same modules under the reference's import
paths (``spin_glass_rl.core.ising_model`` ...) so unchanged callers pick them up.
Importing this package does not load CUDA; the first anneal() / Engine() does, and fails
loudly if libsg_b200.so or the GPU is missing.
"""
__version__ = "0.1.0"

from .annealing import (AnnealingResult, GPUAnnealer, GPUAnnealerConfig, ParallelTempering,
                        ParallelTemperingConfig, ScheduleType, TemperatureScheduler)
from .core import (ComputeMode, EnergyComputer, EnergyStats, IsingModel, IsingModelConfig, SpinDynamics,
                   UpdateRule)
from .api import (BatchProcessor, VectorizedOperations, anneal, batch_energies, batch_local_fields,
                  install_as_spin_glass_rl)

__all__ = ["ComputeMode", "EnergyComputer", "EnergyStats", "BatchProcessor", "VectorizedOperations", "IsingModel", "IsingModelConfig", "SpinDynamics", "UpdateRule", "GPUAnnealer",
           "GPUAnnealerConfig", "ParallelTempering", "ParallelTemperingConfig", "AnnealingResult",
           "ScheduleType", "TemperatureScheduler", "anneal", "batch_energies",
           "batch_local_fields", "install_as_spin_glass_rl"]
