from .gpu_annealer import GPUAnnealer, GPUAnnealerConfig
from .multi_gpu import MultiGPUAnnealer, MultiGPUConfig, global_argmin, shard_replicas
from .parallel_tempering import ParallelTempering, ParallelTemperingConfig
from .result import AnnealingResult
from .temperature_scheduler import (ScheduleConfig, ScheduleType, TemperatureSchedule,
                                    TemperatureScheduler)

__all__ = ["GPUAnnealer", "GPUAnnealerConfig", "ParallelTempering", "ParallelTemperingConfig",
           "AnnealingResult", "ScheduleType", "ScheduleConfig", "TemperatureSchedule",
           "TemperatureScheduler", "MultiGPUAnnealer", "MultiGPUConfig", "global_argmin",
           "shard_replicas"]
