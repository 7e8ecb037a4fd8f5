"""Temperature schedules (host side).

Same names and semantics as the reference's
``spin_glass_rl.annealing.temperature_scheduler`` (ScheduleType :11-21, ScheduleConfig
:24-40, the nine schedule classes :69-269, the TemperatureScheduler factory :272-331,
recommend_schedule :423-462), written here as one closed-form table plus thin named
classes.  Every schedule except ADAPTIVE is a pure function of the sweep index, so the
annealer evaluates it for all sweeps up front (``precompute``) and hands the kernel a
temperature array; ADAPTIVE consumes the running acceptance rate and is evaluated
sweep by sweep.

Quirks kept on purpose (callers and golden traces depend on them): GEOMETRIC ignores
``total_sweeps`` (:119-122); LOGARITHMIC scales c/log(1+t) back by T0/c (:135-142);
LINEAR clamps at final_temp; sweep 0 of LOGARITHMIC/FAST/BOLTZMANN returns T0.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Callable, Dict, List, Optional

import numpy as np


class ScheduleType(Enum):
    LINEAR = "linear"
    EXPONENTIAL = "exponential"
    GEOMETRIC = "geometric"
    LOGARITHMIC = "logarithmic"
    POWER_LAW = "power_law"
    ADAPTIVE = "adaptive"
    FAST = "fast"
    BOLTZMANN = "boltzmann"
    CUSTOM = "custom"


@dataclass
class ScheduleConfig:
    schedule_type: ScheduleType
    initial_temp: float
    final_temp: float
    total_sweeps: int
    alpha: float = 0.95
    k: float = 1.0
    c: float = 1.0
    target_acceptance: float = 0.44
    adaptation_window: int = 100
    adaptation_rate: float = 0.1


def _closed_form(cfg: ScheduleConfig, kind: ScheduleType, sweep: int) -> float:
    T0, Tf = cfg.initial_temp, cfg.final_temp
    if kind is ScheduleType.LINEAR:
        if sweep >= cfg.total_sweeps:
            return Tf
        return max(T0 - (T0 - Tf) * (sweep / cfg.total_sweeps), Tf)
    if kind is ScheduleType.EXPONENTIAL:
        lam = -np.log(Tf / T0) / cfg.total_sweeps if Tf > 0 else 0.01
        return max(T0 * np.exp(-lam * sweep), Tf)
    if kind is ScheduleType.GEOMETRIC:
        return max(T0 * (cfg.alpha ** sweep), Tf)
    if kind is ScheduleType.LOGARITHMIC:
        if sweep == 0:
            return T0
        return max((cfg.c / np.log(1 + sweep)) * T0 / cfg.c, Tf)
    if kind is ScheduleType.POWER_LAW:
        return max(T0 / ((1 + sweep) ** cfg.k), Tf)
    if kind is ScheduleType.FAST:
        return T0 if sweep == 0 else max(T0 / sweep, Tf)
    if kind is ScheduleType.BOLTZMANN:
        return T0 if sweep == 0 else max(T0 / np.log(1 + sweep), Tf)
    raise ValueError(f"no closed form for {kind}")


class TemperatureSchedule:
    """Base class: ``get_temperature(sweep)`` and the stateful ``update(sweep, **kw)``."""

    kind: Optional[ScheduleType] = None
    stateless = True  # T depends on the sweep index only

    def __init__(self, config: ScheduleConfig):
        self.config = config
        self.reset()

    def reset(self) -> None:
        self.current_sweep = 0
        self.current_temp = self.config.initial_temp
        self.temperature_history: List[float] = [self.config.initial_temp]

    def get_temperature(self, sweep: int) -> float:
        return _closed_form(self.config, self.kind, sweep)

    def update(self, sweep: int, **kwargs) -> float:
        self.current_sweep = sweep
        self.current_temp = self.get_temperature(sweep)
        self.temperature_history.append(self.current_temp)
        return self.current_temp

    def precompute(self, n_sweeps: int) -> np.ndarray:
        """T(0..n_sweeps-1) as float64 (only for stateless schedules)."""
        if not self.stateless:
            raise TypeError(f"{type(self).__name__} depends on run-time feedback")
        return np.array([float(self.get_temperature(s)) for s in range(n_sweeps)], dtype=np.float64)


def _named(kind: ScheduleType, doc: str):
    return type(kind.name.title().replace("_", "") + "Schedule", (TemperatureSchedule,),
                {"kind": kind, "__doc__": doc})


LinearSchedule = _named(ScheduleType.LINEAR, "T(t) = T0 - (T0 - Tf) t / total, clamped at Tf.")
ExponentialSchedule = _named(ScheduleType.EXPONENTIAL, "T(t) = T0 exp(-lambda t), lambda from Tf.")
GeometricSchedule = _named(ScheduleType.GEOMETRIC, "T(t) = T0 alpha^t, clamped at Tf.")
LogarithmicSchedule = _named(ScheduleType.LOGARITHMIC, "T(t) = T0 / log(1 + t).")
PowerLawSchedule = _named(ScheduleType.POWER_LAW, "T(t) = T0 / (1 + t)^k.")
FastSchedule = _named(ScheduleType.FAST, "T(t) = T0 / t.")
BoltzmannSchedule = _named(ScheduleType.BOLTZMANN, "T(t) = T0 / log(1 + t).")


class AdaptiveSchedule(TemperatureSchedule):
    """Geometric base, multiplied by (1 -/+ adaptation_rate) once ``adaptation_window``
    acceptance rates have been seen and their mean is above/below ``target_acceptance``."""

    kind = ScheduleType.ADAPTIVE
    stateless = False

    def __init__(self, config: ScheduleConfig):
        super().__init__(config)
        self.acceptance_history: List[float] = []
        self.base_schedule = GeometricSchedule(config)

    def get_temperature(self, sweep: int) -> float:
        return self.current_temp

    def update(self, sweep: int, acceptance_rate: Optional[float] = None, **kwargs) -> float:
        cfg = self.config
        self.current_sweep = sweep
        if acceptance_rate is not None:
            self.acceptance_history.append(acceptance_rate)
        base = self.base_schedule.get_temperature(sweep)
        if len(self.acceptance_history) >= cfg.adaptation_window:
            recent = np.mean(self.acceptance_history[-cfg.adaptation_window:])
            factor = 1.0
            if recent > cfg.target_acceptance:
                factor = 1.0 - cfg.adaptation_rate
            elif recent < cfg.target_acceptance:
                factor = 1.0 + cfg.adaptation_rate
            self.current_temp = max(base * factor, cfg.final_temp)
        else:
            self.current_temp = base
        self.temperature_history.append(self.current_temp)
        return self.current_temp


class CustomSchedule(TemperatureSchedule):
    """User function of the sweep index, clamped at final_temp."""

    kind = ScheduleType.CUSTOM

    def __init__(self, config: ScheduleConfig, schedule_func: Callable[[int], float]):
        self.schedule_func = schedule_func
        super().__init__(config)

    def get_temperature(self, sweep: int) -> float:
        return max(self.schedule_func(sweep), self.config.final_temp)


class TemperatureScheduler:
    """Factory for the schedules above."""

    _schedule_classes: Dict[ScheduleType, type] = {
        ScheduleType.LINEAR: LinearSchedule, ScheduleType.EXPONENTIAL: ExponentialSchedule,
        ScheduleType.GEOMETRIC: GeometricSchedule, ScheduleType.LOGARITHMIC: LogarithmicSchedule,
        ScheduleType.POWER_LAW: PowerLawSchedule, ScheduleType.ADAPTIVE: AdaptiveSchedule,
        ScheduleType.FAST: FastSchedule, ScheduleType.BOLTZMANN: BoltzmannSchedule,
        ScheduleType.CUSTOM: CustomSchedule,
    }

    @classmethod
    def create_schedule(cls, schedule_type: ScheduleType, initial_temp: float, final_temp: float,
                        total_sweeps: int, custom_func: Optional[Callable[[int], float]] = None,
                        **kwargs) -> TemperatureSchedule:
        config = ScheduleConfig(schedule_type=schedule_type, initial_temp=initial_temp,
                                final_temp=final_temp, total_sweeps=total_sweeps, **kwargs)
        klass = cls._schedule_classes[schedule_type]
        if schedule_type is ScheduleType.CUSTOM:
            if custom_func is None:
                raise ValueError("custom_func required for CUSTOM schedule type")
            return klass(config, custom_func)
        return klass(config)

    @classmethod
    def get_available_schedules(cls) -> List[str]:
        return [s.value for s in ScheduleType]

    @classmethod
    def compare_schedules(cls, initial_temp: float, final_temp: float, total_sweeps: int,
                          schedule_types: Optional[List[ScheduleType]] = None) -> dict:
        kinds = schedule_types or [ScheduleType.LINEAR, ScheduleType.EXPONENTIAL,
                                   ScheduleType.GEOMETRIC, ScheduleType.LOGARITHMIC]
        step = max(1, total_sweeps // 100)
        out = {}
        for kind in kinds:
            if kind is ScheduleType.CUSTOM:
                continue
            sched = cls.create_schedule(kind, initial_temp, final_temp, total_sweeps)
            out[kind.value] = [sched.get_temperature(s) for s in range(0, total_sweeps, step)]
        return out

    @classmethod
    def recommend_schedule(cls, problem_size: int, time_budget: int,
                           convergence_preference: str = "balanced"):
        if convergence_preference == "fast":
            return ((ScheduleType.FAST, {"k": 1.0}) if problem_size < 1000
                    else (ScheduleType.EXPONENTIAL, {"alpha": 0.99}))
        if convergence_preference == "quality":
            return ((ScheduleType.LOGARITHMIC, {"c": 10.0}) if time_budget > 10000
                    else (ScheduleType.GEOMETRIC, {"alpha": 0.95}))
        if problem_size < 1000:
            return ScheduleType.GEOMETRIC, {"alpha": 0.95}
        return ScheduleType.ADAPTIVE, {"alpha": 0.95, "target_acceptance": 0.44,
                                       "adaptation_window": 100}

    def __repr__(self) -> str:
        return f"TemperatureScheduler(available_schedules=[{', '.join(self.get_available_schedules())}])"
