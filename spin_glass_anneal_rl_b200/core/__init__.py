from .energy_computer import ComputeMode, EnergyComputer, EnergyStats
from .ising_model import IsingModel, IsingModelConfig
from .spin_dynamics import SpinDynamics, UpdateRule

__all__ = ["ComputeMode", "EnergyComputer", "EnergyStats", "IsingModel", "IsingModelConfig", "SpinDynamics",
           "UpdateRule"]
