from .ising_model import IsingModel, IsingModelConfig
from .spin_dynamics import SpinDynamics, UpdateRule

__all__ = ["IsingModel", "IsingModelConfig", "SpinDynamics", "UpdateRule"]
