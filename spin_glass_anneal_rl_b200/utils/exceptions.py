"""Exception types raised at the annealing boundary.

Mirrors the names callers of the reference catch
(reference utils/exceptions.py:6-58: SpinGlassError, ModelError, AnnealingError,
DeviceError, ValidationError, ConfigurationError).
"""
from typing import Any, Dict, Optional


class SpinGlassError(Exception):
    def __init__(self, message: str, details: Optional[Dict[str, Any]] = None):
        super().__init__(message)
        self.message = message
        self.details = dict(details or {})

    def __str__(self) -> str:
        if not self.details:
            return self.message
        return f"{self.message} (Details: " + ", ".join(f"{k}={v}" for k, v in self.details.items()) + ")"


class ModelError(SpinGlassError):
    pass


class AnnealingError(SpinGlassError):
    pass


class DeviceError(SpinGlassError):
    pass


class ValidationError(SpinGlassError):
    pass


class ConfigurationError(SpinGlassError):
    pass
