"""Drop-in import surface of the reference (``from src.pipeline import FastEditor``, reference ``src/__init__.py:4-7``).
``MetricsCalculator`` is imported lazily: the reference imports it eagerly, which needs torchmetrics (absent here)."""
from .pipeline import FastEditor

__all__ = ["FastEditor", "MetricsCalculator"]


def __getattr__(name):
    if name == "MetricsCalculator":
        from .metrics import MetricsCalculator
        return MetricsCalculator
    raise AttributeError(name)
