"""Drop-in import surface of the reference (``from src import FastEditor, MetricsCalculator``, reference ``src/__init__.py:4-7``)."""
from .pipeline import FastEditor
from .metrics import MetricsCalculator

__all__ = ["FastEditor", "MetricsCalculator"]
