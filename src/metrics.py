"""Drop-in module path of the reference's ``src/metrics.py``: ``from src.metrics import MetricsCalculator`` (``evaluate.py:15``,
``run_batch.py --compute_metrics``) resolves to the B200 implementation."""
from fast_image_editing_with_generative_models_b200.metrics import DinoDistanceMetric, MetricsCalculator  # noqa: F401

__all__ = ["MetricsCalculator", "DinoDistanceMetric"]
