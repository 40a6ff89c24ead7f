"""``from src.pipeline import FastEditor`` — the reference's import path (``run_single_image.py:14``, ``run_batch.py:15``),
served by the B200-native engine."""
from fast_image_editing_with_generative_models_b200.editor import FastEditor

__all__ = ["FastEditor"]
