"""B200-native (sm_100a) hot path of vismaychuriwala/Fast-Image-Editing-with-Generative-Models:
the guided img2img edit behind ``FastEditor.edit`` (reference ``src/pipeline.py:212-274``).

Host code is Python/PyTorch (device memory, streams, torch.distributed); every hot op is a hand-written CUDA
kernel behind the C-ABI in ``include/fie_b200.h`` (``libfie_b200.so``).  There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
