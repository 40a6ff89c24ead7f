"""CPU: the C-ABI library builds, loads, and exports every symbol include/fie_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from fast_image_editing_with_generative_models_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(fie_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_gpu_calls_needed_for_queries():
    L = _lib.lib()
    assert L.fie_version() >= 100
    assert L.fie_geglu_block_n(10240) == 256 and L.fie_geglu_block_n(512) == 256 and L.fie_geglu_block_n(384) == 128
    assert L.fie_canny_workspace_bytes(1, 1024, 1024) >= 6 * 1024 * 1024


def test_epilogue_struct_layout():
    # fie_epilogue: 7 x 8-byte fields + float + 2 ints = 68 -> 72; + gn_stats pointer, gn_groups int (+4 pad), gn_rows_per_image = 96;
    # + ln_stats_out, ln_stats_in pointers, ln_eps float, ln_dim int = 120; + row_scale pointer = 128
    assert ctypes.sizeof(_lib.Epilogue) == 128


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fast_image_editing_with_generative_models_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
