"""CPU: the drop-in boundary of the reference's plugin class.  ``FastEditor`` must keep the reference's names, argument order,
defaults and error behaviour (reference ``src/pipeline.py:17-293``; the expected values below are restated from it, and are
cross-checked against the reference source itself when it is mounted)."""
import ast
import inspect
import os

import pytest
import torch

REF = "/root/reference/src/pipeline.py"

INIT = [("model_name", "sdxl"), ("device", "cuda"), ("dtype", torch.float16), ("enable_cpu_offload", True),
        ("use_full_precision", False), ("use_full_controlnet", False)]                           # src/pipeline.py:45-46
EDIT = [("image", inspect._empty), ("prompt", inspect._empty), ("negative_prompt", ""), ("strength", 0.80), ("num_inference_steps", 4),
        ("guidance_scale", 1.5), ("controlnet_conditioning_scale", 0.5), ("canny_low_threshold", 100), ("canny_high_threshold", 200),
        ("seed", None)]                                                                          # src/pipeline.py:212-224
PRE = [("image", inspect._empty), ("low_threshold", 100), ("high_threshold", 200)]             # src/pipeline.py:183


def _positional(fn):
    ps = [p for p in inspect.signature(fn).parameters.values() if p.name != "self" and p.kind == p.POSITIONAL_OR_KEYWORD]
    return [(p.name, p.default) for p in ps]


def test_fasteditor_signatures_match_the_reference():
    from src.pipeline import FastEditor                  # the reference's own import path (run_batch.py:15)
    assert _positional(FastEditor.__init__) == INIT     # extensions (state=, prompt_encoder=, tiny=, verbose=) are keyword-only
    assert _positional(FastEditor.edit) == EDIT
    assert _positional(FastEditor.preprocess_image) == PRE
    assert _positional(FastEditor.clear_memory) == [] and _positional(FastEditor.get_memory_usage) == []
    assert set(FastEditor.MODEL_CONFIGS) == {"sdxl", "ssd-1b"}
    assert FastEditor.MODEL_CONFIGS["ssd-1b"]["use_full_lcm"] is True and FastEditor.MODEL_CONFIGS["sdxl"]["use_full_lcm"] is False


@pytest.mark.skipif(not os.path.exists(REF), reason="reference source not mounted")
def test_expected_signatures_are_the_reference_ones():
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "FastEditor")
    fns = {n.name: n for n in cls.body if isinstance(n, ast.FunctionDef)}
    for name, expected in (("__init__", INIT), ("edit", EDIT), ("preprocess_image", PRE)):
        args = [a.arg for a in fns[name].args.args if a.arg != "self"]
        assert args == [e[0] for e in expected], name
        defaults = fns[name].args.defaults
        got = [ast.unparse(d) for d in defaults]
        exp = [e[1] for e in expected if e[1] is not inspect._empty]
        assert len(got) == len(exp)
        for g, e in zip(got, exp):
            assert g == "torch.float16" if e is torch.float16 else ast.literal_eval(g) == e, (name, g, e)
    assert {"clear_memory", "get_memory_usage"} <= set(fns)


def test_unknown_model_is_a_value_error_before_anything_else():
    from src.pipeline import FastEditor
    with pytest.raises(ValueError, match="Unknown model"):                                     # src/pipeline.py:61-62
        FastEditor(model_name="sd15")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from src.pipeline import FastEditor
    with pytest.raises(RuntimeError):
        FastEditor(model_name="ssd-1b", device="cpu", verbose=False)
    with pytest.raises(RuntimeError):
        FastEditor(model_name="ssd-1b", device="cuda", verbose=False)                          # no CUDA device here: loud failure


def test_cli_flags_match_the_reference():
    import run_batch
    import run_single_image
    single = {a.dest for a in run_single_image.build_parser()._actions}
    for flag in ("image", "prompt", "model", "negative_prompt", "steps", "guidance", "control_scale", "canny_low", "canny_high", "seed",
                 "output_dir", "no_cpu_offload", "quality_mode", "full_precision", "full_controlnet", "compute_metrics", "show_plot"):
        assert flag in single, flag                                                             # run_single_image.py:18-60 of the reference
    batch = {a.dest for a in run_batch.build_parser()._actions}
    for flag in ("mapping_file", "source_dir", "output_dir", "model", "num_images", "editing_types", "image_ids", "steps", "guidance",
                 "control_scale", "canny_low", "canny_high", "seed", "no_cpu_offload", "quality_mode", "full_precision", "full_controlnet",
                 "skip_existing", "save_comparisons"):
        assert flag in batch, flag                                                              # run_batch.py:19-100 of the reference


# ---- the evaluation side of the drop-in (reference src/metrics.py:113-386) ----------------------------------------------------------------
REF_METRICS = "/root/reference/src/metrics.py"
METRIC_METHODS = {"__init__": [("device", "cuda")],                                             # src/metrics.py:163
                  "calculate_ssim": ["img1", "img2"], "calculate_lpips": ["img1", "img2"], "calculate_clip_score": ["img", "text"],
                  "calculate_psnr": ["img1", "img2"], "calculate_mse": ["img1", "img2"],
                  "calculate_all_metrics": ["source_img", "edited_img", "prompt"], "clear_memory": []}
DINO_INIT = [("device", inspect._empty), ("model_name", "dino_vitb8"), ("resize_to", 224), ("layer", 11)]   # src/metrics.py:116


def test_metrics_calculator_signatures_match_the_reference():
    from src.metrics import DinoDistanceMetric, MetricsCalculator
    for name, expected in METRIC_METHODS.items():
        got = _positional(getattr(MetricsCalculator, name))
        if name == "__init__":
            assert got == expected                          # checkpoints= / networks= are keyword arguments after the reference's one
        else:
            assert [g[0] for g in got] == expected, name
    got = _positional(DinoDistanceMetric.__init__)
    assert [g[0] for g in got[:4]] == [e[0] for e in DINO_INIT] and [g[1] for g in got[1:4]] == [e[1] for e in DINO_INIT[1:]]
    assert [g[0] for g in _positional(DinoDistanceMetric.calculate_distance)] == ["source_img", "edited_img"]


@pytest.mark.skipif(not os.path.exists(REF_METRICS), reason="reference source not mounted")
def test_expected_metric_signatures_are_the_reference_ones():
    tree = ast.parse(open(REF_METRICS).read())
    classes = {n.name: {f.name: f for f in n.body if isinstance(f, ast.FunctionDef)} for n in tree.body if isinstance(n, ast.ClassDef)}
    calc = classes["MetricsCalculator"]
    for name, expected in METRIC_METHODS.items():
        args = [a.arg for a in calc[name].args.args if a.arg != "self"]
        assert args == [e[0] if isinstance(e, tuple) else e for e in expected], name
    assert [ast.literal_eval(d) for d in calc["__init__"].args.defaults] == ["cuda"]
    dino = classes["DinoDistanceMetric"]
    assert [a.arg for a in dino["__init__"].args.args if a.arg != "self"] == [e[0] for e in DINO_INIT]
    assert [ast.literal_eval(d) for d in dino["__init__"].args.defaults] == [e[1] for e in DINO_INIT[1:]]
    assert [a.arg for a in dino["calculate_distance"].args.args if a.arg != "self"] == ["source_img", "edited_img"]
