"""CPU: the host-side pieces of bench.py — the measured CPU reference leg (the fp32 oracle, stage split) and the argument handling
that maps BASELINE.json configs[] to workloads.  The GPU legs are exercised by the driver's bench run."""
import sys

import pytest


def test_cpu_reference_edit_runs_the_whole_path_and_splits_stages():
    import bench
    from fast_image_editing_with_generative_models_b200 import model_zoo
    state = model_zoo.synthetic_state("sdxl", tiny=True)
    t, stages = bench.cpu_reference_edit(state, 64, 2)
    assert t > 0 and {"canny", "vae_encode", "controlnet_step", "unet_step", "vae_decode"} <= set(stages)
    assert abs(sum(stages.values()) - t) < 0.05 * t + 0.05                      # the split accounts for the measured time


@pytest.mark.parametrize("argv,model,batch,impl", [([], "sdxl", 8, "b200"), (["--config", "2"], "ssd-1b", 1, "b200"), (["--config", "1"], "sdxl", 32, "b200"),
                                                  (["--config", "0"], "ssd-1b", 1, "reference"), (["--impl", "reference"], "sdxl", 8, "reference"),
                                                  (["--model", "ssd-1b", "--batch", "4"], "ssd-1b", 4, "b200")])
def test_config_switch(monkeypatch, argv, model, batch, impl):
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    a = bench.parse()
    assert (a.model, a.batch, a.impl) == (model, batch, impl)
    assert "1024x1024" in bench.workload_name(a)
