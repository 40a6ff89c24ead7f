"""CPU: host logic of the batch CLI (reference ``run_batch.py:176-261``) with a fake editor — micro-batching through
``FastEditor.edit_many``, ``--skip_existing``, per-image isolation (a failing batch is retried image by image; a failing image
is counted and the loop continues), missing files / prompts, path traversal."""
import json
import os

import pytest
from PIL import Image


class FakeEditor:
    instances = []

    def __init__(self, **kw):
        self.kw, self.batches, self.singles, self.outputs = kw, [], [], []
        self.fail_batch_with = None
        self.fail_prompt = None
        FakeEditor.instances.append(self)

    def edit_many(self, images, prompts, **kw):
        self.batches.append(list(prompts))
        if self.fail_batch_with is not None and any(self.fail_batch_with in p for p in prompts):
            raise RuntimeError("boom (batch)")
        outs = [Image.new("RGB", (1024, 1024), (i, 0, 0)) for i, _ in enumerate(images)]
        self.outputs.append(kw.get("output", "pil"))
        if kw.get("output") == "jpeg":                    # the GPU encoder's contract: the bytes PIL's own save() would write
            import io
            files = []
            for o in outs:
                b = io.BytesIO()
                o.save(b, "JPEG")
                files.append(b.getvalue())
            return files
        return outs

    def edit(self, image, prompt, **kw):
        self.singles.append(prompt)
        if self.fail_prompt is not None and self.fail_prompt in prompt:
            raise RuntimeError("boom (image)")
        return Image.new("RGB", (1024, 1024), (7, 7, 7))

    def clear_memory(self):
        pass


@pytest.fixture
def dataset(tmp_path):
    src = tmp_path / "src"
    (src / "a").mkdir(parents=True)
    mapping = {}
    for i in range(11):
        name = f"a/img{i:02d}.jpg"
        Image.new("RGB", (64, 48), (i * 20, 10, 10)).save(src / name)
        mapping[f"{i:03d}"] = {"image_path": name, "editing_prompt": f"prompt {i}", "editing_type_id": str(i % 3)}
    mapping["missing"] = {"image_path": "a/nope.jpg", "editing_prompt": "x", "editing_type_id": "0"}
    mapping["noprompt"] = {"image_path": "a/img00.jpg", "editing_prompt": "", "editing_type_id": "0"}
    mapping["evil"] = {"image_path": "../outside.jpg", "editing_prompt": "x", "editing_type_id": "0"}
    mp = tmp_path / "mapping.json"
    mp.write_text(json.dumps(mapping))
    return tmp_path, str(mp), str(src)


def _run(monkeypatch, dataset, extra=(), prepare=None):
    import run_batch
    tmp, mp, src = dataset
    FakeEditor.instances.clear()

    def make(**kw):
        e = FakeEditor(**kw)
        if prepare:
            prepare(e)
        return e
    monkeypatch.setattr(run_batch, "FastEditor", make)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    out = tmp / "out"
    run_batch.main(["--mapping_file", mp, "--source_dir", src, "--output_dir", str(out), "--model", "ssd-1b", "--micro_batch", "4", *extra])
    return FakeEditor.instances[0], out / "batch" / "edited" / "ssd-1b_fp16"


def test_micro_batches_and_outputs(monkeypatch, dataset, capsys):
    ed, outdir = _run(monkeypatch, dataset)
    assert [len(b) for b in ed.batches] == [4, 4, 3] and not ed.singles          # 11 valid images in groups of 4
    assert sorted(os.listdir(outdir / "a")) == [f"img{i:02d}.jpg" for i in range(11)]
    txt = capsys.readouterr().out
    assert "Processed:  11 images" in txt and "Failed:     3 images" in txt     # missing file, empty prompt, path traversal


def test_skip_existing(monkeypatch, dataset, capsys):
    _run(monkeypatch, dataset)
    capsys.readouterr()
    ed, _ = _run(monkeypatch, dataset, extra=["--skip_existing"])
    assert not ed.batches
    assert "Skipped:    12 images" in capsys.readouterr().out                      # 11 outputs + the empty-prompt entry's existing output


def test_failed_batch_falls_back_to_per_image(monkeypatch, dataset, capsys):
    def prepare(e):
        e.fail_batch_with = "prompt 5"
        e.fail_prompt = "prompt 5"
    ed, outdir = _run(monkeypatch, dataset, prepare=prepare)
    assert ed.singles == ["prompt 4", "prompt 5", "prompt 6", "prompt 7"]           # only the failing group is retried image by image
    assert "img05.jpg" not in os.listdir(outdir / "a") and len(os.listdir(outdir / "a")) == 10
    txt = capsys.readouterr().out
    assert "Processed:  10 images" in txt and "Failed:     4 images" in txt


def test_selection_filters(monkeypatch, dataset):
    ed, _ = _run(monkeypatch, dataset, extra=["--editing_types", "1", "--num_images", "3"])
    assert sum(len(b) for b in ed.batches) == 3
    ed, _ = _run(monkeypatch, dataset, extra=["--image_ids", "002", "007"])
    assert ed.batches == [["prompt 2", "prompt 7"]]


def test_jpg_outputs_are_written_from_gpu_encoded_bytes(monkeypatch, dataset):
    ed, outdir = _run(monkeypatch, dataset)
    assert set(ed.outputs) == {"jpeg"}                                               # *.jpg targets: files arrive already encoded
    assert Image.open(outdir / "a" / "img03.jpg").size == (1024, 1024)
    ed, _ = _run(monkeypatch, dataset, extra=["--no_gpu_jpeg"])
    assert set(ed.outputs) == {"pil"}
