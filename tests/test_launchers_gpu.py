"""The experiments/ launchers run end to end on one GPU (tiny synthetic configurations)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, cwd):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "experiments"))
    return subprocess.run([sys.executable] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=240)


def test_train_baseline_synthetic(tmp_path):
    r = run([os.path.join(ROOT, "experiments", "train_baseline.py"), "--synthetic", "--epochs", "2", "--size", "32",
             "--train-samples", "32", "--val-samples", "8", "--batch-size", "8"], tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Val PSNR" in r.stdout and os.path.exists(tmp_path / "checkpoints" / "best_model.pt")


@pytest.mark.parametrize("strategy", ["ewc", "replay"])
def test_train_continual(tmp_path, strategy):
    r = run([os.path.join(ROOT, "experiments", "train_continual.py"), "--strategy", strategy, "--samples", "16",
             "--size", "32", "--epochs", "1", "--batch-size", "8"], tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Training complete" in r.stdout and os.path.exists(tmp_path / "checkpoints" / "continual_model.pt")


@pytest.mark.parametrize("extra", [[], ["--recovery"], ["--scale", "4", "--sr-window", "2"], ["--lightweight"]])
def test_infer_video(tmp_path, extra):
    """Clip-sharded inference driver: host clips -> enhance_video -> host, one JSON line with the frame rate."""
    import json
    r = run([os.path.join(ROOT, "experiments", "infer_video.py"), "--clips", "2", "--frames", "6", "--height", "64",
             "--width", "96", "--batch-size", "4", "--save", str(tmp_path / "out")] + extra, tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["metric"] == "enhanced_frames_per_sec" and line["value"] > 0 and line["clips"] == 2
    assert os.path.exists(tmp_path / "out" / "clip_1.pt")
