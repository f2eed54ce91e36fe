"""pytest configuration: `gpu` marker, import paths, shared fixtures/helpers."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "continual-learning-for-dynamic-video-quality-enhancement_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  -- the fp32-path criterion of BASELINE.json (<= 1e-4)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = float(((a.detach().double().cpu() - b.detach().double().cpu()) ** 2).mean())
    return 10.0 * np.log10(1.0 / max(mse, 1e-30))


def build_case(meta, device="cpu"):
    """Re-create the module of a golden SR case from its seed (weights are a function of the seed:
    the drop-in module builds the same torch layers in the same order as the reference)."""
    from nerve_cl_b200.models import SuperResolutionNet
    scale, feats, blocks, tw, b, h, w, seed, training = [int(v) for v in meta]
    torch.manual_seed(seed)
    model = SuperResolutionNet(scale_factor=scale, num_features=feats, num_residual_blocks=blocks,
                               temporal_window=tw)
    g = torch.Generator().manual_seed(seed + 1)
    for i in range(3):
        bn = model.feature_extractor.body[i].bn
        with torch.no_grad():
            bn.running_mean.copy_(0.05 * torch.randn(feats, generator=g))
            bn.running_var.copy_(1.0 + 0.2 * torch.rand(feats, generator=g))
    model.train(bool(training))
    return model.to(device), scale, bool(training)
