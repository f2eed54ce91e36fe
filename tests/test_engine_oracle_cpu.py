"""CPU: oracle/recovery_oracle.py (FrameRecoveryNet, LightweightSuperResolution, EnhancementEngine restatements)
against the golden fixtures generated from the live reference (tests/golden/make_engine_golden.py), and the drop-in
modules' parameter trees against the fixtures' shapes.  No kernel is launched."""
import numpy as np
import torch

from conftest import load_golden, relerr
from oracle import recovery_oracle as ro


def perturb_bn(module, seed):
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            with torch.no_grad():
                m.running_mean.copy_(0.05 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(1.0 + 0.2 * torch.rand(m.num_features, generator=g))


def build_recovery():
    from nerve_cl_b200.models import FrameRecoveryNet
    g = load_golden("recovery_b16.npz")
    seed, bn_seed, bc, tw = [int(v) for v in g["meta"]]
    torch.manual_seed(seed)
    net = FrameRecoveryNet(base_channels=bc, temporal_window=tw).eval()
    perturb_bn(net, bn_seed)
    return g, net


def build_lightweight():
    from nerve_cl_b200.models import LightweightSuperResolution
    g = load_golden("lightweight_x2.npz")
    seed, bn_seed, scale = [int(v) for v in g["meta"]]
    torch.manual_seed(seed)
    net = LightweightSuperResolution(scale_factor=scale)
    perturb_bn(net, bn_seed)
    return g, net, scale


def build_engines():
    from nerve_cl_b200.models import EnhancementConfig, EnhancementEngine
    g = load_golden("engine_small.npz")
    s0, b0, s1, b1 = [int(v) for v in g["meta"]]
    torch.manual_seed(s0)
    eng = EnhancementEngine(EnhancementConfig(frame_recovery_enabled=True, recovery_base_channels=16, recovery_temporal_window=2,
                                              super_resolution_enabled=True, scale_factor=2, sr_num_features=16,
                                              sr_num_residual_blocks=1, sr_temporal_window=1)).eval()
    perturb_bn(eng, b0)
    torch.manual_seed(s1)
    eng2 = EnhancementEngine(EnhancementConfig(frame_recovery_enabled=False, sr_num_features=16, sr_num_residual_blocks=1)).eval()
    perturb_bn(eng2, b1)
    return g, eng, eng2


def test_recovery_oracle_golden():
    g, net = build_recovery()
    sd = net.state_dict()                      # holders build the reference's layers in the reference's order
    for tag in ("a", "b", "c"):
        out = ro.frame_recovery_forward(sd, *(torch.from_numpy(g[f"{tag}/{k}"]) for k in ("frame", "refs", "mask")))
        assert relerr(out, torch.from_numpy(g[f"{tag}/out"])) <= 1e-5, tag


def test_lightweight_oracle_golden():
    g, net, scale = build_lightweight()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.from_numpy(g["x"])
    assert relerr(ro.lightweight_forward(sd, x, scale), torch.from_numpy(g["out_eval"])) <= 1e-6
    assert relerr(ro.lightweight_forward(sd, x, scale, training=True), torch.from_numpy(g["out_train"])) <= 1e-6
    for k in sd:
        if "running" in k:
            assert relerr(sd[k], torch.from_numpy(g["bn1/" + k])) <= 1e-6


def test_engine_oracle_golden():
    g, eng, eng2 = build_engines()
    rec_sd, sr_sd = eng.frame_recovery.state_dict(), eng.super_resolution.state_dict()
    frames, mask = torch.from_numpy(g["frames"]), torch.from_numpy(g["mask"])
    cases = {"plain": {}, "masked": {"mask": mask}, "s06": {"mask": mask, "strength": 0.6},
             "edge": {"center_idx": 0, "mask": mask, "strength": 0.8}}
    with torch.no_grad():
        for tag, kw in cases.items():
            res = ro.engine_forward(sr_sd, rec_sd, frames, 2, 1, **kw)
            assert sorted(res) == list(g[f"fwd/{tag}/keys"])
            assert relerr(res["enhanced"], torch.from_numpy(g[f"fwd/{tag}/enhanced"])) <= 2e-5, tag
        table = ro.engine_window_table(7, 2, 1)
        assert np.array_equal(np.array(table), g["window_table"])
        assert eng.window_table(7) == table                              # the product's host-side window logic
        video, masks = torch.from_numpy(g["video"]), torch.from_numpy(g["masks"])
        outs = [ro.engine_forward(sr_sd, rec_sd, video[None, s:e], 2, 1, c, masks[t:t + 1])["enhanced"]
                for t, (s, e, c) in enumerate(table)]
        assert relerr(torch.stack(outs, 1)[0][..., ::2, ::2], torch.from_numpy(g["video_out"])) <= 2e-5
        sr2 = eng2.super_resolution.state_dict()
        assert relerr(ro.engine_forward(sr2, None, frames, 2, 1)["enhanced"], torch.from_numpy(g["sronly/fwd"])) <= 2e-5


def test_engine_contract_cpu():
    """Constructor / attribute / error contract that needs no GPU."""
    from nerve_cl_b200.models import EnhancementConfig, EnhancementEngine, FrameRecoveryNet, LightweightSuperResolution
    import pytest
    eng = EnhancementEngine(EnhancementConfig(recovery_base_channels=16, sr_num_features=16, sr_num_residual_blocks=1))
    info = eng.get_model_info()
    assert info["parameters"]["total"] == sum(p.numel() for p in eng.parameters())
    assert set(info["parameters"]) == {"total", "trainable", "frame_recovery", "super_resolution"}
    eng.set_enhancement_mode("sr_only")
    assert not eng.config.frame_recovery_enabled and eng.config.super_resolution_enabled
    assert "enhancement_strength" in eng.state_dict()
    with pytest.raises(RuntimeError, match="CUDA"):
        LightweightSuperResolution()(torch.rand(1, 3, 8, 8))
    net = FrameRecoveryNet(base_channels=16).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.rand(1, 3, 64, 64), torch.rand(1, 2, 3, 64, 64))
    with pytest.raises(NotImplementedError):
        net.train()(torch.rand(1, 3, 64, 64), torch.rand(1, 2, 3, 64, 64))
