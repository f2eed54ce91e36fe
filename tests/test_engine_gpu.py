"""The rows either side of the SR hot path on the CUDA kernels: FrameRecoveryNet (inference),
LightweightSuperResolution (forward + backward), EnhancementEngine.forward / enhance_video -- against golden fixtures
generated from the live reference (tests/golden/make_engine_golden.py).  fp32 path rel-err <= 1e-4; bf16 path PSNR."""
import numpy as np
import pytest
import torch

from conftest import load_golden, psnr, relerr
from test_engine_oracle_cpu import build_engines, build_lightweight, build_recovery

pytestmark = pytest.mark.gpu
TOL = 1e-4


def cu(a):
    return torch.from_numpy(a).cuda()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_frame_recovery_fp32_matches_reference(tag):
    g, net = build_recovery()
    net = net.cuda()
    net.compute_dtype = torch.float32
    with torch.no_grad():
        out = net(cu(g[f"{tag}/frame"]), cu(g[f"{tag}/refs"]), cu(g[f"{tag}/mask"]))
    assert relerr(out, torch.from_numpy(g[f"{tag}/out"])) <= TOL
    with torch.no_grad():                                   # no mask: the frame comes back untouched
        assert torch.equal(net(cu(g[f"{tag}/frame"]), cu(g[f"{tag}/refs"]), None), cu(g[f"{tag}/frame"]))


def test_frame_recovery_bf16_tcgen05():
    """bf16 activations (tcgen05 convs where the shape qualifies): the recovered region within 35 dB of the fp32
    reference output (it is a tanh image in [-1, 1] pasted into the mask), everything outside the mask exact."""
    g, net = build_recovery()
    net = net.cuda()
    net.compute_dtype = torch.bfloat16
    for tag in ("a", "b"):
        frame, refs, mask = cu(g[f"{tag}/frame"]), cu(g[f"{tag}/refs"]), cu(g[f"{tag}/mask"])
        with torch.no_grad():
            out = net(frame, refs, mask)
        ref = torch.from_numpy(g[f"{tag}/out"])
        assert psnr(out, ref) >= 35.0
        keep = (mask == 0).expand_as(frame)
        assert torch.equal(out[keep], frame[keep])


def test_frame_recovery_full_size_geometry():
    """540 x 960 (BASELINE configs[4]): 34 x 60 bottleneck, decoder output 544 x 960 resized to 540 x 960, wide-Cout
    splits (464 / 1024 channels); bf16 vs fp32 path of the same weights."""
    from nerve_cl_b200.models import FrameRecoveryNet
    torch.manual_seed(3)
    net = FrameRecoveryNet().cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(4)
    frame = torch.rand(1, 3, 540, 960, device="cuda", generator=g)
    refs = torch.rand(1, 4, 3, 540, 960, device="cuda", generator=g)
    mask = torch.zeros(1, 1, 540, 960, device="cuda")
    mask[:, :, 100:400, 200:700] = 1
    outs = {}
    for dt in (torch.float32, torch.bfloat16):
        net.compute_dtype = dt
        with torch.no_grad():
            outs[dt] = net(frame, refs, mask)
    assert outs[torch.float32].shape == frame.shape
    assert psnr(outs[torch.bfloat16], outs[torch.float32]) >= 35.0


def test_lightweight_sr_forward_backward_matches_reference():
    g, net, scale = build_lightweight()
    net = net.cuda()
    net.compute_dtype = torch.float32
    x, tgt = cu(g["x"]), cu(g["target"])
    net.eval()
    with torch.no_grad():
        assert relerr(net(x), torch.from_numpy(g["out_eval"])) <= TOL          # BatchNorm-folded inference path
    assert relerr(net(x), torch.from_numpy(g["out_eval"])) <= TOL              # eval with grad: unfolded path
    net.train()
    out = net(x)
    loss = torch.nn.functional.mse_loss(out, tgt)
    loss.backward()
    assert relerr(out, torch.from_numpy(g["out_train"])) <= TOL
    assert abs(float(loss.detach()) - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    for n, p in net.named_parameters():
        ref = torch.from_numpy(g["g/" + n])
        assert relerr(p.grad, ref) <= 2 * TOL or float((p.grad.cpu() - ref).abs().max()) < 1e-9, n
    for k, v in net.state_dict().items():
        if "running" in k or "tracked" in k:
            assert relerr(v.float(), torch.from_numpy(g["bn1/" + k]).float()) <= TOL, k
    # bf16 storage: output within 40 dB of the reference output
    net.compute_dtype = torch.bfloat16
    net.eval()
    with torch.no_grad():
        assert psnr(net(x), torch.from_numpy(g["out_eval"])) >= 40.0


def test_engine_forward_matches_reference():
    g, eng, eng2 = build_engines()
    eng, eng2 = eng.cuda(), eng2.cuda()
    for m in (eng.super_resolution, eng.frame_recovery, eng2.super_resolution):
        m.compute_dtype = torch.float32
    eng.super_resolution.warp_div_mode = eng2.super_resolution.warp_div_mode = 1      # golden ran on ATen-CPU
    frames, mask = cu(g["frames"]), cu(g["mask"])
    cases = {"plain": {}, "masked": {"corruption_mask": mask}, "s06": {"corruption_mask": mask, "enhancement_strength": 0.6},
             "edge": {"center_idx": 0, "corruption_mask": mask, "enhancement_strength": 0.8}}
    with torch.no_grad():
        for tag, kw in cases.items():
            res = eng(frames, **kw)
            assert sorted(res) == list(g[f"fwd/{tag}/keys"]), tag
            assert relerr(res["enhanced"], torch.from_numpy(g[f"fwd/{tag}/enhanced"])) <= TOL, tag
        assert relerr(eng(frames, corruption_mask=mask)["recovered"], torch.from_numpy(g["fwd/masked/recovered"])) <= TOL
        # sync-free handling of an all-zero mask: 'enhanced' identical, 'recovered' == the raw centre frame
        res = eng(frames, corruption_mask=torch.zeros_like(mask))
        assert torch.equal(res["recovered"], frames[:, 2])
        assert relerr(res["enhanced"], torch.from_numpy(g["fwd/plain/enhanced"])) <= TOL
        eng.sync_free = False
        assert "recovered" not in eng(frames, corruption_mask=torch.zeros_like(mask))
        eng.sync_free = True
        assert relerr(eng2(frames)["enhanced"], torch.from_numpy(g["sronly/fwd"])) <= TOL
    # the learnable strength parameter (cached against its version counter: no per-call .item())
    with torch.no_grad():
        eng.enhancement_strength.fill_(0.6)
        assert relerr(eng(frames, corruption_mask=mask)["enhanced"], torch.from_numpy(g["fwd/s06/enhanced"])) <= TOL
        eng.enhancement_strength.fill_(1.0)
        assert relerr(eng(frames)["enhanced"], torch.from_numpy(g["fwd/plain/enhanced"])) <= TOL


@pytest.mark.parametrize("batch_size", [1, 4, 16])
def test_enhance_video_matches_reference_engine(batch_size):
    """Batched enhance_video == the reference engine's frame-by-frame loop (golden from the live reference), with and
    without corruption masks, for the full and the SR-only engine."""
    g, eng, eng2 = build_engines()
    eng, eng2 = eng.cuda(), eng2.cuda()
    for m in (eng.super_resolution, eng.frame_recovery, eng2.super_resolution):
        m.compute_dtype = torch.float32
    eng.super_resolution.warp_div_mode = eng2.super_resolution.warp_div_mode = 1
    video, masks = cu(g["video"]), cu(g["masks"])
    out = eng.enhance_video(video, masks, batch_size=batch_size)
    assert out.shape == (7, 3, 128, 128)
    assert relerr(out[..., ::2, ::2], torch.from_numpy(g["video_out"])) <= TOL
    assert relerr(eng.enhance_video(video, batch_size=batch_size)[..., ::2, ::2], torch.from_numpy(g["video_out_nomask"])) <= TOL
    assert relerr(eng2.enhance_video(video, batch_size=batch_size)[..., ::2, ::2], torch.from_numpy(g["sronly/video_out"])) <= TOL
    # (B, T, C, H, W) input: clips are independent
    two = torch.stack([video, video.flip(0)])
    o2 = eng2.enhance_video(two, batch_size=batch_size)
    assert relerr(o2[0][..., ::2, ::2], torch.from_numpy(g["sronly/video_out"])) <= TOL
    # the free-function SR-only path (nerve_cl_b200.inference.enhance_video) gives the same video
    from nerve_cl_b200.inference import enhance_video
    o3 = enhance_video(eng2.super_resolution, video, batch_size=batch_size)
    assert relerr(o3[..., ::2, ::2], torch.from_numpy(g["sronly/video_out"])) <= TOL


def test_enhance_video_sr_only_path_equals_the_window_loop():
    """The SR-only fast path (all frames' windows batched uniformly) == the reference's loop structure (one forward per
    frame on its clipped window, `enhancement_engine.py:214-240`), including the strength blend and the lightweight net."""
    from nerve_cl_b200.models import EnhancementConfig, EnhancementEngine
    torch.manual_seed(5)
    video = torch.rand(9, 3, 24, 40, device="cuda")
    for light in (False, True):
        eng = EnhancementEngine(EnhancementConfig(frame_recovery_enabled=False, sr_num_features=16, sr_num_residual_blocks=1,
                                                  use_lightweight_sr=light)).cuda().eval()
        eng.super_resolution.compute_dtype = torch.float32
        for strength in (1.0, 0.6):
            with torch.no_grad():
                eng.enhancement_strength.fill_(strength)
                fast = eng.enhance_video(video, batch_size=4)
                loop = torch.stack([eng(video[a:b].unsqueeze(0), center_idx=c)["enhanced"][0]
                                    for a, b, c in eng.window_table(video.shape[0])])
            assert fast.shape == loop.shape
            assert relerr(fast, loop) <= 1e-6, (light, strength)


def test_engine_trains_through_the_sr_path():
    """train_continual.py's configuration: SR-only engine, loss on results['enhanced'], gradients reach the SR
    parameters (and only them), with and without the strength blend."""
    from nerve_cl_b200.models import EnhancementConfig, EnhancementEngine
    torch.manual_seed(9)
    eng = EnhancementEngine(EnhancementConfig(frame_recovery_enabled=False, sr_num_features=16, sr_num_residual_blocks=1)).cuda().train()
    eng.super_resolution.compute_dtype = torch.float32
    x, t = torch.rand(2, 3, 3, 16, 24, device="cuda"), torch.rand(2, 3, 32, 48, device="cuda")
    torch.nn.functional.mse_loss(eng(x)["enhanced"], t).backward()
    g1 = {n: p.grad.clone() for n, p in eng.named_parameters() if p.grad is not None}
    assert "enhancement_strength" not in g1 and len(g1) == len(list(eng.super_resolution.parameters()))
    eng.zero_grad()
    out = eng(x, enhancement_strength=0.5)["enhanced"]
    torch.nn.functional.mse_loss(out, t).backward()
    sr = eng.super_resolution(x).detach()
    bic = torch.nn.functional.interpolate(x[:, 1], scale_factor=2, mode="bicubic", align_corners=False)
    assert relerr(out, 0.5 * sr + 0.5 * bic) <= 1e-5
