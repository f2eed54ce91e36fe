"""Per-op parity of the CUDA kernels (called through torch.ops.nervecl -> C ABI) against the oracle /
plain ATen fp32 on the same seeded inputs.  fp32 tolerance: rel-err <= 1e-4 (BASELINE.json)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-4


def nv():
    from nerve_cl_b200 import ops
    return ops.nv


def nhwc(x, dtype=torch.float32, pad_to=0):
    """NCHW cpu/cuda tensor -> pitched NHWC cuda view."""
    n, c, h, w = x.shape
    buf = torch.zeros((n, h, w, max(c, pad_to)), device="cuda", dtype=dtype)
    buf[..., :c] = x.permute(0, 2, 3, 1).to("cuda", dtype)
    return buf[..., :c]


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def pack(w, dtype=torch.float32, flip=False):
    o, i, k, _ = w.shape
    rows, cols = (i, o) if flip else (o, i)
    dst = torch.empty((k * k, rows, (cols + 7) // 8 * 8), device="cuda", dtype=dtype)
    nv().pack_conv_weight(w.cuda().contiguous(), dst, flip)
    return dst


CONV_SHAPES = [  # (N, H, W, Cin, Cout, K)
    (2, 10, 12, 3, 16, 3), (1, 32, 32, 2, 32, 3), (1, 17, 45, 32, 2, 3), (2, 9, 33, 81, 128, 3),
    (1, 12, 40, 64, 32, 3), (1, 8, 32, 224, 64, 1), (1, 20, 20, 2, 1, 7), (1, 33, 65, 64, 12, 3),
    (1, 16, 64, 96, 32, 3), (1, 40, 24, 64, 3, 3),
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_fwd_simt_epilogue(shape):
    from nerve_cl_b200 import ops
    n, h, w, cin, cout, k = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    res = torch.randn(n, cout, h, w, generator=g)
    ref = F.relu(F.conv2d(x, wt, b, 1, k // 2)) * 0.2 + res
    out = torch.empty((n, h, w, cout), device="cuda")
    nv().conv2d_fwd(nhwc(x, pad_to=cin + 5), pack(wt), b.cuda(), nhwc(res), None, None, out, cout, True, False,
                    cout, 0, 0.2, ops.CONV_SIMT)
    assert relerr(nchw(out), ref) <= TOL


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_dgrad_accumulate_mask(shape):
    """conv2d_fwd on dY with transpose_flip weights == ATen's input gradient; accumulate + ReLU mask."""
    from nerve_cl_b200 import ops
    n, h, w, cin, cout, k = shape
    g = torch.Generator().manual_seed(sum(shape) + 1)
    x = torch.randn(n, cin, h, w, generator=g, requires_grad=True)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    dy = torch.randn(n, cout, h, w, generator=g)
    prev = torch.randn(n, cin, h, w, generator=g)
    act = torch.randn(n, cin, h, w, generator=g)
    (dx,) = torch.autograd.grad(F.conv2d(x, wt, None, 1, k // 2), x, dy)
    c0 = cin // 2
    ref = dx + prev
    ref[:, c0:] = ref[:, c0:] * (act[:, c0:] > 0)
    out = nhwc(prev, pad_to=cin + 3)
    nv().conv2d_fwd(nhwc(dy), pack(wt, flip=True), None, None, nhwc(act), None, out, cin, False, True, 0, c0, 1.0,
                    ops.CONV_SIMT)
    assert relerr(nchw(out), ref) <= TOL


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_wgrad_simt(shape):
    from nerve_cl_b200 import ops
    n, h, w, cin, cout, k = shape
    g = torch.Generator().manual_seed(sum(shape) + 2)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.zeros(cout, cin, k, k, requires_grad=True)
    b = torch.zeros(cout, requires_grad=True)
    dy = torch.randn(n, cout, h, w, generator=g)
    F.conv2d(x, wt, b, 1, k // 2).backward(dy)
    dw = torch.zeros_like(wt, device="cuda").detach()
    db = torch.zeros(cout, device="cuda")
    nv().conv2d_wgrad(nhwc(x, pad_to=cin + 4), nhwc(dy), dw, db, 0.5, ops.CONV_SIMT)
    assert relerr(dw, 0.5 * wt.grad) <= TOL
    assert relerr(db, 0.5 * b.grad) <= TOL


@pytest.mark.parametrize("shape", [(2, 64, 13, 40), (3, 64, 50, 70), (1, 32, 9, 130)])
def test_depthwise_masked(shape):
    """y = (dwconv3x3(x, flipped filter) + add) * (mask > 0): the extractor's first-layer data gradient + skip gradient +
    head ReLU in one pass == the three separate ATen ops (ragged tiles in both directions)."""
    n, c, h, w = shape
    g = torch.Generator().manual_seed(sum(shape))
    x, add, mask = (bf(torch.randn(n, c, h, w, generator=g)) for _ in range(3))
    wt = torch.randn(c, 1, 3, 3, generator=g)
    for flip in (False, True):
        ref = (F.conv2d(x, wt.flip(2, 3) if flip else wt, None, 1, 1, 1, c) + add) * (mask > 0)
        y = torch.full((n, h, w, c + 8), 7.0, device="cuda", dtype=torch.bfloat16)
        nv().dwconv3x3_fwd_masked(nhwc(x, torch.bfloat16), wt.cuda(), nhwc(add, torch.bfloat16, pad_to=c + 8),
                                  nhwc(mask, torch.bfloat16), y[..., :c], flip)
        assert relerr(nchw(y[..., :c]), ref) <= 2e-2
        assert float((y[..., c:].float() - 7).abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_depthwise(dtype):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 16, 13, 19, generator=g)
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    wt = torch.randn(16, 1, 3, 3, generator=g, requires_grad=True)
    xr = x.clone().requires_grad_(True)
    y = F.conv2d(xr, wt, None, 1, 1, 1, 16)
    dy = torch.randn_like(y)
    if dtype == torch.bfloat16:
        dy = dy.bfloat16().float()
    y.backward(dy)
    tol = TOL if dtype == torch.float32 else 2e-2
    xo = nhwc(x, dtype)
    yo = torch.empty_like(xo)
    nv().dwconv3x3_fwd(xo, wt.detach().cuda(), yo, False, False)
    assert relerr(nchw(yo), y) <= tol
    dxo = torch.empty_like(xo)
    nv().dwconv3x3_fwd(nhwc(dy, dtype), wt.detach().cuda(), dxo, True, False)
    assert relerr(nchw(dxo), xr.grad) <= tol
    dw = torch.zeros(16, 1, 3, 3, device="cuda")
    nv().dwconv3x3_wgrad(xo, nhwc(dy, dtype), dw)
    assert relerr(dw, wt.grad) <= tol


@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_relu_groups(training):
    """T frame groups with separate batch statistics + running-stat update in group order."""
    T, B, C, H, W = 3, 2, 16, 9, 11
    g = torch.Generator().manual_seed(4)
    x = torch.randn(T * B, C, H, W, generator=g) * 2 + 0.5
    gamma = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.1).requires_grad_(True)
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    res = torch.randn(T * B, C, H, W, generator=g)
    dy = torch.randn(T * B, C, H, W, generator=g)
    xr = x.clone().requires_grad_(True)
    rm_r, rv_r = rm.clone(), rv.clone()
    ys = [F.relu(F.batch_norm(xr[t * B:(t + 1) * B], rm_r, rv_r, gamma, beta, training, 0.1, 1e-5)) for t in range(T)]
    y = torch.cat(ys) + res
    y.backward(dy)

    xo = nhwc(x)
    sums = torch.zeros(T, C, 2, device="cuda", dtype=torch.float64)
    stat = torch.empty(T, C, 2, device="cuda")
    rmc, rvc = rm.cuda(), rv.cuda()
    nbt = torch.zeros((), device="cuda", dtype=torch.int64)
    if training:
        nv().bn_stats(xo, T, sums)
    nv().bn_finalize(sums if training else None, stat, rmc, rvc, nbt, B * H * W, T, 0.1, 1e-5, training)
    yo = torch.empty_like(xo)
    nv().bn_relu_fwd(xo, stat, gamma.detach().cuda(), beta.detach().cuda(), nhwc(res), yo, T)
    assert relerr(nchw(yo), y) <= TOL
    assert relerr(rmc, rm_r) <= TOL and relerr(rvc, rv_r) <= TOL
    assert int(nbt) == (T if training else 0)
    bs = torch.zeros(T, C, 2, device="cuda", dtype=torch.float64)
    dyo = nhwc(dy)
    nv().bn_relu_bwd_reduce(xo, dyo, stat, gamma.detach().cuda(), beta.detach().cuda(), T, bs)
    dx = torch.empty_like(xo)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    nv().bn_relu_bwd_apply(xo, dyo, stat, gamma.detach().cuda(), beta.detach().cuda(), bs, dx, dg, db, T, training)
    assert relerr(nchw(dx), xr.grad) <= TOL
    assert relerr(dg, gamma.grad) <= TOL and relerr(db, beta.grad) <= TOL


def test_correlation_golden_and_grad():
    from oracle import sr_oracle
    gold = load_golden("corr_case.npz")
    x1, x2 = torch.from_numpy(gold["x1"]), torch.from_numpy(gold["x2"])
    out = torch.empty((2, 11, 14, 96), device="cuda")
    nv().corr_fwd(nhwc(x1), nhwc(x2), out)
    assert relerr(nchw(out[..., :81]), torch.from_numpy(gold["out"])) <= TOL
    assert float(out[..., 81:].abs().max()) == 0.0
    a, b = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    dy = torch.randn(2, 81, 11, 14, generator=torch.Generator().manual_seed(8))
    sr_oracle.correlation(a, b).backward(dy)
    d1, d2 = torch.empty((2, 11, 14, 8), device="cuda"), torch.empty((2, 11, 14, 8), device="cuda")
    nv().corr_bwd(nhwc(x1), nhwc(x2), nhwc(dy, pad_to=96), d1, False, d2, False)
    assert relerr(nchw(d1), a.grad) <= TOL and relerr(nchw(d2), b.grad) <= TOL


@pytest.mark.parametrize("shape", [(2, 37, 150), (1, 9, 64), (1, 50, 70)])
def test_correlation_tiled_bf16(shape):
    """bf16 / C=64 shared-memory kernels (forward, both gradients, accumulate flags, ragged strips and row
    segments) against the oracle on bf16-rounded inputs."""
    from oracle import sr_oracle
    n, h, w = shape
    g = torch.Generator().manual_seed(sum(shape))
    x1 = torch.randn(n, 64, h, w, generator=g).bfloat16().float()
    x2 = torch.randn(n, 64, h, w, generator=g).bfloat16().float()
    a, b = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    ref = sr_oracle.correlation(a, b)
    out = torch.full((n, h, w, 104), 7.0, device="cuda", dtype=torch.bfloat16)
    nv().corr_fwd(nhwc(x1, torch.bfloat16, pad_to=72), nhwc(x2, torch.bfloat16), out[..., :96])
    assert relerr(nchw(out[..., :81]), ref) <= 6e-3
    assert float(out[..., 81:96].float().abs().max()) == 0.0 and float((out[..., 96:].float() - 7).abs().max()) == 0.0
    dy = torch.randn(n, 81, h, w, generator=g).bfloat16().float()
    ref.backward(dy)
    p1 = torch.randn(n, 64, h, w, generator=g).bfloat16().float()
    d1 = nhwc(p1, torch.bfloat16, pad_to=72)
    d2 = torch.empty((n, h, w, 64), device="cuda", dtype=torch.bfloat16)
    nv().corr_bwd(nhwc(x1, torch.bfloat16), nhwc(x2, torch.bfloat16), nhwc(dy, torch.bfloat16, pad_to=96), d1, True,
                  d2, False)
    assert relerr(nchw(d1), a.grad + p1) <= 8e-3 and relerr(nchw(d2), b.grad) <= 8e-3


@pytest.mark.parametrize("shape", [(2, 37, 150), (1, 8, 16), (1, 50, 70), (3, 64, 128), (1, 21, 333)])
def test_correlation_tcgen05_tiles(shape):
    """The tcgen05 2-D tile correlation kernels (csrc/motion_tc.cu: dense Gram product per 8 x 16 tile, diagonal
    extraction through shared memory; gradients as one GEMM per tile over the scattered gradient operand, the second
    operand's gradient through the transposed-gradient workspace): ragged tiles in both directions, accumulate flags."""
    from oracle import sr_oracle
    n, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 3)
    x1 = torch.randn(n, 64, h, w, generator=g).bfloat16().float()
    x2 = torch.randn(n, 64, h, w, generator=g).bfloat16().float()
    a, b = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    ref = sr_oracle.correlation(a, b)
    out = torch.full((n, h, w, 96), 7.0, device="cuda", dtype=torch.bfloat16)
    nv().corr_fwd(nhwc(x1, torch.bfloat16, pad_to=72), nhwc(x2, torch.bfloat16), out)
    assert relerr(nchw(out[..., :81]), ref) <= 6e-3
    assert float(out[..., 81:].float().abs().max()) == 0.0
    dy = torch.randn(n, 81, h, w, generator=g).bfloat16().float()
    ref.backward(dy)
    p1 = torch.randn(n, 64, h, w, generator=g).bfloat16().float()
    d1 = nhwc(p1, torch.bfloat16)
    d2 = torch.full((n, h, w, 64), 3.0, device="cuda", dtype=torch.bfloat16)
    ws = torch.empty((n, h, w, 96), device="cuda", dtype=torch.bfloat16)
    gy = nhwc(dy, torch.bfloat16, pad_to=96)
    nv().corr_bwd(nhwc(x1, torch.bfloat16), nhwc(x2, torch.bfloat16), gy, d1, True, d2, False, ws)
    assert relerr(nchw(d1), a.grad + p1) <= 8e-3 and relerr(nchw(d2), b.grad) <= 8e-3
    # the workspace path and the banded mma.sync path (no workspace) agree
    e1, e2 = torch.empty_like(d2), torch.empty_like(d2)
    nv().corr_bwd(nhwc(x1, torch.bfloat16), nhwc(x2, torch.bfloat16), gy, e1, False, e2, False)
    assert relerr(e2, d2.float()) <= 8e-3 and relerr(nchw(e1), a.grad) <= 8e-3


def test_warp_indices_bit_exact_vs_cpu_golden():
    """div_mode=1 replays ATen-CPU's division: indices must equal the golden (CPU reference) ones exactly."""
    from nerve_cl_b200.models import warp_indices, warp_features
    gold = load_golden("warp_cases.npz")
    for hw in ("9x13", "36x64"):
        feat = torch.from_numpy(gold[f"{hw}/feat"]).cuda()
        feat8 = torch.cat([feat, feat], 1)                      # kernels need C % 8 == 0
        for i in range(10):
            flow = torch.from_numpy(gold[f"{hw}/{i}/flow"]).cuda()
            idx = warp_indices(flow, div_mode=1).cpu().numpy()
            want = gold[f"{hw}/{i}/idx"].astype(np.int64)
            # indices far outside the image all mean "zero padding"; compare after clamping to [-2, size+1]
            h, w = [int(v) for v in hw.split("x")]
            lo, hi = np.array([-2, -2]), np.array([w + 1, h + 1])
            assert np.array_equal(np.clip(idx, lo, hi), np.clip(want, lo, hi)), (hw, i)
            out = warp_features(feat8, flow, div_mode=1)
            ref = torch.from_numpy(gold[f"{hw}/{i}/out"])
            assert relerr(out[:, :4], ref) <= TOL or float((out[:, :4].cpu() - ref).abs().max()) < 1e-6, (hw, i)


def structured_flow(b, h, w):
    ys = torch.arange(h).view(1, h, 1).expand(b, h, w)
    xs = torch.arange(w).view(1, 1, w).expand(b, h, w)
    fx = ((xs * 7 + ys * 3) % 33 - 16).float() * 0.25 + 2.0 ** -12
    fy = ((xs * 5 + ys * 11) % 29 - 14).float() * 0.125 - 2.0 ** -11
    return torch.stack([fx, fy], 1)


def test_warp_indices_full_size_cpu_golden_and_live_cuda():
    """360x640 (BASELINE cfg 2 size): bit-exact vs the CPU-reference golden (div_mode=1) AND vs the oracle
    run by ATen-CUDA on this GPU (div_mode=0, the reciprocal-multiply path)."""
    from oracle import sr_oracle
    from nerve_cl_b200.models import warp_indices
    gold = load_golden("warp_cases.npz")
    h, w = 360, 640
    lo, hi = np.array([-2, -2]), np.array([w + 1, h + 1])
    flows = {"zero": torch.zeros(1, 2, h, w), "structured": structured_flow(1, h, w)}
    for name, flow in flows.items():
        got = warp_indices(flow.cuda(), div_mode=1).cpu().numpy()
        assert np.array_equal(np.clip(got, lo, hi), np.clip(gold[f"{h}x{w}/{name}/idx"].astype(np.int64), lo, hi)), name
    g = torch.Generator().manual_seed(77)
    rnd = 3 * torch.randn(2, 2, h, w, generator=g)
    for flow in list(flows.values()) + [rnd, rnd.round(), rnd.round() + 0.5]:
        live = sr_oracle.warp_corner_indices(flow.cuda()).cpu().numpy()          # ATen-CUDA
        got = warp_indices(flow.cuda(), div_mode=0).cpu().numpy()
        assert np.array_equal(np.clip(got, lo, hi), np.clip(live, lo, hi))


def test_warp_backward():
    from oracle import sr_oracle
    from nerve_cl_b200.models import warp_features
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(2, 16, 12, 15, generator=g)
    flow = 2.5 * torch.randn(2, 2, 12, 15, generator=g) + 0.013
    dy = torch.randn(2, 16, 12, 15, generator=g)
    a, b = feat.clone().requires_grad_(True), flow.clone().requires_grad_(True)
    sr_oracle.warp(a, b).backward(dy)
    fa, fb = feat.cuda().requires_grad_(True), flow.cuda().requires_grad_(True)
    warp_features(fa, fb, div_mode=1).backward(dy.cuda())
    assert relerr(fa.grad, a.grad) <= TOL
    assert relerr(fb.grad, b.grad) <= TOL


def test_warp_backward_bf16_packed_reductions():
    """warp_bwd_lp scatters into a bf16 gradient with packed 8 x bf16 reductions: same result as the fp32 scatter
    of the same bf16 operands up to bf16 rounding of the (<= ~8) partial sums per element."""
    g = torch.Generator().manual_seed(19)
    n, c, h, w = 2, 64, 21, 70
    feat = nhwc(torch.randn(n, c, h, w, generator=g), torch.bfloat16)
    dy = nhwc(torch.randn(n, c, h, w, generator=g), torch.bfloat16)
    flow = (1.7 * torch.randn(n, h, w, 2, generator=g) + 0.013).cuda()
    d32 = torch.zeros((n, h, w, c), device="cuda")
    dfl32 = torch.zeros_like(flow)
    nv().warp_bwd(feat, flow, dy, d32, dfl32, 0)
    d16 = torch.zeros((n, h, w, c), device="cuda", dtype=torch.bfloat16)
    dfl16 = torch.zeros_like(flow)
    nv().warp_bwd_lp(feat, flow, dy, d16, dfl16, 0)
    assert relerr(d16.float(), d32) <= 1e-2
    assert torch.equal(dfl16, dfl32)


@pytest.mark.parametrize("shape", [(3, 2, 16, 7, 9), (5, 1, 64, 21, 37), (1, 2, 8, 3, 5), (8, 1, 32, 9, 11)])
def test_temporal_fusion(shape):
    T, B, C, H, W = shape
    g = torch.Generator().manual_seed(10)
    feats = torch.randn(B, T, C, H, W, generator=g, requires_grad=True)
    logits = torch.randn(B, T, H, W, generator=g, requires_grad=True)
    bias = torch.randn(B, C, generator=g)
    dy = torch.randn(B, C, H, W, generator=g)
    attn = torch.softmax(logits, 1)
    out = (feats * attn.unsqueeze(2)).sum(1)
    out.backward(dy + bias[:, :, None, None])
    cat = nhwc(feats.detach().reshape(B, T * C, H, W))
    lg = logits.detach().permute(0, 2, 3, 1).contiguous().cuda()
    at = torch.empty_like(lg)
    o = torch.empty((B, H, W, C), device="cuda")
    nv().tfuse_fwd(cat, lg, at, o)
    assert relerr(nchw(o), out) <= TOL and relerr(at.permute(0, 3, 1, 2), attn) <= TOL
    dcat, dlg = torch.empty_like(cat), torch.empty_like(lg)
    nv().tfuse_bwd(cat, at, nhwc(dy), bias.cuda(), dcat, dlg)
    assert relerr(nchw(dcat), feats.grad.reshape(B, T * C, H, W)) <= TOL
    assert relerr(dlg.permute(0, 3, 1, 2), logits.grad) <= TOL


@pytest.mark.parametrize("shape", [(2, 32, 11, 13), (1, 64, 19, 45), (2, 8, 33, 70)])
def test_cbam_forward_backward(shape):
    from oracle import sr_oracle
    B, C, H, W = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, C, H, W, generator=g, requires_grad=True)
    sd = {"temporal_aggregator.refine.channel_attention.fc.0.weight": torch.randn(2, C, generator=g).requires_grad_(True),
          "temporal_aggregator.refine.channel_attention.fc.2.weight": torch.randn(C, 2, generator=g).requires_grad_(True),
          "temporal_aggregator.refine.spatial_attention.conv.weight": (0.2 * torch.randn(1, 2, 7, 7, generator=g)).requires_grad_(True)}
    w1, w2, w7 = sd.values()
    dy = torch.randn(B, C, H, W, generator=g)
    y = sr_oracle.spatial_attention(sd, sr_oracle.channel_attention(sd, x))
    y.backward(dy)

    xo = nhwc(x.detach())
    pool = torch.zeros(B, C, device="cuda")
    nv().chan_sum(xo, 1.0 / (H * W), pool)
    hidden, gate = torch.empty(B, 2, device="cuda"), torch.empty(B, C, device="cuda")
    w1c, w2c, w7c = w1.detach().cuda(), w2.detach().cuda(), w7.detach().cuda()
    nv().ca_gate_fwd(pool, w1c, w2c, hidden, gate)
    stats = torch.empty(B, H, W, 2, device="cuda")
    nv().cbam_stats_fwd(xo, gate, stats)
    sg, out = torch.empty(B, H, W, device="cuda"), torch.empty_like(xo)
    nv().cbam_apply_fwd(xo, gate, stats, w7c, sg, out)
    assert relerr(nchw(out), y) <= TOL
    dz = torch.empty(B, H, W, device="cuda")
    dyo = nhwc(dy)
    nv().cbam_bwd_dz(xo, gate, sg, dyo, dz)
    dstats, dw7 = torch.empty_like(stats), torch.zeros_like(w7c)
    nv().cbam_bwd_spatial(dz, stats, w7c, dstats, dw7)
    dx, dgate = torch.empty_like(xo), torch.zeros(B, C, device="cuda")
    nv().cbam_bwd_dx(xo, gate, sg, stats, dstats, dyo, dx, dgate)
    dpool, dw1, dw2 = torch.empty(B, C, device="cuda"), torch.zeros_like(w1c), torch.zeros_like(w2c)
    nv().ca_gate_bwd(pool, w1c, w2c, hidden, gate, dgate, dpool, dw1, dw2)
    full_dx = nchw(dx) + (dpool / (H * W))[:, :, None, None]
    assert relerr(full_dx, x.grad) <= TOL
    assert relerr(dw1, w1.grad) <= TOL and relerr(dw2, w2.grad) <= TOL and relerr(dw7, w7.grad) <= TOL


@pytest.mark.parametrize("shape", [(2, 9, 12), (1, 37, 70), (3, 5, 131)])
@pytest.mark.parametrize("s", [2, 3, 4])
def test_output_stage(s, shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(12 + s)
    conv = (0.3 * torch.randn(B, 3 * s * s, H, W, generator=g)).requires_grad_(True)
    lr = torch.rand(B, 5, 3, H, W, generator=g)[:, 2]            # strided view like lr_frames[:, mid]
    dy = torch.randn(B, 3, H * s, W * s, generator=g)
    ref = torch.clamp(F.interpolate(lr, scale_factor=s, mode="bicubic", align_corners=False)
                      + F.pixel_shuffle(conv, s), 0, 1)
    ref.backward(dy)
    co = conv.detach().permute(0, 2, 3, 1).contiguous().cuda()
    lrc = torch.rand(B, 5, 3, H, W, device="cuda")
    lrc[:, 2] = lr.cuda()
    out = torch.empty(B, 3, H * s, W * s, device="cuda")
    nv().upfinish_fwd(co, lrc[:, 2], out, s)
    assert relerr(out, ref) <= TOL
    dconv = torch.empty_like(co)
    nv().upfinish_bwd(co, lrc[:, 2], dy.cuda(), dconv, s)
    assert relerr(dconv.permute(0, 3, 1, 2), conv.grad) <= TOL


def test_elementwise_helpers():
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, 16, 5, 7, generator=g)
    y = torch.randn(2, 16, 5, 7, generator=g)
    o = nhwc(y, torch.bfloat16, pad_to=24)
    nv().axpy(nhwc(x), o, 0.5, True)
    assert relerr(nchw(o), y.bfloat16().float() + 0.5 * x) <= 1e-2
    o3 = torch.zeros(2, 5, 7, 8, device="cuda", dtype=torch.bfloat16)
    x3 = torch.randn(2, 3, 5, 7, generator=g)
    nv().axpy(nhwc(x3), o3[..., :3], 1.0, False)                 # scalar path (C % 4 != 0)
    assert relerr(nchw(o3[..., :3]), x3) <= 1e-2
    # 8-wide path: C % 8 == 0, strided source and destination rows, copy and accumulate, bf16 and fp32
    x8, y8 = torch.randn(3, 64, 9, 21, generator=g), torch.randn(3, 64, 9, 21, generator=g)
    for dt, tol in ((torch.float32, 1e-6), (torch.bfloat16, 1e-2)):
        sbuf = torch.zeros((3, 9, 21, 192), device="cuda", dtype=dt)
        dbuf = torch.zeros((3, 9, 21, 80), device="cuda", dtype=dt)
        sbuf[..., 64:128] = x8.permute(0, 2, 3, 1).to("cuda", dt)
        dbuf[..., :64] = y8.permute(0, 2, 3, 1).to("cuda", dt)
        src, dst = sbuf[..., 64:128], dbuf[..., :64]
        ref_y = nchw(dst).clone()
        nv().axpy(src, dst, -1.5, True)
        assert relerr(nchw(dst), ref_y - 1.5 * nchw(src)) <= tol
        nv().axpy(src, dst, 1.0, False)
        assert torch.equal(dst, src) and float(dbuf[..., 64:].abs().max()) == 0.0
    out = torch.empty(2, 5, 7, 16, device="cuda")
    nv().relu_bwd(nhwc(x), nhwc(y), None, out)
    assert torch.equal(nchw(out).cpu(), x * (y > 0))
    a, b = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    loss, dg = torch.zeros(1, device="cuda"), torch.empty(1000, device="cuda")
    nv().mse_fwd_bwd(a.cuda(), b.cuda(), dg, loss, 1.0 / 1000)
    assert abs(float(loss) - float(F.mse_loss(a, b))) <= 1e-6
    assert relerr(dg, 2 * (a - b) / 1000) <= 1e-6


def test_ops_reject_cpu_tensors():
    with pytest.raises((RuntimeError, NotImplementedError)):
        nv().fill_zero(torch.zeros(4))


# ---------------------------------------------------------------------------------------------
# tcgen05 implicit-GEMM engine (bf16): same contract as the SIMT engine, checked against fp32 ATen on
# bf16-rounded operands (error budget: fp32 accumulation order + one bf16 rounding of the output)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pack_frames_unfold3(dtype):
    """3x3 unfold of strided (B,T,C,H,W) frames into channels c*9 + tap (== the OIHW filter flattening)."""
    g = torch.Generator().manual_seed(3)
    b, t, c, h, w = 2, 3, 3, 13, 37
    big = torch.randn(b, t, c, h + 2, w + 5, generator=g).cuda()
    src = big[:, :, :, 1:h + 1, 2:w + 2]                       # non-contiguous rows / frames
    dst = torch.full((t * b, h, w, 32), 9.0, device="cuda", dtype=dtype)
    nv().pack_frames_unfold3(src, dst)
    ref = F.unfold(src.permute(1, 0, 2, 3, 4).reshape(t * b, c, h, w), 3, padding=1)      # [TB, c*9, h*w]
    ref = ref.view(t * b, 27, h, w).permute(0, 2, 3, 1)
    if dtype == torch.bfloat16:
        ref = ref.bfloat16()
    assert torch.equal(dst[..., :27].float(), ref.float())
    assert float(dst[..., 27:].float().abs().max()) == 0.0


@pytest.mark.parametrize("c", [2, 3])
def test_unfold3_grad_gives_small_cout_weight_gradient(c):
    """dW of a 3x3 conv with 2-3 output channels == x^T unfold3_grad(dy), bias gradient == its centre-tap columns."""
    g = torch.Generator().manual_seed(30 + c)
    n, cin, h, w = 2, 8, 11, 19
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(c, cin, 3, 3, generator=g, requires_grad=True)
    dy = torch.randn(n, c, h, w, generator=g)
    F.conv2d(x, wt, None, 1, 1).backward(dy)
    src = nhwc(dy, torch.float32, pad_to=16)[..., :c]
    unf = torch.full((n, h, w, 32), 5.0, device="cuda")
    nv().unfold3_grad(src, unf)
    xm = nhwc(x).reshape(-1, cin)
    dw1 = unf.reshape(-1, 32).t() @ xm                                   # [o*9+tap, cin]
    dw = dw1[:c * 9].view(c, 9, cin).transpose(1, 2).reshape(c, cin, 3, 3)
    assert relerr(dw, wt.grad) <= TOL
    assert relerr(unf.reshape(-1, 32).sum(0)[:c * 9].view(c, 9)[:, 4], dy.sum((0, 2, 3))) <= TOL
    assert float(unf[..., c * 9:].abs().max()) == 0.0


TC_SHAPES = [  # (N, H, W, Cin, Cout, K)
    (1, 16, 128, 64, 32, 3), (2, 20, 72, 96, 32, 3), (1, 9, 200, 192, 32, 3), (1, 24, 40, 64, 64, 3),
    (1, 16, 64, 224, 64, 1), (1, 12, 136, 32, 192, 3), (2, 8, 16, 128, 64, 3), (1, 33, 65, 96, 128, 3),
    (1, 30, 257, 160, 32, 3), (3, 11, 23, 64, 64, 1), (1, 16, 32, 32, 224, 1),
    # row-streaming kernel: merged N=192, two chunks, per-ky N=96 with a 5-slot ring, channel-group split,
    # 1/2-row tail segments, TMEM ring wrap (> 16 rows per item)
    (1, 20, 130, 64, 64, 3), (2, 7, 300, 128, 64, 3), (1, 5, 64, 32, 96, 3), (1, 40, 128, 192, 64, 3),
    (1, 19, 140, 96, 128, 3), (1, 64, 128, 64, 32, 3), (1, 3, 640, 64, 48, 3),
    # 1x1 through the row-streaming kernel (centre tap only, zero-filled ky = 0 / 2 weight blocks)
    (1, 20, 130, 64, 64, 1), (2, 9, 200, 64, 32, 1), (1, 16, 128, 128, 64, 1), (1, 12, 640, 32, 48, 1),
    # the benchmark's own geometry (cfg 2: 360 x 640 -> 5 column strips, 360-row items, accumulator-ring wrap)
    (1, 360, 640, 64, 32, 3), (1, 360, 640, 160, 32, 3), (2, 360, 640, 64, 64, 1),
]
BF16_TOL = 6e-3


def bf(x):
    return x.bfloat16().float()


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_fwd_tcgen05_epilogue(shape):
    from nerve_cl_b200 import ops
    n, h, w, cin, cout, k = shape
    g = torch.Generator().manual_seed(sum(shape) + 7)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = bf(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, generator=g)
    res = bf(torch.randn(n, cout, h, w, generator=g))
    ref = F.relu(F.conv2d(x, wt, b, 1, k // 2)) * 0.2 + res
    out = torch.full((n, h, w, cout + 8), 7.0, device="cuda", dtype=torch.bfloat16)
    nv().conv2d_fwd(nhwc(x, torch.bfloat16, pad_to=cin + 8), pack(wt, torch.bfloat16), b.cuda(),
                    nhwc(res, torch.bfloat16), None, None, out[..., :cout], cout, True, False, cout, 0, 0.2,
                    ops.CONV_TC)
    assert relerr(nchw(out[..., :cout]), ref) <= BF16_TOL
    assert float((out[..., cout:].float() - 7.0).abs().max()) == 0.0          # never writes outside its slice


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_dgrad_tcgen05_accumulate_mask(shape):
    from nerve_cl_b200 import ops
    n, h, w, cout, cin, k = shape          # roles swapped: dY has `cout` channels, dX has `cin`
    g = torch.Generator().manual_seed(sum(shape) + 8)
    x = torch.randn(n, cin, h, w, generator=g, requires_grad=True)
    wt = bf(torch.randn(cout, cin, k, k, generator=g) / (cout * k * k) ** 0.5)
    dy = bf(torch.randn(n, cout, h, w, generator=g))
    prev = bf(torch.randn(n, cin, h, w, generator=g))
    act = bf(torch.randn(n, cin, h, w, generator=g))
    (dx,) = torch.autograd.grad(F.conv2d(x, wt, None, 1, k // 2), x, dy)
    c0 = (cin // 2) // 8 * 8 + 3
    ref = 0.5 * dx + prev
    ref[:, c0:] = ref[:, c0:] * (act[:, c0:] > 0)
    out = nhwc(prev, torch.bfloat16, pad_to=cin + 8)
    nv().conv2d_fwd(nhwc(dy, torch.bfloat16), pack(wt, torch.bfloat16, flip=True), None, None,
                    nhwc(act, torch.bfloat16), None, out, cin, False, True, 0, c0, 0.5, ops.CONV_TC)
    assert relerr(nchw(out), ref) <= BF16_TOL


@pytest.mark.parametrize("k", [1, 3])
@pytest.mark.parametrize("cout", [32, 64, 96, 128, 192])
def test_conv_rows_lean_epilogues(k, cout):
    """The row kernel's lean epilogues (bias + ReLU -> bf16; alpha * acc gated by the ReLU mask) for 3x3 and 1x1."""
    from nerve_cl_b200 import ops
    n, h, w, cin = 2, 21, 200, 64
    g = torch.Generator().manual_seed(5 * k + cout)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = bf(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, generator=g)
    act = bf(torch.randn(n, cout, h, w, generator=g))
    xo, wp = nhwc(x, torch.bfloat16), pack(wt, torch.bfloat16)
    out = torch.full((n, h, w, cout + 8), 7.0, device="cuda", dtype=torch.bfloat16)
    nv().conv2d_fwd(xo, wp, b.cuda(), None, None, None, out[..., :cout], cout, True, False, 0, 0, 1.0, ops.CONV_TC)
    assert relerr(nchw(out[..., :cout]), F.relu(F.conv2d(x, wt, b, 1, k // 2))) <= BF16_TOL
    assert float((out[..., cout:].float() - 7.0).abs().max()) == 0.0
    nv().conv2d_fwd(xo, wp, None, None, nhwc(act, torch.bfloat16), None, out[..., :cout], cout, False, False, 0, 0,
                    0.25, ops.CONV_TC)
    assert relerr(nchw(out[..., :cout]), 0.25 * F.conv2d(x, wt, None, 1, k // 2) * (act > 0)) <= BF16_TOL
    assert float((out[..., cout:].float() - 7.0).abs().max()) == 0.0
    nv().conv2d_fwd(xo, wp, None, nhwc(act, torch.bfloat16), None, None, out[..., :cout], cout, False, False, cout, 0,
                    0.5, ops.CONV_TC)
    assert relerr(nchw(out[..., :cout]), 0.5 * F.conv2d(x, wt, None, 1, k // 2) + act) <= BF16_TOL
    assert float((out[..., cout:].float() - 7.0).abs().max()) == 0.0
    nv().conv2d_fwd(xo, wp, b.cuda(), nhwc(act, torch.bfloat16), None, None, out[..., :cout], cout, True, False, cout,
                    0, 1.0, ops.CONV_TC)                             # bias + ReLU, then the residual
    assert relerr(nchw(out[..., :cout]), F.relu(F.conv2d(x, wt, b, 1, k // 2)) + act) <= BF16_TOL
    assert float((out[..., cout:].float() - 7.0).abs().max()) == 0.0
    out[..., :cout] = nhwc(act, torch.bfloat16)                      # plain accumulate: out += alpha * conv
    nv().conv2d_fwd(xo, wp, None, None, None, None, out[..., :cout], cout, False, True, 0, 0, 0.5, ops.CONV_TC)
    assert relerr(nchw(out[..., :cout]), 0.5 * F.conv2d(x, wt, None, 1, k // 2) + act) <= BF16_TOL
    assert float((out[..., cout:].float() - 7.0).abs().max()) == 0.0


def test_conv_rows_fused_column_sums():
    """colsum: per-channel sum of the mask-gated values the row kernel writes (the previous layer's bias gradient)."""
    from nerve_cl_b200 import ops
    n, h, w, cin, cout = 2, 37, 200, 96, 32
    g = torch.Generator().manual_seed(77)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = bf(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    act = bf(torch.randn(n, cout, h, w, generator=g))
    out = torch.empty((n, h, w, cout), device="cuda", dtype=torch.bfloat16)
    cs = torch.full((cout,), 3.0, device="cuda")
    nv().conv2d_fwd(nhwc(x, torch.bfloat16), pack(wt, torch.bfloat16), None, None, nhwc(act, torch.bfloat16), None, out,
                    cout, False, False, 0, 0, 0.5, ops.CONV_TC, None, False, cs)
    ref = 0.5 * F.conv2d(x, wt, None, 1, 1) * (act > 0)
    assert relerr(nchw(out), ref) <= BF16_TOL
    assert relerr(cs.cpu() - 3.0, ref.sum((0, 2, 3))) <= 2e-3
    with pytest.raises(RuntimeError):            # not a mask-gated <= 32-channel bf16 output: refused, never ignored
        nv().conv2d_fwd(nhwc(x, torch.bfloat16), pack(wt, torch.bfloat16), None, None, None, None, out, cout, False,
                        False, 0, 0, 1.0, ops.CONV_TC, None, False, cs)


def test_conv_tc_matches_simt_bitwise_inputs():
    """Same bf16 operands through both engines: results agree to bf16 rounding (different summation order)."""
    from nerve_cl_b200 import ops
    g = torch.Generator().manual_seed(99)
    x = torch.randn(1, 96, 40, 136, generator=g)
    wt = torch.randn(32, 96, 3, 3, generator=g) / 30
    xo, wp = nhwc(x, torch.bfloat16), pack(wt, torch.bfloat16)
    o1 = torch.empty((1, 40, 136, 32), device="cuda", dtype=torch.bfloat16)
    o2 = torch.empty_like(o1)
    nv().conv2d_fwd(xo, wp, None, None, None, None, o1, 32, False, False, 0, 0, 1.0, ops.CONV_SIMT)
    nv().conv2d_fwd(xo, wp, None, None, None, None, o2, 32, False, False, 0, 0, 1.0, ops.CONV_TC)
    assert relerr(o2, o1.float()) <= BF16_TOL


@pytest.mark.parametrize("shape", [(1, 24, 136, 64, 12, 3), (2, 17, 40, 32, 2, 3), (1, 20, 64, 64, 3, 3),
                                   (2, 12, 48, 8, 64, 3)])
def test_conv_fwd_tcgen05_fp32_out_small_channels(shape):
    """fp32 output (flow / logits / upsampler) and channel counts below one MMA chunk (Cout < 16, Cin = 8)."""
    from nerve_cl_b200 import ops
    n, h, w, cin, cout, k = shape
    g = torch.Generator().manual_seed(sum(shape) + 17)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = bf(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, wt, b, 1, k // 2)
    out = torch.full((n, h, w, cout), 7.0, device="cuda", dtype=torch.float32)
    nv().conv2d_fwd(nhwc(x, torch.bfloat16), pack(wt, torch.bfloat16), b.cuda(), None, None, None, out, cout,
                    False, False, 0, 0, 1.0, ops.CONV_TC)
    assert relerr(nchw(out), ref) <= 1e-3


WG_SHAPES = [  # (N, H, W, Cin, Cout, K)
    (1, 20, 72, 64, 12, 3), (2, 17, 40, 32, 2, 3), (1, 24, 64, 3, 64, 3),
    (1, 16, 128, 64, 32, 3), (2, 20, 72, 96, 32, 3), (1, 9, 200, 192, 32, 3), (1, 24, 40, 64, 64, 3),
    (1, 16, 64, 224, 64, 1), (2, 8, 16, 128, 64, 3), (1, 33, 65, 81, 128, 3), (1, 30, 257, 160, 32, 3),
    (3, 11, 23, 64, 64, 1), (1, 40, 136, 192, 64, 3), (2, 90, 160, 64, 32, 3), (1, 360, 640, 64, 64, 3),
    (1, 360, 640, 224, 64, 1),
]


@pytest.mark.parametrize("shape", WG_SHAPES)
def test_conv_wgrad_tcgen05(shape):
    from nerve_cl_b200 import ops
    n, h, w, cin, cout, k = shape
    g = torch.Generator().manual_seed(sum(shape) + 9)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    dy = bf(torch.randn(n, cout, h, w, generator=g))
    wt = torch.zeros(cout, cin, k, k, requires_grad=True)
    b = torch.zeros(cout, requires_grad=True)
    F.conv2d(x, wt, b, 1, k // 2).backward(dy)
    dw = torch.zeros(cout, cin, k, k, device="cuda")
    db = torch.zeros(cout, device="cuda")
    pad = (cin + 7) // 8 * 8 + 8
    nv().conv2d_wgrad(nhwc(x, torch.bfloat16, pad_to=pad), nhwc(dy, torch.bfloat16, pad_to=(cout + 7) // 8 * 8 + 8), dw, db, 0.5,
                      ops.CONV_TC)
    assert relerr(dw, 0.5 * wt.grad) <= 1e-3          # operands are exact bf16, accumulation is fp32
    assert relerr(db, 0.5 * b.grad) <= 1e-3


@pytest.mark.parametrize("descending", [False, True])
@pytest.mark.parametrize("shape", [(1, 20, 200), (2, 33, 128), (1, 9, 640), (1, 360, 640)])
def test_conv3x3_wgrad_grouped_dense_block(shape, descending):
    """One GEMM for the weight/bias gradients of the five dense-block layers (channel-prefix inputs of one
    buffer, adjacent output-gradient slices) == five separate ATen convolution_backward calls.  ``descending``: the
    layers' gradient slices in descending layer order (the engine's gradient-buffer layout)."""
    n, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 31)
    buf = bf(torch.randn(n, 224, h, w, generator=g))
    dy = bf(torch.randn(n, 160, h, w, generator=g))
    dws = [torch.zeros(32, 64 + 32 * i, 3, 3, device="cuda") for i in range(5)]
    dbs = [torch.zeros(32, device="cuda") for _ in range(5)]
    xb = nhwc(buf, torch.bfloat16, pad_to=256)
    col0 = [32 * (4 - i) if descending else 32 * i for i in range(5)]
    dy_buf = torch.cat([dy[:, 32 * i:32 * i + 32] for i in sorted(range(5), key=lambda i: col0[i])], 1)
    gb = nhwc(torch.cat([torch.zeros(n, 64, h, w), dy_buf], 1), torch.bfloat16, pad_to=256)
    nv().conv3x3_wgrad_grouped(xb[..., :192], gb[..., 64:224], dws, dbs, col0, 0.5)
    for i in range(5):
        cin = 64 + 32 * i
        wt = torch.zeros(32, cin, 3, 3, requires_grad=True)
        b = torch.zeros(32, requires_grad=True)
        F.conv2d(buf[:, :cin], wt, b, 1, 1).backward(dy[:, 32 * i:32 * i + 32])
        assert relerr(dws[i], 0.5 * wt.grad) <= 1e-3, i
        assert relerr(dbs[i], 0.5 * b.grad) <= 1e-3, i


@pytest.mark.parametrize("shape", [(1, 20, 130, 96, 32, 64), (2, 9, 200, 160, 64, 64), (1, 40, 128, 32, 32, 64),
                                   (1, 12, 136, 64, 48, 32), (1, 360, 640, 96, 32, 64), (1, 360, 640, 160, 64, 64)])
@pytest.mark.parametrize("center", [True, False])
def test_conv_rows_second_input(shape, center):
    """Virtual channel concat [x | x2] (row-streaming tcgen05 engine): 3x3 over x plus x2 through all taps or
    through the centre tap only, with the ReLU-mask epilogue (prefetched mask path when Cout <= 32)."""
    from nerve_cl_b200 import ops
    n, h, w, cin, cout, cin2 = shape
    g = torch.Generator().manual_seed(sum(shape) + 41)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    x2 = bf(torch.randn(n, cin2, h, w, generator=g))
    w1 = bf(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    w2 = bf(torch.randn(cout, cin2, 3, 3, generator=g) / (cin2 * 9) ** 0.5)
    if center:
        keep = torch.zeros(3, 3)
        keep[1, 1] = 1
        w2 = w2 * keep
    act = bf(torch.randn(n, cout, h, w, generator=g))
    ref = (F.conv2d(x, w1, None, 1, 1) + F.conv2d(x2, w2, None, 1, 1)) * 0.5 * (act > 0)
    cpad = (cin + 63) // 64 * 64
    comb = torch.zeros(cout, cpad + cin2, 3, 3)
    comb[:, :cin] = w1
    comb[:, cpad:] = w2
    out = torch.full((n, h, w, cout + 8), 7.0, device="cuda", dtype=torch.bfloat16)
    nv().conv2d_fwd(nhwc(x, torch.bfloat16, pad_to=cin + 8), pack(comb, torch.bfloat16), None, None,
                    nhwc(act, torch.bfloat16), None, out[..., :cout], cout, False, False, 0, 0, 0.5, ops.CONV_TC,
                    nhwc(x2, torch.bfloat16), center)
    assert relerr(nchw(out[..., :cout]), ref) <= BF16_TOL
    assert float((out[..., cout:].float() - 7.0).abs().max()) == 0.0
