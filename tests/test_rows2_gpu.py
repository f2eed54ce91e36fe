"""The CTA-pair (tcgen05 cta_group::2, M = 256) row-streaming convolution (csrc/conv_tc_rows2.cu) against ATen fp32 on
the same bf16 operands and against the 1-CTA row kernel: every lean epilogue, second-input variants, channel-group
splits, partial column strips, accumulator-ring laps with mirror slots, lock-step pieces that cross image / strip
borders, dead tail ranges."""
import pytest
import torch
import torch.nn.functional as F

from conftest import relerr
from test_ops_gpu import BF16_TOL, bf, nchw, nhwc, nv, pack

pytestmark = pytest.mark.gpu

SHAPES = [  # (N, H, W, Cin, Cout)  -- all have >= 1184 output rows of 128 pixels, so the pair kernel takes them
    (2, 160, 640, 64, 32), (1, 360, 640, 96, 32), (2, 130, 520, 192, 32), (4, 90, 300, 160, 64), (2, 150, 640, 128, 64),
    (1, 360, 640, 64, 128), (2, 200, 384, 96, 96), (2, 200, 640, 16, 64), (3, 140, 320, 32, 32), (2, 123, 650, 64, 80),
    (16, 37, 130, 64, 16),
]


def engines():
    from nerve_cl_b200 import ops
    return ops.CONV_TC, ops.CONV_TC_ROWS1


@pytest.mark.parametrize("shape", SHAPES)
def test_pair_kernel_lean_epilogues(shape):
    n, h, w, cin, cout = shape
    tc, rows1 = engines()
    g = torch.Generator().manual_seed(sum(shape))
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = bf(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    b = torch.randn(cout, generator=g)
    act = bf(torch.randn(n, cout, h, w, generator=g))
    conv = F.conv2d(x, wt, None, 1, 1)
    xo, wp, ao = nhwc(x, torch.bfloat16, pad_to=(cin + 63) // 64 * 64), pack(wt, torch.bfloat16), nhwc(act, torch.bfloat16)
    out = torch.full((n, h, w, cout + 16), 7.0, device="cuda", dtype=torch.bfloat16)
    o1 = torch.empty((n, h, w, cout), device="cuda", dtype=torch.bfloat16)

    def check(ref, *args, **kw):
        nv().conv2d_fwd(xo, wp, *args, tc, **kw)
        assert relerr(nchw(out[..., :cout]), ref) <= BF16_TOL
        assert float((out[..., cout:].float() - 7.0).abs().max()) == 0.0           # never writes outside its slice
        a1 = list(args)
        a1[4] = o1                                                                # same call on the 1-CTA kernel
        nv().conv2d_fwd(xo, wp, *a1, rows1, **kw)
        assert relerr(out[..., :cout].float(), o1.float()) <= 8e-3        # (one bf16 ulp at the largest value: summation order)

    ov = out[..., :cout]
    check(F.relu(conv + b.view(1, -1, 1, 1)), b.cuda(), None, None, None, ov, cout, 1, False, 0, 0, 1.0)        # bias + ReLU
    check(F.relu(conv + b.view(1, -1, 1, 1)) + act, b.cuda(), ao, None, None, ov, cout, 1, False, cout, 0, 1.0)  # ... + residual
    check(conv, None, None, None, None, ov, cout, 0, False, 0, 0, 1.0)                                          # plain
    check(0.25 * conv * (act > 0), None, None, ao, None, ov, cout, 0, False, 0, 0, 0.25)                        # mask-gated
    check(0.5 * conv + act, None, ao, None, None, ov, cout, 0, False, cout, 0, 0.5)                             # alpha * acc + res
    ov.copy_(ao)
    o1.copy_(ao)
    nv().conv2d_fwd(xo, wp, None, None, None, None, ov, cout, 0, True, 0, 0, 0.5, tc)                           # out += alpha * acc
    assert relerr(nchw(ov), 0.5 * conv + act) <= BF16_TOL
    assert float((out[..., cout:].float() - 7.0).abs().max()) == 0.0
    if cout <= 32:                                                                                              # fused column sums
        cs = torch.full((cout,), 3.0, device="cuda")
        nv().conv2d_fwd(xo, wp, None, None, ao, None, ov, cout, 0, False, 0, 0, 0.5, tc, None, False, cs)
        ref = 0.5 * conv * (act > 0)
        assert relerr(nchw(ov), ref) <= BF16_TOL
        assert relerr(cs.cpu() - 3.0, ref.sum((0, 2, 3))) <= 3e-3


@pytest.mark.parametrize("shape", [(2, 160, 640, 96, 32, 64), (1, 360, 640, 160, 64, 64), (2, 200, 384, 32, 32, 64),
                                   (4, 100, 300, 64, 48, 32)])
@pytest.mark.parametrize("center", [True, False])
def test_pair_kernel_second_input(shape, center):
    """Virtual concat [x | x2] with x2 through all taps or the centre tap only (the fused dense-block slice gradient)."""
    n, h, w, cin, cout, cin2 = shape
    tc, rows1 = engines()
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
                    nhwc(act, torch.bfloat16), None, out[..., :cout], cout, 0, False, 0, 0, 0.5, tc,
                    nhwc(x2, torch.bfloat16), center)
    assert relerr(nchw(out[..., :cout]), ref) <= BF16_TOL
    assert float((out[..., cout:].float() - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(2, 160, 640, 64, 32), (1, 360, 640, 160, 32), (3, 140, 320, 32, 16), (2, 123, 650, 96, 32)])
def test_pair_kernel_sign_bits(shape):
    """ABI v7 packed ReLU signs: a forward conv (bias + ReLU) writes one bit per channel of word (group, pixel) <-> out[pixel, channel] > 0;
    a data-gradient conv that reads them as its mask == the same conv gated by the bf16 activation (bit-identical), with
    and without the centre-tap second input and the fused column sums."""
    n, h, w, cin, cout = shape
    tc, _ = engines()
    g = torch.Generator().manual_seed(sum(shape) + 5)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = bf(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    b = torch.randn(cout, generator=g)
    xo, wp = nhwc(x, torch.bfloat16, pad_to=(cin + 63) // 64 * 64), pack(wt, torch.bfloat16)
    act = torch.empty((n, h, w, cout), device="cuda", dtype=torch.bfloat16)
    bits = torch.zeros((cout // 16, n, h, w), device="cuda", dtype=torch.int16)
    nv().conv2d_fwd(xo, wp, b.cuda(), None, None, None, act, cout, True, False, 0, 0, 1.0, tc, None, False, None, bits, 1)
    ref = F.relu(F.conv2d(x, wt, b, 1, 1))
    assert relerr(nchw(act), ref) <= BF16_TOL
    want = (act.view(n, h, w, cout // 16, 16) > 0).to(torch.int32)
    pos = torch.tensor([15 - c // 2 if c & 1 else 7 - c // 2 for c in range(16)], device="cuda", dtype=torch.int32)
    want = (want << pos).sum(-1).permute(3, 0, 1, 2)
    assert torch.equal(bits.to(torch.int32) & 0xFFFF, want)
    # mode 2: gradient conv gated by the bits vs gated by the activation itself
    dy = bf(torch.randn(n, cin, h, w, generator=g))
    x2 = bf(torch.randn(n, 64, h, w, generator=g))
    cpad = (cin + 63) // 64 * 64
    comb = torch.zeros(cout, cpad + 64, 3, 3)
    comb[:, :cin] = bf(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    comb[:, cpad:, 1, 1] = bf(torch.randn(cout, 64, generator=g) / 8)
    wc = pack(comb, torch.bfloat16)
    dyo, x2o = nhwc(dy, torch.bfloat16, pad_to=cpad), nhwc(x2, torch.bfloat16)
    for second in (False, True):
        a_out = torch.empty((n, h, w, cout), device="cuda", dtype=torch.bfloat16)
        b_out = torch.empty_like(a_out)
        cs_a, cs_b = torch.zeros(cout, device="cuda"), torch.zeros(cout, device="cuda")
        nv().conv2d_fwd(dyo, wc, None, None, act, None, a_out, cout, False, False, 0, 0, 0.5, tc,
                        x2o if second else None, second, cs_a)
        nv().conv2d_fwd(dyo, wc, None, None, None, None, b_out, cout, False, False, 0, 0, 0.5, tc,
                        x2o if second else None, second, cs_b, bits, 2)
        assert torch.equal(a_out, b_out)
        assert relerr(cs_b, cs_a) <= 1e-5
