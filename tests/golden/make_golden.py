"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

For each case it runs the unmodified reference (``nerve_cl`` imported from /root/reference) on seeded
inputs, checks that ``oracle/`` reproduces it, and stores inputs + reference outputs as ``.npz``.
The fixtures are what pins the oracle (tests/test_oracle_golden.py) and, on the GPU, the CUDA path.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from nerve_cl.models import SuperResolutionNet  # noqa: E402  (the reference)
from nerve_cl.models.super_resolution import warp_features  # noqa: E402
from nerve_cl.models.layers import LiteFlowNetCorrelation  # noqa: E402
from nerve_cl.continual import EWC  # noqa: E402
from nerve_cl.continual.ewc import SynapticIntelligence  # noqa: E402

from oracle import sr_oracle, ewc_oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def relerr(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(f"wrote {name}: {os.path.getsize(path)/1024:.1f} KiB")


def synth_frames(b, t, h, w, seed):
    """Centre frame uniform; neighbours = centre rolled by an integer shift + small noise."""
    g = torch.Generator().manual_seed(seed)
    centre = torch.rand(b, 3, h, w, generator=g)
    frames = []
    for i in range(t):
        if i == t // 2:
            frames.append(centre)
            continue
        dx = int(torch.randint(-3, 4, (1,), generator=g))
        dy = int(torch.randint(-3, 4, (1,), generator=g))
        f = torch.roll(centre, (dy, dx), (2, 3)) + 0.01 * torch.randn(b, 3, h, w, generator=g)
        frames.append(f.clamp(0, 1))
    return torch.stack(frames, 1)


def sr_case(name, scale, feats, blocks, tw, b, h, w, seed, training, store_weights):
    torch.manual_seed(seed)
    ref = SuperResolutionNet(scale_factor=scale, num_features=feats,
                             num_residual_blocks=blocks, temporal_window=tw)
    # non-trivial BN buffers so eval-mode is a real test
    g = torch.Generator().manual_seed(seed + 1)
    for i in range(3):
        bn = ref.feature_extractor.body[i].bn
        bn.running_mean.copy_(0.05 * torch.randn(feats, generator=g))
        bn.running_var.copy_(1.0 + 0.2 * torch.rand(feats, generator=g))
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    t = 2 * tw + 1
    x = synth_frames(b, t, h, w, seed + 2)
    target = torch.rand(b, 3, h * scale, w * scale, generator=torch.Generator().manual_seed(seed + 3))

    ref.train(training)
    out, inter = ref(x, return_intermediate=True)
    loss = torch.nn.functional.mse_loss(out, target)
    loss.backward()
    ref_grads = {n: p.grad.clone() for n, p in ref.named_parameters()}
    sd1 = ref.state_dict()

    # oracle must agree with the live reference
    osd = {k: v.clone() for k, v in sd0.items()}
    o_out, o_loss, o_grads = sr_oracle.train_step_grads(osd, x, target, scale, training)
    e_out = relerr(o_out, out)
    e_grad = max(relerr(o_grads[n], ref_grads[n]) for n in ref_grads)
    e_bn = max(relerr(osd[k].float(), sd1[k].float()) for k in sd1 if "running" in k or "tracked" in k)
    print(f"{name}: oracle-vs-reference out {e_out:.2e} grad {e_grad:.2e} bn {e_bn:.2e}")
    assert e_out < 1e-6 and e_grad < 1e-5 and e_bn < 1e-6

    arrays = dict(
        meta=np.array([scale, feats, blocks, tw, b, h, w, seed, int(training)], dtype=np.int64),
        lr_frames=x, target=target, out=out, loss=loss,
        aggregated=inter["aggregated"],
        feat_centre=inter["features"][t // 2],
        aligned0=inter["aligned"][0],
    )
    for k, v in sd1.items():
        if "running" in k or "tracked" in k:
            arrays["bn1/" + k] = v
    if store_weights:
        for k, v in sd0.items():
            arrays["w/" + k] = v
        for k, v in ref_grads.items():
            arrays["g/" + k] = v
    else:
        # full-size model: weights are re-created from the seed by the module under test; keep
        # checksums of weights and a norm + strided sample of every gradient instead of 8 MB blobs.
        arrays["w_sum"] = np.array([float(v.double().sum()) for k, v in sd0.items()])
        arrays["w_abs"] = np.array([float(v.double().abs().sum()) for k, v in sd0.items()])
        arrays["g_norm"] = np.array([float(v.double().norm()) for v in ref_grads.values()])
        arrays["g_head"] = np.concatenate([v.flatten()[:16] for v in ref_grads.values()])
    save(name, **arrays)


def structured_flow(b, h, w):
    """Exactly representable, regenerable flow for the full-size index test (no blob stored)."""
    ys = torch.arange(h).view(1, h, 1).expand(b, h, w)
    xs = torch.arange(w).view(1, 1, w).expand(b, h, w)
    fx = ((xs * 7 + ys * 3) % 33 - 16).float() * 0.25 + 2.0 ** -12
    fy = ((xs * 5 + ys * 11) % 29 - 14).float() * 0.125 - 2.0 ** -11
    return torch.stack([fx, fy], 1)


def warp_case():
    """Dedicated warp vectors: zero flow, exact integers, +-0.5, +-(1-2^-20), out of range, edges."""
    arrays = {}
    for (h, w) in [(9, 13), (36, 64)]:
        g = torch.Generator().manual_seed(h * 1000 + w)
        b, c = 2, 4
        base = [torch.zeros(b, 2, h, w)]
        base.append(torch.randint(-3, 4, (b, 2, h, w), generator=g).float())
        base.append(torch.full((b, 2, h, w), 0.5))
        base.append(torch.full((b, 2, h, w), -0.5))
        base.append(torch.full((b, 2, h, w), 1 - 2.0 ** -20))
        base.append(torch.full((b, 2, h, w), -(1 - 2.0 ** -20)))
        base.append(torch.full((b, 2, h, w), 8.0))
        base.append(torch.full((b, 2, h, w), -8.0))
        base.append(4 * torch.randn(b, 2, h, w, generator=g))
        edge = torch.zeros(b, 2, h, w)   # land exactly on W-1 / H-1
        edge[:, 0] = (w - 1) - torch.arange(w).float()[None, None, :]
        edge[:, 1] = (h - 1) - torch.arange(h).float()[None, :, None]
        base.append(edge)
        feat = torch.randn(b, c, h, w, generator=g)
        arrays[f"{h}x{w}/feat"] = feat
        for i, flow in enumerate(base):
            ref = warp_features(feat, flow)
            orc = sr_oracle.warp(feat, flow)
            assert torch.equal(ref, orc), (h, w, i)
            key = f"{h}x{w}/{i}"
            arrays[key + "/flow"] = flow
            arrays[key + "/idx"] = sr_oracle.warp_corner_indices(flow).to(torch.int16)
            arrays[key + "/out"] = ref
    # full-size (360x640) index vectors: zero flow and the structured flow, CPU-ATen semantics
    h, w = 360, 640
    for name, flow in (("zero", torch.zeros(1, 2, h, w)), ("structured", structured_flow(1, h, w))):
        arrays[f"{h}x{w}/{name}/idx"] = sr_oracle.warp_corner_indices(flow).to(torch.int16)
    save("warp_cases.npz", **arrays)


def corr_case():
    g = torch.Generator().manual_seed(7)
    x1 = torch.randn(2, 8, 11, 14, generator=g)
    x2 = torch.randn(2, 8, 11, 14, generator=g)
    ref = LiteFlowNetCorrelation(4)(x1, x2)
    orc = sr_oracle.correlation(x1, x2)
    assert relerr(orc, ref) < 1e-6
    save("corr_case.npz", x1=x1, x2=x2, out=ref)


def ewc_case():
    """EWC on nn.Linear + TensorDataset, like the reference's tests/test_continual.py:60-89."""
    torch.manual_seed(11)
    model = torch.nn.Linear(10, 10)
    xs, ys = torch.randn(40, 10), torch.randn(40, 10)
    loader = [(xs[i:i + 8], ys[i:i + 8]) for i in range(0, 40, 8)]
    ewc = EWC(model, ewc_lambda=5000.0, mode="online", decay=0.9)
    w0 = torch.cat([p.detach().flatten() for p in model.parameters()]).clone()
    ewc.register_task(0, loader)
    f0 = torch.cat([ewc.fisher_dict[n].flatten() for n, _ in model.named_parameters()]).clone()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.1 * torch.randn_like(p))
    w1 = torch.cat([p.detach().flatten() for p in model.parameters()]).clone()
    model.zero_grad()                 # compute_fisher leaves .grad dirty (ewc.py:139)
    pen1 = ewc.penalty()
    pen1.backward()
    gpen = torch.cat([p.grad.flatten() for p in model.parameters()]).clone()
    ewc.register_task(1, loader)
    f1 = torch.cat([ewc.fisher_dict[n].flatten() for n, _ in model.named_parameters()]).clone()

    # oracle agreement
    def batch_grads(wflat):
        m = torch.nn.Linear(10, 10)
        with torch.no_grad():
            m.weight.copy_(wflat[:100].view(10, 10)); m.bias.copy_(wflat[100:])
        out = []
        for xb, yb in loader:
            m.zero_grad()
            torch.nn.functional.mse_loss(m(xb), yb).backward()
            out.append(torch.cat([p.grad.flatten() for p in m.parameters()]).numpy().copy())
        return out
    of0 = ewc_oracle.fisher_from_batches(batch_grads(w0), [8] * 5)
    assert np.allclose(of0, f0.numpy(), rtol=1e-6, atol=0)
    op = ewc_oracle.penalty(w1.numpy(), of0, w0.numpy(), 5000.0)
    assert abs(op - float(pen1.detach())) / abs(float(pen1.detach())) < 1e-5
    og = ewc_oracle.penalty_grad(w1.numpy(), of0, w0.numpy(), 5000.0)
    assert np.allclose(og, gpen.numpy(), rtol=1e-5, atol=1e-7)
    of1 = ewc_oracle.consolidate(of0, ewc_oracle.fisher_from_batches(batch_grads(w1), [8] * 5), 0.9)
    assert np.allclose(of1, f1.numpy(), rtol=1e-5, atol=1e-9)
    save("ewc_linear.npz", xs=xs, ys=ys, w0=w0, w1=w1, fisher0=f0, fisher1=f1,
         penalty1=pen1.detach(), penalty_grad1=gpen)


def _flat_params(model):
    return torch.cat([p.detach().flatten() for p in model.parameters()]).clone()


def _flat_dict(d, model):
    return torch.cat([d[n].flatten() for n, _ in model.named_parameters()]).clone()


def ewc_separate_case():
    """'separate' mode (ewc.py:172-178, 213-223): two tasks on different data, then the summed penalty and its
    gradient at a third parameter point -- from the live reference."""
    torch.manual_seed(12)
    model = torch.nn.Linear(10, 10)
    xa, ya = torch.randn(40, 10), torch.randn(40, 10)
    xb, yb = torch.randn(24, 10) * 2.0, torch.randn(24, 10)
    la = [(xa[i:i + 8], ya[i:i + 8]) for i in range(0, 40, 8)]
    lb = [(xb[i:i + 8], yb[i:i + 8]) for i in range(0, 24, 8)]
    ewc = EWC(model, ewc_lambda=300.0, mode="separate")
    w0 = _flat_params(model)
    ewc.register_task(0, la)
    f0 = _flat_dict(ewc.task_fisher[0], model)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.1 * torch.randn_like(p))
    w1 = _flat_params(model)
    ewc.register_task(1, lb)
    f1 = _flat_dict(ewc.task_fisher[1], model)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    w2 = _flat_params(model)
    model.zero_grad()
    pen = ewc.penalty()
    pen.backward()
    gpen = torch.cat([p.grad.flatten() for p in model.parameters()]).clone()
    op = ewc_oracle.penalty_separate(w2.numpy(), [f0.numpy(), f1.numpy()], [w0.numpy(), w1.numpy()], 300.0)
    assert abs(op - float(pen.detach())) <= 1e-5 * abs(float(pen.detach()))
    og = ewc_oracle.penalty_separate_grad(w2.numpy(), [f0.numpy(), f1.numpy()], [w0.numpy(), w1.numpy()], 300.0)
    assert np.allclose(og, gpen.numpy(), rtol=1e-5, atol=1e-7)
    save("ewc_separate.npz", xa=xa, ya=ya, xb=xb, yb=yb, w0=w0, w1=w1, w2=w2, fisher0=f0, fisher1=f1,
         penalty2=pen.detach(), penalty_grad2=gpen)


def si_case():
    """SynapticIntelligence (ewc.py:306-379) driven like a training loop: 6 SGD steps with update_importance after
    each, register_task, 3 more steps (the last one WITHOUT update_importance, so register_task sees a non-zero
    delta), register_task, then penalty + gradient after a final parameter move.  One parameter (the bias) has
    grad=None during step 2 to pin the 'skipped tensor keeps W and p_old' behaviour."""
    torch.manual_seed(13)
    model = torch.nn.Linear(10, 10)
    xs, ys = torch.randn(72, 10), torch.randn(72, 10)
    si = SynapticIntelligence(model, si_lambda=0.7, damping=0.1)
    w_init = _flat_params(model)
    lr = 0.05
    W = np.zeros(110, np.float32); po = w_init.numpy().copy(); om = np.zeros(110, np.float32)
    grads, skip_bias = [], []

    def sgd_step(i, update=True, drop_bias=False):
        nonlocal W, po
        model.zero_grad()
        torch.nn.functional.mse_loss(model(xs[8 * i:8 * i + 8]), ys[8 * i:8 * i + 8]).backward()
        g = torch.cat([p.grad.flatten() for p in model.parameters()]).clone()
        with torch.no_grad():
            for p in model.parameters():
                p.add_(-lr * p.grad)
        if drop_bias:
            model.bias.grad = None
        if update:
            si.update_importance()
            th = _flat_params(model).numpy()
            if drop_bias:
                Wn, pn = ewc_oracle.si_update(W[:100], po[:100], th[:100], g.numpy()[:100])
                W = np.concatenate([Wn, W[100:]]); po = np.concatenate([pn, po[100:]])
            else:
                W, po = ewc_oracle.si_update(W, po, th, g.numpy())
        grads.append(g); skip_bias.append(drop_bias)

    for i in range(6):
        sgd_step(i, drop_bias=(i == 2))
    si.register_task()
    W, po, om = ewc_oracle.si_register(W, po, om, _flat_params(model).numpy(), 0.1)
    omega_a = _flat_dict(si.omega, model)
    assert np.allclose(om, omega_a.numpy(), rtol=1e-5, atol=1e-9)
    for i in range(6, 9):
        sgd_step(i, update=(i != 8))
    si.register_task()
    W, po, om = ewc_oracle.si_register(W, po, om, _flat_params(model).numpy(), 0.1)
    omega_b = _flat_dict(si.omega, model)
    pold_b = _flat_dict(si.p_old, model)
    assert np.allclose(om, omega_b.numpy(), rtol=1e-5, atol=1e-9) and np.array_equal(po, pold_b.numpy())
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    w_final = _flat_params(model)
    model.zero_grad()
    pen = si.penalty()
    pen.backward()
    gpen = torch.cat([p.grad.flatten() for p in model.parameters()]).clone()
    op = ewc_oracle.si_penalty(w_final.numpy(), om, po, 0.7)
    assert abs(op - float(pen.detach())) <= 1e-5 * abs(float(pen.detach()))
    assert np.allclose(ewc_oracle.si_penalty_grad(w_final.numpy(), om, po, 0.7), gpen.numpy(), rtol=1e-5, atol=1e-8)
    save("si_linear.npz", xs=xs, ys=ys, w_init=w_init, lr=np.float32(lr), omega_a=omega_a, omega_b=omega_b,
         p_old_b=pold_b, w_final=w_final, penalty=pen.detach(), penalty_grad=gpen)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    only = ap.parse_args().only
    if only:
        for nm in only.split(","):
            globals()[nm]()
        sys.exit(0)
    ewc_separate_case()
    si_case()
    warp_case()
    corr_case()
    ewc_case()
    # tiny model, weights stored: train-mode and eval-mode, x2 / x3 / x4
    sr_case("sr_tiny_x2_train.npz", 2, 16, 1, 1, 2, 10, 12, 100, True, True)
    sr_case("sr_tiny_x2_eval.npz", 2, 16, 1, 1, 2, 10, 12, 100, False, False)
    sr_case("sr_tiny_x3_train.npz", 3, 16, 1, 1, 1, 9, 11, 200, True, False)
    sr_case("sr_tiny_x4_t5_train.npz", 4, 16, 1, 2, 1, 12, 10, 300, True, False)
    # the real default model (64 feat / 8 blocks), BASELINE config-1-like but smaller spatially
    sr_case("sr_default_x2_train.npz", 2, 64, 8, 1, 2, 24, 32, 0, True, False)
