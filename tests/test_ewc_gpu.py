"""EWC on the CUDA kernels vs the golden fixture (live reference on nn.Linear, as the reference's own
tests/test_continual.py:60-89 does) and vs the oracle on the real SR model."""
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu


def make_linear(w):
    m = torch.nn.Linear(10, 10)
    with torch.no_grad():
        m.weight.copy_(torch.from_numpy(w[:100]).view(10, 10))
        m.bias.copy_(torch.from_numpy(w[100:]))
    return m.cuda()


def flat(d, model):
    return torch.cat([d[n].flatten() for n, _ in model.named_parameters()]).cpu().numpy()


@pytest.mark.parametrize("mode", ["online", "separate"])
def test_linear_matches_reference(mode):
    from nerve_cl_b200.continual import EWC
    g = load_golden("ewc_linear.npz")
    xs, ys = torch.from_numpy(g["xs"]), torch.from_numpy(g["ys"])
    loader = [(xs[i:i + 8], ys[i:i + 8]) for i in range(0, 40, 8)]
    model = make_linear(g["w0"])
    ewc = EWC(model, ewc_lambda=5000.0, mode=mode, decay=0.9)
    assert ewc.penalty() == 0.0
    ewc.register_task(0, loader)
    assert ewc.num_tasks == 1 and not model.training                 # compute_fisher leaves eval() (ewc.py:99)
    f0 = flat(ewc.fisher_dict if mode == "online" else ewc.task_fisher[0], model)
    np.testing.assert_allclose(f0, g["fisher0"], rtol=1e-5, atol=1e-12)
    before = float(ewc.penalty())
    assert before == 0.0
    with torch.no_grad():
        w1 = torch.from_numpy(g["w1"]).cuda()
        model.weight.copy_(w1[:100].view(10, 10))
        model.bias.copy_(w1[100:])
    model.zero_grad()
    pen = ewc.penalty()
    assert float(pen) > before                                       # the reference's own assertion (:71-89)
    assert abs(float(pen) - float(g["penalty1"])) <= 1e-5 * abs(float(g["penalty1"]))
    pen.backward()
    grad = torch.cat([p.grad.flatten() for p in model.parameters()]).cpu().numpy()
    np.testing.assert_allclose(grad, g["penalty_grad1"], rtol=1e-5, atol=1e-7)
    (2.0 * ewc.penalty()).backward()                                 # upstream gradient scaling + accumulation
    grad3 = torch.cat([p.grad.flatten() for p in model.parameters()]).cpu().numpy()
    np.testing.assert_allclose(grad3, 3 * g["penalty_grad1"], rtol=1e-5, atol=1e-6)
    ewc.register_task(1, loader)
    if mode == "online":
        np.testing.assert_allclose(flat(ewc.fisher_dict, model), g["fisher1"], rtol=1e-5, atol=1e-9)
    else:
        assert set(ewc.task_fisher) == {0, 1}
        assert float(ewc.penalty()) >= 0.0
    # checkpoint round trip keeps the penalty
    state = ewc.state_dict()
    assert all(v.device.type == "cpu" for v in state["fisher_dict"].values())
    e2 = EWC(model, mode=mode)
    e2.load_state_dict(state)
    with torch.no_grad():
        model.weight.add_(0.05)
    assert abs(float(e2.penalty()) - float(ewc.penalty())) <= 1e-6 * abs(float(ewc.penalty())) + 1e-12
    stats = ewc.get_importance_stats()
    assert set(stats["weight"]) == {"mean", "max", "std", "nonzero"}


def test_sr_model_fisher_and_penalty_vs_oracle():
    """compute_fisher on the SR module (gradients from the CUDA engine) vs the oracle's Fisher definition
    applied to oracle gradients; penalty/gradient vs oracle formulas on the flat vectors."""
    from oracle import ewc_oracle, sr_oracle
    from nerve_cl_b200.continual import EWC
    from nerve_cl_b200.models import SuperResolutionNet
    torch.manual_seed(21)
    model = SuperResolutionNet(num_features=16, num_residual_blocks=1).cuda()
    model.compute_dtype = torch.float32
    model.warp_div_mode = 1
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(22)
    batches = [(torch.rand(2, 3, 3, 12, 16, generator=g), torch.rand(2, 3, 24, 32, generator=g)) for _ in range(3)]
    ewc = EWC(model, ewc_lambda=100.0)
    ewc.register_task(0, batches)
    names = [n for n, _ in model.named_parameters()]
    grads = []
    for x, t in batches:
        _, _, gr = sr_oracle.train_step_grads({k: v.clone() for k, v in sd.items()}, x, t, 2, training=False)
        grads.append(np.concatenate([gr[n].numpy().ravel() for n in names]))
    want = ewc_oracle.fisher_from_batches(grads, [2, 2, 2])
    got = np.concatenate([ewc.fisher_dict[n].cpu().numpy().ravel() for n in names])
    assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.01 * torch.randn_like(p))
    theta = np.concatenate([p.detach().cpu().numpy().ravel() for p in model.parameters()])
    star = np.concatenate([ewc.optpar_dict[n].cpu().numpy().ravel() for n in names])
    model.zero_grad()
    pen = ewc.penalty()
    wantp = ewc_oracle.penalty(theta, got, star, 100.0)
    assert abs(float(pen) - wantp) <= 1e-5 * abs(wantp)
    pen.backward()
    gpen = np.concatenate([p.grad.cpu().numpy().ravel() for p in model.parameters()])
    np.testing.assert_allclose(gpen, ewc_oracle.penalty_grad(theta, got, star, 100.0), rtol=1e-5, atol=1e-9)


def test_large_flat_buffer_properties():
    """Size-independent checks at a bandwidth-relevant size (2^26 params): penalty is quadratic in the
    displacement and linear in F; Fisher accumulation of k identical gradients equals k*g^2."""
    from nerve_cl_b200 import ops
    n = 1 << 26
    g = torch.Generator(device="cuda").manual_seed(1)
    theta = torch.randn(n, device="cuda", generator=g)
    star = torch.randn(n, device="cuda", generator=g)
    fisher = torch.rand(n, device="cuda", generator=g)
    out = torch.zeros(3, device="cuda")
    ops.nv.ewc_penalty_fwd([theta], fisher, star, 1.0, out[0:1])
    ops.nv.ewc_penalty_fwd([star + 2 * (theta - star)], fisher, star, 1.0, out[1:2])
    ops.nv.ewc_penalty_fwd([theta], 3 * fisher, star, 1.0, out[2:3])
    ref = float((fisher.double() * (theta.double() - star.double()) ** 2).sum())
    assert abs(float(out[0]) - ref) <= 1e-5 * ref
    assert abs(float(out[1]) / float(out[0]) - 4.0) <= 1e-4
    assert abs(float(out[2]) / float(out[0]) - 3.0) <= 1e-4
    acc = torch.zeros(n, device="cuda")
    for _ in range(3):
        ops.nv.ewc_fisher_accum(acc, [theta], [n], 1.0)
    assert relerr(acc[:1 << 20], 3 * theta[:1 << 20] ** 2) <= 1e-6
    grad = torch.zeros(n, device="cuda")
    ops.nv.ewc_penalty_bwd([theta], [grad], fisher, star, 2.0, None)
    assert relerr(grad[:1 << 20], (2.0 * fisher * (theta - star))[:1 << 20]) <= 1e-6


def test_adamw_matches_torch():
    from nerve_cl_b200 import ops
    g = torch.Generator().manual_seed(3)
    p = torch.randn(1001, generator=g)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-2, weight_decay=1e-2)
    pc, m, v = p.cuda(), torch.zeros(1001, device="cuda"), torch.zeros(1001, device="cuda")
    for step in range(1, 4):
        grad = torch.randn(1001, generator=g)
        ref.grad = grad.clone()
        opt.step()
        ops.nv.adamw_step(pc, grad.cuda(), m, v, 1e-2, 0.9, 0.999, 1e-8, 1e-2, step, 1.0)
    assert relerr(pc, ref.detach()) <= 1e-5
