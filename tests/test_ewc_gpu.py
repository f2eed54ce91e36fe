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


def test_separate_mode_matches_reference():
    """Two tasks in 'separate' mode: per-task Fishers, the summed penalty (reference ewc.py:213-223) and its gradient
    against the golden generated from the live reference (tests/golden/make_golden.py::ewc_separate_case)."""
    from nerve_cl_b200.continual import EWC
    g = load_golden("ewc_separate.npz")
    xa, ya, xb, yb = (torch.from_numpy(g[k]) for k in ("xa", "ya", "xb", "yb"))
    la = [(xa[i:i + 8], ya[i:i + 8]) for i in range(0, 40, 8)]
    lb = [(xb[i:i + 8], yb[i:i + 8]) for i in range(0, 24, 8)]
    model = make_linear(g["w0"])
    ewc = EWC(model, ewc_lambda=300.0, mode="separate")
    ewc.register_task(0, la)
    set_linear(model, g["w1"])
    ewc.register_task(1, lb)
    np.testing.assert_allclose(flat(ewc.task_fisher[0], model), g["fisher0"], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(flat(ewc.task_fisher[1], model), g["fisher1"], rtol=1e-5, atol=1e-12)
    np.testing.assert_array_equal(flat(ewc.task_optpar[0], model), g["w0"])
    np.testing.assert_array_equal(flat(ewc.task_optpar[1], model), g["w1"])
    set_linear(model, g["w2"])
    model.zero_grad()
    pen = ewc.penalty()
    assert abs(float(pen) - float(g["penalty2"])) <= 1e-5 * abs(float(g["penalty2"]))
    pen.backward()
    grad = torch.cat([p.grad.flatten() for p in model.parameters()]).cpu().numpy()
    np.testing.assert_allclose(grad, g["penalty_grad2"], rtol=1e-5, atol=1e-7)
    # state_dict round trip keeps both tasks
    e2 = EWC(model, mode="separate")
    e2.load_state_dict(ewc.state_dict())
    assert abs(float(e2.penalty()) - float(g["penalty2"])) <= 1e-5 * abs(float(g["penalty2"]))


def set_linear(model, w):
    with torch.no_grad():
        w = torch.from_numpy(w).cuda()
        model.weight.copy_(w[:100].view(10, 10))
        model.bias.copy_(w[100:])


def test_synaptic_intelligence_matches_reference():
    """SynapticIntelligence on the fused kernels replays the golden's training loop (live reference,
    tests/golden/make_golden.py::si_case): omega after each register_task, p_old, penalty and its gradient."""
    from nerve_cl_b200.continual import SynapticIntelligence
    from test_oracle_golden import si_replay
    g = load_golden("si_linear.npz")
    xs, ys, lr = torch.from_numpy(g["xs"]).cuda(), torch.from_numpy(g["ys"]).cuda(), float(g["lr"])
    model = make_linear(g["w_init"])
    si = SynapticIntelligence(model, si_lambda=0.7, damping=0.1)
    assert set(si.W) == {"weight", "bias"} and float(si.penalty()) == 0.0

    def step_grad(i):
        model.zero_grad()
        torch.nn.functional.mse_loss(model(xs[8 * i:8 * i + 8]), ys[8 * i:8 * i + 8]).backward()
        return None

    def apply_step(_):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(-lr * p.grad)

    def update(_, drop_bias):
        if drop_bias:
            model.bias.grad = None
        si.update_importance()

    for tag in si_replay(g, update, si.register_task, None, step_grad, apply_step):
        # (the gradients here come from ATen-CUDA matmuls, the golden's from ATen-CPU: a few ulp apart)
        np.testing.assert_allclose(flat(si.omega, model), g["omega_" + tag], rtol=5e-4, atol=1e-7)
    np.testing.assert_allclose(flat(si.p_old, model), g["p_old_b"], rtol=1e-5, atol=1e-6)
    assert float(torch.cat([v.flatten() for v in si.W.values()]).abs().max()) == 0.0
    # penalty / gradient with the golden's own state (so the check is not loosened by the trajectory noise)
    with torch.no_grad():
        si._omega.flat.copy_(torch.from_numpy(g["omega_b"]))
        si._p_old.flat.copy_(torch.from_numpy(g["p_old_b"]))
    set_linear(model, g["w_final"])
    model.zero_grad()
    pen = si.penalty()
    assert abs(float(pen) - float(g["penalty"])) <= 1e-5 * abs(float(g["penalty"]))
    pen.backward()
    grad = torch.cat([p.grad.flatten() for p in model.parameters()]).cpu().numpy()
    np.testing.assert_allclose(grad, g["penalty_grad"], rtol=1e-5, atol=1e-8)


def test_flat_kernels_skip_none_grads_beyond_32_tensors():
    """More than 32 parameter tensors with some ``grad is None`` / zero-size entries (reference ewc.py:140 skips them):
    every later tensor must still land at its own offset of the flat buffers (the host chunking once re-visited
    tensor 32 and shifted everything after it)."""
    from nerve_cl_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(5)
    sizes = [10, 0, 7, 33] + [5 + (i % 9) for i in range(70)]
    none_at = {0, 2, 31, 32, 40, 73}
    grads = [None if i in none_at else torch.randn(n, device="cuda", generator=gen) for i, n in enumerate(sizes)]
    theta = [torch.randn(n, device="cuda", generator=gen) for n in sizes]
    total = sum(sizes)
    fisher = torch.rand(total, device="cuda", generator=gen)
    want = fisher.clone()
    off = 0
    for gsl, n in zip(grads, sizes):
        if gsl is not None:
            want[off:off + n] += 0.5 * gsl * gsl
        off += n
    ops.nv.ewc_fisher_accum(fisher, grads, sizes, 0.5)
    assert relerr(fisher, want) <= 1e-6
    # SI update with the same None pattern
    W, po = torch.zeros(total, device="cuda"), torch.randn(total, device="cuda", generator=gen)
    wantW, wantpo = W.clone(), po.clone()
    off = 0
    for t, gsl, n in zip(theta, grads, sizes):
        if gsl is not None:
            wantW[off:off + n] += -gsl * (t - po[off:off + n])
            wantpo[off:off + n] = t
        off += n
    ops.nv.si_update(theta, grads, W, po)
    assert relerr(W, wantW) <= 1e-6 and torch.equal(po, wantpo)
    # penalty over > 32 tensors (none NULL allowed there), incl. the zero-size one
    star = torch.randn(total, device="cuda", generator=gen)
    out = torch.zeros(1, device="cuda")
    ops.nv.ewc_penalty_fwd(theta, fisher, star, 1.5, out)
    th = torch.cat(theta)
    ref = 1.5 * float((fisher.double() * (th.double() - star.double()) ** 2).sum())
    assert abs(float(out) - ref) <= 1e-5 * abs(ref)
    gr = [torch.zeros(n, device="cuda") for n in sizes]
    ops.nv.ewc_penalty_bwd(theta, gr, fisher, star, 3.0, None)
    assert relerr(torch.cat(gr), 3.0 * fisher * (th - star)) <= 1e-6


def test_penalty_gradient_keeps_the_optimizer_step_fused():
    """loss = mse + ewc.penalty() (cfg 4): autograd sums the two gradients of every leaf out of place, so no param.grad
    aliases the engine's flat buffer; FlatAdamW must still step in ONE adamw launch (after a 5-launch gather), and
    the result must equal torch.optim.AdamW on the same gradients."""
    from nerve_cl_b200 import ops
    from nerve_cl_b200.continual import EWC
    from nerve_cl_b200.models import SuperResolutionNet
    from nerve_cl_b200.optim import FlatAdamW
    torch.manual_seed(3)
    model = SuperResolutionNet(num_features=16, num_residual_blocks=1).cuda().train()
    model.compute_dtype = torch.float32
    opt = FlatAdamW(model, lr=1e-3, weight_decay=1e-2)
    x, t = torch.rand(2, 3, 3, 12, 16, device="cuda"), torch.rand(2, 3, 24, 32, device="cuda")
    ewc = EWC(model, ewc_lambda=50.0)
    ewc.register_task(0, [(x, t)])
    model.train()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.01)
    opt.zero_grad()
    torch.nn.functional.mse_loss(model(x), t).backward()
    g1 = [p.grad.clone() for p in model.parameters()]
    assert model.last_flat_grad() is not None
    opt.zero_grad()
    ewc.penalty().backward()
    g2 = [p.grad.clone() for p in model.parameters()]
    opt.zero_grad()
    (torch.nn.functional.mse_loss(model(x), t) + ewc.penalty()).backward()
    for p, a, b in zip(model.parameters(), g1, g2):
        assert relerr(p.grad, a + b) <= 1e-5 or float((p.grad - a - b).abs().max()) < 1e-9
    ref = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
    for r, p in zip(ref, model.parameters()):
        r.grad = p.grad.clone()
    topt = torch.optim.AdamW(ref, lr=1e-3, weight_decay=1e-2)
    topt.step()
    n1 = ops.LAUNCHES[0]
    opt.step()
    launches = ops.LAUNCHES[0] - n1
    assert launches <= 1 + (len(ref) + 31) // 32, launches          # gather (131 tensors: 5) + ONE adamw launch
    for r, p in zip(ref, model.parameters()):
        assert relerr(p, r) <= 1e-6
    # optimiser checkpoints are torch.optim.AdamW's format, both ways
    sd = opt.state_dict()
    topt2 = torch.optim.AdamW([p.detach().clone().requires_grad_(True) for p in model.parameters()], lr=5e-4)
    topt2.load_state_dict(sd)
    assert topt2.param_groups[0]["weight_decay"] == 1e-2 and topt2.param_groups[0]["lr"] == 1e-3
    opt2 = FlatAdamW(model, lr=7e-4)
    opt2.load_state_dict(topt.state_dict())
    assert opt2.step_count == 1 and opt2.param_groups[0]["weight_decay"] == 1e-2
    assert relerr(opt2.exp_avg, opt.exp_avg) <= 1e-5 and relerr(opt2.exp_avg_sq, opt.exp_avg_sq) <= 1e-4
