"""CUDA-graph replay of a whole training step (nerve_cl_b200.graphs.GraphedTrainStep) == the same steps run eagerly."""
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_graphed_train_step_matches_eager(dtype):
    from nerve_cl_b200.graphs import GraphedTrainStep
    from nerve_cl_b200.models import SuperResolutionNet
    from nerve_cl_b200.optim import FlatAdamW
    g = torch.Generator().manual_seed(1)
    batches = [(torch.rand(4, 3, 3, 32, 72, generator=g).cuda(), torch.rand(4, 3, 64, 144, generator=g).cuda()) for _ in range(5)]
    models, opts = [], []
    for _ in range(2):
        torch.manual_seed(7)
        m = SuperResolutionNet(scale_factor=2, num_features=32, num_residual_blocks=2).cuda().train()
        m.compute_dtype = dtype
        models.append(m)
        opts.append(FlatAdamW(m, lr=1e-3, weight_decay=1e-5))
    # eager: the capture's 3 warm-up steps are real steps on the first batch, so the eager run takes them too
    losses_e = []
    for lr, hr in [batches[0]] * 3 + batches:
        opts[0].zero_grad()
        loss = torch.nn.functional.mse_loss(models[0](lr), hr)
        loss.backward()
        opts[0].step()
        losses_e.append(float(loss))
    step = GraphedTrainStep(models[1], opts[1], warmup=3)
    losses_g = [float(step(lr, hr)) for lr, hr in batches]
    assert opts[1].step_count == opts[0].step_count == 8 and int(opts[1].step_dev) == 8
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for a, b in zip(losses_e[3:], losses_g):
        assert abs(a - b) <= tol * abs(a), (losses_e, losses_g)
    # AdamW turns the last-bit noise of atomically accumulated gradients into +-lr steps wherever a gradient is near zero
    # (flow_net.0 at this size: two EAGER runs differ the same way), so a relative bound per tensor is a coin toss.  What
    # must hold: no element moves apart faster than two full AdamW steps per iteration, and almost none moves at all.
    n_steps, lr_ = 8, 1e-3
    for (n, p), q in zip(models[0].named_parameters(), models[1].parameters()):
        d = (q.detach() - p.detach()).abs()
        assert float(d.max()) <= 2.02 * n_steps * lr_, n
        assert float(d.mean()) <= (0.05 if dtype == torch.float32 else 0.25) * n_steps * lr_, n
    ptol = 5e-3 if dtype == torch.float32 else 5e-2
    sd0, sd1 = models[0].state_dict(), models[1].state_dict()
    for k in sd0:
        if "tracked" in k:
            assert int(sd0[k]) == int(sd1[k]) == 24        # 8 steps x T = 3 extractor calls each
        elif "running" in k:
            assert relerr(sd1[k], sd0[k]) <= ptol, k
