"""CPU: the oracle (oracle/) reproduces the golden fixtures generated from the LIVE reference
(tests/golden/make_golden.py).  This is what pins the oracle; it needs neither a GPU nor /root/reference."""
import numpy as np
import pytest
import torch

from conftest import build_case, load_golden, relerr
from oracle import ewc_oracle, sr_oracle

SR_CASES = ["sr_tiny_x2_train.npz", "sr_tiny_x2_eval.npz", "sr_tiny_x3_train.npz", "sr_tiny_x4_t5_train.npz"]


@pytest.mark.parametrize("name", SR_CASES)
def test_sr_oracle_matches_reference_outputs(name):
    g = load_golden(name)
    model, scale, training = build_case(g["meta"], "cpu")        # same seed => same weights as the reference
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    if "w_sum" in g:
        sums = np.array([float(v.double().sum()) for v in sd.values()])
        np.testing.assert_allclose(sums, g["w_sum"], rtol=1e-12, atol=1e-12)
    x, tgt = torch.from_numpy(g["lr_frames"]), torch.from_numpy(g["target"])
    out, loss, grads = sr_oracle.train_step_grads(sd, x, tgt, scale, training)
    assert relerr(out, torch.from_numpy(g["out"])) <= 1e-6
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    for k in sd:
        if "running" in k or "tracked" in k:
            assert relerr(sd[k].float(), torch.from_numpy(g["bn1/" + k]).float()) <= 1e-6, k
    if "g_norm" in g:
        norms = np.array([float(v.double().norm()) for v in grads.values()])
        np.testing.assert_allclose(norms, g["g_norm"], rtol=1e-4, atol=1e-6 * float(g["g_norm"].max()))
    else:
        for n, v in grads.items():
            assert relerr(v, torch.from_numpy(g["g/" + n])) <= 1e-5, n


def test_sr_oracle_default_model_forward():
    """The real 64-feature / 8-block network (forward only, to keep the CPU suite short)."""
    g = load_golden("sr_default_x2_train.npz")
    model, scale, training = build_case(g["meta"], "cpu")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    np.testing.assert_allclose(np.array([float(v.double().sum()) for v in sd.values()]), g["w_sum"], rtol=1e-12)
    with torch.no_grad():
        out, inter = sr_oracle.sr_forward(sd, torch.from_numpy(g["lr_frames"]), scale, training, True)
    assert relerr(out, torch.from_numpy(g["out"])) <= 1e-6
    assert relerr(inter["aggregated"], torch.from_numpy(g["aggregated"])) <= 1e-6


def test_warp_oracle_golden():
    g = load_golden("warp_cases.npz")
    for hw in ("9x13", "36x64"):
        feat = torch.from_numpy(g[f"{hw}/feat"])
        for i in range(10):
            flow = torch.from_numpy(g[f"{hw}/{i}/flow"])
            assert torch.equal(sr_oracle.warp(feat, flow), torch.from_numpy(g[f"{hw}/{i}/out"]))
            assert np.array_equal(sr_oracle.warp_corner_indices(flow).numpy(), g[f"{hw}/{i}/idx"].astype(np.int32))


def test_warp_round_trip_is_not_identity():
    """with zero flow floor(ix) != x for some columns at W=640 -- the oracle keeps that."""
    idx = load_golden("warp_cases.npz")["360x640/zero/idx"]
    xs = np.arange(640)[None, None, :]
    frac = float((idx[..., 0] == xs).mean())
    assert 0.5 < frac < 1.0


def test_correlation_oracle_golden():
    g = load_golden("corr_case.npz")
    out = sr_oracle.correlation(torch.from_numpy(g["x1"]), torch.from_numpy(g["x2"]))
    assert relerr(out, torch.from_numpy(g["out"])) <= 1e-6


def test_ewc_oracle_golden():
    g = load_golden("ewc_linear.npz")
    xs, ys = torch.from_numpy(g["xs"]), torch.from_numpy(g["ys"])

    def batch_grads(wflat):
        m = torch.nn.Linear(10, 10)
        with torch.no_grad():
            m.weight.copy_(torch.from_numpy(wflat[:100]).view(10, 10))
            m.bias.copy_(torch.from_numpy(wflat[100:]))
        out = []
        for i in range(0, 40, 8):
            m.zero_grad()
            torch.nn.functional.mse_loss(m(xs[i:i + 8]), ys[i:i + 8]).backward()
            out.append(torch.cat([p.grad.flatten() for p in m.parameters()]).numpy().copy())
        return out
    f0 = ewc_oracle.fisher_from_batches(batch_grads(g["w0"]), [8] * 5)
    np.testing.assert_allclose(f0, g["fisher0"], rtol=1e-6)
    pen = ewc_oracle.penalty(g["w1"], f0, g["w0"], 5000.0)
    assert abs(pen - float(g["penalty1"])) <= 1e-5 * abs(float(g["penalty1"]))
    np.testing.assert_allclose(ewc_oracle.penalty_grad(g["w1"], f0, g["w0"], 5000.0), g["penalty_grad1"],
                               rtol=1e-5, atol=1e-7)
    f1 = ewc_oracle.consolidate(f0, ewc_oracle.fisher_from_batches(batch_grads(g["w1"]), [8] * 5), 0.9)
    np.testing.assert_allclose(f1, g["fisher1"], rtol=1e-5, atol=1e-9)


def _linear_batch_grads(wflat, xs, ys, bs=8):
    m = torch.nn.Linear(10, 10)
    with torch.no_grad():
        m.weight.copy_(torch.from_numpy(wflat[:100]).view(10, 10))
        m.bias.copy_(torch.from_numpy(wflat[100:]))
    out = []
    for i in range(0, xs.shape[0], bs):
        m.zero_grad()
        torch.nn.functional.mse_loss(m(xs[i:i + bs]), ys[i:i + bs]).backward()
        out.append(torch.cat([p.grad.flatten() for p in m.parameters()]).numpy().copy())
    return out


def test_ewc_separate_oracle_golden():
    """'separate'-mode penalty (reference ewc.py:213-223) from the live reference: two tasks, summed penalty."""
    g = load_golden("ewc_separate.npz")
    f0 = ewc_oracle.fisher_from_batches(_linear_batch_grads(g["w0"], torch.from_numpy(g["xa"]), torch.from_numpy(g["ya"])), [8] * 5)
    f1 = ewc_oracle.fisher_from_batches(_linear_batch_grads(g["w1"], torch.from_numpy(g["xb"]), torch.from_numpy(g["yb"])), [8] * 3)
    np.testing.assert_allclose(f0, g["fisher0"], rtol=1e-6)
    np.testing.assert_allclose(f1, g["fisher1"], rtol=1e-6)
    pen = ewc_oracle.penalty_separate(g["w2"], [f0, f1], [g["w0"], g["w1"]], 300.0)
    assert abs(pen - float(g["penalty2"])) <= 1e-5 * abs(float(g["penalty2"]))
    np.testing.assert_allclose(ewc_oracle.penalty_separate_grad(g["w2"], [f0, f1], [g["w0"], g["w1"]], 300.0),
                               g["penalty_grad2"], rtol=1e-5, atol=1e-7)


def si_replay(g, update, register, params, step_grad, apply_step):
    """The SI golden's training loop (tests/golden/make_golden.py::si_case) over abstract callbacks."""
    for i in range(6):
        gr = step_grad(i)
        apply_step(gr)
        update(gr, drop_bias=(i == 2))
    register()
    yield "a"
    for i in range(6, 9):
        gr = step_grad(i)
        apply_step(gr)
        if i != 8:
            update(gr, drop_bias=False)
    register()
    yield "b"


def test_si_oracle_golden():
    """SynapticIntelligence (reference ewc.py:306-379) restated in oracle/ewc_oracle.py vs the live-reference golden."""
    g = load_golden("si_linear.npz")
    xs, ys, lr = torch.from_numpy(g["xs"]), torch.from_numpy(g["ys"]), float(g["lr"])
    st = {"th": g["w_init"].copy(), "W": np.zeros(110, np.float32), "po": g["w_init"].copy(), "om": np.zeros(110, np.float32)}

    def step_grad(i):
        return _linear_batch_grads(st["th"], xs[8 * i:8 * i + 8], ys[8 * i:8 * i + 8])[0]

    def apply_step(gr):
        st["th"] = (st["th"] + np.float32(-lr) * gr).astype(np.float32)

    def update(gr, drop_bias):
        n = 100 if drop_bias else 110
        W, po = ewc_oracle.si_update(st["W"][:n], st["po"][:n], st["th"][:n], gr[:n])
        st["W"] = np.concatenate([W, st["W"][n:]])
        st["po"] = np.concatenate([po, st["po"][n:]])

    def register():
        st["W"], st["po"], st["om"] = ewc_oracle.si_register(st["W"], st["po"], st["om"], st["th"], 0.1)

    for tag in si_replay(g, update, register, None, step_grad, apply_step):
        np.testing.assert_allclose(st["om"], g["omega_" + tag], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(st["po"], g["p_old_b"], rtol=1e-5, atol=1e-7)
    pen = ewc_oracle.si_penalty(g["w_final"], g["omega_b"], g["p_old_b"], 0.7)
    assert abs(pen - float(g["penalty"])) <= 1e-5 * abs(float(g["penalty"]))
    np.testing.assert_allclose(ewc_oracle.si_penalty_grad(g["w_final"], g["omega_b"], g["p_old_b"], 0.7),
                               g["penalty_grad"], rtol=1e-5, atol=1e-8)
