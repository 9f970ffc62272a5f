"""GPU parity of the full Model (silent_speech_b200.architecture) against the golden vectors
produced by the executed reference and against the torch oracle: eval/train outputs, loss,
every parameter gradient, BatchNorm running statistics, the in-place input shift, and
`relative_positional.embeddings.grad is None`.
North-star tolerance: 1e-3 relative fp32.  This fp32 path is asserted at 1e-4."""
import os
import random

import numpy as np
import pytest
import torch
from absl import flags

from make_golden_model import CASES, grad_fingerprint, make_input, scalar_loss
from oracle import model as om

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def max_rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "model_golden.npz"))


def build(D, NL, dropout=0.0):
    from silent_speech_b200 import architecture as A
    F = flags.FLAGS
    if not F.is_parsed():
        F(["test"])
    F.model_size, F.num_layers, F.dropout = D, NL, dropout
    m = A.Model(112, 80, 48)
    sd = om.formula_state_dict(D, NL)
    assert sorted(m.state_dict().keys()) == sorted(sd.keys())      # checkpoint contract
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    return m.cuda()


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_model_matches_reference_golden(golden, ci, monkeypatch):
    """Semantics parity with the executed reference, on the exact-fp32 CUDA-core engine (the
    formula-weight cases amplify rounding by ~600x at D >= 64, see DESIGN.md 'Accuracy')."""
    monkeypatch.setenv("SSB_GEMM", "simt")
    name, D, NL, B, L, pyseed, inp = CASES[ci]
    m = build(D, NL)
    x = make_input(B, L, inp)

    m.eval()
    with torch.no_grad():
        pred, aux = m(None, x.clone().cuda(), None)
    assert pred.is_contiguous() and aux.is_contiguous()
    for got, key in ((pred, "eval_pred"), (aux, "eval_aux")):
        want = golden[f"{name}_{key}"]
        assert rel_l2(got.cpu().numpy(), want) < TOL and max_rel(got.cpu().numpy(), want) < TOL, key

    m.train()
    random.seed(pyseed)
    xt = x.clone().cuda()
    pred, aux = m(None, xt, None)
    loss = scalar_loss(pred.cpu(), aux.cpu())
    loss.backward()
    assert rel_l2(pred.detach().cpu().numpy(), golden[f"{name}_train_pred"]) < TOL
    assert rel_l2(aux.detach().cpu().numpy(), golden[f"{name}_train_aux"]) < TOL
    np.testing.assert_array_equal(xt[:, -9:, :].cpu().numpy(), golden[f"{name}_train_x_after"])
    assert abs(loss.item() - float(golden[f"{name}_train_loss"])) < 1e-4 * max(1.0, abs(loss.item()))
    grads = {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}
    fp = grad_fingerprint(grads)
    n = 0
    for key in golden.files:
        if key.startswith(f"{name}_grad::"):
            k = key.split("::", 1)[1]
            want = golden[key]
            if k.startswith("conv_blocks") and k.endswith(("conv1.bias", "conv2.bias",
                                                           "residual_path.bias")):
                # a bias in front of a training-mode BatchNorm has an analytically ZERO gradient;
                # the reference's value is rounding noise (1e-7) -- only require ours to be noise too
                assert fp[k][0] < 1e-4, (k, fp[k][0])
                n += 1
                continue
            # fingerprint = [L2 norm, sum, 4 samples]; compare norm tightly, samples vs the norm scale
            assert abs(fp[k][0] - want[0]) <= 2e-4 * want[0] + 1e-9, (k, fp[k][0], want[0])
            scale = want[0] / np.sqrt(max(grads[k].numel(), 1)) + 1e-12
            assert np.abs(fp[k][2:] - want[2:]).max() < 1e-2 * scale + 1e-7, (k, fp[k], want)
            n += 1
    assert n >= 40
    for k, p in m.named_parameters():
        if k.endswith("relative_positional.embeddings"):
            assert p.grad is None
    sd = m.state_dict()
    for k in sd:
        if "running_" in k:
            assert rel_l2(sd[k].cpu().numpy(), golden[f"{name}_buf::{k}"]) < 1e-5, k
        if "num_batches" in k:
            assert int(sd[k]) == int(golden[f"{name}_buf::{k}"])


def random_case(D, NL, B, L, seed):
    """Well-conditioned problem: the Model's own default initialisation (seeded) and randn
    inputs; fp32 vs fp64 CPU oracle agree to 1e-6 on it (checked when the test was written)."""
    from silent_speech_b200 import architecture as A
    F = flags.FLAGS
    if not F.is_parsed():
        F(["test"])
    F.model_size, F.num_layers, F.dropout = D, NL, 0.0
    torch.manual_seed(seed)
    m = A.Model(112, 80, 48)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randn(B, L, 8, generator=torch.Generator().manual_seed(seed + 1))
    return m.cuda().train(), sd0, x


def oracle_grads(sd0, x, pyseed, gp, ga):
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
              else v.clone()) for k, v in sd0.items()}
    random.seed(pyseed)
    op, oa = om.model_forward(sd, x.clone(), training=True, dropout_p=0.0)
    ((op * gp).sum() + (oa * ga).sum()).backward()
    return op.detach(), oa.detach(), sd


@pytest.mark.parametrize("engine,D", [("simt", 64), ("tc", 64), ("simt", 128), ("tc", 128)])
def test_random_init_forward_backward_vs_oracle(engine, D, monkeypatch):
    """Outputs and every parameter gradient vs the CPU fp32 oracle, T = 150 (> band).
    CUDA-core engine: outputs 1e-4, gradients 1e-3.  tcgen05 bf16x3 engine: outputs 1e-4; gradients are
    limited by ReLU kinks (a forward perturbation of 4e-6 flips ~1e-5 of the masks, which moves
    gradients by ~sqrt(that) ~ 2e-3), so they are held to 2e-2 rel-L2 and cosine > 0.9995."""
    monkeypatch.setenv("SSB_GEMM", engine)
    m, sd0, x = random_case(D, 2, 3, 1200, seed=0)
    random.seed(11)
    pred, aux = m(None, x.clone().cuda(), None)
    gp = torch.randn(pred.shape, generator=torch.Generator().manual_seed(5))
    ga = torch.randn(aux.shape, generator=torch.Generator().manual_seed(6))
    ((pred * gp.cuda()).sum() + (aux * ga.cuda()).sum()).backward()
    op, oa, sd = oracle_grads(sd0, x, 11, gp, ga)
    assert rel_l2(pred.detach().cpu().numpy(), op.numpy()) < TOL
    assert max_rel(aux.detach().cpu().numpy(), oa.numpy()) < TOL
    tol = 1e-3 if engine == "simt" else 2e-2   # even 1e-7 forward noise flips a few ReLU masks
    worst, worst_cos = ("", 0.0), ("", 1.0)
    for k, p in m.named_parameters():
        if sd[k].grad is None:
            assert p.grad is None, k
            continue
        if k.startswith("conv_blocks") and k.endswith(("conv1.bias", "conv2.bias",
                                                       "residual_path.bias")):
            continue                                   # analytically zero (BatchNorm follows)
        a, b = p.grad.cpu().double().flatten(), sd[k].grad.double().flatten()
        r = ((a - b).norm() / (b.norm() + 1e-30)).item()
        c = (a @ b / (a.norm() * b.norm() + 1e-30)).item()
        if r > worst[1]:
            worst = (k, r)
        if c < worst_cos[1]:
            worst_cos = (k, c)
    assert worst[1] < tol, worst
    assert worst_cos[1] > 0.9995, worst_cos


def test_cfg1_width_forward_parity_tensor_cores():
    """BASELINE cfg-1 architecture (768-dim, 6 layers, L = 4000 -> T = 500) on the tcgen05
    engine vs the CPU fp32 oracle, eval and train mode.  North-star bar: 1e-3 relative."""
    m, sd0, x = random_case(768, 6, 2, 4000, seed=1)
    m.eval()
    with torch.no_grad():
        pred, aux = m(None, x.clone().cuda(), None)
        sd = {k: v.clone() for k, v in sd0.items()}
        op, oa = om.model_forward(sd, x.clone(), training=False)
    for got, want in ((pred, op), (aux, oa)):
        assert rel_l2(got.cpu().numpy(), want.numpy()) < 2e-4
        assert max_rel(got.cpu().numpy(), want.numpy()) < 1e-3
    m.train()
    random.seed(2)
    pred, aux = m(None, x.clone().cuda(), None)
    random.seed(2)
    op, oa = om.model_forward(sd, x.clone(), training=True, dropout_p=0.0)
    assert rel_l2(pred.detach().cpu().numpy(), op.detach().numpy()) < 2e-4
    assert max_rel(aux.detach().cpu().numpy(), oa.detach().numpy()) < 1e-3


def test_train_mode_dropout_runs_and_differs():
    m = build(32, 1, dropout=0.2)
    m.train()
    x = make_input(2, 400, 3).cuda()
    torch.manual_seed(0)
    random.seed(0)
    a, _ = m(None, x.clone(), None)
    random.seed(0)
    b, _ = m(None, x.clone(), None)
    assert not torch.equal(a, b)                 # fresh dropout seed per forward
    torch.manual_seed(0)
    random.seed(0)
    c, _ = m(None, x.clone(), None)
    assert torch.equal(a, c)                     # torch.manual_seed controls it
    a.sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
