"""Training support against records of the UNMODIFIED reference's own training step
(tests/golden/make_golden.py:train_case -- train.py:93-123: model.train(), forward, deepVCP_loss(alpha=0.5),
backward, Adam.step). The index stages run on the sm_100a kernels, the differentiable stages through autograd;
held here: the train-mode forward (batch-statistics BatchNorm), the loss value, the gradient of every parameter,
which parameters get none, and the whole state_dict after one optimiser step (running statistics included)."""
import importlib

import pytest
import torch

from conftest import PKG, golden_state_dict, load_golden

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda"


@pytest.fixture(scope="module")
def dv():
    return importlib.import_module(PKG)


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build(dv, g):
    N = int(g["n_points"])
    src, tgt, R, t = T(g["src"]), T(g["tgt"]), T(g["R"]), T(g["t"])
    model = dv.DeepVCP(use_normal=src.shape[1] == 6, npoint=N, r=float(g["r"]), s=float(g["s"]))
    model.load_state_dict(golden_state_dict(g))
    model = model.to(DEV)
    starts = tuple(torch.tensor([int(v)]) for v in g["starts"])
    return model, src.to(DEV), tgt.to(DEV), R.to(DEV), t.to(DEV), starts


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", ["train_modelnet_n512_g5", "train_modelnet_f64_n512_g5"])
def test_training_step_vs_reference_record(dv, name, fused):
    """fused=True: the target-side embedding runs as one kernel forward (dvcp_dfe_tgt_fused) and one kernel backward
    (dvcp_dfe_tgt_backward); False: the same stage through autograd on the materialised float64 tensor."""
    g = load_golden(name)
    model, src, tgt, R, t, starts = build(dv, g)
    model.train()
    lr = float(g["lr"])
    optim = torch.optim.Adam(model.parameters(), lr=lr)
    with dv.training.fp32_math():
        # the key-point choice is the reference's own draw from its scores (top-k ties are unspecified)
        kp, vcp = dv.training.forward(model, src, tgt, R, torch.zeros(1, 3), starts=starts, keep_stages=True,
                                      topk_override=T(g["topk_idx"]).long().view(1, -1), fused_embedding=fused)
        assert vcp.requires_grad and kp.dtype == src.dtype and vcp.dtype == torch.float32
        # train-mode features (batch statistics) and the forward outputs
        assert rel_err(model.last["src_fe_feat"], T(g["src_fe_feat"])) < 1e-5
        assert torch.equal(kp.cpu(), T(g["src_keypts"]))
        assert (vcp.detach().cpu() - T(g["vcp"])).abs().max() < 2e-5
        # the free-running top-k picks the same score values
        sc = model.last["scores"][0, :, 0]
        mine = torch.topk(sc, 64).values
        assert torch.allclose(mine, sc[T(g["topk_idx"]).long().to(DEV).view(-1)], rtol=1e-6, atol=0)
        optim.zero_grad()
        loss, Rp, tp = dv.deepVCP_loss(kp, vcp, R, t.view(1, 3, 1), alpha=0.5)
        loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-6 * max(1.0, abs(float(g["loss"])))
    assert (Rp.detach().cpu() - T(g["R_pred"])).abs().max() < 1e-6
    assert (tp.detach().cpu() - T(g["t_pred"])).abs().max() < 1e-5
    with_grad = {k[5:] for k in g if k.startswith("grad/")}
    assert len(with_grad) == 24
    for k, p in model.named_parameters():
        if k in with_grad:
            assert p.grad is not None, k
            ref = T(g["grad/" + k])
            if ("mlp_convs" in k and k.endswith(".bias")) or k == "cpg.conv3.bias":
                # a bias in front of a train-mode BatchNorm has gradient zero (the mean is subtracted), and so has
                # the bias in front of the softmax over candidates (shift invariance): both sides hold rounding
                # noise only, far below the gradient of the weight beside it
                scale = float(T(g["grad/" + k[:-4] + "weight"]).abs().max())
                assert float(ref.abs().max()) < 1e-2 * scale and float(p.grad.abs().max()) < 1e-2 * scale, k
                continue
            e = rel_err(p.grad, ref)
            assert e < 2e-3, "gradient of %s: %g" % (k, e)
        else:
            # weighting layer (indices only), sa2 / sa3 / fc (never called: SURVEY Q1)
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
    optim.step()
    after = golden_state_dict(g, "sd_after/")
    for k, v in model.state_dict().items():
        ref = after[k]
        if ref.dtype in (torch.int64, torch.int32):
            assert int(v) == int(ref), k                        # num_batches_tracked
        elif "running_" in k:
            assert torch.allclose(v.detach().cpu().double(), ref.double(), rtol=1e-5, atol=1e-7), k
        else:
            # Adam's first step moves every element by lr * g / (|g| + eps): an element whose gradient is rounding
            # noise (the zero-gradient biases above, |g| near eps) may land anywhere within +-lr of its start on
            # either side; everything else must agree closely
            d = (v.detach().cpu().double() - ref.double()).abs()
            assert float(d.max()) <= 2 * lr + 1e-6, k
            noisy = ("mlp_convs" in k and k.endswith(".bias")) or k == "cpg.conv3.bias"
            if not noisy:
                frac = float((d < 2e-5).double().mean())
                assert frac > 0.97, "%s: %g of the elements agree" % (k, frac)


def test_set_abstraction_module_in_train_mode_uses_batch_statistics(dv):
    g = load_golden("train_modelnet_n512_g5")
    model, src, tgt, R, t, starts = build(dv, g)
    sa = model.FE1.sa1
    sa.train()
    before = sa.mlp_bns[0].running_mean.clone()
    with dv.training.fp32_math():
        _, feats = sa(src[:, :3, :], src[:, 3:, :], start=starts[0])
    assert feats.requires_grad and feats.shape == (1, 32, 512)
    assert rel_err(feats.permute(0, 2, 1), T(g["src_fe_feat"])) < 1e-5
    assert not torch.equal(before, sa.mlp_bns[0].running_mean)
    sa.eval()
    _, feats_eval = sa(src[:, :3, :], src[:, 3:, :], start=starts[0])
    assert not feats_eval.requires_grad


def test_training_batch_of_pairs_runs_and_reduces_the_loss(dv, synthetic):
    """B = 2 (independent pairs, batch statistics over both): a few Adam steps on a fixed batch lower the loss."""
    torch.manual_seed(0)
    N = 512
    src, tgt, R, t = synthetic.make_batch("modelnet", [0, 1], N)
    model = dv.DeepVCP(use_normal=True, npoint=N, r=0.8, s=0.4).to(DEV).train()
    optim = torch.optim.Adam(model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(1)
    starts = (torch.randint(0, N, (2,), generator=g), torch.randint(0, 64, (2,), generator=g),
              torch.randint(0, N, (2,), generator=g))
    losses = []
    for _ in range(6):
        kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts)
        optim.zero_grad()
        loss, Rp, tp = dv.deepVCP_loss(kp, vcp, R.to(DEV), t.view(2, 3, 1).to(DEV), alpha=0.5)
        loss.backward()
        optim.step()
        losses.append(loss.item())
        rot, trans = dv.metrics.registration_errors(Rp.detach(), tp.detach(), R.to(DEV), t.view(2, 3, 1).to(DEV))
        assert torch.isfinite(rot).all() and torch.isfinite(trans).all()
    assert all(l == l for l in losses) and min(losses[1:]) < losses[0]
    model.eval()
    with torch.no_grad():
        kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3), starts=starts)
    assert torch.isfinite(vcp).all()


@pytest.mark.parametrize("quirks", [63, 0, 63 - 32, 63 - 2 - 4])
def test_autograd_forward_equals_the_inference_kernels_in_eval_mode(dv, synthetic, quirks):
    """training.forward follows the module's mode: with eval-mode BatchNorm it must reproduce the fused inference
    kernels (reference mode and the intended-semantics switches alike), stage by stage."""
    N = 1024
    src, tgt, R, t = synthetic.make_batch("modelnet", [41, 42], N)
    torch.manual_seed(13)
    model = dv.DeepVCP(use_normal=True, npoint=N, r=0.8, s=0.4, quirks=quirks).to(DEV).eval()
    starts = (torch.tensor([1, 2]), torch.tensor([3, 4]), torch.tensor([5, 6]))
    t_init = torch.tensor([[0.25, -0.5, 0.125]])
    kp, vcp = model(src.to(DEV), tgt.to(DEV), R.to(DEV), t_init, starts=starts, keep_stages=True)
    A = dict(model.last)
    with dv.training.fp32_math():
        kp2, vcp2 = dv.training.forward(model, src.to(DEV), tgt.to(DEV), R.to(DEV), t_init, starts=starts,
                                        keep_stages=True, topk_override=A["topk_idx"])
    Bm = model.last
    assert vcp2.requires_grad
    assert torch.equal(A["src_fps"].long(), Bm["src_fps"]) and torch.equal(A["tgt_fps"].long(), Bm["tgt_fps"])
    assert rel_err(Bm["src_fe_feat"], A["src_fe_feat"]) < 1e-5
    assert torch.equal(kp, kp2) and torch.equal(A["picked_idx"], Bm["picked_idx"])
    assert torch.equal(A["candidates"], Bm["candidates"]) and torch.equal(A["knn_idx"], Bm["knn_idx"])
    assert rel_err(Bm["src_dfe"], A["src_dfe"]) < 1e-5
    assert rel_err(Bm["tgt_dfe"], A["tgt_dfe"]) < 2e-5
    assert (vcp2 - vcp).abs().max() < 5e-5
