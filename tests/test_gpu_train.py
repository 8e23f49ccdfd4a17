"""Second-stage training step on the GPU (BASELINE configs[3]): loss and every parameter gradient of the native step against the
reference's own loss.backward() (golden fixtures), the optimizer against torch.optim.Adam(amsgrad=True), and a short descent."""
import pytest
import torch

from conftest import golden
from util import O, maxabs

pytestmark = pytest.mark.gpu


def _trainer(cfg, sd, precision, max_batch):
    import ipoke_b200 as ipk
    c = dict(cfg); c.update(ipk_precision=precision, ipk_max_batch=max_batch)
    m = ipk.SupervisedMacowTransformer(c)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    return m, ipk.FlowTrainer(m, max_batch=max_batch, precision=precision)


# stated tolerances: loss relative 1e-5; gradients relative to each tensor's own max-abs:
# fp32_simt 2e-4, fp32 (bf16x3 tensor cores) 1e-3
@pytest.mark.parametrize("precision,gtol", [("fp32_simt", 2e-4), ("fp32", 1e-3)])
@pytest.mark.parametrize("name", ["flowgrad_tiny", "flowgrad_c32_hd128", "flowgrad_c64_hd128"])
def test_training_step_matches_reference_autograd(name, precision, gtol):
    fx = golden(name)
    cfg = O.flow_config(**fx["cfg_kwargs"])
    sd = O.synth_flow_state_dict(cfg, seed=fx["wseed"])
    x, cond, _ = O.synth_inputs(fx["B"], cfg["flow_in_channels"], cfg["h_channels"], 8, seed=fx["iseed"])
    m, tr = _trainer(cfg, sd, precision, fx["B"])
    loss, z, ld = tr.step((x * 0.8).cuda(), cond.cuda(), return_latent=True)
    rel = abs(loss.item() - fx["loss"]) / abs(fx["loss"])
    # the training forward agrees with the inference forward of the same module
    z2, ld2 = m((x * 0.8).cuda(), cond.cuda())
    assert maxabs(z, z2) < 5e-4 and maxabs(ld, ld2) < 5e-3
    worst, worst_k, worst_n = 0.0, None, 0.0
    for i, k in enumerate(fx["keys"]):
        g = tr.grad(k).detach().cpu()
        scale = fx["maxabs"][i].item() + 1e-12
        e = (g.flatten()[fx["idx"][i]] - fx["val"][i]).abs().max().item() / scale
        en = abs(g.double().norm().item() - fx["norms"][i].item()) / (fx["norms"][i].item() + 1e-12)
        if e > worst:
            worst, worst_k = e, k
        worst_n = max(worst_n, en)
    print(f"{name}[{precision}]: loss {loss.item():.6f} vs {fx['loss']:.6f} (rel {rel:.2e}); {len(fx['keys'])} gradients, worst sampled error "
          f"{worst:.2e} of max-abs ({worst_k}), worst norm error {worst_n:.2e}")
    assert rel < 1e-5
    assert worst < gtol and worst_n < gtol
    # the step is replayed as a CUDA graph from its third call on: same loss, same gradients
    g1 = tr.flat_grads.clone()
    for _ in range(3):
        loss3 = tr.step((x * 0.8).cuda(), cond.cuda())
    assert abs(loss3.item() - loss.item()) < 1e-6 * abs(loss.item())
    assert (tr.flat_grads - g1).abs().max().item() <= 1e-6 * g1.abs().max().item()


def test_adam_matches_torch_and_loss_decreases():
    cfg = O.flow_config(flow_in_channels=16, flow_mid_channels=64, h_channels=16, num_steps=[2, 1, 1], factor=4)
    sd = O.synth_flow_state_dict(cfg, seed=1)
    x, cond, _ = O.synth_inputs(4, 16, 16, 8, seed=11)
    m, tr = _trainer(cfg, sd, "fp32_simt", 4)
    ref_p = tr.flat_params[:tr.numel].clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-5, amsgrad=True)
    losses = []
    for it in range(4):
        losses.append(tr.step((x * 0.8).cuda(), cond.cuda()).item())
        ref_p.grad = tr.flat_grads[:tr.numel].clone()
        opt.step()
        tr.optimizer_step(lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-5, amsgrad=True)
        assert (tr.flat_params[:tr.numel] - ref_p.detach()).abs().max().item() < 1e-6, it
    print("losses", losses)
    assert losses[-1] < losses[0]
    # the drop-in module samples with the updated parameters (they are views of the flat buffer)
    z = torch.randn(2, 16, 8, 8).cuda()
    xs = m(z, cond[:2].cuda(), reverse=True)
    z2, _ = m(xs, cond[:2].cuda())
    assert maxabs(z2, z) < 1e-3
