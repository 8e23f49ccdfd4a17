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
# flowgrad_full_c64 is the actual h36m_128 shape of BASELINE configs[3] (C0 = 64, Hd = 2048, 1.24 B parameters, 5 305 gradient tensors), B = 2
@pytest.mark.parametrize("name,precision,gtol", [(n, p, t) for n in ("flowgrad_tiny", "flowgrad_c32_hd128", "flowgrad_c64_hd128")
                                                 for p, t in (("fp32_simt", 2e-4), ("fp32", 1e-3))] + [("flowgrad_full_c64", "fp32", 1e-3)])
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


def _fresh_module(fx, precision):
    """Our module in the fixture's pre-init state: reference ActNorm parameters / permutations, own random conv weights, flags 0."""
    import ipoke_b200 as ipk
    cfg = O.flow_config(**fx["cfg_kwargs"])
    c = dict(cfg); c.update(ipk_precision=precision, ipk_max_batch=fx["B"])
    torch.manual_seed(5)
    m = ipk.SupervisedMacowTransformer(c)
    sd = m.state_dict()
    for k, v in fx["sd0"].items():
        assert sd[k].shape == v.shape, k
        sd[k] = v.clone()
    for k in list(sd):
        if k.endswith("weight_g"):
            sd[k] = torch.full_like(sd[k], 0.3)          # anything non-zero: the init pass must overwrite it with 0
        if k.endswith("initialized"):
            sd[k] = torch.zeros_like(sd[k])
    m.load_state_dict(sd, strict=True)
    return cfg, m.cuda()


@pytest.mark.parametrize("via", ["forward", "trainer"])
@pytest.mark.parametrize("name", ["flowinit_tiny", "flowinit_c32"])
def test_data_dependent_init_matches_reference(name, via):
    """First density-direction call of a fresh flow: ActNorm2dFlow.init (macow2.py:503-505,526-539) and Conv2dWeightNorm.init
    (macow_utils.py:231-250) against what the reference's first train() forward left in its state-dict, and its outputs."""
    import ipoke_b200 as ipk
    fx = golden(name)
    cfg, m = _fresh_module(fx, "fp32_simt")
    x, cond = fx["x"].cuda(), fx["cond"].cuda()
    if via == "forward":
        with torch.no_grad():
            z, ld = m.train()(x, cond)
    else:
        tr = ipk.FlowTrainer(m.train(), max_batch=fx["B"], precision="fp32_simt")
        _, z, ld = tr.step(x, cond, return_latent=True)
    sd1 = m.state_dict()
    worst = 0.0
    for k, v in fx["changed"].items():
        worst = max(worst, maxabs(sd1[k].float(), v.float()))
    print(f"{name}[{via}]: {len(fx['changed'])} tensors changed by the reference, worst abs difference {worst:.2e}; z {maxabs(z, fx['z']):.2e} "
          f"logdet {maxabs(ld, fx['logdet']):.2e}")
    assert worst < 2e-5                                   # stated: 2e-5 absolute on log_scale / bias / weight_g (fp64 vs fp32 statistics)
    assert all(int(v) == 1 for k, v in sd1.items() if k.endswith("initialized"))
    assert maxabs(z, fx["z"]) < 1e-4 and maxabs(ld, fx["logdet"]) < 2e-3 * 1.0 + 1e-6 * abs(fx["logdet"][0].item()) * 100
    # the second call is a plain forward with the initialised parameters
    with torch.no_grad():
        z2, ld2 = m.eval()(x, cond)
    assert maxabs(z2, z) < 1e-4


def test_reverse_of_uninitialised_flow_zeroes_weightnorm_only():
    """Sampling through a never-initialised flow: Conv2dWeightNorm.init fires in either direction (macow_utils.py:248-250, zero_init ->
    g = 0, bias = 0) while ActNorm2dFlow initialises only in the density direction (macow2.py:503)."""
    fx = golden("flowinit_tiny")
    cfg, m = _fresh_module(fx, "fp32_simt")
    z = fx["x"].cuda()
    with torch.no_grad():
        x = m.eval()(z, fx["cond"].cuda(), reverse=True)
    sd1 = m.state_dict()
    assert all(float(v.abs().max()) == 0.0 for k, v in sd1.items() if k.endswith("conv.weight_g"))
    assert all(int(v) == 1 for k, v in sd1.items() if k.endswith("initialized") and "actnorm" not in k)
    flags_act = [int(v) for k, v in sd1.items() if k.endswith("initialized") and "actnorm" in k]
    assert flags_act and not any(flags_act)
    # with identity couplings the inverse is the chain of ActNorm inverses / un-shuffles of the oracle on the same state
    sd_cpu = {k: v.cpu() for k, v in sd1.items()}
    want = O.flow_reverse(sd_cpu, cfg, fx["x"], fx["cond"])
    assert maxabs(x, want) < 1e-4


@pytest.mark.parametrize("precision,gtol", [("fp32_simt", 2e-4), ("fp32", 1e-3)])
def test_autograd_transparent_forward_fills_parameter_grads(precision, gtol):
    """`out, logdet = flow(x, cond); loss, log = FlowLoss()(out, logdet); loss.backward()` -- the reference's training_step
    (second_stage_video.py:409-415) -- against the reference's own gradients."""
    import ipoke_b200 as ipk
    fx = golden("flowgrad_tiny")
    cfg = O.flow_config(**fx["cfg_kwargs"])
    sd = O.synth_flow_state_dict(cfg, seed=fx["wseed"])
    x, cond, _ = O.synth_inputs(fx["B"], cfg["flow_in_channels"], cfg["h_channels"], 8, seed=fx["iseed"])
    c = dict(cfg); c.update(ipk_precision=precision, ipk_max_batch=fx["B"])
    m = ipk.SupervisedMacowTransformer(c)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-5, amsgrad=True)     # built BEFORE the first forward
    crit = ipk.FlowLoss(spatial_mean=False, logdet_weight=1.0)
    xin = (x * 0.8).cuda().requires_grad_(True)
    out, logdet = m(xin, cond.cuda())
    assert out.grad_fn is not None and logdet.grad_fn is not None
    loss, log = crit(out, logdet)
    assert set(log) == {"flow_loss", "reference_nll_loss", "nlogdet_loss", "nll_loss", "logdet_weight"}
    loss.backward()
    assert abs(loss.item() - fx["loss"]) < 1e-5 * abs(fx["loss"])
    named = dict(m.named_parameters())
    worst = 0.0
    for i, k in enumerate(fx["keys"]):
        g = named[k].grad
        assert g is not None, k
        e = (g.detach().cpu().flatten()[fx["idx"][i]] - fx["val"][i]).abs().max().item() / (fx["maxabs"][i].item() + 1e-12)
        worst = max(worst, e)
    assert worst < gtol
    # input gradient: against the oracle's autograd
    xo = (x * 0.8).clone().requires_grad_(True)
    zo, ldo = O.flow_forward(sd, cfg, xo, cond)
    O.flow_nll(zo, ldo).backward()
    assert maxabs(xin.grad, xo.grad) < gtol * xo.grad.abs().max().item()
    # gradient accumulation over two backward passes doubles the gradient (fresh storage per backward)
    g0 = named[fx["keys"][3]].grad.clone()
    out, logdet = m(xin, cond.cuda())
    crit(out, logdet)[0].backward()
    assert maxabs(named[fx["keys"][3]].grad, 2 * g0) < 1e-5 * g0.abs().max().item() + 1e-12
    # an optimizer created before the first forward still drives the parameters the native plan reads
    before = m(xin.detach(), cond.cuda())[0].detach().clone()
    opt.step()
    after = m(xin.detach(), cond.cuda())[0].detach()
    assert maxabs(before, after) > 0
    print(f"autograd path[{precision}]: worst sampled gradient error {worst:.2e} of max-abs")


def test_adam_at_reference_hyperparameters_and_trainer_checkpoint():
    """configure_optimizers (second_stage_video.py:647-648): Adam(betas=(0.9, 0.999), weight_decay, amsgrad=True); the sharded state
    survives a state_dict() / load_state_dict() round trip into a fresh trainer."""
    cfg = O.flow_config(flow_in_channels=16, flow_mid_channels=64, h_channels=16, num_steps=[2, 1, 1], factor=4)
    sd = O.synth_flow_state_dict(cfg, seed=1)
    x, cond, _ = O.synth_inputs(4, 16, 16, 8, seed=11)
    m, tr = _trainer(cfg, sd, "fp32_simt", 4)
    ref_p = tr.flat_params[:tr.numel].clone().requires_grad_(True)
    hp = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, amsgrad=True)
    opt = torch.optim.Adam([ref_p], **hp)
    for it in range(3):
        tr.step((x * 0.8).cuda(), cond.cuda())
        ref_p.grad = tr.flat_grads[:tr.numel].clone()
        opt.step()
        tr.optimizer_step(**hp)
        assert (tr.flat_params[:tr.numel] - ref_p.detach()).abs().max().item() < 2e-7, it
    st = tr.state_dict()
    ost = opt.state[ref_p]
    assert st["step"] == 3
    for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
        assert maxabs(st[k], ost[k]) <= 1e-6 * float(ost[k].abs().max())
    # resume in a fresh trainer on a fresh module: the next step equals the uninterrupted run
    m2, tr2 = _trainer(cfg, {k: v.detach().cpu() for k, v in m.state_dict().items()}, "fp32_simt", 4)
    tr2.load_state_dict(st)
    for t in (tr, tr2):
        t.step((x * 0.8).cuda(), cond.cuda())
        t.optimizer_step(**hp)
    assert maxabs(tr.flat_params, tr2.flat_params) == 0.0
    per = tr2.per_parameter_state()
    assert set(per) == set(tr2.names) and per[tr2.names[0]]["exp_avg"].shape == dict(m2.named_parameters())[tr2.names[0]].shape


def test_trainer_step_returns_flowloss_log():
    cfg = O.flow_config(flow_in_channels=16, flow_mid_channels=64, h_channels=16, num_steps=[2, 1, 1], factor=4)
    sd = O.synth_flow_state_dict(cfg, seed=1)
    x, cond, _ = O.synth_inputs(4, 16, 16, 8, seed=11)
    m, tr = _trainer(cfg, sd, "fp32_simt", 4)
    torch.manual_seed(9); torch.cuda.manual_seed(9)
    loss, log = tr.step((x * 0.8).cuda(), cond.cuda(), return_log=True)
    torch.manual_seed(9); torch.cuda.manual_seed(9)
    z, ld = m((x * 0.8).cuda(), cond.cuda())
    want_ref = (0.5 * (torch.randn_like(z) ** 2).flatten(1).sum(dim=1)).mean()
    assert abs(log["reference_nll_loss"].item() - want_ref.item()) < 1e-4 * want_ref.item()
    assert abs((log["nll_loss"] + log["nlogdet_loss"]).item() - loss.item()) < 1e-5 * abs(loss.item())
    assert log["logdet_weight"] == 1.0
