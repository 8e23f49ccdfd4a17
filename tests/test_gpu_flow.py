"""Flow parity on the GPU: drop-in SupervisedMacowTransformer vs the golden fixtures produced by the reference modules
and vs the CPU oracle; bit-exactness of the shuffle path; invertibility at full size."""
import pytest
import torch

from conftest import golden
from util import O, make_flow, maxabs

pytestmark = pytest.mark.gpu

# tolerances: fp32 engines must sit inside the fp32 noise of the reference itself (fixtures store ref-fp32 vs oracle-fp64)
CASES = [("flow_tiny_even", ["fp32_simt", "fp32", "bf16"]), ("flow_tiny_odd", ["fp32_simt"]),
         ("flow_c32_hd128", ["fp32_simt", "fp32", "bf16"]), ("flow_c64_hd128", ["fp32_simt", "fp32"])]
TOL = {"fp32_simt": 2e-4, "fp32": 3e-4, "bf16": 6e-2}


@pytest.mark.parametrize("name,precs", CASES)
def test_flow_matches_reference_golden(name, precs):
    fx = golden(name)
    cfg = O.flow_config(**fx["cfg_kwargs"])
    sd = O.synth_flow_state_dict(cfg, seed=fx["wseed"])
    z, cond, _ = O.synth_inputs(fx["B"], cfg["flow_in_channels"], cfg["h_channels"], 8, seed=fx["iseed"])
    for prec in precs:
        m = make_flow(cfg, sd, prec)
        x = m(z.cuda(), cond.cuda(), reverse=True)
        z2, ld = m(fx["x_rev"].cuda(), cond.cuda())
        e_rev, e_fwd = maxabs(x, fx["x_rev"]), maxabs(z2, fx["z_fwd"])
        e_ld = maxabs(ld, fx["logdet"])
        print(f"{name} {prec}: rev {e_rev:.2e} fwd {e_fwd:.2e} logdet {e_ld:.2e}")
        assert e_rev < TOL[prec] and e_fwd < TOL[prec], (name, prec, e_rev, e_fwd)
        assert e_ld < TOL[prec] * 50 * max(1.0, fx["logdet"].abs().max().item() / 50)
        # sample() draws from the CPU generator like the reference (INN.py:479)
        torch.manual_seed(5)
        zz = torch.randn(z.shape)
        torch.manual_seed(5)
        xs = m.sample(tuple(z.shape), cond.cuda(), device="cuda")
        assert maxabs(xs, m.reverse(zz.cuda(), cond.cuda())) == 0.0


def test_shuffle_path_bit_exact():
    """With identity couplings (g = 0, zero biases, zero ActNorm) the flow is a pure composition of channel permutations:
    the CUDA result must equal the reference gather bit for bit, forward and inverse."""
    cfg = O.flow_config(flow_in_channels=32, flow_mid_channels=64, h_channels=16, num_steps=[2, 1, 1, 1], factor=16)
    sd = O.synth_flow_state_dict(cfg, seed=77)
    for k in list(sd):
        if k.endswith(("weight_g", "conv.bias", "log_scale")) or k.endswith("actnorm1.bias") or k.endswith("actnorm2.bias") or k.endswith("actnorm.bias"):
            sd[k] = torch.zeros_like(sd[k])
    z, cond, _ = O.synth_inputs(3, 32, 16, 8, seed=3)
    x_ref = O.flow_reverse(sd, cfg, z, cond)
    zf_ref, ld_ref = O.flow_forward(sd, cfg, z, cond)
    assert sorted(x_ref.flatten().tolist()) == sorted(z.flatten().tolist())      # it really is a permutation
    for prec in ("fp32_simt", "fp32"):
        m = make_flow(cfg, sd, prec)
        x = m(z.cuda(), cond.cuda(), reverse=True).cpu()
        zf, ld = m(z.cuda(), cond.cuda())
        assert torch.equal(x, x_ref), prec
        assert torch.equal(zf.cpu(), zf_ref), prec
        assert torch.equal(ld.cpu(), torch.zeros(3))


def test_full_size_flow_matches_reference_golden():
    """iper_128 flow (C0=32, Hd=2048, 15 levels, 1.05 B parameters) against the reference's own output (B=2)."""
    fx = golden("flow_full_c32")
    cfg = O.flow_config(**fx["cfg_kwargs"])
    sd = O.synth_flow_state_dict(cfg, seed=fx["wseed"])
    z, cond, _ = O.synth_inputs(fx["B"], 32, 128, 8, seed=fx["iseed"])
    for prec, tol in (("fp32", 5e-4), ("bf16", 1e-1)):
        m = make_flow(cfg, sd, prec, max_batch=4)
        x = m(z.cuda(), cond.cuda(), reverse=True)
        z2, ld = m(fx["x_rev"].cuda(), cond.cuda())
        e1, e2, e3 = maxabs(x, fx["x_rev"]), maxabs(z2, fx["z_fwd"]), maxabs(ld, fx["logdet"])
        print(f"full flow {prec}: rev {e1:.2e} fwd {e2:.2e} logdet {e3:.2e} (ref fp32 vs fp64: {fx['ref_fp32_vs_oracle_fp64']:.2e})")
        assert e1 < tol and e2 < tol
        # invertibility at full size (size-independent property)
        assert maxabs(m(x, cond.cuda())[0], z) < tol
        del m
        torch.cuda.empty_cache()
