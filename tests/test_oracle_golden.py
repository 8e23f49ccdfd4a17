"""The CPU oracle (oracle/ipoke_oracle.py) against the golden fixtures produced by the UNMODIFIED reference
modules (tests/golden/make_golden.py).  Pins the oracle; runs anywhere (no GPU, no /root/reference)."""
import pytest
import torch

from conftest import golden
from oracle import ipoke_oracle as O

FLOW = ["flow_tiny_even", "flow_tiny_odd", "flow_c32_hd128", "flow_c64_hd128"]
FS = ["fs_64", "fs_128", "fs_64_z64"]
TOL = 2e-4  # fp32 re-association noise of a 1000-layer flow; fixtures record ref-fp32 vs oracle-fp64 of the same size


@pytest.mark.parametrize("name", FLOW)
def test_flow_oracle_matches_reference(name):
    fx = golden(name)
    cfg = O.flow_config(**fx["cfg_kwargs"])
    sd = O.synth_flow_state_dict(cfg, seed=fx["wseed"])
    z, cond, _ = O.synth_inputs(fx["B"], cfg["flow_in_channels"], cfg["h_channels"], 8, seed=fx["iseed"])
    with torch.no_grad():
        x = O.flow_reverse(sd, cfg, z, cond)
        z2, ld = O.flow_forward(sd, cfg, fx["x_rev"], cond)
    assert (x - fx["x_rev"]).abs().max().item() < TOL
    assert (z2 - fx["z_fwd"]).abs().max().item() < TOL
    assert (ld - fx["logdet"]).abs().max().item() < 1e-3 * max(1.0, fx["logdet"].abs().max().item())
    # invertibility (SURVEY.md section 8c known-answer property i)
    assert (z2 - z).abs().max().item() < 1e-3


@pytest.mark.parametrize("name", FS)
def test_first_stage_oracle_matches_reference(name):
    fx = golden(name)
    cfg = O.first_stage_config(**fx["cfg_kwargs"])
    sd = O.synth_first_stage_state_dict(cfg, seed=fx["wseed"])
    g = torch.Generator().manual_seed(fx["iseed"])
    motion = torch.randn((fx["B"], cfg["z_dim"], 8, 8), generator=g) * 1.3
    x0 = torch.rand((fx["B"], 3, cfg["spatial"], cfg["spatial"]), generator=g) * 2 - 1
    with torch.no_grad():
        frames = O.decode_first_stage(sd, cfg, motion, x0, fx["T"])
    assert frames.shape == fx["frames"].shape
    assert (frames - fx["frames"]).abs().max().item() < 5e-5


@pytest.mark.parametrize("name", ["enc_64", "enc_128"])
def test_encoder_oracle_matches_reference(name):
    """3-D conv video encoder (training path): oracle restatement vs the reference's resnet18_alternative output."""
    fx = golden(name)
    cfg = O.encoder_config(**fx["cfg_kwargs"])
    sd = O.synth_encoder_state_dict(cfg, seed=fx["wseed"])
    g = torch.Generator().manual_seed(fx["iseed"])
    X = torch.rand((fx["B"], 3, fx["T"], cfg["img_size"], cfg["img_size"]), generator=g) * 2 - 1
    with torch.no_grad():
        z, mu, lv = O.encoder_forward(sd, cfg, X, fx["eps"])
    assert z.shape == fx["z"].shape == (fx["B"], cfg["z_dim"], 8, 8)
    for a, b in ((z, fx["z"]), (mu, fx["mu"]), (lv, fx["logvar"])):
        assert (a - b).abs().max().item() < 5e-5


@pytest.mark.parametrize("name", ["cenc_poke_128", "cenc_img_64"])
def test_cond_encoder_oracle_matches_reference(name):
    fx = golden(name)
    cfg = O.cond_encoder_config(**fx["cfg_kwargs"])
    sd = O.synth_cond_encoder_state_dict(cfg, seed=fx["wseed"])
    g = torch.Generator().manual_seed(fx["iseed"])
    x = torch.rand((fx["B"], cfg["nf_in"], cfg["spatial"], cfg["spatial"]), generator=g) * 2 - 1
    with torch.no_grad():
        out, mean = O.cond_encoder_forward(sd, cfg, x)
    assert (out - fx["out"]).abs().max().item() < 5e-5 and (mean - fx["mean"]).abs().max().item() < 5e-5


def test_shuffle_roundtrip_exact():
    cfg = O.flow_config(flow_in_channels=16, flow_mid_channels=16, h_channels=4, num_steps=[1], factor=2)
    sd = O.synth_flow_state_dict(cfg, seed=9)
    x = torch.randn(2, 16, 8, 8)
    p = "flow.layers.0.0.conv1x1."
    assert torch.equal(O.shuffle_bwd(sd, p, O.shuffle_fwd(sd, p, x)), x)


def test_identity_init_logdet_is_actnorm_sum():
    """Known-answer property ii: with weight_g = 0 and zero biases every coupling is the identity and
    logdet = H*W*sum(ActNorm log_scale), identical for all samples."""
    cfg = O.flow_config(flow_in_channels=16, flow_mid_channels=16, h_channels=4, num_steps=[1, 1], factor=4)
    sd = O.synth_flow_state_dict(cfg, seed=3)
    tot = 0.0
    for k in sd:
        if k.endswith("weight_g") or k.endswith("conv.bias"):
            sd[k] = torch.zeros_like(sd[k])
        if k.endswith("log_scale"):
            tot += 64.0 * sd[k].sum().item()
    z, cond, _ = O.synth_inputs(3, 16, 4, 8, seed=1)
    _, ld = O.flow_forward(sd, cfg, z, cond)
    assert torch.allclose(ld, torch.full((3,), tot), atol=1e-4)


@pytest.mark.slow
def test_full_size_flow_oracle_matches_reference():
    fx = golden("flow_full_c32")
    cfg = O.flow_config(**fx["cfg_kwargs"])
    sd = O.synth_flow_state_dict(cfg, seed=fx["wseed"])
    z, cond, _ = O.synth_inputs(fx["B"], 32, 128, 8, seed=fx["iseed"])
    with torch.no_grad():
        x = O.flow_reverse(sd, cfg, z, cond)
    assert (x - fx["x_rev"]).abs().max().item() < 5e-4


@pytest.mark.parametrize("name", ["i3d_t10", "i3d_t16"])
def test_fvd_chain_oracle_matches_reference(name):
    """FVD chain (I3D logits, preprocess, Frechet distance) restated in oracle/fvd_oracle.py vs the reference's
    utils/metrics.py run on the same seeded weights and clips (tests/golden/make_golden.py: run_i3d_case)."""
    import numpy as np
    from oracle import fvd_oracle as FO
    fx = golden(name)
    sd = FO.synth_i3d_state_dict(fx["wseed"])
    g = torch.Generator().manual_seed(fx["iseed"])
    vid_a = torch.rand((fx["B"], fx["T"], 3, fx["S"], fx["S"]), generator=g) * 2 - 1
    vid_b = torch.rand((fx["B"], fx["T"], 3, fx["S"], fx["S"]), generator=g)
    pa, pb = FO.preprocess(vid_a), FO.preprocess(vid_b)
    assert abs(pa.double().sum().item() - fx["pre_a_sum"]) < 1e-6 * abs(fx["pre_a_sum"])
    assert abs(pb.double().sum().item() - fx["pre_b_sum"]) < 1e-6 * abs(fx["pre_b_sum"])       # [0,1] input is not denormed
    assert pa.min() >= 0 and pa.max() <= 1
    act = FO.activations(sd, pa, batch_size=fx["B"])
    assert np.abs(act - fx["act_a"].numpy()).max() < 1e-3 * fx["logit_std"]
    with torch.no_grad():
        lb = FO.i3d_logits(sd, pb.permute(0, 2, 1, 3, 4))
    assert (lb - fx["logits_b"]).abs().max().item() < 1e-3 * fx["logit_std"]
    assert abs(FO.fvd_from_activations(fx["f1"], fx["f2"]) - fx["fd"]) < 1e-6 * fx["fd"]


def test_frechet_distance_known_answers():
    """identical Gaussians -> 0, unit mean shift in 3-D with equal covariance -> 3.0."""
    import numpy as np
    from oracle import fvd_oracle as FO
    s = np.diag([1.0, 2.0, 0.5])
    assert abs(FO.frechet_distance(np.zeros(3), s, np.zeros(3), s)) < 1e-9
    assert abs(FO.frechet_distance(np.zeros(3), s, np.ones(3), s) - 3.0) < 1e-9


def test_synth_pokes_layout():
    from oracle import fvd_oracle as FO
    poke, centres = FO.synth_pokes(4, 64, seed=5)
    assert poke.shape == (4, 2, 64, 64) and centres.shape == (4, 1, 2)
    for b in range(4):
        nz = (poke[b].abs().sum(0) > 0)
        assert nz.sum().item() == 25
        ys, xs = torch.nonzero(nz, as_tuple=True)
        assert ys.float().mean().item() == centres[b, 0, 0].item() and xs.float().mean().item() == centres[b, 0, 1].item()
        assert 5 <= centres[b, 0, 0] < 59 and 5 <= centres[b, 0, 1] < 59


@pytest.mark.parametrize("name", ["flowgrad_tiny", "flowgrad_c32_hd128"])
def test_training_step_oracle_matches_reference_autograd(name):
    """Second-stage training step (forward_density + FlowLoss + backward): the oracle's autograd over the restated forward vs
    loss.backward() of the reference flow + FlowLoss (tests/golden/make_golden.py: run_grad_case)."""
    fx = golden(name)
    cfg = O.flow_config(**fx["cfg_kwargs"])
    sd = O.synth_flow_state_dict(cfg, seed=fx["wseed"])
    x, cond, _ = O.synth_inputs(fx["B"], cfg["flow_in_channels"], cfg["h_channels"], 8, seed=fx["iseed"])
    loss, grads = O.flow_loss_and_grads(sd, cfg, x * 0.8, cond)
    assert abs(loss.item() - fx["loss"]) < 1e-5 * abs(fx["loss"])
    assert list(grads) == fx["keys"]
    for i, k in enumerate(fx["keys"]):
        g = grads[k]
        scale = fx["maxabs"][i].item() + 1e-12
        assert abs(g.double().norm().item() - fx["norms"][i].item()) < 1e-4 * fx["norms"][i].item() + 1e-9, k
        assert (g.flatten()[fx["idx"][i]] - fx["val"][i]).abs().max().item() < 1e-4 * scale, k


def _fresh_state(fx):
    """Pre-init state of the fixture's flow: the reference's ActNorm parameters / permutations, the oracle's synthetic conv weights (the
    init pass does not depend on them: every weight-normed conv is zero_init), every `initialized` flag 0."""
    cfg = O.flow_config(**fx["cfg_kwargs"])
    sd = O.synth_flow_state_dict(cfg, seed=0)
    for k, v in fx["sd0"].items():
        assert sd[k].shape == v.shape, k
        sd[k] = v.clone()
    for k in list(sd):
        if k.endswith("initialized"):
            sd[k] = torch.tensor(0, dtype=torch.uint8)
    return cfg, sd


@pytest.mark.parametrize("name", ["flowinit_tiny", "flowinit_c32"])
def test_data_dependent_init_oracle_matches_reference(name):
    """ActNorm2dFlow.init / Conv2dWeightNorm.init as fired by the reference's first train() forward (macow2.py:503-505,526-539,
    macow_utils.py:231-250): every tensor the reference changed, and the outputs of that forward."""
    fx = golden(name)
    cfg, sd = _fresh_state(fx)
    with torch.no_grad():
        sd1, z, ld = O.flow_data_init(sd, cfg, fx["x"], fx["cond"])
    assert (z - fx["z"]).abs().max().item() < 1e-5 and (ld - fx["logdet"]).abs().max().item() < 1e-3
    for k, v in fx["changed"].items():
        assert (sd1[k].float() - v.float()).abs().max().item() < 1e-5, k
    # nothing else moved, and every flag is set
    for k, v in sd1.items():
        if k.endswith("initialized"):
            assert int(v) == 1
        elif k.endswith(".conv.bias"):
            assert float(v.abs().max()) == 0.0, k      # the reference's fresh bias is already 0, so it is not among `changed`
        elif k not in fx["changed"]:
            assert torch.equal(v, sd[k]), k


def test_flow_loss_log_dict_and_rng_consumption():
    """FlowLoss.forward (loss.py:13-31): all entries of the log dict, and the randn_like draw it takes from the default generator."""
    fx = golden("flowloss")
    import ipoke_b200 as ipk
    for sm in (False, True):
        want = fx["out"][sm]
        for impl in ("oracle", "product"):
            torch.manual_seed(123)
            if impl == "oracle":
                loss, log = O.flow_loss_log(fx["z"], fx["logdet"], spatial_mean=sm)
            else:
                loss, log = ipk.FlowLoss(spatial_mean=sm, logdet_weight=1.0)(fx["z"], fx["logdet"])     # host arithmetic on tiny tensors
            nxt = torch.randn(3)
            assert abs(loss.item() - want["loss"]) < 1e-4 * abs(want["loss"])
            assert set(log) == set(want["log"])
            for k, v in want["log"].items():
                got = log[k].item() if torch.is_tensor(log[k]) else log[k]
                assert abs(got - v) <= 1e-5 * max(1.0, abs(v)), (impl, k)
            assert torch.equal(nxt, want["next_randn"]), impl
