"""3-D conv video encoder (second-stage training path) on the GPU against the reference's golden outputs and the oracle."""
import pytest
import torch

from conftest import golden
from util import O, maxabs

pytestmark = pytest.mark.gpu


def _make(cfg, sd, max_batch=2, precision="fp32"):
    import ipoke_b200 as ipk
    m = ipk.ResNetMotionEncoder(dict(cfg, ipk_max_batch=max_batch, ipk_precision=precision))
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


# stated tolerances: fp32 (bf16x3 tcgen05 Conv3d) and fp32_simt (FFMA) 2e-4 -- 17 conv + GroupNorm layers, reference fp32 vs fp64 is
# 2.6e-6; bf16 (single-pass tcgen05) 1e-1 on latents of unit scale
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("fp32_simt", 2e-4), ("bf16", 1e-1)])
@pytest.mark.parametrize("name", ["enc_64", "enc_128"])
def test_encoder_matches_reference_golden(name, precision, tol):
    fx = golden(name)
    cfg = O.encoder_config(**fx["cfg_kwargs"])
    sd = O.synth_encoder_state_dict(cfg, seed=fx["wseed"])
    g = torch.Generator().manual_seed(fx["iseed"])
    X = torch.rand((fx["B"], 3, fx["T"], cfg["img_size"], cfg["img_size"]), generator=g) * 2 - 1
    m = _make(cfg, sd, max_batch=fx["B"], precision=precision)
    z, mu, lv = m(X.cuda(), eps=fx["eps"])
    e = [maxabs(z, fx["z"]), maxabs(mu, fx["mu"]), maxabs(lv, fx["logvar"])]
    print(f"{name}[{precision}]: z/mu/logvar max-abs {e[0]:.2e} {e[1]:.2e} {e[2]:.2e}")
    assert z.shape == (fx["B"], cfg["z_dim"], 8, 8)
    assert max(e) < tol


def test_encoder_default_eps_matches_reference_rng_and_is_batch_invariant():
    """eps defaults to the CPU-generator draw the reference makes (motion_encoder.py:220); samples do not interact."""
    cfg = O.encoder_config(z_dim=32, img_size=64, max_frames=10)
    sd = O.synth_encoder_state_dict(cfg, seed=7)
    g = torch.Generator().manual_seed(3)
    X = torch.rand((3, 3, 11, 64, 64), generator=g) * 2 - 1
    m = _make(cfg, sd, max_batch=3)
    torch.manual_seed(9)
    z, mu, lv = m(X.cuda())
    torch.manual_seed(9)
    eps = torch.FloatTensor(torch.Size((3, 32, 8, 8))).normal_()
    with torch.no_grad():
        z_ref, mu_ref, lv_ref = O.encoder_forward(sd, cfg, X, eps)
    assert maxabs(z, z_ref) < 2e-4 and maxabs(mu, mu_ref) < 2e-4 and maxabs(lv, lv_ref) < 2e-4
    _, mu1, _ = m(X[1:2].cuda(), eps=eps[1:2])
    assert maxabs(mu1, mu[1:2]) < 1e-6
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 11, 32, 32, device="cuda"))


@pytest.mark.parametrize("name", ["cenc_poke_128", "cenc_img_64"])
def test_cond_encoder_matches_reference_golden(name):
    """Frozen ConvEncoder of the poke embedder / image conditioner (make_flow_input) vs the reference's output."""
    import ipoke_b200 as ipk
    fx = golden(name)
    cfg = O.cond_encoder_config(**fx["cfg_kwargs"])
    sd = O.synth_cond_encoder_state_dict(cfg, seed=fx["wseed"])
    g = torch.Generator().manual_seed(fx["iseed"])
    x = torch.rand((fx["B"], cfg["nf_in"], cfg["spatial"], cfg["spatial"]), generator=g) * 2 - 1
    m = ipk.ConvEncoder(cfg["nf_in"], cfg["nf_max"], cfg["n_stages"], ipk_max_batch=fx["B"])
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    out, mean, logstd = m(x.cuda())
    e = (maxabs(out, fx["out"]), maxabs(mean, fx["mean"]))
    print(f"{name}: out/mean max-abs {e[0]:.2e} {e[1]:.2e}")
    assert logstd is None and max(e) < 1e-4


def test_make_cond_feeds_the_flow():
    """make_flow_input's conditioning: cat[conditioner(x0), poke_embedder(poke)] has the flow's h_channels = 128."""
    import ipoke_b200 as ipk
    ic, pc = O.cond_encoder_config(nf_in=3, spatial=64), O.cond_encoder_config(nf_in=2, spatial=64)
    img = ipk.ConvEncoder(3, 64, ic["n_stages"]); img.load_state_dict(O.synth_cond_encoder_state_dict(ic, seed=1))
    pk = ipk.ConvEncoder(2, 64, pc["n_stages"]); pk.load_state_dict(O.synth_cond_encoder_state_dict(pc, seed=2))
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand((2, 3, 64, 64), generator=g) * 2 - 1
    poke = torch.zeros(2, 2, 64, 64); poke[:, :, 30:35, 20:25] = 0.7
    cond = ipk.make_cond(img.cuda().eval(), pk.cuda().eval(), x0.cuda(), poke.cuda())
    with torch.no_grad():
        ref = torch.cat([O.cond_encoder_forward(O.synth_cond_encoder_state_dict(ic, seed=1), ic, x0)[0],
                         O.cond_encoder_forward(O.synth_cond_encoder_state_dict(pc, seed=2), pc, poke)[0]], dim=1)
    assert cond.shape == (2, 128, 8, 8) and maxabs(cond, ref) < 1e-4
