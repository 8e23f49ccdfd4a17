"""3-D conv video encoder (second-stage training path) on the GPU against the reference's golden outputs and the oracle."""
import pytest
import torch

from conftest import golden
from util import O, maxabs

pytestmark = pytest.mark.gpu


def _make(cfg, sd, max_batch=2):
    import ipoke_b200 as ipk
    m = ipk.ResNetMotionEncoder(dict(cfg, ipk_max_batch=max_batch))
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("name", ["enc_64", "enc_128"])
def test_encoder_matches_reference_golden(name):
    fx = golden(name)
    cfg = O.encoder_config(**fx["cfg_kwargs"])
    sd = O.synth_encoder_state_dict(cfg, seed=fx["wseed"])
    g = torch.Generator().manual_seed(fx["iseed"])
    X = torch.rand((fx["B"], 3, fx["T"], cfg["img_size"], cfg["img_size"]), generator=g) * 2 - 1
    m = _make(cfg, sd, max_batch=fx["B"])
    z, mu, lv = m(X.cuda(), eps=fx["eps"])
    e = [maxabs(z, fx["z"]), maxabs(mu, fx["mu"]), maxabs(lv, fx["logvar"])]
    print(f"{name}: z/mu/logvar max-abs {e[0]:.2e} {e[1]:.2e} {e[2]:.2e}")
    assert z.shape == (fx["B"], cfg["z_dim"], 8, 8)
    assert max(e) < 2e-4          # fp32 FFMA, 17 conv + GroupNorm layers; reference fp32 vs fp64 is 2.6e-6


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
