"""Native I3D / FVD chain against the reference's utils/metrics.py: the fixtures tests/golden/i3d_*.pt hold the
UNMODIFIED reference I3D's logits on seeded clips and weights, its preprocess() sums and a calculate_frechet_distance value
(tests/golden/make_golden.py: run_i3d_case); the oracle restatement (oracle/fvd_oracle.py) is checked beside it."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import fvd_oracle as FO

pytestmark = pytest.mark.gpu


def _clips(fx):
    g = torch.Generator().manual_seed(fx["iseed"])
    B, T, S = fx["B"], fx["T"], fx["S"]
    vid_a = torch.rand((B, T, 3, S, S), generator=g) * 2 - 1
    vid_b = torch.rand((B, T, 3, S, S), generator=g)                     # already in [0,1]: preprocess must not denorm it
    return vid_a, vid_b


# stated tolerance: logits of scale `logit_std` (~1) after 22 conv layers in bf16x3: 2e-3 max-abs (measured ~1e-4); bf16: 2.5e-1
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-3), ("bf16", 2.5e-1)])
@pytest.mark.parametrize("name", ["i3d_t10", "i3d_t16"])
def test_i3d_logits_match_reference(name, precision, tol):
    import ipoke_b200 as ipk
    fx = golden(name)
    vid_a, vid_b = _clips(fx)
    m = ipk.I3D(400, "rgb", ipk_max_batch=fx["B"], ipk_max_frames=fx["T"], ipk_precision=precision)
    m.load_state_dict(FO.synth_i3d_state_dict(fx["wseed"]), strict=True)
    m = m.cuda().eval()
    pa, pb = ipk.i3d.preprocess(vid_a.cuda(), vid_b.cuda())
    # preprocess: same tensors as the reference's (sums recorded in the fixture), set b is left in [0, 1], set a mapped from [-1, 1]
    assert abs(pa.double().sum().item() - fx["pre_a_sum"]) < 1e-6 * abs(fx["pre_a_sum"]) + 1.0
    assert abs(pb.double().sum().item() - fx["pre_b_sum"]) < 1e-6 * abs(fx["pre_b_sum"]) + 1.0
    assert (pa.cpu() - FO.preprocess(vid_a)).abs().max().item() < 2e-6 and (pb.cpu() - FO.preprocess(vid_b)).abs().max().item() < 2e-6
    assert float(pa.min()) >= 0.0 and float(pa.max()) <= 1.0
    act = ipk.i3d.get_activations(pa, m, batch_size=fx["B"])
    assert act.shape == (fx["B"], 400) and act.dtype == np.float64
    ea = np.abs(act - fx["act_a"].numpy()).max()
    soft, logits = m(pb.permute(0, 2, 1, 3, 4))
    eb = (logits.cpu() - fx["logits_b"]).abs().max().item()
    print(f"{name}[{precision}]: logits max-abs {ea:.2e} / {eb:.2e} (logit std {fx['logit_std']:.3f})")
    assert ea < tol * max(1.0, fx["logit_std"]) and eb < tol * max(1.0, fx["logit_std"])
    assert torch.allclose(soft.sum(dim=1), torch.ones(fx["B"], device=soft.device), atol=1e-5)
    # Frechet distance (host linear algebra, as in the reference) on the fixture's seeded Gaussians
    f1, f2 = fx["f1"], fx["f2"]
    fd = ipk.i3d.calculate_frechet_distance(f1.mean(0), np.cov(f1, rowvar=False), f2.mean(0), np.cov(f2, rowvar=False))
    assert abs(fd - fx["fd"]) < 1e-6 * max(1.0, abs(fx["fd"]))


def test_i3d_odd_and_even_frame_counts_and_batches():
    """TF-SAME depth padding depends on T mod stride (utils/metrics.py:927-931, :955-958): T = 9..16 against the oracle restatement;
    batch sizes that do not fill the 128-voxel boxes of the late layers."""
    import ipoke_b200 as ipk
    sd = FO.synth_i3d_state_dict(7)
    m = ipk.I3D(400, "rgb", ipk_max_batch=3, ipk_max_frames=16)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    sd_dev = {k: v.cuda() for k, v in sd.items()}
    g = torch.Generator().manual_seed(3)
    for T, B in ((9, 1), (11, 3), (12, 2), (13, 1), (16, 3)):
        x = torch.rand((B, 3, T, 224, 224), generator=g).cuda()
        with torch.no_grad():
            want = FO.i3d_logits(sd_dev, x)
        got = m(x)[1]
        err = (got - want).abs().max().item()
        print(f"T={T} B={B}: max-abs {err:.2e} (logit std {want.std().item():.3f})")
        assert err < 2e-3 * max(1.0, want.std().item())


def test_calculate_fvd_end_to_end():
    """calculate_FVD (utils/metrics.py:773-780) on two small video sets against the oracle chain with the same seeded I3D."""
    import ipoke_b200 as ipk
    sd = FO.synth_i3d_state_dict(5)
    m = ipk.I3D(400, "rgb", ipk_max_batch=8, ipk_max_frames=10)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(11)
    gen = (torch.rand((16, 10, 3, 64, 64), generator=g) * 2 - 1).cuda()
    orig = (torch.rand((16, 10, 3, 64, 64), generator=g) * 1.6 - 0.8).cuda()
    # features only (the 400-d covariance of 16 samples is singular: the Frechet matrix square root is not meaningful at this size)
    pg, po = ipk.i3d.preprocess(gen, orig)
    a_g, a_o = ipk.i3d.get_activations(pg, m, batch_size=8), ipk.i3d.get_activations(po, m, batch_size=8)
    sd_dev = {k: v.cuda() for k, v in sd.items()}
    w_g = FO.activations(sd_dev, FO.preprocess(gen), batch_size=8)
    w_o = FO.activations(sd_dev, FO.preprocess(orig), batch_size=8)
    assert np.abs(a_g - w_g).max() < 2e-3 and np.abs(a_o - w_o).max() < 2e-3
    # and the distance itself on a well-conditioned pair of Gaussians
    rng = np.random.RandomState(0)
    f1, f2 = rng.randn(500, 8), rng.randn(600, 8) * 1.3 + 0.2
    d = ipk.i3d.calculate_frechet_distance(f1.mean(0), np.cov(f1, rowvar=False), f2.mean(0), np.cov(f2, rowvar=False))
    assert abs(d - FO.fvd_from_activations(f1, f2)) < 1e-8 * max(1.0, d)
