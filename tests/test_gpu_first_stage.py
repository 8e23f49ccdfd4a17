"""ConvGRU + SPADE decoder parity on the GPU against the reference's golden frames and the CPU oracle."""
import pytest
import torch

from conftest import golden
from util import O, make_first_stage, make_flow, maxabs

pytestmark = pytest.mark.gpu

TOL = {"fp32_simt": 1e-4, "fp32": 1e-3, "bf16": 1.5e-1}


def _inputs(fx, cfg):
    g = torch.Generator().manual_seed(fx["iseed"])
    motion = torch.randn((fx["B"], cfg["z_dim"], 8, 8), generator=g) * 1.3
    x0 = torch.rand((fx["B"], 3, cfg["spatial"], cfg["spatial"]), generator=g) * 2 - 1
    return motion, x0


@pytest.mark.parametrize("name", ["fs_64", "fs_128", "fs_64_z64"])
@pytest.mark.parametrize("prec", ["fp32_simt", "fp32", "bf16"])
def test_decode_matches_reference_golden(name, prec):
    fx = golden(name)
    cfg = O.first_stage_config(**fx["cfg_kwargs"])
    sd = O.synth_first_stage_state_dict(cfg, seed=fx["wseed"])
    motion, x0 = _inputs(fx, cfg)
    m = make_first_stage(cfg, sd, prec, max_batch=fx["B"], max_frames=fx["T"])
    frames = m.decode(motion.cuda(), x0.cuda(), fx["T"])
    err = maxabs(frames, fx["frames"])
    print(f"{name} {prec}: frames max-abs {err:.2e}")
    assert frames.shape == fx["frames"].shape
    assert err < TOL[prec]


def test_rnn_and_gen_interfaces_and_chunking():
    """first_stage_model.rnn(x, hidden) / .gen([h], x0) keep the reference call signatures; chunked decoding (one video per
    pass) gives the same frames as decoding all videos at once."""
    fx = golden("fs_64")
    cfg = O.first_stage_config(**fx["cfg_kwargs"])
    sd = O.synth_first_stage_state_dict(cfg, seed=fx["wseed"])
    motion, x0 = _inputs(fx, cfg)
    m = make_first_stage(cfg, sd, "fp32_simt", max_batch=fx["B"], max_frames=fx["T"])
    B = fx["B"]
    hidden = [motion.cuda()] * m.n_layers
    in_rnn = torch.cat([m.motion_bias] * B, dim=0)
    frames = []
    for _ in range(fx["T"]):
        hidden = m.rnn(in_rnn, hidden)
        lst = [hidden[-1]]
        frames.append(m.gen(lst, x0.cuda(), del_shape=True))
        assert not lst
    frames = torch.stack(frames, dim=1)
    assert maxabs(frames, fx["frames"]) < 1e-4
    assert maxabs(hidden[-1], fx["hidden_last"]) < 1e-4
    m1 = make_first_stage(cfg, sd, "fp32_simt", max_batch=B, max_frames=fx["T"], chunk_videos=1)
    assert maxabs(m1.decode(motion.cuda(), x0.cuda(), fx["T"]), frames) < 1e-5


def test_sample_end_to_end_device_and_host_paths():
    """flow inverse -> decode through ipk_sample (device buffers) and ipk_sample_host (host buffers) vs the oracle."""
    import ipoke_b200 as ipk
    fcfg = O.flow_config(flow_in_channels=32, flow_mid_channels=128, h_channels=128)
    fsd = O.synth_flow_state_dict(fcfg, seed=3)
    dcfg = O.first_stage_config(z_dim=32, spatial=64)
    dsd = O.synth_first_stage_state_dict(dcfg, seed=21)
    z, cond, x0 = O.synth_inputs(2, 32, 128, 64, seed=42)
    with torch.no_grad():
        ref = O.sample_videos(fsd, fcfg, dsd, dcfg, z, cond, x0, 3)
    for prec, tol in (("fp32_simt", 2e-4), ("fp32", 1e-3)):
        s = ipk.PokeMotionSampler(make_flow(fcfg, fsd, prec), make_first_stage(dcfg, dsd, prec, max_batch=2, max_frames=3))
        dev = s.sample(z.cuda(), cond.cuda(), x0.cuda(), 3)
        host = s.sample_host(z, cond, x0, 3)
        e = maxabs(dev, ref)
        print(f"sample {prec}: frames max-abs vs oracle {e:.2e}")
        assert e < tol
        assert maxabs(host, dev) == 0.0
        X = torch.cat([x0.unsqueeze(1)] * 4, dim=1).cuda()
        vids = s.forward_sample(X, cond.cuda(), n_samples=2, n_logged_vids=1)
        assert len(vids) == 2 and vids[0].shape == (1, 3, 3, 64, 64) and vids[0].device.type == "cpu"
