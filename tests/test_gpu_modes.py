"""SURVEY.md section 8f rows 2 and 4 on the GPU: uint8 post-processing (bit-exact), transfer and control-sensitivity modes
(compositions of the flow forward / inverse, the encoders and the decoder) against the same compositions of the oracle."""
import numpy as np
import pytest
import torch

from oracle import fvd_oracle as FO
from oracle import ipoke_oracle as O
from util import maxabs

pytestmark = pytest.mark.gpu


def _ref_u8(frames):
    """second_stage_video.py:673-675 on the host, verbatim."""
    return ((frames + 1.) * 127.5).permute(0, 1, 3, 4, 2).numpy().astype(np.uint8)


def test_frames_to_uint8_bit_exact():
    import ipoke_b200 as ipk
    g = torch.Generator().manual_seed(3)
    x = torch.tanh(torch.randn((3, 5, 3, 64, 64), generator=g) * 2)
    # every representable output level and its neighbourhood: k / 127.5 - 1 and the floats just below / above
    k = torch.arange(0, 256, dtype=torch.float32) / 127.5 - 1.0
    edge = torch.cat([k, torch.nextafter(k, torch.full_like(k, -2.0)), torch.nextafter(k, torch.full_like(k, 2.0))]).clamp(-1, 1)
    x.view(-1)[:edge.numel()] = edge
    x.view(-1)[-2:] = torch.tensor([-1.0, 1.0])
    out = ipk.PokeMotionSampler.to_uint8(x.cuda()).cpu().numpy()
    ref = _ref_u8(x)
    assert out.shape == ref.shape == (3, 5, 64, 64, 3) and out.dtype == np.uint8
    assert np.array_equal(out, ref)


def _build(c0=32, spatial=64, B=4, T=10, precision="fp32"):
    import ipoke_b200 as ipk
    fcfg = O.flow_config(flow_in_channels=c0, flow_mid_channels=128, h_channels=128, num_steps=[2, 1, 1] + [1] * 12)
    dcfg = O.first_stage_config(z_dim=c0, spatial=spatial)
    ecfg = O.encoder_config(z_dim=c0, img_size=spatial, max_frames=T)
    icfg, pcfg = O.cond_encoder_config(nf_in=3, spatial=spatial), O.cond_encoder_config(nf_in=2, spatial=spatial)
    sds = dict(flow=O.synth_flow_state_dict(fcfg, seed=5), fs=O.synth_first_stage_state_dict(dcfg, seed=6),
               enc=O.synth_encoder_state_dict(ecfg, seed=7), img=O.synth_cond_encoder_state_dict(icfg, seed=8),
               poke=O.synth_cond_encoder_state_dict(pcfg, seed=9))
    fc = dict(fcfg); fc.update(ipk_precision=precision, ipk_max_batch=3 * B)
    dc = dict(dcfg); dc.update(ipk_precision=precision, ipk_max_batch=3 * B, ipk_max_frames=T)
    flow = ipk.SupervisedMacowTransformer(fc); flow.load_state_dict(sds["flow"], strict=True)
    fs = ipk.SpadeCondMotionDecoder(dc); fs.load_state_dict(sds["fs"], strict=True)
    img = ipk.ConvEncoder(3, 64, icfg["n_stages"], ipk_max_batch=3 * B); img.load_state_dict(sds["img"], strict=True)
    pk = ipk.ConvEncoder(2, 64, pcfg["n_stages"], ipk_max_batch=3 * B); pk.load_state_dict(sds["poke"], strict=True)
    enc = ipk.ResNetMotionEncoder(dict(ecfg, ipk_max_batch=B)); enc.load_state_dict(sds["enc"], strict=True)
    s = ipk.PokeMotionSampler(flow.cuda().eval(), fs.cuda().eval(), img.cuda().eval(), pk.cuda().eval())
    return s, enc.cuda().eval(), sds, dict(flow=fcfg, fs=dcfg, enc=ecfg, img=icfg, poke=pcfg)


def _cond(sds, cfgs, x0, poke):
    return torch.cat([O.cond_encoder_forward(sds["img"], cfgs["img"], x0)[0], O.cond_encoder_forward(sds["poke"], cfgs["poke"], poke)[0]], dim=1)


def test_transfer_matches_oracle_composition():
    B, T, S, c0 = 2, 10, 64, 32
    s, enc, sds, cfgs = _build(c0, S, B, T)
    g = torch.Generator().manual_seed(17)
    X_1 = torch.rand((B, T + 1, 3, S, S), generator=g) * 2 - 1
    X_2 = torch.rand((B, T + 1, 3, S, S), generator=g) * 2 - 1
    poke_1, _ = FO.synth_pokes(B, S, seed=4)
    eps = torch.randn((B, c0, 8, 8), generator=g)
    res = torch.randn((B, c0, 8, 8), generator=g)
    v1, v2, r1 = s.transfer(X_1.cuda(), X_2.cuda(), poke_1.cuda(), enc, eps=eps.cuda(), residual=res.cuda())
    with torch.no_grad():
        z1, _, _ = O.encoder_forward(sds["enc"], cfgs["enc"], X_1.transpose(1, 2), eps)
        c1, c2 = _cond(sds, cfgs, X_1[:, 0], poke_1), _cond(sds, cfgs, X_2[:, 0], poke_1)
        r1_ref, _ = O.flow_forward(sds["flow"], cfgs["flow"], z1, c1)
        v1_ref = O.decode_first_stage(sds["fs"], cfgs["fs"], O.flow_reverse(sds["flow"], cfgs["flow"], r1_ref, c2), X_2[:, 0], T)
        v2_ref = O.decode_first_stage(sds["fs"], cfgs["fs"], O.flow_reverse(sds["flow"], cfgs["flow"], res, c2), X_2[:, 0], T)
    e = (maxabs(r1, r1_ref), maxabs(v1, v1_ref), maxabs(v2, v2_ref))
    print("transfer: residual / transferred / random max-abs", e)
    assert v1.shape == (B, T, 3, S, S) and e[0] < 5e-4 and e[1] < 1e-3 and e[2] < 1e-3


def test_control_sensitivity_matches_per_poke_forward_sample():
    """The batched sweep equals the reference's loop: one forward_sample per poke, noise drawn in the same order."""
    B, T, S, c0, P = 2, 4, 64, 32, 3
    s, _, sds, cfgs = _build(c0, S, B, T)
    g = torch.Generator().manual_seed(23)
    X = torch.rand((B, T + 1, 3, S, S), generator=g) * 2 - 1
    pokes = torch.stack([FO.synth_pokes(B, S, seed=30 + p)[0] for p in range(P)])
    torch.manual_seed(77)
    out = s.control_sensitivity(X.cuda(), pokes.cuda(), max_batch=4)
    assert out.shape == (B, P, T, 3, S, S) and not out.is_cuda
    torch.manual_seed(77)
    with torch.no_grad():
        for p in range(P):
            z = torch.randn((B, c0, 8, 8))                         # make_flow_input(reverse=True) :300, one draw per forward_sample
            ref = O.sample_videos(sds["flow"], cfgs["flow"], sds["fs"], cfgs["fs"], z, _cond(sds, cfgs, X[:, 0], pokes[p]), X[:, 0], T)
            assert maxabs(out[:, p], ref) < 1e-3


def test_sample_host_uint8_matches_float_path():
    B, T, S, c0 = 3, 4, 64, 32
    s, _, _, _ = _build(c0, S, B, T)
    z, cond, x0 = O.synth_inputs(B, c0, 128, S, seed=5)
    f32 = s.sample_host(z, cond, x0, T).clone()
    u8 = s.sample_host(z, cond, x0, T, uint8=True)
    assert u8.dtype == torch.uint8 and tuple(u8.shape) == (B, T, S, S, 3)
    assert np.array_equal(u8.numpy(), _ref_u8(f32))


def test_training_step_forward_half_loss_parity():
    """BASELINE configs[3] (h36m shapes: C0 = 64), forward half of the second-stage training step
    (second_stage_video.py:345-350, loss.py:13-31): X -> enc_motion (no grad) -> flow forward + log-det -> FlowLoss.
    Stated tolerance: loss relative error 1e-5 in fp32 mode; measured ~1e-6."""
    import ipoke_b200 as ipk
    B, T, S, c0 = 2, 10, 64, 64
    s, enc, sds, cfgs = _build(c0, S, B, T)
    g = torch.Generator().manual_seed(41)
    X = torch.rand((B, T + 1, 3, S, S), generator=g) * 2 - 1
    poke, _ = FO.synth_pokes(B, S, seed=8)
    eps = torch.randn((B, c0, 8, 8), generator=g)
    z_in, _ = ipk.encode_first_stage(enc, X.cuda(), eps=eps.cuda())
    cond = s.make_cond(X[:, 0].cuda(), poke.cuda())
    out, logdet = s.forward_density(z_in, cond)
    loss = ipk.flow_nll(out, logdet).item()
    with torch.no_grad():
        z_ref, _, _ = O.encoder_forward(sds["enc"], cfgs["enc"], X.transpose(1, 2), eps)
        out_ref, ld_ref = O.flow_forward(sds["flow"], cfgs["flow"], z_ref, _cond(sds, cfgs, X[:, 0], poke))
        loss_ref = O.flow_nll(out_ref, ld_ref).item()
    rel = abs(loss - loss_ref) / abs(loss_ref)
    print(f"training forward half: loss {loss:.6f} vs {loss_ref:.6f} (rel {rel:.2e}), logdet max-abs {maxabs(logdet, ld_ref):.2e}")
    assert rel < 1e-5


def test_poke_and_image_embedder_five_channels():
    """embed_poke_and_image (second_stage_video.py:265-266): the poke embedder sees cat[poke, X[:,0]] = 5 channels."""
    import ipoke_b200 as ipk
    cfg = O.cond_encoder_config(nf_in=5, spatial=64)
    sd = O.synth_cond_encoder_state_dict(cfg, seed=12)
    g = torch.Generator().manual_seed(2)
    x0 = torch.rand((3, 3, 64, 64), generator=g) * 2 - 1
    poke, _ = FO.synth_pokes(3, 64, seed=6)
    inp = torch.cat([poke, x0], dim=1)
    m = ipk.ConvEncoder(5, 64, cfg["n_stages"], ipk_max_batch=3)
    m.load_state_dict(sd, strict=True)
    out, mean, _ = m.cuda().eval()(inp.cuda())
    with torch.no_grad():
        ref_out, ref_mean = O.cond_encoder_forward(sd, cfg, inp)
    assert maxabs(out, ref_out) < 1e-4 and maxabs(mean, ref_mean) < 1e-4


def test_config1_plants64_full_size_sampling_pipeline():
    """BASELINE configs[0]: plants_64 `--test samples` plumbing at full model size -- 10-frame 64x64, batch 2, C0 = 32, Hd = 2048
    (1.05 B parameters), seed 42: poke + start frame -> conditioning encoders -> z ~ N(0, I) from the CPU generator -> flow inverse
    -> ConvGRU + SPADE decoder, through PokeMotionSampler.forward_sample, against the CPU reference path (oracle)."""
    import ipoke_b200 as ipk
    B, T, S = 2, 10, 64
    fcfg = O.flow_config(flow_in_channels=32, flow_mid_channels=2048, h_channels=128)
    dcfg = O.first_stage_config(z_dim=32, spatial=S)
    icfg, pcfg = O.cond_encoder_config(nf_in=3, spatial=S), O.cond_encoder_config(nf_in=2, spatial=S)
    fsd, dsd = O.synth_flow_state_dict(fcfg, seed=0), O.synth_first_stage_state_dict(dcfg, seed=1)
    isd, psd = O.synth_cond_encoder_state_dict(icfg, seed=2), O.synth_cond_encoder_state_dict(pcfg, seed=3)
    fc = dict(fcfg); fc.update(ipk_precision="fp32", ipk_max_batch=B)
    dc = dict(dcfg); dc.update(ipk_precision="fp32", ipk_max_batch=B, ipk_max_frames=T)
    flow = ipk.SupervisedMacowTransformer(fc); flow.load_state_dict(fsd, strict=True)
    fs = ipk.SpadeCondMotionDecoder(dc); fs.load_state_dict(dsd, strict=True)
    img = ipk.ConvEncoder(3, 64, icfg["n_stages"], ipk_max_batch=B); img.load_state_dict(isd, strict=True)
    pk = ipk.ConvEncoder(2, 64, pcfg["n_stages"], ipk_max_batch=B); pk.load_state_dict(psd, strict=True)
    s = ipk.PokeMotionSampler(flow.cuda().eval(), fs.cuda().eval(), img.cuda().eval(), pk.cuda().eval())
    g = torch.Generator().manual_seed(42)
    X = torch.rand((B, T + 1, 3, S, S), generator=g) * 2 - 1
    poke, _ = FO.synth_pokes(B, S, seed=42)
    torch.manual_seed(42)
    vids = s.forward_sample(X.cuda(), poke=poke.cuda(), n_samples=2, n_logged_vids=B)
    torch.manual_seed(42)
    with torch.no_grad():
        cond = _cond(dict(img=isd, poke=psd), dict(img=icfg, poke=pcfg), X[:, 0], poke)
        for v in vids:
            z = torch.randn((B, 32, 8, 8))
            ref = O.sample_videos(fsd, fcfg, dsd, dcfg, z, cond, X[:, 0], T)
            assert tuple(v.shape) == (B, T, 3, S, S) and not v.is_cuda
            e = maxabs(v, ref)
            print(f"config 1 sample: per-frame max-abs vs CPU reference path {e:.2e}")
            assert e < 1e-3
