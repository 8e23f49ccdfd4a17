"""CPU ORACLE for the iPOKE sampling hot path -- TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch CPU tensors, fp32 or fp64, explicit loops where the reference
has them) of the reference's algorithm for the hot path of SURVEY.md section 8.  It is NOT part of the
product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it, and only as the checker / the timed CPU baseline.  The product (`ipoke_b200`) never
imports this module and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no golden vectors or tests for this path, so the
oracle is pinned against the UNMODIFIED reference modules imported from /root/reference in the build
container: `tests/golden/make_golden.py` loads the synthetic state-dicts produced here into the
reference's own `SupervisedMacowTransformer`, `ConvGRU` and `SpadeCondConvDecoder`
(`load_state_dict(strict=True)`), runs them, and commits their outputs under `tests/golden/*.pt`;
`tests/test_oracle_golden.py` re-checks the oracle against those fixtures everywhere.

All citations are relative to /root/reference.  Leaf arithmetic (conv2d, conv_transpose2d, group_norm,
instance_norm, bilinear interpolate) is PyTorch ATen, exactly what the reference calls.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

DEFAULT_NUM_STEPS = [10, 5, 5, 4, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1, 1]  # config/second_stage.yaml:58-68


# ----------------------------------------------------------------------------------------------
# configs
# ----------------------------------------------------------------------------------------------
def flow_config(flow_in_channels=32, flow_mid_channels=2048, h_channels=128, num_steps=None, factor=16,
                kernel_size=(2, 3)) -> dict:
    """Keys read by SupervisedMacowTransformer.__init__ (models/modules/INN/INN.py:448-467)."""
    return dict(flow_in_channels=flow_in_channels, flow_mid_channels=flow_mid_channels, h_channels=h_channels,
                num_steps=list(DEFAULT_NUM_STEPS if num_steps is None else num_steps), factor=factor,
                transform="affine", prior_transform="affine", kernel_size=list(kernel_size),
                coupling_type="conv", activation="elu", condition_nice=False, attention=False,
                flow_attn_heads=4, cond_conv=False, cond_conv_hidden_channels=256, p_dropout=0.0)


def first_stage_config(z_dim=32, spatial=128, dec_channels=None, n_gru_layers=4, min_spatial_size=8) -> dict:
    """config['architecture'] of the first stage (config/first_stage.yaml:50-63)."""
    if dec_channels is None:
        dec_channels = [256, 256, 256, 128, 64] if spatial == 128 else [256, 256, 128, 64]
    return dict(z_dim=z_dim, norm="group", spectral_norm=True, running_stats=False, n_gru_layers=n_gru_layers,
                dec_channels=list(dec_channels), min_spatial_size=min_spatial_size, motion_bias=True,
                spatial=spatial)


def flow_levels(cfg: dict) -> List[dict]:
    """Channel bookkeeping of MultiScaleInternal.__init__ (macow2.py:825-871)."""
    C = cfg["flow_in_channels"]
    factor = cfg["factor"]
    assert len(cfg["num_steps"]) < factor  # macow2.py:834
    step = C // factor
    out = []
    for L, n in enumerate(cfg["num_steps"]):
        prior_out = C // factor                 # MultiScalePrior: in_channels // factor (macow2.py:564)
        out.append(dict(level=L, C=C, steps=n, prior_factor=factor, prior_out=prior_out, z1=C - prior_out))
        C = C - step
        assert C == out[-1]["z1"]               # macow2.py:868
        factor -= 1
    return out


# ----------------------------------------------------------------------------------------------
# synthetic checkpoints (reference state-dict layout, SURVEY.md section 5 "Checkpoint / resume")
# ----------------------------------------------------------------------------------------------
def _uni(gen, shape, bound, dtype=torch.float32):
    return (torch.rand(shape, generator=gen, dtype=dtype) * 2.0 - 1.0) * bound


def _nrm(gen, shape, std, dtype=torch.float32):
    return torch.randn(shape, generator=gen, dtype=dtype) * std


def synth_flow_state_dict(cfg: dict, seed: int = 0, g_scale: float = 0.05) -> Dict[str, Tensor]:
    """Seeded synthetic flow checkpoint with the reference's exact keys/shapes/dtypes.

    Fresh-init reference weights make every coupling the identity (zero_init, macow_utils.py:281,423),
    so parity on them is vacuous; we draw non-trivial but stable values instead (SURVEY.md section 7
    "Synthetic weights"): plain convs U(+-1/sqrt(fan_in)) (the nn.Conv2d default), weight_v N(0,0.05)
    (macow_utils.py:222), weight_g U(0, g_scale), biases N(0,0.02), ActNorm log_scale N(0,0.01) /
    bias N(0,0.02), `initialized`=1 so no data-dependent init fires (macow2.py:503, macow_utils.py:248).
    """
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    Hd, hc = cfg["flow_mid_channels"], cfg["h_channels"]
    kH, kW = cfg["kernel_size"]
    one = torch.tensor(1, dtype=torch.uint8)

    def actnorm(p, C):
        sd[p + "log_scale"] = _nrm(gen, (C, 1, 1), 0.01)
        sd[p + "bias"] = _nrm(gen, (C, 1, 1), 0.02)
        sd[p + "initialized"] = one.clone()

    def shuffle(p, C):
        idx = torch.randperm(C, generator=gen)
        sd[p + "forward_shuffle_idx"] = idx
        sd[p + "backward_shuffle_idx"] = torch.argsort(idx)

    def wn_conv(p, cout, cin, kh, kw):
        sd[p + "initialized"] = one.clone()
        sd[p + "conv.bias"] = _nrm(gen, (cout,), 0.02)
        sd[p + "conv.weight_g"] = torch.rand((cout, 1, 1, 1), generator=gen) * g_scale
        sd[p + "conv.weight_v"] = _nrm(gen, (cout, cin, kh, kw), 0.05)

    def mcf(p, C, kh, kw):
        hid = 4 * C if C <= 96 else min(2 * C, 512)          # macow2.py:36-40
        sd[p + "net.shift_conv.weight"] = _uni(gen, (hid, C, kh, kw), 1.0 / math.sqrt(C * kh * kw))
        wn_conv(p + "net.conv1x1.", 2 * C, hid + hc, 1, 1)

    def nice(p, C, factor):
        cout = C // factor
        cin = C - cout
        sd[p + "net.conv1.weight"] = _uni(gen, (Hd, cin, 3, 3), 1.0 / math.sqrt(cin * 9))
        sd[p + "net.conv2.weight"] = _uni(gen, (Hd, Hd, 1, 1), 1.0 / math.sqrt(Hd))
        wn_conv(p + "net.conv3.", 2 * cout, Hd, 3, 3)

    def unit(p, C):
        mcf(p + "conv1.", C, kH, kW)
        mcf(p + "conv2.", C, kH, kW)
        actnorm(p + "actnorm1.", C)
        mcf(p + "conv3.", C, kW, kH)
        mcf(p + "conv4.", C, kW, kH)
        actnorm(p + "actnorm2.", C)

    for lv in flow_levels(cfg):
        L, C = lv["level"], lv["C"]
        for s in range(lv["steps"]):
            p = f"flow.layers.{L}.{s}."
            actnorm(p + "actnorm1.", C)
            shuffle(p + "conv1x1.", C)
            unit(p + "units1.0.", C)
            unit(p + "units1.1.", C)
            nice(p + "coupling1_up.", C, 2)
            nice(p + "coupling1_dn.", C, 2)
            actnorm(p + "actnorm2.", C)
            unit(p + "units2.0.", C)
            unit(p + "units2.1.", C)
            nice(p + "coupling2_up.", C, 2)
            nice(p + "coupling2_dn.", C, 2)
        p = f"flow.priors.{L}."
        shuffle(p + "conv1x1.", C)
        nice(p + "coupling.", C, lv["prior_factor"])
        actnorm(p + "actnorm.", lv["prior_out"])
        shuffle(f"flow.shuffle_layers.{L}.", C)
    return sd


def synth_first_stage_state_dict(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded synthetic first-stage decoder checkpoint: keys of SpadeCondMotionModel.{rnn,gen,motion_bias}
    (first_stage_motion_model.py:482-496).  Spectral-normed convs carry weight_orig/weight_u/weight_v
    (legacy torch.nn.utils.spectral_norm, util.py:51-54,251-254)."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    z = cfg["z_dim"]
    dec = cfg["dec_channels"]

    def conv(p, cout, cin, k=3, snorm=False, transposed=False, bias=True):
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        fan_in = (cout if transposed else cin) * k * k
        w = _uni(gen, shape, 1.0 / math.sqrt(fan_in))
        if snorm:
            sd[p + "weight_orig"] = w
            dim = 1 if transposed else 0                  # torch spectral_norm default dim for ConvTranspose
            hmat = shape[dim]
            wmat = w.numel() // hmat
            u = F.normalize(_nrm(gen, (hmat,), 1.0), dim=0, eps=1e-12)
            v = F.normalize(_nrm(gen, (wmat,), 1.0), dim=0, eps=1e-12)
            sd[p + "weight_u"] = u
            sd[p + "weight_v"] = v
        else:
            sd[p + "weight"] = w
        if bias:
            sd[p + "bias"] = _nrm(gen, (cout,), 0.05)

    # ConvGRU (rnn.py:12-28)
    for i in range(cfg["n_gru_layers"]):
        for gate in ("reset_gate", "update_gate", "out_gate"):
            conv(f"rnn.cells.{i}.{gate}.", z, 2 * z)
    sd["motion_bias"] = _nrm(gen, (1, z, cfg["min_spatial_size"], cfg["min_spatial_size"]), 1.0)

    # in_block: ResBlock(z, dec[0], snorm, norm='group') (fully_conv_models.py:151, util.py:140-182)
    p = "gen.in_block."
    conv(p + "conv1.conv.", dec[0], z, snorm=True)
    sd[p + "conv1.norm.weight"] = 1.0 + _nrm(gen, (dec[0],), 0.1)
    sd[p + "conv1.norm.bias"] = _nrm(gen, (dec[0],), 0.1)
    conv(p + "conv2.conv.", dec[0], dec[0], snorm=True)
    sd[p + "conv2.norm.weight"] = 1.0 + _nrm(gen, (dec[0],), 0.1)
    sd[p + "conv2.norm.bias"] = _nrm(gen, (dec[0],), 0.1)
    if z != dec[0]:
        conv(p + "res_conv.conv.", dec[0], z, snorm=True)
    # up blocks + SPADE (fully_conv_models.py:153-161)
    for i, nf in enumerate(dec[1:]):
        p = f"gen.blocks.{i}."
        conv(p + "conv1.conv.", nf, dec[i], snorm=True, transposed=True)
        conv(p + "conv2.conv.", nf, nf, snorm=True)
        conv(p + "res_conv.conv.", nf, dec[i], snorm=True, transposed=True)
        p = f"gen.spade_blocks.{i}."
        conv(p + "conv.", 128, 3)
        conv(p + "conv_gamma.", nf, 128)
        conv(p + "conv_beta.", nf, 128)
    conv("gen.out_conv.conv.", 3, dec[-1])
    return sd


# ----------------------------------------------------------------------------------------------
# flow building blocks
# ----------------------------------------------------------------------------------------------
def weight_norm_weight(v: Tensor, g: Tensor) -> Tensor:
    """Legacy torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||_2 over all dims but 0
    (macow_utils.py:229; recomputed on every call by the forward pre-hook)."""
    norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


def affine_params(params: Tensor) -> Tuple[Tensor, Tensor]:
    """Affine.calc_params (macow_utils.py:49-52), alpha = 1."""
    mu, log_scale = params.chunk(2, dim=1)
    scale = torch.tanh(log_scale * 0.5) * 1.0 + 1.0
    return mu, scale


def actnorm_data_init(sd, p, x, init_scale=1.0):
    """ActNorm2dFlow.init (macow2.py:526-539): per-channel mean / UNBIASED std of x * exp(ls0) + b0 over (B, H, W); writes
    log_scale = log(init_scale / (std + 1e-6)), bias = -mean * inv_stdv and the `initialized` flag into sd (in place)."""
    ls, b = sd[p + "log_scale"], sd[p + "bias"]
    C = x.shape[1]
    out = x * ls.exp() + b
    out = out.transpose(0, 1).contiguous().view(C, -1)
    mean = out.mean(dim=1).view(C, 1, 1)
    std = out.std(dim=1).view(C, 1, 1)
    inv_stdv = init_scale / (std + 1e-6)
    sd[p + "log_scale"] = inv_stdv.log().to(ls.dtype)
    sd[p + "bias"] = (-mean * inv_stdv).to(b.dtype)
    sd[p + "initialized"] = torch.tensor(1, dtype=torch.uint8)


def weightnorm_data_init(sd, p, x, padding, zero_init=True):
    """p = prefix of the Conv2dWeightNorm module (its nn.Conv2d lives at p + 'conv.').  Conv2dWeightNorm.init (macow_utils.py:231-246) as called from forward (:248-250) with init_scale = 0 for zero_init layers (every
    weight-normed conv of the flow: macow_utils.py:281,423): weight_g = init_scale / (std + 1e-6), bias = -mean * inv_stdv."""
    v, g, b = sd[p + "conv.weight_v"], sd[p + "conv.weight_g"], sd[p + "conv.bias"]
    out = F.conv2d(x, weight_norm_weight(v, g).to(x.dtype), b.to(x.dtype), padding=padding)
    n = out.shape[1]
    out = out.transpose(0, 1).contiguous().view(n, -1)
    mean, std = out.mean(dim=1), out.std(dim=1)
    inv_stdv = (0.0 if zero_init else 1.0) / (std + 1e-6)
    sd[p + "conv.weight_g"] = inv_stdv.view(n, 1, 1, 1).to(g.dtype)
    sd[p + "conv.bias"] = (-mean * inv_stdv).to(b.dtype)
    sd[p + "initialized"] = torch.tensor(1, dtype=torch.uint8)


def _uninitialised(sd, p):
    return (p + "initialized") in sd and int(sd[p + "initialized"]) == 0


def actnorm_fwd(sd, p, x):
    """ActNorm2dFlow.forward (macow2.py:507-513); fires the data-dependent init first when `initialized` is 0 (:503-505)."""
    if _uninitialised(sd, p):
        actnorm_data_init(sd, p, x)
    ls, b = sd[p + "log_scale"].to(x.dtype), sd[p + "bias"].to(x.dtype)
    B, C, H, W = x.shape
    out = x * ls.exp() + b
    logdet = ls.view(1, C).sum(dim=1) * (H * W)
    return out, logdet * torch.ones(B, dtype=x.dtype, device=x.device)


def actnorm_bwd(sd, p, y):
    """ActNorm2dFlow.forward(reverse=True) (macow2.py:515-520): divide by exp(ls)+1e-8."""
    ls, b = sd[p + "log_scale"].to(y.dtype), sd[p + "bias"].to(y.dtype)
    return (y - b) / (ls.exp() + 1e-8)


def shuffle_fwd(sd, p, x):
    """Shuffle.forward (flow_blocks.py:322-324)."""
    return x[:, sd[p + "forward_shuffle_idx"]]


def shuffle_bwd(sd, p, x):
    """Shuffle.forward(reverse=True) (flow_blocks.py:325-326)."""
    return x[:, sd[p + "backward_shuffle_idx"]]


_SHIFT = {  # ShiftedConv2d pad/cut table (macow_utils.py:465-484): (pad l,r,t,b), cut (t,b,l,r)
    "A": lambda kh, kw: (((kw - 1) // 2, (kw - 1) // 2, kh, 0), (0, -1, 0, 0)),
    "B": lambda kh, kw: (((kw - 1) // 2, (kw - 1) // 2, 0, kh), (1, 0, 0, 0)),
    "C": lambda kh, kw: ((kw, 0, (kh - 1) // 2, (kh - 1) // 2), (0, 0, 0, -1)),
    "D": lambda kh, kw: ((0, kw, (kh - 1) // 2, (kh - 1) // 2), (0, 0, 1, 0)),
}


def mcf_block(sd, p, x, h, order, shifted=True):
    """MCFBlock.forward (macow_utils.py:427-434) with ShiftedConv2d.forward (macow_utils.py:492-499)."""
    w = sd[p + "net.shift_conv.weight"].to(x.dtype)
    kh, kw = w.shape[2], w.shape[3]
    if shifted:
        pad, cut = _SHIFT[order](kh, kw)
        x = F.pad(x, pad)
        t, b, l, r = cut
        x = x[:, :, t:x.shape[2] + b, l:x.shape[3] + r]
    c = F.conv2d(x, w)
    c = torch.cat([c, h], dim=1)
    c = F.elu(c)
    if _uninitialised(sd, p + "net.conv1x1."):
        weightnorm_data_init(sd, p + "net.conv1x1.", c, 0)
    w1 = weight_norm_weight(sd[p + "net.conv1x1.conv.weight_v"], sd[p + "net.conv1x1.conv.weight_g"]).to(x.dtype)
    return F.conv2d(c, w1, sd[p + "net.conv1x1.conv.bias"].to(x.dtype))


def mcf_fwd(sd, p, x, h, order):
    """MaskedConvFlow.forward (macow2.py:113-116) + Affine.fwd (macow_utils.py:54-59)."""
    mu, scale = affine_params(mcf_block(sd, p, x, h, order))
    out = scale * x + mu
    return out, scale.log().reshape(x.shape[0], -1).sum(dim=1)


def mcf_bwd(sd, p, y, h, order):
    """MaskedConvFlow.backward_height / backward_width (macow2.py:212-231, 269-288): sequential rows/cols."""
    B, C, H, W = y.shape
    w = sd[p + "net.shift_conv.weight"]
    kH, kW = w.shape[2], w.shape[3]
    if order in ("A", "B"):
        reverse = order == "B"
        cW = kW // 2
        out = y.new_zeros(B, C, H + kH, W + 2 * cW)
        for r in (reversed(range(H)) if reverse else range(H)):
            curr = r if reverse else r + kH
            s = r + 1 if reverse else r
            t = r + kH + 1 if reverse else r + kH
            params = mcf_block(sd, p, out[:, :, s:t], h[:, :, r:r + 1], order, shifted=False)
            mu, scale = affine_params(params.squeeze(2))
            out[:, :, curr, cW:W + cW] = (y[:, :, r] - mu) / (scale + 1e-12)      # Affine.bwd macow_utils.py:62-64
        return out[:, :, :H, cW:cW + W] if reverse else out[:, :, kH:, cW:cW + W]
    reverse = order == "D"
    cH = kH // 2
    out = y.new_zeros(B, C, H + 2 * cH, W + kW)
    for c in (reversed(range(W)) if reverse else range(W)):
        curr = c if reverse else c + kW
        s = c + 1 if reverse else c
        t = c + kW + 1 if reverse else c + kW
        params = mcf_block(sd, p, out[:, :, :, s:t], h[:, :, :, c:c + 1], order, shifted=False)
        mu, scale = affine_params(params.squeeze(3))
        out[:, :, cH:H + cH, curr] = (y[:, :, :, c] - mu) / (scale + 1e-12)
    return out[:, :, cH:cH + H, :W] if reverse else out[:, :, cH:cH + H, kW:]


def nice_net(sd, p, z):
    """NICEConvBlock.forward (macow_utils.py:313-337), normalize=None, no h (condition_nice: false)."""
    out = F.conv2d(z, sd[p + "net.conv1.weight"].to(z.dtype), padding=1)
    out = F.elu(out)
    out = F.conv2d(out, sd[p + "net.conv2.weight"].to(z.dtype))
    out = F.elu(out)
    if _uninitialised(sd, p + "net.conv3."):
        weightnorm_data_init(sd, p + "net.conv3.", out, 1)
    w3 = weight_norm_weight(sd[p + "net.conv3.conv.weight_v"], sd[p + "net.conv3.conv.weight_g"]).to(z.dtype)
    return F.conv2d(out, w3, sd[p + "net.conv3.conv.bias"].to(z.dtype), padding=1)


def nice_split_indices(C: int, factor: int, split_type: str, up: bool):
    """Channel index lists (z = network input, zp = transformed part) of NICE2d.split/unsplit
    (macow2.py:301-317,364-388).  Returns (idx_z, idx_zp) as python lists into the C input channels."""
    if split_type == "skip" and C % 2 == 1:
        split_type = "continuous"                      # macow2.py:303-307
    cout = C // factor
    cin = C - cout
    z1c = cin if up else cout
    if split_type == "continuous":
        i1, i2 = list(range(0, z1c)), list(range(z1c, C))
    else:
        i1, i2 = list(range(0, C, 2)), list(range(1, C, 2))
    return (i1, i2) if up else (i2, i1)


def nice_apply(sd, p, x, factor, split_type, up, reverse):
    """NICE2d.forward / backward_analytic (macow2.py:395-448)."""
    C = x.shape[1]
    iz, ip = nice_split_indices(C, factor, split_type, up)
    z, zp = x[:, iz], x[:, ip]
    mu, scale = affine_params(nice_net(sd, p, z))
    out = x.clone()
    if not reverse:
        out[:, ip] = scale * zp + mu
        return out, scale.log().reshape(x.shape[0], -1).sum(dim=1)
    out[:, ip] = (zp - mu) / (scale + 1e-12)
    return out


_UNIT_FWD = (("mcf", "conv1.", "A"), ("mcf", "conv2.", "B"), ("act", "actnorm1.", None),
             ("mcf", "conv3.", "C"), ("mcf", "conv4.", "D"), ("act", "actnorm2.", None))


def unit_fwd(sd, p, x, h):
    """MaCowUnit.forward (macow2.py:962-980)."""
    ld = x.new_zeros(x.shape[0])
    for kind, name, order in _UNIT_FWD:
        x, l = mcf_fwd(sd, p + name, x, h, order) if kind == "mcf" else actnorm_fwd(sd, p + name, x)
        ld = ld + l
    return x, ld


def unit_bwd(sd, p, x, h):
    """MaCowUnit.forward(reverse=True) (macow2.py:981-995)."""
    for kind, name, order in reversed(_UNIT_FWD):
        x = mcf_bwd(sd, p + name, x, h, order) if kind == "mcf" else actnorm_bwd(sd, p + name, x)
    return x


def step_fwd(sd, p, x, h):
    """MaCowStep.forward (macow2.py:1066-1091)."""
    x, ld = actnorm_fwd(sd, p + "actnorm1.", x)
    x = shuffle_fwd(sd, p + "conv1x1.", x)
    for u in ("units1.0.", "units1.1."):
        x, l = unit_fwd(sd, p + u, x, h); ld = ld + l
    x, l = nice_apply(sd, p + "coupling1_up.", x, 2, "continuous", True, False); ld = ld + l
    x, l = nice_apply(sd, p + "coupling1_dn.", x, 2, "continuous", False, False); ld = ld + l
    x, l = actnorm_fwd(sd, p + "actnorm2.", x); ld = ld + l
    for u in ("units2.0.", "units2.1."):
        x, l = unit_fwd(sd, p + u, x, h); ld = ld + l
    x, l = nice_apply(sd, p + "coupling2_up.", x, 2, "skip", True, False); ld = ld + l
    x, l = nice_apply(sd, p + "coupling2_dn.", x, 2, "skip", False, False); ld = ld + l
    return x, ld


def step_bwd(sd, p, x, h):
    """MaCowStep.forward(reverse=True) (macow2.py:1092-1117)."""
    x = nice_apply(sd, p + "coupling2_dn.", x, 2, "skip", False, True)
    x = nice_apply(sd, p + "coupling2_up.", x, 2, "skip", True, True)
    for u in ("units2.1.", "units2.0."):
        x = unit_bwd(sd, p + u, x, h)
    x = actnorm_bwd(sd, p + "actnorm2.", x)
    x = nice_apply(sd, p + "coupling1_dn.", x, 2, "continuous", False, True)
    x = nice_apply(sd, p + "coupling1_up.", x, 2, "continuous", True, True)
    for u in ("units1.1.", "units1.0."):
        x = unit_bwd(sd, p + u, x, h)
    x = shuffle_bwd(sd, p + "conv1x1.", x)
    return actnorm_bwd(sd, p + "actnorm1.", x)


def prior_fwd(sd, p, x, lv):
    """MultiScalePrior.forward (macow2.py:569-581)."""
    x = shuffle_fwd(sd, p + "conv1x1.", x)
    x, ld = nice_apply(sd, p + "coupling.", x, lv["prior_factor"], "continuous", True, False)
    x1, x2 = x[:, :lv["z1"]], x[:, lv["z1"]:]
    x2, l = actnorm_fwd(sd, p + "actnorm.", x2)
    return torch.cat([x1, x2], dim=1), ld + l


def prior_bwd(sd, p, x, lv):
    """MultiScalePrior.forward(reverse=True) (macow2.py:582-593)."""
    x1, x2 = x[:, :lv["z1"]], x[:, lv["z1"]:]
    x = torch.cat([x1, actnorm_bwd(sd, p + "actnorm.", x2)], dim=1)
    x = nice_apply(sd, p + "coupling.", x, lv["prior_factor"], "continuous", True, True)
    return shuffle_bwd(sd, p + "conv1x1.", x)


def flow_forward(sd, cfg, x: Tensor, cond: Tensor) -> Tuple[Tensor, Tensor]:
    """SupervisedMacowTransformer.forward (INN.py:469-473) -> MultiScaleInternal.forward (macow2.py:873-900)."""
    ld = x.new_zeros(x.shape[0])
    outputs = []
    out = x
    for lv in flow_levels(cfg):
        L = lv["level"]
        for s in range(lv["steps"]):
            out, l = step_fwd(sd, f"flow.layers.{L}.{s}.", out, cond); ld = ld + l
        out, l = prior_fwd(sd, f"flow.priors.{L}.", out, lv); ld = ld + l
        out = shuffle_fwd(sd, f"flow.shuffle_layers.{L}.", out)
        outputs.append(out[:, lv["z1"]:])
        out = out[:, :lv["z1"]]
    outputs.append(out)
    outputs.reverse()
    return torch.cat(outputs, dim=1), ld


def flow_reverse(sd, cfg, z: Tensor, cond: Tensor) -> Tensor:
    """SupervisedMacowTransformer.reverse (INN.py:475-476) -> MultiScaleInternal reverse (macow2.py:901-920)."""
    levels = flow_levels(cfg)
    out = z
    outputs = []
    for lv in levels:
        outputs.append(out[:, lv["z1"]:])
        out = out[:, :lv["z1"]]
    for lv in reversed(levels):
        L = lv["level"]
        out = torch.cat([out, outputs.pop()], dim=1)
        out = shuffle_bwd(sd, f"flow.shuffle_layers.{L}.", out)
        out = prior_bwd(sd, f"flow.priors.{L}.", out, lv)
        for s in reversed(range(lv["steps"])):
            out = step_bwd(sd, f"flow.layers.{L}.{s}.", out, cond)
    assert not outputs
    return out


def flow_nll(z: Tensor, logdet: Tensor) -> Tensor:
    """FlowLoss.forward (loss.py:13-31) with logdet_weight = 1, spatial_mean = False: mean(0.5*sum z^2) - mean(logdet)."""
    return (0.5 * (z ** 2).flatten(1).sum(dim=1)).mean() - logdet.mean()


def flow_loss_log(z: Tensor, logdet: Tensor, spatial_mean: bool = False, logdet_weight: float = 1.0):
    """FlowLoss.forward with its log dict (loss.py:13-31, nll :75-79); draws torch.randn_like(z) from the default generator."""
    def nll(s):
        return 0.5 * torch.sum(torch.mean(s ** 2, dim=[2, 3]), dim=1) if spatial_mean else 0.5 * torch.sum(s ** 2, dim=[1, 2, 3])
    nll_loss = torch.mean(nll(z))
    nlogdet_loss = -torch.mean(logdet) / (z.shape[-2] * z.shape[-1]) if spatial_mean else -torch.mean(logdet)
    loss = nll_loss + logdet_weight * nlogdet_loss
    ref = torch.mean(nll(torch.randn_like(z)))
    return loss, {"flow_loss": loss, "reference_nll_loss": ref, "nlogdet_loss": nlogdet_loss, "nll_loss": nll_loss, "logdet_weight": logdet_weight}


def flow_data_init(sd, cfg, x: Tensor, cond: Tensor):
    """First density-direction forward of a flow whose `initialized` buffers are 0: returns (initialised copy of sd, z, logdet).
    The init fires layer by layer inside the forward, exactly where the reference's modules fire it."""
    sd = dict(sd)
    z, ld = flow_forward(sd, cfg, x, cond)
    return sd, z, ld


def flow_trainable_keys(sd) -> List[str]:
    """The parameters the second-stage optimizer updates (flow.parameters(), second_stage_video.py:633): every floating-point
    tensor of the flow checkpoint; the uint8 `initialized` flags and int64 shuffle indices are buffers."""
    return [k for k, v in sd.items() if v.is_floating_point()]


def flow_loss_and_grads(sd, cfg, x: Tensor, cond: Tensor) -> Tuple[Tensor, Dict[str, Tensor]]:
    """Second-stage training step without the optimizer (second_stage_video.py:345-350 forward_density, loss.py:13-31 FlowLoss,
    then loss.backward()): loss = mean_B(0.5 * sum z^2) - mean_B(logdet) and its gradient w.r.t. every trainable flow tensor,
    by autograd over the restated forward (the reference differentiates the same ops)."""
    leaf = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    z, ld = flow_forward(leaf, cfg, x, cond)
    loss = flow_nll(z, ld)
    keys = flow_trainable_keys(sd)
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys], allow_unused=True)
    return loss.detach(), {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(keys, grads)}


# ----------------------------------------------------------------------------------------------
# first-stage decoder (ConvGRU + SPADE decoder)
# ----------------------------------------------------------------------------------------------
def spectral_weight(sd, p, transposed=False) -> Tensor:
    """Legacy torch.nn.utils.spectral_norm in eval mode: W / sigma with sigma = u^T W_mat v from the STORED
    u, v (no power iteration when module.training is False).  dim=1 for ConvTranspose2d (util.py:51-54)."""
    if p + "weight" in sd:
        return sd[p + "weight"]
    w = sd[p + "weight_orig"]
    dim = 1 if transposed else 0
    wm = w if dim == 0 else w.permute(dim, *[d for d in range(w.dim()) if d != dim])
    wm = wm.reshape(wm.shape[0], -1)
    sigma = torch.dot(sd[p + "weight_u"], torch.mv(wm, sd[p + "weight_v"]))
    return w / sigma


def gru_cell(sd, p, x, h):
    """ConvGRUCell.forward (rnn.py:32-56)."""
    dt = x.dtype
    st = torch.cat([x, h], dim=1)
    u = torch.sigmoid(F.conv2d(st, sd[p + "update_gate.weight"].to(dt), sd[p + "update_gate.bias"].to(dt), padding=1))
    r = torch.sigmoid(F.conv2d(st, sd[p + "reset_gate.weight"].to(dt), sd[p + "reset_gate.bias"].to(dt), padding=1))
    o = torch.tanh(F.conv2d(torch.cat([x, h * r], dim=1), sd[p + "out_gate.weight"].to(dt),
                            sd[p + "out_gate.bias"].to(dt), padding=1))
    return h * (1 - u) + o * u


def gru_step(sd, cfg, x, hidden: List[Tensor]) -> List[Tensor]:
    """ConvGRU.forward (rnn.py:104-133)."""
    new = []
    inp = x
    for i in range(cfg["n_gru_layers"]):
        hi = gru_cell(sd, f"rnn.cells.{i}.", inp, hidden[i])
        new.append(hi)
        inp = hi
    return new


def _conv_block(sd, p, x, norm, act, dt, stride=1):
    """Conv2dBlock.forward (util.py:256-273): ZeroPad(1) + 3x3 conv [+ norm] [+ activation]."""
    w = spectral_weight(sd, p + "conv.").to(dt)
    x = F.conv2d(x, w, sd[p + "conv.bias"].to(dt), stride=stride, padding=1)
    if norm == "group":
        x = F.group_norm(x, 16, sd[p + "norm.weight"].to(dt), sd[p + "norm.bias"].to(dt), eps=1e-5)
    elif norm == "in":
        x = F.instance_norm(x, eps=1e-5)
    if act == "elu":
        x = F.elu(x)                                   # Conv2dBlock maps "elu" -> nn.ELU (util.py:245-246)
    elif act == "tanh":
        x = torch.tanh(x)
    return x


def _convT_block(sd, p, x, norm, dt):
    """Conv2dTransposeBlock.forward (util.py:56-73): ConvTranspose2d(3, s2, p1, op1) [+ IN] + ReLU
    (activation "elu" maps to nn.ReLU here, util.py:41-42)."""
    w = spectral_weight(sd, p + "conv.", transposed=True).to(dt)
    x = F.conv_transpose2d(x, w, sd[p + "conv.bias"].to(dt), stride=2, padding=1, output_padding=1)
    if norm == "in":
        x = F.instance_norm(x, eps=1e-5)
    return F.relu(x)


def spade(sd, p, x, x0):
    """Spade.forward (util.py:494-500), GroupNorm(16, affine=False)."""
    dt = x.dtype
    C = x.shape[1]
    g = 16
    while C % g != 0:
        g -= 1
    normalized = F.group_norm(x, g, eps=1e-5)
    y = F.interpolate(x0, mode="bilinear", size=x.shape[-2:], align_corners=True)
    y = F.leaky_relu(F.conv2d(y, sd[p + "conv.weight"].to(dt), sd[p + "conv.bias"].to(dt), padding=1), 0.2)
    gamma = F.conv2d(y, sd[p + "conv_gamma.weight"].to(dt), sd[p + "conv_gamma.bias"].to(dt), padding=1)
    beta = F.conv2d(y, sd[p + "conv_beta.weight"].to(dt), sd[p + "conv_beta.bias"].to(dt), padding=1)
    return normalized * (1 + gamma) + beta


def decoder_forward(sd, cfg, h: Tensor, x0: Tensor) -> Tensor:
    """SpadeCondConvDecoder.forward (fully_conv_models.py:166-177) incl. ResBlock.forward (util.py:185-192)."""
    dt = h.dtype
    p = "gen.in_block."
    res = h
    if (p + "res_conv.conv.weight_orig") in sd or (p + "res_conv.conv.weight") in sd:
        res = _conv_block(sd, p + "res_conv.", h, "in", "elu", dt)
    x = _conv_block(sd, p + "conv1.", h, "group", "elu", dt)
    x = _conv_block(sd, p + "conv2.", x, "group", "none", dt)
    x = x + res
    for i in range(len(cfg["dec_channels"]) - 1):
        p = f"gen.blocks.{i}."
        res = _convT_block(sd, p + "res_conv.", x, "in", dt)
        y = _convT_block(sd, p + "conv1.", x, "none", dt)
        y = _conv_block(sd, p + "conv2.", y, "none", "none", dt)
        x = y + res
        x = spade(sd, f"gen.spade_blocks.{i}.", x, x0)
    return _conv_block(sd, "gen.out_conv.", x, "none", "tanh", dt)


def decode_first_stage(sd, cfg, motion: Tensor, x0: Tensor, length: int) -> Tensor:
    """PokeMotionModel.decode_first_stage (second_stage_video.py:361-382) for SpadeCondMotionModel."""
    B = motion.shape[0]
    hidden = [motion] * cfg["n_gru_layers"]
    in_rnn = torch.cat([sd["motion_bias"].to(motion.dtype)] * B, dim=0)
    frames = []
    for _ in range(length):
        hidden = gru_step(sd, cfg, in_rnn, hidden)
        frames.append(decoder_forward(sd, cfg, hidden[-1], x0))
    return torch.stack(frames, dim=1)


def sample_videos(flow_sd, flow_cfg, fs_sd, fs_cfg, z: Tensor, cond: Tensor, x0: Tensor, length: int) -> Tensor:
    """forward_sample body (second_stage_video.py:333-337) with z and cond supplied by the caller."""
    motion = flow_reverse(flow_sd, flow_cfg, z, cond)
    return decode_first_stage(fs_sd, fs_cfg, motion, x0, length)


# ----------------------------------------------------------------------------------------------
# first-stage 3-D conv video encoder (training path: PokeMotionModel.encode_first_stage,
# second_stage_video.py:352-359 -> ResNetMotionEncoder, models/modules/motion_models/motion_encoder.py:150-241)
# ----------------------------------------------------------------------------------------------
def encoder_config(z_dim=32, img_size=128, max_frames=10, full_seq=True, channels=None, min_spatial_size=8) -> dict:
    """The keys ResNetMotionEncoder.__init__ reads from dic (motion_encoder.py:152-160); ENC_M_channels as in
    config/first_stage.yaml:60."""
    if channels is None:
        channels = [64, 128, 256, 256, 256] if img_size == 128 else [64, 128, 256, 256]
    return dict(z_dim=z_dim, img_size=img_size, max_frames=max_frames, full_seq=full_seq, ENC_M_channels=list(channels),
                min_spatial_size=min_spatial_size, deterministic=False)


def encoder_layers(cfg: dict) -> List[dict]:
    """Layer plan of ResNetMotionEncoder.__init__ (motion_encoder.py:161-190): list of residual stages
    dict(name, inplanes, planes, stride=(st,sy,sx), blocks=2)."""
    ch = list(cfg["ENC_M_channels"])
    max_frames = cfg["max_frames"]
    first_block_down = (len(ch) - 1 < int(math.ceil(math.log2(max_frames)))) or cfg["full_seq"]      # :164-166
    s1 = (2, 1, 1) if first_block_down else (1, 1, 1)
    stages = [dict(name="layer1", inplanes=ch[0], planes=ch[1], stride=s1),
              dict(name="layer2", inplanes=ch[1], planes=ch[2], stride=(2, 2, 2)),
              dict(name="layer3", inplanes=ch[2], planes=ch[3], stride=(2, 2, 2))]
    stride4 = (2, 1, 1) if (cfg["full_seq"] and max_frames >= 16) else None                             # :172
    if cfg["img_size"] // 2 ** 3 > cfg["min_spatial_size"]:                                             # :174-175
        stride4 = (2, 2, 2)
    if stride4 is not None:
        if len(ch) < 5:
            ch.append(ch[-1])                                                                           # :179-181
        stages.append(dict(name="layer4", inplanes=ch[3], planes=ch[4], stride=stride4))
    if cfg["img_size"] // 2 ** 4 > cfg["min_spatial_size"]:                                             # :184-186
        stages.append(dict(name="layer5", inplanes=ch[-2] if len(ch) > 5 else ch[4], planes=ch[5], stride=(2, 2, 2)))
    return stages


def synth_encoder_state_dict(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded synthetic encoder checkpoint with the reference's keys/shapes: Conv3d kaiming-normal fan_out
    (motion_encoder.py:192-194), GroupNorm weight ~ 1 + 0.1 N(0,1), bias ~ 0.05 N(0,1), 2-D heads nn.Conv2d default."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv3(name, cout, cin, k):
        fan_out = cout * k[0] * k[1] * k[2]
        sd[name] = _nrm(gen, (cout, cin) + tuple(k), math.sqrt(2.0 / fan_out))

    def gn(p, c):
        sd[p + "weight"] = 1.0 + _nrm(gen, (c,), 0.1)
        sd[p + "bias"] = _nrm(gen, (c,), 0.05)

    ch0 = cfg["ENC_M_channels"][0]
    conv3("conv1.weight", ch0, 3, (3, 7, 7))
    gn("bn1.", ch0)
    stages = encoder_layers(cfg)
    for st in stages:
        inp = st["inplanes"]
        for b in range(2):
            p = f"{st['name']}.{b}."
            conv3(p + "conv1.weight", st["planes"], inp, (3, 3, 3))
            gn(p + "bn1.", st["planes"])
            conv3(p + "conv2.weight", st["planes"], st["planes"], (3, 3, 3))
            gn(p + "bn2.", st["planes"])
            if b == 0 and (st["stride"] != (1, 1, 1) or inp != st["planes"]):
                conv3(p + "downsample.0.weight", st["planes"], inp, (1, 1, 1))
                gn(p + "downsample.1.", st["planes"])
            inp = st["planes"]
    last = stages[-1]["planes"]
    for head in ("conv_mu", "conv_var"):
        b = 1.0 / math.sqrt(last * 9)
        sd[head + ".weight"] = _uni(gen, (cfg["z_dim"], last, 3, 3), b)
        sd[head + ".bias"] = _uni(gen, (cfg["z_dim"],), b)
    return sd


def encoder_forward(sd, cfg: dict, X: Tensor, eps: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """ResNetMotionEncoder.forward (motion_encoder.py:224-241) with the reparameterisation noise `eps` supplied by the
    caller (the reference draws it on the CPU generator, :220).  X: [B,3,T,H,W] -> (z, mu, logvar) each [B,z,8,8]."""
    dt = X.dtype
    w = lambda k: sd[k].to(dt)

    def gnorm(x, p):
        return F.group_norm(x, 16, w(p + "weight"), w(p + "bias"), 1e-5)

    x = F.relu(gnorm(F.conv3d(X, w("conv1.weight"), None, stride=(2, 2, 2), padding=(1, 3, 3)), "bn1."))
    for st in encoder_layers(cfg):
        for b in range(2):
            p = f"{st['name']}.{b}."
            stride = st["stride"] if b == 0 else (1, 1, 1)
            out = F.relu(gnorm(F.conv3d(x, w(p + "conv1.weight"), None, stride=stride, padding=1), p + "bn1."))      # BasicBlock :56-74
            out = gnorm(F.conv3d(out, w(p + "conv2.weight"), None, stride=1, padding=1), p + "bn2.")
            res = x
            if (p + "downsample.0.weight") in sd:
                res = gnorm(F.conv3d(x, w(p + "downsample.0.weight"), None, stride=stride), p + "downsample.1.")
            x = F.relu(out + res)
    assert x.shape[2] == 1, f"temporal extent {x.shape[2]} != 1 before squeeze (motion_encoder.py:241)"
    emb = x.squeeze(2)
    mu = F.conv2d(emb, w("conv_mu.weight"), w("conv_mu.bias"), padding=1)
    logvar = F.conv2d(emb, w("conv_var.weight"), w("conv_var.bias"), padding=1)
    z = eps.to(dt) * torch.exp(0.5 * logvar) + mu                                                                    # :218-222
    return z, mu, logvar


# ----------------------------------------------------------------------------------------------
# conditioning encoders (poke embedder / image conditioner): ConvEncoder, fully_conv_models.py:28-94
# (next-row component of SURVEY.md section 8f rank 1; used by make_flow_input, second_stage_video.py:268-287,311)
# ----------------------------------------------------------------------------------------------
def cond_encoder_config(nf_in=3, nf_max=64, spatial=128, min_spatial_size=8) -> dict:
    """FirstStageWrapper wiring (fully_conv_models.py:9-22): n_stages = log2(spatial / min_spatial_size); deterministic."""
    return dict(nf_in=nf_in, nf_max=nf_max, spatial=spatial, min_spatial_size=min_spatial_size,
                n_stages=int(math.log2(spatial // min_spatial_size)))


def cond_encoder_widths(cfg: dict) -> List[int]:
    """Channel widths after each stage of ConvEncoder.__init__ (fully_conv_models.py:38-61): 32, then min(2*nf, nf_max)."""
    w = [32]
    for _ in range(cfg["n_stages"] - 1):
        w.append(min(w[-1] * 2, cfg["nf_max"]))
    return w


def synth_cond_encoder_state_dict(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded checkpoint of ConvEncoder(nf_in, nf_max, n_stages, variational=False) with the reference's keys: stride-2
    blocks carry legacy spectral norm (weight_orig / weight_u / weight_v), the bottleneck ResBlock does not."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(p, cout, cin, sn):
        b = 1.0 / math.sqrt(cin * 9)
        sd[p + "bias"] = _uni(gen, (cout,), b)
        if sn:
            sd[p + "weight_orig"] = _uni(gen, (cout, cin, 3, 3), b)
            sd[p + "weight_u"] = F.normalize(torch.randn(cout, generator=gen), dim=0, eps=1e-12)
            sd[p + "weight_v"] = F.normalize(torch.randn(cin * 9, generator=gen), dim=0, eps=1e-12)
        else:
            sd[p + "weight"] = _uni(gen, (cout, cin, 3, 3), b)

    def gn(p, c):
        sd[p + "weight"] = 1.0 + _nrm(gen, (c,), 0.1)
        sd[p + "bias"] = _nrm(gen, (c,), 0.05)

    widths = cond_encoder_widths(cfg)
    gn("model.0.norm.", widths[0])
    conv("model.0.conv.", widths[0], cfg["nf_in"], True)
    for i in range(1, len(widths)):
        p = f"model.{i}."
        gn(p + "conv1.norm.", widths[i]); conv(p + "conv1.conv.", widths[i], widths[i - 1], True)
        gn(p + "conv2.norm.", widths[i]); conv(p + "conv2.conv.", widths[i], widths[i], True)
        conv(p + "res_conv.conv.", widths[i], widths[i - 1], True)
    nf, nmax = widths[-1], cfg["nf_max"]
    p = "bottleneck.0."
    gn(p + "conv1.norm.", nmax); conv(p + "conv1.conv.", nmax, nf, False)
    gn(p + "conv2.norm.", nmax); conv(p + "conv2.conv.", nmax, nmax, False)
    if nf != nmax:
        conv(p + "res_conv.conv.", nmax, nf, False)
    return sd


def cond_encoder_forward(sd, cfg: dict, x: Tensor) -> Tuple[Tensor, Tensor]:
    """ConvEncoder.forward, variational=False (fully_conv_models.py:74-88): returns (out, mean) with mean = the activations
    before the bottleneck; callers take [0] (second_stage_video.py:274,281)."""
    dt = x.dtype
    x = _conv_block(sd, "model.0.", x, "group", "elu", dt, stride=2)
    widths = cond_encoder_widths(cfg)
    for i in range(1, len(widths)):
        p = f"model.{i}."
        res = _conv_block(sd, p + "res_conv.", x, "in", "elu", dt, stride=2)          # ResBlock :170-176, 185-192
        out = _conv_block(sd, p + "conv1.", x, "group", "elu", dt, stride=2)
        out = _conv_block(sd, p + "conv2.", out, "group", "none", dt)
        x = out + res
    mean = x
    p = "bottleneck.0."
    res = _conv_block(sd, p + "res_conv.", x, "in", "elu", dt) if (p + "res_conv.conv.bias") in sd else x
    out = _conv_block(sd, p + "conv1.", x, "group", "elu", dt)
    out = _conv_block(sd, p + "conv2.", out, "group", "none", dt)
    return out + res, mean


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synth_inputs(B: int, C0: int, h_channels: int, spatial: int, seed: int = 42):
    """z ~ N(0,I) from the CPU generator (second_stage_video.py:300), cond ~ 0.5*N(0,1) standing in for the frozen
    conditioning encoders' output (next-row component), x0 ~ U(-1,1)."""
    gen = torch.Generator().manual_seed(seed)
    z = torch.randn((B, C0, 8, 8), generator=gen)
    cond = torch.randn((B, h_channels, 8, 8), generator=gen) * 0.5
    x0 = torch.rand((B, 3, spatial, spatial), generator=gen) * 2.0 - 1.0
    return z, cond, x0
