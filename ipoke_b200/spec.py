"""Parameter inventories of the reference modules on the hot path: names, shapes, dtypes and default initialisers,
exactly as they appear in a reference checkpoint, so `load_state_dict` of a reference
state-dict works unchanged on the drop-in modules.

flow:        SupervisedMacowTransformer            models/modules/INN/INN.py:446-467, macow2.py, macow_utils.py
first stage: SpadeCondMotionModel.{rnn,gen,motion_bias}   models/first_stage_motion_model.py:482-496
"""
import math

import torch

DEFAULT_NUM_STEPS = [10, 5, 5, 4, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1, 1]   # config/second_stage.yaml:58-68


def flow_levels(cfg):
    """MultiScaleInternal.__init__ channel bookkeeping (macow2.py:825-871)."""
    C, factor = cfg["flow_in_channels"], cfg["factor"]
    if not len(cfg["num_steps"]) < factor:
        raise AssertionError("num_layers < factor (macow2.py:834)")
    step = C // factor
    out = []
    for L, n in enumerate(cfg["num_steps"]):
        po = C // factor
        out.append(dict(level=L, C=C, steps=n, prior_factor=factor, prior_out=po, z1=C - po))
        C -= step
        assert C == out[-1]["z1"]
        factor -= 1
    return out, C


def flow_param_spec(cfg):
    """Yield (name, shape, dtype, is_buffer, init) for every tensor of the flow's state-dict."""
    Hd, hc = cfg["flow_mid_channels"], cfg["h_channels"]
    kH, kW = cfg["kernel_size"]
    f32, u8, i64 = torch.float32, torch.uint8, torch.int64

    def actnorm(p, C):
        yield p + "log_scale", (C, 1, 1), f32, False, ("normal", 0.05)      # macow2.py:486-488
        yield p + "bias", (C, 1, 1), f32, False, ("zeros",)
        yield p + "initialized", (), u8, True, ("zeros",)

    def shuffle(p, C):
        yield p + "forward_shuffle_idx", (C,), i64, True, ("perm",)          # flow_blocks.py:317-320
        yield p + "backward_shuffle_idx", (C,), i64, True, ("argsort_prev",)

    def wn(p, cout, cin, kh, kw):
        yield p + "initialized", (), u8, True, ("zeros",)
        yield p + "conv.bias", (cout,), f32, False, ("zeros",)
        yield p + "conv.weight_g", (cout, 1, 1, 1), f32, False, ("wn_g",)
        yield p + "conv.weight_v", (cout, cin, kh, kw), f32, False, ("normal", 0.05)   # macow_utils.py:222

    def mcf(p, C, kh, kw):
        hid = 4 * C if C <= 96 else min(2 * C, 512)
        yield p + "net.shift_conv.weight", (hid, C, kh, kw), f32, False, ("conv", C * kh * kw)
        yield from wn(p + "net.conv1x1.", 2 * C, hid + hc, 1, 1)

    def nice(p, C, factor):
        cout = C // factor
        cin = C - cout
        yield p + "net.conv1.weight", (Hd, cin, 3, 3), f32, False, ("conv", cin * 9)
        yield p + "net.conv2.weight", (Hd, Hd, 1, 1), f32, False, ("conv", Hd)
        yield from wn(p + "net.conv3.", 2 * cout, Hd, 3, 3)

    def unit(p, C):
        yield from mcf(p + "conv1.", C, kH, kW)
        yield from mcf(p + "conv2.", C, kH, kW)
        yield from actnorm(p + "actnorm1.", C)
        yield from mcf(p + "conv3.", C, kW, kH)
        yield from mcf(p + "conv4.", C, kW, kH)
        yield from actnorm(p + "actnorm2.", C)

    levels, _ = flow_levels(cfg)
    for lv in levels:
        L, C = lv["level"], lv["C"]
        for s in range(lv["steps"]):
            p = f"flow.layers.{L}.{s}."
            yield from actnorm(p + "actnorm1.", C)
            yield from shuffle(p + "conv1x1.", C)
            yield from unit(p + "units1.0.", C)
            yield from unit(p + "units1.1.", C)
            yield from nice(p + "coupling1_up.", C, 2)
            yield from nice(p + "coupling1_dn.", C, 2)
            yield from actnorm(p + "actnorm2.", C)
            yield from unit(p + "units2.0.", C)
            yield from unit(p + "units2.1.", C)
            yield from nice(p + "coupling2_up.", C, 2)
            yield from nice(p + "coupling2_dn.", C, 2)
    for lv in levels:
        L, C = lv["level"], lv["C"]
        p = f"flow.priors.{L}."
        yield from shuffle(p + "conv1x1.", C)
        yield from nice(p + "coupling.", C, lv["prior_factor"])
        yield from actnorm(p + "actnorm.", lv["prior_out"])
    for lv in levels:
        yield from shuffle(f"flow.shuffle_layers.{lv['level']}.", lv["C"])


def first_stage_param_spec(cfg):
    """(name, shape, dtype, is_buffer, init) for rnn.*, motion_bias and gen.* of SpadeCondMotionModel."""
    z, dec = cfg["z_dim"], cfg["dec_channels"]
    f32 = torch.float32
    snorm = cfg.get("spectral_norm", True)

    def conv(p, cout, cin, sn=False, transposed=False):
        shape = (cin, cout, 3, 3) if transposed else (cout, cin, 3, 3)
        fan_in = (cout if transposed else cin) * 9
        if sn:
            yield p + "bias", (cout,), f32, False, ("conv_bias", fan_in)
            yield p + "weight_orig", shape, f32, False, ("conv", fan_in)
            yield p + "weight_u", (cout,), f32, True, ("unit",)
            yield p + "weight_v", (cin * 9,), f32, True, ("unit",)
        else:
            yield p + "weight", shape, f32, False, ("conv", fan_in)
            yield p + "bias", (cout,), f32, False, ("conv_bias", fan_in)

    for i in range(cfg["n_gru_layers"]):
        for gate in ("reset_gate", "update_gate", "out_gate"):
            yield f"rnn.cells.{i}.{gate}.weight", (z, 2 * z, 3, 3), f32, False, ("conv", 2 * z * 9)   # orthogonal in the reference
            yield f"rnn.cells.{i}.{gate}.bias", (z,), f32, False, ("zeros",)
    ms = cfg.get("min_spatial_size", 8)
    yield "motion_bias", (1, z, ms, ms), f32, False, ("normal", 1.0)
    p = "gen.in_block."
    yield from conv(p + "conv1.conv.", dec[0], z, snorm)
    yield p + "conv1.norm.weight", (dec[0],), f32, False, ("ones",)
    yield p + "conv1.norm.bias", (dec[0],), f32, False, ("zeros",)
    yield from conv(p + "conv2.conv.", dec[0], dec[0], snorm)
    yield p + "conv2.norm.weight", (dec[0],), f32, False, ("ones",)
    yield p + "conv2.norm.bias", (dec[0],), f32, False, ("zeros",)
    if z != dec[0]:
        yield from conv(p + "res_conv.conv.", dec[0], z, snorm)
    for i, nf in enumerate(dec[1:]):
        p = f"gen.blocks.{i}."
        yield from conv(p + "conv1.conv.", nf, dec[i], snorm, transposed=True)
        yield from conv(p + "conv2.conv.", nf, nf, snorm)
        yield from conv(p + "res_conv.conv.", nf, dec[i], snorm, transposed=True)
    for i, nf in enumerate(dec[1:]):
        p = f"gen.spade_blocks.{i}."
        yield from conv(p + "conv.", 128, 3)
        yield from conv(p + "conv_gamma.", nf, 128)
        yield from conv(p + "conv_beta.", nf, 128)
    yield from conv("gen.out_conv.conv.", 3, dec[-1])


def encoder_stages(cfg):
    """Layer plan of ResNetMotionEncoder.__init__ (motion_encoder.py:161-190): [(name, inplanes, planes, stride)]."""
    ch = list(cfg["ENC_M_channels"])
    first_down = (len(ch) - 1 < int(math.ceil(math.log2(cfg["max_frames"])))) or cfg["full_seq"]
    st = [("layer1", ch[0], ch[1], (2, 1, 1) if first_down else (1, 1, 1)), ("layer2", ch[1], ch[2], (2, 2, 2)),
          ("layer3", ch[2], ch[3], (2, 2, 2))]
    stride4 = (2, 1, 1) if (cfg["full_seq"] and cfg["max_frames"] >= 16) else None
    if cfg["img_size"] // 8 > cfg.get("min_spatial_size", 8):
        stride4 = (2, 2, 2)
    if stride4 is not None:
        if len(ch) < 5:
            ch.append(ch[-1])
        st.append(("layer4", ch[3], ch[4], stride4))
    if cfg["img_size"] // 16 > cfg.get("min_spatial_size", 8):
        st.append(("layer5", ch[4], ch[5], (2, 2, 2)))
    return st


def encoder_param_spec(cfg):
    """(name, shape, dtype, is_buffer, init) for every tensor of ResNetMotionEncoder's state-dict."""
    f32 = torch.float32
    ch0 = cfg["ENC_M_channels"][0]

    def gn(p, c):
        yield p + "weight", (c,), f32, False, ("ones",)
        yield p + "bias", (c,), f32, False, ("zeros",)

    yield "conv1.weight", (ch0, 3, 3, 7, 7), f32, False, ("normal", math.sqrt(2.0 / (ch0 * 147)))      # kaiming_normal_, fan_out
    yield from gn("bn1.", ch0)
    stages = encoder_stages(cfg)
    for name, inplanes, planes, stride in stages:
        inp = inplanes
        for b in range(2):
            p = f"{name}.{b}."
            yield p + "conv1.weight", (planes, inp, 3, 3, 3), f32, False, ("normal", math.sqrt(2.0 / (planes * 27)))
            yield from gn(p + "bn1.", planes)
            yield p + "conv2.weight", (planes, planes, 3, 3, 3), f32, False, ("normal", math.sqrt(2.0 / (planes * 27)))
            yield from gn(p + "bn2.", planes)
            if b == 0 and (stride != (1, 1, 1) or inp != planes):
                yield p + "downsample.0.weight", (planes, inp, 1, 1, 1), f32, False, ("normal", math.sqrt(2.0 / planes))
                yield from gn(p + "downsample.1.", planes)
            inp = planes
    last = stages[-1][2]
    for head in ("conv_mu", "conv_var"):
        yield head + ".weight", (cfg["z_dim"], last, 3, 3), f32, False, ("conv", last * 9)
        yield head + ".bias", (cfg["z_dim"],), f32, False, ("conv_bias", last * 9)


def cond_encoder_param_spec(nf_in, nf_max, n_stages):
    """(name, shape, dtype, is_buffer, init) for ConvEncoder(nf_in, nf_max, n_stages, variational=False)
    (models/modules/autoencoders/fully_conv_models.py:28-72): spectral-normed stride-2 blocks + a plain bottleneck ResBlock."""
    f32 = torch.float32
    widths = [32]
    for _ in range(n_stages - 1):
        widths.append(min(widths[-1] * 2, nf_max))

    def gn(p, c):
        yield p + "weight", (c,), f32, False, ("ones",)
        yield p + "bias", (c,), f32, False, ("zeros",)

    def conv(p, cout, cin, sn):
        if sn:
            yield p + "bias", (cout,), f32, False, ("conv_bias", cin * 9)
            yield p + "weight_orig", (cout, cin, 3, 3), f32, False, ("conv", cin * 9)
            yield p + "weight_u", (cout,), f32, True, ("unit",)
            yield p + "weight_v", (cin * 9,), f32, True, ("unit",)
        else:
            yield p + "weight", (cout, cin, 3, 3), f32, False, ("conv", cin * 9)
            yield p + "bias", (cout,), f32, False, ("conv_bias", cin * 9)

    yield from gn("model.0.norm.", widths[0])
    yield from conv("model.0.conv.", widths[0], nf_in, True)
    for i in range(1, len(widths)):
        p = f"model.{i}."
        yield from gn(p + "conv1.norm.", widths[i])
        yield from conv(p + "conv1.conv.", widths[i], widths[i - 1], True)
        yield from gn(p + "conv2.norm.", widths[i])
        yield from conv(p + "conv2.conv.", widths[i], widths[i], True)
        yield from conv(p + "res_conv.conv.", widths[i], widths[i - 1], True)
    p = "bottleneck.0."
    yield from gn(p + "conv1.norm.", nf_max)
    yield from conv(p + "conv1.conv.", nf_max, widths[-1], False)
    yield from gn(p + "conv2.norm.", nf_max)
    yield from conv(p + "conv2.conv.", nf_max, nf_max, False)
    if widths[-1] != nf_max:
        yield from conv(p + "res_conv.conv.", nf_max, widths[-1], False)


def init_tensor(shape, dtype, init, prev=None):
    kind = init[0]
    if kind == "zeros":
        return torch.zeros(shape, dtype=dtype)
    if kind == "ones":
        return torch.ones(shape, dtype=dtype)
    if kind == "normal":
        return torch.randn(shape, dtype=dtype) * init[1]
    if kind == "conv":
        b = 1.0 / math.sqrt(init[1])
        return (torch.rand(shape, dtype=dtype) * 2 - 1) * b
    if kind == "conv_bias":
        b = 1.0 / math.sqrt(init[1])
        return (torch.rand(shape, dtype=dtype) * 2 - 1) * b
    if kind == "wn_g":
        return torch.zeros(shape, dtype=dtype)         # zero_init couplings: identity at initialisation (macow_utils.py:281)
    if kind == "perm":
        return torch.randperm(shape[0])
    if kind == "argsort_prev":
        return torch.argsort(prev)
    if kind == "unit":
        return torch.nn.functional.normalize(torch.randn(shape, dtype=dtype), dim=0, eps=1e-12)
    raise ValueError(kind)
