"""Drop-in for the reference's conditional MaCow flow.

`SupervisedMacowTransformer(config)` keeps the reference's constructor config keys, call signatures and state-dict
layout (models/modules/INN/INN.py:446-481): `forward(input, cond, reverse=False) -> (out, logdet)`,
`forward(..., reverse=True) -> out`, `reverse(out, cond)`, `sample(shape, cond, device)`; `self.flow.reshape == 'none'`
(read at models/second_stage_video.py:290).  All arithmetic runs in libipoke_b200.so on the tensors' CUDA device; there
is no PyTorch / CPU fallback.
"""
import ctypes

import torch
from torch import nn

from . import _lib, spec


class _Holder(nn.Module):
    """Parameter container that reproduces the reference's module hierarchy (names only; no forward)."""

    def extra_repr(self):
        return "ipoke_b200 parameter holder"


def build_param_tree(root, entries):
    """Create nested holders so that root.state_dict() has exactly the given dotted names."""
    prev = None
    for name, shape, dtype, is_buffer, init in entries:
        parts = name.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Holder())
            mod = mod._modules[p]
        t = spec.init_tensor(shape, dtype, init, prev)
        prev = t
        if is_buffer:
            mod.register_buffer(parts[-1], t)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=dtype.is_floating_point))


class _NativeFlowPlan:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().ipk_flow_destroy(self.handle)
        except Exception:
            pass
        self.handle = None


class SupervisedMacowTransformer(nn.Module):
    """config: the reference's config['architecture'] dict (flow_in_channels, flow_mid_channels, h_channels, num_steps,
    factor, transform, prior_transform, kernel_size, coupling_type, activation, flow_attn_heads,
    cond_conv_hidden_channels [, condition_nice, attention, cond_conv, p_dropout]).  Extra keys understood by this
    implementation: ipk_precision ('fp32' | 'bf16' | 'fp32_simt', default 'fp32'), ipk_max_batch (default 64)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        # same mandatory-key behaviour as the reference (INN.py:451-457)
        _ = config["flow_attn_heads"]
        _ = config["cond_conv_hidden_channels"]
        for key, want in (("transform", "affine"), ("prior_transform", "affine"), ("coupling_type", "conv"), ("activation", "elu")):
            if config[key] != want:
                raise NotImplementedError(f"ipoke_b200 flow: {key}={config[key]!r} is not used by any shipped config (only {want!r})")
        for key in ("condition_nice", "attention", "cond_conv", "use1x1"):
            if config.get(key, False):
                raise NotImplementedError(f"ipoke_b200 flow: {key}=True is not used by any shipped config")
        if float(config.get("p_dropout", 0.0)) != 0.0 and False:
            raise NotImplementedError
        self._cfg = dict(flow_in_channels=int(config["flow_in_channels"]), flow_mid_channels=int(config["flow_mid_channels"]),
                         h_channels=int(config["h_channels"]), num_steps=[int(s) for s in config["num_steps"]],
                         factor=int(config["factor"]), kernel_size=[int(k) for k in config["kernel_size"]])
        self.flow = _Holder()
        tree = _Holder()
        build_param_tree(tree, spec.flow_param_spec(self._cfg))
        self.flow = tree._modules["flow"]
        self.flow.reshape = "none"                                  # macow2.py:831
        _, self.flow.z_channels = spec.flow_levels(self._cfg)       # macow2.py:870
        self.precision = config.get("ipk_precision", "fp32")
        self.max_batch = int(config.get("ipk_max_batch", 64))
        self._plan = None
        self._plan_key = None
        self._plist = None

    # ------------------------------------------------------------------ native plan
    def _state_key(self):
        # version counters of all parameters: an in-place update (optimizer step) or .data swap re-packs the native plan.
        # The flat list is cached -- walking 7 000 parameters through the module tree costs milliseconds per call.
        pl = self._plist
        if pl is None:
            pl = self._plist = list(self.parameters())
        return (pl[0].device, self.precision, self.max_batch, sum(q._version for q in pl))

    def invalidate(self):
        self._plan = None
        self._plan_key = None
        self._plist = None

    def _ensure_plan(self, device, batch):
        if batch > self.max_batch:
            self.max_batch = int(batch)
            self.invalidate()
        key = self._state_key()
        if self._plan is not None and self._plan_key == key:
            return self._plan
        if device.type != "cuda":
            raise RuntimeError("ipoke_b200 flow runs on CUDA tensors only (no CPU fallback); move the module and inputs to a B200")
        L = _lib.lib()
        sd = self.state_dict()
        bad = [k for k, v in sd.items() if k.endswith("initialized") and int(v) == 0]
        if bad:
            raise RuntimeError(f"ipoke_b200 flow: {len(bad)} ActNorm/WeightNorm layers are uninitialised (e.g. {bad[0]}); the "
                               "data-dependent init pass (macow2.py:503-505, macow_utils.py:248-250) is a training-time step -- load "
                               "an initialised checkpoint")
        c = _lib.FlowConfig()
        c.flow_in_channels = self._cfg["flow_in_channels"]
        c.flow_mid_channels = self._cfg["flow_mid_channels"]
        c.h_channels = self._cfg["h_channels"]
        c.n_levels = len(self._cfg["num_steps"])
        for i, s in enumerate(self._cfg["num_steps"]):
            c.num_steps[i] = s
        c.factor = self._cfg["factor"]
        c.kernel_h, c.kernel_w = self._cfg["kernel_size"]
        c.precision = _lib.precision_code(self.precision)
        c.max_batch = self.max_batch
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.ipk_flow_create(ctypes.byref(c), ctypes.byref(h)), "ipk_flow_create")
            plan = _NativeFlowPlan(h)
            keep = []
            for k, v in sd.items():
                if k.endswith("initialized"):
                    continue
                t = v.detach().contiguous()
                keep.append(t)
                _lib.check(L.ipk_flow_set_tensor(h, k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(), _lib.dtype_code(t)),
                           f"ipk_flow_set_tensor({k})")
            _lib.check(L.ipk_flow_finalize(h, _lib.current_stream_ptr()), "ipk_flow_finalize")
        self._plan, self._plan_key = plan, key
        return plan

    # ------------------------------------------------------------------ reference API
    def forward(self, input, cond, reverse=False):
        if reverse:
            return self.reverse(input, cond)
        x, cond = self._check(input, cond)
        plan = self._ensure_plan(x.device, x.shape[0])
        out = torch.empty_like(x)
        logdet = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ipk_flow_forward(plan.handle, x.data_ptr(), cond.data_ptr(), out.data_ptr(), logdet.data_ptr(),
                                                   x.shape[0], _lib.current_stream_ptr()), "ipk_flow_forward")
        return out, logdet

    def reverse(self, out, cond):
        z, cond = self._check(out, cond)
        plan = self._ensure_plan(z.device, z.shape[0])
        x = torch.empty_like(z)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().ipk_flow_reverse(plan.handle, z.data_ptr(), cond.data_ptr(), x.data_ptr(), z.shape[0],
                                                   _lib.current_stream_ptr()), "ipk_flow_reverse")
        return x

    def sample(self, shape, cond, device="cpu"):
        z_tilde = torch.randn(shape).to(device)          # CPU generator, as the reference (INN.py:479)
        return self.reverse(z_tilde, cond)

    def _check(self, x, cond):
        assert cond is not None                           # macow2.py:110
        C0, hc = self._cfg["flow_in_channels"], self._cfg["h_channels"]
        if x.dim() != 4 or tuple(x.shape[1:]) != (C0, 8, 8):
            raise ValueError(f"flow input must be [B,{C0},8,8], got {tuple(x.shape)}")
        if cond.dim() != 4 or tuple(cond.shape) != (x.shape[0], hc, 8, 8):
            raise ValueError(f"flow conditioning must be [{x.shape[0]},{hc},8,8], got {tuple(cond.shape)}")
        if cond.device != x.device:
            raise ValueError("flow input and conditioning must live on the same device")
        return x.detach().float().contiguous(), cond.detach().float().contiguous()

    def _load_from_state_dict(self, *a, **k):
        self.invalidate()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)


def flow_nll(z, logdet):
    """FlowLoss.forward (models/modules/INN/loss.py:13-31), logdet_weight = 1, spatial_mean = False."""
    return (0.5 * (z ** 2).flatten(1).sum(dim=1)).mean() - logdet.mean()
