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
        if float(config.get("p_dropout", 0.0)) != 0.0:
            # the reference applies nn.Dropout inside MCFBlock / NICEConvBlock when p_dropout > 0 (macow_utils.py:284,425); no shipped
            # config sets it, and silently training without dropout would be a different model
            raise NotImplementedError("ipoke_b200 flow: p_dropout > 0 is not used by any shipped config and is not implemented")
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
        self._init_state = None          # None = unknown, else "all" / "none" / "wn_only" / "partial" (the `initialized` buffers)
        self._autograd_trainer = None    # FlowTrainer behind the autograd-transparent density direction

    # ------------------------------------------------------------------ native plan
    def _state_key(self):
        # version counters of all parameters (an in-place update such as an optimizer step re-packs the native plan) plus their storage
        # addresses (a `.data` swap does not bump the version counter).  Buffers (shuffle indices) only change through load_state_dict /
        # _apply, which invalidate explicitly.  The flat list is cached -- walking 7 000 parameters through the module tree costs
        # milliseconds per call.
        pl = self._plist
        if pl is None:
            pl = self._plist = list(self.parameters())
        return (pl[0].device, self.precision, self.max_batch) + _lib.tensors_key(pl)

    def invalidate(self, flags=False):
        """Drop the packed native plan (it is rebuilt from the current parameters on the next call).  flags=True also forgets the cached
        view of the `initialized` buffers (after load_state_dict / .to())."""
        self._plan = None
        self._plan_key = None
        self._plist = None
        if flags:
            self._init_state = None

    # ------------------------------------------------------------------ data-dependent initialisation
    def _init_flags(self):
        return [(k, b) for k, b in self.named_buffers() if k.endswith("initialized")]

    def _initialization(self):
        """'all' / 'none' / 'wn_only' (weight-norm layers initialised, ActNorms not) / 'partial'; one device reduction, cached until the
        next load_state_dict / .to() / invalidate()."""
        if self._init_state is None:
            flags = self._init_flags()
            # ActNorm2dFlow flags are `...actnorm*.initialized`; Conv2dWeightNorm flags `...net.conv1x1.initialized` / `...net.conv3.initialized`
            act = torch.stack([b.reshape(()) for k, b in flags if "actnorm" in k]).bool()
            wn = torch.stack([b.reshape(()) for k, b in flags if "actnorm" not in k]).bool()
            a_all, a_any, w_all, w_any = (bool(v) for v in torch.stack([act.all(), act.any(), wn.all(), wn.any()]).tolist())
            if a_all and w_all:
                self._init_state = "all"
            elif not a_any and not w_any:
                self._init_state = "none"
            elif w_all and not a_any:
                self._init_state = "wn_only"
            else:
                self._init_state = "partial"
        return self._init_state

    @torch.no_grad()
    def _prepare_initialization(self, reverse):
        """The reference initialises lazily inside forward(): ActNorm2dFlow only in the density direction (macow2.py:503-505),
        Conv2dWeightNorm on its first call in either direction, training or eval (macow_utils.py:248-250).  Returns True when the
        density-direction init pass (native, needs the batch) still has to run while the plan is built."""
        st = self._initialization()
        if st == "all":
            return False
        if st == "partial":
            raise RuntimeError("ipoke_b200 flow: the checkpoint is partially initialised (some ActNorm / WeightNorm `initialized` buffers are 0, "
                               "others 1); the data-dependent init pass is supported for a fresh model (all 0) only")
        if reverse:
            if st == "none":
                # sampling through a never-initialised flow: every weight-normed conv is zero_init (macow_utils.py:281,423), so its init
                # sets weight_g = init_scale / (std + 1e-6) = 0 and bias = -mean * 0 = 0; ActNorms keep their random log_scale
                for k, p in self.named_parameters():
                    if k.endswith(".conv.weight_g") or k.endswith(".conv.bias"):
                        p.zero_()
                for k, b in self._init_flags():
                    if "actnorm" not in k:
                        b.fill_(1)
                self._init_state = "wn_only"
            return False
        return True

    def _ensure_plan(self, device, batch, init_input=None, reverse=True):
        if device.type != "cuda":
            raise RuntimeError("ipoke_b200 flow runs on CUDA tensors only (no CPU fallback); move the module and inputs to a B200")
        if batch > self.max_batch:
            self.max_batch = int(batch)
            self.invalidate()
        need_init = self._prepare_initialization(reverse)
        key = self._state_key()
        if self._plan is not None and self._plan_key == key and not need_init:
            return self._plan
        L = _lib.lib()
        sd = self.state_dict()
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
            if need_init:
                # data-dependent init (macow2.py:526-539, macow_utils.py:231-246): the native pass writes log_scale / bias / weight_g
                # straight into the parameters registered above, then the plan is finalised from the initialised values
                if init_input is None:
                    raise RuntimeError("ipoke_b200 flow: the model is uninitialised and needs a density-direction batch for its data-dependent init")
                _lib.check(L.ipk_flow_data_init(h, init_input.data_ptr(), init_input.shape[0], _lib.current_stream_ptr()), "ipk_flow_data_init")
                with torch.no_grad():
                    for _, b in self._init_flags():
                        b.fill_(1)
                self._init_state = "all"
            _lib.check(L.ipk_flow_finalize(h, _lib.current_stream_ptr()), "ipk_flow_finalize")
        self._plan, self._plan_key = plan, key
        return plan

    def data_init(self, input, cond=None):
        """Run the data-dependent initialisation explicitly on a batch (what the reference's first training forward does)."""
        x = input.detach().float().contiguous()
        if self._prepare_initialization(False):
            self._plan = None
            self._ensure_plan(x.device, x.shape[0], init_input=x, reverse=False)
            if self._autograd_trainer is not None:
                self._autograd_trainer.invalidate_plan()

    # ------------------------------------------------------------------ reference API
    def forward(self, input, cond, reverse=False):
        if reverse:
            return self.reverse(input, cond)
        x, cond = self._check(input, cond)
        if torch.is_grad_enabled() and self._wants_grad(input):
            return self._forward_autograd(x, cond, input)
        plan = self._ensure_plan(x.device, x.shape[0], init_input=x, reverse=False)
        out = torch.empty_like(x)
        logdet = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ipk_flow_forward(plan.handle, x.data_ptr(), cond.data_ptr(), out.data_ptr(), logdet.data_ptr(),
                                                   x.shape[0], _lib.current_stream_ptr()), "ipk_flow_forward")
        return out, logdet

    # ------------------------------------------------------------------ autograd-transparent density direction
    def _wants_grad(self, input):
        # a graph is recorded when the input asks for a gradient, or in train() mode with trainable parameters (Lightning's
        # training_step); eval()-mode calls without an input gradient stay on the inference plan even outside torch.no_grad()
        if input.requires_grad:
            return True
        if not self.training:
            return False
        if self._plist is None:
            self._plist = list(self.parameters())
        return any(p.requires_grad for p in self._plist)

    def _forward_autograd(self, x, cond, input):
        """`out, logdet = self.flow(x, cond)` with a grad_fn: `loss.backward()` fills `p.grad` of every flow parameter (and of the input),
        so training_step / configure_optimizers of the reference (second_stage_video.py:409-415, 633-650) run unmodified.  The native
        training plan behind it keeps the parameters in one flat fp32 buffer (see train.FlowTrainer)."""
        from .train import FlowTrainer, FlowDensityFunction
        if self._prepare_initialization(False):
            self.data_init(x)
        tr = self._autograd_trainer
        if tr is None or tr.max_batch < x.shape[0]:
            tr = self._autograd_trainer = FlowTrainer(self, max_batch=max(x.shape[0], 1), precision=self.precision, distributed=False)
        params = [p for p in tr.param_list]
        return FlowDensityFunction.apply(tr, x, cond, input, *params)

    def reverse(self, out, cond):
        z, cond = self._check(out, cond)
        plan = self._ensure_plan(z.device, z.shape[0], reverse=True)
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
        self.invalidate(flags=True)
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate(flags=True)
        return super()._apply(fn, *a, **k)


def flow_nll(z, logdet):
    """FlowLoss.forward (models/modules/INN/loss.py:13-31), logdet_weight = 1, spatial_mean = False."""
    return (0.5 * (z ** 2).flatten(1).sum(dim=1)).mean() - logdet.mean()


def _nll(sample, spatial_mean=False):
    # models/modules/INN/loss.py:75-79
    if spatial_mean:
        return 0.5 * torch.sum(torch.mean(torch.pow(sample, 2), dim=[2, 3]), dim=1)
    return 0.5 * torch.sum(torch.pow(sample, 2), dim=[1, 2, 3])


class FlowLoss(nn.Module):
    """models/modules/INN/loss.py:6-31, same constructor, same `(loss, log)` return: the log dict carries `flow_loss`,
    `reference_nll_loss` (the NLL of a fresh `torch.randn_like(sample)` draw -- it consumes the RNG of the sample's device exactly like
    the reference), `nlogdet_loss`, `nll_loss` and `logdet_weight`.  Plain tensor arithmetic on the (tiny) outputs of the native flow."""

    def __init__(self, spatial_mean=False, logdet_weight=1.):
        super().__init__()
        self.spatial_mean = spatial_mean
        self.logdet_weight = logdet_weight

    def forward(self, sample, logdet):
        nll_loss = torch.mean(_nll(sample, spatial_mean=self.spatial_mean))
        assert len(logdet.shape) == 1
        if self.spatial_mean:
            h, w = sample.shape[-2:]
            nlogdet_loss = -torch.mean(logdet) / (h * w)
        else:
            nlogdet_loss = -torch.mean(logdet)
        loss = nll_loss + self.logdet_weight * nlogdet_loss
        reference_nll_loss = torch.mean(_nll(torch.randn_like(sample), spatial_mean=self.spatial_mean))
        log = {"flow_loss": loss, "reference_nll_loss": reference_nll_loss, "nlogdet_loss": nlogdet_loss, "nll_loss": nll_loss,
               "logdet_weight": self.logdet_weight}
        return loss, log
