"""Drop-in for the frozen conditioning encoders of the second stage: `ConvEncoder`
(models/modules/autoencoders/fully_conv_models.py:28-94) as used through `FirstStageWrapper.encoder` by
`PokeMotionModel.make_flow_input` (models/second_stage_video.py:268-287): `poke_embedder.encoder(poke)` and
`conditioner.encoder(X[:, 0])`, both deterministic (`config/poke_encoder.yaml:62`, `config/img_encoder.yaml:56`).
Same constructor signature, state-dict layout and return tuple; all arithmetic in libipoke_b200.so."""
import ctypes

import torch
from torch import nn

from . import _lib, spec
from .flow import _Holder, build_param_tree


class _NativeCencPlan:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().ipk_cenc_destroy(self.handle)
        except Exception:
            pass
        self.handle = None


class ConvEncoder(nn.Module):
    """ConvEncoder(nf_in, nf_max, n_stages, variational=False): x [B,nf_in,S,S] with S = 8 << n_stages -> (out, mean, None),
    out = bottleneck(model(x)), mean = model(x) (fully_conv_models.py:74-88).  variational=True (sampling the embedding) is
    not used by the second stage and is rejected."""

    def __init__(self, nf_in, nf_max, n_stages, variational=False, norm_layer="group", layers=None, spectral_norm=True,
                 min_spatial_size=8, ipk_max_batch=64):
        super().__init__()
        if variational or layers is not None or norm_layer != "group" or not spectral_norm:
            raise NotImplementedError("ipoke_b200 ConvEncoder: only the deterministic, group-norm, spectral-norm configuration "
                                      "of the shipped poke / image encoders is implemented")
        self.variational = False
        self.nf_in, self.nf_max, self.n_stages = int(nf_in), int(nf_max), int(n_stages)
        self.min_spatial_size = int(min_spatial_size)
        self.spatial = self.min_spatial_size << self.n_stages
        widths = [32]
        for _ in range(self.n_stages - 1):
            widths.append(min(widths[-1] * 2, self.nf_max))
        self.depths = list(reversed(widths))                      # fully_conv_models.py:46,61
        self.nf_in_bn = widths[-1]
        tree = _Holder()
        build_param_tree(tree, spec.cond_encoder_param_spec(self.nf_in, self.nf_max, self.n_stages))
        for k, m in tree._modules.items():
            self.add_module(k, m)
        self.max_batch = int(ipk_max_batch)
        self._plan = None
        self._plan_key = None
        self._plist = None

    def invalidate(self):
        self._plan = None
        self._plan_key = None
        self._plist = None

    def _ensure_plan(self, device, batch):
        if batch > self.max_batch:
            self.max_batch = int(batch)
            self.invalidate()
        if self._plist is None:
            self._plist = list(self.parameters()) + list(self.buffers())      # buffers: the spectral-norm u / v vectors
        key = (device, self.max_batch) + _lib.tensors_key(self._plist)
        if self._plan is not None and self._plan_key == key:
            return self._plan
        if device.type != "cuda":
            raise RuntimeError("ipoke_b200 ConvEncoder runs on CUDA tensors only (no CPU fallback)")
        L = _lib.lib()
        c = _lib.CencConfig()
        c.nf_in, c.nf_max, c.spatial, c.min_spatial_size = self.nf_in, self.nf_max, self.spatial, self.min_spatial_size
        c.n_stages, c.max_batch = self.n_stages, self.max_batch
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.ipk_cenc_create(ctypes.byref(c), ctypes.byref(h)), "ipk_cenc_create")
            plan = _NativeCencPlan(h)
            keep = []
            for k, v in self.state_dict().items():
                t = v.detach().float().contiguous()
                keep.append(t)
                _lib.check(L.ipk_cenc_set_tensor(h, k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(), _lib.DT_F32),
                           f"ipk_cenc_set_tensor({k})")
            _lib.check(L.ipk_cenc_finalize(h, _lib.current_stream_ptr()), "ipk_cenc_finalize")
        self._plan, self._plan_key = plan, key
        return plan

    def forward(self, input, sample_prior=False):
        if sample_prior:
            raise NotImplementedError("sample_prior needs the variational head")
        if input.dim() != 4 or tuple(input.shape[1:]) != (self.nf_in, self.spatial, self.spatial):
            raise ValueError(f"ConvEncoder input must be [B,{self.nf_in},{self.spatial},{self.spatial}], got {tuple(input.shape)}")
        x = input.detach().float().contiguous()
        B = x.shape[0]
        plan = self._ensure_plan(x.device, B)
        out = torch.empty((B, self.nf_max, self.min_spatial_size, self.min_spatial_size), device=x.device, dtype=torch.float32)
        mean = torch.empty((B, self.nf_in_bn, self.min_spatial_size, self.min_spatial_size), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ipk_cenc_forward(plan.handle, x.data_ptr(), out.data_ptr(), mean.data_ptr(), B,
                                                   _lib.current_stream_ptr()), "ipk_cenc_forward")
        return out, mean, None

    def _load_from_state_dict(self, *a, **k):
        self.invalidate()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)


def make_cond(conditioner_encoder, poke_encoder, x0, poke):
    """The conditioning half of PokeMotionModel.make_flow_input (second_stage_video.py:273-287,311):
    cond = cat([conditioner.encoder(X[:, 0])[0], poke_embedder.encoder(poke)[0]], dim=1)  ->  [B, 2*nf_max, 8, 8]."""
    with torch.no_grad():
        poke_emb, *_ = poke_encoder(poke)
        cond, *_ = conditioner_encoder(x0)
    return torch.cat([cond, poke_emb], dim=1)
