"""Drop-in for the reference's in-repo Frechet-video-distance chain (utils/metrics.py):

    I3D(num_classes, modality)          utils/metrics.py:999-1105   same constructor, state-dict keys and `(softmax, logits)` return
    preprocess(data_gen, data_orig)     :786-802                    bilinear 224x224 (align_corners) + [-1,1] -> [0,1] per set
    get_activations(data, model, ...)   :681-733                    logits of full batches only, as a numpy array
    calculate_activation_statistics     :743-770
    calculate_frechet_distance          :625-678                    host linear algebra on 400 x 400 matrices (numpy / scipy, like the reference)
    calculate_FVD(model, gen, orig, bs) :773-780

The network -- 57 TF-SAME Conv3d + BatchNorm + ReLU units, the TF-padded max-pools, the Inception concats, the head -- runs in
libipoke_b200.so on the tcgen05 Conv3d engine (csrc/i3d.cu); there is no PyTorch / CPU fallback for it.  It is validation-time tooling
(second_stage_video.py:558-576 runs it every validation epoch), not part of the sampling step.
"""
import ctypes

import numpy as np
import torch
from torch import nn

from . import _lib
from .flow import _Holder

# (name, in_channels, [b0, b1a, b1b, b2a, b2b, b3]) of the nine Inception blocks (utils/metrics.py:1047-1066)
_MIXED = [("mixed_3b", 192, [64, 96, 128, 16, 32, 32]), ("mixed_3c", 256, [128, 128, 192, 32, 96, 64]), ("mixed_4b", 480, [192, 96, 208, 16, 48, 64]),
          ("mixed_4c", 512, [160, 112, 224, 24, 64, 64]), ("mixed_4d", 512, [128, 128, 256, 24, 64, 64]), ("mixed_4e", 512, [112, 144, 288, 32, 64, 64]),
          ("mixed_4f", 528, [256, 160, 320, 32, 128, 128]), ("mixed_5b", 832, [256, 160, 320, 32, 128, 128]), ("mixed_5c", 832, [384, 192, 384, 48, 128, 128])]


def i3d_units(num_classes=400):
    """(state-dict prefix, cin, cout, kernel, use_bn, use_bias) of every Unit3Dpy in module order."""
    u = [("conv3d_1a_7x7", 3, 64, 7, True, False), ("conv3d_2b_1x1", 64, 64, 1, True, False), ("conv3d_2c_3x3", 64, 192, 3, True, False)]
    for name, cin, o in _MIXED:
        u += [(f"{name}.branch_0", cin, o[0], 1, True, False), (f"{name}.branch_1.0", cin, o[1], 1, True, False),
              (f"{name}.branch_1.1", o[1], o[2], 3, True, False), (f"{name}.branch_2.0", cin, o[3], 1, True, False),
              (f"{name}.branch_2.1", o[3], o[4], 3, True, False), (f"{name}.branch_3.1", cin, o[5], 1, True, False)]
    u.append(("conv3d_0c_1x1", 1024, num_classes, 1, False, True))
    return u


class _NativeI3dPlan:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().ipk_i3d_destroy(self.handle)
        except Exception:
            pass
        self.handle = None


class I3D(nn.Module):
    """I3D(num_classes, modality='rgb', dropout_prob=0, name='inception') -- utils/metrics.py:999-1077.  Parameters and buffers carry the
    reference's names (`<unit>.conv3d.weight`, `<unit>.batch3d.{weight,bias,running_mean,running_var,num_batches_tracked}`), all frozen
    (`requires_grad=False`, :1073-1074).  Extra keyword arguments: ipk_max_batch, ipk_max_frames, ipk_precision ('fp32' = bf16x3, 'bf16')."""

    def __init__(self, num_classes, modality="rgb", dropout_prob=0, name="inception", ipk_max_batch=50, ipk_max_frames=16, ipk_precision="fp32"):
        super().__init__()
        if modality != "rgb":
            raise NotImplementedError("ipoke_b200 I3D: only the 'rgb' modality is used by the reference's FVD (utils/metrics.py:805)")
        if dropout_prob != 0:
            raise NotImplementedError("ipoke_b200 I3D: inference only (dropout_prob = 0)")
        self.name, self.num_classes, self.modality = name, num_classes, modality
        for p, cin, cout, k, bn, bias in i3d_units(num_classes):
            mod = self
            for part in p.split("."):
                if part not in mod._modules:
                    mod.add_module(part, _Holder())
                mod = mod._modules[part]
            conv = _Holder()
            fan_in = cin * k ** 3
            conv.register_parameter("weight", nn.Parameter(torch.randn((cout, cin, k, k, k)) * (2.0 / fan_in) ** 0.5, requires_grad=False))
            if bias:
                conv.register_parameter("bias", nn.Parameter(torch.zeros(cout), requires_grad=False))
            mod.add_module("conv3d", conv)
            if bn:
                b = _Holder()
                b.register_parameter("weight", nn.Parameter(torch.ones(cout), requires_grad=False))
                b.register_parameter("bias", nn.Parameter(torch.zeros(cout), requires_grad=False))
                b.register_buffer("running_mean", torch.zeros(cout))
                b.register_buffer("running_var", torch.ones(cout))
                b.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
                mod.add_module("batch3d", b)
        self.max_batch, self.max_frames, self.precision = int(ipk_max_batch), int(ipk_max_frames), ipk_precision
        self._plan = None
        self._plan_key = None
        self._plist = None

    def invalidate(self):
        self._plan = None
        self._plan_key = None
        self._plist = None

    def _ensure_plan(self, device, batch, frames):
        if device.type != "cuda":
            raise RuntimeError("ipoke_b200 I3D runs on CUDA tensors only (no CPU fallback)")
        if batch > self.max_batch or frames > self.max_frames:
            self.max_batch, self.max_frames = max(self.max_batch, int(batch)), max(self.max_frames, int(frames))
            self.invalidate()
        if self._plist is None:
            self._plist = list(self.parameters()) + [b for b in self.buffers() if b.dtype.is_floating_point]
        key = (device, self.max_batch, self.max_frames, self.precision) + _lib.tensors_key(self._plist)
        if self._plan is not None and self._plan_key == key:
            return self._plan
        L = _lib.lib()
        c = _lib.I3dConfig()
        c.num_classes, c.max_batch, c.max_frames = self.num_classes, self.max_batch, self.max_frames
        c.precision = _lib.precision_code(self.precision)
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.ipk_i3d_create(ctypes.byref(c), ctypes.byref(h)), "ipk_i3d_create")
            plan = _NativeI3dPlan(h)
            keep = []
            for k, v in self.state_dict().items():
                if not v.dtype.is_floating_point:
                    continue
                t = v.detach().float().contiguous()
                keep.append(t)
                _lib.check(L.ipk_i3d_set_tensor(h, k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(), _lib.DT_F32), f"ipk_i3d_set_tensor({k})")
            _lib.check(L.ipk_i3d_finalize(h, _lib.current_stream_ptr()), "ipk_i3d_finalize")
        self._plan, self._plan_key = plan, key
        return plan

    def forward(self, inp):
        """inp: [B, 3, T, 224, 224] -> (softmax, logits), each [B, num_classes] (utils/metrics.py:1079-1105)."""
        if inp.dim() != 5 or inp.shape[1] != 3 or inp.shape[3] != 224 or inp.shape[4] != 224:
            raise ValueError(f"I3D input must be [B,3,T,224,224], got {tuple(inp.shape)}")
        x = inp.detach().float().contiguous()
        B, T = x.shape[0], x.shape[2]
        plan = self._ensure_plan(x.device, B, T)
        logits = torch.empty((B, self.num_classes), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ipk_i3d_forward(plan.handle, x.data_ptr(), logits.data_ptr(), B, T, _lib.current_stream_ptr()), "ipk_i3d_forward")
        return torch.softmax(logits, dim=1), logits

    def _load_from_state_dict(self, *a, **k):
        self.invalidate()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)


def _preprocess_one(videos):
    """[N, T, 3, H, W] CUDA tensor -> [N, T, 3, 224, 224] in [0, 1] when the set has negative values (utils/metrics.py:786-802)."""
    if not videos.is_cuda:
        raise RuntimeError("ipoke_b200 preprocess runs on CUDA tensors (no CPU fallback)")
    v = videos.detach().float().contiguous()
    N, T, C, H, W = v.shape
    if C != 3 or H != W:
        raise ValueError(f"preprocess expects [N,T,3,S,S] videos, got {tuple(v.shape)}")
    out = torch.empty((N, T, 3, 224, 224), device=v.device, dtype=torch.float32)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().ipk_i3d_preprocess(v.data_ptr(), out.data_ptr(), N * T, H, _lib.current_stream_ptr()), "ipk_i3d_preprocess")
    return out


def preprocess(data_gen, data_orig):
    return _preprocess_one(data_gen), _preprocess_one(data_orig)


def get_activations(data, model, batch_size=50, cuda=False, verbose=False):
    """utils/metrics.py:681-733: the remainder of an incomplete last batch is dropped; returns float64 [n_used, 400]."""
    model.eval()
    n_samples = data.size(0)
    if batch_size > n_samples:
        batch_size = n_samples
    n_batches = n_samples // batch_size
    pred_arr = np.empty((n_batches * batch_size, 400))
    dev = next(model.parameters()).device
    for i in range(n_batches):
        batch = data[i * batch_size:(i + 1) * batch_size].to(dev)
        with torch.no_grad():
            pred = model(batch.permute(0, 2, 1, 3, 4))[1]
        pred_arr[i * batch_size:(i + 1) * batch_size] = pred.cpu().data.numpy().reshape(batch_size, -1)
    return pred_arr


def calculate_activation_statistics(data, model, batch_size=50, cuda=True, verbose=False):
    act = get_activations(data, model, batch_size, cuda, verbose)
    act = act[np.flatnonzero(np.logical_not(np.isnan(act)).any(axis=-1))]
    return np.mean(act, axis=0), np.cov(act, rowvar=False)


def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """utils/metrics.py:625-678: |mu1 - mu2|^2 + Tr(S1 + S2 - 2 sqrt(S1 S2)); 400 x 400 host linear algebra."""
    from scipy import linalg
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape, "Training and test mean vectors have different lengths"
    assert sigma1.shape == sigma2.shape, "Training and test covariances have different dimensions"
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if isinstance(covmean, tuple):          # SciPy < 1.18 with disp=False semantics
        covmean = covmean[0]
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean)


def calculate_FVD(model, data_gen, data_orig, batch_size, cuda=True):
    data_gen, data_orig = preprocess(data_gen, data_orig)
    m1, s1 = calculate_activation_statistics(data_gen, model, batch_size, cuda)
    m2, s2 = calculate_activation_statistics(data_orig, model, batch_size, cuda)
    return calculate_frechet_distance(m1, s1, m2, s2)
