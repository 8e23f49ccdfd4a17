"""Drop-in for the first stage's 3-D conv video encoder `SpadeCondMotionModel.enc_motion`
(`resnet18_alternative`, models/modules/motion_models/motion_encoder.py:150-241), used on the second-stage TRAINING path:
`PokeMotionModel.encode_first_stage` (models/second_stage_video.py:352-359) calls `enc_motion(X.transpose(1, 2))` under
no_grad and takes `(motion, mu, cov)`.  Same constructor dict keys, state-dict layout, call signature and the
`be_determinstic` [sic] attribute read at second_stage_video.py:198; all arithmetic in libipoke_b200.so (no CPU fallback)."""
import ctypes

import torch
from torch import nn

from . import _lib, spec
from .flow import _Holder, build_param_tree


class _NativeEncPlan:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().ipk_enc_destroy(self.handle)
        except Exception:
            pass
        self.handle = None


class ResNetMotionEncoder(nn.Module):
    """dic: the first stage's config['architecture'] (+ img_size, max_frames, full_seq as SpadeCondMotionModel sets them,
    first_stage_motion_model.py:478-480): z_dim, ENC_M_channels, img_size, max_frames, full_seq [, min_spatial_size,
    deterministic].  Extra keys: ipk_max_batch, ipk_precision ("fp32" = bf16x3 tcgen05 Conv3d (default), "bf16", "fp32_simt")."""

    def __init__(self, dic):
        super().__init__()
        self.be_determinstic = bool(dic.get("deterministic", False))
        self._cfg = dict(z_dim=int(dic["z_dim"]), img_size=int(dic["img_size"]), max_frames=int(dic["max_frames"]),
                         full_seq=bool(dic["full_seq"]), ENC_M_channels=[int(c) for c in dic["ENC_M_channels"]],
                         min_spatial_size=int(dic.get("min_spatial_size", 8)))
        self.spatial_size = self._cfg["img_size"]
        tree = _Holder()
        build_param_tree(tree, spec.encoder_param_spec(self._cfg))
        for k, m in tree._modules.items():
            self.add_module(k, m)
        self.max_batch = int(dic.get("ipk_max_batch", 32))
        self.precision = dic.get("ipk_precision", "fp32")
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
            self._plist = list(self.parameters()) + list(self.buffers())
        key = (device, self.max_batch, self.precision) + _lib.tensors_key(self._plist)
        if self._plan is not None and self._plan_key == key:
            return self._plan
        if device.type != "cuda":
            raise RuntimeError("ipoke_b200 encoder runs on CUDA tensors only (no CPU fallback)")
        L = _lib.lib()
        c = _lib.EncConfig()
        c.z_dim, c.img_size, c.max_frames = self._cfg["z_dim"], self._cfg["img_size"], self._cfg["max_frames"]
        c.full_seq = 1 if self._cfg["full_seq"] else 0
        c.n_channels = len(self._cfg["ENC_M_channels"])
        for i, ch in enumerate(self._cfg["ENC_M_channels"]):
            c.channels[i] = ch
        c.min_spatial_size, c.max_batch = self._cfg["min_spatial_size"], self.max_batch
        c.precision = _lib.precision_code(self.precision)
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.ipk_enc_create(ctypes.byref(c), ctypes.byref(h)), "ipk_enc_create")
            plan = _NativeEncPlan(h)
            keep = []
            for k, v in self.state_dict().items():
                t = v.detach().float().contiguous()
                keep.append(t)
                _lib.check(L.ipk_enc_set_tensor(h, k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(), _lib.DT_F32),
                           f"ipk_enc_set_tensor({k})")
            _lib.check(L.ipk_enc_finalize(h, _lib.current_stream_ptr()), "ipk_enc_finalize")
        self._plan, self._plan_key = plan, key
        return plan

    def forward(self, x, eps=None):
        """x: [B,3,T,H,W] -> (z, mu, logvar), each [B,z_dim,8,8].  `eps` defaults to a draw from the CPU generator, exactly like
        the reference (`torch.FloatTensor(size).normal_()`, motion_encoder.py:220)."""
        if x.dim() != 5 or x.shape[1] != 3 or x.shape[3] != self.spatial_size or x.shape[4] != self.spatial_size:
            raise ValueError(f"encoder input must be [B,3,T,{self.spatial_size},{self.spatial_size}], got {tuple(x.shape)}")
        B, T, z = x.shape[0], x.shape[2], self._cfg["z_dim"]
        if eps is None:
            eps = torch.FloatTensor(torch.Size((B, z, 8, 8))).normal_()
        xx = x.detach().float().contiguous()
        ee = eps.detach().float().to(xx.device).contiguous()
        plan = self._ensure_plan(xx.device, B)
        zo = torch.empty((B, z, 8, 8), device=xx.device, dtype=torch.float32)
        mu, lv = torch.empty_like(zo), torch.empty_like(zo)
        with torch.cuda.device(xx.device):
            _lib.check(_lib.lib().ipk_enc_forward(plan.handle, xx.data_ptr(), ee.data_ptr(), zo.data_ptr(), mu.data_ptr(), lv.data_ptr(),
                                                  B, T, _lib.current_stream_ptr()), "ipk_enc_forward")
        if self.be_determinstic:
            return mu, mu, mu                                   # motion_encoder.py:237-239
        return zo, mu, lv

    def _load_from_state_dict(self, *a, **k):
        self.invalidate()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)


def encode_first_stage(enc_motion, X, full_sequence=True, max_frames=10, eps=None, full_seq=True):
    """PokeMotionModel.encode_first_stage (models/second_stage_video.py:352-359): X [B,T,3,H,W] -> (motion, mu).
    The frame selection follows the reference exactly:
        full_seq (the second stage's flag):  X if the first stage was trained on full sequences or max_frames < 16, else X[:, :-1]
        not full_seq:                        X if the first stage was trained on full sequences, else X[:, 1:]
    full_sequence = first_stage_model.full_sequence, max_frames = config['data']['max_frames']."""
    with torch.no_grad():
        if full_seq:
            X_in = X if (full_sequence or max_frames < 16) else X[:, :-1]
        else:
            X_in = X if full_sequence else X[:, 1:]
        motion, mu, _ = enc_motion(X_in.transpose(1, 2), eps=eps)
    return motion, mu
