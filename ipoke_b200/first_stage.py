"""Drop-ins for the first-stage decoder side of SpadeCondMotionModel (models/first_stage_motion_model.py:469-496):
`.rnn` (ConvGRU, models/modules/motion_models/rnn.py:59-133), `.gen` (SpadeCondConvDecoder,
models/modules/autoencoders/fully_conv_models.py:135-177) and `.motion_bias`, plus
`decode_first_stage(motion, X, length)` (models/second_stage_video.py:361-382).  State-dict layout is the reference's
(legacy spectral-norm weight_orig / weight_u / weight_v included)."""
import ctypes

import torch
from torch import nn

from . import _lib, spec
from .flow import _Holder, build_param_tree


class _NativeFsPlan:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().ipk_fs_destroy(self.handle)
        except Exception:
            pass
        self.handle = None


class _RnnView(nn.Module):
    """first_stage_model.rnn(x, hidden: list[n_layers]) -> list[n_layers]   (rnn.py:104)"""

    def __init__(self, owner, tree):
        super().__init__()
        self._owner = [owner]
        self.cells = tree._modules["cells"]
        self.n_layers = owner.n_layers

    def forward(self, x, hidden=None):
        return self._owner[0]._gru_step(x, hidden)


class _GenView(nn.Module):
    """first_stage_model.gen(actual_frame: list[Tensor], start_frame, del_shape=True) -> [B,3,H,W]
    (fully_conv_models.py:166-177; pops the list like the reference)."""

    def __init__(self, owner, tree):
        super().__init__()
        self._owner = [owner]
        for k, m in tree._modules.items():
            self.add_module(k, m)

    def forward(self, actual_frame, start_frame, del_shape=True):
        h = actual_frame.pop() if del_shape else actual_frame[-1]
        if del_shape:
            assert not actual_frame
        return self._owner[0]._gen(h, start_frame)


class SpadeCondMotionDecoder(nn.Module):
    """The decoder half of SpadeCondMotionModel: attributes n_layers, use_motion_bias, motion_bias, full_sequence, rnn, gen.

    config: the first stage's config['architecture'] (z_dim, dec_channels, n_gru_layers, min_spatial_size, norm,
    spectral_norm, motion_bias) plus 'spatial' (output resolution; in the reference it comes from
    config['data']['spatial_size']).  Extra keys: ipk_precision, ipk_max_batch, ipk_max_frames, ipk_chunk_videos."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.z_dim = int(config["z_dim"])
        self.dec_channels = [int(c) for c in config["dec_channels"]]
        self.n_layers = int(config.get("n_gru_layers", 4))
        ms = int(config.get("min_spatial_size", 8))
        if ms != 8:
            raise NotImplementedError("ipoke_b200 decoder: min_spatial_size must be 8 (config/first_stage.yaml:61)")
        self.spatial = int(config.get("spatial", ms << (len(self.dec_channels) - 1)))
        if config.get("norm", "group") not in ("group", "Group"):
            raise NotImplementedError("ipoke_b200 decoder: only norm='group' (config/first_stage.yaml:54)")
        self.use_motion_bias = bool(config.get("motion_bias", True))
        if not self.use_motion_bias:
            raise NotImplementedError("ipoke_b200 decoder: motion_bias=False is not used by any shipped config")
        self.full_sequence = bool(config.get("full_seq", True))
        self._cfg = dict(z_dim=self.z_dim, dec_channels=self.dec_channels, n_gru_layers=self.n_layers, min_spatial_size=ms,
                         spectral_norm=bool(config.get("spectral_norm", True)))
        tree = _Holder()
        build_param_tree(tree, spec.first_stage_param_spec(self._cfg))
        self.motion_bias = tree._parameters["motion_bias"]
        self.rnn = _RnnView(self, tree._modules["rnn"])
        self.gen = _GenView(self, tree._modules["gen"])
        self.precision = config.get("ipk_precision", "fp32")
        self.max_batch = int(config.get("ipk_max_batch", 64))
        self.max_frames = int(config.get("ipk_max_frames", 16))
        self.chunk_videos = int(config.get("ipk_chunk_videos", 0))
        self._plan = None
        self._plan_key = None
        self._plist = None

    def invalidate(self):
        self._plan = None
        self._plan_key = None
        self._plist = None

    def _ensure_plan(self, device, batch, frames=1):
        if batch > self.max_batch or frames > self.max_frames:
            self.max_batch = max(self.max_batch, int(batch))
            self.max_frames = max(self.max_frames, int(frames))
            self.invalidate()
        if self._plist is None:
            self._plist = list(self.parameters()) + list(self.buffers())      # buffers: the spectral-norm u / v vectors
        key = (device, self.precision, self.max_batch, self.max_frames) + _lib.tensors_key(self._plist)
        if self._plan is not None and self._plan_key == key:
            return self._plan
        if device.type != "cuda":
            raise RuntimeError("ipoke_b200 decoder runs on CUDA tensors only (no CPU fallback)")
        L = _lib.lib()
        c = _lib.FsConfig()
        c.z_dim, c.spatial, c.n_gru_layers, c.n_dec = self.z_dim, self.spatial, self.n_layers, len(self.dec_channels)
        for i, ch in enumerate(self.dec_channels):
            c.dec_channels[i] = ch
        c.precision = _lib.precision_code(self.precision)
        c.max_batch, c.max_frames, c.chunk_videos = self.max_batch, self.max_frames, self.chunk_videos
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.ipk_fs_create(ctypes.byref(c), ctypes.byref(h)), "ipk_fs_create")
            plan = _NativeFsPlan(h)
            keep = []
            for k, v in self.state_dict().items():
                t = v.detach().float().contiguous()
                keep.append(t)
                _lib.check(L.ipk_fs_set_tensor(h, k.encode(), ctypes.c_void_p(t.data_ptr()), t.numel(), _lib.DT_F32),
                           f"ipk_fs_set_tensor({k})")
            _lib.check(L.ipk_fs_finalize(h, _lib.current_stream_ptr()), "ipk_fs_finalize")
        self._plan, self._plan_key = plan, key
        return plan

    # ------------------------------------------------------------------ reference API
    def decode(self, motion, start_frame, length):
        """PokeMotionModel.decode_first_stage body (second_stage_video.py:361-382): [B,z,8,8], [B,3,S,S] -> [B,T,3,S,S]."""
        motion = motion.detach().float().contiguous()
        x0 = start_frame.detach().float().contiguous()
        B = motion.shape[0]
        if tuple(motion.shape[1:]) != (self.z_dim, 8, 8) or tuple(x0.shape) != (B, 3, self.spatial, self.spatial):
            raise ValueError(f"decode: bad shapes {tuple(motion.shape)} / {tuple(x0.shape)}")
        plan = self._ensure_plan(motion.device, B, length)
        out = torch.empty((B, int(length), 3, self.spatial, self.spatial), device=motion.device, dtype=torch.float32)
        with torch.cuda.device(motion.device):
            _lib.check(_lib.lib().ipk_fs_decode(plan.handle, motion.data_ptr(), x0.data_ptr(), out.data_ptr(), B, int(length),
                                                _lib.current_stream_ptr()), "ipk_fs_decode")
        return out

    def _gru_step(self, x, hidden):
        B = x.shape[0]
        if hidden is None:
            hidden = [None] * self.n_layers
        hs = [h if h is not None else torch.zeros_like(x) for h in hidden]       # rnn.py:41-47
        plan = self._ensure_plan(x.device, B)
        hin = torch.stack([h.detach().float() for h in hs], dim=0).contiguous()
        hout = torch.empty_like(hin)
        xx = x.detach().float().contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ipk_fs_gru_step(plan.handle, xx.data_ptr(), hin.data_ptr(), hout.data_ptr(), B,
                                                  _lib.current_stream_ptr()), "ipk_fs_gru_step")
        return [hout[i] for i in range(self.n_layers)]

    def _gen(self, h, start_frame):
        B = h.shape[0]
        plan = self._ensure_plan(h.device, B)
        hh = h.detach().float().contiguous()
        x0 = start_frame.detach().float().contiguous()
        out = torch.empty((B, 3, self.spatial, self.spatial), device=h.device, dtype=torch.float32)
        with torch.cuda.device(h.device):
            _lib.check(_lib.lib().ipk_fs_gen(plan.handle, hh.data_ptr(), x0.data_ptr(), out.data_ptr(), B,
                                             _lib.current_stream_ptr()), "ipk_fs_gen")
        return out

    def _load_from_state_dict(self, *a, **k):
        self.invalidate()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)


def decode_first_stage(first_stage_model, motion, X, length=None):
    """PokeMotionModel.decode_first_stage (second_stage_video.py:361-382): only X[:,0] and X.size(1) are used."""
    if length is None:
        length = X.size(1) - 1
    return first_stage_model.decode(motion, X[:, 0], length)
