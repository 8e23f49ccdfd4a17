"""Sampling plumbing of PokeMotionModel around the two native plans.

Mirrors models/second_stage_video.py: make_flow_input(reverse=True) :289-300 (noise from the CPU generator),
forward_sample :326-343, forward_density :345-350, decode_first_stage :361-382.  The conditioning tensor `cond`
(= cat[conditioner(X[:,0]), poke_embedder(poke)], :311) is an input here: the frozen ConvEncoders are the "next" row
of SURVEY.md section 8f and stay in the caller.
"""
import ctypes

import torch

from . import _lib
from .first_stage import SpadeCondMotionDecoder
from .flow import SupervisedMacowTransformer


class PokeMotionSampler:
    def __init__(self, flow: SupervisedMacowTransformer, first_stage_model: SpadeCondMotionDecoder, conditioner=None, poke_embedder=None):
        """conditioner / poke_embedder: optional ipoke_b200.ConvEncoder drop-ins for `self.conditioner.encoder` and
        `self.poke_embedder.encoder` (second_stage_video.py:274,281); without them the caller passes `cond` itself."""
        self.flow = flow
        self.first_stage_model = first_stage_model
        self.conditioner = conditioner
        self.poke_embedder = poke_embedder
        self._pinned = {}

    def make_cond(self, x0, poke):
        """Conditioning half of make_flow_input (second_stage_video.py:268-287,311): cat[conditioner(x0), poke_embedder(poke)]."""
        from .cond_encoder import make_cond
        if self.conditioner is None or self.poke_embedder is None:
            raise RuntimeError("PokeMotionSampler.make_cond needs the conditioner and poke embedder encoders")
        return make_cond(self.conditioner, self.poke_embedder, x0, poke)

    # -- second_stage_video.py:289-300: z ~ N(0, I) drawn on the CPU default generator, then moved to the device
    def draw_noise(self, batch_size, device=None, generator=None):
        C0 = self.flow._cfg["flow_in_channels"]
        z = torch.randn((batch_size, C0, 8, 8), generator=generator)
        return z if device is None else z.to(device)

    def sample(self, z, cond, x0, length):
        """flow inverse -> ConvGRU + decoder on DEVICE tensors: [B,C0,8,8], [B,h,8,8], [B,3,S,S] -> [B,T,3,S,S].
        Uses the fused C entry (the sampled latent stays NHWC on the device between the two stages)."""
        z = z.detach().float().contiguous()
        cond = cond.detach().float().contiguous()
        x0 = x0.detach().float().contiguous()
        B = z.shape[0]
        fs = self.first_stage_model
        fplan = self.flow._ensure_plan(z.device, B)
        dplan = fs._ensure_plan(z.device, B, length)
        out = torch.empty((B, int(length), 3, fs.spatial, fs.spatial), device=z.device, dtype=torch.float32)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().ipk_sample(fplan.handle, dplan.handle, z.data_ptr(), cond.data_ptr(), x0.data_ptr(), out.data_ptr(),
                                             B, int(length), _lib.current_stream_ptr()), "ipk_sample")
        return out

    def _pin(self, name, shape):
        t = self._pinned.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=torch.float32).pin_memory()
            self._pinned[name] = t
        return t

    def sample_host(self, z, cond, x0, length, device=None):
        """Same step with HOST tensors: inputs are staged through pinned memory, copied to the device, and the frames are
        copied back into a pinned host tensor (returned).  Synchronous."""
        device = torch.device(device if device is not None else "cuda:0")
        B = z.shape[0]
        fs = self.first_stage_model
        def staged(name, t):
            # already-pinned fp32 host tensors go to the device as they are; anything else is staged through a pinned buffer
            if t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous() and t.is_pinned():
                return t
            buf = self._pin(name, t.shape)
            buf.copy_(t)
            return buf
        zp, cp, xp = staged("z", z), staged("cond", cond), staged("x0", x0)
        out = self._pin("frames", (B, int(length), 3, fs.spatial, fs.spatial))
        fplan = self.flow._ensure_plan(device, B)
        dplan = fs._ensure_plan(device, B, length)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().ipk_sample_host(fplan.handle, dplan.handle, zp.data_ptr(), cp.data_ptr(), xp.data_ptr(), out.data_ptr(),
                                                  B, int(length), _lib.current_stream_ptr()), "ipk_sample_host")
        return out

    def forward_sample(self, X, cond=None, n_samples=1, n_logged_vids=1, length=None, add_first_frame=False, poke=None):
        """PokeMotionModel.forward_sample (second_stage_video.py:326-343): returns a list of n_samples CPU tensors.
        Pass `cond`, or `poke` to have it computed by the conditioning encoders -- ONCE for all samples (the reference
        re-encodes the same batch for every sample, second_stage_video.py:332-333; the result is sample-invariant)."""
        videos = []
        if length is None:
            length = X.size(1) - 1
        if cond is None:
            cond = self.make_cond(X[:, 0], poke)
        with torch.no_grad():
            for _ in range(n_samples):
                z = self.draw_noise(X.size(0)).type_as(X)
                out = self.sample(z, cond, X[:, 0], length)
                if add_first_frame:
                    out = torch.cat([X[:, 0].unsqueeze(1), out], dim=1)
                videos.append(out[:n_logged_vids].cpu())
        return videos

    def forward_density(self, flow_input, cond):
        """PokeMotionModel.forward_density (second_stage_video.py:345-350) with the encoded latent supplied by the caller."""
        return self.flow(flow_input.detach(), cond, reverse=False)
