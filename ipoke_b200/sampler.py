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
        # frame-selection flags of encode_first_stage (second_stage_video.py:352-359): the second stage's `full_seq`, the first stage's
        # `full_sequence`, config['data']['max_frames']
        self.full_seq = True
        self.full_sequence = bool(getattr(first_stage_model, "full_sequence", True))
        self.max_frames = 10

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

    def _pin(self, name, shape, dtype=torch.float32):
        t = self._pinned.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=dtype).pin_memory()
            self._pinned[name] = t
        return t

    def sample_host(self, z, cond, x0, length, device=None, uint8=False, reuse=False):
        """Same step with HOST tensors: inputs are staged through pinned memory, copied to the device, and the frames are
        copied back into a pinned host tensor (returned).  Synchronous.
        The returned tensor is a FRESH pinned allocation per call (torch's caching host allocator recycles freed blocks, so a loop
        that drops its results pays the page-locking once); reuse=True returns one cached buffer per output shape instead, which the
        next call with the same shape OVERWRITES -- only for callers that consume the frames before calling again.
        uint8=True returns the post-processed sample instead: uint8 [B,T,S,S,3] = ((x + 1) * 127.5).permute(0,1,3,4,2) as
        the callers of forward_sample compute it on the host (second_stage_video.py:673-675), converted on the device."""
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
        fplan = self.flow._ensure_plan(device, B)
        dplan = fs._ensure_plan(device, B, length)
        if uint8:
            shape = (B, int(length), fs.spatial, fs.spatial, 3)
            out = self._pin("frames_u8", shape, torch.uint8) if reuse else torch.empty(shape, dtype=torch.uint8, pin_memory=True)
            with torch.cuda.device(device):
                _lib.check(_lib.lib().ipk_sample_host_u8(fplan.handle, dplan.handle, zp.data_ptr(), cp.data_ptr(), xp.data_ptr(), out.data_ptr(),
                                                         B, int(length), _lib.current_stream_ptr()), "ipk_sample_host_u8")
            return out
        shape = (B, int(length), 3, fs.spatial, fs.spatial)
        out = self._pin("frames", shape) if reuse else torch.empty(shape, dtype=torch.float32, pin_memory=True)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().ipk_sample_host(fplan.handle, dplan.handle, zp.data_ptr(), cp.data_ptr(), xp.data_ptr(), out.data_ptr(),
                                                  B, int(length), _lib.current_stream_ptr()), "ipk_sample_host")
        return out

    @staticmethod
    def to_uint8(frames):
        """Device-side `((frames + 1.) * 127.5).permute(0, 1, 3, 4, 2) -> uint8` (second_stage_video.py:673-675): CUDA fp32
        [..., 3, S, S] -> CUDA uint8 [..., S, S, 3]; `.cpu().numpy()` of the result is what the reference hands to save_video."""
        if not frames.is_cuda:
            raise RuntimeError("ipoke_b200: to_uint8 runs on the device (no CPU fallback)")
        f = frames.detach().float().contiguous()
        S = f.shape[-1]
        if f.shape[-3] != 3 or f.shape[-2] != S:
            raise ValueError(f"to_uint8 expects [..., 3, S, S] frames, got {tuple(f.shape)}")
        out = torch.empty((*f.shape[:-3], S, S, 3), device=f.device, dtype=torch.uint8)
        n = f.numel() // (3 * S * S)
        if n:
            with torch.cuda.device(f.device):
                _lib.check(_lib.lib().ipk_frames_to_u8(f.data_ptr(), out.data_ptr(), n, S, _lib.current_stream_ptr()), "ipk_frames_to_u8")
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

    def control_sensitivity(self, X, pokes, length=None, max_batch=None):
        """Compute core of PokeMotionModel._control_sensitivity (second_stage_video.py:786-848): one sample per poke for a
        stack of pokes `pokes` [P, B, 2, H, W] on the same start frames X[:, 0] -> [B, P, T, 3, H, W] on the host.
        The reference calls forward_sample once per poke (P x the image conditioner on the same frames, P small flow
        batches); here the image conditioning is encoded once, the P pokes go through the poke embedder as one batch, and
        the P*B latents -- drawn in the reference's order, one randn(B, C0, 8, 8) per poke -- run as one sampling pass
        (chunks of `max_batch` videos)."""
        if self.conditioner is None or self.poke_embedder is None:
            raise RuntimeError("control_sensitivity needs the conditioner and poke embedder encoders")
        P, B = pokes.shape[0], pokes.shape[1]
        if length is None:
            length = X.size(1) - 1
        x0 = X[:, 0]
        with torch.no_grad():
            img_cond, *_ = self.conditioner(x0)
            poke_emb, *_ = self.poke_embedder(pokes.reshape(P * B, *pokes.shape[2:]).type_as(X))
            cond = torch.cat([img_cond.repeat(P, 1, 1, 1), poke_emb], dim=1)             # row p*B + b
            z = torch.cat([self.draw_noise(B) for _ in range(P)]).type_as(X)
            x0r = x0.repeat(P, 1, 1, 1)
            step = int(max_batch) if max_batch else P * B
            outs = [self.sample(z[i:i + step], cond[i:i + step], x0r[i:i + step], length).cpu() for i in range(0, P * B, step)]
        out = torch.cat(outs)                                                            # [P*B, T, 3, H, W]
        return out.reshape(P, B, *out.shape[1:]).transpose(0, 1).contiguous()

    def transfer(self, X_1, X_2, poke_1, enc_motion, eps=None, residual=None, embed_poke_and_image=False, length=None):
        """Compute core of PokeMotionModel._test_transfer (second_stage_video.py:948-1015): the motion residual of video
        X_1 under its own conditioning is re-synthesised on the start frame of X_2.
            z_1 = enc_motion(X_1); cond_1 = cond(X_1[:,0], poke_1); cond_2 = cond(X_2[:,0], poke_1 on source 2)
            r_1 = flow(z_1, cond_1);  z_r1_c2 = flow^-1(r_1, cond_2);  z_rand_c2 = flow^-1(N(0,I), cond_2)
        Returns (vid_r1_c2, vid_random_cond2, r_1): two [B,T,3,H,W] device tensors decoded on X_2[:,0] and the residual.
        `eps` is the encoder's reparameterisation noise, `residual` the N(0,I) draw (randn_like(r1), :1000); both default to
        the CPU generator.  X_*: [B, T+1, 3, H, W]; enc_motion: ipoke_b200.ResNetMotionEncoder."""
        from .encoder import encode_first_stage
        if self.conditioner is None or self.poke_embedder is None:
            raise RuntimeError("transfer needs the conditioner and poke embedder encoders")
        fs = self.first_stage_model
        if length is None:
            length = X_2.size(1) - 1
        with torch.no_grad():
            z_1, *_ = encode_first_stage(enc_motion, X_1, full_sequence=self.full_sequence, max_frames=self.max_frames, eps=eps,
                                         full_seq=self.full_seq)
            p1 = torch.cat([poke_1, X_1[:, 0]], dim=1) if embed_poke_and_image else poke_1
            p1_src2 = torch.cat([poke_1, X_2[:, 0]], dim=1) if embed_poke_and_image else poke_1
            cond_1 = self.make_cond(X_1[:, 0], p1)
            cond_2 = self.make_cond(X_2[:, 0], p1_src2)
            r1, _ = self.flow(z_1, cond_1, reverse=False)
            if residual is None:
                residual = torch.randn(tuple(r1.shape)).to(r1.device)
            z_r1_c2 = self.flow(r1, cond_2, reverse=True)
            z_rand_c2 = self.flow(residual.type_as(r1), cond_2, reverse=True)
            from .first_stage import decode_first_stage
            vid_r1_c2 = decode_first_stage(fs, z_r1_c2, X_2, length)
            vid_random = decode_first_stage(fs, z_rand_c2, X_2, length)
        return vid_r1_c2, vid_random, r1

    def forward_density(self, flow_input, cond):
        """PokeMotionModel.forward_density (second_stage_video.py:345-350) with the encoded latent supplied by the caller."""
        return self.flow(flow_input.detach(), cond, reverse=False)
