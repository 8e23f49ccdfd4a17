"""Second-stage training step of the flow on the native library (BASELINE configs[3]).

Mirrors PokeMotionModel.training_step (models/second_stage_video.py:345-350, 596-631): `forward_density` -> FlowLoss
(models/modules/INN/loss.py:13-31) -> backward -> Adam(amsgrad) (`configure_optimizers`, :633-660).  The reference runs this under
PyTorch-Lightning DDP (bucketed gradient all-reduce, replicated optimizer state); here the flow's parameters live in ONE flat fp32
buffer, the native step (`ipk_flowtrain_step`) writes the flat gradient, and the optimizer is sharded: one `reduce_scatter` of the
flat gradient over NCCL, Adam on each rank's shard (`ipk_adam_step`), one `all_gather` of the parameters.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .flow import SupervisedMacowTransformer


class _NativeTrainPlan:
    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().ipk_flowtrain_destroy(self.handle)
        except Exception:
            pass
        self.handle = None


def shard_range(n, world, rank):
    """Contiguous equal shards of a flat buffer padded to a multiple of `world`: returns (padded_n, lo, hi)."""
    per = (n + world - 1) // world
    return per * world, rank * per, (rank + 1) * per


def sharded_update(flat_params, flat_grads, lo, hi, world, apply_fn, shard_grad=None, group=None):
    """One data-parallel optimizer step on a flat parameter buffer (length a multiple of `world`): sum-reduce the flat gradient so that
    each rank holds the sum of its shard [lo, hi), let `apply_fn(param_shard, grad_shard_sum)` update that shard in place (it scales
    by 1 / world itself), then all-gather the shards so that every rank holds identical parameters.  NCCL: one reduce_scatter + one
    all_gather; backends without reduce_scatter (gloo, the CPU tests) all-reduce and slice."""
    if world == 1:
        apply_fn(flat_params[lo:hi], flat_grads[lo:hi])
        return
    if dist.get_backend(group) == "nccl":
        if shard_grad is None:
            shard_grad = torch.empty(hi - lo, device=flat_grads.device, dtype=flat_grads.dtype)
        dist.reduce_scatter_tensor(shard_grad, flat_grads, op=dist.ReduceOp.SUM, group=group)
        g = shard_grad
    else:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
        g = flat_grads[lo:hi]
    p = flat_params[lo:hi]
    apply_fn(p, g)
    dist.all_gather_into_tensor(flat_params, p.contiguous(), group=group)


class FlowDensityFunction(torch.autograd.Function):
    """`out, logdet = flow(x, cond)` as ONE autograd node over the native training plan: forward = ipk_flowtrain_forward (tape kept on
    the device), backward = ipk_flowtrain_backward with the upstream (dz, dlogdet).  The flow's parameters are passed as inputs so that
    autograd accumulates into their `.grad` (what `loss.backward()` does for models/second_stage_video.py:409-415)."""

    @staticmethod
    def forward(ctx, trainer, x, cond, x_orig, *params):
        z, logdet = trainer.forward_native(x, cond)
        ctx.trainer = trainer
        ctx.batch = x.shape[0]
        ctx.serial = trainer.mark_forward()
        ctx.mark_non_differentiable()
        return z, logdet

    @staticmethod
    def backward(ctx, dz, dlogdet):
        tr = ctx.trainer
        if ctx.serial != tr.forward_serial:
            raise RuntimeError("ipoke_b200 flow: backward() of a stale forward -- the native tape holds the most recent density-direction call only "
                               "(one forward, then its backward)")
        dx = tr.backward_native(dz, dlogdet, ctx.batch, want_input_grad=ctx.needs_input_grad[3])
        g = tr.flat_grads[:tr.numel].clone()          # fresh storage per backward: autograd may keep (or accumulate into) these
        grads = tuple(g[o:o + n].view(shape) for (o, n, shape) in tr.param_slices)
        return (None, None, None, dx) + grads


class FlowTrainer:
    """flow: ipoke_b200.SupervisedMacowTransformer on a CUDA device.  After construction the module's parameters are views into
    `self.flat_params`, so sampling through the same module sees every optimizer update.

    Checkpointing: `state_dict()` / `load_state_dict()` carry the sharded Adam(amsgrad) state as FULL-size tensors (gathered over the
    ranks), so a run resumes on any number of ranks; the parameters themselves travel in the module's own state-dict.  With more than
    one rank the constructor broadcasts rank 0's parameters (what DDP does at wrap time).
    Caveat: the parameters' `.data` are views of `flat_params`; `flow.to(...)` or anything else that re-allocates parameters after
    construction breaks that relation -- construct the trainer last (load_state_dict copies in place and is fine)."""

    def __init__(self, flow: SupervisedMacowTransformer, max_batch=32, precision=None, group=None, distributed=True):
        dev = next(flow.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("ipoke_b200 FlowTrainer runs on CUDA only (no CPU fallback)")
        self.flow, self.device, self.group = flow, dev, group
        self.world = dist.get_world_size(group) if (distributed and dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.precision = precision or flow.precision
        self.max_batch = int(max_batch)
        named = [(k, p) for k, p in flow.named_parameters() if p.dtype.is_floating_point]
        self.names = [k for k, _ in named]
        n = sum(p.numel() for _, p in named)
        self.numel = n
        self.padded, self.lo, self.hi = shard_range(n, self.world, self.rank)
        self.flat_params = torch.zeros(self.padded, device=dev, dtype=torch.float32)
        self.flat_grads = torch.zeros(self.padded, device=dev, dtype=torch.float32)
        self.offsets, off = {}, 0
        for k, p in named:
            view = self.flat_params[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            self.offsets[k] = (off, p.numel(), tuple(p.shape))
            off += p.numel()
        self.param_list = [p for _, p in named]
        self.param_slices = [self.offsets[k] for k in self.names]
        if self.world > 1:
            dist.broadcast(self.flat_params, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        flow.invalidate()
        self.forward_serial = 0
        self.exp_avg = torch.zeros(self.hi - self.lo, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        self.max_exp_avg_sq = torch.zeros_like(self.exp_avg)
        self.shard_grad = torch.zeros_like(self.exp_avg) if self.world > 1 else None
        self.steps = 0
        self._plan = None

    def grad(self, name):
        off, n, shape = self.offsets[name]
        return self.flat_grads[off:off + n].view(shape)

    def invalidate_plan(self):
        self._plan = None

    def mark_forward(self):
        self.forward_serial += 1
        return self.forward_serial

    # ------------------------------------------------------------------ checkpointing (torch.optim.Adam-compatible full-size state)
    def _gather_full(self, shard):
        if self.world == 1:
            return shard[: self.numel].clone()
        full = torch.empty(self.padded, device=self.device, dtype=shard.dtype)
        dist.all_gather_into_tensor(full, shard.contiguous(), group=self.group)
        return full[: self.numel]

    def state_dict(self):
        """Collective when world > 1 (every rank must call it; every rank gets the full state)."""
        return {"step": int(self.steps), "numel": int(self.numel), "names": list(self.names),
                "exp_avg": self._gather_full(self.exp_avg).cpu(), "exp_avg_sq": self._gather_full(self.exp_avg_sq).cpu(),
                "max_exp_avg_sq": self._gather_full(self.max_exp_avg_sq).cpu()}

    def load_state_dict(self, sd):
        if int(sd["numel"]) != self.numel or list(sd["names"]) != list(self.names):
            raise ValueError("FlowTrainer.load_state_dict: the optimizer state belongs to a different flow (parameter list mismatch)")
        self.steps = int(sd["step"])
        hi = min(self.hi, self.numel)
        for name in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
            dst = getattr(self, name)
            dst.zero_()
            if hi > self.lo:
                dst[: hi - self.lo].copy_(sd[name][self.lo:hi].to(self.device))

    def per_parameter_state(self):
        """The optimizer state in torch.optim.Adam's layout ({name: {step, exp_avg, exp_avg_sq, max_exp_avg_sq}}); collective like state_dict."""
        sd = self.state_dict()
        out = {}
        for k in self.names:
            off, n, shape = self.offsets[k]
            out[k] = {"step": torch.tensor(float(sd["step"])), "exp_avg": sd["exp_avg"][off:off + n].view(shape),
                      "exp_avg_sq": sd["exp_avg_sq"][off:off + n].view(shape), "max_exp_avg_sq": sd["max_exp_avg_sq"][off:off + n].view(shape)}
        return out

    def _ensure_plan(self):
        if self._plan is not None:
            return self._plan
        L, cfg = _lib.lib(), self.flow._cfg
        c = _lib.FlowConfig()
        c.flow_in_channels, c.flow_mid_channels, c.h_channels = cfg["flow_in_channels"], cfg["flow_mid_channels"], cfg["h_channels"]
        c.n_levels = len(cfg["num_steps"])
        for i, s in enumerate(cfg["num_steps"]):
            c.num_steps[i] = s
        c.factor = cfg["factor"]
        c.kernel_h, c.kernel_w = cfg["kernel_size"]
        c.precision = _lib.precision_code(self.precision)
        c.max_batch = self.max_batch
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.ipk_flowtrain_create(ctypes.byref(c), ctypes.byref(h)), "ipk_flowtrain_create")
            plan = _NativeTrainPlan(h)
            self._keep = []
            for k, v in self.flow.state_dict().items():
                if k.endswith("initialized"):
                    if int(v) == 0:
                        raise RuntimeError(f"ipoke_b200 FlowTrainer: '{k}' is 0 -- run the data-dependent init of the reference first or load an "
                                           "initialised checkpoint (macow2.py:503-505, macow_utils.py:248-250)")
                    continue
                if k in self.offsets:
                    off, n, _ = self.offsets[k]
                    pp, gp = self.flat_params[off:off + n], self.flat_grads[off:off + n]
                    _lib.check(L.ipk_flowtrain_set_tensor(h, k.encode(), ctypes.c_void_p(pp.data_ptr()), ctypes.c_void_p(gp.data_ptr()), n, _lib.DT_F32),
                               f"ipk_flowtrain_set_tensor({k})")
                else:
                    t = v.detach().contiguous()
                    self._keep.append(t)
                    _lib.check(L.ipk_flowtrain_set_tensor(h, k.encode(), ctypes.c_void_p(t.data_ptr()), None, t.numel(), _lib.dtype_code(t)),
                               f"ipk_flowtrain_set_tensor({k})")
            _lib.check(L.ipk_flowtrain_finalize(h, _lib.current_stream_ptr()), "ipk_flowtrain_finalize")
        self._plan = plan
        return plan

    def step(self, flow_input, cond, return_latent=False, return_log=False):
        """forward_density + FlowLoss + backward: returns the loss (0-dim CUDA tensor); gradients land in `flat_grads`.
        return_log=True also returns FlowLoss's log dict (models/modules/INN/loss.py:23-30), including the `reference_nll_loss` of a
        fresh `torch.randn_like(z)` draw (which consumes the device RNG exactly like the reference)."""
        x = flow_input.detach().float().contiguous()
        c = cond.detach().float().contiguous()
        B = x.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds the trainer's max_batch {self.max_batch}")
        if self.flow._prepare_initialization(False):
            self.flow.data_init(x)                     # first training forward of a fresh model (macow2.py:503-505, macow_utils.py:248-250)
            self._plan = None
        plan = self._ensure_plan()
        loss = torch.zeros((), device=self.device, dtype=torch.float32)
        want = return_latent or return_log
        z = torch.empty_like(x) if want else None
        ld = torch.empty(B, device=self.device, dtype=torch.float32) if want else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ipk_flowtrain_step(plan.handle, x.data_ptr(), c.data_ptr(), loss.data_ptr(), z.data_ptr() if z is not None else None,
                                                     ld.data_ptr() if ld is not None else None, B, _lib.current_stream_ptr()), "ipk_flowtrain_step")
        self.forward_serial += 1
        out = (loss, z, ld) if return_latent else (loss,)
        if return_log:
            nll_loss = (0.5 * (z ** 2).flatten(1).sum(dim=1)).mean()
            nlogdet_loss = -ld.mean()
            reference_nll_loss = (0.5 * (torch.randn_like(z) ** 2).flatten(1).sum(dim=1)).mean()
            out = out + ({"flow_loss": loss, "reference_nll_loss": reference_nll_loss, "nlogdet_loss": nlogdet_loss, "nll_loss": nll_loss,
                          "logdet_weight": 1.0},)
        return out[0] if len(out) == 1 else out

    # ------------------------------------------------------------------ the two halves behind FlowDensityFunction
    def forward_native(self, x, cond):
        B = x.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds the trainer's max_batch {self.max_batch}")
        plan = self._ensure_plan()
        z = torch.empty_like(x)
        ld = torch.empty(B, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ipk_flowtrain_forward(plan.handle, x.data_ptr(), cond.data_ptr(), z.data_ptr(), ld.data_ptr(), B,
                                                        _lib.current_stream_ptr()), "ipk_flowtrain_forward")
        return z, ld

    def backward_native(self, dz, dlogdet, B, want_input_grad=False):
        plan = self._ensure_plan()
        dzc = dz.detach().float().contiguous() if dz is not None else None
        dlc = dlogdet.detach().float().contiguous() if dlogdet is not None else None
        dx = torch.empty((B, self.flow._cfg["flow_in_channels"], 8, 8), device=self.device, dtype=torch.float32) if want_input_grad else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ipk_flowtrain_backward(plan.handle, dzc.data_ptr() if dzc is not None else None,
                                                         dlc.data_ptr() if dlc is not None else None, dx.data_ptr() if dx is not None else None, B,
                                                         _lib.current_stream_ptr()), "ipk_flowtrain_backward")
        return dx

    def optimizer_step(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True):
        """torch.optim.Adam semantics on the flat parameter buffer, sharded over the ranks of `group`.  The reference's optimizer is
        Adam(lr=cfg.training.lr, betas=(0.9, 0.999), weight_decay=cfg.training.weight_decay, amsgrad=True)
        (models/second_stage_video.py:647-648); bias corrections are evaluated in double precision on the host like torch does."""
        self.steps += 1
        L = _lib.lib()

        def adam(p, g):
            with torch.cuda.device(self.device):
                _lib.check(L.ipk_adam_step(p.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                           self.max_exp_avg_sq.data_ptr() if amsgrad else None, p.numel(), float(lr), float(betas[0]), float(betas[1]),
                                           float(eps), float(weight_decay), int(self.steps), 1.0 / self.world, _lib.current_stream_ptr()), "ipk_adam_step")

        sharded_update(self.flat_params, self.flat_grads, self.lo, self.hi, self.world, adam, shard_grad=self.shard_grad, group=self.group)
        self.flow.invalidate()          # the inference plan re-packs from the updated parameters on its next use
