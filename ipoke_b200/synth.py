"""Synthetic, seeded checkpoints for benchmarking (no weights ship with the reference).

Fills the drop-in modules' parameters in place, on their device, with non-trivial but numerically stable values:
plain convs U(+-1/sqrt(fan_in)), weight-norm v ~ N(0, 0.05), g ~ U(0, 0.05), small biases, ActNorm log_scale ~ N(0, 0.01),
random channel permutations, unit spectral-norm vectors.  Fresh reference init would make every coupling the identity
(zero_init), which would make the benchmark's arithmetic trivially compressible.
"""
import math

import torch


@torch.no_grad()
def fill_flow_(flow_module, seed=0, g_scale=0.05):
    dev = next(flow_module.parameters()).device
    gen = torch.Generator(device=dev).manual_seed(seed)
    for name, p in flow_module.named_parameters():
        if name.endswith("log_scale"):
            p.copy_(torch.randn(p.shape, generator=gen, device=dev) * 0.01)
        elif name.endswith("weight_g"):
            p.copy_(torch.rand(p.shape, generator=gen, device=dev) * g_scale)
        elif name.endswith("weight_v"):
            p.copy_(torch.randn(p.shape, generator=gen, device=dev) * 0.05)
        elif name.endswith("bias"):
            p.copy_(torch.randn(p.shape, generator=gen, device=dev) * 0.02)
        elif name.endswith("weight"):
            fan_in = p[0].numel()
            p.copy_((torch.rand(p.shape, generator=gen, device=dev) * 2 - 1) / math.sqrt(fan_in))
    prev = None
    for name, b in flow_module.named_buffers():
        if name.endswith("forward_shuffle_idx"):
            prev = torch.randperm(b.numel(), generator=gen, device=dev)
            b.copy_(prev)
        elif name.endswith("backward_shuffle_idx"):
            b.copy_(torch.argsort(prev))
        elif name.endswith("initialized"):
            b.fill_(1)
    flow_module.invalidate(flags=True)
    return flow_module


@torch.no_grad()
def fill_first_stage_(fs_module, seed=0):
    dev = next(fs_module.parameters()).device
    gen = torch.Generator(device=dev).manual_seed(seed)
    for name, p in fs_module.named_parameters():
        if name == "motion_bias":
            p.copy_(torch.randn(p.shape, generator=gen, device=dev))
        elif name.endswith("norm.weight"):
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=gen, device=dev))
        elif name.endswith("bias"):
            p.copy_(torch.randn(p.shape, generator=gen, device=dev) * 0.05)
        elif name.endswith("weight") or name.endswith("weight_orig"):
            transposed = ".blocks." in name and (".conv1.conv." in name or ".res_conv.conv." in name)
            fan_in = (p.shape[1] if transposed else p.shape[1]) * 9 if not transposed else p.shape[1] * 9
            p.copy_((torch.rand(p.shape, generator=gen, device=dev) * 2 - 1) / math.sqrt(fan_in))
    for name, b in fs_module.named_buffers():
        if name.endswith("weight_u") or name.endswith("weight_v"):
            b.copy_(torch.nn.functional.normalize(torch.randn(b.shape, generator=gen, device=dev), dim=0, eps=1e-12))
    fs_module.invalidate()
    return fs_module


@torch.no_grad()
def fill_encoder_(enc_module, seed=0):
    """3-D conv video encoder (motion_encoder.py:150-241): Conv3d kaiming-normal fan_out (:192-194), GroupNorm weight ~ 1 + 0.1 N(0,1),
    bias ~ 0.05 N(0,1), 2-D heads nn.Conv2d default."""
    dev = next(enc_module.parameters()).device
    gen = torch.Generator(device=dev).manual_seed(seed)
    for name, p in enc_module.named_parameters():
        if p.dim() == 5:
            fan_out = p.shape[0] * p.shape[2] * p.shape[3] * p.shape[4]
            p.copy_(torch.randn(p.shape, generator=gen, device=dev) * math.sqrt(2.0 / fan_out))
        elif p.dim() == 4:
            p.copy_((torch.rand(p.shape, generator=gen, device=dev) * 2 - 1) / math.sqrt(p.shape[1] * 9))
        elif name.startswith(("conv_mu", "conv_var")):
            p.copy_((torch.rand(p.shape, generator=gen, device=dev) * 2 - 1) / math.sqrt(256 * 9))
        elif name.endswith("weight"):
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=gen, device=dev))
        else:
            p.copy_(0.05 * torch.randn(p.shape, generator=gen, device=dev))
    enc_module.invalidate()
    return enc_module
