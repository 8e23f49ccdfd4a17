"""Process-local shims that let the UNMODIFIED reference hot-path sub-modules import on CPU.

Only used by tests/golden/make_golden.py (fixture generation, run in the build container where
/root/reference exists).  Nothing here is imported by the product, by `-m gpu` tests, or by bench.py.

Shims:
  * `opt_einsum` is imported by models/modules/INN/modules.py:9 but only used by attention blocks
    (attention: false in every shipped config) -> empty stub module.
  * `Spade.forward` hard-codes `.cuda()` (models/modules/autoencoders/util.py:496) and so does
    `ResNetMotionEncoder.reparameterize` (motion_encoder.py:220) -> Tensor.cuda becomes identity.
"""
import sys
import types

REF = "/root/reference"


def install():
    import torch
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "opt_einsum" not in sys.modules:
        m = types.ModuleType("opt_einsum")
        m.contract = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("opt_einsum stub"))
        sys.modules["opt_einsum"] = m
    if not getattr(torch.Tensor, "_ipk_cuda_shim", False):
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.Tensor._ipk_cuda_shim = True


def flow_cls():
    install()
    from models.modules.INN.INN import SupervisedMacowTransformer
    return SupervisedMacowTransformer


def first_stage_parts():
    install()
    from models.modules.motion_models.rnn import ConvGRU
    from models.modules.autoencoders.fully_conv_models import SpadeCondConvDecoder
    return ConvGRU, SpadeCondConvDecoder


def encoder_fn():
    install()
    from models.modules.motion_models.motion_encoder import resnet18_alternative
    return resnet18_alternative


def cond_encoder_cls():
    install()
    from models.modules.autoencoders.fully_conv_models import ConvEncoder
    return ConvEncoder
