"""Process-local shims that let the UNMODIFIED reference hot-path sub-modules import on CPU.

Only used by tests/golden/make_golden.py (fixture generation, run in the build container where
/root/reference exists).  Nothing here is imported by the product, by `-m gpu` tests, or by bench.py.

Shims:
  * `opt_einsum` is imported by models/modules/INN/modules.py:9 but only used by attention blocks
    (attention: false in every shipped config) -> empty stub module.
  * `utils/metrics.py` imports pytorch_lightning.metrics(.functional), lpips and utils.logging (:3,4,14,17), none of
    which the FVD chain uses -> stub modules; `scipy.linalg.sqrtm(disp=False)` (:660) was removed from SciPy ->
    a wrapper restoring the (result, errest) return.
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


def metrics_module():
    """utils.metrics of the reference (I3D, preprocess, get_activations, calculate_frechet_distance)."""
    install()
    import scipy.linalg
    for name in ("pytorch_lightning", "pytorch_lightning.metrics", "pytorch_lightning.metrics.functional", "lpips"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch_lightning.metrics"].Metric = object
    sys.modules["pytorch_lightning.metrics.functional"].ssim = None
    sys.modules["pytorch_lightning.metrics.functional"].psnr = None
    sys.modules["lpips"].LPIPS = object
    import utils as ref_utils                       # the reference's package (REF is first on sys.path)
    if "utils.logging" not in sys.modules:
        m = types.ModuleType("utils.logging")
        m.make_nn_var_plot = None
        sys.modules["utils.logging"] = m
        ref_utils.logging = m
    if not getattr(scipy.linalg, "_ipk_sqrtm_shim", False):
        _sqrtm = scipy.linalg.sqrtm
        def sqrtm(a, disp=True, **k):
            r = _sqrtm(a, **k)
            return r if disp else (r, 0.0)
        scipy.linalg.sqrtm = sqrtm
        scipy.linalg._ipk_sqrtm_shim = True
    import utils.metrics as M
    return M
