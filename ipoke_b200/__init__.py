"""ipoke_b200: B200-native stochastic-video sampling hot path of CompVis/ipoke (conditional MaCow flow inverse ->
latent ConvGRU + SPADE decoder) behind the reference's module interfaces.  All compute is hand-written sm_100a CUDA in
libipoke_b200.so; there is no CPU / PyTorch fallback."""
from . import _lib  # noqa: F401
from .cond_encoder import ConvEncoder, make_cond  # noqa: F401
from .encoder import ResNetMotionEncoder, encode_first_stage  # noqa: F401
from .first_stage import SpadeCondMotionDecoder, decode_first_stage  # noqa: F401
from .flow import SupervisedMacowTransformer, FlowLoss, flow_nll  # noqa: F401
from . import i3d  # noqa: F401
from .i3d import I3D  # noqa: F401
from .parallel import shard_bounds, sharded_sample, global_noise  # noqa: F401
from .sampler import PokeMotionSampler  # noqa: F401

__version__ = "0.1.0"
from .train import FlowDensityFunction, FlowTrainer, shard_range, sharded_update  # noqa: F401
