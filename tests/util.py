import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ipoke_oracle as O  # noqa: E402


def ref_flow_cfg(cfg, precision, max_batch=8):
    """oracle flow cfg -> config dict of the drop-in module"""
    c = dict(cfg)
    c["ipk_precision"] = precision
    c["ipk_max_batch"] = max_batch
    return c


def make_flow(cfg, sd, precision, device="cuda:0", max_batch=8):
    import ipoke_b200 as ipk
    m = ipk.SupervisedMacowTransformer(ref_flow_cfg(cfg, precision, max_batch))
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def make_first_stage(cfg, sd, precision, device="cuda:0", max_batch=4, max_frames=4, chunk_videos=0):
    import ipoke_b200 as ipk
    c = dict(cfg)
    c.update(ipk_precision=precision, ipk_max_batch=max_batch, ipk_max_frames=max_frames, ipk_chunk_videos=chunk_videos)
    m = ipk.SpadeCondMotionDecoder(c)
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def maxabs(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item()
