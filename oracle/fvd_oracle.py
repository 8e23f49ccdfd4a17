"""TEST INFRASTRUCTURE -- restatement of the reference's in-repo Frechet-video-distance chain (torch ops, any device).

Only tests/ (incl. the sweep tests/fvd_parity.py) may import this; the product never does.

What "FVD" means in this repo: Frechet distance (utils/metrics.py:625-678) between the 400-d
logits (get_activations takes output [1], utils/metrics.py:726) of the reference's PyTorch I3D (utils/metrics.py:999-1105)
on videos resized to 224x224 and mapped to [0,1] (preprocess, utils/metrics.py:786-802).  No pretrained I3D weights ship
with the reference (utils/metrics.py:806 points at a path outside the repo), so the weights here are seeded synthetic
ones (He-normal so that the logits keep O(1) scale through 22 conv layers).

Pinned by tests/golden/i3d_*.pt: the synthetic state-dict is loaded with strict=True into the UNMODIFIED reference I3D,
run on seeded clips, and the logits / preprocess outputs / Frechet distances are stored (tests/golden/make_golden.py).
"""
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# (in_channels, [b0, b1a, b1b, b2a, b2b, b3]) of the nine Inception blocks, utils/metrics.py:1047-1066
MIXED = [
    ("mixed_3b", 192, [64, 96, 128, 16, 32, 32]),
    ("mixed_3c", 256, [128, 128, 192, 32, 96, 64]),
    ("mixed_4b", 480, [192, 96, 208, 16, 48, 64]),
    ("mixed_4c", 512, [160, 112, 224, 24, 64, 64]),
    ("mixed_4d", 512, [128, 128, 256, 24, 64, 64]),
    ("mixed_4e", 512, [112, 144, 288, 32, 64, 64]),
    ("mixed_4f", 528, [256, 160, 320, 32, 128, 128]),
    ("mixed_5b", 832, [256, 160, 320, 32, 128, 128]),
    ("mixed_5c", 832, [384, 192, 384, 48, 128, 128]),
]


def i3d_units(num_classes: int = 400) -> List[Tuple[str, int, int, Tuple[int, int, int], bool, bool]]:
    """(state-dict prefix, cin, cout, kernel, use_bn, use_bias) of every Unit3Dpy, in module order."""
    u = [("conv3d_1a_7x7", 3, 64, (7, 7, 7), True, False),
         ("conv3d_2b_1x1", 64, 64, (1, 1, 1), True, False),
         ("conv3d_2c_3x3", 64, 192, (3, 3, 3), True, False)]
    for name, cin, o in MIXED:
        u += [(f"{name}.branch_0", cin, o[0], (1, 1, 1), True, False),
              (f"{name}.branch_1.0", cin, o[1], (1, 1, 1), True, False),
              (f"{name}.branch_1.1", o[1], o[2], (3, 3, 3), True, False),
              (f"{name}.branch_2.0", cin, o[3], (1, 1, 1), True, False),
              (f"{name}.branch_2.1", o[3], o[4], (3, 3, 3), True, False),
              (f"{name}.branch_3.1", cin, o[5], (1, 1, 1), True, False)]
    u.append(("conv3d_0c_1x1", 1024, num_classes, (1, 1, 1), False, True))
    return u


def synth_i3d_state_dict(seed: int = 0, num_classes: int = 400) -> Dict[str, Tensor]:
    """Seeded synthetic I3D weights with the reference's exact key names (Unit3Dpy: conv3d / batch3d, :884-918)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for p, cin, cout, k, bn, bias in i3d_units(num_classes):
        fan_in = cin * k[0] * k[1] * k[2]
        sd[f"{p}.conv3d.weight"] = torch.randn((cout, cin, *k), generator=g) * (2.0 / fan_in) ** 0.5
        if bias:
            sd[f"{p}.conv3d.bias"] = torch.randn((cout,), generator=g) * 0.1
        if bn:
            sd[f"{p}.batch3d.weight"] = 1.0 + 0.1 * torch.randn((cout,), generator=g)
            sd[f"{p}.batch3d.bias"] = 0.1 * torch.randn((cout,), generator=g)
            sd[f"{p}.batch3d.running_mean"] = 0.1 * torch.randn((cout,), generator=g)
            sd[f"{p}.batch3d.running_var"] = 0.5 + torch.rand((cout,), generator=g)
            sd[f"{p}.batch3d.num_batches_tracked"] = torch.tensor(1, dtype=torch.long)
    return sd


def tf_same_pad(kernel, stride, mod: int = 0) -> Tuple[int, ...]:
    """get_padding_shape (utils/metrics.py:814-842): per dim pad_along = max(k - s, 0) (depth: k - mod when mod != 0),
    split low/high, then the depth pair is moved to the END of the tuple -- which ConstantPad3d / F.pad read as
    (last-dim lo, hi, middle lo, hi, depth lo, hi).  Kept as is (H and W pads swap, all kernels here are square)."""
    out = []
    for i, (k, s) in enumerate(zip(kernel, stride)):
        m = mod if i == 0 else 0
        along = max(k - m, 0) if m else max(k - s, 0)
        lo = along // 2
        out += [lo, along - lo]
    return tuple(out[2:] + out[:2])


def unit3d(sd, p: str, x: Tensor, kernel, stride=(1, 1, 1), relu=True, bn=True) -> Tensor:
    """Unit3Dpy.forward (utils/metrics.py:926-937): TF-SAME padded Conv3d -> BatchNorm3d(eval, eps 1e-3) -> ReLU."""
    pad = tf_same_pad(kernel, stride)
    w = sd[p + ".conv3d.weight"]
    b = sd.get(p + ".conv3d.bias")
    if all(v == pad[0] for v in pad):                       # simplify_padding :845-851 -> symmetric conv padding
        out = F.conv3d(x, w, b, stride=stride, padding=pad[0])
    else:                                                    # runtime pad chosen by T mod stride_T (:927-931)
        pad = tf_same_pad(kernel, stride, x.shape[2] % stride[0]) if stride[0] > 1 else pad
        out = F.conv3d(F.pad(x, pad), w, b, stride=stride)
    if bn:
        out = F.batch_norm(out, sd[p + ".batch3d.running_mean"], sd[p + ".batch3d.running_var"],
                           sd[p + ".batch3d.weight"], sd[p + ".batch3d.bias"], training=False, eps=1e-3)
    return F.relu(out) if relu else out


def maxpool_tf(x: Tensor, kernel, stride) -> Tensor:
    """MaxPool3dTFPadding.forward (utils/metrics.py:955-960): zero pad (index T mod stride_T), MaxPool3d(ceil_mode=True)."""
    pad = tf_same_pad(kernel, stride, x.shape[2] % stride[0]) if stride[0] > 1 else tf_same_pad(kernel, stride)
    return F.max_pool3d(F.pad(x, pad), kernel, stride, ceil_mode=True)


def mixed(sd, p: str, x: Tensor) -> Tensor:
    """Mixed.forward (utils/metrics.py:991-997)."""
    b0 = unit3d(sd, p + ".branch_0", x, (1, 1, 1))
    b1 = unit3d(sd, p + ".branch_1.1", unit3d(sd, p + ".branch_1.0", x, (1, 1, 1)), (3, 3, 3))
    b2 = unit3d(sd, p + ".branch_2.1", unit3d(sd, p + ".branch_2.0", x, (1, 1, 1)), (3, 3, 3))
    b3 = unit3d(sd, p + ".branch_3.1", maxpool_tf(x, (3, 3, 3), (1, 1, 1)), (1, 1, 1))
    return torch.cat((b0, b1, b2, b3), 1)


def i3d_logits(sd, x: Tensor) -> Tensor:
    """I3D.forward (utils/metrics.py:1079-1105), second output (logits).  x: [B, 3, T, 224, 224] in [0, 1]."""
    out = unit3d(sd, "conv3d_1a_7x7", x, (7, 7, 7), (2, 2, 2))
    out = maxpool_tf(out, (1, 3, 3), (1, 2, 2))
    out = unit3d(sd, "conv3d_2b_1x1", out, (1, 1, 1))
    out = unit3d(sd, "conv3d_2c_3x3", out, (3, 3, 3))
    out = maxpool_tf(out, (1, 3, 3), (1, 2, 2))
    out = mixed(sd, "mixed_3b", out)
    out = mixed(sd, "mixed_3c", out)
    out = maxpool_tf(out, (3, 3, 3), (2, 2, 2))
    for n in ("mixed_4b", "mixed_4c", "mixed_4d", "mixed_4e", "mixed_4f"):
        out = mixed(sd, n, out)
    out = maxpool_tf(out, (2, 2, 2), (2, 2, 2))
    out = mixed(sd, "mixed_5b", out)
    out = mixed(sd, "mixed_5c", out)
    out = F.avg_pool3d(out, (2, 7, 7), (1, 1, 1))
    out = unit3d(sd, "conv3d_0c_1x1", out, (1, 1, 1), relu=False, bn=False)
    return out.squeeze(3).squeeze(3).mean(2)


def preprocess(videos: Tensor) -> Tensor:
    """preprocess (utils/metrics.py:786-802) for one set: [N,T,3,H,W] -> bilinear 224x224 (align_corners) -> [0,1] when
    the set has negative values."""
    v = F.interpolate(videos.reshape(-1, *videos.shape[2:]), mode="bilinear", size=(224, 224), align_corners=True)
    v = v.reshape(*videos.shape[:2], 3, 224, 224)
    if v.min() < 0:
        v = (v + 1.0) / 2.0
    return v


def activations(sd, videos: Tensor, batch_size: int = 50) -> np.ndarray:
    """get_activations (utils/metrics.py:681-733): logits of full batches only (the remainder is dropped, :709-710),
    input permuted to [B,3,T,H,W] (:726).  `videos` already preprocessed, on the device of `sd`."""
    n = videos.shape[0]
    batch_size = min(batch_size, n)
    nb = n // batch_size
    out = np.empty((nb * batch_size, 400))
    with torch.no_grad():
        for i in range(nb):
            b = videos[i * batch_size:(i + 1) * batch_size]
            out[i * batch_size:(i + 1) * batch_size] = i3d_logits(sd, b.permute(0, 2, 1, 3, 4)).double().cpu().numpy().reshape(batch_size, -1)
    return out


def moments(act: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """calculate_activation_statistics tail (utils/metrics.py:765-770): drop all-NaN rows, mean and np.cov."""
    act = act[np.flatnonzero(np.logical_not(np.isnan(act)).any(axis=-1))]
    return np.mean(act, axis=0), np.cov(act, rowvar=False)


def frechet_distance(mu1, sigma1, mu2, sigma2, eps: float = 1e-6) -> float:
    """calculate_frechet_distance (utils/metrics.py:625-678): |mu1-mu2|^2 + Tr(S1 + S2 - 2 sqrt(S1 S2))."""
    from scipy import linalg
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape and sigma1.shape == sigma2.shape
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if isinstance(covmean, tuple):
        covmean = covmean[0]
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


def fvd_from_activations(a1: np.ndarray, a2: np.ndarray) -> float:
    """calculate_FVD (utils/metrics.py:773-780) after the activations."""
    m1, s1 = moments(a1)
    m2, s2 = moments(a2)
    return frechet_distance(m1, s1, m2, s2)


def synth_pokes(B: int, spatial: int, seed: int, poke_size: int = 5, max_mag: float = 1.0):
    """One synthetic poke per sample in the layout of data/base_dataset.py:612-648: a zero map
    [B,2,S,S] with one poke_size x poke_size patch holding a constant flow vector, centred in [poke_size, S - poke_size);
    centres returned as [B,1,2]."""
    g = torch.Generator().manual_seed(seed)
    poke = torch.zeros((B, 2, spatial, spatial))
    cy = torch.randint(poke_size, spatial - poke_size, (B,), generator=g)
    cx = torch.randint(poke_size, spatial - poke_size, (B,), generator=g)
    vec = (torch.rand((B, 2), generator=g) * 2 - 1) * max_mag
    h = poke_size // 2
    for b in range(B):
        poke[b, :, cy[b] - h:cy[b] + h + 1, cx[b] - h:cx[b] + h + 1] = vec[b].view(2, 1, 1)
    return poke, torch.stack([cy, cx], dim=1).unsqueeze(1)
