"""FVD parity sweep (BASELINE.json configs[4]: plants_128 diversity, n_samples per poke x synthetic pokes).

    python tests/fvd_parity.py [--pokes 1000 --samples 5 --frames 10 --spatial 128 --c0 64 --hd 2048] [--out profiles/x.json]

For every synthetic poke p (image x0_p ~ U(-1,1), one 5x5 poke patch) the same `n_samples` latents
z ~ N(0,I) (CPU generator seeded 42 + p, second_stage_video.py:289-300) go through

  * the product path: native ConvEncoders (make_cond, once per poke batch) -> ipk_sample (flow inverse -> GRU + decoder),
  * the reference path: the oracle restatement (oracle/ipoke_oracle.py) of the same modules, fp32 torch ops with TF32
    off, on `--oracle-device` (cuda for the full sweep: the checker is allowed to be fast; cpu for the small test),

with identical synthetic weights.  Both video sets are reduced on the fly to the 400-d I3D logits of the reference's FVD
chain with a seeded I3D -- by default through the NATIVE I3D (ipoke_b200.i3d, itself pinned on the reference's logits by
tests/test_gpu_i3d.py), `native_i3d=False` / `--oracle-i3d` through the oracle restatement (oracle/fvd_oracle.py).  Reported:
  fvd_ours_vs_ref        Frechet distance between the two sets on identical seeds/pokes (0 for identical videos),
  fvd_ref_split          reference(even pokes) vs reference(odd pokes): the ref-vs-ref baseline,
  fvd_ours_split         ours(even pokes) vs reference(odd pokes);  parity = |fvd_ours_split - fvd_ref_split| <= 1.0,
  max_abs_frames         per-frame max-abs between the two paths over the whole sweep.
This file is test infrastructure: it is the only place besides tests/ that runs the oracle next to the product.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import fvd_oracle as FO  # noqa: E402
from oracle import ipoke_oracle as O  # noqa: E402


def _to(sd, dev):
    return {k: v.to(dev) for k, v in sd.items()}


def run_sweep(n_pokes=1000, n_samples=5, frames=10, spatial=128, c0=64, hd=2048, batch_pokes=50, precision="fp32",
              oracle_device="cuda", i3d_batch=50, num_steps=None, verbose=True, native_i3d=True):
    import ipoke_b200 as ipk
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    odev = torch.device(oracle_device)
    B = batch_pokes * n_samples
    kw = dict(flow_in_channels=c0, flow_mid_channels=hd, h_channels=128)
    if num_steps is not None:
        kw["num_steps"] = num_steps
    fcfg = O.flow_config(**kw)
    dcfg = O.first_stage_config(z_dim=c0, spatial=spatial)
    icfg, pcfg = O.cond_encoder_config(nf_in=3, spatial=spatial), O.cond_encoder_config(nf_in=2, spatial=spatial)
    fsd, dsd = O.synth_flow_state_dict(fcfg, seed=0), O.synth_first_stage_state_dict(dcfg, seed=1)
    isd, psd = O.synth_cond_encoder_state_dict(icfg, seed=2), O.synth_cond_encoder_state_dict(pcfg, seed=3)
    i3d = _to(FO.synth_i3d_state_dict(seed=4), dev)
    i3d_native = None
    if native_i3d:
        i3d_native = ipk.I3D(400, "rgb", ipk_max_batch=i3d_batch, ipk_max_frames=max(frames, 9))
        i3d_native.load_state_dict(FO.synth_i3d_state_dict(seed=4), strict=True)
        i3d_native = i3d_native.to(dev).eval()

    fc = dict(fcfg); fc.update(ipk_precision=precision, ipk_max_batch=B)
    dc = dict(dcfg); dc.update(ipk_precision=precision, ipk_max_batch=B, ipk_max_frames=frames)
    flow = ipk.SupervisedMacowTransformer(fc); flow.load_state_dict(fsd, strict=True)
    fs = ipk.SpadeCondMotionDecoder(dc); fs.load_state_dict(dsd, strict=True)
    img = ipk.ConvEncoder(3, 64, icfg["n_stages"], ipk_max_batch=batch_pokes); img.load_state_dict(isd, strict=True)
    pk = ipk.ConvEncoder(2, 64, pcfg["n_stages"], ipk_max_batch=batch_pokes); pk.load_state_dict(psd, strict=True)
    sampler = ipk.PokeMotionSampler(flow.to(dev).eval(), fs.to(dev).eval(), img.to(dev).eval(), pk.to(dev).eval())
    o_fsd, o_dsd, o_isd, o_psd = _to(fsd, odev), _to(dsd, odev), _to(isd, odev), _to(psd, odev)

    def features(videos):
        out = []
        for i in range(0, videos.shape[0], i3d_batch):
            if i3d_native is not None:
                v = ipk.i3d._preprocess_one(videos[i:i + i3d_batch].to(dev))
                out.append(ipk.i3d.get_activations(v, i3d_native, batch_size=v.shape[0]))
                continue
            v = FO.preprocess(videos[i:i + i3d_batch].to(dev))
            out.append(FO.activations(i3d, v, batch_size=v.shape[0]))
        return np.concatenate(out, 0)

    f_ours, f_ref, poke_id = [], [], []
    max_abs, t_ours, t_ref = 0.0, 0.0, 0.0
    for p0 in range(0, n_pokes, batch_pokes):
        npk = min(batch_pokes, n_pokes - p0)
        g = torch.Generator().manual_seed(1000 + p0)
        x0 = torch.rand((npk, 3, spatial, spatial), generator=g) * 2 - 1
        poke, _ = FO.synth_pokes(npk, spatial, seed=2000 + p0)
        z = torch.cat([torch.randn((n_samples, c0, 8, 8), generator=torch.Generator().manual_seed(42 + p0 + i)) for i in range(npk)])
        # -- product path: conditioning encoded once per poke, repeated over its samples
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cond = sampler.make_cond(x0.to(dev), poke.to(dev)).repeat_interleave(n_samples, dim=0)
        ours = sampler.sample(z.to(dev), cond, x0.to(dev).repeat_interleave(n_samples, dim=0), frames)
        torch.cuda.synchronize(); t_ours += time.perf_counter() - t0
        # -- reference path (oracle)
        t0 = time.perf_counter()
        with torch.no_grad():
            x0o, pko = x0.to(odev), poke.to(odev)
            cond_o = torch.cat([O.cond_encoder_forward(o_isd, icfg, x0o)[0], O.cond_encoder_forward(o_psd, pcfg, pko)[0]], dim=1)
            ref = O.sample_videos(o_fsd, fcfg, o_dsd, dcfg, z.to(odev), cond_o.repeat_interleave(n_samples, dim=0),
                                  x0o.repeat_interleave(n_samples, dim=0), frames)
        if odev.type == "cuda":
            torch.cuda.synchronize()
        t_ref += time.perf_counter() - t0
        max_abs = max(max_abs, (ours - ref.to(dev)).abs().max().item())
        f_ours.append(features(ours)); f_ref.append(features(ref))
        poke_id.append(np.repeat(np.arange(p0, p0 + npk), n_samples))
        if verbose:
            print(f"pokes {p0}..{p0 + npk}: max-abs {max_abs:.2e}  ours {t_ours:.1f}s  oracle {t_ref:.1f}s", flush=True)
    f_ours, f_ref, poke_id = np.concatenate(f_ours), np.concatenate(f_ref), np.concatenate(poke_id)
    even, odd = poke_id % 2 == 0, poke_id % 2 == 1
    res = {
        "config": f"C0={c0} Hd={hd} {frames}-frame {spatial}x{spatial}, {n_pokes} pokes x {n_samples} samples, precision {precision}",
        "videos_per_set": int(f_ours.shape[0]), "feature_dim": int(f_ours.shape[1]),
        "max_abs_frames": max_abs, "max_abs_features": float(np.abs(f_ours - f_ref).max()),
        "feature_std": float(f_ref.std(axis=0).mean()),
        "fvd_ours_vs_ref": FO.fvd_from_activations(f_ours, f_ref),
        "fvd_ref_split": FO.fvd_from_activations(f_ref[even], f_ref[odd]),
        "fvd_ours_split": FO.fvd_from_activations(f_ours[even], f_ref[odd]),
        "seconds_ours": t_ours, "seconds_oracle": t_ref, "oracle_device": str(odev), "i3d": "native (ipoke_b200.i3d)" if native_i3d else "oracle (torch ops)",
    }
    res["delta_fvd"] = res["fvd_ours_split"] - res["fvd_ref_split"]
    res["fvd_parity"] = bool(abs(res["delta_fvd"]) <= 1.0 and abs(res["fvd_ours_vs_ref"]) <= 1.0)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--pokes", type=int, default=1000)
    ap.add_argument("--samples", type=int, default=5)
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--spatial", type=int, default=128)
    ap.add_argument("--c0", type=int, default=64)
    ap.add_argument("--hd", type=int, default=2048)
    ap.add_argument("--batch-pokes", type=int, default=50)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--oracle-device", default="cuda")
    ap.add_argument("--out", default=None)
    ap.add_argument("--oracle-i3d", action="store_true", help="reduce the videos with the oracle's torch-op I3D instead of the native one")
    a = ap.parse_args()
    r = run_sweep(a.pokes, a.samples, a.frames, a.spatial, a.c0, a.hd, a.batch_pokes, a.precision, a.oracle_device, native_i3d=not a.oracle_i3d)
    print(json.dumps(r))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(r, f, indent=1)
