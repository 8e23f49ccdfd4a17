"""Generate the golden fixtures in tests/golden/*.pt by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py [--full]

For every case the synthetic state-dict from oracle/ipoke_oracle.py is loaded into the reference's own
modules with load_state_dict(strict=True) (so key names / shapes are proven identical to the reference
checkpoint layout), the reference is run on seeded inputs, and its outputs are stored.  The oracle is
compared against the reference in the same run and the max-abs differences are stored beside the outputs.
Inputs and weights are NOT stored: they regenerate from the seeds recorded in each fixture.
"""
import argparse
import os
import sys
import time
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_import  # noqa: E402
from oracle import ipoke_oracle as O  # noqa: E402

FLOW_CASES = {
    # name: (cfg kwargs, B, weight seed, input seed)
    "flow_tiny_even": (dict(flow_in_channels=16, flow_mid_channels=64, h_channels=16, num_steps=[2, 1, 1], factor=4), 3, 1, 11),
    "flow_tiny_odd": (dict(flow_in_channels=16, flow_mid_channels=48, h_channels=8, num_steps=[1, 1], factor=3), 2, 2, 12),
    "flow_c32_hd128": (dict(flow_in_channels=32, flow_mid_channels=128, h_channels=128), 2, 3, 13),
    "flow_c64_hd128": (dict(flow_in_channels=64, flow_mid_channels=128, h_channels=128, num_steps=[1] * 15), 2, 4, 14),
}
FULL_FLOW_CASES = {
    "flow_full_c32": (dict(flow_in_channels=32, flow_mid_channels=2048, h_channels=128), 2, 5, 42),
}
FS_CASES = {
    # name: (cfg kwargs, B, T, weight seed, input seed)
    "fs_64": (dict(z_dim=32, spatial=64), 2, 3, 21, 31),
    "fs_128": (dict(z_dim=32, spatial=128), 1, 2, 22, 32),
    "fs_64_z64": (dict(z_dim=64, spatial=64), 1, 2, 23, 33),
}


ENC_CASES = {
    # name: (cfg kwargs, B, T, weight seed, input seed)
    "enc_64": (dict(z_dim=32, img_size=64, max_frames=10), 2, 11, 41, 51),
    "enc_128": (dict(z_dim=32, img_size=128, max_frames=10), 1, 11, 42, 52),
}


def run_enc_case(name, spec, out_dir):
    kw, B, T, wseed, iseed = spec
    cfg = O.encoder_config(**kw)
    sd = O.synth_encoder_state_dict(cfg, seed=wseed)
    m = ref_import.encoder_fn()(dic=dict(cfg, ENC_M_channels=list(cfg["ENC_M_channels"])))
    m.load_state_dict(sd, strict=True)
    m.eval()
    g = torch.Generator().manual_seed(iseed)
    X = torch.rand((B, 3, T, cfg["img_size"], cfg["img_size"]), generator=g) * 2 - 1
    with torch.no_grad():
        torch.manual_seed(iseed + 1)            # the reference draws eps on the default CPU generator (motion_encoder.py:220)
        z_ref, mu_ref, lv_ref = m(X)
        torch.manual_seed(iseed + 1)
        eps = torch.FloatTensor(mu_ref.size()).normal_()
        z_or, mu_or, lv_or = O.encoder_forward(sd, cfg, X, eps)
        z64, mu64, lv64 = O.encoder_forward(sd, cfg, X.double(), eps.double())
    fix = dict(kind="encoder", cfg_kwargs=kw, B=B, T=T, wseed=wseed, iseed=iseed, z=z_ref.clone(), mu=mu_ref.clone(), logvar=lv_ref.clone(),
               eps=eps.clone(),
               oracle_vs_ref=max((z_or - z_ref).abs().max().item(), (mu_or - mu_ref).abs().max().item(), (lv_or - lv_ref).abs().max().item()),
               ref_fp32_vs_oracle_fp64=(mu64.float() - mu_ref).abs().max().item(), torch_version=torch.__version__)
    torch.save(fix, os.path.join(out_dir, name + ".pt"))
    print(f"{name}: z {tuple(z_ref.shape)} oracle-vs-ref {fix['oracle_vs_ref']:.2e} fp64 {fix['ref_fp32_vs_oracle_fp64']:.2e} "
          f"mu std {mu_ref.std().item():.3f} logvar std {lv_ref.std().item():.3f}")


CENC_CASES = {
    # name: (cfg kwargs, B, weight seed, input seed)
    "cenc_poke_128": (dict(nf_in=2, spatial=128), 2, 61, 71),
    "cenc_img_64": (dict(nf_in=3, spatial=64), 2, 62, 72),
}


def run_cenc_case(name, spec, out_dir):
    kw, B, wseed, iseed = spec
    cfg = O.cond_encoder_config(**kw)
    sd = O.synth_cond_encoder_state_dict(cfg, seed=wseed)
    m = ref_import.cond_encoder_cls()(nf_in=cfg["nf_in"], nf_max=cfg["nf_max"], n_stages=cfg["n_stages"], variational=False)
    m.load_state_dict(sd, strict=True)
    m.eval()
    g = torch.Generator().manual_seed(iseed)
    x = torch.rand((B, cfg["nf_in"], cfg["spatial"], cfg["spatial"]), generator=g) * 2 - 1
    with torch.no_grad():
        out_ref, mean_ref, logstd = m(x)
        out_or, mean_or = O.cond_encoder_forward(sd, cfg, x)
        out64, _ = O.cond_encoder_forward(sd, cfg, x.double())
    assert logstd is None
    fix = dict(kind="cond_encoder", cfg_kwargs=kw, B=B, wseed=wseed, iseed=iseed, out=out_ref.clone(), mean=mean_ref.clone(),
               oracle_vs_ref=max((out_or - out_ref).abs().max().item(), (mean_or - mean_ref).abs().max().item()),
               ref_fp32_vs_oracle_fp64=(out64.float() - out_ref).abs().max().item(), torch_version=torch.__version__)
    torch.save(fix, os.path.join(out_dir, name + ".pt"))
    print(f"{name}: out {tuple(out_ref.shape)} oracle-vs-ref {fix['oracle_vs_ref']:.2e} fp64 {fix['ref_fp32_vs_oracle_fp64']:.2e} out std {out_ref.std().item():.3f}")


I3D_CASES = {
    # name: (B, T, S, weight seed, input seed)
    "i3d_t10": (2, 10, 64, 61, 71),
    "i3d_t16": (1, 16, 128, 62, 72),
}


def run_i3d_case(name, spec, out_dir):
    """FVD chain: synthetic I3D weights -> reference I3D (strict load) on preprocess()ed clips; Frechet distance of
    seeded Gaussians through the reference's calculate_frechet_distance."""
    import numpy as np
    from oracle import fvd_oracle as FO
    M = ref_import.metrics_module()
    B, T, S, wseed, iseed = spec
    sd = FO.synth_i3d_state_dict(wseed)
    m = M.I3D(400, "rgb")
    m.load_state_dict(sd, strict=True)
    m.eval()
    g = torch.Generator().manual_seed(iseed)
    vid_a = torch.rand((B, T, 3, S, S), generator=g) * 2 - 1
    vid_b = torch.rand((B, T, 3, S, S), generator=g)                     # already in [0,1]: preprocess must not denorm it
    pa, pb = M.preprocess(vid_a, vid_b)
    act_ref = M.get_activations(pa, m, batch_size=B)
    with torch.no_grad():
        soft, logits = m(pb.permute(0, 2, 1, 3, 4))
    rng = np.random.RandomState(iseed)
    f1 = rng.randn(64, 12) * 1.5 + 0.3
    f2 = rng.randn(80, 12) @ (np.eye(12) + 0.2 * rng.randn(12, 12))
    fd_ref = M.calculate_frechet_distance(f1.mean(0), np.cov(f1, rowvar=False), f2.mean(0), np.cov(f2, rowvar=False))
    # oracle in the same run
    opa, opb = FO.preprocess(vid_a), FO.preprocess(vid_b)
    act_or = FO.activations(sd, opa, batch_size=B)
    fix = dict(B=B, T=T, S=S, wseed=wseed, iseed=iseed, act_a=torch.from_numpy(act_ref), logits_b=logits,
               pre_a_sum=pa.double().sum().item(), pre_b_sum=pb.double().sum().item(), f1=f1, f2=f2, fd=float(fd_ref),
               oracle_vs_ref=float(np.abs(act_or - act_ref).max()), logit_std=float(act_ref.std()))
    torch.save(fix, os.path.join(out_dir, name + ".pt"))
    print(f"{name}: logits std {fix['logit_std']:.3f} oracle-vs-ref {fix['oracle_vs_ref']:.2e} "
          f"pre {(opa - pa).abs().max().item():.1e} fd {fd_ref:.4f} vs oracle {FO.fvd_from_activations(f1, f2):.4f}")


GRAD_CASES = {
    # name: (cfg kwargs, B, weight seed, input seed)   second-stage training step: loss + gradients
    "flowgrad_tiny": (dict(flow_in_channels=16, flow_mid_channels=64, h_channels=16, num_steps=[2, 1, 1], factor=4), 3, 1, 11),
    "flowgrad_c32_hd128": (dict(flow_in_channels=32, flow_mid_channels=128, h_channels=128, num_steps=[2, 1, 1] + [1] * 12), 4, 3, 13),
    "flowgrad_c64_hd128": (dict(flow_in_channels=64, flow_mid_channels=128, h_channels=128, num_steps=[1] * 15), 2, 4, 14),
    # the actual h36m_128 shape of BASELINE configs[3] (C0 = 64, Hd = 2048, shipped num_steps: 1.24 B parameters); reference only
    # (`skip_oracle`: the oracle's second autograd graph would double the ~25 GB this case needs)
    "flowgrad_full_c64": (dict(flow_in_channels=64, flow_mid_channels=2048, h_channels=128), 2, 5, 15, "skip_oracle"),
}


def run_grad_case(name, spec, out_dir):
    """loss.backward() of the reference flow + FlowLoss on a seeded latent; all gradients are kept as (norm, 64 sampled entries)
    per tensor plus a handful in full, and compared with the oracle's autograd in the same run."""
    import importlib
    kw, B, wseed, iseed = spec[:4]
    skip_oracle = len(spec) > 4 and spec[4] == "skip_oracle"
    cfg = O.flow_config(**kw)
    sd = O.synth_flow_state_dict(cfg, seed=wseed)
    x, cond, _ = O.synth_inputs(B, cfg["flow_in_channels"], cfg["h_channels"], 8, seed=iseed)
    x = x * 0.8
    m = ref_flow(cfg, sd)
    ref_import.install()
    FlowLoss = importlib.import_module("models.modules.INN.loss").FlowLoss
    crit = FlowLoss(spatial_mean=False, logdet_weight=1.0)
    out, logdet = m(x, cond, reverse=False)
    loss, _ = crit(out, logdet)
    loss.backward()
    ref_grads = {k: p.grad.detach() for k, p in m.named_parameters() if p.grad is not None}
    keys = O.flow_trainable_keys(sd)
    assert set(ref_grads) == set(keys), (set(keys) ^ set(ref_grads))
    if skip_oracle:
        loss_or, worst = loss.detach(), float("nan")
    else:
        loss_or, grads_or = O.flow_loss_and_grads(sd, cfg, x, cond)
        worst = max(((grads_or[k] - ref_grads[k]).abs().max().item() / (ref_grads[k].abs().max().item() + 1e-12)) for k in keys)
    # compact summary: per tensor the L2 norm, the max-abs and NS sampled entries (flat index, value), stored as three arrays
    NS = 16
    gsel = torch.Generator().manual_seed(5)
    norms = torch.tensor([ref_grads[k].double().norm().item() for k in keys], dtype=torch.float64)
    maxabs = torch.tensor([ref_grads[k].abs().max().item() for k in keys], dtype=torch.float32)
    idx = torch.stack([torch.randint(0, ref_grads[k].numel(), (NS,), generator=gsel) for k in keys])
    val = torch.stack([ref_grads[k].flatten()[idx[i]] for i, k in enumerate(keys)])
    fix = dict(kind="flowgrad", cfg_kwargs=kw, B=B, wseed=wseed, iseed=iseed, loss=loss.item(), keys=keys, norms=norms, maxabs=maxabs, idx=idx, val=val,
               oracle_loss=loss_or.item(), oracle_vs_ref_worst_rel=worst, torch_version=torch.__version__)
    torch.save(fix, os.path.join(out_dir, name + ".pt"))
    print(f"{name}: loss {loss.item():.6f} oracle {loss_or.item():.6f}; {len(keys)} tensors, oracle-vs-ref worst relative grad error {worst:.2e}")


INIT_CASES = {
    # name: (cfg kwargs, B, construction seed, input seed)   data-dependent init of a FRESH reference flow (first train() forward)
    "flowinit_tiny": (dict(flow_in_channels=16, flow_mid_channels=64, h_channels=16, num_steps=[2, 1, 1], factor=4), 5, 7, 17),
    "flowinit_c32": (dict(flow_in_channels=32, flow_mid_channels=64, h_channels=8, num_steps=[1] * 15), 3, 8, 18),
}


def run_init_case(name, spec, out_dir):
    """A freshly constructed reference flow (every `initialized` buffer 0) runs its first density-direction forward in train() mode:
    ActNorm2dFlow.init (macow2.py:526-539) and Conv2dWeightNorm.init (macow_utils.py:231-246) fire layer by layer.  Stored: the
    state-dict before (fp16-exact weights are not needed -- it is small), every tensor the pass changed, and the outputs."""
    kw, B, cseed, iseed = spec
    cfg = O.flow_config(**kw)
    Flow = ref_import.flow_cls()
    torch.manual_seed(cseed)
    m = Flow(dict(cfg))
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(iseed)
    x = torch.randn((B, cfg["flow_in_channels"], 8, 8), generator=g) * 1.7 + 0.3
    cond = torch.randn((B, cfg["h_channels"], 8, 8), generator=g) * 0.5
    m.train()
    with torch.no_grad():
        z, logdet = m(x, cond, reverse=False)
        z2, logdet2 = m(x, cond, reverse=False)          # second call: already initialised, same result
    assert torch.equal(z, z2) and torch.equal(logdet, logdet2)
    sd1 = m.state_dict()
    changed = {k: v.detach().clone() for k, v in sd1.items() if not torch.equal(v, sd0[k])}
    n_flags = sum(1 for k in changed if k.endswith("initialized"))
    # the pass only depends on the ActNorm parameters and the channel permutations (every coupling / MCF becomes the identity: zero_init),
    # so only those are stored from the pre-init state; the test keeps its own random values for the conv weights
    sd0 = {k: v for k, v in sd0.items() if k.endswith(("log_scale", "shuffle_idx")) or (k.endswith(".bias") and "actnorm" in k)}
    fix = dict(kind="flowinit", cfg_kwargs=kw, B=B, cseed=cseed, iseed=iseed, sd0=sd0, changed=changed, x=x, cond=cond, z=z.clone(), logdet=logdet.clone(),
               torch_version=torch.__version__)
    torch.save(fix, os.path.join(out_dir, name + ".pt"))
    kinds = sorted({k.rsplit(".", 1)[-1] for k in changed})
    print(f"{name}: {len(sd0)} pre-init tensors kept, {len(changed)} changed ({n_flags} flags; kinds {kinds}); z std {z.std().item():.3f} logdet {logdet.tolist()}; "
          f"file {os.path.getsize(os.path.join(out_dir, name + '.pt')) / 1e6:.2f} MB")


def run_flowloss_case(out_dir):
    """FlowLoss.forward (loss.py:13-31) incl. the RNG it consumes: values of the log dict under torch.manual_seed(123) and the next
    randn draw after the call."""
    import importlib
    ref_import.install()
    FlowLoss = importlib.import_module("models.modules.INN.loss").FlowLoss
    g = torch.Generator().manual_seed(3)
    z = torch.randn((4, 16, 8, 8), generator=g) * 1.2
    logdet = torch.randn((4,), generator=g) * 5 - 20
    out = {}
    for sm in (False, True):
        torch.manual_seed(123)
        loss, log = FlowLoss(spatial_mean=sm, logdet_weight=1.0)(z, logdet)
        nxt = torch.randn(3)
        out[sm] = dict(loss=loss.item(), log={k: (v.item() if torch.is_tensor(v) else v) for k, v in log.items()}, next_randn=nxt)
    torch.save(dict(kind="flowloss", z=z, logdet=logdet, out=out, torch_version=torch.__version__), os.path.join(out_dir, "flowloss.pt"))
    print("flowloss:", out[False]["log"])


def ref_flow(cfg, sd):
    Flow = ref_import.flow_cls()
    m = Flow(dict(cfg))
    m.load_state_dict(sd, strict=True)
    return m.eval()


def run_flow_case(name, spec, out_dir):
    kw, B, wseed, iseed = spec
    cfg = O.flow_config(**kw)
    t0 = time.time()
    sd = O.synth_flow_state_dict(cfg, seed=wseed)
    z, cond, _ = O.synth_inputs(B, cfg["flow_in_channels"], cfg["h_channels"], 8, seed=iseed)
    m = ref_flow(cfg, sd)
    with torch.no_grad():
        x_ref = m(z, cond, reverse=True)                 # sampling direction
        z2_ref, ld_ref = m(x_ref, cond, reverse=False)   # density direction on the sampled latent
        x_or = O.flow_reverse(sd, cfg, z, cond)
        z2_or, ld_or = O.flow_forward(sd, cfg, x_ref, cond)
        x64 = O.flow_reverse(sd, cfg, z.double(), cond.double())
    fix = dict(kind="flow", cfg_kwargs=kw, B=B, wseed=wseed, iseed=iseed,
               x_rev=x_ref.clone(), z_fwd=z2_ref.clone(), logdet=ld_ref.clone(),
               oracle_vs_ref=dict(rev=(x_or - x_ref).abs().max().item(), fwd=(z2_or - z2_ref).abs().max().item(),
                                  logdet=(ld_or - ld_ref).abs().max().item()),
               ref_fp32_vs_oracle_fp64=(x64.float() - x_ref).abs().max().item(),
               roundtrip=(z2_ref - z).abs().max().item(),
               x_std=x_ref.std().item(), torch_version=torch.__version__)
    torch.save(fix, os.path.join(out_dir, name + ".pt"))
    print(f"{name}: {time.time() - t0:.1f}s  oracle-vs-ref {fix['oracle_vs_ref']}  fp64 {fix['ref_fp32_vs_oracle_fp64']:.2e} "
          f"roundtrip {fix['roundtrip']:.2e}  x std {fix['x_std']:.3f}  logdet {ld_ref.tolist()}")


def run_fs_case(name, spec, out_dir):
    kw, B, T, wseed, iseed = spec
    cfg = O.first_stage_config(**kw)
    sd = O.synth_first_stage_state_dict(cfg, seed=wseed)
    ConvGRU, Dec = ref_import.first_stage_parts()
    rnn = ConvGRU(input_size=cfg["z_dim"], hidden_sizes=cfg["z_dim"], n_layers=cfg["n_gru_layers"], kernel_sizes=3,
                  upsampling=[False] * cfg["n_gru_layers"])
    gen = Dec(dict(cfg))
    rnn.load_state_dict({k[len("rnn."):]: v for k, v in sd.items() if k.startswith("rnn.")}, strict=True)
    gen.load_state_dict({k[len("gen."):]: v for k, v in sd.items() if k.startswith("gen.")}, strict=True)
    rnn.eval(); gen.eval()
    g = torch.Generator().manual_seed(iseed)
    motion = torch.randn((B, cfg["z_dim"], 8, 8), generator=g) * 1.3
    x0 = torch.rand((B, 3, cfg["spatial"], cfg["spatial"]), generator=g) * 2 - 1
    with torch.no_grad():
        # decode_first_stage, second_stage_video.py:361-382
        hidden = [motion] * cfg["n_gru_layers"]
        in_rnn = torch.cat([sd["motion_bias"]] * B, dim=0)
        frames, hid_last = [], None
        for _ in range(T):
            hidden = rnn(in_rnn, hidden)
            frames.append(gen([hidden[-1]], x0, del_shape=True))
        ref = torch.stack(frames, dim=1)
        orc = O.decode_first_stage(sd, cfg, motion, x0, T)
        orc64 = O.decode_first_stage(sd, cfg, motion.double(), x0.double(), T)
    fix = dict(kind="first_stage", cfg_kwargs=kw, B=B, T=T, wseed=wseed, iseed=iseed, frames=ref.clone(),
               hidden_last=hidden[-1].clone(),
               oracle_vs_ref=(orc - ref).abs().max().item(),
               ref_fp32_vs_oracle_fp64=(orc64.float() - ref).abs().max().item(), torch_version=torch.__version__)
    torch.save(fix, os.path.join(out_dir, name + ".pt"))
    print(f"{name}: oracle-vs-ref {fix['oracle_vs_ref']:.2e}  fp64 {fix['ref_fp32_vs_oracle_fp64']:.2e} "
          f"frames std {ref.std().item():.3f} range [{ref.min().item():.3f},{ref.max().item():.3f}]")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also run the full-size (1.05 B parameter) flow case")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    torch.manual_seed(0)
    for n, s in FLOW_CASES.items():
        if a.only in (None, n):
            run_flow_case(n, s, HERE)
    for n, s in FS_CASES.items():
        if a.only in (None, n):
            run_fs_case(n, s, HERE)
    for n, s in CENC_CASES.items():
        if a.only in (None, n):
            run_cenc_case(n, s, HERE)
    for n, s in ENC_CASES.items():
        if a.only in (None, n):
            run_enc_case(n, s, HERE)
    for n, s in GRAD_CASES.items():
        if a.only in (None, n):
            run_grad_case(n, s, HERE)
    for n, s in I3D_CASES.items():
        if a.only in (None, n):
            run_i3d_case(n, s, HERE)
    for n, s in INIT_CASES.items():
        if a.only in (None, n):
            run_init_case(n, s, HERE)
    if a.only in (None, "flowloss"):
        run_flowloss_case(HERE)
    if a.full:
        for n, s in FULL_FLOW_CASES.items():
            if a.only in (None, n):
                run_flow_case(n, s, HERE)
