"""Times the 3-D conv video encoder (training path) and the flow density direction on one GPU: python profiles/encoder_probe.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ipoke_b200 as ipk
from ipoke_b200 import synth
from oracle import ipoke_oracle as O

dev = torch.device("cuda:0")
B = int(os.environ.get("PROBE_BATCH", "32"))
cfg = O.encoder_config(z_dim=32, img_size=128, max_frames=10)
X = (torch.rand((B, 3, 11, 128, 128)) * 2 - 1).to(dev)
eps = torch.randn((B, 32, 8, 8))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for prec in ("fp32_simt", "fp32", "bf16"):
    enc = ipk.ResNetMotionEncoder(dict(cfg, ipk_max_batch=B, ipk_precision=prec))
    enc.load_state_dict(O.synth_encoder_state_dict(cfg, seed=1))
    enc = enc.to(dev).eval()
    for _ in range(2):
        z, mu, lv = enc(X, eps=eps)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        z, mu, lv = enc(X, eps=eps)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ipk._lib.prof_enable(True)
    enc(X, eps=eps)
    rep = ipk._lib.prof_report()
    ipk._lib.prof_enable(False)
    print(f"encoder [{prec}]: B={B} 11x128x128 -> {ms:.2f} ms/step, {B / ms * 1e3:.1f} clips/s, {69.73 * B / ms:.2f} TFLOP/s algorithmic; "
          + ", ".join(f"{k} {t:.2f} ms ({c})" for k, (c, t) in rep.items()), flush=True)
    del enc
B = min(B, 16)
z = z[:B]

fcfg = dict(O.flow_config(), ipk_precision="fp32", ipk_max_batch=B)
with torch.device(dev):
    flow = ipk.SupervisedMacowTransformer(fcfg)
flow = synth.fill_flow_(flow.to(dev).eval(), seed=0)
cond = (torch.randn((B, 128, 8, 8)) * 0.5).to(dev)
for _ in range(2):
    zz, ld = flow(z, cond)
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    zz, ld = flow(z, cond)
e1.record(); torch.cuda.synchronize()
ms2 = e0.elapsed_time(e1) / 3
print(f"flow forward + logdet (density direction): B={B} -> {ms2:.2f} ms/step; nll {ipk.flow_nll(zz, ld).item():.3f}")
