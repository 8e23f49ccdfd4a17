"""Times the full-size C0 = 64 flow (plants_128 / h36m shapes: 1.24 B parameters) in both directions with the per-phase split:
python profiles/flow_c64_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ipoke_b200 as ipk
from ipoke_b200 import synth
from oracle import ipoke_oracle as O

dev = torch.device("cuda:0")
B = int(os.environ.get("PROBE_BATCH", "64"))
C0 = int(os.environ.get("PROBE_C0", "64"))
fcfg = dict(O.flow_config(flow_in_channels=C0), ipk_precision="fp32", ipk_max_batch=B)
with torch.device(dev):
    flow = ipk.SupervisedMacowTransformer(fcfg)
flow = synth.fill_flow_(flow.to(dev).eval(), seed=0)
z = torch.randn((B, C0, 8, 8)).to(dev)
cond = (torch.randn((B, 128, 8, 8)) * 0.5).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("reverse (sampling)", lambda: flow(z, cond, reverse=True)), ("forward + logdet (density)", lambda: flow(z, cond))):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ipk._lib.prof_enable(True)
    fn()
    rep = ipk._lib.prof_report()
    ipk._lib.prof_enable(False)
    print(f"flow C0={C0} {name}: B={B} {ms:.2f} ms/step ({B / ms * 1e3:.0f} samples/s); " + ", ".join(f"{k} {t:.2f}" for k, (c, t) in rep.items()), flush=True)
x = flow(z, cond, reverse=True)
z2, ld = flow(x, cond)
print(f"round trip max-abs {(z2 - z).abs().max().item():.2e}, logdet mean {ld.mean().item():.3f}")
