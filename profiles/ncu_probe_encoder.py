"""ncu target: one pass of the video encoder (tcgen05 Conv3d + GroupNorm passes) and the uint8 post-processing kernel.
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_enc python profiles/ncu_probe_encoder.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ipoke_b200 as ipk
from oracle import ipoke_oracle as O

dev = torch.device("cuda:0")
B = 32
cfg = O.encoder_config(z_dim=32, img_size=128, max_frames=10)
enc = ipk.ResNetMotionEncoder(dict(cfg, ipk_max_batch=B, ipk_precision="fp32"))
enc.load_state_dict(O.synth_encoder_state_dict(cfg, seed=1))
enc = enc.to(dev).eval()
X = (torch.rand((B, 3, 11, 128, 128)) * 2 - 1).to(dev)
eps = torch.randn((B, 32, 8, 8))
frames = torch.tanh(torch.randn((64, 16, 3, 128, 128), device=dev))
for _ in range(2):
    enc(X, eps=eps)
    ipk.PokeMotionSampler.to_uint8(frames)
torch.cuda.synchronize()
torch.cuda.profiler.start()
enc(X, eps=eps)
ipk.PokeMotionSampler.to_uint8(frames)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
