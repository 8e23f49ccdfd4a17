"""Second-stage training step (BASELINE configs[3]: h36m shapes, C0 = 64, Hd = 2048, 1.24 B parameters) on 1..N GPUs.

    python profiles/train_probe.py [--batch 32] [--steps 3] [--small]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/train_probe.py

One step = video clip -> 3-D conv encoder (no grad) -> flow forward + log-det -> FlowLoss -> backward -> reduce-scatter of the flat
gradient -> Adam(amsgrad) on each rank's shard -> all-gather of the parameters.  Prints one JSON line on rank 0.  With --check (small
model) every rank also verifies that the sharded update equals a single-process update on the global batch."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import ipoke_b200 as ipk
from ipoke_b200 import synth
from oracle import ipoke_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32, help="samples per GPU")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--small", action="store_true", help="Hd = 128 flow (plumbing check)")
ap.add_argument("--check", action="store_true", help="compare the sharded update with a single-process update on the global batch (use with --small)")
ap.add_argument("--precision", default="fp32")
ap.add_argument("--ncu", action="store_true", help="one warm step, then one step between cudaProfilerStart/Stop (run under ncu --profile-from-start off)")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, C0 = a.batch, 64
fcfg = dict(O.flow_config(flow_in_channels=C0, flow_mid_channels=128 if a.small else 2048), ipk_precision=a.precision, ipk_max_batch=B * (world if a.check else 1))
with torch.device(dev):
    flow = ipk.SupervisedMacowTransformer(fcfg)
flow = synth.fill_flow_(flow.to(dev).eval(), seed=0)
ecfg = O.encoder_config(z_dim=C0, img_size=128, max_frames=10)
enc = ipk.ResNetMotionEncoder(dict(ecfg, ipk_max_batch=B, ipk_precision=a.precision))
enc.load_state_dict(O.synth_encoder_state_dict(ecfg, seed=1))
enc = enc.to(dev).eval()
g = torch.Generator().manual_seed(7)
Xg = torch.rand((B * world, 11, 3, 128, 128), generator=g) * 2 - 1          # global batch, sliced per rank
condg = torch.randn((B * world, 128, 8, 8), generator=g) * 0.5
epsg = torch.randn((B * world, C0, 8, 8), generator=g)
sl = slice(rank * B, (rank + 1) * B)
X, cond, eps = Xg[sl].to(dev), condg[sl].to(dev), epsg[sl].to(dev)
ref_after = None
if a.check:
    # single-process reference: same model, global batch, unsharded Adam
    with torch.device(dev):
        flow2 = ipk.SupervisedMacowTransformer(fcfg)
    flow2.load_state_dict(flow.state_dict())
    tr2 = ipk.FlowTrainer(flow2.to(dev).eval(), max_batch=B * world, precision=a.precision, distributed=False)
tr = ipk.FlowTrainer(flow, max_batch=B * (world if a.check else 1), precision=a.precision)
torch.cuda.synchronize()

def one_step(trainer, Xs, conds, epss):
    z_in, _ = ipk.encode_first_stage(enc, Xs, eps=epss)
    loss = trainer.step(z_in, conds)
    trainer.optimizer_step(lr=1e-4, betas=(0.9, 0.99), amsgrad=True)
    return loss

if a.check:
    zs = []
    for r in range(world):                                                   # encoder batch stays <= B
        zr, _ = ipk.encode_first_stage(enc, Xg[r * B:(r + 1) * B].to(dev), eps=epsg[r * B:(r + 1) * B].to(dev))
        zs.append(zr)
    loss2 = tr2.step(torch.cat(zs), condg.to(dev))
    tr2.optimizer_step(lr=1e-4, betas=(0.9, 0.99), amsgrad=True)
    loss1 = one_step(tr, X, cond, eps)
    lg = loss1.clone()
    if world > 1:
        dist.all_reduce(lg); lg /= world
    # the reduced gradient of this rank's shard equals the single-process gradient of the global batch (mean of per-rank means)
    hi = min(tr.hi, tr.numel)
    g_sh = tr.shard_grad[:hi - tr.lo] / world if world > 1 else tr.flat_grads[tr.lo:hi]
    g_ref = tr2.flat_grads[tr.lo:hi]
    dg = (g_sh - g_ref).abs().max().item() / (g_ref.abs().max().item() + 1e-12)
    # parameters: every rank holds the same values after the all-gather; the first Adam step moves each entry by at most lr
    dp = (tr.flat_params[:tr.numel] - tr2.flat_params[:tr2.numel]).abs().max().item()
    chk = tr.flat_params[:tr.numel].double().sum()
    chks = [torch.zeros_like(chk) for _ in range(world)]
    if world > 1:
        dist.all_gather(chks, chk)
    same = all(torch.equal(c, chks[0]) for c in chks) if world > 1 else True
    print(f"rank {rank}: reduced-gradient shard vs single-process gradient: max rel err {dg:.3e}; max |dparam| {dp:.3e} (lr 1e-4); "
          f"parameters identical on all ranks: {same}; mean loss over ranks {lg.item():.6f} vs global-batch loss {loss2.item():.6f}", flush=True)
    assert dg < 1e-4 and dp <= 2.1e-4 and same and abs(lg.item() - loss2.item()) < 1e-4 * abs(loss2.item())
elif a.ncu:
    one_step(tr, X, cond, eps)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    one_step(tr, X, cond, eps)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    for _ in range(2):
        loss = one_step(tr, X, cond, eps)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = []
    for _ in range(a.steps):
        losses.append(one_step(tr, X, cond, eps))
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    phases = None
    if rank == 0 and world == 1:
        ipk._lib.prof_enable(True)
        one_step(tr, X, cond, eps)
        phases = {k: round(t, 2) for k, (c, t) in ipk._lib.prof_report().items() if k.startswith(("train.", "enc."))}
        ipk._lib.prof_enable(False)
    if rank == 0:
        flops = (69.75 + 3 * 158.35) * 1e9 * B * world                        # encoder + flow forward + ~2x backward
        print(json.dumps({"workload": "h36m_128 second-stage training step (config 4)" + (" SMALL Hd=128" if a.small else ""), "n_gpus": world,
                          "per_gpu_batch": B, "ms_per_step": ms.item(), "samples_per_s": B * world / ms.item() * 1e3,
                          "algorithmic_tflops": flops / ms.item() / 1e9, "params": tr.numel, "losses": [l.item() for l in losses],
                          "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "phases_ms": phases, "precision": a.precision}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
