#!/usr/bin/env python
"""bench.py -- iPOKE stochastic-video sampling hot path: 16-frame 128x128 videos/s.

One "step" = one pass of the hot path over one batch of synthetic input per GPU:
    z ~ N(0,I) [B,32,8,8], cond [B,128,8,8], x0 [B,3,128,128]
      -> conditional MaCow flow inverse (1.05 B parameters)  -> latent ConvGRU (T steps) + SPADE decoder
      -> frames [B,T,3,128,128]   (+ ONE NCCL gather of the frames to rank 0 when N > 1)
Workload at N=1: BASELINE.json configs[1] "iper_128 sampling: 16-frame 128x128, batch 64, 1xB200 fp32".
Weak scaling: every rank runs the same per-GPU batch (samples are independent).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16|fp32_simt]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line on rank 0.  `--impl reference` times the reference algorithm's CPU path (the oracle port: the
reference is Python and /root/reference does not exist on the GPU box) on the host cores, bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "videos_per_sec_16f_128x128"
UNIT = "videos/s"
# algorithmic work: per 16-frame 128^2 video, C0 = 32
GFLOP_PER_VIDEO_MIN = 260.8
NICE_CONV2_FLOP_PER_PIXEL = 2.0 * 2048 * 2048          # one 1x1 2048->2048 conv output pixel (2 flop / MAC)
# dram__bytes_read.sum + dram__bytes_write.sum of one NICE conv2 launch at B=64, fp32 mode, from the ncu --set full capture
# summarised in profiles/r01_ncu_summary.md (50.5 MB read + 2.2..3.1 MB written; algorithmic reads are 50.4 MB)
NCU_CONV2_TRAFFIC_BYTES = {("fp32", 64): 53.1e6}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp32_simt"])
    ap.add_argument("--batch", type=int, default=64, help="videos per GPU per step")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--spatial", type=int, default=128)
    ap.add_argument("--cpu-sample", type=int, default=2, help="videos per CPU-baseline step")
    ap.add_argument("--chunk-videos", type=int, default=0, help="videos decoded per decoder pass (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-phases", action="store_true")
    ap.add_argument("--profile-mode", action="store_true", help="for ncu runs only: honour --warmup < 3 and skip e2e / phases / CPU baseline")
    a = ap.parse_args()
    if a.profile_mode:
        a.no_e2e = a.no_phases = a.no_cpu_baseline = True
    else:
        a.warmup = max(a.warmup, 3)
    return a


def workload_config(a, world):
    return {
        "workload": f"iper_128 sampling (BASELINE configs[1]): flow inverse C0=32 Hd=2048 15 levels (1.05 B params) -> ConvGRU x{a.frames} -> "
                    f"SPADE decoder {a.spatial}x{a.spatial}; {a.frames}-frame videos",
        "per_gpu_batch": a.batch, "global_batch": a.batch * world, "frames": a.frames, "spatial": a.spatial,
        "precision": {"fp32": "fp32 via bf16x3 error-compensated tcgen05 MMA, fp32 state/norms",
                      "bf16": "bf16 tcgen05 operands, fp32 accumulate/state/norms", "fp32_simt": "fp32 FFMA"}[a.precision],
        "sharding": f"dp{world}: batch-sharded, full weight replica per GPU, one NCCL gather of frames to rank 0" if world > 1 else "single GPU",
        "l2": "inputs larger than L2: every step streams 4.2 GB of packed flow weights + ~GBs of decoder activations (L2 = 126 MB); no explicit flush",
        "weights": "synthetic seeded (no checkpoints ship with the reference)",
    }


# ------------------------------------------------------------------------------------------------------------------
# CPU reference path (oracle port), used by cpu_baseline and by --impl reference
# ------------------------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, a, n_videos):
        import torch
        from oracle import ipoke_oracle as O
        self.torch, self.O = torch, O
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.fcfg = O.flow_config(flow_in_channels=32, flow_mid_channels=2048, h_channels=128)
        self.dcfg = O.first_stage_config(z_dim=32, spatial=a.spatial)
        self.fsd = O.synth_flow_state_dict(self.fcfg, seed=0)
        self.dsd = O.synth_first_stage_state_dict(self.dcfg, seed=1)
        self.n, self.T = n_videos, a.frames
        self.z, self.cond, self.x0 = O.synth_inputs(n_videos, 32, 128, a.spatial, seed=42)

    def step(self):
        with self.torch.no_grad():
            return self.O.sample_videos(self.fsd, self.fcfg, self.dsd, self.dcfg, self.z, self.cond, self.x0, self.T)

    def sample_desc(self):
        return (f"{self.n} videos x {self.T} frames per step at full size (same flow/decoder shapes as the GPU workload), "
                f"torch CPU fp32 oracle port of the reference modules, {self.cores} threads")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(a, a.cpu_sample)
    for _ in range(max(1, min(a.warmup, 1))):     # one CPU warm-up pass is ~8 s; more would not change the number
        ref.step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ref.step()
    dt = time.perf_counter() - t0
    v = a.cpu_sample * a.steps / dt
    world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(a, world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": ref.sample_desc()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            out["reasons"] = ["nvidia-smi unavailable"]
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm), reasons=sorted(reasons))
        return out


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import ipoke_b200 as ipk
    from ipoke_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()     # fail loudly here if the native library is missing

    B, T, S = a.batch, a.frames, a.spatial
    fcfg = dict(flow_in_channels=32, flow_mid_channels=2048, h_channels=128, num_steps=[10, 5, 5, 4, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1, 1],
                factor=16, transform="affine", prior_transform="affine", kernel_size=[2, 3], coupling_type="conv", activation="elu",
                condition_nice=False, attention=False, flow_attn_heads=4, cond_conv=False, cond_conv_hidden_channels=256, p_dropout=0.0,
                ipk_precision=a.precision, ipk_max_batch=B)
    dec = [256, 256, 256, 128, 64] if S == 128 else [256, 256, 128, 64]
    dcfg = dict(z_dim=32, norm="group", spectral_norm=True, n_gru_layers=4, dec_channels=dec, min_spatial_size=8, motion_bias=True,
                spatial=S, ipk_precision=a.precision, ipk_max_batch=B, ipk_max_frames=T, ipk_chunk_videos=a.chunk_videos)
    torch.manual_seed(1234)
    with torch.device(dev):
        flow = ipk.SupervisedMacowTransformer(fcfg)
        fs = ipk.SpadeCondMotionDecoder(dcfg)
    flow = synth.fill_flow_(flow.to(dev).eval(), seed=0)
    fs = synth.fill_first_stage_(fs.to(dev).eval(), seed=1)
    sampler = ipk.PokeMotionSampler(flow, fs)

    # synthetic inputs: global noise drawn once on the CPU generator and sliced per rank (second_stage_video.py:300)
    zg = ipk.global_noise(B * world, 32, seed=42)
    lo, hi = ipk.shard_bounds(B * world, world, rank)
    g = torch.Generator().manual_seed(100 + rank)
    z_h = zg[lo:hi].contiguous().pin_memory()
    cond_h = (torch.randn((B, 128, 8, 8), generator=g) * 0.5).pin_memory()
    x0_h = (torch.rand((B, 3, S, S), generator=g) * 2 - 1).pin_memory()
    z, cond, x0 = z_h.to(dev), cond_h.to(dev), x0_h.to(dev)
    gather_buf = [torch.empty((B, T, 3, S, S), device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    pending = []      # (NCCL work handle, frames tensor it reads) of the previous step's gather

    def drain():
        while pending:
            w, _keep = pending.pop(0)
            w.wait()

    def step():
        out = sampler.sample(z, cond, x0, T)
        if world > 1:
            # the single collective on the data path.  It is issued asynchronously: the gather of step i runs on NCCL's stream while
            # step i+1 computes (the frames of a step are a fresh tensor; the receive buffer is reused, so gathers stay in order)
            drain()
            pending.append((dist.gather(out, gather_buf, dst=0, async_op=True), out))
        return out

    def sync():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        out = step()
    sync()
    if not torch.isfinite(out).all():
        raise RuntimeError("bench: non-finite frames")

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # ---- timed region: device-resident inputs
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.launch_count_reset()
    sync()
    if a.profile_mode:
        torch.cuda.profiler.start()          # ncu --profile-from-start off: capture the timed steps only
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    sync()
    if a.profile_mode:
        torch.cuda.profiler.stop()
    launches = _lib.launch_count()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    # ---- end to end through the public host-buffer API: H2D of z/cond/x0 and D2H of the frames inside the timed region
    e2e = None
    if not a.no_e2e:
        for _ in range(2):
            sampler.sample_host(z_h, cond_h, x0_h, T, device=dev)
        sync()
        t_e0, t_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_e0.record()
        for _ in range(a.steps):
            frames_h = sampler.sample_host(z_h, cond_h, x0_h, T, device=dev)     # synchronous: returns with frames in pinned host memory
        t_e1.record()
        sync()
        ems = torch.tensor([t_e0.elapsed_time(t_e1)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        h2d = (z_h.numel() + cond_h.numel() + x0_h.numel()) * 4 * world
        d2h = frames_h.numel() * 4 * world
        e2e = {"value": B * world * a.steps / (float(ems.item()) / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "PokeMotionSampler.sample_host -> ipk_sample_host (pinned host buffers; frames delivered to host per rank)"}
        # same step with the post-processed sample (uint8 NTHWC, second_stage_video.py:673-675) leaving the device: 1/4 of the D2H bytes
        for _ in range(2):
            sampler.sample_host(z_h, cond_h, x0_h, T, device=dev, uint8=True)
        sync()
        u_e0, u_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u_e0.record()
        for _ in range(a.steps):
            frames_u8 = sampler.sample_host(z_h, cond_h, x0_h, T, device=dev, uint8=True)
        u_e1.record()
        sync()
        ums = torch.tensor([u_e0.elapsed_time(u_e1)], device=dev)
        if world > 1:
            dist.all_reduce(ums, op=dist.ReduceOp.MAX)
        e2e["uint8_frames"] = {"value": B * world * a.steps / (float(ums.item()) / 1e3), "unit": UNIT, "d2h_bytes_per_step": frames_u8.numel() * world,
                               "api": "PokeMotionSampler.sample_host(uint8=True) -> ipk_sample_host_u8"}
    clk = clocks.stop() if rank == 0 else None

    # ---- per-phase device times (separate untimed pass, CUDA events on the launch stream around every phase)
    phases, roofline = None, None
    if rank == 0 and not a.no_phases:
        _lib.prof_enable(True)
        nprof = 2
        for _ in range(nprof):
            sampler.sample(z, cond, x0, T)
        rep = _lib.prof_report()
        _lib.prof_enable(False)
        phases = {k: {"launch_groups": c // nprof, "ms_per_step": round(t / nprof, 4)} for k, (c, t) in rep.items()}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1400 TFLOP/s sustained"
        if "flow.nice.conv2" in rep:
            c, t = rep["flow.nice.conv2"]
            flop = NICE_CONV2_FLOP_PER_PIXEL * B * 64            # algorithmic: 2*M*N*K with M = B*64 pixels
            ach = flop / (t / c * 1e-3) / 1e12
            roofline = {"kernel": "conv_tc_kernel (NICE coupling conv2: 1x1 2048->2048 implicit GEMM, M=B*64)", "bound": "tensor",
                        "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": NCU_CONV2_TRAFFIC_BYTES.get((a.precision, B)),
                        "launches_per_step": c // nprof, "avg_launch_ms": t / c, "peak_source": peak_src,
                        "note": "algorithmic FLOPs (1x); fp32 mode issues 3 bf16 MMAs per product (bf16x3), so its ceiling is 1/3 of the bf16 peak"
                                if a.precision == "fp32" else "algorithmic FLOPs"}

    cpu_baseline = None
    if rank == 0 and not a.no_cpu_baseline:
        ref = CpuReference(a, a.cpu_sample)
        ref.step()
        t0 = time.perf_counter()
        ref.step()
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": a.cpu_sample / dt, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": ref.sample_desc()}

    if rank == 0:
        vps = B * world * a.steps / (ms_total / 1e3)
        line = {
            "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16": "bf16", "fp32_simt": "f32"}[a.precision], "data": "synthetic",
            "config": workload_config(a, world), "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "phases_ms": phases,
            "whole_step_tflops": GFLOP_PER_VIDEO_MIN * vps / 1e3,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
