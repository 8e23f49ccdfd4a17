#!/usr/bin/env python
"""bench.py -- iPOKE stochastic-video sampling hot path: 16-frame 128x128 videos/s.

One "step" = one pass of the hot path over one batch of synthetic input per GPU:
    z ~ N(0,I) [B,32,8,8], cond [B,128,8,8], x0 [B,3,128,128]
      -> conditional MaCow flow inverse (1.05 B parameters)  -> latent ConvGRU (T steps) + SPADE decoder
      -> frames [B,T,3,128,128]   (+ ONE NCCL gather of the frames to rank 0 when N > 1)
Workload at N=1: BASELINE.json configs[1] "iper_128 sampling: 16-frame 128x128, batch 64, 1xB200 fp32".
Weak scaling: every rank runs the same per-GPU batch (samples are independent).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload sample|train]
                    [--precision fp32|bf16|fp32_simt] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line on rank 0.  Besides the primary measurement the default run adds, under "secondary", the two other
GPU configurations BASELINE.json names: configs[2] (taichi_128, bf16 operands, 32 videos per GPU, gather included) and
configs[3] (h36m_128 second-stage training step, 32 samples per GPU, gradient reduce-scatter included); `--workload train`
makes the training step the primary line instead.

Ordering rule (round-1 bug): no rank ever waits inside an NCCL collective while another rank does CPU work.  Everything
that needs the process group (timed region, e2e, secondary workloads) runs first on all ranks; then the group is
destroyed, ranks != 0 exit, and only then rank 0 runs the CPU legs (parity oracle, cpu_baseline) and prints the line.

`--impl reference` times the reference algorithm's CPU path (the oracle port: the reference is Python and
/root/reference does not exist on the GPU box) on the host cores, a bounded sample per step.
"""
import argparse
import gc
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
TRAIN_METRIC = "train_samples_per_sec_h36m_128"
# algorithmic work: per 16-frame 128^2 video, C0 = 32
GFLOP_PER_VIDEO_MIN = 260.8
NICE_CONV2_FLOP_PER_PIXEL = 2.0 * 2048 * 2048          # one 1x1 2048->2048 conv output pixel (2 flop / MAC)
TRAIN_GFLOP_PER_SAMPLE = 69.75 + 3 * 158.35            # config 4: encoder + flow forward + ~2x backward
NUM_STEPS = [10, 5, 5, 4, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1, 1]
DTYPE = {"fp32": "bf16x3", "bf16": "bf16", "fp32_simt": "f32"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sample", choices=["sample", "train"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp32_simt"])
    ap.add_argument("--batch", type=int, default=0, help="videos (samples) per GPU per step; default 64 (sample) / 32 (train)")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--spatial", type=int, default=128)
    ap.add_argument("--cpu-sample", type=int, default=2, help="videos per cpu_baseline step inside the GPU arm's run")
    ap.add_argument("--ref-batch", type=int, default=8, help="videos per step of --impl reference (BASELINE.md section 3: B=8 is the CPU-feasible stand-in)")
    ap.add_argument("--chunk-videos", type=int, default=0, help="videos decoded per decoder pass (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-phases", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[2] / configs[3] secondary measurements")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--profile-mode", action="store_true", help="for ncu runs only: honour --warmup < 3 and skip every optional leg")
    a = ap.parse_args()
    if a.batch <= 0:
        a.batch = 64 if a.workload == "sample" else 32
    if a.profile_mode:
        a.no_e2e = a.no_phases = a.no_cpu_baseline = a.no_secondary = a.no_parity = True
    else:
        a.warmup = max(a.warmup, 3)
    if a.precision != "fp32" or a.workload != "sample":
        a.no_secondary = True          # secondaries accompany the headline configuration only
    return a


def host_threads():
    """Threads the CPU legs may use: the process's CPU affinity, capped by the cgroup quota and by the physical core count
    (hyper-threads slow the dispatch-bound oracle down: 32 threads ran slower than 16 on the round-1 box)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    try:
        import psutil
        phys = psutil.cpu_count(logical=False)
        if phys:
            n = min(n, phys)
    except Exception:
        pass
    return max(1, n)


def flow_cfg(c0, precision, max_batch):
    return dict(flow_in_channels=c0, flow_mid_channels=2048, h_channels=128, num_steps=list(NUM_STEPS), factor=16, transform="affine",
                prior_transform="affine", kernel_size=[2, 3], coupling_type="conv", activation="elu", condition_nice=False, attention=False,
                flow_attn_heads=4, cond_conv=False, cond_conv_hidden_channels=256, p_dropout=0.0, ipk_precision=precision, ipk_max_batch=max_batch)


def workload_config(a, world, batch=None, precision=None):
    batch = batch or a.batch
    precision = precision or a.precision
    return {
        "workload": f"iper_128 sampling (BASELINE configs[1]): flow inverse C0=32 Hd=2048 15 levels (1.05 B params) -> ConvGRU x{a.frames} -> "
                    f"SPADE decoder {a.spatial}x{a.spatial}; {a.frames}-frame videos",
        "per_gpu_batch": batch, "global_batch": batch * world, "frames": a.frames, "spatial": a.spatial,
        "precision": {"fp32": "fp32 via bf16x3 error-compensated tcgen05 MMA, fp32 state/norms",
                      "bf16": "bf16 tcgen05 operands, fp32 accumulate/state/norms", "fp32_simt": "fp32 FFMA"}[precision],
        "sharding": f"dp{world}: batch-sharded, full weight replica per GPU, one NCCL gather of frames to rank 0" if world > 1 else "single GPU",
        "l2": "inputs larger than L2: every step streams 4.2 GB of packed flow weights + ~GBs of decoder activations (L2 = 126 MB); no explicit flush",
        "weights": "synthetic seeded (no checkpoints ship with the reference)",
    }


def train_config(batch, world):
    return {
        "workload": "h36m_128 second-stage training step (BASELINE configs[3]): clip [B,11,3,128,128] -> 3-D conv encoder (no grad) -> flow forward + "
                    "log-det (C0=64, Hd=2048, 1.24 B params) -> FlowLoss -> backward -> reduce-scatter of the 4.95 GB flat gradient -> "
                    "Adam(amsgrad, betas .9/.999, wd 1e-5) on each rank's shard -> all-gather of the parameters",
        "per_gpu_batch": batch, "global_batch": batch * world,
        "precision": "fp32 via bf16x3 error-compensated tcgen05 MMA, fp32 master parameters / state / optimizer",
        "sharding": f"dp{world}: batch-sharded, NCCL reduce-scatter + all-gather" if world > 1 else "single GPU",
        "l2": "inputs larger than L2: every step streams 4.95 GB of parameters, gradients and ~10 GB of kept activations; no explicit flush",
        "weights": "synthetic seeded",
    }


# ------------------------------------------------------------------------------------------------------------------
# CPU reference path (oracle port), used by parity, cpu_baseline and by --impl reference
# ------------------------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, a, n_videos, flow_sd=None, fs_sd=None, inputs=None, device="cpu"):
        import torch
        from oracle import ipoke_oracle as O
        self.torch, self.O = torch, O
        self.cores = host_threads()
        torch.set_num_threads(self.cores)
        self.device = device
        self.fcfg = O.flow_config(flow_in_channels=32, flow_mid_channels=2048, h_channels=128)
        self.dcfg = O.first_stage_config(z_dim=32, spatial=a.spatial)
        self.fsd = flow_sd if flow_sd is not None else O.synth_flow_state_dict(self.fcfg, seed=0)
        self.dsd = fs_sd if fs_sd is not None else O.synth_first_stage_state_dict(self.dcfg, seed=1)
        self.n, self.T = n_videos, a.frames
        if inputs is None:
            inputs = O.synth_inputs(n_videos, 32, 128, a.spatial, seed=42)
        self.z, self.cond, self.x0 = inputs
        if device != "cpu":
            self.fsd = {k: v.to(device) for k, v in self.fsd.items()}
            self.dsd = {k: v.to(device) for k, v in self.dsd.items()}
            self.z, self.cond, self.x0 = (t.to(device) for t in (self.z, self.cond, self.x0))

    def step(self):
        with self.torch.no_grad():
            return self.O.sample_videos(self.fsd, self.fcfg, self.dsd, self.dcfg, self.z, self.cond, self.x0, self.T)

    def sample_desc(self):
        return (f"{self.n} videos x {self.T} frames per step at full size (same flow/decoder shapes as the GPU workload), "
                f"torch fp32 oracle port of the reference modules, " + (f"{self.cores} host threads" if self.device == "cpu" else f"eager on {self.device}"))


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nb = a.ref_batch
    ref = CpuReference(a, nb)
    ref.step()                                     # one CPU warm-up pass (~10 s); more would not change the number
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ref.step()
    dt = time.perf_counter() - t0
    v = nb * a.steps / dt
    cfg = workload_config(a, 1, batch=nb)
    cfg["sharding"] = "host CPU, rank 0 only"
    cfg["precision"] = "fp32 (torch CPU)"
    cfg["note"] = (f"the CPU arm runs {nb} videos per step (BASELINE.md section 3: B=8 is the CPU-feasible stand-in for the GPU arm's per-GPU batch); "
                   "videos/s is the comparable quantity")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": 1,
        "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def ncu_traffic(kernel_key, precision, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture summarised in profiles/roofline_traffic.json (written by profiles/ncu_traffic.py from the raw ncu page)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        e = t[kernel_key]
        if e.get("precision") == precision and int(e.get("batch", -1)) == int(batch):
            return float(e["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------------------------
# distributed helpers
# ------------------------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py (impl ours) needs a CUDA device: the product has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def sync(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, value):
        t = self.torch.tensor([float(value)], device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok):
        """True only if every rank reports ok (so that ranks take the same branch around collectives)."""
        t = self.torch.tensor([1.0 if ok else 0.0], device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def shutdown(self):
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()
            self.dist.destroy_process_group()
            self.dist = None


# ------------------------------------------------------------------------------------------------------------------
# sampling workload
# ------------------------------------------------------------------------------------------------------------------
class SamplingRun:
    """Builds flow + first stage for one precision / per-GPU batch and times steps (device-resident and end to end)."""

    def __init__(self, a, D, precision, batch):
        import torch
        import ipoke_b200 as ipk
        from ipoke_b200 import synth
        self.a, self.D, self.ipk, self.torch = a, D, ipk, torch
        self.B, self.T, self.S = batch, a.frames, a.spatial
        dev = D.dev
        dec = [256, 256, 256, 128, 64] if self.S == 128 else [256, 256, 128, 64]
        dcfg = dict(z_dim=32, norm="group", spectral_norm=True, n_gru_layers=4, dec_channels=dec, min_spatial_size=8, motion_bias=True,
                    spatial=self.S, ipk_precision=precision, ipk_max_batch=batch, ipk_max_frames=self.T, ipk_chunk_videos=a.chunk_videos)
        torch.manual_seed(1234)
        with torch.device(dev):
            flow = ipk.SupervisedMacowTransformer(flow_cfg(32, precision, batch))
            fs = ipk.SpadeCondMotionDecoder(dcfg)
        self.flow = synth.fill_flow_(flow.to(dev).eval(), seed=0)
        self.fs = synth.fill_first_stage_(fs.to(dev).eval(), seed=1)
        self.sampler = ipk.PokeMotionSampler(self.flow, self.fs)
        # synthetic inputs: global noise drawn once on the CPU generator and sliced per rank (second_stage_video.py:300)
        zg = ipk.global_noise(batch * D.world, 32, seed=42)
        lo, hi = ipk.shard_bounds(batch * D.world, D.world, D.rank)
        g = torch.Generator().manual_seed(100 + D.rank)
        self.z_h = zg[lo:hi].contiguous().pin_memory()
        self.cond_h = (torch.randn((batch, 128, 8, 8), generator=g) * 0.5).pin_memory()
        self.x0_h = (torch.rand((batch, 3, self.S, self.S), generator=g) * 2 - 1).pin_memory()
        self.z, self.cond, self.x0 = self.z_h.to(dev), self.cond_h.to(dev), self.x0_h.to(dev)
        self.gather_buf = ([torch.empty((batch, self.T, 3, self.S, self.S), device=dev) for _ in range(D.world)]
                           if (D.world > 1 and D.rank == 0) else None)
        self.pending = []      # (NCCL work handle, frames tensor it reads) of the previous step's gather
        self.last = None

    def drain(self):
        while self.pending:
            w, _keep = self.pending.pop(0)
            w.wait()

    def step(self):
        out = self.sampler.sample(self.z, self.cond, self.x0, self.T)
        if self.D.dist is not None:
            # the single collective on the data path.  It is issued asynchronously: the gather of step i runs on NCCL's stream while
            # step i+1 computes (the frames of a step are a fresh tensor; the receive buffer is reused, so gathers stay in order)
            self.drain()
            self.pending.append((self.D.dist.gather(out, self.gather_buf, dst=0, async_op=True), out))
        self.last = out
        return out

    def sync(self):
        self.drain()
        self.D.sync()

    def timed(self, steps, warmup, profile=False):
        torch, _lib = self.torch, self.ipk._lib
        for _ in range(warmup):
            out = self.step()
        self.sync()
        if not torch.isfinite(out).all():
            raise RuntimeError("bench: non-finite frames")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.launch_count_reset()
        self.sync()
        if profile:
            torch.cuda.profiler.start()          # ncu --profile-from-start off: capture the timed steps only
        e0.record()
        for _ in range(steps):
            self.step()
        e1.record()
        self.sync()
        if profile:
            torch.cuda.profiler.stop()
        launches = _lib.launch_count()
        return self.D.max_over_ranks(e0.elapsed_time(e1)), int(launches)

    def e2e(self, steps):
        """End to end through the public host-buffer API: H2D of z/cond/x0 and D2H of the frames inside the timed region."""
        torch, D = self.torch, self.D
        out = {}
        for key, kw in (("fp32", {}), ("uint8_frames", {"uint8": True})):
            for _ in range(2):
                self.sampler.sample_host(self.z_h, self.cond_h, self.x0_h, self.T, device=D.dev, reuse=True, **kw)
            self.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                frames_h = self.sampler.sample_host(self.z_h, self.cond_h, self.x0_h, self.T, device=D.dev, reuse=True, **kw)    # synchronous
            e1.record()
            self.sync()
            ms = D.max_over_ranks(e0.elapsed_time(e1))
            out[key] = (self.B * D.world * steps / (ms / 1e3), frames_h.numel() * frames_h.element_size() * D.world)
        h2d = (self.z_h.numel() + self.cond_h.numel() + self.x0_h.numel()) * 4 * D.world
        e = {"value": out["fp32"][0], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out["fp32"][1],
             "api": "PokeMotionSampler.sample_host -> ipk_sample_host (pinned host buffers; frames delivered to host per rank)",
             # same step with the post-processed sample (uint8 NTHWC, second_stage_video.py:673-675) leaving the device: 1/4 of the D2H bytes
             "uint8_frames": {"value": out["uint8_frames"][0], "unit": UNIT, "d2h_bytes_per_step": out["uint8_frames"][1],
                              "api": "PokeMotionSampler.sample_host(uint8=True) -> ipk_sample_host_u8"}}
        return e

    def phases(self, nprof=2):
        """Per-phase device times (separate untimed pass, CUDA events on the launch stream around every phase); rank-local."""
        _lib = self.ipk._lib
        _lib.prof_enable(True)
        for _ in range(nprof):
            self.sampler.sample(self.z, self.cond, self.x0, self.T)
        rep = _lib.prof_report()
        _lib.prof_enable(False)
        return rep, nprof

    def latency_b1(self, reps=5):
        """GUI call shape (testing/gui.py:139-148): one video, flow inverse + decode, device-resident; ms per sample."""
        torch = self.torch
        z1, c1, x1 = self.z[:1].contiguous(), self.cond[:1].contiguous(), self.x0[:1].contiguous()
        for _ in range(2):
            self.sampler.sample(z1, c1, x1, self.T)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            self.sampler.sample(z1, c1, x1, self.T)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def free(self):
        self.drain()
        self.sampler = self.flow = self.fs = self.gather_buf = self.last = None
        self.z = self.cond = self.x0 = self.z_h = self.cond_h = self.x0_h = None
        gc.collect()
        self.torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------------------------
# training workload (config 4)
# ------------------------------------------------------------------------------------------------------------------
class TrainRun:
    def __init__(self, a, D, precision, batch):
        import torch
        import ipoke_b200 as ipk
        from ipoke_b200 import synth
        self.a, self.D, self.ipk, self.torch, self.B = a, D, ipk, torch, batch
        dev, C0 = D.dev, 64
        with torch.device(dev):
            flow = ipk.SupervisedMacowTransformer(flow_cfg(C0, precision, batch))
        self.flow = synth.fill_flow_(flow.to(dev).eval(), seed=0)
        ecfg = dict(z_dim=C0, img_size=128, max_frames=10, full_seq=True, ENC_M_channels=[64, 128, 256, 256, 256], min_spatial_size=8,
                    ipk_max_batch=batch, ipk_precision=precision)
        enc = ipk.ResNetMotionEncoder(ecfg)
        enc = synth.fill_encoder_(enc.to(dev).eval(), seed=1)
        self.enc = enc
        g = torch.Generator().manual_seed(7 + D.rank)
        self.X_h = (torch.rand((batch, 11, 3, 128, 128), generator=g) * 2 - 1).pin_memory()
        self.cond_h = (torch.randn((batch, 128, 8, 8), generator=g) * 0.5).pin_memory()
        self.eps_h = torch.randn((batch, C0, 8, 8), generator=g).pin_memory()
        self.X, self.cond, self.eps = self.X_h.to(dev), self.cond_h.to(dev), self.eps_h.to(dev)
        self.tr = ipk.FlowTrainer(self.flow, max_batch=batch, precision=precision)
        torch.cuda.synchronize()

    def step(self, X=None, cond=None, eps=None):
        X = self.X if X is None else X
        cond = self.cond if cond is None else cond
        eps = self.eps if eps is None else eps
        z_in, _ = self.ipk.encode_first_stage(self.enc, X, eps=eps)
        loss = self.tr.step(z_in, cond)
        # configure_optimizers (second_stage_video.py:647-648): Adam(betas=(0.9, 0.999), weight_decay=cfg, amsgrad=True)
        self.tr.optimizer_step(lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-5, amsgrad=True)
        return loss

    def timed(self, steps, warmup, profile=False):
        torch, D, _lib = self.torch, self.D, self.ipk._lib
        for _ in range(max(warmup, 3)):            # the step is captured as a CUDA graph on its 2nd call and replayed from the 3rd
            loss = self.step()
        D.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.launch_count_reset()
        if profile:
            torch.cuda.profiler.start()
        e0.record()
        losses = []
        for _ in range(steps):
            losses.append(self.step())
        e1.record()
        D.sync()
        if profile:
            torch.cuda.profiler.stop()
        launches = _lib.launch_count()
        lv = [float(l.item()) for l in losses]
        if not all(v == v and abs(v) < 1e30 for v in lv):
            raise RuntimeError(f"bench: non-finite training loss {lv}")
        return D.max_over_ranks(e0.elapsed_time(e1)), int(launches), lv

    def e2e(self, steps):
        torch, D = self.torch, self.D
        dev = D.dev
        def one():
            X, c, e = self.X_h.to(dev, non_blocking=True), self.cond_h.to(dev, non_blocking=True), self.eps_h.to(dev, non_blocking=True)
            return float(self.step(X, c, e).item())           # D2H read of the loss: synchronous
        one()
        D.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        D.sync()
        ms = D.max_over_ranks(e0.elapsed_time(e1))
        h2d = (self.X_h.numel() + self.cond_h.numel() + self.eps_h.numel()) * 4 * D.world
        return {"value": self.B * D.world * steps / (ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * D.world,
                "api": "encode_first_stage -> FlowTrainer.step -> FlowTrainer.optimizer_step (pinned host clip / cond / eps in, loss out)"}

    def phases(self):
        _lib = self.ipk._lib
        _lib.prof_enable(True)
        self.step()
        rep = {k: {"n": c, "ms": round(t, 3)} for k, (c, t) in _lib.prof_report().items() if k.startswith(("train.", "enc."))}
        _lib.prof_enable(False)
        return rep

    def free(self):
        self.tr = self.flow = self.enc = None
        self.X = self.cond = self.eps = self.X_h = self.cond_h = self.eps_h = None
        gc.collect()
        self.torch.cuda.empty_cache()


def train_summary(a, D, run, steps, warmup, with_e2e, with_phases):
    ms, launches, losses = run.timed(steps, warmup, profile=a.profile_mode)
    sps = run.B * D.world * steps / (ms / 1e3)
    peaks = load_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    ach = TRAIN_GFLOP_PER_SAMPLE * sps / 1e3
    out = {"metric": TRAIN_METRIC, "value": sps, "unit": "samples/s", "n_gpus": D.world, "steps": steps, "warmup": max(warmup, 3),
           "ms_per_step": ms / steps, "config": train_config(run.B, D.world), "gpu_launches": launches, "losses": losses,
           "algorithmic_tflops": ach,
           "roofline": {"kernel": "whole training step (545 GFLOP algorithmic per sample: encoder + flow forward + 2x backward)", "bound": "tensor",
                        "achieved": ach / D.world, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / D.world / peak_tf, "traffic": None,
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400 TFLOP/s sustained",
                        "note": "per-GPU algorithmic FLOPs (1x); fp32 mode issues 3 bf16 MMAs per product"}}
    if with_e2e:
        out["e2e"] = run.e2e(steps)
    if with_phases and D.rank == 0:
        out["phases_ms"] = run.phases()
    return out


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    from ipoke_b200 import _lib
    D = Dist()
    _lib.lib()     # fail loudly here if the native library is missing
    if a.workload == "train":
        return run_ours_train(a, D)

    B, T, S = a.batch, a.frames, a.spatial
    run = SamplingRun(a, D, a.precision, B)
    clocks = ClockSampler(D.local)
    # warm-up happens inside timed(); the clock sampler covers warm-up + timed region + e2e
    if D.rank == 0:
        clocks.start()
    ms_total, launches = run.timed(a.steps, a.warmup, profile=a.profile_mode)
    parity_gpu = None
    if D.rank == 0 and not a.no_parity:
        parity_gpu = run.last[:2].detach().cpu()          # first two videos of the timed batch (samples are independent)
    e2e = None if a.no_e2e else run.e2e(a.steps)
    clk = clocks.stop() if D.rank == 0 else None

    phases, roofline, latency = None, None, None
    if D.rank == 0 and not a.no_phases:
        # rank-local GPU work (2 untimed steps, ~0.2 s): the other ranks go on to the secondary workloads' set-up meanwhile
        rep, nprof = run.phases()
        phases = {k: {"launch_groups": c // nprof, "ms_per_step": round(t / nprof, 4)} for k, (c, t) in rep.items()}
        peaks = load_peaks()
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1400 TFLOP/s sustained"
        if "flow.nice.conv2" in rep:
            c, t = rep["flow.nice.conv2"]
            flop = NICE_CONV2_FLOP_PER_PIXEL * B * 64            # algorithmic: 2*M*N*K with M = B*64 pixels
            ach = flop / (t / c * 1e-3) / 1e12
            roofline = {"kernel": "conv_tc_kernel (NICE coupling conv2: 1x1 2048->2048 implicit GEMM, M=B*64)", "bound": "tensor",
                        "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": ncu_traffic("nice_conv2", a.precision, B),
                        "launches_per_step": c // nprof, "avg_launch_ms": t / c, "peak_source": peak_src,
                        "note": "algorithmic FLOPs (1x); fp32 mode issues 3 bf16 MMAs per product (bf16x3), so its ceiling is 1/3 of the bf16 peak; 1.73 waves of the persistent grid (tensor pipe 76 % busy on the two-unit SMs), see DESIGN.md 7b"
                                if a.precision == "fp32" else "algorithmic FLOPs"}
        latency = {"ms_per_sample_b1": round(run.latency_b1(), 3), "note": "B=1, T=%d, device-resident (GUI call shape, testing/gui.py:139-148)" % T}

    # state needed by the CPU legs, taken before the GPU objects are released
    cpu_state = None
    if D.rank == 0 and not (a.no_parity and (a.no_cpu_baseline or D.world > 1)):
        cpu_state = ({k: v.detach().cpu() for k, v in run.flow.state_dict().items()}, {k: v.detach().cpu() for k, v in run.fs.state_dict().items()},
                     (run.z_h[:2].clone(), run.cond_h[:2].clone(), run.x0_h[:2].clone()))

    # ---- secondary workloads (all ranks; need the process group): configs[2] bf16 sampling, configs[3] training step
    secondary = None
    if not a.no_secondary:
        run.free()
        secondary = {}
        ksteps = max(3, min(a.steps, 10))
        for name, fn in (("config3_taichi_128_bf16", lambda: secondary_bf16(a, D, ksteps)), ("config4_h36m_128_train", lambda: secondary_train(a, D, min(ksteps, 5)))):
            err, res = None, None
            try:
                res = fn()
            except Exception as e:       # noqa: BLE001 -- a failing secondary must not cost the primary line
                err = f"{type(e).__name__}: {e}"[:300]
            if not D.all_ok(err is None):
                res = {"error": err or "failed on another rank"}
            secondary[name] = res
            gc.collect()
            torch.cuda.empty_cache()
    else:
        run.free()

    D.shutdown()
    if D.rank != 0:
        return

    # ---- CPU legs, rank 0 only, no process group alive
    parity, cpu_baseline, gpu_eager = None, None, None
    if cpu_state is not None and parity_gpu is not None:
        fsd, dsd, inp = cpu_state
        ref = CpuReference(a, 2, flow_sd=fsd, fs_sd=dsd, inputs=inp)
        t0 = time.perf_counter()
        want = ref.step()
        dt = time.perf_counter() - t0
        err = (parity_gpu.double() - want.double()).abs()
        tol = {"fp32": 1e-3, "fp32_simt": 1e-3, "bf16": 1.5e-1}[a.precision]
        parity = {"max_abs": float(err.max()), "max_abs_per_frame_mean": float(err.flatten(2).max(dim=2).values.mean()), "tol": tol,
                  "ok": bool(err.max() <= tol), "videos": 2,
                  "what": f"first 2 videos of the timed B={B} T={T} {S}x{S} batch vs the CPU fp32 oracle on the same weights / z / cond / x0"}
        if D.world == 1 and not a.no_cpu_baseline:
            # the same pass doubles as the cpu_baseline warm-up; time a second one
            t0 = time.perf_counter()
            ref.step()
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": 2 / dt, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": ref.sample_desc()}
            try:
                gref = CpuReference(a, 8, flow_sd=fsd, fs_sd=dsd, device=f"cuda:{D.local}")
                gref.step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                gref.step()
                torch.cuda.synchronize()
                gpu_eager = {"value": 8 / (time.perf_counter() - t0), "unit": UNIT, "sample": gref.sample_desc(),
                             "note": "the oracle's torch ops executed eagerly on the same B200 (BASELINE.md section 3 'GPU reference point'); a baseline, never the product path"}
                del gref
            except Exception as e:       # noqa: BLE001
                gpu_eager = {"error": f"{type(e).__name__}: {e}"[:200]}
        if not parity["ok"]:
            print(f"bench: PARITY FAILED: max-abs {parity['max_abs']:.3e} > {tol}", file=sys.stderr, flush=True)
    elif D.world == 1 and not a.no_cpu_baseline:
        ref = CpuReference(a, a.cpu_sample)
        ref.step()
        t0 = time.perf_counter()
        ref.step()
        cpu_baseline = {"value": a.cpu_sample / (time.perf_counter() - t0), "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": ref.sample_desc()}

    vps = B * D.world * a.steps / (ms_total / 1e3)
    line = {
        "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": D.world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE[a.precision], "data": "synthetic",
        "config": workload_config(a, D.world), "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity, "gpu_eager_baseline": gpu_eager, "latency": latency,
        "secondary": secondary, "phases_ms": phases, "whole_step_tflops": GFLOP_PER_VIDEO_MIN * vps / 1e3,
    }
    print(json.dumps(line), flush=True)
    if parity is not None and not parity["ok"]:
        sys.exit(3)


def secondary_bf16(a, D, steps):
    """BASELINE configs[2]: taichi_128 shapes (= iper_128 shapes, bn32 first stage), bf16 operands, 32 videos per GPU, gather included."""
    run = SamplingRun(a, D, "bf16", 32)
    try:
        ms, launches = run.timed(steps, 3)
        vps = 32 * D.world * steps / (ms / 1e3)
        return {"metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": D.world, "steps": steps, "warmup": 3, "ms_per_step": ms / steps, "dtype": "bf16",
                "config": workload_config(a, D.world, batch=32, precision="bf16"), "gpu_launches": launches,
                "tolerance": "bf16 operands: stated 1e-1 latent / 1.5e-1 frames max-abs vs the fp32 reference (tests/test_gpu_flow.py, test_gpu_first_stage.py)"}
    finally:
        run.free()


def secondary_train(a, D, steps):
    run = TrainRun(a, D, "fp32", 32)
    try:
        return train_summary(a, D, run, steps, 3, with_e2e=False, with_phases=False)
    finally:
        run.free()


def run_ours_train(a, D):
    run = TrainRun(a, D, a.precision, a.batch)
    clocks = ClockSampler(D.local)
    if D.rank == 0:
        clocks.start()
    s = train_summary(a, D, run, a.steps, a.warmup, with_e2e=not a.no_e2e, with_phases=not a.no_phases and D.world == 1)
    clk = clocks.stop() if D.rank == 0 else None
    run.free()
    D.shutdown()
    if D.rank != 0:
        return
    line = {"metric": s["metric"], "value": s["value"], "unit": s["unit"], "n_gpus": D.world, "steps": a.steps, "warmup": s["warmup"],
            "ms_per_step": s["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE[a.precision],
            "data": "synthetic", "config": s["config"], "e2e": s.get("e2e"), "gpu_launches": s["gpu_launches"], "clocks": clk,
            "roofline": s["roofline"], "cpu_baseline": None, "losses": s["losses"], "phases_ms": s.get("phases_ms"),
            "algorithmic_tflops": s["algorithmic_tflops"]}
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
