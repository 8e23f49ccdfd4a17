"""Timeline probe of the tcgen05 conv engine inside one sampling step of the bench workload (B = 64, T = 16, 128x128, fp32 = bf16x3).

For every (N, Kpad) given on the command line the step is run once with ipk_tc_trace_enable(N, Kpad); the LAST conv_tc launch with that
weight shape leaves 32 SM-clock stamps per CTA, which are printed as medians over the CTAs (cycles since the CTA's entry).

usage (GPU box, repo root): python profiles/tc_trace_probe.py 2048:192 2048:2048 128:128 > gpurun_out/trace.txt
"""
import ctypes
import sys
import types

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

SLOTS = {0: "entry", 1: "prologue done", 2: "producer: pdl_wait passed (before first TMA)", 3: "producer: first stage issued", 4: "producer: tile 0 issued",
         5: "producer: done", 6: "mma: tile 0 first operands", 7: "mma: tile 0 committed", 8: "mma: tile 1 first operands",
         9: "mma: tile 1 committed", 10: "epi w2: tile 0 accumulator ready", 11: "epi w2: tile 0 chunk 0 done (fp32 path)",
         12: "epi w2: tile 0 done", 13: "epi w2: tile 1 accumulator ready", 14: "epi w2: tile 1 done", 15: "epi w9: tile 0 done",
         16: "epi w9: tile 1 done", 17: "all warps joined", 19: "producer: after first TMA", 26: "epi w2: last tile>1 accumulator ready", 28: "mma: last tile>1 first operands",
         29: "mma: last tile>1 committed"}


def main():
    a = types.SimpleNamespace(frames=16, spatial=128, chunk_videos=0)
    D = bench.Dist()
    run = bench.SamplingRun(a, D, "fp32", 64)
    for _ in range(2):
        run.step()
    torch.cuda.synchronize()
    L = run.ipk._lib
    lib = L.lib()
    for spec in sys.argv[1:]:
        n, k = (int(v) for v in spec.split(":"))
        L.check(lib.ipk_tc_trace_enable(n, k), "trace_enable")
        run.step()
        torch.cuda.synchronize()
        buf = np.zeros((148, 32), dtype=np.int64)
        L.check(lib.ipk_tc_trace_read(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 148), "trace_read")
        np.save(f"gpurun_out/trace_{n}_{k}.npy", buf)
        act = buf[buf[:, 0] != 0]
        print(f"== N={n} Kpad={k}: {len(act)} CTAs traced; globaltimer spread of CTA entries {int(act[:, 30].max() - act[:, 30].min())} ns")
        if len(act) == 0:
            continue
        for s, name in SLOTS.items():
            col = act[:, s]
            ok = col != 0
            if ok.sum() == 0:
                continue
            d = (col[ok] - act[ok, 0]).astype(np.float64)
            print(f"  slot {s:2d} {name:45s} n={int(ok.sum()):3d} median {np.median(d):9.0f} cyc ({np.median(d) / 1.92e3:6.2f} us)  min {d.min():9.0f}  max {d.max():9.0f}")
    lib.ipk_tc_trace_enable(-1, -1)


if __name__ == "__main__":
    main()
