"""Kernel micro-benchmarks (CUDA events, L2-flush between iterations). Not the driver's bench.py.

  python tools/microbench.py k0            # world=1 cast/scale kernel, HBM roofline
  python tools/microbench.py loopback      # W ranks on one GPU (protocol overhead only, no NVLink)
  torchrun --nproc-per-node N tools/microbench.py sweep   # real multi-GPU sweep vs NCCL
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_lightning_b200 import _b2d  # noqa: E402
from ray_lightning_b200.comm import LoopbackGroup  # noqa: E402


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return json.load(open(p))
    except Exception:
        return {"hbm_gbs": 6650.0, "fallback": True}


def time_ms(fn, iters, flush=None):
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def k0():
    hbm = peaks()["hbm_gbs"]
    ctx = _b2d.Context(0, 1, 0, 1 << 20)
    flush = torch.empty(256 << 20, dtype=torch.float32, device="cuda")  # 1 GiB > 126 MB L2
    st = torch.cuda.current_stream()
    for mib in (1, 8, 30, 98, 418, 1354):
        n = mib * (1 << 20) // 4
        buf = torch.randn(n, device="cuda")
        for wire in ("bf16", "fp32"):
            f = lambda: ctx.allreduce_bucket(0, buf.data_ptr(), n, _b2d.WIRE_NAMES[wire], 1.0, 0, st, st)
            for _ in range(3):
                f()
            med, best = time_ms(f, 10, flush)
            print(json.dumps({"bench": "k0", "wire": wire, "MiB": mib, "ms_med": round(med, 4), "ms_best": round(best, 4),
                              "GBps_med": round(8 * n / med / 1e6, 1), "frac_of_measured_hbm": round(8 * n / med / 1e6 / hbm, 3),
                              "grid": ctx.plan(n, 1)[1]}), flush=True)
        del buf
    ctx.destroy()


def loopback():
    for world in (2, 4, 8):
        g = LoopbackGroup(world, 0, arena_bytes=512 << 20, timeout_ms=20000)
        for mib in (0.0625, 1, 30):
            n = int(mib * (1 << 20)) // 4
            bufs = [torch.randn(n, device="cuda") for _ in range(world)]
            for algo in ("one_shot", "two_shot"):
                def f():
                    g.allreduce_(bufs, bucket_idx=int(mib * 100) + (algo == "one_shot"), wire="bf16", algo=algo)
                    g.join_current_stream()
                for _ in range(3):
                    f()
                med, best = time_ms(f, 10)
                print(json.dumps({"bench": "loopback", "world": world, "MiB": mib, "algo": algo,
                                  "ms_med": round(med, 4), "ms_best": round(best, 4)}), flush=True)
        g.close()


if __name__ == "__main__":
    {"k0": k0, "loopback": loopback}[sys.argv[1]]()
