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


def sweep():
    """Real multi-GPU sweep (one process per GPU, launched by torchrun): libb2d vs NCCL."""
    import torch.distributed as dist
    from ray_lightning_b200.comm import Communicator
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mem = os.environ.get("B2D_MEM", "vmm")
    max_ctas = int(os.environ.get("B2D_MAX_CTAS", "128"))
    comm = Communicator(rank, world, local, 3 << 30, mem=mem, max_ctas=max_ctas, timeout_ms=20000, nvls="auto")
    if os.environ.get("B2D_TMA_CTAS"):
        comm.ctx.set_tma_ctas(int(os.environ["B2D_TMA_CTAS"]))
    if rank == 0:
        print(json.dumps({"bench": "sweep_setup", "world": world, "mem": mem, "nvls": comm.nvls, "max_ctas": max_ctas}), flush=True)
    side = torch.cuda.Stream()
    sizes = [64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20]  # bytes of WIRE payload (bf16)
    iters = int(os.environ.get("B2D_ITERS", "30"))

    def timed(fn, n_it):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_it):
            fn()
        b.record()
        b.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n_it], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for wire_bytes in sizes:
        n = wire_bytes // 2
        buf = torch.randn(n, device="cuda") * 0.01
        ref = buf.clone()
        bus = 2 * (world - 1) / world * wire_bytes
        rows = {}
        algos = ["one_shot", "two_shot", "two_shot_tma"] + (["nvls"] if comm.nvls and os.environ.get("B2D_SKIP_NVLS") != "1" else [])
        for algo in algos:
            if algo == "one_shot" and wire_bytes > (16 << 20):
                continue
            key = sizes.index(wire_bytes) * 10 + algos.index(algo)
            f = lambda: comm.allreduce_(buf, bucket_idx=key, wire="bf16", algo=algo)
            rows[algo] = timed(f, iters)
            if os.environ.get("B2D_TRACE") == "1":
                comm.ctx.trace(True)
                torch.cuda.synchronize(); dist.barrier()
                f()
                ph = comm.ctx.trace(True, read=True)
                if rank == 0:
                    print(json.dumps({"bench": "trace", "world": world, "wire_bytes": wire_bytes, "algo": algo,
                                      "grid": comm.ctx.stats()["last_grid"],
                                      "phase_us(stage,barrierA,reduce,barrierB,gather,...,span)": [round(x, 2) for x in ph]}), flush=True)
                comm.ctx.trace(False)
        # the reference GPU path with the bf16 hook: cast+div, ncclAllReduce(bf16), copy back
        def nccl_bf16():
            c = buf.to(torch.bfloat16).div_(world)
            dist.all_reduce(c)
            buf.copy_(c)
        if os.environ.get("B2D_SKIP_NCCL") != "1":
            rows["nccl_bf16_hook_seq"] = timed(nccl_bf16, iters)
            cb = buf.to(torch.bfloat16)
            rows["nccl_bf16_allreduce_only"] = timed(lambda: dist.all_reduce(cb), iters)
        if os.environ.get("B2D_SKIP_FP32") != "1":
            rows["nccl_fp32_allreduce"] = timed(lambda: dist.all_reduce(buf), iters)
        buf.copy_(ref)
        if rank == 0:
            out = {"bench": "sweep", "world": world, "wire_bytes": wire_bytes, "n": n}
            for k, ms in rows.items():
                out[k + "_ms"] = round(ms, 4)
                w = 4 * n if k == "nccl_fp32_allreduce" else wire_bytes
                out[k + "_busGBps"] = round(2 * (world - 1) / world * w / ms / 1e6, 1)
            print(json.dumps(out), flush=True)
    # correctness against NCCL on the same inputs
    g = torch.randn(1 << 20, device="cuda", generator=torch.Generator("cuda").manual_seed(rank))
    mine = g.clone()
    comm.allreduce_(mine, bucket_idx=999, wire="fp32", algo="two_shot")
    theirs = g / world
    dist.all_reduce(theirs)
    torch.cuda.synchronize()
    ok = torch.allclose(mine, theirs, rtol=1e-3, atol=1e-5)   # north-star tolerance vs the reference's fp32 path
    # bf16 wire: NCCL rounds every partial sum, libb2d once — compare both with the exact fp64 sum of the wire values
    mine2 = g.clone()
    comm.allreduce_(mine2, bucket_idx=998, wire="bf16", algo="two_shot")
    c = g.to(torch.bfloat16).div_(world)
    parts = [torch.empty_like(c) for _ in range(world)]
    dist.all_gather(parts, c)
    exact = sum(p.double() for p in parts)
    dist.all_reduce(c)
    torch.cuda.synchronize()
    err_mine = float((mine2.double() - exact).abs().max())
    err_nccl = float((c.double() - exact).abs().max())
    if rank == 0:
        print(json.dumps({"bench": "sweep_parity_vs_nccl", "fp32_allclose_rtol1e-3_atol1e-5": bool(ok),
                          "bf16_max_err_vs_exact_libb2d": err_mine, "bf16_max_err_vs_exact_nccl": err_nccl,
                          "bf16_libb2d_not_worse": err_mine <= err_nccl + 1e-12,
                          "bf16_max_abs_diff_vs_nccl": float((mine2 - c.float()).abs().max())}), flush=True)
    comm.close()
    dist.destroy_process_group()


def ncu_target():
    """A tiny multi-rank workload meant to run with EVERY rank under its own ncu (single-pass metrics, no kernel
    replay): a few allreduce calls of one size.  See tools/gpu_runs/ncu_multirank.sh."""
    import torch.distributed as dist
    from ray_lightning_b200.comm import Communicator
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")          # control plane only; keeps NCCL kernels out of the capture
    wire_bytes = int(os.environ.get("B2D_NCU_WIRE_BYTES", str(16 << 20)))
    algo = os.environ.get("B2D_NCU_ALGO", "two_shot")
    comm = Communicator(rank, world, local, 1 << 30, mem="vmm", timeout_ms=5000, nvls="auto",
                        max_ctas=int(os.environ.get("B2D_MAX_CTAS", "64")))
    buf = torch.randn(wire_bytes // 2, device="cuda") * 0.01
    for i in range(6):
        comm.allreduce_(buf, bucket_idx=0, wire="bf16", algo=algo)
        torch.cuda.synchronize()
        dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    {"k0": k0, "loopback": loopback, "sweep": sweep, "ncu_target": ncu_target}[sys.argv[1]]()
