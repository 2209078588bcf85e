#!/usr/bin/env python
"""bench.py — ResNet-50 training under RayStrategy, gradient sync on libb2d (driver contract).

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU/gloo path

Workload (BASELINE.json configs[1]): torchvision resnet50 (random init, seed 0), synthetic
224x224 images, per-GPU batch 64, bf16 autocast, SGD momentum, DDP with
find_unused_parameters=False / gradient_as_bucket_view=True / bucket_cap_mb=25 — driven through
``RayStrategy``'s worker-side path (the same calls RayLauncher._wrapping_function makes), whose
DDP comm hook is libb2d's fused allreduce.  One "step" = forward + backward (+ per-bucket
gradient sync, overlapped) + optimizer step on every rank.  Weak scaling: per-GPU batch fixed.

Output: ONE JSON line on rank 0 (keys documented in DESIGN.md §8).
"""
import argparse
import json
import os
import socket
import statistics
import subprocess
import sys
import threading
import time
from contextlib import closing

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec, ResNet-50 RayStrategy (+ allreduce bus GB/s)"
NVLINK_PEAK_GBS = 770.0   # /opt/skills/guides/B200_PROFILING.md: measured peer copy per direction (fallback: not in MEASURED_PEAKS.json)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: 64 resnet, 16 bert, 4 gpt2)")
    ap.add_argument("--bucket-cap-mb", type=int, default=25)
    ap.add_argument("--wire", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--mem", default="vmm", choices=["vmm", "ipc"])
    ap.add_argument("--max-ctas", type=int, default=None)
    ap.add_argument("--model", default="resnet50", choices=["resnet50", "resnet18", "bert-base", "gpt2-medium"])
    ap.add_argument("--strategy", default="ddp", choices=["ddp", "sharded"],
                    help="ddp = RayStrategy; sharded = RayShardedStrategy (fused reduce-scatter + Adam + all-gather)")
    ap.add_argument("--hook", default="b200", choices=["b200", "nccl_bf16", "nccl_fp32"],
                    help="b200 = libb2d; nccl_* = the reference's GPU path through the same strategy (A/B)")
    ap.add_argument("--seq", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=8)
    return ap.parse_args()


def free_port():
    with closing(socket.socket(socket.AF_INET, socket.SOCK_STREAM)) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (profiling recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def usable_cores():
    """Host threads this process can really run: affinity mask, capped by the cgroup CPU quota, and by
    B2D_CPU_THREADS when set (oversubscribing a quota makes the CPU arm pathologically slow)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    cap = int(os.environ.get("B2D_CPU_THREADS", "64"))   # beyond ~64 threads torch's CPU conv backward stops scaling
    return max(1, min(n, cap))


# ---- the reference arm / cpu baseline: torch DDP over gloo on the host cores ---------------------
def cpu_reference(world, batch, steps, warmup, model="resnet50", budget_s=20.0):
    """What RayStrategy(num_workers=world, use_gpu=False) executes in its workers (oracle/reference_ddp.py)."""
    from oracle import reference_ddp
    cores = usable_cores()
    cfg = {"model": model, "batch": batch, "steps": steps, "warmup": warmup, "threads_total": cores,
           "time_budget_s": budget_s, "ddp_kwargs": {"find_unused_parameters": False, "gradient_as_bucket_view": True}}
    t0 = time.time()
    res = reference_ddp.run_training(world, cfg)
    ms = 1e3 * statistics.mean(res["times"])
    return {"value": world * batch / (ms / 1e3), "unit": "images/sec", "cores": cores, "kind": "reference",
            "steps_done": len(res["times"]),
            "sample": "torch DDP/gloo fp32 (the implementation ray_lightning's use_gpu=False path dispatches to; Ray actors "
                      "replaced by torch.multiprocessing), %s, %d worker(s) x batch %d, %d threads/worker, %d warm-up + %d "
                      "timed steps, %.0f s wall" % (model, world, batch, res["threads_per_rank"], warmup, len(res["times"]), time.time() - t0),
            "ms_per_step": ms}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # --steps / --warmup are honoured; each step is a bounded sample (per-worker batch --cpu-batch) and the
    # run stops early once ~150 s of wall clock are spent, reporting the steps actually timed
    steps, warm = max(1, args.steps), max(1, min(args.warmup, 3))
    try:
        cb = cpu_reference(args.gpus, args.cpu_batch, steps, warm, args.model, budget_s=150.0)
        steps = cb["steps_done"]
    except Exception as e:  # the oracle always exists; a failure here is a bug worth seeing
        print(json.dumps({"impl": "reference", "unavailable": "cpu reference failed: %r" % (e,)}))
        return
    line = {"impl": "reference", "metric": METRIC, "value": round(cb["value"], 2), "unit": "images/sec",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": round(cb["ms_per_step"], 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "%s synthetic 224x224, RayStrategy(num_workers=%d, use_gpu=False) == torch DDP/gloo on host cores"
                                   % (args.model, args.gpus), "per_worker_batch": args.cpu_batch,
                       "global_batch": args.cpu_batch * args.gpus, "parallelism": "dp%d" % args.gpus,
                       "bounded_sample": "%d timed steps" % steps},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": round(cb["value"], 2), "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---- our arm -----------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    from ray_lightning_b200 import RayStrategy
    from ray_lightning_b200._compat import LightningModule

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 gradient-sync path has no CPU fallback "
                         "(use --impl reference for the CPU/gloo arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torchrun --nproc-per-node %d" % (args.gpus, world, args.gpus))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(free_port()))
    os.environ.setdefault("PL_TORCH_DISTRIBUTED_BACKEND", "nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True

    import torchvision
    torch.manual_seed(0)
    if args.model.startswith("resnet"):
        B = args.batch or 64
        unit = "images"

        class Net(LightningModule):
            def __init__(self):
                super().__init__()
                self.net = getattr(torchvision.models, args.model)()

            def training_step(self, batch, batch_idx):
                x, y = batch
                return F.cross_entropy(self.net(x), y)

            def configure_optimizers(self):
                return torch.optim.SGD(self.parameters(), lr=0.05, momentum=0.9)

        def host_batch(g):
            return (torch.randn(B, 3, 224, 224, generator=g).contiguous(memory_format=torch.channels_last),
                    torch.randint(0, 1000, (B,), generator=g))
    else:
        import transformers
        unit = "sequences"
        if args.model == "bert-base":      # BASELINE.json configs[2]
            B, S = args.batch or 16, args.seq or 512
            cfg = transformers.BertConfig()
            make = lambda: transformers.BertForMaskedLM(cfg)
        else:                               # gpt2-medium, BASELINE.json configs[3]
            B, S = args.batch or 4, args.seq or 1024
            cfg = transformers.GPT2Config(n_embd=1024, n_layer=24, n_head=16)
            make = lambda: transformers.GPT2LMHeadModel(cfg)
        vocab = cfg.vocab_size

        class Net(LightningModule):
            def __init__(self):
                super().__init__()
                self.net = make()

            def training_step(self, batch, batch_idx):
                ids, = batch
                return self.net(input_ids=ids, labels=ids).loss

            def configure_optimizers(self):
                return (torch.optim.Adam if args.strategy == "sharded" else torch.optim.AdamW)(self.parameters(), lr=1e-4)

        def host_batch(g):
            return (torch.randint(0, vocab, (B, S), generator=g),)

    # the worker-side call sequence of RayLauncher._wrapping_function (launchers/ray_launcher.py)
    from ray_lightning_b200 import RayShardedStrategy
    kw = dict(num_workers=world, use_gpu=True, b200_wire=args.wire, b200_algo=args.algo, b200_mem=args.mem,
              b200_timing=True, b200_max_ctas=args.max_ctas)
    if args.strategy == "sharded":
        strategy = RayShardedStrategy(**kw)
    else:
        kw.update(find_unused_parameters=False, gradient_as_bucket_view=True, bucket_cap_mb=args.bucket_cap_mb)
        if args.hook != "b200":
            kw["b200_enable"] = False
            if args.hook == "nccl_bf16":
                from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
                kw["ddp_comm_hook"] = default_hooks.bf16_compress_hook
        strategy = RayStrategy(**kw)
    strategy.precision = "bf16"
    strategy.set_remote(True)
    strategy.set_global_to_local([(i, 0) for i in range(world)])
    strategy.root_device = dev
    strategy._worker_setup(process_idx=rank)
    model = Net()
    if args.model.startswith("resnet"):
        model = model.to(memory_format=torch.channels_last)
    strategy.connect(model)
    strategy.model_to_device()
    strategy.configure_ddp()

    class _T:  # the two Trainer attributes setup_optimizers looks at
        pass
    strategy.setup_optimizers(_T())
    opt = strategy.optimizers[0]
    n_params = sum(p.numel() for p in model.parameters())

    g = torch.Generator().manual_seed(1000 + rank)
    host = tuple(t.pin_memory() for t in host_batch(g))
    devb = tuple(t.to(dev, non_blocking=True) for t in host)

    def step(batch, i):
        opt.zero_grad(set_to_none=True) if args.strategy == "ddp" else opt.zero_grad()
        loss = strategy.training_step(batch, i)
        strategy.backward(loss)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, e2e):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        last = None
        for i in range(n):
            if e2e:
                batch = tuple(t.to(dev, non_blocking=True) for t in host)
                last = float(step(batch, i))  # device->host read of the step's result, every step
            else:
                last = step(devb, i)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t), last

    for i in range(args.warmup):
        step(devb, i)
    barrier()
    state = strategy.b200_state if args.strategy == "ddp" else None
    comm = state.comm if state is not None else getattr(strategy, "_comm", None)
    if comm is not None:
        comm.ctx.reset_stats()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")
    phys = int(vis[local]) if len(vis) > local and vis[local].isdigit() else local
    sampler = ClockSampler(phys)
    if rank == 0:
        sampler.start()
    total_ms, _ = timed(args.steps, e2e=False)
    clocks = sampler.stop() if rank == 0 else None
    torch.cuda.synchronize()
    st = comm.stats() if comm is not None else {"launches": 0, "timed_ms": 0.0, "timed_launches": 0}
    launches_timed, kernel_ms = int(st["launches"]), float(st["timed_ms"])
    timed_launches = int(st["timed_launches"])
    e2e_ms, last_loss = timed(args.steps, e2e=True)

    # The same buckets once more, ISOLATED (no backward running, ranks aligned by a barrier): what the
    # kernel does when it is not waiting for SMs or for a slower peer.  Through the same hook entry point.
    isolated = None
    if state is not None and comm is not None and getattr(state, "seen", None):
        isolated = []
        for idx, n in sorted(state.seen.items()):
            buf = torch.randn(n, device=dev) * 0.01
            barrier()
            for _ in range(3):
                comm.allreduce_(buf, bucket_idx=idx, wire=args.wire, algo=args.algo, wait_stream=state.stream,
                                comm_stream=state.stream)
            barrier()
            it = 20
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(state.stream):
                a.record()
                for _ in range(it):
                    comm.allreduce_(buf, bucket_idx=idx, wire=args.wire, algo=args.algo, wait_stream=state.stream,
                                    comm_stream=state.stream)
                b.record()
            b.synchronize()
            t = torch.tensor([a.elapsed_time(b) / it], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            isolated.append((idx, n, float(t)))
        barrier()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        ms_per_step = total_ms / args.steps
        value = world * B / (ms_per_step / 1e3)
        e2e_value = world * B / (e2e_ms / args.steps / 1e3)
        wire_w = 2 if args.wire == "bf16" else 4
        per_launch_ms = kernel_ms / max(timed_launches, 1)
        buckets_per_step = launches_timed / args.steps
        if world == 1:
            # K0: 4 B read + 4 B write per gradient element, nothing else
            alg_bytes_step = 8.0 * n_params
            if args.strategy == "sharded":   # stage (4 r + w w), reduce (w r), Adam p/m/v r+w + p (28), DESIGN.md §4
                alg_bytes_step = (4.0 + 2 * wire_w + 28.0) * n_params
            bound, peak, runit = "hbm", float(peaks["hbm_gbs"]), "GB/s"
            peak_note = "MEASURED_PEAKS.json hbm_gbs (%s)" % peak_src
        else:
            alg_bytes_step = 2.0 * (world - 1) / world * n_params * wire_w   # NCCL-tests bus-bandwidth convention
            if args.strategy == "sharded":   # reduce-scatter at wire width + fp32 parameter all-gather
                alg_bytes_step = (world - 1) / world * n_params * (wire_w + 4.0)
            bound, peak, runit = "nvlink", NVLINK_PEAK_GBS, "GB/s"
            peak_note = "fallback: B200_PROFILING.md measured peer copy 770 GB/s per direction (not in MEASURED_PEAKS.json)"
        achieved = alg_bytes_step * args.steps / (kernel_ms / 1e3) / 1e9 if kernel_ms > 0 else None
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            if world == 1 and "k0_dram_bytes_per_element" in tj:
                traffic = tj["k0_dram_bytes_per_element"] * n_params / max(buckets_per_step, 1)
        except Exception:
            pass
        line = {
            "metric": METRIC if unit == "images" else "%s/sec, %s %s" % (unit, args.model, "RayShardedStrategy" if args.strategy == "sharded" else "RayStrategy"),
            "value": round(value, 2), "unit": unit + "/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.wire == "bf16" else "fp32", "data": "synthetic",
            "config": {"workload": "%s synthetic %s %s(num_workers=%d, use_gpu=True) bf16-autocast, gradient sync = %s"
                                   % (args.model, "224x224" if unit == "images" else "token ids",
                                      "RayShardedStrategy" if args.strategy == "sharded" else "RayStrategy", world,
                                      ("libb2d (%s wire)" % args.wire) if args.hook == "b200" else args.hook),
                       "global_batch": B * world, "per_gpu_batch": B, "parallelism": "dp%d" % world,
                       "bucket_cap_mb": args.bucket_cap_mb, "grad_elements": n_params,
                       "l2_policy": "inputs larger than L2 (activations + 97.5 MiB of gradients per step >> 126 MB)",
                       "algo": args.algo, "mem": args.mem, "nvls_bound": bool(getattr(comm, "nvls", False)),
                       "strategy": args.strategy, "hook": args.hook},
            "e2e": {"value": round(e2e_value, 2), "unit": unit + "/sec", "ms_per_step": round(e2e_ms / args.steps, 3),
                    "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in host)) * world,
                    "d2h_bytes_per_step": 4 * world,
                    "api": "RayStrategy worker path: training_step/backward/optimizer.step with pinned-host batches, loss read back"},
            "gpu_launches": launches_timed,
            "roofline": {"bound": bound, "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": runit,
                         "frac": round(achieved / peak, 4) if achieved else None, "traffic": traffic,
                         "kernel": ("k456_sharded_kernel" if args.strategy == "sharded" else
                                    "k0_cast_scale_kernel<bf16>" if world == 1 else "k1/k2 fused allreduce") if args.hook == "b200" else None,
                         "algorithmic_bytes_per_step": alg_bytes_step, "launches_per_step": buckets_per_step,
                         "avg_launch_ms": round(per_launch_ms, 5), "kernel_share_of_step": round(kernel_ms / total_ms, 5),
                         "peak_source": peak_note,
                         "note": "launch durations from CUDA events on the comm stream inside the timed region (overlapped with backward)"},
            "clocks": clocks, "final_loss": last_loss if isinstance(last_loss, float) else float(last_loss),
            "allreduce_isolated": None if not isolated else {
                "note": "same bucket sizes, back to back on the comm stream with no backward running (L2-warm), max over ranks",
                "buckets": [{"index": i, "elements": n, "ms": round(ms, 5),
                             "GBps": round((8.0 * n if world == 1 else 2.0 * (world - 1) / world * n * (2 if args.wire == "bf16" else 4)) / ms / 1e6, 1)}
                            for i, n, ms in isolated],
                "GBps_total": round(sum((8.0 * n if world == 1 else 2.0 * (world - 1) / world * n * (2 if args.wire == "bf16" else 4))
                                        for _, n, _ in isolated) / sum(ms for _, _, ms in isolated) / 1e6, 1),
                "unit": "HBM GB/s (8 B/element)" if world == 1 else "NVLink bus GB/s (2(W-1)/W x wire bytes)"},
        }
        if not args.no_cpu_baseline and world == 1 and args.model.startswith("resnet"):
            try:
                cb = cpu_reference(1, args.cpu_batch, 2, 1, args.model)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "images/sec", "cores": None, "kind": "reference",
                                        "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    barrier()
    strategy.teardown_worker()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
