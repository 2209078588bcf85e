"""Run the reference's ACTUAL gradient-sync implementation on CPU: torch DDP over gloo.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/ddp_oracle.py header).

What ``RayStrategy(num_workers=W, use_gpu=False)`` executes inside each Ray actor is
``init_process_group("gloo", rank, world, init_method="env://")`` (ray_lightning/ray_ddp.py:192-196)
followed by PL wrapping the module in ``DistributedDataParallel(module, **ddp_kwargs)``
(ray_lightning/ray_ddp.py:112-116).  ray / pytorch_lightning are not installable here, so the W
workers are started with ``torch.multiprocessing`` instead of Ray actors; the code below the
launcher is the same torch code the reference reaches.

Used for: (a) tests/golden fixtures (oracle/make_golden.py), (b) bench.py's ``cpu_baseline`` and
``--impl reference`` legs.
"""
import os
import socket
import time
from contextlib import closing

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn
from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
from torch.nn.parallel import DistributedDataParallel as DDP


def _init_gloo(rank, world, port):
    """The reference's rendezvous is env:// (ray_ddp.py:192-196); here the same TCPStore rendezvous is addressed
    explicitly so that it also works when this runner is itself started under torchrun (whose
    TORCHELASTIC_USE_AGENT_STORE would make env:// look for the launcher's store instead of ours)."""
    for k in list(os.environ):
        if k.startswith("TORCHELASTIC_"):
            os.environ.pop(k)
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method="tcp://127.0.0.1:%d" % port)


def free_port():
    with closing(socket.socket(socket.AF_INET, socket.SOCK_STREAM)) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# ---- models ---------------------------------------------------------------------------------
class SmallMLP(nn.Module):
    """Odd sizes on purpose: ragged bucket tails, several buckets at a tiny cap."""

    def __init__(self):
        super().__init__()
        self.l1 = nn.Linear(37, 53)
        self.l2 = nn.Linear(53, 29)
        self.l3 = nn.Linear(29, 11)

    def forward(self, x):
        return self.l3(torch.relu(self.l2(torch.relu(self.l1(x)))))


class MNISTClassifier(nn.Module):
    """Network of the reference's example/test model (ray_lightning/tests/utils.py:99-123,
    config ray_lightning/examples/ray_ddp_example.py:167): 784 -> 32 -> 64 -> 10."""

    def __init__(self, layer_1=32, layer_2=64):
        super().__init__()
        self.layer_1 = nn.Linear(28 * 28, layer_1)
        self.layer_2 = nn.Linear(layer_1, layer_2)
        self.layer_3 = nn.Linear(layer_2, 10)

    def forward(self, x):
        x = x.view(x.size(0), -1)
        x = torch.relu(self.layer_1(x))
        x = torch.relu(self.layer_2(x))
        return torch.log_softmax(self.layer_3(x), dim=1)


def make_model(name):
    if name == "small_mlp":
        return SmallMLP()
    if name == "mnist":
        return MNISTClassifier()
    if name in ("resnet50", "resnet18"):
        import torchvision
        return getattr(torchvision.models, name)()
    raise ValueError(name)


def make_batch(name, batch, seed):
    g = torch.Generator().manual_seed(seed)
    if name == "small_mlp":
        return torch.randn(batch, 37, generator=g), torch.randint(0, 11, (batch,), generator=g)
    if name == "mnist":
        return torch.rand(batch, 1, 28, 28, generator=g), torch.randint(0, 10, (batch,), generator=g)
    if name in ("resnet50", "resnet18"):
        return torch.randn(batch, 3, 224, 224, generator=g), torch.randint(0, 1000, (batch,), generator=g)
    raise ValueError(name)


def loss_fn(name, out, y):
    if name == "mnist":
        return nn.functional.nll_loss(out, y)
    return nn.functional.cross_entropy(out, y)


HOOKS = {
    "default": None,
    "allreduce_hook": default_hooks.allreduce_hook,
    "bf16_compress_hook": default_hooks.bf16_compress_hook,
}


# ---- one gradient sync, recorded --------------------------------------------------------------
def _grad_sync_worker(rank, world, port, cfg, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    _init_gloo(rank, world, port)
    try:
        name = cfg["model"]
        torch.manual_seed(cfg.get("seed", 0))
        model = make_model(name)
        x, y = make_batch(name, cfg.get("batch", 4), 1000 + rank)

        # local (pre-sync) gradients of this rank
        model.zero_grad()
        loss_fn(name, model(x), y).backward()
        local = [p.grad.detach().clone() for p in model.parameters()]
        model.zero_grad()

        out = {"local_grads": [g.numpy() for g in local]}
        for mode in cfg.get("modes", list(HOOKS)):
            torch.manual_seed(cfg.get("seed", 0))
            m2 = make_model(name)
            m2.load_state_dict(model.state_dict())
            ddp = DDP(m2, **cfg.get("ddp_kwargs", {}))
            layout = []
            hook = HOOKS[mode]

            def recording(state, bucket, _hook=hook, _layout=layout):
                # bucket descriptor handed to every comm hook (comm.hpp:20-98)
                params = bucket.parameters()
                ids = [p.data_ptr() for p in m2.parameters()]
                _layout.append({
                    "index": bucket.index(),
                    "numel": bucket.buffer().numel(),
                    "param_ids": [ids.index(p.data_ptr()) for p in params],
                    "lengths": [g.numel() for g in bucket.gradients()],
                    "is_last": bucket.is_last(),
                    "local_flat": bucket.buffer().detach().clone().numpy(),
                })
                if _hook is None:
                    return default_hooks.allreduce_hook(state, bucket)
                return _hook(state, bucket)

            # "default" = no hook registered at all (the C++ built-in path); its layout is recorded
            # through the allreduce_hook run, which is bit-identical
            if mode != "default":
                ddp.register_comm_hook(None, recording)
            loss_fn(name, ddp(x), y).backward()
            out[mode] = {"grads": [p.grad.detach().clone().numpy() for p in m2.parameters()],
                         "layout": layout}
            del ddp
        if rank in cfg.get("return_ranks", [0]):
            ret[rank] = out
        dist.barrier()
    finally:
        dist.destroy_process_group()


def run_grad_sync(world, cfg):
    """Run one DDP backward per mode on ``world`` gloo ranks; returns {rank: record}."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = free_port()
    cfg = dict(cfg)
    cfg.setdefault("return_ranks", list(range(world)))
    mp.spawn(_grad_sync_worker, args=(world, port, cfg, ret), nprocs=world, join=True)
    return {r: ret[r] for r in sorted(ret.keys())}


# ---- timed training steps (cpu baseline / reference arm) ----------------------------------------
def _train_worker(rank, world, port, cfg, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    threads = max(1, cfg.get("threads_total", os.cpu_count() or 1) // world)
    torch.set_num_threads(threads)
    _init_gloo(rank, world, port)
    try:
        name = cfg["model"]
        torch.manual_seed(0)
        model = make_model(name)
        ddp = DDP(model, **cfg.get("ddp_kwargs", {}))
        hook = HOOKS[cfg.get("hook", "default")]
        if hook is not None:
            ddp.register_comm_hook(None, hook)
        opt = torch.optim.SGD(ddp.parameters(), lr=0.01, momentum=0.9)
        batch = cfg.get("batch", 8)
        x, y = make_batch(name, batch, 1000 + rank)
        times = []
        budget = float(cfg.get("time_budget_s", 1e9))
        t_start = time.perf_counter()
        stop = torch.zeros(1)
        for step in range(cfg.get("warmup", 1) + cfg.get("steps", 2)):
            dist.barrier()
            t0 = time.perf_counter()
            opt.zero_grad(set_to_none=True)
            loss = loss_fn(name, ddp(x), y)
            loss.backward()
            opt.step()
            float(loss.detach())
            dist.barrier()
            if step >= cfg.get("warmup", 1):
                times.append(time.perf_counter() - t0)
            # bounded sample: rank 0 ends the run once the wall-clock budget is spent (>= 1 timed step)
            if rank == 0 and times and time.perf_counter() - t_start > budget:
                stop.fill_(1)
            dist.broadcast(stop, src=0)
            if stop.item() > 0:
                break
        if rank == 0:
            ret["times"] = times
            ret["threads_per_rank"] = threads
    finally:
        dist.destroy_process_group()


def run_training(world, cfg):
    """Timed DDP/gloo training steps on CPU. Returns dict(times=[s per step], threads_per_rank)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_train_worker, args=(world, free_port(), dict(cfg), ret), nprocs=world, join=True)
    return dict(ret)


def _allreduce_worker(rank, world, port, cfg, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    _init_gloo(rank, world, port)
    try:
        res = {}
        for nbytes in cfg["sizes"]:
            t = torch.randn(nbytes // 4)
            for _ in range(cfg.get("warmup", 2)):
                dist.all_reduce(t)
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(cfg.get("iters", 5)):
                t.div_(world)
                dist.all_reduce(t)
            dist.barrier()
            res[nbytes] = (time.perf_counter() - t0) / cfg.get("iters", 5)
        if rank == 0:
            ret["times"] = res
    finally:
        dist.destroy_process_group()


def run_allreduce_sweep(world, sizes, iters=5, warmup=2):
    """fp32 divide+allreduce over gloo, seconds per call per payload size (bytes)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_allreduce_worker, args=(world, free_port(), {"sizes": list(sizes), "iters": iters, "warmup": warmup}, ret),
             nprocs=world, join=True)
    return dict(ret["times"])
