"""The one-process-per-GPU deployment (Ray-actor model) on whatever GPUs the box has: W worker
processes, control plane over gloo, arenas shared through CUDA IPC (and VMM fds); when the box has
fewer GPUs than ranks the ranks share cuda:0 — the IPC/handle/barrier path is the same."""
import os
import socket
from contextlib import closing

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _port():
    with closing(socket.socket(socket.AF_INET, socket.SOCK_STREAM)) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mem, ret):
    import numpy as np
    import torch.distributed as dist
    from oracle import ddp_oracle
    from ray_lightning_b200.comm import Communicator
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method="env://")
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    comm = Communicator(rank, world, dev, 128 << 20, mem=mem, timeout_ms=20000, max_ctas=128 // world, nvls="auto")
    ok = True
    try:
        for k, n in enumerate([5, 4099, 1 << 20]):
            per_rank = [torch.randn(n, generator=torch.Generator().manual_seed(77 + r + 10 * k)) * 0.05
                        for r in range(world)]
            for wire in ("bf16", "fp32"):
                for ai, algo in enumerate(("one_shot", "two_shot", "two_shot_tma", "staged")):
                    if algo == "two_shot_tma" and wire == "fp32":
                        continue
                    buf = per_rank[rank].cuda()
                    comm.allreduce_(buf, bucket_idx=k * 10 + (wire == "bf16") * 4 + ai, wire=wire, algo=algo)
                    torch.cuda.synchronize()
                    fn = ddp_oracle.allreduce_bf16_wire if wire == "bf16" else ddp_oracle.allreduce_fp32_wire
                    want = fn(per_rank)
                    ok = ok and torch.equal(buf.cpu().view(torch.int32), want.view(torch.int32))
                if comm.nvls:   # real NVSwitch multicast between the worker processes: tolerance contract
                    buf = per_rank[rank].cuda()
                    comm.allreduce_(buf, bucket_idx=900 + k * 2 + (wire == "bf16"), wire=wire, algo="nvls")
                    torch.cuda.synchronize()
                    fn = ddp_oracle.allreduce_bf16_wire if wire == "bf16" else ddp_oracle.allreduce_fp32_wire
                    ok = ok and torch.allclose(buf.cpu(), fn(per_rank), rtol=1e-2 if wire == "bf16" else 1e-5, atol=1e-5)
        st = comm.stats()
        ret[rank] = {"ok": ok, "nvls": comm.nvls, "launches": st["launches"], "mem_kind": st["mem_kind"]}
    finally:
        comm.close()
        dist.destroy_process_group()


@pytest.mark.parametrize("mem", ["ipc", "vmm"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_worker_processes_share_arenas(world, mem):
    if world == 8 and torch.cuda.device_count() < 8:
        pytest.skip("world 8 runs on the 8-GPU box only (8 CUDA contexts on one device add nothing over world 4)")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _port(), mem, ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == list(range(world))
    for r in range(world):
        assert ret[r]["ok"], (r, dict(ret[r]))
        # 3 sizes x (5 single-kernel calls + staged bf16 + staged fp32 = 2 x (stage, exchange, wait, write-back)) [+ NVLS]
        assert ret[r]["launches"] == 3 * (5 + 8) + (3 * 8 if ret[r]["nvls"] else 0)
        assert ret[r]["mem_kind"] == (1 if mem == "vmm" else 0)


def _ddp_worker(rank, world, port, ret):
    """torch DDP with the libb2d hook vs (a) stock DDP, (b) the oracle on the local gradients — several
    iterations, several buckets, including the bucket re-layout after iteration 1."""
    import torch.distributed as dist
    import torch.nn as nn
    from torch.nn.parallel import DistributedDataParallel as DDP
    from oracle import ddp_oracle
    from ray_lightning_b200.comm import B200HookState, b200_allreduce_hook
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method="env://")
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)

    def make():
        torch.manual_seed(0)
        return nn.Sequential(nn.Linear(37, 531), nn.ReLU(), nn.Linear(531, 257), nn.ReLU(), nn.Linear(257, 11)).to(dev)

    def batch(it):
        g = torch.Generator().manual_seed(100 * it + rank)
        return torch.randn(8, 37, generator=g).to(dev), torch.randint(0, 11, (8,), generator=g).to(dev)

    ok, info = True, {}
    states = []
    try:
        for wire in ("fp32", "bf16"):
            kw = dict(device_ids=[dev.index], bucket_cap_mb=0.25, find_unused_parameters=False)
            ours, stock, plain = DDP(make(), **kw), DDP(make(), **kw), make()
            st = B200HookState(wire=wire, total_grad_elems=sum(p.numel() for p in plain.parameters()), mem="ipc",
                               max_ctas=32)
            states.append(st)
            ours.register_comm_hook(st, b200_allreduce_hook)
            for it in range(4):
                x, y = batch(it)
                for m in (ours, stock, plain):
                    m.zero_grad(set_to_none=True)
                    nn.functional.cross_entropy(m(x), y).backward()
                torch.cuda.synchronize()
                local = [p.grad.detach().cpu().clone() for p in plain.parameters()]
                gathered = [None] * world
                dist.all_gather_object(gathered, local)
                for i, (po, ps) in enumerate(zip(ours.parameters(), stock.parameters())):
                    per_rank = [gathered[r][i].reshape(-1) for r in range(world)]
                    if wire == "fp32":
                        # world 2: one fp32 add per element -> identical to stock DDP (gloo) bit for bit
                        ok = ok and torch.equal(po.grad, ps.grad)
                        want = ddp_oracle.allreduce_fp32_wire(per_rank)
                    else:
                        want = ddp_oracle.allreduce_bf16_wire(per_rank)
                    ok = ok and torch.equal(po.grad.cpu().reshape(-1).view(torch.int32), want.view(torch.int32))
            info[wire] = {"calls": st.calls, "buckets_seen": len(st.seen)}
        ret[rank] = {"ok": bool(ok), **info}
    finally:
        for st in states:
            st.close()
        dist.destroy_process_group()


def test_ddp_with_the_hook_matches_stock_ddp_and_the_oracle():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, _port(), ret), nprocs=2, join=True)
    for r in range(2):
        assert ret[r]["ok"], dict(ret[r])
        assert ret[r]["fp32"]["calls"] >= 4 and ret[r]["fp32"]["buckets_seen"] >= 2   # several buckets after the re-layout


def _bn_worker(rank, world, port, ret):
    """DDP's per-forward buffer broadcast (BatchNorm statistics from rank 0) through libb2d's peer
    stores (b200_buffer_hook) instead of a coalesced ncclBroadcast — same values as stock DDP on every rank, and the
    gradient buckets living in the arena (fp32 wire, exchanged in place) at the same time."""
    import torch.distributed as dist
    import torch.nn as nn
    from torch.nn.parallel import DistributedDataParallel as DDP
    from torch.nn.parallel.distributed import _BufferCommHookLocation
    from ray_lightning_b200.comm import ArenaBufferSync, B200HookState, b200_allreduce_hook, b200_buffer_hook
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method="env://")
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)

    def make():
        torch.manual_seed(0)
        return nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.ReLU(), nn.Conv2d(8, 4, 3), nn.BatchNorm2d(4),
                             nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(4, 5)).to(dev)

    st = None
    try:
        n_el = sum(p.numel() for p in make().parameters())
        st = B200HookState(wire="fp32", total_grad_elems=n_el, mem="ipc", arena_buckets=True, arena_extra_bytes=8 << 20)
        st.ensure(dev)
        with st.allocate_in_arena():
            ours = DDP(make(), device_ids=[dev.index], gradient_as_bucket_view=True)
        symmetric = st.verify_symmetric_buckets()
        stock = DDP(make(), device_ids=[dev.index], gradient_as_bucket_view=True)
        ours.register_comm_hook(st, b200_allreduce_hook)
        bst = ArenaBufferSync(st)
        ours._register_buffer_comm_hook(bst, b200_buffer_hook, comm_hook_location=_BufferCommHookLocation.PRE_FORWARD)
        opts = [torch.optim.SGD(m.parameters(), lr=0.1) for m in (ours, stock)]
        ok = True
        for it in range(5):
            g = torch.Generator().manual_seed(100 * it + rank)
            x, y = torch.randn(6, 3, 12, 12, generator=g).to(dev), torch.randint(0, 5, (6,), generator=g).to(dev)
            if it == 1:                           # DDP lays its buckets out anew in the second forward: do it in the arena
                with st.allocate_in_arena():
                    ours.reducer._rebuild_buckets()
                symmetric = symmetric and st.verify_symmetric_buckets()
            for m, o in zip((ours, stock), opts):
                o.zero_grad(set_to_none=False)
                nn.functional.cross_entropy(m(x), y).backward()
                o.step()
            torch.cuda.synchronize()
        why = []
        for (n1, b1), (_, b2) in zip(ours.module.named_buffers(), stock.module.named_buffers()):
            if not st.comm.owns(b1):
                why.append("buffer %s is not in the arena" % n1)
            if not torch.allclose(b1.float(), b2.float(), rtol=1e-5, atol=1e-7):
                why.append("buffer %s differs by %g" % (n1, float((b1.float() - b2.float()).abs().max())))
        for (n1, p1), (_, p2) in zip(ours.named_parameters(), stock.named_parameters()):
            if not torch.allclose(p1, p2, rtol=1e-5, atol=1e-7):      # world 2, fp32: one add per element
                why.append("parameter %s differs by %g" % (n1, float((p1 - p2).abs().max())))
        ok = not why
        ret[rank] = {"ok": bool(ok), "why": why[:6], "buffer_syncs": bst.calls, "symmetric": bool(symmetric),
                     "in_arena": dict(st.in_arena), "algo": st.comm.stats()["last_algo"]}
        del ours
    finally:
        if st is not None:
            import gc
            gc.collect()
            st.close()
        dist.destroy_process_group()


def test_buffer_broadcast_and_arena_buckets_match_stock_ddp():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bn_worker, args=(2, _port(), ret), nprocs=2, join=True)
    for r in range(2):
        assert ret[r]["ok"], dict(ret[r])
        assert ret[r]["buffer_syncs"] == 5
        assert ret[r]["symmetric"] and all(ret[r]["in_arena"].values()), dict(ret[r])


def _inbw_worker(rank, world, port, ret):
    """the optimizer step applied per DDP bucket right behind its allreduce (K14) == stock DDP followed by
    torch's own optimizer.step(), across DDP's bucket re-layout, for SGD(momentum, weight decay) and AdamW."""
    import torch.distributed as dist
    import torch.nn as nn
    from torch.nn.parallel import DistributedDataParallel as DDP
    from ray_lightning_b200.comm import B200HookState, InBackwardOptimizer, b200_allreduce_hook
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method="env://")
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)

    def make_mlp():
        torch.manual_seed(0)
        return nn.Sequential(nn.Linear(37, 531), nn.ReLU(), nn.Linear(531, 257), nn.ReLU(), nn.Linear(257, 11)).to(dev)

    class Conv(nn.Module):
        def __init__(self):
            super().__init__()
            self.c1, self.c2, self.fc = nn.Conv2d(4, 16, 3, padding=1), nn.Conv2d(16, 8, 3, padding=1), nn.Linear(8, 11)

        def forward(self, x):
            x = x[:, :36].reshape(-1, 4, 3, 3).contiguous(memory_format=torch.channels_last)
            return self.fc(torch.relu(self.c2(torch.relu(self.c1(x)))).mean(dim=(2, 3)))

    def make_conv():     # channels_last weights: parameter, gradient view and state share the NHWC memory order
        torch.manual_seed(0)
        m = Conv().to(dev).to(memory_format=torch.channels_last)
        assert not m.c1.weight.is_contiguous() and m.c1.weight.is_contiguous(memory_format=torch.channels_last)
        return m

    ok, info, states = True, {}, []
    try:
        for name, mk, make in (("sgd", lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9, weight_decay=1e-2), make_mlp),
                               ("adamw", lambda ps: torch.optim.AdamW(ps, lr=1e-2, weight_decay=0.05), make_mlp),
                               ("sgd_nhwc", lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9), make_conv)):
            kw = dict(device_ids=[dev.index], bucket_cap_mb=0.25, gradient_as_bucket_view=True)
            ours, stock = DDP(make(), **kw), DDP(make(), **kw)
            st = B200HookState(wire="fp32", total_grad_elems=sum(p.numel() for p in stock.parameters()), mem="ipc")
            states.append(st)
            ours.register_comm_hook(st, b200_allreduce_hook)
            opt = InBackwardOptimizer(mk(ours.parameters()), st)
            sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
            ref_opt = mk(stock.parameters())
            ref_sched = torch.optim.lr_scheduler.StepLR(ref_opt, step_size=2, gamma=0.5)
            for it in range(6):
                g = torch.Generator().manual_seed(100 * it + rank)
                x, y = torch.randn(8, 37, generator=g).to(dev), torch.randint(0, 11, (8,), generator=g).to(dev)
                before = [p.detach().clone() for p in ours.parameters()]
                opt.zero_grad(set_to_none=False)
                nn.functional.cross_entropy(ours(x), y).backward()
                torch.cuda.synchronize()
                moved = all(not torch.equal(a, b) for a, b in zip(before, ours.parameters()))   # updated INSIDE backward
                ok = ok and moved
                opt.step(); sched.step()
                ref_opt.zero_grad(set_to_none=False)
                nn.functional.cross_entropy(stock(x), y).backward()
                ref_opt.step(); ref_sched.step()
                torch.cuda.synchronize()
            for a, b in zip(ours.parameters(), stock.parameters()):
                ok = ok and torch.allclose(a, b, rtol=2e-5, atol=2e-6)
            sd, ref_sd = opt.state_dict(), ref_opt.state_dict()
            ok = ok and sorted(sd["state"].keys()) == sorted(ref_sd["state"].keys())
            for i, s in ref_sd["state"].items():
                for k, v in s.items():
                    if isinstance(v, torch.Tensor) and v.dim() > 0:
                        ok = ok and torch.allclose(sd["state"][i][k], v, rtol=2e-5, atol=2e-6)
            info[name] = {"applied": opt.applied, "buckets": len(st.seen)}
        ret[rank] = {"ok": bool(ok), **info}
    finally:
        for st in states:
            st.close()
        dist.destroy_process_group()


def test_optimizer_in_backward_matches_stock_ddp_plus_optimizer():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_inbw_worker, args=(2, _port(), ret), nprocs=2, join=True)
    for r in range(2):
        assert ret[r]["ok"], dict(ret[r])
        # one bucket in the first iteration, two after DDP's re-layout
        assert ret[r]["sgd"]["applied"] >= 6 and ret[r]["adamw"]["applied"] >= 6 and ret[r]["sgd_nhwc"]["applied"] >= 6
