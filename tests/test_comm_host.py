"""Host-side logic of the N>1 path on CPU (gloo, world size 2 and 3): the SCM_RIGHTS fd exchange that
carries VMM arena handles between worker processes, and the sizing / option plumbing of the hook state."""
import os
import socket
from contextlib import closing

import pytest
import torch
import torch.multiprocessing as mp


def _port():
    with closing(socket.socket(socket.AF_INET, socket.SOCK_STREAM)) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fd_worker(rank, world, port, ret):
    import torch.distributed as dist
    from ray_lightning_b200.comm import _exchange_fds
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method="tcp://127.0.0.1:%d" % port)
    try:
        # every rank owns an anonymous file; its fd travels to every peer, who must then read the owner's bytes
        fd = os.memfd_create("b2d-test-%d" % rank)
        os.write(fd, b"from-%d" % rank)
        got = _exchange_fds(None, rank, world, "t%d" % port, fd)
        assert sorted(got) == [p for p in range(world) if p != rank]
        seen = {p: os.pread(f, 64, 0) for p, f in got.items()}
        # the one-sender variant (the multicast handle goes from rank 0 to everybody)
        fd2 = os.memfd_create("b2d-test-mc")
        os.write(fd2, b"mc")
        got2 = _exchange_fds(None, rank, world, "m%d" % port, fd2 if rank == 0 else -1, senders=[0])
        mc = os.pread(got2[0], 8, 0) if rank != 0 else b""
        ret[rank] = {"seen": seen, "mc": mc, "n2": len(got2)}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_fd_exchange_between_worker_processes(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fd_worker, args=(world, _port(), ret), nprocs=world, join=True)
    for rank in range(world):
        assert ret[rank]["seen"] == {p: b"from-%d" % p for p in range(world) if p != rank}
        assert ret[rank]["n2"] == (0 if rank == 0 else 1)
        assert ret[rank]["mc"] == (b"" if rank == 0 else b"mc")


def test_hook_state_sizing_and_pickling():
    import pickle
    from ray_lightning_b200.comm import B200HookState, arena_bytes_for
    assert arena_bytes_for(25557032) == 12 * 25557032 + (64 << 20)                                   # fp32 wire, staged
    assert arena_bytes_for(25557032, wire="bf16") == 6 * 25557032 + (64 << 20)
    assert arena_bytes_for(25557032, wire="fp32", arena_buckets=True, extra_bytes=5) == 8 * 25557032 + (320 << 20) + 5
    st = B200HookState(wire="fp32", algo="two_shot", total_grad_elems=1000, max_ctas=32)
    st.calls, st.seen = 5, {0: 10}
    st2 = pickle.loads(pickle.dumps(st))
    assert st2.comm is None and st2.stream is None and st2.wire == "fp32" and st2.max_ctas == 32
    with pytest.raises(ValueError):
        B200HookState().ensure(type("D", (), {"index": 0})())   # no size information -> refuse before touching CUDA


class _FakeCtx:
    def __init__(self, stats):
        self._stats, self.inplace, self.registered, self.applied = stats, None, [], []

    def set_inplace(self, v):
        self.inplace = bool(v)

    def optim_register(self, *a):
        self.registered.append(a)

    def bucket_optim(self, *a):
        self.applied.append(a)


class _FakeComm:
    """Host-logic double of comm.Communicator (CPU tensors, no libb2d): what the hook-side helpers touch."""
    world, rank, device_index = 1, 0, 0

    def __init__(self, stats=None):
        self.ctx = _FakeCtx(stats or {"pool_allocs": 3, "pool_digest": 99})
        self.pushed = []

    def stats(self):
        return self.ctx._stats

    def arena_tensor(self, numel, dtype=torch.float32):
        return torch.zeros(numel, dtype=dtype)

    def adam_push_(self, flat, m, v, red, shard_off, groups, **kw):
        self.pushed.append((flat.data_ptr(), list(shard_off), list(groups)))


def test_buffer_sync_moves_mixed_dtype_buffers_into_one_flat_region(monkeypatch):
    """f-3 host logic: BatchNorm-style buffers (fp32 vectors, an int64 scalar) become views of ONE flat region, values
    intact, every start 32-byte aligned; the push names rank 0 as the owner of everything."""
    from ray_lightning_b200.comm import ArenaBufferSync, B200HookState
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *_a, **_k: None)
    st = B200HookState(wire="fp32", total_grad_elems=10)
    st.comm = _FakeComm()
    bn = torch.nn.BatchNorm1d(5)
    bn.running_mean.copy_(torch.arange(5.0))
    bn.num_batches_tracked.fill_(7)
    bufs = dict(bn.named_buffers())
    sync = ArenaBufferSync(st)
    sync.sync(bufs)
    assert sync.nbytes % 32 == 0 and sync.flat.numel() * 4 == sync.nbytes
    base = sync.flat.data_ptr()
    for b in bufs.values():
        assert base <= b.data_ptr() < base + sync.nbytes and (b.data_ptr() - base) % 32 == 0
    assert torch.equal(bn.running_mean, torch.arange(5.0)) and int(bn.num_batches_tracked) == 7
    bn.running_var.mul_(3.0)                                    # in-place updates land in the flat region
    off = (bn.running_var.data_ptr() - base) // 4
    assert torch.equal(sync.flat[off:off + 5], torch.full((5,), 3.0))
    assert st.comm.pushed[0][1] == [0, sync.nbytes // 4] and st.comm.pushed[0][2] == [] and sync.calls == 1


def test_in_place_exchange_is_switched_off_when_ranks_disagree(monkeypatch):
    """f-1 host logic: the digest of pool allocations decides whether arena-resident buckets may be exchanged in place."""
    import torch.distributed as dist
    from ray_lightning_b200.comm import B200HookState
    st = B200HookState(wire="fp32", total_grad_elems=10)
    st.comm = _FakeComm()
    assert st.verify_symmetric_buckets() is True and st.comm.ctx.inplace is True
    st.comm.world = 2
    monkeypatch.setattr(dist, "is_initialized", lambda: True)
    monkeypatch.setattr(dist, "all_gather_object", lambda out, obj, group=None: out.__setitem__(slice(None), [obj, (obj[0], obj[1] + 1)]))
    assert st.verify_symmetric_buckets() is False and st.comm.ctx.inplace is False


def test_in_backward_optimizer_state_dict_round_trip():
    """f-2 host logic: per-parameter state in torch's own state-dict layout; step count, lr schedulers, load."""
    from ray_lightning_b200.comm import B200HookState, InBackwardOptimizer
    st = B200HookState(wire="fp32", total_grad_elems=10)
    model = torch.nn.Linear(4, 3)
    opt = InBackwardOptimizer(torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.1), st)
    assert st.in_backward is opt and opt.kind == 1 and opt.param_groups[0]["weight_decay"] == 0.1
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)
    assert opt.state_dict()["state"] == {}
    for p in model.parameters():                     # what apply_bucket would have filled during backward
        s1, s2 = opt._states(p)
        s1.fill_(0.25); s2.fill_(0.5)
    opt.step(); sched.step()
    sd = opt.state_dict()
    assert sd["param_groups"][0]["lr"] == 5e-3 and sd["param_groups"][0]["params"] == [0, 1]
    assert float(sd["state"][1]["step"]) == 1.0 and torch.equal(sd["state"][0]["exp_avg"], torch.full((3, 4), 0.25))
    ref = torch.optim.AdamW(torch.nn.Linear(4, 3).parameters(), lr=1e-2)
    ref.load_state_dict(sd)                          # torch's own optimizer accepts it
    opt2 = InBackwardOptimizer(torch.optim.AdamW(torch.nn.Linear(4, 3).parameters(), lr=1e-2), B200HookState(wire="fp32", total_grad_elems=1))
    opt2.load_state_dict(sd)
    assert opt2._steps == 1 and opt2.param_groups[0]["lr"] == 5e-3
    assert torch.equal(opt2.state_dict()["state"][0]["exp_avg_sq"], torch.full((3, 4), 0.5))
    with pytest.raises(ValueError):
        InBackwardOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True), st)
    with pytest.raises(ValueError):
        InBackwardOptimizer(torch.optim.RMSprop(model.parameters(), lr=0.1), st)
