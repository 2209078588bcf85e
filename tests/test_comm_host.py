"""Host-side logic of the N>1 path on CPU (gloo, world size 2 and 3): the SCM_RIGHTS fd exchange that
carries VMM arena handles between worker processes, and the sizing / option plumbing of the hook state."""
import os
import socket
from contextlib import closing

import pytest
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
