"""Host logic of the sharded path (sharded.py: flat layout, grad views, fused and generic optimizer steps, lr
schedulers, consolidated checkpoints) on CPU, with the communicator replaced by a TEST DOUBLE that implements
the three collectives with torch ops between W ranks running as threads.  The kernels themselves are covered by
test_kernel_emulation.py (CPU) and test_gpu_sharded.py / test_gpu_strategy.py (GPU); what is checked here is the
Python around them — in particular the generic (non-Adam) optimizer path."""
import threading

import numpy as np
import pytest
import torch

from ray_lightning_b200.sharded import FlatShards, ShardedOptimizer


class _Net:
    """World shared by the FakeComm instances of one test."""

    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.arenas = [dict() for _ in range(world)]   # rank -> {alloc index: tensor}


class FakeComm:
    """Stand-in for comm.Communicator: same method names and semantics, CPU tensors, threads as ranks."""
    is_test_double = True

    def __init__(self, net, rank):
        self.net, self.rank, self.world = net, rank, net.world
        self._allocs = 0

    def arena_tensor(self, numel, dtype=torch.float32):
        t = torch.zeros(numel, dtype=dtype)
        self.net.arenas[self.rank][self._allocs] = t
        self._allocs += 1
        t._b2d_alloc = self._allocs - 1
        return t

    def _exchange(self, obj):
        self.net.slots[self.rank] = obj
        self.net.barrier.wait()
        got = list(self.net.slots)
        self.net.barrier.wait()
        return got

    def reduce_scatter(self, grads, out, shard_off, wire="fp32", scale=None, slot=0, wait_stream=None, comm_stream=None):
        allg = self._exchange(grads)
        lo, hi = shard_off[self.rank], shard_off[self.rank + 1]
        acc = sum(g[lo:hi] * (1.0 / self.world) for g in allg)
        out[:hi - lo] = acc
        self.net.barrier.wait()
        return out

    def allgather_(self, buf, shard_off, wait_stream=None, comm_stream=None):
        allb = self._exchange(buf)
        for r in range(self.world):
            if r != self.rank:
                buf[shard_off[r]:shard_off[r + 1]] = allb[r][shard_off[r]:shard_off[r + 1]]
        self.net.barrier.wait()
        return buf

    # ---- the overlapped path: reduce buckets to their owners, then step + push ----------------------------
    def register_bucket(self, bucket_id, segs, wire="bf16"):
        if not hasattr(self, "_buckets"):
            self._buckets = {}
        self._buckets[bucket_id] = [tuple(s) for s in segs]

    def device_barrier(self, stream=None):
        self.net.barrier.wait()

    def reduce_to_owner(self, bucket_id, grads, reduced, shard_off, scale=None, zero_grads=True, accumulate=False,
                        nvls=False, wait_stream=None, comm_stream=None, phases=3):
        allg = self._exchange(grads.clone())
        lo = shard_off[self.rank]
        for off, n, owner in self._buckets[bucket_id]:
            if owner == self.rank:
                acc = sum(g[off:off + n] * (1.0 / self.world) for g in allg)
                if accumulate:
                    reduced[off - lo:off - lo + n] += acc
                else:
                    reduced[off - lo:off - lo + n] = acc
            if zero_grads:
                grads[off:off + n] = 0
        self.net.barrier.wait()

    def adam_push_(self, params, exp_avg, exp_avg_sq, reduced, shard_off, groups, nvls=False, wait_stream=None,
                   comm_stream=None, phases=6):
        lo, hi = shard_off[self.rank], shard_off[self.rank + 1]
        for glo, ghi, a in groups:
            p = torch.nn.Parameter(params[lo + glo:lo + ghi].clone())
            opt = (torch.optim.AdamW if a["adamw"] else torch.optim.Adam)(
                [p], lr=a["lr"], betas=(a["beta1"], a["beta2"]), eps=a["eps"], weight_decay=a["weight_decay"])
            opt.state[p] = {"step": torch.tensor(float(a["step"] - 1)), "exp_avg": exp_avg[glo:ghi].clone(),
                            "exp_avg_sq": exp_avg_sq[glo:ghi].clone()}
            p.grad = reduced[glo:ghi].clone()
            opt.step()
            params[lo + glo:lo + ghi] = p.detach()
            exp_avg[glo:ghi] = opt.state[p]["exp_avg"]
            exp_avg_sq[glo:ghi] = opt.state[p]["exp_avg_sq"]
        self.allgather_(params, shard_off)

    def sharded_step_(self, grads, params, exp_avg, exp_avg_sq, shard_off, step, lr, betas=(0.9, 0.999), eps=1e-8,
                      weight_decay=0.0, adamw=False, zero_grads=False, wire="bf16", scale=None, slot=0, wait_stream=None,
                      comm_stream=None):
        lo, hi = shard_off[self.rank], shard_off[self.rank + 1]
        g = torch.zeros(hi - lo)
        self.reduce_scatter(grads, g, shard_off)
        p = torch.nn.Parameter(params[lo:hi].clone())
        opt = (torch.optim.AdamW if adamw else torch.optim.Adam)([p], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        opt.state[p] = {"step": torch.tensor(float(step - 1)), "exp_avg": exp_avg[:hi - lo].clone(),
                        "exp_avg_sq": exp_avg_sq[:hi - lo].clone()}
        p.grad = g
        opt.step()
        params[lo:hi] = p.detach()
        exp_avg[:hi - lo] = opt.state[p]["exp_avg"]
        exp_avg_sq[:hi - lo] = opt.state[p]["exp_avg_sq"]
        self.allgather_(params, shard_off)


def make_model(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(13, 29), torch.nn.Tanh(), torch.nn.Linear(29, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))


def run_ranks(world, fn):
    out, err = [None] * world, []

    def wrap(r):
        try:
            out[r] = fn(r)
        except BaseException as e:   # surface worker failures in the test thread
            err.append(e)
            net_abort()

    net_abort = lambda: None
    ts = [threading.Thread(target=wrap, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    if err:
        raise err[0]
    return out


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("opt_name", ["adam", "adamw", "sgd_momentum", "rmsprop"])
def test_sharded_optimizer_equals_the_plain_optimizer_on_averaged_grads(world, opt_name):
    """W ranks with different batches + ShardedOptimizer  ==  one replica stepping the plain optimizer on the mean
    gradient; every rank ends with the same, whole parameters; a StepLR on the wrapper drives the shard's lr."""
    net = _Net(world)
    mk = {"adam": lambda ps: torch.optim.Adam(ps, lr=1e-2),
          "adamw": lambda ps: torch.optim.AdamW(ps, lr=1e-2, weight_decay=0.05),
          "sgd_momentum": lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9),
          "rmsprop": lambda ps: torch.optim.RMSprop(ps, lr=1e-3)}[opt_name]
    ref = make_model()
    ref_opt = mk(ref.parameters())
    ref_sched = torch.optim.lr_scheduler.StepLR(ref_opt, step_size=2, gamma=0.5)
    data = [[(torch.randn(6, 13, generator=torch.Generator().manual_seed(100 * s + r)),
              torch.randn(6, 3, generator=torch.Generator().manual_seed(7 + 100 * s + r))) for r in range(world)]
            for s in range(5)]

    models = [make_model() for _ in range(world)]   # built in the main thread: the global RNG is not thread safe

    def rank_fn(r):
        model = models[r]
        comm = FakeComm(net, r)
        base = mk(model.parameters())
        shards = FlatShards(model, comm, wire="fp32", reduce_bucket_mb=0.001)    # several reduce buckets
        assert len(shards.buckets) > 1
        sopt = ShardedOptimizer(base, shards, wire="fp32", overlap=bool(world % 2))   # with and without backward overlap
        assert sopt.fused == (opt_name in ("adam", "adamw"))
        sched = torch.optim.lr_scheduler.StepLR(sopt, step_size=2, gamma=0.5)
        for s in range(5):
            sopt.zero_grad()
            x, y = data[s][r]
            torch.nn.functional.mse_loss(model(x), y).backward()
            assert all(p.grad.data_ptr() == shards.flat_grads[o:o + n].data_ptr()       # autograd accumulated IN the flat buffer
                       for p, o, n in zip(shards.params, shards.offsets, shards.numels))
            if sopt.overlap:   # every bucket went to its owner during backward, and the staged gradients were zeroed
                assert sopt._pass_done and float(shards.flat_grads.abs().max()) == 0.0
            sopt.step()
            sched.step()
        sd = sopt.consolidated_state_dict()
        return [p.detach().clone() for p in model.parameters()], sd, sopt.param_groups[0]["lr"]

    outs = run_ranks(world, rank_fn)
    for s in range(5):
        ref_opt.zero_grad()
        for r in range(world):
            x, y = data[s][r]
            (torch.nn.functional.mse_loss(ref(x), y) / world).backward()
        ref_opt.step()
        ref_sched.step()
    for r in range(world):
        params, sd, lr = outs[r]
        assert lr == ref_opt.param_groups[0]["lr"]
        for a, b in zip(params, ref.parameters()):
            torch.testing.assert_close(a, b.detach(), rtol=1e-5, atol=1e-6)
        for a, b in zip(params, outs[0][0]):
            assert torch.equal(a, b)
        # consolidated optimizer state == the stock optimizer's state, parameter by parameter
        ref_sd = ref_opt.state_dict()
        assert sorted(sd["state"].keys()) == sorted(ref_sd["state"].keys())
        for i, st in ref_sd["state"].items():
            for k, v in st.items():
                if isinstance(v, torch.Tensor) and v.dim() > 0:
                    torch.testing.assert_close(sd["state"][i][k], v, rtol=1e-5, atol=1e-6)


def test_flat_layout_views_and_rebind():
    net = _Net(1)
    model = make_model()
    before = [p.detach().clone() for p in model.parameters()]
    sh = FlatShards(model, FakeComm(net, 0))
    assert sh.total % 8 == 0 and all(o % 8 == 0 for o in sh.offsets)
    for p, b, o, n in zip(model.parameters(), before, sh.offsets, sh.numels):
        assert torch.equal(p.detach(), b) and p.data_ptr() == sh.flat_params[o:o + n].data_ptr()
    for p in model.parameters():
        p.grad = None                      # what optimizer.zero_grad(set_to_none=True) does
    sh.rebind_grads()
    assert all(p.grad is not None and p.grad.data_ptr() == sh.flat_grads[o:o + n].data_ptr()
               for p, o, n in zip(sh.params, sh.offsets, sh.numels))
    with pytest.raises(RuntimeError, match="CUDA"):
        class RealLookingComm(FakeComm):
            is_test_double = False
        FlatShards(make_model(), RealLookingComm(net, 0))


def _param_groups(model):
    """The usual decay / no-decay split: two groups with different hyper-parameters."""
    decay = [p for n, p in model.named_parameters() if n.endswith("weight")]
    no_decay = [p for n, p in model.named_parameters() if not n.endswith("weight")]
    return [{"params": no_decay, "weight_decay": 0.0, "lr": 2e-2}, {"params": decay, "weight_decay": 0.1}]


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("opt_name", ["adamw", "sgd_momentum"])
def test_several_parameter_groups(world, opt_name):
    """Two optimizer parameter groups (different lr / weight decay): the flat layout keeps every group's share of a
    shard contiguous, the step applies each group's own constants, the consolidated state dict has torch's numbering."""
    net = _Net(world)
    mk = {"adamw": lambda gs: torch.optim.AdamW(gs, lr=1e-2),
          "sgd_momentum": lambda gs: torch.optim.SGD(gs, lr=0.05, momentum=0.9)}[opt_name]
    ref = make_model()
    ref_opt = mk(_param_groups(ref))
    data = [[(torch.randn(6, 13, generator=torch.Generator().manual_seed(100 * s + r)),
              torch.randn(6, 3, generator=torch.Generator().manual_seed(7 + 100 * s + r))) for r in range(world)]
            for s in range(4)]
    models = [make_model() for _ in range(world)]

    def rank_fn(r):
        from ray_lightning_b200.sharded import group_index_of
        model = models[r]
        base = mk(_param_groups(model))
        params = [p for p in model.parameters() if p.requires_grad]
        shards = FlatShards(model, FakeComm(net, r), wire="fp32", group_of=group_index_of(params, base), reduce_bucket_mb=0.001)
        sopt = ShardedOptimizer(base, shards, wire="fp32")
        assert len(sopt.param_groups) == 2 and sopt.param_groups[0]["lr"] == 2e-2
        for s in range(4):
            sopt.zero_grad()
            x, y = data[s][r]
            torch.nn.functional.mse_loss(model(x), y).backward()
            sopt.step()
        return [p.detach().clone() for p in model.parameters()], sopt.consolidated_state_dict()

    outs = run_ranks(world, rank_fn)
    for s in range(4):
        ref_opt.zero_grad()
        for r in range(world):
            x, y = data[s][r]
            (torch.nn.functional.mse_loss(ref(x), y) / world).backward()
        ref_opt.step()
    ref_sd = ref_opt.state_dict()
    for r in range(world):
        params, sd = outs[r]
        for a, b in zip(params, ref.parameters()):
            torch.testing.assert_close(a, b.detach(), rtol=1e-5, atol=1e-6)
        assert [g["params"] for g in sd["param_groups"]] == [g["params"] for g in ref_sd["param_groups"]]
        assert [g["weight_decay"] for g in sd["param_groups"]] == [0.0, 0.1]
        for i, st in ref_sd["state"].items():
            for k, v in st.items():
                if isinstance(v, torch.Tensor) and v.dim() > 0:
                    torch.testing.assert_close(sd["state"][i][k], v, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("opt_name", ["adam", "sgd_momentum"])
def test_resume_with_fewer_workers(opt_name):
    """The reference's contract (ray_lightning/tests/test_ddp_sharded.py:118-137): train on 2 workers, save the
    consolidated state, resume on 1 — the continued run equals an uninterrupted single-replica run."""
    mk = {"adam": lambda ps: torch.optim.Adam(ps, lr=1e-2),
          "sgd_momentum": lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9)}[opt_name]
    data = [[(torch.randn(6, 13, generator=torch.Generator().manual_seed(100 * s + r)),
              torch.randn(6, 3, generator=torch.Generator().manual_seed(7 + 100 * s + r))) for r in range(2)]
            for s in range(6)]
    ref = make_model()
    ref_opt = mk(ref.parameters())
    for s in range(6):
        ref_opt.zero_grad()
        world = 2 if s < 3 else 1
        for r in range(world):
            x, y = data[s][r]
            (torch.nn.functional.mse_loss(ref(x), y) / world).backward()
        ref_opt.step()

    net2 = _Net(2)
    models = [make_model() for _ in range(2)]

    def first_leg(r):
        model = models[r]
        shards = FlatShards(model, FakeComm(net2, r), wire="fp32")
        sopt = ShardedOptimizer(mk(model.parameters()), shards, wire="fp32")
        for s in range(3):
            sopt.zero_grad()
            x, y = data[s][r]
            torch.nn.functional.mse_loss(model(x), y).backward()
            sopt.step()
        return {k: v.detach().clone() for k, v in model.state_dict().items()}, sopt.consolidated_state_dict()

    weights, sd = run_ranks(2, first_leg)[0]

    net1 = _Net(1)
    model = make_model(seed=5)
    model.load_state_dict(weights)
    shards = FlatShards(model, FakeComm(net1, 0), wire="fp32")
    sopt = ShardedOptimizer(mk(model.parameters()), shards, wire="fp32")
    sopt.load_state_dict(sd)
    for s in range(3, 6):
        sopt.zero_grad()
        x, y = data[s][0]
        torch.nn.functional.mse_loss(model(x), y).backward()
        sopt.step()
    for a, b in zip(model.parameters(), ref.parameters()):
        torch.testing.assert_close(a.detach(), b.detach(), rtol=1e-5, atol=1e-6)


def test_gradient_accumulation_adds_on_the_owner():
    """Two backward passes before one step: the second reduce accumulates into the owner's shard (after a fence)."""
    world = 2
    net = _Net(world)
    data = [[(torch.randn(6, 13, generator=torch.Generator().manual_seed(10 * s + r)),
              torch.randn(6, 3, generator=torch.Generator().manual_seed(5 + 10 * s + r))) for r in range(world)]
            for s in range(2)]
    ref = make_model()
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for s in range(2):
        for r in range(world):
            x, y = data[s][r]
            (torch.nn.functional.mse_loss(ref(x), y) / world).backward()
    ref_opt.step()
    models = [make_model() for _ in range(world)]

    def rank_fn(r):
        model = models[r]
        shards = FlatShards(model, FakeComm(net, r), wire="fp32", reduce_bucket_mb=0.001)
        sopt = ShardedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), shards, wire="fp32")
        sopt.zero_grad()
        for s in range(2):
            x, y = data[s][r]
            torch.nn.functional.mse_loss(model(x), y).backward()
        sopt.step()
        return [p.detach().clone() for p in model.parameters()]

    for params in run_ranks(world, rank_fn):
        for a, b in zip(params, ref.parameters()):
            torch.testing.assert_close(a, b.detach(), rtol=1e-5, atol=1e-6)
