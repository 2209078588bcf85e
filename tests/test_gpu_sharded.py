"""Parity of the sharded path (K4 reduce-scatter to owner, K5 partitioned Adam, K6 parameter
all-gather) against the oracle: torch.optim.Adam applied to the averaged gradients, which is what
RayShardedStrategy's FairScale OSS computes in aggregate (ray_lightning/ray_ddp_sharded.py:12-13)."""
import numpy as np
import pytest
import torch

from oracle import ddp_oracle

pytestmark = pytest.mark.gpu

_groups = {}


def group(world):
    from ray_lightning_b200.comm import LoopbackGroup
    if world not in _groups:
        _groups[world] = LoopbackGroup(world, 0, arena_bytes=256 << 20, timeout_ms=20000)
    return _groups[world]


def teardown_module(module):
    for g in _groups.values():
        g.close()
    _groups.clear()


def layout(world, seed):
    rng = np.random.default_rng(seed)
    numels = [int(x) for x in rng.integers(1, 3000, size=37)] + [40000, 8, 1]
    owner = ddp_oracle.partition_fairscale(numels, world)
    offs, shard_off, total = ddp_oracle.shard_layout(numels, owner, world)
    return numels, owner, offs, shard_off, total


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("wire", ["fp32", "bf16"])
def test_reduce_scatter_bit_exact(world, wire):
    g = group(world)
    _, _, _, shard_off, total = layout(world, 1)
    per_rank = [torch.randn(total, generator=torch.Generator().manual_seed(r)) * 0.1 for r in range(world)]
    grads = [t.cuda() for t in per_rank]
    outs = [torch.zeros(shard_off[r + 1] - shard_off[r], device="cuda") for r in range(world)]
    g.reduce_scatter(grads, outs, shard_off, wire=wire, slot=1 + (wire == "bf16"))
    g.synchronize()
    scale = float(np.float32(1.0) / np.float32(world))
    if wire == "fp32":
        want = ddp_oracle.allreduce_fp32_wire(per_rank, scale)
    else:  # bf16 wire, fp32 accumulate, NOT rounded again: the sum never goes back on the wire
        want = None
        for t in per_rank:
            c = ddp_oracle.wire_bf16(t, scale)
            want = c if want is None else want + c
    for r in range(world):
        assert torch.equal(outs[r].cpu(), want[shard_off[r]:shard_off[r + 1]]), (world, wire, r)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_allgather_bit_exact(world):
    g = group(world)
    _, _, _, shard_off, total = layout(world, 2)
    full = torch.randn(total, generator=torch.Generator().manual_seed(5))
    bufs = []
    for r, rk in enumerate(g.ranks):
        b = rk.arena_tensor(total)
        b.zero_()
        b[shard_off[r]:shard_off[r + 1]] = full[shard_off[r]:shard_off[r + 1]].cuda()
        bufs.append(b)
    torch.cuda.synchronize()
    g.allgather_(bufs, shard_off)
    g.synchronize()
    for r in range(world):
        assert torch.equal(bufs[r].cpu(), full), r


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("wire", ["fp32", "bf16"])
@pytest.mark.parametrize("wd,adamw", [(0.0, False), (0.01, True)])
def test_sharded_step_matches_adam_on_averaged_grads(world, wire, wd, adamw):
    g = group(world)
    _, _, _, shard_off, total = layout(world, 3)
    p0 = torch.randn(total, generator=torch.Generator().manual_seed(11))
    params, ms, vs = [], [], []
    for r, rk in enumerate(g.ranks):
        p = rk.arena_tensor(total)
        p.copy_(p0.cuda())
        params.append(p)
        n_own = shard_off[r + 1] - shard_off[r]
        ms.append(torch.zeros(n_own, device="cuda"))
        vs.append(torch.zeros(n_own, device="cuda"))
    ref_p = torch.nn.Parameter(p0.clone())
    cls = torch.optim.AdamW if adamw else torch.optim.Adam
    opt = cls([ref_p], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    scale = float(np.float32(1.0) / np.float32(world))
    for step in range(1, 4):
        per_rank = [torch.randn(total, generator=torch.Generator().manual_seed(100 * step + r)) * 0.1
                    for r in range(world)]
        grads = [t.cuda() for t in per_rank]
        torch.cuda.synchronize()
        g.sharded_step_(grads, params, ms, vs, shard_off, step=step, lr=1e-2, betas=(0.9, 0.999), eps=1e-8,
                        weight_decay=wd, adamw=adamw, zero_grads=(step == 2), wire=wire, slot=10 + (wire == "bf16"))
        g.synchronize()
        if wire == "fp32":
            avg = ddp_oracle.allreduce_fp32_wire(per_rank, scale)
        else:
            avg = sum(ddp_oracle.wire_bf16(t, scale) for t in per_rank)
        ref_p.grad = avg.clone()
        opt.step()
        # every rank holds the same, whole parameter vector (bit-identical: it is a copy)
        for r in range(1, world):
            assert torch.equal(params[r], params[0]), (step, r)
        torch.testing.assert_close(params[0].cpu(), ref_p.detach(), rtol=2e-5, atol=2e-6)
        if step == 2:
            assert all(float(gr.abs().max()) == 0.0 for gr in grads)
    # optimizer state of the owned shard == the reference state of that slice
    st = opt.state[ref_p]
    for r in range(world):
        sl = slice(shard_off[r], shard_off[r + 1])
        torch.testing.assert_close(ms[r].cpu(), st["exp_avg"][sl], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(vs[r].cpu(), st["exp_avg_sq"][sl], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("world", [4, 8])
@pytest.mark.parametrize("wire", ["fp32", "bf16"])
def test_sharded_step_with_skewed_ranks(world, wire):
    """One rank at a time is held back by ~1 ms of device sleep before its kernel: every cross-rank
    ordering assumption (stage -> reduce, Adam -> gather, step k gather -> step k+1 Adam) is exercised with
    a straggler.  All ranks must still end every step with identical parameters equal to the reference."""
    g = group(world)
    _, _, _, shard_off, total = layout(world, 7)
    p0 = torch.randn(total, generator=torch.Generator().manual_seed(21))
    params, ms, vs = [], [], []
    for r, rk in enumerate(g.ranks):
        p = rk.arena_tensor(total)
        p.copy_(p0.cuda())
        params.append(p)
        n_own = shard_off[r + 1] - shard_off[r]
        ms.append(torch.zeros(max(n_own, 8), device="cuda"))
        vs.append(torch.zeros(max(n_own, 8), device="cuda"))
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-2)
    scale = float(np.float32(1.0) / np.float32(world))
    steps = 8
    all_grads = [[torch.randn(total, generator=torch.Generator().manual_seed(1000 * s + r)) * 0.1 for r in range(world)]
                 for s in range(steps)]
    dev_grads = [[t.cuda() for t in per_rank] for per_rank in all_grads]
    torch.cuda.synchronize()
    for s in range(steps):                       # launched back to back, no host sync in between
        with torch.cuda.stream(g.ranks[s % world].stream):
            torch.cuda._sleep(2_000_000)
        g.sharded_step_(dev_grads[s], params, ms, vs, shard_off, step=s + 1, lr=1e-2, wire=wire, slot=40 + (wire == "bf16"))
    g.synchronize()
    for s in range(steps):
        if wire == "fp32":
            avg = ddp_oracle.allreduce_fp32_wire(all_grads[s], scale)
        else:
            avg = sum(ddp_oracle.wire_bf16(t, scale) for t in all_grads[s])
        ref_p.grad = avg.clone()
        opt.step()
    for r in range(1, world):
        assert torch.equal(params[r], params[0]), r
    torch.testing.assert_close(params[0].cpu(), ref_p.detach(), rtol=1e-4, atol=1e-5)


# ---- the backward-overlapped path: reduce buckets to their owners (K11 + K12), then Adam + push (K13) ----------
def _buckets(numels, owner, offs, nb):
    """Cut the parameters into nb reduce buckets in reverse declaration order (what FlatShards does)."""
    idx = list(reversed(range(len(numels))))
    per = -(-len(idx) // nb)
    return [[(offs[i], -(-numels[i] // 8) * 8, owner[i]) for i in idx[k * per:(k + 1) * per]] for k in range(nb) if idx[k * per:(k + 1) * per]]


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("wire", ["fp32", "bf16"])
def test_reduce_to_owner_bit_exact_and_zeroes_the_gradients(world, wire):
    g = group(world)
    numels, owner, offs, shard_off, total = layout(world, 11)
    buckets = _buckets(numels, owner, offs, 4)
    for b, segs in enumerate(buckets):
        g.register_bucket(100 * (wire == "bf16") + b, segs, wire)
    scale = float(np.float32(1.0) / np.float32(world))
    for rep in range(2):     # twice: the second pass accumulates on the owner
        per_rank = [torch.randn(total, generator=torch.Generator().manual_seed(40 + r + 10 * rep)) * 0.1 for r in range(world)]
        grads = [t.cuda() for t in per_rank]
        if rep == 0:
            reduced = [torch.full((max(shard_off[r + 1] - shard_off[r], 8),), 7.0, device="cuda") for r in range(world)]
            first = None
        torch.cuda.synchronize()
        for b in range(len(buckets)):
            g.reduce_to_owner(100 * (wire == "bf16") + b, grads, reduced, shard_off, zero_grads=True, accumulate=rep == 1)
        g.synchronize()
        if wire == "fp32":
            want = ddp_oracle.allreduce_fp32_wire(per_rank, scale)
        else:
            want = None
            for t in per_rank:
                c = ddp_oracle.wire_bf16(t, scale)
                want = c if want is None else want + c
        # elements that belong to no parameter (alignment gaps) carry whatever the flat buffer held: compare real elements
        mask = torch.zeros(total, dtype=torch.bool)
        for o, n in zip(offs, numels):
            mask[o:o + n] = True
        if rep == 1:
            want = first + want
        else:
            first = want.clone()
        for r in range(world):
            lo, hi = shard_off[r], shard_off[r + 1]
            got = reduced[r][:hi - lo].cpu()
            assert torch.equal(got[mask[lo:hi]], want[lo:hi][mask[lo:hi]]), (world, wire, rep, r)
            assert float(grads[r][mask.cuda()].abs().max()) == 0.0


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("adamw", [False, True])
def test_adam_push_matches_torch_and_leaves_every_rank_whole(world, adamw):
    """K13 with TWO parameter groups per shard (different lr / weight decay) over three steps: every rank ends with
    the same, whole parameter vector == torch.optim.Adam(W) with the same groups on the reduced gradients."""
    g = group(world)
    _, _, _, shard_off, total = layout(world, 12)
    p0 = torch.randn(total, generator=torch.Generator().manual_seed(5))
    params = []
    for rk in g.ranks:
        p = rk.arena_tensor(total)
        p.copy_(p0.cuda())
        params.append(p)
    n_own = [shard_off[r + 1] - shard_off[r] for r in range(world)]
    ms = [torch.zeros(max(n, 8), device="cuda") for n in n_own]
    vs = [torch.zeros(max(n, 8), device="cuda") for n in n_own]
    split = [(n // 16) * 8 for n in n_own]                     # group 0: [0, split), group 1: [split, n)
    hyp = [dict(lr=1e-2, weight_decay=0.0), dict(lr=3e-3, weight_decay=0.1)]
    refs, opts = [], []
    for r in range(world):
        a = torch.nn.Parameter(p0[shard_off[r]:shard_off[r] + split[r]].clone())
        b = torch.nn.Parameter(p0[shard_off[r] + split[r]:shard_off[r + 1]].clone())
        refs.append((a, b))
        cls = torch.optim.AdamW if adamw else torch.optim.Adam
        opts.append(cls([{"params": [a], **hyp[0]}, {"params": [b], **hyp[1]}]))
    for step in range(1, 4):
        red = [torch.randn(max(n, 8), generator=torch.Generator().manual_seed(100 * step + r)) * 0.1 for r, n in enumerate(n_own)]
        reduced = [t.cuda() for t in red]
        groups = []
        for r in range(world):
            gs = []
            for gi, (lo, hi) in enumerate(((0, split[r]), (split[r], n_own[r]))):
                if hi > lo:
                    gs.append((lo, hi, dict(lr=hyp[gi]["lr"], beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=hyp[gi]["weight_decay"],
                                            step=step, adamw=int(adamw))))
            groups.append(gs)
        torch.cuda.synchronize()
        g.adam_push_(params, ms, vs, reduced, shard_off, groups)
        g.synchronize()
        for r in range(world):
            a, b = refs[r]
            a.grad = red[r][:split[r]].clone()
            b.grad = red[r][split[r]:n_own[r]].clone()
            opts[r].step()
        want = torch.cat([torch.cat([a.detach(), b.detach()]) for a, b in refs])
        for r in range(world):
            np.testing.assert_allclose(params[r].cpu().numpy(), want.numpy(), rtol=2e-5, atol=2e-6)
            assert torch.equal(params[r], params[0])
    # push only (a generic optimizer updated the shard): every rank receives every owner's values bit for bit
    for r in range(world):
        params[r][shard_off[r]:shard_off[r + 1]] = float(r + 1)
    torch.cuda.synchronize()
    g.adam_push_(params, None, None, None, shard_off, [[] for _ in range(world)])
    g.synchronize()
    for r in range(world):
        for o in range(world):
            assert bool((params[r][shard_off[o]:shard_off[o + 1]] == float(o + 1)).all())
