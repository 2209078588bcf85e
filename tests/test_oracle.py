"""The oracle against (a) fixtures written by the reference's real implementation (torch DDP over
gloo, oracle/make_golden.py), (b) live torch functions it restates, (c) its own C twin."""
import math

import numpy as np
import pytest
import torch
import torch.distributed as dist

from conftest import load_golden
from oracle import c_oracle, ddp_oracle

MiB = 1024 * 1024


def _locals(g, i):
    return [torch.from_numpy(g["b%d_local_r%d" % (i, r)]) for r in range(int(g["world"]))]


@pytest.mark.parametrize("name", ["small_mlp_w2", "small_mlp_w4", "mnist_w2"])
def test_golden_bucket_layout_is_bit_exact(name):
    g = load_golden(name)
    numels = [int(x) for x in g["param_numels"]]
    limits = [int(0.004 * MiB)] if name.startswith("small_mlp") else [1 * MiB, 25 * MiB]
    mine = ddp_oracle.bucket_assignment(numels, limits=limits)
    theirs = [list(map(int, g["b%d_param_ids" % i])) for i in range(int(g["n_buckets"]))]
    assert mine == theirs
    assert c_oracle.bucket_assignment(numels, limits) == theirs
    for i, b in enumerate(theirs):
        _, lens = ddp_oracle.bucket_offsets(numels, b)
        assert lens == list(map(int, g["b%d_lengths" % i]))


@pytest.mark.parametrize("name", ["small_mlp_w2", "mnist_w2"])
def test_golden_w2_bit_exact(name):
    """World 2: one add per element, so the reference's result is defined to the bit."""
    g = load_golden(name)
    for i in range(int(g["n_buckets"])):
        loc = _locals(g, i)
        fp32 = ddp_oracle.allreduce_fp32_wire(loc).numpy()
        assert np.array_equal(fp32, g["b%d_out_default" % i])
        assert np.array_equal(g["b%d_out_allreduce_hook" % i], g["b%d_out_default" % i])
        bf16 = ddp_oracle.allreduce_bf16_wire(loc).numpy()
        assert np.array_equal(bf16, g["b%d_out_bf16_compress_hook" % i])
        assert np.array_equal(c_oracle.allreduce([t.numpy() for t in loc], "bf16"), bf16)
        assert np.array_equal(c_oracle.allreduce([t.numpy() for t in loc], "fp32"), fp32)


def test_golden_w4_within_reference_rounding():
    """World 4: gloo adds in ring order and (bf16) rounds every hop; our contract accumulates in fp32
    and rounds once.  fp32 wire: rtol 1e-3 / atol 1e-5 (north star).  bf16 wire: our error against
    the exact sum must not exceed the reference's own."""
    g = load_golden("small_mlp_w4")
    for i in range(int(g["n_buckets"])):
        loc = _locals(g, i)
        fp32 = ddp_oracle.allreduce_fp32_wire(loc).numpy()
        np.testing.assert_allclose(fp32, g["b%d_out_default" % i], rtol=1e-3, atol=1e-5)
        ours = ddp_oracle.allreduce_bf16_wire(loc).double().numpy()
        ref = g["b%d_out_bf16_compress_hook" % i].astype(np.float64)
        exact = ddp_oracle.allreduce_exact_f64(loc).numpy()
        assert np.abs(ours - exact).max() <= np.abs(ref - exact).max() + 1e-12
        # the reference rounds every partial sum to bf16 (half an ulp = 2^-9 relative, of the
        # partial sum), so the two may differ by that much of sum_r |c_r| — not of the result
        mag = sum(ddp_oracle.wire_bf16(t, 0.25).abs() for t in loc).double().numpy()
        assert (np.abs(ours - ref) <= 2.0 ** -7 * mag + 1e-30).all()


def test_bucket_assignment_matches_live_torch():
    rng = np.random.default_rng(0)
    for trial in range(20):
        numels = [int(x) for x in rng.integers(1, 400000, size=int(rng.integers(1, 60)))]
        params = [torch.empty(n, device="meta") for n in numels]
        for limits in ([1 * MiB, 25 * MiB], [1 * MiB, 1 * MiB], [5 * MiB], [2 ** 62]):
            want, _ = dist._compute_bucket_assignment_by_size(params, limits, [False] * len(params))
            assert ddp_oracle.bucket_assignment(numels, limits=limits, reverse=False) == want
            assert c_oracle.bucket_assignment(numels, limits, reverse=False) == want


def test_resnet50_bucket_sizes():
    """161 tensors, 25 557 032 elements, 5 buckets at the default cap."""
    import torchvision
    with torch.device("meta"):
        m = torchvision.models.resnet50()
    numels = [p.numel() for p in m.parameters()]
    assert len(numels) == 161 and sum(numels) == 25557032
    b = ddp_oracle.bucket_assignment(numels)
    assert len(b) == 5
    mib = [sum(numels[i] for i in bb) * 4 / MiB for bb in b]
    assert [round(x, 2) for x in mib] == [11.84, 30.04, 28.29, 25.77, 1.55]  # reversed: last layers first


def test_partition_rules():
    numels = [5, 3, 8, 1, 1, 9, 2]
    assert ddp_oracle.partition_fairscale(numels, 3) == [0, 1, 2, 1, 1, 0, 1]
    assert c_oracle.partition_fairscale(numels, 3) == [0, 1, 2, 1, 1, 0, 1]
    # ZeRO rule == what ZeroRedundancyOptimizer computes (sorted largest first)
    owner = ddp_oracle.partition_zero(numels, 3)
    sizes = [sum(n for n, o in zip(numels, owner) if o == r) for r in range(3)]
    assert sorted(sizes) == [9, 10, 10]
    offs, shard_off, total = ddp_oracle.shard_layout(numels, ddp_oracle.partition_fairscale(numels, 3), 3)
    assert shard_off[0] == 0 and shard_off[-1] == total and all(o % 8 == 0 for o in offs + shard_off)


def test_partition_zero_matches_live_torch(tmp_path):
    from torch.distributed.optim import ZeroRedundancyOptimizer
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="file://%s" % (tmp_path / "pg"), rank=0, world_size=1)
    try:
        rng = np.random.default_rng(1)
        numels = [int(x) for x in rng.integers(1, 5000, size=23)]
        params = [torch.nn.Parameter(torch.zeros(n)) for n in numels]
        z = ZeroRedundancyOptimizer(params, optimizer_class=torch.optim.Adam, lr=0.1)
        z.world_size = 4
        z._partition_parameters_cache.clear()
        parts = z._partition_parameters()
        owner = ddp_oracle.partition_zero(numels, 4)
        for r, groups in enumerate(parts):
            got = sorted(next(i for i, q in enumerate(params) if q is p) for gr in groups for p in gr["params"])
            assert got == [i for i, o in enumerate(owner) if o == r]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("wd,adamw", [(0.0, False), (0.01, False), (0.01, True)])
def test_adam_restatements_match_torch(wd, adamw):
    torch.manual_seed(0)
    p0 = torch.randn(1000)
    p = torch.nn.Parameter(p0.clone())
    cls = torch.optim.AdamW if adamw else torch.optim.Adam
    opt = cls([p], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    pn, mn, vn = p0.numpy().copy(), np.zeros(1000, np.float32), np.zeros(1000, np.float32)
    pc, mc, vc = p0.numpy().copy(), np.zeros(1000, np.float32), np.zeros(1000, np.float32)
    for step in range(1, 6):
        g = torch.randn(1000) * 0.1
        p.grad = g.clone()
        opt.step()
        ddp_oracle.adam_step(pn, g.numpy(), mn, vn, step, 1e-2, 0.9, 0.999, 1e-8, wd, adamw)
        c_oracle.adam(pc, np.ascontiguousarray(g.numpy()), mc, vc, step, 1e-2, 0.9, 0.999, 1e-8, wd, adamw)
        np.testing.assert_allclose(pn, p.detach().numpy(), rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(pc, p.detach().numpy(), rtol=2e-6, atol=1e-7)


def test_c_oracle_is_the_python_oracle_bit_for_bit():
    rng = np.random.default_rng(7)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-40, -1e-40, 3.3895e38, 1.0, 1.00390625,
                        1.005859375, 2 ** -126, 2 ** -133], dtype=np.float32)
    for world in (1, 2, 3, 5, 8):
        per = [np.concatenate([special, (rng.standard_normal(4099) * 2.0 ** rng.integers(-20, 6)).astype(np.float32)])
               for _ in range(world)]
        for wire in ("bf16", "fp32"):
            fn = ddp_oracle.allreduce_bf16_wire if wire == "bf16" else ddp_oracle.allreduce_fp32_wire
            want = fn([torch.from_numpy(x) for x in per]).numpy()
            got = c_oracle.allreduce(per, wire)
            assert np.array_equal(want.view(np.uint32) & 0xffbfffff, got.view(np.uint32) & 0xffbfffff) or \
                np.array_equal(np.isnan(want), np.isnan(got)) and np.array_equal(want[~np.isnan(want)], got[~np.isnan(got)])


def test_bf16_rounding_points():
    lib = c_oracle.lib()
    # ties to even at the bf16 boundary (8 significant bits)
    assert lib.oracle_bf16_round(1.00390625) == 1.0          # 1 + 2^-8: tie -> even (1.0)
    assert lib.oracle_bf16_round(1.01171875) == 1.015625     # 1 + 3*2^-8: tie -> even (1 + 2^-6)
    assert lib.oracle_bf16_round(1.005859375) == 1.0078125   # above the tie -> up
    assert math.isnan(lib.oracle_bf16_round(float("nan")))
    x = torch.tensor([1.00390625, 1.01171875, 1.005859375])
    assert ddp_oracle.bf16_round(x).tolist() == [1.0, 1.015625, 1.0078125]
