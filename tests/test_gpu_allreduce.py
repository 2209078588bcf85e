"""Parity of the CUDA allreduce path (K0/K1/K2 and the staged exchange K7-K10) against the oracle, through the C ABI.

Every test drives libb2d exactly as the DDP hook does (b2d_allreduce_bucket on flat fp32 buckets).
Multi-rank cases run W "loopback" ranks on the one GPU of the test box: separate contexts, arenas,
signal pads and streams — the complete inter-rank protocol, minus the NVLink wires.
Bar: bit-exact against oracle.ddp_oracle (both wires); golden fixtures from the reference's own
torch-DDP run reproduced bit-exactly at W=2.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ddp_oracle

pytestmark = pytest.mark.gpu

SIZES = [1, 7, 8, 9, 1000, 4099, 65536 + 3, (1 << 20) + 5]
_groups = {}


def group(world):
    from ray_lightning_b200.comm import LoopbackGroup
    if world not in _groups:
        _groups[world] = LoopbackGroup(world, 0, arena_bytes=768 << 20, timeout_ms=20000)
    return _groups[world]


def teardown_module(module):
    for g in _groups.values():
        g.close()
    _groups.clear()


def rank_inputs(world, n, seed=0, scale_pow=-4):
    out = []
    for r in range(world):
        gen = torch.Generator().manual_seed(1234 + r + 1000 * seed)
        out.append(torch.randn(n, generator=gen) * 2.0 ** scale_pow)
    return out


def oracle(per_rank, wire):
    fn = ddp_oracle.allreduce_bf16_wire if wire == "bf16" else ddp_oracle.allreduce_fp32_wire
    return fn(per_rank)


def same_bits(a, b):
    a = a.detach().cpu().contiguous().view(torch.int32)
    b = b.detach().cpu().contiguous().view(torch.int32)
    return torch.equal(a, b)


def run(world, per_rank, wire, algo, bucket_idx):
    g = group(world)
    bufs = [t.cuda() for t in per_rank]
    g.allreduce_(bufs, bucket_idx=bucket_idx, wire=wire, algo=algo)
    g.synchronize()
    return bufs


@pytest.mark.parametrize("wire", ["bf16", "fp32"])
def test_k0_world1_is_the_cast_roundtrip(wire):
    from ray_lightning_b200 import _b2d
    ctx = _b2d.Context(0, 1, 0, 1 << 20)
    try:
        for n in SIZES:
            x = rank_inputs(1, n)[0]
            buf = x.cuda()
            st = torch.cuda.current_stream()
            ctx.allreduce_bucket(0, buf.data_ptr(), n, _b2d.WIRE_NAMES[wire], 1.0, 0, st, st)
            torch.cuda.synchronize()
            assert same_bits(buf, oracle([x], wire)), (wire, n)
            if wire == "bf16":  # and it is what torch's own hook prologue+epilogue computes on this GPU
                ref = x.cuda().to(torch.bfloat16).div_(1).to(torch.float32)
                assert same_bits(buf, ref)
        assert ctx.stats()["launches"] == len(SIZES)
    finally:
        ctx.destroy()


@pytest.mark.parametrize("world", [3, 6])
def test_wire_value_matches_torch_cuda_division(world):
    """The oracle's claim about `buffer.to(bf16).div_(W)` on CUDA (multiply by fp32 reciprocal)."""
    x = rank_inputs(1, 100003)[0]
    ref = x.cuda().to(torch.bfloat16).div_(world).to(torch.float32).cpu()
    mine = ddp_oracle.wire_bf16(x, float(np.float32(1.0) / np.float32(world)))
    assert same_bits(ref, mine)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("wire", ["bf16", "fp32"])
@pytest.mark.parametrize("algo", ["one_shot", "two_shot", "staged"])
def test_allreduce_bit_exact_vs_oracle(world, wire, algo):
    for k, n in enumerate(SIZES):
        per_rank = rank_inputs(world, n, seed=k)
        want = oracle(per_rank, wire)
        idx = 100 * k + (10 if wire == "bf16" else 20) + ["one_shot", "two_shot", "staged"].index(algo) + 1
        bufs = run(world, per_rank, wire, algo, idx)
        for r in range(world):
            assert same_bits(bufs[r], want), (world, wire, algo, n, r)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_tma_staged_two_shot_bit_exact_vs_oracle(world):
    """K2T (cp.async.bulk into a shared-memory ring): same bits as the load/store kernels and the oracle.
    Sizes cover one partial macro tile, ragged last tiles, many tiles per block, and the fall-back shapes."""
    g = group(world)
    for k, n in enumerate([8, 64, 4096, 65536 + 8, (1 << 20) + 8, 7874560, 1000, 4099]):
        per_rank = rank_inputs(world, n, seed=20 + k)
        want = oracle(per_rank, "bf16")
        bufs = run(world, per_rank, "bf16", "two_shot_tma", 5000 + k)
        for r in range(world):
            assert same_bits(bufs[r], want), (world, n, r)
        algo_used = g.ranks[0].ctx.stats()["last_algo"]
        assert algo_used == (4 if n % 8 == 0 else 2)
    # fp32 wire falls back to the load/store two-shot
    per_rank = rank_inputs(world, 4096, seed=99)
    bufs = run(world, per_rank, "fp32", "two_shot_tma", 5100)
    assert same_bits(bufs[0], oracle(per_rank, "fp32"))


@pytest.mark.parametrize("world", [4, 8])
def test_tma_back_to_back_steps(world):
    g = group(world)
    n = 200000
    steps = [rank_inputs(world, n, seed=70 + s) for s in range(5)]
    bufs = [[t.cuda() for t in per_rank] for per_rank in steps]
    torch.cuda.synchronize()
    for s in range(5):
        g.allreduce_(bufs[s], bucket_idx=5200 + world, wire="bf16", algo="two_shot_tma")
    g.synchronize()
    for s in range(5):
        want = oracle(steps[s], "bf16")
        for r in range(world):
            assert same_bits(bufs[s][r], want), (s, r)


@pytest.mark.parametrize("world", [2, 8])
def test_auto_picks_one_shot_then_the_staged_exchange(world):
    from ray_lightning_b200 import _b2d
    g = group(world)
    ctx = g.ranks[0].ctx
    assert ctx.plan(1024, _b2d.WIRE_BF16)[0] == _b2d.ALGO_ONE_SHOT
    # no multicast object on one GPU: the P2P variant; at world 2 one-shot moves the same bytes and keeps winning up to 16 MiB
    assert ctx.plan(8 << 20, _b2d.WIRE_BF16)[0] == (_b2d.ALGO_ONE_SHOT if world == 2 else _b2d.ALGO_STAGED)
    assert ctx.plan(32 << 20, _b2d.WIRE_BF16)[0] == _b2d.ALGO_STAGED
    per_rank = rank_inputs(world, 1200001)       # 2.3 MiB of bf16 wire: above the one-shot range from world 4 up
    bufs = run(world, per_rank, "bf16", "auto", 7001)
    assert same_bits(bufs[0], oracle(per_rank, "bf16"))
    assert ctx.stats()["last_algo"] == (_b2d.ALGO_STAGED if world != 2 else _b2d.ALGO_ONE_SHOT)
    for rk in g.ranks:
        rk.ctx.set_auto_profile(_b2d.PROFILE_LATENCY)      # an isolated call: the single-kernel two-shot
    try:
        assert ctx.plan(32 << 20, _b2d.WIRE_BF16)[0] == _b2d.ALGO_TWO_SHOT
    finally:
        for rk in g.ranks:
            rk.ctx.set_auto_profile(_b2d.PROFILE_OVERLAP)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_staged_exchange_chunk_pipeline_bit_exact(world):
    """Several pipeline chunks per bucket (chunk = 64 KiB of wire): stage(c+1) | exchange(c) | write-back(c-1) on the
    three internal streams, both wires, ragged tail in the last chunk, twice on the same slot."""
    g = group(world)
    for rk in g.ranks:
        rk.ctx.set_chunk_bytes(64 << 10)
    try:
        for wire in ("bf16", "fp32"):
            for k, n in enumerate([(1 << 20) + 5, 300007]):
                for rep in range(2):
                    per_rank = rank_inputs(world, n, seed=600 + 10 * k + rep)
                    bufs = run(world, per_rank, wire, "staged", 7100 + k + 10 * (wire == "bf16"))
                    want = oracle(per_rank, wire)
                    for r in range(world):
                        assert same_bits(bufs[r], want), (world, wire, n, rep, r)
    finally:
        for rk in g.ranks:
            rk.ctx.set_chunk_bytes(32 << 20)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_fp32_buckets_that_live_in_the_arena_are_exchanged_in_place(world):
    """a bucket whose storage is arena memory needs no stage and no write-back; same bits as the
    staged path, ragged last pack included, bytes past the end untouched."""
    g = group(world)
    for k, n in enumerate([1, 6, 4099, 300001, (1 << 20) + 3]):
        per_rank = rank_inputs(world, n, seed=700 + k)
        bufs = []
        for rk, t in zip(g.ranks, per_rank):
            a = rk.arena_tensor(n + 4)
            a.fill_(777.0)
            a[:n].copy_(t)
            bufs.append(a)
        torch.cuda.synchronize()
        before = g.ranks[0].ctx.stats()["launches"]
        g.allreduce_([b[:n] for b in bufs], bucket_idx=7200 + k, wire="fp32", algo="staged")
        g.synchronize()
        # in place: arrive + exchange + wait per chunk — no stage / unstage kernels
        assert g.ranks[0].ctx.stats()["launches"] - before == 3
        want = oracle(per_rank, "fp32")
        for r in range(world):
            assert same_bits(bufs[r][:n], want), (world, n, r)
            assert float(bufs[r][n]) == 777.0 and float(bufs[r][n + 3]) == 777.0


@pytest.mark.parametrize("world", [2, 4])
def test_special_values_propagate(world):
    special = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), float("nan"), 1e-40, -1e-40, 3.3895e38,
                            1.00390625, 1.005859375, 2.0 ** -126, 2.0 ** -133, 65504.0, -1.0])
    per_rank = [torch.cat([special.roll(r), rank_inputs(world, 50)[r]]) for r in range(world)]
    for wire in ("bf16", "fp32"):
        for ai, algo in enumerate(("one_shot", "two_shot", "staged")):
            bufs = run(world, per_rank, wire, algo, 8000 + (wire == "bf16") * 3 + ai)
            want = oracle(per_rank, wire)
            got = bufs[0].cpu()
            assert torch.equal(torch.isnan(got), torch.isnan(want))
            ok = ~torch.isnan(want)
            assert same_bits(got[ok], want[ok])


@pytest.mark.parametrize("world,algo", [(2, "two_shot"), (4, "two_shot"), (8, "two_shot"), (4, "one_shot"),
                                        (2, "staged"), (4, "staged"), (8, "staged")])
def test_back_to_back_steps_without_host_sync(world, algo):
    """Five consecutive steps on the same bucket slot, launched back to back: exercises the double
    buffering that replaces a trailing barrier."""
    g = group(world)
    n = 200003
    steps = [rank_inputs(world, n, seed=50 + s) for s in range(5)]
    bufs = [[t.cuda() for t in per_rank] for per_rank in steps]
    torch.cuda.synchronize()
    for s in range(5):
        g.allreduce_(bufs[s], bucket_idx=9000 + world + 10 * (algo == "staged"), wire="bf16", algo=algo)
    g.synchronize()
    for s in range(5):
        want = oracle(steps[s], "bf16")
        for r in range(world):
            assert same_bits(bufs[s][r], want), (s, r)


@pytest.mark.parametrize("name", ["small_mlp_w2", "mnist_w2"])
def test_golden_fixtures_w2_bit_exact(name):
    """Fixtures written by the reference's real path (torch DDP over gloo): same buckets in, same bits out."""
    gold = load_golden(name)
    for i in range(int(gold["n_buckets"])):
        per_rank = [torch.from_numpy(gold["b%d_local_r%d" % (i, r)]) for r in range(2)]
        bf = run(2, per_rank, "bf16", "auto", 9500 + i)
        assert same_bits(bf[0], torch.from_numpy(gold["b%d_out_bf16_compress_hook" % i]))
        assert same_bits(bf[1], bf[0])
        fp = run(2, per_rank, "fp32", "auto", 9600 + i)
        assert same_bits(fp[0], torch.from_numpy(gold["b%d_out_default" % i]))


def test_golden_fixture_w4_within_tolerance():
    gold = load_golden("small_mlp_w4")
    for i in range(int(gold["n_buckets"])):
        per_rank = [torch.from_numpy(gold["b%d_local_r%d" % (i, r)]) for r in range(4)]
        fp = run(4, per_rank, "fp32", "two_shot", 9700 + i)
        # north star: rtol 1e-3 / atol 1e-5 against the reference's fp32 DDP path
        torch.testing.assert_close(fp[0].cpu(), torch.from_numpy(gold["b%d_out_default" % i]), rtol=1e-3, atol=1e-5)
        bf = run(4, per_rank, "bf16", "two_shot", 9800 + i)
        ref = torch.from_numpy(gold["b%d_out_bf16_compress_hook" % i]).double()
        exact = ddp_oracle.allreduce_exact_f64(per_rank)
        assert (bf[0].cpu().double() - exact).abs().max() <= (ref - exact).abs().max() + 1e-12


@pytest.mark.parametrize("world", [8])
def test_full_size_resnet50_bucket_properties(world):
    """BASELINE config 2's largest bucket (30.04 MiB fp32 = 7 874 560 elements) at world 8: direct
    oracle comparison plus size-independent properties."""
    n = 7874560
    per_rank = rank_inputs(world, n, seed=3)
    bufs = run(world, per_rank, "bf16", "two_shot", 9900)
    want = oracle(per_rank, "bf16")
    for r in range(world):
        assert same_bits(bufs[r], want)
    staged = run(world, per_rank, "bf16", "staged", 9901)
    for r in range(world):
        assert same_bits(staged[r], want)
    out = bufs[0].clone()
    # (1) results are bf16-representable: rounding again changes nothing
    assert same_bits(out, out.to(torch.bfloat16).float())
    # (2) linearity under exact (power-of-two) scaling
    bufs2 = run(world, [t * 4.0 for t in per_rank], "bf16", "two_shot", 9900)
    assert same_bits(bufs2[0], out * 4.0)
    # (3) identical ranks, constant input: the mean is the constant
    ones = [torch.full((n,), 0.5) for _ in range(world)]
    bufs3 = run(world, ones, "bf16", "two_shot", 9900)
    assert bool((bufs3[world - 1] == 0.5).all())
    # (4) a checksum of checksums: sum of the output equals the oracle's, bit for bit in fp64
    assert float(bufs[0].double().sum()) == float(want.double().sum())


@pytest.mark.parametrize("world", [4, 8])
@pytest.mark.parametrize("algo", ["one_shot", "two_shot", "two_shot_tma", "staged"])
def test_allreduce_with_skewed_ranks(world, algo):
    """A rotating straggler (about 1 ms of device sleep before its kernel) over 8 back-to-back steps on the
    same slot: barriers, arrival-order gather and double buffering under skew."""
    g = group(world)
    n = 120008
    steps = [rank_inputs(world, n, seed=200 + s) for s in range(8)]
    bufs = [[t.cuda() for t in per_rank] for per_rank in steps]
    torch.cuda.synchronize()
    producers = [torch.cuda.Stream() for _ in range(world)]   # the staged algorithms consume in wait-stream order
    for s in range(8):
        slow = (3 * s) % world
        with torch.cuda.stream(g.ranks[slow].stream):
            torch.cuda._sleep(2_000_000)
        with torch.cuda.stream(producers[slow]):
            torch.cuda._sleep(2_000_000)
        g.allreduce_(bufs[s], bucket_idx=9950 + ["one_shot", "two_shot", "two_shot_tma", "staged"].index(algo), wire="bf16",
                     algo=algo, wait_streams=producers if algo == "staged" else None)
    g.synchronize()
    for s in range(8):
        want = oracle(steps[s], "bf16")
        for r in range(world):
            assert same_bits(bufs[s][r], want), (s, r)


# ---- NVLS: needs an NVSwitch box with at least two GPUs -------------------------------------------------------
def _nvls_group():
    from ray_lightning_b200 import _b2d
    from ray_lightning_b200.comm import LoopbackGroup
    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("NVLS needs at least two GPUs behind an NVSwitch")
    world = 8 if nd >= 8 else (4 if nd >= 4 else 2)
    if "nvls" not in _groups:
        try:
            _groups["nvls"] = LoopbackGroup(world, devices=list(range(world)), arena_bytes=256 << 20, timeout_ms=20000,
                                            mem="vmm", nvls=True)
        except _b2d.B2DError as e:
            pytest.skip("multicast not available on this box: %s" % e)
    return _groups["nvls"]


def _nvls_run(g, per_rank, wire, algo, idx):
    bufs = [t.to("cuda:%d" % rk.device_index) for rk, t in zip(g.ranks, per_rank)]
    for rk in g.ranks:
        torch.cuda.synchronize(rk.device_index)
    g.allreduce_(bufs, bucket_idx=idx, wire=wire, algo=algo)
    g.synchronize()
    return bufs


@pytest.mark.parametrize("algo", ["nvls", "nvls_fused"])
def test_nvls_matches_the_exact_sum_to_one_rounding(algo):
    """The in-switch reduction (multimem.ld_reduce / multimem.st) on real NVLink.  Contract: every rank
    receives the SAME bits; fp32 wire within rtol 1e-6 of the rank-ordered oracle (north star: 1e-3 / 1e-5); bf16 wire:
    the result is one of the two bf16 neighbours of the exact (fp64) sum of the wire values — one rounding, as the oracle."""
    g = _nvls_group()
    world = g.world
    for k, n in enumerate([8, 4099, 300007, (1 << 20) + 5, 7874560]):
        per_rank = rank_inputs(world, n, seed=800 + k)
        fp = _nvls_run(g, per_rank, "fp32", algo, 7300 + k)
        want = oracle(per_rank, "fp32")
        for r in range(world):
            assert same_bits(fp[r], fp[0]), (n, r)
        torch.testing.assert_close(fp[0].cpu(), want, rtol=1e-5, atol=1e-8)
        bf = _nvls_run(g, per_rank, "bf16", algo, 7400 + k)
        for r in range(world):
            assert same_bits(bf[r], bf[0]), (n, r)
        scale = float(np.float32(1.0) / np.float32(world))
        exact = sum(ddp_oracle.wire_bf16(t, scale).double() for t in per_rank)
        got = bf[0].cpu().double()
        ulp = torch.maximum(got.abs(), exact.abs()) * 2.0 ** -7     # spacing of bf16 at that magnitude (upper bound)
        assert bool(((got - exact).abs() <= ulp + 1e-30).all())
        assert same_bits(bf[0], bf[0].to(torch.bfloat16).float())
        # and it is not further from the exact sum than the oracle's own rank-ordered result by more than one step
        mine = (got - exact).abs().max()
        ref = (oracle(per_rank, "bf16").double() - exact).abs().max()
        assert float(mine) <= 2.0 * float(ref) + 1e-12


def test_auto_prefers_nvls_from_four_ranks_up():
    from ray_lightning_b200 import _b2d
    g = _nvls_group()
    a = g.ranks[0].ctx.plan(32 << 20, _b2d.WIRE_BF16)[0]        # 64 MiB of wire: above every one-shot range
    assert a == (_b2d.ALGO_NVLS if g.world >= 4 else _b2d.ALGO_STAGED)
    for rk in g.ranks:
        rk.ctx.set_nvls_auto(False)
    try:
        assert g.ranks[0].ctx.plan(32 << 20, _b2d.WIRE_BF16)[0] == _b2d.ALGO_STAGED
    finally:
        for rk in g.ranks:
            rk.ctx.set_nvls_auto(True)


def test_staged_p2p_across_real_gpus_bit_exact():
    """The P2P variant over NVLink (peer loads + peer stores between distinct devices), bit-exact like on one GPU."""
    g = _nvls_group()
    for k, n in enumerate([9, 300007, 7874560]):
        per_rank = rank_inputs(g.world, n, seed=850 + k)
        for wire in ("bf16", "fp32"):
            bufs = _nvls_run(g, per_rank, wire, "staged", 7500 + 2 * k + (wire == "bf16"))
            want = oracle(per_rank, wire)
            for r in range(g.world):
                assert same_bits(bufs[r], want), (n, wire, r)


def test_nvls_in_place_fp32():
    g = _nvls_group()
    n = 300001
    per_rank = rank_inputs(g.world, n, seed=870)
    bufs = []
    for rk, t in zip(g.ranks, per_rank):
        a = rk.arena_tensor(n + 4)
        a.fill_(777.0)
        a[:n].copy_(t.to(a.device))
        bufs.append(a)
    for rk in g.ranks:
        torch.cuda.synchronize(rk.device_index)
    g.allreduce_([b[:n] for b in bufs], bucket_idx=7600, wire="fp32", algo="nvls")
    g.synchronize()
    want = oracle(per_rank, "fp32")
    for r in range(g.world):
        torch.testing.assert_close(bufs[r][:n].cpu(), want, rtol=1e-5, atol=1e-8)
        assert same_bits(bufs[r][:n], bufs[0][:n])
        assert float(bufs[r][n]) == 777.0
