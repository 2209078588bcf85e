"""K2P (role-decoupled chunk-pipelined two-shot, csrc/b2d_pipe.cuh) on a real GPU.

The kernel's logic is validated on CPU threads (tests/test_kernel_emulation.py) and its protocol by exhaustive
interleaving (tests/test_protocol_model.py), but it was written after round 1's GPU budget was spent and has not
run on hardware yet.  These tests therefore only run when B2D_TEST_EXPERIMENTAL=1; the algorithm is opt-in
(`two_shot_pipe` / `nvls_pipe`) and unreachable from AUTO."""
import os

import pytest
import torch

from test_gpu_allreduce import group, oracle, rank_inputs, run, same_bits

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B2D_TEST_EXPERIMENTAL") != "1",
                                 reason="experimental kernel, not yet validated on hardware (set B2D_TEST_EXPERIMENTAL=1)")]


def teardown_module(module):
    from test_gpu_allreduce import teardown_module as td
    td(module)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("wire", ["bf16", "fp32"])
def test_pipelined_two_shot_bit_exact_vs_oracle(world, wire):
    """K2P (role-decoupled chunk pipeline): same bits as K1/K2 and the oracle; sizes from less than one run
    to many chunks per block, ragged tails included."""
    for k, n in enumerate([1, 9, 1000, 4099, 65536 + 3, (1 << 20) + 5, 7874560]):
        per_rank = rank_inputs(world, n, seed=300 + k)
        want = oracle(per_rank, wire)
        bufs = run(world, per_rank, wire, "two_shot_pipe", 6000 + 10 * k + (wire == "bf16"))
        for r in range(world):
            assert same_bits(bufs[r], want), (world, wire, n, r)
    assert group(world).ranks[0].ctx.stats()["last_algo"] == 5


def test_pipelined_falls_back_for_other_world_sizes():
    per_rank = rank_inputs(3, 4099, seed=77)
    bufs = run(3, per_rank, "bf16", "two_shot_pipe", 6900)
    assert same_bits(bufs[0], oracle(per_rank, "bf16"))
    assert group(3).ranks[0].ctx.stats()["last_algo"] == 2


@pytest.mark.parametrize("world", [4, 8])
def test_pipelined_back_to_back_and_skewed(world):
    g = group(world)
    n = 700008          # several chunks per block at the loopback CTA budget
    steps = [rank_inputs(world, n, seed=400 + s) for s in range(8)]
    bufs = [[t.cuda() for t in per_rank] for per_rank in steps]
    torch.cuda.synchronize()
    for s in range(8):
        with torch.cuda.stream(g.ranks[(5 * s) % world].stream):
            torch.cuda._sleep(1_000_000)
        g.allreduce_(bufs[s], bucket_idx=6950, wire="bf16", algo="two_shot_pipe")
    g.synchronize()
    for s in range(8):
        want = oracle(steps[s], "bf16")
        for r in range(world):
            assert same_bits(bufs[s][r], want), (s, r)
