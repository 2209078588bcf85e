"""CPU restatement of the reference's gradient-sync arithmetic.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package; nothing under ``ray_lightning_b200/`` does (tests/test_no_oracle_in_product.py
enforces it).

The reference (ray-project/ray_lightning @ 24f5922) has no arithmetic of its own on this path:
``RayStrategy`` forwards ``**ddp_kwargs`` to torch's ``DistributedDataParallel``
(ray_lightning/ray_ddp.py:75,112-116) and the numbers are produced by torch:

  [R1] torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:18-54   allreduce_hook / _allreduce_fut
       (divide by world size, then SUM allreduce)  == C++ default, default_comm_hooks.hpp:36-51
  [R2] torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:57-93,116-134  bf16_compress_hook
       (buffer.to(bf16).div_(W) -> allreduce SUM in bf16 -> copy back to the fp32 buffer)
  [R3] torch/nn/parallel/distributed.py:1183-1281  _ddp_init_helper: bucket assignment by size
       ([1 MiB, cap], in declaration order, list reversed), reducer.hpp:30-31
  [R4] torch/distributed/optim/zero_redundancy_optimizer.py:651-722,870-894  greedy partition
       (sorted largest-first) and FairScale OSS.partition_parameters (declaration order; recalled)
  [R5] torch/optim/adam.py:347-547  single-tensor Adam (:530-547 non-capturable branch)

Parity pinning: the reference's own tests hold no golden vector for this path.
The restatement is therefore pinned against OUTPUTS OF THE REFERENCE'S IMPLEMENTATION RUN HERE —
real torch DDP over gloo, fixtures in tests/golden/ written by oracle/make_golden.py — and
against live torch calls (``dist._compute_bucket_assignment_by_size``, ``torch.optim.Adam``,
``ZeroRedundancyOptimizer``) in tests/test_oracle.py.
"""
import math

import numpy as np
import torch

MiB = 1024 * 1024


# ---- [R2]/[R1] per-bucket arithmetic --------------------------------------------------------
def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 (round to nearest even, NaN kept) -> fp32."""
    return x.to(torch.bfloat16).to(torch.float32)


def wire_bf16(g: torch.Tensor, scale: float) -> torch.Tensor:
    """[R2] ``buffer.to(torch.bfloat16).div_(world)``, the value one rank puts on the wire.

    On CUDA torch divides a tensor by a python scalar by multiplying with the fp32 reciprocal
    (ATen BinaryDivTrueKernel.cu, div_true_kernel_cuda, is_cpu_scalar branch), i.e.
    bf16(fp32(bf16(g)) * (1.0f / W)).  For power-of-two W this equals true division exactly.
    Returned as fp32 holding bf16-representable values.
    """
    s = torch.tensor(scale, dtype=torch.float32)
    return bf16_round(bf16_round(g.to(torch.float32)) * s)


def allreduce_bf16_wire(per_rank, scale=None) -> torch.Tensor:
    """Contract of libb2d's bf16 wire (B2D_WIRE_BF16) == [R2] with an exact (fp32-accumulate,
    round-once) SUM:  out = fp32( bf16( sum_{r=0..W-1, in order} wire_bf16(g_r) ) ).

    NCCL's bf16 SUM rounds at algorithm-dependent points (ring: every hop), so [R2] itself is
    only defined up to those roundings; gloo's W=2 result IS this formula (one add, one round),
    which is what tests/golden pins bit-exactly.
    """
    world = len(per_rank)
    scale = 1.0 / world if scale is None else scale
    acc = None
    for g in per_rank:  # strictly sequential fp32 adds, rank order
        c = wire_bf16(g, scale)
        acc = c if acc is None else acc + c
    return bf16_round(acc)


def allreduce_fp32_wire(per_rank, scale=None) -> torch.Tensor:
    """Contract of B2D_WIRE_FP32 == [R1]: out = sum_{r in order} (g_r * fp32(1/W)) in fp32.
    (reducer.cpp multiplies by 1/div_factor when copying grads into the bucket; the SUM order
    is backend-defined — gloo/NCCL differ from rank order by fp32 rounding only.)"""
    world = len(per_rank)
    scale = 1.0 / world if scale is None else scale
    s = torch.tensor(scale, dtype=torch.float32)
    acc = None
    for g in per_rank:
        c = g.to(torch.float32) * s
        acc = c if acc is None else acc + c
    return acc


def allreduce_exact_f64(per_rank, scale=None) -> torch.Tensor:
    """fp64 value of sum_r wire_bf16(g_r): the error yardstick of SURVEY §7.2(1)(ii)."""
    world = len(per_rank)
    scale = 1.0 / world if scale is None else scale
    acc = torch.zeros_like(per_rank[0], dtype=torch.float64)
    for g in per_rank:
        acc += wire_bf16(g, scale).to(torch.float64)
    return acc


# ---- [R3] bucket layout ---------------------------------------------------------------------
def bucket_assignment(param_numels, elem_size=4, bucket_cap_mb=25, first_bucket_mb=1, reverse=True, limits=None):
    """Greedy in-order packing of parameter indices into buckets, single dtype/device.

    Restates ``dist._compute_bucket_assignment_by_size(params, limits)`` followed by the
    ``list(reversed(bucket_indices))`` of [R3]: a bucket is closed as soon as its byte size
    reaches the current limit; the limit list advances by one per closed bucket and its last
    entry repeats; a tensor larger than the limit ends up alone.
    ``limits`` (bytes) defaults to DDP's ``[1 MiB, bucket_cap]`` (reducer.hpp:30-31); torch 2.11
    passes ``[cap]`` alone when ``bucket_cap_mb`` was given explicitly and ``[sys.maxsize]`` for
    the first iteration when ``find_unused_parameters=False`` (distributed.py:1224-1235).
    Returns a list of lists of parameter indices (bit-exact integer contract).
    """
    if limits is None:
        limits = [int(first_bucket_mb * MiB), int(bucket_cap_mb * MiB)]
    buckets, cur, cur_bytes, li = [], [], 0, 0
    for i, n in enumerate(param_numels):
        cur.append(i)
        cur_bytes += n * elem_size
        if cur_bytes >= limits[min(li, len(limits) - 1)]:
            buckets.append(cur)
            cur, cur_bytes = [], 0
            li += 1
    if cur:
        buckets.append(cur)
    return list(reversed(buckets)) if reverse else buckets


def bucket_offsets(param_numels, bucket):
    """(offsets, lengths) of each parameter inside its flat bucket (GradBucket layout, comm.hpp:20-98)."""
    offs, o = [], 0
    for i in bucket:
        offs.append(o)
        o += param_numels[i]
    return offs, [param_numels[i] for i in bucket]


# ---- [R4] owner partition -------------------------------------------------------------------
def partition_fairscale(param_numels, world):
    """FairScale ``OSS.partition_parameters``: declaration order, each parameter goes to the rank
    with the smallest running size, first minimum wins (recalled).  Returns owner[i]."""
    sizes = [0] * world
    owner = []
    for n in param_numels:
        r = sizes.index(min(sizes))
        owner.append(r)
        sizes[r] += n
    return owner


def partition_zero(param_numels, world):
    """torch ``ZeroRedundancyOptimizer._partition_parameters`` [R4]: parameters sorted by size,
    largest first (stable), then greedy smallest-rank-first with first-minimum tie-break."""
    order = sorted(range(len(param_numels)), key=lambda i: param_numels[i], reverse=True)
    sizes = [0] * world
    owner = [None] * len(param_numels)
    for i in order:
        r = sizes.index(min(sizes))
        owner[i] = r
        sizes[r] += param_numels[i]
    return owner


def shard_layout(param_numels, owner, world, align=8):
    """Flat layout used by b2d_sharded_step: parameters grouped by owner (declaration order inside
    an owner), every parameter start aligned to ``align`` elements.  Returns
    (param_offset[i], shard_off[world+1], total)."""
    offs = [0] * len(param_numels)
    shard_off = [0]
    cur = 0
    for r in range(world):
        for i, n in enumerate(param_numels):
            if owner[i] == r:
                offs[i] = cur
                cur += (n + align - 1) // align * align
        shard_off.append(cur)
    return offs, shard_off, cur


# ---- [R5] Adam -----------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, adamw=False):
    """One single-tensor Adam update in fp32 numpy, same operation order as [R5]
    (lerp, mul/addcmul, python-float bias corrections, sqrt/div/add eps, addcdiv).
    Mutates and returns (p, m, v)."""
    f = np.float32
    p = p.astype(f, copy=False); g = g.astype(f); m = m.astype(f, copy=False); v = v.astype(f, copy=False)
    if adamw:
        p *= f(1.0 - lr * weight_decay)
    elif weight_decay != 0.0:
        g = g + f(weight_decay) * p
    m += f(1.0 - beta1) * (g - m)
    v *= f(beta2)
    v += f(1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    denom = np.sqrt(v) / f(bc2_sqrt) + f(eps)
    p += f(-step_size) * (m / denom)
    return p, m, v
