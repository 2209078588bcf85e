"""Integer bookkeeping of the sharded path: who owns which parameter, and where it lives in the
flat buffers that b2d_sharded_step works on.  Bit-exact contracts (tests compare them with the
oracle's independent restatement and with torch's ZeroRedundancyOptimizer).

Ownership follows FairScale ``OSS.partition_parameters`` — what ``RayShardedStrategy`` reaches
through PL's DDPSpawnShardedStrategy (ray_lightning/ray_ddp_sharded.py:12-13): parameters in
declaration order, each to the rank with the smallest running element count, first minimum wins.
``rule="zero"`` gives torch's ZeroRedundancyOptimizer order (largest first) instead.
"""
from typing import List, Sequence, Tuple

ALIGN = 8  # elements: one 16-byte bf16 pack; every parameter and shard starts on it


def partition_parameters(numels: Sequence[int], world: int, rule: str = "fairscale") -> List[int]:
    if world < 1:
        raise ValueError("world must be >= 1")
    order = range(len(numels))
    if rule == "zero":
        order = sorted(order, key=lambda i: numels[i], reverse=True)
    elif rule != "fairscale":
        raise ValueError("unknown partition rule %r" % rule)
    sizes = [0] * world
    owner = [0] * len(numels)
    for i in order:
        r = min(range(world), key=lambda k: (sizes[k], k))
        owner[i] = r
        sizes[r] += numels[i]
    return owner


def flat_layout(numels: Sequence[int], owner: Sequence[int], world: int, align: int = ALIGN,
                group_of: Sequence[int] = None) -> Tuple[List[int], List[int], int]:
    """Parameters grouped by owner rank; inside a rank by optimizer parameter group (``group_of[i]``, default one
    group), then declaration order; each start aligned.  An owner's shard and, inside it, every parameter group's
    share are therefore contiguous.  Returns (param_offset[i], shard_off[0..world], total_elements)."""
    offsets = [0] * len(numels)
    shard_off = [0]
    cur = 0
    groups = sorted(set(group_of)) if group_of is not None else [0]
    for r in range(world):
        for g in groups:
            for i, n in enumerate(numels):
                if owner[i] == r and (group_of is None or group_of[i] == g):
                    offsets[i] = cur
                    cur += -(-n // align) * align
        shard_off.append(cur)
    return offsets, shard_off, cur
