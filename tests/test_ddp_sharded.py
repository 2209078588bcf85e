"""RayShardedStrategy plumbing on CPU workers (gloo): the contracts of
ray_lightning/tests/test_ddp_sharded.py — strategy recognised, checkpoint round trip equals the live
parameters, resume, resume with fewer workers.  On CPU the sharded state comes from torch's
ZeroRedundancyOptimizer (FairScale is not installable); the GPU path is tests/test_gpu_strategy.py."""
import os

import pytest
import torch

from ray_lightning_b200 import RayShardedStrategy, RayStrategy
from ray_lightning_b200._compat import ray
from ray_lightning_b200.partition import flat_layout, partition_parameters
from oracle import ddp_oracle
from utils import AdamBoringModel, BoringModel, get_trainer


@pytest.fixture
def ray_start_2_cpus():
    ray.init(num_cpus=2)
    yield
    ray.shutdown()


def test_strategy_identity():
    s = RayShardedStrategy(num_workers=2)
    assert isinstance(s, RayStrategy) and s.strategy_name == "ddp_sharded_ray"
    assert RayStrategy.strategy_name == "ddp_ray"
    mro = [c.__name__ for c in type(s).__mro__]
    assert mro.index("RayStrategy") < mro.index("DDPSpawnShardedStrategy") < mro.index("DDPSpawnStrategy")


def test_partition_is_the_oracles_bit_for_bit():
    import numpy as np
    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 8):
        numels = [int(x) for x in rng.integers(1, 100000, size=61)]
        for rule, ref in (("fairscale", ddp_oracle.partition_fairscale), ("zero", ddp_oracle.partition_zero)):
            owner = partition_parameters(numels, world, rule)
            assert owner == ref(numels, world)
            assert flat_layout(numels, owner, world) == ddp_oracle.shard_layout(numels, owner, world)


def test_checkpoint_equals_live_parameters_and_resume(tmpdir, ray_start_2_cpus):
    """(reference test_ddp_sharded.py:46-63, 83-104, 118-137)"""
    model = AdamBoringModel()
    trainer = get_trainer(tmpdir, strategy=RayShardedStrategy(num_workers=2), max_epochs=1)
    trainer.fit(model)
    path = trainer.checkpoint_callback.best_model_path
    ckpt = torch.load(path, weights_only=False)
    for (k, v), p in zip(ckpt["state_dict"].items(), model.state_dict().values()):
        assert torch.equal(v, p)
    assert len(ckpt["optimizer_states"][0]["state"]) == 2   # consolidated: both parameters present
    # resume with FEWER workers (2 -> 1) for one more epoch
    model2 = AdamBoringModel()
    trainer2 = get_trainer(os.path.join(str(tmpdir), "resume"), strategy=RayShardedStrategy(num_workers=1),
                           max_epochs=2, resume_from_checkpoint=path)
    trainer2.fit(model2)
    assert trainer2.state.finished
    resumed = AdamBoringModel.load_from_checkpoint(trainer2.checkpoint_callback.best_model_path)
    assert resumed.val_epoch == 2   # epoch 0 came from the checkpoint, epoch 1 ran after the resume


def test_test_without_fit(tmpdir, ray_start_2_cpus):
    trainer = get_trainer(tmpdir, strategy=RayShardedStrategy(num_workers=1))
    trainer.test(BoringModel())
    assert trainer.state.finished


def test_flat_layout_keeps_shards_and_parameter_groups_contiguous():
    """Integer contract of the group-aware flat layout (partition.flat_layout): every owner's shard is one contiguous,
    8-aligned range; inside it every optimizer parameter group is one contiguous range; nothing overlaps; with a single
    group the layout is the oracle's."""
    import numpy as np
    rng = np.random.default_rng(11)
    for world in (1, 2, 3, 8):
        numels = [int(x) for x in rng.integers(1, 5000, size=47)]
        owner = partition_parameters(numels, world)
        group_of = [int(x) for x in rng.integers(0, 3, size=len(numels))]
        offs, shard_off, total = flat_layout(numels, owner, world, group_of=group_of)
        assert shard_off[0] == 0 and shard_off[-1] == total and all(o % 8 == 0 for o in offs + shard_off)
        spans = sorted((o, o + -(-n // 8) * 8) for o, n in zip(offs, numels))
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))             # no overlap
        for r in range(world):
            mine = [i for i in range(len(numels)) if owner[i] == r]
            assert all(shard_off[r] <= offs[i] and offs[i] + numels[i] <= shard_off[r + 1] for i in mine)
            seen = []
            for i in sorted(mine, key=lambda i: offs[i]):
                if not seen or seen[-1] != group_of[i]:
                    seen.append(group_of[i])
            assert len(seen) == len(set(seen))                                   # each group appears as ONE run
        assert flat_layout(numels, owner, world, group_of=[0] * len(numels)) == ddp_oracle.shard_layout(numels, owner, world)
