"""Plumbing contracts of RayStrategy / RayLauncher on CPU workers (gloo), world size 1-2 — the
behaviours the reference's own suite pins (ray_lightning/tests/test_ddp.py), ported onto this
repo's actor runtime.  These cover the host side around the hot path: ranks, resources, samplers,
state hand-back, metrics; the arithmetic is covered by test_oracle.py / test_gpu_*.py."""
import os

import pytest
import torch
from torch.utils.data import DistributedSampler

from ray_lightning_b200 import RayStrategy
from ray_lightning_b200._compat import Callback, EarlyStopping, ray
from ray_lightning_b200.launchers.ray_launcher import RayLauncher
from utils import (BoringModel, LightningMNISTClassifier, MNISTDataModule, XORDataModule, XORModel, get_trainer,
                   load_test, predict_test, train_test)


@pytest.fixture
def ray_start_2_cpus():
    ray.init(num_cpus=2)
    yield
    ray.shutdown()


@pytest.fixture
def ray_start_4_cpus_extra():
    ray.init(num_cpus=4, resources={"extra": 4})
    yield
    ray.shutdown()


@pytest.fixture
def seed():
    from ray_lightning_b200._compat import seed_everything
    seed_everything(0)
    yield
    os.environ.pop("PL_GLOBAL_SEED", None)


@pytest.mark.parametrize("num_workers", [1, 2])
def test_actor_creation(tmpdir, ray_start_2_cpus, num_workers):
    """One actor per worker while training runs; none afterwards (reference :65-77)."""
    model = BoringModel()

    def check_num_actor():
        assert len(ray._state.actors) == num_workers

    class Check(Callback):
        pass

    strategy = RayStrategy(num_workers=num_workers)
    trainer = get_trainer(tmpdir, strategy=strategy)
    launcher = strategy.launcher
    orig = launcher.run_function_on_workers

    def spy(*a, **k):
        check_num_actor()
        return orig(*a, **k)

    launcher.run_function_on_workers = spy
    trainer.fit(model)
    assert len(ray._state.actors) == 0
    assert 2.0 == ray.available_resources()["CPU"]


def test_global_local_ranks(ray_start_2_cpus):
    """global -> (local, node) rank map with fake actors reporting fake IPs (reference :80-114)."""

    class Node1Actor:
        def get_node_ip(self):
            return "1"

    class Node2Actor:
        def get_node_ip(self):
            return "2"

    strategy = RayStrategy(num_workers=4, use_gpu=False)
    launcher = RayLauncher(strategy)
    launcher._workers = [ray.remote(Node1Actor).options(num_cpus=0).remote(), ray.remote(Node2Actor).options(num_cpus=0).remote(),
                         ray.remote(Node1Actor).options(num_cpus=0).remote(), ray.remote(Node2Actor).options(num_cpus=0).remote()]
    ranks = launcher.get_local_ranks()
    assert ranks == [(0, 0), (0, 1), (1, 0), (1, 1)]
    strategy.set_remote(True)
    strategy.set_global_to_local(ranks)
    strategy.set_world_ranks(3)
    assert (strategy.global_rank, strategy.local_rank, strategy.node_rank) == (3, 1, 1)


def test_custom_resources_and_overrides(ray_start_4_cpus_extra):
    """Resource arithmetic of the constructor (reference :117-176)."""
    s = RayStrategy(num_workers=2, num_cpus_per_worker=2, resources_per_worker={"extra": 1})
    assert s.additional_resources_per_worker == {"extra": 1} and s.num_cpus_per_worker == 2
    s = RayStrategy(num_workers=1, num_cpus_per_worker=2, resources_per_worker={"CPU": 3})
    assert s.num_cpus_per_worker == 3
    s = RayStrategy(num_workers=1, use_gpu=True, resources_per_worker={"GPU": 0})
    assert s.num_gpus_per_worker == 0 and not s.use_gpu
    s = RayStrategy(num_workers=1, use_gpu=False, resources_per_worker={"GPU": 1})
    assert s.num_gpus_per_worker == 1 and s.use_gpu
    s = RayStrategy(num_workers=1, use_gpu=False, resources_per_worker={"GPU": 2})
    assert s.num_gpus_per_worker == 2 and s.use_gpu
    with pytest.warns(UserWarning):
        s = RayStrategy(num_workers=2, use_gpu=True, resources_per_worker={"GPU": 0.5})
    assert s.num_gpus_per_worker == 0.5 and s.use_gpu
    # the custom resource is really reserved per actor
    strategy = RayStrategy(num_workers=2, num_cpus_per_worker=1, resources_per_worker={"extra": 1})
    launcher = RayLauncher(strategy)
    launcher.setup_workers(tune_enabled=False)
    assert ray.available_resources().get("extra") == 2.0
    launcher.teardown_workers()
    assert ray.available_resources().get("extra") == 4.0


def test_default_contract_is_the_references_and_bf16_is_asked_for_the_references_way():
    """No hook given -> DDP's default arithmetic (divide, fp32 SUM), like the reference (ray_lightning/ray_ddp.py:75,112-116
    registers none).  `ddp_comm_hook=bf16_compress_hook` — how one asks the reference for bf16 compression — selects the
    fused bf16-wire path instead of torch's cast + allreduce + copy sequence."""
    from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
    s = RayStrategy(num_workers=2, use_gpu=True)
    assert s._ddp_comm_hook.__name__ == "b200_allreduce_hook" and s._ddp_comm_state.wire == "fp32"
    assert s._ddp_comm_state.arena_buckets                       # fp32: DDP's buckets go into the arena
    b = RayStrategy(num_workers=2, use_gpu=True, ddp_comm_hook=default_hooks.bf16_compress_hook)
    assert b._ddp_comm_hook.__name__ == "b200_allreduce_hook" and b._ddp_comm_state.wire == "bf16"
    assert not b._ddp_comm_state.arena_buckets
    # with a wrapper around it the user's composition is left alone
    w = RayStrategy(num_workers=2, use_gpu=True, ddp_comm_hook=default_hooks.bf16_compress_hook,
                    ddp_comm_wrapper=default_hooks.fp16_compress_wrapper)
    assert w._ddp_comm_hook is default_hooks.bf16_compress_hook
    # and on CPU workers nothing is injected at all
    c = RayStrategy(num_workers=2, use_gpu=False)
    assert c._ddp_comm_hook is None


def test_strategy_is_picklable_and_kwargs_pass_through():
    import cloudpickle
    s = RayStrategy(num_workers=2, use_gpu=True, bucket_cap_mb=5, find_unused_parameters=False,
                    gradient_as_bucket_view=True, b200_wire="fp32", b200_algo="two_shot")
    assert s._ddp_kwargs == dict(bucket_cap_mb=5, find_unused_parameters=False, gradient_as_bucket_view=True)
    assert s._ddp_comm_hook.__name__ == "b200_allreduce_hook" and s._ddp_comm_state.wire == "fp32"
    s2 = cloudpickle.loads(cloudpickle.dumps(s))
    assert s2._ddp_comm_state.comm is None and s2._ddp_comm_state.algo == "two_shot"
    with pytest.raises(TypeError):
        RayStrategy(num_workers=1, b200_nonsense=1)
    # a user-supplied hook wins
    def my_hook(state, bucket):
        return None
    s3 = RayStrategy(num_workers=1, use_gpu=True, ddp_comm_hook=my_hook)
    assert s3._ddp_comm_hook is my_hook


def test_distributed_sampler(tmpdir, ray_start_2_cpus):
    """DistributedSampler(num_replicas=2, rank=global_rank), shuffling on train only (reference :179-211)."""
    model = BoringModel()
    assert not isinstance(model.train_dataloader().sampler, DistributedSampler)

    class SamplerCheck(Callback):
        def on_train_start(self, trainer, pl_module):
            s = trainer.train_dataloader.sampler
            assert isinstance(s, DistributedSampler) and s.shuffle
            assert s.num_replicas == 2 and s.rank == trainer.global_rank

        def on_validation_start(self, trainer, pl_module):
            s = trainer.val_dataloaders[0].sampler
            assert isinstance(s, DistributedSampler) and not s.shuffle
            assert s.num_replicas == 2 and s.rank == trainer.global_rank

    trainer = get_trainer(tmpdir, strategy=RayStrategy(num_workers=2), callbacks=[SamplerCheck()])
    trainer.fit(model)


@pytest.mark.parametrize("num_workers", [1, 2])
def test_train(tmpdir, ray_start_2_cpus, num_workers):
    train_test(get_trainer(tmpdir, strategy=RayStrategy(num_workers=num_workers)), BoringModel())


def test_test_without_fit(tmpdir, ray_start_2_cpus, seed):
    trainer = get_trainer(tmpdir, limit_train_batches=20, max_epochs=1, strategy=RayStrategy(num_workers=1, use_gpu=False))
    trainer.test(BoringModel())
    assert trainer.state.finished


@pytest.mark.parametrize("num_workers", [1, 2])
def test_load(tmpdir, ray_start_2_cpus, num_workers):
    load_test(get_trainer(tmpdir, strategy=RayStrategy(num_workers=num_workers, use_gpu=False)), BoringModel())


@pytest.mark.parametrize("num_workers", [1, 2])
def test_predict(tmpdir, ray_start_2_cpus, seed, num_workers):
    config = {"layer_1": 32, "layer_2": 32, "lr": 1e-2, "batch_size": 32}
    trainer = get_trainer(tmpdir, limit_train_batches=20, max_epochs=1,
                          strategy=RayStrategy(num_workers=num_workers, use_gpu=False))
    predict_test(trainer, LightningMNISTClassifier(config, tmpdir), MNISTDataModule(batch_size=32))


def test_early_stop(tmpdir, ray_start_2_cpus):
    """val_loss is constant: stop after patience + 1 validation epochs (reference :289-308)."""
    patience = 2
    trainer = get_trainer(tmpdir, max_epochs=500, strategy=RayStrategy(num_workers=1, use_gpu=False),
                          callbacks=[EarlyStopping(monitor="val_loss", patience=patience)],
                          limit_train_batches=1.0, limit_val_batches=1.0)
    trainer.fit(BoringModel())
    trained = BoringModel.load_from_checkpoint(trainer.checkpoint_callback.best_model_path)
    assert trained.val_epoch == patience + 1, trained.val_epoch


def test_early_stop_with_rank_dependent_metrics(tmpdir, ray_start_2_cpus):
    """Ranks that disagree about the monitored metric must still leave the fit loop together (PL's
    reduce_boolean_decision): rank 1 sees a val_loss that keeps improving, rank 0 a constant one.  Without the
    reduction rank 0 stops alone and rank 1 hangs in its next allreduce."""

    class Skewed(BoringModel):
        def validation_step(self, batch, batch_idx):
            self.layer(batch)
            rank = self.trainer.strategy.global_rank
            loss = torch.tensor(1.0 if rank == 0 else 1.0 / (1.0 + self.val_epoch))
            self.log("val_loss", loss)
            return {"x": loss}

    patience = 2
    trainer = get_trainer(tmpdir, max_epochs=50, strategy=RayStrategy(num_workers=2, use_gpu=False),
                          callbacks=[EarlyStopping(monitor="val_loss", patience=patience)],
                          limit_train_batches=2, limit_val_batches=2)
    trainer.fit(Skewed())
    assert trainer.state.finished
    trained = Skewed.load_from_checkpoint(trainer.checkpoint_callback.best_model_path)
    assert trained.val_epoch == patience + 1, trained.val_epoch     # rank 0's decision, taken by both


def test_unused_parameters(tmpdir, ray_start_2_cpus):
    """find_unused_parameters=False reaches torch DDP (reference :311-323); default is PL's True."""

    class Check(Callback):
        def __init__(self, want):
            self.want = want

        def on_train_start(self, trainer, pl_module):
            assert trainer.strategy.model.find_unused_parameters is self.want

    trainer = get_trainer(tmpdir, strategy=RayStrategy(num_workers=2, use_gpu=False, find_unused_parameters=False),
                          callbacks=[Check(False)])
    trainer.fit(BoringModel())


def test_metrics(tmpdir, ray_start_2_cpus):
    """Metrics logged in the workers come back to the driver, constants intact, `_step` forks only in
    logged_metrics (reference :326-352)."""
    trainer = get_trainer(tmpdir, strategy=RayStrategy(num_workers=2, find_unused_parameters=False), max_epochs=1,
                          limit_train_batches=2, limit_val_batches=2)
    trainer.fit(XORModel(), XORDataModule())
    cm, lm = trainer.callback_metrics, trainer.logged_metrics
    assert cm["avg_val_loss"] == lm["avg_val_loss"]
    assert lm["val_foo"] == torch.tensor(1.234) and cm["val_foo"] == torch.tensor(1.234)
    assert "val_loss_step" in lm and lm["val_bar_step"] == torch.tensor(5.678)
    assert "val_loss_step" not in cm and "val_bar_step" not in cm


def test_worker_failure_surfaces_on_the_driver(tmpdir, ray_start_2_cpus):
    class Exploding(BoringModel):
        def training_step(self, batch, batch_idx):
            raise ValueError("boom in worker")

    trainer = get_trainer(tmpdir, strategy=RayStrategy(num_workers=1))
    with pytest.raises(Exception) as ei:
        trainer.fit(Exploding())
    assert "boom in worker" in str(ei.value)
    assert len(ray._state.actors) == 0


def test_launch_without_trainer_is_rejected(ray_start_2_cpus):
    launcher = RayLauncher(RayStrategy(num_workers=1))

    class FakeTrainer:
        model = None

    with pytest.raises((NotImplementedError, AttributeError)):
        launcher.launch(lambda: None, trainer=None)


@pytest.mark.parametrize("num_gpus_per_worker,expect", [(0.4, [["0"], ["0"]]), (0.5, [["0"], ["0"]]),
                                                         (1, [["0"], ["1"]]), (2, [["0", "1"], ["2", "3"]])])
def test_gpu_ids_and_shared_visibility(num_gpus_per_worker, expect):
    """GPU bin packing + CUDA_VISIBLE_DEVICES sharing (reference tests/test_ddp_gpu.py:82-123 and
    launchers/ray_launcher.py:177-219), exercised without GPUs: ids are bookkeeping, not devices."""
    ray.init(num_cpus=2, num_gpus=4)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            strategy = RayStrategy(num_workers=2, use_gpu=True, resources_per_worker={"GPU": num_gpus_per_worker})
        launcher = RayLauncher(strategy)
        launcher.setup_workers(tune_enabled=False)
        ids = [ray.get(w.get_node_and_gpu_ids.remote())[1] for w in launcher._workers]
        assert ids == expect
        union = []
        for g in sum(expect, []):
            if g not in union:
                union.append(g)
        vis = ray.get([w.execute.remote(lambda: os.environ.get("CUDA_VISIBLE_DEVICES")) for w in launcher._workers])
        assert vis == [",".join(union)] * 2
        assert ray.get(launcher._workers[0].execute.remote(lambda: os.environ.get("CUDA_DEVICE_ORDER"))) == "PCI_BUS_ID"
        assert launcher.get_local_ranks() == [(0, 0), (1, 0)]
        launcher.teardown_workers()
    finally:
        ray.shutdown()


def test_root_device_is_the_position_in_the_shared_list(monkeypatch):
    """root_device.index = index of the worker's GPU id inside CUDA_VISIBLE_DEVICES (reference ray_ddp.py:259-304)."""
    s = RayStrategy(num_workers=2, use_gpu=True)
    assert s.root_device == torch.device("cpu") or s.root_device == torch.device("cuda:0")   # driver side
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    assert s.root_device == torch.device("cuda:0")             # driver: any device
    s.set_remote(True)
    monkeypatch.setattr(ray, "get_gpu_ids", lambda: [2])
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "0,1,2,3")
    assert s.root_device == torch.device("cuda:2")
    monkeypatch.setattr(ray, "get_gpu_ids", lambda: ["3", "1"])  # several: the first one wins
    assert s.root_device == torch.device("cuda:3")
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "0,1")
    with pytest.raises(RuntimeError, match="CUDA_VISIBLE_DEVICES set incorrectly"):
        s.root_device
    s.root_device = torch.device("cuda:1")                      # the launcher pins it (ray_launcher.py:296)
    assert s.root_device == torch.device("cuda:1")
    assert s.distributed_sampler_kwargs == dict(num_replicas=2, rank=0)
