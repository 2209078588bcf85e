"""BASELINE.json config #1: MNISTClassifier under RayStrategy(num_workers=2, use_gpu=False), CPU/gloo —
the reference's own runnable configuration (ray_lightning/examples/ray_ddp_example.py), synthetic data."""
import pytest

from ray_lightning_b200._compat import ray
from ray_lightning_b200.examples.ray_ddp_example import MNISTClassifier, train_mnist


@pytest.fixture
def ray_start_2_cpus():
    ray.init(num_cpus=2)
    yield
    ray.shutdown()


def test_mnist_example_two_cpu_workers(tmpdir, ray_start_2_cpus):
    config = {"layer_1": 32, "layer_2": 64, "lr": 1e-2, "batch_size": 32}
    assert sum(p.numel() for p in MNISTClassifier(config).parameters()) == 27882   # 109 KiB, one bucket
    trainer, model = train_mnist(config, num_epochs=1, num_workers=2, use_gpu=False, root=str(tmpdir))
    assert trainer.state.finished
    assert float(trainer.callback_metrics["ptl/val_accuracy"]) >= 0.5   # the reference's bar (tests/utils.py:256-272)
    assert "ptl/train_loss" in trainer.callback_metrics


def test_sharded_example_callback_keeps_its_two_scalar_allreduces(tmpdir, ray_start_2_cpus, capfd):
    """the example's CUDACallback averages epoch time and peak memory over the workers with two scalar
    allreduces (ray_lightning/examples/ray_ddp_sharded_example.py:33-36) — kept, on the control-plane group."""
    from ray_lightning_b200 import RayShardedStrategy
    from ray_lightning_b200._compat import Trainer
    from ray_lightning_b200.examples.ray_ddp_sharded_example import CUDACallback, TinyGPT
    trainer = Trainer(default_root_dir=str(tmpdir), max_epochs=1, limit_train_batches=2, callbacks=[CUDACallback()],
                      strategy=RayShardedStrategy(num_workers=2, use_gpu=False))
    trainer.fit(TinyGPT(d=32, layers=1, heads=2))
    assert trainer.state.finished
