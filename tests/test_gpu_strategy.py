"""End to end on the GPU box: Trainer.fit through RayStrategy -> RayLauncher -> actor worker(s) ->
torch DDP -> libb2d comm hook, compared with the reference's own DDP path on the same inputs
(ray_lightning/tests/test_ddp_gpu.py is the model: train on GPU workers, parameters on CUDA,
root_device == current device)."""
import os

import pytest
import torch

from ray_lightning_b200 import RayShardedStrategy, RayStrategy
from ray_lightning_b200._compat import Callback, ray
from utils import AdamBoringModel, BoringModel, get_trainer, train_test

pytestmark = pytest.mark.gpu


@pytest.fixture
def ray_gpu():
    n = torch.cuda.device_count()
    ray.init(num_cpus=4, num_gpus=n)
    yield n
    ray.shutdown()
    os.environ.pop("PL_TORCH_DISTRIBUTED_BACKEND", None)


class Probe(Callback):
    """Runs in the worker; reports through logged metrics (they travel back in _RayOutput)."""

    def on_train_end(self, trainer, pl_module):
        st = trainer.strategy.b200_state if hasattr(trainer.strategy, "b200_state") else None
        dev = next(pl_module.parameters()).device
        pl_module._current_fx = "training_step"
        pl_module.log("probe_param_is_cuda", float(dev.type == "cuda"), on_step=True, on_epoch=False)
        pl_module.log("probe_root_is_current", float(trainer.strategy.root_device.index == torch.cuda.current_device()),
                      on_step=True, on_epoch=False)
        pl_module.log("probe_hook_calls", float(st.calls if st is not None else -1), on_step=True, on_epoch=False)
        if st is not None and st.comm is not None:
            pl_module.log("probe_kernel_launches", float(st.comm.stats()["launches"]), on_step=True, on_epoch=False)


def test_fit_on_one_gpu_worker(tmpdir, ray_gpu):
    model = BoringModel()
    trainer = get_trainer(tmpdir, strategy=RayStrategy(num_workers=1, use_gpu=True), callbacks=[Probe()])
    train_test(trainer, model)
    m = trainer.logged_metrics
    assert m["probe_param_is_cuda"] == 1.0 and m["probe_root_is_current"] == 1.0
    assert m["probe_hook_calls"] >= 10 and m["probe_kernel_launches"] >= 10   # the CUDA path really ran


def _fit_weights(tmpdir, sub, strategy, model_cls=BoringModel):
    torch.manual_seed(0)
    model = model_cls()
    trainer = get_trainer(os.path.join(str(tmpdir), sub), strategy=strategy, limit_train_batches=8, limit_val_batches=1,
                          callbacks=[Probe()])
    trainer.fit(model)
    return [p.detach().clone() for p in model.parameters()], trainer.logged_metrics


def test_two_workers_match_the_reference_ddp_path_bit_for_bit(tmpdir, ray_gpu):
    """Same seeds, same data: weights after 8 SGD steps with libb2d's fp32-wire hook == weights with
    stock DDP (no hook) — at world 2 the fp32 sum is order independent, so the match is exact."""
    n = ray_gpu
    share = {"GPU": 1} if n >= 2 else {"GPU": 0.5}
    if n < 2:
        os.environ["PL_TORCH_DISTRIBUTED_BACKEND"] = "gloo"   # two workers on one device: NCCL refuses
    common = dict(num_workers=2, use_gpu=True, resources_per_worker=dict(share), find_unused_parameters=False)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ours, m = _fit_weights(tmpdir, "ours", RayStrategy(b200_wire="fp32", **common))
        ref, m_ref = _fit_weights(tmpdir, "ref", RayStrategy(b200_enable=False, **common))
    assert m["probe_hook_calls"] >= 8 and m_ref["probe_hook_calls"] == -1
    for a, b in zip(ours, ref):
        assert torch.equal(a, b)
    # bf16 wire: within the north star's tolerance of the fp32 reference path
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bf, _ = _fit_weights(tmpdir, "bf16", RayStrategy(b200_wire="bf16", **common))
    for a, b in zip(bf, ref):
        torch.testing.assert_close(a, b, rtol=2e-2, atol=2e-3)


def test_sharded_strategy_two_workers(tmpdir, ray_gpu):
    """RayShardedStrategy: fused reduce-scatter + Adam + all-gather; the driver gets whole, moved weights
    and a consolidated optimizer state in the checkpoint (reference tests/test_ddp_sharded.py:46-63)."""
    n = ray_gpu
    share = {"GPU": 1} if n >= 2 else {"GPU": 0.5}
    if n < 2:
        os.environ["PL_TORCH_DISTRIBUTED_BACKEND"] = "gloo"
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        strategy = RayShardedStrategy(num_workers=2, use_gpu=True, resources_per_worker=dict(share), b200_wire="fp32")
        model = AdamBoringModel()
        trainer = get_trainer(tmpdir, strategy=strategy)
        train_test(trainer, model)
    ckpt = torch.load(trainer.checkpoint_callback.best_model_path, weights_only=False)
    for (k, v), p in zip(ckpt["state_dict"].items(), model.state_dict().values()):
        assert torch.equal(v, p.cpu())
    st = ckpt["optimizer_states"][0]["state"]
    assert set(st.keys()) == {0, 1} and st[0]["exp_avg"].shape == (2, 32) and st[1]["exp_avg_sq"].shape == (2,)


def test_sharded_resume_with_fewer_workers_on_gpu(tmpdir, ray_gpu):
    """The reference's downsize contract on the GPU path (ray_lightning/tests/test_ddp_sharded.py:118-137): fit with 2
    sharded workers, then resume the consolidated checkpoint with ONE worker (a different flat layout and owner table)."""
    n = ray_gpu
    share = {"GPU": 1} if n >= 2 else {"GPU": 0.5}
    if n < 2:
        os.environ["PL_TORCH_DISTRIBUTED_BACKEND"] = "gloo"
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = AdamBoringModel()
        trainer = get_trainer(tmpdir, strategy=RayShardedStrategy(num_workers=2, use_gpu=True, resources_per_worker=dict(share)),
                              max_epochs=1)
        trainer.fit(model)
        path = trainer.checkpoint_callback.best_model_path
        ckpt = torch.load(path, weights_only=False)
        assert len(ckpt["optimizer_states"][0]["state"]) == 2 and ckpt["optimizer_states"][0]["state"][0]["step"] > 0
        model2 = AdamBoringModel()
        trainer2 = get_trainer(os.path.join(str(tmpdir), "resume"), strategy=RayShardedStrategy(num_workers=1, use_gpu=True),
                               max_epochs=2, resume_from_checkpoint=path)
        trainer2.fit(model2)
    assert trainer2.state.finished
    resumed = AdamBoringModel.load_from_checkpoint(trainer2.checkpoint_callback.best_model_path)
    assert resumed.val_epoch == 2
    ck2 = torch.load(trainer2.checkpoint_callback.best_model_path, weights_only=False)
    # Adam's step count continued from the checkpoint instead of restarting
    assert float(ck2["optimizer_states"][0]["state"][0]["step"]) > float(ckpt["optimizer_states"][0]["state"][0]["step"])
