"""Control-plane helpers around the launcher (ray_lightning/util.py, session.py): the worker -> driver
closure queue, the state-dict byte stream, the Unavailable sentinel."""
import pytest
import torch

from ray_lightning_b200 import session, tune, util
from ray_lightning_b200._compat import ray
from ray_lightning_b200.launchers.utils import RayExecutor, _RayOutput, find_free_port


def test_state_stream_round_trip():
    sd = {"w": torch.randn(3, 4), "b": torch.arange(5), "nested": {"x": torch.tensor(1.5)}}
    blob = util.to_state_stream(sd)
    assert isinstance(blob, bytes)
    back = util.load_state_stream(blob, to_gpu=False)
    assert torch.equal(back["w"], sd["w"]) and torch.equal(back["b"], sd["b"]) and back["nested"]["x"] == 1.5
    back2 = util.load_state_stream(blob, to_gpu=True)   # to_gpu is honoured only when CUDA exists (ref util.py:80-92)
    assert back2["w"].device.type == ("cuda" if torch.cuda.is_available() else "cpu")


def test_unavailable_and_tune_fallbacks():
    with pytest.raises(RuntimeError, match="should never be instantiated"):
        util.Unavailable()
    assert tune.TUNE_INSTALLED is False and tune.is_session_enabled() is False
    with pytest.raises(RuntimeError):
        tune.TuneReportCallback()
    with pytest.raises(RuntimeError):
        tune.get_tune_resources()


def test_session_singleton_contract():
    session.shutdown_session()
    with pytest.raises(ValueError, match="outside an Pytorch Lightning run"):
        session.get_actor_rank()
    session.init_session(rank=3, queue=None)
    try:
        assert session.get_actor_rank() == 3
        with pytest.raises(ValueError, match="twice"):
            session.init_session(rank=0, queue=None)
        with pytest.raises(ValueError, match="queue was not initialized"):
            session.put_queue(lambda: None)
    finally:
        session.shutdown_session()


def _worker_reports(rank, queue):
    """Runs inside an actor: what a TuneReportCallback does on rank 0 (ref tune.py:130-134)."""
    from ray_lightning_b200 import session as s
    s.shutdown_session()
    s.init_session(rank=rank, queue=queue)
    box = {"rank": rank}
    s.put_queue(lambda: box)      # a closure, executed later in the DRIVER process
    return rank * 10


def test_process_results_drains_the_worker_queue():
    ray.init(num_cpus=2)
    try:
        q = ray.util.queue.Queue()
        workers = [RayExecutor.options(num_cpus=1).remote() for _ in range(2)]
        futures = [w.execute.remote(_worker_reports, r, q) for r, w in enumerate(workers)]
        seen = []
        orig = util._handle_queue

        def spy(queue):
            while not queue.empty():
                rank, item = queue.get()
                seen.append((rank, item()))

        util._handle_queue = spy
        try:
            out = util.process_results(futures, q)
        finally:
            util._handle_queue = orig
        assert out == [0, 10]
        assert sorted(r for r, _ in seen) == [0, 1] and all(v["rank"] == r for r, v in seen)
        for w in workers:
            ray.kill(w)
        q.shutdown()
    finally:
        ray.shutdown()


def test_executor_env_and_ports():
    ray.init(num_cpus=1)
    try:
        w = RayExecutor.options(num_cpus=1).remote()
        ray.get(w.set_env_vars.remote(["B2D_A", "B2D_B"], ["1", None]))
        import os
        assert ray.get(w.execute.remote(lambda: (os.environ.get("B2D_A"), os.environ.get("B2D_B")))) == ("1", None)
        assert ray.get(w.get_node_ip.remote()) == "127.0.0.1"
        node, gpus = ray.get(w.get_node_and_gpu_ids.remote())
        assert isinstance(node, str) and gpus == []
        p = ray.get(w.execute.remote(find_free_port))
        assert 1024 < p < 65536
        ray.kill(w)
    finally:
        ray.shutdown()
    assert _RayOutput._fields == ("best_model_path", "weights_path", "trainer_state", "trainer_results",
                                  "callback_metrics", "logged_metrics")
