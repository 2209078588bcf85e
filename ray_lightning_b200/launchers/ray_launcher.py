"""RayLauncher: one actor per worker, one GPU per actor.

Same role, method names and contracts as ray_lightning/launchers/ray_launcher.py:27-379 — the
drop-in boundary for ``strategy.launcher.launch(trainer._fit_impl, model, trainer=trainer)`` —
re-implemented for this repo's data path:

* every worker's CUDA_VISIBLE_DEVICES is widened to all GPUs its node's workers own
  (reference :177-219).  For the reference that lets NCCL reach the peers; here it is the
  PRECONDITION of libb2d's peer mapping (CUDA IPC / VMM import needs the peer device visible).
* MASTER_ADDR/MASTER_PORT are published for the ``env://`` rendezvous of the control-plane
  process group (reference :85-91,159-175) over which arena handles are exchanged.
* the model travels through the object store, the trainer inside the pickled bound method
  (reference :221-250,252-310); rank 0 returns a ``_RayOutput`` (reference :312-349).
"""
import os
from collections import defaultdict
from typing import Any, Callable, List, Optional, Tuple

import numpy as np
import torch

from .._compat import _Launcher, apply_to_collection, move_data_to_device, rank_zero_debug, ray
from ..session import init_session, shutdown_session
from ..util import load_state_stream, process_results, set_cuda_device_if_used, to_state_stream
from .utils import RayExecutor, _RayOutput, find_free_port

_FORWARDED_ENV = ("PL_GLOBAL_SEED", "PL_TORCH_DISTRIBUTED_BACKEND", "MASTER_ADDR", "MASTER_PORT")


def _tune_session_enabled():
    from ..tune import TUNE_INSTALLED, is_session_enabled
    return TUNE_INSTALLED and is_session_enabled()


class RayLauncher(_Launcher):
    def __init__(self, strategy) -> None:
        self._strategy = strategy
        self._start_method = "ray"
        self._workers = []
        self._futures = []
        self._master_addr = None
        self._master_port = None
        self._global_to_local = None
        self.tune_queue = None
        if not ray.is_initialized():
            ray.init()

    def is_interactive_compatible(self) -> bool:
        return True

    # ---- driver side ---------------------------------------------------------------------------
    def launch(self, function: Callable, *args: Any, trainer=None, **kwargs: Any) -> Any:
        """Spawn the workers, run ``function`` on all of them, pull rank 0's results back into
        ``trainer`` and tear everything down.  (reference :48-69)"""
        self.setup_workers()
        try:
            ray_output = self.run_function_on_workers(function, *args, trainer=trainer, **kwargs)
            if trainer is None:
                raise NotImplementedError("Ray launcher does not support trainer is None!")
            self._recover_results_in_main_process(ray_output, trainer)
            return ray_output.trainer_results
        finally:
            self.teardown_workers()
            self._strategy.teardown()

    def setup_workers(self, tune_enabled: bool = True) -> None:
        """Create the actors and give them a common environment.  (reference :71-103)"""
        self._workers = [self._create_worker() for _ in range(self._strategy.num_workers)]
        if self._strategy.init_hook:
            ray.get([w.execute.remote(self._strategy.init_hook) for w in self._workers])
        head = self._workers[0]
        self._master_addr = ray.get(head.get_node_ip.remote())
        self._master_port = str(ray.get(head.execute.remote(find_free_port)))
        self._setup_env_vars()
        if self._strategy.use_gpu:
            self._share_cuda_visible_devices()
        self._global_to_local = self.get_local_ranks()
        if tune_enabled and _tune_session_enabled():
            self.tune_queue = ray.util.queue.Queue(actor_options={"num_cpus": 0})

    def _create_worker(self):
        """One actor with the strategy's per-worker resources.  (reference :105-114)"""
        s = self._strategy
        return RayExecutor.options(num_cpus=s.num_cpus_per_worker, num_gpus=s.num_gpus_per_worker,
                                   resources=s.additional_resources_per_worker).remote()

    def teardown_workers(self):
        """Kill the actors (no restart).  (reference :116-128)"""
        if self.tune_queue:
            self.tune_queue.shutdown()
            self.tune_queue = None
        for w in self._workers:
            ray.kill(w, no_restart=True)
        self._workers = []

    def get_local_ranks(self) -> List[Optional[Tuple[int, int]]]:
        """global rank -> (local rank, node rank): nodes are numbered in order of first appearance
        of their IP, local ranks count up per IP.  (reference :130-157, pinned by
        ray_lightning/tests/test_ddp.py:80-114)"""
        ips = ray.get([w.get_node_ip.remote() for w in self._workers])
        node_of, seen_on = {}, defaultdict(int)
        mapping = []
        for ip in ips[:self._strategy.num_workers]:
            node_of.setdefault(ip, len(node_of))
            mapping.append((seen_on[ip], node_of[ip]))
            seen_on[ip] += 1
        return mapping + [None] * (self._strategy.num_workers - len(mapping))

    def _setup_env_vars(self):
        """Publish the rendezvous address and forward the PL env knobs.  (reference :159-175)"""
        os.environ["MASTER_ADDR"] = self._master_addr
        os.environ["MASTER_PORT"] = self._master_port
        keys = list(_FORWARDED_ENV)
        values = [os.getenv(k) for k in keys]
        ray.get([w.set_env_vars.remote(keys, values) for w in self._workers])

    def _share_cuda_visible_devices(self):
        """Every worker sees all GPUs owned by the workers of its node.  (reference :177-219)

        e.g. node A: w0 {0,1}, w1 {2,3}; node B: w2 {0,1}  ->  w0,w1: "0,1,2,3"; w2: "0,1"."""
        info = ray.get([w.get_node_and_gpu_ids.remote() for w in self._workers])
        workers_on, gpus_on = defaultdict(list), defaultdict(list)
        for wid, (node, gpu_ids) in enumerate(info):
            workers_on[node].append(wid)
            for g in gpu_ids:
                if g not in gpus_on[node]:
                    gpus_on[node].append(g)
        pending = []
        for node, gpu_ids in gpus_on.items():
            visible = ",".join(str(g) for g in gpu_ids)

            def export(visible=visible):
                os.environ["CUDA_DEVICE_ORDER"] = "PCI_BUS_ID"
                os.environ["CUDA_VISIBLE_DEVICES"] = visible

            pending += [self._workers[wid].execute.remote(export) for wid in workers_on[node]]
        ray.get(pending)

    def run_function_on_workers(self, function: Callable, *args: Any, trainer=None, **kwargs: Any):
        """Ship model + function to every worker and wait for rank 0's output.  (reference :221-250)"""
        model = trainer.model
        model_ref = ray.put(model)
        trainer.model = None  # the model goes through the object store, not inside the trainer pickle
        rest = tuple([None] + list(args[1:]))
        try:
            self._futures = [
                w.execute.remote(self._wrapping_function, rank, self._global_to_local, function, model_ref, rest,
                                 kwargs, self.tune_queue) for rank, w in enumerate(self._workers)
            ]
        finally:
            trainer.model = model
        return process_results(self._futures, self.tune_queue)[0]

    # ---- worker side ------------------------------------------------------------------------------
    def _wrapping_function(self, global_rank: int, global_to_local, function: Callable, model_ref, args: Any,
                           kwargs: Any, tune_queue) -> Any:
        """Runs inside the actor: bind ranks + device, join the process group, run ``function``
        (== trainer._fit_impl of the unpickled trainer copy).  (reference :252-310)"""
        strategy = self._strategy
        strategy.set_remote(True)
        strategy.set_global_to_local(global_to_local)

        trainer = function.__self__  # the bound method carries this worker's own trainer copy
        model = ray.get(model_ref) if isinstance(model_ref, getattr(ray, "ObjectRef", ())) else model_ref
        trainer.model = model
        args = tuple([model] + list(args[1:]))

        trainer._data_connector.prepare_data()
        if tune_queue is not None:
            shutdown_session()
            init_session(rank=global_rank, queue=tune_queue)

        strategy._worker_setup(process_idx=global_rank)
        trainer.strategy.root_device = strategy.root_device
        trainer.strategy.global_rank = strategy.global_rank
        trainer.strategy.local_rank = strategy.local_rank
        set_cuda_device_if_used(trainer.strategy)

        results = function(*args, **kwargs)
        out = self._collect_rank_zero_results(trainer, results)
        trainer.strategy.teardown_worker()
        return out

    def _collect_rank_zero_results(self, trainer, results: Any) -> Optional[_RayOutput]:
        """Rank 0 packs weights (as a byte stream — a temp file would not survive multi-node),
        trainer state and metrics.  (reference :312-349)"""
        rank_zero_debug("Finalizing the Ray launcher environment.")
        ckpt_cb = trainer.checkpoint_callback
        best_model_path = ckpt_cb.best_model_path if ckpt_cb else None
        state_dict = trainer.lightning_module.state_dict()
        if self._strategy.global_rank != 0:
            return None
        state_dict = move_data_to_device(state_dict, "cpu")
        stream = to_state_stream(state_dict)

        def as_numpy(t):  # numpy, not tensors: no shared-memory handles cross the actor boundary
            return t.cpu().numpy()

        callback_metrics = apply_to_collection(dict(trainer.callback_metrics), torch.Tensor, as_numpy)
        logged_metrics = apply_to_collection(dict(trainer.logged_metrics), torch.Tensor, as_numpy)
        return _RayOutput(best_model_path, stream, trainer.state, results, callback_metrics, logged_metrics)

    def _recover_results_in_main_process(self, ray_output: _RayOutput, trainer) -> None:
        """Driver: adopt rank 0's weights, state and metrics.  (reference :351-379)"""
        if trainer.checkpoint_callback:
            trainer.checkpoint_callback.best_model_path = str(ray_output.best_model_path)
        if ray_output.weights_path is not None:
            state_dict = load_state_stream(ray_output.weights_path, to_gpu=self._strategy.use_gpu)
            trainer.lightning_module.load_state_dict(state_dict)
        trainer.state = ray_output.trainer_state

        def as_tensor(a):
            return torch.tensor(a)

        trainer.callback_metrics.update(apply_to_collection(ray_output.callback_metrics, np.ndarray, as_tensor))
        trainer.logged_metrics.update(apply_to_collection(ray_output.logged_metrics, np.ndarray, as_tensor))
