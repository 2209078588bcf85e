"""The generic actor the launcher spawns, and the rank-0 -> driver result tuple.
Mirrors ray_lightning/launchers/utils.py:12-69."""
import os
import socket
from contextlib import closing
from typing import Any, Callable, Dict, List, NamedTuple, Optional

from .._compat import ray


def find_free_port():
    """Find a free port on the machine (ray_lightning/launchers/utils.py:12-17)."""
    with closing(socket.socket(socket.AF_INET, socket.SOCK_STREAM)) as s:
        s.bind(("", 0))
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        return s.getsockname()[1]


def get_executable_cls():
    # Only used for testing purposes (ray_lightning/launchers/utils.py:20-24).
    return None


class _RayExecutorImpl:
    """A class to execute any arbitrary function remotely: one instance == one worker process == one
    GPU when ``use_gpu`` (ray_lightning/launchers/utils.py:27-52)."""

    def set_env_var(self, key: str, value: str):
        if value is not None:
            value = str(value)
            os.environ[key] = value

    def set_env_vars(self, keys: List[str], values: List[str]):
        assert len(keys) == len(values)
        for key, value in zip(keys, values):
            self.set_env_var(key, value)

    def get_node_ip(self):
        return ray.util.get_node_ip_address()

    def get_node_and_gpu_ids(self):
        return ray.get_runtime_context().node_id.hex(), ray.get_gpu_ids()

    def execute(self, fn: Callable, *args, **kwargs):
        return fn(*args, **kwargs)


RayExecutor = ray.remote(_RayExecutorImpl)


class _RayOutput(NamedTuple):
    """What rank 0 returns to the driver (ray_lightning/launchers/utils.py:55-69)."""
    best_model_path: Optional[str]
    weights_path: Optional[Any]   # the state-dict byte stream (name kept from the reference)
    trainer_state: Any
    trainer_results: Any
    callback_metrics: Dict[str, Any]
    logged_metrics: Dict[str, Any]
