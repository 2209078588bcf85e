"""What the launcher instantiates in every worker process, and what rank 0 sends home.

Same surface as ray_lightning/launchers/utils.py:12-69:
  RayExecutor   the generic actor (one instance == one worker process == one GPU when ``use_gpu``) with
                ``set_env_var(s)``, ``get_node_ip``, ``get_node_and_gpu_ids``, ``execute``
  _RayOutput    the rank-0 -> driver result tuple (field names kept, ``weights_path`` is a byte stream)
  find_free_port, get_executable_cls
"""
import os
import socket
from typing import Any, Callable, Dict, List, NamedTuple, Optional

from .._compat import ray


def find_free_port() -> int:
    """Ask the kernel for an unused TCP port (called ON the rank-0 worker, whose address is MASTER_ADDR)."""
    sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    try:
        sock.bind(("", 0))
        sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        return sock.getsockname()[1]
    finally:
        sock.close()


def get_executable_cls():
    """Test hook of the Horovod launcher; nothing to override here."""
    return None


class _RayExecutorImpl:
    """Runs arbitrary callables inside its own process and answers a few questions about where it lives."""

    def set_env_var(self, key: str, value: str) -> None:
        if value is not None:           # None means "not set on the driver": leave the worker's value alone
            os.environ[key] = str(value)

    def set_env_vars(self, keys: List[str], values: List[str]) -> None:
        if len(keys) != len(values):
            raise AssertionError("keys and values differ in length")
        for k, v in zip(keys, values):
            self.set_env_var(k, v)

    def get_node_ip(self) -> str:
        return ray.util.get_node_ip_address()

    def get_node_and_gpu_ids(self):
        return ray.get_runtime_context().node_id.hex(), ray.get_gpu_ids()

    def execute(self, fn: Callable, *args, **kwargs):
        return fn(*args, **kwargs)


RayExecutor = ray.remote(_RayExecutorImpl)


class _RayOutput(NamedTuple):
    best_model_path: Optional[str]      # ModelCheckpoint.best_model_path in the worker
    weights_path: Optional[Any]         # state-dict byte stream (name kept from the reference)
    trainer_state: Any
    trainer_results: Any
    callback_metrics: Dict[str, Any]    # tensors converted to numpy for the trip
    logged_metrics: Dict[str, Any]
