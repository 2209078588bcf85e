"""Driver-side helpers of the launcher: result pump, state-dict byte stream, device binding.
Mirrors ray_lightning/util.py:42-102 (control plane; no gradient bytes pass here)."""
import io
from typing import Callable

import torch

from ._compat import ray, rank_zero_info


class Unavailable:
    """No object should be instance of this class (ray_lightning/util.py:42-46)."""

    def __init__(self, *args, **kwargs):
        raise RuntimeError("This class should never be instantiated.")


def _handle_queue(queue):
    """Run the closures workers have queued for the driver (ray_lightning/util.py:49-54)."""
    while not queue.empty():
        (actor_rank, item) = queue.get()
        if isinstance(item, Callable):
            item()


def process_results(training_result_futures, queue=None):
    """Drain the queue while the worker futures are outstanding, then return their results
    (ray_lightning/util.py:57-70).  A failed worker surfaces here as the exception of ray.get."""
    not_ready = training_result_futures
    ready = []
    while not_ready:
        if queue:
            _handle_queue(queue)
        ready, not_ready = ray.wait(not_ready, timeout=0)
        ray.get(ready)
    ray.get(ready)
    if queue:
        _handle_queue(queue)
    return ray.get(training_result_futures)


def to_state_stream(model_state_dict):
    """state dict -> bytes (torch.save), the driver<-rank-0 wire format (ray_lightning/util.py:73-77)."""
    _buffer = io.BytesIO()
    torch.save(model_state_dict, _buffer)
    return _buffer.getvalue()


def load_state_stream(state_stream, to_gpu):
    """bytes -> state dict on cpu, or on the current GPU when ``to_gpu`` and CUDA is available
    (ray_lightning/util.py:80-92)."""
    _buffer = io.BytesIO(state_stream)
    to_gpu = to_gpu and torch.cuda.is_available()
    return torch.load(_buffer, map_location=("cpu" if not to_gpu else lambda storage, loc: storage.cuda()),
                      weights_only=False)


def set_cuda_device_if_used(strategy) -> None:
    """Bind the worker process to its root device (ray_lightning/util.py:95-102)."""
    if strategy.use_gpu:
        rank_zero_info("GPU available: True (cuda), used: True")
        torch.cuda.set_device(strategy.root_device)
