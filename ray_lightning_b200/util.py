"""Driver-side helpers of the launcher — control plane only, no gradient bytes pass here.

Counterparts of ray_lightning/util.py:42-102 with the same names:
  process_results      poll the worker futures while running the closures they queue (ref :57-70)
  to_state_stream /    the rank-0 -> driver wire format of the trained weights: ``torch.save`` bytes
  load_state_stream    (ref :73-92; a temp file would not survive a multi-node run)
  set_cuda_device_if_used   late device binding in the worker (ref :95-102)
  Unavailable          placeholder for optional integrations that are not installed (ref :42-46)
"""
import io

import torch

from ._compat import rank_zero_info, ray


class Unavailable:
    """Stands in for a class of a missing optional dependency; cannot be instantiated."""

    def __init__(self, *_args, **_kwargs):
        raise RuntimeError("This class should never be instantiated.")


def _handle_queue(queue) -> None:
    """Execute every closure currently queued by the workers (items are ``(rank, callable)``)."""
    while not queue.empty():
        _rank, item = queue.get()
        if callable(item):
            item()


def process_results(training_result_futures, queue=None):
    """Block until every worker future is done, draining ``queue`` meanwhile; a failed worker raises
    here (through ``ray.get``).  Returns the list of worker results."""
    pending = list(training_result_futures)
    while pending:
        if queue:
            _handle_queue(queue)
        finished, pending = ray.wait(pending, timeout=0)
        ray.get(finished)            # surfaces a worker exception as soon as it exists
    if queue:
        _handle_queue(queue)         # whatever was queued right before the last worker returned
    return ray.get(list(training_result_futures))


def to_state_stream(model_state_dict) -> bytes:
    buf = io.BytesIO()
    torch.save(model_state_dict, buf)
    return buf.getvalue()


def load_state_stream(state_stream: bytes, to_gpu: bool):
    """Bytes -> state dict, on the current CUDA device when ``to_gpu`` and CUDA exists, else on CPU."""
    on_gpu = bool(to_gpu) and torch.cuda.is_available()
    where = (lambda storage, _loc: storage.cuda()) if on_gpu else "cpu"
    return torch.load(io.BytesIO(state_stream), map_location=where, weights_only=False)


def set_cuda_device_if_used(strategy) -> None:
    if not strategy.use_gpu:
        return
    rank_zero_info("GPU available: True (cuda), used: True")
    torch.cuda.set_device(strategy.root_device)
