"""Worker-side session: which rank am I, and where do closures for the driver go.

Same public names and error behaviour as ray_lightning/session.py:6-63 (the Tune callbacks of
ray_lightning/tune.py and the launcher rely on them); one module-level slot instead of the
reference's global juggling.  The queue carries ``(rank, callable)`` pairs that the driver's
``util.process_results`` executes — control plane only, no tensor ever goes through it.
"""
from dataclasses import dataclass
from typing import Any, Callable, Optional

_HINT = ("\nFIX THIS by calling function in `session.py` like `get_actor_rank()` only from within an "
         "Pytorch Lightning actor session.")


@dataclass
class RayLightningSession:
    """State of one worker process for the duration of one ``launch``."""
    _rank: int
    _queue: Optional[Any] = None

    def __init__(self, rank: int, queue: Optional[Any]):
        self._rank, self._queue = rank, queue

    def get_actor_rank(self) -> int:
        return self._rank

    def set_queue(self, queue) -> None:
        self._queue = queue

    def put_queue(self, item: Callable) -> None:
        q = self._queue
        if q is None:
            raise ValueError("Trying to put something into session queue, but queue was not initialized. "
                             "This is probably a bug.")
        q.put((self._rank, item))


_slot = {"session": None}


def init_session(*args, **kwargs) -> None:
    if _slot["session"] is not None:
        raise ValueError("Trying to initialize RayLightningSession twice."
                         "\nFIX THIS by not calling `init_session()` manually.")
    _slot["session"] = RayLightningSession(*args, **kwargs)


def shutdown_session() -> None:
    """Forget the session (a worker process is reused for consecutive ``launch`` calls in tests)."""
    _slot["session"] = None


def get_session() -> RayLightningSession:
    s = _slot["session"]
    if not isinstance(s, RayLightningSession):
        raise ValueError("Trying to access RayLightningSession from outside an Pytorch Lightning run." + _HINT)
    return s


def set_session_queue(queue) -> None:
    get_session().set_queue(queue)


def get_actor_rank() -> int:
    return get_session().get_actor_rank()


def put_queue(*args, **kwargs) -> None:
    get_session().put_queue(*args, **kwargs)
