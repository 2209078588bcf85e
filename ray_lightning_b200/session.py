"""Per-worker session: rank + the queue through which a worker hands closures to the driver.
Mirrors ray_lightning/session.py:6-63 (same names, same error behaviour)."""
from typing import Optional


class RayLightningSession:
    def __init__(self, rank: int, queue: Optional[object]):
        self._rank = rank
        self._queue = queue

    def get_actor_rank(self):
        return self._rank

    def set_queue(self, queue):
        self._queue = queue

    def put_queue(self, item):
        if self._queue is None:
            raise ValueError("Trying to put something into session queue, but queue was not initialized. "
                             "This is probably a bug.")
        self._queue.put((self._rank, item))


_session = None


def init_session(*args, **kwargs):
    global _session
    if _session:
        raise ValueError("Trying to initialize RayLightningSession twice."
                         "\nFIX THIS by not calling `init_session()` manually.")
    _session = RayLightningSession(*args, **kwargs)


def shutdown_session():
    global _session
    _session = None


def get_session() -> RayLightningSession:
    if not _session or not isinstance(_session, RayLightningSession):
        raise ValueError("Trying to access RayLightningSession from outside an Pytorch Lightning run."
                         "\nFIX THIS by calling function in `session.py` like `get_actor_rank()` only from "
                         "within an Pytorch Lightning actor session.")
    return _session


def set_session_queue(queue):
    get_session().set_queue(queue)


def get_actor_rank() -> int:
    return get_session().get_actor_rank()


def put_queue(*args, **kwargs):
    get_session().put_queue(*args, **kwargs)
