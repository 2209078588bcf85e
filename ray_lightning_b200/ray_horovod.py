"""HorovodRayStrategy — constructor / property surface of ray_lightning/ray_horovod.py:31-183.

Horovod itself is absent from this image, so — exactly like the reference when Horovod is not
installed (ray_lightning/ray_horovod.py:78-79) — constructing the strategy raises
``RuntimeError("Please intall Horovod to use this strategy.")``.  With Horovod present the class
exposes the reference's rank properties; the gradient sync would remain Horovod's own.
"""
import torch

from ._compat import HorovodStrategy, ParallelStrategy, ray
from .util import Unavailable

try:  # pragma: no cover - horovod is not installable here
    import horovod.torch as hvd
    from horovod.ray import RayExecutor
except (ModuleNotFoundError, ImportError):
    HOROVOD_AVAILABLE = False
    RayExecutor = Unavailable
    hvd = Unavailable
else:  # pragma: no cover
    HOROVOD_AVAILABLE = True

from .launchers.ray_horovod_launcher import RayHorovodLauncher


def get_executable_cls():
    # Only used for testing purposes (ray_lightning/ray_horovod.py:25-28).
    return None


class HorovodRayStrategy(HorovodStrategy):
    strategy_name = "horovod_ray"

    def __init__(self, num_workers: int, num_cpus_per_worker: int = 1, use_gpu: bool = False):
        if not HOROVOD_AVAILABLE:
            raise RuntimeError("Please intall Horovod to use this strategy.")
        if not ray.is_initialized():  # pragma: no cover
            ray.init()
        ParallelStrategy.__init__(self, accelerator="_gpu" if use_gpu else "cpu")  # pragma: no cover
        self.num_workers = num_workers  # pragma: no cover
        self.cpus_per_worker = num_cpus_per_worker  # pragma: no cover
        self.use_gpu = use_gpu  # pragma: no cover
        self.executor = None  # pragma: no cover
        self._exit_stack = None  # pragma: no cover
        self._local_rank = 0  # pragma: no cover
        self._is_remote = False  # pragma: no cover

    def _configure_launcher(self):  # pragma: no cover
        settings = RayExecutor.create_settings(timeout_s=30)
        self.executor = RayExecutor(settings, num_workers=self.num_workers, cpus_per_worker=self.cpus_per_worker,
                                    use_gpu=self.use_gpu)
        self._launcher = RayHorovodLauncher(self)

    @property
    def global_rank(self) -> int:  # pragma: no cover
        return hvd.rank() if hvd.is_initialized() else 0

    @property
    def local_rank(self) -> int:  # pragma: no cover
        return hvd.local_rank() if hvd.is_initialized() else 0

    @property
    def world_size(self) -> int:  # pragma: no cover
        return hvd.size() if hvd.is_initialized() else self.num_workers

    def teardown(self) -> None:  # pragma: no cover
        self.join()
        self.accelerator = None
        super().teardown()

    @property
    def is_distributed(self):  # pragma: no cover
        return True

    def set_remote(self, remote: bool):  # pragma: no cover
        self._is_remote = remote

    @property
    def root_device(self):  # pragma: no cover
        if self.use_gpu and torch.cuda.is_available():
            return torch.device("cuda", hvd.local_rank() if hvd.is_initialized() else 0)
        return torch.device("cpu")
